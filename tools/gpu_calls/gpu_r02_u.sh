#!/bin/bash
# round 2, GPU call U (1 GPU): dynamic tile assignment + rolled descriptor loop as the default - parity, then A/B against GDB_K3_STATIC=1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_network_gpu.py -x -q -m gpu 2>&1 | tail -4 | cut -c1-300
for wl in dtu llff nerf; do
  for prec in 1 2; do
    timeout 120 python tools/bench_k3.py --workload $wl --precisions $prec --iters 10 2>&1 | grep precision | sed "s/^/dyn    /"
    GDB_K3_STATIC=1 timeout 120 python tools/bench_k3.py --workload $wl --precisions $prec --iters 10 2>&1 | grep precision | sed "s/^/static /"
  done
done | tee gpurun_out/u_k3_dyn.log
for wl in dtu llff; do
  GDB_SKIP_DIGEST_CHECK=1 timeout 120 python tools/bench_k3.py --workload $wl --precisions 2 --iters 10 --lib gdb_nerf_b200/variants/lib_p1roll4.so 2>&1 | grep precision | sed "s/^/p1roll4 /"
done | tee -a gpurun_out/u_k3_dyn.log
