#!/bin/bash
# round 2, GPU call W (1 GPU): K3 experiments - unconditional gathers (no zero initialisation), GEMM 4 rounds rolled
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export GDB_SKIP_DIGEST_CHECK=1
for rep in 1 2; do
for wl in dtu llff nerf; do
  timeout 120 python tools/bench_k3.py --workload $wl --precisions 1 --iters 10 2>&1 | grep precision | sed "s/^/base   /"
  for v in ${VARIANTS}; do
    timeout 120 python tools/bench_k3.py --workload $wl --precisions 1 --iters 10 --lib gdb_nerf_b200/variants/lib_$v.so 2>&1 | grep precision | sed "s/^/$v   /"
  done
done
done | tee gpurun_out/w_k3_${TAG:-x}.log
