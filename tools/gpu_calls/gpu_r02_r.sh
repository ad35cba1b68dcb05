#!/bin/bash
# round 2, GPU call R (1 GPU): split-fp16 K3 in the third-generation layout (gdb_render_tc4.cu) - parity and A/B against the first generation
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "split_precision or render_fused" 2>&1 | tail -15 | cut -c1-300
for wl in dtu llff; do
  timeout 120 python tools/bench_k3.py --workload $wl --precisions 2 --iters 8 2>&1 | grep precision
  GDB_K3_SPLIT_GEN1=1 timeout 120 python tools/bench_k3.py --workload $wl --precisions 2 --iters 8 2>&1 | grep precision | sed 's/^/gen1 /'
done | tee gpurun_out/r_bench_k3_split.log
