#!/bin/bash
# round 2, GPU call K (1 GPU): K3 generation 4 with the warp-uniform second-level skip
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k render_fused 2>&1 | tail -2 | cut -c1-250
for wl in dtu nerf llff; do python tools/bench_k3.py --workload $wl --precisions 1 --iters 10 2>&1 | grep precision; done | tee gpurun_out/k_bench_k3.log
