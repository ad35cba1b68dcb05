#!/bin/bash
# round 2, GPU call G (1 GPU): K3 generation 4 - parity, A/B against generation 3 (precision 5), ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_network_gpu.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/g_pytest.log
for wl in dtu nerf llff; do
  python tools/bench_k3.py --workload $wl --precisions 5,1 --iters 8 2>&1 | grep precision
done | tee gpurun_out/g_bench_k3.log
python tools/bench_k3.py --workload dtu --precisions 5,1 --iters 8 --narrow 2>&1 | grep precision | sed 's/^/narrow /' | tee -a gpurun_out/g_bench_k3.log
ncu --set full --import-source on --clock-control none -k regex:render_tc3 --launch-skip 2 -c 1 -f -o gpurun_out/prof_k3_g4_dtu python tools/bench_k3.py --workload dtu --precisions 1 --iters 1 > gpurun_out/g_ncu_dtu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:render_tc3 --launch-skip 2 -c 1 -f -o gpurun_out/prof_k3_g4_nerf python tools/bench_k3.py --workload nerf --precisions 1 --iters 1 > gpurun_out/g_ncu_nerf.log 2>&1
ls -la gpurun_out/prof_k3_g4*
