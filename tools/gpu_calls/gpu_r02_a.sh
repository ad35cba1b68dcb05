#!/bin/bash
# round 2, GPU call A: parity of the batched-gather K3 + A/B timings + one ncu capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "render_fused" 2>&1 | tail -15 > gpurun_out/a_pytest_k3.log
cat gpurun_out/a_pytest_k3.log
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/a_pytest_all.log
cat gpurun_out/a_pytest_all.log
for wl in dtu nerf llff; do
  for fb in 1 0 2; do
    GDB_K3_FB=$fb python tools/bench_k3.py --workload $wl --precisions 4,1 --iters 8 2>&1 | grep precision
  done
done | tee gpurun_out/a_bench_k3.log
ncu --set full --import-source on --clock-control none -k regex:render_tc2 --launch-skip 2 -c 1 -f -o gpurun_out/prof_k3_g3_dtu python tools/bench_k3.py --workload dtu --precisions 1 --iters 1 > gpurun_out/a_ncu_dtu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:render_tc2 --launch-skip 2 -c 1 -f -o gpurun_out/prof_k3_g3_nerf python tools/bench_k3.py --workload nerf --precisions 1 --iters 1 > gpurun_out/a_ncu_nerf.log 2>&1
ls -la gpurun_out/prof_k3_g3*
