#!/bin/bash
# round 2, GPU call B: slot-major rows; parity + A/B timings + ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "render_fused" 2>&1 | tail -5 | tee gpurun_out/b_pytest_k3.log
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/b_pytest_all.log
for wl in dtu nerf llff; do
  for fb in 0 1; do
    GDB_K3_FB=$fb python tools/bench_k3.py --workload $wl --precisions 4,1 --iters 8 2>&1 | grep precision
  done
done | tee gpurun_out/b_bench_k3.log
ncu --set full --import-source on --clock-control none -k regex:render_tc2 --launch-skip 2 -c 1 -f -o gpurun_out/prof_k3_g3b_dtu python tools/bench_k3.py --workload dtu --precisions 1 --iters 1 > gpurun_out/b_ncu_dtu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:render_tc2 --launch-skip 2 -c 1 -f -o gpurun_out/prof_k3_g3b_nerf python tools/bench_k3.py --workload nerf --precisions 1 --iters 1 > gpurun_out/b_ncu_nerf.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; tail -c 3000 gpurun_out/b_bench.json; tail -5 gpurun_out/b_bench.err
