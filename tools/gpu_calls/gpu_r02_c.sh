#!/bin/bash
# round 2, GPU call C (1 GPU): new tests, smooth-depth A/B, real-forward ncu, matrix tool smoke, train bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/c_pytest_all.log
for wl in dtu nerf llff; do
  python tools/bench_k3.py --workload $wl --precisions 4,1 --iters 8 2>&1 | grep precision
done | tee gpurun_out/c_bench_k3.log
timeout 300 python tools/multi_gpu_matrix.py --steps 3 --out gpurun_out/c_matrix_1gpu.jsonl > gpurun_out/c_matrix.log 2>&1; tail -25 gpurun_out/c_matrix.log | cut -c1-600
timeout 300 python bench.py --mode train --steps 20 > gpurun_out/c_train.json 2> gpurun_out/c_train.err; cat gpurun_out/c_train.json | cut -c1-1200; tail -3 gpurun_out/c_train.err
ncu --set full --import-source on --clock-control none -k regex:render_tc2 --launch-skip 3 -c 1 -f -o gpurun_out/prof_k3_g3c_dtu python bench.py --steps 1 --warmup 3 --lean > gpurun_out/c_ncu_dtu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:render_tc2 --launch-skip 3 -c 1 -f -o gpurun_out/prof_k3_g3c_nerf python bench.py --workload nerf --steps 1 --warmup 3 --lean > gpurun_out/c_ncu_nerf.log 2>&1
for wl in nerf llff; do python bench.py --workload $wl --steps 10 --lean 2>/dev/null | cut -c1-900; done | tee gpurun_out/c_bench_lean.log
