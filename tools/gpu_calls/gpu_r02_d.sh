#!/bin/bash
# round 2, GPU call D (1 GPU): state of HEAD after re-entry: all GPU tests, K3 A/B, bench line, launch list, ncu full of K3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/d_pytest_all.log
for wl in dtu nerf llff; do
  python tools/bench_k3.py --workload $wl --precisions 4,1,2 --iters 8 2>&1 | grep precision
done | tee gpurun_out/d_bench_k3.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; cut -c1-3000 gpurun_out/d_bench.json; tail -3 gpurun_out/d_bench.err
timeout 300 python bench.py --mode train --steps 20 > gpurun_out/d_train.json 2> gpurun_out/d_train.err; cut -c1-1200 gpurun_out/d_train.json; tail -3 gpurun_out/d_train.err
for wl in nerf llff; do timeout 300 python bench.py --workload $wl --steps 10 --lean 2>/dev/null | cut -c1-900; done | tee gpurun_out/d_bench_lean.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/d_launches_dtu.csv python bench.py --steps 1 --warmup 3 --lean > gpurun_out/d_ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:render_tc2 --launch-skip 3 -c 1 -f -o gpurun_out/prof_k3_g3_dtu python bench.py --steps 1 --warmup 3 --lean > gpurun_out/d_ncu_dtu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:render_tc2 --launch-skip 3 -c 1 -f -o gpurun_out/prof_k3_g3_nerf python bench.py --workload nerf --steps 1 --warmup 3 --lean > gpurun_out/d_ncu_nerf.log 2>&1
ls -la gpurun_out/
