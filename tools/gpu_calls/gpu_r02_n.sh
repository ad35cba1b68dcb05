#!/bin/bash
# round 2, GPU call N (1 GPU): everything at HEAD - all GPU tests, smoke, bench line, K3 capture for roofline.traffic
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/n_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/n_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; cut -c1-300 gpurun_out/n_bench.json; tail -2 gpurun_out/n_bench.err
ncu --set full --import-source on --clock-control none -k regex:render_tc2 --launch-skip 3 -c 1 -f -o gpurun_out/prof_k3_final_dtu python bench.py --steps 1 --warmup 3 --lean > gpurun_out/n_ncu_dtu.log 2>&1
ls -la gpurun_out/prof_k3_final_dtu.ncu-rep
