#!/bin/bash
# round 2, GPU call X (1 GPU): everything at HEAD - all GPU tests, smoke, bench line, NVTX-filtered launch list, ncu --set full of both production K3 kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/x_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/x_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err; cut -c1-300 gpurun_out/x_bench.json; tail -2 gpurun_out/x_bench.err
timeout 600 ncu --nvtx --nvtx-include "gdb_timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x_launches_dtu.csv python bench.py --steps 1 --warmup 3 --lean > gpurun_out/x_ncu_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:render_tc2 --launch-skip 3 -c 1 -f -o gpurun_out/prof_k3_s3_dtu python bench.py --steps 1 --warmup 3 --lean > gpurun_out/x_ncu_dtu.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:render_tc4 --launch-skip 1 -c 1 -f -o gpurun_out/prof_k3_tc4_dtu python tools/bench_k3.py --workload dtu --precisions 2 --iters 3 > gpurun_out/x_ncu_tc4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
