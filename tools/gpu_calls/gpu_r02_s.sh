#!/bin/bash
# round 2, GPU call S (1 GPU): everything at HEAD (precision 2 now runs gdb_render_tc4.cu) - all GPU tests, smoke, bench line, launch list, ncu of the split kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/s_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/s_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; cut -c1-300 gpurun_out/s_bench.json; tail -2 gpurun_out/s_bench.err
timeout 300 ncu --set full --import-source on --clock-control none -k regex:render_tc4 --launch-skip 1 -c 1 -f -o gpurun_out/prof_k3_tc4_dtu python tools/bench_k3.py --workload dtu --precisions 2 --iters 3 > gpurun_out/s_ncu_tc4.log 2>&1
ls -la gpurun_out/prof_k3_tc4_dtu.ncu-rep
