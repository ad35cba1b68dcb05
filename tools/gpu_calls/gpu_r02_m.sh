#!/bin/bash
# round 2, GPU call M (1 GPU): all GPU tests; probability head hand-staged vs TMA-fed; bench line; NVTX-filtered launch list; ncu of the TMA kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/m_pytest_all.log
python tools/bench_ph.py 2>&1 | grep "prob head" | tee gpurun_out/m_bench_ph.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; cut -c1-700 gpurun_out/m_bench.json; tail -2 gpurun_out/m_bench.err
ncu --nvtx --nvtx-include "gdb_timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/m_launches_dtu.csv python bench.py --steps 1 --warmup 3 --lean > gpurun_out/m_ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:prob_head_tma --launch-skip 3 -c 1 -f -o gpurun_out/prof_ph_tma python tools/bench_ph.py --iters 2 > gpurun_out/m_ncu_ph.log 2>&1
ls -la gpurun_out | tail -8
