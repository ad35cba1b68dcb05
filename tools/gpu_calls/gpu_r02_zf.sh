#!/bin/bash
# round 2, GPU call ZF (1 GPU): the new K3 defaults (single-MUFU reciprocal / square roots, composed colour-pass projection) at HEAD:
# all GPU tests, smoke, ncu --set full of K3 inside bench.py (-> profiles/k3_dram_traffic.json), the bench line, the NVTX-filtered
# launch list, K3 alone default vs the -DGDB_X_EXACT build (precisions 1 and 2), errors against the oracle, lean llff / nerf lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/k3truth
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 | cut -c1-400 | tee gpurun_out/zf_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/zf_smoke.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:render_tc2 --launch-skip 3 -c 1 -f -o gpurun_out/prof_k3_s4_dtu python bench.py --steps 1 --warmup 3 --lean > gpurun_out/zf_ncu_dtu.log 2>&1
ncu -i gpurun_out/prof_k3_s4_dtu.ncu-rep --page raw --csv > gpurun_out/zf_k3_raw.csv 2>/dev/null && python tools/update_k3_traffic.py gpurun_out/zf_k3_raw.csv | tail -4; cp profiles/k3_dram_traffic.json gpurun_out/zf_k3_dram_traffic.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/zf_bench.json 2> gpurun_out/zf_bench.err; cut -c1-300 gpurun_out/zf_bench.json; tail -2 gpurun_out/zf_bench.err
timeout 600 ncu --nvtx --nvtx-include "gdb_timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/zf_launches_dtu.csv python bench.py --steps 1 --warmup 3 --lean > gpurun_out/zf_ncu_launches.log 2>&1
export GDB_SKIP_DIGEST_CHECK=1 GDB_K3_TRUTH_CACHE=/tmp/k3truth GDB_K3_PRECS=1,2
for rep in 1 2; do
  for w in dtu nerf llff; do
    pr=1,2; [ $w = nerf ] && pr=1
    timeout 120 python tools/bench_k3.py --workload $w --precisions $pr --iters 10 2>&1 | grep precision | sed "s/^/default /"
    timeout 120 python tools/bench_k3.py --workload $w --precisions $pr --iters 10 --lib gdb_nerf_b200/variants/lib_exact.so 2>&1 | grep precision | sed "s/^/exact   /"
  done
done | tee gpurun_out/zf_k3_default_vs_exact.log
for w in dtu nerf; do
  timeout 300 python tools/k3_errors.py $w 2>&1 | grep "precision" | sed "s/^/default /"
  GDB_K3_LIB=gdb_nerf_b200/variants/lib_exact.so timeout 300 python tools/k3_errors.py $w 2>&1 | grep "precision" | sed "s/^/exact   /"
done | tee -a gpurun_out/zf_k3_default_vs_exact.log
unset GDB_SKIP_DIGEST_CHECK
timeout 600 python bench.py --workload nerf --steps 10 --warmup 3 --lean > gpurun_out/zf_bench_nerf.json 2>> gpurun_out/zf_bench.err; cut -c1-200 gpurun_out/zf_bench_nerf.json
timeout 600 python bench.py --workload llff --steps 10 --warmup 3 --lean > gpurun_out/zf_bench_llff.json 2>> gpurun_out/zf_bench.err; cut -c1-200 gpurun_out/zf_bench_llff.json
