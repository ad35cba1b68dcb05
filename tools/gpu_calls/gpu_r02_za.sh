#!/bin/bash
# round 2, GPU call ZA (1 GPU): all GPU tests at HEAD; K1 alone, default against the three measured variants (GDB_K1_VARIANT: right column
# by shuffle from the x-adjacent lane / projections without descriptor shuffles / both; bit-identity checked by the tool) at the three
# workloads; the default bench line; lean bench lines of llff / nerf; ncu --set full of every K1 variant
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/za_pytest_all.log
for w in dtu llff nerf; do timeout 200 python tools/bench_k1.py --workload $w --iters 20 2>&1 | grep "^K1"; done | tee gpurun_out/za_k1_variants.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/za_bench.json 2> gpurun_out/za_bench.err; cut -c1-400 gpurun_out/za_bench.json; tail -2 gpurun_out/za_bench.err
timeout 600 python bench.py --workload llff --steps 10 --warmup 3 --lean > gpurun_out/za_bench_llff.json 2>> gpurun_out/za_bench.err; cut -c1-300 gpurun_out/za_bench_llff.json
timeout 600 python bench.py --workload nerf --steps 10 --warmup 3 --lean > gpurun_out/za_bench_nerf.json 2>> gpurun_out/za_bench.err; cut -c1-300 gpurun_out/za_bench_nerf.json
timeout 400 ncu --set full --import-source on --clock-control none -k regex:warp_variance8 -c 32 -f -o gpurun_out/prof_k1_variants python tools/bench_k1.py --workload dtu --iters 1 > gpurun_out/za_ncu_k1.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
