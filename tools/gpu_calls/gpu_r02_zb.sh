#!/bin/bash
# round 2, GPU call ZB (1 GPU): K1 vertical-pair variant (GDB_K1_VARIANT=4) against the default and variants 1-3 at the three workloads
# (bit-identity checked by the tool, also on an odd-sized map), the K1 / network parity tests with the variant, ncu --set full of it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for w in dtu llff nerf; do timeout 200 python tools/bench_k1.py --workload $w --iters 20 2>&1 | grep "^K1"; done | tee gpurun_out/zb_k1_variants.log
GDB_K1_VARIANT=4 timeout 600 python -m pytest tests -q -m gpu -k "warp_variance or network_forward_matches_reference or benched_math_mode or view_counts" 2>&1 | tail -3 | cut -c1-300 | tee gpurun_out/zb_pytest_vpair.log
GDB_K1_VARIANT=4 timeout 600 python bench.py --steps 20 --warmup 3 --lean > gpurun_out/zb_bench_vpair.json 2> gpurun_out/zb_bench_vpair.err; cut -c1-300 gpurun_out/zb_bench_vpair.json
timeout 600 python bench.py --steps 20 --warmup 3 --lean > gpurun_out/zb_bench_default.json 2> gpurun_out/zb_bench_default.err; cut -c1-300 gpurun_out/zb_bench_default.json
timeout 300 ncu --set full --import-source on --clock-control none -k regex:vpair -c 4 -f -o gpurun_out/prof_k1_vpair python tools/bench_k1.py --workload dtu --iters 1 > gpurun_out/zb_ncu_k1.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
