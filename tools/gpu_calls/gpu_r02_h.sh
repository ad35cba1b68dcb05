#!/bin/bash
# round 2, GPU call H (1 GPU): K3 generation-4 variants (taps in flight, tile slots per SM, the shared-memory bound)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for var in base pb ng3 ng3pb ldsbound; do
  for wl in dtu nerf llff; do
    GDB_K3_VARIANT=$var python tools/bench_k3.py --workload $wl --precisions 1 --iters 8 2>&1 | grep precision | sed "s/^/$var /"
  done
done | tee gpurun_out/h_bench_k3_variants.log
GDB_K3_VARIANT=ng3pb ncu --set full --import-source on --clock-control none -k regex:render_tc3 --launch-skip 2 -c 1 -f -o gpurun_out/prof_k3_g4_ng3pb_dtu python tools/bench_k3.py --workload dtu --precisions 1 --iters 1 > gpurun_out/h_ncu.log 2>&1
GDB_K3_VARIANT=ldsbound ncu --set full --import-source on --clock-control none -k regex:render_tc3 --launch-skip 2 -c 1 -f -o gpurun_out/prof_k3_g4_ldsbound_dtu python tools/bench_k3.py --workload dtu --precisions 1 --iters 1 >> gpurun_out/h_ncu.log 2>&1
