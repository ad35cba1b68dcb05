#!/bin/bash
# round 2, GPU call ZJ (1 GPU): ncu --set full of the final 4x4 kernel inside bench.py --workload nerf
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:render_tc2 --launch-skip 3 -c 1 -f -o gpurun_out/prof_k3_s4_nerf python bench.py --workload nerf --steps 1 --warmup 3 --lean > gpurun_out/zj_ncu_nerf.log 2>&1
ls -la gpurun_out/prof_k3_s4_nerf.ncu-rep
