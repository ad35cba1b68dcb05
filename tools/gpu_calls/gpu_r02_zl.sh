#!/bin/bash
# round 2, GPU call ZL (1 GPU): benched-mode / fp32-class end-to-end error against the oracle at full size for nerf and llff
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 220 python tools/benched_parity.py nerf llff 2>&1 | tail -6 | cut -c1-900 | tee gpurun_out/zl_parity.log
