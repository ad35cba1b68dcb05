#!/bin/bash
# round 2, GPU call ZM (1 GPU): the whole GPU suite at the final HEAD (151 tests) and smoke
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/zm_pytest_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/zm_smoke.log
