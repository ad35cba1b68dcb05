#!/bin/bash
# round 2, GPU call E (1 GPU): all GPU tests after the FlatAdam test fix; 1-GPU smoke of the multi-GPU matrix tool; smoke()
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/e_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/e_smoke.log
timeout 600 python tools/multi_gpu_matrix.py --steps 3 --out gpurun_out/e_matrix_1gpu.jsonl > gpurun_out/e_matrix.log 2>&1; tail -30 gpurun_out/e_matrix.log | cut -c1-700
