#!/bin/bash
# round 2, GPU call ZG (1 GPU): the one failing test of call ZF, with its assertion message
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "batched_gather_kernel_bit_identical" 2>&1 | grep -v "^$" | tail -40 | cut -c1-400 | tee gpurun_out/zg_pytest.log
