#!/bin/bash
# round 2, GPU call J (2 GPUs): K3 generation 4 with padded chunks (parity, A/B); train section of the matrix at N = 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_network_gpu.py -x -q -m gpu 2>&1 | tail -4 | cut -c1-250
for wl in dtu nerf llff; do python tools/bench_k3.py --workload $wl --precisions 5,1 --iters 8 2>&1 | grep precision; done | tee gpurun_out/j_bench_k3_padded.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  tools/multi_gpu_matrix.py --steps 10 --sections train --workloads dtu --out gpurun_out/j_matrix_2gpu.jsonl > gpurun_out/j_matrix_2.log 2>&1
echo "matrix rc=$?"; cut -c1-900 gpurun_out/j_matrix_2gpu.jsonl; grep -v "^$" gpurun_out/j_matrix_2.log | grep -v '^{' | tail -6 | cut -c1-300
