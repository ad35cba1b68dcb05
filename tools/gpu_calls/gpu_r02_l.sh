#!/bin/bash
# round 2, GPU call L (8 GPUs): BASELINE configs 2-5 at N = 1, 2, 4, 8 in ONE torchrun (tools/multi_gpu_matrix.py)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/l_topo_8.txt 2>&1
NCCL_DEBUG=WARN timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  tools/multi_gpu_matrix.py --steps 10 --out gpurun_out/r02_matrix_8gpu.jsonl > gpurun_out/l_matrix_8.log 2>&1
echo "matrix rc=$?"; grep -c . gpurun_out/r02_matrix_8gpu.jsonl; cut -c1-330 gpurun_out/r02_matrix_8gpu.jsonl; grep -v "^$" gpurun_out/l_matrix_8.log | grep -v '^{' | tail -5 | cut -c1-300
