#!/bin/bash
# round 2, GPU call I (2 GPUs): the sections of the multi-GPU matrix that died at N = 2 (train step capture, sweep), short watchdog
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  tools/multi_gpu_matrix.py --steps 10 --sections sweep,train --workloads nerf --out gpurun_out/i_matrix_2gpu.jsonl > gpurun_out/i_matrix_2.log 2>&1
echo "matrix rc=$?"; cut -c1-600 gpurun_out/i_matrix_2gpu.jsonl; grep -v "^$" gpurun_out/i_matrix_2.log | grep -v '^{' | tail -12 | cut -c1-300
python -m pytest tests/test_network_gpu.py -x -q -m gpu -k "sweep" 2>&1 | tail -3
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "render_fused" 2>&1 | tail -3
for wl in dtu nerf llff; do python tools/bench_k3.py --workload $wl --precisions 5,1 --iters 8 2>&1 | grep precision; done | tee gpurun_out/i_bench_k3_padded.log
