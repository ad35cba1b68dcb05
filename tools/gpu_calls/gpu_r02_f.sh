#!/bin/bash
# round 2, GPU call F (N GPUs, N = $1): the multi-GPU matrix (BASELINE configs 2-5 at N = 1..$1) in ONE torchrun, then bench.py at N
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/f_topo_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  tools/multi_gpu_matrix.py --steps 20 --out gpurun_out/r02_matrix_${N}gpu.jsonl > gpurun_out/f_matrix_$N.log 2>&1
echo "matrix rc=$?"; grep -c . gpurun_out/r02_matrix_${N}gpu.jsonl; cut -c1-420 gpurun_out/r02_matrix_${N}gpu.jsonl; tail -5 gpurun_out/f_matrix_$N.log | cut -c1-300
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/f_bench_$N.json 2> gpurun_out/f_bench_$N.err
echo "bench rc=$?"; grep '^{' gpurun_out/f_bench_$N.json | cut -c1-1500; grep -c "NCCL INFO" gpurun_out/f_bench_$N.err; grep -m3 "NVLS\|nRanks" gpurun_out/f_bench_$N.err | cut -c1-200
