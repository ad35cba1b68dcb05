#!/bin/bash
# round 2, GPU call O (8 GPUs): the 200-view sweep again (pinned-buffer pool), bench.py at N = 8 (three steps in flight end to end)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  tools/multi_gpu_matrix.py --steps 10 --sections sweep --workloads nerf --out gpurun_out/o_matrix_sweep_8gpu.jsonl > gpurun_out/o_matrix_8.log 2>&1
echo "matrix rc=$?"; cut -c1-700 gpurun_out/o_matrix_sweep_8gpu.jsonl
NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/o_bench_8.json 2> gpurun_out/o_bench_8.err
echo "bench rc=$?"; grep '^{' gpurun_out/o_bench_8.json | cut -c1-260; python - <<'PY'
import json
for l in open('gpurun_out/o_bench_8.json'):
    if l.startswith('{'):
        d=json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
