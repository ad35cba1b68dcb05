#!/bin/bash
# round 2, GPU call Q (8 GPUs): bench.py at N = 8, end to end eager vs CUDA-graph replay
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/q_bench_8.json 2> gpurun_out/q_bench_8.err
echo "bench rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/q_bench_8.json'):
    if l.startswith('{'):
        d=json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', json.dumps(d['e2e'])[:200]); print(d['e2e']['forward'], d['e2e']['eager'], d['e2e']['cuda_graph'])
PY
tail -3 gpurun_out/q_bench_8.err | cut -c1-300
