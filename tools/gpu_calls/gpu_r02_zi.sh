#!/bin/bash
# round 2, GPU call ZI (2 GPUs): the driver's multi-GPU launch of bench.py with the final code (torchrun, one rank per GPU), both arms
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/zi_bench_2gpu.json 2> gpurun_out/zi_bench_2gpu.err; cut -c1-300 gpurun_out/zi_bench_2gpu.json; tail -3 gpurun_out/zi_bench_2gpu.err
python - <<'P'
import json
b = json.load(open("gpurun_out/zi_bench_2gpu.json")); e = b["e2e"]
print("n_gpus", b["n_gpus"], "value", b["value"], "ms", b["ms_per_step"], "e2e", e["value"], e["ms_per_step"], e["forward"][:50]); print(e["eager"], e["cuda_graph"], e["cuda_graph_copy_pipeline"])
P
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | cut -c1-300 | tee gpurun_out/zi_bench_ref_2gpu.json
