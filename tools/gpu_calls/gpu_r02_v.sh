#!/bin/bash
# round 2, GPU call V (1 GPU): operand chunks 2048 + pad bytes apart (bank conflicts of the fetch lanes' stores), feature-fetch batching modes again under the dynamic tile assignment
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export GDB_SKIP_DIGEST_CHECK=1
for rep in 1 2; do
for wl in dtu llff nerf; do
  timeout 120 python tools/bench_k3.py --workload $wl --precisions 1 --iters 10 2>&1 | grep precision | sed "s/^/base   /"
  for v in pad64 pad80 pad32; do
    timeout 120 python tools/bench_k3.py --workload $wl --precisions 1 --iters 10 --lib gdb_nerf_b200/variants/lib_$v.so 2>&1 | grep precision | sed "s/^/$v   /"
  done
done
done | tee gpurun_out/v_k3_pad.log
for fb in 1 2; do GDB_K3_FB=$fb timeout 120 python tools/bench_k3.py --workload dtu --precisions 1 --iters 10 2>&1 | grep precision | sed "s/^/fb$fb /"; done | tee -a gpurun_out/v_k3_pad.log
