#!/bin/bash
# round 2, GPU call Y (1 GPU): colour pass of the 4x4 kernel with 2 / 4 (row, ray) items in flight
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export GDB_SKIP_DIGEST_CHECK=1
for rep in 1 2; do
  timeout 120 python tools/bench_k3.py --workload nerf --precisions 1 --iters 10 2>&1 | grep precision | sed "s/^/base   /"
  for v in p6u2 p6u4; do
    timeout 120 python tools/bench_k3.py --workload nerf --precisions 1 --iters 10 --lib gdb_nerf_b200/variants/lib_$v.so 2>&1 | grep precision | sed "s/^/$v   /"
  done
done | tee gpurun_out/y_k3_p6u.log
