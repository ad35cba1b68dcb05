#!/bin/bash
# round 2, GPU call ZK (1 GPU): feature fetch with the second mip level blended branch-free (-DGDB_X_MIXSEL, an alternative build)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export GDB_SKIP_DIGEST_CHECK=1
for rep in 1 2; do
  for w in dtu nerf llff; do
    timeout 120 python tools/bench_k3.py --workload $w --precisions 1 --iters 10 2>&1 | grep precision | sed "s/^/base   /"
    timeout 120 python tools/bench_k3.py --workload $w --precisions 1 --iters 10 --lib gdb_nerf_b200/variants/lib_mixsel.so 2>&1 | grep precision | sed "s/^/mixsel /"
  done
done | tee gpurun_out/zk_k3_mixsel.log
