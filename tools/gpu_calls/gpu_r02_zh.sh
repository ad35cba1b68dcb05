#!/bin/bash
# round 2, GPU call ZH (1 GPU): all GPU tests at HEAD (the relaxed variant-vs-default tolerance of the bit-identity test included);
# colour pass of the 4x4 kernel unrolled over 2 / 4 (row, ray) items per trip again, now that no branch sits between its gathers
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/zh_pytest_all.log
export GDB_SKIP_DIGEST_CHECK=1
for rep in 1 2; do
  timeout 120 python tools/bench_k3.py --workload nerf --precisions 1 --iters 10 2>&1 | grep precision | sed "s/^/base /"
  for v in p6u2 p6u4; do
    timeout 120 python tools/bench_k3.py --workload nerf --precisions 1 --iters 10 --lib gdb_nerf_b200/variants/lib_$v.so 2>&1 | grep precision | sed "s/^/$v /"
  done
done | tee gpurun_out/zh_k3_p6u.log
