#!/bin/bash
# round 2, GPU call ZE (1 GPU): K3 with MUFU.RCP alone in the colour pass (-DGDB_X_RCPA) and additionally single-MUFU square roots /
# reciprocal in P0 / P1 (-DGDB_X_P1A) as alternative builds against the default build: time at the three workloads, error against the oracle
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/k3truth
export GDB_SKIP_DIGEST_CHECK=1 GDB_K3_TRUTH_CACHE=/tmp/k3truth GDB_K3_PRECS=1
for rep in 1 2; do
  for w in dtu nerf llff; do
    timeout 120 python tools/bench_k3.py --workload $w --precisions 1 --iters 10 2>&1 | grep precision | sed "s/^/base    /"
    for v in rcpa rcpap1; do
      timeout 120 python tools/bench_k3.py --workload $w --precisions 1 --iters 10 --lib gdb_nerf_b200/variants/lib_$v.so 2>&1 | grep precision | sed "s/^/$v /"
    done
  done
done | tee gpurun_out/ze_k3_rcpa.log
for w in dtu nerf; do
  timeout 300 python tools/k3_errors.py $w 2>&1 | grep "precision" | sed "s/^/base    /"
  for v in rcpa rcpap1; do
    GDB_K3_LIB=gdb_nerf_b200/variants/lib_$v.so timeout 300 python tools/k3_errors.py $w 2>&1 | grep "precision" | sed "s/^/$v /"
  done
done | tee -a gpurun_out/ze_k3_rcpa.log
