#!/bin/bash
# round 2, GPU call T (1 GPU): K3 experiments as alternative builds (tools/build_variant.sh): probe hint on the mbarrier wait, packed FFMA2 in the
# ReLU-dot epilogues, tiles handed out by an atomic counter, the descriptor loop rolled over views
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export GDB_SKIP_DIGEST_CHECK=1
for rep in 1 2; do
for wl in dtu nerf; do
  timeout 120 python tools/bench_k3.py --workload $wl --precisions 1 --iters 10 2>&1 | grep precision | sed "s/^/base   /"
  for v in ${VARIANTS:-spin fma2 dyn p1roll}; do
    timeout 120 python tools/bench_k3.py --workload $wl --precisions 1 --iters 10 --lib gdb_nerf_b200/variants/lib_$v.so 2>&1 | grep precision | sed "s/^/$v   /"
  done
done
done | tee gpurun_out/t_k3_variants.log
