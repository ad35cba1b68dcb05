#!/bin/bash
# round 2, GPU call ZC (1 GPU): everything at HEAD - all GPU tests, smoke, the box's pinned copy rates, the bench line (now also timing
# the end-to-end step as a copy pipeline around one compute stream), lean bench lines of llff / nerf
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/zc_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/zc_smoke.log
timeout 120 python tools/pcie_probe.py 2>&1 | tail -1 | tee gpurun_out/zc_pcie_probe.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/zc_bench.json 2> gpurun_out/zc_bench.err; cut -c1-300 gpurun_out/zc_bench.json; tail -2 gpurun_out/zc_bench.err
timeout 600 python bench.py --workload nerf --steps 10 --warmup 3 --lean > gpurun_out/zc_bench_nerf.json 2>> gpurun_out/zc_bench.err; cut -c1-200 gpurun_out/zc_bench_nerf.json
timeout 600 python bench.py --workload llff --steps 10 --warmup 3 --lean > gpurun_out/zc_bench_llff.json 2>> gpurun_out/zc_bench.err; cut -c1-200 gpurun_out/zc_bench_llff.json
python - <<'P'
import json
for f in ("zc_bench", "zc_bench_nerf", "zc_bench_llff"):
    try:
        b = json.load(open(f"gpurun_out/{f}.json")); e = b["e2e"]
        print(f, "dev", round(b["ms_per_step"], 3), "e2e", round(e["ms_per_step"], 3), "eager", e["eager"]["ms_per_step"], "graph", e["cuda_graph"], "pipe", e["cuda_graph_copy_pipeline"])
    except Exception as ex:
        print(f, "unreadable", ex)
P
