#!/usr/bin/env python
"""Pinned host <-> device copy rates of the box, one GPU: the denominator behind DESIGN.md section 7's reading of the end-to-end
numbers (bytes per step / step time).  Sizes are the benchmark's per-step uploads / downloads.  python tools/pcie_probe.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

dev = "cuda"
out = {}
for name, nbytes in (("dtu_d2h_44.6MB", 44_564_480), ("nerf_d2h_86MB", 86_016_000), ("dtu_h2d_23.6MB", 23_596_224), ("256MB", 256 << 20)):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    gpu = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for direction in ("d2h", "h2d"):
        ts = []
        for i in range(8):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            if direction == "d2h":
                host.copy_(gpu, non_blocking=True)
            else:
                gpu.copy_(host, non_blocking=True)
            e.record(); torch.cuda.synchronize()
            if i >= 2: ts.append(s.elapsed_time(e))
        out[f"{name}_{direction}_GBps"] = round(nbytes / (sum(ts) / len(ts)) / 1e6, 2)
# both directions at once on two streams (what three steps in flight do)
nbytes = 86_016_000
h1, h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.empty(nbytes, dtype=torch.uint8).pin_memory()
g1, g2 = torch.empty(nbytes, dtype=torch.uint8, device=dev), torch.empty(nbytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): h1.copy_(g1, non_blocking=True)
    with torch.cuda.stream(s2): g2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
out["86MB_both_directions_each_GBps"] = round(10 * nbytes / dt / 1e9, 2)
print(json.dumps(out))
