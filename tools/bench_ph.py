#!/usr/bin/env python
"""Times the fused probability head + depth-range kernel (K2) alone: hand-staged split kernel vs the TMA-fed second generation.
DTU stage-0 shape by default (8 views, D = 64 planes of 64x80 voxels x 8 channels).  python tools/bench_ph.py [--B 8 --D 64 --h 64 --w 80]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gdb_nerf_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8); ap.add_argument("--D", type=int, default=64)
ap.add_argument("--h", type=int, default=64); ap.add_argument("--w", type=int, default=80)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
dev = "cuda"
g = torch.Generator().manual_seed(0)
y = (torch.randn(a.B, 8, a.D, a.h, a.w, generator=g) * 0.7).to(dev).contiguous(memory_format=torch.channels_last_3d)
wt = (torch.randn(1, 8, 3, 3, 3, generator=g) * 0.2).to(dev)
rng = torch.tensor([[425.0, 905.0]] * a.B).view(a.B, 2, 1, 1).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
nbytes = y.numel() * 4
for name, kw in (("hand-staged split kernel", dict(tma=False)), ("TMA-fed kernel", dict(tma=True)), ("TMA-fed kernel, depth unsplit", dict(tma=True, split=False))):
    ts = []
    for i in range(a.iters + 3):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); out = ops.prob_head_depth_range(y, wt, rng, 1.0, True, **kw); e.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(s.elapsed_time(e))
    ms = sum(ts) / len(ts)
    print(f"prob head {name}: {ms:.4f} ms (min {min(ts):.4f}) for {a.B}x{a.D}x{a.h}x{a.w} voxels = {nbytes / ms / 1e6:.0f} GB/s of volume bytes", flush=True)
