#!/usr/bin/env python
"""Times K1 (homography warp + variance, gdb_warp_variance_fwd) alone at the two cascade stages of a workload, L2 flushed between
launches, and A/Bs the kernel variants the library selects by environment (read per call):

    default              second generation: 8 channels per thread, 256-bit loads
    GDB_K1_VARIANT=1     the right column of a pixel's 2x2 footprint taken from the x-adjacent lane by shuffle where it coincides
                         with that lane's left column
    GDB_K1_VARIANT=2     every lane projects its pixel into all views itself (no descriptor shuffles)
    GDB_K1_VARIANT=3     both
    GDB_K1_VARIANT=4     vertical pixel pairs per thread: the lower pixel's top footprint row taken from the upper pixel's registers
                         where the two coincide (6 loads instead of 8 per pair and view)
all of them must be bit-identical to the default (also checked on an odd-sized, ragged map)

    python tools/bench_k1.py [--workload dtu|llff|nerf] [--B 8] [--iters 20]
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gdb_nerf_b200 import ops
from gdb_nerf_b200.synthetic import camera_rig

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="dtu"); ap.add_argument("--B", type=int, default=8)
ap.add_argument("--V", type=int, default=3); ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
H, W, near, far, focal, stages = {"dtu": (512, 640, 425.0, 905.0, 1446.0, ((0.125, 32, 64), (0.5, 16, 8))),
                                  "llff": (640, 960, 2.0, 12.0, 850.0, ((0.125, 32, 36), (0.5, 16, 8))),
                                  "nerf": (800, 800, 2.5, 5.5, 1111.1, ((0.125, 32, 64), (0.25, 32, 8)))}[a.workload]
dev = "cuda"
g = torch.Generator().manual_seed(0)
rig = camera_rig(a.B, a.V, H, W, near, far, focal, tilt=0.02)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for si, (scale, C, D) in enumerate(stages):
    Hs, Ws = int(H * scale), int(W * scale)
    feat = (torch.randn(a.B, a.V, Hs, Ws, C, generator=g) * 0.5).to(dev)
    proj = ops.homography_mats(rig["src_exts"].to(dev), rig["src_ints"].to(dev), rig["tar_exts"].to(dev), rig["tar_ints"].to(dev), scale, scale)
    if si == 0:
        rng = torch.tensor([[near, far]] * a.B).view(a.B, 2, 1, 1).to(dev)
    else:   # a narrow per-pixel interval around a smooth depth map, as the first stage hands it on
        yy, xx = torch.meshgrid(torch.linspace(0, 1, Hs), torch.linspace(0, 1, Ws), indexing="ij")
        mid = near + (far - near) * (0.35 + 0.3 * (0.5 + 0.5 * torch.sin(6 * xx) * torch.cos(5 * yy)))
        half = (far - near) / 64 * 1.5
        rng = torch.stack((mid - half, mid + half)).unsqueeze(0).repeat(a.B, 1, 1, 1).to(dev)
    nbytes = feat.numel() * 4 + a.B * D * Hs * Ws * C * 4
    ref = None
    for name, env in (("default", {}), ("xshare", {"GDB_K1_VARIANT": "1"}), ("noshfl", {"GDB_K1_VARIANT": "2"}), ("xshare+noshfl", {"GDB_K1_VARIANT": "3"}), ("vpair", {"GDB_K1_VARIANT": "4"})):
        os.environ.pop("GDB_K1_VARIANT", None)
        os.environ.update(env)
        ts = []
        for i in range(a.iters + 3):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); out = ops.warp_variance(feat, proj, rng, D, Hs, Ws, False, out_channels_last=True, depth_folded=(si == 1)); e.record()
            torch.cuda.synchronize()
            if i >= 3: ts.append(s.elapsed_time(e))
        ms = sum(ts) / len(ts)
        same = "" if ref is None else f"  bit-identical to default: {torch.equal(out, ref)}"
        if ref is None: ref = out.clone()
        print(f"K1 {a.workload} stage {si} (C={C} D={D} {Hs}x{Ws}, {a.B} views) {name}: {ms:.4f} ms (min {min(ts):.4f}) = "
              f"{nbytes / ms / 1e6:.0f} GB/s algorithmic{same}", flush=True)
    os.environ.pop("GDB_K1_VARIANT", None)

# odd-sized ragged map, per-pixel ranges, every variant against the default
Hs, Ws, C, D = 37, 53, 16, 5
feat = (torch.randn(2, a.V, Hs, Ws, C, generator=g) * 0.5).to(dev)
rig2 = camera_rig(2, a.V, Hs * 2, Ws * 2, near, far, focal * Hs * 2 / H, tilt=0.05)
proj = ops.homography_mats(rig2["src_exts"].to(dev), rig2["src_ints"].to(dev), rig2["tar_exts"].to(dev), rig2["tar_ints"].to(dev), 0.5, 0.5)
mid = near + (far - near) * (0.2 + 0.6 * torch.rand(2, 1, Hs, Ws, generator=g))
rng = torch.cat((mid * 0.97, mid * 1.03), 1).to(dev)
outs = []
for v in range(5):
    os.environ["GDB_K1_VARIANT"] = str(v)
    outs.append([ops.warp_variance(feat, proj, rng, D, Hs, Ws, False, out_channels_last=True, depth_folded=f).clone() for f in (False, True)])
os.environ.pop("GDB_K1_VARIANT", None)
print("K1 odd-sized map (37x53, C=16, D=5): variants 1-4 bit-identical to default:",
      [all(torch.equal(x, y) for x, y in zip(outs[v], outs[0])) for v in range(1, 5)], flush=True)
