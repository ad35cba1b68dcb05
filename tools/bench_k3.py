#!/usr/bin/env python
"""Times the fused render kernel alone (fp32 SIMT and tcgen05 variants) on full-size synthetic inputs.
Development tool for ncu captures: python tools/bench_k3.py [--precision P] [--B 8] [--iters 5]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gdb_nerf_b200 import ops
from gdb_nerf_b200.config import make_cfg
from gdb_nerf_b200.nerf import NeRF
from gdb_nerf_b200.synthetic import WORKLOADS, camera_rig, smooth_images

ap = argparse.ArgumentParser()
ap.add_argument("--precision", type=int, default=-1)
ap.add_argument("--precisions", default="", help="comma-separated list, e.g. 4,1 (4 = round-1 kernel, 1 = batched-gather kernel)")
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--workload", default="dtu")
ap.add_argument("--V", type=int, default=3, help="source views (2, 3 or 4)")
ap.add_argument("--lib", default="", help="alternative build of the library (A/B measurements)")
ap.add_argument("--folded", action="store_true", help="volume in the depth-folded (B,Hb,Wb,D,12) layout of the 2-D cost-regularisation head")
ap.add_argument("--noisy-depth", action="store_true", help="per-bundle white-noise depth (adversarial: no coherence between neighbouring bundles) "
                "instead of the smooth depth field a cost-volume network produces")
ap.add_argument("--narrow", action="store_true", help="narrow depth ranges (adaptive counts 1..max) instead of saturated")
args = ap.parse_args()
if args.lib:
    from gdb_nerf_b200 import _lib as _L
    _L.LIB_PATH = os.path.abspath(args.lib)
w = WORKLOADS[args.workload]; cfg = make_cfg(w["recipe"]); b = cfg.nerf.bundle_size
H, W, V, B = w["H"], w["W"], args.V, args.B
Hb, Wb = H // b, W // b
dev = "cuda"
g = torch.Generator().manual_seed(0)
lvl = 0
while cfg.fpn.feat_scales[lvl] < 1.0 / b: lvl += 1
fd = cfg.fpn.feat_dims[lvl]
rig = camera_rig(B, V, H, W, w["near"], w["far"], w["focal"], tilt=0.03)
rgb = smooth_images(B, V, H, W).to(dev)
feat = (torch.randn(B, V, fd, Hb, Wb, generator=g) * 0.5).to(dev)
vol = (torch.randn(B, 8, 8, Hb, Wb, generator=g) * 0.5).to(dev)
min_iv = (w["far"] - w["near"]) / cfg.nerf.global_num_depth
if args.noisy_depth:
    mid = w["near"] + (w["far"] - w["near"]) * (0.3 + 0.4 * torch.rand(B, 1, Hb, Wb, generator=g))
else:       # smooth field: bicubic up-sampling of a coarse random grid
    coarse = torch.rand(B, 1, max(Hb // 32, 2), max(Wb // 32, 2), generator=g)
    mid = w["near"] + (w["far"] - w["near"]) * (0.3 + 0.4 * torch.nn.functional.interpolate(coarse, size=(Hb, Wb), mode="bicubic", align_corners=True).clamp(0, 1))
half = torch.rand(B, 1, Hb, Wb, generator=g) * (0.55 * cfg.nerf.max_num_samples * min_iv) if args.narrow else torch.full((B, 1, Hb, Wb), 4 * min_iv)
dr = torch.cat((mid - half, mid + half), 1).to(dev); vr = torch.cat((mid - 2.5 * min_iv, mid + 2.5 * min_iv), 1).to(dev)
torch.manual_seed(0)
mlp = ops.pack_mlp({k: v.detach() for k, v in NeRF(64, fd, 8, True).state_dict().items()}, fd, device=dev)
cam = ops.camera_block(rig["tar_exts"].to(dev), rig["tar_ints"].to(dev), rig["src_exts"].to(dev), rig["src_ints"].to(dev), rig["near_far"].to(dev), b, cfg.nerf.global_num_depth, False)
src = ops.prepare_sources(feat, rgb, b, cfg.nerf.max_mipmap_level)
vol_cl = ops.to_channels_last(vol, 8)
if args.folded:
    buf = torch.zeros(B, Hb, Wb, vol_cl.shape[1], 12, device=dev)
    buf[..., :8] = vol_cl.permute(0, 2, 3, 1, 4)
    vol_cl = buf[..., :8].permute(0, 3, 1, 2, 4)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
plist = [int(x) for x in args.precisions.split(",")] if args.precisions else ([0, 1, 2, 3, 4] if args.precision < 0 else [args.precision])
for prec in plist:
    ts = []
    for i in range(args.iters + 2):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = ops.render_fused(src, vol_cl, dr, vr, cam, mlp, B, V, H, W, b, cfg.nerf.max_num_samples, False, True, precision=prec, out_channels_last=True, pad_dec=True, dec_one=True)
        e.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(s.elapsed_time(e))
    ms = sum(ts) / len(ts)
    print(f"{args.workload} V={V} precision {prec} (GDB_K3_FB={os.environ.get('GDB_K3_FB', '-')}): {ms:.4f} ms per launch ({B} views) = {ms / B * 1e3:.1f} us/view, min {min(ts):.4f}", flush=True)
