#!/usr/bin/env python
"""Error of every MLP arithmetic of the fused render kernel against the CPU oracle at full size (development tool).
Prints max / mean absolute error of the fine colours, the decoder features and the depth per precision."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_network_gpu as T
from gdb_nerf_b200 import ops
from oracle import gdb_oracle as O

if os.environ.get("GDB_K3_LIB"):          # alternative build of the library (tools/build_variant.sh), as tools/bench_k3.py --lib
    from gdb_nerf_b200 import _lib as _L
    _L.LIB_PATH = os.path.abspath(os.environ["GDB_K3_LIB"])
PRECS = tuple(int(x) for x in os.environ.get("GDB_K3_PRECS", "0,1,2,3").split(","))
DEV = "cuda"
for workload in sys.argv[1:] or ["dtu"]:
    cfg, w, rig, data, mlp, feat_dim = T._full_size_inputs(workload)
    b = cfg.nerf.bundle_size; H, W = w["H"], w["W"]; R = 3 * b * b
    cache = os.path.join(os.environ.get("GDB_K3_TRUTH_CACHE", ""), f"k3_truth_{workload}.pt") if os.environ.get("GDB_K3_TRUTH_CACHE") else ""
    if cache and os.path.exists(cache):          # the oracle's answer for these seeded inputs, computed by an earlier invocation of this tool
        truth = torch.load(cache)
    else:
        truth = O.render_bundles(mlp, feat_dim, data["rgb"], data["feat"], data["vol"], data["depth_range"], data["vol_range"],
                                 rig["src_exts"], rig["src_ints"], rig["tar_exts"], rig["tar_ints"], rig["near_far"], b,
                                 cfg.nerf.max_num_samples, cfg.nerf.global_num_depth, cfg.nerf.max_mipmap_level, False, True)
        if cache:
            torch.save({k: truth[k] for k in ("bundle_feat", "bundle_depth")}, cache)
    cam = ops.camera_block(rig["tar_exts"].to(DEV), rig["tar_ints"].to(DEV), rig["src_exts"].to(DEV), rig["src_ints"].to(DEV),
                           rig["near_far"].to(DEV), b, cfg.nerf.global_num_depth, False)
    src = ops.prepare_sources(data["feat"].to(DEV), data["rgb"].to(DEV), b, cfg.nerf.max_mipmap_level)
    vol_cl = ops.to_channels_last(data["vol"].to(DEV), 8)
    for prec in PRECS:
        out = ops.render_fused(src, vol_cl, data["depth_range"].to(DEV), data["vol_range"].to(DEV), cam,
                               ops.pack_mlp(mlp, feat_dim, device=DEV), 1, 3, H, W, b, cfg.nerf.max_num_samples, False, True, precision=prec)
        e = (out["feat"].cpu().double() - truth["bundle_feat"].double()).abs()
        ed = (out["depth"].cpu().double() - truth["bundle_depth"].double()).abs() / (w["far"] - w["near"])
        print(f"{workload} precision {prec}: fine rgb max {e[:, :R].max():.2e} mean {e[:, :R].mean():.2e} | decoder feat max {e[:, R:].max():.2e} "
              f"mean {e[:, R:].mean():.2e} | depth (normalised) max {ed.max():.2e} mean {ed.mean():.2e}")
