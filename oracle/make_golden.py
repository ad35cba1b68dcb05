#!/usr/bin/env python
"""Golden-vector generator - TEST INFRASTRUCTURE, runs only in the build container.

Imports the UNMODIFIED reference from /root/reference (read-only), with three
stand-ins on sys.path for modules that do not exist offline (``imp``,
``nvdiffrast``, ``nerfacc``; see oracle/ref_shims), feeds it the deterministic
synthetic workloads of ``gdb_nerf_b200.synthetic`` and dumps every intermediate
of the star-marked path into ``tests/golden/<case>.npz``.

The reference cannot travel to the GPU box, so the vectors are committed.
Re-generate with:   python oracle/make_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("GDB_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "ref_shims"), REF, ROOT]

from gdb_nerf_b200.config import make_cfg  # noqa: E402
from gdb_nerf_b200.synthetic import make_batch, synth_state_dict  # noqa: E402

CASES = {
    # name: recipe, overrides, B, V, H, W, near, far, focal, images, tilt, train
    "dtu_b2": dict(recipe="dtu_eval", B=1, V=3, H=32, W=32, near=425.0, far=905.0, focal=90.0, images="noise", tilt=0.0, train=False),
    "nerf_b4": dict(recipe="nerf_eval_4x4", B=2, V=3, H=32, W=32, near=2.5, far=5.5, focal=44.0, images="smooth", tilt=0.06, train=False),
    "train_b2": dict(recipe="dtu_pretrain", B=2, V=2, H=32, W=32, near=425.0, far=905.0, focal=90.0, images="smooth", tilt=0.04, train=True),
}


def _np(t):
    if isinstance(t, (list, tuple)):
        return [_np(x) for x in t]
    return t.detach().cpu().numpy()


def run_case(name: str, spec: dict) -> None:
    import networks.gdb_nerf.network as ref_network
    import networks.gdb_nerf.depth_net as ref_depth
    import networks.gdb_nerf.utils as ref_utils

    cfg = make_cfg(spec["recipe"])
    torch.manual_seed(0)
    net = ref_network.Network(cfg)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(synth_state_dict(shapes, seed=1), strict=True)
    net.train(spec["train"])
    batch = make_batch(spec["B"], spec["V"], spec["H"], spec["W"], spec["near"], spec["far"], spec["focal"],
                       seed=3, images=spec["images"], tilt=spec["tilt"])

    rec = {}
    stage = {"warp": 0, "reg": 0, "dv": 0}

    def wrap(owner, attr, fn):
        orig = getattr(owner, attr)

        def inner(*a, **k):
            out = orig(*a, **k)
            fn(a, k, out)
            return out

        setattr(owner, attr, inner)
        return orig

    undo = []

    def on_warp(a, k, out):
        i = stage["warp"]; stage["warp"] += 1
        src_feat, src_exts, src_ints, tar_exts, tar_ints, depth_values, inv = a
        rec[f"s{i}_src_feat"] = _np(src_feat)
        rec[f"s{i}_src_ints"] = _np(src_ints)
        rec[f"s{i}_tar_ints"] = _np(tar_ints)
        rec[f"s{i}_depth_values"] = _np(depth_values.contiguous())
        rec[f"s{i}_variance"] = _np(out)

    def on_reg(a, k, out):
        i = stage["reg"]; stage["reg"] += 1
        depth_values, prob, ci_scale, inv = a
        rec[f"s{i}_prob"] = _np(prob)
        rec[f"s{i}_depth"] = _np(out[0])
        rec[f"s{i}_ci"] = _np(out[1])

    def on_dv(a, k, out):
        i = stage["dv"]; stage["dv"] += 1
        rec[f"s{i}_range_in"] = _np(a[0].contiguous())

    undo.append((ref_depth, "build_feature_volume", wrap(ref_depth, "build_feature_volume", on_warp)))
    undo.append((ref_depth, "depth_regression", wrap(ref_depth, "depth_regression", on_reg)))
    undo.append((ref_depth, "get_depth_values", wrap(ref_depth, "get_depth_values", on_dv)))

    def on_sample(a, k, out):
        rec["depth_range"] = _np(a[0]); rec["vol_range"] = _np(a[1])
        for key, val in zip(("rays_xyz", "uvd", "z_vals", "ball_radii", "indices", "samples_per_batch", "samples_per_bundle"), out):
            rec[key] = _np(val)

    def on_encode(a, k, out):
        rec["tex_nchw"] = _np(a[1]); rec["feat_volume"] = _np(a[2])
        rec["rgbs_feat_dir"] = _np(out[0]); rec["vox_feat"] = _np(out[1])

    def on_mlp(a, k, out):
        rec["sigma"] = _np(out[0]); rec["feat"] = _np(out[1])

    def on_weights(a, k, out):
        rec["weights"] = _np(out[0])

    def on_acc(a, k, out):
        rec["bundle_feat"] = _np(out[0]); rec["bundle_depth"] = _np(out[1]); rec["bundle_opacity"] = _np(out[2])

    wrap(net.sampler, "sample", on_sample)
    wrap(net.sampler, "encode", on_encode)
    wrap(net.nerf, "forward", on_mlp)
    undo.append((ref_utils, "render_weight_from_density", wrap(ref_utils, "render_weight_from_density", on_weights)))
    undo.append((ref_utils, "accumulate_value_along_rays", wrap(ref_utils, "accumulate_value_along_rays", on_acc)))

    def on_coarse(a, k, out):
        rec["coarse_rays"] = _np(a[0]); rec["coarse_tex_nchw"] = _np(a[2]); rec["coarse_rgb"] = _np(out)

    if spec["train"]:
        wrap(net.depth_net, "_render_rays", on_coarse)

    with torch.no_grad():
        ret, mvs_depths, blend = net(batch)
    for k, v in ret.items():
        rec["ret_" + k] = _np(v)
    for i, d in enumerate(mvs_depths):
        rec[f"mvs_depth_{i}"] = _np(d)
    for i, d in enumerate(blend):
        rec[f"blend_rgb_{i}"] = _np(d)

    # --- second pass through the sampler with an injected, narrow depth range so that
    # every adaptive count 1..max occurs (random weights always saturate at max) ---
    b = cfg.nerf.bundle_size
    g = torch.Generator().manual_seed(11)
    dr = torch.from_numpy(rec["depth_range"]); vr = torch.from_numpy(rec["vol_range"])
    min_iv = (spec["far"] - spec["near"]) / cfg.nerf.global_num_depth
    mid = 0.5 * (vr[:, :1] + vr[:, 1:])
    half = torch.rand(mid.shape, generator=g) * (0.55 * cfg.nerf.max_num_samples * min_iv)
    inj = torch.cat((mid - half, mid + half), 1)
    first = dict(rec)
    net.sampler.build_rays(batch["tar_views"]["extrinsics"], batch["tar_views"]["intrinsics"], (spec["H"], spec["W"]),
                           batch["near_far"][:, 0], batch["near_far"][:, 1])
    with torch.no_grad():
        rays_xyz, uvd, z_vals, ball, idx, per_batch, per_bundle = net.sampler.sample(
            inj, vr, b, cfg.nerf.max_num_samples, cfg.mvs.inv_depth[-1], True)
        rfd, vox = net.sampler.encode(batch["src_views"]["rgb"], torch.from_numpy(first["tex_nchw"]),
                                      torch.from_numpy(first["feat_volume"]), rays_xyz, uvd, ball,
                                      batch["src_views"]["extrinsics"], batch["src_views"]["intrinsics"],
                                      batch["tar_views"]["extrinsics"], per_batch)
        net.render_bundles(rfd, vox, z_vals, idx, per_bundle)
    # the wrappers recorded the second pass over the first-pass keys: file them under inj_* and restore
    for k in list(rec):
        if k in first and rec[k] is not first[k]:
            rec["inj_" + k] = rec[k]
            rec[k] = first[k]
    # inputs + the MLP parameters (the only weights the star path owns)
    rec["in_rgb"] = _np(batch["src_views"]["rgb"])
    rec["in_src_exts"] = _np(batch["src_views"]["extrinsics"]); rec["in_src_ints"] = _np(batch["src_views"]["intrinsics"])
    rec["in_tar_exts"] = _np(batch["tar_views"]["extrinsics"]); rec["in_tar_ints"] = _np(batch["tar_views"]["intrinsics"])
    rec["in_near_far"] = _np(batch["near_far"])
    for k, v in net.nerf.state_dict().items():
        rec["mlp_" + k] = _np(v)
    if spec["train"]:
        for k, v in net.depth_net.nerfs[0].state_dict().items():
            rec["coarse_mlp_" + k] = _np(v)
    rec["state_dict_keys"] = np.array(sorted(shapes))
    rec["state_dict_shapes"] = np.array([",".join(map(str, shapes[k])) for k in sorted(shapes)])

    for owner, attr, orig in undo:
        setattr(owner, attr, orig)
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, **rec)
    hist = np.bincount(rec["inj_samples_per_bundle"].astype(np.int64))
    print(f"{name}: {len(rec)} arrays, {os.path.getsize(path) / 1e6:.2f} MB, S={rec['indices'].shape[0]}, "
          f"injected S={rec['inj_indices'].shape[0]} count hist={hist.tolist()}")


GRAD_KEYS = ("nerf.", "depth_net.nerfs.0.", "depth_net.cost_regs.0.prob_head.weight", "depth_net.cost_regs.1.prob_head.weight",
             "depth_net.cost_regs.1.feat_head.weight", "depth_net.cost_regs.1.conv0.0.weight", "feature_net.conv0.0.0.weight",
             "feature_net.out1.weight", "feature_net.out0.weight", "upsampler.in_conv.weight")


def run_grad_case(name: str, spec: dict) -> None:
    """Training step of the UNMODIFIED reference under autograd (train mode, batch-norm batch statistics):
    loss = mean(rgb^2) + sum_i mean(blend_i^2), gradients of a representative parameter subset -> <name>_grad.npz."""
    import networks.gdb_nerf.network as ref_network

    cfg = make_cfg(spec["recipe"])
    torch.manual_seed(0)
    net = ref_network.Network(cfg)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(synth_state_dict(shapes, seed=1), strict=True)
    net.train(True)
    batch = make_batch(spec["B"], spec["V"], spec["H"], spec["W"], spec["near"], spec["far"], spec["focal"],
                       seed=3, images=spec["images"], tilt=spec["tilt"])
    ret, mvs_depths, blend = net(batch)
    loss = ret["rgb"].square().mean() + sum(b.square().mean() for b in blend)
    loss.backward()
    rec = {"loss": np.array(float(loss)), "ret_rgb": _np(ret["rgb"]), "ret_nerf_depth": _np(ret["nerf_depth"])}
    for i, b in enumerate(blend):
        rec[f"blend_rgb_{i}"] = _np(b)
    for k, p in net.named_parameters():
        if any(k.startswith(pref) for pref in GRAD_KEYS):
            rec["grad_" + k] = _np(p.grad) if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
    path = os.path.join(ROOT, "tests", "golden", name + "_grad.npz")
    np.savez_compressed(path, **rec)
    print(f"{name}_grad: loss={float(loss):.6f}, {len(rec)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    only = sys.argv[1:]
    for case, spec in CASES.items():
        if not only or case in only:
            run_case(case, spec)
        if spec["train"] and (not only or case + "_grad" in only):
            run_grad_case(case, spec)
