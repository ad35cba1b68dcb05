"""Deterministic synthetic workloads.

No dataset or checkpoint is reachable offline, so every test, the golden
generator and ``bench.py`` draw inputs from here.  Everything is a pure function
of its arguments: tensors are generated on the CPU with explicit generators so
the values are identical in this container and on the GPU box.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Mapping, Tuple

import torch

# Named workloads: recipe, H, W, near, far, focal  
WORKLOADS: Dict[str, Dict] = {
    "dtu": dict(recipe="dtu_eval", H=512, W=640, near=425.0, far=905.0, focal=1446.2),
    "llff": dict(recipe="llff_eval", H=640, W=960, near=2.0, far=12.0, focal=850.0),
    "nerf": dict(recipe="nerf_eval_4x4", H=800, W=800, near=2.5, far=5.5, focal=1111.1),
    "dtu_train": dict(recipe="dtu_pretrain", H=64, W=64, near=425.0, far=905.0, focal=1446.2),
}


def smooth_images(B: int, V: int, H: int, W: int, seed: int = 0, cutoff: int = 8) -> torch.Tensor:
    """Band-limited random images in [0, 1]: bilinear up-sampling of a coarse
    random grid plus a gentle ramp.  Used where parity has to be judged below the
    reference's own fp32 noise floor (white noise has unit gradient per pixel and
    turns 1e-5 px of coordinate rounding into 1e-5 of colour)."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.rand(B * V, 3, max(H // cutoff, 2), max(W // cutoff, 2), generator=g)
    img = torch.nn.functional.interpolate(coarse, size=(H, W), mode="bicubic", align_corners=True)
    return img.clamp_(0.0, 1.0).view(B, V, 3, H, W).contiguous()


def camera_rig(
    B: int, V: int, H: int, W: int, near: float, far: float, focal: float, view_offset: int = 0, tilt: float = 0.0
) -> Dict[str, torch.Tensor]:
    """Target camera at the origin looking down +z; V source cameras translated
    on a circle of radius 0.08*(near+far)/2 in the image plane (angle
    2*pi*v/V + view index).  ``tilt`` (radians) additionally yaws each source
    camera toward the scene centre so the homography is not a pure shift."""
    K = torch.tensor([[focal, 0.0, W / 2.0], [0.0, focal, H / 2.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    mid = 0.5 * (near + far)
    src_E = torch.zeros(B, V, 4, 4, dtype=torch.float64)
    for b in range(B):
        for v in range(V):
            ang = 2.0 * math.pi * v / V + (b + view_offset)
            tx, ty = 0.08 * mid * math.cos(ang), 0.08 * mid * math.sin(ang)
            E = torch.eye(4, dtype=torch.float64)
            if tilt != 0.0:
                yaw = tilt * math.cos(ang)
                pitch = -tilt * math.sin(ang)
                cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
                Ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=torch.float64)
                Rx = torch.tensor([[1, 0, 0], [0, cp, -sp], [0, sp, cp]], dtype=torch.float64)
                E[:3, :3] = Rx @ Ry
            E[0, 3], E[1, 3] = tx, ty
            src_E[b, v] = E
    tar_E = torch.eye(4, dtype=torch.float64).expand(B, 4, 4).clone()
    return {
        "src_exts": src_E.float(),
        "src_ints": K.float().expand(B, V, 3, 3).clone(),
        "tar_exts": tar_E.float(),
        "tar_ints": K.float().expand(B, 3, 3).clone(),
        "near_far": torch.tensor([[near, far]] * B, dtype=torch.float32),
    }


def make_batch(
    B: int, V: int, H: int, W: int, near: float, far: float, focal: float,
    seed: int = 0, images: str = "noise", view_offset: int = 0, tilt: float = 0.0,
) -> Dict:
    """Batch dict in the reference's layout (networks/gdb_nerf/network.py:96-103)."""
    if images == "noise":
        rgb = torch.rand(B, V, 3, H, W, generator=torch.Generator().manual_seed(seed))
    elif images == "noise8":
        # 8-bit white noise converted as the reference's loaders convert image files (dtu.py:135: astype(float32) / 255.)
        rgb = torch.randint(0, 256, (B, V, 3, H, W), generator=torch.Generator().manual_seed(seed), dtype=torch.uint8).float() / 255.0
    elif images == "smooth":
        rgb = smooth_images(B, V, H, W, seed)
    else:
        raise ValueError(images)
    rig = camera_rig(B, V, H, W, near, far, focal, view_offset, tilt)
    return {
        "src_views": {"rgb": rgb, "extrinsics": rig["src_exts"], "intrinsics": rig["src_ints"]},
        "tar_views": {"extrinsics": rig["tar_exts"], "intrinsics": rig["tar_ints"]},
        "near_far": rig["near_far"],
    }


def workload_batch(name: str, B: int = 1, V: int = 3, seed: int = 0, images: str = "noise", view_offset: int = 0) -> Dict:
    w = WORKLOADS[name]
    return make_batch(B, V, w["H"], w["W"], w["near"], w["far"], w["focal"], seed, images, view_offset)


def with_uint8_images(batch: Mapping) -> Dict:
    """The same batch with the source images as 8-bit samples (exact for images that came from 8-bit files or from
    ``images="noise8"``: x -> round(255 x)); ``Network.forward`` converts them back on the device, bit-identically."""
    out = {k: (dict(v) if isinstance(v, Mapping) else v) for k, v in batch.items()}
    rgb = batch["src_views"]["rgb"]
    u8 = torch.round(rgb * 255.0).clamp_(0, 255).to(torch.uint8)
    if not torch.equal(u8.float() / 255.0, rgb):
        raise ValueError("source images are not 8-bit samples / 255: an 8-bit copy would change the result")
    out["src_views"]["rgb"] = u8
    return out


def batch_to(batch: Mapping, device, non_blocking: bool = False) -> Dict:
    out = {}
    for k, v in batch.items():
        if isinstance(v, Mapping):
            out[k] = batch_to(v, device, non_blocking)
        elif torch.is_tensor(v):
            out[k] = v.to(device, non_blocking=non_blocking)
        else:
            out[k] = v
    return out


def synth_state_dict(shapes: Mapping[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Parameter values that depend only on (key, shape, seed) - not on module
    construction order - so the reference network (golden generator) and this
    package's network (tests) can be given identical weights without shipping a
    4 MB checkpoint.  Scale follows PyTorch's default fan-in uniform init."""
    out: Dict[str, torch.Tensor] = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[key] = torch.zeros(shape, dtype=torch.int64)
        elif leaf == "running_mean":
            out[key] = (torch.rand(shape, generator=g) - 0.5) * 0.1
        elif leaf == "running_var":
            out[key] = 0.75 + 0.5 * torch.rand(shape, generator=g)
        elif len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 1.0 / math.sqrt(fan_in)
            out[key] = (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
        elif leaf == "weight":  # norm scale
            out[key] = 0.9 + 0.2 * torch.rand(shape, generator=g)
        else:  # bias
            out[key] = (torch.rand(shape, generator=g) * 2.0 - 1.0) * 0.1
    return out
