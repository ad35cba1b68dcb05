"""CPU oracle for the GDB-NeRF per-target-view rendering path.

TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module, and only as the checker / the CPU
baseline.  The product (``gdb_nerf_b200``) never imports it and has no CPU
fallback.

What it is: a restatement, in plain CPU tensor arithmetic (explicit index
arithmetic for every gather - no ``grid_sample``, no third-party CUDA
libraries), of the algorithm the reference implements in
``networks/gdb_nerf/{depth_net,bundle_sampler,nerf,utils,network}.py``.  Each
function cites the reference lines it follows.  Every function takes the
arithmetic ``dtype`` from its inputs: run it in float32 to mirror the
reference's rounding, or float64 to get a "truth" both the reference and the
CUDA kernels are compared against ("noise floor").

Pinning: ``tests/test_oracle_golden.py`` checks every function below against
tensors dumped from the UNMODIFIED reference executed in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``).  Two arithmetic
boundaries of the reference live in third-party libraries that are neither
vendored nor pinned nor installable offline - ``nvdiffrast.torch.texture``
(bundle_sampler.py:355) and ``nerfacc.volrend`` (utils.py:35,110).  For those
the golden vectors were produced through stand-ins that restate the libraries'
published behaviour (``oracle/ref_shims``): PARITY UNPINNED at those two
boundaries (mip-mapped texture fetch, exclusive transmittance product); pinned
by the reference's own code everywhere else.
"""
from __future__ import annotations

import math
from typing import Dict, List, Mapping, NamedTuple, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------
def _linspace01(n: int, like: Tensor) -> Tensor:
    # torch.linspace(0, 1, n) is evaluated in float32 by the reference
    # (depth_net.py:419 passes no dtype) and only then promoted.
    return torch.linspace(0.0, 1.0, n, dtype=torch.float32).to(like.dtype)


def _pixel_centres(H: int, W: int, dtype) -> Tuple[Tensor, Tensor]:
    xs = torch.arange(W, dtype=dtype) + 0.5
    ys = torch.arange(H, dtype=dtype) + 0.5
    return xs.view(1, W).expand(H, W), ys.view(H, 1).expand(H, W)


def _gather_hw(img: Tensor, iy: Tensor, ix: Tensor) -> Tensor:
    """img (C,H,W); iy, ix integer tensors of any (equal) shape, in range.
    Returns (C, *shape)."""
    C, H, W = img.shape
    flat = img.reshape(C, H * W)
    return flat[:, (iy * W + ix).reshape(-1)].reshape(C, *ix.shape)


def bilinear_2d(img: Tensor, gx: Tensor, gy: Tensor, padding: str) -> Tensor:
    """Bilinear fetch with normalised coordinates in [-1, 1], align_corners =
    False, ``padding`` 'zeros' or 'border' - the semantics of
    ``F.grid_sample`` relied on at depth_net.py:472 and
    bundle_sampler.py:336.  img (C,H,W); gx, gy any shape -> (C, *shape)."""
    C, H, W = img.shape
    ix = ((gx + 1.0) * W - 1.0) / 2.0
    iy = ((gy + 1.0) * H - 1.0) / 2.0
    if padding == "border":
        ix = ix.clamp(0.0, W - 1.0)
        iy = iy.clamp(0.0, H - 1.0)
    x0 = ix.floor()
    y0 = iy.floor()
    x1 = x0 + 1.0
    y1 = y0 + 1.0
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    out = torch.zeros((C,) + tuple(gx.shape), dtype=img.dtype)
    for xs, ys, w in ((x0, y0, w_nw), (x1, y0, w_ne), (x0, y1, w_sw), (x1, y1, w_se)):
        inside = (xs >= 0) & (xs <= W - 1) & (ys >= 0) & (ys <= H - 1)
        xi = xs.clamp(0, W - 1).long()
        yi = ys.clamp(0, H - 1).long()
        val = _gather_hw(img, yi, xi)
        out = out + val * (w * inside.to(img.dtype))
    return out


def trilinear_border(vol: Tensor, gx: Tensor, gy: Tensor, gz: Tensor) -> Tensor:
    """3-D ``grid_sample(bilinear, border, align_corners=False)``
    (bundle_sampler.py:323).  vol (C,D,H,W); coordinates (S,) -> (S, C)."""
    C, D, H, W = vol.shape
    ix = (((gx + 1.0) * W - 1.0) / 2.0).clamp(0.0, W - 1.0)
    iy = (((gy + 1.0) * H - 1.0) / 2.0).clamp(0.0, H - 1.0)
    iz = (((gz + 1.0) * D - 1.0) / 2.0).clamp(0.0, D - 1.0)
    x0, y0, z0 = ix.floor(), iy.floor(), iz.floor()
    fx, fy, fz = ix - x0, iy - y0, iz - z0
    flat = vol.reshape(C, D * H * W)
    out = torch.zeros((C, gx.numel()), dtype=vol.dtype)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xs, ys, zs = x0 + dx, y0 + dy, z0 + dz
                w = (fx if dx else 1.0 - fx) * (fy if dy else 1.0 - fy) * (fz if dz else 1.0 - fz)
                inside = (xs <= W - 1) & (ys <= H - 1) & (zs <= D - 1)
                idx = (zs.clamp(0, D - 1).long() * H + ys.clamp(0, H - 1).long()) * W + xs.clamp(0, W - 1).long()
                out = out + flat[:, idx] * (w * inside.to(vol.dtype))
    return out.t().contiguous()


# --------------------------------------------------------------------------
# a1  depth hypotheses            depth_net.py:399-421
# --------------------------------------------------------------------------
def depth_hypotheses(depth_range: Tensor, num_depth: int, inv_depth: bool) -> Tensor:
    """(B,2,h,w) near/far -> (B,D,h,w) values, linear in depth or in disparity."""
    near, far = depth_range[:, :1], depth_range[:, -1:]
    if inv_depth:
        near, far = 1.0 / near, 1.0 / far
    steps = _linspace01(num_depth, depth_range).view(1, num_depth, 1, 1)
    return near + (far - near) * steps


# --------------------------------------------------------------------------
# a2  homography warp + variance  depth_net.py:424-476
# --------------------------------------------------------------------------
def homography_matrices(src_exts: Tensor, src_ints: Tensor, tar_exts: Tensor, tar_ints: Tensor) -> Tensor:
    """(B,V,3,4) matrices taking [x*depth, y*depth, depth, 1] of the target
    view to homogeneous source pixels (depth_net.py:453-457)."""
    src_proj = src_ints @ src_exts[..., :3, :]
    tar_proj = tar_ints @ tar_exts[..., :3, :]
    B = tar_proj.shape[0]
    full = torch.zeros(B, 4, 4, dtype=tar_proj.dtype)
    full[:, :3] = tar_proj
    full[:, 3, 3] = 1.0
    return src_proj @ torch.linalg.inv(full).unsqueeze(1)


def warp_variance(src_feat: Tensor, proj: Tensor, depth_values: Tensor, inv_depth: bool) -> Tensor:
    """src_feat (B,V,C,Hs,Ws), proj (B,V,3,4), depth_values (B,D,Ht,Wt) ->
    population variance over views of the warped features (B,C,D,Ht,Wt)."""
    B, V, C, Hs, Ws = src_feat.shape
    _, D, Ht, Wt = depth_values.shape
    dt = src_feat.dtype
    depth = 1.0 / depth_values if inv_depth else depth_values
    px, py = _pixel_centres(Ht, Wt, dt)
    out = torch.empty(B, C, D, Ht, Wt, dtype=dt)
    for b in range(B):
        warped = []
        for v in range(V):
            P = proj[b, v]
            # rotation part applied to (x, y, 1), scaled by depth, plus translation column
            rx = P[0, 0] * px + P[0, 1] * py + P[0, 2]
            ry = P[1, 0] * px + P[1, 1] * py + P[1, 2]
            rz = P[2, 0] * px + P[2, 1] * py + P[2, 2]
            X = rx.unsqueeze(0) * depth[b] + P[0, 3]
            Y = ry.unsqueeze(0) * depth[b] + P[1, 3]
            Z = (rz.unsqueeze(0) * depth[b] + P[2, 3]).clamp_min(1e-6)
            gx = 2.0 * (X / Z) / Ws - 1.0
            gy = 2.0 * (Y / Z) / Hs - 1.0
            warped.append(bilinear_2d(src_feat[b, v], gx, gy, "zeros"))  # (C,D,Ht,Wt)
        stack = torch.stack(warped, 0)
        mean = stack.mean(0, keepdim=True)
        out[b] = ((stack - mean) ** 2).mean(0)
    return out


# --------------------------------------------------------------------------
# a3  depth regression -> confidence interval     depth_net.py:479-514
# --------------------------------------------------------------------------
def depth_interval(depth_values: Tensor, depth_prob: Tensor, ci_scale: float, inv_depth: bool) -> Tuple[Tensor, Tensor]:
    """-> depth (B,1,h,w), ci (B,2,h,w) [near, far] in depth units."""
    mean = (depth_prob * depth_values).sum(1, keepdim=True)
    var = (depth_prob * (depth_values - mean) ** 2).sum(1, keepdim=True)
    half = ci_scale * var.clamp_min(1e-12).sqrt()
    first, last = depth_values[:, :1], depth_values[:, -1:]
    if inv_depth:
        ci = 1.0 / torch.cat((torch.minimum(mean + half, first), torch.maximum(mean - half, last)), 1)
        return 1.0 / mean, ci
    ci = torch.cat((torch.maximum(mean - half, first), torch.minimum(mean + half, last)), 1)
    return mean, ci


def upsample_bilinear(x: Tensor, out_h: int, out_w: int) -> Tensor:
    """``F.interpolate(mode='bilinear', align_corners=False)`` for (B,C,h,w)
    (depth_net.py:195-196, network.py:151-152,176-177)."""
    B, C, h, w = x.shape

    def axis(n_in, n_out):
        scale = n_in / n_out
        src = ((torch.arange(n_out, dtype=x.dtype) + 0.5) * scale - 0.5).clamp_min(0.0)
        i0 = src.floor().long().clamp(max=n_in - 1)
        i1 = (i0 + 1).clamp(max=n_in - 1)
        f = src - i0.to(x.dtype)
        return i0, i1, f

    y0, y1, fy = axis(h, out_h)
    x0, x1, fx = axis(w, out_w)
    rows = x[:, :, y0] * (1.0 - fy).view(1, 1, -1, 1) + x[:, :, y1] * fy.view(1, 1, -1, 1)
    return rows[:, :, :, x0] * (1.0 - fx) + rows[:, :, :, x1] * fx


# --------------------------------------------------------------------------
# a4/a5  target rays and bundle table     bundle_sampler.py:30-120
# --------------------------------------------------------------------------
class Rays(NamedTuple):
    origin: Tensor        # (B,3)
    dirs: Tensor          # (B,H,W,3) un-normalised
    uv: Tensor            # (H,W,2) in [-1,1]
    z_axis: Tensor        # (B,3)
    pixel_radius: Tensor  # (B,)


def target_rays(tar_exts: Tensor, tar_ints: Tensor, H: int, W: int) -> Rays:
    dt = tar_exts.dtype
    px, py = _pixel_centres(H, W, dt)
    uv = torch.stack((2.0 * px / W - 1.0, 2.0 * py / H - 1.0), -1)
    c2w = torch.linalg.inv(tar_exts)
    M = c2w[:, :3, :3] @ torch.linalg.inv(tar_ints)                      # (B,3,3)
    xyz = torch.stack((px.reshape(-1), py.reshape(-1), torch.ones(H * W, dtype=dt)), 1)
    dirs = (xyz @ M.transpose(-2, -1)).view(-1, H, W, 3)
    radius = 1.0 / torch.sqrt(tar_ints[:, 0, 0] * tar_ints[:, 1, 1] * math.pi)
    return Rays(c2w[:, :3, 3], dirs, uv, c2w[:, :3, 2], radius)


class Bundles(NamedTuple):
    origin: Tensor     # (B,3)
    ray_dirs: Tensor   # (B,Hb,Wb,3,b*b)  component-major, then by*b+bx
    uv: Tensor         # (Hb,Wb,2)
    disk_radius: Tensor  # (B,)
    cos: Tensor        # (B,Hb,Wb)


def assemble_bundles(rays: Rays, b: int) -> Bundles:
    B, H, W, _ = rays.dirs.shape
    Hb, Wb = H // b, W // b
    d = rays.dirs.view(B, Hb, b, Wb, b, 3)
    centre = d.mean(dim=(2, 4))                                           # (B,Hb,Wb,3)
    per_ray = d.permute(0, 1, 3, 5, 2, 4).reshape(B, Hb, Wb, 3, b * b)
    cos = (centre * rays.z_axis.view(B, 1, 1, 3)).sum(-1) / torch.linalg.vector_norm(centre, dim=-1)
    uv = rays.uv.view(Hb, b, Wb, b, 2).mean(dim=(1, 3))
    return Bundles(rays.origin, per_ray, uv, b * rays.pixel_radius, cos)


# --------------------------------------------------------------------------
# a6/a7  depth-guided sampling    bundle_sampler.py:122-265
# --------------------------------------------------------------------------
class Samples(NamedTuple):
    rays_xyz: Tensor          # (S,3,b*b)
    uvd: Tensor               # (S,3)
    z_vals: Tensor            # (S,)
    ball_radii: Tensor        # (S,)
    indices: Tensor           # (S,) int64, sorted
    samples_per_batch: Tensor  # (B,)
    samples_per_bundle: Tensor  # (NB,)


def sample_counts(near: Tensor, far: Tensor, min_interval: Tensor, max_num: int) -> Tensor:
    """Adaptive count per bundle (bundle_sampler.py:179).  Evaluated in the
    dtype given - callers that need the reference's exact integer result pass
    float32."""
    return torch.ceil((far - near).abs() / min_interval).clamp(1, max_num)


def sample_bundles(
    bundles: Bundles, depth_range: Tensor, vol_range: Tensor, scene_near: Tensor, scene_far: Tensor,
    global_num_depth: int, max_num: int, inv_depth: bool, adaptive: bool,
) -> Samples:
    """depth_range / vol_range (B,2,Hb,Wb).  Mirrors BundleSampler.sample."""
    B, _, Hb, Wb = depth_range.shape
    dt = depth_range.dtype
    NB = B * Hb * Wb
    if inv_depth:
        depth_range = 1.0 / depth_range
        vol_range = 1.0 / vol_range
        min_interval = (1.0 / scene_near - 1.0 / scene_far) / global_num_depth
    else:
        min_interval = (scene_far - scene_near) / global_num_depth
    near = depth_range[:, 0].reshape(NB)
    far = depth_range[:, 1].reshape(NB)
    vnear = vol_range[:, 0].reshape(NB)
    vfar = vol_range[:, 1].reshape(NB)
    min_interval = min_interval.view(B, 1).expand(B, Hb * Wb).reshape(NB)

    if adaptive:
        n = sample_counts(near, far, min_interval, max_num)               # float
        per_bundle = n
    else:
        n = torch.full((NB,), float(max_num), dtype=dt)
        per_bundle = torch.full((NB,), max_num, dtype=torch.int32)
    n_int = n.long()
    # row-major enumeration (bundle, slot) of the valid slots == boolean-mask order
    indices = torch.repeat_interleave(torch.arange(NB), n_int)
    first = torch.cumsum(n_int, 0) - n_int
    slot = torch.arange(indices.numel()) - first[indices]
    step = (far - near) / n                                                # (NB,)
    t0 = near[indices] + step[indices] * slot.to(dt)
    t1 = near[indices] + step[indices] * (slot + 1).to(dt)
    z = 0.5 * (t0 + t1)
    d = 2.0 * (z - vnear[indices]) / (vfar[indices] - vnear[indices]) - 1.0
    uv = bundles.uv.view(1, Hb, Wb, 2).expand(B, Hb, Wb, 2).reshape(NB, 2)[indices]
    uvd = torch.cat((uv, d.unsqueeze(1)), 1)
    if inv_depth:
        z = 1.0 / z
    batch_of = torch.div(indices, Hb * Wb, rounding_mode="floor")
    origin = bundles.origin[batch_of]                                       # (S,3)
    dirs = bundles.ray_dirs.reshape(NB, 3, -1)[indices]                     # (S,3,bb)
    rays_xyz = origin.unsqueeze(-1) + dirs * z.view(-1, 1, 1)
    centre = rays_xyz.mean(-1)
    dist = torch.linalg.vector_norm(centre - origin, dim=-1)
    cos = bundles.cos.reshape(NB)
    disk = bundles.disk_radius.view(B, 1).expand(B, Hb * Wb).reshape(NB)
    unit_ball = disk * cos / torch.sqrt((torch.sqrt((1.0 / cos.square() - 1.0).clamp_min(1e-12)) - disk).square() + 1.0)
    ball = dist * unit_ball[indices]
    per_batch = per_bundle.view(B, -1).sum(1)
    return Samples(rays_xyz, uvd, z, ball, indices, per_batch, per_bundle)


# --------------------------------------------------------------------------
# a8  feature+rgb texture and its mip chain   network.py:159-164, nvdiffrast
# --------------------------------------------------------------------------
def downsample_rgb(src_images: Tensor, Hb: int, Wb: int) -> Tensor:
    """``F.interpolate(size=(Hb,Wb), bilinear, align_corners=False)`` of
    (N,3,H,W) (network.py:163): no anti-aliasing."""
    return upsample_bilinear(src_images, Hb, Wb)


def build_mips(tex: Tensor, max_level: int) -> List[Tensor]:
    """tex (V,Hb,Wb,F) channels-last; 2x2 box chain (nvdiffrast MipBuildKernel)."""
    levels = [tex]
    for _ in range(max_level):
        t = levels[-1]
        V, h, w, F = t.shape
        if h % 2 or w % 2:
            raise ValueError("mip dims must stay even")
        q = t.view(V, h // 2, 2, w // 2, 2, F)
        levels.append(0.25 * ((q[:, :, 0, :, 0] + q[:, :, 0, :, 1]) + (q[:, :, 1, :, 0] + q[:, :, 1, :, 1])))
    return levels


def _clamped_bilerp(level: Tensor, u01: Tensor, v01: Tensor) -> Tensor:
    """level (h,w,F); uv in [0,1] (S,) -> (S,F).  nvdiffrast indexTextureLinear
    with boundary 'clamp' + bilerp."""
    h, w, F = level.shape
    u = (u01 * w - 0.5).clamp(0.0, w - 1.0)
    v = (v01 * h - 0.5).clamp(0.0, h - 1.0)
    iu0, iv0 = u.floor().long(), v.floor().long()
    iu1 = iu0 + (~((u == 0.0) | (u == w - 1.0))).long()
    iv1 = iv0 + (~((v == 0.0) | (v == h - 1.0))).long()
    fu = (u - iu0.to(u.dtype)).unsqueeze(1)
    fv = (v - iv0.to(v.dtype)).unsqueeze(1)
    flat = level.reshape(h * w, F)
    a00, a10 = flat[iv0 * w + iu0], flat[iv0 * w + iu1]
    a01, a11 = flat[iv1 * w + iu0], flat[iv1 * w + iu1]
    top = a00 + fu * (a10 - a00)
    bot = a01 + fu * (a11 - a01)
    return top + fv * (bot - top)


def mip_fetch(levels: Sequence[Tensor], view: int, u01: Tensor, v01: Tensor, lod_bias: Tensor) -> Tensor:
    """Linear-mipmap-linear fetch with bias-only level of detail -> (S,F)."""
    L = len(levels) - 1
    lod = lod_bias.clamp(0.0, float(L))
    l0 = lod.floor()
    l1 = (l0 + 1.0).clamp(max=float(L))
    frac = (lod - l0).unsqueeze(1)
    per_level = [_clamped_bilerp(lv[view], u01, v01) for lv in levels]
    a = torch.zeros_like(per_level[0])
    b = torch.zeros_like(per_level[0])
    for k, val in enumerate(per_level):
        a = torch.where((l0 == k).unsqueeze(1), val, a)
        b = torch.where((l1 == k).unsqueeze(1), val, b)
    return torch.where((lod > 0.0).unsqueeze(1), a + frac * (b - a), a)


# --------------------------------------------------------------------------
# a9  sphere-based multi-view encoding    bundle_sampler.py:267-371
# --------------------------------------------------------------------------
def _unit(x: Tensor) -> Tensor:
    # F.normalize(p=2, eps=1e-12)
    return x / torch.linalg.vector_norm(x, dim=-1, keepdim=True).clamp_min(1e-12)


def encode_samples(
    src_images: Tensor, tex_levels: Sequence[Sequence[Tensor]], feat_volume: Tensor, samples: Samples,
    src_exts: Tensor, src_ints: Tensor, tar_exts: Tensor, b: int,
) -> Tuple[Tensor, Tensor]:
    """src_images (B,V,3,H,W); tex_levels[b] = mip chain of (V,Hb,Wb,F);
    feat_volume (B,8,D,Hb,Wb) -> rgbs_feat_dir (V,S,3b^2+F+4), vox_feat (S,8)."""
    B, V, _, H, W = src_images.shape
    dt = src_images.dtype
    S, _, bb = samples.rays_xyz.shape
    Hb, Wb, F = tex_levels[0][0].shape[1:]
    tar_centre = torch.linalg.inv(tar_exts)[:, :3, 3]                       # (B,3)
    src_centre = torch.linalg.inv(src_exts)[..., :3, 3]                     # (B,V,3)
    Ks = src_ints.clone()
    Ks[..., :2, :] = Ks[..., :2, :] / b
    src_radius = 1.0 / torch.sqrt(Ks[..., 0, 0] * Ks[..., 1, 1] * math.pi)  # (B,V)
    out = torch.empty(V, S, 3 * bb + F + 4, dtype=dt)
    vox = torch.empty(S, feat_volume.shape[1], dtype=dt)
    start = 0
    for bi in range(B):
        n = int(samples.samples_per_batch[bi])
        sl = slice(start, start + n)
        uvd = samples.uvd[sl]
        vox[sl] = trilinear_border(feat_volume[bi], uvd[:, 0], uvd[:, 1], uvd[:, 2])
        pts = samples.rays_xyz[sl].permute(0, 2, 1)                         # (n,bb,3)
        centre_w = samples.rays_xyz[sl].mean(-1)                            # (n,3)
        t_hat = _unit(centre_w - tar_centre[bi])
        for v in range(V):
            E, K = src_exts[bi, v], src_ints[bi, v]
            cam = pts @ E[:3, :3].t() + E[:3, 3]                            # (n,bb,3)
            img = cam @ K.t()
            z = img[..., 2].clamp_min(1e-6)
            gx = 2.0 * (img[..., 0] / z) / W - 1.0
            gy = 2.0 * (img[..., 1] / z) / H - 1.0
            col = bilinear_2d(src_images[bi, v], gx, gy, "border")          # (3,n,bb)
            out[v, sl, : 3 * bb] = col.permute(1, 0, 2).reshape(n, 3 * bb)
            # sphere centre in camera space and its projected footprint -> level of detail
            c = cam.mean(-2)                                                # (n,3)
            dist = torch.linalg.vector_norm(c, dim=-1)
            sec_sq = (dist / c[:, 2]).square()
            foot = sec_sq / (
                torch.sqrt(((dist / samples.ball_radii[sl]).square() - 1.0).clamp_min(1e-12))
                + torch.sqrt((sec_sq - 1.0).clamp_min(1e-12))
            )
            lod = torch.log2(foot / src_radius[bi, v])
            cimg = c @ Ks[bi, v].t()
            cz = cimg[:, 2].clamp_min(1e-6)
            u01 = cimg[:, 0] / cz / Wb
            v01 = cimg[:, 1] / cz / Hb
            out[v, sl, 3 * bb: 3 * bb + F] = mip_fetch(tex_levels[bi], v, u01, v01, lod)
            s_hat = _unit(centre_w - src_centre[bi, v])
            out[v, sl, 3 * bb + F: 3 * bb + F + 3] = _unit(t_hat - s_hat)
            out[v, sl, 3 * bb + F + 3] = (t_hat * s_hat).sum(-1)
        start += n
    return out, vox


# --------------------------------------------------------------------------
# a10  aggregation / radiance MLP      nerf.py:58-115
# --------------------------------------------------------------------------
MLP_LAYERS = ("view_fc.0", "global_fc.0", "agg_w_fc.0", "fc.0", "lr0.0", "sigma.0", "weight.0", "weight.2", "feat_head.0")


def _lin(p: Mapping[str, Tensor], name: str, x: Tensor) -> Tensor:
    return x @ p[name + ".weight"].to(x.dtype).t() + p[name + ".bias"].to(x.dtype)


def radiance_mlp(p: Mapping[str, Tensor], vox: Tensor, rgbs_feat_dir: Tensor, feat_dim: int) -> Tuple[Tensor, Tensor]:
    """p: the ``nerf.*`` slice of the state dict (keys without the prefix).
    vox (S,8); rgbs_feat_dir (V,S,R+F+4) with F = feat_dim+3 ->
    sigma (S,), feat (S,R+F+8)."""
    V = rgbs_feat_dir.shape[0]
    F = feat_dim + 3
    fr_dir = rgbs_feat_dir[..., -(F + 4):]
    x = fr_dir[..., :F] + torch.relu(_lin(p, "view_fc.0", fr_dir[..., F:]))
    mean = x.mean(0, keepdim=True)
    var = ((x - mean) ** 2).sum(0, keepdim=True) / (V - 1)
    g = torch.relu(_lin(p, "global_fc.0", torch.cat((x, var.expand_as(x), mean.expand_as(x)), -1)))
    a = torch.softmax(torch.relu(_lin(p, "agg_w_fc.0", g)), 0)
    img = torch.relu(_lin(p, "fc.0", (g * a).sum(0)))
    vi = torch.cat((vox, img), -1)
    h = torch.relu(_lin(p, "lr0.0", vi))
    sigma = torch.nn.functional.softplus(_lin(p, "sigma.0", h)).squeeze(-1)
    shared = torch.cat((h, vi), -1).unsqueeze(0).expand(V, -1, -1)
    hid = torch.relu(_lin(p, "weight.0", torch.cat((shared, fr_dir), -1)))
    w = torch.softmax(torch.relu(_lin(p, "weight.2", hid)), 0)
    blended = (rgbs_feat_dir[..., :-4] * w).sum(0)
    feat = torch.cat((blended, torch.relu(_lin(p, "feat_head.0", h))), -1)
    return sigma, feat


# --------------------------------------------------------------------------
# a11/a12  compositing             utils.py:19-43,88-121; network.py:54-91
# --------------------------------------------------------------------------
def composite(sigma: Tensor, feat: Tensor, z_vals: Tensor, indices: Tensor, counts: Tensor, inv_depth: bool):
    """counts (NB,) integer samples per bundle (every bundle has >= 1).
    -> weights (S,), bundle feat (NB,C), depth (NB,), opacity (NB,)."""
    NB = counts.numel()
    S = sigma.numel()
    counts = counts.long()
    first = torch.cumsum(counts, 0) - counts
    rank = torch.arange(S) - first[indices]
    alpha = 1.0 - torch.exp(-sigma)
    trans = torch.ones_like(alpha)
    running = torch.ones(NB, dtype=sigma.dtype)
    for k in range(int(counts.max()) if S else 0):
        sel = rank == k
        which = indices[sel]
        trans[sel] = running[which]
        running[which] = running[which] * (1.0 - alpha[sel])
    w = alpha * trans
    total = torch.zeros(NB, dtype=sigma.dtype).index_add_(0, indices, w)
    w = w / total[indices].clamp_min(1e-6)
    zz = 1.0 / z_vals if inv_depth else z_vals
    values = torch.cat((feat, zz.unsqueeze(1), torch.ones_like(zz).unsqueeze(1)), 1)
    acc = torch.zeros(NB, values.shape[1], dtype=sigma.dtype).index_add_(0, indices, w.unsqueeze(1) * values)
    depth = acc[:, -2]
    if inv_depth:
        depth = 1.0 / depth
    return w, acc[:, :-2], depth, acc[:, -1]


# --------------------------------------------------------------------------
# the whole star-marked path after the CNNs   network.py:145-172
# --------------------------------------------------------------------------
def feature_texture(feat_level_maps: Tensor, src_images: Tensor, Hb: int, Wb: int, max_mip: int) -> List[List[Tensor]]:
    """feat_level_maps (B,V,Cf,Hb,Wb), src_images (B,V,3,H,W) -> per batch mip
    chain of channels-last (V,Hb,Wb,Cf+3) textures (network.py:159-164)."""
    B, V = src_images.shape[:2]
    if feat_level_maps.shape[-2:] != (Hb, Wb):
        feat_level_maps = upsample_bilinear(feat_level_maps.flatten(0, 1), Hb, Wb).unflatten(0, (B, V))
    lo = downsample_rgb(src_images.flatten(0, 1), Hb, Wb).unflatten(0, (B, V))
    tex = torch.cat((feat_level_maps, lo), 2).permute(0, 1, 3, 4, 2).contiguous()
    return [build_mips(tex[bi], max_mip) for bi in range(B)]


def render_bundles(
    mlp: Mapping[str, Tensor], feat_dim: int,
    src_images: Tensor, feat_level_maps: Tensor, feat_volume: Tensor, depth_range: Tensor, vol_range: Tensor,
    src_exts: Tensor, src_ints: Tensor, tar_exts: Tensor, tar_ints: Tensor, near_far: Tensor,
    b: int, max_num: int, global_num_depth: int, max_mip: int, inv_depth: bool, adaptive: bool,
) -> Dict[str, Tensor]:
    """Everything the fused CUDA kernel replaces, with all intermediates."""
    B, V, _, H, W = src_images.shape
    Hb, Wb = H // b, W // b
    rays = target_rays(tar_exts, tar_ints, H, W)
    bundles = assemble_bundles(rays, b)
    smp = sample_bundles(bundles, depth_range, vol_range, near_far[:, 0], near_far[:, 1],
                         global_num_depth, max_num, inv_depth, adaptive)
    tex = feature_texture(feat_level_maps, src_images, Hb, Wb, max_mip)
    rfd, vox = encode_samples(src_images, tex, feat_volume, smp, src_exts, src_ints, tar_exts, b)
    sigma, feat = radiance_mlp(mlp, vox, rfd, feat_dim)
    w, bfeat, bdepth, bopac = composite(sigma, feat, smp.z_vals, smp.indices, smp.samples_per_bundle, inv_depth)
    return {
        "indices": smp.indices, "samples_per_bundle": smp.samples_per_bundle, "samples_per_batch": smp.samples_per_batch,
        "z_vals": smp.z_vals, "uvd": smp.uvd, "ball_radii": smp.ball_radii, "rays_xyz": smp.rays_xyz,
        "rgbs_feat_dir": rfd, "vox_feat": vox, "sigma": sigma, "feat": feat, "weights": w,
        "bundle_feat": bfeat.view(B, Hb, Wb, -1).permute(0, 3, 1, 2).contiguous(),
        "bundle_depth": bdepth.view(B, Hb, Wb), "bundle_opacity": bopac.view(B, Hb, Wb),
    }


# --------------------------------------------------------------------------
# whole Network.forward on the CPU (network.py:93-189): the CPU baseline that
# bench.py times (`cpu_baseline`, `--impl reference`).  The convolutional
# networks are PyTorch modules in the reference as well; they are passed in.
# --------------------------------------------------------------------------
def network_forward(net, batch: Mapping, cfg) -> Tuple[Dict[str, Tensor], List[Tensor]]:
    """``net``: any module exposing feature_net / depth_net.cost_regs / nerf /
    upsampler with the reference's parameter names (the product's
    ``gdb_nerf_b200.network.Network`` on the CPU qualifies; only its CNN
    sub-modules and its parameter tensors are used here, never its forward)."""
    import torch.nn.functional as Fn

    src = batch["src_views"]
    images, src_exts, src_ints = src["rgb"], src["extrinsics"], src["intrinsics"]
    tar_exts, tar_ints = batch["tar_views"]["extrinsics"], batch["tar_views"]["intrinsics"]
    near_far = batch["near_far"]
    B, V, _, H, W = images.shape
    b = cfg.nerf.bundle_size
    feats = [f.unflatten(0, (B, V)) for f in net.feature_net(images.flatten(0, 1))]
    depth_range = near_far[..., None, None]
    mvs_depths = []
    n_stage = len(cfg.mvs.vol_levels)
    for s in range(n_stage):
        lvl = cfg.mvs.vol_levels[s]
        fs, vs = cfg.fpn.feat_scales[lvl], cfg.mvs.vol_scales[s]
        Ks = src_ints.clone(); Ks[..., :2, :] *= fs
        Kt = tar_ints.clone(); Kt[:, :2, :] *= vs
        Hi, Wi = int(H * vs), int(W * vs)
        dv = depth_hypotheses(depth_range, cfg.mvs.num_depth[s], cfg.mvs.inv_depth[s]).expand(B, -1, Hi, Wi)
        proj = homography_matrices(src_exts, Ks, tar_exts, Kt)
        variance = warp_variance(feats[lvl], proj, dv, cfg.mvs.inv_depth[s])
        volume, prob = net.depth_net.cost_regs[s](variance)
        depth, ci = depth_interval(dv, prob, cfg.mvs.ci_scales[s], cfg.mvs.inv_depth[s])
        mvs_depths.append(depth.squeeze(1))
        vol_range = dv[:, [0, -1]]
        depth_range = ci
        if s < n_stage - 1:
            up = cfg.mvs.vol_scales[s + 1] / cfg.mvs.vol_scales[s]
            depth_range = upsample_bilinear(ci, int(ci.shape[-2] * up), int(ci.shape[-1] * up))
    Hb, Wb = H // b, W // b
    mvs_depth = mvs_depths[-1]
    if ci.shape[-2:] != (Hb, Wb):
        ci = upsample_bilinear(ci, Hb, Wb)
        vol_range = upsample_bilinear(vol_range, Hb, Wb)
        mvs_depth = Fn.interpolate(mvs_depth.unsqueeze(1), size=(Hb, Wb), mode="nearest").squeeze(1)
    lvl = 0
    while cfg.fpn.feat_scales[lvl] < 1.0 / b:
        lvl += 1
    mlp = {k: v.detach() for k, v in net.nerf.state_dict().items()}
    out = render_bundles(mlp, cfg.fpn.feat_dims[lvl], images, feats[lvl], volume, ci, vol_range, src_exts, src_ints, tar_exts,
                         tar_ints, near_far, b, cfg.nerf.max_num_samples, cfg.nerf.global_num_depth, cfg.nerf.max_mipmap_level,
                         cfg.mvs.inv_depth[-1], cfg.nerf.is_adaptive)
    bf = out["bundle_feat"]
    R = 3 * b * b
    coarse = net.upsampler(bf[:, R:])
    fine = Fn.pixel_shuffle(bf[:, :R], b)
    rgb = coarse + fine
    if cfg.nerf.reweighting:
        rgb = 0.5 * (rgb + fine)
    ret = {
        "rgb": rgb,
        "nerf_depth": upsample_bilinear(out["bundle_depth"].unsqueeze(1), H, W).squeeze(1),
        "opacity": upsample_bilinear(out["bundle_opacity"].unsqueeze(1), H, W).squeeze(1),
        "mvs_depth": mvs_depth,
    }
    return ret, mvs_depths
