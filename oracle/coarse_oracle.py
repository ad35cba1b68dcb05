"""Training-only coarse render of the cascade's first stage (SURVEY.md section 8 row a14;
reference: networks/gdb_nerf/depth_net.py:49-116 `_render_rays`, :301-341 `build_rays`,
:344-396 `get_img_feat_vectorized`).

Status: TEST INFRASTRUCTURE (lives under oracle/; only tests import it).  The product path runs this row on hand-written kernels
(csrc/gdb_coarse.cu through autograd.CoarseRender / coarse_render_train); this PyTorch-operator
restatement is what tests/test_backward_gpu.py evaluates in float64 on the CPU to check the
kernels' outputs and gradients.  The arithmetic follows the reference step by step.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def coarse_render(nerf, feat_volume: Tensor, feats: Tensor, src_images: Tensor, feat_scale: float, src_exts: Tensor,
                  src_ints_stage: Tensor, tar_exts: Tensor, tar_ints_stage: Tensor, ray_range: Tensor, vol_range: Tensor,
                  num_samples: int, inv_depth: bool) -> Tensor:
    """-> (B,3,Hi,Wi) blended colour of one ray per cost-volume pixel.
    feat_volume (B,8,D,Hi,Wi); feats (B,V,C,Hs,Ws) FPN level of this stage; ray_range / vol_range (B,2,Hi,Wi)."""
    B, V = feats.shape[:2]
    Hi, Wi = ray_range.shape[-2:]
    dev, dt = feats.device, feats.dtype
    n = Hi * Wi
    # rays through the pixel centres (depth_net.py:318-331)
    xs = torch.arange(Wi, device=dev, dtype=dt) + 0.5
    ys = torch.arange(Hi, device=dev, dtype=dt) + 0.5
    gx, gy = torch.meshgrid(xs, ys, indexing="xy")
    gx, gy = gx.reshape(-1), gy.reshape(-1)
    pix = torch.stack((gx, gy, torch.ones_like(gx)), 1)                                  # (n,3)
    c2w = torch.inverse(tar_exts)
    origin = c2w[:, None, :3, 3]                                                         # (B,1,3)
    dirs = pix @ (c2w[:, :3, :3] @ torch.inverse(tar_ints_stage)).transpose(-2, -1)      # (B,n,3) un-normalised
    uv = torch.stack((2 * gx / Wi - 1, 2 * gy / Hi - 1), -1)                             # (n,2)
    rr = ray_range.permute(0, 2, 3, 1).reshape(B, n, 2)
    vr = vol_range.permute(0, 2, 3, 1).reshape(B, n, 2)
    r_near, r_far, v_near, v_far = rr[..., :1], rr[..., 1:], vr[..., :1], vr[..., 1:]
    if inv_depth:                                                                        # sample in disparity (:82-84)
        r_near, r_far = 1. / r_far, 1. / r_near
        v_near, v_far = 1. / v_far, 1. / v_near
    t = r_near + (r_far - r_near) * torch.linspace(0., 1., num_samples + 1, device=dev, dtype=dt)
    z = 0.5 * (t[..., :-1] + t[..., 1:])                                                 # (B,n,S)
    d = 2 * (z - v_near) / (v_far - v_near) - 1.
    uvd = torch.cat((uv[None, :, None, :].expand(B, -1, num_samples, -1), d[..., None]), -1)
    if inv_depth:
        z = 1. / z
    xyz = origin[:, :, None, :] + dirs[:, :, None, :] * z[..., None]                     # (B,n,S,3)

    vox = F.grid_sample(feat_volume, uvd.view(B, -1, 1, 1, 3), mode="bilinear", padding_mode="border", align_corners=False)
    vox = vox.flatten(2).permute(0, 2, 1)                                                # (B,n*S,8)
    # source features + low-resolution colours (:184-189)
    rgb_lo = F.interpolate(src_images.flatten(0, 1), scale_factor=feat_scale, mode="bilinear", align_corners=False)
    tex = torch.cat((feats, rgb_lo.unflatten(0, (B, V))), 2)                             # (B,V,C+3,Hs,Ws)
    Hs, Ws = tex.shape[-2:]
    pts = xyz.reshape(B, -1, 3)
    hom = torch.cat((pts, torch.ones_like(pts[..., :1])), -1)
    cam = torch.matmul(torch.matmul(hom[:, None], src_exts.transpose(-2, -1))[..., :3], src_ints_stage.transpose(-2, -1))  # (B,V,N,3)
    behind = cam[..., 2] < 1e-8
    grid = cam[..., :2] / cam[..., 2:3]
    grid = torch.stack((2 * grid[..., 0] / Ws - 1, 2 * grid[..., 1] / Hs - 1), -1)
    grid = torch.where(behind[..., None], torch.full_like(grid, -99.), grid)             # (:369-372)
    samp = F.grid_sample(tex.flatten(0, 1), grid.reshape(B * V, -1, 1, 2), mode="bilinear", padding_mode="border", align_corners=False)
    samp = samp.view(B, V, -1, pts.shape[1]).permute(0, 3, 1, 2)                         # (B,N,V,C+3)
    # direction features (:381-391)
    tar_c = c2w[:, None, :3, 3]
    src_c = torch.inverse(src_exts)[..., :3, 3]
    t_dir = F.normalize(pts - tar_c, dim=-1)
    s_dir = F.normalize(pts[:, :, None] - src_c[:, None], dim=-1)
    diff = F.normalize(t_dir[:, :, None] - s_dir, dim=-1)
    dot = (t_dir[:, :, None] * s_dir).sum(-1, keepdim=True)
    feat_rgb_dir = torch.cat((samp, diff, dot), -1)                                      # (B,N,V,C+3+4)

    sigma, rgb = nerf(vox, feat_rgb_dir)                                                 # (B,N), (B,N,3)
    sigma = sigma.view(B, n, num_samples)
    rgb = rgb.view(B, n, num_samples, 3)
    alpha = 1. - torch.exp(-sigma)
    T = torch.cumprod(1. - alpha + 1e-10, -1)[..., :-1]                                  # no renormalisation here (:109-114)
    T = torch.cat((torch.ones_like(alpha[..., :1]), T), -1)
    out = ((alpha * T)[..., None] * rgb).sum(-2)                                         # (B,n,3)
    return out.permute(0, 2, 1).reshape(B, 3, Hi, Wi)
