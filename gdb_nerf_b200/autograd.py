"""``torch.autograd.Function`` wrappers of the forward/backward kernel pairs used by the
training configuration (configs/dtu_pretrain.yaml; SURVEY.md section 8b "Autograd").

The reference has no hand-written backward: autograd differentiates its Python.  Here
each star-marked forward kernel is paired with the adjoint kernel of
csrc/gdb_costvolume_bwd.cu / csrc/gdb_render_bwd.cu, so gradients reach

  * the FPN feature maps      (warp+variance taps, mip-mapped texture taps),
  * the cost-regularisation probabilities and feature volume,
  * the depth interval / volume range of every stage (through sample positions, tap
    coordinates, mip level, direction features - network.py:149-168 keeps all of
    these in the graph),
  * the aggregation / radiance MLP parameters.

Source images and camera parameters are model inputs and get no gradient.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib
from .ops import CAM_HEAD, CAM_VIEW, Sources, _f32, _p, _stream, padded_feat  # noqa: F401
from . import ops

Tensor = torch.Tensor


class WarpVariance(torch.autograd.Function):
    """K1.  feat_cl (B,V,Hs,Ws,C) channels-last, depth_range (B,2,1|Ht,1|Wt) -> variance (B,C,D,Ht,Wt) planar."""

    @staticmethod
    def forward(ctx, feat_cl: Tensor, proj: Tensor, depth_range: Tensor, num_depth: int, Ht: int, Wt: int, inv_depth: bool) -> Tensor:
        feat_cl, proj, depth_range = _f32(feat_cl.detach()), _f32(proj.detach()), _f32(depth_range.detach())
        out = ops.warp_variance(feat_cl, proj, depth_range, num_depth, Ht, Wt, inv_depth, out_channels_last=False)
        ctx.save_for_backward(feat_cl, proj, depth_range)
        ctx.meta = (num_depth, Ht, Wt, bool(inv_depth))
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        feat_cl, proj, depth_range = ctx.saved_tensors
        D, Ht, Wt, inv = ctx.meta
        B, V, Hs, Ws, Cc = feat_cl.shape
        rh, rw = depth_range.shape[-2:]
        g = _f32(g)
        d_feat = torch.zeros_like(feat_cl)
        want_range = ctx.needs_input_grad[2] and rh != 1
        d_range = torch.zeros_like(depth_range) if want_range else None
        lib = _lib.load()
        _lib.check(lib.gdb_warp_variance_bwd(feat_cl.data_ptr(), proj.data_ptr(), depth_range.data_ptr(), rh, rw, B, V, Cc, Hs, Ws, D,
                                             Ht, Wt, int(inv), g.data_ptr(), d_feat.data_ptr(), _p(d_range), _stream()),
                   "gdb_warp_variance_bwd")
        return d_feat, None, d_range, None, None, None, None


class DepthRange(torch.autograd.Function):
    """K2.  -> depth (B,1,h,w), ci (B,2,h,w), vol_range (B,2,h,w)."""

    @staticmethod
    def forward(ctx, depth_range: Tensor, prob: Tensor, ci_scale: float, inv_depth: bool) -> Tuple[Tensor, Tensor, Tensor]:
        depth_range, prob = _f32(depth_range.detach()), _f32(prob.detach())
        depth, ci, vol = ops.depth_range_from_prob(depth_range, prob, ci_scale, inv_depth)
        ctx.save_for_backward(depth_range, prob)
        ctx.meta = (float(ci_scale), bool(inv_depth))
        return depth, ci, vol

    @staticmethod
    def backward(ctx, g_depth, g_ci, g_vol):
        depth_range, prob = ctx.saved_tensors
        ci_scale, inv = ctx.meta
        B, D, h, w = prob.shape
        rh, rw = depth_range.shape[-2:]
        d_prob = torch.empty_like(prob)
        want_range = ctx.needs_input_grad[0] and rh != 1
        d_range = torch.zeros_like(depth_range) if want_range else None
        gd = None if g_depth is None else _f32(g_depth)
        gc = None if g_ci is None else _f32(g_ci)
        gv = None if g_vol is None else _f32(g_vol)
        lib = _lib.load()
        _lib.check(lib.gdb_depth_range_bwd(depth_range.data_ptr(), rh, rw, prob.data_ptr(), B, D, h, w, ci_scale, int(inv), _p(gd), _p(gc),
                                           _p(gv), d_prob.data_ptr(), _p(d_range), _stream()), "gdb_depth_range_bwd")
        return d_range, d_prob, None, None


class PrepareSources(torch.autograd.Function):
    """Texture pyramid + RGBA images.  feat (B,V,Cf,Hb,Wb), images (B,V,3,H,W) -> tex (flat mip chain), rgba."""

    @staticmethod
    def forward(ctx, feat: Tensor, images: Tensor, bundle_size: int, max_mip: int) -> Tuple[Tensor, Tensor]:
        src = ops.prepare_sources(feat.detach(), images.detach(), bundle_size, max_mip)
        ctx.meta = (tuple(feat.shape), max_mip)
        ctx.mark_non_differentiable(src.rgba)
        return src.tex, src.rgba

    @staticmethod
    def backward(ctx, g_tex, _g_rgba):
        (B, V, Cf, Hb, Wb), max_mip = ctx.meta
        g = _f32(g_tex).clone()                       # the pull-down works in place
        d_feat = torch.empty((B, V, Cf, Hb, Wb), device=g.device, dtype=torch.float32)
        lib = _lib.load()
        _lib.check(lib.gdb_prepare_sources_bwd(g.data_ptr(), B * V, Cf, Hb, Wb, max_mip, d_feat.data_ptr(), _stream()),
                   "gdb_prepare_sources_bwd")
        return d_feat, None, None, None


class RenderFused(torch.autograd.Function):
    """K3 / K4.  -> feat (B,3b^2+F+8,Hb,Wb) planar, depth (B,Hb,Wb), opacity (B,Hb,Wb)."""

    @staticmethod
    def forward(ctx, tex: Tensor, rgba: Tensor, vol_cl: Tensor, depth_range: Tensor, vol_range: Tensor, cam: Tensor, mlp: Tensor,
                meta: tuple) -> Tuple[Tensor, Tensor, Tensor]:
        feat_dim, max_mip, B, V, H, W, b, max_samples, inv_depth, adaptive = meta
        tex, rgba, vol_cl = tex.detach(), rgba.detach(), _f32(vol_cl.detach())
        depth_range, vol_range, cam, mlp = _f32(depth_range.detach()), _f32(vol_range.detach()), cam.detach(), _f32(mlp.detach())
        out = ops.render_fused(Sources(tex, rgba, feat_dim, max_mip), vol_cl, depth_range, vol_range, cam, mlp, B, V, H, W, b,
                               max_samples, inv_depth, adaptive)
        ctx.save_for_backward(tex, rgba, vol_cl, depth_range, vol_range, cam, mlp)
        ctx.meta = meta
        return out["feat"], out["depth"], out["opacity"]

    @staticmethod
    def backward(ctx, g_feat, g_depth, g_opac):
        tex, rgba, vol_cl, depth_range, vol_range, cam, mlp = ctx.saved_tensors
        feat_dim, max_mip, B, V, H, W, b, max_samples, inv_depth, adaptive = ctx.meta
        D = vol_cl.shape[1]
        g_feat = _f32(g_feat) if g_feat is not None else torch.zeros((B, 3 * b * b + feat_dim + 3 + 8, H // b, W // b), device=tex.device)
        gd = None if g_depth is None else _f32(g_depth)
        go = None if g_opac is None else _f32(g_opac)
        d_mlp = torch.zeros_like(mlp)
        d_tex = torch.zeros_like(tex)
        d_vol = torch.zeros_like(vol_cl)
        d_dr = torch.empty_like(depth_range)
        d_vr = torch.empty_like(vol_range)
        lib = _lib.load()
        _lib.check(lib.gdb_render_fused_bwd(rgba.data_ptr(), tex.data_ptr(), vol_cl.data_ptr(), depth_range.data_ptr(), vol_range.data_ptr(),
                                            cam.data_ptr(), cam.shape[1], mlp.data_ptr(), B, V, H, W, b, feat_dim, D, 8, max_samples,
                                            max_mip, int(inv_depth), int(adaptive), g_feat.data_ptr(), _p(gd), _p(go), d_mlp.data_ptr(),
                                            d_tex.data_ptr(), d_vol.data_ptr(), d_dr.data_ptr(), d_vr.data_ptr(), _stream()),
                   "gdb_render_fused_bwd")
        return d_tex, None, d_vol, d_dr, d_vr, None, d_mlp, None


def render_fused_train(feat: Tensor, images: Tensor, feat_volume: Tensor, depth_range: Tensor, vol_range: Tensor, cam: Tensor,
                       mlp: Tensor, bundle_size: int, max_samples: int, max_mip: int, inv_depth: bool, adaptive: bool):
    """Differentiable fused render: feat (B,V,Cf,Hb,Wb) FPN level, images (B,V,3,H,W), feat_volume (B,8,D,Hb,Wb)."""
    B, V, Cf = feat.shape[:3]
    H, W = images.shape[-2:]
    tex, rgba = PrepareSources.apply(feat, images, bundle_size, max_mip)
    vol_cl = feat_volume.permute(0, 2, 3, 4, 1).contiguous()          # (B,D,Hb,Wb,8); autograd permutes the gradient back
    meta = (Cf, max_mip, B, V, H, W, bundle_size, max_samples, bool(inv_depth), bool(adaptive))
    return RenderFused.apply(tex, rgba, vol_cl, depth_range, vol_range, cam, mlp, meta)


class CoarseRender(torch.autograd.Function):
    """K5 (row a14).  tex: level-0 texture of the stage (flat), vol_cl (B,D,Hi,Wi,8), ray_range / vol_range (B,2,Hi,Wi)
    -> rgb (B,3,Hi,Wi)."""

    @staticmethod
    def forward(ctx, tex: Tensor, vol_cl: Tensor, ray_range: Tensor, vol_range: Tensor, cam: Tensor, mlp: Tensor, meta: tuple) -> Tensor:
        feat_dim, B, V, Hs, Ws, num_samples, inv_depth = meta
        tex, vol_cl = tex.detach(), _f32(vol_cl.detach())
        ray_range, vol_range, cam, mlp = _f32(ray_range.detach()), _f32(vol_range.detach()), cam.detach(), _f32(mlp.detach())
        D, Hi, Wi = vol_cl.shape[1:4]
        rgb = torch.empty((B, 3, Hi, Wi), device=tex.device, dtype=torch.float32)
        lib = _lib.load()
        _lib.check(lib.gdb_coarse_render_fwd(tex.data_ptr(), vol_cl.data_ptr(), ray_range.data_ptr(), vol_range.data_ptr(), cam.data_ptr(),
                                             cam.shape[1], mlp.data_ptr(), B, V, Hi, Wi, Hs, Ws, feat_dim, D, num_samples, int(inv_depth),
                                             rgb.data_ptr(), _stream()), "gdb_coarse_render_fwd")
        ctx.save_for_backward(tex, vol_cl, ray_range, vol_range, cam, mlp)
        ctx.meta = meta
        return rgb

    @staticmethod
    def backward(ctx, g_rgb):
        tex, vol_cl, ray_range, vol_range, cam, mlp = ctx.saved_tensors
        feat_dim, B, V, Hs, Ws, num_samples, inv_depth = ctx.meta
        D, Hi, Wi = vol_cl.shape[1:4]
        g_rgb = _f32(g_rgb)
        d_mlp, d_tex, d_vol = torch.zeros_like(mlp), torch.zeros_like(tex), torch.zeros_like(vol_cl)
        d_rr, d_vr = torch.empty_like(ray_range), torch.empty_like(vol_range)
        lib = _lib.load()
        _lib.check(lib.gdb_coarse_render_bwd(tex.data_ptr(), vol_cl.data_ptr(), ray_range.data_ptr(), vol_range.data_ptr(), cam.data_ptr(),
                                             cam.shape[1], mlp.data_ptr(), B, V, Hi, Wi, Hs, Ws, feat_dim, D, num_samples, int(inv_depth),
                                             g_rgb.data_ptr(), d_mlp.data_ptr(), d_tex.data_ptr(), d_vol.data_ptr(), d_rr.data_ptr(),
                                             d_vr.data_ptr(), _stream()), "gdb_coarse_render_bwd")
        return d_tex, d_vol, d_rr, d_vr, None, d_mlp, None


def coarse_render_train(nerf, feat_volume: Tensor, feats: Tensor, src_images: Tensor, src_exts: Tensor, src_ints_stage: Tensor,
                        tar_exts: Tensor, tar_ints_stage: Tensor, near_far: Tensor, ray_range: Tensor, vol_range: Tensor,
                        num_samples: int, inv_depth: bool) -> Tensor:
    """Differentiable coarse render of one cascade stage (depth_net.py:181-191): same arguments as ``coarse.coarse_render``
    (the PyTorch restatement kept as the test reference), arithmetic in gdb_coarse.cu."""
    from .mlp_pack import pack_mlp
    B, V, Cf, Hs, Ws = feats.shape
    H = src_images.shape[-2]
    if H % Hs:
        raise ValueError("image height must be a multiple of the feature-map height")
    tex, _rgba = PrepareSources.apply(feats, src_images, H // Hs, 0)
    vol_cl = feat_volume.permute(0, 2, 3, 4, 1).contiguous()
    cam = ops.camera_block(tar_exts, tar_ints_stage, src_exts, src_ints_stage, near_far, 1, 1, inv_depth)
    named = dict(nerf.named_parameters())
    zeros_w = torch.zeros(8, 64, device=feats.device)
    params = {k: v for k, v in named.items() if not k.startswith("color.")}
    params.update({"weight.0.weight": named["color.0.weight"], "weight.0.bias": named["color.0.bias"],
                   "weight.2.weight": named["color.2.weight"], "weight.2.bias": named["color.2.bias"],
                   "feat_head.0.weight": zeros_w, "feat_head.0.bias": zeros_w[:, 0]})
    mlp = pack_mlp(params, Cf, detach=False)
    meta = (Cf, B, V, Hs, Ws, num_samples, bool(inv_depth))
    return CoarseRender.apply(tex, vol_cl, ray_range, vol_range, cam, mlp, meta)
