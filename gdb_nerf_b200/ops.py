"""Host-side operators: thin wrappers that hand PyTorch-owned device memory and
the current CUDA stream to the C-ABI kernels.  PyTorch is plumbing here
(allocator + stream); all arithmetic of the path happens in the kernels.

Every operator refuses CPU tensors: there is no fallback implementation.
"""
from __future__ import annotations

import os

import ctypes as C
from typing import Dict, NamedTuple, Optional, Tuple

import torch

from . import _lib
from .mlp_pack import pack_mlp  # noqa: F401  (re-export)

Tensor = torch.Tensor
CAM_HEAD = 32
CAM_VIEW = 32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev(*tensors: Tensor) -> None:
    """Every tensor handed to the library must live on the CURRENT CUDA device: the kernels launch on that device's current
    stream (one process per GPU is the deployment; a caller driving several GPUs wraps calls in ``torch.cuda.device``)."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.GdbError("gdb_nerf_b200 operators need CUDA tensors (no CPU fallback exists)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise _lib.GdbError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: "
                                "wrap the call in `with torch.cuda.device(tensor.device):`")


def _f32(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _p(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def padded_feat(feat_dim: int) -> int:
    return (feat_dim + 3 + 3) & ~3


# ------------------------------------------------------------------ layout --
def to_channels_last(x: Tensor, c_pad: Optional[int] = None) -> Tensor:
    """(N, C, *spatial) planar -> (N, *spatial, Cpad) channels-last."""
    _dev(x)
    N, Cc = x.shape[:2]
    spatial = x.shape[2:]
    if x.dtype == torch.float32 and (c_pad is None or c_pad == Cc) and Cc % 4 == 0:
        perm = x.permute(0, *range(2, x.dim()), 1)
        if perm.is_contiguous():          # already channels-last in memory (cuDNN NHWC / NDHWC output): zero-copy
            return perm
    x = _f32(x)
    S = 1
    for s in spatial:
        S *= s
    c_pad = c_pad or ((Cc + 3) & ~3)
    out = torch.empty((N, *spatial, c_pad), device=x.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.gdb_planar_to_channels_last(x.data_ptr(), out.data_ptr(), N, Cc, S, c_pad, _stream()), "gdb_planar_to_channels_last")
    return out


def u8_to_unit(x: Tensor) -> Tensor:
    """uint8 image samples -> float32 in [0, 1] (x / 255 in IEEE float32, as the reference's loaders compute on the host)."""
    _dev(x)
    if x.dtype != torch.uint8:
        raise _lib.GdbError(f"u8_to_unit needs a uint8 tensor, got {x.dtype}")
    x = x.contiguous()
    out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    if x.numel():
        lib = _lib.load()
        _lib.check(lib.gdb_u8_to_unit_f32(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "gdb_u8_to_unit_f32")
    return out


# ------------------------------------------------------------- cost volume --
def homography_mats(src_exts: Tensor, src_ints: Tensor, tar_exts: Tensor, tar_ints: Tensor, src_scale: float, tar_scale: float) -> Tensor:
    _dev(src_exts, src_ints, tar_exts, tar_ints)
    B, V = src_exts.shape[:2]
    proj = torch.empty((B, V, 3, 4), device=src_exts.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.gdb_homography_mats(_f32(src_exts).data_ptr(), _f32(src_ints).data_ptr(), _f32(tar_exts).data_ptr(),
                                       _f32(tar_ints).data_ptr(), float(src_scale), float(tar_scale), B, V, proj.data_ptr(), _stream()),
               "gdb_homography_mats")
    return proj


def depth_values(depth_range: Tensor, num_depth: int, Ht: int, Wt: int, inv_depth: bool) -> Tensor:
    _dev(depth_range)
    depth_range = _f32(depth_range)
    B, _, rh, rw = depth_range.shape
    out = torch.empty((B, num_depth, Ht, Wt), device=depth_range.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.gdb_depth_values(depth_range.data_ptr(), rh, rw, B, num_depth, Ht, Wt, int(inv_depth), out.data_ptr(), _stream()),
               "gdb_depth_values")
    return out


def warp_variance(feat_cl: Tensor, proj: Tensor, depth_range: Tensor, num_depth: int, Ht: int, Wt: int, inv_depth: bool,
                  out_channels_last: bool = False, depth_folded: bool = False) -> Tensor:
    """feat_cl (B,V,Hs,Ws,C) channels-last -> variance volume of SHAPE (B,C,D,Ht,Wt); with ``out_channels_last``
    the memory behind it is (B,D,Ht,Wt,C), i.e. torch.channels_last_3d strides; with ``depth_folded`` the result is
    the (B,D*C,Ht,Wt)-shaped channels-last 2-D map over (B,Ht,Wt,D,C) memory (channel = d*C + c)."""
    _dev(feat_cl, proj, depth_range)
    feat_cl, proj, depth_range = _f32(feat_cl), _f32(proj), _f32(depth_range)
    B, V, Hs, Ws, Cc = feat_cl.shape
    _, _, rh, rw = depth_range.shape
    if depth_folded:
        out = torch.empty((B, Ht, Wt, num_depth * Cc), device=feat_cl.device, dtype=torch.float32)
    elif out_channels_last:
        out = torch.empty((B, num_depth, Ht, Wt, Cc), device=feat_cl.device, dtype=torch.float32)
    else:
        out = torch.empty((B, Cc, num_depth, Ht, Wt), device=feat_cl.device, dtype=torch.float32)
    lib = _lib.load()
    layout = 2 if depth_folded else int(out_channels_last)
    _lib.check(lib.gdb_warp_variance_fwd(feat_cl.data_ptr(), proj.data_ptr(), depth_range.data_ptr(), rh, rw, B, V, Cc, Hs, Ws,
                                         num_depth, Ht, Wt, int(inv_depth), layout, out.data_ptr(), _stream()),
               "gdb_warp_variance_fwd")
    if depth_folded:
        return out.permute(0, 3, 1, 2)
    return out.permute(0, 4, 1, 2, 3) if out_channels_last else out


def depth_range_from_prob(depth_range: Tensor, prob: Tensor, ci_scale: float, inv_depth: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """-> depth (B,1,h,w), ci (B,2,h,w), vol_range (B,2,h,w)."""
    _dev(depth_range, prob)
    depth_range, prob = _f32(depth_range), _f32(prob)
    B, D, h, w = prob.shape
    _, _, rh, rw = depth_range.shape
    depth = torch.empty((B, 1, h, w), device=prob.device, dtype=torch.float32)
    ci = torch.empty((B, 2, h, w), device=prob.device, dtype=torch.float32)
    vol = torch.empty((B, 2, h, w), device=prob.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.gdb_depth_range_fwd(depth_range.data_ptr(), rh, rw, prob.data_ptr(), B, D, h, w, float(ci_scale), int(inv_depth),
                                       depth.data_ptr(), ci.data_ptr(), vol.data_ptr(), _stream()), "gdb_depth_range_fwd")
    return depth, ci, vol


def depth_range_from_logits(depth_range: Tensor, logits: Tensor, ci_scale: float, inv_depth: bool,
                            want_prob: bool = False) -> Tuple[Tensor, Tensor, Tensor, Optional[Tensor]]:
    """K2 fused with the probability head's soft-max.  ``logits`` is a (B,D,h,w) VIEW with any batch/depth strides whose
    pixels are equally spaced (e.g. one channel of a channels-last convolution output): no copy is made.
    -> depth (B,1,h,w), ci (B,2,h,w), vol_range (B,2,h,w), prob (B,D,h,w) or None."""
    _dev(depth_range, logits)
    depth_range = _f32(depth_range)
    if logits.dtype != torch.float32:
        logits = logits.float()
    B, D, h, w = logits.shape
    sb, sd, sy, sx = logits.stride()
    if sy != sx * w:
        logits = logits.contiguous()
        sb, sd, sy, sx = logits.stride()
    _, _, rh, rw = depth_range.shape
    dev = logits.device
    depth = torch.empty((B, 1, h, w), device=dev, dtype=torch.float32)
    ci = torch.empty((B, 2, h, w), device=dev, dtype=torch.float32)
    vol = torch.empty((B, 2, h, w), device=dev, dtype=torch.float32)
    prob = torch.empty((B, D, h, w), device=dev, dtype=torch.float32) if want_prob else None
    lib = _lib.load()
    _lib.check(lib.gdb_depth_range_from_logits_fwd(depth_range.data_ptr(), rh, rw, logits.data_ptr(), sb, sd, sx, B, D, h, w,
                                                   float(ci_scale), int(inv_depth), depth.data_ptr(), ci.data_ptr(), vol.data_ptr(),
                                                   _p(prob), _stream()), "gdb_depth_range_from_logits_fwd")
    return depth, ci, vol, prob


_PH_WEIGHTS: Dict = {}


def _prob_head_chunks(D: int) -> int:
    """Depth chunks of the split probability-head kernel: ~16 planes each, none empty."""
    n = max(1, D // 16)
    while n > 1 and (n - 1) * ((D + n - 1) // n) >= D:
        n -= 1
    return n


def prob_head_depth_range(y: Tensor, weight: Tensor, depth_range: Tensor, ci_scale: float, inv_depth: bool,
                          want_prob: bool = False, split: bool = True, tma: Optional[bool] = None) -> Tuple[Tensor, Tensor, Tensor, Optional[Tensor]]:
    """K2 fused with the probability head: ``y`` is the U-Net's last feature volume, (B,8,D,h,w)-shaped over channels-last
    (B,D,h,w,8) memory, ``weight`` the head's (1,8,3,3,3) Conv3d weight (padding 1, no bias).
    -> depth (B,1,h,w), ci (B,2,h,w), vol_range (B,2,h,w), prob (B,D,h,w) or None.
    ``tma``: the TMA-fed second-generation kernel (default; ``GDB_PH_TMA=0`` or ``tma=False`` selects the hand-staged split kernel)."""
    _dev(y, weight, depth_range)
    if not (_is_cl(y) and y.shape[1] == 8 and tuple(weight.shape) == (1, 8, 3, 3, 3)):
        raise _lib.GdbError("prob_head_depth_range needs a dense channels-last (B,8,D,h,w) volume and a (1,8,3,3,3) weight")
    depth_range = _f32(depth_range)
    B, Cc, D, h, w = y.shape
    _, _, rh, rw = depth_range.shape
    wkey = (weight.data_ptr(), weight._version, str(weight.device))
    hit = _PH_WEIGHTS.get(wkey)
    if hit is None:                                                   # [kd][ky][kx][c], re-laid out once per weight version
        if len(_PH_WEIGHTS) >= 8:
            _PH_WEIGHTS.clear()
        # the entry keeps `weight` itself alive: its address cannot be handed to another tensor while the key exists
        hit = _PH_WEIGHTS[wkey] = (weight, _f32(weight.detach())[0].permute(1, 2, 3, 0).contiguous())
    wk = hit[1]
    dev = y.device
    depth = torch.empty((B, 1, h, w), device=dev, dtype=torch.float32)
    ci = torch.empty((B, 2, h, w), device=dev, dtype=torch.float32)
    vol = torch.empty((B, 2, h, w), device=dev, dtype=torch.float32)
    prob = torch.empty((B, D, h, w), device=dev, dtype=torch.float32) if want_prob else None
    lib = _lib.load()
    nch = _prob_head_chunks(D) if (split and not want_prob) else 1
    if tma is None:
        tma = os.environ.get("GDB_PH_TMA", "1") != "0"
    if tma and not want_prob and B * nch <= 65535:
        # haloed planes by TMA (out-of-bounds zero fill = the padding), each plane read once, weights as constant operands
        scratch = torch.empty(int(lib.gdb_prob_head_tma_scratch_floats(B, h, w, nch)), device=dev, dtype=torch.float32)
        counters = torch.empty(int(lib.gdb_prob_head_tma_counters(B, h, w)), device=dev, dtype=torch.int32)
        _lib.check(lib.gdb_prob_head_depth_range_tma_fwd(y.data_ptr(), wk.data_ptr(), depth_range.data_ptr(), rh, rw, B, Cc, D, h, w,
                                                         nch, float(ci_scale), int(inv_depth), scratch.data_ptr(), counters.data_ptr(),
                                                         depth.data_ptr(), ci.data_ptr(), vol.data_ptr(), _stream()),
                   "gdb_prob_head_depth_range_tma_fwd")
        return depth, ci, vol, None
    if nch > 1 and B * nch <= 65535:
        # depth axis split over nch CTAs per tile (on-line soft-max partials merged by the tile's last CTA)
        scratch = torch.empty(int(lib.gdb_prob_head_split_scratch_floats(B, h, w, nch)), device=dev, dtype=torch.float32)
        # arrival counters: owned by this call, zeroed by the library (cudaMemsetAsync) in front of the kernel
        counters = torch.empty(int(lib.gdb_prob_head_split_counters(B, h, w)), device=dev, dtype=torch.int32)
        _lib.check(lib.gdb_prob_head_depth_range_split_fwd(y.data_ptr(), wk.data_ptr(), depth_range.data_ptr(), rh, rw, B, Cc, D, h, w,
                                                           nch, float(ci_scale), int(inv_depth), scratch.data_ptr(), counters.data_ptr(),
                                                           depth.data_ptr(), ci.data_ptr(), vol.data_ptr(), _stream()),
                   "gdb_prob_head_depth_range_split_fwd")
        return depth, ci, vol, None
    _lib.check(lib.gdb_prob_head_depth_range_fwd(y.data_ptr(), wk.data_ptr(), depth_range.data_ptr(), rh, rw, B, Cc, D, h, w,
                                                 float(ci_scale), int(inv_depth), depth.data_ptr(), ci.data_ptr(), vol.data_ptr(),
                                                 _p(prob), _stream()), "gdb_prob_head_depth_range_fwd")
    return depth, ci, vol, prob


# -------------------------------------------------------------------- glue --
def _is_cl(x: Tensor) -> bool:
    """Shape (N, C, *spatial) over channels-last memory (N, *spatial, C), dense."""
    return x.dtype == torch.float32 and x.is_cuda and x.permute(0, *range(2, x.dim()), 1).is_contiguous()


def bias_act_add(x: Tensor, bias: Optional[Tensor], skip: Optional[Tensor], relu: bool, skip_up2: bool = False) -> Tensor:
    """out = skip + act(x + bias) on channels-last tensors of shape (N, C, *spatial); with ``skip_up2`` skip is the
    half-resolution (N, C, H/2, W/2) map, nearest-neighbour up-sampled on the fly."""
    _dev(x, bias, skip)
    Cc = x.shape[1]
    if not (_is_cl(x) and (skip is None or _is_cl(skip)) and Cc % 4 == 0):
        raise _lib.GdbError("bias_act_add needs dense channels-last fp32 tensors with C % 4 == 0")
    N = x.shape[0]
    S = x.numel() // (N * Cc)
    Hs = Ws = 0
    if skip_up2:
        Hs, Ws = skip.shape[-2:]
        if x.dim() != 4 or tuple(x.shape[-2:]) != (2 * Hs, 2 * Ws):
            raise ValueError("skip_up2: x must be (N,C,2Hs,2Ws)")
    elif skip is not None and skip.shape != x.shape:
        raise ValueError("skip shape mismatch")
    out = torch.empty_like(x)             # preserves the channels-last strides
    lib = _lib.load()
    _lib.check(lib.gdb_bias_act_add(x.data_ptr(), _p(None if bias is None else _f32(bias)), _p(skip), N, S, Cc, int(relu),
                                    int(skip_up2), Hs, Ws, out.data_ptr(), _stream()), "gdb_bias_act_add")
    return out


def gate_add(x: Tensor, y: Tensor, gate: Tensor, extra: Optional[Tensor] = None) -> Tensor:
    """out = x + y * gate[n, c] (+ extra) (channels-last (N,C,H,W)-shaped x, y, extra; gate (N,C))."""
    _dev(x, y, gate, extra)
    if extra is not None and not (_is_cl(extra) and extra.shape == x.shape):
        raise _lib.GdbError("gate_add: extra must be a dense channels-last fp32 tensor of x's shape")
    N, Cc = x.shape[:2]
    if not (_is_cl(x) and _is_cl(y) and Cc % 4 == 0 and x.shape == y.shape):
        raise _lib.GdbError("gate_add needs dense channels-last fp32 tensors with C % 4 == 0")
    S = x.numel() // (N * Cc)
    out = torch.empty_like(x)
    lib = _lib.load()
    _lib.check(lib.gdb_gate_add(x.data_ptr(), y.data_ptr(), _f32(gate).data_ptr(), _p(extra), N, S, Cc, out.data_ptr(), _stream()), "gdb_gate_add")
    return out


def se_gate_add(x: Tensor, y: Tensor, w1: Tensor, w2: Tensor, extra: Optional[Tensor] = None, chunks: int = 64,
                cat_out: Optional[Tensor] = None) -> Tensor:
    """out = x + y * sigmoid(w2 @ relu(w1 @ mean_hw(y))) (+ extra): the squeeze-excite gate of a dense block and its residual
    (decoder_rdn.py:31-41) as two launches - channel sums of y, then one kernel that finishes the gate in its prologue."""
    _dev(x, y, w1, w2, extra)
    N, Cc = x.shape[:2]
    R = w1.shape[0]
    if not (_is_cl(x) and _is_cl(y) and Cc % 4 == 0 and x.shape == y.shape and (extra is None or (_is_cl(extra) and extra.shape == x.shape))):
        raise _lib.GdbError("se_gate_add needs dense channels-last fp32 tensors of one shape with C % 4 == 0")
    if tuple(w1.shape) != (R, Cc) or tuple(w2.shape) != (Cc, R):
        raise _lib.GdbError("se_gate_add: w1 must be (R, C) and w2 (C, R)")
    S = x.numel() // (N * Cc)
    partial = torch.empty((N, chunks, Cc), device=x.device, dtype=torch.float32)
    out = torch.empty_like(x)
    lib = _lib.load()
    cat_c = 0
    if cat_out is not None:      # (N, Ccat, *spatial) channels-last: the result is also written into its first C channels
        _dev(cat_out)
        if not (_is_cl(cat_out) and cat_out.shape[0] == N and cat_out.shape[2:] == x.shape[2:] and cat_out.shape[1] >= Cc and cat_out.shape[1] % 4 == 0):
            raise _lib.GdbError("se_gate_add: cat_out must be a dense channels-last fp32 buffer of x's batch / spatial size with >= C channels")
        cat_c = cat_out.shape[1]
    _lib.check(lib.gdb_se_gate_add_cat(x.data_ptr(), y.data_ptr(), _f32(w1).data_ptr(), _f32(w2).data_ptr(), R, _p(extra), N, S, Cc, chunks,
                                       partial.data_ptr(), out.data_ptr(), _p(cat_out), cat_c, _stream()), "gdb_se_gate_add")
    return out


def concat_into(buf: Tensor, offset: int, a: Tensor, b: Tensor) -> Tensor:
    """buf[:, offset : offset + Ca + Cb] = cat((a, b), 1) for dense channels-last fp32 maps (buf's leading channels were written
    by the producer of that tensor, see se_gate_add(cat_out=...)).  Returns buf."""
    _dev(buf, a, b)
    if not all(_is_cl(t) and t.shape[1] % 4 == 0 and t.shape[0] == buf.shape[0] and t.shape[2:] == buf.shape[2:] for t in (buf, a, b)) \
            or offset % 4 or offset + a.shape[1] + b.shape[1] > buf.shape[1]:
        raise _lib.GdbError("concat_into needs dense channels-last fp32 maps of one size with C % 4 == 0 that fit the buffer")
    npix = buf.numel() // buf.shape[1]
    _lib.check(_lib.load().gdb_concat2_into(a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1], npix, buf.data_ptr(), buf.shape[1], offset,
                                            _stream()), "gdb_concat2_into")
    return buf


def concat_channels(a: Tensor, b: Tensor, c: Optional[Tensor] = None) -> Tensor:
    """torch.cat((a, b[, c]), 1) for dense channels-last (N,C,H,W)-shaped fp32 maps, one streaming pass."""
    _dev(a, b, c)
    ts = [t for t in (a, b, c) if t is not None]
    if not all(_is_cl(t) and t.shape[1] % 4 == 0 and t.shape[0] == a.shape[0] and t.shape[2:] == a.shape[2:] for t in ts):
        raise _lib.GdbError("concat_channels needs dense channels-last fp32 maps of one size with C % 4 == 0")
    N, _, H, W = a.shape
    Ct = sum(t.shape[1] for t in ts)
    out = torch.empty((N, H, W, Ct), device=a.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.gdb_concat3(a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1], _p(c), 0 if c is None else c.shape[1], N * H * W,
                               out.data_ptr(), _stream()), "gdb_concat3")
    return out.permute(0, 3, 1, 2)


def pixel_shuffle2_bias(x: Tensor, bias: Optional[Tensor]) -> Tensor:
    """F.pixel_shuffle(x + bias, 2) for a dense channels-last (N,4C,H,W)-shaped fp32 map -> (N,C,2H,2W) channels-last."""
    _dev(x, bias)
    N, C4, H, W = x.shape
    if not (_is_cl(x) and C4 % 16 == 0):
        raise _lib.GdbError("pixel_shuffle2_bias needs a dense channels-last fp32 map with C % 16 == 0")
    Cc = C4 // 4
    out = torch.empty((N, 2 * H, 2 * W, Cc), device=x.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.gdb_pixel_shuffle2(x.data_ptr(), _p(None if bias is None else _f32(bias)), N, H, W, Cc, out.data_ptr(), _stream()),
               "gdb_pixel_shuffle2")
    return out.permute(0, 3, 1, 2)


def channel_mean(x: Tensor, chunks: int = 64) -> Tensor:
    """x.mean((2, 3)) for a dense channels-last (N,C,H,W)-shaped fp32 map -> (N, C); fixed summation order."""
    _dev(x)
    if not (_is_cl(x) and x.shape[1] % 4 == 0):
        raise _lib.GdbError("channel_mean needs a dense channels-last fp32 map with C % 4 == 0")
    N, Cc, H, W = x.shape
    partial = torch.empty((N, chunks, Cc), device=x.device, dtype=torch.float32)
    out = torch.empty((N, Cc), device=x.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.gdb_channel_mean(x.data_ptr(), N, H * W, Cc, chunks, partial.data_ptr(), out.data_ptr(), _stream()), "gdb_channel_mean")
    return out


# ---------------------------------------------------------------- sampling --
def camera_block(tar_exts: Tensor, tar_ints: Tensor, src_exts: Tensor, src_ints: Tensor, near_far: Tensor, bundle_size: int,
                 global_num_depth: int, inv_depth: bool) -> Tensor:
    _dev(tar_exts, tar_ints, src_exts, src_ints, near_far)
    B, V = src_exts.shape[:2]
    cam = torch.empty((B, CAM_HEAD + CAM_VIEW * V), device=tar_exts.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.gdb_camera_block(_f32(tar_exts).data_ptr(), _f32(tar_ints).data_ptr(), _f32(src_exts).data_ptr(),
                                    _f32(src_ints).data_ptr(), _f32(near_far).data_ptr(), B, V, bundle_size, global_num_depth,
                                    int(inv_depth), cam.data_ptr(), _stream()), "gdb_camera_block")
    return cam


class SampleList(NamedTuple):
    counts: Tensor       # (NB,) int32
    offsets: Tensor      # (NB+1,) int32, offsets[NB] = S
    total: int           # S (host value: one read-back)
    indices: Optional[Tensor]
    z_vals: Optional[Tensor]
    uvd: Optional[Tensor]
    ball_radii: Optional[Tensor]
    rays_xyz: Optional[Tensor]


def bundle_counts(depth_range: Tensor, cam: Tensor, max_samples: int, inv_depth: bool, adaptive: bool) -> Tuple[Tensor, Tensor]:
    """-> counts (NB,) int32 and offsets (NB+1,) int32 (exclusive scan), no host sync."""
    _dev(depth_range, cam)
    depth_range = _f32(depth_range)
    B, _, Hb, Wb = depth_range.shape
    NB = B * Hb * Wb
    counts = torch.empty(NB, device=cam.device, dtype=torch.int32)
    offsets = torch.empty(NB + 1, device=cam.device, dtype=torch.int32)
    block_sums = torch.empty((NB + 4095) // 4096, device=cam.device, dtype=torch.int32)
    lib = _lib.load()
    _lib.check(lib.gdb_bundle_count(depth_range.data_ptr(), cam.data_ptr(), cam.shape[1], B, Hb, Wb, max_samples, int(inv_depth),
                                    int(adaptive), counts.data_ptr(), block_sums.data_ptr(), _stream()), "gdb_bundle_count")
    _lib.check(lib.gdb_bundle_scan(counts.data_ptr(), block_sums.data_ptr(), NB, offsets.data_ptr(), _stream()), "gdb_bundle_scan")
    return counts, offsets


def sample_bundles(depth_range: Tensor, vol_range: Tensor, cam: Tensor, bundle_size: int, max_samples: int, inv_depth: bool,
                   adaptive: bool, want_rays: bool = True) -> SampleList:
    """The reference's packed sample list (bundle_sampler.py:193-265).  Reads S
    back to the host once to size the outputs (the fused render never needs it)."""
    _dev(depth_range, vol_range, cam)
    depth_range, vol_range = _f32(depth_range), _f32(vol_range)
    B, _, Hb, Wb = depth_range.shape
    counts, offsets = bundle_counts(depth_range, cam, max_samples, inv_depth, adaptive)
    S = int(offsets[-1].item())
    dev = cam.device
    bb = bundle_size * bundle_size
    indices = torch.empty(S, device=dev, dtype=torch.int64)
    z_vals = torch.empty(S, device=dev, dtype=torch.float32)
    uvd = torch.empty((S, 3), device=dev, dtype=torch.float32)
    ball = torch.empty(S, device=dev, dtype=torch.float32)
    rays = torch.empty((S, 3, bb), device=dev, dtype=torch.float32) if want_rays else None
    lib = _lib.load()
    _lib.check(lib.gdb_bundle_emit(depth_range.data_ptr(), vol_range.data_ptr(), cam.data_ptr(), cam.shape[1], counts.data_ptr(),
                                   offsets.data_ptr(), B, Hb, Wb, bundle_size, int(inv_depth), indices.data_ptr(), z_vals.data_ptr(),
                                   uvd.data_ptr(), ball.data_ptr(), _p(rays), _stream()), "gdb_bundle_emit")
    return SampleList(counts, offsets, S, indices, z_vals, uvd, ball, rays)


# ----------------------------------------------------------------- sources --
class Sources(NamedTuple):
    tex: Tensor      # flat mip chain (floats)
    rgba: Tensor     # (B*V, H, W, 4)
    feat_dim: int
    max_mip: int


def prepare_sources(feat: Tensor, images: Tensor, bundle_size: int, max_mip: int) -> Sources:
    """feat (B,V,Cf,Hb,Wb) FPN level (planar or channels-last memory), images (B,V,3,H,W)."""
    _dev(feat, images)
    B, V, Cf, Hb, Wb = feat.shape
    feat_cl = feat.dtype == torch.float32 and Cf % 4 == 0 and feat.permute(0, 1, 3, 4, 2).is_contiguous()
    if not feat_cl:
        feat = _f32(feat)
    images = _f32(images)
    H, W = images.shape[-2:]
    if (H, W) != (Hb * bundle_size, Wb * bundle_size):
        raise ValueError(f"feature map {Hb}x{Wb} x bundle {bundle_size} != image {H}x{W}")
    lib = _lib.load()
    n = lib.gdb_texture_floats(B * V, Hb, Wb, Cf, max_mip)
    tex = torch.empty(n, device=feat.device, dtype=torch.float32)
    rgba = torch.empty((B * V, H, W, 4), device=feat.device, dtype=torch.float32)
    _lib.check(lib.gdb_prepare_sources(feat.data_ptr(), int(feat_cl), images.data_ptr(), B * V, Cf, Hb, Wb, bundle_size, max_mip,
                                       tex.data_ptr(), rgba.data_ptr(), _stream()), "gdb_prepare_sources")
    return Sources(tex, rgba, Cf, max_mip)


def texture_level(src: Sources, BV: int, Hb: int, Wb: int, level: int) -> Tensor:
    """View of one mip level as (BV, h, w, FP)."""
    FP = padded_feat(src.feat_dim)
    off = 0
    for k in range(level):
        off += BV * (Hb >> k) * (Wb >> k) * FP
    h, w = Hb >> level, Wb >> level
    return src.tex[off: off + BV * h * w * FP].view(BV, h, w, FP)


# ------------------------------------------------------------------ render --
def render_fused(src: Sources, vol_cl: Tensor, depth_range: Tensor, vol_range: Tensor, cam: Tensor, mlp: Tensor,
                 B: int, V: int, H: int, W: int, bundle_size: int, max_samples: int, inv_depth: bool, adaptive: bool,
                 taps: Optional[SampleList] = None, precision: int = 0, out_channels_last: bool = False,
                 pad_dec: bool = False, dec_one: bool = False, rows: Optional[Tuple[int, int]] = None) -> Dict[str, Tensor]:
    """-> {'feat' (B,CT,Hb,Wb), 'depth' (B,Hb,Wb), 'opacity' (B,Hb,Wb)} and, when
    ``taps`` (a SampleList) is given, the reference's packed intermediates.
    With ``out_channels_last``: {'fine' (B,Hb,Wb,3b^2), 'dec_in' (B,Hb,Wb,F+8), 'depth', 'opacity'}; ``pad_dec`` rounds the
    channels of 'dec_in' up to a multiple of 4 (zero pad channels: a float4-aligned pixel for the decoder's first convolution);
    ``dec_one`` writes 1.0 into the first pad channel (a constant-one channel that can carry that convolution's bias; needs a
    pad channel, i.e. (F + 8) % 4 != 0).  ``rows = (lo, hi)`` renders only the bundle-map rows [lo, hi) of every view (the
    image-tile split of a view over several GPUs): the outputs keep their full shape, rows outside the range stay
    uninitialised."""
    _dev(src.tex, src.rgba, vol_cl, depth_range, vol_range, cam, mlp)
    depth_range, vol_range = _f32(depth_range), _f32(vol_range)
    Hb, Wb = H // bundle_size, W // bundle_size
    D = vol_cl.shape[1]
    # a (B,D,Hb,Wb,8) view whose voxels are vol_stride floats apart (the leading channels of a wider channels-last
    # convolution output) is consumed in place, in either memory order: (B,D,Hb,Wb,.) or depth-folded (B,Hb,Wb,D,.)
    vol_layout = 0
    ok = vol_cl.dtype == torch.float32 and vol_cl.shape[-1] == 8 and vol_cl.stride(4) == 1 and vol_cl.data_ptr() % 16 == 0
    vs = vol_cl.stride(3)
    if ok and vs >= 8 and vs % 4 == 0 and vol_cl.stride(2) == vs * Wb and vol_cl.stride(1) == vs * Wb * Hb and vol_cl.stride(0) == vs * Wb * Hb * D:
        vol_layout = 0
    elif ok and vol_cl.stride(1) >= 8 and vol_cl.stride(1) % 4 == 0 and vol_cl.stride(3) == vol_cl.stride(1) * D \
            and vol_cl.stride(2) == vol_cl.stride(3) * Wb and vol_cl.stride(0) == vol_cl.stride(2) * Hb:
        vol_layout, vs = 1, vol_cl.stride(1)
    else:
        vol_cl = _f32(vol_cl)
        vs = 8
    bb = bundle_size * bundle_size
    F = src.feat_dim + 3
    CT = 3 * bb + F + 8
    dev = cam.device
    out_depth = torch.empty((B, Hb, Wb), device=dev, dtype=torch.float32)
    out_opac = torch.empty((B, Hb, Wb), device=dev, dtype=torch.float32)
    if out_channels_last:
        out_feat = torch.empty((B, Hb, Wb, 3 * bb), device=dev, dtype=torch.float32)
        dec_c = (F + 8 + 3) & ~3 if pad_dec else F + 8
        if dec_one and dec_c == F + 8:
            raise _lib.GdbError("render_fused: dec_one needs a pad channel (pad_dec and (F + 8) % 4 != 0)")
        out_dec = torch.empty((B, Hb, Wb, dec_c), device=dev, dtype=torch.float32)
        res = {"fine": out_feat, "dec_in": out_dec, "depth": out_depth, "opacity": out_opac}
    else:
        out_feat = torch.empty((B, CT, Hb, Wb), device=dev, dtype=torch.float32)
        out_dec = None
        res = {"feat": out_feat, "depth": out_depth, "opacity": out_opac}
    tp = None
    if taps is not None:
        S = taps.total
        res["rgbs_feat_dir"] = torch.zeros((V, S, 3 * bb + F + 4), device=dev, dtype=torch.float32)
        res["vox_feat"] = torch.zeros((S, 8), device=dev, dtype=torch.float32)
        res["sigma"] = torch.zeros(S, device=dev, dtype=torch.float32)
        res["sample_feat"] = torch.zeros((S, CT), device=dev, dtype=torch.float32)
        res["weights"] = torch.zeros(S, device=dev, dtype=torch.float32)
        tp = _lib.RenderTaps(taps.offsets.data_ptr(), S, res["rgbs_feat_dir"].data_ptr(), res["vox_feat"].data_ptr(),
                             res["sigma"].data_ptr(), res["sample_feat"].data_ptr(), res["weights"].data_ptr())
    lib = _lib.load()
    _lib.check(lib.gdb_render_fused_fwd(src.rgba.data_ptr(), src.tex.data_ptr(), vol_cl.data_ptr(), depth_range.data_ptr(),
                                        vol_range.data_ptr(), cam.data_ptr(), cam.shape[1], mlp.data_ptr(), B, V, H, W, bundle_size,
                                        src.feat_dim, D, vs, vol_layout, max_samples, src.max_mip, int(inv_depth), int(adaptive), precision,
                                        (2 if dec_one else 1) if out_channels_last else 0, out_dec.shape[-1] if out_dec is not None else 0, out_feat.data_ptr(), _p(out_dec), out_depth.data_ptr(), out_opac.data_ptr(),
                                        rows[0] if rows else 0, rows[1] if rows else Hb, C.byref(tp) if tp is not None else None, _stream()), "gdb_render_fused_fwd")
    return res


def assemble_output(feat: Tensor, dec: Tensor, bdepth: Tensor, bopac: Tensor, bundle_size: int, reweighting: bool,
                    feat_channels_last: bool = False, dec_pre_shuffle: bool = False, dec_bias: Optional[Tensor] = None):
    """rgb = dec + pixel_shuffle(feat[:, :3b^2]) (network.py:175-182).  ``feat`` is (B,CT,Hb,Wb) planar or, with
    ``feat_channels_last``, (B,Hb,Wb,C) holding at least the 3b^2 fine-colour channels; ``dec`` (B,3,H,W) in either
    memory format."""
    _dev(feat, dec, bdepth, bopac)
    layout = 1 if feat_channels_last else 0
    if dec_pre_shuffle:
        # dec: (B, H/2, W/2, 12) dense channels-last output of the decoder's composed last convolution
        if not (dec.dtype == torch.float32 and dec.is_contiguous() and dec.shape[-1] == 12):
            raise ValueError("dec_pre_shuffle needs a contiguous (B,H/2,W/2,12) fp32 tensor")
        layout |= 4
    if dec_bias is not None and not (layout & 4):
        raise ValueError("dec_bias is the bias of the pre-shuffle decoder output (dec_pre_shuffle=True)")
    elif dec.dtype == torch.float32 and not dec.is_contiguous() and dec.permute(0, 2, 3, 1).is_contiguous():
        layout |= 2
    else:
        dec = _f32(dec)
    feat, bdepth, bopac = _f32(feat), _f32(bdepth), _f32(bopac)
    if feat_channels_last:
        B, Hb, Wb, CT = feat.shape
    else:
        B, CT, Hb, Wb = feat.shape
    H, W = Hb * bundle_size, Wb * bundle_size
    rgb = torch.empty((B, 3, H, W), device=feat.device, dtype=torch.float32)
    depth = torch.empty((B, H, W), device=feat.device, dtype=torch.float32)
    opac = torch.empty((B, H, W), device=feat.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.gdb_assemble_output(feat.data_ptr(), CT, dec.data_ptr(), bdepth.data_ptr(), bopac.data_ptr(), B, Hb, Wb, bundle_size,
                                       int(reweighting), layout, _p(None if dec_bias is None else _f32(dec_bias)), rgb.data_ptr(), depth.data_ptr(),
                                       opac.data_ptr(), _stream()),
               "gdb_assemble_output")
    return rgb, depth, opac
