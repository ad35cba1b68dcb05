"""The convolutional networks that stay in PyTorch / cuDNN (out of scope of the
hand-written path): 2-D feature pyramid, the two 3-D
cost-regularisation U-Nets and the residual-dense up-sampling decoder.

They are re-declared here only so that checkpoints of the reference load with
``strict=True``: parameter names, shapes and arithmetic follow
networks/gdb_nerf/{feature_net.py:12-64, cost_reg_net.py:8-117,
decoder_rdn.py:7-82, modules.py:5-57}.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# ---------------------------------------------------------------------------
# Inference-time execution of a conv -> batch-norm -> ReLU block: the batch-norm
# affine map is folded into the convolution (exact algebra, eval mode only) and
# the block runs as ONE cuDNN call (conv + bias + ReLU) on channels-last data.
# Host-side plumbing around library kernels; the arithmetic is still cuDNN's.
# ---------------------------------------------------------------------------
def _folded(block: nn.Sequential):
    conv, bn = block[0], block[1]
    key = (conv.weight.data_ptr(), conv.weight._version, bn.weight._version, bn.bias._version, bn.running_mean._version,
           bn.running_var._version, bn.weight.data_ptr())
    cache = getattr(block, "_gdb_fold", None)
    if cache is not None and cache[0] == key:
        return cache[1], cache[2]
    with torch.no_grad():
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        shift = (bn.bias - bn.running_mean * scale).contiguous()
        w = conv.weight
        if isinstance(conv, (nn.ConvTranspose2d, nn.ConvTranspose3d)):
            w = w * scale.view(1, -1, *([1] * (w.dim() - 2)))
        else:
            w = w * scale.view(-1, *([1] * (w.dim() - 1)))
        fmt = torch.channels_last if w.dim() == 4 else torch.channels_last_3d
        w = w.contiguous(memory_format=fmt)
    block._gdb_fold = (key, w, shift)
    return w, shift


def _wview(x: torch.Tensor, f: int) -> torch.Tensor:
    """(N, C, *sp, W) over channels-last memory -> (N, f*C, *sp, W/f) over the SAME memory (f = 1/k un-folds by k)."""
    n, c, w = x.shape[0], x.shape[1], x.shape[-1]
    nd = x.dim()
    cl = x.permute(0, *range(2, nd), 1)
    if f >= 1:
        cl = cl.reshape(n, *x.shape[2:-1], w // f, f * c)
    else:
        k = round(1 / f)
        cl = cl.reshape(n, *x.shape[2:-1], w * k, c // k)
    return cl.permute(0, nd - 1, *range(1, nd - 1))


def _wfolded(block: nn.Sequential, f_in: int):
    """Width-folded weight / bias of a conv+BN+ReLU block (see fold_width_weight), cached."""
    conv = block[0]
    w, shift = _folded(block)

    def make():
        s = conv.stride[-1]
        w2, kw = fold_width_weight(w, f_in, s)
        fmt = torch.channels_last if w2.dim() == 4 else torch.channels_last_3d
        return w2.contiguous(memory_format=fmt), shift.repeat(f_in // s).contiguous(), kw

    return _cached(block, f"_gdb_wfold{f_in}", (w, shift), make)


def can_wfold(x: torch.Tensor, f: int) -> bool:
    return x.is_cuda and x.dtype in (torch.float32, torch.float16) and x.shape[-1] % f == 0 and x.permute(0, *range(2, x.dim()), 1).is_contiguous()


def run_block(block: nn.Sequential, x: torch.Tensor, wfold: int = 0) -> torch.Tensor:
    """conv+BN+ReLU block in eval mode, BN folded, fused where cuDNN offers it.  With ``wfold`` = f the input arrives in
    the f-fold width view (N, f*Ci, .., W/f) and the output leaves in the (f/stride)-fold view: same memory as the plain
    channels-last result, but the 8/16-channel layer runs as a 32-channel one (2-3x faster in cuDNN, tools/fold_sweep.py)."""
    conv = block[0]
    if wfold:
        w, b, kw = _wfolded(block, wfold)
        nd = w.dim() - 2
        stride = tuple(conv.stride[:-1]) + (1,)
        pad = tuple(conv.padding[:-1]) + (kw // 2,)
        global _FUSED_3D
        if nd == 2 or _FUSED_3D is not False:
            try:
                y = torch.cudnn_convolution_relu(x, w, b, stride, pad, (1,) * nd, 1)
                if nd == 3:
                    _FUSED_3D = True
                return y
            except RuntimeError:
                if nd == 2 or _FUSED_3D is True:
                    raise
                _FUSED_3D = False
        return F.conv3d(x, w, b, stride, pad).relu_()
    w, b = _folded(block)
    if isinstance(conv, nn.ConvTranspose3d):
        return F.conv_transpose3d(x, w, b, conv.stride, conv.padding, conv.output_padding).relu_()
    if isinstance(conv, nn.Conv2d):
        return torch.cudnn_convolution_relu(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
    if _FUSED_3D is not False:
        try:
            y = torch.cudnn_convolution_relu(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
            _FUSED_3D = True
            return y
        except RuntimeError:
            if _FUSED_3D is True:
                raise
            _FUSED_3D = False     # this cuDNN build has no fused 3-D conv+bias+ReLU: plain conv + in-place ReLU
    return F.conv3d(x, w, b, conv.stride, conv.padding).relu_()


_FUSED_3D = None


def _cbr(conv: nn.Module, norm: nn.Module) -> nn.Sequential:
    # conv -> batch-norm -> ReLU, indices 0/1/2 as in the reference's block builders
    return nn.Sequential(conv, norm, nn.ReLU(inplace=True))


def _c2(cin: int, cout: int, k: int, stride: int = 1) -> nn.Sequential:
    return _cbr(nn.Conv2d(cin, cout, k, stride, k // 2, bias=False), nn.BatchNorm2d(cout))


def _c3(cin: int, cout: int, stride: int = 1) -> nn.Sequential:
    return _cbr(nn.Conv3d(cin, cout, 3, stride, 1, bias=False), nn.BatchNorm3d(cout))


def _d3(cin: int, cout: int) -> nn.Sequential:
    return _cbr(nn.ConvTranspose3d(cin, cout, 3, 2, 1, 1, bias=False), nn.BatchNorm3d(cout))


class FeatureNet(nn.Module):
    """Three-level FPN; returns [1/4 res, 1/2 res, full res] feature maps."""

    def __init__(self, base_channels: int = 8, out_channels: Sequence[int] = (32, 16, 8)) -> None:
        super().__init__()
        c = base_channels
        self.conv0 = nn.Sequential(_c2(3, c, 3), _c2(c, c, 3))
        self.conv1 = nn.Sequential(_c2(c, 2 * c, 5, 2), _c2(2 * c, 2 * c, 3))
        self.conv2 = nn.Sequential(_c2(2 * c, 4 * c, 5, 2), _c2(4 * c, 4 * c, 3))
        self.out0 = nn.Conv2d(4 * c, out_channels[0], 1)
        self.inner1 = nn.Conv2d(2 * c, 4 * c, 1)
        self.inner2 = nn.Conv2d(c, 4 * c, 1)
        self.out1 = nn.Conv2d(4 * c, out_channels[1], 3, padding=1, bias=False)
        self.out2 = nn.Conv2d(4 * c, out_channels[2], 3, padding=1, bias=False)

    def forward(self, x: torch.Tensor, levels: int = 3) -> List[torch.Tensor]:
        """``levels`` < 3 skips the finer lateral branches nobody consumes
        (the full-resolution level is never read on the rendering path)."""
        f0 = self.conv0(x)
        f1 = self.conv1(f0)
        f2 = self.conv2(f1)
        outs = [self.out0(f2)]
        if levels >= 2:
            top = F.interpolate(f2, size=f1.shape[-2:], mode="nearest") + self.inner1(f1)
            outs.append(self.out1(top))
            if levels >= 3:
                top = F.interpolate(top, size=f0.shape[-2:], mode="nearest") + self.inner2(f0)
                outs.append(self.out2(top))
        return outs


def _cached(mod: nn.Module, name: str, tensors, make):
    """Derived eval-time parameters (composed / concatenated weights), rebuilt when a source tensor changes."""
    key = tuple((t.data_ptr(), t._version) for t in tensors)
    cache = getattr(mod, name, None)
    if cache is not None and cache[0] == key:
        return cache[1]
    with torch.no_grad():
        val = make()
    object.__setattr__(mod, name, (key, val))
    return val


def feature_net_fused(net: "FeatureNet", x: torch.Tensor, levels: int = 3) -> List[torch.Tensor]:
    from . import ops
    blk0 = net.conv0[0]
    c = blk0[0].out_channels
    # width-folded execution of the 8/16-channel layers: every one of them becomes a 32 -> 32 channel convolution over
    # the same channels-last memory (conv0.* fold 4, conv1.0 fold 4 -> 2, conv1.1 fold 2, conv2.0 fold 2 -> 1)
    wf = x.is_cuda and x.dtype == torch.float32 and x.shape[-1] % 4 == 0 and c == 8 and blk0[0].groups == 1
    if x.is_cuda and x.dtype == torch.float32 and x.shape[1] < 8 and blk0[0].groups == 1:
        # RGB planes -> 8-channel channels-last pixels (zero pad) in one pass; the first convolution gets zero weights for
        # the pad: cuDNN then needs no layout conversion and picks a 1.6x faster kernel than for 3 input channels
        c_in = x.shape[1]
        x = ops.to_channels_last(x.contiguous(), 8).permute(0, 3, 1, 2)
        w, b = _folded(blk0)
        w8 = _cached(blk0, "_gdb_pad8", (w,), lambda: F.pad(w, (0, 0, 0, 0, 0, 8 - c_in)).contiguous(memory_format=torch.channels_last))
        conv = blk0[0]
        if wf:
            w32, b32 = _cached(blk0, "_gdb_pad8_wfold4", (w8, b), lambda: (
                fold_width_weight(w8, 4, 1)[0].contiguous(memory_format=torch.channels_last), b.repeat(4).contiguous()))
            f0 = torch.cudnn_convolution_relu(_wview(x, 4), w32, b32, conv.stride, conv.padding, conv.dilation, 1)
        else:
            f0 = torch.cudnn_convolution_relu(x, w8, b, conv.stride, conv.padding, conv.dilation, 1)
    else:
        x = x.contiguous(memory_format=torch.channels_last)
        f0 = run_block(blk0, x)
        if wf:
            f0 = _wview(f0, 4)
    if wf:
        f0 = run_block(net.conv0[1], f0, wfold=4)                                       # (N, 32, H, W/4)   = 8 ch x 4 px
        f1 = run_block(net.conv1[1], run_block(net.conv1[0], f0, wfold=4), wfold=2)     # (N, 32, H/2, W/4) = 16 ch x 2 px
        f2 = run_block(net.conv2[1], run_block(net.conv2[0], f1, wfold=2))              # (N, 32, H/4, W/4)
        lat1 = None
        if levels >= 2:     # lateral 1x1 convolution 16 -> 32 as 32 -> 64 on the fold-2 view
            lat1 = _wview(F.conv2d(f1, _cached(net.inner1, "_gdb_wfold2", (net.inner1.weight,), lambda: fold_width_weight(
                net.inner1.weight, 2, 1)[0].contiguous(memory_format=torch.channels_last))), 0.5)
        f0, f1 = _wview(f0, 0.25), _wview(f1, 0.5)
    else:
        f0 = run_block(net.conv0[1], f0)
        f1 = run_block(net.conv1[1], run_block(net.conv1[0], f0))
        f2 = run_block(net.conv2[1], run_block(net.conv2[0], f1))
        lat1 = None
    if ops._is_cl(f2) and net.out0.out_channels % 4 == 0:      # 1x1 output convolution, bias through the vectorised epilogue
        outs = [ops.bias_act_add(F.conv2d(f2, net.out0.weight), net.out0.bias, None, relu=False)]
    else:
        outs = [net.out0(f2)]
    if levels >= 2:
        # top-down step (feature_net.py:52-58): nearest x2 + lateral 1x1 conv + its bias in one pass
        top = ops.bias_act_add(F.conv2d(f1, net.inner1.weight) if lat1 is None else lat1, net.inner1.bias, f2, relu=False, skip_up2=True)
        outs.append(net.out1(top))
        if levels >= 3:
            top = ops.bias_act_add(F.conv2d(f0, net.inner2.weight), net.inner2.bias, top, relu=False, skip_up2=True)
            outs.append(net.out2(top))
    return outs


def _deconv_skip(block: nn.Sequential, x: torch.Tensor, skip: torch.Tensor, wfold: int = 0) -> torch.Tensor:
    """skip + relu(bn(deconv(x))) (cost_reg_net.py:108-110): library transposed convolution, one fused epilogue.  With
    ``wfold`` = f the input arrives in the f-fold width view and skip / result are in the 2f-fold view."""
    from . import ops
    conv = block[0]
    w, b = _folded(block)
    if wfold:
        w, b = _cached(block, f"_gdb_wfoldT{wfold}", (w, b), lambda: (
            fold_width_weight_transposed(w, wfold)[0].contiguous(memory_format=torch.channels_last_3d), b.repeat(2 * wfold).contiguous()))
        d = F.conv_transpose3d(x, w, None, tuple(conv.stride[:-1]) + (1,), conv.padding, tuple(conv.output_padding[:-1]) + (0,))
    else:
        d = F.conv_transpose3d(x, w, None, conv.stride, conv.padding, conv.output_padding)
    if d.dtype == torch.float32 and ops._is_cl(d) and ops._is_cl(skip) and d.shape[1] % 4 == 0:
        return ops.bias_act_add(d, b, skip, relu=True)
    return skip + (d + b.view(1, -1, 1, 1, 1)).relu_()


def cost_reg_fused(net: "_CostReg", x: torch.Tensor, want_volume: bool = True,
                   defer_prob_head: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Same data flow as CostRegNet(Small).forward with folded blocks; x is NCDHW-shaped, channels_last_3d strides.
    Returns (volume, logits): volume is a (B,D,H,W,8) channels-last VIEW of the joint head output (or None when the
    caller does not consume it: the stage-0 feature head only feeds the training-time coarse render), logits a
    (B,D,H,W) view of the probability head BEFORE its soft-max (the soft-max is fused into the depth-range kernel)."""
    r = run_block
    if can_wfold(x, 4) and net.conv0[0].out_channels == 8 and net.conv0[0].stride[-1] == 1:
        # width-folded execution of the 8/16-channel levels (same memory, 32-channel layers; see fold_width_weight)
        s0 = r(net.conv0, _wview(x, 4), wfold=4)                       # 8 ch x 4 px
        s1 = r(net.conv2, r(net.conv1, s0, wfold=4), wfold=2)          # 16 ch x 2 px
        t3 = r(net.conv3, s1, wfold=2)                                 # 32 ch
        f1, f2 = 1, 2     # the two finest up-sampling layers stay in the folded views (16 ch x 2 px, 8 ch x 4 px)
    else:
        s0 = r(net.conv0, x)
        s1 = r(net.conv2, r(net.conv1, s0))
        t3 = r(net.conv3, s1)
        f1 = f2 = 0
    if isinstance(net, CostRegNetSmall):
        y = r(net.conv4, t3)
        y = _deconv_skip(net.conv5, y, s1, f1)
        y = _deconv_skip(net.conv6, y, s0, f2)
    else:
        s2 = r(net.conv4, t3)
        y = r(net.conv6, r(net.conv5, s2))
        y = _deconv_skip(net.conv7, y, s2)
        y = _deconv_skip(net.conv8, y, s1, f1)
        y = _deconv_skip(net.conv9, y, s0, f2)
    if f2:
        y = _wview(y, 0.25)
    if not want_volume:
        # the caller fuses the 1-channel probability head into the depth-range kernel when it can (ops.prob_head_depth_range)
        if defer_prob_head:
            return None, y
        return None, F.conv3d(y, net.prob_head.weight, None, 1, 1).squeeze(1)
    # both heads as ONE convolution (identical arithmetic per output channel), padded to 12 output channels so the
    # channels-last voxel stays float4-addressable: channels 0-7 feature volume, 8 probability logits
    fw, pw = net.feat_head.weight, net.prob_head.weight
    cout = fw.shape[0]
    cpad = (cout + 1 + 3) & ~3

    def make():
        w = torch.zeros((cpad, *fw.shape[1:]), device=fw.device, dtype=fw.dtype)
        w[:cout] = fw
        w[cout] = pw[0]
        return w.contiguous(memory_format=torch.channels_last_3d)

    w = _cached(net, "_gdb_heads", (fw, pw), make)
    out = F.conv3d(y, w, None, 1, 1)                             # (B,cpad,D,H,W) over NDHWC memory
    cl = out.permute(0, 2, 3, 4, 1)
    if not cl.is_contiguous():
        cl = cl.contiguous()
    return cl[..., :cout], cl[..., cout]


# ---------------------------------------------------------------------------
# Depth-folded execution of the 3-D U-Net (eval mode): with D <= 8 hypotheses the depth axis is folded into the
# channels (channel = d*C + c) and every 3x3x3 convolution becomes a 3x3 2-D convolution whose weight is the
# block-Toeplitz expansion of the 3-D kernel along depth (zero blocks where |d_in - s*d_out| > 1, which is also the
# zero padding).  Same products, same BN folding; the 64-128 channel 2-D shapes are the ones cuDNN runs efficiently,
# the 8/16-channel 3-D shapes are not (stage-1 conv0: 0.95 ms as Conv3d 16->8, 0.16 ms as Conv2d 128->64).
# ---------------------------------------------------------------------------
def fold_depth_weight(w3: torch.Tensor, d_in: int, stride: int, transposed: bool) -> Tuple[torch.Tensor, int]:
    """(Co,Ci,3,3,3) Conv3d weight (padding 1, isotropic ``stride``) -> (d_out*Co, d_in*Ci, 3, 3) Conv2d weight, or
    (Ci,Co,3,3,3) ConvTranspose3d weight (stride 2, padding 1, output_padding 1) -> (d_in*Ci, d_out*Co, 3, 3)
    ConvTranspose2d weight.  Returns (weight, d_out)."""
    if transposed:
        ci, co = w3.shape[:2]
        d_out = 2 * d_in
        w2 = w3.new_zeros((d_in, ci, d_out, co, 3, 3))
        for di in range(d_in):
            for kd in range(3):
                do = 2 * di - 1 + kd
                if 0 <= do < d_out:
                    w2[di, :, do, :] = w3[:, :, kd]
        return w2.reshape(d_in * ci, d_out * co, 3, 3), d_out
    co, ci = w3.shape[:2]
    d_out = (d_in + 2 - 3) // stride + 1
    w2 = w3.new_zeros((d_out, co, d_in, ci, 3, 3))
    for do in range(d_out):
        for kd in range(3):
            di = do * stride + kd - 1
            if 0 <= di < d_in:
                w2[do, :, di, :] = w3[:, :, kd]
    return w2.reshape(d_out * co, d_in * ci, 3, 3), d_out


def fold_width_weight(w: torch.Tensor, f_in: int, stride: int = 1) -> Tuple[torch.Tensor, int]:
    """Width-folded form of a convolution weight.  ``f_in`` neighbouring pixels of an input row are read as ``f_in*Ci``
    channels of one pixel (channel = r*Ci + ci; for channels-last data this is the SAME memory, (N,H,W,C) viewed as
    (N,H,W/f_in,f_in*C)), the output likewise with f_out = f_in/stride pixels per group.  (Co,Ci,[kd,]kh,kw) with
    padding k//2 and ``stride`` along the width -> (f_out*Co, f_in*Ci, [kd,]kh, KW): the block-Toeplitz expansion of the
    kernel along the row, zero where a (group, pixel) pair lies outside the kernel's reach - which is also the zero
    padding, so W % f_in == 0 is the only condition.  The folded convolution keeps the original stride/padding on every
    other axis and runs with stride 1 and padding KW//2 along the folded width.  Same products as the original (plus
    exact zeros); 8/16-channel layers become 32/64-channel layers, the shapes cuDNN has efficient kernels for.
    Returns (weight, KW)."""
    if f_in % stride:
        raise ValueError("fold_width_weight: the stride must divide the fold")
    co, ci, k = w.shape[0], w.shape[1], w.shape[-1]
    p, f_out = k // 2, f_in // stride
    lo = (-p) // f_in                                   # group offsets reached by the kernel: floor(t / f_in)
    hi = (stride * (f_out - 1) + k - 1 - p) // f_in
    half = max(-lo, hi)
    kw2 = 2 * half + 1
    mid = w.shape[2:-1]
    w2 = w.new_zeros((f_out, co, f_in, ci, *mid, kw2))
    for q in range(f_out):
        for kx in range(k):
            t = stride * q + kx - p
            dg, r = t // f_in, t % f_in
            w2[q, :, r, :, ..., dg + half] = w[..., kx]
    return w2.reshape(f_out * co, f_in * ci, *mid, kw2), kw2


def fold_width_weight_transposed(w: torch.Tensor, f_in: int) -> Tuple[torch.Tensor, int]:
    """Width-folded form of a transposed-convolution weight (Ci,Co,[kd,]kh,3) with stride 2, padding 1, output_padding 1
    (the U-Net's up-sampling layers): input groups of ``f_in`` pixels map to output groups of 2*f_in pixels, so along the
    folded width the layer is a stride-1 transposed convolution with a 3-tap kernel (padding 1, output_padding 0; one tap
    stays zero).  Returns ((f_in*Ci, 2*f_in*Co, [kd,]kh,3), 3)."""
    ci, co, k = w.shape[0], w.shape[1], w.shape[-1]
    if k != 3:
        raise ValueError("fold_width_weight_transposed: 3-tap kernels only")
    f_out = 2 * f_in
    mid = w.shape[2:-1]
    w2 = w.new_zeros((f_in, ci, f_out, co, *mid, 3))
    for r in range(f_in):
        for kx in range(3):
            t = 2 * r - 1 + kx                      # output pixel relative to the start of the input group's output group
            dh, q = t // f_out, t % f_out
            w2[r, :, q, :, ..., dh + 1] = w[..., kx]
    return w2.reshape(f_in * ci, f_out * co, *mid, 3), 3


def _folded2d(block: nn.Sequential, d_in: int):
    conv = block[0]
    w3, shift = _folded(block)
    transposed = isinstance(conv, nn.ConvTranspose3d)

    def make():
        w2, d_out = fold_depth_weight(w3, d_in, conv.stride[0], transposed)
        return w2.contiguous(memory_format=torch.channels_last), shift.repeat(d_out).contiguous(), d_out

    return _cached(block, f"_gdb_dfold{d_in}", (w3, shift), make)


def cost_reg_folded(net: "_CostReg", x: torch.Tensor, D: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """CostRegNet(Small).forward on a depth-folded volume: x is (B, D*C, H, W)-shaped over (B,H,W,D,C) memory.
    Returns (volume, logits) as (B,D,H,W,8) / (B,D,H,W) VIEWS of the joint head output, whose memory is (B,H,W,D,12)."""
    from . import ops

    def r(block, t, d):
        w, b, d_out = _folded2d(block, d)
        s = block[0].stride[0]
        return torch.cudnn_convolution_relu(t, w, b, (s, s), (1, 1), (1, 1), 1), d_out

    def up(block, t, skip, d):
        w, b, d_out = _folded2d(block, d)
        y = F.conv_transpose2d(t, w, None, 2, 1, 1)
        return ops.bias_act_add(y, b, skip, relu=True), d_out

    s0, d0 = r(net.conv0, x, D)
    t, d1 = r(net.conv1, s0, d0)
    s1, d1 = r(net.conv2, t, d1)
    if isinstance(net, CostRegNetSmall):
        t, d2 = r(net.conv3, s1, d1)
        y, d2 = r(net.conv4, t, d2)
        y, d = up(net.conv5, y, s1, d2)
        y, d = up(net.conv6, y, s0, d)
    else:
        t, d2 = r(net.conv3, s1, d1)
        s2, d2 = r(net.conv4, t, d2)
        t, d3 = r(net.conv5, s2, d2)
        y, d3 = r(net.conv6, t, d3)
        y, d = up(net.conv7, y, s2, d3)
        y, d = up(net.conv8, y, s1, d)
        y, d = up(net.conv9, y, s0, d)
    if d != D:
        raise ValueError(f"depth-folded cost regularisation needs D divisible by the U-Net's total stride (D = {D})")
    fw, pw = net.feat_head.weight, net.prob_head.weight
    cout = fw.shape[0]
    cpad = (cout + 1 + 3) & ~3

    def make():
        w = torch.zeros((cpad, *fw.shape[1:]), device=fw.device, dtype=fw.dtype)
        w[:cout] = fw
        w[cout] = pw[0]
        return fold_depth_weight(w, D, 1, False)[0].contiguous(memory_format=torch.channels_last)

    w = _cached(net, f"_gdb_heads2d{D}", (fw, pw), make)
    out = F.conv2d(y, w, None, 1, 1)                              # (B, D*cpad, H, W) over (B,H,W,D,cpad) memory
    B, _, H, W = out.shape
    cl = out.permute(0, 2, 3, 1)
    if not cl.is_contiguous():
        cl = cl.contiguous()
    cl = cl.view(B, H, W, D, cpad)
    return cl[..., :cout].permute(0, 3, 1, 2, 4), cl[..., cout].permute(0, 3, 1, 2)


class _CostReg(nn.Module):
    def _heads(self, c: int, cout: int) -> None:
        self.feat_head = nn.Conv3d(c, cout, 3, padding=1, bias=False)
        self.prob_head = nn.Conv3d(c, 1, 3, padding=1, bias=False)

    def _finish(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.feat_head(x), torch.softmax(self.prob_head(x).squeeze(1), dim=1)


class CostRegNetSmall(_CostReg):
    """Two-level 3-D U-Net (stage 0)."""

    def __init__(self, cin: int, cout: int, c: int) -> None:
        super().__init__()
        self.conv0 = _c3(cin, c)
        self.conv1 = _c3(c, 2 * c, 2)
        self.conv2 = _c3(2 * c, 2 * c)
        self.conv3 = _c3(2 * c, 4 * c, 2)
        self.conv4 = _c3(4 * c, 4 * c)
        self.conv5 = _d3(4 * c, 2 * c)
        self.conv6 = _d3(2 * c, c)
        self._heads(c, cout)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        s0 = self.conv0(x)
        s1 = self.conv2(self.conv1(s0))
        y = self.conv4(self.conv3(s1))
        y = s1 + self.conv5(y)
        y = s0 + self.conv6(y)
        return self._finish(y)


class CostRegNet(_CostReg):
    """Three-level 3-D U-Net (later stages)."""

    def __init__(self, cin: int, cout: int, c: int) -> None:
        super().__init__()
        self.conv0 = _c3(cin, c)
        self.conv1 = _c3(c, 2 * c, 2)
        self.conv2 = _c3(2 * c, 2 * c)
        self.conv3 = _c3(2 * c, 4 * c, 2)
        self.conv4 = _c3(4 * c, 4 * c)
        self.conv5 = _c3(4 * c, 8 * c, 2)
        self.conv6 = _c3(8 * c, 8 * c)
        self.conv7 = _d3(8 * c, 4 * c)
        self.conv8 = _d3(4 * c, 2 * c)
        self.conv9 = _d3(2 * c, c)
        self._heads(c, cout)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        s0 = self.conv0(x)
        s1 = self.conv2(self.conv1(s0))
        s2 = self.conv4(self.conv3(s1))
        y = self.conv6(self.conv5(s2))
        y = s2 + self.conv7(y)
        y = s1 + self.conv8(y)
        y = s0 + self.conv9(y)
        return self._finish(y)


class _SqueezeExcite(nn.Module):
    def __init__(self, channels: int, reduction: int = 16) -> None:
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(channels, channels // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(channels // reduction, channels, bias=False), nn.Sigmoid())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        gate = self.fc(self.avg_pool(x).flatten(1))
        return x * gate[:, :, None, None]


class _DenseBlock(nn.Module):
    def __init__(self, channels: int, growth: int = 32) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(channels, growth, 3, padding=1, bias=False)
        self.conv2 = nn.Conv2d(channels + growth, growth, 3, padding=1, bias=False)
        self.conv3 = nn.Conv2d(channels + 2 * growth, channels, 3, padding=1, bias=False)
        self.se = _SqueezeExcite(channels)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        a = F.relu(self.conv1(x))
        b = F.relu(self.conv2(torch.cat((x, a), 1)))
        return x + self.se(self.conv3(torch.cat((x, a, b), 1)))


class Decoder(nn.Module):
    """Residual-dense network + pixel-shuffle up-sampler (x upscale_factor)."""

    def __init__(self, in_channels: int, out_channels: int, num_feats: int, num_layers: int, upscale_factor: int) -> None:
        super().__init__()
        if upscale_factor <= 0 or upscale_factor & (upscale_factor - 1):
            raise ValueError("`upscale_factor` must be a power of 2.")
        self.upscale_factor = upscale_factor
        self.in_conv = nn.Conv2d(in_channels, num_feats, 3, padding=1)
        self.blocks = nn.Sequential(*(_DenseBlock(num_feats) for _ in range(num_layers)))
        ups: List[nn.Module] = []
        for _ in range(int(round(math.log2(upscale_factor)))):
            ups += [nn.Conv2d(num_feats, 4 * num_feats, 3, padding=1), nn.PixelShuffle(2)]
        self.up = nn.Sequential(*ups)
        self.out_conv = nn.Conv2d(num_feats, out_channels, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.in_conv(x)
        return self.out_conv(self.up(y + self.blocks(y)))


def decoder_fused(dec: "Decoder", x: torch.Tensor, one_channel: int = -1) -> torch.Tensor:
    """Eval-time decoder (decoder_rdn.py:44-82) on channels-last data.  ``one_channel``: index of a constant-one pad channel of x
    that carries the first convolution's bias (-1: none, the bias is added by a separate pass).  Returns (out, bias): the output BEFORE the last pixel
    shuffle and WITHOUT its bias as (B, H/2, W/2, 12) channels-last plus that bias (12,), applied by ``gdb_assemble_output``: the last up-sampling convolution, the pixel shuffle and the 1x1
    output convolution are all linear, so the 1x1 is folded into the 3x3 (64 -> 3*4 channels instead of 64 -> 256,
    no 256-channel intermediate); ``gdb_assemble_output`` performs the pending shuffle."""
    from . import ops
    if ops._is_cl(x) or x.is_cuda:
        w_in = dec.in_conv.weight
        cin = w_in.shape[1]
        if x.shape[1] > cin:        # decoder input with pad channels (float4-aligned pixels from the render kernel)
            extra = x.shape[1] - cin

            def make():
                w = F.pad(dec.in_conv.weight, (0, 0, 0, 0, 0, extra))
                if one_channel >= cin:
                    # channel `one_channel` of x is the constant 1 (render kernel, out_channels_last = 2): its centre tap carries
                    # the bias - the centre tap always lies inside the image, so the zero padding does not touch it
                    kh, kw = w.shape[2] // 2, w.shape[3] // 2
                    w[:, one_channel, kh, kw] = dec.in_conv.bias
                return w.contiguous(memory_format=torch.channels_last)

            w_in = _cached(dec, f"_gdb_inpad{extra}_{one_channel}", (w_in, dec.in_conv.bias), make)
        y = F.conv2d(x, w_in, None, 1, dec.in_conv.padding)
        if not (x.shape[1] > cin and one_channel >= cin):
            y = ops.bias_act_add(y, dec.in_conv.bias, None, relu=False) if ops._is_cl(y) and y.shape[1] % 4 == 0 else y + dec.in_conv.bias.view(1, -1, 1, 1)
    else:
        y = dec.in_conv(x)
    h = y
    cat_buf = None
    for blk in dec.blocks:
        zb = _cached(blk, "_gdb_zero", (blk.conv1.weight,), lambda: torch.zeros(blk.conv1.weight.shape[0], device=x.device))
        a = torch.cudnn_convolution_relu(h, blk.conv1.weight, zb, (1, 1), (1, 1), (1, 1), 1)
        if ops._is_cl(h) and ops._is_cl(a) and h.shape[1] % 4 == 0 and a.shape[1] % 4 == 0:
            # conv2(cat(h, a)) = relu(conv(a; W_a) + conv(h; W_h)): the second term is cuDNN's fused conv + add + ReLU, so
            # the first concatenation of the block is never materialised; the second one is one streaming pass
            ch = h.shape[1]
            wh, wa = _cached(blk, "_gdb_split2", (blk.conv2.weight,), lambda: (
                blk.conv2.weight[:, :ch].contiguous(memory_format=torch.channels_last),
                blk.conv2.weight[:, ch:].contiguous(memory_format=torch.channels_last)))
            z = torch.cudnn_convolution(h, wh, (1, 1), (1, 1), (1, 1), 1, False, False, torch.backends.cudnn.allow_tf32)
            b = torch.cudnn_convolution_add_relu(a, wa, z, 1.0, zb, (1, 1), (1, 1), (1, 1), 1)
            if cat_buf is not None:      # h already sits in the buffer's leading channels (written by the previous block)
                c = blk.conv3(ops.concat_into(cat_buf, ch, a, b))
            else:
                c = blk.conv3(ops.concat_channels(h, a, b))
            cat_buf = None
            gate = None if ops._is_cl(c) else blk.se.fc(c.mean((2, 3)))
        else:
            cat_buf = None
            b = torch.cudnn_convolution_relu(torch.cat((h, a), 1), blk.conv2.weight, zb, (1, 1), (1, 1), (1, 1), 1)
            c = blk.conv3(torch.cat((h, a, b), 1))
            gate = blk.se.fc(c.mean((2, 3)))
        last = blk is dec.blocks[-1]
        if ops._is_cl(h) and ops._is_cl(c):
            # the last block also adds the decoder's global residual y (decoder_rdn.py: y + blocks(y)) in the same pass
            extra = y if last and ops._is_cl(y) else None
            if gate is None:      # squeeze-excite gate finished in the prologue of the residual kernel
                if not last:      # ... which also writes the next block's h slice of its concatenation buffer
                    N_, C_, H_, W_ = h.shape
                    cat_buf = torch.empty((N_, H_, W_, C_ + 2 * blk.conv1.weight.shape[0]), device=h.device, dtype=h.dtype).permute(0, 3, 1, 2)
                h = ops.se_gate_add(h, c, blk.se.fc[0].weight, blk.se.fc[2].weight, extra=extra, cat_out=cat_buf)
            else:
                h = ops.gate_add(h, c, gate, extra=extra)
            if last and not ops._is_cl(y):
                h = h + y
        else:
            h = h + c * gate[:, :, None, None]
            if last:
                h = h + y
    y = h
    mods = list(dec.up)
    if not mods:
        raise ValueError("decoder_fused needs upscale_factor >= 2")
    for i in range(0, len(mods) - 2, 2):
        conv = mods[i]
        if ops._is_cl(y) and conv.weight.shape[0] % 16 == 0:
            # up-sampling stage: convolution without its bias, then bias + PixelShuffle(2) in one channels-last pass
            y = ops.pixel_shuffle2_bias(F.conv2d(y, conv.weight, None, 1, 1), conv.bias)
        else:
            y = mods[i + 1](conv(y))
    up, oc = mods[-2], dec.out_conv

    def make():
        co, ci = oc.weight.shape[:2]
        wu = up.weight.view(ci, 4, *up.weight.shape[1:])                          # [c][k] <- channel c*4 + k
        w = torch.einsum("oc,ckihw->okihw", oc.weight.view(co, ci), wu).reshape(co * 4, *up.weight.shape[1:])
        b = torch.einsum("oc,ck->ok", oc.weight.view(co, ci), up.bias.view(ci, 4)) + oc.bias[:, None]
        return w.contiguous(memory_format=torch.channels_last), b.reshape(-1).contiguous()

    w, b = _cached(dec, "_gdb_tail", (up.weight, up.bias, oc.weight, oc.bias), make)
    out = F.conv2d(y, w, None, 1, 1)                                              # (B,12,h,w) over NHWC memory; bias applied by the consumer
    return out.permute(0, 2, 3, 1).contiguous(), b

