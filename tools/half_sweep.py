#!/usr/bin/env python
"""Stage-0 cost regularisation (3-D U-Net, cuDNN) with fp32 activations / TF32 math against fp16 activations: is the half path
worth building?  python tools/half_sweep.py"""
import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gdb_nerf_b200.cnn import CostRegNet, cost_reg_fused
torch.backends.cudnn.benchmark = True
dev = "cuda"
torch.manual_seed(0)
net = CostRegNet(32, 8, 8).to(dev).eval()
x = torch.randn(8, 32, 64, 64, 80, device=dev).contiguous(memory_format=torch.channels_last_3d)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return sum(ts) / len(ts)
with torch.no_grad():
    t32 = timed(lambda: cost_reg_fused(net, x, want_volume=False, defer_prob_head=True))
    y32 = cost_reg_fused(net, x, want_volume=False, defer_prob_head=True)[1]
    neth = copy.deepcopy(net).half(); xh = x.half()
    t16 = timed(lambda: cost_reg_fused(neth, xh, want_volume=False, defer_prob_head=True))
    y16 = cost_reg_fused(neth, xh, want_volume=False, defer_prob_head=True)[1]
    torch.backends.cudnn.allow_tf32 = False
    yref = cost_reg_fused(net, x, want_volume=False, defer_prob_head=True)[1]
print(f"stage-0 cost regularisation, 8 views: fp32 activations / TF32 {t32:.3f} ms, fp16 activations {t16:.3f} ms")
print(f"max |err| vs true fp32: TF32 {float((y32 - yref).abs().max()):.3e}, fp16 {float((y16.float() - yref).abs().max()):.3e}; max |y| {float(yref.abs().max()):.3f}")

# ---- decoder (residual dense blocks, 2-D convolutions at bundle-map resolution) and FPN: plain module forward, TF32 vs fp16
from gdb_nerf_b200.cnn import Decoder, FeatureNet, decoder_fused, feature_net_fused
torch.backends.cudnn.allow_tf32 = True
torch.manual_seed(0)
dec = Decoder(16 + 3 + 8, 3, num_feats=64, num_layers=3, upscale_factor=2).to(dev).eval().to(memory_format=torch.channels_last)
xd = torch.randn(8, 27, 256, 320, device=dev).contiguous(memory_format=torch.channels_last)
with torch.no_grad():
    t_plain32 = timed(lambda: dec(xd))
    dech = copy.deepcopy(dec).half(); xdh = xd.half()
    t_plain16 = timed(lambda: dech(xdh))
    xpad = torch.zeros(8, 28, 256, 320, device=dev).contiguous(memory_format=torch.channels_last); xpad[:, :27] = xd
    t_fused32 = timed(lambda: decoder_fused(dec, xpad))
    e16 = float((dech(xdh).float() - dec(xd)).abs().max())
print(f"decoder, 8 views 256x320 -> 512x640: plain module TF32 {t_plain32:.3f} ms, plain module fp16 {t_plain16:.3f} ms, fused path (fp32 / TF32) {t_fused32:.3f} ms; max |fp16 - TF32| {e16:.2e}")
fpn = FeatureNet().to(dev).eval().to(memory_format=torch.channels_last)
xi = torch.rand(24, 3, 512, 640, device=dev).contiguous(memory_format=torch.channels_last)
with torch.no_grad():
    t_f32 = timed(lambda: fpn(xi))
    fpnh = copy.deepcopy(fpn).half(); xih = xi.half()
    t_f16 = timed(lambda: fpnh(xih))
    xi8 = torch.zeros(24, 8, 512, 640, device=dev).contiguous(memory_format=torch.channels_last); xi8[:, :3] = xi
    try:
        t_ff = timed(lambda: feature_net_fused(fpn, xi8, levels=2))
    except Exception as exc:
        t_ff = float("nan"); print("fused FPN:", type(exc).__name__, str(exc)[:200])
print(f"FPN, 24 images 512x640: plain module TF32 {t_f32:.3f} ms, plain module fp16 {t_f16:.3f} ms, fused path (levels=2) {t_ff:.3f} ms")
