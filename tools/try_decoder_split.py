#!/usr/bin/env python
"""Experiment: dense block's conv2(cat(h, a)) as conv(h; W_h) followed by cuDNN's fused conv(a; W_a) + z + bias + ReLU,
which removes the first torch.cat of the block; and the concatenation of (h, a, b) timed against torch.cat."""
import torch, torch.nn.functional as F
dev = "cuda"
torch.backends.cudnn.benchmark = True
N, H, W = 8, 256, 320
cl = torch.channels_last
h = torch.randn(N, 64, H, W, device=dev).contiguous(memory_format=cl)
a = torch.randn(N, 32, H, W, device=dev).contiguous(memory_format=cl).relu_()
b = torch.randn(N, 32, H, W, device=dev).contiguous(memory_format=cl).relu_()
w2 = (torch.randn(32, 96, 3, 3, device=dev) * 0.05).contiguous(memory_format=cl)
zb = torch.zeros(32, device=dev)
wh, wa = w2[:, :64].contiguous(memory_format=cl), w2[:, 64:].contiguous(memory_format=cl)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=8):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts) * 1e3

ref = lambda: torch.cudnn_convolution_relu(torch.cat((h, a), 1), w2, zb, (1, 1), (1, 1), (1, 1), 1)
def split():
    z = torch.cudnn_convolution(h, wh, (1, 1), (1, 1), (1, 1), 1, False, False, True)
    return torch.cudnn_convolution_add_relu(a, wa, z, 1.0, zb, (1, 1), (1, 1), (1, 1), 1)
print("max diff", float((ref() - split()).abs().max()))
print("cat + conv2 %.0f us, split %.0f us" % (timeit(ref), timeit(split)))
print("cat(h,a) %.0f us, cat(h,a,b) %.0f us" % (timeit(lambda: torch.cat((h, a), 1)), timeit(lambda: torch.cat((h, a, b), 1))))
w1 = (torch.randn(32, 64, 3, 3, device=dev) * 0.05).contiguous(memory_format=cl)
w1h = torch.cat((w1, wh), 0).contiguous(memory_format=cl)
print("conv1 64->32 relu %.0f us, conv 64->32 %.0f us, joint 64->64 %.0f us" % (
    timeit(lambda: torch.cudnn_convolution_relu(h, w1, zb, (1, 1), (1, 1), (1, 1), 1)),
    timeit(lambda: torch.cudnn_convolution(h, wh, (1, 1), (1, 1), (1, 1), 1, False, False, True)),
    timeit(lambda: torch.cudnn_convolution(h, w1h, (1, 1), (1, 1), (1, 1), 1, False, False, True))))
