#!/usr/bin/env python
"""Experiment: can cuDNN write a convolution's output into a channel slice of a wider channels-last buffer
(aten::cudnn_convolution.out), so the dense block's torch.cat copies disappear?"""
import torch, time
dev = "cuda"
torch.backends.cudnn.benchmark = True
N, H, W = 8, 256, 320
buf = torch.empty(N, H, W, 128, device=dev)                     # NHWC memory
x = torch.randn(N, 64, H, W, device=dev).contiguous(memory_format=torch.channels_last)
w1 = torch.randn(32, 64, 3, 3, device=dev).contiguous(memory_format=torch.channels_last) * 0.05
nchw = buf.permute(0, 3, 1, 2)                                  # (N,128,H,W) view over NHWC memory
nchw[:, :64].copy_(x)
ref = torch.cudnn_convolution(x, w1, (1, 1), (1, 1), (1, 1), 1, False, False, True)
out_slice = nchw[:, 64:96]
try:
    torch.ops.aten.cudnn_convolution.out(x, w1, (1, 1), (1, 1), (1, 1), 1, False, False, True, out=out_slice)
    torch.cuda.synchronize()
    print("out= into slice: max diff", float((out_slice - ref).abs().max()), "strides", out_slice.stride(), "same storage", out_slice.data_ptr() == buf.data_ptr() + 64 * 4)
except Exception as e:
    print("out= into slice FAILED:", repr(e)[:300])
# input as a channel-slice view (non-dense channels-last)
w2 = torch.randn(32, 96, 3, 3, device=dev).contiguous(memory_format=torch.channels_last) * 0.05
inp = nchw[:, :96]
try:
    a = torch.cudnn_convolution(inp, w2, (1, 1), (1, 1), (1, 1), 1, False, False, True)
    b = torch.cudnn_convolution(inp.contiguous(memory_format=torch.channels_last), w2, (1, 1), (1, 1), (1, 1), 1, False, False, True)
    torch.cuda.synchronize()
    print("strided input: max diff", float((a - b).abs().max()))
    for fn, name in ((lambda: torch.cudnn_convolution(inp, w2, (1, 1), (1, 1), (1, 1), 1, False, False, True), "strided-input conv"),
                     (lambda: torch.cudnn_convolution(torch.cat((x, ref), 1), w2, (1, 1), (1, 1), (1, 1), 1, False, False, True), "cat + conv")):
        for _ in range(3): fn()
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(20): fn()
        torch.cuda.synchronize(); print(name, (time.perf_counter() - t) / 20 * 1e3, "ms")
except Exception as e:
    print("strided input FAILED:", repr(e)[:300])
