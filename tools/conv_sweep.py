#!/usr/bin/env python
"""Experiment: do the thin (<= 16 channel) full-resolution convolutions of the FPN / cost-regularisation nets run
faster through cuDNN with zero-padded channel counts?  (Same arithmetic: padded weights are zero.)"""
import itertools, sys, torch, torch.nn.functional as F
dev = "cuda"
torch.backends.cudnn.benchmark = True
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=6):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts) * 1e3

LAYERS = [  # name, dims, N, cin, cout, spatial, k, stride
    ("cr0.prob_head", 3, 8, 8, 1, (64, 64, 80), 3, 1),
    ("cr0.conv6 (deconv out) as conv", 3, 8, 8, 8, (64, 64, 80), 3, 1),
    ("fpn.conv0.0", 2, 24, 3, 8, (512, 640), 3, 1),
    ("fpn.conv0.1", 2, 24, 8, 8, (512, 640), 3, 1),
    ("fpn.conv1.0", 2, 24, 8, 16, (512, 640), 5, 2),
    ("cr0.conv0", 3, 8, 32, 8, (64, 64, 80), 3, 1),
    ("cr1.conv0", 3, 8, 16, 8, (8, 256, 320), 3, 1),
    ("cr1.conv1", 3, 8, 8, 16, (8, 256, 320), 3, 2),
    ("cr1.heads", 3, 8, 8, 12, (8, 256, 320), 3, 1),
]
import sys
for name, nd, N, cin, cout, sp, k, st in LAYERS[:int(sys.argv[1]) if len(sys.argv) > 1 else None]:
    conv = F.conv2d if nd == 2 else F.conv3d
    fmt = torch.channels_last if nd == 2 else torch.channels_last_3d
    res = []
    for ci, co in itertools.product(sorted({cin, (cin + 3) // 4 * 4, (cin + 7) // 8 * 8, (cin + 15) // 16 * 16}), sorted({cout, (cout + 3) // 4 * 4, (cout + 7) // 8 * 8, (cout + 15) // 16 * 16, (cout + 31) // 32 * 32})):
        for cl in (True, False):
            x = torch.randn(N, ci, *sp, device=dev)
            w = torch.randn(co, ci, *([k] * nd), device=dev) * 0.05
            b = torch.zeros(co, device=dev)
            if cl:
                x = x.contiguous(memory_format=fmt); w = w.contiguous(memory_format=fmt)
            try:
                t = timeit(lambda: conv(x, w, b, st, k // 2))
                tr = timeit(lambda: torch.cudnn_convolution_relu(x, w, b, (st,) * nd, (k // 2,) * nd, (1,) * nd, 1)) if nd == 2 or True else None
            except Exception as ex:
                t, tr = float("nan"), float("nan")
            res.append((min(t, tr), t, tr, ci, co, cl))
            del x, w
    res.sort(key=lambda r: r[0])
    base = [r for r in res if r[3] == cin and r[4] == cout and r[5]][0]
    print(f"{name}: as is (channels-last) {base[1]:.0f} / fused-relu {base[2]:.0f} us; best: " + "; ".join(f"cin {r[3]} cout {r[4]} {'cl' if r[5] else 'nchw'} {r[1]:.0f}/{r[2]:.0f}" for r in res[:4]), flush=True)
