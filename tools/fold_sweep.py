#!/usr/bin/env python
"""Development tool: times the small-channel convolutions of the FPN / stage-0 cost regularisation as cuDNN runs them
today against their width-folded form (cnn.fold_width_weight: f neighbouring pixels of a row become f*C channels of one
pixel - the SAME channels-last memory - and the kernel its block-Toeplitz expansion along the row).
Writes gpurun_out/fold_sweep.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from gdb_nerf_b200.cnn import fold_width_weight, fold_width_weight_transposed  # noqa: E402


def timed(fn, n=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n


def main():
    torch.backends.cudnn.benchmark = True
    dev = "cuda"
    cases = [  # name, input (N,C,[D],H,W), Co, k, stride, folds
        ("fpn.conv0.1 8->8 3x3", (24, 8, 512, 640), 8, 3, 1, (2, 4, 8)),
        ("fpn.conv1.0 8->16 5x5 s2", (24, 8, 512, 640), 16, 5, 2, (4, 8)),
        ("fpn.conv1.1 16->16 3x3", (24, 16, 256, 320), 16, 3, 1, (2, 4)),
        ("fpn.conv2.0 16->32 5x5 s2", (24, 16, 256, 320), 32, 5, 2, (2, 4)),
        ("fpn.inner1 16->32 1x1", (24, 16, 256, 320), 32, 1, 1, (2, 4)),
        ("fpn.out1 32->16 3x3", (24, 32, 256, 320), 16, 3, 1, (2, 4)),
        ("cost0.conv0 32->8 3x3x3", (8, 32, 64, 64, 80), 8, 3, 1, (2, 4)),
        ("cost0.conv1 8->16 3x3x3 s2", (8, 8, 64, 64, 80), 16, 3, 2, (2, 4, 8)),
        ("cost0.conv2 16->16 3x3x3", (8, 16, 32, 32, 40), 16, 3, 1, (2, 4)),
        ("cost0.head 8->12 3x3x3", (8, 8, 64, 64, 80), 12, 3, 1, (2, 4)),
        ("dec.in 28->64 3x3", (8, 28, 256, 320), 64, 3, 1, (2,)),
        ("dec.tail 64->12 3x3", (8, 64, 256, 320), 12, 3, 1, (2, 4)),
    ]
    out = []
    if "--transposed-only" in sys.argv:
        cases = [("cost0.conv3 16->32 3x3x3 s2", (8, 16, 32, 32, 40), 32, 3, 2, (2,))]
    for name, shp, co, k, s, folds in cases:
        nd = len(shp) - 2
        fmt = torch.channels_last if nd == 2 else torch.channels_last_3d
        x = torch.randn(shp, device=dev).contiguous(memory_format=fmt)
        w = (torch.randn((co, shp[1]) + (k,) * nd, device=dev) * 0.05).contiguous(memory_format=fmt)
        b = torch.randn(co, device=dev)
        conv = F.conv2d if nd == 2 else F.conv3d
        p = k // 2
        relu_fused = k > 1

        def base():
            if relu_fused:
                return torch.cudnn_convolution_relu(x, w, b, (s,) * nd, (p,) * nd, (1,) * nd, 1)
            return conv(x, w, None, s, p)

        try:
            ref = base()
        except RuntimeError:
            relu_fused = False
            ref = base()
        row = {"case": name, "base_ms": timed(base)}
        in_b = x.numel() * 4 + ref.numel() * 4
        row["roofline_ms"] = in_b / 6524.9e9 * 1e3
        for f in folds:
            if shp[-1] % f or f % s:
                continue
            w2, kw = fold_width_weight(w, f, s)
            w2 = w2.contiguous(memory_format=fmt)
            b2 = b.repeat(f // s)
            perm = (0, 2, 3, 1) if nd == 2 else (0, 2, 3, 4, 1)
            inv = (0, 3, 1, 2) if nd == 2 else (0, 4, 1, 2, 3)
            xf = x.permute(*perm).reshape(*[shp[0]] + list(shp[2:-1]) + [shp[-1] // f, f * shp[1]]).permute(*inv)
            stride = (s,) * (nd - 1) + (1,)
            pad = (p,) * (nd - 1) + (kw // 2,)

            def folded():
                if relu_fused:
                    return torch.cudnn_convolution_relu(xf, w2, b2, stride, pad, (1,) * nd, 1)
                return conv(xf, w2, None, stride, pad)

            y = folded()
            yr = y.permute(*perm).reshape(ref.permute(*perm).shape)
            err = (yr - ref.permute(*perm)).abs().max().item()
            row[f"fold{f}_ms"] = timed(folded)
            row[f"fold{f}_err"] = err
            row[f"fold{f}_kw"] = kw
        out.append(row)
        print(json.dumps(row), flush=True)
    tcases = [  # transposed convolutions (k 3, stride 2, padding 1, output_padding 1): name, input, Co, folds
        ("cost0.conv9 deconv 16->8", (8, 16, 32, 32, 40), 8, (1, 2, 4)),
        ("cost0.conv8 deconv 32->16", (8, 32, 16, 16, 20), 16, (1, 2)),
        ("cost1.conv6 folded deconv 64->64", (8, 64, 128, 160), 64, (1, 2)),
        ("cost1.conv5 folded deconv 64->64", (8, 64, 64, 80), 64, (1, 2)),
    ]
    for name, shp, co, folds in tcases:
        nd = len(shp) - 2
        fmt = torch.channels_last if nd == 2 else torch.channels_last_3d
        x = torch.randn(shp, device=dev).contiguous(memory_format=fmt)
        w = (torch.randn((shp[1], co) + (3,) * nd, device=dev) * 0.05).contiguous(memory_format=fmt)
        ct = F.conv_transpose2d if nd == 2 else F.conv_transpose3d
        ref = ct(x, w, None, 2, 1, 1)
        row = {"case": name, "base_ms": timed(lambda: ct(x, w, None, 2, 1, 1))}
        row["roofline_ms"] = (x.numel() + ref.numel()) * 4 / 6524.9e9 * 1e3
        perm = (0, 2, 3, 1) if nd == 2 else (0, 2, 3, 4, 1)
        inv = (0, 3, 1, 2) if nd == 2 else (0, 4, 1, 2, 3)
        for f in folds:
            w2 = fold_width_weight_transposed(w, f)[0].contiguous(memory_format=fmt)
            xf = x.permute(*perm).reshape(*[shp[0]] + list(shp[2:-1]) + [shp[-1] // f, f * shp[1]]).permute(*inv)
            fn = lambda: ct(xf, w2, None, (2,) * (nd - 1) + (1,), (1,) * nd, (1,) * (nd - 1) + (0,))  # noqa: E731
            y = fn()
            row[f"fold{f}_err"] = (y.permute(*perm).reshape(ref.permute(*perm).shape) - ref.permute(*perm)).abs().max().item()
            row[f"fold{f}_ms"] = timed(fn)
        out.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fold_sweep.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
