#!/usr/bin/env python
"""Per-segment CUDA-event timing and a kernel table (torch.profiler) of one Network.forward.
Development tool; writes gpurun_out/profile_forward_<tag>.json."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gdb_nerf_b200 import ops  # noqa: E402
from gdb_nerf_b200.config import make_cfg  # noqa: E402
from gdb_nerf_b200.network import Network  # noqa: E402
from gdb_nerf_b200.synthetic import WORKLOADS, batch_to, workload_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="dtu")
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--tag", default="base")
    ap.add_argument("--benchmark", action="store_true")
    ap.add_argument("--channels-last", action="store_true")
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--no-tf32", action="store_true")
    ap.add_argument("--kernels", action="store_true")
    ap.add_argument("--tc", action="store_true")
    ap.add_argument("--ops", action="store_true", help="ATen ops grouped by input shape (finds the glue copies/adds)")
    args = ap.parse_args()
    torch.backends.cudnn.benchmark = args.benchmark
    if args.no_tf32:
        torch.backends.cudnn.allow_tf32 = False
    dev = "cuda"
    wl = WORKLOADS[args.workload]
    cfg = make_cfg(wl["recipe"])
    torch.manual_seed(0)
    net = Network(cfg).to(dev).eval()
    net.mlp_precision = 0 if args.tc is False and False else net.mlp_precision
    if args.channels_last:
        net.feature_net.to(memory_format=torch.channels_last)
        net.upsampler.to(memory_format=torch.channels_last)
        net.depth_net.cost_regs.to(memory_format=torch.channels_last_3d)
    batch = batch_to(workload_batch(args.workload, B=args.B), dev)
    T = {}

    def wrap(obj, attr, label):
        f = getattr(obj, attr)

        def g(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = f(*a, **k)
            e.record()
            T.setdefault(label, []).append((s, e))
            return r

        setattr(obj, attr, g)

    wrap(net.feature_net, "forward", "fpn")
    wrap(net.depth_net.cost_regs[0], "forward", "cost_reg0")
    wrap(net.depth_net.cost_regs[1], "forward", "cost_reg1")
    wrap(net.upsampler, "forward", "decoder")
    for name in ("to_channels_last", "homography_mats", "warp_variance", "depth_range_from_prob", "camera_block", "prepare_sources",
                 "render_fused", "assemble_output"):
        wrap(ops, name, name)

    def run():
        with torch.no_grad():
            if args.bf16:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return net(batch)
            return net(batch)

    for _ in range(4):
        run()
    torch.cuda.synchronize()
    T.clear()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    s.record()
    for _ in range(n):
        run()
    e.record()
    torch.cuda.synchronize()
    seg = {k: sum(a.elapsed_time(b) for a, b in v) / n for k, v in T.items()}
    total = s.elapsed_time(e) / n
    out = {"tag": args.tag, "B": args.B, "total_ms": total, "ms_per_view": total / args.B, "segments_ms": seg,
           "other_ms": total - sum(seg.values())}
    if args.kernels:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            run()
            torch.cuda.synchronize()
        rows = []
        for ev in prof.key_averages():
            if ev.device_type == torch.autograd.DeviceType.CUDA or getattr(ev, "self_device_time_total", 0) > 0:
                rows.append((ev.self_device_time_total / 1e3, ev.count, ev.key[:110]))
        rows.sort(reverse=True)
        out["kernels"] = [{"ms": r[0], "count": r[1], "name": r[2]} for r in rows[:40]]
        out["n_kernel_launches"] = sum(r[1] for r in rows)
    if args.ops:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
            run()
            torch.cuda.synchronize()
        rows = []
        for ev in prof.key_averages(group_by_input_shape=True):
            t = getattr(ev, "self_device_time_total", 0)
            if t > 0 and ev.device_type != torch.autograd.DeviceType.CUDA:
                rows.append((t / 1e3, ev.count, ev.key, str(ev.input_shapes)[:160]))
        rows.sort(reverse=True)
        out["ops"] = [{"ms": r[0], "count": r[1], "op": r[2], "shapes": r[3]} for r in rows[:60]]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"profile_forward_{args.tag}.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k not in ("kernels", "ops")}))
    for r in out.get("ops", [])[:45]:
        print(f"  {r['ms']:8.3f} ms x{r['count']:<3d} {r['op']:<34s} {r['shapes']}")
    if args.kernels:
        for r in out["kernels"][:25]:
            print(f"  {r['ms']:8.3f} ms x{r['count']:<4d} {r['name']}")


if __name__ == "__main__":
    main()
