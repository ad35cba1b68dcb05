#!/usr/bin/env python
"""BASELINE.json configs[1..4] at N = 1, 2, 4, 8 GPUs in ONE launch (a gpurun --gpus 8 box is charged 8x for its whole
lifetime, so the matrix shares one process start, one NCCL initialisation and one cuDNN warm-up per workload):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/multi_gpu_matrix.py [--steps 20] [--out gpurun_out/r02_matrix.jsonl]

For every N the ranks >= N idle at a barrier while ranks < N (an NCCL sub-communicator of exactly N ranks) run the
measurement; timing is the max over the N active ranks of CUDA-event time (device-resident inputs) and of wall clock
(end to end, host buffers), as bench.py does.  One JSON line per (section, workload, N) on rank 0.

Sections: `eval` (target views sharded over N GPUs: dtu, llff, nerf), `sweep` (the 200-view NeRF-synthetic render sweep
through pipeline.render_sweep over all GPUs), `train` (DTU pre-training step with the NCCL all-reduce of the flat gradient),
`tile` (ONE target view split into bundle-row tiles over N GPUs, single-view latency).
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--views-per-step", type=int, default=8)
    ap.add_argument("--sections", default="eval,sweep,train,tile")
    ap.add_argument("--workloads", default="dtu,llff,nerf")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_matrix.jsonl"))
    ap.add_argument("--no-train-graph", action="store_true")
    ap.add_argument("--sweep-views-per-call", type=int, default=5)
    args = ap.parse_args()

    from gdb_nerf_b200.config import make_cfg
    from gdb_nerf_b200.network import Network
    from gdb_nerf_b200.synthetic import WORKLOADS, batch_to, make_batch, with_uint8_images, workload_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = True
    if world > 1:
        import datetime
        # a short watchdog: a rank that dies must not leave the others (and an N-GPU box charged N times) waiting ten minutes
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))
    sizes = [n for n in (1, 2, 4, 8) if n <= world]
    groups = {}
    for n in sizes:                       # every rank takes part in every new_group call
        groups[n] = None if world == 1 else (dist.group.WORLD if n == world else dist.new_group(list(range(n))))
    lines = []

    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        open(args.out, "w").close()

    def emit(d):
        if rank == 0:
            print(json.dumps(d), flush=True)
            lines.append(d)
            with open(args.out, "a") as fh:          # line by line: a later section that dies does not take these with it
                fh.write(json.dumps(d) + "\n")

    def world_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def group_max(vals, n):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if n > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=groups[n])
        return [float(x) for x in t]

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    B = args.views_per_step
    sections = args.sections.split(",")

    # ------------------------------------------------------------------ eval: target views sharded over N GPUs
    for wl in [w for w in args.workloads.split(",") if "eval" in sections or ("sweep" in sections and w == "nerf") or ("tile" in sections and w == "dtu")]:
        w = WORKLOADS[wl]
        cfg = make_cfg(w["recipe"])
        torch.manual_seed(0)
        net = Network(cfg).to(dev).eval()
        H, W = w["H"], w["W"]
        if "eval" in sections:
            host = workload_batch(wl, B=B, V=3, seed=rank, view_offset=rank * B, images="noise8")
            host_u8 = with_uint8_images(host)
            pinned = {k: ({kk: vv.pin_memory() for kk, vv in v.items()} if isinstance(v, dict) else v.pin_memory()) for k, v in host_u8.items()}
            dev_batch = batch_to(host, dev)
            bs = cfg.nerf.bundle_size
            out_host = [{"rgb": torch.empty((B, 3, H, W)).pin_memory(), "nerf_depth": torch.empty((B, H, W)).pin_memory(),
                         "mvs_depth": torch.empty((B, H // bs, W // bs)).pin_memory()} for _ in range(2)]
            side = [torch.cuda.Stream(device=dev) for _ in range(2)]

            def step_dev():
                with torch.no_grad():
                    net(dev_batch)

            def step_e2e(i):
                with torch.cuda.stream(side[i % 2]), torch.no_grad():
                    ret, _, _ = net(batch_to(pinned, dev, non_blocking=True))
                    for k, buf in out_host[i % 2].items():
                        buf.copy_(ret[k], non_blocking=True)

            for _ in range(4):            # every rank warms up (cuDNN autotuning, folded weights)
                step_dev()
            for i in range(4):
                step_e2e(i)
            world_barrier()
            for n in sizes:
                ms_dev = ms_e2e = 0.0
                if rank < n:
                    if n > 1:
                        dist.barrier(group=groups[n])
                    torch.cuda.synchronize()
                    evs = []
                    for _ in range(args.steps):
                        flush.zero_()
                        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        s.record(stream); step_dev(); e.record(stream)
                        evs.append((s, e))
                    torch.cuda.synchronize()
                    ms_dev = sum(s.elapsed_time(e) for s, e in evs)
                    if n > 1:
                        dist.barrier(group=groups[n])
                    flush.zero_()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for i in range(args.steps):
                        step_e2e(i)
                    torch.cuda.synchronize()
                    ms_e2e = 1e3 * (time.perf_counter() - t0)
                    ms_dev, ms_e2e = group_max([ms_dev, ms_e2e], n)
                world_barrier()
                rays = n * B * H * W * args.steps
                emit({"section": "eval", "workload": f"{wl} {H}x{W}, {B} target views per GPU per step, 3 source views", "n_gpus": n,
                      "nccl_ranks_active": n, "value": rays / (ms_dev * 1e-3) if ms_dev else None, "unit": "rays/s", "ms_per_step": ms_dev / args.steps,
                      "e2e": {"value": rays / (ms_e2e * 1e-3) if ms_e2e else None, "unit": "rays/s", "ms_per_step": ms_e2e / args.steps},
                      "steps": args.steps, "scaling": "weak", "how": f"ranks < {n} of a {world}-rank job active, the others idle at a barrier"})
            del dev_batch, pinned, out_host

        # -------------------------------------------------------------- sweep: 200 NeRF-synthetic views over all GPUs
        if "sweep" in sections and wl == "nerf":
            from gdb_nerf_b200.pipeline import render_sweep
            from gdb_nerf_b200.sharding import shard_views
            n_views = 200
            mine = shard_views(n_views, rank, world)
            store = {i: with_uint8_images(workload_batch("nerf", B=1, V=3, seed=i, view_offset=i, images="noise8")) for i in mine}
            vpc = args.sweep_views_per_call
            for _ in render_sweep(net, lambda i: store[i], n_views=min(2 * vpc * world, n_views), rank=rank, world=world, device=dev, keys=("rgb", "nerf_depth"),
                                  views_per_call=vpc, copy=False):
                pass
            world_barrier()
            t0 = time.perf_counter()
            got = 0
            for idx, res in render_sweep(net, lambda i: store[i], n_views=n_views, rank=rank, world=world, device=dev, keys=("rgb", "nerf_depth"),
                                         views_per_call=vpc, copy=False):
                got += 1
            torch.cuda.synchronize()
            sec = group_max([time.perf_counter() - t0], world)[0]
            world_barrier()
            emit({"section": "sweep", "workload": f"NeRF-synthetic 800x800, 4x4 bundles, {n_views}-view render sweep (pipeline.render_sweep: {vpc} views per forward, pinned double-buffered "
                  "uploads of 8-bit images, image + depth of every view copied back into pinned memory, consumed in place)", "n_gpus": world, "views_per_call": vpc, "views_per_gpu": len(mine),
                  "value": n_views * H * W / sec, "unit": "rays/s", "seconds": sec, "ms_per_view_per_gpu": 1e3 * sec / max(len(mine), 1)})
            del store

        # -------------------------------------------------------------- tile: ONE view split into bundle-row tiles over N GPUs
        if "tile" in sections and wl == "dtu":
            one = batch_to(workload_batch(wl, B=1, V=3, seed=0, images="noise8"), dev)
            with torch.no_grad():
                base = net(one)[0]["rgb"].clone()
            for n in sizes:
                ms = 0.0
                same = 1.0
                if rank < n:
                    net.set_tile_split(rank, n, groups[n] if n > 1 else None)
                    with torch.no_grad():
                        for _ in range(5):
                            ret = net(one)[0]
                    same = 1.0 if torch.equal(ret["rgb"], base) else 0.0
                    if n > 1:
                        dist.barrier(group=groups[n])
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    with torch.no_grad():
                        for _ in range(30):
                            net(one)
                    torch.cuda.synchronize()
                    ms = 1e3 * (time.perf_counter() - t0) / 30
                    ms = group_max([ms], n)[0]
                    same = -group_max([-same], n)[0]
                    net.set_tile_split(0, 1)
                world_barrier()
                emit({"section": "tile", "workload": f"{wl} {H}x{W}: ONE target view, fused render kernel on bundle-row tiles of {n} GPU(s), FPN / DepthNet / decoder "
                      "replicated, one all-gather of the row tiles", "n_gpus": n, "ms_per_view": ms, "unit": "ms", "eager launches": True,
                      "image_bit_identical_to_one_gpu_on_every_rank": bool(same == 1.0)})
            del one
        del net
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------ train: DTU pre-training step, all-reduce of the flat gradient
    if "train" in sections:
        from gdb_nerf_b200.optim import FlatAdam
        cfg = make_cfg("dtu_pretrain")
        crop, V = 64, 3
        batch = batch_to(make_batch(1, V, crop, crop, 425.0, 905.0, 1446.0 * crop / 512.0, seed=100 + rank, images="smooth", tilt=0.03), dev)

        def loss_fn(out):
            return out[0]["rgb"].square().mean() + sum(b.square().mean() for b in out[2])

        for n in sizes:
            res = [0.0] * 5
            note = ""
            if rank < n:
                grp = groups[n] if n > 1 else None
                torch.manual_seed(0)
                net = Network(cfg).to(dev)
                if n > 1:
                    net = torch.nn.SyncBatchNorm.convert_sync_batchnorm(net, process_group=grp)
                net.train()
                net_g = None
                if not args.no_train_graph:         # a second instance with the same weights (SyncBatchNorm holds a process group: no deepcopy)
                    torch.manual_seed(0)
                    net_g = Network(cfg).to(dev)
                    if n > 1:
                        net_g = torch.nn.SyncBatchNorm.convert_sync_batchnorm(net_g, process_group=grp)
                    net_g.load_state_dict(net.state_dict())
                    net_g.train()
                opt = FlatAdam(net.parameters(), lr=5e-4)

                def step():
                    opt.zero_grad()
                    loss = loss_fn(net(batch))
                    loss.backward()
                    opt.step(group=grp, world=n)
                    return loss

                def timed(fn, k):
                    if n > 1:
                        dist.barrier(group=grp)
                    torch.cuda.synchronize()
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record(stream)
                    for _ in range(k):
                        fn()
                    e.record(stream)
                    torch.cuda.synchronize()
                    return s.elapsed_time(e) / k

                for _ in range(4):
                    step()
                ms_eager = timed(step, args.steps)
                ms_ar = timed(lambda: dist.all_reduce(opt.grad_flat, group=grp), 50) if n > 1 else 0.0
                ms_tail = timed(lambda: opt.step(group=grp, world=n), 50)
                ms_graph, ok = 0.0, 0.0
                if net_g is not None:
                    try:
                        from gdb_nerf_b200.graphed import GraphedTrainStep
                        opt_g = FlatAdam(net_g.parameters(), lr=5e-4)
                        graphed = GraphedTrainStep(net_g, opt_g, batch, loss_fn, opt_g.params, group=grp, world=n)
                        for _ in range(3):
                            graphed(batch)
                        torch.cuda.synchronize()
                        ok = 1.0
                    except Exception as exc:
                        note = f"{type(exc).__name__}: {str(exc)[:160]}"
                        sys.stderr.write(f"[rank {rank}] train-step capture failed at N={n}: {note}\n")
                okmin = -group_max([-ok], n)[0]
                if okmin > 0:
                    ms_graph = timed(lambda: graphed(batch), args.steps)
                res = group_max([ms_eager, ms_ar, ms_tail, ms_graph, 0.0], n)
                res[4] = okmin
                nbytes = opt.allreduce_bytes
                del net, net_g, opt
                torch.cuda.empty_cache()
            world_barrier()
            if rank == 0:
                ms_eager, ms_ar, ms_tail, ms_graph, okmin = res
                head = ms_graph if okmin > 0 else ms_eager
                emit({"section": "train", "workload": "dtu_pretrain training step (fwd + bwd + all-reduce of the flat gradient + fused clip/Adam), 3 source views, "
                      "6 samples/bundle, one 64x64 crop = 1024 bundles per GPU", "n_gpus": n, "nccl_ranks_active": n, "ms_per_step": head,
                      "value": n * crop * crop / (head * 1e-3), "unit": "rays/s",
                      "step_execution": "one CUDA graph (NCCL all-reduce and SyncBN collectives captured)" if okmin > 0 else f"eager launches ({note or 'capture not attempted'})",
                      "eager_ms_per_step": ms_eager, "graph_ms_per_step": ms_graph if okmin > 0 else None, "allreduce_ms": ms_ar if n > 1 else None,
                      "allreduce_plus_clip_adam_ms": ms_tail, "allreduce_bytes": nbytes})

    # no destroy_process_group(): with a captured graph that holds NCCL kernels the teardown of the communicator never returned
    # (N = 2, round 2); every line is already on disk, leave at once
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
