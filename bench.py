#!/usr/bin/env python
"""Headline benchmark: rays/s of ``Network.forward`` at DTU 512x640, 3 source views,
batch of 8 target views per step per GPU (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload dtu|llff|nerf] [--views-per-step B]

One JSON line on stdout (rank 0).  Timing rules: W >= 3 warm-up steps, every
timed step bracketed by CUDA events on the launching stream, an L2 flush (256 MB
memset) between timed steps outside the events, max over ranks, SM clocks and
throttle reasons sampled with nvidia-smi during the timed region.

``--impl reference`` times the reference's own ``Network.forward`` (the unmodified tree staged under baseline/_ref,
see baseline/README.md; pure-PyTorch stand-ins for nvdiffrast / nerfacc) on the host cores, one target view per step;
when the tree is not staged it falls back to the CPU port (oracle/gdb_oracle.py:network_forward) and says so.
The GPU arm's line also carries ``reference_cuda``: the same unmodified model on the same GPU in the same process.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "rays/sec and ms/target-view at DTU 512x640, 3 src views"
UNIT = "rays/s"

# algorithmic bytes per target view (SURVEY.md section 8d / DESIGN.md): unique bytes the kernel must move
K3_BYTES_PER_VIEW = {"dtu": 66.2e6, "llff": 124.1e6, "nerf": 65.6e6}
K1_BYTES_PER_VIEW = {"dtu": 107.5e6, "llff": 167.1e6, "nerf": 153.6e6}
K3_MLP_FLOP_PER_VIEW = {"dtu": 14.75e9, "llff": 27.7e9, "nerf": 18.3e9}      # the reference MLP's operation count


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic(kernel, workload, views):
    """dram__bytes_read.sum + dram__bytes_write.sum of the kernel from the committed ncu --set full capture (per launch),
    when one exists for this workload, launch size AND this build of the kernels (keyed on the source digest: a capture of
    an older kernel is never reported); None otherwise."""
    path = os.path.join(ROOT, "profiles", "k3_dram_traffic.json")
    if kernel != "gdb_render_fused_fwd" or not os.path.exists(path):
        return None
    with open(path) as fh:
        d = json.load(fh)
    if d.get("workload") != workload or d.get("views_per_launch") != views or d.get("kernel_source_sha256") != _kernel_digest():
        return None
    return d["traffic_bytes_per_launch"]


def _kernel_digest():
    """sha256 of the fused render kernel's sources (the file the ncu capture of profiles/k3_dram_traffic.json profiled)."""
    import hashlib
    h = hashlib.sha256()
    for f in ("gdb_render_tc2.cu", "gdb_render_tc2.cuh", "gdb_render_common.cuh", "gdb_tcgen05.cuh", "gdb_sampling.cuh", "gdb_common.cuh"):
        with open(os.path.join(ROOT, "gdb_nerf_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _parity_note(workload):
    """End-to-end error of the BENCHED math mode (TF32 cuDNN convolutions, fp16-operand MLP) against the CPU oracle at the
    benchmark's full image size, measured by tests/test_network_gpu.py::test_benched_math_mode_against_oracle and committed
    under profiles/ (the test asserts the same bounds on every GPU run)."""
    path = os.path.join(ROOT, "profiles", "r02_benched_mode_parity.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        return json.load(fh).get(workload)


def _tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        for k in ("bf16_tflops_sustained", "bf16_tflops"):      # the kernel is timed inside a long step: sustained figure
            if k in d:
                return float(d[k]), f"measured (MEASURED_PEAKS.json {k})"
    return 1400.6, "fallback (sustained figure recorded in DESIGN.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _reference_forward(cfg, device):
    """(callable batch -> outputs, kind): the unmodified reference model when its tree is staged, else the CPU port."""
    from oracle import ref_runner
    if ref_runner.reference_dir() is not None:
        net = ref_runner.load_reference_network(cfg, device=device, seed=0)
        return (lambda batch: net(batch)), "reference"
    from gdb_nerf_b200.network import Network
    from oracle import gdb_oracle as O
    torch.manual_seed(0)
    net = Network(cfg).eval()
    return (lambda batch: O.network_forward(net, batch, cfg)), "port"


def run_reference(args, wl, cfg):
    """CPU arm: the reference's own Network.forward on the host cores, one target view per step (run.py:53-66 loop body)."""
    from gdb_nerf_b200.synthetic import workload_batch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fwd, kind = _reference_forward(cfg, "cpu")
    batch = workload_batch(args.workload, B=1, V=3, seed=0, images="noise8")
    H, W = batch["src_views"]["rgb"].shape[-2:]
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            fwd(batch)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = args.steps * H * W / total
    what = ("the unmodified reference Network.forward (baseline/_ref/reference, pure-PyTorch stand-ins for nvdiffrast / nerfacc)"
            if kind == "reference" else "CPU port of Network.forward (oracle/gdb_oracle.py; reference tree not staged)")
    sample = f"{what}: {args.steps} steps x 1 target view ({H}x{W}, 3 source views) after {args.warmup} warm-up, torch threads = {cores}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload} {H}x{W} eval forward, 3 source views, 1 target view per step, CPU",
                   "recipe": wl["recipe"], "views_per_step": 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(args, cfg, wl):
    """Bounded CPU sample for the `cpu_baseline` object of our own arm (rank 0, N=1)."""
    from gdb_nerf_b200.synthetic import workload_batch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fwd, kind = _reference_forward(cfg, "cpu")
    batch = workload_batch(args.workload, B=1, V=3, seed=0, images="noise8")
    H, W = batch["src_views"]["rgb"].shape[-2:]
    times = []
    with torch.no_grad():
        for i in range(3):
            t0 = time.perf_counter()
            fwd(batch)
            times.append(time.perf_counter() - t0)
    best = min(times[1:])
    what = "unmodified reference Network.forward (baseline/_ref, stand-ins for nvdiffrast / nerfacc)" if kind == "reference" \
        else "CPU port of Network.forward (oracle/gdb_oracle.py)"
    return {"value": H * W / best, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{what}, 1 target view {H}x{W}, best of 2 after 1 warm-up ({best:.2f} s/view)"}


def reference_cuda_probe(workload, cfg, dev):
    """The reference's own PyTorch-CUDA forward on THIS GPU in THIS process (the denominator of the north star's >= 50x):
    the unmodified model from baseline/_ref with pure-PyTorch stand-ins for nvdiffrast / nerfacc, timed as run.py:59-73
    does (synchronize, wall clock, network(batch), synchronize; first iteration dropped), 1 and 8 target views per call."""
    from gdb_nerf_b200.synthetic import batch_to, workload_batch
    from oracle import ref_runner
    if ref_runner.reference_dir() is None:
        return {"unavailable": "reference tree not staged under baseline/_ref (baseline/README.md)"}
    try:
        net = ref_runner.load_reference_network(cfg, device=dev, seed=0)
        out = {"what": "unmodified reference Network.forward on this GPU (stand-ins: pure-PyTorch nvdiffrast.texture / nerfacc.volrend), "
                       "run.py:59-73 timing, mean of 5 calls after 3 dropped, PyTorch default math (TF32 conv), cudnn.benchmark as our arm"}
        for B in (1, 8):
            batch = batch_to(workload_batch(workload, B=B, V=3, seed=0, images="noise8"), dev)
            H, W = batch["src_views"]["rgb"].shape[-2:]
            times = []
            with torch.no_grad():
                for _ in range(8):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    net(batch)
                    torch.cuda.synchronize()
                    times.append(time.perf_counter() - t0)
            ms = 1e3 * sum(times[3:]) / len(times[3:])
            out[f"views_per_call_{B}"] = {"ms_per_call": ms, "ms_per_view": ms / B, "rays_per_s": B * H * W / (ms * 1e-3)}
            del batch
        del net
        torch.cuda.empty_cache()
        return out
    except Exception as exc:                         # the probe must never take the benchmark down
        torch.cuda.empty_cache()
        return {"unavailable": f"{type(exc).__name__}: {str(exc)[:200]}"}


def run_ours(args, wl, cfg):
    import torch.distributed as dist

    from gdb_nerf_b200 import ops
    from gdb_nerf_b200.network import Network
    from gdb_nerf_b200.synthetic import batch_to, with_uint8_images, workload_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = True      # let cuDNN pick its fastest algorithm per shape (shapes are static)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.views_per_step
    torch.manual_seed(0)
    net = Network(cfg).to(dev).eval()
    # every rank renders its own target views (round-robin shard of the sweep): no data-path collective
    # source images are 8-bit samples / 255 (what the reference's loaders produce from image files, dtu.py:135)
    host_batch = workload_batch(args.workload, B=B, V=3, seed=rank, view_offset=rank * B, images="noise8")
    H, W = host_batch["src_views"]["rgb"].shape[-2:]

    def pin(x):
        return {k: pin(v) for k, v in x.items()} if isinstance(x, dict) else x.pin_memory()

    # end to end the images cross PCIe as the 8-bit samples they are; Network.forward converts them on the device
    # (gdb_u8_to_unit_f32: the loaders' `astype(float32) / 255.`, bit-identical), so both timings render the same pixels
    host_e2e = with_uint8_images(host_batch)
    pinned = pin(host_e2e)
    dev_batch = batch_to(host_batch, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    # ---- launch counter + per-kernel events (on torch's current stream, which is the one the kernels launch on)
    counts = {"n": 0}
    spans = {"gdb_render_fused_fwd": [], "gdb_warp_variance_fwd": []}
    recording = {"on": False}
    launches_per_call = {"u8_to_unit": 1, "to_channels_last": 1, "homography_mats": 1, "depth_values": 1, "warp_variance": 1, "depth_range_from_prob": 1,
                         "depth_range_from_logits": 1, "bias_act_add": 1, "gate_add": 1, "se_gate_add": 2, "concat_into": 1,
                         "camera_block": 1, "prepare_sources": 1, "render_fused": 1, "assemble_output": 1,
                         "prob_head_depth_range": 1, "concat_channels": 1, "channel_mean": 2, "pixel_shuffle2_bias": 1}

    def wrap(name, key=None):
        fn = getattr(ops, name)

        def inner(*a, **k):
            counts["n"] += launches_per_call[name]
            if key and recording["on"]:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(stream)
                out = fn(*a, **k)
                e.record(stream)
                spans[key].append((s, e))
                return out
            return fn(*a, **k)

        setattr(ops, name, inner)

    for nm in launches_per_call:
        wrap(nm, {"render_fused": "gdb_render_fused_fwd", "warp_variance": "gdb_warp_variance_fwd"}.get(nm))

    def step_device():
        with torch.no_grad():
            ret, _, _ = net(dev_batch)
        return ret

    b_sz = cfg.nerf.bundle_size
    out_shapes = {"rgb": (B, 3, H, W), "nerf_depth": (B, H, W), "mvs_depth": (B, H // b_sz, W // b_sz)}
    NFLY = 3                      # steps in flight: the copies of two steps overlap the kernels of the third
    out_host = [{k: torch.empty(shp, dtype=torch.float32).pin_memory() for k, shp in out_shapes.items()} for _ in range(NFLY)]
    side = [torch.cuda.Stream(device=dev) for _ in range(NFLY)]

    def step_e2e(i):
        """One end-to-end step on stream i%2: pinned H2D of the batch, forward, D2H of what the reference's evaluator
        consumes (image + both depth maps, evaluators/gdb_nerf.py:37-39,98-100) into pinned memory.  NFLY steps are in
        flight, so the copies of one overlap the kernels of the others; every step still pays its own copies."""
        st = side[i % NFLY]
        with torch.cuda.stream(st), torch.no_grad():
            ret, _, _ = net(batch_to(pinned, dev, non_blocking=True))
            for k, buf in out_host[i % NFLY].items():
                buf.copy_(ret[k], non_blocking=True)

    runners = []

    def step_e2e_graph(i):
        """The same step with the forward replayed as ONE CUDA graph (gdb_nerf_b200.graphed.GraphedForward, the package's public
        replay API): the pinned batch is copied into the graph's static inputs (H2D), the graph replays, the results come back
        (D2H).  One graph instance per step in flight.  ~150 eager operator calls per step become one launch, which matters
        when eight ranks share the host's cores."""
        st = side[i % NFLY]
        with torch.cuda.stream(st), torch.no_grad():
            ret, _, _ = runners[i % NFLY](pinned)
            for k, buf in out_host[i % NFLY].items():
                buf.copy_(ret[k], non_blocking=True)

    pipe = {}

    def step_e2e_pipe(i):
        """The same step as a copy pipeline around ONE compute stream: uploads on an upload stream, graph replays back to back on
        a compute stream, downloads on a download stream, ordered by events; NFLY graph instances (= input / output buffer sets)
        rotate.  Forwards of consecutive steps never interleave on the SMs (the three-stream scheme above lets them)."""
        if not pipe:
            pipe.update(cin=torch.cuda.Stream(device=dev), comp=torch.cuda.Stream(device=dev), cout=torch.cuda.Stream(device=dev),
                        staged=[torch.cuda.Event() for _ in range(NFLY)], computed=[torch.cuda.Event() for _ in range(NFLY)],
                        drained=[torch.cuda.Event() for _ in range(NFLY)])
            for evs in (pipe["computed"], pipe["drained"]):
                for ev in evs:
                    ev.record(torch.cuda.current_stream(dev))
        k = i % NFLY
        r = runners[k]
        with torch.no_grad():
            with torch.cuda.stream(pipe["cin"]):
                pipe["cin"].wait_event(pipe["computed"][k])       # the previous replay of this instance has read its inputs
                r.stage(pinned)
                pipe["staged"][k].record(pipe["cin"])
            with torch.cuda.stream(pipe["comp"]):
                pipe["comp"].wait_event(pipe["staged"][k])
                pipe["comp"].wait_event(pipe["drained"][k])       # ... and its outputs have been copied out
                ret, _, _ = r.replay()
                pipe["computed"][k].record(pipe["comp"])
            with torch.cuda.stream(pipe["cout"]):
                pipe["cout"].wait_event(pipe["computed"][k])
                for key, buf in out_host[k].items():
                    buf.copy_(ret[key], non_blocking=True)
                pipe["drained"][k].record(pipe["cout"])

    def time_e2e(step=None):
        step = step or step_e2e
        for i in range(2 * NFLY):
            step(i)
        barrier()
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step(i)
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0)
        barrier()
        return ms

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, record, tag="gdb_timed"):
        torch.cuda.nvtx.range_push(tag)       # lets `ncu --nvtx --nvtx-include "<tag>/"` profile exactly the timed launches
        evs = []
        for _ in range(steps):
            flush.zero_()                                   # L2 flush between timed iterations, outside the events
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            recording["on"] = record
            s.record(stream)
            fn()
            e.record(stream)
            recording["on"] = False
            evs.append((s, e))
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        return sum(s.elapsed_time(e) for s, e in evs)       # ms over exactly `steps` steps

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    counts["n"] = 0
    ms_dev = timed(step_device, args.steps, True)
    launches = counts["n"]
    barrier()

    # the same forward with the two other MLP variants of the fused render kernel (not the headline):
    #   precision 0 = fp32 SIMT, 2 = split-fp16 (hi+lo) operands on tcgen05 - both the fp32 class (1e-4)
    spans_main = {k: list(v) for k, v in spans.items()}
    alt_steps = max(3, args.steps // 2)
    alt = {}
    for prec in (() if args.lean else (0, 2)):
        for v in spans.values():
            v.clear()
        net.mlp_precision = prec
        for _ in range(3):
            step_device()
        barrier()
        ms_alt = timed(step_device, alt_steps, True, tag=f"gdb_timed_p{prec}")
        alt[prec] = (ms_alt, {k: list(v) for k, v in spans.items()})
    spans.clear(); spans.update(spans_main)
    net.mlp_precision = 1
    barrier()

    # end to end through the public API with host buffers (pinned H2D inside, D2H of the outputs inside):
    # the headline arithmetic, then the fp32-class MLP (precision 2)
    ms_e2e = time_e2e()
    # ... and with the forward as a CUDA-graph replay (every rank decides alone whether its capture worked; the slowest path of
    # any rank is what the max over ranks reports)
    ms_e2e_graph, ms_e2e_pipe, graph_note = 0.0, 0.0, "not attempted (--no-e2e-graph)"
    if not args.no_e2e_graph:
        try:
            from gdb_nerf_b200.graphed import GraphedForward
            example = batch_to(pinned, dev)
            for _ in range(NFLY):
                runners.append(GraphedForward(net, example))
            del example
            graph_note = "ok"
        except Exception as exc:
            runners.clear()
            graph_note = f"capture failed: {type(exc).__name__}: {str(exc)[:160]}"
        ok_all = torch.tensor([1.0 if runners else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
        if float(ok_all[0]) > 0:
            ms_e2e_graph = time_e2e(step_e2e_graph)
            ms_e2e_pipe = time_e2e(step_e2e_pipe)
        pipe.clear()
        runners.clear()
        torch.cuda.empty_cache()
    ms_e2e_p2 = 0.0
    if not args.lean:
        net.mlp_precision = 2
        ms_e2e_p2 = time_e2e()
        net.mlp_precision = 1
    # the sampler ran through every timed region above (device-timed headline, the MLP variants, end to end): the GPU was
    # under the same load throughout; a default run is too short for nvidia-smi to report during the headline loop alone
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0 and clocks is not None:
        clocks["window"] = "all timed regions of this run (headline steps, MLP-variant steps, end-to-end steps)"

    # latency of ONE target view per call (the "ms/target-view" half of the metric for an interactive caller): eager launches
    # vs the whole forward replayed as one CUDA graph (gdb_nerf_b200/graphed.py); rank 0 only, not the headline
    latency = None
    if rank == 0 and not args.lean:
        from gdb_nerf_b200.graphed import GraphedForward
        one = batch_to(workload_batch(args.workload, B=1, V=3, seed=0), dev)
        with torch.no_grad():
            for _ in range(5):
                net(one)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        with torch.no_grad():
            for _ in range(20):
                net(one)
        torch.cuda.synchronize()
        eager_ms = (time.perf_counter() - t1) / 20 * 1e3
        runner = GraphedForward(net, one)
        for _ in range(3):
            runner(one)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(20):
            runner(one)
        torch.cuda.synchronize()
        latency = {"what": "Network.forward on ONE target view per call, inputs resident, wall clock per call, mean of 20",
                   "eager_ms": eager_ms, "cuda_graph_ms": (time.perf_counter() - t1) / 20 * 1e3}
        del runner
        # and the benched batch itself back to back (no L2 flush between calls, wall clock): eager launches against one graph
        # replay per step - how much of the step is launch gaps rather than kernels
        try:
            with torch.no_grad():
                for _ in range(3):
                    net(dev_batch)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            with torch.no_grad():
                for _ in range(20):
                    net(dev_batch)
            torch.cuda.synchronize()
            step_eager = (time.perf_counter() - t1) / 20 * 1e3
            runner = GraphedForward(net, dev_batch)
            for _ in range(3):
                runner(dev_batch)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(20):
                runner(dev_batch)
            torch.cuda.synchronize()
            latency["batch_step_back_to_back"] = {"views_per_step": B, "eager_ms": step_eager, "cuda_graph_ms": (time.perf_counter() - t1) / 20 * 1e3}
            del runner
        except Exception as exc:                         # never lose the headline to a side measurement
            latency["batch_step_back_to_back"] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
    barrier()

    # the reference's own CUDA forward on this GPU, same process (rank 0 of a single-GPU run only: it needs ~7 GB)
    ref_cuda = None
    if rank == 0 and world == 1 and not args.lean and not args.no_reference_cuda:
        del dev_batch
        torch.cuda.empty_cache()
        ref_cuda = reference_cuda_probe(args.workload, cfg, dev)
    barrier()

    t = torch.tensor([ms_dev, ms_e2e, alt[0][0] if alt else 0.0, alt[2][0] if alt else 0.0, ms_e2e_p2, ms_e2e_graph, ms_e2e_pipe], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_p0, ms_p2, ms_e2e_p2, ms_e2e_graph, ms_e2e_pipe = (float(x) for x in t)
    ms_e2e_eager = ms_e2e
    use_graph = 0.0 < ms_e2e_graph < ms_e2e   # all paths are public API and all were timed: the fastest one is the end-to-end figure
    if use_graph:
        ms_e2e = ms_e2e_graph
    use_pipe = 0.0 < ms_e2e_pipe < ms_e2e
    if use_pipe:
        ms_e2e = ms_e2e_pipe

    if rank == 0:
        rays_per_step = world * B * H * W
        value = rays_per_step * args.steps / (ms_dev * 1e-3)
        e2e_value = rays_per_step * args.steps / (ms_e2e * 1e-3)
        peak, how = _peaks()

        def roof(key, bytes_per_view, table=None, nsteps=None):
            sp = (table or spans)[key]
            if not sp:
                return None
            ms = sum(s.elapsed_time(e) for s, e in sp) / len(sp)
            launches_per_step = len(sp) / (nsteps or args.steps)
            # the cost-volume kernel launches once per cascade stage; its per-view figure covers both stages
            bytes_per_launch = bytes_per_view * B / launches_per_step
            ach = bytes_per_launch / (ms * 1e-3) / 1e9
            return {"kernel": key, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": _ncu_traffic(key, args.workload, B) if table is None else None,
                    "avg_launch_ms": ms, "launches_per_step": launches_per_step, "algorithmic_bytes_per_launch": bytes_per_launch,
                    "peak_source": how}

        def roof_tensor(key, flop_per_view):
            """The MLP of the fused render kernel against the measured dense bf16/fp16 tensor peak: flops are the
            reference's operation count, the time is the whole kernel (gathers included)."""
            sp = spans[key]
            if not sp:
                return None
            ms = sum(s.elapsed_time(e) for s, e in sp) / len(sp)
            tpeak, thow = _tensor_peak()
            ach = flop_per_view * B / (ms * 1e-3) / 1e12
            return {"kernel": key, "bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                    "flop_per_launch": flop_per_view * B, "avg_launch_ms": ms, "peak_source": thow}

        h2d = sum(v.numel() * v.element_size() for d in (host_e2e["src_views"], host_e2e["tar_views"]) for v in d.values())
        h2d += host_e2e["near_far"].numel() * 4
        d2h = sum(buf.numel() * 4 for buf in out_host[0].values())
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev / args.steps, "ms_per_target_view": ms_dev / args.steps / B,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (TF32 cuDNN convolutions as in PyTorch's default; per-point MLP GEMM operands f16, f32 accumulate)",
            "data": "synthetic",
            "config": {"mlp_arithmetic": "MLP GEMMs on tcgen05 with fp16 operands and fp32 accumulation in TMEM (the operand precision of the TF32 "
                                         "convolutions PyTorch runs next to it); gathers, geometry, compositing and everything else fp32",
                       "workload": f"{args.workload} {H}x{W} eval forward, 3 source views, batch of {B} target views per GPU per step "
                                   f"(BASELINE.json configs[1])", "recipe": wl["recipe"], "views_per_step_per_gpu": B,
                       "parallelism": f"target views sharded over {world} GPU(s), no data-path collective",
                       "l2": "256 MB memset between timed steps (outside the CUDA events)", "cnn_math": "cuDNN, PyTorch default (TF32 conv), cudnn.benchmark",
                       "images": "8-bit white noise / 255 (as the loaders decode image files); resident run: float32 on the device, e2e: the 8-bit samples cross PCIe",
                       "parity_of_this_math_mode": _parity_note(args.workload)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "what": "pinned host batch (8-bit source images, cameras) -> H2D -> Network.forward -> D2H of ret['rgb'], "
                            "ret['nerf_depth'], ret['mvs_depth'] into pinned memory, every step; three steps in flight on three streams (copies overlap kernels); "
                            "wall clock over all steps",
                    "forward": ("copy pipeline around one compute stream: GraphedForward.stage on an upload stream, .replay back to back on a compute "
                                "stream, the downloads on a third, ordered by events (three instances rotate)" if use_pipe else
                                "one CUDA-graph replay per step (gdb_nerf_b200.graphed.GraphedForward, one instance per step in flight)"
                                if use_graph else "eager operator calls (Network.forward)"),
                    "eager": {"value": rays_per_step * args.steps / (ms_e2e_eager * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_eager / args.steps},
                    "cuda_graph": ({"value": rays_per_step * args.steps / (ms_e2e_graph * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_graph / args.steps}
                                   if ms_e2e_graph > 0.0 else graph_note),
                    "cuda_graph_copy_pipeline": ({"value": rays_per_step * args.steps / (ms_e2e_pipe * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_pipe / args.steps}
                                                 if ms_e2e_pipe > 0.0 else graph_note)},
            "gpu_launches": launches,
            "single_view_latency": latency,
            "roofline": roof("gdb_render_fused_fwd", K3_BYTES_PER_VIEW[args.workload]),
            "roofline_warp_variance": dict(roof("gdb_warp_variance_fwd", K1_BYTES_PER_VIEW[args.workload]) or {},
                                           binding_pipe="L1 data pipe (l1tex__data_pipe_lsu_wavefronts 84-85 % of peak over elapsed, 88-91 % over "
                                                        "active cycles: ncu --set full, profiles/r02_ncu_full_k1_variants.json), not HBM"),
            "roofline_mlp_tensor": roof_tensor("gdb_render_fused_fwd", K3_MLP_FLOP_PER_VIEW[args.workload]),
            "reference_cuda": ref_cuda,
            "mlp_variants": None if args.lean else {
                "headline": "gdb_render_fused_fwd precision=1: MLP GEMMs on tcgen05, fp16 operands, fp32 accumulators in TMEM. Measured against "
                            "the oracle at full size (tools/k3_errors.py): fine rgb <= 1.3e-5, depth <= 1.3e-6 of the range, decoder "
                            "features <= 3.4e-4 - inside the north star's 2e-3 class for a reduced-precision MLP, rgb/depth inside its 1e-4 class",
                "fp32_simt": {"what": "precision=0, fp32 SIMT MLP (fp32 class, validation variant)",
                              "value": rays_per_step * alt_steps / (ms_p0 * 1e-3), "unit": UNIT, "ms_per_step": ms_p0 / alt_steps, "steps": alt_steps,
                              "roofline": roof("gdb_render_fused_fwd", K3_BYTES_PER_VIEW[args.workload], alt[0][1], alt_steps)},
                "tcgen05_split_fp16": {"what": "precision=2, operands split into two fp16 planes (hi+lo, 22 bits), three MMAs per K step (fp32 class: "
                                               "1e-4 vs the reference, 2e-5 vs the SIMT kernel)",
                                       "value": rays_per_step * alt_steps / (ms_p2 * 1e-3), "unit": UNIT, "ms_per_step": ms_p2 / alt_steps, "steps": alt_steps,
                                       "e2e": {"value": rays_per_step * args.steps / (ms_e2e_p2 * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_p2 / args.steps,
                                               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                                       "roofline": roof("gdb_render_fused_fwd", K3_BYTES_PER_VIEW[args.workload], alt[2][1], alt_steps)},
            },
        }
        if ref_cuda and "views_per_call_8" in ref_cuda:
            line["reference_cuda"]["speedup_device_timed"] = value / ref_cuda["views_per_call_8"]["rays_per_s"]
            line["reference_cuda"]["speedup_fp32_class"] = (rays_per_step * alt_steps / (ms_p2 * 1e-3)) / ref_cuda["views_per_call_8"]["rays_per_s"]
        if world == 1 and not args.no_cpu_baseline and not args.lean:
            line["cpu_baseline"] = cpu_baseline_sample(args, cfg, wl)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train(args):
    """BASELINE.json configs[4]: DTU pre-training step (fwd + bwd of the fused render path and the rest of the network,
    gradient all-reduce, clip, Adam) on a 64x64 crop = 1024 bundles per GPU.  One JSON line: training steps are not the
    headline metric; rays/s here counts the rays of the crops all ranks trained on.  The step tail runs on ONE flat buffer
    (gdb_nerf_b200/optim.py: a single NCCL all-reduce issued on the flat gradient, then the fused average + clip + Adam
    kernel); the whole step, collective included, is also captured as a CUDA graph."""
    import copy

    import torch.distributed as dist

    from gdb_nerf_b200.config import make_cfg
    from gdb_nerf_b200.network import Network
    from gdb_nerf_b200.optim import FlatAdam
    from gdb_nerf_b200.synthetic import batch_to, make_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --mode train needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = make_cfg("dtu_pretrain")
    torch.manual_seed(0)
    net = Network(cfg).to(dev)
    if world > 1:
        net = torch.nn.SyncBatchNorm.convert_sync_batchnorm(net)            # as trainer.py:16
    net.train()
    # the captured model never runs eagerly on the default stream (its AccumulateGrad nodes must live on the capture stream)
    net_g = None if args.no_graph else copy.deepcopy(net)
    opt = FlatAdam(net.parameters(), lr=5e-4)
    crop, B, V = 64, 1, 3
    batch = batch_to(make_batch(B, V, crop, crop, 425.0, 905.0, 1446.0 * crop / 512.0, seed=100 + rank, images="smooth", tilt=0.03), dev)
    stream = torch.cuda.current_stream()

    def loss_fn(out):
        return out[0]["rgb"].square().mean() + sum(b.square().mean() for b in out[2])

    def step():
        opt.zero_grad()
        loss = loss_fn(net(batch))
        loss.backward()
        opt.step()                                                         # all-reduce + clip(40) + Adam: trainer.py:63-65
        return loss

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        for _ in range(steps):
            out = fn()
        e.record(stream)
        torch.cuda.synchronize()
        return s.elapsed_time(e) / steps, out

    for _ in range(max(args.warmup, 3)):
        loss = step()
    ms_eager, loss = timed(step, args.steps)
    # the exchange step alone: one all-reduce of the flat gradient (what DDP spreads over its buckets)
    ms_ar = None
    if world > 1:
        ms_ar, _ = timed(lambda: dist.all_reduce(opt.grad_flat), 50)
    ms_tail, _ = timed(lambda: opt.step(), 50)                             # all-reduce + fused clip/Adam kernel (+ counter)
    graphed, graph_note = None, "not attempted (--no-graph)"
    if net_g is not None:
        try:
            from gdb_nerf_b200.graphed import GraphedTrainStep
            opt_g = FlatAdam(net_g.parameters(), lr=5e-4)
            graphed = GraphedTrainStep(net_g, opt_g, batch, loss_fn, opt_g.params)
            for _ in range(3):
                graphed(batch)
            torch.cuda.synchronize()
            graph_note = "whole step (forward, loss, backward, all-reduce of the flat gradient, fused clip + Adam) replayed as one CUDA graph"
        except Exception:                                                  # capture is an optimisation, never a requirement
            import traceback
            sys.stderr.write("[bench] CUDA-graph capture of the training step failed, timing eager launches:\n" + traceback.format_exc()[-1500:] + "\n")
            graphed, graph_note = None, "capture failed (see stderr): eager launches"
    ok = torch.tensor([1.0 if graphed is not None else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)                          # replay only if every rank captured
    torch.cuda.nvtx.range_push("gdb_timed")
    if float(ok[0]) > 0:
        ms_head, loss = timed(lambda: graphed(batch), args.steps)
    else:
        ms_head, loss = ms_eager, loss
        if net_g is not None and graphed is not None:
            graph_note = "another rank failed to capture: eager launches"
    torch.cuda.nvtx.range_pop()
    t = torch.tensor([ms_head, ms_eager, ms_ar or 0.0, ms_tail], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms_step = float(t[0])
        print(json.dumps({
            "metric": "training step (fwd+bwd+allreduce+Adam), DTU pretrain, 64x64 crop = 1024 bundles per GPU", "value": world * B * crop * crop / (ms_step * 1e-3),
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "loss": float(loss.detach()),
            "step_execution": graph_note if float(ok[0]) > 0 else f"eager launches ({graph_note})", "eager_ms_per_step": float(t[1]),
            "allreduce_ms": float(t[2]) if world > 1 else None, "allreduce_plus_clip_adam_ms": float(t[3]),
            "config": {"workload": "dtu_pretrain training step, 3 source views, fixed 6 samples/bundle, 1 crop of 64x64 px per GPU "
                                   "(BASELINE.json configs[4])", "allreduce_bytes": opt.allreduce_bytes,
                       "optimizer": "FlatAdam: one flat fp32 buffer, gdb_adam_clip_step (average + clip 40 + Adam) in one kernel",
                       "parallelism": f"data parallel over {world} GPU(s): one NCCL all-reduce issued on the flat gradient per step, SyncBN"},
        }), flush=True)
    if world > 1:
        # destroying a communicator whose kernels sit in a captured graph never returned at N = 2 (round 2): leave at once
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["dtu", "llff", "nerf"], default="dtu")
    ap.add_argument("--views-per-step", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true", help="skip the reference's own CUDA forward (reference_cuda object)")
    ap.add_argument("--lean", action="store_true", help="headline + e2e only (no MLP variants, latency, CPU / reference-CUDA legs): multi-GPU matrix runs")
    ap.add_argument("--no-graph", action="store_true", help="train mode: time eager launches only")
    ap.add_argument("--no-e2e-graph", action="store_true", help="eval mode: end to end with eager operator calls only")
    ap.add_argument("--mode", choices=["eval", "train"], default="eval", help="train: BASELINE.json configs[4] (not the headline)")
    args = ap.parse_args()
    if args.mode == "train":
        if args.steps is None:
            args.steps = 20
        return run_train(args)
    from gdb_nerf_b200.config import make_cfg
    from gdb_nerf_b200.synthetic import WORKLOADS
    wl = WORKLOADS[args.workload]
    cfg = make_cfg(wl["recipe"])
    if args.impl == "reference":
        if args.steps is None:
            args.steps = 3
        run_reference(args, wl, cfg)
    else:
        if args.steps is None:
            args.steps = 30
        run_ours(args, wl, cfg)


if __name__ == "__main__":
    main()
