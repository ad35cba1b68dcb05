"""CUDA-graph capture of the eval forward.

``Network.forward`` in eval mode enqueues ~95 kernels and never synchronises with the host (no ``.item()``, no mask
compaction, no ``torch.inverse``: the 4x4 / 3x3 inverses of the reference run on the device in fp64 inside the camera
kernels), so for a fixed batch geometry the whole step can be captured once and replayed as ONE graph launch.  That
removes the host-side launch cost (~0.5 ms per step, what bounds the latency of a single small view) without touching
a kernel.

    runner = GraphedForward(net, example_batch)       # captures after warm-up
    ret, mvs_depths, blend_rgbs = runner(batch)       # copies the batch into the static inputs, replays, returns views

The outputs are the graph's static buffers: they are overwritten by the next call (clone what must survive).
"""
from __future__ import annotations

from typing import Any, Dict, List, Mapping, Tuple

import torch


def _flatten(batch: Mapping, prefix: str = "") -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    for k, v in batch.items():
        if isinstance(v, Mapping):
            out.update(_flatten(v, f"{prefix}{k}."))
        elif torch.is_tensor(v):
            out[f"{prefix}{k}"] = v
    return out


def _rebuild(batch: Mapping, flat: Dict[str, torch.Tensor], prefix: str = "") -> Dict[str, Any]:
    out: Dict[str, Any] = {}
    for k, v in batch.items():
        if isinstance(v, Mapping):
            out[k] = _rebuild(v, flat, f"{prefix}{k}.")
        elif torch.is_tensor(v):
            out[k] = flat[f"{prefix}{k}"]
        else:
            out[k] = v
    return out


class GraphedForward:
    """Replays ``net(batch)`` (eval, no grad) as a CUDA graph for batches with the shapes of ``example_batch``."""

    USED = ("src_views.rgb", "src_views.extrinsics", "src_views.intrinsics", "tar_views.extrinsics", "tar_views.intrinsics", "near_far")

    def __init__(self, net: torch.nn.Module, example_batch: Mapping, warmup: int = 3) -> None:
        if net.training:
            raise ValueError("GraphedForward captures the eval forward; call net.eval() first")
        if "render_scale" in example_batch:
            raise ValueError("render_scale is read on the host (reference network.py:125-126): not capturable")
        flat = {k: v for k, v in _flatten(example_batch).items() if k in self.USED}
        missing = [k for k in self.USED if k not in flat]
        if missing:
            raise KeyError(f"batch lacks {missing}")
        if not all(v.is_cuda for v in flat.values()):
            raise ValueError("GraphedForward needs CUDA tensors (no CPU fallback exists)")
        self.net = net
        self._static_in = {k: v.clone() for k, v in flat.items()}
        self._batch = _rebuild({k: example_batch[k] for k in ("src_views", "tar_views", "near_far")}, self._static_in)
        self._batch["src_views"] = {k: v for k, v in self._batch["src_views"].items() if f"src_views.{k}" in self._static_in}
        self._batch["tar_views"] = {k: v for k, v in self._batch["tar_views"].items() if f"tar_views.{k}" in self._static_in}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):              # cuDNN autotuning, cached folded weights, lazy library init
                net(self._batch)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self._out = net(self._batch)

    def __call__(self, batch: Mapping) -> Tuple[Dict[str, torch.Tensor], List[torch.Tensor], List[torch.Tensor]]:
        self.stage(batch)
        return self.replay()

    def stage(self, batch: Mapping) -> None:
        """Copies ``batch`` (host or device tensors) into the graph's static inputs on the CURRENT stream.  With ``replay`` this
        lets a caller run uploads, replays and downloads of consecutive batches on three streams ordered by events (a copy
        pipeline around ONE compute stream) instead of whole steps on several streams."""
        flat = _flatten(batch)
        for k, dst in self._static_in.items():
            src = flat[k]
            if src.shape != dst.shape:
                raise ValueError(f"{k}: shape {tuple(src.shape)} differs from the captured {tuple(dst.shape)}")
            dst.copy_(src, non_blocking=True)

    def replay(self) -> Tuple[Dict[str, torch.Tensor], List[torch.Tensor], List[torch.Tensor]]:
        """Replays the captured forward on the current stream; the returned tensors are the graph's static outputs
        (overwritten by the next replay of this instance)."""
        self.graph.replay()
        return self._out


class GraphedTrainStep:
    """One whole training step - forward, loss, backward, gradient all-reduce, value clipping, optimiser update
    (train/trainers/trainer.py:44-66) - captured once and replayed as a CUDA graph.  At the reference's training size
    (a 64x64 crop: 1024 bundles, ~1 000 small kernels) the step is bound by host launch overhead, not by the GPU; the
    replay removes it.  Requirements: static batch geometry, an optimiser constructed with ``capturable=True``, no host
    synchronisation in the step (the forward / backward kernel pairs of this package have none), and a model that has not
    run a backward pass on the default stream before (autograd creates each parameter's AccumulateGrad node on the stream of
    its first backward; the capture needs them on its own stream - construct the stepper first, it warms up by itself).

    ``group`` / ``world``: the process group (and its size) the gradient all-reduce runs over; the default is the whole job.

        stepper = GraphedTrainStep(net, opt, example_batch, loss_fn, params)
        loss = stepper(batch)          # device scalar, overwritten by the next call
    """

    def __init__(self, net: torch.nn.Module, opt, example_batch: Mapping, loss_fn, params,
                 clip_value: float = 40.0, allreduce=None, warmup: int = 3, group=None, world=None) -> None:
        if not net.training:
            raise ValueError("GraphedTrainStep captures the training step; call net.train() first")
        from .optim import FlatAdam
        flat = {k: v for k, v in _flatten(example_batch).items() if torch.is_tensor(v) and v.is_cuda}
        self._static_in = {k: v.clone() for k, v in flat.items()}
        self._batch = _rebuild(example_batch, {**_flatten(example_batch), **self._static_in})
        self.params = list(params)
        fused = isinstance(opt, FlatAdam)

        def one_step():
            if fused:
                opt.zero_grad()                       # a memset node of the graph: the flat gradient is static memory
            out = net(self._batch)
            loss = loss_fn(out)
            loss.backward()
            if fused:
                opt.step(group=group, world=world)    # all-reduce on the flat gradient + fused average / clip / Adam
            else:
                if allreduce is not None:
                    allreduce(self.params)
                torch.nn.utils.clip_grad_value_(self.params, clip_value)          # trainer.py:64
                opt.step()
            return loss

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                if not fused:
                    opt.zero_grad(set_to_none=True)
                one_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        if not fused:
            opt.zero_grad(set_to_none=True)
        # thread_local: the NCCL watchdog thread of a multi-GPU run polls CUDA events while this thread captures; under the
        # default "global" mode those calls invalidate the capture (round 1: a 2-GPU capture never completed)
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self._loss = one_step()

    def __call__(self, batch: Mapping) -> torch.Tensor:
        flat = _flatten(batch)
        for k, dst in self._static_in.items():
            dst.copy_(flat[k], non_blocking=True)
        self.graph.replay()
        return self._loss
