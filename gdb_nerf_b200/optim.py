"""Flat-buffer training-step tail.

The reference's trainer ends every iteration with ``clip_grad_value_(network.parameters(), 40)`` and ``optimizer.step()``
(train/trainers/trainer.py:63-65) on an Adam with one parameter group per tensor (train/optimizer.py:13-29), after
DistributedDataParallel has averaged the gradients bucket by bucket (trainer.py:16-22).  For this model that is 205 small
tensors: ~600 kernel launches per step that move 15 MB.  ``FlatAdam`` keeps parameters, gradients and both Adam moments in
ONE contiguous fp32 buffer each:

* every ``p.data`` / ``p.grad`` is a view into the flat buffers (autograd accumulates straight into the flat gradient),
* the data-parallel exchange is ONE ``all_reduce`` issued on the flat gradient itself (no gather / scatter copies),
* averaging, value clipping and the Adam update are ONE kernel (``gdb_adam_clip_step``), the step counter lives on the
  device, so the whole training step - NCCL all-reduce included - can be captured into a CUDA graph.

Semantics: ``torch.optim.Adam`` (betas, eps, weight_decay as given; no amsgrad).  One deliberate difference: a parameter that
receives no gradient in a step is treated as having a zero gradient (its moments decay) instead of being skipped; for
parameters that NEVER receive one (feature_net.inner2 / out2 in this model) both leave the parameter untouched when
weight_decay = 0 (the reference's setting, configs/dtu_pretrain.yaml:59); with weight decay the flat update decays them too.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from . import _lib


class FlatAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, clip_value: float = 40.0) -> None:
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatAdam: no trainable parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise _lib.GdbError("FlatAdam needs CUDA parameters (no CPU fallback exists)")
        if any(p.dtype != torch.float32 or p.device != dev for p in self.params):
            raise ValueError("FlatAdam: all parameters must be float32 on one device")
        self.lr, self.betas, self.eps, self.weight_decay, self.clip_value = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay), float(clip_value)
        sizes = [p.numel() for p in self.params]
        n = (sum(sizes) + 3) & ~3
        self.param_flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad_flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.state = torch.zeros(4, dtype=torch.float32, device=dev)         # [0] = completed steps
        self.numel = sum(sizes)
        off = 0
        with torch.no_grad():
            for p, k in zip(self.params, sizes):
                view = self.param_flat[off: off + k].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad_flat[off: off + k].view(p.shape)
                off += k

    @property
    def allreduce_bytes(self) -> int:
        return self.grad_flat.numel() * 4

    def zero_grad(self, set_to_none: bool = False) -> None:
        """Zeroes the flat gradient; the ``p.grad`` views stay attached (``set_to_none`` is accepted and ignored)."""
        self.grad_flat.zero_()
        for p, g in zip(self.params, self._grad_views()):
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g

    def _grad_views(self):
        off = 0
        for p in self.params:
            k = p.numel()
            yield self.grad_flat[off: off + k].view(p.shape)
            off += k

    def step(self, group: Optional["dist.ProcessGroup"] = None, world: Optional[int] = None) -> None:
        """All-reduce (sum) of the flat gradient over ``group`` when torch.distributed is initialised, then the fused
        average + clip + Adam kernel."""
        scale = 1.0
        if dist.is_available() and dist.is_initialized():
            w = world if world is not None else dist.get_world_size(group)
            if w > 1:
                dist.all_reduce(self.grad_flat, op=dist.ReduceOp.SUM, group=group)
                scale = 1.0 / w
        lib = _lib.load()
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.gdb_adam_clip_step(self.param_flat.data_ptr(), self.grad_flat.data_ptr(), self.exp_avg.data_ptr(),
                                          self.exp_avg_sq.data_ptr(), self.state.data_ptr(), self.param_flat.numel(), self.lr,
                                          self.betas[0], self.betas[1], self.eps, self.weight_decay, self.clip_value, scale, st),
                   "gdb_adam_clip_step")
        _lib.check(lib.gdb_adam_advance(self.state.data_ptr(), st), "gdb_adam_advance")
