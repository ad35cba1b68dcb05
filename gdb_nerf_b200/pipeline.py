"""Input pipeline and sweep driver around ``Network.forward``.

Mirrors the three pieces of the reference a caller of ``run.py evaluate`` touches outside the model:

* ``to_cuda``        - utils/data_utils.py:579-596 (recursive move of the batch dict, ``meta`` left on the host)
* ``load_network``   - utils/net_utils.py:79-111 (checkpoint file or directory with ``latest.pth`` / ``<epoch>.pth``,
                       weights under the ``'net'`` key, strict loading)
* the evaluation loop of run.py:53-66 as ``render_sweep``: target views sharded round-robin over ranks
  (sharding.shard_views), batches staged through pinned host memory and uploaded on a side stream so the copy of view
  i+1 overlaps the kernels of view i, results returned on the host.

Nothing here computes: it is host plumbing around the hand-written path.
"""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, Iterable, Iterator, List, Mapping, Optional, Sequence, Tuple

import torch

from .sharding import shard_views


def to_cuda(batch: Any, device: torch.device | str = "cuda:0", non_blocking: bool = False) -> Any:
    """utils/data_utils.py:579-596: tuples / lists / dicts are walked, tensors moved, ``batch['meta']`` is kept as is."""
    if isinstance(batch, (tuple, list)):
        return [to_cuda(b, device, non_blocking) for b in batch]
    if isinstance(batch, Mapping):
        return {k: (v if k == "meta" else to_cuda(v, device, non_blocking)) for k, v in batch.items()}
    if torch.is_tensor(batch):
        return batch.to(device, non_blocking=non_blocking)
    return batch


def _checkpoint_path(model_dir: str, epoch: int = -1) -> Optional[str]:
    if not os.path.isdir(model_dir):
        return model_dir if os.path.exists(model_dir) else None
    names = os.listdir(model_dir)
    pths = [int(n.split(".")[0]) for n in names if n.endswith(".pth") and n != "latest.pth" and n.split(".")[0].isdigit()]
    if not pths and "latest.pth" not in names:
        return None
    if epoch == -1:
        pth = "latest" if "latest.pth" in names else str(max(pths))
    else:
        pth = str(epoch)
    return os.path.join(model_dir, f"{pth}.pth")


def load_network(net: torch.nn.Module, model_dir: str, resume: bool = True, epoch: int = -1, strict: bool = True) -> int:
    """utils/net_utils.py:79-111.  Returns the epoch to resume from (0 when nothing was loaded).  The file is read with
    ``weights_only=True`` (tensors and plain containers only): a checkpoint is data, not code."""
    if not resume:
        return 0
    path = _checkpoint_path(model_dir, epoch)
    if path is None:
        return 0
    ckpt = torch.load(path, map_location="cpu", weights_only=True)
    state = ckpt["net"] if isinstance(ckpt, Mapping) and "net" in ckpt else ckpt
    net.load_state_dict(state, strict=strict)
    return int(ckpt["epoch"]) + 1 if isinstance(ckpt, Mapping) and "epoch" in ckpt else 0


_PINNED: Dict[Tuple, List[torch.Tensor]] = {}


def _pinned(tag: Tuple, shape, dtype: torch.dtype) -> torch.Tensor:
    """Page-locked host buffer from a process-wide pool keyed by (owner tag, shape, dtype): page-locking costs milliseconds per
    buffer (and serialises between the ranks of a node), so a sweep re-uses the buffers of the sweeps before it."""
    key = (tag, tuple(shape), dtype)
    buf = _PINNED.get(key)
    if buf is None:
        buf = _PINNED[key] = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
    return buf


def _record_stream(batch: Any, stream: torch.cuda.Stream) -> None:
    if isinstance(batch, Mapping):
        for k, v in batch.items():
            if k != "meta":
                _record_stream(v, stream)
    elif isinstance(batch, (tuple, list)):
        for v in batch:
            _record_stream(v, stream)
    elif torch.is_tensor(batch) and batch.is_cuda:
        batch.record_stream(stream)


class PinnedUploader:
    """Double-buffered staging of batch dicts: tensors are copied into page-locked host buffers (allocated once per shape)
    and uploaded with ``non_blocking=True`` on a side stream; ``upload`` returns the device batch and the event that marks
    its arrival.  Non-tensor entries and ``meta`` pass through."""

    def __init__(self, device: torch.device | str = "cuda:0", slots: int = 2) -> None:
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = slots
        torch.cuda.synchronize(self.device)     # the pooled staging buffers may still feed copies of an abandoned sweep
        self._free: List[Optional[torch.cuda.Event]] = [None] * slots
        self._turn = 0

    def _stage(self, slot: int, key: str, t: torch.Tensor) -> torch.Tensor:
        buf = _pinned(("up", str(self.device), slot, key), t.shape, t.dtype)
        buf.copy_(t)
        return buf

    def _stage_cat(self, slot: int, key: str, ts: Sequence[torch.Tensor]) -> torch.Tensor:
        shape = (sum(int(t.shape[0]) for t in ts),) + tuple(ts[0].shape[1:])
        buf = _pinned(("up", str(self.device), slot, key), shape, ts[0].dtype)
        torch.cat([t.cpu() for t in ts], 0, out=buf)
        return buf

    def _walk(self, slot: int, batch: Any, prefix: str) -> Any:
        if isinstance(batch, _CatList):
            return self._stage_cat(slot, prefix, batch).to(self.device, non_blocking=True)
        if isinstance(batch, Mapping):
            return {k: (v if k == "meta" else self._walk(slot, v, f"{prefix}{k}.")) for k, v in batch.items()}
        if isinstance(batch, (tuple, list)):
            return [self._walk(slot, v, f"{prefix}{i}.") for i, v in enumerate(batch)]
        if torch.is_tensor(batch):
            if batch.is_cuda:
                return batch
            return self._stage(slot, prefix, batch).to(self.device, non_blocking=True)
        return batch

    def upload(self, batch: Mapping) -> Tuple[Dict[str, Any], torch.cuda.Event]:
        slot = self._turn % self.slots
        self._turn += 1
        if self._free[slot] is not None:
            self._free[slot].synchronize()          # the previous upload from this slot's host buffers has completed
        with torch.cuda.stream(self.stream):
            dev = self._walk(slot, batch, "")
            done = torch.cuda.Event()
            done.record(self.stream)
        self._free[slot] = done
        return dev, done


class _CatList(list):
    """Marker for `PinnedUploader`: host tensors to be concatenated along dim 0 straight into the pinned staging buffer."""


def _merge_batches(items: Sequence[Any]) -> Any:
    """Batch dicts of several target views -> one structure whose tensor leaves are `_CatList`s (``meta`` -> list of metas)."""
    first = items[0]
    if isinstance(first, Mapping):
        return {k: ([it[k] for it in items] if k == "meta" else _merge_batches([it[k] for it in items])) for k in first}
    if isinstance(first, (tuple, list)):
        return [_merge_batches([it[j] for it in items]) for j in range(len(first))]
    if torch.is_tensor(first):
        return _CatList(items)
    return first


def render_sweep(net: torch.nn.Module, batches: Sequence[Mapping] | Callable[[int], Mapping], n_views: Optional[int] = None,
                 rank: int = 0, world: int = 1, device: torch.device | str = "cuda:0",
                 keys: Sequence[str] = ("rgb",), views_per_call: int = 1, copy: bool = True) -> Iterator[Tuple[int, Dict[str, torch.Tensor]]]:
    """The evaluation loop of run.py:53-66 for this rank's share of a sweep: yields ``(view_index, {key: host tensor})`` in
    order.  ``batches`` is a sequence of (host) batch dicts of ONE target view each or a function ``index -> batch``; view
    ``i`` goes to rank ``i % world``.  There is no collective: ranks are independent.

    The loop is pipelined so that the GPU never waits for the host: ``views_per_call`` consecutive views of this rank are
    concatenated into one forward (target views are independent, so this is the reference's loop with its iterations
    batched; 1 = the reference's one view per call), the upload of call c+1 and the kernels of call c+1 are enqueued before
    the results of call c are handed out, and results travel through a ring of three pinned buffers per key.  With
    ``copy=False`` the yielded tensors are views of that ring, valid until two further calls have been consumed (an
    evaluator that reduces each view to its metrics or writes it to disk, as evaluators/gdb_nerf.py does, needs no copy)."""
    n = len(batches) if n_views is None else n_views       # type: ignore[arg-type]
    get = batches if callable(batches) else (lambda i: batches[i])      # type: ignore[index]
    mine = shard_views(n, rank, world)
    if not mine:
        return
    vpc = max(1, int(views_per_call))
    calls = [mine[i: i + vpc] for i in range(0, len(mine), vpc)]
    up = PinnedUploader(device)
    compute = torch.cuda.current_stream(torch.device(device))
    ring = 3

    def stage(call):
        items = [get(i) for i in call]
        return up.upload(items[0] if len(items) == 1 else _merge_batches(items))

    def launch(c, staged):
        dev_batch, arrived = staged
        compute.wait_event(arrived)
        _record_stream(dev_batch, compute)      # allocated on the upload stream, consumed on the compute stream
        with torch.no_grad():
            ret, _, _ = net(dev_batch)
        slot = c % ring
        res = {}
        for k in keys:
            buf = _pinned(("down", str(device), slot, k), ret[k].shape, ret[k].dtype)
            buf.copy_(ret[k], non_blocking=True)
            res[k] = buf
        done = torch.cuda.Event()
        done.record(compute)
        return res, done

    staged = stage(calls[0])
    pending = None
    for c, call in enumerate(calls):
        cur = launch(c, staged)
        if c + 1 < len(calls):
            staged = stage(calls[c + 1])
        if pending is not None:
            yield from _hand_out(*pending, copy)
        pending = (call, cur)
    yield from _hand_out(*pending, copy)


def _hand_out(call, cur, copy):
    res, done = cur
    done.synchronize()
    for j, idx in enumerate(call):
        yield idx, {k: (v[j: j + 1].clone() if copy else v[j: j + 1]) for k, v in res.items()}
