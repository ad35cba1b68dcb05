"""Multi-GPU partitioning of the rendering path.

Training (configs/dtu_pretrain.yaml, BASELINE.json configs[4]) is data parallel over
target-view crops; its one real exchange step is the gradient all-reduce the reference
gets from DistributedDataParallel (train/trainers/trainer.py:16-22, with
find_unused_parameters=True): ``allreduce_gradients`` does it as ONE flat bucket
(962 311 fp32 parameters = 3.85 MB: a single NCCL all-reduce over NVLink/NVSwitch,
latency-bound, no per-layer buckets worth overlapping).

Target views are independent units (the reference itself loops over the batch,
bundle_sampler.py:318): a sweep is sharded round-robin over one process per
GPU with no collective in the data path.  The only communication is the timing
reduction of a benchmark (max over ranks) and, optionally, gathering the
rendered images on rank 0.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Indices of the target views rank ``rank`` renders: views[rank::world]."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_views, world))


def shard_rows(n_rows: int, rank: int, world: int, align: int = 8) -> range:
    """Contiguous bundle-row tile of one large view (rows aligned to the mip block)."""
    blocks = (n_rows + align - 1) // align
    lo = (blocks * rank) // world * align
    hi = min((blocks * (rank + 1)) // world * align, n_rows)
    return range(lo, hi)


def gather_row_tiles(tensors: Sequence[torch.Tensor], n_rows: int, world: int, group=None, align: int = 8) -> None:
    """Image-tile split of ONE target view (north star: "partitioned by target views and image tiles"): every rank has
    written the bundle rows ``shard_rows(n_rows, rank, world)`` of each ``(1, n_rows, ...)`` tensor; afterwards every rank
    holds all rows.  Equal tiles travel as one in-place ``all_gather_into_tensor`` per tensor (each rank's tile is already
    in its slot of the output), ragged tiles as one broadcast per rank."""
    if world == 1:
        return
    tiles = [shard_rows(n_rows, r, world, align) for r in range(world)]
    equal = len({len(t) for t in tiles}) == 1 and len(tiles[0]) * world == n_rows
    rank = dist.get_rank(group)
    for t in tensors:
        if t.shape[0] != 1 or t.shape[1] != n_rows or not t.is_contiguous():
            raise ValueError("gather_row_tiles needs contiguous (1, n_rows, ...) tensors (single-view latency mode)")
        if equal:
            mine = t[0, tiles[rank].start: tiles[rank].stop]
            dist.all_gather_into_tensor(t[0], mine, group=group)
        else:
            for r, rows in enumerate(tiles):
                if len(rows):
                    dist.broadcast(t[0, rows.start: rows.stop], src=dist.get_global_rank(group, r) if group is not None else r, group=group)


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """Timing reduction used by bench.py: every rank contributes its elapsed time, all get the max."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_images(local: torch.Tensor, view_ids: Sequence[int], n_views: int) -> Optional[torch.Tensor]:
    """Collect the rendered images of all ranks on rank 0 in view order (views, 3, H, W)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    per_rank = (n_views + world - 1) // world
    pad = torch.zeros((per_rank,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0)
    if rank != 0:
        return None
    out = torch.empty((n_views,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        ids = shard_views(n_views, r, world)
        out[ids] = bufs[r][: len(ids)]
    return out


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world: Optional[int] = None) -> int:
    """Average the gradients of ``params`` over all ranks with one all-reduce of a flat fp32 bucket.
    Parameters that received no gradient on this rank contribute zeros and get the averaged gradient of the others
    (DDP's find_unused_parameters=True semantics, trainer.py:21).  Returns the bucket size in bytes."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    dev = params[0].device
    sizes = [p.numel() for p in params]
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
    off = 0
    for p, n in zip(params, sizes):
        if p.grad is not None:
            flat[off: off + n].copy_(p.grad.reshape(-1))
        off += n
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world or dist.get_world_size())
    off = 0
    for p, n in zip(params, sizes):
        g = flat[off: off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat.numel() * 4
