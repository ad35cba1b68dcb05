"""Packing of the aggregation / radiance MLP (reference: networks/gdb_nerf/nerf.py:6-56)
into the flat parameter block the render kernel stages in shared memory.

Layout (floats; mirrors ``MlpLayout`` in csrc/gdb_common.cuh).  Every linear
layer is transposed to [K][N] (N innermost) so the N weights of one input are
contiguous.  F = feat_dim + 3, FP = roundup4(F).

    view_fc.0    W [4][FP]        b [FP]
    global_fc.0  W [3F][32]       b [32]      rows: x | var | mean
    agg_w_fc.0   w [32]           b [4]       (1 used)
    fc.0         W [32][16]       b [16]
    lr0.0        W [24][64]       b [64]      rows: vox(8) | img(16)
    sigma.0      w [64]           b [4]       (1 used)
    weight.0     W [88+F+4][64]   b [64]      rows: h(64) | vox(8) | img(16) | featrgb(F) | dir(4)
    weight.2     w [64]           b [4]       (1 used)
    feat_head.0  W [64][8]        b [8]
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Tuple

import torch

LAYER_NAMES = ("view_fc.0", "global_fc.0", "agg_w_fc.0", "fc.0", "lr0.0", "sigma.0", "weight.0", "weight.2", "feat_head.0")


def layout(feat_dim: int) -> Tuple[List[Tuple[str, int, Tuple[int, int], int]], int]:
    """-> ([(key, offset, (K, Npad), N)], total floats)."""
    F = feat_dim + 3
    FP = (F + 3) & ~3
    spec = [
        ("view_fc.0.weight", (4, FP), F), ("view_fc.0.bias", (1, FP), F),
        ("global_fc.0.weight", (3 * F, 32), 32), ("global_fc.0.bias", (1, 32), 32),
        ("agg_w_fc.0.weight", (1, 32), 32), ("agg_w_fc.0.bias", (1, 4), 1),
        ("fc.0.weight", (32, 16), 16), ("fc.0.bias", (1, 16), 16),
        ("lr0.0.weight", (24, 64), 64), ("lr0.0.bias", (1, 64), 64),
        ("sigma.0.weight", (1, 64), 64), ("sigma.0.bias", (1, 4), 1),
        ("weight.0.weight", (88 + F + 4, 64), 64), ("weight.0.bias", (1, 64), 64),
        ("weight.2.weight", (1, 64), 64), ("weight.2.bias", (1, 4), 1),
        ("feat_head.0.weight", (64, 8), 8), ("feat_head.0.bias", (1, 8), 8),
    ]
    out, off = [], 0
    for key, shape, n in spec:
        out.append((key, off, shape, n))
        off += shape[0] * shape[1]
    return out, off


def pack_mlp(params: Mapping[str, torch.Tensor], feat_dim: int, device=None, detach: bool = True) -> torch.Tensor:
    """params: the ``nerf.*`` state-dict slice (keys without the prefix).  With ``detach=False`` the block is built
    with differentiable operators (training: the gradient of the block reaches the parameters)."""
    spec, total = layout(feat_dim)
    chunks = []
    for key, _off, (K, Np), n in spec:
        t = params[key].detach() if detach else params[key]
        t = t.to(torch.float32)
        if key.endswith(".weight"):
            t = t.reshape(-1, t.shape[-1]) if t.dim() == 2 else t.reshape(1, -1)
            # nn.Linear stores (out, in); vector heads (out == 1) are kept as a single row of K inputs
            if K == 1:
                mat = t.reshape(1, -1)
            else:
                mat = t.t()                      # (in=K, out=n)
        else:
            mat = t.reshape(1, -1)
        if mat.shape[0] != K or mat.shape[1] != n:
            raise ValueError(f"{key}: expected ({K},{n}) after transpose, got {tuple(mat.shape)}")
        if n < Np:
            mat = torch.nn.functional.pad(mat, (0, Np - n))
        chunks.append(mat.reshape(-1))
    flat = torch.cat(chunks)
    assert flat.numel() == total
    return flat.to(device) if device is not None else flat


def unpack_grad(flat_grad: torch.Tensor, feat_dim: int) -> Dict[str, torch.Tensor]:
    """Inverse of ``pack_mlp`` for gradients: flat block -> per-parameter tensors."""
    spec, _ = layout(feat_dim)
    out = {}
    for key, off, (K, Np), n in spec:
        blk = flat_grad[off: off + K * Np].view(K, Np)[:, :n]
        if key.endswith(".weight"):
            out[key] = blk.reshape(1, -1) if K == 1 else blk.t()
        else:
            out[key] = blk.reshape(-1)
    return out
