"""Stand-in for ``nerfacc.volrend`` - TEST INFRASTRUCTURE ONLY.

nerfacc is an un-vendored, un-pinned third-party CUDA library (absent from the
reference's requirements.txt).  The reference uses two entry points
(networks/gdb_nerf/utils.py:35,110).  Published behaviour restated here so the
unmodified reference can run on the CPU.  PARITY UNPINNED.

  render_weight_from_alpha(alphas, ray_indices=, n_rays=) -> (weights, trans)
      trans[i] = prod_{j<i, same ray} (1 - alphas[j])   (exclusive, no epsilon)
      weights  = alphas * trans
  accumulate_along_rays(weights, values, ray_indices, n_rays)
      zeros(n_rays, C).index_add_(0, ray_indices, weights[:, None] * values)
"""
import torch


def render_weight_from_alpha(alphas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
    if ray_indices is None or packed_info is not None or prefix_trans is not None:
        raise NotImplementedError("stand-in covers the reference's call shape only")
    n = alphas.shape[0]
    counts = torch.bincount(ray_indices, minlength=n_rays)
    first = torch.cumsum(counts, 0) - counts
    rank = torch.arange(n, device=alphas.device) - first[ray_indices]
    longest = int(counts.max()) if n else 0
    if alphas.is_cuda:
        # same sequence of fp32 multiplications per ray as the loop below (a cumulative product over the ray's samples padded
        # to the longest ray), without one host synchronisation per position: used when the reference is timed on a GPU
        padded = torch.ones(n_rays, longest + 1, dtype=alphas.dtype, device=alphas.device)
        padded = padded.index_put((ray_indices, rank + 1), 1.0 - alphas)
        trans = torch.cumprod(padded, dim=1)[ray_indices, rank]
        return alphas * trans, trans
    # sequential per-segment product, one position at a time (segments are short)
    trans = torch.ones_like(alphas)
    running = torch.ones(n_rays, dtype=alphas.dtype, device=alphas.device)
    for k in range(longest):
        sel = rank == k
        rays_k = ray_indices[sel]
        trans[sel] = running[rays_k]
        running[rays_k] = running[rays_k] * (1.0 - alphas[sel])
    return alphas * trans, trans


def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
    src = weights[:, None] * values if values is not None else weights[:, None]
    out = torch.zeros((n_rays, src.shape[-1]), dtype=src.dtype, device=src.device)
    out.index_add_(0, ray_indices, src)
    return out
