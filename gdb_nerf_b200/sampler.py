"""Host mirror of the reference's ``BundleSampler`` (networks/gdb_nerf/bundle_sampler.py:8-371):
same method names, argument meaning, return tuples and error behaviour, with
the arithmetic done by the CUDA kernels.

In the reference, ``sample`` materialises the packed sample list and ``encode``
gathers per-sample inputs for the MLP.  Here the fused render kernel does both
(plus the MLP and the compositing) without materialising anything; the packed
tensors are still available - bit-exact for the integer ones - because they are
part of the parity contract:

    sampler.build_rays(tar_exts, tar_ints, (H, W), near, far)
    rays_xyz, uvd, z, ball, idx, per_batch, per_bundle = sampler.sample(depth_range, vol_range, b, max_n, inv, adaptive)
    rgbs_feat_dir, vox_feat = sampler.encode(src_images, img_feat, feat_volume, rays_xyz, uvd, ball, src_exts, src_ints, tar_exts, per_batch)
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import torch

from . import ops

Tensor = torch.Tensor


class BundleSampler:
    def __init__(self, global_num_depth: int, max_mipmap_level: int) -> None:
        self.global_num_depth = global_num_depth
        self.max_mipmap_level = max_mipmap_level
        self.H_orig: Optional[int] = None
        self.W_orig: Optional[int] = None
        self.near: Optional[Tensor] = None
        self.far: Optional[Tensor] = None
        self._tar_exts: Optional[Tensor] = None
        self._tar_ints: Optional[Tensor] = None
        self._last: Optional[Dict] = None

    # -- bundle_sampler.py:30-74.  The rays themselves are never materialised: the kernels
    #    rebuild them from the 3x3 ray matrix held in the camera block.
    def build_rays(self, tar_exts: Tensor, tar_ints: Tensor, im_size: Union[Tuple[int, int], List[int]], near: Tensor, far: Tensor) -> None:
        self.H_orig, self.W_orig = im_size
        self.near, self.far = near, far
        self._tar_exts, self._tar_ints = tar_exts, tar_ints
        self._last = None

    def camera_block(self, src_exts: Tensor, src_ints: Tensor, b_size: int, inv_depth: bool) -> Tensor:
        if self._tar_exts is None:
            raise ValueError("Rays have not been built yet. Please call build_rays() first.")
        near_far = torch.stack((self.near, self.far), dim=1)
        return ops.camera_block(self._tar_exts, self._tar_ints, src_exts, src_ints, near_far, b_size, self.global_num_depth, inv_depth)

    # -- bundle_sampler.py:193-265
    def sample(self, depth_range: Tensor, vol_range: Tensor, b_size: int, max_num_samples: int, inv_depth: bool = False,
               is_adaptive: bool = False):
        if self._tar_exts is None:
            raise ValueError("Rays have not been built yet. Please call build_rays() first.")
        B = depth_range.shape[0]
        dev = depth_range.device
        # a camera block without source views is enough for sampling
        eye_e = torch.eye(4, device=dev).expand(B, 1, 4, 4).contiguous()
        eye_k = torch.eye(3, device=dev).expand(B, 1, 3, 3).contiguous()
        cam = self.camera_block(eye_e, eye_k, b_size, inv_depth)
        sl = ops.sample_bundles(depth_range, vol_range, cam, b_size, max_num_samples, inv_depth, is_adaptive)
        if is_adaptive:   # dtypes as in the reference (float32 adaptive, int32/int64 fixed)
            per_bundle = sl.counts.to(torch.float32)
            per_batch = per_bundle.view(B, -1).sum(1)
        else:
            per_bundle = sl.counts
            per_batch = per_bundle.view(B, -1).sum(1)
        self._last = dict(depth_range=depth_range, vol_range=vol_range, b=b_size, max_n=max_num_samples, inv=inv_depth,
                          adaptive=is_adaptive, samples=sl)
        return sl.rays_xyz, sl.uvd, sl.z_vals, sl.ball_radii, sl.indices, per_batch, per_bundle

    # -- bundle_sampler.py:267-371 (+ the MLP and compositing: one kernel)
    def render(self, src_images: Tensor, img_feat: Tensor, feat_volume: Tensor, src_exts: Tensor, src_ints: Tensor, mlp: Tensor,
               with_intermediates: bool = True) -> Dict[str, Tensor]:
        """Fused render of the bundles of the last ``sample`` call.  ``img_feat``
        is the (B,V,feat_dim,Hb,Wb) FPN level WITHOUT the rgb channels (the
        low-resolution colours are appended by the source-preparation kernel)."""
        if self._last is None:
            raise ValueError("call sample() first")
        L = self._last
        B, V, _, H, W = src_images.shape
        cam = self.camera_block(src_exts, src_ints, L["b"], L["inv"])
        src = ops.prepare_sources(img_feat, src_images, L["b"], self.max_mipmap_level)
        vol_cl = ops.to_channels_last(feat_volume, 8)
        return ops.render_fused(src, vol_cl, L["depth_range"], L["vol_range"], cam, mlp, B, V, H, W, L["b"], L["max_n"],
                                L["inv"], L["adaptive"], taps=L["samples"] if with_intermediates else None)

    def encode(self, src_images: Tensor, img_feat: Tensor, feat_volume: Tensor, rays_xyz: Tensor, uvd: Tensor, ball_radii: Tensor,
               src_exts: Tensor, src_ints: Tensor, tar_exts: Tensor, samples_per_batch: Tensor, mlp: Optional[Tensor] = None):
        """Reference signature.  ``img_feat`` carries feat_dim+3 channels as in the
        reference (network.py:162-164); the rgb channels are recomputed by the
        kernel and must be the bilinear down-sampling of ``src_images``.  The
        sample tensors must be the ones returned by the last ``sample`` call."""
        if self._last is None or rays_xyz is not self._last["samples"].rays_xyz:
            raise ValueError("encode() gathers for the samples of the last sample() call; pass its outputs unchanged")
        feat_dim = img_feat.shape[2] - 3
        if mlp is None:
            lib_n = ops._lib.load().gdb_mlp_param_floats(feat_dim)
            mlp = torch.zeros(lib_n, device=src_images.device)
        out = self.render(src_images, img_feat[:, :, :feat_dim], feat_volume, src_exts, src_ints, mlp)
        return out["rgbs_feat_dir"], out["vox_feat"]
