"""Drop-in ``Network`` for the reference's plugin loader.

The reference selects its model with ``cfg.network_module`` and instantiates
``Network(cfg)`` from that file (networks/make_network.py:5-9); callers then use
``forward(batch) -> (ret, mvs_depths, blend_rgbs)`` (networks/gdb_nerf/network.py:93-189).
This class keeps that contract - constructor keys, batch-dict layout, return
structure and ``state_dict`` names/shapes (strict checkpoint loading) - and runs
the star-marked path of SURVEY.md section 8 on hand-written sm_100a kernels:

    FPN (cuDNN) -> [K1 warp+variance -> CostReg (cuDNN) -> K2 depth range] x stages
                -> source preparation (texture pyramid, RGBA) -> K3 fused render
                -> decoder (cuDNN) -> output assembly kernel

There is no CPU path: ``forward`` needs CUDA tensors and the built C-ABI library.
"""
from __future__ import annotations

from operator import itemgetter
from types import SimpleNamespace
from typing import Any, Dict, List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import autograd as AG
from . import ops
from .cnn import CostRegNet, CostRegNetSmall, Decoder, FeatureNet, cost_reg_folded, cost_reg_fused, decoder_fused, feature_net_fused
from .nerf import CoarseNeRF, NeRF
from .sampler import BundleSampler


def _as_tuple(x):
    return tuple(x) if isinstance(x, (list, tuple)) else (x,)


class DepthNet(nn.Module):
    """Cascade cost-volume depth estimator (networks/gdb_nerf/depth_net.py:10-198):
    owns the cost-regularisation CNNs; warp/variance and depth regression are kernels."""

    def __init__(self, config: SimpleNamespace) -> None:
        super().__init__()
        base = config.fpn.base_channels
        self.vol_levels = list(config.mvs.vol_levels)
        self.vol_scales = list(config.mvs.vol_scales)
        self.num_stages = len(self.vol_levels)
        self.feat_scales = _as_tuple(itemgetter(*self.vol_levels)(config.fpn.feat_scales))
        self.feat_dims = _as_tuple(itemgetter(*self.vol_levels)(config.fpn.feat_dims))
        self.ci_scales = list(config.mvs.ci_scales)
        self.num_depth = list(config.mvs.num_depth)
        self.inv_depth = list(config.mvs.inv_depth)
        voxel_dim = config.mvs.voxel_dim
        # NB: the reference indexes feat_dims (already gathered by level) with vol_levels again (depth_net.py:32,36)
        self.cost_regs = nn.ModuleList([CostRegNetSmall(self.feat_dims[self.vol_levels[0]], voxel_dim, base)])
        for i in range(1, self.num_stages):
            self.cost_regs.append(CostRegNet(self.feat_dims[self.vol_levels[i]], voxel_dim, base))
        # coarse NeRFs exist in every reference checkpoint (module is in training mode at construction, :40-47)
        self.num_samples = list(config.mvs.num_samples)
        self.fold_depth = True          # eval: stages with <= 8 hypotheses run their U-Net depth-folded on 2-D kernels (cnn.py)
        self.nerfs = nn.ModuleList([
            CoarseNeRF(config.nerf.nerf_hidden_dims, voxel_dim, self.feat_dims[i], config.nerf.viewdir_agg)
            for i in range(self.num_stages - 1)])

    def forward_train(self, src_images, ms_feats, src_exts, src_ints, tar_exts, tar_ints, near_far, coarse: bool = True):
        """Training-mode cascade (depth_net.py:118-198 with self.training): the CNNs run as plain modules (batch-norm
        in batch-statistics mode, autograd), K1 / K2 through their forward+backward kernel pairs, and every stage but
        the last also renders the coarse supervision image (row a14)."""
        B, V, _, H, W = src_images.shape
        mvs_depths, range_list, vol_list, volume_list, blend_rgbs = [], [], [], [], []
        depth_range = near_far[..., None, None].contiguous()
        for s in range(self.num_stages):
            feats = ms_feats[self.vol_levels[s]]
            Hi, Wi = int(H * self.vol_scales[s]), int(W * self.vol_scales[s])
            proj = ops.homography_mats(src_exts, src_ints, tar_exts, tar_ints, self.feat_scales[s], self.vol_scales[s])
            feat_cl = feats.permute(0, 1, 3, 4, 2).contiguous()                       # autograd permutes the gradient back
            variance = AG.WarpVariance.apply(feat_cl, proj, depth_range, self.num_depth[s], Hi, Wi, self.inv_depth[s])
            volume, prob = self.cost_regs[s](variance)
            depth, ci, vol_range = AG.DepthRange.apply(depth_range, prob, self.ci_scales[s], self.inv_depth[s])
            mvs_depths.append(depth.squeeze(1))
            range_list.append(ci)
            vol_list.append(vol_range)
            volume_list.append(volume)
            depth_range = ci
            if s < self.num_stages - 1 and coarse:
                src_ints_s = src_ints.clone()
                src_ints_s[..., :2, :] *= self.feat_scales[s]
                tar_ints_s = tar_ints.clone()
                tar_ints_s[:, :2, :] *= self.vol_scales[s]
                blend_rgbs.append(AG.coarse_render_train(self.nerfs[s], volume, feats, src_images, src_exts, src_ints_s, tar_exts,
                                                         tar_ints_s, near_far, ci, vol_range, self.num_samples[s], self.inv_depth[s]))
            if s < self.num_stages - 1:
                up = self.vol_scales[s + 1] / self.vol_scales[s]
                depth_range = F.interpolate(ci, scale_factor=up, mode="bilinear", align_corners=False)
        return mvs_depths, range_list, vol_list, volume_list, blend_rgbs

    def forward(self, src_images, ms_feats, src_exts, src_ints, tar_exts, tar_ints, near_far, fused_cnn: bool = False):
        B, V, _, H, W = src_images.shape
        mvs_depths: List[torch.Tensor] = []
        range_list: List[torch.Tensor] = []
        vol_list: List[torch.Tensor] = []
        volume_list: List[torch.Tensor] = []
        depth_range = near_far[..., None, None].contiguous()                  # (B,2,1,1)
        for s in range(self.num_stages):
            feats = ms_feats[self.vol_levels[s]]                              # (B,V,C,Hs,Ws)
            Hi, Wi = int(H * self.vol_scales[s]), int(W * self.vol_scales[s])
            proj = ops.homography_mats(src_exts, src_ints, tar_exts, tar_ints, self.feat_scales[s], self.vol_scales[s])
            feat_cl = ops.to_channels_last(feats.flatten(0, 1)).unflatten(0, (B, V))
            # few hypotheses (the refinement stages): depth folded into the channels, the U-Net runs on 2-D convolutions
            folded = fused_cnn and self.fold_depth and self.num_depth[s] <= 8 and self.num_depth[s] % 8 == 0
            variance = ops.warp_variance(feat_cl, proj, depth_range, self.num_depth[s], Hi, Wi, self.inv_depth[s],
                                         out_channels_last=fused_cnn, depth_folded=folded)
            if folded:
                volume, logits = cost_reg_folded(self.cost_regs[s], variance, self.num_depth[s])
                depth, ci, vol_range, _ = ops.depth_range_from_logits(depth_range, logits, self.ci_scales[s], self.inv_depth[s])
            elif fused_cnn:
                # the feature head of every stage but the last only feeds the training-time coarse render
                last = s == self.num_stages - 1
                head = self.cost_regs[s].prob_head
                fuse_head = (not last) and head.weight.shape[1] == 8 and self.num_depth[s] <= 256
                volume, logits = cost_reg_fused(self.cost_regs[s], variance, want_volume=last, defer_prob_head=fuse_head)
                if fuse_head:      # `logits` is the U-Net's last feature volume: 1-channel head + soft-max + regression in one kernel
                    depth, ci, vol_range, _ = ops.prob_head_depth_range(logits, head.weight, depth_range, self.ci_scales[s], self.inv_depth[s])
                else:
                    depth, ci, vol_range, _ = ops.depth_range_from_logits(depth_range, logits, self.ci_scales[s], self.inv_depth[s])
            else:
                volume, prob = self.cost_regs[s](variance)
                depth, ci, vol_range = ops.depth_range_from_prob(depth_range, prob, self.ci_scales[s], self.inv_depth[s])
            mvs_depths.append(depth.squeeze(1))
            range_list.append(ci)
            vol_list.append(vol_range)
            volume_list.append(volume)
            depth_range = ci
            if s < self.num_stages - 1:
                up = self.vol_scales[s + 1] / self.vol_scales[s]
                depth_range = F.interpolate(ci, scale_factor=up, mode="bilinear", align_corners=False)
        return mvs_depths, range_list, vol_list, volume_list, []


class Network(nn.Module):
    def __init__(self, config: SimpleNamespace) -> None:
        super().__init__()
        self.feature_net = FeatureNet(config.fpn.base_channels, config.fpn.feat_dims)
        self.voxel_dim = config.mvs.voxel_dim
        self.depth_net = DepthNet(config)

        self.max_num_samples = config.nerf.max_num_samples
        self.b_size = config.nerf.bundle_size
        if self.b_size <= 0 or (self.b_size & (self.b_size - 1)) != 0:
            raise ValueError('`Bundle size` must be a power of 2.')
        self.inv_depth = config.mvs.inv_depth[-1]
        self.is_adaptive = config.nerf.is_adaptive
        self.sampler = BundleSampler(config.nerf.global_num_depth, config.nerf.max_mipmap_level)

        self.feat_level = 0
        while self.feat_level < len(config.fpn.feat_scales) and config.fpn.feat_scales[self.feat_level] < 1. / self.b_size:
            self.feat_level += 1
        feat_dim = config.fpn.feat_dims[self.feat_level]
        self.nerf_hidden_dims = config.nerf.nerf_hidden_dims
        self.viewdir_agg = config.nerf.viewdir_agg
        self.render_scale = 1.
        self.nerf = NeRF(self.nerf_hidden_dims, feat_dim, self.voxel_dim, self.viewdir_agg)

        self.dec_layers = config.nerf.dec_layers
        self.upsampler = Decoder(feat_dim + 3 + self.voxel_dim, 3, num_feats=64, num_layers=self.dec_layers, upscale_factor=self.b_size)
        self.reweighting = config.nerf.reweighting
        self._fpn_levels = max(max(self.depth_net.vol_levels), self.feat_level) + 1
        # "fused": eval-time execution of the cuDNN networks with batch-norm folded into the convolutions and
        # channels-last tensors end to end (the kernels read/write those layouts directly);
        # "modules": the plain nn.Module graph in the reference's NCHW layout.  Same arithmetic either way.
        self.cnn_mode = "fused"
        self._cl_ready = False
        # MLP arithmetic inside the fused render kernel:
        #   1 (default): tcgen05 tensor cores, single fp16 operands, fp32 accumulation in TMEM - the north star's
        #                reduced-precision class (2e-3; measured: fine rgb ~1e-5, decoder features ~3e-4)
        #   2: tcgen05, every operand split into two fp16 planes (hi + lo = 22 bits), three MMAs per K step - the north
        #      star's fp32 class (1e-4), agrees with 0 to ~1e-5
        #   0: fp32 SIMT (the validation variant of the fp32 class)
        self.mlp_precision = 1
        # image-tile split of a single target view over several GPUs (set_tile_split): FPN / DepthNet replicated, the fused
        # render kernel on this rank's bundle rows, one all-gather of the row tiles, decoder replicated
        self._tile = None
        # derived eval-time parameters (folded batch-norm, width/depth-folded kernels, the packed MLP block) are cached on
        # the modules keyed on the parameters' version counters; in-place `.data` writes do not bump those, so a
        # checkpoint load drops every cache explicitly
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_caches())

    def set_tile_split(self, rank: int = 0, world: int = 1, group=None) -> None:
        """Render bundle rows ``sharding.shard_rows(Hb, rank, world)`` of every view on this rank and all-gather the tiles
        (eval path, one target view per call); ``world = 1`` switches the split off."""
        self._tile = None if world <= 1 else (int(rank), int(world), group)

    def invalidate_caches(self) -> None:
        """Drop every derived-parameter cache (call after writing parameters through `.data` / `copy_` under no_grad)."""
        for m in self.modules():
            for name in [k for k in vars(m) if k.startswith("_gdb_")]:
                object.__delattr__(m, name)
            if isinstance(m, NeRF):
                m._packed, m._packed_key = None, None
        ops._PH_WEIGHTS.clear()

    def _channels_last_params(self) -> None:
        if not self._cl_ready:
            self.feature_net.to(memory_format=torch.channels_last)
            self.upsampler.to(memory_format=torch.channels_last)
            self.depth_net.cost_regs.to(memory_format=torch.channels_last_3d)
            self._cl_ready = True

    def _apply(self, fn, *args, **kwargs):
        self._cl_ready = False      # .to()/.cuda() may re-create parameter storage
        return super()._apply(fn, *args, **kwargs)

    def _forward_train(self, src_images, src_exts, src_ints, tar_exts, tar_ints, near_far):
        """Differentiable forward (train_net.py / trainer.py:44-66 call sites): every star-marked step runs through a
        forward + backward kernel pair (autograd.py); the CNNs are plain modules; the output assembly uses PyTorch's
        pixel-shuffle / bilinear operators (pure data movement with a trivial adjoint)."""
        B, V, _, H, W = src_images.shape
        feats = self.feature_net(src_images.flatten(0, 1), levels=self._fpn_levels)
        ms_feats = [f.unflatten(0, (B, V)) for f in feats]
        # the coarse supervision render only exists in training mode (depth_net.py:181)
        mvs_depths, range_list, vol_list, volume_list, blend_rgbs = self.depth_net.forward_train(
            src_images, ms_feats, src_exts, src_ints, tar_exts, tar_ints, near_far, coarse=self.training)
        depth_range, vol_range, feat_volume, mvs_depth = range_list[-1], vol_list[-1], volume_list[-1], mvs_depths[-1]
        b = self.b_size
        Hb, Wb = H // b, W // b
        if depth_range.shape[2:] != (Hb, Wb):
            depth_range = F.interpolate(depth_range, size=(Hb, Wb), mode='bilinear', align_corners=False)
            vol_range = F.interpolate(vol_range, size=(Hb, Wb), mode='bilinear', align_corners=False)
            mvs_depth = F.interpolate(mvs_depth.unsqueeze(1), size=(Hb, Wb), mode='nearest').squeeze(1)
        if feat_volume.shape[-2:] != (Hb, Wb):
            raise ValueError("feature volume resolution must equal the bundle map (true for every shipped recipe)")
        img_feat = ms_feats[self.feat_level]
        if img_feat.shape[-2:] != (Hb, Wb):
            img_feat = F.interpolate(img_feat.flatten(0, 1), size=(Hb, Wb), mode='bilinear', align_corners=False).unflatten(0, (B, V))
        self.sampler.build_rays(tar_exts, tar_ints, (H, W), near_far[:, 0], near_far[:, 1])
        cam = self.sampler.camera_block(src_exts, src_ints, b, self.inv_depth)
        feat, bdepth, bopac = AG.render_fused_train(img_feat, src_images, feat_volume, depth_range, vol_range, cam,
                                                    self.nerf.packed_autograd(), b, self.max_num_samples,
                                                    self.sampler.max_mipmap_level, self.inv_depth, self.is_adaptive)
        R = 3 * b * b
        rgb_f = F.pixel_shuffle(feat[:, :R], b)
        rgb_c = self.upsampler(feat[:, R:])
        nerf_depth = F.interpolate(bdepth.unsqueeze(1), scale_factor=b, mode='bilinear', align_corners=False).squeeze(1)
        nerf_opacity = F.interpolate(bopac.unsqueeze(1), scale_factor=b, mode='bilinear', align_corners=False).squeeze(1)
        rgb = rgb_c + rgb_f
        if self.reweighting:
            rgb = 0.5 * (rgb + rgb_f)
        ret = {'rgb': rgb, 'nerf_depth': nerf_depth, 'mvs_depth': mvs_depth, 'opacity': nerf_opacity}
        return ret, mvs_depths, blend_rgbs

    def forward(self, batch: Dict[str, Any]) -> Tuple[Dict[str, torch.Tensor], List[torch.Tensor], List[torch.Tensor]]:
        src_views, tar_views = batch['src_views'], batch['tar_views']
        near_far = batch['near_far']
        src_images = src_views['rgb']
        if not src_images.is_cuda:
            raise ops._lib.GdbError("gdb_nerf_b200.Network.forward needs CUDA tensors (no CPU fallback exists)")
        if src_images.dtype == torch.uint8:
            # 8-bit source images (what the loaders read from disk, dtu.py:135) converted on the device: x / 255 in IEEE
            # float32, bit-identical to the loaders' host-side conversion, a quarter of the PCIe bytes
            src_images = ops.u8_to_unit(src_images)
        B, V, _, H, W = src_images.shape
        src_exts = src_views['extrinsics']
        src_ints = src_views['intrinsics'].clone()
        tar_exts = tar_views['extrinsics']
        tar_ints = tar_views['intrinsics'].clone()

        if 'render_scale' in batch:
            rs = batch['render_scale']
            self.render_scale = rs[0].item() if torch.is_tensor(rs) else float(rs[0] if isinstance(rs, (list, tuple)) else rs)
        if self.render_scale != 1.:
            src_images = F.interpolate(src_images.flatten(0, 1), scale_factor=self.render_scale, mode='bilinear',
                                       align_corners=False).unflatten(0, (B, V))
            H, W = src_images.shape[-2:]
            src_ints[..., :2, :] *= self.render_scale
            tar_ints[:, :2, :] *= self.render_scale

        # differentiable route: training mode, or an eval-mode caller that asked for gradients (grad mode on and parameters
        # that require them) - same kernels' forward+backward pairs, fp32 SIMT MLP; eval under torch.no_grad() (run.py:56)
        # takes the fast path below
        if self.training or torch.is_grad_enabled() and any(p.requires_grad for p in self.nerf.parameters()):
            return self._forward_train(src_images, src_exts, src_ints, tar_exts, tar_ints, near_far)

        fused = self.cnn_mode == "fused"
        if fused:
            self._channels_last_params()
            feats = feature_net_fused(self.feature_net, src_images.flatten(0, 1), levels=self._fpn_levels)
        else:
            feats = self.feature_net(src_images.flatten(0, 1), levels=self._fpn_levels)
        ms_feats = [f.unflatten(0, (B, V)) for f in feats]
        mvs_depths, range_list, vol_list, volume_list, blend_rgbs = self.depth_net(
            src_images, ms_feats, src_exts, src_ints, tar_exts, tar_ints, near_far, fused_cnn=fused)
        depth_range, vol_range, feat_volume, mvs_depth = range_list[-1], vol_list[-1], volume_list[-1], mvs_depths[-1]

        b = self.b_size
        Hb, Wb = H // b, W // b
        if depth_range.shape[2:] != (Hb, Wb):
            depth_range = F.interpolate(depth_range, size=(Hb, Wb), mode='bilinear', align_corners=False)
            vol_range = F.interpolate(vol_range, size=(Hb, Wb), mode='bilinear', align_corners=False)
            mvs_depth = F.interpolate(mvs_depth.unsqueeze(1), size=(Hb, Wb), mode='nearest').squeeze(1)
        vol_hw = feat_volume.shape[2:4] if fused else feat_volume.shape[-2:]
        if tuple(vol_hw) != (Hb, Wb):
            raise ValueError("feature volume resolution must equal the bundle map (true for every shipped recipe)")

        img_feat = ms_feats[self.feat_level]
        if img_feat.shape[-2:] != (Hb, Wb):
            img_feat = F.interpolate(img_feat.flatten(0, 1), size=(Hb, Wb), mode='bilinear', align_corners=False).unflatten(0, (B, V))

        self.sampler.build_rays(tar_exts, tar_ints, (H, W), near_far[:, 0], near_far[:, 1])
        cam = self.sampler.camera_block(src_exts, src_ints, b, self.inv_depth)
        sources = ops.prepare_sources(img_feat, src_images, b, self.sampler.max_mipmap_level)
        vol_cl = feat_volume if fused else ops.to_channels_last(feat_volume, 8)    # fused: already a (B,D,Hb,Wb,8) view
        dec_c = self.nerf.feat_dim + 3 + self.voxel_dim                              # channels of the decoder's input
        rows = None
        if self._tile is not None:
            from .sharding import gather_row_tiles, shard_rows
            if B != 1 or not fused:
                raise ValueError("the image-tile split renders one target view per call on the fused eval path")
            tile = shard_rows(Hb, self._tile[0], self._tile[1])
            rows = (tile.start, tile.stop)
        out = ops.render_fused(sources, vol_cl, depth_range, vol_range, cam, self.nerf.packed(), B, V, H, W, b,
                               self.max_num_samples, self.inv_depth, self.is_adaptive, out_channels_last=fused,
                               precision=self.mlp_precision, pad_dec=fused, dec_one=fused and dec_c % 4 != 0, rows=rows)
        if rows is not None:
            gather_row_tiles([out['fine'], out['dec_in'], out['depth'], out['opacity']], Hb, self._tile[1], self._tile[2])
        if fused:
            # NCHW shape over channels-last memory; the first pad channel (constant 1) carries the first convolution's bias
            dec12, dec_b = decoder_fused(self.upsampler, out['dec_in'].permute(0, 3, 1, 2), one_channel=dec_c if dec_c % 4 != 0 else -1)
            rgb, nerf_depth, nerf_opacity = ops.assemble_output(out['fine'], dec12, out['depth'], out['opacity'], b,
                                                                self.reweighting, feat_channels_last=True, dec_pre_shuffle=True, dec_bias=dec_b)
        else:
            rgb_c = self.upsampler(out['feat'][:, 3 * b * b:])
            rgb, nerf_depth, nerf_opacity = ops.assemble_output(out['feat'], rgb_c, out['depth'], out['opacity'], b, self.reweighting)
        ret = {'rgb': rgb, 'nerf_depth': nerf_depth, 'mvs_depth': mvs_depth, 'opacity': nerf_opacity}
        return ret, mvs_depths, blend_rgbs
