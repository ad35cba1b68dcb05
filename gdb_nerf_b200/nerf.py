"""Parameter containers of the two small MLPs.  Same attribute / state-dict
names as the reference (networks/gdb_nerf/nerf.py:6-56 and
networks/gdb_nerf/depth_net.py:201-246) so checkpoints load with strict=True;
the arithmetic itself runs inside the fused CUDA kernel, which reads the
parameters from a packed block (``mlp_pack.pack_mlp``).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from .mlp_pack import pack_mlp


def _lin_relu(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(cin, cout), nn.ReLU(inplace=True))


class NeRF(nn.Module):
    """Aggregation / radiance MLP of the fine (bundle) renderer."""

    def __init__(self, hid_dim: int = 64, feat_dim: int = 16, voxel_dim: int = 8, viewdir_agg: bool = True) -> None:
        super().__init__()
        if hid_dim != 64 or voxel_dim != 8 or not viewdir_agg:
            raise ValueError("the fused kernel is built for hid_dim=64, voxel_dim=8, viewdir_agg=True (all shipped configs)")
        self.feat_dim = feat_dim
        self.viewdir_agg = viewdir_agg
        F = feat_dim + 3
        self.view_fc = _lin_relu(4, F)
        self.global_fc = _lin_relu(3 * F, 32)
        self.agg_w_fc = _lin_relu(32, 1)
        self.fc = _lin_relu(32, 16)
        self.lr0 = _lin_relu(voxel_dim + 16, hid_dim)
        self.sigma = nn.Sequential(nn.Linear(hid_dim, 1), nn.Softplus())
        self.weight = nn.Sequential(nn.Linear(hid_dim + voxel_dim + 16 + F + 4, hid_dim), nn.ReLU(inplace=True),
                                    nn.Linear(hid_dim, 1), nn.ReLU(inplace=True))
        self.feat_head = _lin_relu(hid_dim, voxel_dim)
        self._packed: Optional[torch.Tensor] = None
        self._packed_key: Optional[Tuple] = None

    def packed_autograd(self) -> torch.Tensor:
        """Packed block built with differentiable operators: gradients of the block flow back to the parameters."""
        return pack_mlp(dict(self.named_parameters()), self.feat_dim, detach=False)

    def packed(self) -> torch.Tensor:
        """Flat parameter block on the parameters' device, re-packed only when a
        parameter was modified (optimizer step, load_state_dict, .to())."""
        params = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._packed is None or key != self._packed_key:
            sd = {k: v for k, v in self.state_dict().items()}
            self._packed = pack_mlp(sd, self.feat_dim, device=params[0].device)
            self._packed_key = key
        return self._packed

    def forward(self, *args, **kwargs):  # pragma: no cover - the arithmetic lives in the fused kernel
        raise RuntimeError("gdb_nerf_b200.NeRF holds parameters only; it is evaluated inside gdb_render_fused_fwd "
                           "(see Network.forward / BundleSampler.render)")


class CoarseNeRF(nn.Module):
    """Training-only MLP of the 1/8-resolution coarse renderer
    (depth_net.py:201-246).  Declared so that ``depth_net.nerfs.0.*`` exists in
    the state dict exactly as in the reference."""

    def __init__(self, hid_dim: int = 64, voxel_dim: int = 8, feat_dim: int = 16, viewdir_agg: bool = True) -> None:
        super().__init__()
        self.hid_dim = hid_dim
        self.viewdir_agg = viewdir_agg
        F = feat_dim + 3
        if viewdir_agg:
            self.view_fc = _lin_relu(4, F)
        self.global_fc = _lin_relu(3 * F, 32)
        self.agg_w_fc = _lin_relu(32, 1)
        self.fc = _lin_relu(32, 16)
        self.lr0 = _lin_relu(voxel_dim + 16, hid_dim)
        self.sigma = nn.Sequential(nn.Linear(hid_dim, 1), nn.Softplus())
        self.color = nn.Sequential(nn.Linear(hid_dim + voxel_dim + 16 + F + 4, hid_dim), nn.ReLU(inplace=True),
                                   nn.Linear(hid_dim, 1), nn.ReLU(inplace=True))

    def forward(self, vox: torch.Tensor, feat_rgb_dir: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """vox (B,N,8), feat_rgb_dir (B,N,V,F+4) -> sigma (B,N), rgb (B,N,3)  (depth_net.py:248-298).
        PyTorch operators: see coarse.py for the status of this training-only row."""
        V = feat_rgb_dir.shape[-2]
        x = feat_rgb_dir[..., :-4]
        if self.viewdir_agg:
            x = x + self.view_fc(feat_rgb_dir[..., -4:])
        var, mean = torch.var_mean(x, dim=-2, keepdim=True)
        g = self.global_fc(torch.cat((x, var.expand(-1, -1, V, -1), mean.expand(-1, -1, V, -1)), -1))
        a = torch.softmax(self.agg_w_fc(g), dim=-2)
        img = self.fc((g * a).sum(-2))
        vi = torch.cat((vox, img), -1)
        h = self.lr0(vi)
        sigma = self.sigma(h)
        hv = torch.cat((h, vi), -1)[..., None, :].expand(-1, -1, V, -1)
        w = torch.softmax(self.color(torch.cat((hv, feat_rgb_dir), -1)), dim=-2)
        rgb = (feat_rgb_dir[..., -7:-4] * w).sum(-2)
        return sigma.squeeze(-1), rgb
