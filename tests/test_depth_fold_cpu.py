"""Depth-folded execution of the cost-regularisation U-Net (gdb_nerf_b200/cnn.py): the block-Toeplitz 2-D weights
reproduce the 3-D convolutions of networks/gdb_nerf/cost_reg_net.py exactly (float64 algebra, CPU)."""
import pytest
import torch
import torch.nn.functional as F

from gdb_nerf_b200.cnn import fold_depth_weight


def _fold_x(x):      # (B,C,D,H,W) -> (B, D*C, H, W), channel = d*C + c
    B, C, D, H, W = x.shape
    return x.permute(0, 2, 1, 3, 4).reshape(B, D * C, H, W)


def _unfold_y(y, C):  # (B, D*C, H, W) -> (B,C,D,H,W)
    B, DC, H, W = y.shape
    return y.view(B, DC // C, C, H, W).permute(0, 2, 1, 3, 4)


@pytest.mark.parametrize("D", [8, 4, 2, 1])
@pytest.mark.parametrize("stride", [1, 2])
def test_folded_conv3d(D, stride):
    g = torch.Generator().manual_seed(D * 10 + stride)
    x = torch.randn(2, 5, D, 6, 8, generator=g, dtype=torch.float64)
    w = torch.randn(7, 5, 3, 3, 3, generator=g, dtype=torch.float64)
    want = F.conv3d(x, w, None, stride, 1)
    w2, d_out = fold_depth_weight(w, D, stride, False)
    got = _unfold_y(F.conv2d(_fold_x(x), w2, None, stride, 1), 7)
    assert d_out == want.shape[2]
    assert torch.allclose(got, want, atol=1e-12, rtol=0)


@pytest.mark.parametrize("D", [1, 2, 4])
def test_folded_conv_transpose3d(D):
    g = torch.Generator().manual_seed(D)
    x = torch.randn(2, 6, D, 5, 7, generator=g, dtype=torch.float64)
    w = torch.randn(6, 4, 3, 3, 3, generator=g, dtype=torch.float64)
    want = F.conv_transpose3d(x, w, None, 2, 1, 1)
    w2, d_out = fold_depth_weight(w, D, 2, True)
    got = _unfold_y(F.conv_transpose2d(_fold_x(x), w2, None, 2, 1, 1), 4)
    assert d_out == want.shape[2] == 2 * D
    assert torch.allclose(got, want, atol=1e-12, rtol=0)
