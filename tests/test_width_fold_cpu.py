"""Width-folded execution of the small-channel convolutions (gdb_nerf_b200/cnn.py): f neighbouring pixels of a row read as
f*C channels of one pixel (the same channels-last memory) with the block-Toeplitz expansion of the kernel along the row
reproduce the original Conv2d / Conv3d / ConvTranspose of networks/gdb_nerf/{feature_net.py:12-64, cost_reg_net.py:8-117}
exactly (float64 algebra, CPU), including the layer chains the forward uses."""
import pytest
import torch
import torch.nn.functional as F

from gdb_nerf_b200.cnn import _wview, fold_width_weight, fold_width_weight_transposed


def _cl(x):
    return x.contiguous(memory_format=torch.channels_last if x.dim() == 4 else torch.channels_last_3d)


def _conv(nd):
    return F.conv2d if nd == 2 else F.conv3d


@pytest.mark.parametrize("shape,co,k,s,f", [
    ((2, 8, 12, 16), 8, 3, 1, 2), ((2, 8, 12, 16), 8, 3, 1, 4), ((2, 8, 12, 16), 8, 3, 1, 8),
    ((2, 8, 12, 16), 16, 5, 2, 4), ((2, 8, 12, 16), 16, 5, 2, 2), ((2, 8, 11, 16), 16, 5, 2, 8),
    ((2, 5, 10, 16), 7, 1, 1, 4), ((2, 8, 12, 16), 8, 7, 1, 2),
    ((1, 4, 6, 6, 8), 3, 3, 1, 2), ((1, 4, 6, 6, 8), 3, 3, 2, 4), ((1, 4, 5, 7, 8), 3, 3, 2, 2),
])
def test_folded_conv_equals_conv(shape, co, k, s, f):
    g = torch.Generator().manual_seed(sum(shape) + co + k + s + f)
    nd = len(shape) - 2
    x = _cl(torch.randn(shape, generator=g, dtype=torch.float64))
    w = torch.randn((co, shape[1]) + (k,) * nd, generator=g, dtype=torch.float64)
    want = _conv(nd)(x, w, None, s, k // 2)
    w2, kw = fold_width_weight(w, f, s)
    got = _conv(nd)(_wview(x, f), w2, None, (s,) * (nd - 1) + (1,), (k // 2,) * (nd - 1) + (kw // 2,))
    assert got.shape[1] == co * f // s
    got = _wview(_cl(got), s / f) if f != s else got
    assert got.shape == want.shape
    assert torch.allclose(got, want, atol=1e-12, rtol=0)


@pytest.mark.parametrize("shape,co,f", [((2, 6, 5, 8), 4, 1), ((2, 6, 5, 8), 4, 2), ((2, 6, 5, 8), 4, 4),
                                        ((1, 4, 3, 5, 8), 3, 1), ((1, 4, 3, 5, 8), 3, 2)])
def test_folded_conv_transpose_equals_conv_transpose(shape, co, f):
    g = torch.Generator().manual_seed(sum(shape) + co + f)
    nd = len(shape) - 2
    x = _cl(torch.randn(shape, generator=g, dtype=torch.float64))
    w = torch.randn((shape[1], co) + (3,) * nd, generator=g, dtype=torch.float64)
    ct = F.conv_transpose2d if nd == 2 else F.conv_transpose3d
    want = ct(x, w, None, 2, 1, 1)
    w2, kw = fold_width_weight_transposed(w, f)
    got = ct(_wview(x, f), w2, None, (2,) * (nd - 1) + (1,), (1,) * nd, (1,) * (nd - 1) + (0,))
    got = _wview(_cl(got), 1 / (2 * f))
    assert got.shape == want.shape
    assert torch.allclose(got, want, atol=1e-12, rtol=0)


def test_view_roundtrip_shares_memory():
    x = _cl(torch.arange(2 * 8 * 4 * 16, dtype=torch.float32).view(2, 8, 4, 16))
    v = _wview(x, 4)
    assert v.shape == (2, 32, 4, 4) and v.data_ptr() == x.data_ptr()
    # channel r*C + c of group g is channel c of pixel 4g + r
    assert torch.equal(v[:, 8 * 3 + 5, :, 2], x[:, 5, :, 4 * 2 + 3])
    back = _wview(v, 0.25)
    assert back.data_ptr() == x.data_ptr() and torch.equal(back, x)


def test_fpn_chain_in_folded_views():
    """conv0.1 (fold 4) -> conv1.0 (5x5 stride 2, fold 4 -> 2) -> conv1.1 (fold 2) -> conv2.0 (5x5 stride 2, fold 2 -> 1) chained
    without leaving the folded views equals the plain layer sequence of feature_net.py:40-50."""
    g = torch.Generator().manual_seed(7)
    x = _cl(torch.randn(2, 8, 12, 16, generator=g, dtype=torch.float64))
    ws = [torch.randn(8, 8, 3, 3, generator=g, dtype=torch.float64), torch.randn(16, 8, 5, 5, generator=g, dtype=torch.float64),
          torch.randn(16, 16, 3, 3, generator=g, dtype=torch.float64), torch.randn(32, 16, 5, 5, generator=g, dtype=torch.float64)]
    strides = [1, 2, 1, 2]
    want = x
    for w, s in zip(ws, strides):
        want = F.relu(F.conv2d(want, w, None, s, w.shape[-1] // 2))
    got, f = _wview(x, 4), 4
    for w, s in zip(ws, strides):
        w2, kw = fold_width_weight(w, f, s)
        got = _cl(F.relu(F.conv2d(got, w2, None, (s, 1), (w.shape[-1] // 2, kw // 2))))
        f //= s
    assert f == 1 and got.shape == want.shape
    assert torch.allclose(got, want, atol=0, rtol=1e-12) or (got - want).abs().max() <= 1e-13 * want.abs().max()


def test_stride_must_divide_fold():
    with pytest.raises(ValueError):
        fold_width_weight(torch.zeros(4, 4, 3, 3), 1, 2)


def test_prob_head_depth_chunks_are_never_empty():
    """ops._prob_head_chunks: chunks of ~16 planes; the C side counts one arrival per chunk, so none may be empty."""
    from gdb_nerf_b200.ops import _prob_head_chunks
    for D in range(1, 300):
        n = _prob_head_chunks(D)
        per = (D + n - 1) // n
        assert n >= 1 and (n - 1) * per < D <= n * per
        if D >= 32:
            assert n >= 2
    assert _prob_head_chunks(64) == 4 and _prob_head_chunks(36) == 2 and _prob_head_chunks(8) == 1
