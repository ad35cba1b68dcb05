"""CPU-side checks of the drop-in boundary: state-dict contract, C-ABI symbols,
parameter packing and the no-fallback rule.  No GPU, no compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import CASE_SPECS, ROOT, load_golden
from gdb_nerf_b200 import _lib, mlp_pack
from gdb_nerf_b200.config import make_cfg
from gdb_nerf_b200.network import Network
from gdb_nerf_b200.synthetic import make_batch, synth_state_dict


def _built():
    from gdb_nerf_b200.build import build
    return build()


@pytest.mark.parametrize("case", list(CASE_SPECS))
def test_state_dict_matches_reference(case):
    """Names and shapes of all 205 tensors equal the reference's (recorded by the golden generator)."""
    g = load_golden(case)
    net = Network(make_cfg(CASE_SPECS[case]["recipe"]))
    mine = {k: ",".join(map(str, v.shape)) for k, v in net.state_dict().items()}
    ref = dict(zip(g.np("state_dict_keys").tolist(), g.np("state_dict_shapes").tolist()))
    assert sorted(mine) == sorted(ref)
    assert mine == ref
    assert len(ref) == (207 if CASE_SPECS[case]["recipe"] == "nerf_eval_4x4" else 205)   # b=4 adds one up-sampling conv
    # strict loading of a reference-shaped checkpoint
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(synth_state_dict(shapes, seed=1), strict=True)
    if CASE_SPECS[case]["recipe"] != "nerf_eval_4x4":
        assert sum(p.numel() for p in net.parameters()) == 962311


def test_constructor_errors():
    with pytest.raises(ValueError, match="power of 2"):
        Network(make_cfg("dtu_eval", [("nerf.bundle_size", 3)]))


def test_header_symbols_are_exported_and_typed():
    lib_path = _built()
    header = open(os.path.join(ROOT, "include", "gdb_nerf_b200.h")).read()
    declared = set(re.findall(r"\b(gdb_[a-z0-9_]+)\s*\(", header))
    declared -= {"gdb_render_taps"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), name
    typed = _lib.load()
    assert typed.gdb_abi_version() == _lib.ABI_VERSION
    assert typed.gdb_last_error_string() is not None


@pytest.mark.parametrize("feat_dim", [16, 32])
def test_mlp_pack_layout(feat_dim):
    lib = _lib.load()
    spec, total = mlp_pack.layout(feat_dim)
    assert lib.gdb_mlp_param_floats(feat_dim) == total
    assert lib.gdb_mlp_param_floats(5) < 0
    from gdb_nerf_b200.nerf import NeRF
    net = NeRF(64, feat_dim, 8, True)
    flat = mlp_pack.pack_mlp(net.state_dict(), feat_dim)
    assert flat.numel() == total
    # every offset is float4-aligned and the transpose really is (in, out)
    for key, off, (K, Np), n in spec:
        assert off % 4 == 0 and Np % 4 == 0
    w0 = net.weight[0].weight
    off = dict((k, o) for k, o, _, _ in spec)["weight.0.weight"]
    assert torch.equal(flat[off: off + 64], w0[:, 0])
    back = mlp_pack.unpack_grad(flat, feat_dim)
    for k, v in net.state_dict().items():
        assert torch.equal(back[k].reshape(v.shape), v), k


def test_forward_refuses_cpu_and_training():
    spec = CASE_SPECS["dtu_b2"]
    net = Network(make_cfg(spec["recipe"])).eval()
    batch = make_batch(spec["B"], spec["V"], spec["H"], spec["W"], spec["near"], spec["far"], spec["focal"])
    with pytest.raises(_lib.GdbError, match="CUDA"):
        net(batch)
    with pytest.raises(_lib.GdbError, match="CUDA"):      # the training path has no CPU fallback either
        net.train()(batch)


def test_ops_refuse_cpu_tensors():
    from gdb_nerf_b200 import ops
    with pytest.raises(_lib.GdbError, match="CUDA"):
        ops.to_channels_last(torch.zeros(1, 8, 4, 4))
    with pytest.raises(_lib.GdbError, match="CUDA"):
        ops.depth_values(torch.ones(1, 2, 1, 1), 8, 4, 4, False)


def test_c_abi_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call, so it can be exercised here."""
    lib = _lib.load()
    assert lib.gdb_warp_variance_fwd(None, None, None, 1, 1, 1, 3, 32, 8, 8, 8, 4, 4, 0, 0, None, None) == -1
    assert b"null" in lib.gdb_last_error_string()
    assert lib.gdb_texture_floats(3, 256, 320, 16, 3) == 3 * 20 * (256 * 320 + 128 * 160 + 64 * 80 + 32 * 40)
