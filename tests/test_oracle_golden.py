"""Pin the CPU oracle against tensors dumped from the unmodified reference.

Every star-marked function of SURVEY.md section 8a is checked segment by segment on
identical inputs (the reference's own intermediates), in float32 (mirrors the
reference's rounding) and in float64 (the "truth" the CUDA kernels are also
compared with).  Integer outputs must be bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import CASE_SPECS
from gdb_nerf_b200.config import make_cfg
from oracle import gdb_oracle as O


def _maxdiff(a, b):
    return float((a.double() - b.double()).abs().max())


def _cfg(golden):
    return make_cfg(CASE_SPECS[golden.name]["recipe"])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_depth_hypotheses_and_warp_variance(golden, dtype):
    cfg = _cfg(golden)
    src_exts, tar_exts = golden.t("in_src_exts", dtype), golden.t("in_tar_exts", dtype)
    for s in range(2):
        rng = golden.t(f"s{s}_range_in", dtype)
        dv_ref = golden.t(f"s{s}_depth_values")
        dv = O.depth_hypotheses(rng, cfg.mvs.num_depth[s], cfg.mvs.inv_depth[s]).expand_as(dv_ref)
        assert _maxdiff(dv, dv_ref) <= 1e-6 * float(dv_ref.abs().max())
        proj = O.homography_matrices(src_exts, golden.t(f"s{s}_src_ints", dtype), tar_exts, golden.t(f"s{s}_tar_ints", dtype))
        var = O.warp_variance(golden.t(f"s{s}_src_feat", dtype), proj, dv_ref.to(dtype), cfg.mvs.inv_depth[s])
        ref = golden.t(f"s{s}_variance")
        assert var.shape == ref.shape
        # reference fp32 self-noise on this segment is 2-4e-4 for unit-variance features
        assert _maxdiff(var, ref) <= 5e-4 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_depth_interval(golden, dtype):
    cfg = _cfg(golden)
    for s in range(2):
        dv = golden.t(f"s{s}_depth_values", dtype)
        depth, ci = O.depth_interval(dv, golden.t(f"s{s}_prob", dtype), cfg.mvs.ci_scales[s], cfg.mvs.inv_depth[s])
        scale = float(golden.t(f"s{s}_depth").abs().max())
        assert _maxdiff(depth, golden.t(f"s{s}_depth")) <= 2e-5 * scale
        assert _maxdiff(ci, golden.t(f"s{s}_ci")) <= 2e-5 * scale
    # the x4 / x2 bilinear up-sampling between the stages (depth_net.py:195-196)
    up = O.upsample_bilinear(golden.t("s0_ci", dtype), *golden.np("s1_range_in").shape[-2:])
    assert _maxdiff(up, golden.t("s1_range_in")) <= 2e-5 * float(golden.t("s1_range_in").abs().max())


def _sample(golden, prefix, dtype, adaptive=None):
    spec = CASE_SPECS[golden.name]
    cfg = _cfg(golden)
    rays = O.target_rays(golden.t("in_tar_exts", dtype), golden.t("in_tar_ints", dtype), spec["H"], spec["W"])
    bundles = O.assemble_bundles(rays, cfg.nerf.bundle_size)
    nf = golden.t("in_near_far", dtype)
    if adaptive is None:
        adaptive = cfg.nerf.is_adaptive
    return O.sample_bundles(bundles, golden.t(prefix + "depth_range", dtype), golden.t(prefix + "vol_range", dtype),
                            nf[:, 0], nf[:, 1], cfg.nerf.global_num_depth, cfg.nerf.max_num_samples,
                            cfg.mvs.inv_depth[-1], adaptive)


@pytest.mark.parametrize("prefix", ["", "inj_"])
def test_sampling_indices_bit_exact(golden, prefix):
    smp = _sample(golden, prefix, torch.float32, adaptive=True if prefix else None)
    assert smp.indices.dtype == torch.int64
    assert np.array_equal(smp.indices.numpy(), golden.np(prefix + "indices"))
    ref_counts = golden.np(prefix + "samples_per_bundle")
    assert smp.samples_per_bundle.numpy().dtype == ref_counts.dtype      # int32 fixed / float32 adaptive
    assert np.array_equal(smp.samples_per_bundle.numpy(), ref_counts)
    assert np.array_equal(smp.samples_per_batch.numpy(), golden.np(prefix + "samples_per_batch"))
    if prefix:
        assert len(np.unique(ref_counts)) >= 3   # every count 1..max occurs


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("prefix", ["", "inj_"])
def test_sampling_geometry(golden, prefix, dtype):
    if dtype == torch.float64 and prefix:
        pytest.skip("counts at the ceil() boundary are only defined in float32")
    smp = _sample(golden, prefix, dtype, adaptive=True if prefix else None)
    zs = float(golden.t(prefix + "z_vals").abs().max())
    assert _maxdiff(smp.z_vals, golden.t(prefix + "z_vals")) <= 1e-6 * zs
    assert _maxdiff(smp.uvd, golden.t(prefix + "uvd")) <= 2e-5
    assert _maxdiff(smp.rays_xyz, golden.t(prefix + "rays_xyz")) <= 2e-6 * float(golden.t(prefix + "rays_xyz").abs().max())
    ref_ball = golden.t(prefix + "ball_radii")
    assert _maxdiff(smp.ball_radii, ref_ball) <= 2e-5 * float(ref_ball.abs().max())


def _encode(golden, prefix, dtype):
    cfg = _cfg(golden)
    b = cfg.nerf.bundle_size
    tex_nchw = golden.t("tex_nchw", dtype)                      # (B,V,F,Hb,Wb) as handed to encode()
    levels = [O.build_mips(tex_nchw[i].permute(0, 2, 3, 1).contiguous(), cfg.nerf.max_mipmap_level) for i in range(tex_nchw.shape[0])]
    smp = O.Samples(golden.t(prefix + "rays_xyz", dtype), golden.t(prefix + "uvd", dtype), golden.t(prefix + "z_vals", dtype),
                    golden.t(prefix + "ball_radii", dtype), golden.t(prefix + "indices"),
                    golden.t(prefix + "samples_per_batch"), golden.t(prefix + "samples_per_bundle"))
    return O.encode_samples(golden.t("in_rgb", dtype), levels, golden.t("feat_volume", dtype), smp,
                            golden.t("in_src_exts", dtype), golden.t("in_src_ints", dtype), golden.t("in_tar_exts", dtype), b)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("prefix", ["", "inj_"])
def test_encode(golden, prefix, dtype):
    rfd, vox = _encode(golden, prefix, dtype)
    ref = golden.t(prefix + "rgbs_feat_dir")
    assert rfd.shape == ref.shape
    noise = 3e-4 if CASE_SPECS[golden.name]["images"] == "noise" else 5e-5
    assert _maxdiff(vox, golden.t(prefix + "vox_feat")) <= 5e-5
    assert _maxdiff(rfd[..., :-4], ref[..., :-4]) <= noise
    # direction features: the unit difference vector is ill-conditioned when target and source rays are nearly parallel
    assert _maxdiff(rfd[..., -1], ref[..., -1]) <= 1e-5
    assert _maxdiff(rfd[..., -4:-1], ref[..., -4:-1]) <= 2e-3


def test_feature_texture_matches_reference(golden):
    cfg = _cfg(golden)
    spec = CASE_SPECS[golden.name]
    b = cfg.nerf.bundle_size
    lvl = 0
    while cfg.fpn.feat_scales[lvl] < 1.0 / b:
        lvl += 1
    # the FPN level used for gathering is the stage whose source features equal that level
    stage = list(cfg.mvs.vol_levels).index(lvl) if lvl in cfg.mvs.vol_levels else None
    if stage is None:
        pytest.skip("gather level not among cost-volume levels")
    feat = golden.t(f"s{stage}_src_feat")
    tex = O.feature_texture(feat, golden.t("in_rgb"), spec["H"] // b, spec["W"] // b, cfg.nerf.max_mipmap_level)
    ref = golden.t("tex_nchw")
    for i in range(ref.shape[0]):
        assert _maxdiff(tex[i][0].permute(0, 3, 1, 2), ref[i]) <= 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("prefix", ["", "inj_"])
def test_mlp_and_composite(golden, prefix, dtype):
    cfg = _cfg(golden)
    feat_dim = golden.np("tex_nchw").shape[2] - 3
    sigma, feat = O.radiance_mlp(golden.mlp(dtype=dtype), golden.t(prefix + "vox_feat", dtype),
                                 golden.t(prefix + "rgbs_feat_dir", dtype), feat_dim)
    assert _maxdiff(sigma, golden.t(prefix + "sigma")) <= 2e-5
    assert _maxdiff(feat, golden.t(prefix + "feat")) <= 2e-5
    w, bfeat, bdepth, bopac = O.composite(golden.t(prefix + "sigma", dtype), golden.t(prefix + "feat", dtype),
                                          golden.t(prefix + "z_vals", dtype), golden.t(prefix + "indices"),
                                          golden.t(prefix + "samples_per_bundle"), cfg.mvs.inv_depth[-1])
    assert _maxdiff(w, golden.t(prefix + "weights")) <= 1e-5
    assert _maxdiff(bfeat, golden.t(prefix + "bundle_feat")) <= 2e-5
    zs = float(golden.t(prefix + "z_vals").abs().max())
    assert _maxdiff(bdepth, golden.t(prefix + "bundle_depth")) <= 2e-6 * zs
    assert _maxdiff(bopac, golden.t(prefix + "bundle_opacity")) <= 1e-5


def test_render_bundles_end_to_end(golden):
    """The whole fused-kernel scope from the reference's own inputs to its bundle maps."""
    cfg = _cfg(golden)
    spec = CASE_SPECS[golden.name]
    b = cfg.nerf.bundle_size
    tex = golden.t("tex_nchw")
    out = O.render_bundles(
        golden.mlp(), tex.shape[2] - 3, golden.t("in_rgb"), tex[:, :, :-3], golden.t("feat_volume"),
        golden.t("inj_depth_range"), golden.t("inj_vol_range"), golden.t("in_src_exts"), golden.t("in_src_ints"),
        golden.t("in_tar_exts"), golden.t("in_tar_ints"), golden.t("in_near_far"), b, cfg.nerf.max_num_samples,
        cfg.nerf.global_num_depth, cfg.nerf.max_mipmap_level, cfg.mvs.inv_depth[-1], True)
    assert np.array_equal(out["indices"].numpy(), golden.np("inj_indices"))
    B, Hb, Wb = spec["B"], spec["H"] // b, spec["W"] // b
    ref_feat = golden.t("inj_bundle_feat").view(B, Hb, Wb, -1).permute(0, 3, 1, 2)
    noise = 3e-4 if spec["images"] == "noise" else 1e-4
    assert _maxdiff(out["bundle_feat"], ref_feat) <= noise
    zs = float(golden.t("inj_z_vals").abs().max())
    assert _maxdiff(out["bundle_depth"].reshape(-1), golden.t("inj_bundle_depth")) <= 1e-5 * zs


@pytest.mark.parametrize("case", ["dtu_b2", "nerf_b4"])
def test_network_forward_port_matches_reference(case):
    """The whole-forward CPU port that bench.py times as `cpu_baseline` reproduces the reference's outputs."""
    from conftest import load_golden
    from gdb_nerf_b200.network import Network
    from gdb_nerf_b200.synthetic import make_batch, synth_state_dict
    g = load_golden(case)
    spec = CASE_SPECS[case]
    cfg = make_cfg(spec["recipe"])
    net = Network(cfg)
    net.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=1), strict=True)
    net.eval()
    batch = make_batch(spec["B"], spec["V"], spec["H"], spec["W"], spec["near"], spec["far"], spec["focal"], seed=3,
                       images=spec["images"], tilt=spec["tilt"])
    with torch.no_grad():
        ret, mvs = O.network_forward(net, batch, cfg)
    scale = spec["far"] - spec["near"]
    assert _maxdiff(ret["rgb"], g.t("ret_rgb")) <= 1e-5
    assert _maxdiff(ret["opacity"], g.t("ret_opacity")) <= 1e-5
    assert _maxdiff(ret["nerf_depth"], g.t("ret_nerf_depth")) <= 1e-5 * scale
    assert _maxdiff(ret["mvs_depth"], g.t("ret_mvs_depth")) <= 1e-5 * scale
    for i, d in enumerate(mvs):
        assert _maxdiff(d, g.t(f"mvs_depth_{i}")) <= 1e-5 * scale
