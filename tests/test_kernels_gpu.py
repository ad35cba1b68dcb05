"""Parity of the CUDA kernels (called through the C ABI) against the CPU oracle
and the golden vectors dumped from the reference.  Integer outputs bit-exact;
floating point within the tolerances of BASELINE.json's north star, written in
each assertion (fp32: 1e-4 absolute on rgb/depth-like quantities in normalised
units; white-noise inputs get the reference's own fp32 noise floor instead)."""
import numpy as np
import pytest
import torch

from conftest import CASE_SPECS, load_golden
from gdb_nerf_b200 import ops
from gdb_nerf_b200.config import make_cfg
from oracle import gdb_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _md(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def _cfg(g):
    return make_cfg(CASE_SPECS[g.name]["recipe"])


def _scales(cfg, s):
    fs = cfg.fpn.feat_scales[cfg.mvs.vol_levels[s]]
    return fs, cfg.mvs.vol_scales[s]


# ------------------------------------------------------------------ K1 / a1 / a2
def test_homography_and_depth_values(golden):
    cfg = _cfg(golden)
    se, si = golden.t("in_src_exts"), golden.t("in_src_ints")
    te, ti = golden.t("in_tar_exts"), golden.t("in_tar_ints")
    for s in range(2):
        fs, vs = _scales(cfg, s)
        proj = ops.homography_mats(se.to(DEV), si.to(DEV), te.to(DEV), ti.to(DEV), fs, vs)
        ref = O.homography_matrices(se.double(), golden.t(f"s{s}_src_ints").double(), te.double(), golden.t(f"s{s}_tar_ints").double())
        assert _md(proj, ref) <= 2e-6 * float(ref.abs().max())
        rng = golden.t(f"s{s}_range_in")
        dv_ref = golden.t(f"s{s}_depth_values")
        Ht, Wt = dv_ref.shape[-2:]
        dv = ops.depth_values(rng.to(DEV), cfg.mvs.num_depth[s], Ht, Wt, cfg.mvs.inv_depth[s])
        want = O.depth_hypotheses(rng, cfg.mvs.num_depth[s], cfg.mvs.inv_depth[s]).expand_as(dv_ref)
        assert _md(dv, want) <= 2.5e-7 * float(dv_ref.abs().max())             # 2 ulp: ATen's vectorised linspace may fuse
        assert _md(dv, dv_ref) <= 1e-6 * float(dv_ref.abs().max())


def test_warp_variance(golden):
    cfg = _cfg(golden)
    se, te = golden.t("in_src_exts"), golden.t("in_tar_exts")
    for s in range(2):
        feat = golden.t(f"s{s}_src_feat")                                      # (B,V,C,Hs,Ws)
        B, V = feat.shape[:2]
        proj = O.homography_matrices(se, golden.t(f"s{s}_src_ints"), te, golden.t(f"s{s}_tar_ints"))
        rng = golden.t(f"s{s}_range_in")
        ref = golden.t(f"s{s}_variance")
        D, Ht, Wt = ref.shape[2:]
        feat_cl = ops.to_channels_last(feat.flatten(0, 1).to(DEV)).unflatten(0, (B, V))
        assert torch.equal(feat_cl.cpu(), feat.permute(0, 1, 3, 4, 2).contiguous())
        var = ops.warp_variance(feat_cl, proj.to(DEV), rng.to(DEV), D, Ht, Wt, cfg.mvs.inv_depth[s])
        assert var.shape == ref.shape
        truth = O.warp_variance(feat.double(), proj.double(), O.depth_hypotheses(rng.double(), D, cfg.mvs.inv_depth[s]).expand(B, D, Ht, Wt),
                                cfg.mvs.inv_depth[s])
        scale = max(1.0, float(ref.abs().max()))
        ref_noise = _md(ref, truth)                                             # the reference's own fp32 error
        assert _md(var, truth) <= max(1e-4 * scale, 2.0 * ref_noise)
        assert _md(var, ref) <= 5e-4 * scale
        # channels-last output (what the cuDNN NDHWC path consumes): same values, other memory order
        var_cl = ops.warp_variance(feat_cl, proj.to(DEV), rng.to(DEV), D, Ht, Wt, cfg.mvs.inv_depth[s], out_channels_last=True)
        assert var_cl.shape == var.shape and var_cl.permute(0, 2, 3, 4, 1).is_contiguous()
        assert torch.equal(var_cl.contiguous(), var)


# ------------------------------------------------------------------ K2 / a3
def test_depth_range(golden):
    cfg = _cfg(golden)
    for s in range(2):
        rng, prob = golden.t(f"s{s}_range_in"), golden.t(f"s{s}_prob")
        depth, ci, vol = ops.depth_range_from_prob(rng.to(DEV), prob.to(DEV), cfg.mvs.ci_scales[s], cfg.mvs.inv_depth[s])
        scale = float(golden.t(f"s{s}_depth").abs().max())
        assert _md(depth, golden.t(f"s{s}_depth")) <= 2e-5 * scale
        assert _md(ci, golden.t(f"s{s}_ci")) <= 2e-5 * scale
        dv = golden.t(f"s{s}_depth_values")
        assert _md(vol, dv[:, [0, -1]]) <= 1e-6 * float(dv.abs().max())


# ------------------------------------------------------------------ a4-a7 sampling
def _cam(golden, cfg, V_src=True):
    return ops.camera_block(golden.t("in_tar_exts").to(DEV), golden.t("in_tar_ints").to(DEV), golden.t("in_src_exts").to(DEV),
                            golden.t("in_src_ints").to(DEV), golden.t("in_near_far").to(DEV), cfg.nerf.bundle_size,
                            cfg.nerf.global_num_depth, cfg.mvs.inv_depth[-1])


@pytest.mark.parametrize("prefix", ["", "inj_"])
def test_sampling_bit_exact_indices(golden, prefix):
    cfg = _cfg(golden)
    adaptive = True if prefix else cfg.nerf.is_adaptive
    cam = _cam(golden, cfg)
    sl = ops.sample_bundles(golden.t(prefix + "depth_range").to(DEV), golden.t(prefix + "vol_range").to(DEV), cam,
                            cfg.nerf.bundle_size, cfg.nerf.max_num_samples, cfg.mvs.inv_depth[-1], adaptive)
    assert sl.indices.dtype == torch.int64
    assert np.array_equal(sl.indices.cpu().numpy(), golden.np(prefix + "indices"))                  # bit-exact
    assert np.array_equal(sl.counts.cpu().numpy().astype(np.int64), golden.np(prefix + "samples_per_bundle").astype(np.int64))
    assert sl.total == golden.np(prefix + "indices").shape[0]
    offs = sl.offsets.cpu().numpy()
    assert offs[0] == 0 and np.array_equal(np.diff(offs), sl.counts.cpu().numpy())
    zs = float(golden.t(prefix + "z_vals").abs().max())
    assert _md(sl.z_vals, golden.t(prefix + "z_vals")) <= 1e-6 * zs
    assert _md(sl.uvd, golden.t(prefix + "uvd")) <= 2e-5
    assert _md(sl.rays_xyz, golden.t(prefix + "rays_xyz")) <= 2e-6 * float(golden.t(prefix + "rays_xyz").abs().max())
    rb = golden.t(prefix + "ball_radii")
    assert _md(sl.ball_radii, rb) <= 2e-5 * float(rb.abs().max())


def test_sampler_mirror_api(golden):
    """BundleSampler keeps the reference's call sequence, return tuple and dtypes."""
    from gdb_nerf_b200.sampler import BundleSampler
    cfg = _cfg(golden)
    spec = CASE_SPECS[golden.name]
    smp = BundleSampler(cfg.nerf.global_num_depth, cfg.nerf.max_mipmap_level)
    with pytest.raises(ValueError, match="build_rays"):
        smp.sample(golden.t("inj_depth_range").to(DEV), golden.t("inj_vol_range").to(DEV), cfg.nerf.bundle_size, cfg.nerf.max_num_samples)
    nf = golden.t("in_near_far").to(DEV)
    smp.build_rays(golden.t("in_tar_exts").to(DEV), golden.t("in_tar_ints").to(DEV), (spec["H"], spec["W"]), nf[:, 0], nf[:, 1])
    out = smp.sample(golden.t("inj_depth_range").to(DEV), golden.t("inj_vol_range").to(DEV), cfg.nerf.bundle_size,
                     cfg.nerf.max_num_samples, cfg.mvs.inv_depth[-1], True)
    rays_xyz, uvd, z, ball, idx, per_batch, per_bundle = out
    assert per_bundle.dtype == torch.float32 and per_batch.dtype == torch.float32     # adaptive mode dtypes of the reference
    assert np.array_equal(per_batch.cpu().numpy(), golden.np("inj_samples_per_batch"))
    assert np.array_equal(idx.cpu().numpy(), golden.np("inj_indices"))


# ------------------------------------------------------------------ a8 sources
def test_prepare_sources(golden):
    cfg = _cfg(golden)
    spec = CASE_SPECS[golden.name]
    b, L = cfg.nerf.bundle_size, cfg.nerf.max_mipmap_level
    tex_ref = golden.t("tex_nchw")                                             # (B,V,F,Hb,Wb)
    B, V, F, Hb, Wb = tex_ref.shape
    rgb = golden.t("in_rgb")
    src = ops.prepare_sources(tex_ref[:, :, :F - 3].contiguous().to(DEV), rgb.to(DEV), b, L)
    lvl0 = ops.texture_level(src, B * V, Hb, Wb, 0).cpu()
    want = tex_ref.flatten(0, 1).permute(0, 2, 3, 1)
    assert torch.equal(lvl0[..., :F], want.contiguous())                       # exact: pure data movement + exact 2x2 mean
    assert float(lvl0[..., F:].abs().max()) == 0.0 if lvl0.shape[-1] > F else True
    mips = O.build_mips(want.contiguous(), L)
    for k in range(1, L + 1):
        got = ops.texture_level(src, B * V, Hb, Wb, k).cpu()[..., :F]
        assert torch.equal(got, mips[k])                                       # same association as nvdiffrast's box filter
    assert torch.equal(src.rgba.cpu()[..., :3], rgb.flatten(0, 1).permute(0, 2, 3, 1).contiguous())


# ------------------------------------------------------------------ K3 / a9-a12
def _render(golden, prefix, adaptive):
    cfg = _cfg(golden)
    spec = CASE_SPECS[golden.name]
    b = cfg.nerf.bundle_size
    tex_ref = golden.t("tex_nchw")
    B, V, F, Hb, Wb = tex_ref.shape
    feat_dim = F - 3
    cam = _cam(golden, cfg)
    dr, vr = golden.t(prefix + "depth_range").to(DEV), golden.t(prefix + "vol_range").to(DEV)
    sl = ops.sample_bundles(dr, vr, cam, b, cfg.nerf.max_num_samples, cfg.mvs.inv_depth[-1], adaptive, want_rays=False)
    src = ops.prepare_sources(tex_ref[:, :, :feat_dim].contiguous().to(DEV), golden.t("in_rgb").to(DEV), b, cfg.nerf.max_mipmap_level)
    vol_cl = ops.to_channels_last(golden.t("feat_volume").to(DEV), 8)
    mlp = ops.pack_mlp(golden.mlp(), feat_dim, device=DEV)
    out = ops.render_fused(src, vol_cl, dr, vr, cam, mlp, B, V, spec["H"], spec["W"], b, cfg.nerf.max_num_samples,
                           cfg.mvs.inv_depth[-1], adaptive, taps=sl)
    plain = ops.render_fused(src, vol_cl, dr, vr, cam, mlp, B, V, spec["H"], spec["W"], b, cfg.nerf.max_num_samples,
                             cfg.mvs.inv_depth[-1], adaptive)
    for k in ("feat", "depth", "opacity"):
        assert torch.equal(out[k], plain[k])                                   # taps do not change the result
    split = ops.render_fused(src, vol_cl, dr, vr, cam, mlp, B, V, spec["H"], spec["W"], b, cfg.nerf.max_num_samples,
                             cfg.mvs.inv_depth[-1], adaptive, out_channels_last=True)
    R = 3 * b * b                                                              # channels-last split output: same values
    assert torch.equal(split["fine"].permute(0, 3, 1, 2), plain["feat"][:, :R])
    assert torch.equal(split["dec_in"].permute(0, 3, 1, 2), plain["feat"][:, R:])
    assert torch.equal(split["depth"], plain["depth"])
    # padded decoder input: zero pad channels, or a constant-one first pad channel that carries the first convolution's bias
    for prec in (0, 1, 2):
        pad0 = ops.render_fused(src, vol_cl, dr, vr, cam, mlp, B, V, spec["H"], spec["W"], b, cfg.nerf.max_num_samples,
                                cfg.mvs.inv_depth[-1], adaptive, out_channels_last=True, pad_dec=True, precision=prec)
        pad1 = ops.render_fused(src, vol_cl, dr, vr, cam, mlp, B, V, spec["H"], spec["W"], b, cfg.nerf.max_num_samples,
                                cfg.mvs.inv_depth[-1], adaptive, out_channels_last=True, pad_dec=True, dec_one=True, precision=prec)
        nd = F + 8
        assert pad0["dec_in"].shape[-1] == (nd + 3) // 4 * 4 > nd
        assert torch.equal(pad0["dec_in"][..., :nd], pad1["dec_in"][..., :nd]) and torch.equal(pad0["fine"], pad1["fine"])
        assert float(pad0["dec_in"][..., nd:].abs().max()) == 0.0
        assert torch.equal(pad1["dec_in"][..., nd], torch.ones_like(pad1["dec_in"][..., nd]))
        assert float(pad1["dec_in"][..., nd + 1:].abs().sum()) == 0.0
        if prec == 0:
            assert torch.equal(pad0["dec_in"][..., :nd], split["dec_in"])
    # the decoder with its first bias inside the convolution (constant-one channel) = with the separate bias pass
    from gdb_nerf_b200.cnn import Decoder, decoder_fused
    torch.manual_seed(2)
    dec = Decoder(nd, 3, num_feats=64, num_layers=1, upscale_factor=b).to(DEV).eval().to(memory_format=torch.channels_last)
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            ya, ba = decoder_fused(dec, pad0["dec_in"].permute(0, 3, 1, 2))
            yb, bb = decoder_fused(dec, pad1["dec_in"].permute(0, 3, 1, 2), one_channel=nd)
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32
    assert torch.equal(ba, bb) and _md(ya, yb) <= 1e-5 * max(1.0, float(ya.abs().max()))
    # channels-last feature input to the source preparation: same texture
    feat_nhwc = tex_ref[:, :, :feat_dim].to(DEV).permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    src2 = ops.prepare_sources(feat_nhwc, golden.t("in_rgb").to(DEV), b, cfg.nerf.max_mipmap_level)
    assert torch.equal(src2.tex, src.tex) and torch.equal(src2.rgba, src.rgba)
    return out, (B, Hb, Wb)


@pytest.mark.parametrize("prefix", ["", "inj_"])
def test_render_fused_against_reference(golden, prefix):
    cfg = _cfg(golden)
    spec = CASE_SPECS[golden.name]
    adaptive = True if prefix else cfg.nerf.is_adaptive
    out, (B, Hb, Wb) = _render(golden, prefix, adaptive)
    noise = 3e-4 if spec["images"] == "noise" else 1e-4
    # gathered inputs of the MLP
    rfd_ref = golden.t(prefix + "rgbs_feat_dir")
    assert out["rgbs_feat_dir"].shape == rfd_ref.shape
    assert _md(out["vox_feat"], golden.t(prefix + "vox_feat")) <= 1e-4
    assert _md(out["rgbs_feat_dir"][..., :-4], rfd_ref[..., :-4]) <= noise
    assert _md(out["rgbs_feat_dir"][..., -1], rfd_ref[..., -1]) <= 1e-5
    assert _md(out["rgbs_feat_dir"][..., -4:-1], rfd_ref[..., -4:-1]) <= 2e-3   # unit difference of nearly parallel rays
    # MLP outputs, weights, composited maps
    assert _md(out["sigma"], golden.t(prefix + "sigma")) <= 1e-4
    assert _md(out["sample_feat"], golden.t(prefix + "feat")) <= noise
    assert _md(out["weights"], golden.t(prefix + "weights")) <= 1e-4
    ref_feat = golden.t(prefix + "bundle_feat").view(B, Hb, Wb, -1).permute(0, 3, 1, 2)
    assert _md(out["feat"], ref_feat) <= noise
    zs = float(golden.t(prefix + "z_vals").abs().max())
    # depth in normalised units (d - near) / (far - near): 1e-4 absolute
    assert _md(out["depth"].reshape(-1), golden.t(prefix + "bundle_depth")) <= 1e-4 * (spec["far"] - spec["near"])
    assert _md(out["depth"].reshape(-1), golden.t(prefix + "bundle_depth")) <= 4e-6 * zs                 # and a few ulp of z
    assert _md(out["opacity"].reshape(-1), golden.t(prefix + "bundle_opacity")) <= 1e-5


def test_render_fused_against_fp64_oracle(golden):
    """CUDA fp32 vs the float64 oracle on identical inputs: error no larger than ~the reference's own fp32 error."""
    cfg = _cfg(golden)
    spec = CASE_SPECS[golden.name]
    out, (B, Hb, Wb) = _render(golden, "inj_", True)
    tex = golden.t("tex_nchw", torch.float64)
    truth = O.render_bundles(
        golden.mlp(dtype=torch.float64), tex.shape[2] - 3, golden.t("in_rgb", torch.float64), tex[:, :, :-3], golden.t("feat_volume", torch.float64),
        golden.t("inj_depth_range"), golden.t("inj_vol_range"), golden.t("in_src_exts"), golden.t("in_src_ints"),
        golden.t("in_tar_exts"), golden.t("in_tar_ints"), golden.t("in_near_far"), cfg.nerf.bundle_size, cfg.nerf.max_num_samples,
        cfg.nerf.global_num_depth, cfg.nerf.max_mipmap_level, cfg.mvs.inv_depth[-1], True)
    # NB: sampling runs in float32 in both (counts are only defined there); everything downstream in float64
    ref_feat = golden.t("inj_bundle_feat").view(B, Hb, Wb, -1).permute(0, 3, 1, 2)
    ref_err = _md(ref_feat, truth["bundle_feat"])
    mine = _md(out["feat"], truth["bundle_feat"])
    assert mine <= max(1e-4, 2.0 * ref_err), (mine, ref_err)


def test_output_assembly(golden):
    cfg = _cfg(golden)
    spec = CASE_SPECS[golden.name]
    b = cfg.nerf.bundle_size
    B, Hb, Wb = spec["B"], spec["H"] // b, spec["W"] // b
    feat = golden.t("bundle_feat").view(B, Hb, Wb, -1).permute(0, 3, 1, 2).contiguous()
    bd, bo = golden.t("bundle_depth").view(B, Hb, Wb), golden.t("bundle_opacity").view(B, Hb, Wb)
    g = torch.Generator().manual_seed(5)
    dec = torch.rand(B, 3, spec["H"], spec["W"], generator=g)
    for rew in (False, True):
        rgb, depth, opac = ops.assemble_output(feat.to(DEV), dec.to(DEV), bd.to(DEV), bo.to(DEV), b, rew)
        fine = torch.nn.functional.pixel_shuffle(feat[:, :3 * b * b], b)
        want = dec + fine
        if rew:
            want = 0.5 * (want + fine)
        assert _md(rgb, want) <= 1e-6
        assert _md(depth, torch.nn.functional.interpolate(bd[:, None], scale_factor=b, mode="bilinear", align_corners=False)[:, 0]) <= 1e-6 * float(bd.abs().max())
        assert _md(opac, torch.nn.functional.interpolate(bo[:, None], scale_factor=b, mode="bilinear", align_corners=False)[:, 0]) <= 1e-6


def test_c_abi_rejects_unsupported_combinations():
    from gdb_nerf_b200 import _lib
    lib = _lib.load()
    z = torch.zeros(64, device=DEV)
    code = lib.gdb_warp_variance_fwd(z.data_ptr(), z.data_ptr(), z.data_ptr(), 1, 1, 1, 5, 32, 2, 2, 4, 2, 2, 0, 0, z.data_ptr(), None)
    assert code == -2 and b"not instantiated" in lib.gdb_last_error_string()
    code = lib.gdb_warp_variance_fwd(z.data_ptr(), z.data_ptr(), z.data_ptr(), 3, 1, 1, 3, 32, 2, 2, 4, 2, 2, 0, 0, z.data_ptr(), None)
    assert code == -1


# ------------------------------------------------------------------ K3, tensor-core MLP variant (precision = 1)
@pytest.mark.parametrize("prefix", ["", "inj_"])
def test_render_fused_tensor_core_variant(golden, prefix):
    """fp16-operand tcgen05 MLP: same indices/geometry, rgb/depth within the 2e-3 class of BASELINE.json's north star."""
    cfg = _cfg(golden)
    spec = CASE_SPECS[golden.name]
    b = cfg.nerf.bundle_size
    adaptive = True if prefix else cfg.nerf.is_adaptive
    tex_ref = golden.t("tex_nchw")
    B, V, F, Hb, Wb = tex_ref.shape
    feat_dim = F - 3
    cam = _cam(golden, cfg)
    dr, vr = golden.t(prefix + "depth_range").to(DEV), golden.t(prefix + "vol_range").to(DEV)
    sl = ops.sample_bundles(dr, vr, cam, b, cfg.nerf.max_num_samples, cfg.mvs.inv_depth[-1], adaptive, want_rays=False)
    src = ops.prepare_sources(tex_ref[:, :, :feat_dim].contiguous().to(DEV), golden.t("in_rgb").to(DEV), b, cfg.nerf.max_mipmap_level)
    vol_cl = ops.to_channels_last(golden.t("feat_volume").to(DEV), 8)
    mlp = ops.pack_mlp(golden.mlp(), feat_dim, device=DEV)
    args = (src, vol_cl, dr, vr, cam, mlp, B, V, spec["H"], spec["W"], b, cfg.nerf.max_num_samples, cfg.mvs.inv_depth[-1], adaptive)
    tc = ops.render_fused(*args, taps=sl, precision=1)
    ref32 = ops.render_fused(*args, taps=sl, precision=0)
    torch.cuda.synchronize()
    # gathers are the same code in fp32: identical inputs to the MLP
    assert _md(tc["rgbs_feat_dir"], ref32["rgbs_feat_dir"]) <= 1e-5
    assert _md(tc["vox_feat"], ref32["vox_feat"]) <= 1e-5
    # MLP in half precision operands / fp32 accumulation
    assert _md(tc["sigma"], golden.t(prefix + "sigma")) <= 2e-3
    assert _md(tc["sample_feat"], golden.t(prefix + "feat")) <= 2e-3
    assert _md(tc["weights"], golden.t(prefix + "weights")) <= 2e-3
    ref_feat = golden.t(prefix + "bundle_feat").view(B, Hb, Wb, -1).permute(0, 3, 1, 2)
    assert _md(tc["feat"], ref_feat) <= 2e-3
    assert _md(tc["depth"].reshape(-1), golden.t(prefix + "bundle_depth")) <= 2e-3 * (spec["far"] - spec["near"])
    assert _md(tc["opacity"].reshape(-1), golden.t(prefix + "bundle_opacity")) <= 1e-5
    # channels-last outputs of the tensor-core variant
    split = ops.render_fused(*args, precision=1, out_channels_last=True)
    R = 3 * b * b
    assert torch.equal(split["fine"].permute(0, 3, 1, 2), tc["feat"][:, :R])
    assert torch.equal(split["dec_in"].permute(0, 3, 1, 2), tc["feat"][:, R:])


@pytest.mark.parametrize("prefix", ["", "inj_"])
def test_render_fused_batched_gather_kernel_bit_identical(golden, prefix):
    """precision=1 (third generation, the default: batched predicated gathers, descriptors through shared memory, float4
    compositing) performs the same operations in the same order as the round-1 kernel kept as precision=4: every output, tap and
    layout is bit-identical.  The fourth generation (precision=6, a measured variant: weighted-sum taps instead of nested lerps)
    differs from both only by fp32 rounding of the gathered features in front of their fp16 rounding."""
    cfg = _cfg(golden)
    spec = CASE_SPECS[golden.name]
    b = cfg.nerf.bundle_size
    adaptive = True if prefix else cfg.nerf.is_adaptive
    tex_ref = golden.t("tex_nchw")
    B, V, F, Hb, Wb = tex_ref.shape
    feat_dim = F - 3
    cam = _cam(golden, cfg)
    dr, vr = golden.t(prefix + "depth_range").to(DEV), golden.t(prefix + "vol_range").to(DEV)
    sl = ops.sample_bundles(dr, vr, cam, b, cfg.nerf.max_num_samples, cfg.mvs.inv_depth[-1], adaptive, want_rays=False)
    src = ops.prepare_sources(tex_ref[:, :, :feat_dim].contiguous().to(DEV), golden.t("in_rgb").to(DEV), b, cfg.nerf.max_mipmap_level)
    vol_cl = ops.to_channels_last(golden.t("feat_volume").to(DEV), 8)
    mlp = ops.pack_mlp(golden.mlp(), feat_dim, device=DEV)
    args = (src, vol_cl, dr, vr, cam, mlp, B, V, spec["H"], spec["W"], b, cfg.nerf.max_num_samples, cfg.mvs.inv_depth[-1], adaptive)
    for kw in (dict(taps=sl), dict(), dict(out_channels_last=True), dict(out_channels_last=True, pad_dec=True, dec_one=True)):
        new = ops.render_fused(*args, precision=1, **kw)
        old = ops.render_fused(*args, precision=4, **kw)
        cur = ops.render_fused(*args, precision=6, **kw)
        torch.cuda.synchronize()
        assert set(new) == set(old) == set(cur)
        for k in new:
            assert torch.equal(new[k], old[k]), (k, kw.keys(), _md(new[k], old[k]))
            if k in ("rgbs_feat_dir", "vox_feat"):      # fp32 gathers: same taps, different summation order
                # the colour taps inside rgbs_feat_dir additionally come through two equivalent projection chains since the
                # last session of round 2 (precision 1 / 4: composed matrix + MUFU.RCP, gdb_render_tc2.cu; precision 6: step by
                # step + correctly rounded reciprocal): a few 1e-5 pixels of tap position, 4.8e-6 of colour measured
                tol = (1e-5 if k == "rgbs_feat_dir" else 2e-6) * (1.0 + float(old[k].abs().max()))
                assert _md(cur[k], old[k]) <= tol, (k, _md(cur[k], old[k]), tol)
            else:                                        # downstream of fp16 operand rounding (2^-11 relative steps)
                tol = 1e-4 * (float(dr.max()) - float(dr.min())) if k == "depth" else 4e-4 * (1.0 + float(old[k].abs().max()))
                assert _md(cur[k], old[k]) <= tol, (k, _md(cur[k], old[k]), tol)


def test_render_fused_row_range_stitches_bit_exact(golden):
    """Image-tile split (north star: work partitioned by target views AND image tiles): rendering the bundle rows of a view
    in pieces writes exactly the bytes of the one-piece render, and nothing outside the requested rows."""
    cfg = _cfg(golden)
    spec = CASE_SPECS[golden.name]
    b = cfg.nerf.bundle_size
    tex_ref = golden.t("tex_nchw")
    B, V, F, Hb, Wb = tex_ref.shape
    feat_dim = F - 3
    cam = _cam(golden, cfg)
    dr, vr = golden.t("inj_depth_range").to(DEV), golden.t("inj_vol_range").to(DEV)
    src = ops.prepare_sources(tex_ref[:, :, :feat_dim].contiguous().to(DEV), golden.t("in_rgb").to(DEV), b, cfg.nerf.max_mipmap_level)
    vol_cl = ops.to_channels_last(golden.t("feat_volume").to(DEV), 8)
    mlp = ops.pack_mlp(golden.mlp(), feat_dim, device=DEV)
    args = (src, vol_cl, dr, vr, cam, mlp, B, V, spec["H"], spec["W"], b, cfg.nerf.max_num_samples, cfg.mvs.inv_depth[-1], True)
    kw = dict(precision=1, out_channels_last=True, pad_dec=True, dec_one=True)
    full = ops.render_fused(*args, **kw)
    cuts = [0, 1, Hb // 3, Hb // 3 + 5, Hb]
    stitched = {k: torch.full_like(v, float("nan")) for k, v in full.items()}
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        part = ops.render_fused(*args, rows=(lo, hi), **kw)
        for k in full:
            stitched[k][:, lo:hi] = part[k][:, lo:hi]
    torch.cuda.synchronize()
    for k in full:
        assert torch.equal(stitched[k], full[k]), k
    # a piece writes nothing outside its rows
    canary = {k: None for k in full}
    lo, hi = cuts[1], cuts[2]
    import gdb_nerf_b200.ops as _ops
    real_empty = torch.empty
    try:
        torch.empty = lambda *a, **k: real_empty(*a, **k).fill_(-7.0) if k.get("dtype") == torch.float32 else real_empty(*a, **k)
        part = _ops.render_fused(*args, rows=(lo, hi), **kw)
    finally:
        torch.empty = real_empty
    torch.cuda.synchronize()
    for k, v in part.items():
        assert bool((v[:, :lo] == -7.0).all()) and bool((v[:, hi:] == -7.0).all()), k
    # partial ranges exist for the production kernel only
    with pytest.raises(ops._lib.GdbError, match="partial row range"):
        ops.render_fused(*args, rows=(0, 1), precision=0)


# ------------------------------------------------------------------ glue next to the path
def test_depth_range_from_logits_strided(golden):
    """Soft-max fused into K2, logits read in place as channel 8 of a 12-channel channels-last head output."""
    cfg = _cfg(golden)
    g = torch.Generator().manual_seed(11)
    for s in range(2):
        rng, prob = golden.t(f"s{s}_range_in"), golden.t(f"s{s}_prob")
        B, D, h, w = prob.shape
        logits = torch.log(prob.clamp_min(1e-30)) + torch.randn(B, 1, h, w, generator=g)      # softmax(logits) == prob
        heads = torch.randn(B, D, h, w, 12, generator=g)
        heads[..., 8] = logits
        hd = heads.to(DEV)
        depth, ci, vol, pr = ops.depth_range_from_logits(rng.to(DEV), hd[..., 8], cfg.mvs.ci_scales[s], cfg.mvs.inv_depth[s], want_prob=True)
        assert _md(pr, torch.softmax(logits.double(), 1)) <= 5e-7
        scale = float(golden.t(f"s{s}_depth").abs().max())
        assert _md(depth, golden.t(f"s{s}_depth")) <= 2e-5 * scale
        assert _md(ci, golden.t(f"s{s}_ci")) <= 2e-5 * scale
        d2, c2, v2 = ops.depth_range_from_prob(rng.to(DEV), pr, cfg.mvs.ci_scales[s], cfg.mvs.inv_depth[s])
        assert torch.equal(d2, depth) and torch.equal(c2, ci) and torch.equal(v2, vol)           # same arithmetic after the soft-max


def test_glue_epilogues_match_torch():
    g = torch.Generator().manual_seed(3)
    F_ = torch.nn.functional
    x = torch.randn(2, 8, 4, 6, 10, generator=g).to(DEV).contiguous(memory_format=torch.channels_last_3d)
    skip = torch.randn(2, 8, 4, 6, 10, generator=g).to(DEV).contiguous(memory_format=torch.channels_last_3d)
    bias = torch.randn(8, generator=g).to(DEV)
    out = ops.bias_act_add(x, bias, skip, relu=True)
    assert torch.equal(out, skip + (x + bias.view(1, -1, 1, 1, 1)).relu())
    lat = torch.randn(3, 32, 8, 12, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    top = torch.randn(3, 32, 4, 6, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    b2 = torch.randn(32, generator=g).to(DEV)
    out = ops.bias_act_add(lat, b2, top, relu=False, skip_up2=True)
    want = F_.interpolate(top, size=(8, 12), mode="nearest") + (lat + b2.view(1, -1, 1, 1))
    assert _md(out, want) <= 1e-6
    y = torch.randn(3, 32, 8, 12, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    gate = torch.rand(3, 32, generator=g).to(DEV)
    assert _md(ops.gate_add(lat, y, gate), lat + y * gate[:, :, None, None]) <= 1e-6
    assert _md(ops.gate_add(lat, y, gate, extra=y), lat + y * gate[:, :, None, None] + y) <= 1e-6
    with pytest.raises(ops._lib.GdbError):
        ops.bias_act_add(lat.contiguous(), b2, None, relu=False)                                 # planar memory is refused, not converted
    # dense-block concatenation and the squeeze of the squeeze-excite gate (decoder_rdn.py:36-41, modules.py)
    a8 = torch.randn(3, 8, 8, 12, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    cat3 = ops.concat_channels(lat, a8, y)
    assert torch.equal(cat3, torch.cat((lat, a8, y), 1)) and ops._is_cl(cat3)
    assert torch.equal(ops.concat_channels(lat, a8), torch.cat((lat, a8), 1))
    ps_in = torch.randn(2, 32, 6, 10, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    ps_b = torch.randn(32, generator=g).to(DEV)
    ps = ops.pixel_shuffle2_bias(ps_in, ps_b)
    assert torch.equal(ps, F_.pixel_shuffle(ps_in + ps_b.view(1, -1, 1, 1), 2)) and ops._is_cl(ps)
    big = torch.randn(2, 64, 50, 70, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    assert _md(ops.channel_mean(big), big.double().mean((2, 3))) <= 1e-6
    assert _md(ops.channel_mean(big, chunks=7), big.double().mean((2, 3))) <= 1e-6
    # squeeze-excite gate finished inside the residual kernel (modules.py:5-20 + decoder_rdn.py:41)
    hx = torch.randn(2, 64, 50, 70, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    w1 = (torch.randn(4, 64, generator=g) * 0.5).to(DEV)
    w2 = (torch.randn(64, 4, generator=g) * 0.5).to(DEV)
    gd = torch.sigmoid(torch.relu(big.double().mean((2, 3)) @ w1.double().t()) @ w2.double().t())
    assert _md(ops.se_gate_add(hx, big, w1, w2), hx.double() + big.double() * gd[:, :, None, None]) <= 2e-6
    assert _md(ops.se_gate_add(hx, big, w1, w2, extra=big, chunks=5), hx.double() + big.double() * gd[:, :, None, None] + big.double()) <= 2e-6
    # ... and written a second time into the leading channels of the next block's concatenation buffer, whose other slices
    # concat_into fills
    buf = torch.full((2, 50, 70, 64 + 8 + 32), float("nan"), device=DEV).permute(0, 3, 1, 2)
    out_c = ops.se_gate_add(hx, big, w1, w2, cat_out=buf)
    assert torch.equal(out_c, ops.se_gate_add(hx, big, w1, w2)) and torch.equal(buf[:, :64], out_c)
    a8b = torch.randn(2, 8, 50, 70, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    b32 = torch.randn(2, 32, 50, 70, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    assert ops.concat_into(buf, 64, a8b, b32) is buf and torch.equal(buf, torch.cat((out_c, a8b, b32), 1))
    w1b, w2b = (torch.randn(2, 32, generator=g) * 0.5).to(DEV), (torch.randn(32, 2, generator=g) * 0.5).to(DEV)      # R = 2: padded hid
    gd = torch.sigmoid(torch.relu(y.double().mean((2, 3)) @ w1b.double().t()) @ w2b.double().t())
    assert _md(ops.se_gate_add(lat, y, w1b, w2b), lat.double() + y.double() * gd[:, :, None, None]) <= 2e-6


def test_assemble_output_pre_shuffle_decoder(golden):
    cfg = _cfg(golden)
    spec = CASE_SPECS[golden.name]
    b = cfg.nerf.bundle_size
    B, H, W = spec["B"], spec["H"], spec["W"]
    Hb, Wb = H // b, W // b
    feat = golden.t("bundle_feat").view(B, Hb, Wb, -1).contiguous()
    bd, bo = golden.t("bundle_depth").view(B, Hb, Wb), golden.t("bundle_opacity").view(B, Hb, Wb)
    dec12 = torch.rand(B, H // 2, W // 2, 12, generator=torch.Generator().manual_seed(6))
    dec = torch.nn.functional.pixel_shuffle(dec12.permute(0, 3, 1, 2), 2)
    fine = torch.nn.functional.pixel_shuffle(feat.permute(0, 3, 1, 2)[:, :3 * b * b], b)
    for rew in (False, True):
        rgb, _, _ = ops.assemble_output(feat.to(DEV), dec12.to(DEV), bd.to(DEV), bo.to(DEV), b, rew, feat_channels_last=True,
                                        dec_pre_shuffle=True)
        want = dec + fine
        if rew:
            want = 0.5 * (want + fine)
        assert _md(rgb, want) <= 1e-6
        # with the composed convolution's bias applied inside the assembly
        bias12 = torch.rand(12, generator=torch.Generator().manual_seed(7))
        rgb_b, _, _ = ops.assemble_output(feat.to(DEV), dec12.to(DEV), bd.to(DEV), bo.to(DEV), b, rew, feat_channels_last=True,
                                          dec_pre_shuffle=True, dec_bias=bias12.to(DEV))
        want_b = torch.nn.functional.pixel_shuffle((dec12 + bias12).permute(0, 3, 1, 2), 2) + fine
        if rew:
            want_b = 0.5 * (want_b + fine)
        assert _md(rgb_b, want_b) <= 1e-6


def test_render_fused_reads_strided_volume(golden):
    """The feature volume may be the leading 8 channels of a wider channels-last tensor (vol_stride = 12)."""
    cfg = _cfg(golden)
    spec = CASE_SPECS[golden.name]
    b = cfg.nerf.bundle_size
    B, V, H, W = spec["B"], spec["V"], spec["H"], spec["W"]
    cam = _cam(golden, cfg)
    tex_ref = golden.t("tex_nchw")
    fd = tex_ref.shape[2] - 3
    src = ops.prepare_sources(tex_ref[:, :, :fd].contiguous().to(DEV), golden.t("in_rgb").to(DEV), b, cfg.nerf.max_mipmap_level)
    vol = ops.to_channels_last(golden.t("feat_volume").to(DEV), 8)
    wide = torch.randn(*vol.shape[:-1], 12, device=DEV)
    wide[..., :8] = vol
    mlp = ops.pack_mlp(golden.mlp(), fd, device=DEV)
    args = (golden.t("depth_range").to(DEV), golden.t("vol_range").to(DEV), cam, mlp, B, V, H, W, b, cfg.nerf.max_num_samples,
            cfg.mvs.inv_depth[-1], cfg.nerf.is_adaptive)
    for prec in (0, 1):
        a = ops.render_fused(src, vol, *args, precision=prec)
        c = ops.render_fused(src, wide[..., :8], *args, precision=prec)
        assert torch.equal(a["feat"], c["feat"]) and torch.equal(a["depth"], c["depth"])


@pytest.mark.parametrize("prefix", ["", "inj_"])
def test_render_fused_split_precision_tensor_core(golden, prefix):
    """precision=2: MLP GEMMs on tcgen05 with every operand split into two fp16 planes (hi + lo, three MMAs per K step,
    fp32 accumulation in TMEM) - the fp32 class of BASELINE.json's north star: same tolerances as the SIMT fp32 kernel."""
    cfg = _cfg(golden)
    spec = CASE_SPECS[golden.name]
    b = cfg.nerf.bundle_size
    adaptive = True if prefix else cfg.nerf.is_adaptive
    tex_ref = golden.t("tex_nchw")
    B, V, F, Hb, Wb = tex_ref.shape
    feat_dim = F - 3
    cam = _cam(golden, cfg)
    dr, vr = golden.t(prefix + "depth_range").to(DEV), golden.t(prefix + "vol_range").to(DEV)
    sl = ops.sample_bundles(dr, vr, cam, b, cfg.nerf.max_num_samples, cfg.mvs.inv_depth[-1], adaptive, want_rays=False)
    src = ops.prepare_sources(tex_ref[:, :, :feat_dim].contiguous().to(DEV), golden.t("in_rgb").to(DEV), b, cfg.nerf.max_mipmap_level)
    vol_cl = ops.to_channels_last(golden.t("feat_volume").to(DEV), 8)
    mlp = ops.pack_mlp(golden.mlp(), feat_dim, device=DEV)
    args = (src, vol_cl, dr, vr, cam, mlp, B, V, spec["H"], spec["W"], b, cfg.nerf.max_num_samples, cfg.mvs.inv_depth[-1], adaptive)
    sp = ops.render_fused(*args, taps=sl, precision=2)
    ref32 = ops.render_fused(*args, taps=sl, precision=0)
    torch.cuda.synchronize()
    noise = 3e-4 if spec["images"] == "noise" else 1e-4
    assert _md(sp["sigma"], golden.t(prefix + "sigma")) <= 1e-4
    assert _md(sp["sample_feat"], golden.t(prefix + "feat")) <= noise
    assert _md(sp["weights"], golden.t(prefix + "weights")) <= 1e-4
    ref_feat = golden.t(prefix + "bundle_feat").view(B, Hb, Wb, -1).permute(0, 3, 1, 2)
    assert _md(sp["feat"], ref_feat) <= noise
    assert _md(sp["depth"].reshape(-1), golden.t(prefix + "bundle_depth")) <= 1e-4 * (spec["far"] - spec["near"])
    assert _md(sp["opacity"].reshape(-1), golden.t(prefix + "bundle_opacity")) <= 1e-5
    # and it agrees with the SIMT fp32 kernel to a few fp32 ulps of the activations
    assert _md(sp["sigma"], ref32["sigma"]) <= 2e-5
    assert _md(sp["feat"], ref32["feat"]) <= 2e-5
    # without the parity taps the production kernel of this arithmetic runs: for 2x2 bundles and three source views that is the
    # split-fp16 kernel in the third-generation layout (gdb_render_tc4.cu: `mean` folded into extra MMAs, var[16:19] in the spare
    # K slots of [x_v | 1]), elsewhere the first-generation split kernel itself.  Same class, not the same summation order.
    R = 3 * b * b
    exact = not (b == 2 and feat_dim == 16 and V == 3)
    for kw in (dict(), dict(out_channels_last=True), dict(out_channels_last=True, pad_dec=True, dec_one=True)):
        prod = ops.render_fused(*args, precision=2, **kw)
        torch.cuda.synchronize()
        if kw:
            fine, dec = prod["fine"].permute(0, 3, 1, 2), prod["dec_in"].permute(0, 3, 1, 2)[:, :sp["feat"].shape[1] - R]
        else:
            fine, dec = prod["feat"][:, :R], prod["feat"][:, R:]
        if exact:
            assert torch.equal(fine, sp["feat"][:, :R]) and torch.equal(dec, sp["feat"][:, R:])
        else:
            assert _md(fine, sp["feat"][:, :R]) <= 2e-5 and _md(dec, sp["feat"][:, R:]) <= 2e-5, (kw, _md(fine, sp["feat"][:, :R]), _md(dec, sp["feat"][:, R:]))
            assert _md(fine, ref_feat[:, :R]) <= noise and _md(dec, ref_feat[:, R:]) <= noise
        assert _md(prod["depth"].reshape(-1), golden.t(prefix + "bundle_depth")) <= 1e-4 * (spec["far"] - spec["near"])
        assert _md(prod["depth"], sp["depth"]) <= 2e-6 * (spec["far"] - spec["near"])
        assert _md(prod["opacity"].reshape(-1), golden.t(prefix + "bundle_opacity")) <= 1e-5


@pytest.mark.parametrize("D,h,w,inv,per_pixel", [(64, 12, 40, True, False), (36, 9, 33, True, False), (8, 10, 70, False, True)])
def test_prob_head_fused_into_depth_range(D, h, w, inv, per_pixel):
    """1-channel probability head (Conv3d 8 -> 1, 3x3x3, padding 1) + soft-max + depth regression in one kernel against the
    two-step path (true-fp32 cuDNN convolution, then the logits variant of K2), including ragged tiles."""
    g = torch.Generator().manual_seed(D + h)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        B = 2
        y = (torch.randn(B, 8, D, h, w, generator=g) * 0.7).to(DEV).contiguous(memory_format=torch.channels_last_3d)
        wt = (torch.randn(1, 8, 3, 3, 3, generator=g) * 0.2).to(DEV)
        if per_pixel:
            near = 425.0 + 100.0 * torch.rand(B, 1, h, w, generator=g)
            rng = torch.cat((near, near + 40.0 + 40.0 * torch.rand(B, 1, h, w, generator=g)), 1).to(DEV)
        else:
            rng = torch.tensor([[425.0, 905.0]] * B).view(B, 2, 1, 1).to(DEV)
        logits = torch.nn.functional.conv3d(y, wt, None, 1, 1).squeeze(1)
        want = ops.depth_range_from_logits(rng, logits, 1.0, inv, want_prob=True)
        got = ops.prob_head_depth_range(y, wt, rng, 1.0, inv, want_prob=True)
        torch.cuda.synchronize()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert _md(got[3], want[3]) <= 2e-6                               # probabilities
    for k in range(3):                                                # depth, confidence interval, volume range
        assert _md(got[k], want[k]) <= 1e-5 * 480.0, k
    # depth axis split over several CTAs per tile (on-line soft-max partials); twice: the arrival counters reset themselves
    for _ in range(2):
        sp = ops.prob_head_depth_range(y, wt, rng, 1.0, inv, tma=False)
        assert sp[3] is None
        for k in range(3):
            assert _md(sp[k], want[k]) <= 1e-5 * 480.0, k
    one = ops.prob_head_depth_range(y, wt, rng, 1.0, inv, split=False, tma=False)
    for k in range(3):
        assert torch.equal(one[k], got[k])
    # second generation (the default): planes by TMA with the out-of-bounds zero fill as padding, each plane read once and
    # scattered into three output planes, weights as constant operands; depth axis split and unsplit, twice each
    for split in (True, False):
        for _ in range(2):
            tm = ops.prob_head_depth_range(y, wt, rng, 1.0, inv, split=split, tma=True)
            assert tm[3] is None
            for k in range(3):
                assert _md(tm[k], want[k]) <= 1e-5 * 480.0, (split, k, _md(tm[k], want[k]))
    # other weights on the same stream: the constant-memory copy is ordered with the launches
    wt2 = (torch.randn(1, 8, 3, 3, 3, generator=g) * 0.2).to(DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        want2 = ops.depth_range_from_logits(rng, torch.nn.functional.conv3d(y, wt2, None, 1, 1).squeeze(1), 1.0, inv)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    tm2 = ops.prob_head_depth_range(y, wt2, rng, 1.0, inv, tma=True)
    tm1 = ops.prob_head_depth_range(y, wt, rng, 1.0, inv, tma=True)
    for k in range(3):
        assert _md(tm2[k], want2[k]) <= 1e-5 * 480.0 and _md(tm1[k], want[k]) <= 1e-5 * 480.0, k
