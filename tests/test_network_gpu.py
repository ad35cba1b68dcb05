"""End-to-end drop-in check and full-size (BASELINE.json configs[1]) parity / property tests."""
import numpy as np
import pytest
import torch

from conftest import CASE_SPECS, load_golden
from gdb_nerf_b200 import ops
from gdb_nerf_b200.config import make_cfg
from gdb_nerf_b200.network import Network
from gdb_nerf_b200.synthetic import WORKLOADS, batch_to, camera_rig, make_batch, synth_state_dict
from oracle import gdb_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _md(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


@pytest.fixture(autouse=True)
def _fp32_cnn():
    # parity runs keep the cuDNN networks in true fp32 (TF32 is PyTorch's default for convolutions)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("cnn_mode", ["fused", "modules"])
@pytest.mark.parametrize("case", ["dtu_b2", "nerf_b4"])
def test_network_forward_matches_reference(case, cnn_mode):
    """Same weights, same batch dict -> same (ret, mvs_depths, blend_rgbs) as the reference's Network.forward."""
    g = load_golden(case)
    spec = CASE_SPECS[case]
    net = Network(make_cfg(spec["recipe"]))
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(synth_state_dict(shapes, seed=1), strict=True)
    net = net.to(DEV).eval()
    net.cnn_mode = cnn_mode
    batch = make_batch(spec["B"], spec["V"], spec["H"], spec["W"], spec["near"], spec["far"], spec["focal"], seed=3,
                       images=spec["images"], tilt=spec["tilt"])
    assert np.array_equal(batch["src_views"]["rgb"].numpy(), g.np("in_rgb"))
    keep = {k: v.clone() for k, v in batch["src_views"].items()}
    with torch.no_grad():
        ret, mvs_depths, blend = net(batch_to(batch, DEV))
    assert blend == [] and len(mvs_depths) == 2
    for k, v in keep.items():                                                   # inputs are not mutated
        assert torch.equal(batch["src_views"][k], v)
    assert set(ret) == {"rgb", "nerf_depth", "mvs_depth", "opacity"}
    for k in ret:
        assert ret[k].shape == g.t("ret_" + k).shape, k
    depth_scale = spec["far"] - spec["near"]
    # cuDNN vs MKL-DNN convolutions differ in summation order; 1e-3 covers the CNNs, the star path itself is
    # checked to 1e-4 on identical inputs in test_kernels_gpu.py
    assert _md(ret["rgb"], g.t("ret_rgb")) <= 1e-3
    assert _md(ret["opacity"], g.t("ret_opacity")) <= 1e-4
    assert _md(ret["mvs_depth"], g.t("ret_mvs_depth")) <= 1e-3 * depth_scale
    assert _md(ret["nerf_depth"], g.t("ret_nerf_depth")) <= 1e-3 * depth_scale
    for i, d in enumerate(mvs_depths):
        assert _md(d, g.t(f"mvs_depth_{i}")) <= 1e-3 * depth_scale
    mse = float(((ret["rgb"].cpu().double() - g.t("ret_rgb").double()) ** 2).mean())
    assert mse < 1e-8           # PSNR delta against the reference image far below 0.01 dB


def _full_size_inputs(name="dtu", V=3, seed=0, hw=None):
    w = dict(WORKLOADS[name])
    cfg = make_cfg(w["recipe"])
    b = cfg.nerf.bundle_size
    if hw is not None:                      # reduced image, same field of view
        w["focal"] = w["focal"] * hw[0] / w["H"]
        w["H"], w["W"] = hw
    H, W = w["H"], w["W"]
    Hb, Wb = H // b, W // b
    g = torch.Generator().manual_seed(seed)
    lvl = 0
    while cfg.fpn.feat_scales[lvl] < 1.0 / b:
        lvl += 1
    feat_dim = cfg.fpn.feat_dims[lvl]
    rig = camera_rig(1, V, H, W, w["near"], w["far"], w["focal"], tilt=0.03)
    from gdb_nerf_b200.synthetic import smooth_images
    data = dict(
        rgb=smooth_images(1, V, H, W, seed=seed),
        feat=torch.randn(1, V, feat_dim, Hb, Wb, generator=g) * 0.5,
        vol=torch.randn(1, 8, 8, Hb, Wb, generator=g) * 0.5,
    )
    min_iv = (w["far"] - w["near"]) / cfg.nerf.global_num_depth
    mid = w["near"] + (w["far"] - w["near"]) * (0.3 + 0.4 * torch.rand(1, 1, Hb, Wb, generator=g))
    half = torch.rand(1, 1, Hb, Wb, generator=g) * (0.55 * cfg.nerf.max_num_samples * min_iv)
    data["depth_range"] = torch.cat((mid - half, mid + half), 1)
    data["vol_range"] = torch.cat((mid - 2.5 * min_iv, mid + 2.5 * min_iv), 1)
    from gdb_nerf_b200.nerf import NeRF
    torch.manual_seed(seed)
    mlp = {k: v.detach() for k, v in NeRF(64, feat_dim, 8, True).state_dict().items()}
    return cfg, w, rig, data, mlp, feat_dim


def test_full_size_sampling_properties_and_counts():
    cfg, w, rig, data, mlp, feat_dim = _full_size_inputs()
    b = cfg.nerf.bundle_size
    cam = ops.camera_block(rig["tar_exts"].to(DEV), rig["tar_ints"].to(DEV), rig["src_exts"].to(DEV), rig["src_ints"].to(DEV),
                           rig["near_far"].to(DEV), b, cfg.nerf.global_num_depth, False)
    sl = ops.sample_bundles(data["depth_range"].to(DEV), data["vol_range"].to(DEV), cam, b, cfg.nerf.max_num_samples, False, True)
    idx = sl.indices.cpu()
    counts = sl.counts.cpu().long()
    NB = counts.numel()
    assert NB == 81920
    assert int(counts.min()) >= 1 and int(counts.max()) <= cfg.nerf.max_num_samples and len(torch.unique(counts)) == cfg.nerf.max_num_samples
    assert bool((idx[1:] >= idx[:-1]).all())                                    # sorted by bundle
    assert torch.equal(torch.bincount(idx, minlength=NB), counts)               # every bundle appears count times
    assert sl.total == int(counts.sum()) == idx.numel()
    # bit-exact against the oracle at full size
    nf = rig["near_far"]
    want = O.sample_counts(data["depth_range"][:, 0].reshape(-1), data["depth_range"][:, 1].reshape(-1),
                           ((nf[:, 1] - nf[:, 0]) / cfg.nerf.global_num_depth).expand(NB), cfg.nerf.max_num_samples)
    assert torch.equal(counts, want.long())
    assert torch.equal(idx, torch.repeat_interleave(torch.arange(NB), counts))


@pytest.mark.parametrize("workload,precision", [("dtu", 0), ("dtu", 1), ("dtu", 2), ("nerf", 0), ("nerf", 1), ("nerf", 2), ("llff", 1), ("llff", 2)])
def test_full_size_render_against_oracle_and_view_symmetry(workload, precision):
    """BASELINE.json sizes (DTU 512x640 2x2 bundles, NeRF-synthetic 800x800 4x4 bundles): parity with the oracle on
    identical inputs plus size-independent properties.  precision 1 = tensor-core MLP (2e-3 class)."""
    cfg, w, rig, data, mlp, feat_dim = _full_size_inputs(workload)
    b = cfg.nerf.bundle_size
    H, W = w["H"], w["W"]
    tol = 2e-3 if precision == 1 else 1e-4

    def run(order):
        o = torch.tensor(order)
        cam = ops.camera_block(rig["tar_exts"].to(DEV), rig["tar_ints"].to(DEV), rig["src_exts"][:, o].to(DEV),
                               rig["src_ints"][:, o].to(DEV), rig["near_far"].to(DEV), b, cfg.nerf.global_num_depth, False)
        src = ops.prepare_sources(data["feat"][:, o].contiguous().to(DEV), data["rgb"][:, o].contiguous().to(DEV), b, cfg.nerf.max_mipmap_level)
        vol_cl = ops.to_channels_last(data["vol"].to(DEV), 8)
        return ops.render_fused(src, vol_cl, data["depth_range"].to(DEV), data["vol_range"].to(DEV), cam,
                                ops.pack_mlp(mlp, feat_dim, device=DEV), 1, 3, H, W, b, cfg.nerf.max_num_samples, False, True,
                                precision=precision)

    out = run([0, 1, 2])
    # property: aggregation over source views is symmetric
    perm = run([2, 0, 1])
    assert _md(out["feat"], perm["feat"]) <= (1e-3 if precision == 1 else 2e-5)
    assert _md(out["depth"], perm["depth"]) <= tol * (w["far"] - w["near"])
    # property: weights are renormalised per bundle -> opacity == 1
    assert _md(out["opacity"], torch.ones_like(out["opacity"])) <= 1e-5
    dr = data["depth_range"]
    slack = 1e-5 * w["far"]
    assert bool(((out["depth"].cpu() >= dr[:, 0] - slack) & (out["depth"].cpu() <= dr[:, 1] + slack)).all())
    # full-size parity with the oracle (float32, identical inputs)
    truth = O.render_bundles(mlp, feat_dim, data["rgb"], data["feat"], data["vol"], data["depth_range"], data["vol_range"],
                             rig["src_exts"], rig["src_ints"], rig["tar_exts"], rig["tar_ints"], rig["near_far"], b,
                             cfg.nerf.max_num_samples, cfg.nerf.global_num_depth, cfg.nerf.max_mipmap_level, False, True)
    assert _md(out["feat"], truth["bundle_feat"]) <= tol
    assert _md(out["depth"], truth["bundle_depth"]) <= tol * (w["far"] - w["near"])


@pytest.mark.parametrize("precision", [1, 2])
def test_dynamic_tile_assignment_on_concurrent_streams_and_graph_replays(precision):
    """The tensor-core kernels of 2x2 bundles draw their tiles from an atomic counter of the library: one counter per stream
    (launches on a stream are ordered), a fresh one per launch recorded in a stream capture.  Launches in flight on several
    streams, and a graph replay overlapping eager launches, must never share a counter - a shared one would skip tiles and
    leave `torch.empty` garbage in the outputs.  Every result equals the serial one bit for bit (tiles are independent)."""
    cfg, w, rig, data, mlp, feat_dim = _full_size_inputs("dtu")
    b = cfg.nerf.bundle_size
    cam = ops.camera_block(rig["tar_exts"].to(DEV), rig["tar_ints"].to(DEV), rig["src_exts"].to(DEV), rig["src_ints"].to(DEV),
                           rig["near_far"].to(DEV), b, cfg.nerf.global_num_depth, False)
    src = ops.prepare_sources(data["feat"].to(DEV), data["rgb"].to(DEV), b, cfg.nerf.max_mipmap_level)
    vol_cl = ops.to_channels_last(data["vol"].to(DEV), 8)
    dr, vr, packed = data["depth_range"].to(DEV), data["vol_range"].to(DEV), ops.pack_mlp(mlp, feat_dim, device=DEV)

    def run():
        return ops.render_fused(src, vol_cl, dr, vr, cam, packed, 1, 3, w["H"], w["W"], b, cfg.nerf.max_num_samples, False, True,
                                precision=precision)

    def same(a, bb):
        return all(torch.equal(a[k], bb[k]) for k in ("feat", "depth", "opacity"))

    ref = run()
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(3)]
    outs = []
    for _ in range(4):
        for s in streams:
            with torch.cuda.stream(s):
                outs.append(run())
    torch.cuda.synchronize()
    assert all(same(o, ref) for o in outs)
    # a captured launch replayed on the current stream while eager launches run on another one
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        run()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        gout = run()
    for _ in range(3):
        for v in gout.values():
            if torch.is_tensor(v):
                v.fill_(float("nan"))
        torch.cuda.synchronize()
        graph.replay()
        with torch.cuda.stream(side):        # the capturing stream's own counter is not the graph's
            eager = run()
        torch.cuda.synchronize()
        assert same(gout, ref) and same(eager, ref)


@pytest.mark.parametrize("workload,hw", [("dtu", (64, 96)), ("nerf", (128, 160))])
@pytest.mark.parametrize("V", [2, 4])
def test_view_counts_two_and_four(workload, hw, V):
    """the kernels must cover V in {2, 3, 4} (training draws 2/3/4 views, fine-tuned evaluation uses 4).  Every MLP
    arithmetic of the fused render kernel and the cost-volume kernel against the oracle on identical inputs."""
    cfg, w, rig, data, mlp, feat_dim = _full_size_inputs(workload, V=V, seed=3, hw=hw)
    b = cfg.nerf.bundle_size
    H, W = w["H"], w["W"]
    truth = O.render_bundles(mlp, feat_dim, data["rgb"], data["feat"], data["vol"], data["depth_range"], data["vol_range"],
                             rig["src_exts"], rig["src_ints"], rig["tar_exts"], rig["tar_ints"], rig["near_far"], b,
                             cfg.nerf.max_num_samples, cfg.nerf.global_num_depth, cfg.nerf.max_mipmap_level, False, True)
    cam = ops.camera_block(rig["tar_exts"].to(DEV), rig["tar_ints"].to(DEV), rig["src_exts"].to(DEV), rig["src_ints"].to(DEV),
                           rig["near_far"].to(DEV), b, cfg.nerf.global_num_depth, False)
    src = ops.prepare_sources(data["feat"].to(DEV), data["rgb"].to(DEV), b, cfg.nerf.max_mipmap_level)
    vol_cl = ops.to_channels_last(data["vol"].to(DEV), 8)
    sl = ops.sample_bundles(data["depth_range"].to(DEV), data["vol_range"].to(DEV), cam, b, cfg.nerf.max_num_samples, False, True)
    assert torch.equal(sl.indices.cpu(), truth["indices"])
    for precision in (0, 1, 2):
        tol = 2e-3 if precision == 1 else 1e-4
        out = ops.render_fused(src, vol_cl, data["depth_range"].to(DEV), data["vol_range"].to(DEV), cam,
                               ops.pack_mlp(mlp, feat_dim, device=DEV), 1, V, H, W, b, cfg.nerf.max_num_samples, False, True,
                               precision=precision)
        assert _md(out["feat"], truth["bundle_feat"]) <= tol, precision
        assert _md(out["feat"][:, :3 * b * b], truth["bundle_feat"][:, :3 * b * b]) <= 1e-4, precision      # fine colours: fp32 in every variant
        assert _md(out["depth"], truth["bundle_depth"]) <= 1e-4 * (w["far"] - w["near"]), precision
        assert _md(out["opacity"], torch.ones_like(out["opacity"])) <= 1e-5
    # cost volume over V views (stage-1 geometry of the recipe: feature level of the bundle map)
    Hb, Wb = H // b, W // b
    proj = ops.homography_mats(rig["src_exts"].to(DEV), rig["src_ints"].to(DEV), rig["tar_exts"].to(DEV), rig["tar_ints"].to(DEV), 1.0 / b, 1.0 / b)
    feat_cl = ops.to_channels_last(data["feat"].flatten(0, 1).to(DEV)).unflatten(0, (1, V))
    var = ops.warp_variance(feat_cl, proj, data["depth_range"].to(DEV), 8, Hb, Wb, False)
    tv = O.warp_variance(data["feat"], proj.cpu(), O.depth_hypotheses(data["depth_range"], 8, False), False)
    assert _md(var, tv) <= 1e-4


@pytest.mark.parametrize("V", [2, 4])
def test_full_size_view_counts_on_the_dynamic_tile_path(V):
    """Two and four source views at the full DTU size: more tiles than tile slots, so the 2x2 tensor-core kernels (four
    tile slots per SM with two views, three with four) draw tiles from the atomic counter.  Checked against the fp32 SIMT
    kernel, which has no counter: a skipped or repeated tile would leave garbage / stale rows far outside the class."""
    cfg, w, rig, data, mlp, feat_dim = _full_size_inputs("dtu", V=V)
    b = cfg.nerf.bundle_size
    cam = ops.camera_block(rig["tar_exts"].to(DEV), rig["tar_ints"].to(DEV), rig["src_exts"].to(DEV), rig["src_ints"].to(DEV),
                           rig["near_far"].to(DEV), b, cfg.nerf.global_num_depth, False)
    src = ops.prepare_sources(data["feat"].to(DEV), data["rgb"].to(DEV), b, cfg.nerf.max_mipmap_level)
    vol_cl = ops.to_channels_last(data["vol"].to(DEV), 8)
    args = (src, vol_cl, data["depth_range"].to(DEV), data["vol_range"].to(DEV), cam, ops.pack_mlp(mlp, feat_dim, device=DEV),
            1, V, w["H"], w["W"], b, cfg.nerf.max_num_samples, False, True)
    ref = ops.render_fused(*args, precision=0)
    for _ in range(2):                                    # the second launch re-uses the stream's counter
        out = ops.render_fused(*args, precision=1)
        assert _md(out["feat"], ref["feat"]) <= 2e-3
        assert _md(out["feat"][:, :3 * b * b], ref["feat"][:, :3 * b * b]) <= 1e-4
        assert _md(out["depth"], ref["depth"]) <= 1e-4 * (w["far"] - w["near"])
        assert _md(out["opacity"], torch.ones_like(out["opacity"])) <= 1e-5


def test_full_size_warp_variance_against_oracle():
    w = WORKLOADS["dtu"]
    cfg = make_cfg(w["recipe"])
    H, W, V = w["H"], w["W"], 3
    rig = camera_rig(1, V, H, W, w["near"], w["far"], w["focal"], tilt=0.03)
    g = torch.Generator().manual_seed(2)
    for s in range(2):
        fs, vs = cfg.fpn.feat_scales[cfg.mvs.vol_levels[s]], cfg.mvs.vol_scales[s]
        C = cfg.fpn.feat_dims[cfg.mvs.vol_levels[s]]
        Hs, Ws, Ht, Wt = int(H * fs), int(W * fs), int(H * vs), int(W * vs)
        D = cfg.mvs.num_depth[s]
        # smooth features: judged below the fp32 noise floor of white-noise inputs
        coarse = torch.randn(V, C, Hs // 8, Ws // 8, generator=g)
        feat = torch.nn.functional.interpolate(coarse, size=(Hs, Ws), mode="bicubic", align_corners=True)[None]
        if s == 0:
            rng = rig["near_far"][..., None, None].contiguous()
        else:
            mid = w["near"] + (w["far"] - w["near"]) * (0.3 + 0.4 * torch.rand(1, 1, Ht, Wt, generator=g))
            rng = torch.cat((mid - 20.0, mid + 20.0), 1)
        proj = ops.homography_mats(rig["src_exts"].to(DEV), rig["src_ints"].to(DEV), rig["tar_exts"].to(DEV), rig["tar_ints"].to(DEV), fs, vs)
        feat_cl = ops.to_channels_last(feat.flatten(0, 1).to(DEV)).unflatten(0, (1, V))
        var = ops.warp_variance(feat_cl, proj, rng.to(DEV), D, Ht, Wt, cfg.mvs.inv_depth[s])
        dv = O.depth_hypotheses(rng, D, cfg.mvs.inv_depth[s]).expand(1, D, Ht, Wt)
        truth = O.warp_variance(feat, proj.cpu(), dv, cfg.mvs.inv_depth[s])
        assert _md(var, truth) <= 1e-4 * max(1.0, float(truth.abs().max()))


@pytest.mark.parametrize("workload", ["dtu", "nerf", "llff"])
def test_default_mlp_arithmetic_meets_fp32_class_end_to_end(workload):
    """The default MLP arithmetic (tcgen05, fp16 operands, fp32 accumulation) against the fp32 SIMT arithmetic through the
    whole Network.forward at BASELINE.json sizes: the north star's fp32-class tolerance (1e-4 absolute on rgb and depth)
    holds on the outputs, PSNR delta far below 0.01 dB."""
    from gdb_nerf_b200.synthetic import workload_batch
    w = WORKLOADS[workload]
    cfg = make_cfg(w["recipe"])
    torch.manual_seed(0)
    net = Network(cfg).to(DEV).eval()
    assert net.mlp_precision == 1
    batch = batch_to(workload_batch(workload, B=1), DEV)
    outs = {}
    with torch.no_grad():
        for prec in (0, 1):
            net.mlp_precision = prec
            outs[prec] = net(batch)[0]
    assert _md(outs[1]["rgb"], outs[0]["rgb"]) <= 1e-4
    assert _md(outs[1]["nerf_depth"], outs[0]["nerf_depth"]) <= 1e-4 * (w["far"] - w["near"])
    assert _md(outs[1]["mvs_depth"], outs[0]["mvs_depth"]) <= 1e-5 * (w["far"] - w["near"])     # same CNNs; cuDNN run-to-run noise only
    mse = float(((outs[1]["rgb"].double() - outs[0]["rgb"].double()) ** 2).mean())
    assert mse < 1e-9


def test_cuda_graph_forward_matches_eager():
    """the eval forward has no host synchronisation, so it captures into one CUDA graph; the replay on new
    inputs reproduces the eager forward (same kernels, same order)."""
    from gdb_nerf_b200.graphed import GraphedForward
    cfg = make_cfg("dtu_eval")
    torch.manual_seed(0)
    net = Network(cfg).to(DEV).eval()
    mk = lambda seed: batch_to(make_batch(2, 3, 64, 96, 425.0, 905.0, 180.0, seed=seed, images="smooth", tilt=0.03), DEV)
    runner = GraphedForward(net, mk(0))
    for seed in (1, 2):
        batch = mk(seed)
        with torch.no_grad():
            want = net(batch)[0]
        got = runner(batch)[0]
        torch.cuda.synchronize()
        for k in ("rgb", "nerf_depth", "mvs_depth", "opacity"):
            scale = 480.0 if "depth" in k else 1.0
            assert _md(got[k], want[k]) <= 1e-5 * scale, k


@pytest.mark.parametrize("max_samples,adaptive", [(1, False), (2, True), (4, True), (5, False), (8, True), (32, False)])
def test_sample_counts_and_ragged_tiles(max_samples, adaptive):
    """Every legal samples-per-bundle limit (the C ABI takes 1..32; the shipped recipes use 3 and 6), a bundle map whose
    size is not a multiple of the kernel's tile (partial last tile), two target views with different cameras in one launch,
    fixed and adaptive sampling: every MLP arithmetic against the oracle, indices bit-exact."""
    cfg, w, rig, data, mlp, feat_dim = _full_size_inputs("dtu", V=3, seed=11, hw=(48, 112))
    b = cfg.nerf.bundle_size
    H, W = w["H"], w["W"]
    Hb, Wb = H // b, W // b
    B = 2
    rig = camera_rig(B, 3, H, W, w["near"], w["far"], w["focal"], tilt=0.03)
    g = torch.Generator().manual_seed(5)
    data = {k: torch.cat((v, v.flip(-1) * 0.9 + 0.05 * torch.rand(v.shape, generator=g)), 0) for k, v in data.items()}
    min_iv = (w["far"] - w["near"]) / cfg.nerf.global_num_depth
    mid = w["near"] + (w["far"] - w["near"]) * (0.3 + 0.4 * torch.rand(B, 1, Hb, Wb, generator=g))
    half = torch.rand(B, 1, Hb, Wb, generator=g) * (0.55 * max_samples * min_iv)
    data["depth_range"] = torch.cat((mid - half, mid + half), 1)
    data["vol_range"] = torch.cat((mid - 2.5 * min_iv, mid + 2.5 * min_iv), 1)
    truth = O.render_bundles(mlp, feat_dim, data["rgb"], data["feat"], data["vol"], data["depth_range"], data["vol_range"],
                             rig["src_exts"], rig["src_ints"], rig["tar_exts"], rig["tar_ints"], rig["near_far"], b,
                             max_samples, cfg.nerf.global_num_depth, cfg.nerf.max_mipmap_level, False, adaptive)
    cam = ops.camera_block(rig["tar_exts"].to(DEV), rig["tar_ints"].to(DEV), rig["src_exts"].to(DEV), rig["src_ints"].to(DEV),
                           rig["near_far"].to(DEV), b, cfg.nerf.global_num_depth, False)
    src = ops.prepare_sources(data["feat"].to(DEV), data["rgb"].to(DEV), b, cfg.nerf.max_mipmap_level)
    vol_cl = ops.to_channels_last(data["vol"].to(DEV), 8)
    sl = ops.sample_bundles(data["depth_range"].to(DEV), data["vol_range"].to(DEV), cam, b, max_samples, False, adaptive)
    assert torch.equal(sl.indices.cpu(), truth["indices"])
    for precision in (0, 1, 2):
        out = ops.render_fused(src, vol_cl, data["depth_range"].to(DEV), data["vol_range"].to(DEV), cam,
                               ops.pack_mlp(mlp, feat_dim, device=DEV), B, 3, H, W, b, max_samples, False, adaptive, precision=precision)
        assert _md(out["feat"], truth["bundle_feat"]) <= (2e-3 if precision == 1 else 1e-4), precision
        assert _md(out["depth"], truth["bundle_depth"]) <= 1e-4 * (w["far"] - w["near"]), precision
        assert _md(out["opacity"], torch.ones_like(out["opacity"])) <= 1e-5


def test_render_sweep_matches_direct_forward():
    """gdb_nerf_b200.pipeline.render_sweep (run.py:53-66 mirror): pinned double-buffered uploads on a side stream, this rank's
    share of the views, results identical to calling the network on each batch."""
    from gdb_nerf_b200.pipeline import render_sweep
    cfg = make_cfg("dtu_eval")
    torch.manual_seed(0)
    net = Network(cfg).to(DEV).eval()
    batches = [make_batch(1, 3, 64, 96, 425.0, 905.0, 180.0, seed=s, images="smooth", tilt=0.03) for s in range(5)]
    for b in batches:
        b["meta"] = {"scene": "synthetic"}
    want = []
    with torch.no_grad():
        for b in batches:
            want.append(net(batch_to(b, DEV))[0]["rgb"].cpu())
    got = dict(render_sweep(net, batches, device=DEV))
    assert sorted(got) == [0, 1, 2, 3, 4]
    for i in range(5):
        assert _md(got[i]["rgb"], want[i]) <= 1e-5
    part = dict(render_sweep(net, batches, rank=1, world=2, device=DEV))
    assert sorted(part) == [1, 3] and _md(part[3]["rgb"], want[3]) <= 1e-5
    # several views per forward (the reference's loop with its independent iterations batched), zero-copy results: a yielded
    # tensor is valid until two further calls have been consumed, so it is compared as it arrives
    seen = []
    for idx, res in render_sweep(net, batches, device=DEV, views_per_call=2, copy=False, keys=("rgb", "nerf_depth")):
        assert res["rgb"].shape == want[idx].shape and res["nerf_depth"].shape[0] == 1
        assert _md(res["rgb"], want[idx]) <= 1e-4, idx          # cuDNN may pick another algorithm for the larger batch
        seen.append(idx)
    assert seen == [0, 1, 2, 3, 4]


@pytest.mark.parametrize("recipe,hw", [("dtu_eval", (64, 96)), ("nerf_eval_4x4", (128, 160))])
def test_depth_folded_cost_regularisation_matches_3d(recipe, hw):
    """Stage-1 U-Net run depth-folded on 2-D convolutions (K1 layout 2, block-Toeplitz weights, K3 vol_layout 1) against the same
    network on cuDNN's 3-D convolutions: same products, only the summation order differs (true-fp32 convolutions here)."""
    cfg = make_cfg(recipe)
    torch.manual_seed(0)
    net = Network(cfg).to(DEV).eval()
    H, W = hw
    batch = batch_to(make_batch(2, 3, H, W, 425.0, 905.0, 180.0 * H / 64, seed=4, images="smooth", tilt=0.03), DEV)
    outs = []
    with torch.no_grad():
        for fold in (True, False):
            net.depth_net.fold_depth = fold
            ret, mvs, _ = net(batch)
            outs.append((ret, mvs))
    (a, ma), (b_, mb) = outs
    assert _md(a["rgb"], b_["rgb"]) <= 2e-5
    assert _md(a["nerf_depth"], b_["nerf_depth"]) <= 2e-5 * 480.0
    for x, y in zip(ma, mb):
        assert _md(x, y) <= 2e-5 * 480.0


@pytest.mark.parametrize("levels", [1, 2, 3])
@pytest.mark.parametrize("hw", [(24, 40), (32, 48), (20, 36), (64, 64)])
def test_width_folded_fpn_matches_modules(hw, levels):
    """feature_net_fused (BN folded, width-folded 8/16-channel layers when W % 4 == 0, plain layers otherwise) equals the
    plain module sequence of feature_net.py:40-64 in fp32."""
    from gdb_nerf_b200.cnn import FeatureNet, feature_net_fused
    torch.manual_seed(5)
    net = FeatureNet().to(DEV).eval()
    for m in net.modules():                                       # non-trivial batch-norm statistics
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.2)
    x = torch.rand(3, 3, *hw, device=DEV)
    with torch.no_grad():
        want = net(x, levels=levels)
        got = feature_net_fused(net, x, levels=levels)
    assert len(got) == len(want) == levels
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert _md(a, b) <= 2e-5 * max(1.0, float(b.abs().max()))


@pytest.mark.parametrize("small,dhw", [(True, (8, 8, 12)), (True, (16, 8, 16)), (True, (8, 12, 4)), (False, (8, 8, 16)),
                                       (False, (16, 8, 8)), (False, (8, 16, 24))])
def test_width_folded_cost_regularisation_matches_modules(small, dhw):
    """cost_reg_fused (width-folded levels when W % 4 == 0) equals CostRegNet(Small).forward (cost_reg_net.py:40-117): feature
    volume and the probability head before its soft-max."""
    from gdb_nerf_b200.cnn import CostRegNet, CostRegNetSmall, cost_reg_fused
    torch.manual_seed(6)
    net = (CostRegNetSmall if small else CostRegNet)(32, 8, 8).to(DEV).eval()
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm3d):
            m.running_mean.normal_(0, 0.2)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.2)
    x = torch.randn(2, 32, *dhw, device=DEV).contiguous(memory_format=torch.channels_last_3d)
    with torch.no_grad():
        feat, prob = net(x)
        vol, logits = cost_reg_fused(net, x, want_volume=True)
    assert _md(vol.permute(0, 4, 1, 2, 3), feat) <= 2e-5 * max(1.0, float(feat.abs().max()))
    assert _md(torch.softmax(logits, 1), prob) <= 2e-5


# ------------------------------------------------------------------ the BENCHED math mode against the oracle, full size
def _psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 99.0 if mse == 0 else -10.0 * np.log10(mse)


@pytest.mark.parametrize("workload", ["dtu", "nerf", "llff"])
def test_benched_math_mode_against_oracle(workload):
    """bench.py's configuration - TF32 cuDNN convolutions (PyTorch's default, also what the reference's own CUDA forward
    runs), cudnn.benchmark, fp16-operand MLP (precision 1), 8-bit source images converted on the device - at the benchmark's
    full image size against the CPU oracle's fp32 forward on the same batch and weights.  The measured errors are written
    to gpurun_out/ (committed as profiles/r02_benched_mode_parity.json, quoted in the bench line's config); the same run
    in the fp32 class (TF32 off, precision 2) is reported next to it."""
    import json
    import os
    from gdb_nerf_b200.synthetic import with_uint8_images, workload_batch
    w = WORKLOADS[workload]
    cfg = make_cfg(w["recipe"])
    torch.manual_seed(0)
    net = Network(cfg).eval()
    batch = workload_batch(workload, B=1, V=3, seed=0, images="noise8")
    with torch.no_grad():
        truth, _ = O.network_forward(net, batch, cfg)
    net = net.to(DEV)
    scale = w["far"] - w["near"]
    report = {}
    old = torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark
    try:
        for mode, tf32, prec in (("benched (TF32 conv, fp16-operand MLP)", True, 1), ("fp32 class (fp32 conv, split-fp16 MLP)", False, 2)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cudnn.benchmark = tf32
            net.mlp_precision = prec
            with torch.no_grad():
                ret, _, _ = net(batch_to(with_uint8_images(batch), DEV))
            torch.cuda.synchronize()
            rgb, dep = ret["rgb"].cpu(), ret["nerf_depth"].cpu()
            report[mode] = {
                "rgb_max_abs": _md(rgb, truth["rgb"]), "rgb_mean_abs": float((rgb - truth["rgb"]).abs().mean()),
                "rgb_psnr_vs_oracle_db": _psnr(rgb, truth["rgb"]),
                "nerf_depth_max_abs_normalised": _md(dep, truth["nerf_depth"]) / scale,
                "nerf_depth_mean_abs_normalised": float((dep - truth["nerf_depth"]).abs().mean()) / scale,
                "mvs_depth_max_abs_normalised": _md(ret["mvs_depth"], truth["mvs_depth"]) / scale,
            }
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = old
        net.mlp_precision = 1
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"benched_mode_parity_{workload}.json"), "w") as fh:
        json.dump({workload: report, "what": f"{workload} {w['H']}x{w['W']}, 1 target view, 3 source views, 8-bit white-noise images, "
                                             "random-init weights (seed 0): Network.forward on the GPU vs oracle.network_forward (CPU fp32)"}, fh, indent=1)
    print(json.dumps(report))
    fp32 = report["fp32 class (fp32 conv, split-fp16 MLP)"]
    bench = report["benched (TF32 conv, fp16-operand MLP)"]
    # fp32 class: cuDNN vs MKL-DNN summation order through ~40 convolution layers on white-noise images
    assert fp32["rgb_max_abs"] <= 2e-3 and fp32["nerf_depth_max_abs_normalised"] <= 2e-3, fp32
    # benched mode: TF32 convolutions (10-bit operand mantissa) dominate; the image must stay visually identical
    assert bench["rgb_psnr_vs_oracle_db"] >= 45.0 and bench["nerf_depth_mean_abs_normalised"] <= 5e-3, bench


def test_network_tile_split_plumbing(monkeypatch):
    """Image-tile split through Network.forward: each emulated rank renders its bundle rows (the fused kernel launch is
    restricted to them) and receives the other rows from the gather - here fed from a one-piece run, whose row tiles are
    bit-identical to what the other ranks' launches write (tests/test_kernels_gpu.py::test_render_fused_row_range_...).
    The final image equals the one-piece image bit for bit."""
    import gdb_nerf_b200.sharding as sharding
    cfg = make_cfg("dtu_eval")
    torch.manual_seed(0)
    net = Network(cfg).to(DEV).eval()
    batch = batch_to(make_batch(1, 3, 128, 160, 425.0, 905.0, 360.0, seed=2, images="smooth"), DEV)
    captured = {}
    real = ops.render_fused

    def spy(*a, **k):
        out = real(*a, **k)
        captured["out"] = {n: t.clone() for n, t in out.items()}
        captured["rows"] = k.get("rows")
        return out

    monkeypatch.setattr(ops, "render_fused", spy)
    with torch.no_grad():
        full, _, _ = net(batch)
    whole = captured["out"]
    assert captured["rows"] is None
    world = 4
    seen = []

    def fake_gather(tensors, n_rows, w, group=None, align=8):
        tile = sharding.shard_rows(n_rows, net._tile[0], w, align)
        seen.append((tile.start, tile.stop))
        for t, name in zip(tensors, ("fine", "dec_in", "depth", "opacity")):
            mine = t[:, tile.start: tile.stop].clone()
            assert torch.equal(mine, whole[name][:, tile.start: tile.stop]), name
            t.copy_(whole[name])
            t[:, tile.start: tile.stop] = mine

    monkeypatch.setattr(sharding, "gather_row_tiles", fake_gather)
    for r in range(world):
        net.set_tile_split(r, world)
        with torch.no_grad():
            ret, _, _ = net(batch)
        assert captured["rows"] == seen[-1]
        for k in full:
            assert torch.equal(ret[k], full[k]), (r, k)
    net.set_tile_split(0, 1)
    assert seen == [(0, 16), (16, 32), (32, 48), (48, 64)]
