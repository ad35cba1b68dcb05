"""Gradient parity of the backward kernels (training configuration, SURVEY.md 8b "Autograd") against
torch.autograd run on the float64 CPU oracle, through the C ABI.  Tolerances are relative to the largest
gradient entry of each tensor (fp32 kernels + atomics vs an fp64 reference): 2e-3."""
import pytest
import torch

from conftest import CASE_SPECS, load_golden
from gdb_nerf_b200 import autograd as AG
from gdb_nerf_b200 import ops
from gdb_nerf_b200.config import make_cfg
from gdb_nerf_b200.mlp_pack import unpack_grad
from oracle import gdb_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
DD = torch.float64
RTOL = 2e-3


def _md(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def _rel(got, want, floor=1e-30):
    """max |got - want| relative to the largest reference entry (``floor``: smallest scale that counts, for
    gradients that vanish analytically, e.g. the bias in front of a soft-max over views)."""
    want = want.detach().double().cpu()
    got = got.detach().double().cpu()
    return float((got - want).abs().max()) / max(float(want.abs().max()), floor)


@pytest.mark.parametrize("case", ["train_b2", "dtu_b2", "nerf_b4"])
@pytest.mark.parametrize("stage", [0, 1])
def test_warp_variance_backward(case, stage):
    g = load_golden(case)
    cfg = make_cfg(CASE_SPECS[case]["recipe"])
    s = stage
    feat = g.t(f"s{s}_src_feat")
    rng = g.t(f"s{s}_range_in")
    ref = g.t(f"s{s}_variance")
    D, Ht, Wt = ref.shape[2:]
    inv = cfg.mvs.inv_depth[s]
    proj = O.homography_matrices(g.t("in_src_exts"), g.t(f"s{s}_src_ints"), g.t("in_tar_exts"), g.t(f"s{s}_tar_ints"))
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(s))
    # oracle, float64
    f64 = feat.to(DD).requires_grad_(True)
    r64 = rng.to(DD).requires_grad_(True)
    var64 = O.warp_variance(f64, proj.to(DD), O.depth_hypotheses(r64, D, inv).expand(-1, -1, Ht, Wt), inv)
    (var64 * w.to(DD)).sum().backward()
    # kernels
    fcl = feat.to(DEV).permute(0, 1, 3, 4, 2).contiguous().requires_grad_(True)
    rd = rng.to(DEV).requires_grad_(True)
    var = AG.WarpVariance.apply(fcl, proj.to(DEV), rd, D, Ht, Wt, inv)
    assert _rel(var, var64) <= 1e-4
    (var * w.to(DEV)).sum().backward()
    assert _rel(fcl.grad.permute(0, 1, 4, 2, 3), f64.grad) <= RTOL
    if rng.shape[-1] != 1:
        assert _rel(rd.grad, r64.grad) <= RTOL
    else:
        assert rd.grad is None                       # near/far of the batch are inputs


@pytest.mark.parametrize("case", ["train_b2", "dtu_b2", "nerf_b4"])
@pytest.mark.parametrize("stage", [0, 1])
def test_depth_range_backward(case, stage):
    g = load_golden(case)
    cfg = make_cfg(CASE_SPECS[case]["recipe"])
    s = stage
    rng, prob = g.t(f"s{s}_range_in"), g.t(f"s{s}_prob")
    B, D, h, w = prob.shape
    inv = cfg.mvs.inv_depth[s]
    gen = torch.Generator().manual_seed(7 + s)
    wd, wc, wv = torch.randn(B, 1, h, w, generator=gen), torch.randn(B, 2, h, w, generator=gen), torch.randn(B, 2, h, w, generator=gen)
    for ci_scale in (cfg.mvs.ci_scales[s], 0.05):     # the small scale keeps the interval strictly inside the clamps
        r64 = rng.to(DD).requires_grad_(True)
        p64 = prob.to(DD).requires_grad_(True)
        dv = O.depth_hypotheses(r64, D, inv).expand(-1, -1, h, w)
        d64, c64 = O.depth_interval(dv, p64, ci_scale, inv)
        v64 = dv[:, [0, -1]]
        ((d64 * wd.to(DD)).sum() + (c64 * wc.to(DD)).sum() + (v64 * wv.to(DD)).sum()).backward()
        rd = rng.to(DEV).requires_grad_(True)
        pd = prob.to(DEV).requires_grad_(True)
        dep, ci, vol = AG.DepthRange.apply(rd, pd, ci_scale, inv)
        ((dep * wd.to(DEV)).sum() + (ci * wc.to(DEV)).sum() + (vol * wv.to(DEV)).sum()).backward()
        assert _rel(pd.grad, p64.grad) <= RTOL
        if rng.shape[-1] != 1:
            assert _rel(rd.grad, r64.grad) <= RTOL


def _render_inputs(g, cfg):
    tex = g.t("tex_nchw")
    fd = tex.shape[2] - 3
    return dict(fd=fd, feat=tex[:, :, :fd].contiguous(), rgb=g.t("in_rgb"), vol=g.t("feat_volume"))


@pytest.mark.parametrize("case,prefix", [("train_b2", ""), ("train_b2", "inj_"), ("dtu_b2", "inj_"), ("nerf_b4", ""), ("nerf_b4", "inj_")])
def test_render_fused_backward(case, prefix):
    """dL/d(features, feature volume, depth interval, volume range, MLP parameters) of the fused render."""
    g = load_golden(case)
    spec = CASE_SPECS[case]
    cfg = make_cfg(spec["recipe"])
    b = cfg.nerf.bundle_size
    adaptive = True if prefix else cfg.nerf.is_adaptive
    inv = cfg.mvs.inv_depth[-1]
    x = _render_inputs(g, cfg)
    fd = x["fd"]
    dr, vr = g.t(prefix + "depth_range"), g.t(prefix + "vol_range")
    B, V = x["feat"].shape[:2]
    H, W = spec["H"], spec["W"]
    Hb, Wb = H // b, W // b
    CT = 3 * b * b + fd + 3 + 8
    gen = torch.Generator().manual_seed(21)
    wf = torch.randn(B, CT, Hb, Wb, generator=gen)
    wd = torch.randn(B, Hb, Wb, generator=gen) * (1.0 / (spec["far"] - spec["near"]))
    wo = torch.randn(B, Hb, Wb, generator=gen)

    # ---- oracle in float64 with autograd
    f64 = x["feat"].to(DD).requires_grad_(True)
    v64 = x["vol"].to(DD).requires_grad_(True)
    dr64 = dr.to(DD).requires_grad_(True)
    vr64 = vr.to(DD).requires_grad_(True)
    mlp64 = {k: v.to(DD).requires_grad_(True) for k, v in g.mlp().items()}
    out = O.render_bundles(mlp64, fd, x["rgb"].to(DD), f64, v64, dr64, vr64, g.t("in_src_exts").to(DD), g.t("in_src_ints").to(DD),
                           g.t("in_tar_exts").to(DD), g.t("in_tar_ints").to(DD), g.t("in_near_far").to(DD), b,
                           cfg.nerf.max_num_samples, cfg.nerf.global_num_depth, cfg.nerf.max_mipmap_level, inv, adaptive)
    loss = (out["bundle_feat"] * wf.to(DD)).sum() + (out["bundle_depth"] * wd.to(DD)).sum() + (out["bundle_opacity"] * wo.to(DD)).sum()
    loss.backward()

    # ---- kernels
    fdv = x["feat"].to(DEV).requires_grad_(True)
    vdv = x["vol"].to(DEV).requires_grad_(True)
    drd = dr.to(DEV).requires_grad_(True)
    vrd = vr.to(DEV).requires_grad_(True)
    params = {k: v.to(DEV).requires_grad_(True) for k, v in g.mlp().items()}
    mlp = ops.pack_mlp(params, fd, detach=False)
    cam = ops.camera_block(g.t("in_tar_exts").to(DEV), g.t("in_tar_ints").to(DEV), g.t("in_src_exts").to(DEV), g.t("in_src_ints").to(DEV),
                           g.t("in_near_far").to(DEV), b, cfg.nerf.global_num_depth, inv)
    feat, depth, opac = AG.render_fused_train(fdv, x["rgb"].to(DEV), vdv, drd, vrd, cam, mlp, b, cfg.nerf.max_num_samples,
                                              cfg.nerf.max_mipmap_level, inv, adaptive)
    assert _rel(feat, out["bundle_feat"]) <= 1e-4
    ((feat * wf.to(DEV)).sum() + (depth * wd.to(DEV)).sum() + (opac * wo.to(DEV)).sum()).backward()

    assert _rel(fdv.grad, f64.grad) <= RTOL, "feature-map gradient"
    assert _rel(vdv.grad, v64.grad) <= RTOL, "feature-volume gradient"
    assert _rel(drd.grad, dr64.grad) <= RTOL, "depth-interval gradient"
    assert _rel(vrd.grad, vr64.grad) <= RTOL, "volume-range gradient"
    scale = max(float(v.grad.abs().max()) for v in mlp64.values())
    for k in params:
        assert _rel(params[k].grad, mlp64[k].grad, floor=1e-3 * scale) <= RTOL, k


def test_unpack_grad_roundtrip():
    g = load_golden("dtu_b2")
    p = g.mlp()
    flat = ops.pack_mlp(p, 16)
    back = unpack_grad(flat, 16)
    for k in p:
        assert torch.equal(back[k].reshape(p[k].shape), p[k])


# ------------------------------------------------------------------ whole training step vs the reference under autograd
def _train_net(case="train_b2"):
    from gdb_nerf_b200.network import Network
    from gdb_nerf_b200.synthetic import batch_to, make_batch, synth_state_dict
    spec = CASE_SPECS[case]
    net = Network(make_cfg(spec["recipe"]))
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(synth_state_dict(shapes, seed=1), strict=True)
    net = net.to(DEV).train()
    batch = make_batch(spec["B"], spec["V"], spec["H"], spec["W"], spec["near"], spec["far"], spec["focal"], seed=3,
                       images=spec["images"], tilt=spec["tilt"])
    return net, batch_to(batch, DEV), spec


def test_training_step_matches_reference_gradients():
    """fwd + bwd of the whole Network in train mode (BASELINE.json configs[4] code path at test size): loss, outputs and
    parameter gradients against the UNMODIFIED reference under torch.autograd (tests/golden/train_b2_grad.npz)."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        gg = load_golden("train_b2_grad")
        net, batch, spec = _train_net()
        ret, mvs_depths, blend = net(batch)
        assert len(blend) == 1 and blend[0].shape == gg.t("blend_rgb_0").shape
        loss = ret["rgb"].square().mean() + sum(b.square().mean() for b in blend)
        loss.backward()
        assert _rel(ret["rgb"], gg.t("ret_rgb")) <= 2e-3
        assert _rel(blend[0], gg.t("blend_rgb_0")) <= 2e-3
        assert abs(float(loss) - float(gg.np("loss"))) <= 2e-3 * float(gg.np("loss"))
        params = dict(net.named_parameters())
        checked = 0
        for key in gg._z.files:
            if not key.startswith("grad_"):
                continue
            name = key[5:]
            want = gg.t(key)
            got = params[name].grad
            if float(want.abs().max()) == 0.0:
                assert got is None or float(got.abs().max()) == 0.0, name
                continue
            assert got is not None, f"{name} got no gradient"
            assert _rel(got, want, floor=1e-6) <= 2e-2, name       # cuDNN vs MKL-DNN batch-norm / conv backward noise included
            checked += 1
        assert checked >= 30
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_training_step_updates_every_trainable_parameter():
    net, batch, _ = _train_net()
    ret, _, blend = net(batch)
    (ret["rgb"].square().mean() + blend[0].square().mean()).backward()
    missing = [k for k, p in net.named_parameters() if p.grad is None]
    # the reference's own autograd leaves exactly the unused full-resolution FPN branch without gradient
    assert all(k.startswith(("feature_net.inner2", "feature_net.out2")) for k in missing), missing
    for k, p in net.named_parameters():
        if p.grad is not None:
            assert torch.isfinite(p.grad).all(), k


# ------------------------------------------------------------------ row a14: coarse render kernels vs the PyTorch restatement
@pytest.mark.parametrize("inv_depth,V", [(True, 3), (False, 2), (True, 4)])
def test_coarse_render_forward_and_backward(inv_depth, V):
    """gdb_coarse_render_fwd/bwd against coarse.coarse_render (depth_net.py:49-116 restated with PyTorch operators,
    itself pinned end-to-end by the reference's blend_rgbs) evaluated in float64 on the CPU."""
    from oracle.coarse_oracle import coarse_render
    from gdb_nerf_b200.nerf import CoarseNeRF
    from gdb_nerf_b200.synthetic import camera_rig, smooth_images
    B, H, W, Cf, D, S = 2, 32, 40, 32, 16, 8
    Hs, Ws, Hi, Wi = H // 4, W // 4, H // 8, W // 8
    near, far = 2.5, 5.5
    gen = torch.Generator().manual_seed(5)
    rig = camera_rig(B, V, H, W, near, far, 44.0, tilt=0.05)
    images = smooth_images(B, V, H, W, seed=2)
    feats = torch.randn(B, V, Cf, Hs, Ws, generator=gen) * 0.5
    volume = torch.randn(B, 8, D, Hi, Wi, generator=gen) * 0.5
    mid = near + (far - near) * (0.3 + 0.4 * torch.rand(B, 1, Hi, Wi, generator=gen))
    half = 0.1 + 0.3 * torch.rand(B, 1, Hi, Wi, generator=gen)
    ray_range = torch.cat((mid - half, mid + half), 1)
    vol_range = torch.cat((torch.full_like(mid, near), torch.full_like(mid, far)), 1)
    if inv_depth:        # a disparity-spaced volume lists its hypotheses far -> near in depth? no: first = near, last = far (depth units)
        pass
    src_ints_s = rig["src_ints"].clone(); src_ints_s[..., :2, :] *= 0.25
    tar_ints_s = rig["tar_ints"].clone(); tar_ints_s[:, :2, :] *= 0.125
    torch.manual_seed(3)
    nerf = CoarseNeRF(64, 8, Cf, True)
    wout = torch.randn(B, 3, Hi, Wi, generator=gen)

    # ---- reference: PyTorch operators, float64, CPU
    n64 = CoarseNeRF(64, 8, Cf, True).double()
    n64.load_state_dict({k: v.double() for k, v in nerf.state_dict().items()})
    f64 = feats.double().requires_grad_(True)
    v64 = volume.double().requires_grad_(True)
    rr64 = ray_range.double().requires_grad_(True)
    vr64 = vol_range.double().requires_grad_(True)
    want = coarse_render(n64, v64, f64, images.double(), 0.25, rig["src_exts"].double(), src_ints_s.double(), rig["tar_exts"].double(),
                         tar_ints_s.double(), rr64, vr64, S, inv_depth)
    (want * wout.double()).sum().backward()

    # ---- kernels
    nd = CoarseNeRF(64, 8, Cf, True).to(DEV)
    nd.load_state_dict(nerf.state_dict())
    fd = feats.to(DEV).requires_grad_(True)
    vd = volume.to(DEV).requires_grad_(True)
    rrd = ray_range.to(DEV).requires_grad_(True)
    vrd = vol_range.to(DEV).requires_grad_(True)
    got = AG.coarse_render_train(nd, vd, fd, images.to(DEV), rig["src_exts"].to(DEV), src_ints_s.to(DEV), rig["tar_exts"].to(DEV),
                                 tar_ints_s.to(DEV), rig["near_far"].to(DEV), rrd, vrd, S, inv_depth)
    assert got.shape == want.shape
    assert float((got.detach().cpu().double() - want.detach()).abs().max()) <= 1e-4
    (got * wout.to(DEV)).sum().backward()
    assert _rel(fd.grad, f64.grad) <= RTOL, "feature gradient"
    assert _rel(vd.grad, v64.grad) <= RTOL, "volume gradient"
    assert _rel(rrd.grad, rr64.grad) <= RTOL, "ray-range gradient"
    assert _rel(vrd.grad, vr64.grad) <= RTOL, "volume-range gradient"
    ref_p = dict(n64.named_parameters())
    scale = max(float(p.grad.abs().max()) for p in ref_p.values())
    for k, p in nd.named_parameters():
        assert _rel(p.grad, ref_p[k].grad, floor=1e-3 * scale) <= RTOL, k


def test_training_step_as_cuda_graph_matches_eager():
    """The whole training step (forward, loss, backward, clip, Adam; trainer.py:44-66) captured as one CUDA graph reproduces the
    eager step from the same state on new batches (atomics in the weight-gradient accumulation: agreement to fp32 noise)."""
    import copy
    from gdb_nerf_b200.config import make_cfg
    from gdb_nerf_b200.graphed import GraphedTrainStep
    from gdb_nerf_b200.network import Network
    from gdb_nerf_b200.synthetic import batch_to, make_batch
    cfg = make_cfg("dtu_pretrain")
    torch.manual_seed(0)
    net_a = Network(cfg).to("cuda").train()
    net_b = copy.deepcopy(net_a)
    loss_fn = lambda out: out[0]["rgb"].square().mean() + sum(b.square().mean() for b in out[2])
    mk = lambda s: batch_to(make_batch(1, 3, 64, 64, 425.0, 905.0, 1446.0 * 64 / 512.0, seed=s, images="smooth", tilt=0.03), "cuda")
    pa = [p for p in net_a.parameters() if p.requires_grad]
    pb = [p for p in net_b.parameters() if p.requires_grad]
    opt_a = torch.optim.Adam(pa, lr=5e-4)
    opt_b = torch.optim.Adam(pb, lr=5e-4, capturable=True)

    def eager(batch):
        opt_a.zero_grad(set_to_none=True)
        loss = loss_fn(net_a(batch))
        loss.backward()
        torch.nn.utils.clip_grad_value_(pa, 40)
        opt_a.step()
        return float(loss.detach())

    for _ in range(3):                 # the stepper's three warm-up steps on its example batch (capturing executes nothing)
        eager(mk(0))
    stepper = GraphedTrainStep(net_b, opt_b, mk(0), loss_fn, pb)
    for s in (1, 2):
        la, lb = eager(mk(s)), float(stepper(mk(s)))
        assert abs(la - lb) <= 2e-3 * abs(la), (s, la, lb)


# ------------------------------------------------------------------ flat-buffer clip + Adam (trainer.py:63-65, optimizer.py:13-29)
def test_flat_adam_matches_torch_adam_with_value_clipping():
    """gdb_adam_clip_step on one flat buffer against clip_grad_value_(40) + torch.optim.Adam on the same parameters and
    gradients over several steps, including clipped gradients, weight decay and a parameter that never receives a gradient."""
    from gdb_nerf_b200.optim import FlatAdam
    g = torch.Generator().manual_seed(5)
    shapes = [(64, 24), (64,), (8, 16, 3, 3), (1,), (7, 5)]
    for wd in (0.0, 0.01):
        ref = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
        mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
        opt_ref = torch.optim.Adam([{"params": [p], "lr": 5e-4, "weight_decay": wd, "eps": 1e-8} for p in ref], 5e-4, weight_decay=wd, eps=1e-8)
        opt = FlatAdam(mine, lr=5e-4, eps=1e-8, weight_decay=wd, clip_value=40.0)
        for p, q in zip(ref, mine):
            assert torch.equal(p.data, q.data)
        for step in range(6):
            opt_ref.zero_grad()
            opt.zero_grad()
            for i, (p, q) in enumerate(zip(ref, mine)):
                if i == len(shapes) - 1:
                    continue                                              # never gets a gradient
                grad = torch.randn(p.shape, generator=g).to(DEV) * (100.0 if step % 2 == 0 else 1e-3)   # clipped / tiny
                p.grad = grad.clone()
                q.grad.add_(grad)                                         # autograd accumulates into the flat view
            torch.nn.utils.clip_grad_value_(ref, 40)
            opt_ref.step()
            opt.step()
        torch.cuda.synchronize()
        assert float(opt.state[0]) == 6.0
        for p, q in zip(ref[:-1], mine[:-1]):
            assert _md(p, q) <= 2e-6 * (1.0 + float(p.abs().max())), (wd, tuple(p.shape), _md(p, q))
        if wd == 0.0:      # the reference's setting (dtu_pretrain.yaml:59): a parameter without gradient stays untouched, as in torch
            assert torch.equal(mine[-1].data, ref[-1].data)
        else:              # documented difference (optim.py): torch skips it, the flat update applies the weight decay to it
            assert _md(mine[-1], ref[-1]) <= 6 * 5e-4 * 1.01
        assert all(q.data.data_ptr() >= opt.param_flat.data_ptr() for q in mine)


def test_flat_adam_training_step_is_graph_capturable():
    """The whole training step (forward, loss, backward into the flat gradient, fused clip + Adam) replays as one CUDA graph
    and produces the same parameters as the same steps launched eagerly."""
    import copy
    from gdb_nerf_b200.graphed import GraphedTrainStep
    from gdb_nerf_b200.optim import FlatAdam
    net_a, batch, _ = _train_net()
    net_b = copy.deepcopy(net_a)

    def loss_fn(out):
        return out[0]["rgb"].square().mean() + sum(b.square().mean() for b in out[2])

    opt_b = FlatAdam(net_b.parameters(), lr=5e-4)
    stepper = GraphedTrainStep(net_b, opt_b, batch, loss_fn, opt_b.params, warmup=2)      # 2 warm-up steps + capture (no update)
    for _ in range(3):
        loss_b = stepper(batch)
    torch.cuda.synchronize()
    opt_a = FlatAdam(net_a.parameters(), lr=5e-4)
    before = opt_a.param_flat.clone()
    for _ in range(5):
        opt_a.zero_grad()
        loss_a = loss_fn(net_a(batch))
        loss_a.backward()
        opt_a.step()
    torch.cuda.synchronize()
    assert float(opt_a.state[0]) == float(opt_b.state[0]) == 5.0
    assert abs(float(loss_a.detach()) - float(loss_b)) <= 1e-3 * abs(float(loss_a.detach()))
    # atomics in the backward kernels are not bit-stable and Adam turns a gradient of either sign into a step of size lr, so
    # single parameters with a gradient near zero may walk apart by up to 2 * steps * lr; what must agree is the bulk of the update
    diff = (opt_a.param_flat - opt_b.param_flat).abs()
    moved = (opt_a.param_flat - before).abs()
    assert float(diff.mean()) <= 0.1 * float(moved.mean()), (float(diff.mean()), float(moved.mean()))
    assert float(diff.max()) <= 2 * 5 * 5e-4 * 1.001
