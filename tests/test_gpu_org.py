"""GPU (-m gpu): the VunetOrg family (models/vunets.py:18-106, 151-219, 600-783) on the CUDA path — module API ->
torch ops -> C ABI — against the golden vectors generated from the UNMODIFIED reference class
(oracle/gen_golden_org.py) and against the CPU oracle at the real 256x256 geometry of config/vunet.yaml.

Tolerances (BASELINE.json north_star): relative L2 <= 1e-2 in bf16, <= 1e-4 in the fp32 validation mode (images,
latents and gradients); latent noise bit-exact with torch's generator in the reference's draw order.
"""
import contextlib
import io

import pytest
import torch

import vunet_oracle as O

pytestmark = pytest.mark.gpu

BF16_TOL = 1e-2
FP32_TOL = 1e-4
ARCH = dict(n_latent_scales=2, conv_layer_type="l1", subpixel_upsampling=True, n_scales=0, nf_start=32, nf_max=128)
D256 = dict(ARCH, spatial_size=256, bottleneck_factor=2, box_factor=2, n_channels_x=30)   # config/vunet.yaml
H256 = dict(ARCH, spatial_size=256, bottleneck_factor=2, box_factor=2, n_channels_x=3)


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def build(cfg, sd=None):
    import bdvs_b200 as B
    cfg = {k: v for k, v in cfg.items() if k != "batch"}
    with contextlib.redirect_stdout(io.StringIO()):
        m = B.VunetOrg(**cfg)
    if sd is None:
        sd = O.det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()})
    m.load_state_dict(sd)
    return m.cuda().eval(), sd


def cuda(ts):
    return [t.cuda() for t in ts]


def cuda2(tss):
    return [[t.cuda() for t in ts] for ts in tss]


@pytest.mark.parametrize("name", ["org_tiny3", "org_tiny30", "org_m128"])
def test_model_matches_reference_golden(golden, name):
    import bdvs_b200 as B
    fx = golden(name)
    cfg = fx["config"]
    m, sd = build(cfg)
    bs = cfg["batch"]
    x, c, eps, _ = O.make_inputs(cfg, bs)
    geps = O.make_group_eps(cfg, bs)
    with torch.no_grad():
        imgs, qs, ps, acts = m(x.cuda(), c.cuda(), eps=cuda(eps), eps_groups=cuda2(geps))
        timgs = m.transfer(x.cuda(), c.cuda(), eps=cuda(eps), eps_groups=cuda2(geps))
    errs = {"imgs": rel(imgs, fx["imgs"]), "transfer": rel(timgs, fx["transfer"])}
    for i in range(2):
        errs[f"q{i}"] = rel(qs[i], fx["qs"][i])
        errs[f"p{i}"] = rel(ps[i], fx["ps"][i])
    print(name, {k: f"{v:.2e}" for k, v in errs.items()})
    assert imgs.shape == fx["imgs"].shape and imgs.dtype == torch.float32
    assert max(errs.values()) < BF16_TOL, errs
    # return structure of the reference: activations = (hs left after ed's pops, qs, gs, ds)
    assert len(acts) == 4 and len(acts[0]) == fx["n_hs_left"] and len(acts[2]) == fx["n_activations"][2]
    assert acts[1] is qs
    kl = float(B.compute_kl_loss(ps, qs))
    assert abs(kl - fx["kl"]) < 2e-2 * abs(fx["kl"])
    if "test_forward" in fx:
        with torch.no_grad():
            pimgs = m.test_forward(c.cuda(), eps_groups=cuda2(geps))
        assert rel(pimgs, fx["test_forward"]) < BF16_TOL


@pytest.mark.parametrize("name", ["org_tiny3"])
def test_submodules_prior_samples_and_draw_order(golden, name):
    """``dd(gs, [], training=False)`` returns the merged prior means / samples of the reference, ``ed(eu(x))`` pops in
    place and returns (hs, qs, zs); the noise is drawn with torch.randn in the reference's order and shapes, so a
    seeded run consumes the global generator exactly like the reference does."""
    fx = golden(name)
    cfg = fx["config"]
    m, sd = build(cfg)
    bs = cfg["batch"]
    x, c, eps, _ = O.make_inputs(cfg, bs)
    geps = O.make_group_eps(cfg, bs)
    import bdvs_b200.vunets as V
    orig = V._draw_eps
    shapes = []

    def injected(draws):
        it = iter(draws)

        def f(shape, device):
            shapes.append(tuple(shape))
            return next(it)
        return f

    with torch.no_grad():
        V._draw_eps = injected(cuda([e for grp in geps for e in grp]))
        try:
            imgs, hs, tps, tzs = m.dd(m.du(c.cuda()), [], False)
        finally:
            V._draw_eps = orig
        assert rel(imgs, fx["test_forward"]) < BF16_TOL
        for i in range(2):
            assert rel(tps[i], fx["test_ps"][i]) < BF16_TOL
            assert rel(tzs[i], fx["test_zs"][i]) < BF16_TOL
        assert shapes == [tuple(e.shape) for grp in geps for e in grp]   # 2 x 4 quarter-resolution draws
        shapes.clear()
        V._draw_eps = injected(cuda(eps))
        try:
            hs_in = m.eu(x.cuda())
            n_in = len(hs_in)
            ed_hs, qs, zs = m.ed(hs_in)
        finally:
            V._draw_eps = orig
        assert n_in - len(hs_in) == 5 and len(ed_hs) == 5
        assert shapes == [tuple(e.shape) for e in eps]
        for i in range(2):
            assert rel(zs[i], fx["zs"][i]) < BF16_TOL
            assert rel(qs[i], fx["qs"][i]) < BF16_TOL
        # seeded: forward() draws ed's two tensors, then dd's eight -- the same calls torch.randn_like makes
        torch.manual_seed(7)
        ref_draws = [torch.randn(tuple(e.shape), device="cuda") for e in eps] + \
                    [torch.randn(tuple(e.shape), device="cuda") for grp in geps for e in grp]
        after_ref = torch.randn(4, device="cuda")
        torch.manual_seed(7)
        m(x.cuda(), c.cuda())
        after = torch.randn(4, device="cuda")
        assert torch.equal(after, after_ref)  # the generator advanced by exactly the reference's draws
        torch.manual_seed(7)
        a = m(x.cuda(), c.cuda())[0]
        b = m(x.cuda(), c.cuda(), eps=ref_draws[:2], eps_groups=[ref_draws[2:6], ref_draws[6:]])[0]
        assert torch.equal(a, b)


@pytest.mark.parametrize("cfg,batch", [(D256, 2), (H256, 8)])
def test_full_size_matches_oracle(cfg, batch):
    """config/vunet.yaml geometry (256x256, 128 -> 512 / 1024-channel prior convs on the 4x4 and 8x8 latent maps: the
    split-K cluster kernel) against the CPU oracle."""
    m, sd = build(cfg)
    x, c, eps, _ = O.make_inputs(cfg, batch)
    geps = O.make_group_eps(cfg, batch)
    with torch.no_grad():
        imgs, qs, ps, _ = m(x.cuda(), c.cuda(), eps=cuda(eps), eps_groups=cuda2(geps))
        timgs = m.transfer(x.cuda(), c.cuda(), eps=cuda(eps), eps_groups=cuda2(geps))
        pimgs = m.test_forward(c.cuda(), eps_groups=cuda2(geps))
        o_imgs, o_qs, o_ps, _, _, _ = O.org_forward(sd, x, c, eps, geps)
        o_t = O.org_transfer(sd, x, c, eps, geps)
        o_p = O.org_test_forward(sd, c, geps)
    errs = {"imgs": rel(imgs, o_imgs), "transfer": rel(timgs, o_t), "test_forward": rel(pimgs, o_p)}
    for i in range(2):
        errs[f"q{i}"] = rel(qs[i], o_qs[i])
        errs[f"p{i}"] = rel(ps[i], o_ps[i])
    worst_frame = max(rel(timgs[b], o_t[b]) for b in range(batch))
    print(f"VunetOrg {cfg['n_channels_x']}ch B={batch}:", {k: f"{v:.2e}" for k, v in errs.items()},
          f"worst frame {worst_frame:.2e}")
    assert max(errs.values()) < BF16_TOL, errs
    assert worst_frame < BF16_TOL
    # uint8 frame epilogue and batch independence of the split-K layers (chosen from the layer geometry only)
    with torch.no_grad():
        u8 = m.transfer(x.cuda(), c.cuda(), eps=cuda(eps), eps_groups=cuda2(geps), uint8_frames=True)
        one = m.transfer(x[:1].cuda(), c[:1].cuda(), eps=[e[:1].cuda() for e in eps],
                         eps_groups=[[e[:1].cuda() for e in grp] for grp in geps])
    ref8 = (((timgs + 1.0) / 2.0).clamp(0.0, 1.0) * 255.0).permute(0, 2, 3, 1).to(torch.uint8)
    assert torch.equal(u8, ref8)
    assert torch.equal(one[0], timgs[0])


@pytest.mark.parametrize("name", ["org_tiny3", "org_tiny30", "org_m128"])
def test_fp32_validation_mode_matches_reference_to_1e4(golden, name):
    """Forward, transfer, test_forward, KL and ALL parameter gradients of the fused schedule in fp32 against the
    reference's own tensors and the oracle's fp32 autograd."""
    import bdvs_b200 as B
    fx = golden(name)
    cfg = fx["config"]
    m, sd = build(cfg)
    bs = cfg["batch"]
    x, c, eps, _ = O.make_inputs(cfg, bs)
    geps = O.make_group_eps(cfg, bs)
    target = O.det_uniform("target", (bs, 3, cfg["spatial_size"], cfg["spatial_size"]))
    with torch.no_grad(), B.fp32_validation(m):
        imgs, qs, ps, _ = m(x.cuda(), c.cuda(), eps=cuda(eps), eps_groups=cuda2(geps))
        timgs = m.transfer(x.cuda(), c.cuda(), eps=cuda(eps), eps_groups=cuda2(geps))
        pimgs = m.test_forward(c.cuda(), eps_groups=cuda2(geps))
    errs = {"imgs": rel(imgs, fx["imgs"]), "transfer": rel(timgs, fx["transfer"])}
    for i in range(2):
        errs[f"q{i}"] = rel(qs[i], fx["qs"][i])
        errs[f"p{i}"] = rel(ps[i], fx["ps"][i])
    errs["kl"] = abs(float(B.compute_kl_loss(ps, qs)) - fx["kl"]) / abs(fx["kl"])
    if "test_forward" in fx:
        errs["test_forward"] = rel(pimgs, fx["test_forward"])
    print(name, {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < FP32_TOL, errs
    # ---- gradients ----
    m.train()
    with B.fp32_validation(m):
        imgs, qs, ps, _ = m(x.cuda(), c.cuda(), eps=cuda(eps), eps_groups=cuda2(geps))
        assert imgs.requires_grad and qs[0].requires_grad and ps[1].requires_grad
        loss = torch.mean((imgs - target.cuda()) ** 2) + 1e-3 * B.compute_kl_loss(ps, qs)
        loss.backward()
    torch.cuda.synchronize()
    got = {k: (p.grad.detach().float().cpu() if p.grad is not None else None) for k, p in m.named_parameters()}
    sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o_imgs, o_qs, o_ps, _, _, _ = O.org_forward(sd_g, x, c, eps, geps)
    o_loss = torch.mean((o_imgs - target) ** 2) + 1e-3 * O.compute_kl_loss(o_ps, o_qs)
    o_loss.backward()
    ref = {k: v.grad for k, v in sd_g.items()}
    assert abs(float(loss) - fx["loss"]) < 1e-4 * abs(fx["loss"])
    assert sorted(k for k, g in got.items() if g is None) == fx["grad_is_none"]
    keys = [k for k, g in ref.items() if g is not None and float(g.abs().max()) > 0.0]
    num = sum(float((got[k] - ref[k]).pow(2).sum()) for k in keys)
    den = sum(float(ref[k].pow(2).sum()) for k in keys)
    total = (num / den) ** 0.5
    worst_k = max(keys, key=lambda k: float((got[k] - ref[k]).norm() / ref[k].norm().clamp_min(1e-20)))
    worst = float((got[worst_k] - ref[worst_k]).norm() / ref[worst_k].norm().clamp_min(1e-20))
    fixture = max(float((got[k] - g).norm() / g.norm()) for k, g in fx["grad_full"].items())
    print(f"{name}: fp32 validation gradients: concatenated rel-L2 {total:.2e}, worst tensor {worst_k} {worst:.2e}, "
          f"vs reference fixture tensors {fixture:.2e}")
    assert total < FP32_TOL
    assert worst < 20 * FP32_TOL
    assert fixture < FP32_TOL * 10
    for k, n_ref in fx["grad_norms"].items():
        assert abs(float(got[k].norm()) - n_ref) < 1e-3 * n_ref + 1e-9, k


# ---------------------------------------------------------------------------------------------------------------
# bf16 training: forward + the hand-written backward schedule (tcgen05 dgrad / wgrad, 128-channel slices for the
# 256 / 512-channel prior convs) against the oracle's autograd and the reference's own gradient fixtures
# ---------------------------------------------------------------------------------------------------------------
GRAD_TOL = 1e-2


def _bf16_grads(m, sd, cfg, bs, tag, fx=None):
    import bdvs_b200 as B
    m.train()
    x, c, eps, _ = O.make_inputs(cfg, bs)
    geps = O.make_group_eps(cfg, bs)
    target = O.det_uniform("target", (bs, 3, cfg["spatial_size"], cfg["spatial_size"]))
    imgs, qs, ps, _ = m(x.cuda(), c.cuda(), eps=cuda(eps), eps_groups=cuda2(geps))
    assert imgs.requires_grad and qs[0].requires_grad and ps[1].requires_grad
    loss = torch.mean((imgs - target.cuda()) ** 2) + 1e-3 * B.compute_kl_loss(ps, qs)
    loss.backward()
    torch.cuda.synchronize()
    got = {k: (p.grad.detach().float().cpu() if p.grad is not None else None) for k, p in m.named_parameters()}
    sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o_imgs, o_qs, o_ps, _, _, _ = O.org_forward(sd_g, x, c, eps, geps)
    o_loss = torch.mean((o_imgs - target) ** 2) + 1e-3 * O.compute_kl_loss(o_ps, o_qs)
    o_loss.backward()
    ref = {k: v.grad for k, v in sd_g.items()}
    assert abs(float(loss) - float(o_loss)) < 2e-2 * abs(float(o_loss))
    none_ref = sorted(k for k, g in ref.items() if g is None or float(g.abs().max()) == 0.0)
    assert sorted(k for k, g in got.items() if g is None) == none_ref
    keys = [k for k in ref if k not in none_ref]
    for k in keys:
        assert torch.isfinite(got[k]).all(), k

    def crel(ks):
        num = sum(float((got[k] - ref[k]).pow(2).sum()) for k in ks)
        den = sum(float(ref[k].pow(2).sum()) for k in ks)
        return (num / den) ** 0.5

    total = crel(keys)
    per_owner = {o: crel([k for k in keys if k.startswith(o + ".")]) for o in ("eu", "ed", "du", "dd")}
    prior = crel([k for k in keys if ".auto_lp." in k or ".auto_blocks." in k or ".latent_nins." in k])
    worst = max(keys, key=lambda k: float((got[k] - ref[k]).norm() / ref[k].norm().clamp_min(1e-12)))
    worst_err = float((got[worst] - ref[worst]).norm() / ref[worst].norm().clamp_min(1e-12))
    print(f"{tag}: bf16 concatenated grad rel-L2 {total:.3e}, per sub-network {per_owner}, prior convs {prior:.3e}, "
          f"worst tensor {worst} {worst_err:.3e}")
    assert total < GRAD_TOL and prior < 2 * GRAD_TOL
    assert all(v < 2 * GRAD_TOL for v in per_owner.values()), per_owner
    assert worst_err < 0.15, (worst, worst_err)
    if fx is not None:
        assert abs(float(loss) - fx["loss"]) < 2e-2 * abs(fx["loss"])
        assert sorted(k for k, g in got.items() if g is None) == fx["grad_is_none"]
        num = den = 0.0
        for k, n_ref in fx["grad_norms"].items():
            num += (float(got[k].norm()) - n_ref) ** 2
            den += n_ref ** 2
        assert (num / den) ** 0.5 < GRAD_TOL
        for k, g_ref in fx["grad_full"].items():
            assert float((got[k] - g_ref).norm() / g_ref.norm()) < 5e-2, k
    return got


@pytest.mark.parametrize("name", ["org_tiny3", "org_tiny30", "org_m128"])
def test_bf16_gradients_match_oracle_and_reference_fixture(golden, name):
    fx = golden(name)
    cfg = fx["config"]
    m, sd = build(cfg)
    got = _bf16_grads(m, sd, cfg, cfg["batch"], name, fx)
    # gradients accumulate over two backward passes like autograd's, and an optimizer step moves the loss
    import bdvs_b200 as B
    x, c, eps, _ = O.make_inputs(cfg, cfg["batch"])
    geps = O.make_group_eps(cfg, cfg["batch"])
    target = O.det_uniform("target", (cfg["batch"], 3, cfg["spatial_size"], cfg["spatial_size"])).cuda()
    imgs, qs, ps, _ = m(x.cuda(), c.cuda(), eps=cuda(eps), eps_groups=cuda2(geps))
    (torch.mean((imgs - target) ** 2) + 1e-3 * B.compute_kl_loss(ps, qs)).backward()
    k = "dd.auto_lp.l_1.2.conv.weight_v"
    g2 = dict(m.named_parameters())[k].grad.float().cpu()
    assert float((g2 - 2 * got[k]).norm() / (2 * got[k]).norm()) < 1e-3
    opt = torch.optim.Adam(m.parameters(), lr=2e-5, betas=(0.5, 0.9))  # small steps: a first-order decrease
    losses = []
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        imgs, qs, ps, _ = m(x.cuda(), c.cuda(), eps=cuda(eps), eps_groups=cuda2(geps))
        loss = torch.mean((imgs - target) ** 2) + 1e-3 * B.compute_kl_loss(ps, qs)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]


def test_bf16_gradients_full_size():
    """config/vunet.yaml geometry (256x256, 30 appearance channels), batch 2."""
    m, sd = build(D256)
    _bf16_grads(m, sd, D256, 2, "VunetOrg D256 B=2")


def test_sequence_split_equals_transfer(golden):
    """encode_appearance once + decode_pose == transfer on the repeated appearance, also through the
    CUDA-graph synthesizers."""
    from bdvs_b200.synthesis import FrameSynthesizer, SequenceSynthesizer
    fx = golden("org_tiny3")
    cfg = fx["config"]
    m, sd = build(cfg)
    x, c, eps, _ = O.make_inputs(cfg, 1)
    c4 = O.det_uniform("c4", (4, 3, cfg["spatial_size"], cfg["spatial_size"]))
    with torch.no_grad():
        app = m.encode_appearance(x.cuda(), eps=cuda(eps))
        a = m.decode_pose(c4.cuda(), app, frames_per_appearance=4)
        b = m.transfer(x.cuda().repeat_interleave(4, 0), c4.cuda(), eps=[e.cuda().repeat_interleave(4, 0) for e in eps])
    assert torch.equal(a, b)
    torch.manual_seed(3)
    fs = FrameSynthesizer(m, 4)
    out = fs(x.cuda().repeat_interleave(4, 0), c4.cuda()).clone()
    torch.manual_seed(3)
    with torch.no_grad():
        ref = m.transfer(x.cuda().repeat_interleave(4, 0), c4.cuda())
    assert out.shape == ref.shape and torch.isfinite(out).all()
    assert rel(out, ref) < 5e-2  # different noise draws inside the graph: same distribution, not the same bits
    ss = SequenceSynthesizer(m, 1, 4)
    assert ss(x.cuda(), c4.cuda()).shape == ref.shape


def test_graphed_train_step_with_fused_adam(golden):
    """bdvs_b200.FusedAdam + GraphedTrainStep on VunetOrg (loss_fn receives (imgs, qs, ps, target)): the replayed
    graph equals the same steps issued launch by launch."""
    import copy
    import bdvs_b200 as B
    fx = golden("org_tiny3")
    cfg = fx["config"]
    a, sd = build(cfg)
    a.train()
    b = copy.deepcopy(a)
    bs = 4
    x, c, eps, _ = O.make_inputs(cfg, bs)
    target = O.det_uniform("target", (bs, 3, cfg["spatial_size"], cfg["spatial_size"])).cuda()
    x, c, eps = x.cuda(), c.cuda(), cuda(eps)

    def loss_fn(imgs, qs, ps, tgt):
        return ((imgs - tgt) ** 2).mean() + 1e-3 * B.compute_kl_loss(ps, qs)

    def groups(m):
        return [{"params": list(getattr(m, n).parameters()), "name": n} for n in ("eu", "ed", "du", "dd")]

    opt_a = B.FusedAdam(groups(a), lr=1e-4, betas=(0.5, 0.9), model=a)
    opt_b = B.FusedAdam(groups(b), lr=1e-4, betas=(0.5, 0.9), model=b)
    losses_a = []
    for _ in range(3):
        imgs, qs, ps, _acts = a(x, c, eps=eps)
        loss = loss_fn(imgs, qs, ps, target)
        opt_a.zero_grad()
        loss.backward()
        opt_a.step()
        losses_a.append(float(loss))
    step = B.GraphedTrainStep(b, opt_b, loss_fn, x, c, target, eps=eps)
    losses_b = [float(step(x, c, target)) for _ in range(3)]
    for la, lb in zip(losses_a, losses_b):
        assert abs(la - lb) <= 2e-3 * abs(la), (losses_a, losses_b)
    assert losses_b[-1] < losses_b[0]
    for (n, pa), pb in zip(a.named_parameters(), b.parameters()):
        torch.testing.assert_close(pa, pb, rtol=1e-3, atol=1e-4, msg=n)
    # the prior convs were stepped too
    k = "dd.auto_lp.l_1.2.conv.weight_v"
    assert float((dict(b.named_parameters())[k].detach().cpu() - sd[k]).abs().max()) > 1e-5
