"""FusedAdam (vub_adam_wn_step) against torch.optim.Adam on the same gradients (SURVEY.md §8f-4;
experiments/shape_and_pose_net.py:237-246, 429-432)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


ARCH = dict(n_latent_scales=2, conv_layer_type="l1", subpixel_upsampling=True, n_scales=0, nf_start=32, nf_max=64,
            spatial_size=32, bottleneck_factor=1, box_factor=1, n_channels_x=3)


def _model(seed=0):
    import contextlib
    import io

    import bdvs_b200
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = bdvs_b200.VunetAlter(**ARCH)
    return m.to("cuda:0").train()


def _groups(m):
    return [{"params": list(getattr(m, n).parameters()), "name": n} for n in ("eu", "ed", "du", "dd")]


def _loss(m, x, c, eps):
    import bdvs_b200
    imgs, means, logstds, _ = m(x, c, eps=eps)
    return ((imgs - x) ** 2).mean() + 1e-3 * bdvs_b200.compute_kl_with_prior(means, logstds)


def _eps(m, x, c):
    with torch.no_grad():
        _, means, _, _ = m(x, c)
    g = torch.Generator(device="cuda:0").manual_seed(7)
    return [torch.randn(t.shape, device=t.device, generator=g) for t in means]


def test_fused_adam_matches_torch_adam():
    import bdvs_b200
    a = _model()
    b = copy.deepcopy(a)
    g = torch.Generator(device="cuda:0").manual_seed(1)
    x = torch.randn(4, 3, 32, 32, device="cuda:0", generator=g).clamp(-1, 1)
    c = torch.randn(4, 3, 32, 32, device="cuda:0", generator=g).clamp(-1, 1)
    eps = _eps(a, x, c)
    opt_a = bdvs_b200.FusedAdam(_groups(a), lr=2e-4, betas=(0.5, 0.9), model=a)
    opt_b = torch.optim.Adam(_groups(b), lr=2e-4, betas=(0.5, 0.9))
    for step in range(4):
        la = _loss(a, x, c, eps)
        opt_a.zero_grad()
        la.backward()
        # feed the SAME gradients to torch's Adam so that only the optimizer arithmetic is compared
        for pa, pb in zip(a.parameters(), b.parameters()):
            pb.grad = None if pa.grad is None else pa.grad.clone()
        opt_a.step()
        opt_b.step()
        worst = 0.0
        for (n, pa), pb in zip(a.named_parameters(), b.parameters()):
            err = (pa - pb).abs().max().item()
            worst = max(worst, err)
            assert err <= 2e-6 + 1e-5 * pb.abs().max().item(), (step, n, err)
        # the prepared weights were re-emitted by the fused launch: a forward of `a` must equal a forward of a
        # freshly prepared copy of the same parameters (b, through vub_wn_prep)
        with torch.no_grad():
            b.load_state_dict(a.state_dict())   # exact same parameters; bumps versions -> vub_wn_prep
            ia = a(x, c, eps=eps)[0]
            ib = b(x, c, eps=eps)[0]
        assert torch.equal(ia, ib), step
    sa, sb = opt_a.state_dict(), opt_b.state_dict()
    assert set(sa["state"].keys()) == set(sb["state"].keys())
    for k in sb["state"]:
        assert float(sa["state"][k]["step"]) == float(sb["state"][k]["step"])
        torch.testing.assert_close(sa["state"][k]["exp_avg"], sb["state"][k]["exp_avg"], rtol=1e-5, atol=1e-9)
        torch.testing.assert_close(sa["state"][k]["exp_avg_sq"], sb["state"][k]["exp_avg_sq"], rtol=1e-5, atol=1e-12)


def test_fused_adam_state_dict_round_trip_and_lr_change():
    import bdvs_b200
    a = _model(3)
    g = torch.Generator(device="cuda:0").manual_seed(2)
    x = torch.randn(2, 3, 32, 32, device="cuda:0", generator=g).clamp(-1, 1)
    c = torch.randn(2, 3, 32, 32, device="cuda:0", generator=g).clamp(-1, 1)
    eps = _eps(a, x, c)
    opt = bdvs_b200.FusedAdam(_groups(a), lr=1e-4, betas=(0.5, 0.9), model=a)
    ref = torch.optim.Adam(_groups(copy.deepcopy(a)), lr=1e-4, betas=(0.5, 0.9))
    for _ in range(2):
        opt.zero_grad()
        _loss(a, x, c, eps).backward()
        opt.step()
    sd = opt.state_dict()
    ref_groups = ref.state_dict()["param_groups"]
    assert [g_["params"] for g_ in sd["param_groups"]] == [g_["params"] for g_ in ref_groups]
    # torch's Adam accepts our checkpoint and vice versa
    ref.load_state_dict(sd)
    b = copy.deepcopy(a)
    opt2 = bdvs_b200.FusedAdam(_groups(b), lr=1e-4, betas=(0.5, 0.9), model=b)
    opt2.load_state_dict(ref.state_dict())
    for o in (opt, opt2):
        for g_ in o.param_groups:
            g_["lr"] = 5e-5 if g_["name"] == "dd" else 2e-5   # scheduler-style change, per group
    for o, m in ((opt, a), (opt2, b)):
        o.zero_grad()
        _loss(m, x, c, eps).backward()
        o.step()
    # not bit-equal: the stem weight gradient accumulates with float atomics (order varies run to run)
    for (n, pa), pb in zip(a.named_parameters(), b.named_parameters()):
        torch.testing.assert_close(pa, pb[1], rtol=1e-5, atol=2e-7, msg=n)


def test_fused_adam_invalidates_other_weight_caches():
    """Sub-module banks (vunet.eu called on its own) must see parameters updated through raw pointers."""
    import bdvs_b200
    a = _model(5)
    g = torch.Generator(device="cuda:0").manual_seed(3)
    x = torch.randn(2, 3, 32, 32, device="cuda:0", generator=g).clamp(-1, 1)
    c = torch.randn(2, 3, 32, 32, device="cuda:0", generator=g).clamp(-1, 1)
    eps = _eps(a, x, c)
    with torch.no_grad():
        before = [t.float().clone() for t in a.eu(x)]
    opt = bdvs_b200.FusedAdam(_groups(a), lr=1e-2, betas=(0.5, 0.9), model=a)
    opt.zero_grad()
    _loss(a, x, c, eps).backward()
    opt.step()
    b = copy.deepcopy(a)
    with torch.no_grad():
        b.load_state_dict(a.state_dict())
        after = [t.float() for t in a.eu(x)]
        fresh = [t.float() for t in b.eu(x)]
    assert any(not torch.equal(p, q) for p, q in zip(before, after))
    for p, q in zip(after, fresh):
        assert torch.equal(p, q)


def test_graphed_train_step_matches_eager_steps():
    """training.GraphedTrainStep (whole step in one CUDA graph, device-side Adam step count) against the same steps
    issued launch by launch."""
    import bdvs_b200
    a = _model(11)
    b = copy.deepcopy(a)
    g = torch.Generator(device="cuda:0").manual_seed(4)
    x = torch.randn(4, 3, 32, 32, device="cuda:0", generator=g).clamp(-1, 1)
    c = torch.randn(4, 3, 32, 32, device="cuda:0", generator=g).clamp(-1, 1)
    tgt = torch.randn(4, 3, 32, 32, device="cuda:0", generator=g).clamp(-1, 1)
    eps = _eps(a, x, c)

    def loss_fn(imgs, means, logstds, target):
        return ((imgs - target) ** 2).mean() + 1e-3 * bdvs_b200.compute_kl_with_prior(means, logstds)

    opt_a = bdvs_b200.FusedAdam(_groups(a), lr=2e-4, betas=(0.5, 0.9), model=a)
    opt_b = bdvs_b200.FusedAdam(_groups(b), lr=2e-4, betas=(0.5, 0.9), model=b)
    losses_a = []
    for it in range(4):
        if it == 2:  # a learning-rate schedule step, per group (shape_and_pose_net.py:311-319)
            for g_ in opt_a.param_groups:
                g_["lr"] = 5e-5 if g_["name"] == "dd" else 1e-4
        imgs, means, logstds, _ = a(x, c, eps=eps)
        loss = loss_fn(imgs, means, logstds, tgt)
        opt_a.zero_grad()
        loss.backward()
        opt_a.step()
        losses_a.append(float(loss))
    before = [p.detach().clone() for p in b.parameters()]
    step = bdvs_b200.GraphedTrainStep(b, opt_b, loss_fn, x, c, tgt, eps=eps)
    for p, q in zip(b.parameters(), before):   # warm-up ran on a snapshot
        assert torch.equal(p, q)
    assert opt_b._step == 0 and int(opt_b._step_dev) == 0
    losses_b = []
    for it in range(4):
        if it == 2:  # the captured graph must pick the new rates up (device table updated in place)
            for g_ in opt_b.param_groups:
                g_["lr"] = 5e-5 if g_["name"] == "dd" else 1e-4
        losses_b.append(float(step(x, c, tgt)))
    assert opt_b._step == 4 and int(opt_b._step_dev) == 4
    # not bit-equal: the stem weight gradient accumulates with float atomics, and Adam turns a last-bit difference
    # of a near-zero gradient into a visible fraction of lr
    for la, lb in zip(losses_a, losses_b):
        assert abs(la - lb) <= 1e-3 * abs(la), (losses_a, losses_b)
    assert losses_b[-1] < losses_b[0]
    for (n, pa), pb in zip(a.named_parameters(), b.parameters()):
        torch.testing.assert_close(pa, pb, rtol=1e-3, atol=1e-4, msg=n)   # 4 steps of lr = 2e-4 move a weight by <= 8e-4
    moved = max(float((pa - q).abs().max()) for pa, q in zip(a.parameters(), before))
    assert moved > 2e-4
    sa, sb = opt_a.state_dict(), opt_b.state_dict()
    assert set(sa["state"].keys()) == set(sb["state"].keys())
