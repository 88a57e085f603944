"""CPU: the VunetOrg family (models/vunets.py:18-106, 151-219, 600-783) — oracle restatement against the golden vectors
generated from the UNMODIFIED reference class (oracle/gen_golden_org.py), module structure / state_dict layout / init
stream of the host mirror, the KL terms of lib/losses.py:26-65, and loud failure without a GPU."""
import contextlib
import io
import os
import sys

import pytest
import torch

import vunet_oracle as O
from conftest import REFERENCE_ROOT

import bdvs_b200 as B

NAMES = ["org_tiny3", "org_tiny30", "org_m128"]
H256 = dict(spatial_size=256, bottleneck_factor=2, n_scales=0, box_factor=2, n_latent_scales=2,
            conv_layer_type="l1", nf_start=32, nf_max=128, subpixel_upsampling=True)


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_golden(golden, name):
    fx = golden(name)
    cfg = fx["config"]
    sd = O.det_state_dict(dict(fx["param_shapes"]))
    x, c, eps, _ = O.make_inputs(cfg, cfg["batch"])
    geps = O.make_group_eps(cfg, cfg["batch"])
    with torch.no_grad():
        imgs, qs, ps, zpost, zprior, hs = O.org_forward(sd, x, c, eps, geps)
        timgs = O.org_transfer(sd, x, c, eps, geps)
    assert rel(imgs, fx["imgs"]) < 1e-5
    assert rel(timgs, fx["transfer"]) < 1e-5
    for i in range(2):
        assert rel(qs[i], fx["qs"][i]) < 1e-5
        assert rel(ps[i], fx["ps"][i]) < 1e-5
    assert len(hs) == fx["n_hs_left"]
    kl = float(O.compute_kl_loss(ps, qs))
    assert abs(kl - fx["kl"]) / abs(fx["kl"]) < 1e-5
    if "test_forward" in fx:
        with torch.no_grad():
            gs = O.enc_up(sd, "du", c)
            timg, _, tps, tzs = O.dec_down(sd, "dd", gs, [], False, geps)
        assert rel(timg, fx["test_forward"]) < 1e-5
        for i in range(2):
            assert rel(zpost[i], fx["zs"][i]) < 1e-5
            assert rel(tps[i], fx["test_ps"][i]) < 1e-5
            assert rel(tzs[i], fx["test_zs"][i]) < 1e-5


@pytest.mark.parametrize("name", ["org_tiny3", "org_tiny30"])
def test_oracle_gradients_match_reference_golden(golden, name):
    fx = golden(name)
    cfg = fx["config"]
    sd = {k: v.requires_grad_(True) for k, v in O.det_state_dict(dict(fx["param_shapes"])).items()}
    x, c, eps, _ = O.make_inputs(cfg, cfg["batch"])
    geps = O.make_group_eps(cfg, cfg["batch"])
    target = O.det_uniform("target", (cfg["batch"], 3, cfg["spatial_size"], cfg["spatial_size"]))
    imgs, qs, ps, _, _, _ = O.org_forward(sd, x, c, eps, geps)
    loss = torch.mean((imgs - target) ** 2) + 1e-3 * O.compute_kl_loss(ps, qs)
    loss.backward()
    assert abs(float(loss) - fx["loss"]) / fx["loss"] < 1e-5
    none = sorted(k for k, v in sd.items() if v.grad is None or float(v.grad.abs().max()) == 0.0)
    assert none == fx["grad_is_none"]  # the dead branch: ed.blocks.3 / ed.ups.1 / ed.fin_block
    for k, n in fx["grad_norms"].items():
        assert abs(float(sd[k].grad.norm()) - n) <= 2e-4 * n + 1e-9, k
    for k, g in fx["grad_full"].items():
        assert rel(sd[k].grad, g) < 1e-4, k


def test_space_to_depth_restatement():
    x = O.det_uniform("s2d", (2, 8, 4, 6))
    y = O.space_to_depth(x)
    assert torch.equal(y, B.SpaceToDepth(2)(x))
    assert torch.equal(O.depth_to_space(y), x)
    # group (gi, gj) of SpaceToDepth is the parity sub-grid: the fact the NHWC kernels are built on
    for g, blk in enumerate(torch.split(y, 8, dim=1)):
        assert torch.equal(blk, x[:, :, g // 2::2, g % 2::2])


@pytest.mark.parametrize("name", NAMES)
def test_state_dict_layout_matches_reference_fixture(golden, name):
    fx = golden(name)
    cfg = dict(fx["config"])
    cfg.pop("batch")
    m = quiet(B.VunetOrg, **cfg)
    got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert got == [(k, tuple(s)) for k, s in fx["param_shapes"]]  # names, shapes AND registration order
    sd = O.det_state_dict(dict(fx["param_shapes"]))
    m.load_state_dict(sd, strict=True)
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k])


def test_ctor_contract():
    m = quiet(B.VunetOrg, n_channels_x=3, unknown_key=1, dataset="DeepFashion", dropout_prob=0.05, **H256)
    assert (m.n_scales, m.n_scales_x, m.spatial_size, m.n_channels_x) == (7, 7, 256, 3)
    # +33 M parameters of the 4-group prior on top of the shared encoder / decoder
    assert sum(v.numel() for v in m.state_dict().values()) > 47_000_000
    assert any(mod.p == 0.05 for mod in m.modules() if isinstance(mod, torch.nn.Dropout))
    for sub in ("eu", "ed", "du", "dd"):
        assert hasattr(m, sub)
    assert isinstance(m.ed, B.EncDown) and isinstance(m.dd, B.DecDown)
    with pytest.raises(KeyError):
        quiet(B.VunetOrg, n_channels_x=3)
    with pytest.raises(NotImplementedError):
        quiet(B.VunetOrg, **dict(H256, conv_layer_type="instance"))


@pytest.mark.skipif(not os.path.isdir(REFERENCE_ROOT), reason="reference tree only exists in the build container")
def test_same_seed_gives_reference_initialisation():
    sys.path.insert(0, REFERENCE_ROOT)
    from models.vunets import VunetOrg as RefVunet
    cfg = dict(H256, spatial_size=64, nf_max=64)
    torch.manual_seed(42)
    ref = quiet(RefVunet, n_channels_x=30, **cfg)
    torch.manual_seed(42)
    mine = quiet(B.VunetOrg, n_channels_x=30, **cfg)
    rsd, msd = ref.state_dict(), mine.state_dict()
    assert list(rsd.keys()) == list(msd.keys())
    for k in rsd:
        assert torch.equal(rsd[k], msd[k]), k  # bit-identical: same RNG stream, same formulas


def test_kl_terms_match_oracle():
    p = [O.det_normal("p0", (3, 8, 2, 2)), O.det_normal("p1", (3, 8, 4, 4))]
    q = [O.det_normal("q0", (3, 8, 2, 2)), O.det_normal("q1", (3, 8, 4, 4))]
    assert torch.allclose(B.latent_kl(p[0], q[0]), O.latent_kl(p[0], q[0]), rtol=1e-6)
    assert torch.allclose(B.compute_kl_loss(p, q), O.compute_kl_loss(p, q), rtol=1e-6)


def test_no_cpu_fallback():
    m = quiet(B.VunetOrg, n_channels_x=3, **dict(H256, spatial_size=32, bottleneck_factor=1, nf_max=64))
    x = torch.zeros(1, 3, 32, 32)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(x, x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.transfer(x, x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.test_forward(x)
