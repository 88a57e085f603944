"""CPU: the oracle restatement against the committed golden vectors (generated from the unmodified
reference by oracle/gen_golden.py) and, when /root/reference is present, against the live reference."""
import contextlib
import io
import os
import sys

import pytest
import torch

import vunet_oracle as O
from conftest import REFERENCE_ROOT


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("name", ["tiny3", "tiny30", "market128"])
def test_oracle_matches_golden(golden, name):
    fx = golden(name)
    cfg = fx["config"]
    sd = O.det_state_dict(dict(fx["param_shapes"]))
    x, c, eps, prior = O.make_inputs(cfg, cfg["batch"])
    with torch.no_grad():
        imgs, means, logstds, zs, hs = O.forward(sd, x, c, eps)
        timgs = O.transfer(sd, x, c, eps)
    assert rel(imgs, fx["imgs"]) < 1e-5
    assert rel(timgs, fx["transfer"]) < 1e-5
    for i in range(2):
        assert rel(means[i], fx["means"][i]) < 1e-5
        assert rel(logstds[i], fx["logstds"][i]) < 1e-5
    if "zs" in fx:
        for i in range(2):
            assert rel(zs[i], fx["zs"][i]) < 1e-5
        with torch.no_grad():
            assert rel(O.test_forward(sd, c, prior), fx["test_forward"]) < 1e-5
    # ed pops 5 of eu's outputs
    assert len(hs) == fx["n_hs_left"]
    assert [tuple(h.shape) for h in hs] == fx["hs_left_shapes"]
    kl = float(O.compute_kl_with_prior(means, logstds))
    assert abs(kl - fx["kl"]) / abs(fx["kl"]) < 1e-5


@pytest.mark.parametrize("name", ["tiny3", "tiny30"])
def test_oracle_gradients_match_golden(golden, name):
    fx = golden(name)
    cfg = fx["config"]
    sd = {k: v.requires_grad_(True) for k, v in O.det_state_dict(dict(fx["param_shapes"])).items()}
    x, c, eps, _ = O.make_inputs(cfg, cfg["batch"])
    target = O.det_uniform("target", (cfg["batch"], 3, cfg["spatial_size"], cfg["spatial_size"]))
    imgs, means, logstds, _, _ = O.forward(sd, x, c, eps)
    loss = torch.mean((imgs - target) ** 2) + 1e-3 * O.compute_kl_with_prior(means, logstds)
    loss.backward()
    assert abs(float(loss) - fx["loss"]) / fx["loss"] < 1e-5
    none = sorted(k for k, v in sd.items() if v.grad is None or float(v.grad.abs().max()) == 0.0)
    assert none == fx["grad_is_none"]  # the 25 dead parameters of ed.blocks.3 / ed.ups.1 / ed.fin_block
    for k, n in fx["grad_norms"].items():
        assert abs(float(sd[k].grad.norm()) - n) <= 2e-4 * n + 1e-9, k
    for k, g in fx["grad_full"].items():
        assert rel(sd[k].grad, g) < 1e-4, k


def test_depth_to_space_is_dcr_not_pixel_shuffle():
    x = torch.arange(2 * 8 * 3 * 3, dtype=torch.float32).reshape(2, 8, 3, 3)
    y = O.depth_to_space(x)
    n, c, h, w = x.shape
    ref = x.view(n, 2, 2, c // 4, h, w).permute(0, 3, 4, 1, 5, 2).reshape(n, c // 4, 2 * h, 2 * w)
    assert torch.equal(y, ref)
    assert not torch.equal(y, torch.nn.functional.pixel_shuffle(x, 2))


def test_det_data_is_stable():
    # the hashed test data must never change silently: fixtures depend on it
    t = O.det_uniform("x", (5,))
    pinned = torch.tensor([0.4835610091686249, 0.6520955562591553, -0.5568044781684875, 0.5867308378219604,
                           0.42399054765701294])
    assert torch.allclose(t, pinned, atol=1e-7)
    a = O.det_uniform("abc", (4, 3))
    b = O.det_uniform("abc", (4, 3))
    assert torch.equal(a, b)
    assert not torch.equal(a, O.det_uniform("abd", (4, 3)))
    n = O.det_normal("n", (20000,))
    assert abs(float(n.mean())) < 0.03 and abs(float(n.std()) - 1.0) < 0.03


@pytest.mark.skipif(not os.path.isdir(REFERENCE_ROOT), reason="reference tree only exists in the build container")
def test_oracle_matches_live_reference():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
    import gen_golden as G
    cfg = dict(G.CONFIGS["tiny30"], batch=2, nf_max=32)
    with contextlib.redirect_stdout(io.StringIO()):
        model = G.build_reference(REFERENCE_ROOT, cfg)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = O.det_state_dict(shapes)
    model.load_state_dict(sd)
    x, c, eps, prior = O.make_inputs(cfg, 2, tag="live")
    with torch.no_grad():
        with G.injected_randn(eps):
            imgs, means, logstds, acts = model(x, c)
        with G.injected_randn(prior):
            pimgs = model.test_forward(c)
        o = O.forward(sd, x, c, eps)
        assert rel(o[0], imgs) < 1e-5
        assert rel(o[1][1], means[1]) < 1e-5 and rel(o[2][0], logstds[0]) < 1e-5
        assert rel(O.test_forward(sd, c, prior), pimgs) < 1e-5
