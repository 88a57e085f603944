"""CPU: the VGG19 perceptual loss (lib/losses.py:81-119, models/imagenet_pretrained.py:8-61) — oracle restatement against
the golden vector generated from the UNMODIFIED reference functions (oracle/gen_golden_vgg.py), structure of the host
mirror, loud failure without a GPU."""
import pytest
import torch

import vgg_oracle as V
import vunet_oracle as O

import bdvs_b200 as B

torchvision = pytest.importorskip("torchvision")


def make_vgg():
    from torchvision.models import vgg19
    vgg = vgg19(weights=None).eval()
    vgg.features.load_state_dict(V.det_vgg_state_dict(), strict=True)
    return vgg


def test_oracle_matches_reference_golden(golden):
    fx = golden("vgg_tiny")
    sd = V.det_vgg_state_dict()
    Bs, S = fx["batch"], fx["size"]
    target = O.det_uniform("vgg_target", (Bs, 3, S, S))
    pred = (O.det_uniform("vgg_pred", (Bs, 3, S, S)) * 0.9).requires_grad_(True)
    losses = V.vgg_loss(sd, target, pred, fx["weights"])
    assert list(losses.keys()) == list(fx["losses"].keys()) == list(V.TAP_NAMES)
    for k, v in fx["losses"].items():
        assert abs(float(losses[k]) - v) <= 1e-5 * abs(v), k
        assert tuple(losses[k].shape) == fx["loss_shapes"][k] == (1,)
    sum(c * losses[k] for c, k in zip(fx["coeff"], losses)).sum().backward()
    assert float((pred.grad - fx["d_pred"]).norm() / fx["d_pred"].norm()) < 1e-5
    feats = V.perceptual_features(sd, target)
    assert {k: tuple(v.shape) for k, v in feats.items()} == fx["feat_shapes"]
    assert float((feats["relu5_2"] - fx["relu5_2"]).norm() / fx["relu5_2"].norm()) < 1e-5


def test_feature_layout_is_torchvision_vgg19():
    vgg = make_vgg()
    lay = V.feature_layout()
    mods = list(vgg.features)
    assert len(lay) == len(mods) == 37
    for (idx, kind, cin, cout), m in zip(lay, mods):
        if kind == "conv":
            assert isinstance(m, torch.nn.Conv2d) and (m.in_channels, m.out_channels) == (cin, cout)
        elif kind == "relu":
            assert isinstance(m, torch.nn.ReLU)
        else:
            assert isinstance(m, torch.nn.MaxPool2d)


def test_module_structure(golden):
    fx = golden("vgg_tiny")
    vgg = make_vgg()
    m = B.PerceptualVGG(vgg, fx["weights"])
    assert list(m.state_dict().keys()) == fx["state_dict_keys"]   # mean, std, vgg_layers.N.{weight,bias}
    assert m.vgg_layers is vgg.features and m.loss_weights == fx["weights"]
    kinds = [p[0] for p in m._plan]
    assert kinds.count("conv") == 14 and kinds.count("pool") == 4   # up to relu5_2; conv5_3 / conv5_4 / pool5 are dead
    assert [p[2] for p in m._plan if p[0] == "conv" and p[2]] == list(V.TAP_NAMES[1:])
    wrapped = B.PerceptualVGG(torch.nn.DataParallel(vgg), fx["weights"])
    assert wrapped.vgg_layers is vgg.features
    bad = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 5, padding=2), torch.nn.ReLU())

    class Fake:
        features = bad
    with pytest.raises(NotImplementedError):
        B.PerceptualVGG(Fake(), fx["weights"])


def test_no_cpu_fallback(golden):
    fx = golden("vgg_tiny")
    m = B.PerceptualVGG(make_vgg(), fx["weights"])
    x = torch.zeros(1, 3, 32, 32)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        B.vgg_loss(m, x, x)
    with pytest.raises(TypeError):
        B.vgg_loss(torch.nn.Identity(), x, x)
