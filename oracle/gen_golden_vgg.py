"""Generates ``tests/golden/vgg_tiny.pt`` from the UNMODIFIED reference ``PerceptualVGG`` / ``vgg_loss`` — test
infrastructure, build container only.

``lib/losses.py`` and ``models/imagenet_pretrained.py`` import helper modules that need packages absent here (kornia,
matplotlib, ... through ``lib/utils.py``) — none of them on the path under test: they are replaced by stubs that provide
the three names those files import (``get_member`` restated from lib/utils.py:37-43), and the two reference files are
executed as they are.  The VGG19 itself is torchvision's (the reference's own dependency) with deterministic hashed
weights (``vgg_oracle.det_vgg_state_dict``; ImageNet weights do not exist offline).

Usage:  python oracle/gen_golden_vgg.py [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import vgg_oracle as V  # noqa: E402
import vunet_oracle as O  # noqa: E402


def import_reference(ref_root: str):
    def get_member(model, name):  # lib/utils.py:37-43
        module = model.module if isinstance(model, nn.DataParallel) else model
        return getattr(module, name)

    utils = types.ModuleType("lib.utils")
    utils.get_member = get_member
    utils.bounding_box_batch = utils.toggle_grad = utils.scale_img = None   # imported by name only, not on this path
    disc = types.ModuleType("models.pose_discriminator")
    disc.MIDisc = disc.MIDiscConv1 = None
    sys.modules["lib.utils"] = utils
    sys.modules["models.pose_discriminator"] = disc
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    from lib.losses import vgg_loss          # the reference's own functions, unmodified
    from models.imagenet_pretrained import PerceptualVGG
    return PerceptualVGG, vgg_loss


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    args = ap.parse_args()
    from torchvision.models import vgg19
    PerceptualVGG, ref_vgg_loss = import_reference(args.ref)
    torch.manual_seed(0)
    vgg = vgg19(weights=None).eval()
    sd = V.det_vgg_state_dict()
    vgg.features.load_state_dict(sd, strict=True)
    weights = [1.0, 0.5, 2.0, 1.5, 1.0, 3.0]
    custom = PerceptualVGG(vgg, weights).eval()
    B, S = 2, 32
    target = O.det_uniform("vgg_target", (B, 3, S, S))
    pred = (O.det_uniform("vgg_pred", (B, 3, S, S)) * 0.9).requires_grad_(True)
    losses = ref_vgg_loss(custom, target, pred)
    coeff = [1.0, 2.0, 0.5, 1.5, 1.0, 0.7]   # unequal upstream gradients per term
    total = sum(c * losses[k] for c, k in zip(coeff, losses)).sum()
    total.backward()
    feats = custom(target)
    # ---- oracle must agree ----
    p2 = pred.detach().clone().requires_grad_(True)
    o_losses = V.vgg_loss(sd, target, p2, weights)
    sum(c * o_losses[k] for c, k in zip(coeff, o_losses)).sum().backward()
    o_feats = V.perceptual_features(sd, target)
    errs = {k: abs(float(o_losses[k]) - float(losses[k])) / abs(float(losses[k])) for k in losses}
    errs["d_pred"] = float((p2.grad - pred.grad).norm() / pred.grad.norm())
    for k in o_feats:
        errs["feat_" + k] = float((o_feats[k] - feats[k]).norm() / feats[k].norm())
    assert list(losses.keys()) == list(V.TAP_NAMES)
    print("oracle vs reference:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < 1e-5, errs
    fixture = {
        "weights": weights, "coeff": coeff, "batch": B, "size": S,
        "state_dict_keys": list(custom.state_dict().keys()),
        "losses": {k: float(v) for k, v in losses.items()},
        "loss_shapes": {k: tuple(v.shape) for k, v in losses.items()},
        "d_pred": pred.grad.clone(),
        "feat_shapes": {k: tuple(v.shape) for k, v in feats.items()},
        "feat_mean_abs": {k: float(v.abs().mean()) for k, v in feats.items()},
        "relu5_2": feats["relu5_2"].detach().clone(),
        "oracle_vs_reference": errs, "torch_version": torch.__version__,
    }
    path = os.path.join(args.out, "vgg_tiny.pt")
    torch.save(fixture, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB); losses {fixture['losses']}")


if __name__ == "__main__":
    main()
