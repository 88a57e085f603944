"""CPU oracle for the VGG19 perceptual loss — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Functional fp32 restatement (plain torch CPU ops on a ``vgg19.features``-layout state dict ``{"N.weight", "N.bias"}``)
of the reference's ``models/imagenet_pretrained.py:8-61`` (``PerceptualVGG.forward``: normalisation, feature taps after
features[3, 8, 13, 22, 31]) and ``lib/losses.py:81-119`` (``vgg_loss``: ``loss_weights[i] * mean|tf - pf|`` per tap).
The feature stack itself is torchvision's ``vgg19().features`` (the reference instantiates ``vgg19(pretrained=True)``,
``experiments/shape_and_pose_net.py:222``; cfg "E": 64 64 M 128 128 M 256x4 M 512x4 M 512x4 M, 3x3 convs padding 1 + ReLU,
2x2 max-pooling) — third-party code, restated here from its published architecture.

Parity pinning: ``oracle/gen_golden_vgg.py`` runs the reference's UNMODIFIED ``PerceptualVGG`` and ``vgg_loss`` (imported
from /root/reference with the unrelated, un-importable helper modules stubbed) on a torchvision ``vgg19`` with deterministic
hashed weights and checks this restatement against them, losses and the gradient w.r.t. the predicted image
(``tests/golden/vgg_tiny.pt``).  The pretrained ImageNet weights do not exist offline, so the NUMBERS of a real training
run cannot be reproduced; the arithmetic is weight-independent.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

import vunet_oracle as O

CFG_E = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]
TARGET_LAYERS = {3: "relu1_2", 8: "relu2_2", 13: "relu3_2", 22: "relu4_2", 31: "relu5_2"}
TAP_NAMES = ("input", "relu1_2", "relu2_2", "relu3_2", "relu4_2", "relu5_2")
MEAN = torch.tensor([0.485, 0.456, 0.406]).reshape(1, 3, 1, 1)
STD = torch.tensor([0.229, 0.224, 0.225]).reshape(1, 3, 1, 1)


def feature_layout() -> List[tuple]:
    """(features index, kind, cin, cout) of torchvision's vgg19().features."""
    out, idx, cin = [], 0, 3
    for v in CFG_E:
        if v == "M":
            out.append((idx, "pool", cin, cin))
            idx += 1
        else:
            out.append((idx, "conv", cin, v))
            out.append((idx + 1, "relu", v, v))
            idx += 2
            cin = v
    return out


def det_vgg_state_dict() -> Dict[str, torch.Tensor]:
    """Deterministic hashed weights: w ~ U(+-sqrt(6 / fan_in)) (keeps ReLU activations O(1) through 16 layers),
    b ~ U(+-0.05)."""
    sd = {}
    for idx, kind, cin, cout in feature_layout():
        if kind != "conv":
            continue
        bound = float(np.sqrt(6.0 / (cin * 9)))
        sd[f"{idx}.weight"] = O.det_uniform(f"vgg.{idx}.weight", (cout, cin, 3, 3), -bound, bound)
        sd[f"{idx}.bias"] = O.det_uniform(f"vgg.{idx}.bias", (cout,), -0.05, 0.05)
    return sd


def perceptual_features(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> Dict[str, torch.Tensor]:
    """PerceptualVGG.forward — models/imagenet_pretrained.py:45-61 (only the layers up to the last tap)."""
    x = (x + 1.0) / 2.0
    x = (x - MEAN) / STD
    out = {"input": x}
    for idx, kind, _, _ in feature_layout():
        if idx > max(TARGET_LAYERS):
            break
        if kind == "conv":
            x = F.conv2d(x, sd[f"{idx}.weight"], sd[f"{idx}.bias"], padding=1)
        elif kind == "relu":
            x = F.relu(x)
        else:
            x = F.max_pool2d(x, kernel_size=2, stride=2)
        if idx in TARGET_LAYERS:
            out[TARGET_LAYERS[idx]] = x
    return out


def vgg_loss(sd: Dict[str, torch.Tensor], target: torch.Tensor, pred: torch.Tensor,
             loss_weights: Sequence[float]) -> Dict[str, torch.Tensor]:
    """lib/losses.py:81-104 (the ``weights is None`` branch, the only one a shipped caller takes)."""
    tf = perceptual_features(sd, target)
    pf = perceptual_features(sd, pred)
    return {name: loss_weights[i] * torch.mean(torch.abs(tf[name] - pf[name])).unsqueeze(dim=-1)
            for i, name in enumerate(TAP_NAMES)}
