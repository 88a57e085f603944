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


class _RoundBF16(torch.autograd.Function):
    """Value AND gradient rounded to bf16: what storing an activation / its gradient as bf16 does."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def perceptual_features(sd: Dict[str, torch.Tensor], x: torch.Tensor, bf16_model=False) -> Dict[str, torch.Tensor]:
    """PerceptualVGG.forward — models/imagenet_pretrained.py:45-61 (only the layers up to the last tap).

    ``bf16_model``: the same graph with the storage precision of a bf16 tensor-core evaluation — weights, the
    normalised image and every ReLU output (and their gradients) rounded to bf16, fp32 accumulation.  ``True`` rounds
    the weights as ``bf16(w)``; ``"wn"`` as ``bf16(w / ||w||) * ||w||`` per output channel (a plain conv pushed through
    a weight-norm representation with an fp32 scale, which is what the product's conv kernels consume).  The L1 terms
    make the GRADIENT a sign function of feature differences, so ANY two bf16 evaluations that round differently —
    these two models included — differ from each other and from the fp32 graph by O(0.2) in relative L2 (gradients
    only; the loss values agree to 1e-3).  A bf16 implementation is therefore checked against the model that rounds
    where it rounds."""
    x = (x + 1.0) / 2.0
    x = (x - MEAN) / STD
    out = {"input": x}
    if bf16_model:
        x = _RoundBF16.apply(x)
    for idx, kind, _, _ in feature_layout():
        if idx > max(TARGET_LAYERS):
            break
        if kind == "conv":
            w = sd[f"{idx}.weight"]
            if bf16_model == "wn":
                n = w.reshape(w.shape[0], -1).norm(dim=1).reshape(-1, 1, 1, 1)
                w = (w / n).bfloat16().float() * n
            elif bf16_model:
                w = w.bfloat16().float()
            x = F.conv2d(x, w, sd[f"{idx}.bias"], padding=1)
        elif kind == "relu":
            x = F.relu(x)
            if bf16_model:
                x = _RoundBF16.apply(x)
        else:
            x = F.max_pool2d(x, kernel_size=2, stride=2)
        if idx in TARGET_LAYERS:
            out[TARGET_LAYERS[idx]] = x
    return out


def vgg_loss(sd: Dict[str, torch.Tensor], target: torch.Tensor, pred: torch.Tensor,
             loss_weights: Sequence[float], bf16_model=False) -> Dict[str, torch.Tensor]:
    """lib/losses.py:81-104 (the ``weights is None`` branch, the only one a shipped caller takes)."""
    tf = perceptual_features(sd, target, bf16_model)
    pf = perceptual_features(sd, pred, bf16_model)
    return {name: loss_weights[i] * torch.mean(torch.abs(tf[name] - pf[name])).unsqueeze(dim=-1)
            for i, name in enumerate(TAP_NAMES)}
