"""VGG19 perceptual loss — host-side mirror of the reference's ``models/imagenet_pretrained.py:8-61``
(``PerceptualVGG``) and ``lib/losses.py:81-119`` (``vgg_loss``), the likelihood term of the VUnet objective
(``experiments/shape_and_pose_net.py:222-235, 388-395``): L1 distances between VGG19 features of the target and of the
predicted image at the taps ``input, relu1_2, relu2_2, relu3_2, relu4_2, relu5_2``.

``PerceptualVGG(vgg, weights)`` takes the same ``torchvision.models.vgg19`` object (or ``nn.DataParallel`` around it) and
keeps its ``features`` as ``vgg_layers`` — identical ``state_dict`` — but runs them on the sm_100a kernels: every
``Conv2d(3x3) + ReLU`` pair is one tcgen05 implicit-GEMM launch with a ReLU epilogue (``VUB_CONV_RELU``; the plain conv
``w, b`` is fed to the weight-norm preparation as ``v = w, g = ||w||, gamma = 1, beta = 0``), feature maps stay NHWC bf16,
max-pooling / normalisation / L1 reductions are vectorised memory-bound kernels.  :func:`vgg_loss` is ONE autograd node:
its backward walks the feature stack in reverse with ``vub_conv_dgrad`` (ReLU' from the stored activations, the tap
gradients entering as epilogue addends), fused max-pool + ReLU backward, and hands ``d loss / d pred`` (NCHW fp32) to the
VUnet's own backward node.  The VGG weights are frozen: no weight gradient is computed (the reference never optimises
them either).  Layers after ``relu5_2`` (``conv5_3``, ``conv5_4``, the last pool), which the reference evaluates although
nothing reads them, are skipped.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
from torch import nn

from . import ops
from .modules import WeightBank

TAP_NAMES = ("input", "relu1_2", "relu2_2", "relu3_2", "relu4_2", "relu5_2")


class _Params:
    __slots__ = ("weight_v", "weight_g", "bias")


class _PlainConv:
    """A frozen ``nn.Conv2d(3x3, padding 1)`` presented to :class:`~.modules.WeightBank` as a NormConv2d whose
    normalisation is the identity: ``v = w``, ``g = ||w||`` per output channel, ``gamma = 1``, ``beta = 0``."""

    _vub_stem = False
    _vub_split = 1
    kernel_size = 3
    stride = 1

    def __init__(self, conv: nn.Conv2d, device: torch.device):
        w = conv.weight.detach().to(device=device, dtype=torch.float32)
        cout, cin = w.shape[0], w.shape[1]
        if cin < 32:  # conv1_1: 3 input channels, padded to the kernels' smallest K block with zero weights
            wp = torch.zeros((cout, 32, 3, 3), dtype=torch.float32, device=device)
            wp[:, :cin] = w
            w, cin = wp, 32
        self.in_channels, self.out_channels = cin, cout
        self.conv = _Params()
        self.conv.weight_v = w.contiguous()
        self.conv.weight_g = w.reshape(cout, -1).norm(dim=1).contiguous()
        b = conv.bias.detach() if conv.bias is not None else torch.zeros(cout)
        self.conv.bias = b.to(device=device, dtype=torch.float32).contiguous()
        self.gamma = torch.ones(cout, dtype=torch.float32, device=device)
        self.beta = torch.zeros(cout, dtype=torch.float32, device=device)


class PerceptualVGG(nn.Module):
    """models/imagenet_pretrained.py:8-61.  ``forward(x)`` returns the reference's dict of feature taps (inference);
    training goes through :func:`vgg_loss`."""

    def __init__(self, vgg, weights):
        super().__init__()
        self.vgg_layers = vgg.module.features if isinstance(vgg, nn.DataParallel) else vgg.features
        self.loss_weights = weights
        self.register_buffer("mean", torch.tensor([0.485, 0.456, 0.406], dtype=torch.float)
                             .unsqueeze(dim=0).unsqueeze(dim=-1).unsqueeze(dim=-1))
        self.register_buffer("std", torch.tensor([0.229, 0.224, 0.225], dtype=torch.float)
                             .unsqueeze(dim=0).unsqueeze(dim=-1).unsqueeze(dim=-1))
        self.target_layers = {"3": "relu1_2", "8": "relu2_2", "13": "relu3_2", "22": "relu4_2", "31": "relu5_2"}
        # ---- the schedule: ("conv", features index, tap name or None) / ("pool",), up to the last tap ----
        plan: List[tuple] = []
        items = list(self.vgg_layers._modules.items())
        last = max(int(k) for k in self.target_layers)
        i = 0
        while i < len(items) and int(items[i][0]) <= last:
            name, mod = items[i]
            if isinstance(mod, nn.Conv2d):
                if (mod.kernel_size != (3, 3) or mod.padding != (1, 1) or mod.stride != (1, 1) or mod.groups != 1
                        or i + 1 >= len(items) or not isinstance(items[i + 1][1], nn.ReLU)):
                    raise NotImplementedError("PerceptualVGG: expected Conv2d(3x3, padding 1) + ReLU pairs (vgg19.features)")
                relu_name = items[i + 1][0]
                if name in self.target_layers:
                    raise NotImplementedError("PerceptualVGG: a tap on a pre-activation is not supported")
                plan.append(("conv", name, self.target_layers.get(relu_name)))
                i += 2
            elif isinstance(mod, nn.MaxPool2d):
                ks = mod.kernel_size if isinstance(mod.kernel_size, tuple) else (mod.kernel_size,) * 2
                st = mod.stride if isinstance(mod.stride, tuple) else (mod.stride,) * 2
                if ks != (2, 2) or st != (2, 2) or name in self.target_layers:
                    raise NotImplementedError("PerceptualVGG: expected MaxPool2d(2, 2) between the conv blocks")
                plan.append(("pool",))
                i += 1
            else:
                raise NotImplementedError(f"PerceptualVGG: unexpected layer {type(mod).__name__} in vgg.features")
        self._plan = plan
        self._prep = None

    # ---- prepared (packed bf16) weights: frozen, built once per device / weight version -----------------------------
    def _prepared(self, device: torch.device):
        convs = [self.vgg_layers._modules[p[1]] for p in self._plan if p[0] == "conv"]
        # (not keyed on modules._WEIGHT_EPOCH: FusedAdam bumps it every step — also inside a captured training step —
        # and never touches these frozen weights; after a ``.data`` write to the VGG call ``vgg._prep = None``)
        key = (device, tuple((c.weight.data_ptr(), c.weight._version) for c in convs),
               tuple((c.bias.data_ptr(), c.bias._version) for c in convs if c.bias is not None))
        if self._prep is None or self._prep[0] != key:
            if device.type != "cuda":
                raise RuntimeError("vunet_b200 has no CPU path: move the module and its inputs to a CUDA device")
            adapters = [_PlainConv(c, device) for c in convs]
            bank = WeightBank([[a] for a in adapters])
            bank.refresh(with_wd=True)
            object.__setattr__(self, "_prep", (key, bank, adapters, None))
        if getattr(self, "validation_fp32", False):
            # fp32 validation mode (``with bdvs_b200.fp32_validation(custom_vgg): ...``): the same schedule with fp32
            # feature maps and fp32 weights on the CUDA-core kernels (vub_conv_fwd_f32 / vub_conv_dgrad_f32 with
            # VUB_CONV_RELU) -- a checker: loss terms AND d loss / d pred within 1e-4 of the reference
            if self._prep[3] is None:
                from .validation import WeightBank32
                bank32 = WeightBank32(slots=[[a] for a in self._prep[2]])
                bank32.refresh()
                object.__setattr__(self, "_prep", self._prep[:3] + (bank32,))
            return self._prep[3], self._prep[2]
        return self._prep[1], self._prep[2]

    @staticmethod
    def _conv_relu(bank, cw, h: torch.Tensor) -> torch.Tensor:
        if bank.precision == "fp32":
            return ops.conv_fwd32([h], cw.w32, cw.scale, cw.shift, cout=cw.cout, ksize=3, want_raw=False, want_elu=True,
                                  relu=True)[1]
        return ops.conv_fwd([h], cw.w, cw.scale, cw.shift, ksize=3, want_raw=False, want_elu=True, relu=True)[1]

    @staticmethod
    def _conv_relu_dgrad(bank, cw, dy, dx, mask=None, addend=None) -> None:
        if bank.precision == "fp32":
            ops.conv_dgrad32(dy, cw.w32, cw.scale, [dx], mask=[mask], addend=[addend], cout=cw.cout, ksize=3, relu=True)
        else:
            ops.conv_dgrad(dy, cw.wd, [dx], mask=[mask], addend=[addend], ksize=3, relu=True)

    def _run_features(self, x: torch.Tensor, keep: bool):
        """``x``: NCHW fp32 in [-1, 1].  Returns (taps: NHWC bf16 feature maps relu1_2 .. relu5_2, records for the
        backward pass: per plan entry (kind, conv weights, input activation, output activation))."""
        bank, adapters = self._prepared(x.device)
        h, _ = ops.vgg_normalize(x.float(), f32=bank.precision == "fp32")
        taps, recs = [], []
        ci = 0
        for entry in self._plan:
            if entry[0] == "conv":
                cw = bank.get(adapters[ci])
                ci += 1
                out = self._conv_relu(bank, cw, h)
                if keep:
                    recs.append(("conv", cw, h, out, entry[2]))
                if entry[2] is not None:
                    taps.append(out)
                h = out
            else:
                out = ops.maxpool2_fwd(h)
                if keep:
                    recs.append(("pool", None, h, out, None))
                h = out
        return taps, recs

    def forward(self, x):
        """The reference's dict: ``input`` (normalised image, NCHW fp32) and the five ReLU taps as NCHW-shaped views of
        the NHWC bf16 feature maps.  Inference only (no autograd history): use :func:`vgg_loss` for training."""
        if not x.is_cuda:
            raise RuntimeError("vunet_b200 has no CPU path: inputs must live on a CUDA device")
        with torch.no_grad():
            _, xn = ops.vgg_normalize(x.float(), want_nhwc=False, want_nchw=True)
            taps, _ = self._run_features(x, keep=False)
        out = {"input": xn}
        for name, t in zip(TAP_NAMES[1:], taps):
            out[name] = t.permute(0, 3, 1, 2)   # NCHW-shaped view of the NHWC feature map (bf16; fp32 in validation mode)
        return out


def _get_member(model, name):
    return getattr(model.module if isinstance(model, nn.DataParallel) else model, name)


class _VggLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, vgg: PerceptualVGG):
        w = [float(x) for x in vgg.loss_weights]
        pred_c = pred.detach().float().contiguous()
        target_c = target.detach().float().contiguous()
        keep = bool(ctx.needs_input_grad[0])  # (forward runs with grad mode off: ask autograd, not the grad mode)
        taps_t, _ = vgg._run_features(target_c, keep=False)
        taps_p, recs = vgg._run_features(pred_c, keep=keep)
        losses = [ops.vgg_input_l1(pred_c, target_c).sum().reshape(1) * (w[0] / pred_c.numel())]
        for i, (pf, tf) in enumerate(zip(taps_p, taps_t)):
            losses.append(ops.l1_partial(pf, tf).sum().reshape(1) * (w[i + 1] / pf.numel()))
        ctx.vgg, ctx.w = vgg, w
        ctx.bank = vgg._prepared(pred_c.device)[0]
        ctx.recs = recs if keep else None
        ctx.taps_t = taps_t if keep else None
        ctx.pred, ctx.target = pred_c, target_c
        ctx.set_materialize_grads(False)
        return tuple(losses)

    @staticmethod
    def backward(ctx, *gs):
        recs, w = ctx.recs, ctx.w
        if recs is None:
            raise RuntimeError("vunet_b200: backward through the same vgg_loss twice is not supported")
        gs = [None if g is None else g.detach().float().contiguous() for g in gs]
        tap_target = {name: t for name, t in zip(TAP_NAMES[1:], ctx.taps_t)}
        tap_index = {name: i for i, name in enumerate(TAP_NAMES)}

        def tap_grad(name: Optional[str], pf: torch.Tensor) -> Optional[torch.Tensor]:
            """Gradient of the tap's L1 term w.r.t. the conv's pre-activation (masked by ReLU'), or None."""
            if name is None or gs[tap_index[name]] is None:
                return None
            i = tap_index[name]
            return ops.l1_relu_grad(pf, tap_target[name], gs[i], w[i] / pf.numel())

        # d_out: gradient w.r.t. the OUTPUT activation of the entry being processed — for a conv already multiplied by
        # ReLU' (i.e. w.r.t. its pre-activation), for a pool the raw gradient of the pooled map
        d_out: Optional[torch.Tensor] = None
        for k in range(len(recs) - 1, -1, -1):
            kind, cw, a_in, a_out, tap = recs[k]
            if kind == "conv":
                if k == len(recs) - 1:
                    d_out = tap_grad(tap, a_out)
                if d_out is None:
                    d_in = None
                else:
                    prev = recs[k - 1] if k > 0 else None
                    dx = torch.empty_like(a_in)
                    if prev is not None and prev[0] == "conv":  # a_in = ReLU output of the previous conv
                        PerceptualVGG._conv_relu_dgrad(ctx.bank, cw, d_out, dx, a_in, tap_grad(prev[4], a_in))
                    else:  # a_in = a pooled map, or the normalised image
                        PerceptualVGG._conv_relu_dgrad(ctx.bank, cw, d_out, dx)
                    d_in = dx
                if d_in is None and k > 0 and recs[k - 1][0] == "conv":
                    d_in = tap_grad(recs[k - 1][4], a_in)
                d_out = d_in
            else:  # pool: a_in = ReLU output of the conv before it (possibly a tap), a_out = the pooled map
                prev = recs[k - 1]
                tg = tap_grad(prev[4], a_in)
                if d_out is None:
                    d_out = tg
                else:
                    d_out = ops.maxpool2_relu_bwd(a_in, d_out, tg)
        g0 = gs[0]
        if d_out is None and g0 is None:
            d_pred = torch.zeros_like(ctx.pred)
        else:
            d_pred = ops.vgg_input_bwd(ctx.pred, ctx.target, d_out, g0, w[0] / ctx.pred.numel())
        ctx.recs = ctx.taps_t = None
        return d_pred, None, None


def vgg_loss(custom_vgg, target, pred, weights=None) -> Dict[str, torch.Tensor]:
    """lib/losses.py:81-119: ``{tap name: loss_weights[i] * mean|target_feature - pred_feature|}`` (each of shape
    ``[1]``), differentiable w.r.t. ``pred``.  ``custom_vgg``: :class:`PerceptualVGG` (or ``nn.DataParallel`` around it)."""
    vgg = custom_vgg.module if isinstance(custom_vgg, nn.DataParallel) else custom_vgg
    if not isinstance(vgg, PerceptualVGG):
        raise TypeError("vgg_loss needs bdvs_b200.PerceptualVGG")
    if weights is not None:
        raise NotImplementedError("vunet_b200: per-pixel `weights` of vgg_loss are not fused (no shipped caller passes them)")
    if not (pred.is_cuda and target.is_cuda):
        raise RuntimeError("vunet_b200 has no CPU path: inputs must live on a CUDA device")
    if tuple(pred.shape) != tuple(target.shape) or pred.dim() != 4 or pred.shape[1] != 3:
        raise ValueError("vgg_loss: pred / target must be NCHW images with 3 channels of the same shape")
    if target.requires_grad:
        raise NotImplementedError("vunet_b200: vgg_loss is differentiable w.r.t. pred only")
    losses = _VggLossFn.apply(pred, target, vgg)
    return {name: l for name, l in zip(TAP_NAMES, losses)}
