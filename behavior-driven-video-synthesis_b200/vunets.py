"""The shape-and-posture VUnet — host-side mirror of the reference's ``models/vunets.py``
(``VunetAlter`` 426-515, ``EncUp`` 109-148, ``DecUp`` 222-261, ``EncDownAlter`` 520-597,
``DecDownAlter`` 264-414).

Same constructor (``VunetAlter(init_fn=None, n_channels_x=3, **kwargs)``, unknown kwargs ignored), same
``forward(x, c)`` / ``transfer(x, c)`` / ``test_forward(c)`` signatures and return structure, same
sub-module names and ``state_dict`` layout, same latent-noise draws (``torch.randn`` of the reference's
shapes from the global generator, in the reference's order).  The arithmetic is the fused schedule of
SURVEY.md App. B executed by sm_100a kernels; there is no torch conv and no CPU path.
"""
from __future__ import annotations

import os

from typing import List, Optional, Sequence

import numpy as np
import torch
from torch import nn

from . import ops
from .modules import (lean_elu, Act, DepthToSpace, Downsample, NormConv2d, SpaceToDepth, Upsample, VunetRNB,
                      WeightBank, _solo_bank, bump_weight_epoch, to_act, warn_forward_only)


def _draw_eps(shape, device) -> torch.Tensor:
    """Latent noise exactly as the reference draws it: ``torch.randn_like(std)`` on an NCHW fp32
    tensor of this shape from the global generator (models/vunets.py:596, 352)."""
    return torch.randn(shape, dtype=torch.float32, device=device)


class EncUp(nn.Module):
    """models/vunets.py:109-148 — appearance encoder, bottom-up."""

    def __init__(self, n_scales, n_filters, max_filters, nf_in=3, conv_layer=NormConv2d, dropout_prob=0.0):
        super().__init__()
        self.n_rnb = 2
        self.n_scales = n_scales
        self.nin = conv_layer(in_channels=nf_in, out_channels=n_filters, kernel_size=1)
        self.nin._vub_stem = True
        self.blocks = nn.ModuleList()
        self.downs = nn.ModuleList()
        nf = n_filters
        for i in range(self.n_scales):
            for n in range(self.n_rnb):
                self.blocks.append(VunetRNB(channels=nf, conv_layer=conv_layer, dropout_prob=dropout_prob))
            if i + 1 < self.n_scales:
                out_c = min(2 * nf, max_filters)
                self.downs.append(Downsample(nf, out_c))
                nf = out_c

    def _run(self, bank: WeightBank, x: torch.Tensor) -> List[Act]:
        hs: List[Act] = []
        # ELU copies are skipped where every reader (the next block's 3x3, a skip connection's 1x1) applies ELU on load
        hh, ww = (x.shape[1], x.shape[2]) if x.dtype == torch.uint8 else (x.shape[2], x.shape[3])
        h = self.nin._run_stem(x, bank=bank, want_elu=not lean_elu(bank, self.nin.out_channels, hh, ww))
        for i in range(self.n_scales):
            for n in range(self.n_rnb):
                blk = self.blocks[2 * i + n]
                h = blk._run(bank, h, want_elu=not lean_elu(bank, blk.conv.out_channels, hh, ww))
                hs.append(h)
            if i + 1 < self.n_scales:
                hh, ww = hh // 2, ww // 2
                h = self.downs[i]._run(bank, h, want_elu=not lean_elu(bank, self.downs[i].down.out_channels, hh, ww))
        return hs

    def forward(self, x, **kwargs):
        warn_forward_only(self, type(self).__name__ + ".forward", x)
        bank = _solo_bank(self)
        _require_cuda(x)
        return [h.nchw() for h in self._run(bank, _as_input(x))]


class DecUp(EncUp):
    """models/vunets.py:222-261 — pose (stickman) encoder; identical structure to :class:`EncUp`."""

    def forward(self, c):
        return super().forward(c)


class EncDownAlter(nn.Module):
    """models/vunets.py:520-597 — appearance encoder, top-down, with the per-scale latent heads."""

    def __init__(self, n_filters, nf_in, subpixel_upsampling, n_scales=2, conv_layer=NormConv2d,
                 dropout_prob=0.0):
        super().__init__()
        self.nin = conv_layer(nf_in, n_filters, kernel_size=1)
        self.n_scales = n_scales
        self.n_rnb = 2
        self.blocks = nn.ModuleList()
        self.ups = nn.ModuleList()
        self.make_latent_params = nn.ModuleList()
        self.make_logstds = nn.ModuleList()
        nf = n_filters
        for i in range(self.n_scales):
            for n in range(self.n_rnb // 2):
                self.blocks.append(VunetRNB(channels=nf, a_channels=nf, residual=True, dropout_prob=dropout_prob))
            self.make_latent_params.append(conv_layer(nf, nf, kernel_size=3, padding=1))
            self.make_logstds.append(conv_layer(nf, nf, kernel_size=3, padding=1))
            for n in range(self.n_rnb // 2):
                self.blocks.append(VunetRNB(channels=nf, a_channels=2 * nf, residual=True))
            self.ups.append(Upsample(nf, nf, subpixel=True))
        self.fin_block = VunetRNB(channels=nf, a_channels=nf, residual=True, dropout_prob=dropout_prob)
        self.squash = nn.Sigmoid()
        # mu and logstd convs read the same activation: one GEMM with N = 2*nf
        self._vub_merged_pairs = [(self.make_latent_params[i], self.make_logstds[i]) for i in range(n_scales)]

    def _latent(self, bank, i, h: Act, eps: Optional[torch.Tensor], want_elu_mu: bool):
        ml, ml_slot = self.make_latent_params[i]._run_f32(bank, [(h, "raw")], "nhwc")  # (mu || logit), one GEMM
        B, hh, ww, c2 = ml.shape
        if eps is None:
            eps = _draw_eps((B, c2 // 2, hh, ww), ml.device)
        if bank.precision == "fp32":  # validation mode: fp32 ELU(z) / ELU(mu) operands
            res = ops.latent_fwd(ml, eps, want_elu_z=False, want_elu_mu=False)
            res["z_act"] = Act(None, ops.nchw_to_nhwc32(res["z"], False, True)[1])
            if want_elu_mu:
                res["mu_act"] = Act(None, ops.nchw_to_nhwc32(res["mu"], False, True)[1])
            if bank.tape is not None:
                bank.tape.latent(ml_slot, res, eps)
            return res
        res = ops.latent_fwd(ml, eps, want_elu_z=True, want_elu_mu=want_elu_mu)
        res["z_act"] = Act(None, res["elu_z"])
        if want_elu_mu:
            res["mu_act"] = Act(None, res["elu_mu"])
        if bank.tape is not None:
            bank.tape.latent(ml_slot, res, eps)
        return res

    def _run(self, bank, gs: List[Act], eps: Optional[Sequence[torch.Tensor]] = None, skip_dead=False,
             want_elu_mu=False):
        """``gs`` is consumed from the back exactly like the reference's ``gs.pop()`` calls.
        Returns (hs, latents) where latents[i] is the dict from ``ops.latent_fwd``."""
        hs: List[Act] = []
        lat = []
        h = self.nin._run(bank, [(gs[-1], "raw")])
        for i in range(self.n_scales):
            h = self.blocks[2 * i]._run(bank, h, gs.pop())
            hs.append(h)
            res = self._latent(bank, i, h, None if eps is None else eps[i], want_elu_mu)
            lat.append(res)
            last = i + 1 == self.n_scales
            if skip_dead and last:
                # the reference computes blocks[2i+1], ups[i] and fin_block here but nothing that
                # VunetAlter.forward/transfer return depends on them; keep its pops
                gs.pop()
                gs.pop()
                break
            h = self.blocks[2 * i + 1]._run(bank, h, [gs.pop(), res["z_act"]], want_elu=False)
            hs.append(h)
            h = self.ups[i]._run(bank, h)
        else:
            h = self.fin_block._run(bank, h, gs.pop())
            hs.append(h)
        return hs, lat

    def forward(self, gs):
        """Reference signature: pops from the caller's list; returns (hs, means, log_stds, zs)."""
        warn_forward_only(self, "EncDownAlter.forward", *gs)
        bank = _solo_bank(self)
        acts = [to_act(g) for g in gs]
        n_before = len(acts)
        hs, lat = self._run(bank, acts)
        del gs[len(gs) - (n_before - len(acts)):]  # mirror the reference's in-place pops
        means = [_tag_kl(r["mu"], r) for r in lat]
        return [h.nchw() for h in hs], means, [r["logstd"] for r in lat], [r["z"] for r in lat]


class DecDownAlter(nn.Module):
    """models/vunets.py:264-414 — image decoder, top-down, skip connections from :class:`DecUp`."""

    def __init__(self, n_scales, nf_in, nf_last, nf_out, subpixel_upsampling, conv_layer=NormConv2d,
                 n_latent_scales=2, dropout_prob=0.0):
        super().__init__()
        self.n_rnb = 2
        self.n_scales = n_scales
        self.n_latent_scales = n_latent_scales
        self.nin = conv_layer(nf_in, nf_in, kernel_size=1)
        self.blocks = nn.ModuleList()
        self.ups = nn.ModuleList()
        self.auto_blocks = nn.ModuleList()
        self.out_conv = conv_layer(nf_last, nf_out, kernel_size=3, padding=1)
        self.depth_to_space = DepthToSpace(block_size=2)
        self.space_to_depth = SpaceToDepth(block_size=2)
        nf = nf_in
        for i in range(self.n_scales):
            for n in range(self.n_rnb // 2):
                self.blocks.append(VunetRNB(channels=nf, a_channels=nf, residual=True, conv_layer=conv_layer,
                                            dropout_prob=dropout_prob))
            if i < self.n_latent_scales:
                self.auto_blocks.append(VunetRNB(channels=nf, a_channels=nf, residual=True, conv_layer=conv_layer,
                                                 dropout_prob=dropout_prob))
            for n in range(self.n_rnb // 2):
                self.blocks.append(VunetRNB(channels=nf, a_channels=nf, residual=True, conv_layer=conv_layer,
                                            dropout_prob=dropout_prob))
            if i + 1 < self.n_scales:
                out_c = min(nf_in, nf_last * 2 ** (n_scales - (i + 2)))
                if not subpixel_upsampling and i >= self.n_latent_scales:
                    raise NotImplementedError("bilinear up-sampling is outside the B200 hot path")
                self.ups.append(Upsample(nf, out_c, subpixel=True))
                nf = out_c

    def _run(self, bank, gs: List[Act], zs: List[Act], training: bool, uint8_frames: bool = False) -> torch.Tensor:
        """``zs[i]``: activation holding ELU(z_i) (posterior sample or mean); ignored when ``training`` is
        False (then z ~ N(0, I) is drawn like the reference, models/vunets.py:352).  ``uint8_frames``: return uint8
        HWC video frames, the reference renderer's ``(scale_img(img) * 255).astype(uint8)`` — stored by ``out_conv``'s
        epilogue where the kernel supports it, else converted from the fp32 image by one more kernel."""
        gs = list(gs)
        zs = list(zs)
        h = self.nin._run(bank, [(gs[-1], "raw")])
        lat_count = 0
        for i in range(self.n_scales):
            B_, hh, ww, cc = h.raw.shape
            h = self.blocks[2 * i]._run(bank, h, gs.pop(), want_elu=not lean_elu(bank, cc, hh, ww, concat=True))
            if i < self.n_latent_scales:
                if training:
                    z = zs.pop(0)
                else:
                    B, hh, ww, cc = h.raw.shape
                    conv = ops.nchw_to_nhwc32 if bank.precision == "fp32" else ops.nchw_to_nhwc
                    z = Act(None, conv(_draw_eps((B, cc, hh, ww), h.raw.device), False, True)[1])
                h = self.auto_blocks[lat_count]._run(bank, h, z)
                lat_count += 1
            last = i + 1 == self.n_scales
            h = self.blocks[2 * i + 1]._run(bank, h, gs.pop(), want_elu=False)
            if not last:
                co = self.ups[i].up.out_channels // 4
                h = self.ups[i]._run(bank, h, want_elu=not lean_elu(bank, co, 2 * hh, 2 * ww, concat=True))
        assert not gs
        if training:
            assert not zs
        if uint8_frames and bank.tape is None and bank.precision == "bf16":
            B_, hh, ww, cc = h.raw.shape
            cw = bank.get(self.out_conv)
            if ops.u8_out_ok(1, cc, cw.cout_pad, cw.ksize, 1, hh, ww):
                return self.out_conv._run_f32(bank, [(h, "raw")], "u8_hwc")[0]
        img, slot = self.out_conv._run_f32(bank, [(h, "raw")], "nchw")
        if bank.tape is not None:
            bank.tape.img_slot = slot
        return ops.frames_to_uint8(img) if uint8_frames else img

    def forward(self, gs, zs_posterior, training):
        warn_forward_only(self, "DecDownAlter.forward", *gs, *zs_posterior)
        bank = _solo_bank(self)
        acts = [to_act(g) for g in gs]
        zs = [to_act(z, need_raw=False, need_elu=True) for z in zs_posterior]
        return self._run(bank, acts, zs, training)


_SIDE_STREAMS: dict = {}  # one side stream per device (not a module attribute: models are deep-copied / pickled)


def _as_input(t: torch.Tensor) -> torch.Tensor:
    """NCHW float tensors like the reference's ([-1, 1]) — or, for the rendering front-end, uint8 HWC image batches
    ``[B, H, W, C]`` exactly as the reference holds them BEFORE ``ToTensor()`` and ``x * 2 - 1``
    (data/__init__.py:15-43): the stem kernel applies that conversion itself, bit-identically."""
    return t if t.dtype == torch.uint8 else t.float()


def _require_cuda(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError("vunet_b200 has no CPU path: inputs and the module must live on a CUDA device")


def _tag_kl(mu: torch.Tensor, res: dict) -> torch.Tensor:
    """Attach the fused KL partial sums to the returned mean so losses.compute_kl_with_prior can use them."""
    mu._vub_kl_partial = res["kl_partial"]
    mu._vub_logstd = res["logstd"]
    return mu


class VunetAlter(nn.Module):
    """models/vunets.py:426-515."""

    def __init__(self, init_fn=None, n_channels_x=3, **kwargs):
        super().__init__()
        self.spatial_size = kwargs["spatial_size"]
        self.n_scales = (
            1 + int(np.round(np.log2(kwargs["spatial_size"]))) - kwargs["bottleneck_factor"]
            if kwargs["n_scales"] < 6 else kwargs["n_scales"]
        )
        self.n_scales_x = self.n_scales - kwargs["box_factor"] if n_channels_x > 3 else self.n_scales
        self.n_channels_x = n_channels_x
        n_latent_scales = kwargs["n_latent_scales"]
        dropout_prob = kwargs["dropout_prob"] if "dropout_prob" in kwargs else 0.0
        if kwargs["conv_layer_type"] != "l1":
            raise NotImplementedError(
                "vunet_b200 implements conv_layer_type 'l1' (NormConv2d), the type every shipped config uses")
        print("Vunet using L1NormConv2d as conv layers.")
        self.eu = EncUp(self.n_scales_x, n_filters=kwargs["nf_start"], max_filters=kwargs["nf_max"],
                        conv_layer=NormConv2d, nf_in=n_channels_x, dropout_prob=dropout_prob)
        self.ed = EncDownAlter(n_filters=kwargs["nf_max"], nf_in=kwargs["nf_max"], conv_layer=NormConv2d,
                               n_scales=n_latent_scales, subpixel_upsampling=kwargs["subpixel_upsampling"],
                               dropout_prob=dropout_prob)
        self.du = DecUp(self.n_scales, n_filters=kwargs["nf_start"], max_filters=kwargs["nf_max"],
                        conv_layer=NormConv2d, dropout_prob=dropout_prob)
        self.dd = DecDownAlter(self.n_scales, kwargs["nf_max"], kwargs["nf_start"], nf_out=3,
                               conv_layer=NormConv2d, n_latent_scales=n_latent_scales,
                               subpixel_upsampling=kwargs["subpixel_upsampling"], dropout_prob=dropout_prob)
        for owner in ("eu", "ed", "du", "dd"):
            for m in getattr(self, owner).modules():
                if isinstance(m, NormConv2d):
                    m._vub_owner = owner

    # ---------------------------------------------------------------------------------------------
    def _bank(self, with_wd: bool = False) -> WeightBank:
        if getattr(self, "validation_fp32", False):
            bank32 = getattr(self, "_vub_bank32", None)
            if bank32 is None:
                from .validation import WeightBank32
                bank32 = WeightBank32(self)
                object.__setattr__(self, "_vub_bank32", bank32)
            bank32.refresh()
            return bank32
        bank = _solo_bank(self)
        if with_wd and not bank._with_wd:
            bank.refresh(with_wd=True)
        return bank

    def enable_data_parallel(self, process_group=None):
        """Average parameter gradients over the ranks of ``process_group`` (default: the world) with one
        all-reduce per sub-network bucket, overlapped with the backward pass.  One process per GPU; call
        after ``torch.distributed.init_process_group``.  Replaces ``nn.DataParallel(vunet)``
        (experiments/shape_and_pose_net.py:213-214).  Returns ``self``."""
        from .training import BucketReducer
        object.__setattr__(self, "_vub_reducer", BucketReducer(process_group))
        return self

    def _wants_grad(self) -> bool:
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def _forward_impl(self, bank, x, c, eps):
        _keep, hs_left, lat, gs = self._encode_and_pose(bank, x, c, eps, want_elu_mu=False)
        imgs = self.dd._run(bank, gs, [r["z_act"] for r in lat], True)
        return imgs, lat, hs_left

    def _encode(self, bank, x, eps, skip_dead, want_elu_mu):
        hs = self.eu._run(bank, _as_input(x))
        hs_all = list(hs)
        es, lat = self.ed._run(bank, hs, eps=eps, skip_dead=skip_dead, want_elu_mu=want_elu_mu)
        return hs_all, hs, lat

    def _encode_and_pose(self, bank, x, c, eps, want_elu_mu):
        """``eu`` -> ``ed`` and ``du``.  ``ed`` (a chain of ~20 kernels on 4x4 .. 16x16 maps, each a few microseconds
        on a fraction of the SMs) depends only on ``eu``; it runs on a second stream beside ``du``, whose large
        kernels hide it.  VUB_FWD_STREAMS=1 keeps one stream."""
        hs = self.eu._run(bank, _as_input(x))
        hs_all = list(hs)  # keeps eu's activations alive until the side stream has read them (ed pops from `hs`)
        if os.environ.get("VUB_FWD_STREAMS", "2") == "1":
            _, lat = self.ed._run(bank, hs, eps=eps, skip_dead=True, want_elu_mu=want_elu_mu)
            gs = self.du._run(bank, _as_input(c))
            return hs_all, hs, lat, gs
        main = torch.cuda.current_stream()
        side = _SIDE_STREAMS.get(main.device)
        if side is None:
            side = _SIDE_STREAMS[main.device] = torch.cuda.Stream(device=main.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _, lat = self.ed._run(bank, hs, eps=eps, skip_dead=True, want_elu_mu=want_elu_mu)
        gs = self.du._run(bank, _as_input(c))
        main.wait_stream(side)
        return hs_all, hs, lat, gs

    def forward(self, x, c, eps: Optional[Sequence[torch.Tensor]] = None):
        """Returns ``(imgs, means, logstds, (hs, means, logstds))`` like the reference.  ``eps`` optionally
        injects the latent noise (one NCHW fp32 tensor per latent scale) instead of drawing it."""
        _require_cuda(x)
        _require_cuda(c)
        if self._wants_grad():
            from .training import train_forward
            imgs, means, logstds, hs = train_forward(self, x, c, eps)
            return imgs, means, logstds, (hs, means, logstds)
        bank = self._bank()
        imgs, lat, hs_left = self._forward_impl(bank, x, c, eps)
        means = [_tag_kl(r["mu"], r) for r in lat]
        logstds = [r["logstd"] for r in lat]
        # like the reference, `hs` is what ed's pops left of eu's outputs
        activations = ([h.nchw() for h in hs_left], means, logstds)
        return imgs, means, logstds, activations

    def invalidate_weights(self) -> None:
        """Call after writing parameters through ``p.data`` (EMA, checkpoint averaging, custom init): autograd's
        version counters do not see such writes, so the prepared (weight-normalised, packed) copies would stay
        stale.  Same as ``bdvs_b200.bump_weight_epoch()``."""
        bump_weight_epoch()

    def test_forward(self, c):
        _require_cuda(c)
        warn_forward_only(self, "VunetAlter.test_forward", c)
        bank = self._bank()
        gs = self.du._run(bank, _as_input(c))
        return self.dd._run(bank, gs, [], False)

    # ---- sequence-level split of `transfer` ----------------------------------------
    def encode_appearance(self, x, eps: Optional[Sequence[torch.Tensor]] = None):
        """The appearance half of :meth:`transfer` — ``eu`` + ``ed`` — run ONCE per appearance image.
        The reference re-encodes the (constant) appearance image for every frame of a sequence
        (``data/data_conversions_3d.py:1130-1165``); the posterior means are all the decoder needs.
        Returns an opaque list to hand to :meth:`decode_pose`."""
        _require_cuda(x)
        bank = self._bank()
        _, _, lat = self._encode(bank, x, eps, skip_dead=True, want_elu_mu=True)
        return [r["mu_act"] for r in lat]

    def decode_pose(self, c, appearance, frames_per_appearance: int = 1, uint8_frames: bool = False):
        """The pose half of :meth:`transfer` — ``du`` + ``dd`` — for ``c`` = ``[S * T, 3, H, W]`` stickmen of ``S``
        sequences with ``T = frames_per_appearance`` frames each (sequence-major), ``appearance`` from
        :meth:`encode_appearance` on the ``S`` appearance images.  Equals
        ``transfer(x.repeat_interleave(T, 0), c)`` with one latent-noise draw per sequence instead of per frame."""
        _require_cuda(c)
        bank = self._bank()
        if frames_per_appearance > 1:  # tiny tensors (128 x 4^2 / 8^2 per sequence): plain torch indexing
            appearance = [Act(None, a.elu.repeat_interleave(frames_per_appearance, dim=0)) for a in appearance]
        if appearance[0].elu.shape[0] != c.shape[0]:
            raise ValueError("decode_pose: c must hold frames_per_appearance frames for every appearance image")
        gs = self.du._run(bank, _as_input(c))
        return self.dd._run(bank, gs, list(appearance), True, uint8_frames=uint8_frames)

    def transfer(self, x, c, eps: Optional[Sequence[torch.Tensor]] = None, uint8_frames: bool = False):
        """``uint8_frames=True`` (extension): return ``[B, H, W, 3]`` uint8 video frames instead of the fp32 NCHW image —
        bit-identical to ``(scale_img(transfer(x, c)) * 255).permute(0, 2, 3, 1).astype(uint8)`` of the reference's
        rendering loop (data/data_conversions_3d.py:1166-1173), written directly by ``out_conv``'s epilogue."""
        _require_cuda(x)
        _require_cuda(c)
        warn_forward_only(self, "VunetAlter.transfer", x, c)
        bank = self._bank()
        _keep, _, lat, gs = self._encode_and_pose(bank, x, c, eps, want_elu_mu=True)
        return self.dd._run(bank, gs, [r["mu_act"] for r in lat], True, uint8_frames=uint8_frames)
