"""The shape-and-posture VUnet — host-side mirror of the reference's ``models/vunets.py``
(``VunetAlter`` 426-515, ``EncUp`` 109-148, ``DecUp`` 222-261, ``EncDownAlter`` 520-597,
``DecDownAlter`` 264-414) and of the ``VunetOrg`` family with its 4-group autoregressive prior
(``VunetOrg`` 18-106, ``latent_sample`` 151-156, ``EncDown`` 159-219, ``DecDown`` 600-783).

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


class _VunetBase(nn.Module):
    """What ``VunetAlter`` and ``VunetOrg`` share on the host: the prepared-weight bank of the whole tree (bf16, or
    fp32 in validation mode), the data-parallel hook and cache invalidation."""

    def _tag_owners(self) -> None:
        for owner in ("eu", "ed", "du", "dd"):
            for m in getattr(self, owner).modules():
                if isinstance(m, NormConv2d):
                    m._vub_owner = owner

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

    def invalidate_weights(self) -> None:
        """Call after writing parameters through ``p.data`` (EMA, checkpoint averaging, custom init): autograd's
        version counters do not see such writes, so the prepared (weight-normalised, packed) copies would stay
        stale.  Same as ``bdvs_b200.bump_weight_epoch()``."""
        bump_weight_epoch()


class VunetAlter(_VunetBase):
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
        self._tag_owners()

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


# ======================================================================================================
# VunetOrg family: the original VUnet with the 4-group autoregressive prior (models/vunets.py:18-106, 151-219,
# 600-783; config/vunet.yaml, experiments/vunet.py)
# ======================================================================================================
def _act_dtype(bank) -> torch.dtype:
    return torch.float32 if bank.precision == "fp32" else torch.bfloat16


def _sample(bank, conv, h: Act, eps: Optional[torch.Tensor], *, owner: str, sub: int = 1, group=(0, 0), mean=None,
            z=None, want_act: bool = True, want_grp: bool = False, draw: bool = True):
    """``latent_sample(conv(h))`` (models/vunets.py:151-156): the parameter conv with an fp32 NHWC epilogue, then ONE
    elementwise launch that adds the noise and writes every copy its consumers read — the NCHW fp32 tensors the module
    returns (``mean`` / ``z``: the merged maps of ``__merge_groups`` when ``sub`` is 2) and the NHWC activations
    (raw + ELU) of the next convs.  Returns (mean, z, z_act, grp_elu)."""
    p, slot = conv._run_f32(bank, [(h, "raw")], "nhwc")
    B, hh, ww, c = p.shape
    if eps is None and draw:
        eps = _draw_eps((B, c, hh, ww), p.device)   # the reference's randn_like(mean), same shape and order
    H, W = hh * sub, ww * sub
    if mean is None:
        mean = torch.empty((B, c, H, W), dtype=torch.float32, device=p.device)
    if z is None:
        z = torch.empty((B, c, H, W), dtype=torch.float32, device=p.device)
    z_raw = z_elu = None
    if want_act:
        z_raw = torch.empty((B, H, W, c), dtype=_act_dtype(bank), device=p.device)
        z_elu = torch.empty_like(z_raw)
    grp = ops.latent_sample(p, eps, sub=sub, group=group, mean=mean, z=z, z_raw=z_raw, z_elu=z_elu, want_grp_elu=want_grp)
    z_act = Act(z_raw, z_elu) if want_act else None
    if bank.tape is not None:
        bank.tape.sample(owner, slot, mean, z_act, sub, group)
    return mean, z, z_act, grp, p


class EncDown(nn.Module):
    """models/vunets.py:159-219 — appearance encoder, top-down: posterior means ``q`` and samples ``z = q + eps``."""

    def __init__(self, n_filters, nf_in, subpixel_upsampling, n_scales=2, conv_layer=NormConv2d, dropout_prob=0.0):
        super().__init__()
        self.nin = conv_layer(nf_in, n_filters, kernel_size=1)
        self.n_scales = n_scales
        self.n_rnb = 2
        self.blocks = nn.ModuleList()
        self.ups = nn.ModuleList()
        self.make_latent_params = nn.ModuleList()
        nf = n_filters
        for i in range(self.n_scales):
            for n in range(self.n_rnb // 2):
                self.blocks.append(VunetRNB(channels=nf, a_channels=nf, residual=True, dropout_prob=dropout_prob))
            self.make_latent_params.append(conv_layer(nf, nf, kernel_size=3, padding=1))
            for n in range(self.n_rnb // 2):
                self.blocks.append(VunetRNB(channels=nf, a_channels=2 * nf, residual=True, dropout_prob=dropout_prob))
            self.ups.append(Upsample(nf, nf, subpixel=True))
        self.fin_block = VunetRNB(channels=nf, a_channels=nf, residual=True, dropout_prob=dropout_prob)

    def _run(self, bank, gs: List[Act], eps: Optional[Sequence[torch.Tensor]] = None, skip_dead=False,
             want_q_act=False):
        """``gs`` is consumed from the back like the reference's ``gs.pop()`` calls.  Returns (hs, lat): ``lat[i]`` =
        dict(q, z: NCHW fp32; z_act: the sample as an activation; q_act: the mean as one, when ``want_q_act``)."""
        hs: List[Act] = []
        lat = []
        h = self.nin._run(bank, [(gs[-1], "raw")])
        for i in range(self.n_scales):
            h = self.blocks[2 * i]._run(bank, h, gs.pop())
            hs.append(h)
            q, z, z_act, _, p = _sample(bank, self.make_latent_params[i], h, None if eps is None else eps[i], owner="ed")
            res = {"q": q, "z": z, "z_act": z_act}
            if want_q_act:  # `transfer` feeds the decoder the posterior means: the same launch without noise
                B, hh, ww, c = p.shape
                q_raw = torch.empty((B, hh, ww, c), dtype=_act_dtype(bank), device=p.device)
                q_elu = torch.empty_like(q_raw)
                ops.latent_sample(p, None, z_raw=q_raw, z_elu=q_elu)
                res["q_act"] = Act(q_raw, q_elu)
            lat.append(res)
            last = i + 1 == self.n_scales
            if skip_dead and last:
                # blocks[2i+1], ups[i] and fin_block feed only `hs`, which VunetOrg.forward / transfer drop
                # (models/vunets.py:87, 104): keep the reference's pops, skip the dead launches
                gs.pop()
                gs.pop()
                break
            h = self.blocks[2 * i + 1]._run(bank, h, [gs.pop(), z_act], want_elu=False)
            hs.append(h)
            h = self.ups[i]._run(bank, h)
        else:
            h = self.fin_block._run(bank, h, gs.pop())
            hs.append(h)
        return hs, lat

    def forward(self, gs):
        """Reference signature: pops from the caller's list; returns (hs, qs, zs)."""
        warn_forward_only(self, "EncDown.forward", *gs)
        bank = _solo_bank(self)
        acts = [to_act(g) for g in gs]
        n_before = len(acts)
        hs, lat = self._run(bank, acts)
        del gs[len(gs) - (n_before - len(acts)):]  # mirror the reference's in-place pops
        return [h.nchw() for h in hs], [r["q"] for r in lat], [r["z"] for r in lat]


class DecDown(nn.Module):
    """models/vunets.py:600-783 — image decoder, top-down, with the autoregressive prior over the four
    SpaceToDepth groups of each latent scale.

    Fused schedule of one latent scale (NHWC; group (gi, gj) = the parity sub-grid ``[:, gi::2, gj::2, :]``):
    ``auto_blocks[0]`` -> one regroup launch (SpaceToDepth, raw + ELU copies) -> four times {``auto_lp[l]`` 3x3 on
    4C channels with an fp32 epilogue -> one sampling launch that scatters ``p_l`` / ``z_l`` to their parity positions
    of the merged maps and emits ELU(feedback) -> ``auto_blocks[l+1]``: 1x1 C->4C and the two-source 3x3 on 4C + 4C
    channels} -> ``latent_nins`` as a two-source 1x1 (no concat tensor).  The 4C / 8C-channel convs run on the
    split-K cluster instantiation of the per-tap tcgen05 kernel."""

    def __init__(self, n_scales, nf_in, nf_last, nf_out, subpixel_upsampling, conv_layer=NormConv2d,
                 n_latent_scales=2, dropout_prob=0.0):
        super().__init__()
        self.n_rnb = 2
        self.n_scales = n_scales
        self.n_latent_scales = n_latent_scales
        self.nin = conv_layer(nf_in, nf_in, kernel_size=1)
        self.blocks = nn.ModuleList()
        self.ups = nn.ModuleList()
        self.latent_nins = nn.ModuleDict()
        self.auto_lp = nn.ModuleDict()
        self.auto_blocks = nn.ModuleDict()
        self.out_conv = conv_layer(nf_last, nf_out, kernel_size=3, padding=1)
        self.depth_to_space = DepthToSpace(block_size=2)
        self.space_to_depth = SpaceToDepth(block_size=2)
        nf_h = nf_in
        nf = nf_in
        for i in range(self.n_scales):
            for n in range(self.n_rnb // 2):
                self.blocks.append(VunetRNB(channels=nf, a_channels=nf, residual=True, conv_layer=conv_layer,
                                            dropout_prob=dropout_prob))
            if i < self.n_latent_scales:
                scale = f"l_{i}"
                nin = conv_layer(nf_h * 2, nf_h, kernel_size=1)
                nin._vub_split = 2  # cat[h, z] is read as two sources
                self.latent_nins.update({scale: nin})
                clp = nn.ModuleList()
                cb = nn.ModuleList()
                for l in range(4):
                    clp.append(conv_layer(4 * nf_h, nf_h, kernel_size=3, padding=1))
                    if l == 0:
                        cb.append(VunetRNB(channels=nf_h, dropout_prob=dropout_prob))
                    else:
                        cb.append(VunetRNB(channels=4 * nf_h, a_channels=nf_h, residual=True,
                                           dropout_prob=dropout_prob))
                self.auto_lp.update({scale: clp})
                self.auto_blocks.update({scale: cb})
            for n in range(self.n_rnb // 2):
                self.blocks.append(VunetRNB(channels=nf, a_channels=nf, residual=True, conv_layer=conv_layer,
                                            dropout_prob=dropout_prob))
            if i + 1 < self.n_scales:
                out_c = min(nf_in, nf_last * 2 ** (n_scales - (i + 2)))
                if not subpixel_upsampling and i >= self.n_latent_scales:
                    raise NotImplementedError("bilinear up-sampling is outside the B200 hot path")
                self.ups.append(Upsample(nf, out_c, subpixel=True))
                nf = out_c

    def _run(self, bank, gs: List[Act], zs_posterior: List[Act], training: bool,
             eps_groups: Optional[Sequence[Sequence[torch.Tensor]]] = None, uint8_frames: bool = False,
             want_prior: bool = True):
        """``zs_posterior[i]``: activation (raw + ELU) of what the reference passes as ``zs_posterior[i]`` — the
        posterior sample (``forward``) or mean (``transfer``); ignored when ``training`` is False (each group is then
        sampled from its prior and fed back).  ``eps_groups[i][l]`` optionally injects the group noise.  With
        ``want_prior`` False the noise the reference draws for the (then unused) prior samples is still drawn — the
        global generator stays in step with the reference — but the samples are not computed.
        Returns (imgs, ps, zs_prior) with the merged NCHW fp32 maps of ``__merge_groups``."""
        gs = list(gs)
        zs_posterior = list(zs_posterior)
        ps: List[torch.Tensor] = []
        zs: List[torch.Tensor] = []
        h = self.nin._run(bank, [(gs[-1], "raw")])
        for i in range(self.n_scales):
            B_, hh, ww, cc = h.raw.shape
            h = self.blocks[2 * i]._run(bank, h, gs.pop(), want_elu=not lean_elu(bank, cc, hh, ww, concat=True))
            if i < self.n_latent_scales:
                scale = f"l_{i}"
                dev = h.raw.device
                groups = None
                if training:
                    zpost = zs_posterior[0]
                    # __split_groups: the four parity sub-grids of ELU(z) (every reader of a group is a 1x1 on ELU)
                    gm, _ = ops.space_to_depth(zpost.elu, group_major=True)
                    groups = [Act(None, gm[g]) for g in range(4)]
                    if bank.tape is not None:
                        bank.tape.regroup("dd", zpost, groups, group_major=True)
                pre = self.auto_blocks[scale][0]._run(bank, h, want_elu=False)
                pf_raw, pf_elu = ops.space_to_depth(pre.raw, want_elu=True)
                p_features = Act(pf_raw, pf_elu)
                if bank.tape is not None:
                    bank.tape.regroup("dd", pre, [p_features], group_major=False)
                p_merged = torch.empty((B_, cc, hh, ww), dtype=torch.float32, device=dev)
                z_merged = torch.empty_like(p_merged) if (want_prior or not training) else None
                z_raw_m = None if training else torch.empty((B_, hh, ww, cc), dtype=_act_dtype(bank), device=dev)
                for l in range(4):
                    e = None if eps_groups is None else eps_groups[i][l]
                    conv = self.auto_lp[scale][l]
                    if z_merged is None:  # prior sample not wanted: keep the reference's draw, skip the arithmetic
                        if e is None:
                            _draw_eps((B_, cc, hh // 2, ww // 2), dev)
                        pg, slot = conv._run_f32(bank, [(p_features, "raw")], "nhwc")
                        ops.latent_sample(pg, None, sub=2, group=(l // 2, l % 2), mean=p_merged)
                        grp = None
                        if bank.tape is not None:
                            bank.tape.sample("dd", slot, p_merged, None, 2, (l // 2, l % 2))
                    else:
                        pg, slot = conv._run_f32(bank, [(p_features, "raw")], "nhwc")
                        if e is None:
                            e = _draw_eps((B_, cc, hh // 2, ww // 2), dev)
                        grp = ops.latent_sample(pg, e, sub=2, group=(l // 2, l % 2), mean=p_merged, z=z_merged,
                                                z_raw=z_raw_m, want_grp_elu=not training)
                        if bank.tape is not None:
                            bank.tape.sample("dd", slot, p_merged, None, 2, (l // 2, l % 2))
                    if l + 1 < 4:
                        feedback = groups[l] if training else Act(None, grp)
                        p_features = self.auto_blocks[scale][l + 1]._run(bank, p_features, feedback,
                                                                         want_elu=l + 2 < 4)
                ps.append(p_merged)
                zs.append(z_merged)
                z = zs_posterior.pop(0) if training else Act(z_raw_m, None)
                h = self.latent_nins[scale]._run(bank, [(h, "raw"), (z, "raw")])
            last = i + 1 == self.n_scales
            h = self.blocks[2 * i + 1]._run(bank, h, gs.pop(), want_elu=False)
            if not last:
                co = self.ups[i].up.out_channels // 4
                h = self.ups[i]._run(bank, h, want_elu=not lean_elu(bank, co, 2 * hh, 2 * ww, concat=True))
        assert not gs
        if training:
            assert not zs_posterior
        if uint8_frames and bank.tape is None and bank.precision == "bf16":
            B_, hh, ww, cc = h.raw.shape
            cw = bank.get(self.out_conv)
            if ops.u8_out_ok(1, cc, cw.cout_pad, cw.ksize, 1, hh, ww):
                return self.out_conv._run_f32(bank, [(h, "raw")], "u8_hwc")[0], ps, zs
        img, slot = self.out_conv._run_f32(bank, [(h, "raw")], "nchw")
        if bank.tape is not None:
            bank.tape.img_slot = slot
        return (ops.frames_to_uint8(img) if uint8_frames else img), ps, zs

    def forward(self, gs, zs_posterior, training):
        """Reference signature; returns (imgs, hs, ps, zs_prior) — ``hs`` (the decoder's activations, read by no
        caller of the reference) is returned empty."""
        warn_forward_only(self, "DecDown.forward", *gs, *zs_posterior)
        bank = _solo_bank(self)
        acts = [to_act(g) for g in gs]
        zs = [to_act(z) for z in zs_posterior]
        imgs, ps, zs_prior = self._run(bank, acts, zs, training)
        return imgs, [], ps, zs_prior


class VunetOrg(_VunetBase):
    """models/vunets.py:18-106: ``forward(x, c) -> (imgs, qs, ps, activations)``, ``transfer``, ``test_forward``.
    The KL term of its objective is ``losses.compute_kl_loss(ps, qs)`` (lib/losses.py:55-65)."""

    def __init__(self, init_fn=None, n_channels_x=3, **kwargs):
        super().__init__()
        self.spatial_size = kwargs["spatial_size"]
        self.n_scales = (
            1 + int(np.round(np.log2(kwargs["spatial_size"]))) - kwargs["bottleneck_factor"]
            if kwargs["n_scales"] < 6 else kwargs["n_scales"]
        )
        self.n_scales_x = self.n_scales - kwargs["box_factor"] if n_channels_x > 3 else self.n_scales
        dropout_prob = kwargs["dropout_prob"] if "dropout_prob" in kwargs else 0.0
        self.n_channels_x = n_channels_x
        n_latent_scales = kwargs["n_latent_scales"]
        if kwargs["conv_layer_type"] != "l1":
            # the reference also offers "l2" (L2NormConv2d) here and raises for anything else (models/vunets.py:38-44)
            raise NotImplementedError("vunet_b200 implements conv_layer_type 'l1' (NormConv2d), the type every shipped "
                                      "config uses")
        print("Vunet using L1NormConv2d as conv layers.")
        self.eu = EncUp(self.n_scales_x, n_filters=kwargs["nf_start"], max_filters=kwargs["nf_max"],
                        conv_layer=NormConv2d, nf_in=self.n_channels_x, dropout_prob=dropout_prob)
        self.ed = EncDown(n_filters=kwargs["nf_max"], nf_in=kwargs["nf_max"], conv_layer=NormConv2d,
                          n_scales=n_latent_scales, subpixel_upsampling=kwargs["subpixel_upsampling"],
                          dropout_prob=dropout_prob)
        self.du = DecUp(self.n_scales, n_filters=kwargs["nf_start"], max_filters=kwargs["nf_max"],
                        conv_layer=NormConv2d, dropout_prob=dropout_prob)
        self.dd = DecDown(self.n_scales, kwargs["nf_max"], kwargs["nf_start"], nf_out=3, conv_layer=NormConv2d,
                          n_latent_scales=n_latent_scales, subpixel_upsampling=kwargs["subpixel_upsampling"],
                          dropout_prob=dropout_prob)
        self._tag_owners()

    def _encode_and_pose(self, bank, x, c, eps, want_q_act):
        """``eu`` -> ``ed`` (second stream, small maps) beside ``du`` — see VunetAlter._encode_and_pose."""
        hs = self.eu._run(bank, _as_input(x))
        hs_all = list(hs)
        if os.environ.get("VUB_FWD_STREAMS", "2") == "1" or bank.tape is not None or bank.precision == "fp32":
            _, lat = self.ed._run(bank, hs, eps=eps, skip_dead=True, want_q_act=want_q_act)
            gs = self.du._run(bank, _as_input(c))
            return hs_all, hs, lat, gs
        main = torch.cuda.current_stream()
        side = _SIDE_STREAMS.get(main.device)
        if side is None:
            side = _SIDE_STREAMS[main.device] = torch.cuda.Stream(device=main.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _, lat = self.ed._run(bank, hs, eps=eps, skip_dead=True, want_q_act=want_q_act)
        gs = self.du._run(bank, _as_input(c))
        main.wait_stream(side)
        return hs_all, hs, lat, gs

    def _forward_impl(self, bank, x, c, eps, eps_groups=None):
        _keep, hs_left, lat, gs = self._encode_and_pose(bank, x, c, eps, want_q_act=False)
        imgs, ps, _ = self.dd._run(bank, gs, [r["z_act"] for r in lat], True, eps_groups=eps_groups, want_prior=False)
        return imgs, lat, ps, hs_left, gs

    def forward(self, x, c, eps: Optional[Sequence[torch.Tensor]] = None,
                eps_groups: Optional[Sequence[Sequence[torch.Tensor]]] = None):
        """Returns ``(imgs, qs, ps, (hs, qs, gs, ds))`` like the reference; ``ds`` (the decoder's activations, which
        no caller of the reference reads) is returned empty.  ``eps`` / ``eps_groups`` optionally inject the latent
        noise: one NCHW fp32 tensor per latent scale, and four quarter-resolution ones per latent scale."""
        _require_cuda(x)
        _require_cuda(c)
        if self._wants_grad():
            from .training import train_forward_org
            imgs, qs, ps, hs, gs = train_forward_org(self, x, c, eps, eps_groups)
            return imgs, qs, ps, (hs, qs, gs, [])
        bank = self._bank()
        imgs, lat, ps, hs_left, gs = self._forward_impl(bank, x, c, eps, eps_groups)
        qs = [r["q"] for r in lat]
        return imgs, qs, ps, ([h.nchw() for h in hs_left], qs, [g.nchw() for g in gs], [])

    def test_forward(self, c, eps_groups: Optional[Sequence[Sequence[torch.Tensor]]] = None):
        _require_cuda(c)
        warn_forward_only(self, "VunetOrg.test_forward", c)
        bank = self._bank()
        gs = self.du._run(bank, _as_input(c))
        return self.dd._run(bank, gs, [], False, eps_groups=eps_groups)[0]

    def transfer(self, x, c, eps: Optional[Sequence[torch.Tensor]] = None,
                 eps_groups: Optional[Sequence[Sequence[torch.Tensor]]] = None, uint8_frames: bool = False):
        """``uint8_frames=True`` (extension): ``[B, H, W, 3]`` uint8 video frames, see :meth:`VunetAlter.transfer`."""
        _require_cuda(x)
        _require_cuda(c)
        warn_forward_only(self, "VunetOrg.transfer", x, c)
        bank = self._bank()
        _keep, _, lat, gs = self._encode_and_pose(bank, x, c, eps, want_q_act=True)
        return self.dd._run(bank, gs, [r["q_act"] for r in lat], True, eps_groups=eps_groups,
                            uint8_frames=uint8_frames, want_prior=False)[0]

    # ---- sequence-level split of `transfer`, as in VunetAlter ------------------------------------
    def encode_appearance(self, x, eps: Optional[Sequence[torch.Tensor]] = None):
        """``eu`` + ``ed`` ONCE per appearance image: the posterior means as activations, for :meth:`decode_pose`."""
        _require_cuda(x)
        bank = self._bank()
        hs = self.eu._run(bank, _as_input(x))
        _, lat = self.ed._run(bank, hs, eps=eps, skip_dead=True, want_q_act=True)
        return [r["q_act"] for r in lat]

    def decode_pose(self, c, appearance, frames_per_appearance: int = 1, uint8_frames: bool = False):
        """``du`` + ``dd`` for ``c`` = ``[S * T, 3, H, W]`` stickmen (sequence-major) against ``S`` encoded appearances:
        equals ``transfer(x.repeat_interleave(T, 0), c)`` with one latent-noise draw per sequence."""
        _require_cuda(c)
        bank = self._bank()
        if frames_per_appearance > 1:  # tiny tensors (128 x 4^2 / 8^2 per sequence): plain torch indexing
            appearance = [Act(a.raw.repeat_interleave(frames_per_appearance, dim=0),
                              a.elu.repeat_interleave(frames_per_appearance, dim=0)) for a in appearance]
        if appearance[0].raw.shape[0] != c.shape[0]:
            raise ValueError("decode_pose: c must hold frames_per_appearance frames for every appearance image")
        gs = self.du._run(bank, _as_input(c))
        return self.dd._run(bank, gs, list(appearance), True, uint8_frames=uint8_frames, want_prior=False)[0]
