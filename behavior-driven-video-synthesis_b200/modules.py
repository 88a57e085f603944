"""Conv building blocks of the VUnet — host-side mirror of the reference's ``lib/modules.py:11-233``.

Same class names, constructor arguments, parameter names/shapes/registration order (so
``state_dict()`` round-trips with reference checkpoints) and the same initialisation stream; the
arithmetic runs in the sm_100a kernels behind ``torch.ops.vunet_b200`` (no torch conv, no CPU path).

Public ``forward`` methods take/return NCHW tensors like the reference.  Internally an activation is
an :class:`Act` (NHWC bf16, raw and/or ELU'd copy); tensors returned by ``forward`` are NCHW *views*
of the raw copy that carry their ``Act`` so that chained blocks never convert layouts.
"""
from __future__ import annotations

import os
import ctypes as C
import math
from typing import Iterable, List, Optional, Sequence, Tuple, Union

import torch
from torch import nn

from . import _native as N
from . import ops

_SUPPORTED_C = (32, 64, 128, 256, 512)  # 256 / 512: the 4-group prior of VunetOrg (smallest maps, stride 1)


class Act:
    """One activation of the path, NHWC bf16: ``raw`` = as produced, ``elu`` = ELU(raw)."""

    __slots__ = ("raw", "elu", "grad", "grad_view")

    def __init__(self, raw: Optional[torch.Tensor], elu: Optional[torch.Tensor]):
        self.raw = raw
        self.elu = elu
        self.grad = None  # training: NHWC bf16 gradient w.r.t. the raw value, accumulated by the backward schedule
        self.grad_view = None  # training: storage the gradient must be created in (a slice of a stacked buffer)

    def pick(self, mode: str) -> torch.Tensor:
        """The copy a consumer reads: ``"raw"`` or ``"elu"``."""
        t = self.raw if mode == "raw" else self.elu
        if t is None:
            raise RuntimeError(f"activation has no {mode} copy")
        return t

    @property
    def any(self) -> torch.Tensor:
        return self.raw if self.raw is not None else self.elu

    def nchw(self) -> torch.Tensor:
        """NCHW-shaped (channels-last strided) view of the raw copy that remembers this Act."""
        if self.raw is None:
            raise RuntimeError("activation has no raw copy")
        t = self.raw.permute(0, 3, 1, 2)
        t._vub_act = self
        return t


def to_act(t: Union[torch.Tensor, Act], need_raw: bool = True, need_elu: bool = True) -> Act:
    """Accepts an Act, a tensor produced by this package, or any NCHW tensor (converted on the GPU)."""
    if isinstance(t, Act):
        act = t
    else:
        act = getattr(t, "_vub_act", None)
    if act is not None and (act.raw is not None or not need_raw) and (act.elu is not None or not need_elu):
        return act
    if act is not None and act.raw is not None:
        t = ops.nhwc_to_nchw(act.raw)
    if not t.is_cuda:
        raise RuntimeError("vunet_b200 has no CPU path: inputs must live on a CUDA device")
    raw, elu = ops.nchw_to_nhwc(t.float(), need_raw, need_elu)
    return Act(raw, elu)


def lean_elu(bank, c: int, h: int, w: int, concat: bool = False) -> bool:
    """True when the ELU copy of a ``[*, h, w, c]`` activation need not be stored: the convs that read ELU(t) in
    the VUnet — a residual block's 3x3 (``concat``: as the first half of the block's two-source 3x3) and a skip
    connection's 1x1 — all run on the row-streaming kernel at this shape, which applies ELU to its input rows in
    shared memory.  Halves the HBM traffic of the wide HBM-bound layers (inference only)."""
    if not getattr(bank, "lean", False):
        return False
    if concat:
        return ops.elu_in_ok(2, c, c, 3, 1, h, w)
    return ops.elu_in_ok(1, c, c, 3, 1, h, w) and ops.elu_in_ok(1, c, c, 1, 1, h, w)


# ------------------------------------------------------------------------------------------------
# Prepared (weight-normalised, packed) weights
# ------------------------------------------------------------------------------------------------
class ConvWeights:
    """Packed bf16 ``v/||v||`` [cout_pad, K] + fp32 scale/shift of one (possibly merged) conv."""

    __slots__ = ("w", "scale", "shift", "cout", "cout_pad", "nsrc", "c_src", "ksize", "stride", "wd", "cout_k",
                 "slot", "w_tiled", "wd_tiled")

    def __init__(self, w, scale, shift, cout, cout_pad, nsrc, c_src, ksize, stride, wd=None, cout_k=0, slot=-1,
                 w_tiled=None, wd_tiled=None):
        # tile-major pre-swizzled copies for the tensor-core-bound layers (1-D bulk weight loads), or None
        self.w_tiled, self.wd_tiled = w_tiled, wd_tiled
        self.w, self.scale, self.shift = w, scale, shift
        self.cout, self.cout_pad = cout, cout_pad
        self.nsrc, self.c_src, self.ksize, self.stride = nsrc, c_src, ksize, stride
        # training: data-gradient weights [nsrc*c_src, k*k*cout_k] (bf16, pre-scaled), cout_k = cout padded to 32
        self.wd, self.cout_k, self.slot = wd, cout_k, slot


def _pad_cout(c: int) -> int:
    for p in (16, 32, 64, 128, 256, 512):
        if c <= p:
            return p
    raise NotImplementedError(f"no kernel for {c} output channels")


# Parameters can also change through raw device pointers (optim.FusedAdam), which autograd version counters do
# not see: such writers bump this epoch, and every bank's cache key includes it.
_WEIGHT_EPOCH = [0]


def bump_weight_epoch() -> int:
    """Invalidate every prepared-weight cache of the process.  REQUIRED after writing parameters in a way autograd's
    version counters do not see — ``p.data.copy_()`` / ``p.data.mul_()`` (EMA, checkpoint averaging, custom init),
    raw-pointer writers — because ``weights_key`` cannot notice those (the reference re-normalises the weights in
    every forward, so the idiom works there).  ``VunetAlter.invalidate_weights()`` is the per-model spelling."""
    _WEIGHT_EPOCH[0] += 1
    return _WEIGHT_EPOCH[0]


def weights_key(params) -> tuple:
    return (_WEIGHT_EPOCH[0],) + tuple(p._version for p in params)


class WeightBank:
    """All tensor-core convs of a module tree, prepared by ONE ``wn_prep`` launch per refresh.

    ``slots`` is a list of lists of :class:`NormConv2d`; the members of a slot share their input and
    are packed one after another along the output-channel axis (the latent mu/logstd pair).
    """

    def __init__(self, slots: Sequence[Sequence["NormConv2d"]]):
        self.slots = [list(s) for s in slots]
        self._key = None
        self._device = None
        self._views = {}
        self._with_wd = False
        self.generation = 0  # bumped when the forward buffers are re-allocated (graphs captured earlier are stale)
        self.tape = None  # training: list the forward pass appends its records to (training.Tape)
        self.precision = "bf16"  # validation.WeightBank32 says "fp32"

    @property
    def lean(self) -> bool:
        """Inference in bf16 (no tape): producers skip the ELU copy of an activation whose consumers can all
        apply ELU while loading (``lean_elu``).  The row-streaming kernel itself is no faster for it (it is bound
        by the shared-memory data pipe, and the in-smem transform adds to that), but the producers write half as
        much: -0.36 ms per 64-frame H256 batch (stems, stride-2 and up-sampling convs).
        ``VUB_LEAN_ACTS=0`` restores (raw, elu) pairs everywhere."""
        return self.tape is None and self.precision == "bf16" and os.environ.get("VUB_LEAN_ACTS", "1") != "0"

    def __deepcopy__(self, memo):
        # a cache of device pointers into ONE module tree: copy.deepcopy(model) starts without it and rebuilds lazily
        return None

    @staticmethod
    def for_module(root: nn.Module) -> "WeightBank":
        merged_second = {}
        for m in root.modules():
            for a, b in getattr(m, "_vub_merged_pairs", ()):
                merged_second[id(b)] = a
        slots, index = [], {}
        for m in root.modules():
            if isinstance(m, NormConv2d) and not m._vub_stem and id(m) not in merged_second:
                index[id(m)] = len(slots)
                slots.append([m])
        for m in root.modules():
            if isinstance(m, NormConv2d) and id(m) in merged_second:
                slots[index[id(merged_second[id(m)])]].append(m)
        return WeightBank(slots)

    # -- allocation -------------------------------------------------------------------------------
    def _layout(self):
        w_off, s_off, wd_off = 0, 0, 0
        layout = []
        for slot in self.slots:
            first = slot[0]
            k = first.in_channels * first.kernel_size ** 2
            cout = sum(m.out_channels for m in slot)
            for m in slot:
                if m.in_channels * m.kernel_size ** 2 != k or m._vub_split != first._vub_split:
                    raise ValueError("merged convs must share their input geometry")
            nsrc = first._vub_split
            c_src = first.in_channels // nsrc
            linear = getattr(first, "_vub_linear", False)  # nn.Linear as a 1x1 conv on a 1x1 map (behavior.py)
            if linear and (c_src % 64 or c_src > 2048):
                raise NotImplementedError(f"linear layers take a multiple of 64 (<= 2048) input features, got {c_src}")
            if not linear and c_src not in _SUPPORTED_C:
                raise NotImplementedError(
                    f"no tensor-core kernel for {first.in_channels} input channels split {nsrc} ways "
                    f"(supported per-source channel counts: {_SUPPORTED_C})")
            cpad = _pad_cout(cout) if cout <= 512 else (cout + 127) // 128 * 128
            cout_k = (cout + 31) // 32 * 32
            layout.append((w_off, s_off, k, cout, cpad, nsrc, c_src, wd_off, cout_k))
            w_off += cpad * k
            s_off += cpad
            wd_off += k * cout_k
        return layout, w_off, s_off, wd_off

    def _allocate(self, device: torch.device, with_wd: bool) -> None:
        """(Re)build the device table of ``wn_prep`` entries.  The packed-weight buffers a captured CUDA graph may
        have baked in (``_w``, ``_scale``, ``_shift``, tile-major copies) depend only on the layer shapes: they are
        allocated ONCE per device and survive parameter-pointer changes and the first training step; the
        data-gradient matrices (``_wd``) are a separate allocation made when training first asks for them.
        ``generation`` counts re-allocations of the forward buffers (device moves only) — graph holders compare it."""
        layout, w_off, s_off, wd_off = self._layout()
        tiled = os.environ.get("VUB_TILED_WEIGHTS", "0") == "1"
        if self._device != device or getattr(self, "_tiled", None) != tiled:
            self._w = torch.zeros(w_off, dtype=torch.bfloat16, device=device)
            # Tile-major pre-swizzled copies (1-D bulk weight loads, vub_wn_entry.w_tiled).  Measured neutral in every
            # kernel although a bulk copy is 1.7x cheaper for the TMA unit than a tensor box
            # (profiles/r01_exp_bulk.log): the weight tile is not what bounds them.  Opt-in (VUB_TILED_WEIGHTS=1) so
            # that the default path does not carry a second copy of the weights.
            self._wt = torch.zeros(w_off if tiled else 0, dtype=torch.bfloat16, device=device)
            self._scale = torch.zeros(s_off, dtype=torch.float32, device=device)
            self._shift = torch.zeros(s_off, dtype=torch.float32, device=device)
            self._wd = torch.zeros(0, dtype=torch.bfloat16, device=device)
            self._wdt = torch.zeros(0, dtype=torch.bfloat16, device=device)
            self._tiled = tiled
            self.generation += 1
        if with_wd and self._wd.numel() == 0:
            self._wd = torch.zeros(wd_off, dtype=torch.bfloat16, device=device)
            self._wdt = torch.zeros(wd_off if tiled else 0, dtype=torch.bfloat16, device=device)
        self._with_wd = with_wd
        entries = []
        self._views = {}
        self._entry_of = {}  # id(NormConv2d) -> its host-side WnEntry (optim.FusedAdam embeds it)
        rows = 0
        for si, (slot, (wo, so, k, cout, cpad, nsrc, c_src, wdo, cout_k)) in enumerate(zip(self.slots, layout)):
            first = slot[0]
            w = self._w[wo:wo + cpad * k].view(cpad, k)
            sc = self._scale[so:so + cpad]
            sh = self._shift[so:so + cpad]
            taps = first.kernel_size ** 2
            wd = self._wd[wdo:wdo + k * cout_k].view(nsrc * c_src, taps * cout_k) if with_wd else None
            wt = self._wt[wo:wo + cpad * k] if (tiled and c_src % 64 == 0) else None
            wdt = self._wdt[wdo:wdo + k * cout_k] if (tiled and with_wd and cout_k % 64 == 0) else None
            self._views[id(first)] = ConvWeights(w, sc, sh, cout, cpad, nsrc, c_src, first.kernel_size,
                                                 first.stride, wd=wd, cout_k=cout_k, slot=si, w_tiled=wt, wd_tiled=wdt)
            r = 0
            for m in slot:
                e = N.WnEntry()
                e.v = m.conv.weight_v.data_ptr()
                e.g = m.conv.weight_g.data_ptr()
                e.gamma = m.gamma.data_ptr()
                e.beta = m.beta.data_ptr()
                e.bias = m.conv.bias.data_ptr()
                e.w_packed = w.data_ptr() + r * k * 2
                e.scale = sc.data_ptr() + r * 4
                e.shift = sh.data_ptr() + r * 4
                e.cout, e.cin, e.ktaps, e.c_src = m.out_channels, m.in_channels, m.kernel_size ** 2, c_src
                e.ktot, e.row_start = k, rows
                e.wd_cout, e.wd_col0 = cout_k, r
                if wt is not None:
                    e.w_tiled, e.tiled_rows = wt.data_ptr(), cpad
                if with_wd:
                    e.wd = wd.data_ptr()
                    if wdt is not None:
                        e.wd_tiled = wdt.data_ptr()
                rows += m.out_channels
                r += m.out_channels
                entries.append(e)
                self._entry_of[id(m)] = e
        arr = (N.WnEntry * len(entries))(*entries)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self._table = host.to(device)
        self._n_entries, self._rows = len(entries), rows
        self._device = device

    def _params(self) -> Iterable[torch.Tensor]:
        for slot in self.slots:
            for m in slot:
                yield m.conv.weight_v
                yield m.conv.weight_g
                yield m.gamma
                yield m.beta
                yield m.conv.bias

    def refresh(self, with_wd: bool = False) -> None:
        """Re-run weight-norm preparation iff a parameter changed (version or storage).  ``with_wd`` also
        prepares the data-gradient weight matrices the training path needs (kept once requested)."""
        params = list(self._params())
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("vunet_b200 has no CPU path: move the module to a CUDA device")
        ptr_key = tuple(p.data_ptr() for p in params)
        with_wd = with_wd or self._with_wd
        if self._device != dev or getattr(self, "_ptr_key", None) != ptr_key or with_wd != self._with_wd:
            self._allocate(dev, with_wd)
            self._ptr_key = ptr_key
            self._key = None
        key = weights_key(params)
        if key != self._key:
            ops.wn_prep(self._table, self._n_entries, self._rows)
            self._key = key

    def mark_current(self) -> None:
        """The prepared weights were just re-emitted by another launch (optim.FusedAdam)."""
        self._key = weights_key(list(self._params()))

    def get(self, conv: "NormConv2d") -> ConvWeights:
        return self._views[id(conv)]


# ------------------------------------------------------------------------------------------------
# Modules (names and signatures follow lib/modules.py)
# ------------------------------------------------------------------------------------------------
class SpaceToDepth(nn.Module):
    """lib/modules.py:11-21 (pure index permutation; only ``VunetOrg`` uses it)."""

    def __init__(self, block_size):
        super().__init__()
        self.bs = block_size

    def forward(self, x):
        n, c, h, w = x.size()
        bs = self.bs
        x = x.reshape(n, c, h // bs, bs, w // bs, bs).permute(0, 3, 5, 1, 2, 4)
        return x.reshape(n, c * bs * bs, h // bs, w // bs)


class DepthToSpace(nn.Module):
    """lib/modules.py:24-34, DCR order: out[n,c,bs*h+i,bs*w+j] = in[n,(i*bs+j)*C'+c,h,w].

    Inside the fused path this permutation is the store addressing of the up-sampling conv
    (``VUB_CONV_D2S``); this module exists for API parity and standalone use.
    """

    def __init__(self, block_size):
        super().__init__()
        self.bs = block_size

    def forward(self, x):
        n, c, h, w = x.size()
        bs = self.bs
        x = x.reshape(n, bs, bs, c // (bs * bs), h, w).permute(0, 3, 4, 1, 5, 2)
        return x.reshape(n, c // (bs * bs), h * bs, w * bs)


class IDAct(nn.Module):
    def forward(self, input):
        return input


class _WeightNormConvParams(nn.Module):
    """Holds ``bias``, ``weight_g``, ``weight_v`` under the reference's names (``*.conv.*``)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding):
        super().__init__()
        # Same initialisation stream as nn.Conv2d + weight_norm in the reference (lib/modules.py:135-138):
        # kaiming-uniform v, uniform bias, g = ||v|| per output channel.
        proto = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding)
        v = proto.weight.detach()
        self.bias = nn.Parameter(proto.bias.detach().clone())
        self.weight_g = nn.Parameter(torch.norm_except_dim(v, 2, 0).clone())
        self.weight_v = nn.Parameter(v.clone())


class NormConv2d(nn.Module):
    """Weight-normalised conv with learned affine — lib/modules.py:120-145.

    ``out = gamma * (conv2d(x, g * v/||v||) + bias) + beta``; parameters ``beta, gamma, conv.bias,
    conv.weight_g, conv.weight_v`` in that order.
    """

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
        super().__init__()
        if kernel_size not in (1, 3) or padding != kernel_size // 2 or stride not in (1, 2):
            raise NotImplementedError("vunet_b200 kernels cover k in {1,3}, padding k//2, stride in {1,2}")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        self.beta = nn.Parameter(torch.zeros([1, out_channels, 1, 1], dtype=torch.float32))
        self.gamma = nn.Parameter(torch.ones([1, out_channels, 1, 1], dtype=torch.float32))
        self.conv = _WeightNormConvParams(in_channels, out_channels, kernel_size, stride, padding)
        self._vub_split = 1      # number of concatenated sources feeding this conv
        self._vub_stem = False   # K too small for tensor cores -> CUDA-core stem kernel
        self._vub_owner = None   # top-level sub-network ("eu" / "ed" / "du" / "dd"): gradient bucket of this conv
        self._solo_bank: Optional[WeightBank] = None

    # fused-path entry points -------------------------------------------------------------------
    def _run(self, bank: WeightBank, srcs: Sequence[Tuple[Act, str]], *, residual: Optional[Act] = None,
             want_raw=True, want_elu=True, d2s=False) -> Act:
        """``srcs``: the concatenated inputs as (activation, "raw" | "elu") pairs; ``residual`` is added raw."""
        cw = bank.get(self)
        if bank.precision == "fp32":  # validation mode: same schedule, fp32 CUDA-core conv
            raw, elu = ops.conv_fwd32([a.pick(m) for a, m in srcs], cw.w32, cw.scale, cw.shift, cout=cw.cout,
                                      ksize=cw.ksize, stride=cw.stride,
                                      residual=None if residual is None else residual.raw, want_raw=want_raw,
                                      want_elu=want_elu, d2s=d2s)
            out = Act(raw, elu)
            if bank.tape is not None:
                bank.tape.conv(self, srcs, residual, out, d2s)
            return out
        xs, elu_in = [], 0
        for s, (a, m) in enumerate(srcs):
            if m == "elu" and a.elu is None and a.raw is not None:
                xs.append(a.raw)     # the producer skipped the ELU copy (lean_elu): ELU is applied on load
                elu_in |= 1 << s
            else:
                xs.append(a.pick(m))
        raw, elu = ops.conv_fwd(xs, cw.w, cw.scale, cw.shift, ksize=cw.ksize,
                                stride=cw.stride, residual=None if residual is None else residual.raw,
                                want_raw=want_raw, want_elu=want_elu, d2s=d2s, elu_in=elu_in, w_tiled=cw.w_tiled)
        out = Act(raw, elu)
        if bank.tape is not None:
            bank.tape.conv(self, srcs, residual, out, d2s)
        return out

    def _run_f32(self, bank: WeightBank, srcs: Sequence[Tuple[Act, str]], layout: str):
        """fp32 output (latent head input NHWC / image NCHW).  Returns (tensor, gradient slot or None)."""
        cw = bank.get(self)
        if bank.precision == "fp32":
            out, _ = ops.conv_fwd32([a.pick(m) for a, m in srcs], cw.w32, cw.scale, cw.shift, cout=cw.cout, ksize=cw.ksize,
                                    want_raw=True, want_elu=False, nchw=layout == "nchw")
            return out, (bank.tape.conv(self, srcs, None, None, False) if bank.tape is not None else None)
        out = ops.conv_fwd_f32([a.pick(m) for a, m in srcs], cw.w, cw.scale, cw.shift, ksize=cw.ksize, layout=layout,
                               w_tiled=cw.w_tiled,
                               cout_valid=cw.cout)
        slot = None
        if bank.tape is not None:
            slot = bank.tape.conv(self, srcs, None, None, False)
        return out, slot

    def _run_stem(self, x: torch.Tensor, want_raw=True, want_elu=True, bank: Optional[WeightBank] = None) -> Act:
        if x.dtype == torch.uint8 and bank is not None and (bank.precision == "fp32" or bank.tape is not None):
            # uint8 HWC images are an inference front-end format; training / validation take the reference's tensors
            x = ((x.float().div(255.0) * 2.0) - 1.0).permute(0, 3, 1, 2).contiguous()
        if bank is not None and bank.precision == "fp32":
            x_nhwc, _ = ops.nchw_to_nhwc32(x, True, False)
            return self._run(bank, [(Act(x_nhwc, None), "raw")], want_raw=want_raw, want_elu=want_elu)
        c = self.conv
        raw, elu = ops.stem_fwd(x, c.weight_v, c.weight_g, self.gamma, self.beta, c.bias, want_raw, want_elu)
        out = Act(raw, elu)
        if bank is not None and bank.tape is not None:
            bank.tape.stem(self, x, out)
        return out

    # reference-compatible standalone call -----------------------------------------------------
    def forward(self, x):
        warn_forward_only(self, "NormConv2d.forward", x)
        if self.kernel_size == 1 and self.in_channels not in _SUPPORTED_C and self.out_channels in (32, 64):
            if not x.is_cuda:
                raise RuntimeError("vunet_b200 has no CPU path: inputs must live on a CUDA device")
            return self._run_stem(x.float()).nchw()
        if self._solo_bank is None:
            self._solo_bank = WeightBank([[self]])
        self._solo_bank.refresh()
        cw = self._solo_bank.get(self)
        act = to_act(x, need_raw=True, need_elu=False)
        if cw.cout_pad == cw.cout and self.stride == 1:
            out = ops.conv_fwd_f32([act.raw], cw.w, cw.scale, cw.shift, ksize=cw.ksize, layout="nhwc")
            return out.permute(0, 3, 1, 2)
        if self.stride == 1:
            return ops.conv_fwd_f32([act.raw], cw.w, cw.scale, cw.shift, ksize=cw.ksize, layout="nchw",
                                    cout_valid=cw.cout)
        return self._run(self._solo_bank, [(act, "raw")]).nchw()


class Downsample(nn.Module):
    """lib/modules.py:148-161: 3x3 stride-2 NormConv2d on the raw activation."""

    def __init__(self, channels, out_channels=None, conv_layer=NormConv2d):
        super().__init__()
        self.down = conv_layer(channels, channels if out_channels is None else out_channels,
                               kernel_size=3, stride=2, padding=1)

    def _run(self, bank, h: Act, want_raw=True, want_elu=True) -> Act:
        return self.down._run(bank, [(h, "raw")], want_raw=want_raw, want_elu=want_elu)

    def forward(self, x):
        warn_forward_only(self, "Downsample.forward", x)
        bank = _solo_bank(self)
        return self._run(bank, to_act(x, True, False)).nchw()


class Upsample(nn.Module):
    """lib/modules.py:164-182 (sub-pixel branch): 3x3 conv to 4*C' channels + DepthToSpace(2),
    fused into one kernel (the shuffle is the epilogue's store addressing)."""

    def __init__(self, in_channels, out_channels, subpixel=True, conv_layer=NormConv2d):
        super().__init__()
        if not subpixel:
            raise NotImplementedError("bilinear Upsample (subpixel=False) is outside the B200 hot path; "
                                      "every shipped config sets subpixel_upsampling: True")
        self.up = conv_layer(in_channels, 4 * out_channels, 3, padding=1)
        self.op2 = DepthToSpace(block_size=2)

    def _run(self, bank, h: Act, want_raw=True, want_elu=True) -> Act:
        return self.up._run(bank, [(h, "raw")], want_raw=want_raw, want_elu=want_elu, d2s=True)

    def forward(self, x):
        warn_forward_only(self, "Upsample.forward", x)
        bank = _solo_bank(self)
        return self._run(bank, to_act(x, True, False)).nchw()


class VunetRNB(nn.Module):
    """Residual block — lib/modules.py:185-233:
    ``x + conv3x3(ELU(cat[x, nin1x1(ELU(a))]))`` (or ``x + conv3x3(ELU(x))`` without ``a``).

    Fused schedule: ``nin`` reads the stored ELU(a) and stores only ELU(n); the 3x3 conv reads the two
    ELU'd sources through two TMA maps (no concat tensor), adds the raw ``x`` in its epilogue and
    stores the result raw and/or ELU'd for its consumers.
    """

    def __init__(self, channels, a_channels=None, residual=False, kernel_size=3, activate=True,
                 conv_layer=NormConv2d, act_fn=None, dropout_prob=0.0):
        super().__init__()
        if not activate or act_fn is not None or kernel_size != 3:
            raise NotImplementedError("vunet_b200 VunetRNB fuses ELU and a 3x3 conv (the only variant the VUnet uses)")
        self.residual = residual
        self.dout = nn.Dropout(p=dropout_prob)
        if self.residual:
            assert a_channels is not None
            self.nin = conv_layer(a_channels, channels, kernel_size=1)
            self.nin._vub_split = max(1, a_channels // channels)
        self.act_fn = nn.ELU()
        in_c = 2 * channels if self.residual else channels
        self.conv = conv_layer(in_channels=in_c, out_channels=channels, kernel_size=kernel_size,
                               padding=kernel_size // 2)
        self.conv._vub_split = 2 if self.residual else 1

    def _run(self, bank, x: Act, a: Union[None, Act, Sequence[Act]] = None, want_raw=True,
             want_elu=True) -> Act:
        if self.dout.p > 0.0 and self.training:
            raise NotImplementedError("dropout_prob > 0 is not fused (the reference never enables it: SURVEY.md §3.3)")
        srcs = [(x, "elu")]
        if a is not None:
            assert self.residual
            a_list = [a] if isinstance(a, Act) else list(a)
            n = self.nin._run(bank, [(t, "elu") for t in a_list], want_raw=False, want_elu=True)
            srcs.append((n, "elu"))
        return self.conv._run(bank, srcs, residual=x, want_raw=want_raw, want_elu=want_elu)

    def forward(self, x, a=None):
        warn_forward_only(self, "VunetRNB.forward", x, a)
        bank = _solo_bank(self)
        xa = to_act(x)
        aa = None if a is None else to_act(a, need_raw=False, need_elu=True)
        if aa is not None and self.nin._vub_split == 2:
            c = self.nin.in_channels // 2
            aa = [Act(None, aa.elu[..., :c].contiguous()), Act(None, aa.elu[..., c:].contiguous())]
        return self._run(bank, xa, aa).nchw()


_WARNED_FWD_ONLY = [False]


def warn_forward_only(module: nn.Module, what: str, *inputs) -> None:
    """Only ``VunetAlter.forward`` records a backward schedule (w.r.t. parameters).  Every other entry point returns
    tensors without autograd history; the reference is differentiable on all of them, so a loss built on them would
    silently train nothing: raise when an INPUT asks for gradients, warn once when only parameters do."""
    if not torch.is_grad_enabled():
        return
    if any(isinstance(t, torch.Tensor) and t.requires_grad for t in inputs):
        raise NotImplementedError(f"vunet_b200: {what} is forward-only; its inputs require grad but no gradient would "
                                  "flow (only VunetAlter.forward is differentiable, w.r.t. parameters)")
    if not _WARNED_FWD_ONLY[0] and any(p.requires_grad for p in module.parameters()):
        _WARNED_FWD_ONLY[0] = True
        import warnings
        warnings.warn(f"vunet_b200: {what} is forward-only (outputs carry no autograd history); only "
                      "VunetAlter.forward is differentiable. Wrap the call in torch.no_grad() to silence this.",
                      stacklevel=3)


def _solo_bank(module: nn.Module) -> WeightBank:
    bank = getattr(module, "_vub_bank", None)
    if bank is None:
        bank = WeightBank.for_module(module)
        object.__setattr__(module, "_vub_bank", bank)
    bank.refresh()
    return bank
