"""fp32 validation mode (BASELINE.json north_star: relative L2 <= 1e-4 against the reference "in an
fp32-accumulate validation mode").

``model.validation_fp32 = True`` (or ``with fp32_validation(model): ...``) makes ``VunetAlter.forward`` /
``transfer`` / ``test_forward`` run the SAME fused schedule — same operand packing order, concat elimination,
raw/ELU activation pairs, DepthToSpace addressing, merged latent head, dead-branch skipping — with fp32
activations and fp32 weights through ``vub_conv_fwd_f32`` (CUDA cores, direct convolution).  It separates
schedule errors from bf16 rounding; it is forward-only and not a fast path.
"""
from __future__ import annotations

import contextlib
from typing import Sequence

import torch
from torch import nn

from . import _native as N
from . import ops
from .modules import weights_key


class ConvWeights32:
    """fp32 ``v/||v||`` [K, pitch] (K order (source, tap, channel)) + scale/shift of one (possibly merged) conv."""

    __slots__ = ("w32", "scale", "shift", "cout", "nsrc", "c_src", "ksize", "stride")

    def __init__(self, w32, scale, shift, cout, nsrc, c_src, ksize, stride):
        self.w32, self.scale, self.shift = w32, scale, shift
        self.cout, self.nsrc, self.c_src, self.ksize, self.stride = cout, nsrc, c_src, ksize, stride


class WeightBank32:
    """Counterpart of ``modules.WeightBank`` for the validation mode: every conv of the tree (stems included),
    prepared by one ``wn_prep`` launch that emits fp32 matrices only."""

    precision = "fp32"

    def __deepcopy__(self, memo):
        return None  # pointer cache of one module tree (see WeightBank.__deepcopy__)

    def __init__(self, root: nn.Module = None, slots=None):
        """``root``: a module tree (every NormConv2d of it, merged pairs in one slot); or ``slots``: an explicit list of
        lists of NormConv2d-like objects (perceptual.PerceptualVGG's frozen plain convs)."""
        from .modules import NormConv2d
        if slots is not None:
            self.slots = [list(sl) for sl in slots]
        else:
            merged_second = {}
            for m in root.modules():
                for a, b in getattr(m, "_vub_merged_pairs", ()):
                    merged_second[id(b)] = a
            self.slots, index = [], {}
            for m in root.modules():
                if isinstance(m, NormConv2d) and id(m) not in merged_second:
                    index[id(m)] = len(self.slots)
                    self.slots.append([m])
            for m in root.modules():
                if isinstance(m, NormConv2d) and id(m) in merged_second:
                    self.slots[index[id(merged_second[id(m)])]].append(m)
        self.tape = None
        self._key = None
        self._ptr_key = None
        self._views = {}
        self.generation = 0  # bumped by every re-allocation (graph holders compare it, see synthesis.py)

    def _params(self):
        for slot in self.slots:
            for m in slot:
                yield from (m.conv.weight_v, m.conv.weight_g, m.gamma, m.beta, m.conv.bias)

    def _allocate(self, device) -> None:
        self.generation += 1
        w_off, s_off = 0, 0
        layout = []
        for slot in self.slots:
            first = slot[0]
            k = first.in_channels * first.kernel_size ** 2
            cout = sum(m.out_channels for m in slot)
            pitch = (cout + 3) // 4 * 4
            layout.append((w_off, s_off, k, cout, pitch))
            w_off += k * pitch
            s_off += pitch
        self._w32 = torch.zeros(w_off, dtype=torch.float32, device=device)
        self._scale = torch.zeros(s_off, dtype=torch.float32, device=device)
        self._shift = torch.zeros(s_off, dtype=torch.float32, device=device)
        entries, rows = [], 0
        self._views = {}
        for slot, (wo, so, k, cout, pitch) in zip(self.slots, layout):
            first = slot[0]
            nsrc = first._vub_split
            c_src = first.in_channels // nsrc
            w = self._w32[wo:wo + k * pitch].view(k, pitch)
            sc, sh = self._scale[so:so + pitch], self._shift[so:so + pitch]
            self._views[id(first)] = ConvWeights32(w, sc, sh, cout, nsrc, c_src, first.kernel_size, first.stride)
            r = 0
            for m in slot:
                e = N.WnEntry()
                e.v, e.g = m.conv.weight_v.data_ptr(), m.conv.weight_g.data_ptr()
                e.gamma, e.beta, e.bias = m.gamma.data_ptr(), m.beta.data_ptr(), m.conv.bias.data_ptr()
                e.w_packed = None
                e.scale, e.shift = sc.data_ptr() + r * 4, sh.data_ptr() + r * 4
                e.cout, e.cin, e.ktaps, e.c_src = m.out_channels, m.in_channels, m.kernel_size ** 2, c_src
                e.ktot, e.row_start = k, rows
                e.wd, e.wd_cout, e.wd_col0, e.w32 = None, pitch, r, w.data_ptr()
                rows += m.out_channels
                r += m.out_channels
                entries.append(e)
        arr = (N.WnEntry * len(entries))(*entries)
        self._table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
        self._n_entries, self._rows = len(entries), rows

    def refresh(self, with_wd: bool = False) -> None:
        params = list(self._params())
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("vunet_b200 has no CPU path: move the module to a CUDA device")
        ptr_key = tuple(p.data_ptr() for p in params)
        if ptr_key != self._ptr_key:
            self._allocate(dev)
            self._ptr_key, self._key = ptr_key, None
        key = weights_key(params)
        if key != self._key:
            ops.wn_prep(self._table, self._n_entries, self._rows)
            self._key = key

    def get(self, conv) -> ConvWeights32:
        return self._views[id(conv)]


@contextlib.contextmanager
def fp32_validation(model):
    """``with fp32_validation(model): imgs = model.transfer(x, c)`` — run the fused schedule in fp32."""
    old = getattr(model, "validation_fp32", False)
    model.validation_fp32 = True
    try:
        yield model
    finally:
        model.validation_fp32 = old
