"""Batched frame synthesis — the forward-only caller of the VUnet.

The reference renders behavior-net pose sequences one frame at a time with batch 1 and a host sync
per frame (``data/data_conversions_3d.py:1130-1173``: ``synth_model.transfer(app_img, stickman)`` then
``.cpu()``).  :class:`FrameSynthesizer` runs the same ``VunetAlter.transfer`` on whole batches of
frames, captured once in a CUDA graph (~90 kernel launches per batch, ~40 of them launch-bound), with a
double-buffered pinned-host pipeline for callers that hold their stickmen on the host.  Frames are independent
units: multi-GPU rendering shards them with :func:`shard_range` and needs no collective.

Data formats at the host boundary follow the reference's rendering loop: the stickman it rasterises is a uint8
HWC image that ``ToTensor()`` + ``x * 2 - 1`` turn into the fp32 NCHW network input, and the frame it keeps is
``(scale_img(rgb) * 255).permute(0, 2, 3, 1).astype(uint8)``.  With ``uint8_inputs`` / ``uint8_frames`` both
conversions run inside the kernels (bit-identically), so 1 byte per value crosses PCIe instead of 4.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import ops


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) share of ``n_items`` frames (or sequences) for ``rank``."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(n_items, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class _IoSet:
    """One set of static graph inputs / output and the graph that reads / writes them."""

    __slots__ = ("x", "c", "out", "graph", "done", "in_ready", "out_free")

    def __init__(self, x, c):
        self.x, self.c = x, c
        self.out = None
        self.graph = None
        self.done = self.in_ready = self.out_free = None


class _GraphedSynthesizer:
    """Shared machinery: static input sets, CUDA-graph capture against the model's prepared-weight buffers,
    re-capture when those buffers move, and the host pipeline (:meth:`render_host`)."""

    def __init__(self, model, x_shape, c_shape, device, use_graph: bool, uint8_frames: bool, uint8_inputs: bool):
        self.model = model.eval()
        p = next(model.parameters())
        self.device = torch.device(device) if device is not None else p.device
        if self.device.type != "cuda":
            raise RuntimeError("vunet_b200 has no CPU path: move the module to a CUDA device")
        # uint8_frames: return HWC uint8 video frames, (scale_img(rgb) * 255).astype(uint8) of the reference's
        # rendering loop (data/data_conversions_3d.py:1166-1173) fused behind the graph: 4x fewer bytes to the host
        self.uint8_frames = uint8_frames
        # uint8_inputs: appearance / stickman batches as uint8 HWC images [B, H, W, C] (what the reference holds before
        # ToTensor() and x * 2 - 1, data/__init__.py:15-43); the stem kernels convert, bit-identically
        self.uint8_inputs = uint8_inputs
        self._x_shape, self._c_shape = tuple(x_shape), tuple(c_shape)
        self._use_graph = use_graph
        self.launches_per_call = 0
        self._pool = None
        self._sets: List[_IoSet] = [self._new_set()]
        self._slot = 0
        self._h2d = self._d2h = None
        self._capture()

    # ---- static inputs --------------------------------------------------------------------------------
    def _new_set(self) -> _IoSet:
        def alloc(shape):
            if self.uint8_inputs:
                b, ch, h, w = shape
                return torch.zeros(b, h, w, ch, dtype=torch.uint8, device=self.device)
            return torch.zeros(shape, dtype=torch.float32, device=self.device)
        return _IoSet(alloc(self._x_shape), alloc(self._c_shape))

    @property
    def x(self) -> torch.Tensor:
        return self._sets[0].x

    @property
    def c(self) -> torch.Tensor:
        return self._sets[0].c

    @property
    def graph(self):
        return self._sets[0].graph

    # ---- capture --------------------------------------------------------------------------------------
    def _forward(self, x, c) -> torch.Tensor:
        raise NotImplementedError

    def _step(self, s: _IoSet) -> torch.Tensor:
        return self._forward(s.x, s.c)

    def _capture_set(self, s: _IoSet) -> None:
        with torch.no_grad():
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(3):
                    before = ops.LAUNCHES
                    s.out = self._step(s)
                    self.launches_per_call = ops.LAUNCHES - before
            torch.cuda.current_stream(self.device).wait_stream(side)
            s.graph = None
            if self._use_graph:
                if self._pool is None:
                    self._pool = torch.cuda.graph_pool_handle()
                s.graph = torch.cuda.CUDAGraph()
                # every set's graph allocates from ONE private pool: the graphs replay one after another on one
                # stream, so their intermediate activations may share memory; the outputs stay alive (held here)
                with torch.cuda.graph(s.graph, pool=self._pool):
                    s.out = self._step(s)

    def _capture(self) -> None:
        """Warm up and (re)capture every set's graph against the model's CURRENT prepared-weight buffers."""
        with torch.no_grad():
            self._bank = self.model._bank()  # weight-norm preparation happens outside the graph
        self._bank_gen = self._bank.generation
        self._pool = None
        for s in self._sets:
            self._capture_set(s)

    def _check_bank(self) -> None:
        # The graphs have the bank's packed-weight buffers baked in.  They survive training steps and parameter
        # re-assignment (modules.WeightBank._allocate keeps them per device); a device move or a rebuilt bank
        # (copy.deepcopy, VUB_TILED_WEIGHTS flip) re-allocates them -> capture again instead of reading freed memory.
        # model._bank() also re-runs weight-norm preparation when a parameter changed, so a replay never reads stale
        # weights (the reference re-normalises in every forward).
        with torch.no_grad():
            bank = self.model._bank()
        if bank is not self._bank or bank.generation != self._bank_gen:
            torch.cuda.current_stream(self.device).synchronize()
            self._capture()

    def refresh_weights(self) -> None:
        """Call after the model's parameters changed: re-runs weight-norm preparation (``run`` does the same check
        itself).  The packed buffers are allocated once per device — training steps, ``FusedAdam`` and
        ``load_state_dict`` write into the same memory — so the captured graphs normally stay valid; if the bank had
        to re-allocate (device move) they are captured again."""
        self._check_bank()

    # ---- device-resident API --------------------------------------------------------------------------
    def _replay(self, s: _IoSet) -> torch.Tensor:
        if s.graph is not None:
            s.graph.replay()
        else:
            with torch.no_grad():
                s.out = self._step(s)
        return s.out

    def run(self) -> torch.Tensor:
        """Synthesize from the static input buffers ``self.x`` / ``self.c``; returns a static output buffer
        (overwritten by the next call)."""
        self._check_bank()
        return self._replay(self._sets[0])

    def __call__(self, x: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.c.copy_(c, non_blocking=True)
        return self.run()

    # ---- host pipeline --------------------------------------------------------------------------------
    def _init_pipeline(self) -> None:
        dev = self.device
        while len(self._sets) < 2:  # second set of static inputs / output with its own graph (same memory pool)
            s = self._new_set()
            self._capture_set(s)
            self._sets.append(s)
        self._h2d = torch.cuda.Stream(device=dev)
        self._d2h = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        for s in self._sets:
            s.done, s.in_ready, s.out_free = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
            s.done.record(cur)
            s.out_free.record(cur)

    def render_host(self, x_host: torch.Tensor, c_host: torch.Tensor, out_host: torch.Tensor) -> None:
        """Host (pinned) buffers in, host (pinned) frames out; dtypes / layouts as the static inputs and the output
        (``uint8_inputs``: uint8 HWC, else fp32 NCHW; ``uint8_frames``: uint8 HWC, else fp32 NCHW).  Asynchronous and
        pipelined over consecutive calls: two input/output sets with one graph each alternate, the host->device
        copies go STRAIGHT into the next graph's static inputs on a copy stream while the current graph runs, and
        the frames leave from the graph's own output buffer on a third stream (PCIe is full duplex).  ``out_host``
        is valid after :meth:`join` + a stream/device synchronize (or after the next-but-one call has been
        synchronized); ``x_host`` / ``c_host`` may be refilled once the call after the next has been issued, or
        after :meth:`join`."""
        self._check_bank()
        if self._h2d is None or len(self._sets) < 2 or self._sets[0].done is None:
            self._init_pipeline()
        s = self._sets[self._slot]
        self._slot = (self._slot + 1) % len(self._sets)
        for name, host, dev_t in (("x_host", x_host, s.x), ("c_host", c_host, s.c), ("out_host", out_host, s.out)):
            if host.dtype != dev_t.dtype or tuple(host.shape) != tuple(dev_t.shape):
                raise ValueError(f"render_host: {name} must be {dev_t.dtype} {tuple(dev_t.shape)}, got {host.dtype} "
                                 f"{tuple(host.shape)}")
        cur = torch.cuda.current_stream(self.device)
        self._h2d.wait_event(s.done)                # the previous replay of this set's graph has read its inputs
        with torch.cuda.stream(self._h2d):
            s.x.copy_(x_host, non_blocking=True)
            s.c.copy_(c_host, non_blocking=True)
            s.in_ready.record(self._h2d)
        cur.wait_event(s.in_ready)
        cur.wait_event(s.out_free)                  # the previous download of this set's output has finished
        out = self._replay(s)
        s.done.record(cur)
        self._d2h.wait_event(s.done)
        with torch.cuda.stream(self._d2h):
            out_host.copy_(out, non_blocking=True)
            s.out_free.record(self._d2h)

    def join(self) -> None:
        """Make the current stream wait for every copy issued by :meth:`render_host`."""
        if self._h2d is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_stream(self._h2d)
            cur.wait_stream(self._d2h)


class FrameSynthesizer(_GraphedSynthesizer):
    """``imgs = FrameSynthesizer(model, batch)(x, c)`` == ``model.transfer(x, c)`` for a fixed batch size.

    x: appearance ``[batch, n_channels_x, Sx, Sx]`` fp32, c: stickmen ``[batch, 3, S, S]`` fp32 (NCHW, values
    in [-1, 1]); returns ``[batch, 3, S, S]`` fp32 (a static buffer, overwritten by the next call).  With
    ``uint8_inputs`` x / c are uint8 HWC image batches ``[batch, Sx, Sx, n_channels_x]`` / ``[batch, S, S, 3]``; with
    ``uint8_frames`` the result is ``[batch, S, S, 3]`` uint8.  The latent noise is drawn inside the graph from
    torch's CUDA generator, like the reference.
    """

    def __init__(self, model, batch: int, device: Optional[torch.device] = None, use_graph: bool = True,
                 uint8_frames: bool = False, uint8_inputs: bool = False):
        self.batch = batch
        s = model.spatial_size
        sx = s // 2 ** (model.n_scales - model.n_scales_x)
        super().__init__(model, (batch, model.n_channels_x, sx, sx), (batch, 3, s, s), device, use_graph, uint8_frames,
                         uint8_inputs)

    def _forward(self, x, c):
        return self.model.transfer(x, c, uint8_frames=self.uint8_frames)


class SequenceSynthesizer(_GraphedSynthesizer):
    """Renders whole pose sequences against one appearance image each (BASELINE.json configs[4]).

    ``n_sequences`` appearance images ``x [S, Cx, Sx, Sx]`` and ``S * T`` stickmen ``c [S * T, 3, H, W]``
    (sequence-major) -> ``[S * T, 3, H, W]`` frames (uint8 HWC variants as for :class:`FrameSynthesizer`).  The
    appearance encoder (``eu`` + ``ed``, 23 % of the FLOPs of ``transfer`` at 256 x 256) runs once per sequence
    instead of once per frame and the appearance image is uploaded once per sequence; both halves are captured in
    one CUDA graph.  The latent noise is drawn once per sequence (the reference's per-frame loop re-draws it for
    every frame; the distribution of a single frame is the same)."""

    def __init__(self, model, n_sequences: int, frames_per_sequence: int, use_graph: bool = True,
                 uint8_frames: bool = False, uint8_inputs: bool = False):
        self.S, self.T = n_sequences, frames_per_sequence
        s = model.spatial_size
        sx = s // 2 ** (model.n_scales - model.n_scales_x)
        super().__init__(model, (self.S, model.n_channels_x, sx, sx), (self.S * self.T, 3, s, s), None, use_graph,
                         uint8_frames, uint8_inputs)

    def _forward(self, x, c):
        app = self.model.encode_appearance(x)
        return self.model.decode_pose(c, app, self.T, uint8_frames=self.uint8_frames)
