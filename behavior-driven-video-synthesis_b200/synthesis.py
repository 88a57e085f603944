"""Batched frame synthesis — the forward-only caller of the VUnet.

The reference renders behavior-net pose sequences one frame at a time with batch 1 and a host sync
per frame (``data/data_conversions_3d.py:1130-1173``: ``synth_model.transfer(app_img, stickman)`` then
``.cpu()``).  :class:`FrameSynthesizer` runs the same ``VunetAlter.transfer`` on whole batches of
frames, captured once in a CUDA graph (~90 kernel launches per batch, ~40 of them launch-bound), with
pinned-host staging for callers that hold their stickmen on the host.  Frames are independent units:
multi-GPU rendering shards them with :func:`shard_range` and needs no collective.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) share of ``n_items`` frames (or sequences) for ``rank``."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(n_items, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class FrameSynthesizer:
    """``imgs = FrameSynthesizer(model, batch)(x, c)`` == ``model.transfer(x, c)`` for a fixed batch size.

    x: appearance ``[batch, n_channels_x, Sx, Sx]`` fp32, c: stickmen ``[batch, 3, S, S]`` fp32 (NCHW, values
    in [-1, 1]); returns ``[batch, 3, S, S]`` fp32 (a static buffer, overwritten by the next call).
    The latent noise is drawn inside the graph from torch's CUDA generator, like the reference.
    """

    def __init__(self, model, batch: int, device: Optional[torch.device] = None, use_graph: bool = True,
                 uint8_frames: bool = False):
        self.model = model.eval()
        self.batch = batch
        # uint8_frames: return HWC uint8 video frames, (scale_img(rgb) * 255).astype(uint8) of the reference's
        # rendering loop (data/data_conversions_3d.py:1166-1173) fused behind the graph: 4x fewer bytes to the host
        self.uint8_frames = uint8_frames
        p = next(model.parameters())
        self.device = torch.device(device) if device is not None else p.device
        if self.device.type != "cuda":
            raise RuntimeError("vunet_b200 has no CPU path: move the module to a CUDA device")
        s = model.spatial_size
        sx = s // 2 ** (model.n_scales - model.n_scales_x)
        self.x = torch.zeros(batch, model.n_channels_x, sx, sx, dtype=torch.float32, device=self.device)
        self.c = torch.zeros(batch, 3, s, s, dtype=torch.float32, device=self.device)
        self.graph = None
        self.launches_per_call = 0
        self._out = None
        # host pipeline (render_host): staging buffers + copy streams so that the host->device copy of batch
        # k+1 and the device->host copy of batch k-1 overlap the kernels of batch k (PCIe is full duplex)
        self._depth = 2
        self._slot = 0
        self._stage_in = None
        with torch.no_grad():
            model._bank()  # weight-norm preparation happens outside the graph
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(3):
                    before = ops.LAUNCHES
                    self._out = self._step()
                    self.launches_per_call = ops.LAUNCHES - before
            torch.cuda.current_stream(self.device).wait_stream(side)
            if use_graph:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._out = self._step()

    def _step(self) -> torch.Tensor:
        imgs = self.model.transfer(self.x, self.c)
        return ops.frames_to_uint8(imgs) if self.uint8_frames else imgs

    def refresh_weights(self) -> None:
        """Call after the model's parameters changed (re-runs weight-norm prep into the same buffers)."""
        self.model._bank()

    def run(self) -> torch.Tensor:
        """Synthesize from the static input buffers ``self.x`` / ``self.c``."""
        if self.graph is not None:
            self.graph.replay()
        else:
            with torch.no_grad():
                self._out = self._step()
        return self._out

    def __call__(self, x: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.c.copy_(c, non_blocking=True)
        return self.run()

    def _init_pipeline(self) -> None:
        dev = self.device
        self._stage_in = [(torch.empty_like(self.x), torch.empty_like(self.c)) for _ in range(self._depth)]
        self._stage_out = [torch.empty_like(self._out) for _ in range(self._depth)]
        self._h2d = torch.cuda.Stream(device=dev)
        self._d2h = torch.cuda.Stream(device=dev)
        mk = lambda: [torch.cuda.Event() for _ in range(self._depth)]  # noqa: E731
        self._in_ready, self._in_free, self._out_ready, self._out_free = mk(), mk(), mk(), mk()
        cur = torch.cuda.current_stream(dev)
        for i in range(self._depth):
            self._in_free[i].record(cur)
            self._out_free[i].record(cur)

    def render_host(self, x_host: torch.Tensor, c_host: torch.Tensor, out_host: torch.Tensor) -> None:
        """Host (pinned) buffers in, host (pinned) frames out.  Asynchronous and pipelined over consecutive
        calls: copies run on their own streams through double-buffered staging tensors, the kernels on the
        current stream.  ``out_host`` is valid after :meth:`join` + a stream/device synchronize (or after the
        next-but-one call has been synchronized); ``x_host`` / ``c_host`` may be refilled once the call after
        the next has been issued, or after :meth:`join`."""
        if self._stage_in is None:
            self._init_pipeline()
        i = self._slot
        self._slot = (i + 1) % self._depth
        cur = torch.cuda.current_stream(self.device)
        sx, sc = self._stage_in[i]
        self._h2d.wait_event(self._in_free[i])          # kernels of the previous user of this slot have read it
        with torch.cuda.stream(self._h2d):
            sx.copy_(x_host, non_blocking=True)
            sc.copy_(c_host, non_blocking=True)
            self._in_ready[i].record(self._h2d)
        cur.wait_event(self._in_ready[i])
        self.x.copy_(sx, non_blocking=True)              # device-to-device into the graph's static inputs
        self.c.copy_(sc, non_blocking=True)
        self._in_free[i].record(cur)
        out = self.run()
        cur.wait_event(self._out_free[i])                # the previous download from this slot has finished
        so = self._stage_out[i]
        so.copy_(out, non_blocking=True)
        self._out_ready[i].record(cur)
        self._d2h.wait_event(self._out_ready[i])
        with torch.cuda.stream(self._d2h):
            out_host.copy_(so, non_blocking=True)
            self._out_free[i].record(self._d2h)

    def join(self) -> None:
        """Make the current stream wait for every copy issued by :meth:`render_host`."""
        if self._stage_in is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_stream(self._h2d)
            cur.wait_stream(self._d2h)


class SequenceSynthesizer:
    """Renders whole pose sequences against one appearance image each (BASELINE.json configs[4]).

    ``n_sequences`` appearance images ``x [S, Cx, Sx, Sx]`` and ``S * T`` stickmen ``c [S * T, 3, H, W]``
    (sequence-major) -> ``[S * T, 3, H, W]`` frames.  The appearance encoder (``eu`` + ``ed``, 23 % of the FLOPs
    of ``transfer`` at 256 x 256) runs once per sequence instead of once per frame; both halves are captured in
    one CUDA graph.  The latent noise is drawn once per sequence (the reference's per-frame loop re-draws it for
    every frame; the distribution of a single frame is the same)."""

    def __init__(self, model, n_sequences: int, frames_per_sequence: int, use_graph: bool = True):
        self.model = model.eval()
        p = next(model.parameters())
        self.device = p.device
        if self.device.type != "cuda":
            raise RuntimeError("vunet_b200 has no CPU path: move the module to a CUDA device")
        self.S, self.T = n_sequences, frames_per_sequence
        s = model.spatial_size
        sx = s // 2 ** (model.n_scales - model.n_scales_x)
        self.x = torch.zeros(self.S, model.n_channels_x, sx, sx, dtype=torch.float32, device=self.device)
        self.c = torch.zeros(self.S * self.T, 3, s, s, dtype=torch.float32, device=self.device)
        self.graph = None
        with torch.no_grad():
            model._bank()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._out = self._step()
            torch.cuda.current_stream(self.device).wait_stream(side)
            if use_graph:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._out = self._step()

    def _step(self):
        app = self.model.encode_appearance(self.x)
        return self.model.decode_pose(self.c, app, self.T)

    def run(self) -> torch.Tensor:
        if self.graph is not None:
            self.graph.replay()
        else:
            with torch.no_grad():
                self._out = self._step()
        return self._out

    def __call__(self, x: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.c.copy_(c, non_blocking=True)
        return self.run()
