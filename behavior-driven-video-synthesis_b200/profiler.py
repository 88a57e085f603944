"""Per-kernel device timing with CUDA events on the launching stream (used by ``bench.py`` for the
roofline numbers; never active inside a timed throughput region)."""
from __future__ import annotations

from collections import defaultdict
from typing import Dict

import torch

from . import ops


class KernelProfiler:
    """``with KernelProfiler() as kp: model.transfer(x, c)`` then ``kp.summary()``.

    Every launch made through ``ops`` is bracketed by two CUDA events; ``summary`` returns, per kernel
    name, the launch count, total device milliseconds and the algorithmic FLOPs / bytes the launches
    covered (each distinct input element read once, each output written once at its stored dtype)."""

    def __init__(self):
        self._records = []

    def add(self, name, ev0, ev1, flops, nbytes):
        self._records.append((name, ev0, ev1, flops, nbytes))

    def __enter__(self):
        ops._PROFILER = self
        return self

    def __exit__(self, *exc):
        ops._PROFILER = None
        return False

    def summary(self) -> Dict[str, dict]:
        torch.cuda.synchronize()
        out = defaultdict(lambda: {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        for name, e0, e1, fl, nb in self._records:
            r = out[name]
            r["launches"] += 1
            r["ms"] += e0.elapsed_time(e1)
            r["flops"] += fl
            r["bytes"] += nb
        return dict(out)
