"""Fused Adam for the VUnet's NormConv2d parameters.

The reference trains with ``torch.optim.Adam`` over four parameter groups eu / ed / du / dd with
``betas = config["training"]["adam_betas"]`` (experiments/shape_and_pose_net.py:237-246) and re-normalises
every weight in the next forward (``torch.nn.utils.weight_norm`` hook, lib/modules.py:135-138).  Here both
are ONE launch of ``vub_adam_wn_step``: a block per output channel updates that channel's ``weight_v`` row
and its four scalars (``weight_g, gamma, beta, bias``) in place and then emits the prepared weights the next
forward / backward read (packed bf16 ``v/||v||``, folded scale / shift, data-gradient weights), so the
separate ``vub_wn_prep`` pass disappears from the training step.

Arithmetic is torch.optim.Adam's (no weight decay, no amsgrad, not maximize):
``m = b1*m + (1-b1)*g; s = b2*s + (1-b2)*g*g; p -= lr/(1-b1^t) * m / (sqrt(s)/sqrt(1-b2^t) + eps)``.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Union

import torch

from . import _native as N
from . import ops
from .modules import NormConv2d, bump_weight_epoch

_FIELDS = ("v", "g", "gamma", "beta", "bias")


def _five(m: NormConv2d) -> List[torch.nn.Parameter]:
    return [m.conv.weight_v, m.conv.weight_g, m.gamma, m.beta, m.conv.bias]


class FusedAdam:
    """``FusedAdam(params, lr=..., betas=..., eps=..., model=vunet)`` — ``params`` as for
    ``torch.optim.Adam`` (an iterable of parameters or of group dicts with ``"params"`` and optional
    ``"lr"`` / ``"name"``); ``model`` is the module tree that owns them (to find each parameter's
    NormConv2d and the model's prepared-weight bank).  ``param_groups[i]["lr"]`` may be changed between
    steps (schedulers); betas / eps are shared by all groups.

    Per step: ``loss.backward(); opt.step(); opt.zero_grad()``.  Convs whose parameters have no gradient
    (not on the executed path) are skipped, like torch does; every stepped conv shares one step counter.
    ``state_dict()`` / ``load_state_dict()`` use torch.optim.Adam's layout, so the reference's optimizer
    checkpoints (``optimizer.load_state_dict(op_ckpt)``, shape_and_pose_net.py:248-249) load.
    """

    def __init__(self, params: Union[Iterable[torch.nn.Parameter], Iterable[dict]], lr: float = 1e-3,
                 betas: Sequence[float] = (0.9, 0.999), eps: float = 1e-8, *, model: torch.nn.Module):
        groups = list(params)
        if not groups:
            raise ValueError("FusedAdam: empty parameter list")
        if not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.defaults = dict(lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps), weight_decay=0,
                             amsgrad=False, maximize=False)
        self.param_groups: List[dict] = []
        for g in groups:
            g = dict(g)
            g["params"] = list(g["params"])
            for k, v in self.defaults.items():
                g.setdefault(k, v)
            if g["weight_decay"] or g["amsgrad"] or g["maximize"]:
                raise NotImplementedError("FusedAdam: weight_decay / amsgrad / maximize are not implemented")
            if tuple(g["betas"]) != self.defaults["betas"] or g["eps"] != self.defaults["eps"]:
                raise NotImplementedError("FusedAdam: betas and eps are shared by all parameter groups")
            self.param_groups.append(g)
        self.model = model
        owner: Dict[int, NormConv2d] = {}
        for m in model.modules():
            if isinstance(m, NormConv2d):
                for p in _five(m):
                    owner[id(p)] = m
        self._params: List[torch.nn.Parameter] = [p for g in self.param_groups for p in g["params"]]
        if len({id(p) for p in self._params}) != len(self._params):
            raise ValueError("FusedAdam: a parameter appears in more than one group")
        self._group_of = {id(p): gi for gi, g in enumerate(self.param_groups) for p in g["params"]}
        self._convs: List[NormConv2d] = []
        seen = set()
        for p in self._params:
            m = owner.get(id(p))
            if m is None:
                raise NotImplementedError("FusedAdam only steps NormConv2d parameters (the VUnet has no others)")
            if id(m) not in seen:
                seen.add(id(m))
                self._convs.append(m)
        for m in self._convs:
            gis = {self._group_of.get(id(p)) for p in _five(m)}
            if len(gis) != 1 or None in gis:
                raise ValueError("FusedAdam: the five parameters of a NormConv2d must share one group")
        dev = self._params[0].device
        if dev.type != "cuda":
            raise RuntimeError("vunet_b200 has no CPU path: move the module to a CUDA device")
        total = sum(p.numel() for p in self._params)
        self._exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self._exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self._m, self._s, o = {}, {}, 0
        for p in self._params:
            self._m[id(p)] = self._exp_avg[o:o + p.numel()].view(p.shape)
            self._s[id(p)] = self._exp_avg_sq[o:o + p.numel()].view(p.shape)
            o += p.numel()
        self._step = 0
        # the step count the kernel uses lives on the device (bias corrections computed in-kernel), so the launch
        # arguments are the same every step and a captured CUDA graph of the step can be replayed
        self._step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._stepped = set()  # id(param) of every parameter that has Adam state
        self._key = None
        self._table = None

    # ---------------------------------------------------------------------------------------------
    def _build(self, live: List[NormConv2d], bank) -> None:
        entries, rows = [], 0
        for m in live:
            e = N.AdamEntry()
            pe = bank._entry_of.get(id(m))
            if pe is not None:  # tensor-core conv: embed its weight-norm entry (prepared outputs included)
                for f, _ in N.WnEntry._fields_:
                    setattr(e.wn, f, getattr(pe, f))
            else:               # stem: the stem kernels read the raw parameters, nothing to prepare
                e.wn.v, e.wn.g = m.conv.weight_v.data_ptr(), m.conv.weight_g.data_ptr()
                e.wn.gamma, e.wn.beta, e.wn.bias = m.gamma.data_ptr(), m.beta.data_ptr(), m.conv.bias.data_ptr()
                e.wn.cout, e.wn.cin, e.wn.ktaps = m.out_channels, m.in_channels, m.kernel_size ** 2
                e.wn.c_src, e.wn.ktot = m.in_channels, m.in_channels * m.kernel_size ** 2
            e.wn.row_start = rows
            rows += m.out_channels
            for f, p in zip(_FIELDS, _five(m)):
                setattr(e, "d_" + f, p.grad.data_ptr())
                setattr(e, "m_" + f, self._m[id(p)].data_ptr())
                setattr(e, "s_" + f, self._s[id(p)].data_ptr())
            e.lr = float(self.param_groups[self._group_of[id(m.gamma)]]["lr"])
            entries.append(e)
        arr = (N.AdamEntry * len(entries))(*entries)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        if self._table is not None and self._table[0].numel() == host.numel():
            # same layout (e.g. only a learning rate changed): update IN PLACE, so a CUDA graph that captured the
            # table's address (training.GraphedTrainStep) sees the new values
            self._table[0].copy_(host)
            self._table = (self._table[0], len(entries), rows)
        else:
            self._table = (host.to(self._exp_avg.device), len(entries), rows)
        self._last_build = (live, bank)
        # algorithmic bytes of one step: p, g, m, s read + p, m, s written (28 B per parameter) + the
        # packed bf16 outputs (forward matrix and, in training, the data-gradient matrix) of the tensor-core convs
        self._nbytes = float(sum(28 * sum(p.numel() for p in _five(m))
                                 + (4 * m.conv.weight_v.numel() if bank._entry_of.get(id(m)) is not None else 0)
                                 for m in live))

    def sync_lr(self) -> None:
        """Push changed ``param_groups[i]["lr"]`` values into the device table without a step (the table is updated
        in place; ``GraphedTrainStep`` calls this before every replay, ``step()`` does it implicitly)."""
        lrs = tuple(float(g["lr"]) for g in self.param_groups)
        if self._key is not None and lrs != self._key[-1] and getattr(self, "_last_build", None) is not None:
            live, bank = self._last_build
            self._build(live, bank)
            self._key = self._key[:-1] + (lrs,)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        live = []
        for m in self._convs:
            have = [p.grad is not None for p in _five(m)]
            if all(have):
                live.append(m)
            elif any(have):
                raise RuntimeError("FusedAdam: a NormConv2d has gradients for only some of its parameters")
        if not live:
            return loss
        bank = self.model._bank(with_wd=True)
        bank.refresh(with_wd=True)  # (re)allocates the prepared buffers if parameters moved; a no-op otherwise
        key = (id(bank), bank._table.data_ptr(),
               tuple(p.data_ptr() for m in live for p in _five(m)),
               tuple(p.grad.data_ptr() for m in live for p in _five(m)),
               tuple(float(g["lr"]) for g in self.param_groups))
        for m in live:
            for p in _five(m):
                if p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                    raise ValueError("FusedAdam: gradients must be contiguous fp32")
        if key != self._key:
            self._build(live, bank)
            self._key = key
            self._stepped.update(id(p) for m in live for p in _five(m))
        self._step += 1
        b1, b2 = self.defaults["betas"]
        table, n, rows = self._table
        ops.adam_wn_step_dev(table, n, rows, b1, b2, self.defaults["eps"], self._step_dev, getattr(self, "_nbytes", 0.0))
        # parameters were written through raw pointers: autograd version counters did not move, so tell every
        # other cache of prepared weights (sub-module banks, the fp32 validation bank) that they are stale; the
        # training bank was re-emitted by this very launch.
        bump_weight_epoch()
        bank.mark_current()  # convs that were not stepped did not change either
        return loss

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self._params:
            if p.grad is None:
                continue
            if set_to_none:
                p.grad = None
            else:
                p.grad.zero_()

    # ---- torch.optim.Adam-compatible checkpoints -------------------------------------------------
    def state_dict(self) -> dict:
        state = {}
        if self._step > 0:
            for i, p in enumerate(self._params):
                if id(p) not in self._stepped:
                    continue
                state[i] = {"step": torch.tensor(float(self._step)), "exp_avg": self._m[id(p)].clone(),
                            "exp_avg_sq": self._s[id(p)].clone()}
        groups, o = [], 0
        for g in self.param_groups:
            d = {k: v for k, v in g.items() if k != "params"}
            d["params"] = list(range(o, o + len(g["params"])))
            o += len(g["params"])
            groups.append(d)
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd: dict) -> None:
        if len(sd["param_groups"]) != len(self.param_groups):
            raise ValueError("FusedAdam.load_state_dict: different number of parameter groups")
        for g, src in zip(self.param_groups, sd["param_groups"]):
            if len(src["params"]) != len(g["params"]):
                raise ValueError("FusedAdam.load_state_dict: parameter group sizes differ")
            if "lr" in src:
                g["lr"] = float(src["lr"])
        steps = set()
        for i, p in enumerate(self._params):
            st = sd["state"].get(i)
            if st is None:
                self._stepped.discard(id(p))
                self._m[id(p)].zero_()
                self._s[id(p)].zero_()
                continue
            self._m[id(p)].copy_(st["exp_avg"].to(self._exp_avg.device, torch.float32))
            self._s[id(p)].copy_(st["exp_avg_sq"].to(self._exp_avg.device, torch.float32))
            steps.add(int(float(st["step"])))
            self._stepped.add(id(p))
        if len(steps) > 1:
            raise NotImplementedError("FusedAdam keeps one step counter; the checkpoint has several")
        self._step = steps.pop() if steps else 0
        self._step_dev.fill_(self._step)
        self._key = None
