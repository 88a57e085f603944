"""Training half of the VUnet path: the backward schedule over the fused forward graph.

The reference trains through autograd (``loss.backward()`` in ``experiments/shape_and_pose_net.py:429-432``):
~600 eager forward kernels leave a graph of conv / ELU / cat / add / weight-norm nodes behind and autograd
replays it.  Here the forward pass of :class:`~.vunets.VunetAlter` appends one record per fused launch to
a :class:`Tape`; ONE ``torch.autograd.Function`` spans the whole model, and its backward walks the tape in
reverse issuing, per conv, one weight-correlation launch (``vub_conv_wgrad``) and one data-gradient launch
(``vub_conv_dgrad``, ELU' masks / residual adds / skip accumulation fused into its epilogue).  Parameter
gradients of all layers of a sub-network come from ONE ``vub_wn_bwd`` launch and are written straight into
flat per-sub-network buckets (``dd``, ``du``, ``ed``, ``eu`` — reverse forward order); ``param.grad`` are
views of those buckets, so the data-parallel gradient exchange is one in-place NCCL all-reduce per bucket
on a side stream, overlapped with the rest of the backward pass (replacing ``nn.DataParallel``'s
per-step broadcast / gather / reduce, ``shape_and_pose_net.py:213-214``).

Gradients flow from ``imgs``, ``means`` and ``logstds``; the ``activations`` list is returned detached
(no caller of the reference reads it).  The 25 parameters of the dead branch
(``ed.blocks[3]``, ``ed.ups[1]``, ``ed.fin_block``) keep ``grad = None`` like in the reference.
"""
from __future__ import annotations

import os

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _native as N
from . import ops
from .modules import Act, NormConv2d, WeightBank

OWNER_ORDER = ("dd", "du", "ed", "eu")  # reverse forward order = order in which backward finishes them


class GradSlot:
    """Gradient holder for a conv output that is not an :class:`Act` (fp32 latent-head input, image)."""

    __slots__ = ("grad",)

    def __init__(self):
        self.grad = None


class Tape:
    """Records of one forward pass, in launch order."""

    def __init__(self):
        self.recs: List[tuple] = []
        self.img_slot: Optional[GradSlot] = None
        self.latents: List[dict] = []
        self.samples: List[dict] = []   # VunetOrg: latent_sample records (posterior means first, then prior groups)

    def conv(self, conv: NormConv2d, srcs, residual: Optional[Act], out: Optional[Act], d2s: bool):
        slot = out if out is not None else GradSlot()
        self.recs.append(("conv", conv, list(srcs), residual, slot, d2s))
        return slot

    def stem(self, conv: NormConv2d, x: torch.Tensor, out: Act):
        self.recs.append(("stem", conv, x.contiguous(), out))

    # ---- VunetOrg family -------------------------------------------------------------------------------
    def sample(self, owner: str, slot: GradSlot, mean: torch.Tensor, z_act: Optional[Act], sub: int, group) -> None:
        """``latent_sample`` of the parameter conv whose output gradient is ``slot``: ``mean`` is the (merged) NCHW
        tensor the module returns, ``z_act`` the (merged) activation of the sample, or None when nothing reads it."""
        rec = {"owner": owner, "slot": slot, "mean": mean, "z_act": z_act, "sub": sub, "group": tuple(group)}
        self.samples.append(rec)
        self.recs.append(("sample", rec))

    def regroup(self, owner: str, src: Act, dsts: List[Act], group_major: bool) -> None:
        """SpaceToDepth of ``src``: ``dsts`` = [the 4C-channel activation] or, group-major, the four group
        activations, whose gradients are then created inside ONE stacked buffer (unread groups stay zero)."""
        buf = None
        if group_major:
            buf = torch.zeros((4,) + tuple(dsts[0].any.shape), dtype=dsts[0].any.dtype, device=dsts[0].any.device)
            for g, d in enumerate(dsts):
                d.grad_view = buf[g]
        self.recs.append(("regroup", owner, src, list(dsts), group_major, buf))

    def latent(self, ml_slot: GradSlot, res: dict, eps: torch.Tensor):
        rec = {"ml_slot": ml_slot, "logstd": res["logstd"], "eps": eps, "z_act": res["z_act"], "d_mu": None,
               "d_ls": None}
        self.latents.append(rec)
        self.recs.append(("latent", rec))


# --------------------------------------------------------------------------------------------------
def _tape_signature(tape: "Tape") -> tuple:
    """Which convs ran on which input shapes: the gradient bank (split counts, buffer sizes) depends on both."""
    sig = []
    for rec in tape.recs:
        if rec[0] == "conv":
            sig.append((id(rec[1]), tuple(rec[2][0][0].pick(rec[2][0][1]).shape)))
        elif rec[0] == "stem":
            sig.append((id(rec[1]), tuple(rec[2].shape)))
    return tuple(sig)


class GradBank:
    """Weight-correlation scratch + flat gradient buckets for the convs that appear on a tape."""

    def __init__(self, model: torch.nn.Module, bank: WeightBank, tape: Tape):
        convs: List[NormConv2d] = []   # slot leaders and stems, in tape order
        seen = set()
        for rec in tape.recs:
            if rec[0] in ("conv", "stem") and id(rec[1]) not in seen:
                seen.add(id(rec[1]))
                convs.append(rec[1])
        self.key = _tape_signature(tape)
        dev = next(model.parameters()).device
        # fp32 validation mode: fp32 activations, every conv (stems included) is a regular conv of the bank, output
        # rows are not padded, pixel splits are chosen here (vub_conv_wgrad_f32 takes any split count)
        fp32 = self.fp32 = bank.precision == "fp32"
        is_stem = (lambda cv: False) if fp32 else (lambda cv: cv._vub_stem)
        rows_of = (lambda cw: cw.cout) if fp32 else (lambda cw: cw.cout_k)
        members: Dict[int, List[NormConv2d]] = {}
        for slot in bank.slots:
            members[id(slot[0])] = slot
        # ---- correlation scratch per conv slot: partial correlations gt [n_split, rows = output channels, K] and
        #      partial sums dy_sum [n_split, rows], written (not accumulated) by vub_conv_wgrad; stems keep one
        #      atomically accumulated slice that is zeroed every step ----
        shapes = {}
        for rec in tape.recs:
            if rec[0] == "conv" and id(rec[1]) not in shapes:
                _, cv, srcs, _, _, d2s = rec
                cw = bank.get(cv)
                xs = tuple(srcs[0][0].pick(srcs[0][1]).shape)
                B_, H_, W_, _ = xs
                ho, wo = H_ // cw.stride, W_ // cw.stride
                if fp32:
                    shapes[id(cv)] = max(1, min(64, (B_ * ho * wo) // 512))
                    continue
                dys = (B_, 2 * ho, 2 * wo, cw.cout_k // 4) if d2s else (B_, ho, wo, cw.cout_k)
                shapes[id(cv)] = ops.wgrad_splits(dys, xs, cw.nsrc, cw.ksize, cw.stride, d2s)
        off, stem_floats = 0, 0
        place = {}
        for cv in convs:  # stems first: they share one zeroed region at the front of the buffer
            if is_stem(cv):
                k, rows = cv.in_channels, cv.out_channels
                place[id(cv)] = (off, k, rows, 1)
                off += (k + 1) * rows
        stem_floats = off
        for cv in convs:
            if not is_stem(cv):
                cw = bank.get(cv)
                k, rows = cw.nsrc * cw.ksize * cw.ksize * cw.c_src, rows_of(cw)
                ns = shapes[id(cv)]
                off = (off + 3) // 4 * 4
                place[id(cv)] = (off, k, rows, ns)
                off += ns * (k + 1) * rows
        self._gbuf = torch.zeros(off, dtype=torch.float32, device=dev)
        self._stem_region = self._gbuf[:stem_floats]
        self.gt, self.dy_sum = {}, {}
        for cv in convs:
            o, k, rows, ns = place[id(cv)]
            self.gt[id(cv)] = self._gbuf[o:o + ns * k * rows].view(ns, rows, k)
            self.dy_sum[id(cv)] = self._gbuf[o + ns * k * rows:o + ns * (k + 1) * rows].view(ns, rows)
        # ---- gradient buckets per owner, parameters in registration order ----
        live = []  # every NormConv2d that receives gradients
        for cv in convs:
            live.extend([cv] if is_stem(cv) else members[id(cv)])
        live_ids = {id(m) for m in live}
        self.buckets: Dict[str, torch.Tensor] = {}
        self.grads: Dict[int, torch.Tensor] = {}   # id(param) -> view of its bucket
        self.params: Dict[str, List[torch.nn.Parameter]] = {}
        self._tables = {}
        leader_of = {}
        for cv in convs:
            for j, m in enumerate([cv] if is_stem(cv) else members[id(cv)]):
                leader_of[id(m)] = cv
        for owner in OWNER_ORDER:
            sub = getattr(model, owner, None)
            if sub is None:
                continue
            mods = [m for m in sub.modules() if isinstance(m, NormConv2d) and id(m) in live_ids]
            if not mods:
                continue
            plist = []
            for m in mods:
                plist.extend([m.beta, m.gamma, m.conv.bias, m.conv.weight_g, m.conv.weight_v])
            total = sum(p.numel() for p in plist)
            flat = torch.zeros(total, dtype=torch.float32, device=dev)
            o = 0
            for p in plist:
                self.grads[id(p)] = flat[o:o + p.numel()].view(p.shape)
                o += p.numel()
            self.buckets[owner] = flat
            self.params[owner] = plist
            entries, rows, nbytes = [], 0, 0
            max_len = max(m.in_channels * m.kernel_size ** 2 for m in mods)
            for m in mods:
                cv = leader_of[id(m)]
                # algorithmic bytes: this conv's split partials read once, v read, dv written
                nbytes += 4 * m.conv.weight_v.numel() * (place[id(cv)][3] + 2)
                e = N.WnbEntry()
                e.v, e.g = m.conv.weight_v.data_ptr(), m.conv.weight_g.data_ptr()
                e.gamma, e.bias = m.gamma.data_ptr(), m.conv.bias.data_ptr()
                e.gt, e.dy_sum = self.gt[id(cv)].data_ptr(), self.dy_sum[id(cv)].data_ptr()
                e.d_v, e.d_g = self.grads[id(m.conv.weight_v)].data_ptr(), self.grads[id(m.conv.weight_g)].data_ptr()
                e.d_gamma, e.d_beta = self.grads[id(m.gamma)].data_ptr(), self.grads[id(m.beta)].data_ptr()
                e.d_bias = self.grads[id(m.conv.bias)].data_ptr()
                if is_stem(m):
                    c_src, col0 = m.in_channels, 0
                else:
                    cw = bank.get(cv)
                    c_src = cw.c_src
                    col0 = sum(x.out_channels for x in members[id(cv)][:members[id(cv)].index(m)])
                e.cout, e.cin, e.ktaps, e.c_src = m.out_channels, m.in_channels, m.kernel_size ** 2, c_src
                e.gt_pitch, e.gt_col0, e.row_start = place[id(cv)][1], col0, rows
                ns = place[id(cv)][3]
                e.n_split, e.split_stride, e.sum_stride = ns, place[id(cv)][1] * place[id(cv)][2], place[id(cv)][2]
                rows += m.out_channels
                entries.append(e)
            arr = (N.WnbEntry * len(entries))(*entries)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            self._tables[owner] = (host.to(dev), len(entries), rows, float(nbytes), max_len)
        self._ptr_key = tuple(p.data_ptr() for ps in self.params.values() for p in ps)

    def __deepcopy__(self, memo):
        return None  # pointer cache of one module tree (see WeightBank.__deepcopy__)

    def matches(self, tape: Tape, model) -> bool:
        if _tape_signature(tape) != self.key:
            return False
        return self._ptr_key == tuple(p.data_ptr() for ps in self.params.values() for p in ps)

    def zero(self) -> None:
        if self._stem_region.numel():
            self._stem_region.zero_()  # only the stems accumulate (atomics); conv partials are overwritten

    def finish_owner(self, owner: str) -> None:
        """All correlations of this sub-network are complete: produce its parameter gradients."""
        if owner not in self._tables:
            return
        plist = self.params[owner]
        # A parameter accumulates iff it still holds OUR view from an earlier backward (no zero_grad in between).
        # Mixed buckets (per-parameter zeroing, frozen subsets): the slices of the parameters that do NOT accumulate
        # are zeroed and the whole bucket runs with accumulate=True, so nobody's accumulated gradient is dropped.
        held = [p.grad is not None and p.grad.data_ptr() == self.grads[id(p)].data_ptr() for p in plist]
        accumulate = any(held)
        if accumulate and not all(held):
            for p, h in zip(plist, held):
                if not h:
                    self.grads[id(p)].zero_()
        extra = [(p, p.grad) for p, h in zip(plist, held) if not h and p.grad is not None]
        table, n, rows, nbytes, max_len = self._tables[owner]
        ops.wn_bwd(table, n, rows, accumulate, nbytes, max_len)
        for p in plist:
            if p.requires_grad:  # frozen parameters keep grad = None like under autograd
                p.grad = self.grads[id(p)]
        for p, old in extra:  # foreign gradient tensors (e.g. zero_grad(set_to_none=False) copies): keep their sum
            if p.requires_grad:
                p.grad.add_(old)


# --------------------------------------------------------------------------------------------------
_WGRAD_STREAMS: dict = {}  # one weight-gradient stream per device


def run_backward(model, bank: WeightBank, tape: Tape, d_imgs: Optional[torch.Tensor], reducer=None) -> None:
    """Walk the tape in reverse; parameter gradients land in ``param.grad`` (views of flat buckets)."""
    fp32 = bank.precision == "fp32"
    gb: Optional[GradBank] = getattr(model, "_vub_grad_bank", None)
    if gb is None or gb.fp32 != fp32 or not gb.matches(tape, model):
        gb = GradBank(model, bank, tape)
        object.__setattr__(model, "_vub_grad_bank", gb)
    gb.zero()
    if d_imgs is not None and tape.img_slot is not None:
        if fp32:
            tape.img_slot.grad = ops.nchw_to_nhwc32(d_imgs.float(), True, False)[0]
        else:
            tape.img_slot.grad = ops.nchw_to_nhwc_pad(d_imgs.float(), 32)
    # The weight gradient and the data gradient of a conv are independent: wgrad runs on a second stream, so the
    # small-map layers (10 us kernels that fill a fraction of the GPU) overlap pairwise and the tails of the large
    # ones interleave.  VUB_BWD_STREAMS=1 keeps everything on one stream.
    main = torch.cuda.current_stream()
    two = os.environ.get("VUB_BWD_STREAMS", "2") != "1" and not fp32
    ws = None
    if two:
        ws = _WGRAD_STREAMS.get(main.device)
        if ws is None:
            ws = _WGRAD_STREAMS[main.device] = torch.cuda.Stream(device=main.device)
    pending: List[tuple] = []  # (dy kept alive, event after its wgrad): dy must not be recycled while wgrad reads it

    def join_wgrads() -> None:
        if ws is not None:
            main.wait_stream(ws)
        pending.clear()

    current = None
    for rec in reversed(tape.recs):
        kind = rec[0]
        if kind == "latent":
            owner = "ed"
        elif kind == "sample":
            owner = rec[1]["owner"]
        elif kind == "regroup":
            owner = rec[1]
        else:
            owner = rec[1]._vub_owner
        if owner != current:
            if current is not None:
                join_wgrads()
                gb.finish_owner(current)
                if reducer is not None:
                    reducer.bucket_ready(current, gb.buckets.get(current))
            current = owner
        if kind == "conv":
            _, conv, srcs, residual, out, d2s = rec
            dy = out.grad
            if dy is None:
                continue
            cw = bank.get(conv)
            if fp32:  # validation mode: CUDA-core fp32 kernels, one stream
                ops.conv_wgrad32(dy, [a.pick(m) for a, m in srcs], gb.gt[id(conv)], gb.dy_sum[id(conv)], cout=cw.cout,
                                 ksize=cw.ksize, stride=cw.stride, d2s=d2s)
            elif ws is not None:
                ws.wait_stream(main)  # dy (and, the first time, the forward activations) are complete
                with torch.cuda.stream(ws):
                    ops.conv_wgrad(dy, [a.pick(m) for a, m in srcs], gb.gt[id(conv)], gb.dy_sum[id(conv)],
                                   ksize=cw.ksize, stride=cw.stride, d2s=d2s)
                    ev = torch.cuda.Event()
                    ev.record(ws)
                pending.append((dy, ev))
                if len(pending) > 6:  # bound the gradients kept alive: the oldest wgrad has long finished
                    main.wait_event(pending[0][1])
                    pending.pop(0)
            else:
                ops.conv_wgrad(dy, [a.pick(m) for a, m in srcs], gb.gt[id(conv)], gb.dy_sum[id(conv)], ksize=cw.ksize,
                               stride=cw.stride, d2s=d2s)
            dx, mask, add1, add2 = [], [], [], []
            for a, m in srcs:
                existed = a.grad is not None
                if not existed:
                    a.grad = a.grad_view if a.grad_view is not None else torch.empty_like(a.pick(m))
                dx.append(a.grad)
                mask.append(a.elu if m == "elu" else None)
                a1 = a.grad if existed else None
                a2 = dy if residual is a else None
                if a1 is None:
                    a1, a2 = a2, None
                add1.append(a1)
                add2.append(a2)
            if fp32:
                ops.conv_dgrad32(dy, cw.w32, cw.scale, dx, mask=mask, addend=add1, addend2=add2, cout=cw.cout,
                                 ksize=cw.ksize, stride=cw.stride, d2s=d2s)
            else:
                ops.conv_dgrad(dy, cw.wd, dx, mask=mask, addend=add1, addend2=add2, ksize=cw.ksize, stride=cw.stride,
                               d2s=d2s, wd_tiled=cw.wd_tiled)
            out.grad = None
        elif kind == "stem":
            _, conv, x, out = rec
            if out.grad is not None:
                ops.stem_bwd(x, out.grad, gb.gt[id(conv)][0], gb.dy_sum[id(conv)][0])
                out.grad = None
        elif kind == "sample":
            r = rec[1]
            dz = None if r["z_act"] is None else r["z_act"].grad
            d_mean = r.get("d_mean")
            if dz is None and d_mean is None:
                continue
            sub = r["sub"]
            hh, ww = r["mean"].shape[2] // sub, r["mean"].shape[3] // sub
            r["slot"].grad = ops.latent_sample_bwd(d_mean, dz, sub=sub, group=r["group"], h=hh, w=ww, f32=fp32)
            if r["z_act"] is not None:
                r["z_act"].grad = None
        elif kind == "regroup":
            _, _, src, dsts, group_major, buf = rec
            if group_major:
                if not any(d.grad is not None for d in dsts):
                    continue
                g = buf  # the groups' gradients live in its four slices (groups nobody read stay zero)
            else:
                g = dsts[0].grad
                if g is None:
                    continue
            if src.grad is None:
                src.grad = ops.space_to_depth(g, group_major=group_major, inverse=True)[0]
            else:
                ops.space_to_depth(g, group_major=group_major, inverse=True, accumulate_into=src.grad)
            for d in dsts:
                d.grad = None
        else:
            r = rec[1]
            dz = r["z_act"].grad
            if dz is None and r["d_mu"] is None and r["d_ls"] is None:
                continue
            r["ml_slot"].grad = (ops.latent_bwd32 if fp32 else ops.latent_bwd)(r["logstd"], r["eps"], r["d_mu"], r["d_ls"], dz)
            r["z_act"].grad = None
    join_wgrads()
    if current is not None:
        gb.finish_owner(current)
        if reducer is not None:
            reducer.bucket_ready(current, gb.buckets.get(current))
    if reducer is not None:
        reducer.finish()


class _VunetTrainFn(torch.autograd.Function):
    """The whole VunetAlter forward as ONE autograd node (see module docstring)."""

    @staticmethod
    def forward(ctx, model, x, c, eps, anchor):
        bank: WeightBank = model._bank(with_wd=True)
        tape = Tape()
        bank.tape = tape
        try:
            imgs, lat, hs_left = model._forward_impl(bank, x, c, eps)
        finally:
            bank.tape = None
        ctx.model, ctx.bank, ctx.tape = model, bank, tape
        ctx.set_materialize_grads(False)
        means = [r["mu"] for r in lat]
        logstds = [r["logstd"] for r in lat]
        hs = [h.nchw() for h in hs_left]
        ctx.mark_non_differentiable(*hs)
        ctx.n_lat = len(lat)
        return (imgs, *means, *logstds, *hs)

    @staticmethod
    def backward(ctx, d_imgs, *rest):
        tape: Tape = ctx.tape
        if tape is None:
            raise RuntimeError("vunet_b200: backward through the same forward pass twice is not supported")
        n = ctx.n_lat
        for i, rec in enumerate(tape.latents):
            d_mu, d_ls = rest[i], rest[n + i]
            rec["d_mu"] = None if d_mu is None else d_mu.float().contiguous()
            rec["d_ls"] = None if d_ls is None else d_ls.float().contiguous()
        run_backward(ctx.model, ctx.bank, tape, None if d_imgs is None else d_imgs.contiguous(),
                     reducer=getattr(ctx.model, "_vub_reducer", None))
        ctx.tape = None  # activations and gradient buffers are released here
        return None, None, None, None, None


class _VunetOrgTrainFn(torch.autograd.Function):
    """The whole VunetOrg forward as ONE autograd node: outputs (imgs, *qs, *ps, *hs, *gs); gradients flow from
    ``imgs``, the posterior means ``qs`` and the prior means ``ps`` (lib/losses.py:55-65 compares the two)."""

    @staticmethod
    def forward(ctx, model, x, c, eps, eps_groups, anchor):
        bank: WeightBank = model._bank(with_wd=True)
        tape = Tape()
        bank.tape = tape
        try:
            imgs, lat, ps, hs_left, gs = model._forward_impl(bank, x, c, eps, eps_groups)
        finally:
            bank.tape = None
        ctx.model, ctx.bank, ctx.tape = model, bank, tape
        ctx.set_materialize_grads(False)
        qs = [r["q"] for r in lat]
        hs = [h.nchw() for h in hs_left]
        gsn = [g.nchw() for g in gs]
        ctx.mark_non_differentiable(*hs, *gsn)
        ctx.n_lat, ctx.n_hs = len(lat), len(hs)
        return (imgs, *qs, *ps, *hs, *gsn)

    @staticmethod
    def backward(ctx, d_imgs, *rest):
        tape: Tape = ctx.tape
        if tape is None:
            raise RuntimeError("vunet_b200: backward through the same forward pass twice is not supported")
        n = ctx.n_lat
        # tape.samples: the n posterior-mean records of `ed`, then four prior-group records per latent scale of `dd`
        for i in range(n):
            d_q, d_p = rest[i], rest[n + i]
            tape.samples[i]["d_mean"] = None if d_q is None else d_q.float().contiguous()
            d_p = None if d_p is None else d_p.float().contiguous()
            for l in range(4):
                tape.samples[n + 4 * i + l]["d_mean"] = d_p
        run_backward(ctx.model, ctx.bank, tape, None if d_imgs is None else d_imgs.contiguous(),
                     reducer=getattr(ctx.model, "_vub_reducer", None))
        ctx.tape = None
        return None, None, None, None, None, None


def train_forward_org(model, x, c, eps, eps_groups):
    """``VunetOrg.forward`` under autograd: returns (imgs, qs, ps, hs_left, gs)."""
    if x.requires_grad or c.requires_grad:
        raise NotImplementedError("vunet_b200: VunetOrg.forward is differentiable w.r.t. its parameters only; "
                                  "gradients w.r.t. the inputs x / c are not produced (detach them)")
    anchor = torch.zeros(1, device=x.device, requires_grad=True)  # fresh leaf per call, see train_forward
    outs = _VunetOrgTrainFn.apply(model, x, c, None if eps is None else list(eps),
                                  None if eps_groups is None else [list(g) for g in eps_groups], anchor)
    n = model.ed.n_scales
    n_hs = len(outs) - 1 - 2 * n - 2 * model.du.n_scales
    imgs, qs, ps = outs[0], list(outs[1:1 + n]), list(outs[1 + n:1 + 2 * n])
    hs = list(outs[1 + 2 * n:1 + 2 * n + n_hs])
    gs = list(outs[1 + 2 * n + n_hs:])
    return imgs, qs, ps, hs, gs


def train_forward(model, x, c, eps: Optional[Sequence[torch.Tensor]]):
    """``VunetAlter.forward`` under autograd: returns (imgs, means, logstds, hs_left)."""
    if x.requires_grad or c.requires_grad:
        raise NotImplementedError("vunet_b200: VunetAlter.forward is differentiable w.r.t. its parameters only; "
                                  "gradients w.r.t. the inputs x / c are not produced (detach them)")
    # A FRESH leaf per call: the autograd engine makes the caller's stream wait for the stream an input's gradient
    # accumulator was created on.  A cached leaf keeps the accumulator of an earlier (still referenced) graph — e.g. made
    # on the default stream — and a backward pass under CUDA-graph capture then fails with "dependency created on
    # uncaptured work in another stream".
    anchor = torch.zeros(1, device=x.device, requires_grad=True)
    outs = _VunetTrainFn.apply(model, x, c, None if eps is None else list(eps), anchor)
    n = model.ed.n_scales
    imgs, means, logstds, hs = outs[0], list(outs[1:1 + n]), list(outs[1 + n:1 + 2 * n]), list(outs[1 + 2 * n:])
    return imgs, means, logstds, hs


# --------------------------------------------------------------------------------------------------
class BucketReducer:
    """Data-parallel gradient exchange: one in-place all-reduce (mean) per flat bucket, issued on a side
    stream the moment the bucket's ``vub_wn_bwd`` has been enqueued, so it overlaps the remaining backward
    launches; ``finish`` makes the compute stream wait for all of them (before ``optimizer.step``).

    On CUDA with the NCCL backend the exchange goes through the package's own C ABI —
    ``vub_allreduce_bucket(flat, n, ncclComm, average, stream)``: ONE ncclAllReduce with ``ncclAvg`` per bucket (the
    1/world scaling happens inside the collective, no separate scaling kernel), on a communicator this object
    creates over the same ranks (the 128-byte id travels through ``torch.distributed``).  Any other backend (gloo
    in the CPU tests and on single-GPU boxes) falls back to ``torch.distributed.all_reduce``."""

    def __init__(self, process_group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        self._stream = None
        self._comm = None
        self.bytes_last = 0
        self.bytes_per_step = 0
        self.native = False
        if (self.world > 1 and dist.get_backend(process_group) == "nccl" and torch.cuda.is_available()
                and os.environ.get("VUB_NATIVE_ALLREDUCE", "1") != "0"):
            self._init_native()

    def _init_native(self) -> None:
        import ctypes as C
        lib = N.lib()
        ident = (C.c_char * 128)()
        if self.rank == 0:
            N.check(lib.vub_nccl_unique_id(C.cast(ident, C.c_void_p)), "vub_nccl_unique_id")
        box = [bytes(ident)]
        src = self.dist.get_global_rank(self.group, 0) if self.group is not None else 0
        self.dist.broadcast_object_list(box, src=src, group=self.group)
        ident = (C.c_char * 128).from_buffer_copy(box[0])
        comm = C.c_void_p()
        N.check(lib.vub_nccl_comm_init(C.byref(comm), self.world, C.cast(ident, C.c_void_p), self.rank),
                "vub_nccl_comm_init")
        self._comm = comm
        self.native = True

    def close(self) -> None:
        """Destroy the communicator.  Call it only after every CUDA graph that captured an exchange on it has been
        destroyed (NCCL keeps captured graphs' references alive); never called implicitly — at interpreter exit the
        process simply ends."""
        if self._comm is not None:
            torch.cuda.synchronize()
            N.check(N.lib().vub_nccl_comm_destroy(self._comm), "vub_nccl_comm_destroy")
            self._comm = None
            self.native = False

    def bucket_ready(self, owner: str, flat: Optional[torch.Tensor]) -> None:
        if flat is None or self.world == 1:
            return
        self.bytes_last += flat.numel() * flat.element_size()
        if flat.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=flat.device)
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(ev)
                if self.native:
                    N.check(N.lib().vub_allreduce_bucket(flat.data_ptr(), flat.numel(), self._comm, 1,
                                                         self._stream.cuda_stream), "vub_allreduce_bucket")
                else:
                    flat.div_(self.world)
                    self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, group=self.group)
            flat.record_stream(self._stream)
        else:
            flat.div_(self.world)
            self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, group=self.group)

    def finish(self) -> None:
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        self.bytes_per_step, self.bytes_last = self.bytes_last, 0


# --------------------------------------------------------------------------------------------------
class GraphedTrainStep:
    """One training step — forward, loss, backward (incl. the data-parallel gradient exchange) and the optimizer
    step — captured in ONE CUDA graph and replayed.

    Why: the step is ~300 launches, many of them 10 us kernels on the small maps, and the Python / ctypes path
    costs more host time per launch than those kernels run — measured with ``tools/train_host_overhead.py``, the
    host's enqueue time EQUALS the step time (the GPU is starved).  Replay removes the host from the loop.

    ``loss_fn(imgs, means, logstds, *extra) -> scalar`` is traced once; ``extra`` are additional static inputs
    (the target image).  The optimizer must be :class:`~bdvs_b200.FusedAdam` (its step count lives on the device,
    so the captured launch is valid for every step).  Warm-up steps needed before capture run on a snapshot: the
    parameters and optimizer state are restored afterwards.  Replaces the body of the reference's ``train_fn``
    (experiments/shape_and_pose_net.py:360-466, minus logging)::

        step = GraphedTrainStep(vunet, opt, loss_fn, x, c, target)
        for x, c, target in loader:
            loss = step(x, c, target)        # copies into the static inputs, replays; ``loss`` is a device scalar
    """

    def __init__(self, model, optimizer, loss_fn, *example_inputs: torch.Tensor, warmup: int = 3,
                 eps: Optional[Sequence[torch.Tensor]] = None):
        """``eps``: optional static latent-noise tensors (``model(x, c, eps=...)``; copy new noise into
        ``self.static_eps`` between steps); default: drawn inside the graph by torch's graph-safe generator."""
        from .modules import bump_weight_epoch
        from .optim import FusedAdam

        if not isinstance(optimizer, FusedAdam):
            raise TypeError("GraphedTrainStep needs bdvs_b200.FusedAdam (device-side step count)")
        if len(example_inputs) < 2:
            raise ValueError("GraphedTrainStep(model, optimizer, loss_fn, x, c, *extra)")
        self.model, self.opt, self.loss_fn = model, optimizer, loss_fn
        self.static_in = [t.detach().clone() for t in example_inputs]
        self.static_eps = None if eps is None else [t.detach().clone() for t in eps]
        params = list(model.parameters())
        snap = [p.detach().clone() for p in params]
        m_snap, s_snap, step_snap = optimizer._exp_avg.clone(), optimizer._exp_avg_sq.clone(), optimizer._step
        stepped_snap = set(optimizer._stepped)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):  # allocates every cache (weight bank, gradient bank, optimizer table)
                self._one_step()
        cur.wait_stream(side)
        with torch.no_grad():
            for p, s in zip(params, snap):
                p.copy_(s)
            optimizer._exp_avg.copy_(m_snap)
            optimizer._exp_avg_sq.copy_(s_snap)
        optimizer._step = step_snap
        optimizer._step_dev.fill_(step_snap)
        optimizer._stepped = stepped_snap
        bump_weight_epoch()
        model._bank(with_wd=True).refresh(with_wd=True)  # re-prepare the restored weights OUTSIDE the graph
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._one_step()
        optimizer._step = step_snap  # capture ran the Python side of step() once without executing a kernel
        # the graph has these objects' device buffers baked in: keep them alive even if an eager step with another
        # batch shape later replaces the model's gradient bank
        self._bank = model._bank(with_wd=True)
        self._grad_bank = getattr(model, "_vub_grad_bank", None)
        self._live_ids = {id(p) for p in optimizer._params if p.grad is not None}

    # ---- host-resident batches: upload batch k+1 while step k runs ------------------------------------
    def prefetch(self, *host_inputs: torch.Tensor) -> None:
        """Enqueue the host -> device copy of the NEXT batch (pinned tensors) on a copy stream; it overlaps the
        step that is running.  Pair every call with one :meth:`run_prefetched`."""
        if len(host_inputs) != len(self.static_in):
            raise ValueError(f"expected {len(self.static_in)} inputs")
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream()
            self._staging = [[torch.empty_like(t) for t in self.static_in] for _ in range(2)]
            self._staged_ev = [None, None]   # upload of set k finished
            self._drained_ev = [None, None]  # set k has been copied into the static inputs
            self._put, self._get = 0, 0
        k = self._put & 1
        with torch.cuda.stream(self._copy_stream):
            if self._drained_ev[k] is not None:
                self._copy_stream.wait_event(self._drained_ev[k])
            for dst, src in zip(self._staging[k], host_inputs):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._staged_ev[k] = ev
        self._put += 1

    def run_prefetched(self) -> torch.Tensor:
        """One step on the oldest prefetched batch."""
        if not hasattr(self, "_copy_stream") or self._get >= self._put:
            raise RuntimeError("run_prefetched() without a pending prefetch()")
        k = self._get & 1
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged_ev[k])
        for dst, src in zip(self.static_in, self._staging[k]):
            dst.copy_(src, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cur)
        self._drained_ev[k] = ev
        self._get += 1
        return self(*self.static_in)

    def _one_step(self) -> torch.Tensor:
        x, c, *extra = self.static_in
        if self.static_eps is None:
            imgs, means, logstds, _ = self.model(x, c)
        else:
            imgs, means, logstds, _ = self.model(x, c, eps=self.static_eps)
        loss = self.loss_fn(imgs, means, logstds, *extra)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        if len(inputs) != len(self.static_in):
            raise ValueError(f"expected {len(self.static_in)} inputs")
        for dst, src in zip(self.static_in, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.opt.sync_lr()  # learning-rate schedules: the optimizer's device table is updated in place
        self.graph.replay()
        # The replayed Adam kernel rewrote the parameters through raw pointers: no autograd version changed.  Every
        # OTHER prepared-weight cache (sub-module solo banks used by ``vunet.ed(vunet.eu(x))``, NormConv2d._solo_bank,
        # the fp32 validation bank) must re-prepare; the training bank itself was re-emitted by the same kernel.
        from .modules import bump_weight_epoch
        bump_weight_epoch()
        self._bank.mark_current()
        self.opt._step += 1
        if self._live_ids is not None:
            self.opt._stepped.update(self._live_ids)
            self._live_ids = None
        return self.static_loss
