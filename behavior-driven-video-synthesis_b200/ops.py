"""Torch-op layer over the C ABI: ``torch.ops.vunet_b200.*``.

Each op validates its tensors, allocates outputs with torch (device memory is torch's job) and
launches the sm_100a kernel on torch's current CUDA stream through ``_native``.  The ops are
registered for the CUDA dispatch key only — calling them with CPU tensors raises, by design.

Activation convention inside the path: NHWC bf16 tensors ``[B, H, W, C]``; an activation usually
exists as a (raw, elu) pair because every consumer wants one of the two.
"""
from __future__ import annotations

import ctypes as C
import functools
from typing import List, Optional, Sequence, Tuple

import torch

from . import _native as N

_LIB = torch.library.Library("vunet_b200", "DEF")

_LIB.define(
    "conv_fwd(Tensor x0, Tensor? x1, Tensor w, Tensor scale, Tensor shift, Tensor? residual, "
    "int ksize, int stride, int cout_valid, bool want_raw, bool want_elu, bool d2s, int f32_mode, "
    "int stages, int elu_in=0, Tensor? w_tiled=None, bool relu=False, int act=0) -> Tensor[]"
)
_LIB.define("wn_prep(Tensor table, int n_entries, int total_rows) -> ()")
_LIB.define(
    "stem_fwd(Tensor x, Tensor v, Tensor g, Tensor gamma, Tensor beta, Tensor bias, bool want_raw, "
    "bool want_elu) -> Tensor[]"
)
_LIB.define("latent_fwd(Tensor ml, Tensor? eps, bool want_elu_z, bool want_elu_mu) -> Tensor[]")
_LIB.define("nchw_to_nhwc(Tensor x, bool want_raw, bool want_elu) -> Tensor[]")
_LIB.define("nhwc_to_nchw(Tensor x) -> Tensor")
_LIB.define("frames_to_uint8(Tensor imgs) -> Tensor")
# ---- fp32 validation path -----------------------------------------------------------------------
_LIB.define(
    "conv_fwd32(Tensor x0, Tensor? x1, Tensor w, Tensor scale, Tensor shift, Tensor? residual, int cout, int ksize, "
    "int stride, bool want_raw, bool want_elu, bool d2s, bool nchw, bool relu=False) -> Tensor[]"
)
_LIB.define("nchw_to_nhwc32(Tensor x, bool want_raw, bool want_elu) -> Tensor[]")
_LIB.define(
    "conv_dgrad32(Tensor dy, Tensor w, Tensor scale, Tensor(a!) dx0, Tensor(b!)? dx1, Tensor? mask0, Tensor? mask1, "
    "Tensor? add0, Tensor? add1, Tensor? add20, Tensor? add21, int cout, int ksize, int stride, bool d2s, "
    "bool relu=False) -> ()"
)
_LIB.define(
    "conv_wgrad32(Tensor dy, Tensor x0, Tensor? x1, Tensor(a!) gt, Tensor(b!) dy_sum, int cout, int ksize, int stride, "
    "bool d2s) -> ()"
)
_LIB.define("latent_bwd32(Tensor logstd, Tensor? eps, Tensor? d_mu, Tensor? d_ls, Tensor? dz) -> Tensor")
# ---- training half (mutating ops: gradient buffers are accumulated in place) ----------------------
_LIB.define(
    "conv_dgrad(Tensor dy, Tensor wd, Tensor(a!) dx0, Tensor(b!)? dx1, Tensor? mask0, Tensor? mask1, "
    "Tensor? add0, Tensor? add1, Tensor? add20, Tensor? add21, int ksize, int stride, bool d2s, "
    "Tensor? wd_tiled=None, bool relu=False) -> ()"
)
_LIB.define(
    "conv_wgrad(Tensor dy, Tensor x0, Tensor? x1, Tensor(a!) gt, Tensor(b!) dy_sum, int ksize, int stride, "
    "bool d2s) -> ()"
)
_LIB.define("wn_bwd(Tensor table, int n_entries, int total_rows, bool accumulate, float nbytes=0.0, "
            "int max_len=0) -> ()")
_LIB.define("adam_wn_step(Tensor table, int n_entries, int total_rows, float beta1, float beta2, float eps, int step, "
            "float nbytes=0.0) -> ()")
_LIB.define("adam_wn_step_dev(Tensor table, int n_entries, int total_rows, float beta1, float beta2, float eps, "
            "Tensor(a!) step, float nbytes=0.0) -> ()")
_LIB.define("stem_bwd(Tensor x, Tensor dy, Tensor(a!) gt, Tensor(b!) dy_sum) -> ()")
_LIB.define("latent_bwd(Tensor logstd, Tensor? eps, Tensor? d_mu, Tensor? d_ls, Tensor? dz) -> Tensor")
_LIB.define("nchw_to_nhwc_pad(Tensor x, int c_pad) -> Tensor")
# ---- VGG19 perceptual loss: memory-bound pieces around the conv kernels ------------------------------
_LIB.define("vgg_normalize(Tensor x, bool want_nhwc, bool want_nchw, bool f32=False) -> Tensor[]")
_LIB.define("vgg_input_l1(Tensor pred, Tensor target) -> Tensor")
_LIB.define("vgg_input_bwd(Tensor pred, Tensor target, Tensor? d_xn, Tensor? g0, float scale0) -> Tensor")
_LIB.define("maxpool2_fwd(Tensor x) -> Tensor")
_LIB.define("maxpool2_relu_bwd(Tensor x, Tensor dy, Tensor? addend) -> Tensor")
_LIB.define("l1_partial(Tensor pf, Tensor tf) -> Tensor")
_LIB.define("l1_relu_grad(Tensor pf, Tensor tf, Tensor g, float scale) -> Tensor")
# ---- VunetOrg family: latent_sample + 4-group prior bookkeeping -----------------------------------
_LIB.define("latent_sample(Tensor p, Tensor? eps, int sub, int gi, int gj, Tensor(a!)? mean, Tensor(b!)? z, "
            "Tensor(c!)? z_raw, Tensor(d!)? z_elu, bool want_grp_elu) -> Tensor[]")
_LIB.define("latent_sample_bwd(Tensor? d_mean, Tensor? dz, int sub, int gi, int gj, int h, int w, bool f32) -> Tensor")
_LIB.define("space_to_depth(Tensor x, bool group_major, bool inverse, bool want_copy, bool want_elu, "
            "Tensor(a!)? accumulate_into=None) -> Tensor[]")

F32_NONE, F32_NHWC, F32_NCHW, U8_HWC = 0, 1, 2, 3

# ---- launch accounting (bench.py's "gpu_launches" and per-kernel roofline) -----------------------
LAUNCHES = 0          # kernels of THIS package launched since import (graph replays are counted by the caller)
_PROFILER = None      # set by profiler.KernelProfiler: records CUDA events around every launch


def _launch(name: str, fn, *args, flops: float = 0.0, bytes_: float = 0.0) -> int:
    """Run one C-ABI launcher; count it and, when a profiler is active, time it with CUDA events
    on the launching stream."""
    global LAUNCHES
    LAUNCHES += 1
    if _PROFILER is None:
        return fn(*args)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    rc = fn(*args)
    ev1.record()
    _PROFILER.add(name, ev0, ev1, flops, bytes_)
    return rc


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _check_act(t: torch.Tensor, name: str) -> None:
    if t.dtype != torch.bfloat16 or t.dim() != 4 or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous NHWC bf16 [B,H,W,C], got {t.dtype} {tuple(t.shape)}")


def _conv_fwd_cuda(x0, x1, w, scale, shift, residual, ksize, stride, cout_valid, want_raw, want_elu,
                   d2s, f32_mode, stages, elu_in=0, w_tiled=None, relu=False, act=0) -> List[torch.Tensor]:
    _check_act(x0, "conv_fwd x0")
    B, H, W, c_src = x0.shape
    nsrc = 1
    if x1 is not None:
        _check_act(x1, "conv_fwd x1")
        if x1.shape != x0.shape:
            raise ValueError("conv_fwd: both sources must have the same shape")
        nsrc = 2
    cout = w.shape[0]
    if w.dtype != torch.bfloat16 or w.shape[1] != nsrc * ksize * ksize * c_src:
        raise ValueError(f"conv_fwd: packed weight {tuple(w.shape)} does not match nsrc={nsrc} k={ksize} c_src={c_src}")
    Ho, Wo = H // stride, W // stride
    d = N.ConvDesc()
    d.x[0] = x0.data_ptr()
    d.x[1] = _p(x1)
    d.w = w.data_ptr()
    if w_tiled is not None:
        if w_tiled.dtype != torch.bfloat16 or w_tiled.numel() != w.numel() or c_src % 64:
            raise ValueError("conv_fwd: w_tiled must be the tile-major copy of w (c_src a multiple of 64)")
        d.w_tiled = w_tiled.data_ptr()
    d.scale = scale.data_ptr()
    d.shift = shift.data_ptr()
    d.residual = _p(residual)
    outs: List[torch.Tensor] = []
    flags = 0
    if f32_mode == F32_NONE:
        if d2s:
            flags |= N.CONV_D2S
            oshape = (B, 2 * Ho, 2 * Wo, cout // 4)
        else:
            oshape = (B, Ho, Wo, cout)
        if residual is not None:
            _check_act(residual, "conv_fwd residual")
            if tuple(residual.shape) != oshape:
                raise ValueError("conv_fwd: residual shape mismatch")
        if want_raw:
            raw = torch.empty(oshape, dtype=torch.bfloat16, device=x0.device)
            d.out_raw = raw.data_ptr()
            outs.append(raw)
        if want_elu:
            elu = torch.empty(oshape, dtype=torch.bfloat16, device=x0.device)
            d.out_elu = elu.data_ptr()
            outs.append(elu)
    else:
        if f32_mode == F32_NHWC:
            flags |= N.CONV_F32_NHWC
            out = torch.empty((B, Ho, Wo, cout), dtype=torch.float32, device=x0.device)
        elif f32_mode == U8_HWC:
            flags |= N.CONV_U8_HWC
            out = torch.empty((B, Ho, Wo, cout_valid), dtype=torch.uint8, device=x0.device)
        else:
            flags |= N.CONV_F32_NCHW
            out = torch.empty((B, cout_valid, Ho, Wo), dtype=torch.float32, device=x0.device)
        d.out_f32 = out.data_ptr()
        outs.append(out)
    d.nsrc, d.c_src, d.batch, d.hin, d.win = nsrc, c_src, B, H, W
    d.cout, d.cout_valid, d.ksize, d.stride = cout, cout_valid, ksize, stride
    if elu_in & 1:
        flags |= N.CONV_ELU_IN0
    if elu_in & 2:
        flags |= N.CONV_ELU_IN1
    if relu:  # the "elu" output holds ReLU(y): VGG19 feature convs of the perceptual loss
        flags |= N.CONV_RELU
    if act == 2:  # LeakyReLU(0.01) / tanh: the MLPs of the behavior net's latent flow (per-tap kernel only)
        flags |= N.CONV_LRELU
    elif act == 3:
        flags |= N.CONV_TANH
    d.flags, d.stages = flags, stages
    bn = min(cout, 256)
    bkc = 64 if c_src >= 64 else 32
    out_bytes = sum(o.numel() * o.element_size() for o in outs)
    # distinct bytes: a residual that IS one of the sources (lean activations: raw x feeds both) is read once
    res_distinct = residual is not None and residual.data_ptr() not in (x0.data_ptr(), -1 if x1 is None else x1.data_ptr())
    in_bytes = nsrc * x0.numel() * 2 + w.numel() * 2 + (residual.numel() * 2 if res_distinct else 0)
    flops = 2.0 * B * Ho * Wo * cout_valid * w.shape[1]
    tag = (f"conv_fwd<{bn},{bkc}> {nsrc}x{c_src}->{cout} k{ksize} s{stride}{' d2s' if d2s else ''}"
           f"{' elu-in' if elu_in else ''} @{H}x{W}")
    N.check(_launch(tag, N.lib().vub_conv_fwd, C.byref(d), _stream(), flops=flops,
                    bytes_=float(in_bytes + out_bytes)), "vub_conv_fwd")
    return outs


def _wn_prep_cuda(table, n_entries, total_rows) -> None:
    N.check(_launch("wn_prep", N.lib().vub_wn_prep, table.data_ptr(), n_entries, total_rows, _stream()),
            "vub_wn_prep")


def _stem_fwd_cuda(x, v, g, gamma, beta, bias, want_raw, want_elu) -> List[torch.Tensor]:
    u8 = x.dtype == torch.uint8
    if x.dim() != 4 or x.dtype not in (torch.float32, torch.uint8):
        raise ValueError("stem_fwd: x must be NCHW fp32, or a uint8 HWC image batch [B, H, W, C]")
    x = x.contiguous()
    if u8:  # uint8 HWC images as the reference holds them before ToTensor() + (x * 2 - 1)
        B, H, W, cx = x.shape
    else:
        B, cx, H, W = x.shape
    if cx != v.shape[1]:
        raise ValueError(f"stem_fwd: input has {cx} channels, the layer expects {v.shape[1]}")
    cout = v.shape[0]
    outs = []
    raw = elu = None
    if want_raw:
        raw = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device=x.device)
        outs.append(raw)
    if want_elu:
        elu = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device=x.device)
        outs.append(elu)
    nbytes = float(x.numel() * x.element_size() + sum(o.numel() * 2 for o in outs))
    N.check(
        _launch(f"stem<{cout}> {cx}->{cout}{' u8' if u8 else ''} @{H}x{W}",
                N.lib().vub_stem_fwd_u8 if u8 else N.lib().vub_stem_fwd, x.data_ptr(), v.data_ptr(), g.data_ptr(), gamma.data_ptr(),
                beta.data_ptr(), bias.data_ptr(), B, cx, H, W, cout, _p(raw), _p(elu), _stream(),
                flops=2.0 * B * H * W * cx * cout, bytes_=nbytes),
        "vub_stem_fwd",
    )
    return outs


def _latent_fwd_cuda(ml, eps, want_elu_z, want_elu_mu) -> List[torch.Tensor]:
    if ml.dtype != torch.float32 or ml.dim() != 4 or not ml.is_contiguous():
        raise ValueError("latent_fwd: ml must be contiguous fp32 NHWC [B,h,w,2C]")
    B, h, w, c2 = ml.shape
    c = c2 // 2
    if eps is not None:
        if eps.dtype != torch.float32 or tuple(eps.shape) != (B, c, h, w) or not eps.is_contiguous():
            raise ValueError("latent_fwd: eps must be contiguous fp32 NCHW [B,C,h,w]")
    dev = ml.device
    mu = torch.empty((B, c, h, w), dtype=torch.float32, device=dev)
    ls = torch.empty_like(mu)
    z = torch.empty_like(mu)
    ez = torch.empty((B, h, w, c), dtype=torch.bfloat16, device=dev) if want_elu_z else None
    em = torch.empty((B, h, w, c), dtype=torch.bfloat16, device=dev) if want_elu_mu else None
    npart = N.lib().vub_latent_num_partials(B, c, h, w)
    kl = torch.empty((npart,), dtype=torch.float32, device=dev)
    n = B * c * h * w
    nbytes = float(n * (8 + 4 + 12) + (n * 2 if want_elu_z else 0) + (n * 2 if want_elu_mu else 0))
    N.check(
        _launch("latent", N.lib().vub_latent_fwd, ml.data_ptr(), _p(eps), B, c, h, w, mu.data_ptr(), ls.data_ptr(),
                z.data_ptr(), _p(ez), _p(em), kl.data_ptr(), _stream(), bytes_=nbytes),
        "vub_latent_fwd",
    )
    outs = [mu, ls, z, kl]
    if ez is not None:
        outs.append(ez)
    if em is not None:
        outs.append(em)
    return outs


def _nchw_to_nhwc_cuda(x, want_raw, want_elu) -> List[torch.Tensor]:
    if x.dtype != torch.float32 or x.dim() != 4:
        raise ValueError("nchw_to_nhwc: x must be NCHW fp32")
    x = x.contiguous()
    B, c, H, W = x.shape
    raw = torch.empty((B, H, W, c), dtype=torch.bfloat16, device=x.device) if want_raw else None
    elu = torch.empty((B, H, W, c), dtype=torch.bfloat16, device=x.device) if want_elu else None
    nbytes = float(x.numel() * 4 + (x.numel() * 2 if want_raw else 0) + (x.numel() * 2 if want_elu else 0))
    N.check(_launch("nchw_to_nhwc", N.lib().vub_nchw_f32_to_nhwc_bf16, x.data_ptr(), B, c, H, W, _p(raw), _p(elu),
                    _stream(), bytes_=nbytes), "vub_nchw_f32_to_nhwc_bf16")
    return [t for t in (raw, elu) if t is not None]


def _nhwc_to_nchw_cuda(x) -> torch.Tensor:
    _check_act(x, "nhwc_to_nchw x")
    B, H, W, c = x.shape
    out = torch.empty((B, c, H, W), dtype=torch.float32, device=x.device)
    N.check(_launch("nhwc_to_nchw", N.lib().vub_nhwc_bf16_to_nchw_f32, x.data_ptr(), B, c, H, W, out.data_ptr(),
                    _stream(), bytes_=float(x.numel() * 6)), "vub_nhwc_bf16_to_nchw_f32")
    return out


def _conv_dgrad_cuda(dy, wd, dx0, dx1, mask0, mask1, add0, add1, add20, add21, ksize, stride, d2s,
                     wd_tiled=None, relu=False) -> None:
    _check_act(dy, "conv_dgrad dy")
    _check_act(dx0, "conv_dgrad dx0")
    B, H, W, c_src = dx0.shape
    nsrc = 1 if dx1 is None else 2
    taps = ksize * ksize
    if wd.dtype != torch.bfloat16 or wd.dim() != 2 or wd.shape[0] != nsrc * c_src or wd.shape[1] % taps:
        raise ValueError(f"conv_dgrad: wd {tuple(wd.shape)} does not match nsrc={nsrc} c_src={c_src} k={ksize}")
    cout_k = wd.shape[1] // taps
    Ho, Wo = H // stride, W // stride
    want = (B, 2 * Ho, 2 * Wo, cout_k // 4) if d2s else (B, Ho, Wo, cout_k)
    if tuple(dy.shape) != want:
        raise ValueError(f"conv_dgrad: dy {tuple(dy.shape)} != {want}")
    for t in (dx1, mask0, mask1, add0, add1, add20, add21):
        if t is not None and (t.shape != dx0.shape or t.dtype != torch.bfloat16 or not t.is_contiguous()):
            raise ValueError("conv_dgrad: dx / mask / addend tensors must be contiguous NHWC bf16 of the input shape")
    d = N.DgradDesc()
    d.dy, d.wd = dy.data_ptr(), wd.data_ptr()
    if wd_tiled is not None:
        if wd_tiled.dtype != torch.bfloat16 or wd_tiled.numel() != wd.numel() or cout_k % 64:
            raise ValueError("conv_dgrad: wd_tiled must be the tile-major copy of wd (cout_k a multiple of 64)")
        d.wd_tiled = wd_tiled.data_ptr()
    d.dx[0], d.dx[1] = dx0.data_ptr(), _p(dx1)
    d.mask_elu[0], d.mask_elu[1] = _p(mask0), _p(mask1)
    d.addend[0], d.addend[1] = _p(add0), _p(add1)
    d.addend2[0], d.addend2[1] = _p(add20), _p(add21)
    d.nsrc, d.c_src, d.batch, d.hin, d.win = nsrc, c_src, B, H, W
    d.cout_k, d.ksize, d.stride, d.flags = cout_k, ksize, stride, (N.CONV_D2S if d2s else 0) | (N.CONV_RELU if relu else 0)
    n_extra = sum(t is not None for t in (mask0, mask1, add0, add1, add20, add21))
    nbytes = float(dy.numel() * 2 + wd.numel() * 2 + dx0.numel() * 2 * (nsrc + n_extra))
    flops = 2.0 * B * Ho * Wo * cout_k * taps * nsrc * c_src
    tag = f"conv_dgrad {nsrc}x{c_src}<-{cout_k} k{ksize} s{stride}{' d2s' if d2s else ''} @{H}x{W}"
    N.check(_launch(tag, N.lib().vub_conv_dgrad, C.byref(d), _stream(), flops=flops,
                    bytes_=nbytes), "vub_conv_dgrad")


def _wgrad_desc(dy_shape, x_shape, nsrc, ksize, stride, d2s) -> "N.WgradDesc":
    B, H, W, c_src = x_shape
    d = N.WgradDesc()
    d.nsrc, d.c_src, d.batch, d.hin, d.win = nsrc, c_src, B, H, W
    d.cout_k = dy_shape[3] * 4 if d2s else dy_shape[3]
    d.ksize, d.stride, d.flags = ksize, stride, (N.CONV_D2S if d2s else 0)
    return d


def wgrad_splits(dy_shape, x_shape, nsrc: int, ksize: int, stride: int, d2s: bool) -> int:
    """Number of pixel splits ``vub_conv_wgrad`` should use for this shape (sizes the partial buffers)."""
    n = N.lib().vub_conv_wgrad_splits(C.byref(_wgrad_desc(dy_shape, x_shape, nsrc, ksize, stride, d2s)))
    if n <= 0:
        N.check(n if n else 1, "vub_conv_wgrad_splits")
    return n


def _conv_wgrad_cuda(dy, x0, x1, gt, dy_sum, ksize, stride, d2s) -> None:
    _check_act(dy, "conv_wgrad dy")
    _check_act(x0, "conv_wgrad x0")
    B, H, W, c_src = x0.shape
    nsrc = 1
    if x1 is not None:
        _check_act(x1, "conv_wgrad x1")
        if x1.shape != x0.shape:
            raise ValueError("conv_wgrad: both sources must have the same shape")
        nsrc = 2
    taps = ksize * ksize
    Ho, Wo = H // stride, W // stride
    cout_k = dy.shape[3] * 4 if d2s else dy.shape[3]
    want = (B, 2 * Ho, 2 * Wo, cout_k // 4) if d2s else (B, Ho, Wo, cout_k)
    if tuple(dy.shape) != want:
        raise ValueError(f"conv_wgrad: dy {tuple(dy.shape)} != {want}")
    if (gt.dtype != torch.float32 or gt.dim() != 3 or gt.shape[1] < cout_k or gt.shape[2] != nsrc * taps * c_src
            or gt.stride(2) != 1 or dy_sum.dtype != torch.float32 or dy_sum.dim() != 2
            or dy_sum.shape[0] != gt.shape[0] or dy_sum.shape[1] < cout_k or dy_sum.stride(1) != 1):
        raise ValueError(f"conv_wgrad: gt {tuple(gt.shape)} / dy_sum {tuple(dy_sum.shape)} do not match "
                         f"[n_split, cout_k={cout_k}, K={nsrc * taps * c_src}]")
    d = _wgrad_desc(dy.shape, x0.shape, nsrc, ksize, stride, d2s)
    d.dy, d.gt, d.dy_sum = dy.data_ptr(), gt.data_ptr(), dy_sum.data_ptr()
    d.x[0], d.x[1] = x0.data_ptr(), _p(x1)
    d.n_split, d.split_stride, d.sum_stride, d.gt_pitch = gt.shape[0], gt.stride(0), dy_sum.stride(0), gt.stride(1)
    nbytes = float(dy.numel() * 2 + nsrc * x0.numel() * 2 + gt.numel() * 4)
    flops = 2.0 * B * Ho * Wo * cout_k * taps * nsrc * c_src
    tag = f"conv_wgrad {nsrc}x{c_src}x{cout_k} k{ksize} s{stride}{' d2s' if d2s else ''} @{H}x{W}"
    N.check(_launch(tag, N.lib().vub_conv_wgrad, C.byref(d), _stream(), flops=flops,
                    bytes_=nbytes), "vub_conv_wgrad")


def _wn_bwd_cuda(table, n_entries, total_rows, accumulate, nbytes=0.0, max_len=0) -> None:
    # nbytes: algorithmic bytes of this table (split partials read once, v read, dv written; training.GradBank)
    N.check(_launch("wn_bwd", N.lib().vub_wn_bwd, table.data_ptr(), n_entries, total_rows, int(accumulate), int(max_len),
                    _stream(),
                    bytes_=float(nbytes)), "vub_wn_bwd")


def _adam_wn_step_cuda(table, n_entries, total_rows, beta1, beta2, eps, step, nbytes=0.0) -> None:
    N.check(_launch("adam_wn_step", N.lib().vub_adam_wn_step, table.data_ptr(), n_entries, total_rows, beta1, beta2, eps,
                    step, _stream(), bytes_=float(nbytes)), "vub_adam_wn_step")


def _adam_wn_step_dev_cuda(table, n_entries, total_rows, beta1, beta2, eps, step, nbytes=0.0) -> None:
    if step.dtype != torch.int32 or step.numel() != 1:
        raise ValueError("adam_wn_step_dev: step must be a 1-element int32 tensor")
    global LAUNCHES
    LAUNCHES += 1  # the counter increment is a second (1-thread) launch
    N.check(_launch("adam_wn_step", N.lib().vub_adam_wn_step_dev, table.data_ptr(), n_entries, total_rows, beta1, beta2,
                    eps, step.data_ptr(), _stream(), bytes_=float(nbytes)), "vub_adam_wn_step_dev")


def _stem_bwd_cuda(x, dy, gt, dy_sum) -> None:
    if x.dtype != torch.float32 or x.dim() != 4 or not x.is_contiguous():
        raise ValueError("stem_bwd: x must be contiguous NCHW fp32")
    _check_act(dy, "stem_bwd dy")
    B, cx, H, W = x.shape
    cout = dy.shape[3]
    if tuple(dy.shape) != (B, H, W, cout) or gt.numel() != cx * cout or dy_sum.numel() != cout:
        raise ValueError("stem_bwd: shape mismatch")
    N.check(_launch(f"stem_bwd<{cout}>", N.lib().vub_stem_bwd, x.data_ptr(), dy.data_ptr(), B, cx, H, W, cout,
                    gt.data_ptr(), dy_sum.data_ptr(), _stream(), flops=2.0 * B * H * W * cx * cout,
                    bytes_=float(x.numel() * 4 + dy.numel() * 2)), "vub_stem_bwd")


def _latent_bwd_cuda(logstd, eps, d_mu, d_ls, dz) -> torch.Tensor:
    if logstd.dtype != torch.float32 or logstd.dim() != 4 or not logstd.is_contiguous():
        raise ValueError("latent_bwd: logstd must be contiguous NCHW fp32")
    B, c, h, w = logstd.shape
    for t in (eps, d_mu, d_ls):
        if t is not None and (t.dtype != torch.float32 or t.shape != logstd.shape or not t.is_contiguous()):
            raise ValueError("latent_bwd: eps / d_mu / d_ls must be contiguous NCHW fp32 like logstd")
    if dz is not None:
        _check_act(dz, "latent_bwd dz")
        if tuple(dz.shape) != (B, h, w, c):
            raise ValueError("latent_bwd: dz shape mismatch")
    out = torch.empty((B, h, w, 2 * c), dtype=torch.bfloat16, device=logstd.device)
    N.check(_launch("latent_bwd", N.lib().vub_latent_bwd, None, logstd.data_ptr(), _p(eps), _p(d_mu), _p(d_ls), _p(dz),
                    B, c, h, w, out.data_ptr(), _stream(), bytes_=float(logstd.numel() * 22)), "vub_latent_bwd")
    return out


def _nchw_to_nhwc_pad_cuda(x, c_pad) -> torch.Tensor:
    if x.dtype != torch.float32 or x.dim() != 4:
        raise ValueError("nchw_to_nhwc_pad: x must be NCHW fp32")
    x = x.contiguous()
    B, c, H, W = x.shape
    out = torch.empty((B, H, W, c_pad), dtype=torch.bfloat16, device=x.device)
    N.check(_launch("nchw_to_nhwc_pad", N.lib().vub_nchw_f32_to_nhwc_bf16_pad, x.data_ptr(), B, c, H, W, c_pad,
                    out.data_ptr(), _stream(), bytes_=float(x.numel() * 4 + out.numel() * 2)),
            "vub_nchw_f32_to_nhwc_bf16_pad")
    return out


def _conv_fwd32_cuda(x0, x1, w, scale, shift, residual, cout, ksize, stride, want_raw, want_elu, d2s, nchw, relu=False):
    for t in (x0, x1, residual):
        if t is not None and (t.dtype != torch.float32 or t.dim() != 4 or not t.is_contiguous()):
            raise ValueError("conv_fwd32: activations must be contiguous NHWC fp32")
    B, H, W, c_src = x0.shape
    nsrc = 1 if x1 is None else 2
    if w.dtype != torch.float32 or w.dim() != 2 or w.shape[0] != nsrc * ksize * ksize * c_src or w.shape[1] < cout:
        raise ValueError(f"conv_fwd32: weight {tuple(w.shape)} does not match K={nsrc * ksize * ksize * c_src}, cout={cout}")
    Ho, Wo = H // stride, W // stride
    oshape = (B, cout, Ho, Wo) if nchw else ((B, 2 * Ho, 2 * Wo, cout // 4) if d2s else (B, Ho, Wo, cout))
    d = N.ConvF32Desc()
    d.x[0], d.x[1] = x0.data_ptr(), _p(x1)
    d.w, d.scale, d.shift, d.residual = w.data_ptr(), scale.data_ptr(), shift.data_ptr(), _p(residual)
    outs = []
    if want_raw:
        raw = torch.empty(oshape, dtype=torch.float32, device=x0.device)
        d.out_raw = raw.data_ptr()
        outs.append(raw)
    if want_elu:
        elu = torch.empty(oshape, dtype=torch.float32, device=x0.device)
        d.out_elu = elu.data_ptr()
        outs.append(elu)
    d.nsrc, d.c_src, d.batch, d.hin, d.win, d.cout = nsrc, c_src, B, H, W, cout
    d.w_pitch, d.ksize, d.stride = w.stride(0), ksize, stride
    d.flags = (N.CONV_D2S if d2s else 0) | (N.CONV_F32_NCHW if nchw else 0) | (N.CONV_RELU if relu else 0)
    N.check(_launch(f"conv_fwd32 {nsrc}x{c_src}->{cout} k{ksize} s{stride} @{H}x{W}", N.lib().vub_conv_fwd_f32, C.byref(d),
                    _stream(), flops=2.0 * B * Ho * Wo * cout * w.shape[0]), "vub_conv_fwd_f32")
    return outs


def _nchw_to_nhwc32_cuda(x, want_raw, want_elu):
    if x.dtype != torch.float32 or x.dim() != 4:
        raise ValueError("nchw_to_nhwc32: x must be NCHW fp32")
    x = x.contiguous()
    B, c, H, W = x.shape
    raw = torch.empty((B, H, W, c), dtype=torch.float32, device=x.device) if want_raw else None
    elu = torch.empty((B, H, W, c), dtype=torch.float32, device=x.device) if want_elu else None
    N.check(_launch("nchw_to_nhwc32", N.lib().vub_nchw_f32_to_nhwc_f32, x.data_ptr(), B, c, H, W, _p(raw), _p(elu),
                    _stream()), "vub_nchw_f32_to_nhwc_f32")
    return [t for t in (raw, elu) if t is not None]


def _check_act32(t: torch.Tensor, name: str) -> None:
    if t.dtype != torch.float32 or t.dim() != 4 or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous NHWC fp32 [B,H,W,C], got {t.dtype} {tuple(t.shape)}")


def _conv_dgrad32_cuda(dy, w, scale, dx0, dx1, mask0, mask1, add0, add1, add20, add21, cout, ksize, stride, d2s,
                       relu=False) -> None:
    _check_act32(dy, "conv_dgrad32 dy")
    _check_act32(dx0, "conv_dgrad32 dx0")
    B, H, W, c_src = dx0.shape
    nsrc = 1 if dx1 is None else 2
    Ho, Wo = H // stride, W // stride
    want = (B, 2 * Ho, 2 * Wo, cout // 4) if d2s else (B, Ho, Wo, cout)
    if tuple(dy.shape) != want:
        raise ValueError(f"conv_dgrad32: dy {tuple(dy.shape)} != {want}")
    if w.dtype != torch.float32 or w.dim() != 2 or w.shape[0] != nsrc * ksize * ksize * c_src or w.shape[1] < cout:
        raise ValueError(f"conv_dgrad32: weight {tuple(w.shape)} does not match K / cout")
    for t in (dx1, mask0, mask1, add0, add1, add20, add21):
        if t is not None and (t.shape != dx0.shape or t.dtype != torch.float32 or not t.is_contiguous()):
            raise ValueError("conv_dgrad32: dx / mask / addend tensors must be contiguous NHWC fp32 of the input shape")
    d = N.DgradF32Desc()
    d.dy, d.w, d.scale = dy.data_ptr(), w.data_ptr(), scale.data_ptr()
    d.dx[0], d.dx[1] = dx0.data_ptr(), _p(dx1)
    d.mask_elu[0], d.mask_elu[1] = _p(mask0), _p(mask1)
    d.addend[0], d.addend[1] = _p(add0), _p(add1)
    d.addend2[0], d.addend2[1] = _p(add20), _p(add21)
    d.nsrc, d.c_src, d.batch, d.hin, d.win, d.cout = nsrc, c_src, B, H, W, cout
    d.w_pitch, d.ksize, d.stride, d.flags = w.stride(0), ksize, stride, (N.CONV_D2S if d2s else 0)
    if relu:
        d.flags |= N.CONV_RELU
    N.check(_launch(f"conv_dgrad32 {nsrc}x{c_src}<-{cout} k{ksize} s{stride} @{H}x{W}", N.lib().vub_conv_dgrad_f32,
                    C.byref(d), _stream(), flops=2.0 * B * Ho * Wo * cout * w.shape[0]), "vub_conv_dgrad_f32")


def _conv_wgrad32_cuda(dy, x0, x1, gt, dy_sum, cout, ksize, stride, d2s) -> None:
    _check_act32(dy, "conv_wgrad32 dy")
    _check_act32(x0, "conv_wgrad32 x0")
    B, H, W, c_src = x0.shape
    nsrc = 1
    if x1 is not None:
        _check_act32(x1, "conv_wgrad32 x1")
        nsrc = 2
    Ho, Wo = H // stride, W // stride
    want = (B, 2 * Ho, 2 * Wo, cout // 4) if d2s else (B, Ho, Wo, cout)
    if tuple(dy.shape) != want:
        raise ValueError(f"conv_wgrad32: dy {tuple(dy.shape)} != {want}")
    K = nsrc * ksize * ksize * c_src
    if (gt.dtype != torch.float32 or gt.dim() != 3 or gt.shape[1] < cout or gt.shape[2] != K or gt.stride(2) != 1
            or dy_sum.dtype != torch.float32 or dy_sum.dim() != 2 or dy_sum.shape[0] != gt.shape[0]
            or dy_sum.shape[1] < cout or dy_sum.stride(1) != 1):
        raise ValueError(f"conv_wgrad32: gt {tuple(gt.shape)} / dy_sum {tuple(dy_sum.shape)} do not match "
                         f"[n_split, cout={cout}, K={K}]")
    d = N.WgradF32Desc()
    d.dy, d.gt, d.dy_sum = dy.data_ptr(), gt.data_ptr(), dy_sum.data_ptr()
    d.x[0], d.x[1] = x0.data_ptr(), _p(x1)
    d.nsrc, d.c_src, d.batch, d.hin, d.win, d.cout = nsrc, c_src, B, H, W, cout
    d.ksize, d.stride, d.flags = ksize, stride, (N.CONV_D2S if d2s else 0)
    d.n_split, d.gt_pitch, d.split_stride, d.sum_stride = gt.shape[0], gt.stride(1), gt.stride(0), dy_sum.stride(0)
    N.check(_launch(f"conv_wgrad32 {nsrc}x{c_src}x{cout} k{ksize} s{stride} @{H}x{W}", N.lib().vub_conv_wgrad_f32,
                    C.byref(d), _stream(), flops=2.0 * B * Ho * Wo * cout * K), "vub_conv_wgrad_f32")


def _latent_bwd32_cuda(logstd, eps, d_mu, d_ls, dz) -> torch.Tensor:
    if logstd.dtype != torch.float32 or logstd.dim() != 4 or not logstd.is_contiguous():
        raise ValueError("latent_bwd32: logstd must be contiguous NCHW fp32")
    B, c, h, w = logstd.shape
    for t in (eps, d_mu, d_ls):
        if t is not None and (t.dtype != torch.float32 or t.shape != logstd.shape or not t.is_contiguous()):
            raise ValueError("latent_bwd32: eps / d_mu / d_ls must be contiguous NCHW fp32 like logstd")
    if dz is not None:
        _check_act32(dz, "latent_bwd32 dz")
        if tuple(dz.shape) != (B, h, w, c):
            raise ValueError("latent_bwd32: dz shape mismatch")
    out = torch.empty((B, h, w, 2 * c), dtype=torch.float32, device=logstd.device)
    N.check(_launch("latent_bwd32", N.lib().vub_latent_bwd_f32, logstd.data_ptr(), _p(eps), _p(d_mu), _p(d_ls), _p(dz),
                    B, c, h, w, out.data_ptr(), _stream()), "vub_latent_bwd_f32")
    return out


def _frames_to_uint8_cuda(imgs):
    if imgs.dtype != torch.float32 or imgs.dim() != 4 or imgs.shape[1] != 3:
        raise ValueError("frames_to_uint8: imgs must be NCHW fp32 with 3 channels")
    imgs = imgs.contiguous()
    B, _, H, W = imgs.shape
    out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=imgs.device)
    N.check(_launch("frames_to_uint8", N.lib().vub_frames_to_uint8, imgs.data_ptr(), B, H, W, out.data_ptr(), _stream(),
                    bytes_=float(imgs.numel() * 5)), "vub_frames_to_uint8")
    return out


def _latent_sample_cuda(p, eps, sub, gi, gj, mean, z, z_raw, z_elu, want_grp_elu) -> List[torch.Tensor]:
    """``vub_latent_sample``: z = p + eps for one parameter map (the whole latent, sub = 1, or prior group (gi, gj),
    sub = 2), written straight to its positions in the merged tensors the caller passes (see include/vunet_b200.h)."""
    if p.dtype != torch.float32 or p.dim() != 4 or not p.is_contiguous():
        raise ValueError("latent_sample: p must be contiguous fp32 NHWC [B,h,w,C]")
    B, h, w, c = p.shape
    if eps is not None and (eps.dtype != torch.float32 or tuple(eps.shape) != (B, c, h, w) or not eps.is_contiguous()):
        raise ValueError("latent_sample: eps must be contiguous fp32 NCHW [B,C,h,w]")
    H, W = h * sub, w * sub
    for t, name in ((mean, "mean"), (z, "z")):
        if t is not None and (t.dtype != torch.float32 or tuple(t.shape) != (B, c, H, W) or not t.is_contiguous()):
            raise ValueError(f"latent_sample: {name} must be contiguous fp32 NCHW [B,C,{H},{W}]")
    act_dtype = None
    for t, name in ((z_raw, "z_raw"), (z_elu, "z_elu")):
        if t is not None:
            if t.dtype not in (torch.bfloat16, torch.float32) or tuple(t.shape) != (B, H, W, c) or not t.is_contiguous():
                raise ValueError(f"latent_sample: {name} must be contiguous NHWC [B,{H},{W},C] bf16 / fp32")
            if act_dtype is not None and t.dtype != act_dtype:
                raise ValueError("latent_sample: activation outputs must share one dtype")
            act_dtype = t.dtype
    if act_dtype is None:
        act_dtype = torch.bfloat16
    grp = torch.empty((B, h, w, c), dtype=act_dtype, device=p.device) if want_grp_elu else None
    n = p.numel()
    asz = 4 if act_dtype == torch.float32 else 2
    nbytes = float(n * 4 * (1 + (eps is not None) + (mean is not None) + (z is not None))
                   + n * asz * ((z_raw is not None) + (z_elu is not None) + want_grp_elu))
    N.check(_launch("latent_sample", N.lib().vub_latent_sample, p.data_ptr(), _p(eps), B, c, h, w, sub, gi, gj, _p(mean),
                    _p(z), _p(z_raw), _p(z_elu), _p(grp), 1 if act_dtype == torch.float32 else 0, _stream(),
                    bytes_=nbytes), "vub_latent_sample")
    return [grp] if grp is not None else []


def _latent_sample_bwd_cuda(d_mean, dz, sub, gi, gj, h, w, f32) -> torch.Tensor:
    """``vub_latent_sample_bwd``: gradient of the parameter map [B,h,w,C] from the merged d_mean (NCHW fp32) / dz (NHWC)."""
    if d_mean is None and dz is None:
        raise ValueError("latent_sample_bwd: no gradient given")
    act = torch.float32 if f32 else torch.bfloat16
    if d_mean is not None:
        if d_mean.dtype != torch.float32 or d_mean.dim() != 4 or d_mean.shape[2] != h * sub or d_mean.shape[3] != w * sub:
            raise ValueError("latent_sample_bwd: d_mean must be fp32 NCHW [B,C,h*sub,w*sub]")
        d_mean = d_mean.contiguous()
        B, c = d_mean.shape[0], d_mean.shape[1]
    if dz is not None:
        if dz.dtype != act or dz.dim() != 4 or dz.shape[1] != h * sub or dz.shape[2] != w * sub or not dz.is_contiguous():
            raise ValueError("latent_sample_bwd: dz must be a contiguous NHWC activation gradient [B,h*sub,w*sub,C]")
        B, c = dz.shape[0], dz.shape[3]
    ref = d_mean if d_mean is not None else dz
    out = torch.empty((B, h, w, c), dtype=act, device=ref.device)
    N.check(_launch("latent_sample_bwd", N.lib().vub_latent_sample_bwd, _p(d_mean), _p(dz), B, c, h, w, sub, gi, gj,
                    out.data_ptr(), int(f32), _stream(), bytes_=float(out.numel() * (8 + 2 * out.element_size()))),
            "vub_latent_sample_bwd")
    return out


def _space_to_depth_cuda(x, group_major, inverse, want_copy, want_elu, accumulate_into=None) -> List[torch.Tensor]:
    """``vub_space_to_depth`` on an NHWC activation (bf16, or fp32 in validation mode).  ``accumulate_into``: add the
    regrouped values to this existing tensor instead of returning a copy (gradient accumulation)."""
    if x.dtype not in (torch.bfloat16, torch.float32) or not x.is_contiguous():
        raise ValueError("space_to_depth: x must be a contiguous bf16 / fp32 activation")
    if inverse:  # x is the grouped tensor
        if group_major:
            if x.dim() != 5 or x.shape[0] != 4:
                raise ValueError("space_to_depth: inverse group-major input must be [4,B,h/2,w/2,C]")
            _, B, h2, w2, c = x.shape
        else:
            if x.dim() != 4 or x.shape[3] % 4:
                raise ValueError("space_to_depth: inverse input must be [B,h/2,w/2,4C]")
            B, h2, w2, c4 = x.shape
            c = c4 // 4
        h, w = 2 * h2, 2 * w2
        oshape = (B, h, w, c)
    else:
        if x.dim() != 4 or x.shape[1] % 2 or x.shape[2] % 2:
            raise ValueError("space_to_depth: x must be NHWC [B,h,w,C] with even h, w")
        B, h, w, c = x.shape
        oshape = (4, B, h // 2, w // 2, c) if group_major else (B, h // 2, w // 2, 4 * c)
    if accumulate_into is not None:
        if want_elu or accumulate_into.dtype != x.dtype or tuple(accumulate_into.shape) != oshape \
                or not accumulate_into.is_contiguous():
            raise ValueError("space_to_depth: accumulate_into must match the output (and excludes want_elu)")
        copy = accumulate_into
    else:
        copy = torch.empty(oshape, dtype=x.dtype, device=x.device) if want_copy else None
    elu = torch.empty(oshape, dtype=x.dtype, device=x.device) if want_elu else None
    nbytes = float(x.numel() * x.element_size() * (1 + (copy is not None) + want_elu + (accumulate_into is not None)))
    N.check(_launch("space_to_depth", N.lib().vub_space_to_depth, x.data_ptr(), B, h, w, c, int(group_major), int(inverse),
                    int(accumulate_into is not None), _p(copy), _p(elu), 1 if x.dtype == torch.float32 else 0, _stream(),
                    bytes_=nbytes), "vub_space_to_depth")
    if accumulate_into is not None:
        return []
    return [t for t in (copy, elu) if t is not None]



# ---- VGG19 perceptual loss ------------------------------------------------------------------------
def _check_img(t: torch.Tensor, name: str) -> None:
    if t.dtype != torch.float32 or t.dim() != 4 or t.shape[1] != 3:
        raise ValueError(f"{name}: expected NCHW fp32 with 3 channels, got {t.dtype} {tuple(t.shape)}")


def _feat_f32(t: torch.Tensor, name: str) -> int:
    """Feature maps are contiguous NHWC bf16, or fp32 in the validation mode: returns the C ABI's f32_acts flag."""
    if t.dtype not in (torch.bfloat16, torch.float32) or t.dim() != 4 or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous NHWC bf16 / fp32 [B,H,W,C], got {t.dtype} {tuple(t.shape)}")
    return 1 if t.dtype == torch.float32 else 0


def _vgg_normalize_cuda(x, want_nhwc, want_nchw, f32=False) -> List[torch.Tensor]:
    _check_img(x, "vgg_normalize x")
    x = x.contiguous()
    B, _, H, W = x.shape
    nhwc = torch.empty((B, H, W, 32), dtype=torch.float32 if f32 else torch.bfloat16, device=x.device) if want_nhwc else None
    nchw = torch.empty_like(x) if want_nchw else None
    N.check(_launch("vgg_normalize", N.lib().vub_vgg_normalize, x.data_ptr(), B, H, W, _p(nhwc), _p(nchw), int(f32), _stream(),
                    bytes_=float(x.numel() * 4 + (B * H * W * 64 if want_nhwc else 0) + (x.numel() * 4 if want_nchw else 0))),
            "vub_vgg_normalize")
    return [t for t in (nhwc, nchw) if t is not None]


def _vgg_input_l1_cuda(pred, target) -> torch.Tensor:
    _check_img(pred, "vgg_input_l1 pred")
    _check_img(target, "vgg_input_l1 target")
    if pred.shape != target.shape:
        raise ValueError("vgg_input_l1: pred / target shapes differ")
    pred, target = pred.contiguous(), target.contiguous()
    B, _, H, W = pred.shape
    part = torch.empty((N.lib().vub_vgg_input_l1_partials(B, H, W),), dtype=torch.float32, device=pred.device)
    N.check(_launch("vgg_input_l1", N.lib().vub_vgg_input_l1, pred.data_ptr(), target.data_ptr(), B, H, W,
                    part.data_ptr(), _stream(), bytes_=float(pred.numel() * 8)), "vub_vgg_input_l1")
    return part


def _vgg_input_bwd_cuda(pred, target, d_xn, g0, scale0) -> torch.Tensor:
    _check_img(pred, "vgg_input_bwd pred")
    pred, target = pred.contiguous(), target.contiguous()
    B, _, H, W = pred.shape
    f32 = 0
    if d_xn is not None:
        f32 = _feat_f32(d_xn, "vgg_input_bwd d_xn")
        if tuple(d_xn.shape) != (B, H, W, 32):
            raise ValueError("vgg_input_bwd: d_xn must be [B,H,W,32]")
    if g0 is not None and (g0.dtype != torch.float32 or g0.numel() != 1):
        raise ValueError("vgg_input_bwd: g0 must be one fp32 value")
    out = torch.empty_like(pred)
    N.check(_launch("vgg_input_bwd", N.lib().vub_vgg_input_bwd, pred.data_ptr(), target.data_ptr(), _p(d_xn), _p(g0),
                    float(scale0), B, H, W, out.data_ptr(), f32, _stream(),
                    bytes_=float(pred.numel() * 12 + (0 if d_xn is None else d_xn.numel() * 2))), "vub_vgg_input_bwd")
    return out


def _maxpool2_fwd_cuda(x) -> torch.Tensor:
    f32 = _feat_f32(x, "maxpool2_fwd x")
    B, H, W, c = x.shape
    y = torch.empty((B, H // 2, W // 2, c), dtype=x.dtype, device=x.device)
    N.check(_launch(f"maxpool2 c{c} @{H}x{W}", N.lib().vub_maxpool2_fwd, x.data_ptr(), B, H, W, c, y.data_ptr(), f32,
                    _stream(), bytes_=float((x.numel() + y.numel()) * x.element_size())), "vub_maxpool2_fwd")
    return y


def _maxpool2_relu_bwd_cuda(x, dy, addend) -> torch.Tensor:
    f32 = _feat_f32(x, "maxpool2_relu_bwd x")
    B, H, W, c = x.shape
    if _feat_f32(dy, "maxpool2_relu_bwd dy") != f32 or tuple(dy.shape) != (B, H // 2, W // 2, c):
        raise ValueError("maxpool2_relu_bwd: dy shape / dtype mismatch")
    if addend is not None and (_feat_f32(addend, "maxpool2_relu_bwd addend") != f32 or addend.shape != x.shape):
        raise ValueError("maxpool2_relu_bwd: addend shape / dtype mismatch")
    dx = torch.empty_like(x)
    es = x.element_size()
    N.check(_launch(f"maxpool2_relu_bwd c{c} @{H}x{W}", N.lib().vub_maxpool2_relu_bwd, x.data_ptr(), dy.data_ptr(),
                    _p(addend), B, H, W, c, dx.data_ptr(), f32, _stream(),
                    bytes_=float(x.numel() * es * (2 + (addend is not None)) + dy.numel() * es)), "vub_maxpool2_relu_bwd")
    return dx


def _l1_partial_cuda(pf, tf) -> torch.Tensor:
    f32 = _feat_f32(pf, "l1_partial pf")
    if _feat_f32(tf, "l1_partial tf") != f32 or pf.shape != tf.shape:
        raise ValueError("l1_partial: shapes / dtypes differ")
    part = torch.empty((N.lib().vub_l1_num_partials(pf.numel()),), dtype=torch.float32, device=pf.device)
    N.check(_launch("l1_partial", N.lib().vub_l1_partial, pf.data_ptr(), tf.data_ptr(), pf.numel(), part.data_ptr(), f32,
                    _stream(), bytes_=float(pf.numel() * 2 * pf.element_size())), "vub_l1_partial")
    return part


def _l1_relu_grad_cuda(pf, tf, g, scale) -> torch.Tensor:
    f32 = _feat_f32(pf, "l1_relu_grad pf")
    if _feat_f32(tf, "l1_relu_grad tf") != f32 or pf.shape != tf.shape or g.dtype != torch.float32 or g.numel() != 1:
        raise ValueError("l1_relu_grad: shapes / dtypes differ, or g is not one fp32 value")
    out = torch.empty_like(pf)
    N.check(_launch("l1_relu_grad", N.lib().vub_l1_relu_grad, pf.data_ptr(), tf.data_ptr(), pf.numel(), g.data_ptr(),
                    float(scale), out.data_ptr(), f32, _stream(), bytes_=float(pf.numel() * 3 * pf.element_size())),
            "vub_l1_relu_grad")
    return out


_LIB.impl("vgg_normalize", _vgg_normalize_cuda, "CUDA")
_LIB.impl("vgg_input_l1", _vgg_input_l1_cuda, "CUDA")
_LIB.impl("vgg_input_bwd", _vgg_input_bwd_cuda, "CUDA")
_LIB.impl("maxpool2_fwd", _maxpool2_fwd_cuda, "CUDA")
_LIB.impl("maxpool2_relu_bwd", _maxpool2_relu_bwd_cuda, "CUDA")
_LIB.impl("l1_partial", _l1_partial_cuda, "CUDA")
_LIB.impl("l1_relu_grad", _l1_relu_grad_cuda, "CUDA")
_LIB.impl("latent_sample", _latent_sample_cuda, "CUDA")
_LIB.impl("space_to_depth", _space_to_depth_cuda, "CUDA")
_LIB.impl("latent_sample_bwd", _latent_sample_bwd_cuda, "CUDA")
_LIB.impl("conv_fwd", _conv_fwd_cuda, "CUDA")
_LIB.impl("frames_to_uint8", _frames_to_uint8_cuda, "CUDA")
_LIB.impl("conv_fwd32", _conv_fwd32_cuda, "CUDA")
_LIB.impl("conv_dgrad32", _conv_dgrad32_cuda, "CUDA")
_LIB.impl("conv_wgrad32", _conv_wgrad32_cuda, "CUDA")
_LIB.impl("latent_bwd32", _latent_bwd32_cuda, "CUDA")
_LIB.impl("nchw_to_nhwc32", _nchw_to_nhwc32_cuda, "CUDA")
_LIB.impl("conv_dgrad", _conv_dgrad_cuda, "CUDA")
_LIB.impl("conv_wgrad", _conv_wgrad_cuda, "CUDA")
_LIB.impl("wn_bwd", _wn_bwd_cuda, "CUDA")
_LIB.impl("adam_wn_step", _adam_wn_step_cuda, "CUDA")
_LIB.impl("adam_wn_step_dev", _adam_wn_step_dev_cuda, "CUDA")
_LIB.impl("stem_bwd", _stem_bwd_cuda, "CUDA")
_LIB.impl("latent_bwd", _latent_bwd_cuda, "CUDA")
_LIB.impl("nchw_to_nhwc_pad", _nchw_to_nhwc_pad_cuda, "CUDA")
_LIB.impl("wn_prep", _wn_prep_cuda, "CUDA")
_LIB.impl("stem_fwd", _stem_fwd_cuda, "CUDA")
_LIB.impl("latent_fwd", _latent_fwd_cuda, "CUDA")
_LIB.impl("nchw_to_nhwc", _nchw_to_nhwc_cuda, "CUDA")
_LIB.impl("nhwc_to_nchw", _nhwc_to_nchw_cuda, "CUDA")

class _Direct:
    """The module path calls the CUDA implementations of the registered ops DIRECTLY when its first tensor lives on
    a CUDA device: the torch dispatcher costs ~10 us per call on a Python-registered op, which matters for the
    launch-by-launch (eager) training step — ~270 launches, host-bound (profiles/r01_train_host_bound.log).  The ops
    stay registered as ``torch.ops.vunet_b200.*`` (CUDA key only) for every other caller; CPU tensors are still
    refused here, loudly.  ``VUB_DISPATCH=torch`` routes everything through the dispatcher again."""

    def __init__(self):
        import os
        self._via_torch = os.environ.get("VUB_DISPATCH", "direct") == "torch"
        self._impl = {}

    def register(self, name, fn):
        self._impl[name] = fn

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if self._via_torch:
            return getattr(torch.ops.vunet_b200, name)
        impl = self._impl[name]

        def call(*args):
            for a in args:
                if isinstance(a, torch.Tensor):
                    if not a.is_cuda:
                        raise NotImplementedError(
                            f"vunet_b200::{name}: no CPU implementation (there is no CPU path); got a {a.device} tensor")
                    break
            return impl(*args)

        self.__dict__[name] = call
        return call


_ops = _Direct()
for _name, _fn in (("conv_fwd", _conv_fwd_cuda), ("frames_to_uint8", _frames_to_uint8_cuda),
                   ("conv_fwd32", _conv_fwd32_cuda), ("conv_dgrad32", _conv_dgrad32_cuda),
                   ("conv_wgrad32", _conv_wgrad32_cuda), ("latent_bwd32", _latent_bwd32_cuda),
                   ("nchw_to_nhwc32", _nchw_to_nhwc32_cuda), ("conv_dgrad", _conv_dgrad_cuda),
                   ("conv_wgrad", _conv_wgrad_cuda), ("wn_bwd", _wn_bwd_cuda), ("adam_wn_step", _adam_wn_step_cuda),
                   ("adam_wn_step_dev", _adam_wn_step_dev_cuda), ("stem_bwd", _stem_bwd_cuda),
                   ("latent_bwd", _latent_bwd_cuda), ("nchw_to_nhwc_pad", _nchw_to_nhwc_pad_cuda),
                   ("wn_prep", _wn_prep_cuda), ("stem_fwd", _stem_fwd_cuda), ("latent_fwd", _latent_fwd_cuda),
                   ("nchw_to_nhwc", _nchw_to_nhwc_cuda), ("nhwc_to_nchw", _nhwc_to_nchw_cuda),
                   ("latent_sample", _latent_sample_cuda), ("space_to_depth", _space_to_depth_cuda),
                   ("latent_sample_bwd", _latent_sample_bwd_cuda), ("vgg_normalize", _vgg_normalize_cuda),
                   ("vgg_input_l1", _vgg_input_l1_cuda), ("vgg_input_bwd", _vgg_input_bwd_cuda),
                   ("maxpool2_fwd", _maxpool2_fwd_cuda), ("maxpool2_relu_bwd", _maxpool2_relu_bwd_cuda),
                   ("l1_partial", _l1_partial_cuda), ("l1_relu_grad", _l1_relu_grad_cuda)):
    _ops.register(_name, _fn)


# ------------------------------------------------------------------------------------------------
# Convenience wrappers used by the modules
# ------------------------------------------------------------------------------------------------
def conv_fwd(x: Sequence[torch.Tensor], w, scale, shift, *, ksize: int, stride: int = 1,
             residual: Optional[torch.Tensor] = None, want_raw: bool = True, want_elu: bool = True,
             d2s: bool = False, cout_valid: Optional[int] = None, stages: int = 0, elu_in: int = 0,
             w_tiled: Optional[torch.Tensor] = None, relu: bool = False,
             act: int = 0) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """bf16 NHWC outputs: returns (raw or None, elu or None).  ``elu_in``: bit s set = source s is a RAW
    activation to which the kernel applies ELU while loading (only where :func:`elu_in_ok`)."""
    cv = w.shape[0] if cout_valid is None else cout_valid
    outs = _ops.conv_fwd(x[0], x[1] if len(x) > 1 else None, w, scale, shift, residual, ksize, stride, cv,
                         want_raw, want_elu, d2s, F32_NONE, stages, elu_in, w_tiled, relu, act)
    it = iter(outs)
    raw = next(it) if want_raw else None
    elu = next(it) if want_elu else None
    return raw, elu


@functools.lru_cache(maxsize=None)
def elu_in_ok(nsrc: int, c_src: int, cout: int, ksize: int, stride: int, h: int, w: int) -> bool:
    """Host query (vub_conv_fwd_elu_in_ok): can a conv of this geometry apply ELU to its sources on load?"""
    return bool(N.lib().vub_conv_fwd_elu_in_ok(nsrc, c_src, cout, ksize, stride, h, w))


@functools.lru_cache(maxsize=None)
def u8_out_ok(nsrc: int, c_src: int, cout: int, ksize: int, stride: int, h: int, w: int) -> bool:
    """Host query (vub_conv_fwd_u8_out_ok): can a conv of this geometry store uint8 HWC video frames itself?"""
    return bool(N.lib().vub_conv_fwd_u8_out_ok(nsrc, c_src, cout, ksize, stride, h, w))


def conv_fwd_f32(x: Sequence[torch.Tensor], w, scale, shift, *, ksize: int, layout: str,
                 cout_valid: Optional[int] = None, stages: int = 0, w_tiled: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 output, ``layout`` in {"nhwc", "nchw"} (latent head input / out_conv image); ``"u8_hwc"``: uint8 HWC
    video frames ``(scale_img(y) * 255).astype(uint8)`` (only where :func:`u8_out_ok`)."""
    cv = w.shape[0] if cout_valid is None else cout_valid
    mode = F32_NHWC if layout == "nhwc" else (U8_HWC if layout == "u8_hwc" else F32_NCHW)
    return _ops.conv_fwd(x[0], x[1] if len(x) > 1 else None, w, scale, shift, None, ksize, 1, cv, False, False,
                         False, mode, stages, 0, w_tiled)[0]


def wn_prep(table: torch.Tensor, n_entries: int, total_rows: int) -> None:
    _ops.wn_prep(table, n_entries, total_rows)


def stem_fwd(x, v, g, gamma, beta, bias, want_raw=True, want_elu=True):
    outs = _ops.stem_fwd(x, v, g, gamma, beta, bias, want_raw, want_elu)
    it = iter(outs)
    return (next(it) if want_raw else None), (next(it) if want_elu else None)


def latent_fwd(ml, eps, want_elu_z=True, want_elu_mu=False):
    """Returns dict(mu, logstd, z, kl_partial[, elu_z][, elu_mu])."""
    outs = _ops.latent_fwd(ml, eps, want_elu_z, want_elu_mu)
    res = {"mu": outs[0], "logstd": outs[1], "z": outs[2], "kl_partial": outs[3]}
    i = 4
    if want_elu_z:
        res["elu_z"] = outs[i]
        i += 1
    if want_elu_mu:
        res["elu_mu"] = outs[i]
    return res


def nchw_to_nhwc(x, want_raw=True, want_elu=True):
    outs = _ops.nchw_to_nhwc(x, want_raw, want_elu)
    it = iter(outs)
    return (next(it) if want_raw else None), (next(it) if want_elu else None)


def nhwc_to_nchw(x):
    return _ops.nhwc_to_nchw(x)


def frames_to_uint8(imgs: torch.Tensor) -> torch.Tensor:
    """``(scale_img(imgs) * 255).permute(0, 2, 3, 1)`` as uint8 (the reference's video-frame conversion)."""
    return _ops.frames_to_uint8(imgs)


# ---- fp32 validation path -------------------------------------------------------------------------
def conv_fwd32(x: Sequence[torch.Tensor], w, scale, shift, *, cout: int, ksize: int, stride: int = 1,
               residual: Optional[torch.Tensor] = None, want_raw: bool = True, want_elu: bool = True,
               d2s: bool = False, nchw: bool = False, relu: bool = False):
    """fp32 NHWC (or NCHW image) outputs: returns (raw or None, elu or None)."""
    outs = _ops.conv_fwd32(x[0], x[1] if len(x) > 1 else None, w, scale, shift, residual, cout, ksize, stride,
                           want_raw, want_elu, d2s, nchw, relu)
    it = iter(outs)
    return (next(it) if want_raw else None), (next(it) if want_elu else None)


def nchw_to_nhwc32(x, want_raw=True, want_elu=True):
    outs = _ops.nchw_to_nhwc32(x, want_raw, want_elu)
    it = iter(outs)
    return (next(it) if want_raw else None), (next(it) if want_elu else None)


def conv_dgrad32(dy, w, scale, dx, *, mask=(None, None), addend=(None, None), addend2=(None, None), cout: int,
                 ksize: int, stride: int = 1, d2s: bool = False, relu: bool = False) -> None:
    """fp32 validation counterpart of :func:`conv_dgrad` (``w``: the forward fp32 matrix ``v/||v||``)."""
    dx = list(dx) + [None] * (2 - len(dx))
    mask = list(mask) + [None] * (2 - len(mask))
    addend = list(addend) + [None] * (2 - len(addend))
    addend2 = list(addend2) + [None] * (2 - len(addend2))
    _ops.conv_dgrad32(dy, w, scale, dx[0], dx[1], mask[0], mask[1], addend[0], addend[1], addend2[0], addend2[1], cout,
                      ksize, stride, d2s, relu)


def conv_wgrad32(dy, x, gt, dy_sum, *, cout: int, ksize: int, stride: int = 1, d2s: bool = False) -> None:
    _ops.conv_wgrad32(dy, x[0], x[1] if len(x) > 1 else None, gt, dy_sum, cout, ksize, stride, d2s)


def latent_bwd32(logstd, eps, d_mu, d_ls, dz) -> torch.Tensor:
    return _ops.latent_bwd32(logstd, eps, d_mu, d_ls, dz)


# ---- training half ------------------------------------------------------------------------------
def conv_dgrad(dy, wd, dx, *, mask=(None, None), addend=(None, None), addend2=(None, None), ksize: int,
               stride: int = 1, d2s: bool = False, wd_tiled: Optional[torch.Tensor] = None, relu: bool = False) -> None:
    """``dx[s] = ELU'(mask[s]) * dgrad_s + addend[s] + addend2[s]`` for every source (in place into dx)."""
    dx = list(dx) + [None] * (2 - len(dx))
    mask = list(mask) + [None] * (2 - len(mask))
    addend = list(addend) + [None] * (2 - len(addend))
    addend2 = list(addend2) + [None] * (2 - len(addend2))
    _ops.conv_dgrad(dy, wd, dx[0], dx[1], mask[0], mask[1], addend[0], addend[1], addend2[0], addend2[1], ksize,
                    stride, d2s, wd_tiled, relu)


def conv_wgrad(dy, x, gt, dy_sum, *, ksize: int, stride: int = 1, d2s: bool = False) -> None:
    _ops.conv_wgrad(dy, x[0], x[1] if len(x) > 1 else None, gt, dy_sum, ksize, stride, d2s)


def wn_bwd(table: torch.Tensor, n_entries: int, total_rows: int, accumulate: bool = False, nbytes: float = 0.0,
           max_len: int = 0) -> None:
    """``max_len``: the largest ``cin * k * k`` among the table's layers (0 = 2304, the VunetAlter maximum)."""
    _ops.wn_bwd(table, n_entries, total_rows, accumulate, nbytes, max_len)


def adam_wn_step(table: torch.Tensor, n_entries: int, total_rows: int, beta1: float, beta2: float, eps: float,
                 step: int, nbytes: float = 0.0) -> None:
    _ops.adam_wn_step(table, n_entries, total_rows, beta1, beta2, eps, step, nbytes)


def adam_wn_step_dev(table: torch.Tensor, n_entries: int, total_rows: int, beta1: float, beta2: float, eps: float,
                     step: torch.Tensor, nbytes: float = 0.0) -> None:
    _ops.adam_wn_step_dev(table, n_entries, total_rows, beta1, beta2, eps, step, nbytes)


def stem_bwd(x, dy, gt, dy_sum) -> None:
    _ops.stem_bwd(x, dy, gt, dy_sum)


def latent_bwd(logstd, eps, d_mu, d_ls, dz) -> torch.Tensor:
    return _ops.latent_bwd(logstd, eps, d_mu, d_ls, dz)


def nchw_to_nhwc_pad(x, c_pad: int) -> torch.Tensor:
    return _ops.nchw_to_nhwc_pad(x, c_pad)


# ---- VunetOrg family ------------------------------------------------------------------------------
def latent_sample(p, eps, *, sub: int = 1, group=(0, 0), mean=None, z=None, z_raw=None, z_elu=None,
                  want_grp_elu: bool = False):
    """``latent_sample`` (models/vunets.py:151-156) on the fp32 NHWC parameter map ``p``; results are written into the
    (merged) tensors passed in.  Returns the compact ELU(z) of this map when ``want_grp_elu``."""
    outs = _ops.latent_sample(p, eps, sub, group[0], group[1], mean, z, z_raw, z_elu, want_grp_elu)
    return outs[0] if want_grp_elu else None


def latent_sample_bwd(d_mean, dz, *, sub: int, group, h: int, w: int, f32: bool) -> torch.Tensor:
    return _ops.latent_sample_bwd(d_mean, dz, sub, group[0], group[1], h, w, f32)


def space_to_depth(x, *, group_major: bool = False, inverse: bool = False, want_copy: bool = True,
                   want_elu: bool = False, accumulate_into=None):
    """SpaceToDepth(2) / its inverse on an NHWC activation: returns (copy or None, elu or None); with
    ``accumulate_into`` the regrouped values are added to that tensor and (None, None) is returned."""
    outs = _ops.space_to_depth(x, group_major, inverse, want_copy, want_elu, accumulate_into)
    if accumulate_into is not None:
        return None, None
    it = iter(outs)
    return (next(it) if want_copy else None), (next(it) if want_elu else None)


# ---- VGG19 perceptual loss ------------------------------------------------------------------------
def vgg_normalize(x, want_nhwc: bool = True, want_nchw: bool = False, f32: bool = False):
    outs = _ops.vgg_normalize(x, want_nhwc, want_nchw, f32)
    it = iter(outs)
    return (next(it) if want_nhwc else None), (next(it) if want_nchw else None)


def vgg_input_l1(pred, target) -> torch.Tensor:
    return _ops.vgg_input_l1(pred, target)


def vgg_input_bwd(pred, target, d_xn, g0, scale0: float) -> torch.Tensor:
    return _ops.vgg_input_bwd(pred, target, d_xn, g0, scale0)


def maxpool2_fwd(x) -> torch.Tensor:
    return _ops.maxpool2_fwd(x)


def maxpool2_relu_bwd(x, dy, addend=None) -> torch.Tensor:
    return _ops.maxpool2_relu_bwd(x, dy, addend)


def l1_partial(pf, tf) -> torch.Tensor:
    return _ops.l1_partial(pf, tf)


def l1_relu_grad(pf, tf, g, scale: float) -> torch.Tensor:
    return _ops.l1_relu_grad(pf, tf, g, scale)
