"""ctypes binding of the C-ABI library ``libvunet_b200.so`` (declared in ``include/vunet_b200.h``).

There is deliberately no fallback: if the library is missing or a launcher returns non-zero, a
``RuntimeError`` is raised.  ``build()`` compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvunet_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "vunet_b200.h")

# vub_conv_desc.flags (include/vunet_b200.h)
CONV_D2S = 1
CONV_F32_NHWC = 2
CONV_F32_NCHW = 4
CONV_ELU_IN0 = 8
CONV_ELU_IN1 = 16
CONV_U8_HWC = 32
CONV_RELU = 64
CONV_LRELU = 128
CONV_TANH = 256

# every extern "C" symbol the header declares; tests check the built library exports all of them
EXPORTED_SYMBOLS = (
    "vub_abi_version",
    "vub_last_error",
    "vub_conv_fwd",
    "vub_conv_fwd_elu_in_ok",
    "vub_conv_fwd_u8_out_ok",
    "vub_wn_prep",
    "vub_stem_fwd",
    "vub_stem_fwd_u8",
    "vub_stickman_render",
    "vub_allreduce_bucket",
    "vub_nccl_version",
    "vub_nccl_unique_id",
    "vub_nccl_comm_init",
    "vub_nccl_comm_destroy",
    "vub_latent_num_partials",
    "vub_latent_fwd",
    "vub_nchw_f32_to_nhwc_bf16",
    "vub_nhwc_bf16_to_nchw_f32",
    "vub_nchw_f32_to_nhwc_bf16_pad",
    "vub_frames_to_uint8",
    "vub_conv_fwd_f32",
    "vub_nchw_f32_to_nhwc_f32",
    "vub_conv_dgrad_f32",
    "vub_conv_wgrad_f32",
    "vub_latent_bwd_f32",
    "vub_conv_dgrad",
    "vub_conv_wgrad",
    "vub_conv_wgrad_splits",
    "vub_wn_bwd",
    "vub_adam_wn_step",
    "vub_adam_wn_step_dev",
    "vub_stem_bwd",
    "vub_latent_bwd",
    "vub_latent_sample",
    "vub_space_to_depth",
    "vub_latent_sample_bwd",
    "vub_vgg_normalize",
    "vub_vgg_input_l1_partials",
    "vub_vgg_input_l1",
    "vub_vgg_input_bwd",
    "vub_maxpool2_fwd",
    "vub_maxpool2_relu_bwd",
    "vub_l1_num_partials",
    "vub_l1_partial",
    "vub_l1_relu_grad",
    "vub_gather_affine",
    "vub_couple_inv",
    "vub_lstm_cell",
    "vub_lstm_input",
    "vub_decoder_step",
    "vub_decoder_init",
)
ABI_VERSION = 5


class ConvDesc(C.Structure):
    """Mirror of ``vub_conv_desc``."""

    _fields_ = [
        ("x", C.c_void_p * 2),
        ("w", C.c_void_p),
        ("scale", C.c_void_p),
        ("shift", C.c_void_p),
        ("residual", C.c_void_p),
        ("out_raw", C.c_void_p),
        ("out_elu", C.c_void_p),
        ("out_f32", C.c_void_p),
        ("nsrc", C.c_int32),
        ("c_src", C.c_int32),
        ("batch", C.c_int32),
        ("hin", C.c_int32),
        ("win", C.c_int32),
        ("cout", C.c_int32),
        ("cout_valid", C.c_int32),
        ("ksize", C.c_int32),
        ("stride", C.c_int32),
        ("flags", C.c_uint32),
        ("stages", C.c_int32),
        ("w_tiled", C.c_void_p),
    ]


class WnEntry(C.Structure):
    """Mirror of ``vub_wn_entry``."""

    _fields_ = [
        ("v", C.c_void_p),
        ("g", C.c_void_p),
        ("gamma", C.c_void_p),
        ("beta", C.c_void_p),
        ("bias", C.c_void_p),
        ("w_packed", C.c_void_p),
        ("scale", C.c_void_p),
        ("shift", C.c_void_p),
        ("cout", C.c_int32),
        ("cin", C.c_int32),
        ("ktaps", C.c_int32),
        ("c_src", C.c_int32),
        ("ktot", C.c_int32),
        ("row_start", C.c_int32),
        ("wd", C.c_void_p),
        ("wd_cout", C.c_int32),
        ("wd_col0", C.c_int32),
        ("w32", C.c_void_p),
        ("w_tiled", C.c_void_p),
        ("wd_tiled", C.c_void_p),
        ("tiled_rows", C.c_int32),
        ("pad_", C.c_int32),
    ]


class DgradF32Desc(C.Structure):
    """Mirror of ``vub_dgrad_f32_desc``."""

    _fields_ = [
        ("dy", C.c_void_p),
        ("w", C.c_void_p),
        ("scale", C.c_void_p),
        ("dx", C.c_void_p * 2),
        ("mask_elu", C.c_void_p * 2),
        ("addend", C.c_void_p * 2),
        ("addend2", C.c_void_p * 2),
        ("nsrc", C.c_int32),
        ("c_src", C.c_int32),
        ("batch", C.c_int32),
        ("hin", C.c_int32),
        ("win", C.c_int32),
        ("cout", C.c_int32),
        ("w_pitch", C.c_int32),
        ("ksize", C.c_int32),
        ("stride", C.c_int32),
        ("flags", C.c_uint32),
    ]


class WgradF32Desc(C.Structure):
    """Mirror of ``vub_wgrad_f32_desc``."""

    _fields_ = [
        ("dy", C.c_void_p),
        ("x", C.c_void_p * 2),
        ("gt", C.c_void_p),
        ("dy_sum", C.c_void_p),
        ("nsrc", C.c_int32),
        ("c_src", C.c_int32),
        ("batch", C.c_int32),
        ("hin", C.c_int32),
        ("win", C.c_int32),
        ("cout", C.c_int32),
        ("ksize", C.c_int32),
        ("stride", C.c_int32),
        ("flags", C.c_uint32),
        ("n_split", C.c_int32),
        ("gt_pitch", C.c_int32),
        ("split_stride", C.c_int64),
        ("sum_stride", C.c_int64),
    ]


class StickmanModel(C.Structure):
    """Mirror of ``vub_stickman_model``."""

    _fields_ = [
        ("n_body", C.c_int32), ("body", C.c_int32 * 8),
        ("n_right", C.c_int32), ("right", (C.c_int32 * 2) * 16),
        ("n_left", C.c_int32), ("left", (C.c_int32 * 2) * 16),
        ("n_head", C.c_int32), ("head", (C.c_int32 * 2) * 8),
        ("n_face", C.c_int32), ("face", (C.c_int32 * 2) * 16),
        ("rshoulder", C.c_int32), ("lshoulder", C.c_int32), ("headup", C.c_int32),
        ("colored", C.c_int32),
        ("right_chan", C.c_int32 * 16), ("right_col", C.c_int32 * 16),
        ("left_chan", C.c_int32 * 16), ("left_col", C.c_int32 * 16),
        ("head_chan", C.c_int32 * 16), ("head_col", C.c_int32 * 16),
    ]


class ConvF32Desc(C.Structure):
    """Mirror of ``vub_conv_f32_desc`` (fp32 validation path)."""

    _fields_ = [
        ("x", C.c_void_p * 2),
        ("w", C.c_void_p),
        ("scale", C.c_void_p),
        ("shift", C.c_void_p),
        ("residual", C.c_void_p),
        ("out_raw", C.c_void_p),
        ("out_elu", C.c_void_p),
        ("nsrc", C.c_int32),
        ("c_src", C.c_int32),
        ("batch", C.c_int32),
        ("hin", C.c_int32),
        ("win", C.c_int32),
        ("cout", C.c_int32),
        ("w_pitch", C.c_int32),
        ("ksize", C.c_int32),
        ("stride", C.c_int32),
        ("flags", C.c_uint32),
    ]


class DgradDesc(C.Structure):
    """Mirror of ``vub_dgrad_desc``."""

    _fields_ = [
        ("dy", C.c_void_p),
        ("wd", C.c_void_p),
        ("dx", C.c_void_p * 2),
        ("mask_elu", C.c_void_p * 2),
        ("addend", C.c_void_p * 2),
        ("addend2", C.c_void_p * 2),
        ("nsrc", C.c_int32),
        ("c_src", C.c_int32),
        ("batch", C.c_int32),
        ("hin", C.c_int32),
        ("win", C.c_int32),
        ("cout_k", C.c_int32),
        ("ksize", C.c_int32),
        ("stride", C.c_int32),
        ("flags", C.c_uint32),
        ("pad_", C.c_int32),
        ("wd_tiled", C.c_void_p),
    ]


class WgradDesc(C.Structure):
    """Mirror of ``vub_wgrad_desc``."""

    _fields_ = [
        ("dy", C.c_void_p),
        ("x", C.c_void_p * 2),
        ("gt", C.c_void_p),
        ("dy_sum", C.c_void_p),
        ("split_stride", C.c_int64),
        ("sum_stride", C.c_int32),
        ("n_split", C.c_int32),
        ("nsrc", C.c_int32),
        ("c_src", C.c_int32),
        ("batch", C.c_int32),
        ("hin", C.c_int32),
        ("win", C.c_int32),
        ("cout_k", C.c_int32),
        ("gt_pitch", C.c_int32),
        ("ksize", C.c_int32),
        ("stride", C.c_int32),
        ("flags", C.c_uint32),
    ]


class WnbEntry(C.Structure):
    """Mirror of ``vub_wnb_entry``."""

    _fields_ = [
        ("v", C.c_void_p),
        ("g", C.c_void_p),
        ("gamma", C.c_void_p),
        ("bias", C.c_void_p),
        ("gt", C.c_void_p),
        ("dy_sum", C.c_void_p),
        ("d_v", C.c_void_p),
        ("d_g", C.c_void_p),
        ("d_gamma", C.c_void_p),
        ("d_beta", C.c_void_p),
        ("d_bias", C.c_void_p),
        ("cout", C.c_int32),
        ("cin", C.c_int32),
        ("ktaps", C.c_int32),
        ("c_src", C.c_int32),
        ("gt_pitch", C.c_int32),
        ("gt_col0", C.c_int32),
        ("row_start", C.c_int32),
        ("n_split", C.c_int32),
        ("split_stride", C.c_int64),
        ("sum_stride", C.c_int32),
        ("pad", C.c_int32),
    ]


class AdamEntry(C.Structure):
    """Mirror of ``vub_adam_entry``."""

    _fields_ = ([("wn", WnEntry)]
                + [(f"{a}_{b}", C.c_void_p) for a in ("d", "m", "s") for b in ("v", "g", "gamma", "beta", "bias")]
                + [("lr", C.c_float), ("pad", C.c_int32)])


_lib = None
_lock = threading.Lock()


def build(verbose: bool = False) -> str:
    """Compile ``csrc/*.cu`` into ``libvunet_b200.so`` (nvcc, -gencode arch=compute_100a,code=sm_100a)."""
    proc = subprocess.run(
        ["make", "-C", CSRC_DIR, "-j8"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True
    )
    if verbose:
        print(proc.stdout)
    if proc.returncode != 0 or not os.path.exists(LIB_PATH):
        raise RuntimeError("building libvunet_b200.so failed:\n" + proc.stdout[-4000:])
    return LIB_PATH


def _declare(lib: C.CDLL) -> None:
    vp, i32, fp = C.c_void_p, C.c_int32, C.c_void_p
    lib.vub_abi_version.restype = C.c_int
    lib.vub_abi_version.argtypes = []
    lib.vub_last_error.restype = C.c_char_p
    lib.vub_last_error.argtypes = []
    lib.vub_conv_fwd.restype = C.c_int
    lib.vub_conv_fwd.argtypes = [C.POINTER(ConvDesc), vp]
    lib.vub_wn_prep.restype = C.c_int
    lib.vub_wn_prep.argtypes = [vp, i32, i32, vp]
    lib.vub_stem_fwd.restype = C.c_int
    lib.vub_stem_fwd.argtypes = [fp, fp, fp, fp, fp, fp, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.vub_allreduce_bucket.restype = C.c_int
    lib.vub_allreduce_bucket.argtypes = [vp, C.c_int64, vp, i32, vp]
    lib.vub_nccl_version.restype = i32
    lib.vub_nccl_version.argtypes = []
    lib.vub_nccl_unique_id.restype = C.c_int
    lib.vub_nccl_unique_id.argtypes = [vp]
    lib.vub_nccl_comm_init.restype = C.c_int
    lib.vub_nccl_comm_init.argtypes = [C.POINTER(vp), i32, vp, i32]
    lib.vub_nccl_comm_destroy.restype = C.c_int
    lib.vub_nccl_comm_destroy.argtypes = [vp]
    lib.vub_stickman_render.restype = C.c_int
    lib.vub_stickman_render.argtypes = [vp, i32, i32, C.POINTER(StickmanModel), i32, i32, vp, vp]
    lib.vub_stem_fwd_u8.restype = C.c_int
    lib.vub_stem_fwd_u8.argtypes = [vp, fp, fp, fp, fp, fp, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.vub_latent_num_partials.restype = i32
    lib.vub_latent_num_partials.argtypes = [i32, i32, i32, i32]
    lib.vub_latent_fwd.restype = C.c_int
    lib.vub_latent_fwd.argtypes = [fp, fp, i32, i32, i32, i32, fp, fp, fp, vp, vp, fp, vp]
    lib.vub_nchw_f32_to_nhwc_bf16.restype = C.c_int
    lib.vub_nchw_f32_to_nhwc_bf16.argtypes = [fp, i32, i32, i32, i32, vp, vp, vp]
    lib.vub_nhwc_bf16_to_nchw_f32.restype = C.c_int
    lib.vub_nhwc_bf16_to_nchw_f32.argtypes = [vp, i32, i32, i32, i32, fp, vp]
    lib.vub_nchw_f32_to_nhwc_bf16_pad.restype = C.c_int
    lib.vub_nchw_f32_to_nhwc_bf16_pad.argtypes = [fp, i32, i32, i32, i32, i32, vp, vp]
    lib.vub_frames_to_uint8.restype = C.c_int
    lib.vub_frames_to_uint8.argtypes = [fp, i32, i32, i32, vp, vp]
    lib.vub_conv_fwd_f32.restype = C.c_int
    lib.vub_conv_fwd_f32.argtypes = [C.POINTER(ConvF32Desc), vp]
    lib.vub_conv_dgrad_f32.restype = C.c_int
    lib.vub_conv_dgrad_f32.argtypes = [C.POINTER(DgradF32Desc), vp]
    lib.vub_conv_wgrad_f32.restype = C.c_int
    lib.vub_conv_wgrad_f32.argtypes = [C.POINTER(WgradF32Desc), vp]
    lib.vub_latent_bwd_f32.restype = C.c_int
    lib.vub_latent_bwd_f32.argtypes = [fp, fp, fp, fp, fp, i32, i32, i32, i32, fp, vp]
    lib.vub_nchw_f32_to_nhwc_f32.restype = C.c_int
    lib.vub_nchw_f32_to_nhwc_f32.argtypes = [fp, i32, i32, i32, i32, fp, fp, vp]
    lib.vub_conv_dgrad.restype = C.c_int
    lib.vub_conv_dgrad.argtypes = [C.POINTER(DgradDesc), vp]
    lib.vub_conv_wgrad_splits.restype = i32
    lib.vub_conv_wgrad_splits.argtypes = [C.POINTER(WgradDesc)]
    lib.vub_conv_wgrad.restype = C.c_int
    lib.vub_conv_wgrad.argtypes = [C.POINTER(WgradDesc), vp]
    lib.vub_conv_fwd_elu_in_ok.restype = C.c_int
    lib.vub_conv_fwd_elu_in_ok.argtypes = [i32] * 7
    lib.vub_conv_fwd_u8_out_ok.restype = C.c_int
    lib.vub_conv_fwd_u8_out_ok.argtypes = [i32] * 7
    lib.vub_adam_wn_step_dev.restype = C.c_int
    lib.vub_adam_wn_step_dev.argtypes = [vp, i32, i32, C.c_float, C.c_float, C.c_float, vp, vp]
    lib.vub_adam_wn_step.restype = C.c_int
    lib.vub_adam_wn_step.argtypes = [vp, i32, i32, C.c_float, C.c_float, C.c_float, i32, vp]
    lib.vub_wn_bwd.restype = C.c_int
    lib.vub_wn_bwd.argtypes = [vp, i32, i32, i32, i32, vp]
    lib.vub_stem_bwd.restype = C.c_int
    lib.vub_stem_bwd.argtypes = [fp, vp, i32, i32, i32, i32, i32, fp, fp, vp]
    lib.vub_latent_bwd.restype = C.c_int
    lib.vub_latent_bwd.argtypes = [fp, fp, fp, fp, fp, vp, i32, i32, i32, i32, vp, vp]
    lib.vub_latent_sample.restype = C.c_int
    lib.vub_latent_sample.argtypes = [fp, fp, i32, i32, i32, i32, i32, i32, i32, fp, fp, vp, vp, vp, i32, vp]
    lib.vub_vgg_normalize.restype = C.c_int
    lib.vub_vgg_normalize.argtypes = [fp, i32, i32, i32, vp, fp, i32, vp]
    lib.vub_vgg_input_l1_partials.restype = i32
    lib.vub_vgg_input_l1_partials.argtypes = [i32, i32, i32]
    lib.vub_vgg_input_l1.restype = C.c_int
    lib.vub_vgg_input_l1.argtypes = [fp, fp, i32, i32, i32, fp, vp]
    lib.vub_vgg_input_bwd.restype = C.c_int
    lib.vub_vgg_input_bwd.argtypes = [fp, fp, vp, fp, C.c_float, i32, i32, i32, fp, i32, vp]
    lib.vub_maxpool2_fwd.restype = C.c_int
    lib.vub_maxpool2_fwd.argtypes = [vp, i32, i32, i32, i32, vp, i32, vp]
    lib.vub_maxpool2_relu_bwd.restype = C.c_int
    lib.vub_maxpool2_relu_bwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, i32, vp]
    lib.vub_l1_num_partials.restype = i32
    lib.vub_l1_num_partials.argtypes = [C.c_int64]
    lib.vub_l1_partial.restype = C.c_int
    lib.vub_l1_partial.argtypes = [vp, vp, C.c_int64, fp, i32, vp]
    lib.vub_l1_relu_grad.restype = C.c_int
    lib.vub_l1_relu_grad.argtypes = [vp, vp, C.c_int64, fp, C.c_float, vp, i32, vp]
    lib.vub_gather_affine.restype = C.c_int
    lib.vub_gather_affine.argtypes = [fp, vp, fp, fp, i32, i32, fp, vp, i32, i32, vp]
    lib.vub_couple_inv.restype = C.c_int
    lib.vub_couple_inv.argtypes = [fp, fp, fp, i32, i32, i32, i32, fp, vp, vp]
    lib.vub_lstm_cell.restype = C.c_int
    lib.vub_lstm_cell.argtypes = [fp, fp, i32, i32, vp, vp, i32, fp, vp]
    lib.vub_lstm_input.restype = C.c_int
    lib.vub_lstm_input.argtypes = [fp, C.c_int64, i32, i32, i32, vp, i32, i32, vp]
    lib.vub_decoder_step.restype = C.c_int
    lib.vub_decoder_step.argtypes = [fp, fp, i32, i32, i32, i32, i32, i32, fp, fp, vp, i32, i32, vp]
    lib.vub_decoder_init.restype = C.c_int
    lib.vub_decoder_init.argtypes = [fp, fp, i32, i32, i32, i32, fp, fp, vp, i32, vp]
    lib.vub_latent_sample_bwd.restype = C.c_int
    lib.vub_latent_sample_bwd.argtypes = [fp, vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp]
    lib.vub_space_to_depth.restype = C.c_int
    lib.vub_space_to_depth.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp]


def lib() -> C.CDLL:
    """The loaded C-ABI library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: the CUDA extension is not built. Run "
                        "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)."
                    )
                handle = C.CDLL(LIB_PATH)
                _declare(handle)
                _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vub_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with code {rc}: {msg}")
