"""CPU: host-side logic — module structure / state_dict layout / init stream / ctor contract, the C-ABI
library (loads, exports every symbol the header declares), and loud failure without a GPU."""
import contextlib
import ctypes
import io
import os
import re
import sys

import pytest
import torch

import vunet_oracle as O
from conftest import REFERENCE_ROOT, ROOT

import bdvs_b200 as B
from bdvs_b200 import _native

H256 = dict(spatial_size=256, bottleneck_factor=2, n_scales=0, box_factor=2, n_latent_scales=2,
            conv_layer_type="l1", nf_start=32, nf_max=128, subpixel_upsampling=True)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.mark.parametrize("name", ["tiny3", "tiny30", "market128"])
def test_state_dict_layout_matches_reference_fixture(golden, name):
    fx = golden(name)
    cfg = dict(fx["config"])
    cfg.pop("batch")
    m = quiet(B.VunetAlter, **cfg)
    got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert got == [(k, tuple(s)) for k, s in fx["param_shapes"]]  # names, shapes AND registration order
    assert all(v.dtype == torch.float32 for v in m.state_dict().values())
    # reference-generated checkpoint loads strictly, and round-trips
    sd = O.det_state_dict(dict(fx["param_shapes"]))
    m.load_state_dict(sd, strict=True)
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k])


def test_ctor_contract():
    m = quiet(B.VunetAlter, n_channels_x=3, unknown_key=1, dataset="Human3.6m", **H256)  # unknown kwargs ignored
    assert (m.n_scales, m.n_scales_x, m.spatial_size, m.n_channels_x) == (7, 7, 256, 3)
    assert len(m.state_dict()) == 495 and sum(v.numel() for v in m.state_dict().values()) == 14601516
    m30 = quiet(B.VunetAlter, n_channels_x=30, **H256)
    assert m30.n_scales_x == 5 and len(m30.state_dict()) == 465
    for sub in ("eu", "ed", "du", "dd"):
        assert hasattr(m, sub)
    with pytest.raises(KeyError):
        quiet(B.VunetAlter, n_channels_x=3)  # required kwargs, like the reference
    with pytest.raises(NotImplementedError):
        quiet(B.VunetAlter, **dict(H256, conv_layer_type="l2"))
    # dropout_prob never reaches the ctor in the reference's callers -> p = 0 everywhere
    assert all(mod.p == 0.0 for mod in m.modules() if isinstance(mod, torch.nn.Dropout))


@pytest.mark.skipif(not os.path.isdir(REFERENCE_ROOT), reason="reference tree only exists in the build container")
def test_same_seed_gives_reference_initialisation():
    sys.path.insert(0, REFERENCE_ROOT)
    from models.vunets import VunetAlter as RefVunet
    cfg = dict(H256, spatial_size=64, nf_max=64)
    torch.manual_seed(42)
    ref = quiet(RefVunet, n_channels_x=30, **cfg)
    torch.manual_seed(42)
    mine = quiet(B.VunetAlter, n_channels_x=30, **cfg)
    rsd, msd = ref.state_dict(), mine.state_dict()
    assert list(rsd.keys()) == list(msd.keys())
    for k in rsd:
        assert torch.equal(rsd[k], msd[k]), k  # bit-identical: same RNG stream, same formulas


def test_no_cpu_fallback():
    m = quiet(B.VunetAlter, n_channels_x=3, **dict(H256, spatial_size=32, bottleneck_factor=1))
    x = torch.zeros(1, 3, 32, 32)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(x, x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.transfer(x, x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        B.NormConv2d(32, 32, 3, padding=1)(torch.zeros(1, 32, 8, 8))
    with pytest.raises(NotImplementedError):  # torch op registered for CUDA only
        torch.ops.vunet_b200.nhwc_to_nchw(torch.zeros(1, 2, 2, 8, dtype=torch.bfloat16))


def test_kl_matches_oracle():
    mu = [O.det_normal("m0", (3, 8, 2, 2)), O.det_normal("m1", (3, 8, 4, 4))]
    ls = [torch.sigmoid(O.det_normal("l0", (3, 8, 2, 2))), torch.sigmoid(O.det_normal("l1", (3, 8, 4, 4)))]
    assert torch.allclose(B.compute_kl_with_prior(mu, ls), O.compute_kl_with_prior(mu, ls), rtol=1e-6)
    assert torch.allclose(B.kl_loss(mu[0], ls[0]), O.kl_loss(mu[0], ls[0]), rtol=1e-6)


def test_space_depth_modules():
    x = O.det_uniform("sd", (2, 8, 4, 6))
    assert torch.equal(B.DepthToSpace(2)(x), O.depth_to_space(x))
    assert torch.equal(B.SpaceToDepth(2)(B.DepthToSpace(2)(x)), x)


# ---- C ABI ------------------------------------------------------------------------------------
def header_symbols():
    text = open(os.path.join(ROOT, "include", "vunet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vub_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_native.EXPORTED_SYMBOLS)


def test_library_loads_and_exports_every_symbol():
    if not os.path.exists(_native.LIB_PATH):
        _native.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    for sym in header_symbols():
        assert hasattr(lib, sym), sym
    assert _native.lib().vub_abi_version() == _native.ABI_VERSION
    # struct mirrors must match the C layout (pointers first, then int32 fields)
    assert ctypes.sizeof(_native.ConvDesc) == 9 * 8 + 11 * 4 + 4 + 8  # padded to 8, + w_tiled
    assert ctypes.sizeof(_native.WnEntry) == 8 * 8 + 6 * 4 + 8 + 2 * 4 + 8 + 2 * 8 + 2 * 4
    assert ctypes.sizeof(_native.ConvF32Desc) == 8 * 8 + 10 * 4
    assert ctypes.sizeof(_native.DgradDesc) == 10 * 8 + 9 * 4 + 4 + 8  # + wd_tiled
    assert ctypes.sizeof(_native.WgradDesc) == 6 * 8 + 12 * 4
    assert ctypes.sizeof(_native.WnbEntry) == 11 * 8 + 8 * 4 + 8 + 2 * 4
    assert ctypes.sizeof(_native.AdamEntry) == ctypes.sizeof(_native.WnEntry) + 15 * 8 + 2 * 4


def test_bad_arguments_are_reported_not_crashed():
    lib = _native.lib()
    assert lib.vub_conv_fwd(None, None) == -1  # VUB_E_NULL
    assert b"NULL" in lib.vub_last_error()
    d = _native.ConvDesc()
    d.x[0] = 16; d.w = 16; d.scale = 16; d.shift = 16
    d.nsrc, d.c_src, d.batch, d.hin, d.win, d.cout, d.cout_valid, d.ksize, d.stride = 1, 48, 1, 8, 8, 32, 32, 3, 1
    assert lib.vub_conv_fwd(ctypes.byref(d), None) == -2  # VUB_E_SHAPE: c_src 48
    d.c_src, d.hin = 32, 12
    assert lib.vub_conv_fwd(ctypes.byref(d), None) == -2  # non power-of-two extent
    d.hin = 8
    assert lib.vub_conv_fwd(ctypes.byref(d), None) == -3  # VUB_E_FLAGS: no output requested
    assert lib.vub_stem_fwd(None, None, None, None, None, None, 1, 3, 8, 8, 32, None, None, None) == -1
    assert lib.vub_wn_prep(None, 1, 1, None) == -1
    assert lib.vub_latent_num_partials(2, 128, 4, 4) == 8
    # training entry points validate before touching the device too
    assert lib.vub_conv_dgrad(None, None) == -1
    assert lib.vub_conv_wgrad(None, None) == -1
    assert lib.vub_wn_bwd(None, 1, 1, 0, 0, None) == -1
    assert lib.vub_stem_bwd(None, None, 1, 3, 8, 8, 32, None, None, None) == -1
    assert lib.vub_latent_bwd(None, None, None, None, None, None, 1, 128, 4, 4, None, None) == -1
    g = _native.DgradDesc()
    g.dy = 16; g.wd = 16; g.dx[0] = 16
    g.nsrc, g.c_src, g.batch, g.hin, g.win, g.cout_k, g.ksize, g.stride = 1, 32, 1, 8, 8, 48, 3, 1
    assert lib.vub_conv_dgrad(ctypes.byref(g), None) == -2  # cout_k 48 not a multiple of 32
    w = _native.WgradDesc()
    w.dy = 16; w.x[0] = 16; w.gt = 16; w.dy_sum = 16
    w.nsrc, w.c_src, w.batch, w.hin, w.win, w.cout_k, w.gt_pitch, w.ksize, w.stride = 1, 32, 1, 8, 8, 64, 32, 3, 1
    w.n_split = 1
    assert lib.vub_conv_wgrad(ctypes.byref(w), None) == -2  # gt_pitch < K = 288
    w.gt_pitch = 288
    assert lib.vub_conv_wgrad_splits(ctypes.byref(w)) == 1  # one 64-pixel tile -> one split (pure host planning)
    w.n_split = 5
    assert lib.vub_conv_wgrad(ctypes.byref(w), None) == -2  # more splits than pixel tiles


def test_host_side_planning_queries():
    """Pure host logic of the C ABI that decides data layouts: no GPU needed."""
    lib = _native.lib()
    ok = lib.vub_conv_fwd_elu_in_ok
    # ELU-on-load exists where the row-streaming kernel runs: width >= 128, <= 64 channels, stride 1, square layers
    assert ok(1, 32, 32, 3, 1, 256, 256) == 1      # residual-block 3x3 at 256^2
    assert ok(1, 64, 64, 3, 1, 128, 128) == 1      # ... at 128^2
    assert ok(1, 32, 32, 1, 1, 256, 256) == 1      # skip-connection 1x1
    assert ok(2, 32, 32, 3, 1, 256, 256) == 1      # concat block, 32 channels per source
    assert ok(2, 64, 64, 3, 1, 128, 128) == 0      # concat block, 64 per source: weights do not fit beside the ring
    assert ok(1, 128, 128, 3, 1, 64, 64) == 0      # narrow maps run on the tile kernels
    assert ok(1, 32, 64, 3, 2, 256, 256) == 0      # stride 2
    assert ok(1, 32, 64, 3, 1, 256, 256) == 0      # not square
    # weight-gradient splits: every split keeps >= 4 pixel k-steps (partials below the layer's own input bytes)
    w = _native.WgradDesc()
    w.nsrc, w.c_src, w.ksize, w.stride = 1, 128, 3, 1
    w.cout_k, w.gt_pitch = 128, 1152
    for batch, hw in ((32, 4), (32, 8), (32, 16), (32, 64)):
        w.batch, w.hin, w.win = batch, hw, hw
        n = lib.vub_conv_wgrad_splits(ctypes.byref(w))
        pixels = batch * hw * hw
        assert 1 <= n <= max(1, pixels // (4 * 64)), (hw, n)
    w.batch, w.hin, w.win = 32, 64, 64
    assert lib.vub_conv_wgrad_splits(ctypes.byref(w)) == 148 // 3   # 3 tap-row groups fill the 148 SMs


def test_lean_activation_policy_follows_the_query():
    from bdvs_b200.modules import lean_elu

    class Bank:  # stands in for a WeightBank
        lean = True

    assert lean_elu(Bank, 32, 256, 256) and lean_elu(Bank, 64, 128, 128)
    assert lean_elu(Bank, 32, 256, 256, concat=True)
    assert not lean_elu(Bank, 64, 128, 128, concat=True)     # dd blocks at 128^2 keep their ELU copies
    assert not lean_elu(Bank, 128, 64, 64)
    Bank.lean = False                                         # training (tape) or fp32 validation: pairs everywhere
    assert not lean_elu(Bank, 32, 256, 256)


def test_fused_adam_argument_checks():
    import torch
    import bdvs_b200 as B
    m = B.NormConv2d(32, 32, 3, padding=1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        B.FusedAdam(m.parameters(), lr=1e-3, model=m)
    with pytest.raises(NotImplementedError):
        B.FusedAdam([{"params": list(m.parameters()), "weight_decay": 0.1}], model=m)
    with pytest.raises(NotImplementedError):
        B.FusedAdam([torch.nn.Parameter(torch.zeros(3))], model=m)
    with pytest.raises(ValueError):
        B.FusedAdam([], model=m)


def test_integration_md_stub_matches_the_binding():
    """The ctypes stub shown to reference maintainers must describe the same struct as the package's binding."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = text[text.index("class ConvDesc(ctypes.Structure)"):text.index("lib.vub_conv_fwd.argtypes")]
    names = re.findall(r'\("(\w+)",\s*ctypes\.', block)
    assert names == [f[0] for f in _native.ConvDesc._fields_]
