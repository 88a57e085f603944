"""Behavior-net sampling (BASELINE.json configs[3]) — host-side mirror of the reference's latent flow and residual LSTM
decoder: ``UnsupervisedTransformer2`` (models/flow/simple_flow.py:136-180; models/flow/blocks.py:95-128, 276-319,
531-559, 692-704; lib/modules.py:236-330) and ``ResidualBehaviorNet`` / ``ResidualRNNDecoder`` / ``BEncoder``
(models/pose_behavior_rnn.py:125-210, 463-626).

Same class names, constructor arguments, parameter names and registration order (``state_dict`` round-trips with the
reference's checkpoints; same-seed initialisation is bit-identical).  What runs here is the SAMPLING path the reference
uses to turn Gaussian noise into 3-D pose sequences (lib/logging.py:280-300): ``b = flow.reverse(z)`` and
``net.generate_seq(b, x_pose, len, start_frame)``.  Every matrix product — 240 linear layers of the 15 coupling blocks,
the LSTM gates and ``n_out`` of every time step — is a "linear layer" launch of the tcgen05 per-tap kernel (1x1 conv on
a [batch, 1, 1, features] tensor, split-K over a thread-block cluster, LeakyReLU epilogue); the glue between them
(shuffle + ActNorm, the coupling update, the LSTM cell, the residual step) are the small kernels of ``csrc/behavior.cu``.
:class:`PoseSequenceSampler` captures noise -> sequences in one CUDA graph.  ``ResidualBehaviorNet.forward`` /
``infer_b`` (the LSTM encoder: behavior transfer and reconstruction) run on the same kernels.  Training of these models
and the flow's density direction are outside the hot path: those entry points raise or are inference-only.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from . import _native as N
from . import ops
from .modules import _WEIGHT_EPOCH, WeightBank


# ------------------------------------------------------------------------------------------------
# parameter holders with the reference's structure (lib/modules.py:236-330, models/flow/blocks.py)
# ------------------------------------------------------------------------------------------------
class BasicFullyConnectedNet(nn.Module):
    """lib/modules.py:236-257."""

    def __init__(self, dim, depth, hidden_dim=256, use_tanh=False, use_bn=False, out_dim=None):
        super().__init__()
        if use_bn:
            raise NotImplementedError("vunet_b200: BatchNorm1d in BasicFullyConnectedNet is not used by the behavior net")
        layers = [nn.Linear(dim, hidden_dim), nn.LeakyReLU()]
        for _ in range(depth):
            layers += [nn.Linear(hidden_dim, hidden_dim), nn.LeakyReLU()]
        layers.append(nn.Linear(hidden_dim, dim if out_dim is None else out_dim))
        if use_tanh:
            layers.append(nn.Tanh())
        self.main = nn.Sequential(*layers)
        self.use_tanh = use_tanh

    def linears(self) -> List[nn.Linear]:
        return [m for m in self.main if isinstance(m, nn.Linear)]

    def forward(self, x):
        raise NotImplementedError("vunet_b200: the coupling MLPs run inside UnsupervisedTransformer2.reverse")


class ActNorm(nn.Module):
    """lib/modules.py:260-330 (parameters only; the reverse pass is fused into the flow's gather kernel)."""

    def __init__(self, num_features, logdet=False, affine=True):
        assert affine
        super().__init__()
        self.logdet = logdet
        self.loc = nn.Parameter(torch.zeros(1, num_features, 1, 1))
        self.scale = nn.Parameter(torch.ones(1, num_features, 1, 1))
        self.register_buffer("initialized", torch.tensor(0, dtype=torch.uint8))


class Shuffle(nn.Module):
    """models/flow/blocks.py:692-704."""

    def __init__(self, in_channels, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        idx = torch.randperm(in_channels)
        self.register_buffer("forward_shuffle_idx", nn.Parameter(idx, requires_grad=False))
        self.register_buffer("backward_shuffle_idx", nn.Parameter(torch.argsort(idx), requires_grad=False))


class DoubleVectorCouplingBlock2(nn.Module):
    """models/flow/blocks.py:276-319."""

    def __init__(self, in_channels, hidden_dim, hidden_depth=2):
        super().__init__()
        dim1 = (in_channels // 2) + (in_channels % 2)
        dim2 = in_channels // 2
        self.s = nn.ModuleList([BasicFullyConnectedNet(dim=dim1, out_dim=dim2, depth=hidden_depth, hidden_dim=hidden_dim,
                                                       use_tanh=True) for _ in range(2)])
        self.t = nn.ModuleList([BasicFullyConnectedNet(dim=dim1, out_dim=dim2, depth=hidden_depth, hidden_dim=hidden_dim,
                                                       use_tanh=False) for _ in range(2)])


class UnconditionalFlatDoubleCouplingFlowBlock2(nn.Module):
    """models/flow/blocks.py:531-559."""

    def __init__(self, in_channels, hidden_dim, hidden_depth):
        super().__init__()
        self.norm_layer = ActNorm(in_channels, logdet=True)
        self.coupling = DoubleVectorCouplingBlock2(in_channels, hidden_dim, hidden_depth)
        self.shuffle = Shuffle(in_channels)


class UnconditionalFlow2(nn.Module):
    """models/flow/blocks.py:95-128."""

    def __init__(self, in_channels, hidden_dim, hidden_depth, n_flows):
        super().__init__()
        self.in_channels = in_channels
        self.mid_channels = hidden_dim
        self.num_blocks = hidden_depth
        self.n_flows = n_flows
        self.sub_layers = nn.ModuleList()
        for _ in range(self.n_flows):
            self.sub_layers.append(UnconditionalFlatDoubleCouplingFlowBlock2(self.in_channels, self.mid_channels,
                                                                             self.num_blocks))


class _LinearAdapter:
    """A frozen-for-this-call ``nn.Linear`` (or a concatenation of them) presented to WeightBank as a 1x1 NormConv2d with
    identity normalisation (v = W, g = ||W_row||, gamma = 1, beta = 0); K padded with zero columns to a multiple of 64,
    N padded to the kernels' tile sizes with rows whose scale (gamma) is 0."""

    _vub_stem = False
    _vub_split = 1
    _vub_linear = True
    kernel_size = 1
    stride = 1

    class _P:
        __slots__ = ("weight_v", "weight_g", "bias")

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor):
        w = weight.detach().float()
        cout, cin = w.shape
        kpad = (cin + 63) // 64 * 64
        npad = next((p for p in (64, 128, 256, 512) if cout <= p), (cout + 127) // 128 * 128)
        wp = torch.zeros((npad, kpad), dtype=torch.float32, device=w.device)
        wp[:cout, :cin] = w
        wp[cout:, 0] = 1.0                      # padded rows: finite v / ||v||, scale 0
        self.in_channels, self.out_channels = kpad, npad
        self.n_valid = cout
        self.conv = self._P()
        self.conv.weight_v = wp.reshape(npad, kpad, 1, 1).contiguous()
        self.conv.weight_g = wp.norm(dim=1).contiguous()
        b = torch.zeros(npad, dtype=torch.float32, device=w.device)
        b[:cout] = bias.detach().float()
        self.conv.bias = b
        g = torch.zeros(npad, dtype=torch.float32, device=w.device)
        g[:cout] = 1.0
        self.gamma = g
        self.beta = torch.zeros(npad, dtype=torch.float32, device=w.device)


def _linear(bank, ad, x: torch.Tensor, *, act: int = 0, f32: bool = False) -> torch.Tensor:
    """One linear layer on ``x`` [B, 1, 1, K] bf16: bf16 output with activation ``act`` (2 = LeakyReLU), or the raw fp32
    output [B, 1, 1, N]."""
    cw = bank.get(ad)
    if f32:
        return ops.conv_fwd_f32([x], cw.w, cw.scale, cw.shift, ksize=1, layout="nhwc")
    return ops.conv_fwd([x], cw.w, cw.scale, cw.shift, ksize=1, want_raw=False, want_elu=True, act=act)[1]


class _Prepared:
    """Packed weights of a set of linear layers, re-prepared when a parameter changes (version / storage /
    ``bump_weight_epoch()``).  The adapters and the bank's packed buffers are allocated ONCE per device and layer
    shapes: a change of the parameters is copied into the same adapter tensors and prepared into the same packed
    buffers, so CUDA graphs that have those addresses baked in (``PoseSequenceSampler``) stay valid; only a device move
    or a change of a layer's shape builds a new bank (``generation`` counts those)."""

    def __init__(self):
        self.key = None
        self.bank = None
        self.ads = None
        self.generation = 0
        self._flat = None

    @staticmethod
    def _tensors(ad):
        return (ad.conv.weight_v, ad.conv.weight_g, ad.conv.bias, ad.gamma, ad.beta)

    def get(self, params: List[torch.Tensor], build):
        key = (_WEIGHT_EPOCH[0],) + tuple((p.data_ptr(), p._version) for p in params)
        if key != self.key:
            if params[0].device.type != "cuda":
                raise RuntimeError("vunet_b200 has no CPU path: move the module to a CUDA device")
            ads = build()
            flat = list(_flatten(ads))
            same = self.bank is not None and self._flat is not None and len(flat) == len(self._flat) and all(
                a.shape == b.shape and a.device == b.device
                for o, n in zip(self._flat, flat) for a, b in zip(self._tensors(o), self._tensors(n)))
            if same:
                for o, n in zip(self._flat, flat):
                    for a, b in zip(self._tensors(o), self._tensors(n)):
                        a.copy_(b)
                self.bank._key = None            # same buffers, new contents: force the preparation pass
            else:
                self.ads, self._flat = ads, flat
                self.bank = WeightBank([[a] for a in flat])
                self.generation += 1
            self.bank.refresh()
            self.key = key
        return self.bank, self.ads


def _flatten(x):
    if isinstance(x, (list, tuple)):
        for y in x:
            yield from _flatten(y)
    elif x is not None:
        yield x


class UnsupervisedTransformer2(nn.Module):
    """models/flow/simple_flow.py:136-180.  ``reverse(z)`` / ``sample(shape, device)`` run on the GPU kernels; the forward
    (density) direction is only needed to TRAIN the flow and raises."""

    def __init__(self, **kwargs):
        super().__init__()
        in_channels = kwargs["flow_in_channels"]
        mid_channels = kwargs["flow_mid_channels"]
        hidden_depth = kwargs["flow_hidden_depth"]
        n_flows = kwargs["n_flows"]
        self.in_channels = in_channels
        if in_channels % 128:
            raise NotImplementedError("vunet_b200: the flow kernels take an even split into multiples of 64 channels "
                                      "(the reference's 'uneven dims' variant is not used: dim_hidden_b = 1024)")
        self.flow = UnconditionalFlow2(in_channels=in_channels, hidden_dim=mid_channels, hidden_depth=hidden_depth,
                                       n_flows=n_flows)
        self._prep = _Prepared()

    def _prepared(self):
        # the Parameter objects are fixed after construction (load_state_dict / .to() / optimisers write into them):
        # walking the 300-module tree costs 1.3 ms per call, the cached list 0.1 ms
        params = self.__dict__.get("_vub_params")
        if params is None:
            params = [p for p in self.parameters()]
            object.__setattr__(self, "_vub_params", params)

        def build():
            ads = []
            for blk in self.flow.sub_layers:
                per = []
                for i in range(2):
                    per.append(([_LinearAdapter(l.weight, l.bias) for l in blk.coupling.s[i].linears()],
                                [_LinearAdapter(l.weight, l.bias) for l in blk.coupling.t[i].linears()]))
                ads.append(per)
            return ads
        return self._prep.get(params, build)

    def _mlp(self, bank, ads, x_bf: torch.Tensor) -> torch.Tensor:
        h = x_bf
        for ad in ads[:-1]:
            h = _linear(bank, ad, h, act=2)
        return _linear(bank, ads[-1], h, f32=True)        # raw fp32 (the s-net's Tanh is applied by the coupling kernel)

    def _s_and_t(self, bank, ads_s, ads_t, x_bf: torch.Tensor):
        """The scale and the translation net of a coupling read the same half and are independent: the t-net runs on
        a second stream (parallel branches of the captured graph), which halves the chain of dependent launches —
        each layer is ~10 us of latency on at most 128 CTAs.  VUB_FLOW_STREAMS=1 keeps one stream."""
        import os
        if os.environ.get("VUB_FLOW_STREAMS", "2") == "1":
            return self._mlp(bank, ads_s, x_bf), self._mlp(bank, ads_t, x_bf)
        main = torch.cuda.current_stream()
        side = _FLOW_STREAMS.get(main.device)
        if side is None:
            side = _FLOW_STREAMS[main.device] = torch.cuda.Stream(device=main.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            t = self._mlp(bank, ads_t, x_bf)
        s_raw = self._mlp(bank, ads_s, x_bf)
        main.wait_stream(side)
        t.record_stream(main)
        return s_raw, t

    def reverse(self, out: torch.Tensor) -> torch.Tensor:
        """``flow(out, reverse=True)``: [B, D] or [B, D, 1, 1] -> [B, D, 1, 1] like the reference."""
        if not out.is_cuda:
            raise RuntimeError("vunet_b200 has no CPU path: inputs must live on a CUDA device")
        x = out.reshape(out.shape[0], -1).float().contiguous()
        B, D = x.shape
        if D != self.in_channels:
            raise ValueError(f"flow.reverse: expected {self.in_channels} channels, got {D}")
        bank, ads = self._prepared()
        lib, st = N.lib(), ops._stream
        d2 = D // 2
        state = x
        prev_norm = None   # ActNorm of the block processed before (its reverse is fused into the next gather)
        for bi in reversed(range(len(self.flow.sub_layers))):
            blk = self.flow.sub_layers[bi]
            # shuffle^-1 (+ the previous block's actnorm^-1): new state and the bf16 copy of its first half
            nxt = torch.empty_like(state)
            xa_bf = torch.empty((B, 1, 1, d2), dtype=torch.bfloat16, device=x.device)
            src = blk.shuffle.backward_shuffle_idx           # int64 buffer: the kernel takes an int32 copy, cached
            cached = getattr(blk, "_vub_idx32", None)
            if cached is None or cached[0] != (src.data_ptr(), src._version, x.device):
                cached = ((src.data_ptr(), src._version, x.device), src.to(device=x.device, dtype=torch.int32))
                object.__setattr__(blk, "_vub_idx32", cached)
            idx = cached[1]
            N.check(ops._launch("gather_affine", lib.vub_gather_affine, state.data_ptr(), idx.data_ptr(),
                                None if prev_norm is None else prev_norm.scale.data_ptr(),
                                None if prev_norm is None else prev_norm.loc.data_ptr(), B, D, nxt.data_ptr(),
                                xa_bf.data_ptr(), 0, d2, st()), "vub_gather_affine")
            state = nxt
            # coupling i = 1: keep = second half, apply = first half
            s_raw, t = self._s_and_t(bank, ads[bi][1][0], ads[bi][1][1], xa_bf)
            xb_bf = torch.empty_like(xa_bf)
            N.check(ops._launch("couple_inv", lib.vub_couple_inv, state.data_ptr(), s_raw.data_ptr(), t.data_ptr(), B, D, d2,
                                d2, state.data_ptr(), xb_bf.data_ptr(), st()), "vub_couple_inv")
            # coupling i = 0 after the swap: apply = (updated) second half, keep = first half; the result
            # [second half | updated first half] is written back in swapped order
            s_raw, t = self._s_and_t(bank, ads[bi][0][0], ads[bi][0][1], xb_bf)
            N.check(ops._launch("couple_inv", lib.vub_couple_inv, state.data_ptr(), s_raw.data_ptr(), t.data_ptr(), B, D, d2,
                                0, state.data_ptr(), None, st()), "vub_couple_inv")
            # state now holds [A'' | B''] in its original column order; the coupling's output order is [B'' | A'']
            state = _swap_halves(state, d2)
            prev_norm = blk.norm_layer
        res = torch.empty_like(state)
        N.check(ops._launch("gather_affine", lib.vub_gather_affine, state.data_ptr(), None, prev_norm.scale.data_ptr(),
                            prev_norm.loc.data_ptr(), B, D, res.data_ptr(), None, 0, 0, st()), "vub_gather_affine")
        return res.reshape(B, D, 1, 1)

    def sample(self, shape, device="cuda"):
        z_tilde = torch.randn(shape).to(device)
        return self.reverse(z_tilde).squeeze(dim=-1).squeeze(dim=-1)

    def forward(self, input, reverse=False, train=False):
        if reverse:
            return self.reverse(input)
        raise NotImplementedError("vunet_b200: the density direction of the flow (training) is outside the sampling path")

    def get_last_layer(self):
        return getattr(self.flow.sub_layers[-1].coupling.t[-1].main[-1], "weight")


def _swap_halves(state: torch.Tensor, d2: int) -> torch.Tensor:
    """[first | second] -> [second | first] on a [B, 2*d2] fp32 row-major tensor (one gather launch, no arithmetic)."""
    key = (state.device, d2)
    idx = _SWAP_IDX.get(key)
    if idx is None:
        idx = torch.cat([torch.arange(d2, 2 * d2), torch.arange(0, d2)]).to(device=state.device, dtype=torch.int32)
        _SWAP_IDX[key] = idx
    out = torch.empty_like(state)
    N.check(ops._launch("gather_affine", N.lib().vub_gather_affine, state.data_ptr(), idx.data_ptr(), None, None,
                        state.shape[0], 2 * d2, out.data_ptr(), None, 0, 0, ops._stream()), "vub_gather_affine")
    return out


_SWAP_IDX: dict = {}
_FLOW_STREAMS: dict = {}  # one side stream per device for the translation nets


# ------------------------------------------------------------------------------------------------
# models/pose_behavior_rnn.py
# ------------------------------------------------------------------------------------------------
class BEncoder(nn.Module):
    """models/pose_behavior_rnn.py:125-210 — the LSTM encoder that infers the behavior code b from a pose sequence
    (behavior transfer / reconstruction; the sampling path draws b from the flow instead)."""

    def __init__(self, n_in, n_layers, dim_hidden, use_linear, dim_linear, ib=False):
        super().__init__()
        if ib:
            raise NotImplementedError("vunet_b200: information_bottleneck is not used by config/behavior_net.yaml")
        self.rnn = nn.LSTM(input_size=n_in, hidden_size=dim_hidden, num_layers=n_layers, batch_first=True)
        self.n_layer = n_layers
        self.dim_hidden = dim_hidden
        self.ib = ib
        self.use_linear = use_linear
        self.linear = nn.Linear(self.dim_hidden, dim_linear) if self.use_linear else None

        self._prep = _Prepared()
        self.hidden = None

    def init_hidden(self, bs, device):
        self.hidden = (torch.zeros((self.n_layer, bs, self.dim_hidden), device=device),
                       torch.zeros((self.n_layer, bs, self.dim_hidden), device=device))

    def set_hidden(self, bs, device, hidden=None, cell=None):
        if hidden is None and cell is None:
            self.init_hidden(bs, device)
        elif hidden is None:
            self.hidden = (torch.zeros_like(cell), cell)
        elif cell is None:
            self.hidden = (hidden, torch.zeros_like(hidden))
        else:
            self.hidden = (hidden, cell)

    def forward(self, x, sample=False):
        """``out, self.hidden = self.rnn(x, self.hidden)``; returns ``pre = self.hidden[0][-1]`` (the behavior code b):
        one gate GEMM on the concatenated [h | x_t] row + one cell launch per time step."""
        if not x.is_cuda:
            raise RuntimeError("vunet_b200 has no CPU path: inputs must live on a CUDA device")
        if self.n_layer != 1:
            raise NotImplementedError("vunet_b200: BEncoder with one LSTM layer (ResidualBehaviorNet builds it with 1)")
        rnn = self.rnn
        n = rnn.input_size

        def build():
            xpad = (n + 63) // 64 * 64
            wx = torch.zeros((rnn.weight_ih_l0.shape[0], xpad), dtype=torch.float32, device=rnn.weight_ih_l0.device)
            wx[:, :n] = rnn.weight_ih_l0.detach()
            return [_LinearAdapter(torch.cat([rnn.weight_hh_l0.detach(), wx], dim=1),
                                   rnn.bias_ih_l0.detach() + rnn.bias_hh_l0.detach())]
        bank, (ad_g,) = self._prep.get([rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0, rnn.bias_hh_l0], build)
        lib, st = N.lib(), ops._stream
        B, T, H = x.shape[0], x.shape[1], self.dim_hidden
        xpad = (n + 63) // 64 * 64
        xs = x.float().contiguous()
        if self.hidden is None or self.hidden[0].shape[1] != B or self.hidden[0].device != x.device:
            self.init_hidden(B, x.device)
        h0, c0 = self.hidden[0][-1].float().contiguous(), self.hidden[1][-1].float().contiguous()
        c = c0.clone()
        h32 = h0.clone()
        cat = torch.zeros((B, 1, 1, H + xpad), dtype=torch.bfloat16, device=x.device)
        cat.view(B, H + xpad)[:, :H] = h0.to(torch.bfloat16)      # initial hidden state (zeros after init_hidden)
        for t in range(T):
            N.check(ops._launch("lstm_input", lib.vub_lstm_input, xs.data_ptr() + t * n * 4, T * n, B, n, xpad, cat.data_ptr(),
                                H + xpad, H, st()), "vub_lstm_input")
            gates = _linear(bank, ad_g, cat, f32=True)
            N.check(ops._launch("lstm_cell", lib.vub_lstm_cell, gates.data_ptr(), c.data_ptr(), B, H, None, cat.data_ptr(),
                                H + xpad, h32.data_ptr(), st()), "vub_lstm_cell")
        self.hidden = (h32.unsqueeze(0), c.unsqueeze(0))
        return h32


class ResidualRNNDecoder(nn.Module):
    """models/pose_behavior_rnn.py:463-535 (LSTM variant)."""

    def __init__(self, n_in_out, n_hidden, rnn_type="lstm", use_nin=False):
        super().__init__()
        if rnn_type != "lstm" or use_nin:
            raise NotImplementedError("vunet_b200: decoder_arch 'lstm' without linear_in_decoder (config/behavior_net.yaml)")
        self.n_in_out = n_in_out
        self.n_hidden = n_hidden
        self.rnn_type = rnn_type
        self.use_nin = use_nin
        self.rnn = nn.LSTMCell(self.n_in_out, self.n_hidden)
        self.n_out = nn.Linear(self.n_hidden, self.n_in_out)


class ResidualBehaviorNet(nn.Module):
    """models/pose_behavior_rnn.py:538-626.  ``generate_seq(b, x_pose, len, start_frame)`` -> (xs, cs, zs, b)."""

    def __init__(self, n_kps, **kwargs):
        super().__init__()
        self.dec_type = kwargs["decoder_arch"] if "decoder_arch" in kwargs else "lstm"
        self.use_nin_dec = kwargs["linear_in_decoder"] if "linear_in_decoder" in kwargs else False
        self.ib = kwargs["information_bottleneck"] if "information_bottleneck" in kwargs else False
        self.dim_hidden_b = kwargs["dim_hidden_b"]
        if self.dim_hidden_b % 64:
            raise NotImplementedError("vunet_b200: dim_hidden_b must be a multiple of 64")
        self.n_kps = n_kps
        self.b_enc = BEncoder(n_kps, 1, self.dim_hidden_b, use_linear=False, dim_linear=1, ib=self.ib)
        self.decoder = ResidualRNNDecoder(n_in_out=n_kps, n_hidden=self.dim_hidden_b, rnn_type=self.dec_type,
                                          use_nin=self.use_nin_dec)
        self._prep = _Prepared()

    def _prepared(self):
        rnn, n_out = self.decoder.rnn, self.decoder.n_out

        def build():
            xpad = (self.n_kps + 63) // 64 * 64
            wx = torch.zeros((rnn.weight_ih.shape[0], xpad), dtype=torch.float32, device=rnn.weight_ih.device)
            wx[:, :self.n_kps] = rnn.weight_ih.detach()
            # gates = [h | x_pad] @ [W_hh | W_ih_pad]^T + (b_ih + b_hh): one GEMM per time step
            gates = _LinearAdapter(torch.cat([rnn.weight_hh.detach(), wx], dim=1), rnn.bias_ih.detach() + rnn.bias_hh.detach())
            return [gates, _LinearAdapter(n_out.weight, n_out.bias)]
        return self._prep.get([rnn.weight_ih, rnn.weight_hh, rnn.bias_ih, rnn.bias_hh, n_out.weight, n_out.bias], build)

    def generate_seq(self, b, x_pose, len, start_frame):
        if not (b.is_cuda and x_pose.is_cuda):
            raise RuntimeError("vunet_b200 has no CPU path: inputs must live on a CUDA device")
        bank, (ad_g, ad_o) = self._prepared()
        lib, st = N.lib(), ops._stream
        B, H, n = b.shape[0], self.dim_hidden_b, self.n_kps
        xpad = (n + 63) // 64 * 64
        dev = b.device
        b32 = b.reshape(B, H).float().contiguous()
        x0 = x_pose[:, start_frame].reshape(B, n).float().contiguous()
        c = torch.empty((B, H), dtype=torch.float32, device=dev)
        x = torch.empty((B, n), dtype=torch.float32, device=dev)
        cat = torch.empty((B, 1, 1, H + xpad), dtype=torch.bfloat16, device=dev)
        h_lin = torch.empty((B, 1, 1, H), dtype=torch.bfloat16, device=dev)
        xs = torch.empty((B, len, n), dtype=torch.float32, device=dev)
        cs = torch.empty((B, len, n), dtype=torch.float32, device=dev)
        N.check(ops._launch("decoder_init", lib.vub_decoder_init, b32.data_ptr(), x0.data_ptr(), B, H, n, xpad, c.data_ptr(),
                            x.data_ptr(), cat.data_ptr(), H + xpad, st()), "vub_decoder_init")
        for t in range(len):
            gates = _linear(bank, ad_g, cat, f32=True)            # [B, 1, 1, 4H] fp32
            N.check(ops._launch("lstm_cell", lib.vub_lstm_cell, gates.data_ptr(), c.data_ptr(), B, H, h_lin.data_ptr(),
                                cat.data_ptr(), H + xpad, None, st()), "vub_lstm_cell")
            v = _linear(bank, ad_o, h_lin, f32=True)              # [B, 1, 1, npad] fp32
            N.check(ops._launch("decoder_step", lib.vub_decoder_step, x.data_ptr(), v.data_ptr(), B, n, v.shape[-1], xpad, t,
                                len, xs.data_ptr(), cs.data_ptr(), cat.data_ptr(), H + xpad, H, st()), "vub_decoder_step")
        return xs, cs, [], b

    def infer_b(self, s, sample):
        """models/pose_behavior_rnn.py:587-601: the behavior code of a sequence, b = last hidden state of the LSTM encoder."""
        self.b_enc.init_hidden(s.shape[0], device=s.device)
        return self.b_enc(s, sample)

    def forward(self, x1, x2, len, start_frame=0, sample=False):
        """models/pose_behavior_rnn.py:574-585: behavior of ``x1`` re-enacted from the start pose of ``x2``:
        returns (xs, cs, zs_gen, b)."""
        b = self.infer_b(x1, sample)
        xs, cs, zs_gen, _ = self.generate_seq(b, x2, len, start_frame=start_frame)
        return xs, cs, zs_gen, b


class PoseSequenceSampler:
    """Noise -> 3-D pose sequences in ONE CUDA graph (lib/logging.py:280-300 done for a whole batch):
    ``xs = sampler(z, x_start)`` == ``net.generate_seq(flow.reverse(z).squeeze(), x_start[:, None], length, 0)[0]``."""

    def __init__(self, flow: UnsupervisedTransformer2, net: ResidualBehaviorNet, batch: int, length: int,
                 use_graph: bool = True):
        self.flow, self.net, self.batch, self.length = flow, net, batch, length
        dev = next(net.parameters()).device
        self.z = torch.zeros((batch, flow.in_channels), dtype=torch.float32, device=dev)
        self.x_start = torch.zeros((batch, 1, net.n_kps), dtype=torch.float32, device=dev)
        self.graph = None
        self._use_graph = use_graph
        self._dev = dev
        self._capture()

    def _generations(self):
        with torch.no_grad():
            self.flow._prepared()
            self.net._prepared()
        return (self.flow._prep.generation, self.net._prep.generation)

    def _capture(self) -> None:
        dev = self._dev
        self._gens = self._generations()         # weight preparation happens outside the graph
        with torch.no_grad():
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    before = ops.LAUNCHES
                    self.out = self._step()
                    self.launches_per_call = ops.LAUNCHES - before
            torch.cuda.current_stream(dev).wait_stream(side)
            if self._use_graph:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self.out = self._step()

    def _step(self):
        b = self.flow.reverse(self.z).reshape(self.batch, -1)
        return self.net.generate_seq(b, self.x_start, self.length, 0)[0]

    def refresh_weights(self) -> None:
        """Re-prepares the packed weights if a parameter of the flow or the decoder changed (``run`` does the same
        check itself).  The packed buffers keep their addresses across parameter updates (``_Prepared``), so the
        captured graph normally stays valid; after a device move or a layer-shape change it is captured again."""
        if self._generations() != self._gens:
            torch.cuda.current_stream(self._dev).synchronize()
            self._capture()

    def run(self) -> torch.Tensor:
        self.refresh_weights()
        if self.graph is not None:
            self.graph.replay()
        else:
            with torch.no_grad():
                self.out = self._step()
        return self.out

    def __call__(self, z: torch.Tensor, x_start: torch.Tensor) -> torch.Tensor:
        self.z.copy_(z.reshape(self.z.shape), non_blocking=True)
        self.x_start.copy_(x_start.reshape(self.x_start.shape), non_blocking=True)
        return self.run()
