"""CPU oracle for behavior-net sampling — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Functional fp32 restatement (plain torch CPU ops on reference-layout state dicts) of the sampling path of BASELINE.json
configs[3]: ``b = flow.reverse(z)`` through ``UnsupervisedTransformer2`` (models/flow/simple_flow.py:136-180 ->
models/flow/blocks.py:95-128 ``UnconditionalFlow2``, 531-559 ``UnconditionalFlatDoubleCouplingFlowBlock2``, 276-319
``DoubleVectorCouplingBlock2``, 692-704 ``Shuffle``; lib/modules.py:236-257 ``BasicFullyConnectedNet``, 260-330
``ActNorm``) and ``ResidualBehaviorNet.generate_seq`` (models/pose_behavior_rnn.py:603-626) with ``ResidualRNNDecoder``
(463-506: ``nn.LSTMCell`` + ``n_out``, ``x_{t+1} = x_t + n_out(h_t)``), plus ``infer_b`` (587-601, the LSTM encoder of
``BEncoder`` 125-201) for behavior transfer.

Parity pinning: ``oracle/gen_golden_behavior.py`` imports the UNMODIFIED reference classes (they import with torch
alone), loads deterministic hashed parameters and checks this restatement against them (``tests/golden/behavior_tiny.pt``).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

import vunet_oracle as O

SD = Dict[str, torch.Tensor]


def mlp(sd: SD, prefix: str, x: torch.Tensor, depth: int, use_tanh: bool) -> torch.Tensor:
    """BasicFullyConnectedNet.forward — lib/modules.py:236-257: Linear, LeakyReLU, depth x (Linear, LeakyReLU), Linear
    [, Tanh]; the Linear layers sit at the even indices of ``main``."""
    n_lin = depth + 2
    for k in range(n_lin):
        x = F.linear(x, sd[f"{prefix}.main.{2 * k}.weight"], sd[f"{prefix}.main.{2 * k}.bias"])
        if k + 1 < n_lin:
            x = F.leaky_relu(x)
    return torch.tanh(x) if use_tanh else x


def coupling_reverse(sd: SD, prefix: str, x: torch.Tensor, depth: int) -> torch.Tensor:
    """DoubleVectorCouplingBlock2.forward(reverse=True) — models/flow/blocks.py:309-319."""
    for i in reversed(range(2)):
        if i % 2 == 0:
            x = torch.cat(torch.chunk(x, 2, dim=1)[::-1], dim=1)
        xa, xk = torch.chunk(x, 2, dim=1)
        xk = (xk - mlp(sd, f"{prefix}.t.{i}", xa, depth, False)) * mlp(sd, f"{prefix}.s.{i}", xa, depth, True).neg().exp()
        x = torch.cat((xa, xk), dim=1)
    return x


def flow_reverse(sd: SD, z: torch.Tensor, n_flows: int, depth: int, prefix: str = "flow.sub_layers") -> torch.Tensor:
    """UnsupervisedTransformer2.reverse — simple_flow.py:171-174 -> UnconditionalFlow2 (blocks.py:122-125): the blocks in
    reverse order, each: shuffle^-1, coupling^-1, actnorm^-1 (blocks.py:551-556).  z, result: [B, D]."""
    h = z
    for i in reversed(range(n_flows)):
        p = f"{prefix}.{i}"
        h = h[:, sd[f"{p}.shuffle.backward_shuffle_idx"]]
        h = coupling_reverse(sd, f"{p}.coupling", h, depth)
        h = h / sd[f"{p}.norm_layer.scale"].reshape(1, -1) - sd[f"{p}.norm_layer.loc"].reshape(1, -1)
    return h


def generate_seq(sd: SD, b: torch.Tensor, x_start: torch.Tensor, length: int, prefix: str = "decoder"):
    """ResidualBehaviorNet.generate_seq — pose_behavior_rnn.py:603-626 with the LSTM decoder: hidden = cell = b; per step
    (h, c) = LSTMCell(x, (h, c)); x <- x + n_out(h).  Returns (xs [B, len, n], cs [B, len, n])."""
    h, c = b, b
    x = x_start
    xs, cs = [], []
    for _ in range(length):
        gates = F.linear(x, sd[f"{prefix}.rnn.weight_ih"], sd[f"{prefix}.rnn.bias_ih"]) + \
            F.linear(h, sd[f"{prefix}.rnn.weight_hh"], sd[f"{prefix}.rnn.bias_hh"])
        i, f, g, o = gates.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        cs.append(x)
        x = x + F.linear(h, sd[f"{prefix}.n_out.weight"], sd[f"{prefix}.n_out.bias"])
        xs.append(x)
    return torch.stack(xs, dim=1), torch.stack(cs, dim=1)


def infer_b(sd: SD, s: torch.Tensor, prefix: str = "b_enc.rnn") -> torch.Tensor:
    """ResidualBehaviorNet.infer_b — pose_behavior_rnn.py:587-601 -> BEncoder.forward (177-201, ib = False): a one-layer
    nn.LSTM from zero state over the sequence s [B, T, n]; returns the last hidden state."""
    B = s.shape[0]
    hdim = sd[f"{prefix}.weight_hh_l0"].shape[1]
    h = torch.zeros(B, hdim)
    c = torch.zeros(B, hdim)
    for t in range(s.shape[1]):
        gates = F.linear(s[:, t], sd[f"{prefix}.weight_ih_l0"], sd[f"{prefix}.bias_ih_l0"]) + \
            F.linear(h, sd[f"{prefix}.weight_hh_l0"], sd[f"{prefix}.bias_hh_l0"])
        i, f, g, o = gates.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
    return h


# ---- deterministic parameters ------------------------------------------------------------------------
def det_flow_state_dict(shapes: Dict[str, Tuple[int, ...]]) -> SD:
    """Hashed parameters for an UnsupervisedTransformer2 state dict: Linear weights / biases ~ U(+-1/sqrt(fan_in)) (the
    last layer of every net scaled by 0.5 so that exp(+-s) stays moderate over 15 blocks), ActNorm loc ~ U(+-0.1),
    scale ~ U(0.8, 1.25), shuffles = the permutation sorting a hashed vector."""
    sd: SD = {}
    for name, shape in shapes.items():
        if name.endswith("initialized"):
            sd[name] = torch.tensor(1, dtype=torch.uint8)
        elif name.endswith("forward_shuffle_idx"):
            sd[name] = torch.argsort(O.det_uniform(name, shape)).to(torch.int64)
        elif name.endswith("backward_shuffle_idx"):
            continue
        elif name.endswith("norm_layer.loc"):
            sd[name] = O.det_uniform(name, shape, -0.1, 0.1)
        elif name.endswith("norm_layer.scale"):
            sd[name] = O.det_uniform(name, shape, 0.8, 1.25)
        elif name.endswith("weight"):
            bound = 1.0 / float(np.sqrt(shape[1]))
            sd[name] = O.det_uniform(name, shape, -bound, bound)
        elif name.endswith("bias"):
            w = shapes[name[:-4] + "weight"]
            bound = 1.0 / float(np.sqrt(w[1]))
            sd[name] = O.det_uniform(name, shape, -bound, bound)
        else:
            raise KeyError(name)
    for name in shapes:
        if name.endswith("backward_shuffle_idx"):
            sd[name] = torch.argsort(sd[name.replace("backward", "forward")])
    return {k: sd[k] for k in shapes}


def det_net_state_dict(shapes: Dict[str, Tuple[int, ...]]) -> SD:
    """Hashed parameters for a ResidualBehaviorNet state dict (torch's own LSTM / Linear init range +-1/sqrt(hidden))."""
    hidden = shapes["decoder.rnn.weight_hh"][1]
    bound = 1.0 / float(np.sqrt(hidden))
    return {k: O.det_uniform("bnet." + k, s, -bound, bound) for k, s in shapes.items()}
