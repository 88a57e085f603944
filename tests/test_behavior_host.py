"""CPU: behavior-net sampling (BASELINE.json configs[3]) — oracle restatement against the golden vector generated from the
UNMODIFIED reference classes (oracle/gen_golden_behavior.py), state_dict layout / init stream of the host mirror, loud
failure without a GPU."""
import os
import sys

import pytest
import torch

import behavior_oracle as BO
import vunet_oracle as O
from conftest import REFERENCE_ROOT

import bdvs_b200 as B


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_oracle_matches_reference_golden(golden):
    fx = golden("behavior_tiny")
    cfg = fx["config"]
    fsd = BO.det_flow_state_dict(dict(fx["flow_shapes"]))
    nsd = BO.det_net_state_dict(dict(fx["net_shapes"]))
    Bs, T, n, D = cfg["batch"], cfg["length"], cfg["n_kps"], cfg["dim_hidden_b"]
    z = O.det_normal("bz", (Bs, D))
    x_pose = O.det_uniform("bx", (Bs, T, n))
    b = BO.flow_reverse(fsd, z, cfg["n_flows"], cfg["flow_hidden_depth"])
    xs, cs = BO.generate_seq(nsd, b, x_pose[:, 0], T)
    assert rel(b, fx["b"]) < 1e-5 and rel(xs, fx["xs"]) < 1e-5 and rel(cs, fx["cs"]) < 1e-5
    assert rel(BO.infer_b(nsd, O.det_uniform("bseq", (Bs, T + 3, n))), fx["infer_b"]) < 1e-5
    # the flow is a bijection: the oracle's forward direction (restated here from blocks.py:540-550) inverts reverse
    h = b
    for i in range(cfg["n_flows"]):
        p = f"flow.sub_layers.{i}"
        h = fsd[f"{p}.norm_layer.scale"].reshape(1, -1) * (h + fsd[f"{p}.norm_layer.loc"].reshape(1, -1))
        for j in range(2):
            if j % 2 != 0:
                h = torch.cat(torch.chunk(h, 2, dim=1)[::-1], dim=1)
            xa, xk = torch.chunk(h, 2, dim=1)
            s = BO.mlp(fsd, f"{p}.coupling.s.{j}", xa, cfg["flow_hidden_depth"], True)
            h = torch.cat((xa, xk * s.exp() + BO.mlp(fsd, f"{p}.coupling.t.{j}", xa, cfg["flow_hidden_depth"], False)), dim=1)
        h = h[:, fsd[f"{p}.shuffle.forward_shuffle_idx"]]
    assert rel(h, z) < 1e-4


def make(cfg):
    flow = B.UnsupervisedTransformer2(flow_in_channels=cfg["dim_hidden_b"], n_flows=cfg["n_flows"],
                                      flow_hidden_depth=cfg["flow_hidden_depth"],
                                      flow_mid_channels=cfg["dim_hidden_b"] * cfg["flow_mid_channels_factor"])
    net = B.ResidualBehaviorNet(n_kps=cfg["n_kps"], dim_hidden_b=cfg["dim_hidden_b"], decoder_arch="lstm",
                                linear_in_decoder=False)
    return flow, net


def test_state_dict_layout_matches_reference_fixture(golden):
    fx = golden("behavior_tiny")
    flow, net = make(fx["config"])
    assert [(k, tuple(v.shape)) for k, v in flow.state_dict().items()] == [(k, tuple(s)) for k, s in fx["flow_shapes"]]
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == [(k, tuple(s)) for k, s in fx["net_shapes"]]
    fsd = BO.det_flow_state_dict(dict(fx["flow_shapes"]))
    nsd = BO.det_net_state_dict(dict(fx["net_shapes"]))
    flow.load_state_dict(fsd, strict=True)
    net.load_state_dict(nsd, strict=True)
    for k, v in flow.state_dict().items():
        assert torch.equal(v, fsd[k]), k


@pytest.mark.skipif(not os.path.isdir(REFERENCE_ROOT), reason="reference tree only exists in the build container")
def test_same_seed_gives_reference_initialisation(golden):
    sys.path.insert(0, REFERENCE_ROOT)
    from models.flow.simple_flow import UnsupervisedTransformer2 as RefFlow
    from models.pose_behavior_rnn import ResidualBehaviorNet as RefNet
    cfg = golden("behavior_tiny")["config"]
    torch.manual_seed(3)
    rf = RefFlow(flow_in_channels=128, n_flows=2, flow_hidden_depth=2, flow_mid_channels=256)
    rn = RefNet(n_kps=51, dim_hidden_b=128, decoder_arch="lstm", linear_in_decoder=False)
    torch.manual_seed(3)
    mf = B.UnsupervisedTransformer2(flow_in_channels=128, n_flows=2, flow_hidden_depth=2, flow_mid_channels=256)
    mn = B.ResidualBehaviorNet(n_kps=51, dim_hidden_b=128, decoder_arch="lstm", linear_in_decoder=False)
    for ref, mine in ((rf, mf), (rn, mn)):
        rsd, msd = ref.state_dict(), mine.state_dict()
        assert list(rsd.keys()) == list(msd.keys())
        for k in rsd:
            assert torch.equal(rsd[k], msd[k]), k   # same RNG stream (Linear / LSTM init, randperm of the shuffles)


def test_no_cpu_fallback_and_out_of_scope_entry_points(golden):
    flow, net = make(golden("behavior_tiny")["config"])
    with pytest.raises(RuntimeError, match="no CPU path"):
        flow.reverse(torch.zeros(2, 128))
    with pytest.raises(RuntimeError, match="no CPU path"):
        net.generate_seq(torch.zeros(2, 128), torch.zeros(2, 3, 51), 3, 0)
    with pytest.raises(NotImplementedError):
        flow(torch.zeros(2, 128))          # density direction: training only
    with pytest.raises(RuntimeError, match="no CPU path"):
        net.infer_b(torch.zeros(2, 3, 51), False)
