"""Generates ``tests/golden/behavior_tiny.pt`` from the UNMODIFIED reference classes ``UnsupervisedTransformer2``
(models/flow/simple_flow.py) and ``ResidualBehaviorNet`` (models/pose_behavior_rnn.py) — test infrastructure, build
container only.  Deterministic hashed parameters, the reference's ``flow.reverse(z)`` and ``generate_seq(b, x, len, 0)``;
the CPU oracle must reproduce both before the fixture is written.

Usage:  python oracle/gen_golden_behavior.py [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import behavior_oracle as BO  # noqa: E402
import vunet_oracle as O  # noqa: E402

CFG = dict(dim_hidden_b=128, flow_mid_channels_factor=2, n_flows=3, flow_hidden_depth=2, n_kps=51, batch=4, length=6)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    from models.flow.simple_flow import UnsupervisedTransformer2   # the reference's own classes
    from models.pose_behavior_rnn import ResidualBehaviorNet
    torch.manual_seed(0)
    D = CFG["dim_hidden_b"]
    flow = UnsupervisedTransformer2(flow_in_channels=D, n_flows=CFG["n_flows"], flow_hidden_depth=CFG["flow_hidden_depth"],
                                    flow_mid_channels=D * CFG["flow_mid_channels_factor"])
    net = ResidualBehaviorNet(n_kps=CFG["n_kps"], dim_hidden_b=D, decoder_arch="lstm", linear_in_decoder=False)
    fshapes = {k: tuple(v.shape) for k, v in flow.state_dict().items()}
    nshapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    fsd, nsd = BO.det_flow_state_dict(fshapes), BO.det_net_state_dict(nshapes)
    flow.load_state_dict(fsd, strict=True)
    net.load_state_dict(nsd, strict=True)
    B, T, n = CFG["batch"], CFG["length"], CFG["n_kps"]
    z = O.det_normal("bz", (B, D))
    x_pose = O.det_uniform("bx", (B, T, n))
    with torch.no_grad():
        b = flow.reverse(z)
        assert tuple(b.shape) == (B, D, 1, 1)
        b2 = b.squeeze(-1).squeeze(-1)
        # the reference asks tensors for their CUDA device (b.get_device()); on CPU that is -1, which `.to(-1)` rejects:
        # set the decoder state the way set_hidden does (pose_behavior_rnn.py:520-535) and run its loop
        net.decoder.hidden = (b2, b2)
        xs, cs = [], []
        x = x_pose[:, 0]
        for _ in range(T):
            x, v = net.decoder(x)
            xs.append(x)
            cs.append(v)
        xs, cs = torch.stack(xs, dim=1), torch.stack(cs, dim=1)
        # behavior inference from a sequence (infer_b asks for s.get_device(): run BEncoder the way it does, on CPU)
        seq = O.det_uniform("bseq", (B, T + 3, n))
        net.b_enc.init_hidden(B, "cpu")
        b_inf = net.b_enc(seq, False)
        o_inf = BO.infer_b(nsd, seq)
        o_b = BO.flow_reverse(fsd, z, CFG["n_flows"], CFG["flow_hidden_depth"])
        o_xs, o_cs = BO.generate_seq(nsd, o_b, x_pose[:, 0], T)
    rel = lambda a, r: float((a - r).norm() / r.norm())
    errs = {"b": rel(o_b, b2), "xs": rel(o_xs, xs), "cs": rel(o_cs, cs), "infer_b": rel(o_inf, b_inf)}
    print("oracle vs reference:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < 1e-5, errs
    fixture = {"config": CFG, "flow_shapes": list(fshapes.items()), "net_shapes": list(nshapes.items()),
               "b": b2.clone(), "xs": xs.clone(), "cs": cs.clone(), "infer_b": b_inf.clone(), "oracle_vs_reference": errs,
               "torch_version": torch.__version__}
    path = os.path.join(args.out, "behavior_tiny.pt")
    torch.save(fixture, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB); |b| {float(b2.abs().mean()):.3f}, |xs| "
          f"{float(xs.abs().mean()):.3f}")


if __name__ == "__main__":
    main()
