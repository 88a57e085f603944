"""Eager behavior-net sampling steps for an ncu capture (tools/r02_ncu_behavior.sh): two warm-up passes, then ONE pass of
the 502 launches of noise -> 256 pose sequences of 50 frames (no graph, one stream so the launch order is the chain)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
os.environ["VUB_FLOW_STREAMS"] = "1"
import torch
import behavior_oracle as BO
import bdvs_b200 as B
from bdvs_b200 import ops

flow = B.UnsupervisedTransformer2(flow_in_channels=1024, n_flows=15, flow_hidden_depth=2, flow_mid_channels=2048)
net = B.ResidualBehaviorNet(n_kps=51, dim_hidden_b=1024, decoder_arch="lstm", linear_in_decoder=False)
flow.load_state_dict(BO.det_flow_state_dict({k: tuple(v.shape) for k, v in flow.state_dict().items()}))
net.load_state_dict(BO.det_net_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}))
flow, net = flow.cuda().eval(), net.cuda().eval()
z = torch.randn(256, 1024, device="cuda")
x0 = torch.rand(256, 1, 51, device="cuda") * 2 - 1
with torch.no_grad():
    flow._prepared()
    net._prepared()
    torch.cuda.synchronize()
    for i in range(3):
        before = ops.LAUNCHES
        xs = net.generate_seq(flow.reverse(z).reshape(256, 1024), x0, 50, 0)[0]
        torch.cuda.synchronize()
        print("pass", i, "launches", ops.LAUNCHES - before, flush=True)
print("finite", bool(torch.isfinite(xs).all()))
