"""Time the graph-captured behavior-net sampler (batch 256, 50 steps): python tools/behavior_perf.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import behavior_oracle as BO
import bdvs_b200 as B

flow = B.UnsupervisedTransformer2(flow_in_channels=1024, n_flows=15, flow_hidden_depth=2, flow_mid_channels=2048)
net = B.ResidualBehaviorNet(n_kps=51, dim_hidden_b=1024, decoder_arch="lstm", linear_in_decoder=False)
flow.load_state_dict(BO.det_flow_state_dict({k: tuple(v.shape) for k, v in flow.state_dict().items()}))
net.load_state_dict(BO.det_net_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}))
flow, net = flow.cuda().eval(), net.cuda().eval()
z = torch.randn(256, 1024, device="cuda")
x0 = torch.rand(256, 1, 51, device="cuda") * 2 - 1


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


smp = B.PoseSequenceSampler(flow, net, 256, 50)
smp(z, x0)
print(f"sampler (graph, {smp.launches_per_call} launches): {timeit(smp.run):.3f} ms per batch of 256")
# the two halves separately, each in its own graph
g1 = torch.cuda.CUDAGraph()
with torch.no_grad():
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        flow.reverse(z)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g1):
        b = flow.reverse(z)
    g2 = torch.cuda.CUDAGraph()
    bb = b.reshape(256, 1024).clone()
    with torch.cuda.stream(s):
        net.generate_seq(bb, x0, 50, 0)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g2):
        xs = net.generate_seq(bb, x0, 50, 0)[0]
print(f"flow.reverse alone: {timeit(g1.replay):.3f} ms; generate_seq alone: {timeit(g2.replay):.3f} ms")
