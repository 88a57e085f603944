"""GPU (-m gpu): data-parallel training equivalence — 2-rank averaged gradients (flat-bucket all-reduce overlapped
with backward, one process per rank) == 1-process gradients on the concatenated batch.

With >= 2 devices the ranks own one GPU each and exchange over NCCL (``gpurun --gpus 2``).  On a 1-GPU box (the
driver's GPU tier) both ranks share cuda:0 and exchange through gloo — NCCL refuses two ranks on one device — which
still runs the whole product path: bucket layout, side-stream ordering, in-place averaging, ``param.grad`` views.
The model is the Market config (BASELINE configs[0], 128x128, 30 appearance channels), not a toy."""
import contextlib
import io
import os
import socket

import pytest
import torch

import vunet_oracle as O

pytestmark = pytest.mark.gpu

CFG = dict(n_latent_scales=2, conv_layer_type="l1", subpixel_upsampling=True, n_scales=0, nf_start=32, nf_max=128,
           spatial_size=128, bottleneck_factor=1, box_factor=1, n_channels_x=30)
S = CFG["spatial_size"]


def _build(device):
    import bdvs_b200 as B
    with contextlib.redirect_stdout(io.StringIO()):
        m = B.VunetAlter(**CFG)
    sd = O.det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()})
    m.load_state_dict(sd)
    return m.to(device).train()


def _grads(m, x, c, eps, target):
    import bdvs_b200 as B
    imgs, means, logstds, _ = m(x, c, eps=eps)
    loss = torch.mean((imgs - target) ** 2) + 1e-3 * B.compute_kl_with_prior(means, logstds)
    loss.backward()
    torch.cuda.synchronize()
    return {k: p.grad.detach().float().cpu().clone() for k, p in m.named_parameters() if p.grad is not None}


def _worker(rank, world, port, batch, out, shared_gpu):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", 0 if shared_gpu else rank)
    torch.cuda.set_device(dev)
    if shared_gpu:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from bdvs_b200.synthesis import shard_range
        m = _build(dev).enable_data_parallel()
        x, c, eps, _ = O.make_inputs(CFG, batch)
        target = O.det_uniform("target", (batch, 3, S, S))
        lo, hi = shard_range(batch, rank, world)
        g = _grads(m, x[lo:hi].to(dev), c[lo:hi].to(dev), [e[lo:hi].to(dev).contiguous() for e in eps],
                   target[lo:hi].to(dev))
        out[rank] = g
        assert m._vub_reducer.bytes_per_step > 0
    finally:
        dist.destroy_process_group()


def test_two_rank_gradients_equal_single_process_on_full_batch():
    import torch.multiprocessing as mp
    world, batch = 2, 4
    shared_gpu = torch.cuda.device_count() < 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, batch, out, shared_gpu), nprocs=world, join=True)
    m = _build(torch.device("cuda", 0))
    x, c, eps, _ = O.make_inputs(CFG, batch)
    target = O.det_uniform("target", (batch, 3, S, S))
    ref = _grads(m, x.cuda(), c.cuda(), [e.cuda() for e in eps], target.cuda())
    for rank in range(world):
        got = out[rank]
        assert sorted(got) == sorted(ref)
        num = sum(float((got[k] - ref[k]).pow(2).sum()) for k in ref)
        den = sum(float(ref[k].pow(2).sum()) for k in ref)
        assert (num / den) ** 0.5 < 5e-3, rank  # bf16 activations differ per shard only through summation order
    for k in ref:
        assert torch.equal(out[0][k], out[1][k]), k  # every rank holds the same averaged gradient
