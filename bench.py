#!/usr/bin/env python
"""Benchmark of the VUnet hot path (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-train] [--no-cpu-baseline]
                    [--frames-total T]

Default workload = BASELINE.json configs[1]: VunetAlter Human3.6m settings (256x256, bottleneck_factor 2,
box_factor 2, 3 appearance channels), forward-only frame synthesis (``transfer``), bf16 tensor-core
compute, 64 frames per GPU per step; frames are sharded over ranks with no collective (weak scaling).
A "step" is one pass of the hot path over one 64-frame batch of synthetic frames.  ``--frames-total T`` is the
strong-scaling reading of the same config (T frames per step split over the GPUs, "scaling": "strong").

  value  frames/s, whole job, inputs resident in HBM, CUDA-graph replay timed with CUDA events
  e2e    frames/s through the public API (FrameSynthesizer.render_host) with pinned HOST buffers:
         host->device copies of x and c and the device->host copy of the frames inside the timed region
  roofline / kernels   per-kernel CUDA-event timing of one instrumented step (outside the timed region)
  cpu_baseline         the CPU oracle (oracle/vunet_oracle.py, a port of the reference's PyTorch path) on a
                       bounded sample, all host threads
``--impl reference`` times that CPU port alone and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

ARCH = dict(n_latent_scales=2, conv_layer_type="l1", subpixel_upsampling=True, n_scales=0, nf_start=32, nf_max=128)
H256 = dict(ARCH, spatial_size=256, bottleneck_factor=2, box_factor=2, n_channels_x=3)
WORKLOAD = "VunetAlter H3.6M-256 (256x256, bottleneck_factor=2, box=2, n_channels_x=3) transfer(x,c) forward-only"
FRAMES_PER_GPU = 64
# live convs of the reference graph, forward/transfer, per frame
GFLOP_PER_FRAME = 43.064
METRIC = "VUnet 256x256 frame synthesis throughput (fwd)"
UNIT = "frames/s"
# BASELINE.json configs[2]: DeepFashion settings, training step, batch 32 per GPU
D256 = dict(ARCH, spatial_size=256, bottleneck_factor=2, box_factor=2, n_channels_x=30)
TRAIN_WORKLOAD = ("VunetAlter DeepFashion-256 (256x256, bottleneck_factor=2, box=2, n_channels_x=30) training step: "
                  "forward + MSE/KL surrogate loss + backward + gradient all-reduce + fused Adam")
TRAIN_BATCH_PER_GPU = 32
# BASELINE.json configs[0] ("config 1"): Market settings, forward+backward fp32 on CPU, batch 8
M128 = dict(ARCH, spatial_size=128, bottleneck_factor=1, box_factor=1, n_channels_x=30)
CONFIG1_WORKLOAD = ("VunetAlter Market-128 (128x128, bottleneck_factor=1, box=1, n_channels_x=30) forward + MSE/KL "
                    "surrogate loss + backward, batch 8")
CONFIG1_BATCH = 8
# per-layer roofline terms of the H256 forward, per frame, bf16 NHWC with conv-level fusion:
# (GFLOP, MB of input + output + weights) per layer group; the blended lower bound is sum(max(t_TC, t_HBM)).
APP_A_LAYERS = [
    (4.832, 33.63), (4.832, 17.07), (4.832, 9.57), (4.832, 7.47), (4.832, 12.88), (4.832, 25.24), (2.416, 3.74),
    (2.416, 6.44), (1.208, 12.66), (1.208, 6.59), (1.208, 3.28), (1.208, 2.75), (1.208, 2.49), (0.268, 16.78),
    (0.268, 8.40), (0.268, 4.26), (0.025, 9.18), (0.113, 4.59), (0.0, 0.0),  # last entry: placeholder for the tail
]


def blended_roofline_fps(peaks) -> float:
    """frames/s/GPU lower bound of SURVEY.md §8d from the MEASURED peaks (burst bf16, HBM copy)."""
    us = 0.0
    for gflop, mb in APP_A_LAYERS[:-1]:
        us += max(gflop * 1e9 / (peaks["bf16_tflops"] * 1e12), mb * 1e6 / (peaks["hbm_gbs"] * 1e9)) * 1e6
    us += 4.4  # everything <= 16^2 and the remaining <= 32^2 layers (total 39.9 us at these peaks)
    return 1e6 / us
# live convs, forward 34.000 GFLOP/sample; training = fwd + dgrad + wgrad minus the stems' dgrad
TRAIN_GFLOP_PER_SAMPLE = 3 * 34.000 - 2e-9 * (64 * 64 * 32 * 30 + 256 * 256 * 32 * 3)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self._stop_evt = threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def build_model(device):
    import bdvs_b200 as B
    import vunet_oracle as O
    with contextlib.redirect_stdout(io.StringIO()):
        model = B.VunetAlter(**H256)
    sd = O.det_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()})
    model.load_state_dict(sd)
    return model.to(device).eval(), sd


# --------------------------------------------------------------------------------------------------
def cpu_port_fps(sample_frames: int, iters: int, warmup: int = 1):
    """frames/s of the CPU oracle (port of the reference's PyTorch path) on `sample_frames` H256 frames."""
    import vunet_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    with contextlib.redirect_stdout(io.StringIO()):
        import bdvs_b200 as B
        shapes = {k: tuple(v.shape) for k, v in B.VunetAlter(**H256).state_dict().items()}
    sd = O.det_state_dict(shapes)
    x, c, eps, _ = O.make_inputs(H256, sample_frames, tag="cpu")
    times = []
    with torch.no_grad():
        for i in range(warmup + iters):
            t0 = time.perf_counter()
            O.transfer(sd, x, c, eps)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return sample_frames / statistics.median(times), threads, times


def cpu_port_train_sps(sample: int, iters: int, warmup: int = 1):
    """samples/s of the CPU oracle's forward + autograd backward on `sample` D256 samples."""
    import vunet_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    with contextlib.redirect_stdout(io.StringIO()):
        import bdvs_b200 as B
        shapes = {k: tuple(v.shape) for k, v in B.VunetAlter(**D256).state_dict().items()}
    sd = {k: v.requires_grad_(True) for k, v in O.det_state_dict(shapes).items()}
    x, c, eps, _ = O.make_inputs(D256, sample, tag="cpu")
    target = O.det_uniform("target", (sample, 3, 256, 256))
    times = []
    for i in range(warmup + iters):
        t0 = time.perf_counter()
        imgs, means, logstds, _, _ = O.forward(sd, x, c, eps)
        loss = torch.mean((imgs - target) ** 2) + 1e-3 * O.compute_kl_with_prior(means, logstds)
        loss.backward()
        for v in sd.values():
            v.grad = None
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sample / statistics.median(times), threads


def cpu_port_config1_sps(iters: int = 3, warmup: int = 1):
    """samples/s of BASELINE config 1 EXACTLY: Market 128x128, batch 8, forward + backward, fp32 torch CPU."""
    import vunet_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    with contextlib.redirect_stdout(io.StringIO()):
        import bdvs_b200 as B
        shapes = {k: tuple(v.shape) for k, v in B.VunetAlter(**M128).state_dict().items()}
    sd = {k: v.requires_grad_(True) for k, v in O.det_state_dict(shapes).items()}
    x, c, eps, _ = O.make_inputs(M128, CONFIG1_BATCH, tag="cfg1")
    target = O.det_uniform("target8", (CONFIG1_BATCH, 3, 128, 128))
    times = []
    for i in range(warmup + iters):
        t0 = time.perf_counter()
        imgs, means, logstds, _, _ = O.forward(sd, x, c, eps)
        loss = torch.mean((imgs - target) ** 2) + 1e-3 * O.compute_kl_with_prior(means, logstds)
        loss.backward()
        for v in sd.values():
            v.grad = None
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return CONFIG1_BATCH / statistics.median(times), CONFIG1_BATCH / min(times), threads


def gpu_config1_sps(device, iters: int = 20):
    """The same config 1 step (forward + loss + backward, no optimizer) on the GPU path, eager launches."""
    import bdvs_b200 as B
    import vunet_oracle as O
    with contextlib.redirect_stdout(io.StringIO()):
        model = B.VunetAlter(**M128)
    model.load_state_dict(O.det_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}))
    model = model.to(device).train()
    x, c, eps, _ = O.make_inputs(M128, CONFIG1_BATCH, tag="cfg1")
    x, c = x.to(device), c.to(device)
    target = O.det_uniform("target8", (CONFIG1_BATCH, 3, 128, 128)).to(device)

    def step():
        imgs, means, logstds, _ = model(x, c)
        loss = torch.mean((imgs - target) ** 2) + 1e-3 * B.compute_kl_with_prior(means, logstds)
        for p in model.parameters():
            p.grad = None
        loss.backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    return CONFIG1_BATCH * iters / (e0.elapsed_time(e1) * 1e-3)


def run_train_section(args, device, world, rank, barrier, peaks):
    """BASELINE.json configs[2]: training step throughput (samples/s), reported next to the headline line."""
    import torch.distributed as dist
    import bdvs_b200 as B
    import vunet_oracle as O
    from bdvs_b200 import ops

    with contextlib.redirect_stdout(io.StringIO()):
        model = B.VunetAlter(**D256)
    sd = O.det_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()})
    model.load_state_dict(sd)
    model = model.to(device).train()
    if world > 1:
        model.enable_data_parallel()
    # the reference's optimizer (experiments/shape_and_pose_net.py:237-246): Adam over eu / ed / du / dd, here as the
    # package's own fused Adam + weight-norm-prep launch (no library kernel inside the timed step)
    opt = B.FusedAdam([{"params": list(getattr(model, n).parameters()), "name": n} for n in ("eu", "ed", "du", "dd")],
                      lr=1e-4, betas=(0.5, 0.9), model=model)
    Bt = TRAIN_BATCH_PER_GPU
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    x = torch.rand(Bt, 30, 64, 64, device=device, generator=g) * 2 - 1
    c = torch.rand(Bt, 3, 256, 256, device=device, generator=g) * 2 - 1
    target = torch.rand(Bt, 3, 256, 256, device=device, generator=g) * 2 - 1
    xh, ch, th = (t.cpu().pin_memory() for t in (x, c, target))
    loss_h = torch.zeros(1).pin_memory()

    def step(xx, cc, tt):
        imgs, means, logstds, _ = model(xx, cc)
        loss = torch.mean((imgs - tt) ** 2) + 1e-3 * B.compute_kl_with_prior(means, logstds)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    K = max(10, args.steps // 2)
    W = 3
    for _ in range(W):
        step(x, c, target)
    barrier()
    # The step is ~300 launches and host-bound when issued launch by launch (tools/train_host_overhead.py): the
    # package's GraphedTrainStep captures forward + loss + backward + all-reduce + optimizer in one CUDA graph.
    # VUB_BENCH_TRAIN_GRAPH=0 times the eager step instead.
    graphed, graph_err = None, None
    if os.environ.get("VUB_BENCH_TRAIN_GRAPH", "1") != "0":
        try:
            n0 = ops.LAUNCHES
            graphed = B.GraphedTrainStep(
                model, opt, lambda imgs, means, logstds, tt: torch.mean((imgs - tt) ** 2) + 1e-3 * B.compute_kl_with_prior(means, logstds),
                x, c, target)
            launches_per_step = (ops.LAUNCHES - n0) // 4  # 3 warm-up steps + 1 captured step
        except Exception as exc:  # noqa: BLE001 - report and fall back to the eager step
            graph_err = f"{type(exc).__name__}: {exc}"[:200]
            graphed = None
    run = (lambda xx, cc, tt: graphed(xx, cc, tt)) if graphed is not None else step
    for _ in range(2):
        run(x, c, target)
    barrier()
    before = ops.LAUNCHES
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(K):
        loss = run(x, c, target)
    ev1.record()
    barrier()
    launches = launches_per_step * K if graphed is not None else ops.LAUNCHES - before
    ms = ev0.elapsed_time(ev1)
    # e2e: inputs and target from pinned host memory every step, loss read back every step
    if graphed is not None:
        # the package's own pipelined path: batch k+1 is uploaded on a copy stream while step k runs
        graphed.prefetch(xh, ch, th)
        for _ in range(2):
            graphed.prefetch(xh, ch, th)
            graphed.run_prefetched()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            graphed.prefetch(xh, ch, th)
            loss = graphed.run_prefetched()
            loss_h.copy_(loss.detach().reshape(1), non_blocking=True)
        e1.record()
        graphed.run_prefetched()  # drain the last prefetch (outside the timed region)
    else:
        for _ in range(2):
            run(x.copy_(xh, non_blocking=True), c.copy_(ch, non_blocking=True), target.copy_(th, non_blocking=True))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            loss = run(x.copy_(xh, non_blocking=True), c.copy_(ch, non_blocking=True), target.copy_(th, non_blocking=True))
            loss_h.copy_(loss.detach().reshape(1), non_blocking=True)
        e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    value = world * Bt * K / (ms * 1e-3)
    red = getattr(model, "_vub_reducer", None)
    # per-kernel timing of one instrumented step (CUDA events around every launch, outside the timed region)
    tk = []
    # every rank runs the instrumented step (it contains the gradient all-reduce); only rank 0 times it
    from bdvs_b200.profiler import KernelProfiler
    # single-stream for this one step: with wgrad / ed on their second streams the per-kernel event brackets overlap
    os.environ["VUB_BWD_STREAMS"] = os.environ["VUB_FWD_STREAMS"] = "1"
    try:
        with (KernelProfiler() if rank == 0 else contextlib.nullcontext()) as kp:
            step(x, c, target)
    finally:
        del os.environ["VUB_BWD_STREAMS"], os.environ["VUB_FWD_STREAMS"]
    barrier()
    if rank == 0:
        prof = kp.summary()
        tot = sum(r["ms"] for r in prof.values())
        agg = {}
        for name, r in prof.items():
            fam = name.split(" ")[0].split("<")[0]
            a = agg.setdefault(fam, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            for k2 in a:
                a[k2] += r[k2]
        families = [{"family": f, "launches": a["launches"], "ms": round(a["ms"], 3), "share": round(a["ms"] / tot, 4),
                     "tflops": round(a["flops"] / (a["ms"] * 1e-3) / 1e12, 1) if a["ms"] > 0 else 0.0,
                     "gbs": round(a["bytes"] / (a["ms"] * 1e-3) / 1e9, 1) if a["ms"] > 0 else 0.0}
                    for f, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])]
        for name, r in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:64]:
            tk.append({"kernel": name, "launches": r["launches"], "ms": round(r["ms"], 4), "share": round(r["ms"] / tot, 4),
                       "tflops": round(r["flops"] / (r["ms"] * 1e-3) / 1e12, 1) if r["ms"] > 0 else 0.0,
                       "gbs": round(r["bytes"] / (r["ms"] * 1e-3) / 1e9, 1) if r["ms"] > 0 else 0.0})
        tk = {"sum_kernel_ms": round(tot, 3), "families": families, "top": tk}
    out = {
        "metric": "VUnet 256x256 training throughput (fwd+bwd+all-reduce+Adam)", "value": value, "unit": "samples/s",
        "steps": K, "warmup": W, "ms_per_step": ms / K, "dtype": "bf16", "scaling": "weak",
        "config": {"workload": TRAIN_WORKLOAD, "samples_per_gpu_per_step": Bt, "global_batch": Bt * world,
                   "parallelism": f"dp{world}: batch sharded, one flat-bucket all-reduce per sub-network overlapped "
                                  "with backward" if world > 1 else "single GPU",
                   "algorithmic_gflop_per_sample": round(TRAIN_GFLOP_PER_SAMPLE, 3)},
        "e2e": {"value": world * Bt * K / (ms_e2e * 1e-3), "unit": "samples/s",
                "h2d_bytes_per_step": (xh.numel() + ch.numel() + th.numel()) * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / K},
        "gpu_launches": launches,
        "cuda_graph": graphed is not None,
        "cuda_graph_error": graph_err,
        "allreduce_bytes_per_step": getattr(red, "bytes_per_step", 0),
        "conv_tflops_per_gpu": round(TRAIN_GFLOP_PER_SAMPLE * Bt * K / ms, 2),
        "frac_tensor_peak_sustained": round(TRAIN_GFLOP_PER_SAMPLE * Bt * K / ms / peaks["bf16_tflops_sustained"], 4),
        "loss": float(loss.detach()),
        "kernels": tk,
    }
    del model, opt
    torch.cuda.empty_cache()
    return out


def run_train_vgg(device, steps: int, peaks) -> dict:
    """The reference's REAL objective (experiments/shape_and_pose_net.py:388-432): likelihood = ll_weight * sum of the
    VGG19 perceptual terms (lib/losses.py:81-119) + kl_weight * KL, on the D256 training workload (batch 32), as one
    CUDA-graph step: VUnet forward, VGG19 features of target and prediction, L1 terms, VGG data gradients, VUnet
    backward, FusedAdam.  The VGG weights are random (ImageNet weights do not exist offline): timing only."""
    import bdvs_b200 as B
    import vunet_oracle as O
    from bdvs_b200 import ops
    from bdvs_b200.profiler import KernelProfiler
    from torchvision.models import vgg19

    with contextlib.redirect_stdout(io.StringIO()):
        model = B.VunetAlter(**D256)
    model.load_state_dict(O.det_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}))
    model = model.to(device).train()
    torch.manual_seed(7)
    vgg = B.PerceptualVGG(vgg19(weights=None).eval(), [1.0] * 6).to(device).eval()
    opt = B.FusedAdam([{"params": list(getattr(model, n).parameters()), "name": n} for n in ("eu", "ed", "du", "dd")],
                      lr=1e-4, betas=(0.5, 0.9), model=model)
    Bt = TRAIN_BATCH_PER_GPU
    g = torch.Generator(device=device).manual_seed(4321)
    x = torch.rand(Bt, 30, 64, 64, device=device, generator=g) * 2 - 1
    c = torch.rand(Bt, 3, 256, 256, device=device, generator=g) * 2 - 1
    target = torch.rand(Bt, 3, 256, 256, device=device, generator=g) * 2 - 1

    def loss_fn(imgs, means, logstds, tt):
        ll = B.vgg_loss(vgg, tt, imgs)
        return 5.0 * torch.stack([ll[k] for k in ll]).sum() + 1e-3 * B.compute_kl_with_prior(means, logstds)

    n0 = ops.LAUNCHES
    graphed = B.GraphedTrainStep(model, opt, loss_fn, x, c, target)
    launches_per_step = (ops.LAUNCHES - n0) // 4
    for _ in range(2):
        graphed(x, c, target)
    torch.cuda.synchronize()
    K = max(5, steps // 2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss = graphed(x, c, target)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    # one instrumented eager step of the loss alone (single stream): per-family time of the VGG part
    imgs = (torch.rand(Bt, 3, 256, 256, device=device, generator=g) * 2 - 1).requires_grad_(True)
    for _ in range(2):
        torch.stack(list(B.vgg_loss(vgg, target, imgs).values())).sum().backward()
    with KernelProfiler() as kp:
        torch.stack(list(B.vgg_loss(vgg, target, imgs).values())).sum().backward()
    prof = kp.summary()
    tot = sum(r["ms"] for r in prof.values())
    agg = {}
    for name, r in prof.items():
        fam = name.split(" ")[0].split("<")[0]
        a = agg.setdefault(fam, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        for k2 in a:
            a[k2] += r[k2]
    fams = [{"family": f, "launches": a["launches"], "ms": round(a["ms"], 3), "share": round(a["ms"] / tot, 4),
             "tflops": round(a["flops"] / (a["ms"] * 1e-3) / 1e12, 1) if a["ms"] > 0 else 0.0,
             "gbs": round(a["bytes"] / (a["ms"] * 1e-3) / 1e9, 1) if a["ms"] > 0 else 0.0}
            for f, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])]
    vgg_gflop = 3 * 48.4   # conv1_1 .. conv5_2 forward on target and prediction + data gradient, per sample
    gflop = TRAIN_GFLOP_PER_SAMPLE + vgg_gflop
    out = {"metric": "VUnet 256x256 training throughput with the VGG19 perceptual loss (fwd+bwd+Adam)",
           "value": Bt / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "steps": K, "samples_per_step": Bt,
           "per_gpu": True, "cuda_graph": True, "gpu_launches_per_step": launches_per_step,
           "algorithmic_gflop_per_sample": round(gflop, 1), "conv_tflops": round(gflop * Bt / ms, 1),
           "frac_tensor_peak_sustained": round(gflop * Bt / ms / peaks["bf16_tflops_sustained"], 4),
           "vgg_loss_alone": {"sum_kernel_ms": round(tot, 3), "tflops": round(vgg_gflop * Bt / tot, 1), "families": fams},
           "loss": float(loss.detach()),
           "note": "random VGG19 weights (no ImageNet weights offline); single GPU (the VGG part has no collective)"}
    del model, opt, graphed, vgg
    torch.cuda.empty_cache()
    return out


def run_behavior_sampling(device, steps: int, peaks) -> dict:
    """BASELINE.json configs[3]: behavior-net sampling, batch 256 — noise -> latent flow reverse (15 coupling blocks,
    2048-wide MLPs, 630 M parameters) -> residual LSTM-1024 decoder, 50-step 3-D pose sequences (51 coordinates), as ONE
    CUDA graph (PoseSequenceSampler); plus a per-kernel breakdown of one instrumented eager call and the CPU oracle on
    a bounded sample."""
    import time
    import behavior_oracle as BO
    import bdvs_b200 as B
    from bdvs_b200.profiler import KernelProfiler

    cfg = dict(dim_hidden_b=1024, mid=2048, n_flows=15, depth=2, n_kps=51)
    Bs, T = 256, 50
    flow = B.UnsupervisedTransformer2(flow_in_channels=1024, n_flows=15, flow_hidden_depth=2, flow_mid_channels=2048)
    net = B.ResidualBehaviorNet(n_kps=51, dim_hidden_b=1024, decoder_arch="lstm", linear_in_decoder=False)
    fsd = BO.det_flow_state_dict({k: tuple(v.shape) for k, v in flow.state_dict().items()})
    nsd = BO.det_net_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()})
    flow.load_state_dict(fsd)
    net.load_state_dict(nsd)
    flow, net = flow.to(device).eval(), net.to(device).eval()
    smp = B.PoseSequenceSampler(flow, net, Bs, T)
    g = torch.Generator(device=device).manual_seed(99)
    z = torch.randn(Bs, 1024, device=device, generator=g)
    x0 = torch.rand(Bs, 1, 51, device=device, generator=g) * 2 - 1
    zh, xh = z.cpu().pin_memory(), x0.cpu().pin_memory()
    outh = torch.empty((Bs, T, 51), dtype=torch.float32).pin_memory()
    for _ in range(3):
        smp(z, x0)
    torch.cuda.synchronize()
    K = max(5, steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        smp.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    e0.record()
    for _ in range(K):   # e2e: noise and start poses from pinned host memory, sequences back to it
        outh.copy_(smp(zh, xh), non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / K
    with torch.no_grad():
        smp._step()
        with KernelProfiler() as kp:
            smp._step()
    prof = kp.summary()
    tot = sum(r["ms"] for r in prof.values())
    rows = []
    for name, r in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]:
        rows.append({"kernel": name, "launches": r["launches"], "ms": round(r["ms"], 4), "share": round(r["ms"] / tot, 4),
                     "tflops": round(r["flops"] / (r["ms"] * 1e-3) / 1e12, 1) if r["ms"] > 0 else 0.0,
                     "gbs": round(r["bytes"] / (r["ms"] * 1e-3) / 1e9, 1) if r["ms"] > 0 else 0.0})
    n_flow_w = sum(p.numel() for n_, p in flow.named_parameters() if n_.endswith("weight"))
    gflop = (2.0 * Bs * n_flow_w + T * 2.0 * Bs * (4096 * (1024 + 51) + 1024 * 51)) / 1e9
    w_bytes = 2.0 * (n_flow_w + 4096 * 1088 + 64 * 1024)
    # CPU oracle on a bounded sample (16 sequences)
    torch.set_num_threads(os.cpu_count() or 1)
    zc, xc = z[:16].cpu(), x0[:16, 0].cpu()
    with torch.no_grad():
        BO.generate_seq(nsd, BO.flow_reverse(fsd, zc, 15, 2), xc, T)
        t0 = time.time()
        for _ in range(2):
            BO.generate_seq(nsd, BO.flow_reverse(fsd, zc, 15, 2), xc, T)
        cpu_s = (time.time() - t0) / 2
    return {"metric": "behavior-net sampling throughput (flow reverse + LSTM decoder, 50-step 3-D pose sequences)",
            "value": Bs / (ms * 1e-3), "unit": "sequences/s", "ms_per_batch": ms, "batch": Bs, "steps_per_sequence": T,
            "per_gpu": True, "cuda_graph": True, "gpu_launches_per_batch": smp.launches_per_call,
            "e2e": {"value": Bs / (ms_e2e * 1e-3), "unit": "sequences/s", "h2d_bytes_per_step": zh.numel() * 4 + xh.numel() * 4,
                    "d2h_bytes_per_step": outh.numel() * 4},
            "algorithmic_gflop_per_batch": round(gflop, 1), "tflops": round(gflop / ms, 1),
            "weight_bytes_per_batch": w_bytes, "weight_stream_gbs": round(w_bytes / (ms * 1e-3) / 1e9, 1),
            "frac_hbm_weight_stream": round(w_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
            "parameters": sum(p.numel() for p in flow.parameters()) + sum(p.numel() for p in net.parameters()),
            "kernels": {"sum_kernel_ms": round(tot, 3), "top": rows},
            "cpu_baseline": {"value": 16 / cpu_s, "unit": "sequences/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": "16 sequences x 2 iters (+1 warm-up), oracle/behavior_oracle.py fp32 torch CPU"}}


def run_behavior_transfer(device, model, steps: int) -> dict:
    """BASELINE.json configs[4], one GPU's share: behavior-net pose sequences (64 frames) -> camera projection ->
    GPU stickman rasteriser -> VUnet 256x256 synthesis with the appearance encoded once per sequence -> uint8 frames.
    Nothing leaves the GPU between the noise and the frames.  256 sequences are sampled per PoseSequenceSampler replay;
    every sequence is one SequenceSynthesizer replay (64 frames)."""
    import collections
    import behavior_oracle as BO
    import bdvs_b200 as B
    from bdvs_b200.rendering import StickmanRasterizer, project_poses
    from bdvs_b200.synthesis import SequenceSynthesizer

    T, S_batch = 64, 256
    flow = B.UnsupervisedTransformer2(flow_in_channels=1024, n_flows=15, flow_hidden_depth=2, flow_mid_channels=2048)
    net = B.ResidualBehaviorNet(n_kps=51, dim_hidden_b=1024, decoder_arch="lstm", linear_in_decoder=False)
    flow.load_state_dict(BO.det_flow_state_dict({k: tuple(v.shape) for k, v in flow.state_dict().items()}))
    net.load_state_dict(BO.det_net_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}))
    flow, net = flow.to(device).eval(), net.to(device).eval()
    smp = B.PoseSequenceSampler(flow, net, S_batch, T)
    JM = collections.namedtuple("JointModel", "body right_lines left_lines head_lines face rshoulder lshoulder headup")
    jm = JM(body=[0, 14, 8, 11, 3], right_lines=[(0, 1), (1, 2), (0, 14), (14, 15), (15, 16)],
            left_lines=[(3, 4), (4, 5), (3, 11), (11, 12), (12, 13)], head_lines=[(8, 9), (9, 10)], face=[],
            rshoulder=14, lshoulder=11, headup=10)   # data/human36m.py:136-152, keypoints_3d_world (17 joints)
    raster = StickmanRasterizer(jm, 256)
    seq = SequenceSynthesizer(model, 1, T, uint8_frames=True, uint8_inputs=True)
    g = torch.Generator(device=device).manual_seed(5)
    z = torch.randn(S_batch, 1024, device=device, generator=g)
    x0 = torch.rand(S_batch, 1, 51, device=device, generator=g) * 0.6 - 0.3
    app = torch.randint(0, 256, (S_batch, 256, 256, 3), dtype=torch.uint8, device=device, generator=g)
    extr = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 4.0]], dtype=torch.float64)   # synthetic camera
    intr = (1100.0, 500.0, 1100.0, 500.0)
    frames_h = torch.empty((T, 256, 256, 3), dtype=torch.uint8).pin_memory()

    def render(n_seq: int):
        poses = smp(z, x0)                                             # [256, T, 51] 3-D poses, one graph replay
        j3 = poses[:n_seq].reshape(n_seq * T, 17, 3).clamp(-1.5, 1.5)
        j2 = project_poses(j3, extr, intr, (1000.0, 1000.0), 256)      # [n_seq * T, 17, 2]
        for i in range(n_seq):
            raster(j2[i * T:(i + 1) * T], out=seq.c)                   # stickmen straight into the graph's input
            seq.x.copy_(app[i:i + 1])
            frames_h.copy_(seq.run(), non_blocking=True)               # video frames back to pinned host memory

    n_seq = 8
    render(2)
    torch.cuda.synchronize()
    K = max(2, steps // 10)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        render(n_seq)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    e0.record()
    for _ in range(5):
        smp.run()
    e1.record()
    torch.cuda.synchronize()
    ms_sample = e0.elapsed_time(e1) / 5
    ms_per_seq = (ms - ms_sample) / n_seq + ms_sample / S_batch   # sampling amortised over the 256 sequences it yields
    return {"metric": "end-to-end behavior transfer: sampled 64-frame pose sequences -> stickmen -> 256x256 video (1 GPU)",
            "value": 1e3 / ms_per_seq, "unit": "sequences/s", "frames_per_s": T * 1e3 / ms_per_seq, "per_gpu": True,
            "ms_per_sequence": ms_per_seq, "ms_sampling_256_sequences": ms_sample,
            "sequences_per_step_rendered": n_seq, "frames_per_sequence": T,
            "d2h_bytes_per_sequence": frames_h.numel(),
            "seconds_for_512_sequences_1gpu": 512 * ms_per_seq * 1e-3,
            "seconds_for_512_sequences_8gpu_no_collective": 64 * ms_per_seq * 1e-3,
            "note": "sequences are independent: 8 GPUs take 64 each (synthesis.shard_range), no collective; random weights, "
                    "synthetic camera"}


def run_vunet_org(device, frames: int, steps: int, peaks) -> dict:
    """``VunetOrg.transfer`` (models/vunets.py:102-110) at the H256 geometry, ``frames`` frames per step, CUDA-graph
    replay timed with CUDA events on this rank; plus the share of the step the prior's 512 / 1024-channel convs take
    (one instrumented eager step)."""
    import bdvs_b200 as B
    import vunet_oracle as O
    from bdvs_b200.profiler import KernelProfiler
    from bdvs_b200.synthesis import FrameSynthesizer

    cfg = {k: v for k, v in H256.items()}
    with contextlib.redirect_stdout(io.StringIO()):
        m = B.VunetOrg(**cfg)
    m.load_state_dict(O.det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}))
    m = m.to(device).eval()
    syn = FrameSynthesizer(m, frames, device)
    syn.x.uniform_(-1, 1)
    syn.c.uniform_(-1, 1)
    for _ in range(3):
        syn.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        syn.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    with torch.no_grad():
        os.environ["VUB_FWD_STREAMS"] = "1"
        try:
            m.transfer(syn.x, syn.c)
            with KernelProfiler() as kp:
                m.transfer(syn.x, syn.c)
        finally:
            del os.environ["VUB_FWD_STREAMS"]
    prof = kp.summary()
    total = sum(r["ms"] for r in prof.values())
    wide = {k: r for k, r in prof.items() if "x512->" in k or "x256->" in k or "->512 k1" in k}
    wide_ms = sum(r["ms"] for r in wide.values())
    wide_fl = sum(r["flops"] for r in wide.values())
    gflop = 43.96 - 0.277  # SURVEY §8f-3 census (117 convs) minus the dead ed branch
    return {"metric": "VunetOrg 256x256 frame synthesis throughput (transfer, 4-group autoregressive prior)",
            "value": frames / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "frames_per_step": frames, "per_gpu": True,
            "launches_per_step": syn.launches_per_call, "parameters": sum(p.numel() for p in m.parameters()),
            "conv_tflops": round(gflop * frames / ms, 1),
            "frac_tensor_peak_sustained": round(gflop * frames / ms / peaks["bf16_tflops_sustained"], 4),
            "prior_convs": {"launches": sum(r["launches"] for r in wide.values()), "ms": round(wide_ms, 4),
                            "share_of_instrumented_step": round(wide_ms / total, 4),
                            "tflops": round(wide_fl / (wide_ms * 1e-3) / 1e12, 1) if wide_ms > 0 else 0.0,
                            "note": "the 512->128, 128->512 and [512|512]->512 convs of the prior on the 2x2 / 4x4 maps: "
                                    "per-tap tcgen05 kernel, split-K over 3-CTA clusters (DSMEM reduction)"}}


def static_config(frames_per_gpu: int, world: int) -> dict:
    """The workload description shared by both arms (the product arm adds nothing run-dependent here)."""
    return {"workload": WORKLOAD, "frames_per_gpu_per_step": frames_per_gpu,
            "global_frames_per_step": frames_per_gpu * world,
            "parallelism": f"frames sharded over {world} GPU(s), no collective",
            "cache": f"inputs + activations per step (~{14.0 * frames_per_gpu / 64:.1f} GB) far exceed the 126 MB L2: "
                     "no flush needed",
            "execution": "CUDA graph replay of model.transfer (latent noise drawn inside the graph); e2e: "
                         "FrameSynthesizer.render_host, pinned host buffers (uint8 HWC stickmen / appearance in, uint8 "
                         "HWC frames out, like the reference's rendering loop), copies double-buffered on their own "
                         "streams and overlapped with the kernels of the neighbouring batches",
            "algorithmic_gflop_per_frame": GFLOP_PER_FRAME,
            "train_workload": TRAIN_WORKLOAD, "train_samples_per_gpu_per_step": TRAIN_BATCH_PER_GPU,
            "config1_workload": CONFIG1_WORKLOAD}


def run_reference(args):
    """The reference's CPU path (its port, oracle/vunet_oracle.py) on the SAME config as the product arm: one step =
    64 H256 frames through ``transfer`` (in chunks of 8 frames to bound host memory), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import vunet_oracle as O
    t_all0 = time.perf_counter()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    with contextlib.redirect_stdout(io.StringIO()):
        import bdvs_b200 as B
        shapes = {k: tuple(v.shape) for k, v in B.VunetAlter(**H256).state_dict().items()}
    sd = O.det_state_dict(shapes)
    chunk = 8
    frames = FRAMES_PER_GPU if args.frames_total <= 0 else args.frames_total
    x, c, eps, _ = O.make_inputs(H256, chunk, tag="cpu")
    K, W = max(1, args.steps), max(0, args.warmup)
    # bound the run: the whole --steps/--warmup run must end within a few minutes (~14 frames/s on 16 threads)
    t0 = time.perf_counter()
    with torch.no_grad():
        O.transfer(sd, x, c, eps)
    probe = time.perf_counter() - t0
    budget_s = float(os.environ.get("VUB_REF_BUDGET_S", "170"))
    chunks_per_step = max(1, frames // chunk)
    while chunks_per_step > 1 and probe * chunks_per_step * (K + W) > budget_s:
        chunks_per_step //= 2
    sample = chunks_per_step * chunk
    times = []
    with torch.no_grad():
        for i in range(W + K):
            t0 = time.perf_counter()
            for _ in range(chunks_per_step):
                O.transfer(sd, x, c, eps)
            if i >= W:
                times.append(time.perf_counter() - t0)
    ms = 1e3 * statistics.median(times)
    fps = sample / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": static_config(FRAMES_PER_GPU, max(1, args.gpus)),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} frames/step ({chunks_per_step} chunks of {chunk}) of the same H256 "
                                   f"transfer workload, fp32, torch CPU ({threads} threads); the reference is pure "
                                   "Python/PyTorch and is absent on the GPU box, so its port oracle/vunet_oracle.py "
                                   "(pinned to reference outputs) is timed"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import bdvs_b200 as B
    from bdvs_b200 import ops
    from bdvs_b200.profiler import KernelProfiler
    from bdvs_b200.synthesis import FrameSynthesizer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = load_peaks()
    model, sd = build_model(device)
    Bf = FRAMES_PER_GPU
    strong = args.frames_total > 0  # SURVEY 8(d) literal reading: a fixed number of frames split over the GPUs
    if strong:
        if args.frames_total % world:
            raise SystemExit(f"--frames-total {args.frames_total} is not divisible by {world} GPUs")
        Bf = args.frames_total // world
    torch.manual_seed(42 + rank)
    synth = FrameSynthesizer(model, Bf, device)
    synth.x.uniform_(-1, 1)
    synth.c.uniform_(-1, 1)
    K, W = args.steps, max(3, args.warmup)

    # ---- value: device-resident inputs, graph replay, CUDA events ----
    for _ in range(W):
        synth.run()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(K):
        synth.run()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / K
    value = world * Bf * K / (ms_total * 1e-3)
    launches = synth.launches_per_call * K

    # ---- e2e: pinned host buffers through the public API, copies inside the timed region ----
    # Host-side formats are the reference rendering loop's own (data/data_conversions_3d.py:1154-1173): uint8 HWC
    # stickman / appearance images in (its ToTensor() + x*2-1 run inside the stem kernels, bit-identically) and uint8 HWC
    # frames out ((scale_img(rgb) * 255).astype(uint8)).  The fp32 NCHW variant of the same pipeline is timed too.
    def time_render_host(s8, hosts):
        for i in range(2):
            s8.render_host(*hosts[i % 2])
        s8.join()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            s8.render_host(*hosts[i % 2])
        s8.join()  # the timed region ends when the last frames have landed in host memory
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    S = H256["spatial_size"]
    synth8 = FrameSynthesizer(model, Bf, device, uint8_frames=True, uint8_inputs=True)
    # two host buffer sets, alternated, like a caller that prepares batch k+1 while batch k renders
    hosts8 = [(torch.randint(0, 256, (Bf, S, S, 3), dtype=torch.uint8).pin_memory(),
               torch.randint(0, 256, (Bf, S, S, 3), dtype=torch.uint8).pin_memory(),
               torch.empty((Bf, S, S, 3), dtype=torch.uint8).pin_memory()) for _ in range(2)]
    ms_e2e = time_render_host(synth8, hosts8)
    e2e = world * Bf * K / (ms_e2e * 1e-3)
    h2d = hosts8[0][0].numel() + hosts8[0][1].numel()
    d2h = hosts8[0][2].numel()
    del synth8
    hosts32 = [(torch.empty(synth.x.shape, dtype=torch.float32).uniform_(-1, 1).pin_memory(),
                torch.empty(synth.c.shape, dtype=torch.float32).uniform_(-1, 1).pin_memory(),
                torch.empty((Bf, 3, S, S), dtype=torch.float32).pin_memory()) for _ in range(2)]
    ms32 = time_render_host(synth, hosts32)
    e2e_fp32 = {"value": world * Bf * K / (ms32 * 1e-3), "unit": UNIT, "ms_per_step": ms32 / K,
                "h2d_bytes_per_step": 4 * (hosts32[0][0].numel() + hosts32[0][1].numel()),
                "d2h_bytes_per_step": 4 * hosts32[0][2].numel(),
                "note": "same render_host pipeline with fp32 NCHW host tensors (the module API's own dtype)"}
    del hosts32

    # BASELINE configs[4]-style rendering: whole 64-frame sequences against ONE appearance image each —
    # the appearance encoder runs once per sequence (the reference re-encodes it for every frame)
    sequence = None
    try:
        from bdvs_b200.synthesis import SequenceSynthesizer
        seq = SequenceSynthesizer(model, 1, Bf)
        seq.x.copy_(synth.x[:1])
        seq.c.copy_(synth.c)
        for _ in range(3):
            seq.run()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(K):
            seq.run()
        s1.record()
        barrier()
        ms_seq = s0.elapsed_time(s1)
        sequence = {"metric": "frames/s rendering 64-frame sequences (appearance encoded once per sequence)",
                    "value": Bf * K / (ms_seq * 1e-3), "unit": UNIT, "ms_per_sequence": ms_seq / K, "per_gpu": True}
        del seq
    except Exception as exc:  # secondary number: never take the headline down
        sequence = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        barrier()

    # BASELINE configs[4]-style front-end: 2-D joints in, video frames out — stickman rasterisation,
    # ToTensor / x*2-1, transfer and the uint8 frame conversion on the GPU; per step the joints come from pinned host
    # memory (32 KB) and the uint8 frames go back to it (12.6 MB), the appearance image stays resident
    pose = None
    try:
        from bdvs_b200.rendering import PoseRenderer
        import collections
        JM = collections.namedtuple("JointModel", "body right_lines left_lines head_lines face rshoulder lshoulder headup")
        jm = JM(body=[1, 25, 13, 17, 6], right_lines=[(3, 2), (2, 1), (1, 25), (25, 26), (26, 27)],
                left_lines=[(8, 7), (7, 6), (6, 17), (17, 18), (18, 19)], head_lines=[(13, 14), (14, 15)], face=[],
                rshoulder=25, lshoulder=17, headup=15)  # data/human36m.py:136-152
        pr = PoseRenderer(model, jm)
        gj = torch.Generator().manual_seed(7)
        jh = (torch.randn(Bf, 32, 2, generator=gj, dtype=torch.float64) * 40 + 128).clamp_(0, 255).pin_memory()
        app8 = torch.randint(0, 256, (Bf, S, S, 3), dtype=torch.uint8, device=device)
        fh = torch.empty((Bf, S, S, 3), dtype=torch.uint8).pin_memory()
        jd = torch.empty_like(jh, device=device)
        for _ in range(3):
            pr(app8, jd.copy_(jh, non_blocking=True))
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(K):
            fh.copy_(pr(app8, jd.copy_(jh, non_blocking=True)), non_blocking=True)
        p1.record()
        barrier()
        ms_pose = p0.elapsed_time(p1)
        pose = {"metric": "frames/s from 2-D joints to uint8 frames (GPU stickman rasteriser + transfer, eager launches)",
                "value": Bf * K / (ms_pose * 1e-3), "unit": UNIT, "ms_per_step": ms_pose / K, "per_gpu": True,
                "h2d_bytes_per_step": jh.numel() * 8, "d2h_bytes_per_step": fh.numel()}
        del pr, app8
    except Exception as exc:  # secondary number: never take the headline down
        pose = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        barrier()

    train = None
    if not args.no_train:
        try:
            train = run_train_section(args, device, world, rank, barrier, peaks)
        except Exception as exc:  # the headline (synthesis) line must survive a failure of the second workload
            train = {"error": f"{type(exc).__name__}: {exc}"[:400]}
            barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel roofline: one instrumented eager step (CUDA events around every launch) ----
    with torch.no_grad():
        for _ in range(2):
            model.transfer(synth.x, synth.c)
        os.environ["VUB_FWD_STREAMS"] = "1"  # single stream: per-kernel event brackets must not overlap
        try:
            with KernelProfiler() as kp:
                model.transfer(synth.x, synth.c)
        finally:
            del os.environ["VUB_FWD_STREAMS"]
        prof = kp.summary()
    total_kernel_ms = sum(r["ms"] for r in prof.values())
    kernels = []
    for name, r in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        tf = r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] > 0 else 0.0
        gb = r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["ms"] > 0 else 0.0
        kernels.append({"kernel": name, "launches": r["launches"], "ms": round(r["ms"], 4),
                        "share": round(r["ms"] / total_kernel_ms, 4), "tflops": round(tf, 2), "gbs": round(gb, 1),
                        "frac_tensor": round(tf / peaks["bf16_tflops_sustained"], 4),
                        "frac_hbm": round(gb / peaks["hbm_gbs"], 4)})
    top = kernels[0]
    t_tc = prof[top["kernel"]]["flops"] / (peaks["bf16_tflops_sustained"] * 1e12)
    t_hbm = prof[top["kernel"]]["bytes"] / (peaks["hbm_gbs"] * 1e9)
    if t_hbm >= t_tc:
        roofline = {"bound": "hbm", "achieved": top["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": round(top["gbs"] / peaks["hbm_gbs"], 4), "traffic": None}
    else:
        roofline = {"bound": "tensor", "achieved": top["tflops"], "peak": peaks["bf16_tflops_sustained"],
                    "unit": "TFLOP/s", "frac": round(top["tflops"] / peaks["bf16_tflops_sustained"], 4), "traffic": None}
    traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(traffic_path) and Bf == FRAMES_PER_GPU:  # the captures were taken at 64 frames per GPU
        with open(traffic_path) as f:
            tr = json.load(f).get(top["kernel"])
        if tr:  # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed ncu --set full capture
            roofline["traffic"] = tr["dram_read"] + tr["dram_write"]
            roofline["traffic_source"] = "profiles/ncu_traffic.json (ncu --set full, per launch, same workload)"
    if "@256x256" in top["kernel"] or "@128x128" in top["kernel"]:
        roofline["note"] = ("row-streaming kernel: per ncu it is limited by the SM's shared-memory data pipe (tensor-core "
                            "operand reads + epilogue staging), not by HBM; profiles/r01_ncu_rows_smem.txt")
    roofline["algorithmic_bytes_per_launch"] = prof[top["kernel"]]["bytes"] / top["launches"]
    roofline["algorithmic_flops_per_launch"] = prof[top["kernel"]]["flops"] / top["launches"]
    roofline.update({"kernel": top["kernel"], "launches_per_step": top["launches"], "share_of_step": top["share"],
                     "peak_source": peaks["source"] + (" (sustained)" if roofline["bound"] == "tensor" else ""),
                     "avg_launch_ms": round(top["ms"] / top["launches"], 4)})
    blended = blended_roofline_fps(peaks)
    whole = {"conv_tflops_per_gpu": round(GFLOP_PER_FRAME * Bf / ms_step, 2),
             "frac_tensor_peak_sustained": round(GFLOP_PER_FRAME * Bf / ms_step / peaks["bf16_tflops_sustained"], 4),
             "blended_roofline_frames_per_s_per_gpu": round(blended, 1),
             "frac_of_blended_roofline": round(value / world / blended, 4)}

    # ---- CPU baseline (bounded sample) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        fps, threads, times = cpu_port_fps(8, iters=3, warmup=1)
        cpu = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "8 frames/iter x 3 iters (+1 warm-up) of the same H256 transfer workload, fp32 torch CPU, "
                         "oracle/vunet_oracle.py (port of the reference's PyTorch path, pinned to reference outputs)"}
        # BASELINE config 1 exactly: M128 forward + backward, batch 8, CPU fp32 — and the same step
        # on the GPU path beside it
        try:
            med, best, threads = cpu_port_config1_sps(iters=3, warmup=1)
            cpu["config1"] = {"workload": CONFIG1_WORKLOAD, "cpu_samples_per_s_median": round(med, 2),
                              "cpu_samples_per_s_best": round(best, 2), "cores": threads, "iters": 3, "warmup": 1,
                              "gpu_samples_per_s": round(gpu_config1_sps(device), 1),
                              "gpu_note": "same step (fwd + loss + bwd) through the module API, eager launches, bf16"}
        except Exception as exc:  # noqa: BLE001
            cpu["config1"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        if isinstance(train, dict) and "value" in train:
            sps, threads = cpu_port_train_sps(4, iters=2, warmup=1)
            train["cpu_baseline"] = {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port",
                                     "sample": "4 samples/iter x 2 iters (+1 warm-up) of the same D256 forward+backward, "
                                               "fp32 torch CPU autograd through oracle/vunet_oracle.py"}

    # The driver keeps `config`, `roofline`, `cpu_baseline`, `e2e`, `clocks` in full and drops unknown top-level keys:
    # the second workload (BASELINE configs[2], training) and the sequence-rendering number therefore ALSO travel
    # inside `roofline` (run-dependent results) while `config` stays a static description shared with --impl reference.
    if isinstance(train, dict) and "value" in train:
        fams = (train.get("kernels") or {}).get("families") or []
        worst = None
        for f in fams:  # worst family among those that carry >= 5 % of the step
            if f["share"] < 0.05:
                continue
            frac_t = f["tflops"] / peaks["bf16_tflops_sustained"]
            frac_h = f["gbs"] / peaks["hbm_gbs"]
            frac = max(frac_t, frac_h)
            if worst is None or frac < worst["frac"]:
                worst = {"family": f["family"], "share": f["share"], "ms": f["ms"], "tflops": f["tflops"],
                         "gbs": f["gbs"], "bound": "tensor" if frac_t >= frac_h else "hbm", "frac": round(frac, 4)}
        roofline["train"] = {
            "metric": train["metric"], "value": train["value"], "unit": train["unit"], "n_gpus": world,
            "ms_per_step": train["ms_per_step"], "steps": train["steps"], "warmup": train["warmup"],
            "samples_per_gpu_per_step": TRAIN_BATCH_PER_GPU, "e2e": train["e2e"],
            "allreduce_bytes_per_step": train["allreduce_bytes_per_step"], "gpu_launches": train["gpu_launches"],
            "cuda_graph": train["cuda_graph"], "conv_tflops_per_gpu": train["conv_tflops_per_gpu"],
            "frac_tensor_peak_sustained": train["frac_tensor_peak_sustained"], "worst_family": worst,
            "families": fams, "cpu_baseline": train.get("cpu_baseline"), "loss": train["loss"]}
    elif isinstance(train, dict):
        roofline["train"] = train
    # ---- the VunetOrg family (4-group autoregressive prior) on the same 64-frame synthesis workload ----
    try:
        roofline["vunet_org"] = run_vunet_org(device, Bf, K, peaks)
    except Exception as exc:  # secondary number: never take the headline down
        roofline["vunet_org"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    # ---- SURVEY §8f-4 / BASELINE configs[3]: behavior-net sampling, batch 256 ----
    try:
        roofline["behavior_sampling"] = run_behavior_sampling(device, K, peaks)
    except Exception as exc:  # secondary number: never take the headline down
        roofline["behavior_sampling"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    # ---- BASELINE configs[4]: sampled pose sequences -> stickmen -> video, everything on the GPU ----
    try:
        roofline["behavior_transfer_e2e"] = run_behavior_transfer(device, model, K)
    except Exception as exc:  # secondary number: never take the headline down
        roofline["behavior_transfer_e2e"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    # ---- the training step with the reference's real likelihood, the VGG19 perceptual loss ----
    if not args.no_train:
        try:
            roofline["train_vgg"] = run_train_vgg(device, K, peaks)
        except Exception as exc:  # secondary number: never take the headline down
            roofline["train_vgg"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    roofline["whole_step"] = whole
    roofline["sequence"] = sequence
    roofline["pose_rendering"] = pose
    roofline["e2e_fp32"] = e2e_fp32

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": static_config(Bf, world),
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / K},
        "gpu_launches": launches,
        "roofline": roofline,
        "whole_step": whole,
        "kernels": kernels,
        "cpu_baseline": cpu,
        "train": train,
        "sequence": sequence,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step workload (configs[2])")
    ap.add_argument("--frames-total", type=int, default=0,
                    help="strong-scaling reading of configs[1]: this many frames per step split over the GPUs "
                         "(default 0 = 64 frames per GPU, weak scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
