"""Generates ``tests/golden/*.pt`` from the UNMODIFIED reference (``/root/reference``) — test
infrastructure, run in the build container only (the GPU box has no reference tree).

For each config it
  1. builds the reference ``models.vunets.VunetAlter`` and loads deterministic hashed parameters
     (``vunet_oracle.det_state_dict``; the fixture stores only names+shapes, the values are recomputed
     from the same integer hash wherever the tests run),
  2. runs the reference ``forward`` / ``transfer`` / ``test_forward`` with the latent noise injected by
     temporarily replacing ``torch.randn_like`` (bit-identical to seeding),
  3. back-propagates a smooth surrogate loss (MSE to a hashed target + 1e-3 * KL) through the
     reference for gradient fixtures,
  4. asserts that the CPU oracle (``oracle/vunet_oracle.py``) reproduces every tensor, and
  5. writes the small result tensors.

Usage:  python oracle/gen_golden.py [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import vunet_oracle as O  # noqa: E402

ARCH = dict(n_latent_scales=2, conv_layer_type="l1", subpixel_upsampling=True, n_scales=0)

CONFIGS = {
    # small nets (nf 32..64, 32x32) that the CUDA kernels support and the oracle runs in < 1 s
    "tiny3": dict(ARCH, spatial_size=32, bottleneck_factor=1, box_factor=1, n_channels_x=3, nf_start=32,
                  nf_max=64, batch=2),
    "tiny30": dict(ARCH, spatial_size=32, bottleneck_factor=1, box_factor=1, n_channels_x=30, nf_start=32,
                   nf_max=64, batch=3),
    # BASELINE.json configs[0]: Market 128x128, bottleneck_factor 1, box_factor 1, 30 appearance channels
    "market128": dict(ARCH, spatial_size=128, bottleneck_factor=1, box_factor=1, n_channels_x=30, nf_start=32,
                      nf_max=128, batch=1),
}

FULL_GRAD_KEYS = ("dd.out_conv.conv.weight_v", "dd.out_conv.gamma", "eu.nin.conv.weight_v",
                  "ed.make_logstds.0.gamma", "ed.make_latent_params.1.beta", "du.blocks.3.conv.conv.weight_g",
                  "dd.auto_blocks.0.nin.conv.bias")


@contextlib.contextmanager
def injected_randn(draws):
    """Make ``torch.randn_like`` hand out the given tensors in order (the reference's only RNG use)."""
    it = iter(draws)
    orig = torch.randn_like

    def fake(t, *a, **k):
        e = next(it)
        assert e.shape == t.shape, (e.shape, t.shape)
        return e.to(t.dtype)

    torch.randn_like = fake
    try:
        yield
    finally:
        torch.randn_like = orig


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def build_reference(ref_root: str, cfg: dict):
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    from models.vunets import VunetAlter  # the reference's own class
    with contextlib.redirect_stdout(io.StringIO()):
        model = VunetAlter(n_channels_x=cfg["n_channels_x"], **{k: v for k, v in cfg.items() if k != "n_channels_x"})
    return model


def generate(name: str, cfg: dict, ref_root: str, out_dir: str) -> None:
    torch.manual_seed(0)
    torch.set_grad_enabled(True)
    model = build_reference(ref_root, cfg)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = O.det_state_dict(shapes)
    model.load_state_dict(sd, strict=True)
    B = cfg["batch"]
    x, c, eps, prior = O.make_inputs(cfg, B)
    target = O.det_uniform("target", (B, 3, cfg["spatial_size"], cfg["spatial_size"]))

    # ---- reference ----
    model.zero_grad()
    with injected_randn(eps):
        imgs, means, logstds, acts = model(x, c)
    # the reference's KL (lib/losses.py is not importable here: restated in the oracle, §8c)
    kl = O.compute_kl_with_prior(means, logstds)
    loss = torch.mean((imgs - target) ** 2) + 1e-3 * kl
    loss.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in model.named_parameters()}
    with torch.no_grad():
        with injected_randn(eps):
            timgs = model.transfer(x, c)
        with injected_randn(prior):
            pimgs = model.test_forward(c)
        with injected_randn(eps):
            ed_hs, ed_means, ed_logstds, ed_zs = model.ed(model.eu(x))

    # ---- oracle must agree ----
    with torch.no_grad():
        o_imgs, o_means, o_logstds, o_zs, o_hs = O.forward(sd, x, c, eps)
        o_t = O.transfer(sd, x, c, eps)
        o_p = O.test_forward(sd, c, prior)
    errs = {"imgs": rel(o_imgs, imgs.detach()), "transfer": rel(o_t, timgs), "test_forward": rel(o_p, pimgs)}
    for i in range(len(means)):
        errs[f"mean{i}"] = rel(o_means[i], means[i].detach())
        errs[f"logstd{i}"] = rel(o_logstds[i], logstds[i].detach())
        errs[f"z{i}"] = rel(o_zs[i], ed_zs[i])
    assert len(o_hs) == len(acts[0]), (len(o_hs), len(acts[0]))
    for i, (a, b) in enumerate(zip(o_hs, acts[0])):
        errs[f"hs{i}"] = rel(a, b.detach())
    sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    g_imgs, g_means, g_logstds, _, _ = O.forward(sd_g, x, c, eps)
    o_loss = torch.mean((g_imgs - target) ** 2) + 1e-3 * O.compute_kl_with_prior(g_means, g_logstds)
    o_loss.backward()
    num = den = 0.0
    for k, g in grads.items():
        og = sd_g[k].grad
        if g is None:
            assert og is None or float(og.abs().max()) == 0.0, k
            continue
        num += float((og - g).pow(2).sum())
        den += float(g.pow(2).sum())
    errs["grad_concat"] = (num / den) ** 0.5
    errs["loss"] = abs(float(o_loss) - float(loss)) / abs(float(loss))
    worst = max(errs.values())
    print(f"[{name}] oracle vs reference: worst rel err {worst:.3e}  ({max(errs, key=errs.get)})")
    assert worst < 2e-5, errs

    fixture = {
        "name": name,
        "config": cfg,
        "param_shapes": list(shapes.items()),
        "imgs": imgs.detach().clone(),
        "means": [m.detach().clone() for m in means],
        "logstds": [l.detach().clone() for l in logstds],
        "zs": [z.clone() for z in ed_zs],
        "transfer": timgs.clone(),
        "test_forward": pimgs.clone(),
        "n_hs_left": len(acts[0]),
        "hs_left_shapes": [tuple(a.shape) for a in acts[0]],
        "hs_left_mean_abs": [float(a.detach().abs().mean()) for a in acts[0]],
        "ed_hs_shapes": [tuple(a.shape) for a in ed_hs],
        "kl": float(kl),
        "loss": float(loss),
        "grad_is_none": sorted(k for k, g in grads.items() if g is None),
        "grad_norms": {k: float(g.norm()) for k, g in grads.items() if g is not None},
        "grad_full": {k: grads[k].clone() for k in FULL_GRAD_KEYS if grads.get(k) is not None},
        "oracle_vs_reference": errs,
        "torch_version": torch.__version__,
    }
    if name == "market128":  # keep the big fixture small: no z/test_forward duplicates needed
        for k in ("zs", "test_forward"):
            fixture.pop(k)
    path = os.path.join(out_dir, f"{name}.pt")
    torch.save(fixture, path)
    print(f"[{name}] wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), kl={float(kl):.4f} loss={float(loss):.5f}, "
          f"{len(fixture['grad_is_none'])} params without grad")


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    for name, cfg in CONFIGS.items():
        if args.only and name != args.only:
            continue
        generate(name, dict(cfg), args.ref, args.out)


if __name__ == "__main__":
    main()
