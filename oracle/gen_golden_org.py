"""Generates ``tests/golden/org_*.pt`` from the UNMODIFIED reference ``models.vunets.VunetOrg`` — test
infrastructure, run in the build container only (the GPU box has no reference tree).

Same recipe as ``gen_golden.py``: deterministic hashed parameters, latent noise injected by replacing
``torch.randn_like`` (the reference's only RNG use; draw order: ``ed`` one per latent scale, then ``dd`` four
quarter-resolution group draws per latent scale — also in training mode, where the prior sample is only returned),
reference ``forward`` / ``transfer`` / ``test_forward`` and a surrogate loss (MSE + 1e-3 * the reference's
``compute_kl_loss``, lib/losses.py:55-65, restated) back-propagated for gradient fixtures; the CPU oracle must
reproduce every tensor before the fixture is written.

Usage:  python oracle/gen_golden_org.py [--ref /root/reference]
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
from gen_golden import ARCH, injected_randn, rel  # noqa: E402

CONFIGS = {
    "org_tiny3": dict(ARCH, spatial_size=32, bottleneck_factor=1, box_factor=1, n_channels_x=3, nf_start=32, nf_max=64,
                      batch=2),
    "org_tiny30": dict(ARCH, spatial_size=32, bottleneck_factor=1, box_factor=1, n_channels_x=30, nf_start=32,
                       nf_max=64, batch=3),
    # config/vunet.yaml geometry at 128x128: nf_max 128 -> the 4-group prior runs on 512 / 1024-channel convs
    "org_m128": dict(ARCH, spatial_size=128, bottleneck_factor=1, box_factor=1, n_channels_x=30, nf_start=32,
                     nf_max=128, batch=1),
}

FULL_GRAD_KEYS = ("dd.out_conv.conv.weight_v", "eu.nin.conv.weight_v", "ed.make_latent_params.1.beta",
                  "dd.auto_lp.l_0.2.gamma", "dd.auto_lp.l_1.3.conv.bias", "dd.auto_blocks.l_1.1.nin.conv.weight_g",
                  "dd.latent_nins.l_0.conv.weight_g", "dd.auto_blocks.l_0.0.conv.conv.bias")


def build_reference(ref_root: str, cfg: dict):
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    from models.vunets import VunetOrg  # the reference's own class
    with contextlib.redirect_stdout(io.StringIO()):
        return VunetOrg(n_channels_x=cfg["n_channels_x"], **{k: v for k, v in cfg.items() if k != "n_channels_x"})


def generate(name: str, cfg: dict, ref_root: str, out_dir: str) -> None:
    torch.manual_seed(0)
    torch.set_grad_enabled(True)
    model = build_reference(ref_root, cfg)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = O.det_state_dict(shapes)
    model.load_state_dict(sd, strict=True)
    B = cfg["batch"]
    x, c, eps, _ = O.make_inputs(cfg, B)
    geps = O.make_group_eps(cfg, B)
    flat_g = [e for grp in geps for e in grp]
    target = O.det_uniform("target", (B, 3, cfg["spatial_size"], cfg["spatial_size"]))

    # ---- reference ----
    model.zero_grad()
    with injected_randn(list(eps) + flat_g):
        imgs, qs, ps, acts = model(x, c)
    kl = O.compute_kl_loss(ps, qs)
    loss = torch.mean((imgs - target) ** 2) + 1e-3 * kl
    loss.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in model.named_parameters()}
    with torch.no_grad():
        with injected_randn(list(eps) + flat_g):
            timgs = model.transfer(x, c)
        with injected_randn(flat_g):
            pimgs = model.test_forward(c)
        with injected_randn(list(eps)):
            _, ed_qs, ed_zs = model.ed(model.eu(x))
        with injected_randn(flat_g):
            _, _, t_ps, t_zs = model.dd(model.du(c), [], training=False)

    # ---- oracle must agree ----
    with torch.no_grad():
        o_imgs, o_qs, o_ps, o_zpost, o_zprior, o_hs = O.org_forward(sd, x, c, eps, geps)
        o_t = O.org_transfer(sd, x, c, eps, geps)
        o_p = O.org_test_forward(sd, c, geps)
        _, _, o_tps, o_tzs = O.dec_down(sd, "dd", O.enc_up(sd, "du", c), [], False, geps)
    errs = {"imgs": rel(o_imgs, imgs.detach()), "transfer": rel(o_t, timgs), "test_forward": rel(o_p, pimgs)}
    for i in range(len(qs)):
        errs[f"q{i}"] = rel(o_qs[i], qs[i].detach())
        errs[f"p{i}"] = rel(o_ps[i], ps[i].detach())
        errs[f"z{i}"] = rel(o_zpost[i], ed_zs[i])
        errs[f"test_p{i}"] = rel(o_tps[i], t_ps[i])
        errs[f"test_z{i}"] = rel(o_tzs[i], t_zs[i])
    assert len(o_hs) == len(acts[0]), (len(o_hs), len(acts[0]))
    sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    g_imgs, g_qs, g_ps, _, _, _ = O.org_forward(sd_g, x, c, eps, geps)
    o_loss = torch.mean((g_imgs - target) ** 2) + 1e-3 * O.compute_kl_loss(g_ps, g_qs)
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
        "qs": [q.detach().clone() for q in qs],
        "ps": [p.detach().clone() for p in ps],
        "zs": [z.clone() for z in ed_zs],
        "transfer": timgs.clone(),
        "test_forward": pimgs.clone(),
        "test_ps": [p.clone() for p in t_ps],
        "test_zs": [z.clone() for z in t_zs],
        "n_hs_left": len(acts[0]),
        "n_activations": [len(a) for a in acts],
        "kl": float(kl),
        "loss": float(loss),
        "grad_is_none": sorted(k for k, g in grads.items() if g is None),
        "grad_norms": {k: float(g.norm()) for k, g in grads.items() if g is not None},
        "grad_full": {k: grads[k].clone() for k in FULL_GRAD_KEYS if grads.get(k) is not None},
        "oracle_vs_reference": errs,
        "torch_version": torch.__version__,
    }
    if name == "org_m128":  # keep the big fixture small
        for k in ("zs", "test_forward", "test_ps", "test_zs"):
            fixture.pop(k)
    path = os.path.join(out_dir, f"{name}.pt")
    torch.save(fixture, path)
    print(f"[{name}] wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), kl={float(kl):.4f} loss={float(loss):.5f}, "
          f"{len(fixture['grad_is_none'])} params without grad, {len(shapes)} tensors")


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
