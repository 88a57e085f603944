"""CPU oracle for the VUnet hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional fp32 restatement (plain torch CPU ops on a reference-layout ``state_dict``) of the
algorithm in the reference's ``lib/modules.py:11-34,120-233`` and ``models/vunets.py:109-148,222-414,
426-597`` plus the KL of ``lib/losses.py:68-78,283-291`` — and of the ``VunetOrg`` family
(``models/vunets.py:18-106,151-219,600-783``, ``lib/losses.py:26-65``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
it — and there only as the checker / the timed CPU baseline, never on the product path.

Parity pinning: the reference ships no tests, golden vectors or known-answer fixtures for this path
(SURVEY.md §4, §8c), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: in the build
container ``oracle/gen_golden.py`` imports ``/root/reference``'s ``models.vunets.VunetAlter``, checks
this restatement against it (fp32, max rel-L2 ~1e-6) and writes ``tests/golden/*.pt``;
``tests/test_oracle.py`` re-checks the restatement against those committed vectors everywhere and
against the live reference when ``/root/reference`` exists.

Everything is written against a flat ``{name: tensor}`` state dict in the reference's checkpoint
layout (``<path>.beta, .gamma, .conv.bias, .conv.weight_g, .conv.weight_v``).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------------------
# operators
# --------------------------------------------------------------------------------------------------
def norm_conv(sd: SD, path: str, x: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """NormConv2d.forward — lib/modules.py:120-145.

    weight = g * v / ||v||_2 (norm over (cin, kh, kw) per output channel: old-style weight_norm, dim 0),
    out = gamma * (conv2d(x, weight) + bias) + beta.  Padding is k // 2 for every call site."""
    v = sd[path + ".conv.weight_v"]
    g = sd[path + ".conv.weight_g"]
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1, 1)
    w = v * (g / norm)
    out = F.conv2d(x, w, sd[path + ".conv.bias"], stride=stride, padding=v.shape[-1] // 2)
    return sd[path + ".gamma"] * out + sd[path + ".beta"]


def depth_to_space(x: torch.Tensor, bs: int = 2) -> torch.Tensor:
    """DepthToSpace — lib/modules.py:24-34: out[n,c,bs*h+i,bs*w+j] = in[n,(i*bs+j)*C'+c,h,w]."""
    n, c, h, w = x.shape
    cq = c // (bs * bs)
    out = x.new_empty(n, cq, h * bs, w * bs)
    for i in range(bs):
        for j in range(bs):
            out[:, :, i::bs, j::bs] = x[:, (i * bs + j) * cq:(i * bs + j + 1) * cq]
    return out


def rnb(sd: SD, path: str, x: torch.Tensor, a: Optional[torch.Tensor] = None) -> torch.Tensor:
    """VunetRNB.forward — lib/modules.py:221-233 (ELU alpha 1, dropout p = 0: SURVEY.md §3.3)."""
    r = x
    if a is not None:
        r = torch.cat([r, norm_conv(sd, path + ".nin", F.elu(a))], dim=1)
    return x + norm_conv(sd, path + ".conv", F.elu(r))


def upsample(sd: SD, path: str, x: torch.Tensor) -> torch.Tensor:
    """Upsample.forward, sub-pixel branch — lib/modules.py:179-182."""
    return depth_to_space(norm_conv(sd, path + ".up", x))


# --------------------------------------------------------------------------------------------------
# sub-networks
# --------------------------------------------------------------------------------------------------
def count_scales(sd: SD, prefix: str) -> int:
    """Number of scales of an EncUp/DecUp = (number of residual blocks) / 2."""
    idx = {int(k[len(prefix) + 8:].split(".")[0]) for k in sd if k.startswith(prefix + ".blocks.")}
    return (max(idx) + 1) // 2


def enc_up(sd: SD, prefix: str, x: torch.Tensor) -> List[torch.Tensor]:
    """EncUp.forward / DecUp.forward — models/vunets.py:133-148, 246-261."""
    n_scales = count_scales(sd, prefix)
    hs = []
    h = norm_conv(sd, prefix + ".nin", x)
    for i in range(n_scales):
        for n in range(2):
            h = rnb(sd, f"{prefix}.blocks.{2 * i + n}", h)
            hs.append(h)
        if i + 1 < n_scales:
            h = norm_conv(sd, f"{prefix}.downs.{i}.down", h, stride=2)
    return hs


def enc_down_alter(sd: SD, prefix: str, gs: List[torch.Tensor], eps: Sequence[torch.Tensor],
                   n_latent_scales: int = 2):
    """EncDownAlter.forward — models/vunets.py:558-597.  Pops from ``gs`` like the reference;
    ``eps[i]`` replaces the ``torch.randn_like`` draw of scale i."""
    hs, means, logstds, zs = [], [], [], []
    h = norm_conv(sd, prefix + ".nin", gs[-1])
    for i in range(n_latent_scales):
        h = rnb(sd, f"{prefix}.blocks.{2 * i}", h, gs.pop())
        hs.append(h)
        mu = norm_conv(sd, f"{prefix}.make_latent_params.{i}", h)
        logstd = torch.sigmoid(norm_conv(sd, f"{prefix}.make_logstds.{i}", h))
        z = eps[i] * torch.exp(logstd) + mu
        means.append(mu)
        logstds.append(logstd)
        zs.append(z)
        h = rnb(sd, f"{prefix}.blocks.{2 * i + 1}", h, torch.cat([gs.pop(), z], dim=1))
        hs.append(h)
        h = upsample(sd, f"{prefix}.ups.{i}", h)
    h = rnb(sd, prefix + ".fin_block", h, gs.pop())
    hs.append(h)
    return hs, means, logstds, zs


def dec_down_alter(sd: SD, prefix: str, gs: Sequence[torch.Tensor], zs: Sequence[torch.Tensor],
                   n_latent_scales: int = 2) -> torch.Tensor:
    """DecDownAlter.forward — models/vunets.py:332-414.  ``zs[i]`` is what the reference feeds
    ``auto_blocks[i]``: the posterior sample (forward), the posterior mean (transfer) or a prior draw
    (test_forward)."""
    gs = list(gs)
    zs = list(zs)
    n_scales = len(gs) // 2
    h = norm_conv(sd, prefix + ".nin", gs[-1])
    for i in range(n_scales):
        h = rnb(sd, f"{prefix}.blocks.{2 * i}", h, gs.pop())
        if i < n_latent_scales:
            h = rnb(sd, f"{prefix}.auto_blocks.{i}", h, zs.pop(0))
        h = rnb(sd, f"{prefix}.blocks.{2 * i + 1}", h, gs.pop())
        if i + 1 < n_scales:
            h = upsample(sd, f"{prefix}.ups.{i}", h)
    assert not gs and not zs
    return norm_conv(sd, prefix + ".out_conv", h)


# --------------------------------------------------------------------------------------------------
# whole model (VunetAlter, models/vunets.py:490-515)
# --------------------------------------------------------------------------------------------------
def forward(sd: SD, x: torch.Tensor, c: torch.Tensor, eps: Sequence[torch.Tensor], n_latent_scales: int = 2):
    """VunetAlter.forward: returns (imgs, means, logstds, zs, hs_left) — hs_left is what remains of
    eu's outputs after ed's pops (the reference returns that truncated list as ``activations[0]``)."""
    hs = enc_up(sd, "eu", x)
    _, means, logstds, zs = enc_down_alter(sd, "ed", hs, eps, n_latent_scales)
    gs = enc_up(sd, "du", c)
    imgs = dec_down_alter(sd, "dd", gs, zs, n_latent_scales)
    return imgs, means, logstds, zs, hs


def transfer(sd: SD, x: torch.Tensor, c: torch.Tensor, eps: Sequence[torch.Tensor], n_latent_scales: int = 2):
    """VunetAlter.transfer: decoder fed with the posterior means (still depends on eps[0])."""
    hs = enc_up(sd, "eu", x)
    _, means, _, _ = enc_down_alter(sd, "ed", hs, eps, n_latent_scales)
    gs = enc_up(sd, "du", c)
    return dec_down_alter(sd, "dd", gs, list(means), n_latent_scales)


def test_forward(sd: SD, c: torch.Tensor, prior: Sequence[torch.Tensor], n_latent_scales: int = 2):
    """VunetAlter.test_forward: ``prior[i]`` replaces the ``randn_like(h)`` draw of latent scale i."""
    gs = enc_up(sd, "du", c)
    return dec_down_alter(sd, "dd", gs, list(prior), n_latent_scales)


test_forward.__test__ = False  # not a pytest test


# --------------------------------------------------------------------------------------------------
# losses (lib/losses.py:68-78, 283-291)
# --------------------------------------------------------------------------------------------------
def kl_loss(mu: torch.Tensor, logstd: torch.Tensor) -> torch.Tensor:
    mu = mu.reshape(mu.shape[0], -1)
    logstd = logstd.reshape(logstd.shape[0], -1)
    dim = mu.shape[1]
    kl = torch.sum(-logstd + 0.5 * (torch.exp(logstd) ** 2 + mu ** 2), dim=-1) - 0.5 * dim
    return kl.mean()


def compute_kl_with_prior(means, logstds) -> torch.Tensor:
    return torch.stack([kl_loss(m, l) for m, l in zip(means, logstds)]).mean()


# --------------------------------------------------------------------------------------------------
# VunetOrg family (models/vunets.py:18-106, 151-219, 600-783; lib/modules.py:11-21; lib/losses.py:26-65)
# --------------------------------------------------------------------------------------------------
def space_to_depth(x: torch.Tensor, bs: int = 2) -> torch.Tensor:
    """SpaceToDepth — lib/modules.py:11-21: out[n,(i*bs+j)*C+c,h,w] = in[n,c,bs*h+i,bs*w+j] (inverse of
    :func:`depth_to_space`)."""
    return torch.cat([x[:, :, i::bs, j::bs] for i in range(bs) for j in range(bs)], dim=1)


def enc_down(sd: SD, prefix: str, gs: List[torch.Tensor], eps: Sequence[torch.Tensor], n_latent_scales: int = 2):
    """EncDown.forward — models/vunets.py:191-219.  Pops from ``gs`` like the reference; ``eps[i]`` replaces the
    ``randn_like`` draw of ``latent_sample`` (151-156: z = q + 1.0 * eps) at scale i."""
    hs, qs, zs = [], [], []
    h = norm_conv(sd, prefix + ".nin", gs[-1])
    for i in range(n_latent_scales):
        h = rnb(sd, f"{prefix}.blocks.{2 * i}", h, gs.pop())
        hs.append(h)
        q = norm_conv(sd, f"{prefix}.make_latent_params.{i}", h)
        qs.append(q)
        z = q + eps[i]
        zs.append(z)
        h = rnb(sd, f"{prefix}.blocks.{2 * i + 1}", h, torch.cat([gs.pop(), z], dim=1))
        hs.append(h)
        h = upsample(sd, f"{prefix}.ups.{i}", h)
    h = rnb(sd, prefix + ".fin_block", h, gs.pop())
    hs.append(h)
    return hs, qs, zs


def dec_down(sd: SD, prefix: str, gs: Sequence[torch.Tensor], zs_posterior: Sequence[torch.Tensor], training: bool,
             eps_groups: Sequence[Sequence[torch.Tensor]], n_latent_scales: int = 2):
    """DecDown.forward — models/vunets.py:696-772: the decoder with the 4-group autoregressive prior.
    ``eps_groups[i][l]`` replaces the ``randn_like`` draw of ``latent_sample(p_group)`` for group l of latent
    scale i (drawn in BOTH modes: with ``training`` the sample is only returned, the posterior groups are fed back).
    Returns (imgs, hs, ps, zs_prior)."""
    gs = list(gs)
    zs_posterior = list(zs_posterior)
    n_scales = len(gs) // 2
    hs, ps, zs = [], [], []
    h = norm_conv(sd, prefix + ".nin", gs[-1])
    for i in range(n_scales):
        h = rnb(sd, f"{prefix}.blocks.{2 * i}", h, gs.pop())
        hs.append(h)
        if i < n_latent_scales:
            scale = f"l_{i}"
            if training:
                cq = zs_posterior[0].shape[1]
                post_groups = list(torch.split(space_to_depth(zs_posterior[0]), cq, dim=1))
            p_groups, z_groups = [], []
            p_features = space_to_depth(rnb(sd, f"{prefix}.auto_blocks.{scale}.0", h))
            for l in range(4):
                p_group = norm_conv(sd, f"{prefix}.auto_lp.{scale}.{l}", p_features)
                p_groups.append(p_group)
                z_group = p_group + eps_groups[i][l]
                z_groups.append(z_group)
                feedback = post_groups.pop(0) if training else z_group
                if l + 1 < 4:
                    p_features = rnb(sd, f"{prefix}.auto_blocks.{scale}.{l + 1}", p_features, feedback)
            ps.append(depth_to_space(torch.cat(p_groups, dim=1)))
            z_prior = depth_to_space(torch.cat(z_groups, dim=1))
            zs.append(z_prior)
            z = zs_posterior.pop(0) if training else z_prior
            h = norm_conv(sd, f"{prefix}.latent_nins.{scale}", torch.cat([h, z], dim=1))
        h = rnb(sd, f"{prefix}.blocks.{2 * i + 1}", h, gs.pop())
        hs.append(h)
        if i + 1 < n_scales:
            h = upsample(sd, f"{prefix}.ups.{i}", h)
    assert not gs
    if training:
        assert not zs_posterior
    return norm_conv(sd, prefix + ".out_conv", hs[-1]), hs, ps, zs


def org_forward(sd: SD, x, c, eps, eps_groups, n_latent_scales: int = 2):
    """VunetOrg.forward — models/vunets.py:83-94: returns (imgs, qs, ps, zs_posterior, zs_prior, hs_left)."""
    hs = enc_up(sd, "eu", x)
    _, qs, zs_post = enc_down(sd, "ed", hs, eps, n_latent_scales)
    gs = enc_up(sd, "du", c)
    imgs, _, ps, zs_prior = dec_down(sd, "dd", gs, zs_post, True, eps_groups, n_latent_scales)
    return imgs, qs, ps, zs_post, zs_prior, hs


def org_transfer(sd: SD, x, c, eps, eps_groups, n_latent_scales: int = 2):
    """VunetOrg.transfer — models/vunets.py:102-110: the decoder is fed the posterior MEANS."""
    hs = enc_up(sd, "eu", x)
    _, qs, _ = enc_down(sd, "ed", hs, eps, n_latent_scales)
    gs = enc_up(sd, "du", c)
    return dec_down(sd, "dd", gs, list(qs), True, eps_groups, n_latent_scales)[0]


def org_test_forward(sd: SD, c, eps_groups, n_latent_scales: int = 2):
    """VunetOrg.test_forward — models/vunets.py:96-100: appearance sampled group by group from the prior."""
    gs = enc_up(sd, "du", c)
    return dec_down(sd, "dd", gs, [], False, eps_groups, n_latent_scales)[0]


def latent_kl(prior_mean: torch.Tensor, posterior_mean: torch.Tensor) -> torch.Tensor:
    """lib/losses.py:26-37."""
    return (0.5 * (prior_mean - posterior_mean) ** 2).sum(dim=[1, 2, 3]).mean()


def compute_kl_loss(prior_means, posterior_means) -> torch.Tensor:
    """lib/losses.py:55-65: sum over the latent scales."""
    return torch.stack([latent_kl(p, q) for p, q in zip(prior_means, posterior_means)]).sum()


def make_group_eps(cfg: dict, batch: int, tag: str = ""):
    """Deterministic draws for the 4 prior groups of every latent scale (shape = a quarter-resolution latent)."""
    ns, _ = n_scales_of(cfg)
    shp = latent_shapes(batch, cfg["spatial_size"], ns, cfg["nf_max"], cfg["n_latent_scales"])
    return [[det_normal(f"geps{i}.{l}{tag}", (b, ch, h // 2, w // 2)) for l in range(4)]
            for i, (b, ch, h, w) in enumerate(shp)]


# --------------------------------------------------------------------------------------------------
# deterministic, portable test data (integer hash -> floats; no dependence on any RNG implementation)
# --------------------------------------------------------------------------------------------------
def _hash01(n: int, seed: int) -> np.ndarray:
    """n floats in [0, 1) from a 32-bit integer mix of (index, seed); exact on every platform."""
    i = np.arange(n, dtype=np.uint64)
    x = (i * np.uint64(2654435761) + np.uint64(seed) * np.uint64(40503) + np.uint64(12345)) & np.uint64(0xFFFFFFFF)
    for _ in range(2):
        x ^= x >> np.uint64(16)
        x = (x * np.uint64(0x45D9F3B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return (x.astype(np.float64) / 4294967296.0)


def _name_seed(name: str) -> int:
    import zlib
    return zlib.crc32(name.encode("utf-8")) & 0x7FFFFFFF


def det_uniform(name: str, shape, lo: float = -1.0, hi: float = 1.0) -> torch.Tensor:
    n = int(np.prod(shape))
    u = _hash01(n, _name_seed(name))
    return torch.from_numpy((lo + (hi - lo) * u).astype(np.float32)).reshape(tuple(shape))


def det_normal(name: str, shape) -> torch.Tensor:
    """Approximately N(0,1): centred sum of 12 hashed uniforms."""
    n = int(np.prod(shape))
    acc = np.zeros(n, dtype=np.float64)
    for k in range(12):
        acc += _hash01(n, _name_seed(f"{name}#{k}"))
    return torch.from_numpy((acc - 6.0).astype(np.float32)).reshape(tuple(shape))


def det_state_dict(shapes: Dict[str, Sequence[int]]) -> SD:
    """Deterministic non-trivial parameters for a reference-layout state dict: v ~ U(+-1/sqrt(fan_in)),
    g = ||v|| * U(.7,1.3), gamma ~ U(.7,1.3), beta ~ U(-.1,.1), bias ~ U(+-1/sqrt(fan_in)) — so the
    affine/weight-norm path is exercised (default init has gamma=1, beta=0, g=||v||)."""
    sd: SD = {}
    for name, shape in shapes.items():
        if name.endswith("weight_v"):
            bound = 1.0 / float(np.sqrt(shape[1] * shape[2] * shape[3]))
            sd[name] = det_uniform(name, shape, -bound, bound)
    for name, shape in shapes.items():
        if name.endswith("weight_v"):
            continue
        if name.endswith("weight_g"):
            v = sd[name[:-1] + "v"]
            sd[name] = v.reshape(v.shape[0], -1).norm(dim=1).reshape(shape) * det_uniform(name, shape, 0.7, 1.3)
        elif name.endswith("gamma"):
            sd[name] = det_uniform(name, shape, 0.7, 1.3)
        elif name.endswith("beta"):
            sd[name] = det_uniform(name, shape, -0.1, 0.1)
        elif name.endswith("bias"):
            v = sd[name[:-4] + "weight_v"]
            bound = 1.0 / float(np.sqrt(v.shape[1] * v.shape[2] * v.shape[3]))
            sd[name] = det_uniform(name, shape, -bound, bound)
        else:
            raise KeyError(f"unexpected parameter {name}")
    return {k: sd[k] for k in shapes}


def latent_shapes(batch: int, spatial_size: int, n_scales: int, nf_max: int, n_latent_scales: int = 2):
    """Shapes of the latent tensors (and of the eps draws) per latent scale."""
    h0 = spatial_size // 2 ** (n_scales - 1)
    return [(batch, nf_max, h0 * 2 ** i, h0 * 2 ** i) for i in range(n_latent_scales)]


def n_scales_of(cfg: dict) -> Tuple[int, int]:
    """(n_scales, n_scales_x) — models/vunets.py:430-441."""
    ns = (1 + int(np.round(np.log2(cfg["spatial_size"]))) - cfg["bottleneck_factor"]
          if cfg["n_scales"] < 6 else cfg["n_scales"])
    nx = ns - cfg["box_factor"] if cfg.get("n_channels_x", 3) > 3 else ns
    return ns, nx


def make_inputs(cfg: dict, batch: int, tag: str = ""):
    """Deterministic x (appearance), c (stickman), eps and prior draws for a config."""
    ns, nx = n_scales_of(cfg)
    s = cfg["spatial_size"]
    sx = s // 2 ** (ns - nx)
    x = det_uniform(f"x{tag}", (batch, cfg.get("n_channels_x", 3), sx, sx))
    c = det_uniform(f"c{tag}", (batch, 3, s, s))
    shp = latent_shapes(batch, s, ns, cfg["nf_max"], cfg["n_latent_scales"])
    eps = [det_normal(f"eps{i}{tag}", sh) for i, sh in enumerate(shp)]
    prior = [det_normal(f"prior{i}{tag}", sh) for i, sh in enumerate(shp)]
    return x, c, eps, prior
