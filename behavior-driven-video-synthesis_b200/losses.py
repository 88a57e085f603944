"""KL terms of the VUnet objective — mirror of the reference's ``lib/losses.py:68-78, 283-291`` (``VunetAlter``)
and ``lib/losses.py:26-65`` (``VunetOrg``: ``latent_kl`` / ``compute_kl_loss`` between prior and posterior means).

``kl_loss(mu, logstd)``: ``mean_b( sum_d(-logstd + 0.5*(exp(logstd)^2 + mu^2)) - 0.5*D )`` and
``compute_kl_with_prior(means, logstds)``: the mean of ``kl_loss`` over latent scales.

When ``means`` come from this package's forward pass they carry the per-block partial sums the fused
latent kernel already produced (``vub_latent_fwd``), so no second pass over the latents is made; for
foreign tensors the same formula is evaluated with torch ops on the tensors' device.
"""
from __future__ import annotations

import torch


def kl_loss(mu: torch.Tensor, logstd: torch.Tensor) -> torch.Tensor:
    part = getattr(mu, "_vub_kl_partial", None)
    differentiable = torch.is_grad_enabled() and (mu.requires_grad or logstd.requires_grad)
    if part is not None and getattr(mu, "_vub_logstd", None) is logstd and not differentiable:
        b = mu.shape[0]
        dim = mu.numel() // b
        return part.sum() / b - 0.5 * dim
    mu2 = mu.reshape(mu.size(0), -1)
    ls2 = logstd.reshape(logstd.size(0), -1)
    dim = mu2.shape[1]
    std = torch.exp(ls2)
    kl = torch.sum(-ls2 + 0.5 * (std ** 2 + mu2 ** 2), dim=-1) - (0.5 * dim)
    return kl.mean()


def compute_kl_with_prior(means, logstds) -> torch.Tensor:
    terms = [kl_loss(m, l).unsqueeze(-1) for m, l in zip(means, logstds)]
    return torch.mean(torch.cat(terms, dim=-1))


def latent_kl(prior_mean: torch.Tensor, posterior_mean: torch.Tensor) -> torch.Tensor:
    """lib/losses.py:26-37: ``mean_b( sum_{c,h,w} 0.5 * (prior_mean - posterior_mean)^2 )`` (unit variances)."""
    kl = 0.5 * torch.pow(prior_mean - posterior_mean, 2)
    return torch.mean(torch.sum(kl, dim=[1, 2, 3]))


def compute_kl_loss(prior_means, posterior_means) -> torch.Tensor:
    """lib/losses.py:55-65: the SUM of ``latent_kl`` over the latent scales (``VunetOrg``'s ``ps`` and ``qs``)."""
    return torch.sum(torch.cat([latent_kl(p, q).unsqueeze(dim=-1) for p, q in zip(prior_means, posterior_means)],
                               dim=-1))
