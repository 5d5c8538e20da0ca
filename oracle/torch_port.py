"""TEST INFRASTRUCTURE ONLY -- torch (CPU) port of the reference's op sequence.

The reference's arithmetic lives in third-party PyTorch eager ops (pinned
pytorch=2.5.1, environment.yml:132; here torch 2.11).  This module issues the same
ops in the same order as the reference call sites, so that timing it on the host
cores is a fair stand-in for "the reference's own CPU path" on a box where
/root/reference does not exist (bench.py cpu_baseline, kind "port"), and so that
autograd through it is a second opinion for oracle/closed_form.py.
"""
import torch
import torch.nn.functional as F


def anchor(I, T, temperature=0.07):
    # sparsify_clip.py:119-132 -- mm, div, two cross-entropies against arange
    z = torch.div(torch.mm(I, T.t()), temperature)
    tgt = torch.arange(I.size(0), device=z.device)
    return (F.cross_entropy(z, tgt) + F.cross_entropy(z.t(), tgt)) / 2


def lunif(x, t=2):
    # sparsify_clip.py:161,164 -- pdist -> square -> *(-t) -> exp -> mean -> log
    return torch.pdist(x, p=2).pow(2).mul(-t).exp().mean().log()


def lalign(x, y, alpha=2):
    # sparsify_clip.py:187
    return (x - y).norm(dim=1).pow(alpha).mean()


def centroids(a, b):
    # sparsify_clip.py:353 then F.normalize at the call site (:804)
    return F.normalize((a + b) / 2.0, dim=-1)


def weighted(I, T, temperature, w_anchor, w_align, w_ui, w_ut, w_uc):
    """Same composition as oracle.closed_form.weighted_loss, built from the ops above."""
    loss = 0
    if w_anchor != 0.0:
        loss = loss + w_anchor * anchor(I, T, temperature)
    if w_align != 0.0:
        loss = loss + w_align * lalign(I, T)
    if w_ui != 0.0:
        loss = loss + w_ui * lunif(I)
    if w_ut != 0.0:
        loss = loss + w_ut * lunif(T)
    if w_uc != 0.0:
        loss = loss + w_uc * lunif(centroids(I, T))
    return loss


def fwd_bwd(I, T, temperature, weights):
    """One forward+backward; returns (loss, dI, dT, dtau or None)."""
    I = I.detach().clone().requires_grad_(True)
    T = T.detach().clone().requires_grad_(True)
    tau = temperature
    if isinstance(tau, torch.Tensor):
        tau = tau.detach().clone().requires_grad_(True)
    loss = weighted(I, T, tau, *weights)
    loss.backward()
    return (loss.detach(), I.grad, T.grad,
            tau.grad if isinstance(tau, torch.Tensor) else None)
