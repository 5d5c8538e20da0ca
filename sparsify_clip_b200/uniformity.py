"""W2 uniformity metrics with the call signatures of the reference's uniformity.py
(torch_uniformity1 :6, torch_uniformity :53, numpy_uniformity :101,
torch_uniformity_equivalent :138, uniformity10 :182).

Evaluation-only and cold: one D x B . B x D covariance (<= 0.1 % of a B x B x D contraction; on CUDA tensors
the library's scb_gram_dd kernel) plus a D x D eigen-decomposition, which is a cuSOLVER library call.
All five variants share one implementation of
    W2 = sqrt(|mu|^2 + 1 + tr(Sigma) - (2/sqrt(D)) tr(Sigma^(1/2)))
and differ only in the details the reference differs in (sign, epsilon, decomposition).
"""
import math

import torch


def _moments(x):
    """(mean [1, D], covariance [D, D]); CUDA inputs run on the library's column-sum and D x D second-moment kernels
    (scb_col_sum, scb_gram_dd: fp32 FMAs, fixed-order reductions), host inputs on the reference's own two lines."""
    n = x.size(0)
    if x.is_cuda and x.dim() == 2:
        from .backend_cuda import get_backend
        be = get_backend()
        xp = be.prep(x)
        mu = be.col_sum(xp, None, 1.0 / n)
        return mu[None, :], be.gram_dd(xp, mu, 1.0 / n)
    mu = x.mean(dim=0, keepdim=True)
    xc = x - mu
    return mu, torch.mm(xc.t(), xc) / n


def _w2(mu, sigma, tr_sqrt, tr_sigma=None):
    d = sigma.shape[0]
    tr_sigma = torch.trace(sigma) if tr_sigma is None else tr_sigma
    return torch.sqrt((mu * mu).sum() + 1 + tr_sigma - (2.0 / math.sqrt(d)) * tr_sqrt)


def torch_uniformity1(features_modality1):
    mu, sigma = _moments(features_modality1)
    u, s, _ = torch.linalg.svd(sigma)
    root = u @ torch.diag(torch.sqrt(torch.clamp(s + 1e-8, min=0))) @ u.T
    return _w2(mu, sigma, torch.trace(root), torch.clamp(torch.trace(sigma), min=0))


def torch_uniformity(features_modality1, features_modality2):
    mu, sigma = _moments(torch.cat([features_modality1, features_modality2], dim=0))
    sigma = sigma + 1e-6
    w, v = torch.linalg.eigh(sigma)
    root = v @ torch.diag(torch.sqrt(torch.clamp(w + 1e-8, min=0))) @ v.T
    return -_w2(mu, sigma, torch.trace(root))


def numpy_uniformity(features_modality1, features_modality2):
    import numpy as np
    x = torch.cat([features_modality1, features_modality2], dim=0)
    if x.is_cuda:                      # GPU-resident: eigenvalues only, no NumPy round trip (see metrics.uniformity)
        from .metrics import uniformity as _gpu_uniformity
        return _gpu_uniformity(features_modality1, features_modality2)
    mu, sigma = _moments(x)
    cov = sigma.detach().cpu().numpy()
    m = mu.detach().cpu().numpy().ravel()
    w, q = np.linalg.eig(cov)
    root = q @ np.sqrt(np.diag((w + 1e-8).clip(min=0))) @ q.T
    val = float(np.sum(m * m)) + 1 + float(np.trace(cov - 2.0 / np.sqrt(x.size(1)) * root).real)
    return -math.sqrt(val)


def torch_uniformity_equivalent(features_modality1):
    mu, sigma = _moments(features_modality1)
    w, v = torch.linalg.eig(sigma)
    w, v = w.real + 1e-8, v.real
    root = v @ torch.sqrt(torch.diag(torch.clamp(w, min=0))) @ v.t()
    return _w2(mu, sigma, torch.trace(root))


def uniformity10(z1):
    mu, sigma = _moments(z1)
    w, v = torch.linalg.eig(sigma)
    w, v = torch.abs(w), torch.abs(v)
    root = v @ torch.sqrt(torch.diag(w)) @ v.T
    return _w2(mu, sigma, torch.trace(root))
