"""TEST INFRASTRUCTURE ONLY -- the closed forms of oracle/closed_form.py, restated in torch fp64 and
evaluated in ROW SLABS so that they run at BASELINE sizes (B = 32768: the dense B x B fp64 matrices of the numpy oracle
would need 8.6 GB each).  Runs on whatever device the inputs live on (the GPU box: fp64 on the B200; here: CPU).

Reference lines restated: sparsify_clip.py:110-132 (anchor), :159-164 (L_unif, strict i<j pairs, t=2), :186-187
(L_align), :334-355 + :804 (centroids + F.normalize eps 1e-12).  Gradients are returned for a SAMPLE of rows only
(every row needs a full pass; 64 rows are what the full-size parity test checks).  Pinned against
oracle/closed_form.py (itself pinned against the unmodified reference) by tests/test_oracle.py::test_chunked_oracle_*.
"""
import math

import torch

F64 = torch.float64


def _slabs(n, slab):
    for lo in range(0, n, slab):
        yield lo, min(n, lo + slab)


def anchor(I, T, tau, rows, slab=2048):
    """-> (loss, dI[rows], dT[rows], dtau); I, T fp64 [B, D]; rows: 1-D LongTensor of sampled row indices."""
    B = I.shape[0]
    dev = I.device
    r = torch.empty(B, dtype=F64, device=dev)
    cm = torch.full((B,), -math.inf, dtype=F64, device=dev)       # online column max / sum
    cl = torch.zeros(B, dtype=F64, device=dev)
    for lo, hi in _slabs(B, slab):
        S = (I[lo:hi] @ T.t()) / tau
        r[lo:hi] = torch.logsumexp(S, dim=1)
        m = torch.maximum(cm, S.max(dim=0).values)
        cl = cl * torch.exp(cm - m) + torch.exp(S - m).sum(dim=0)
        cm = m
    c = cm + torch.log(cl)
    d = (I * T).sum(dim=1) / tau
    loss = ((r - d).sum() + (c - d).sum()) / (2.0 * B)
    dI = torch.zeros(len(rows), I.shape[1], dtype=F64, device=dev)
    dT = torch.zeros(len(rows), I.shape[1], dtype=F64, device=dev)
    dtau = torch.zeros((), dtype=F64, device=dev)
    ar = torch.arange(B, device=dev)
    for lo, hi in _slabs(B, slab):
        S = (I[lo:hi] @ T.t()) / tau
        G = (torch.exp(S - r[lo:hi, None]) + torch.exp(S - c[None, :])) / (2.0 * B)
        G[ar[lo:hi] - lo, ar[lo:hi]] -= 1.0 / B
        dtau -= (G * S).sum() / tau
        dT += (G[:, rows].t() @ I[lo:hi]) / tau                    # columns `rows` of G: sum over this slab's i
        sel = ((rows >= lo) & (rows < hi)).nonzero().flatten()
        if sel.numel():
            dI[sel] = (G[rows[sel] - lo] @ T) / tau
    return loss, dI, dT, dtau


def lunif(X, rows, t=2.0, slab=2048):
    """-> (loss, dX[rows])"""
    B = X.shape[0]
    n = (X * X).sum(dim=1)
    ssum = torch.zeros((), dtype=F64, device=X.device)
    rs = torch.zeros(len(rows), dtype=F64, device=X.device)
    WX = torch.zeros(len(rows), X.shape[1], dtype=F64, device=X.device)
    for lo, hi in _slabs(B, slab):
        d2 = (n[lo:hi, None] + n[None, :] - 2.0 * (X[lo:hi] @ X.t())).clamp_min(0.0)
        W = torch.exp(-t * d2)
        idx = torch.arange(lo, hi, device=X.device)
        W[idx - lo, idx] = 0.0
        ssum += 0.5 * W.sum()
        sel = ((rows >= lo) & (rows < hi)).nonzero().flatten()
        if sel.numel():
            Wr = W[rows[sel] - lo]
            rs[sel] = Wr.sum(dim=1)
            WX[sel] = Wr @ X
    loss = torch.log(ssum / (B * (B - 1) / 2.0))
    dX = (-2.0 * t / ssum) * (rs[:, None] * X[rows] - WX)
    return loss, dX


def weighted(I, T, tau, w_anchor, w_align, w_ui, w_ut, w_uc, rows, t=2.0, slab=2048):
    """Same composition as oracle.closed_form.weighted_loss -> (loss, dI[rows], dT[rows], dtau, terms)."""
    I, T = I.to(F64), T.to(F64)
    B, D = I.shape
    dI = torch.zeros(len(rows), D, dtype=F64, device=I.device)
    dT = torch.zeros_like(dI)
    loss = torch.zeros((), dtype=F64, device=I.device)
    dtau = torch.zeros((), dtype=F64, device=I.device)
    terms = {}
    if w_anchor != 0.0:
        a, gi, gt, gtau = anchor(I, T, tau, rows, slab)
        terms["anchor"] = a.item()
        loss += w_anchor * a; dI += w_anchor * gi; dT += w_anchor * gt; dtau += w_anchor * gtau
    if w_align != 0.0:
        diff = I - T
        a = (diff * diff).sum(dim=1).mean()
        terms["lalign"] = a.item()
        loss += w_align * a; dI += w_align * 2.0 * diff[rows] / B; dT -= w_align * 2.0 * diff[rows] / B
    if w_ui != 0.0:
        a, g = lunif(I, rows, t, slab)
        terms["lunif_img"] = a.item()
        loss += w_ui * a; dI += w_ui * g
    if w_ut != 0.0:
        a, g = lunif(T, rows, t, slab)
        terms["lunif_txt"] = a.item()
        loss += w_ut * a; dT += w_ut * g
    if w_uc != 0.0:
        m = (I + T) / 2.0
        nrm = m.norm(dim=1, keepdim=True).clamp_min(1e-12)
        C = m / nrm
        a, dc = lunif(C, rows, t, slab)
        terms["lunif_centroids"] = a.item()
        cr = C[rows]
        dm = (dc - cr * (cr * dc).sum(dim=1, keepdim=True)) / nrm[rows]
        loss += w_uc * a; dI += w_uc * dm / 2.0; dT += w_uc * dm / 2.0
    return loss.item(), dI, dT, dtau.item(), terms
