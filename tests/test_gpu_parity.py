"""GPU parity (run on the B200 box: pytest -m gpu).  Everything goes through the public Python API,
i.e. through the C ABI of libscb200.so; the oracle (tests/golden + oracle/closed_form.py) is the checker.

Tolerances: losses 1e-5 relative; gradients 1e-5 relative on the fp32 (SIMT, "tf32-off")
path and 1e-3 relative on the bf16 tensor-core path; relative = Frobenius norm of the difference / norm of
the reference (and per sampled row for the fixtures that store sampled rows only)."""
import math

import numpy as np
import pytest
import torch

import sparsify_clip_b200 as scb
from oracle import closed_form as cf
from sparsify_clip_b200 import backend_cuda
from tests import _golden

pytestmark = pytest.mark.gpu
CASES = _golden.names()
LOSS_RTOL = 1e-5
GRAD_RTOL = {"exact": 1e-5, "bf16": 1e-3}


def _dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype)


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-30)


def _run_terms(I, T, tau, gscale=1.0):
    """All terms through the public API; returns dict of losses and fp32 grads (already / gscale)."""
    out = {}
    I = I.clone().requires_grad_(True)
    T = T.clone().requires_grad_(True)
    tp = torch.nn.Parameter(torch.tensor(tau, dtype=torch.float32))      # CPU 0-dim, like sparsify_clip.py:716-717

    def grads(loss, *xs):
        gs = torch.autograd.grad(loss * gscale, xs)
        return [g.double().cpu().numpy() / gscale for g in gs]

    a = scb.contrastive_loss(I, T, tp)
    out["anchor"] = a.item()
    out["anchor_dI"], out["anchor_dT"], dt = grads(a, I, T, tp)
    out["anchor_dtau"] = float(dt)
    al = scb.lalign_loss(I, T)
    out["lalign"] = al.item()
    out["lalign_dI"], out["lalign_dT"] = grads(al, I, T)
    ui = scb.lunif_loss(I)
    out["lunif_img"] = ui.item()
    (out["lunif_img_dX"],) = grads(ui, I)
    ut = scb.lunif_loss(T)
    out["lunif_txt"] = ut.item()
    (out["lunif_txt_dX"],) = grads(ut, T)
    uc = scb.lunif_loss(scb.normalized_centroids(I, T), mma_dtype=scb.centroid_operand_dtype(I))
    out["lunif_cen"] = uc.item()
    out["lunif_cen_dI"], out["lunif_cen_dT"] = grads(uc, I, T)
    w3 = dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0)
    e3 = scb.weighted_loss(I, T, tp, w3)
    out["exp3"] = e3.item()
    out["exp3_dI"], out["exp3_dT"], dt = grads(e3, I, T, tp)
    out["exp3_dtau"] = float(dt)
    return out


def _check_against(z, got, loss_rtol, grad_rtol, cen_loss_rtol=None, row_factor=1.0):
    for k in ("anchor", "lalign", "lunif_img", "lunif_txt"):
        assert _rel(got[k], float(z["f64_" + k])) <= loss_rtol, (z["name"], k, got[k], float(z["f64_" + k]))
    mag = abs(float(z["f64_anchor"])) + abs(float(z["f64_lalign"])) + 0.5 * abs(float(z["f64_lunif_img"])) + \
        0.5 * abs(float(z["f64_lunif_txt"]))
    assert abs(got["exp3"] - float(z["f64_exp3"])) <= loss_rtol * mag, (z["name"], "exp3")
    assert _rel(got["lunif_cen"], float(z["f64_lunif_cen"])) <= (cen_loss_rtol or loss_rtol), (z["name"], "lunif_cen")
    for k in ("anchor_dtau", "exp3_dtau"):
        assert _rel(got[k], float(z["f64_" + k])) <= max(grad_rtol, 1e-5), (z["name"], k, got[k], float(z["f64_" + k]))
    for k in ("anchor_dI", "anchor_dT", "lalign_dI", "lalign_dT", "lunif_img_dX", "lunif_txt_dX", "lunif_cen_dI",
              "lunif_cen_dT", "exp3_dI", "exp3_dT"):
        _golden.grad_check(z, k, got[k], grad_rtol, row_factor)


@pytest.mark.parametrize("name", CASES)
def test_fp32_exact_path_matches_golden(name):
    """fp32 inputs -> SIMT fp32 kernels: 1e-5 on every loss and gradient (the 'tf32-off fp32' gate)."""
    z = _golden.load(name)
    scb.set_fp32_mode("exact")
    got = _run_terms(_dev(z["I"]), _dev(z["T"]), float(z["tau"]), gscale=1024.0)   # GradScaler-like grad_output
    _check_against(z, got, LOSS_RTOL, GRAD_RTOL["exact"])


@pytest.mark.parametrize("tc_flags", [0, 1, 3])      # smem / TMEM weight tile; 3 = CTA-pair gradient kernel (256 < D <= 512)
@pytest.mark.parametrize("name", [n for n in CASES if bool(_golden.load(n)["bf16_exact"])])
def test_bf16_tensor_core_path_matches_golden(name, tc_flags):
    """bf16-exact inputs -> TMA + tcgen05 kernels.  Inputs are handed over as fp32 tensors holding bf16-exact
    values with fp32 mode 'bf16', so the returned gradients are fp32 (not re-rounded to bf16)."""
    z = _golden.load(name)
    assert bool(z["bf16_exact"])
    be = scb.get_backend()
    prev_flags = be.lib.scb_set_tc_flags(tc_flags)
    prev = scb.set_fp32_mode("bf16")
    try:
        got = _run_terms(_dev(z["I"]), _dev(z["T"]), float(z["tau"]), gscale=3.0)
    finally:
        scb.set_fp32_mode(prev)
        be.lib.scb_set_tc_flags(prev_flags)
    # The gate is norm-wise: 1e-3 of the gradient's Frobenius norm, checked in full against the oracle below.
    # Fixtures that only store sampled rows (which include the planted duplicate rows, the worst case at
    # tau = 0.01 where one 2^-9-rounded weight dominates a row) bound each sampled row by 3e-3 of the typical row norm.
    _check_against(z, got, LOSS_RTOL, GRAD_RTOL["bf16"], row_factor=3.0)
    _, dI, dT, _, _ = cf.weighted_loss(z["I"], z["T"], float(z["tau"]), 1.0, 1.0, 0.5, 0.5, 0.0)
    _, aI, aT, _ = cf.contrastive_loss(z["I"], z["T"], float(z["tau"]))
    for key, ref in (("exp3_dI", dI), ("exp3_dT", dT), ("anchor_dI", aI), ("anchor_dT", aT)):
        err = np.linalg.norm(got[key] - ref) / np.linalg.norm(ref)
        assert err <= GRAD_RTOL["bf16"], (name, key, err)


@pytest.mark.parametrize("B,D,tau,kind", [(127, 512, 0.1, "corr"), (129, 1024, 0.07, "cluster"), (1000, 64, 1.0, "corr"),
                                           (4096, 512, 0.1, "corr"), (300, 768, 0.01, "cluster"), (2, 64, 0.1, "iid"),
                                           (3, 8, 0.5, "iid"),
                                           # CTA-pair gradient kernel (256 < D <= 512): every K-chunk count 5..8 (padding
                                           # slots, odd output halves), D tails, odd / single tile counts, row tails
                                           (640, 320, 0.1, "corr"), (385, 384, 0.07, "cluster"), (1300, 448, 0.1, "corr"),
                                           (129, 264, 0.1, "iid"), (128, 512, 0.1, "corr"), (2100, 456, 0.05, "cluster"),
                                           (5000, 512, 0.1, "corr"),
                                           # cluster-of-4 kernel with column groups (512 < D <= 1024): a second group of
                                           # one chunk per CTA (D <= 768, clamped chunks at D = 576) or two (D > 768),
                                           # D tails, odd tile counts, several row blocks per cluster span
                                           (640, 576, 0.1, "corr"), (2100, 768, 0.05, "cluster"), (513, 832, 0.07, "cluster"),
                                           (900, 776, 0.1, "iid"), (1300, 1024, 0.1, "corr"), (4096, 768, 0.1, "corr")])
def test_tensor_core_path_vs_oracle_on_seeded_inputs(B, D, tau, kind):
    from oracle.make_golden import make_inputs
    I, T = make_inputs(kind, B, D, seed=B + D)
    I, T = I.to(torch.bfloat16).float(), T.to(torch.bfloat16).float()
    prev = scb.set_fp32_mode("bf16")
    try:
        Ig, Tg = I.cuda().requires_grad_(True), T.cuda().requires_grad_(True)
        tp = torch.nn.Parameter(torch.tensor(tau))
        loss = scb.weighted_loss(Ig, Tg, tp, dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0))
        loss.backward()
    finally:
        scb.set_fp32_mode(prev)
    ref, dI, dT, dtau, terms = cf.weighted_loss(I.numpy(), T.numpy(), tau, 1.0, 1.0, 0.5, 0.5, 0.0)
    # the composite can cancel to ~0: 1e-5 relative to the magnitude of its terms
    assert abs(loss.item() - ref) <= LOSS_RTOL * sum(abs(v) for v in terms.values())
    assert np.linalg.norm(Ig.grad.double().cpu().numpy() - dI) / np.linalg.norm(dI) <= 1e-3
    assert np.linalg.norm(Tg.grad.double().cpu().numpy() - dT) / np.linalg.norm(dT) <= 1e-3
    assert _rel(tp.grad.item(), dtau) <= 1e-3


@pytest.mark.parametrize("B,D,tau", [(2100, 768, 0.05), (5000, 512, 0.1), (1300, 1024, 0.1), (700, 640, 0.07)])
def test_row_block_aligned_spans_of_the_cluster_kernel(B, D, tau):
    """The work split used when the column operand does not fit in L2 (whole row blocks per cluster, one partial slot;
    tc_flags bit5 forces it at test sizes): same results as the equal-span split, and inside the oracle gate."""
    from oracle.make_golden import make_inputs
    I, T = make_inputs("cluster", B, D, seed=B + D)
    I, T = I.to(torch.bfloat16).float(), T.to(torch.bfloat16).float()      # bf16-exact values, fp32 gradients back
    be = scb.get_backend()
    w = dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0)
    got = {}
    for flags in (31, 63, 15):       # 15: D = 768 in two column groups instead of the single-S-buffer variant (bit4 off)
        prev = be.lib.scb_set_tc_flags(flags)
        prev_mode = scb.set_fp32_mode("bf16")
        try:
            Ig, Tg = I.cuda().requires_grad_(True), T.cuda().requires_grad_(True)
            tp = torch.nn.Parameter(torch.tensor(tau))
            loss = scb.weighted_loss(Ig, Tg, tp, w)
            loss.backward()
            got[flags] = (loss.item(), Ig.grad.double().cpu().numpy(), Tg.grad.double().cpu().numpy(), tp.grad.item())
        finally:
            scb.set_fp32_mode(prev_mode)
            be.lib.scb_set_tc_flags(prev)
    ref, dI, dT, dtau, terms = cf.weighted_loss(I.numpy(), T.numpy(), tau, 1.0, 1.0, 0.5, 0.5, 0.0)
    mag = sum(abs(v) for v in terms.values())
    for flags, (l, gI, gT, gt) in got.items():
        assert abs(l - ref) <= LOSS_RTOL * mag, (flags, l, ref)
        assert np.linalg.norm(gI - dI) / np.linalg.norm(dI) <= 1e-3, flags
        assert np.linalg.norm(gT - dT) / np.linalg.norm(dT) <= 1e-3, flags
        assert _rel(gt, dtau) <= 1e-3, flags
    # the two splits differ only in where the fp32 partial sums are cut
    assert abs(got[31][0] - got[63][0]) <= 1e-6 * mag
    assert np.linalg.norm(got[31][1] - got[63][1]) <= 1e-5 * np.linalg.norm(dI)
    assert np.linalg.norm(got[31][2] - got[63][2]) <= 1e-5 * np.linalg.norm(dT)


def test_full_size_properties_c3():
    """B = 32768, D = 512 (BASELINE c3) -- too big for the dense oracle; size-independent properties:
    translation invariance of L_unif (gradient rows sum to 0), Euler homogeneity of the anchor
    (sum_i I_i.dI_i + tau dtau = 0 and likewise for T), I<->T symmetry, permutation invariance."""
    B, D, tau = 32768, 512, 0.1
    g = torch.Generator(device="cuda").manual_seed(42)
    I0 = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    T0 = torch.nn.functional.normalize(I0 + 0.5 * torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    I = I0.to(torch.bfloat16).float().requires_grad_(True)
    T = T0.to(torch.bfloat16).float().requires_grad_(True)
    prev = scb.set_fp32_mode("bf16")
    try:
        tp = torch.nn.Parameter(torch.tensor(tau))
        a = scb.contrastive_loss(I, T, tp)
        gI, gT, gtau = torch.autograd.grad(a, (I, T, tp))
        e_i = (I.detach().double() * gI.double()).sum().item()
        e_t = (T.detach().double() * gT.double()).sum().item()
        assert abs(e_i + tau * gtau.item()) <= 2e-3 * abs(e_i)
        assert abs(e_t + tau * gtau.item()) <= 2e-3 * abs(e_t)
        a_sw = scb.contrastive_loss(T, I, tp)
        assert _rel(a_sw.item(), a.item()) <= 1e-6
        u = scb.lunif_loss(I)
        (gu,) = torch.autograd.grad(u, (I,))
        assert gu.double().sum(0).norm().item() <= 1e-3 * gu.double().norm().item()
        perm = torch.randperm(B, device="cuda")
        assert _rel(scb.lunif_loss(I.detach()[perm]).item(), u.item()) <= 1e-6
        # iid-like unit vectors in high D: L_unif ~ -4 + O(1/D)  
        assert -4.05 < u.item() < -3.9
        # forward-only sweep (no grad) agrees with the fused forward+backward sweep
        with torch.no_grad():
            assert _rel(scb.lunif_loss(I).item(), u.item()) <= 1e-6
    finally:
        scb.set_fp32_mode(prev)


def test_native_bf16_and_fp16_leaves_and_noncontiguous():
    B, D, tau = 256, 512, 0.1
    g = torch.Generator(device="cuda").manual_seed(1)
    I = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    T = torch.nn.functional.normalize(I + 0.5 * torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    for dt in (torch.bfloat16, torch.float16):
        Iq, Tq = I.to(dt), T.to(dt)
        ref, dI, dT, _, _ = cf.weighted_loss(Iq.float().cpu().numpy(), Tq.float().cpu().numpy(), tau, 1.0, 1.0, 0.5, 0.5, 0.0)
        Ig, Tg = Iq.clone().requires_grad_(True), Tq.clone().requires_grad_(True)
        loss = scb.weighted_loss(Ig, Tg, tau, dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0))
        loss.backward()
        assert loss.dtype == torch.float32 and loss.dim() == 0 and Ig.grad.dtype == dt
        assert _rel(loss.item(), ref) <= LOSS_RTOL
        # gradients come back in the leaf dtype: one rounding to 8 (bf16) / 11 (fp16) bits per summed term
        tol = 6e-3 if dt == torch.bfloat16 else 1.5e-3
        assert np.linalg.norm(Ig.grad.double().cpu().numpy() - dI) / np.linalg.norm(dI) <= tol
    # non-contiguous view (column slice of a wider buffer) and transposed input
    wide = torch.randn(B, 2 * D, device="cuda")
    x = torch.nn.functional.normalize(wide[:, ::2], dim=-1)
    xs = x.t().contiguous().t()                      # stride(1) != 1
    assert _rel(scb.lunif_loss(xs).item(), cf.lunif_loss(x.cpu().numpy(), need_grad=False)) <= LOSS_RTOL


def test_edge_cases():
    x = torch.nn.functional.normalize(torch.randn(1, 64, device="cuda"), dim=-1)
    assert torch.isnan(scb.lunif_loss(x))                      # mean over an empty pdist vector
    x2 = torch.nn.functional.normalize(torch.randn(2, 64, device="cuda"), dim=-1)
    assert _rel(scb.lunif_loss(x2).item(), cf.lunif_loss(x2.cpu().numpy(), need_grad=False)) <= LOSS_RTOL
    y = x2.clone().requires_grad_(True)
    l = scb.lalign_loss(y, x2)
    l.backward()
    assert l.item() == 0.0 and torch.all(y.grad == 0)           # zero (not nan) where x_i == y_i
    with pytest.raises(ValueError):
        scb.lunif_loss(torch.randn(4, 4, 4, device="cuda"))
    # sparsify_loss forward on both paths
    xs = torch.nn.functional.normalize(torch.randn(200, 96, device="cuda"), dim=-1)
    assert _rel(scb.sparsify_loss(xs).item(), cf.sparsify_loss(xs.cpu().numpy(), need_grad=False)) <= LOSS_RTOL
    xb = xs.to(torch.bfloat16)
    assert _rel(scb.sparsify_loss(xb).item(), cf.sparsify_loss(xb.float().cpu().numpy(), need_grad=False)) <= LOSS_RTOL


def test_ladder_on_gpu_matches_oracle():
    g = torch.Generator().manual_seed(5)
    I = torch.nn.functional.normalize(torch.randn(192, 128, generator=g), dim=-1)
    T = torch.nn.functional.normalize(I + 0.5 * torch.randn(192, 128, generator=g), dim=-1)
    base = {"only_lunif_epochs": 1, "beta_warmup_epoch": 20, "beta_decay_epoch": 50, "alpha_warmup_epoch": 50,
            "alpha_increment_epoch": 50}
    for lt in scb.LOSS_TYPES:
        if lt.endswith("[intended]"):
            continue
        for epoch, step in ((0, 10), (2, 600)):
            cfg = dict(base, loss_type=lt)
            Ig, Tg = I.cuda().requires_grad_(True), T.cuda().requires_grad_(True)
            loss = scb.compose_loss(cfg, Ig, Tg, 0.1, epoch=epoch, current_batch=step, t_total=1000)
            loss.backward()
            ref, dI, dT, _, terms = cf.compose_loss(cfg, I.numpy(), T.numpy(), 0.1, epoch, step, 1000)
            assert abs(loss.item() - ref) <= LOSS_RTOL * sum(abs(v) for v in terms.values()), (lt, epoch)
            assert np.linalg.norm(Ig.grad.double().cpu().numpy() - dI) <= 1e-5 * max(np.linalg.norm(dI), 1e-30), (lt, epoch)


@pytest.mark.parametrize("B,D,dtype", [(384, 512, torch.bfloat16), (200, 128, torch.float32), (1000, 768, torch.float16),
                                       (150, 20, torch.float32), (131, 44, torch.bfloat16)])      # D % 8 != 0: scalar paths
def test_fused_composition_equals_the_sum_of_its_terms(B, D, dtype):
    """weighted_loss through the fused autograd node (scb_grad_combine) vs the term-by-term composition: same passes,
    so the loss agrees to fp32 rounding and the gradients to the rounding of the output dtype."""
    g = torch.Generator(device="cuda").manual_seed(7)
    I0 = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    T0 = torch.nn.functional.normalize(I0 + 0.5 * torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    for w in (dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0),
              dict(anchor=1.0, align=1.3, unif_img=0.1, unif_txt=0.1, unif_cen=0.0),
              dict(anchor=0.0, align=0.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0),
              dict(anchor=1.0, align=0.0, unif_img=0.0, unif_txt=0.0, unif_cen=0.0),
              dict(anchor=1.0, align=1.0, unif_img=0.0, unif_txt=0.0, unif_cen=1.0),      # exp 4 / 10: centroid chain
              dict(anchor=1.0, align=0.7, unif_img=0.25, unif_txt=0.25, unif_cen=0.5)):
        res = []
        for fused in (True, False):
            prev = scb.set_fused(fused)
            try:
                I = I0.to(dtype).requires_grad_(True)
                T = T0.to(dtype).requires_grad_(True)
                tp = torch.nn.Parameter(torch.tensor(0.1))
                loss = scb.weighted_loss(I, T, tp, w)
                (loss * 4.0).backward()
                res.append((loss.item(), I.grad.float(), T.grad.float(), None if tp.grad is None else tp.grad.item()))
            finally:
                scb.set_fused(prev)
        (lf, dIf, dTf, dtf), (lm, dIm, dTm, dtm) = res
        assert abs(lf - lm) <= 2e-6 * max(1.0, abs(lm)), (w, lf, lm)
        tol = 1e-5 if dtype == torch.float32 else 8e-3          # one extra rounding per term in the modular path
        for a, b in ((dIf, dIm), (dTf, dTm)):
            assert ((a - b).norm() / b.norm().clamp_min(1e-30)).item() <= tol, (w, dtype)
        if w["anchor"] != 0.0:
            assert abs(dtf - dtm) <= 1e-5 * abs(dtm), (w, dtf, dtm)


class _GuardedAlloc:
    """torch.empty / torch.zeros stand-in: every buffer sits between two 1 KiB canary zones."""
    PAD = 1024      # bytes (multiple of 16: TMA / vector-access alignment is preserved)

    def __init__(self):
        self.live = []

    def _make(self, fill, *shape, dtype=torch.float32, device=None):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        n = 1
        for d in shape:
            n *= int(d)
        esz = torch.empty((), dtype=dtype).element_size()
        nbytes = ((n * esz + 15) // 16) * 16
        raw = torch.full((nbytes + 2 * self.PAD,), 0xA5, dtype=torch.uint8, device=device)
        inner = raw[self.PAD:self.PAD + n * esz].view(dtype).reshape(tuple(shape)) if n else torch.empty(shape, dtype=dtype, device=device)
        if fill is not None and n:
            inner.fill_(fill)
        self.live.append((raw, nbytes))
        return inner

    def empty(self, *shape, **kw):
        return self._make(None, *shape, **kw)

    def zeros(self, *shape, **kw):
        return self._make(0, *shape, **kw)

    def check(self):
        torch.cuda.synchronize()
        bad = 0
        for raw, nbytes in self.live:
            lo, hi = raw[:self.PAD], raw[self.PAD + nbytes:]
            bad += int((lo != 0xA5).sum().item()) + int((hi != 0xA5).sum().item())
        return bad, len(self.live)


@pytest.mark.parametrize("B,D,tau,dtype", [(385, 384, 0.07, torch.bfloat16), (129, 264, 0.1, torch.bfloat16), (640, 320, 0.1, torch.float16),
                                           (300, 512, 0.01, torch.bfloat16), (257, 768, 0.1, torch.bfloat16),
                                           (131, 44, 0.1, torch.float32), (1000, 512, 0.1, torch.bfloat16),
                                           (385, 1000, 0.1, torch.bfloat16), (1000, 768, 0.07, torch.float16),
                                           (300, 584, 0.1, torch.bfloat16)])
def test_kernels_stay_inside_their_buffers(B, D, tau, dtype):
    """No out-of-bounds write by any kernel (compute-sanitizer is closed on this pool): all library-side buffers are
    allocated with canary zones around them; tails in rows, columns and D, every K-chunk count of the pair kernel."""
    g = torch.Generator(device="cuda").manual_seed(B + D)
    I0 = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    T0 = torch.nn.functional.normalize(I0 + 0.5 * torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    ga = _GuardedAlloc()
    prev = backend_cuda.set_allocator(ga.empty, ga.zeros)
    try:
        for w in (dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0),
                  dict(anchor=1.0, align=1.0, unif_img=0.0, unif_txt=0.0, unif_cen=1.0)):
            for fused in (True, False):
                pf = scb.set_fused(fused)
                try:
                    I = I0.to(dtype).requires_grad_(True)
                    T = T0.to(dtype).requires_grad_(True)
                    tp = torch.nn.Parameter(torch.tensor(tau))
                    scb.weighted_loss(I, T, tp, w).backward()
                finally:
                    scb.set_fused(pf)
        bad, n = ga.check()
    finally:
        backend_cuda.set_allocator(*prev)
    assert n > 20 and bad == 0, (bad, n)


@pytest.mark.parametrize("scale_i,scale_t,tau,same", [(4.0, 3.0, 0.1, False), (1.0, 1.0, 0.1, True), (0.05, 20.0, 0.5, False),
                                                      (1.0, 1.0, 0.02, False)])
def test_unnormalised_and_degenerate_inputs(scale_i, scale_t, tau, same):
    """Rows far from unit norm (logits of several hundred: the device-side norm bound must arm the exact column sweep),
    I == T (every diagonal logit is its row's and column's maximum) and a small temperature, against the fp64 oracle."""
    B, D = 300, 512
    g = torch.Generator().manual_seed(3)
    I = torch.randn(B, D, generator=g)
    I = scale_i * I / I.norm(dim=1, keepdim=True) * (0.5 + torch.rand(B, 1, generator=g))
    T = I.clone() if same else torch.randn(B, D, generator=g)
    if not same:
        T = scale_t * T / T.norm(dim=1, keepdim=True) * (0.5 + torch.rand(B, 1, generator=g))
    I, T = I.to(torch.bfloat16).float(), T.to(torch.bfloat16).float()
    prev = scb.set_fp32_mode("bf16")
    try:
        Ig, Tg = I.cuda().requires_grad_(True), T.cuda().requires_grad_(True)
        tp = torch.nn.Parameter(torch.tensor(tau))
        loss = scb.weighted_loss(Ig, Tg, tp, dict(anchor=1.0, align=0.0, unif_img=0.0, unif_txt=0.0, unif_cen=0.0))
        loss.backward()
    finally:
        scb.set_fp32_mode(prev)
    ref, dI, dT, dtau = cf.contrastive_loss(I.numpy(), T.numpy(), tau)
    # the loss is mean(lse - diag): with I == T the two nearly cancel, so the bound is stated against the size of the
    # terms (fp32 tensor-core accumulation truncates, ~1e-6 relative on each logit), not against the difference
    diag_mag = float(np.abs((I.double() * T.double()).sum(1).numpy()).mean()) / tau
    assert np.isfinite(loss.item()) and abs(loss.item() - ref) <= 1e-5 * abs(ref) + 2e-6 * diag_mag, (loss.item(), ref)
    assert np.linalg.norm(Ig.grad.double().cpu().numpy() - dI) <= 2e-3 * np.linalg.norm(dI)
    assert np.linalg.norm(Tg.grad.double().cpu().numpy() - dT) <= 2e-3 * np.linalg.norm(dT)
    assert _rel(tp.grad.item(), dtau) <= 2e-3


@pytest.mark.parametrize("world,n_loc", [(2, 300), (8, 129), (1, 64)])
def test_rank_fold_of_column_partials(world, n_loc):
    """scb_lse2_fold_ranks against the log-sum-exp it restates, on a synthetic packed gather (both flag states)."""
    be = scb.get_backend()
    g = torch.Generator(device="cuda").manual_seed(world * 1000 + n_loc)
    B, NS = world * n_loc, 5
    width = 2 * n_loc + NS + 2 * B
    pack = torch.randn(world, width, generator=g, device="cuda")
    off_ref, off_sum = 2 * n_loc + NS, 2 * n_loc + NS + B
    pack[:, off_ref:off_ref + B] *= 30.0                                   # references tens of log2 units apart
    pack[:, off_sum:off_sum + B] = pack[:, off_sum:off_sum + B].abs() + 0.5
    if world > 1:
        pack[0, off_ref + 3] = float("-inf")                               # a rank that saw nothing for a column
        pack[0, off_sum + 3] = 0.0
    Mr, Lr = pack[:, off_ref:off_ref + B].double(), pack[:, off_sum:off_sum + B].double()
    want = torch.logsumexp(Mr * math.log(2.0) + torch.log(Lr), dim=0)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    got = be.lse2_fold_ranks(pack, n_loc, n_loc, off_ref, off_sum, flag)
    assert (got.double() - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())
    flag.fill_(1)
    got = be.lse2_fold_ranks(pack, n_loc, n_loc, off_ref, off_sum, flag)
    assert torch.equal(got, pack[:, n_loc:2 * n_loc].reshape(-1))
