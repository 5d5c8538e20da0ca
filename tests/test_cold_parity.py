"""CPU parity of the cold, signature-only surface: the package's functions against the UNMODIFIED
reference (loaded through oracle/ref_loader.py when /root/reference is present) and against oracle/closed_form.py
(always).  Reference lines: sparsify_clip.py:135-157 (contrastive_loss_roberta), :166-176 (sparsify_loss), :178-184
(random_alignment_loss), :308-332 (compute_centroids), :334-355 (compute_centroids_only), :487-505
(centroid_alignment_loss); uniformity.py:6, :53, :101, :138, :182 (the five W2 variants)."""
import numpy as np
import pytest
import torch

import sparsify_clip_b200 as scb
from oracle import closed_form as cf
from oracle import ref_loader
from sparsify_clip_b200 import uniformity as pu


def _inputs(B=48, D=24, seed=0, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    I = torch.nn.functional.normalize(torch.randn(B, D, generator=g, dtype=dtype), dim=-1)
    T = torch.nn.functional.normalize(I + 0.5 * torch.randn(B, D, generator=g, dtype=dtype), dim=-1)
    return I, T


def _ref():
    if not ref_loader.available():
        pytest.skip("reference not present on this box")
    return ref_loader.load()


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_cold_losses_match_the_reference(dtype):
    ref, _ = _ref()
    I, T = _inputs(dtype=dtype)
    tol = 1e-12 if dtype == torch.float64 else 1e-5
    R = torch.softmax((T @ T.t()) / 0.1, dim=-1)           # the soft targets of the dead-code call site (:1196-1202)
    for tau in (0.07, 0.5):
        a = scb.contrastive_loss_roberta(I, T, R, tau)
        b = ref.contrastive_loss_roberta(I, T, R, tau)
        assert abs(a.item() - b.item()) <= tol * abs(b.item())
    assert scb.contrastive_loss_roberta(I, T, R).item() == pytest.approx(ref.contrastive_loss_roberta(I, T, R).item(), rel=tol)
    for p in (2, 1, 3):
        a, b = scb.centroid_alignment_loss(I, T, p=p), ref.centroid_alignment_loss(I, T, p=p)
        assert abs(a.item() - b.item()) <= tol * abs(b.item())
    assert scb.centroid_alignment_loss(I, T).item() == pytest.approx(ref.centroid_alignment_loss(I, T).item(), rel=tol)
    n1, c1 = scb.compute_centroids(I[:7], T[:5])
    n2, c2 = ref.compute_centroids(I[:7], T[:5])
    assert n1.shape == (7, 5) and c1.shape == (7, 5, I.shape[1])
    assert torch.equal(n1, n2) and torch.equal(c1, c2)
    assert torch.equal(scb.compute_centroids_only(I, T), ref.compute_centroids_only(I, T))
    # gradients of the two differentiable one-liners through autograd, package vs reference
    for fn_a, fn_b in ((lambda x, y: scb.centroid_alignment_loss(x, y), lambda x, y: ref.centroid_alignment_loss(x, y)),
                       (lambda x, y: scb.contrastive_loss_roberta(x, y, R, 0.2), lambda x, y: ref.contrastive_loss_roberta(x, y, R, 0.2))):
        xa, ya = I.clone().requires_grad_(True), T.clone().requires_grad_(True)
        xb, yb = I.clone().requires_grad_(True), T.clone().requires_grad_(True)
        fn_a(xa, ya).backward()
        fn_b(xb, yb).backward()
        assert (xa.grad - xb.grad).abs().max().item() <= tol and (ya.grad - yb.grad).abs().max().item() <= tol


def test_cold_losses_match_the_closed_forms():
    I, T = _inputs()
    R = torch.softmax((T @ T.t()) / 0.1, dim=-1)
    assert scb.contrastive_loss_roberta(I, T, R, 0.2).item() == pytest.approx(
        cf.contrastive_loss_roberta(I.numpy(), T.numpy(), R.numpy(), 0.2), rel=1e-12)
    for p in (1, 2, 3):
        assert scb.centroid_alignment_loss(I, T, p=p).item() == pytest.approx(cf.centroid_alignment_loss(I.numpy(), T.numpy(), p), rel=1e-12)
    n, c = scb.compute_centroids(I[:6], T[:4])
    n2, c2 = cf.compute_centroids(I[:6].numpy(), T[:4].numpy())
    assert np.abs(n.numpy() - n2).max() <= 1e-14 and np.abs(c.numpy() - c2).max() <= 1e-14


@pytest.mark.parametrize("dtype,B,D", [(torch.float64, 96, 16), (torch.float32, 300, 32), (torch.float64, 40, 64)])
def test_uniformity_variants_match_the_reference(dtype, B, D, capsys):
    """All five W2 variants, package vs the reference's uniformity.py on the same rows (B < D included: a singular
    covariance, where the variants differ in how they clamp)."""
    _, uni = _ref()
    g = torch.Generator().manual_seed(B + D)
    x1 = torch.nn.functional.normalize(torch.randn(B, D, generator=g, dtype=dtype) + 0.3, dim=-1)
    x2 = torch.nn.functional.normalize(torch.randn(B, D, generator=g, dtype=dtype) - 0.2, dim=-1)
    tol = 1e-9 if dtype == torch.float64 else 2e-4
    pairs = [
        (pu.torch_uniformity1(x1), uni.torch_uniformity1(x1)),
        (pu.torch_uniformity(x1, x2), uni.torch_uniformity(x1, x2)),
        (pu.numpy_uniformity(x1, x2), uni.numpy_uniformity(x1, x2)),
        (pu.torch_uniformity_equivalent(x1), uni.torch_uniformity_equivalent(x1)),
    ]
    if B > D:      # uniformity10 takes |Q| of a general eigensolver: only defined up to the solver's vector order/sign
        pairs.append((pu.uniformity10(x1), uni.uniformity10(x1)))   # for a nondegenerate spectrum (B > D)
    capsys.readouterr()
    for a, b in pairs:
        a, b = float(a), float(b)
        assert np.isfinite(b) and abs(a - b) <= tol * max(1.0, abs(b)), (a, b)
    # the sign conventions of the reference: two-modality variants return -W2, one-modality variants +W2
    assert float(pairs[0][0]) > 0 and float(pairs[1][0]) < 0 and float(pairs[2][0]) < 0 and float(pairs[3][0]) > 0


def test_reference_eval_script_uniformity_matches_numpy_uniformity():
    """sparsify_clip.py:459-485 (`uniformity`, used by evaluate_model) is the numpy variant of uniformity.py:101."""
    ref, uni = _ref()
    I, T = _inputs(B=128, D=16)
    assert pu.numpy_uniformity(I, T) == pytest.approx(ref.uniformity(I, T), rel=1e-10)


def test_eval_metric_restatements_match_the_reference(capsys):
    """oracle/closed_form.py's restatements of the evaluation-side consumers (sparsify_clip.py:357-528) against the
    unmodified reference: ranks and R@k (paired ids and a multi-caption id list), gap, mean angular value, true-pair
    cosine, W2 uniformity."""
    ref, _ = _ref()
    I, T = _inputs(B=96, D=16, dtype=torch.float32)
    S = T @ I.t()                                            # [N_text, N_image], as sparsify_clip.py:628
    for ids, ids_txt in ((list(range(96)), list(range(96))),
                         (list(range(48)), [k // 2 for k in range(96)])):      # two captions per image
        Sx = S[:, :len(ids)]
        fwd, bwd = cf.retrieval_ranks(Sx.numpy(), ids, ids_txt)
        assert cf.recall_log(fwd, "forward") == ref.compute_metric_ret(Sx, ids, ids_txt, "forward")
        assert cf.recall_log(bwd, "backward") == ref.compute_metric_ret(Sx, ids, ids_txt, "backward")
    assert cf.compute_gap(I.numpy(), T.numpy()) == pytest.approx(ref.compute_gap(I, T), rel=1e-5)
    assert cf.mean_angular_value(I.numpy()) == pytest.approx(ref.compute_mean_angular_value_of_a_modality(I), rel=1e-4, abs=1e-7)
    assert cf.mean_true_pair_cosine(I.numpy(), T.numpy()) == pytest.approx(ref.mean_distance_of_true_pairs(I, T), rel=1e-5)
    assert cf.w2_uniformity(I.numpy(), T.numpy()) == pytest.approx(ref.uniformity(I, T), rel=1e-5)
    capsys.readouterr()
