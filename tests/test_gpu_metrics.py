"""GPU parity (pytest -m gpu) of the cold loss variants and the evaluation-side consumers on their kernels
(csrc/metrics.cu, the rank-count mode of the sweeps): against tests/golden_eval/*.npz minted from the unmodified
reference (oracle/make_golden_eval.py) and against oracle/closed_form.py.
Reference lines: sparsify_clip.py:166-176 (sparsify_loss), :357-416 (compute_metric_ret), :418-436 (compute_gap), :438-457
(mean angular value), :459-485 (uniformity), :487-505 (centroid_alignment_loss), :508-528 (true-pair cosine);
uniformity.py:6-205."""
import glob
import os

import numpy as np
import pytest
import torch

import sparsify_clip_b200 as scb
from oracle import closed_form as cf
from sparsify_clip_b200 import metrics as M
from sparsify_clip_b200 import uniformity as pu

pytestmark = pytest.mark.gpu
EVAL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_eval")
EVAL_CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(EVAL_DIR, "*.npz")))


def _cuda(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype)


@pytest.mark.parametrize("name", EVAL_CASES)
def test_eval_consumers_match_the_reference_golden(name):
    z = dict(np.load(os.path.join(EVAL_DIR, name + ".npz")))
    img, txt = _cuda(z["img"]), _cuda(z["txt"])
    n = img.shape[0]
    ids = list(range(n))
    # the reference's interface: a materialised score matrix + id lists
    S = txt @ img.t()
    fwd = M.compute_metric_ret(S, ids, ids, direction="forward")
    bwd = M.compute_metric_ret(S, ids, ids, direction="backward")
    keys_f = ["forward_r1", "forward_r5", "forward_r10", "forward_ravg"]
    keys_b = ["backward_r1", "backward_r5", "backward_r10", "backward_ravg"]
    assert [fwd[k] for k in keys_f] == pytest.approx(list(z["forward"]), abs=1e-9)
    assert [bwd[k] for k in keys_b] == pytest.approx(list(z["backward"]), abs=1e-9)
    # the fused form: straight from the features, no N x N matrix (fp32 features -> exact fp32 path)
    both = M.retrieval_metrics(txt, img)
    assert [both[k] for k in keys_f] == pytest.approx(list(z["forward"]), abs=1e-9)
    assert [both[k] for k in keys_b] == pytest.approx(list(z["backward"]), abs=1e-9)
    assert M.compute_gap(img, txt) == pytest.approx(float(z["gap"]), rel=1e-5)
    assert M.compute_mean_angular_value_of_a_modality(img) == pytest.approx(float(z["ang_img"]), rel=1e-4, abs=1e-7)
    assert M.compute_mean_angular_value_of_a_modality(txt) == pytest.approx(float(z["ang_txt"]), rel=1e-4, abs=1e-7)
    assert M.mean_distance_of_true_pairs(img, txt) == pytest.approx(float(z["cos_true"]), rel=1e-5)
    assert M.uniformity(img, txt) == pytest.approx(float(z["unif"]), rel=1e-4)
    # uniformity.py's variants on CUDA tensors (covariance on the library's kernel).  They decompose the covariance in
    # fp32 (svd / eigh / eig), and W2^2 is a difference of O(1) terms: the reference's own variants disagree with each
    # other at the 3e-4 level on this input (LAPACK fp32 vs NumPy fp64), so cuSOLVER vs LAPACK gets 1e-3
    assert float(pu.torch_uniformity1(img)) == pytest.approx(float(z["u1"]), rel=1e-3)
    assert float(pu.torch_uniformity(img, txt)) == pytest.approx(float(z["u2"]), rel=1e-3)
    assert float(pu.torch_uniformity_equivalent(img)) == pytest.approx(float(z["u_eq"]), rel=1e-3)
    assert pu.numpy_uniformity(img, txt) == pytest.approx(float(z["unif"]), rel=1e-4)


@pytest.mark.parametrize("N,D,dtype", [(1000, 512, torch.bfloat16), (4096, 512, torch.bfloat16), (333, 40, torch.float32),
                                       (2048, 768, torch.float16)])
def test_fused_retrieval_ranks_vs_sorted_oracle(N, D, dtype):
    """retrieval_ranks from the features (tensor-core sweep for 16-bit features) against the reference's full sort of the
    fp64 score matrix.  A rank may differ only where another score ties the true pair's within the rounding of the fp32
    accumulation (|s - s_true| <= 2e-6): those near-ties are counted and must be rare; R@k must agree."""
    g = torch.Generator(device="cuda").manual_seed(N + D)
    img = torch.nn.functional.normalize(torch.randn(N, D, generator=g, device="cuda") + 0.3, dim=-1)
    txt = torch.nn.functional.normalize(img + (5.0 / D ** 0.5) * torch.randn(N, D, generator=g, device="cuda"), dim=-1)
    img, txt = img.to(dtype), txt.to(dtype)
    fwd, bwd = M.retrieval_ranks(txt, img)
    S = (txt.double() @ img.double().t()).cpu().numpy()
    ids = list(range(N))
    ofwd, obwd = cf.retrieval_ranks(S, ids, ids)
    for got, want, axis in ((fwd.cpu().numpy(), ofwd, 1), (bwd.cpu().numpy(), obwd, 0)):
        diff = np.nonzero(got != want)[0]
        d = np.diagonal(S)
        for i in diff:        # every disagreement must be explained by a near-tie with the true pair's score
            line = S[i] if axis == 1 else S[:, i]
            near = np.sum(np.abs(line - d[i]) <= 2e-6) - 1
            assert abs(int(got[i]) - int(want[i])) <= near, (i, got[i], want[i], near)
        assert len(diff) <= max(1, N // 200)
        assert cf.recall_log(got, "x") == cf.recall_log(want, "x")
    assert 0.01 < (ofwd < 1).mean() < 0.999           # the case is neither trivial nor hopeless


def test_compute_metric_ret_with_several_captions_per_image():
    """ids_txt maps two captions to each image: 'backward' takes the best rank over an image's captions (:396-400)."""
    g = torch.Generator(device="cuda").manual_seed(5)
    n_img, n_txt, D = 200, 400, 32
    img = torch.nn.functional.normalize(torch.randn(n_img, D, generator=g, device="cuda"), dim=-1)
    txt = torch.nn.functional.normalize(img.repeat_interleave(2, dim=0) + 0.9 * torch.randn(n_txt, D, generator=g, device="cuda"), dim=-1)
    S = txt @ img.t()
    ids, ids_txt = list(range(n_img)), [k // 2 for k in range(n_txt)]
    ofwd, obwd = cf.retrieval_ranks(S.double().cpu().numpy(), ids, ids_txt)
    assert M.compute_metric_ret(S, ids, ids_txt, "forward") == cf.recall_log(ofwd, "forward")
    assert M.compute_metric_ret(S, ids, ids_txt, "backward") == cf.recall_log(obwd, "backward")
    with pytest.raises(AssertionError):
        M.compute_metric_ret(S, ids[:-1], ids_txt)


@pytest.mark.parametrize("B,D,dtype,tol", [(300, 96, torch.float32, 1e-5), (1000, 512, torch.bfloat16, 1e-5), (129, 44, torch.float32, 1e-5),
                                           (2048, 768, torch.float16, 1e-5)])
def test_sparsify_loss_gradient_vs_oracle(B, D, dtype, tol):
    """sparsify_loss backward through the D x D second moment (no second B x B pass): value and gradient against
    cf.sparsify_loss on the same (rounded) rows; gradients returned as fp32 leaves hold the rounded values."""
    g = torch.Generator(device="cuda").manual_seed(B + D)
    x = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device="cuda") + 0.2, dim=-1).to(dtype).float()
    prev = scb.set_fp32_mode("bf16" if dtype == torch.bfloat16 else "exact")
    try:
        xg = (x.to(dtype) if dtype == torch.float16 else x).clone().requires_grad_(True)
        loss = scb.sparsify_loss(xg)
        (loss * 3.0).backward()
    finally:
        scb.set_fp32_mode(prev)
    ref, dX = cf.sparsify_loss(x.double().cpu().numpy())
    assert abs(loss.item() - ref) <= 1e-5 * abs(ref)
    got = xg.grad.double().cpu().numpy() / 3.0
    gtol = tol if xg.grad.dtype == torch.float32 else 1.5e-3        # fp16 leaves: the returned gradient is rounded to 11 bits
    assert np.linalg.norm(got - dX) <= gtol * np.linalg.norm(dX), np.linalg.norm(got - dX) / np.linalg.norm(dX)


@pytest.mark.parametrize("p", [2, 1, 3])
def test_centroid_alignment_loss_value_and_gradient(p):
    g = torch.Generator(device="cuda").manual_seed(11)
    I = torch.nn.functional.normalize(torch.randn(500, 72, generator=g, device="cuda") + 0.1, dim=-1)
    T = torch.nn.functional.normalize(torch.randn(500, 72, generator=g, device="cuda") - 0.1, dim=-1)
    a, b = I.clone().requires_grad_(True), T.clone().requires_grad_(True)
    loss = scb.centroid_alignment_loss(a, b, p=p)
    (loss * 5.0).backward()
    assert loss.item() == pytest.approx(cf.centroid_alignment_loss(I.cpu().numpy(), T.cpu().numpy(), p), rel=1e-5)
    # the reference's own formula through autograd, in fp64
    a2, b2 = I.double().clone().requires_grad_(True), T.double().clone().requires_grad_(True)
    (torch.norm(a2.mean(dim=0) - b2.mean(dim=0), p=p) * 5.0).backward()
    assert ((a.grad.double() - a2.grad).norm() / a2.grad.norm()).item() <= 1e-5
    assert ((b.grad.double() - b2.grad).norm() / b2.grad.norm()).item() <= 1e-5


def test_gram_and_column_kernels_vs_fp64():
    """scb_gram_dd / scb_col_sum / scb_rows_times_dd on ragged shapes (rows and D not multiples of the tiles), all dtypes."""
    be = scb.get_backend()
    g = torch.Generator(device="cuda").manual_seed(3)
    for n, D, dt in ((1000, 100, torch.float32), (37, 8, torch.float32), (5000, 512, torch.bfloat16), (130, 264, torch.float16)):
        x = torch.randn(n, D, generator=g, device="cuda").to(dt)
        y = torch.randn(n, D, generator=g, device="cuda").to(dt)
        xd, yd = x.double(), y.double()
        mu = be.col_sum(x, None, 1.0 / n)
        assert (mu.double() - xd.mean(0)).abs().max().item() <= 1e-5
        assert (be.col_sum(x, y).double() - (xd - yd).sum(0)).abs().max().item() <= 1e-4 * n ** 0.5
        cov = be.gram_dd(x, mu, 1.0 / n)
        xc = xd - mu.double()
        want = xc.t() @ xc / n
        assert (cov.double() - want).abs().max().item() <= 1e-5 * want.abs().max().item()
        Mx = torch.randn(D, D, generator=g, device="cuda")
        out = be.rows_times_dd(x, Mx)
        assert ((out.double() - xd @ Mx.double()).norm() / (xd @ Mx.double()).norm()).item() <= 1e-5
