"""CPU: pin the oracle (closed_form + torch_port) against the golden vectors minted from
the reference, and -- when /root/reference is present -- against the reference itself."""
import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import ref_loader, torch_port
from tests import _golden

CASES = _golden.names()


def test_golden_present():
    assert len(CASES) >= 10


@pytest.mark.parametrize("name", CASES)
def test_closed_form_matches_golden(name):
    z = _golden.load(name)
    I, T, tau = z["I"], z["T"], float(z["tau"])
    a, dI, dT, dtau = cf.contrastive_loss(I, T, tau)
    assert a == pytest.approx(float(z["f64_anchor"]), rel=1e-11, abs=1e-12)
    assert dtau == pytest.approx(float(z["f64_anchor_dtau"]), rel=1e-9, abs=1e-12)
    _golden.grad_check(z, "anchor_dI", dI, 1e-6)
    _golden.grad_check(z, "anchor_dT", dT, 1e-6)
    al, gx, gy = cf.lalign_loss(I, T)
    assert al == pytest.approx(float(z["f64_lalign"]), rel=1e-12)
    _golden.grad_check(z, "lalign_dI", gx, 1e-6)
    _golden.grad_check(z, "lalign_dT", gy, 1e-6)
    for key, x in (("lunif_img", I), ("lunif_txt", T)):
        u, g = cf.lunif_loss(x)
        assert u == pytest.approx(float(z["f64_" + key]), rel=1e-10)
        # duplicate rows: pdist's sqrt has a 0 subgradient there, identical in the closed form
        _golden.grad_check(z, key + "_dX", g, 1e-6)
    c = cf.normalized_centroids(I, T)
    u, g = cf.lunif_loss(c)
    assert u == pytest.approx(float(z["f64_lunif_cen"]), rel=1e-10)
    da, db = cf.normalized_centroids_backward(I, T, g)
    _golden.grad_check(z, "lunif_cen_dI", da, 1e-6)
    _golden.grad_check(z, "lunif_cen_dT", db, 1e-6)
    l3, dI3, dT3, dtau3, _ = cf.weighted_loss(I, T, tau, 1.0, 1.0, 0.5, 0.5, 0.0)
    assert l3 == pytest.approx(float(z["f64_exp3"]), rel=1e-10, abs=1e-12)
    _golden.grad_check(z, "exp3_dI", dI3, 1e-6)
    _golden.grad_check(z, "exp3_dT", dT3, 1e-6)
    assert dtau3 == pytest.approx(float(z["f64_exp3_dtau"]), rel=1e-9, abs=1e-12)
    l4, dI4, dT4, dtau4, _ = cf.weighted_loss(I, T, tau, 1.0, 1.0, 0.0, 0.0, 1.0)
    assert l4 == pytest.approx(float(z["f64_exp4"]), rel=1e-10, abs=1e-12)
    _golden.grad_check(z, "exp4_dI", dI4, 1e-6)
    _golden.grad_check(z, "exp4_dT", dT4, 1e-6)
    assert cf.sparsify_loss(I, need_grad=False) == pytest.approx(float(z["f64_sparsify_img"]), rel=1e-10)


@pytest.mark.parametrize("name", [n for n in CASES if "b128" in n or "b129" in n or "b3_" in n])
def test_torch_port_matches_golden_fp32(name):
    """torch_port in fp32 reproduces what the reference computed in fp32 (same ops, same order)."""
    z = _golden.load(name)
    I, T = torch.from_numpy(z["I"]), torch.from_numpy(z["T"])
    tau = float(z["tau"])
    assert torch_port.anchor(I, T, tau).item() == pytest.approx(float(z["f32_anchor"]), rel=1e-6)
    assert torch_port.lunif(I).item() == pytest.approx(float(z["f32_lunif_img"]), rel=1e-6)
    assert torch_port.lalign(I, T).item() == pytest.approx(float(z["f32_lalign"]), rel=1e-6)
    assert torch_port.lunif(torch_port.centroids(I, T)).item() == pytest.approx(float(z["f32_lunif_cen"]), rel=1e-6)
    loss, dI, dT, dtau = torch_port.fwd_bwd(I.double(), T.double(), torch.tensor(tau, dtype=torch.float64),
                                            (1.0, 1.0, 0.5, 0.5, 0.0))
    assert loss.item() == pytest.approx(float(z["f64_exp3"]), rel=1e-10)
    assert dtau.item() == pytest.approx(float(z["f64_exp3_dtau"]), rel=1e-8)
    _golden.grad_check(z, "exp3_dI", dI.numpy(), 1e-6)


def test_survey_sanity_anchors():
    """SURVEY.md §8c probe values (seed 0, B=128, D=512, tau=0.1)."""
    z = _golden.load("b128_d512_iid_s0")
    assert float(z["f32_anchor"]) == pytest.approx(4.931442, abs=2e-6)
    assert float(z["f64_anchor_dtau"]) == pytest.approx(-1.779671, abs=2e-6)
    assert float(z["f32_lalign"]) == pytest.approx(1.996119, abs=2e-6)
    assert float(z["f32_lunif_img"]) == pytest.approx(-3.985226, abs=2e-6)
    assert float(z["f32_lunif_txt"]) == pytest.approx(-3.984154, abs=2e-6)
    assert np.linalg.norm(z["f64_anchor_dI"]) == pytest.approx(0.881004, abs=2e-6)


def test_schedules_and_edge_cases():
    assert cf.get_beta(10, 1000, 20, 50) == 1.0
    assert cf.get_beta(450, 1000, 20, 50) == pytest.approx(0.5)
    assert cf.get_beta(900, 1000, 20, 50) == 0.0
    assert cf.get_alpha(10, 1000, 50, 50) == 1.0
    assert cf.get_alpha(750, 1000, 50, 50) == pytest.approx(1.5)
    assert cf.get_alpha(2000, 1000, 50, 50) == 2.0
    assert np.isnan(cf.lunif_loss(np.ones((1, 4)), need_grad=False))
    x = np.array([[1.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    _, gx, gy = cf.lalign_loss(x, x)
    assert np.all(gx == 0) and np.all(gy == 0)


needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


@needs_ref
@pytest.mark.parametrize("B,D,tau", [(5, 16, 0.07), (64, 32, 0.1), (33, 24, 1.0)])
def test_closed_form_vs_reference_autograd(B, D, tau):
    ref, _ = ref_loader.load()
    g = torch.Generator().manual_seed(B)
    I = torch.nn.functional.normalize(torch.randn(B, D, generator=g, dtype=torch.float64), dim=-1).requires_grad_(True)
    T = torch.nn.functional.normalize(torch.randn(B, D, generator=g, dtype=torch.float64), dim=-1).requires_grad_(True)
    tp = torch.nn.Parameter(torch.tensor(tau, dtype=torch.float64))
    loss = ref.contrastive_loss(I, T, tp)
    gi, gt, gtau = torch.autograd.grad(loss, (I, T, tp))
    a, dI, dT, dtau = cf.contrastive_loss(I.detach().numpy(), T.detach().numpy(), tau)
    assert a == pytest.approx(loss.item(), rel=1e-13)
    assert np.abs(dI - gi.numpy()).max() < 1e-14 and np.abs(dT - gt.numpy()).max() < 1e-14
    assert dtau == pytest.approx(gtau.item(), rel=1e-11)
    u = ref.lunif_loss(I)
    (gu,) = torch.autograd.grad(u, (I,))
    uu, dX = cf.lunif_loss(I.detach().numpy())
    assert uu == pytest.approx(u.item(), rel=1e-13)
    assert np.abs(dX - gu.numpy()).max() < 1e-13
    assert cf.sparsify_loss(I.detach().numpy(), need_grad=False) == pytest.approx(ref.sparsify_loss(I).item(), rel=1e-12)
    assert cf.centroid_alignment_loss(I.detach().numpy(), T.detach().numpy()) == pytest.approx(
        ref.centroid_alignment_loss(I, T).item(), rel=1e-12)
    R = torch.softmax(torch.randn(B, B, generator=g, dtype=torch.float64), dim=1)
    assert cf.contrastive_loss_roberta(I.detach().numpy(), T.detach().numpy(), R.numpy(), tau) == pytest.approx(
        ref.contrastive_loss_roberta(I, T, R, tau).item(), rel=1e-12)
    for step in (0, 150, 200, 450, 699, 700, 5000):
        assert cf.get_beta(step, 1000, 20, 50) == ref.get_beta(step, 1000, 20, 50)
        assert cf.get_alpha(step, 1000, 50, 50) == ref.get_alpha(step, 1000, 50, 50)


@pytest.mark.parametrize("B,D,slab,w", [(200, 24, 64, (1.0, 1.0, 0.5, 0.5, 0.0)), (131, 16, 50, (1.0, 1.3, 0.0, 0.0, 0.7)),
                                        (96, 8, 4096, (1.0, 0.7, 0.25, 0.25, 0.5))])
def test_chunked_oracle_matches_the_closed_forms(B, D, slab, w):
    """tests/_chunked_oracle.py (the slab-wise fp64 restatement used for full-size GPU parity) against
    oracle/closed_form.py, which is pinned against the reference above: every term, sampled gradient rows, d/dtau;
    slab sizes that do and do not divide B, duplicate rows included."""
    from tests import _chunked_oracle as co
    g = torch.Generator().manual_seed(B + D)
    I = torch.nn.functional.normalize(torch.randn(B, D, generator=g, dtype=torch.float64), dim=-1)
    T = torch.nn.functional.normalize(I + 0.5 * torch.randn(B, D, generator=g, dtype=torch.float64), dim=-1)
    I[5] = I[17]                                   # an exact duplicate pair (d^2 clamps at 0)
    rows = torch.tensor([0, 5, 17, B // 2, B - 1])
    tau = 0.1
    ref, dI, dT, dtau, terms = cf.weighted_loss(I.numpy(), T.numpy(), tau, *w)
    got, gI, gT, gtau, gterms = co.weighted(I, T, tau, *w, rows, slab=slab)
    assert got == pytest.approx(ref, rel=1e-12, abs=1e-12)
    for k, v in terms.items():
        assert gterms[k] == pytest.approx(v, rel=1e-12)
    assert np.abs(gI.numpy() - dI[rows.numpy()]).max() <= 1e-13
    assert np.abs(gT.numpy() - dT[rows.numpy()]).max() <= 1e-13
    assert gtau == pytest.approx(dtau, rel=1e-11)
