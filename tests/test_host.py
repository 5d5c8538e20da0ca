"""CPU suite: C-ABI surface, host logic (ladder, schedules, work split), loud failure without CUDA,
and the sharded (world_size 2, gloo) algebra of losses.py against the oracle."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sparsify_clip_b200 as scb
from oracle import closed_form as cf
from sparsify_clip_b200 import _lib, backend_cuda
from tests._fake_backend import FakeBackend

YAML_LOSS_TYPES = [  # loss_type strings of the 10 experiment + 3 ablation YAMLs (experiments_configs/*.yaml:19)
    ("anchor", 0), ("anchor", 0),
    ("only_lunif_n_then_anchor+lalign+lunif(text)+lunif(img)", 0),
    ("only_lunif_n_then_anchor+lalign+lunif(centroids)", 0),
    ("only_lunif_n_then_anchor+lalign+lunif(text)+lunif(img)", 1),
    ("only_lunif_n_then_anchor+lalign+lunif(centroids)", 1),
    ("only_lunif_n_then_anchor+lalign+BETA*lunif(centroids)", 0),
    ("only_lunif_n_then_anchor+lalign+BETA*lunif(centroids)", 0),
    ("only_lunif_n_then_anchor+ALPHA*lalign+BETA*(lunif(text)+lunif(img))", 0),
    ("only_lunif_n_then_anchor+ALPHA*lalign+BETA*lunif(centroids)", 0),
    ("ANCHOR(IMAGE,TEXT)+LALIGN(IMAGE,TEXT)+LUNIF(CENTROIDS)", 0),
    ("ANCHOR(IMAGE,TEXT)+LALIGN(IMAGE,TEXT)", 0),
    ("ANCHOR(IMAGE,TEXT)+LUNIF(CENTROIDS)", 0),
]


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _lib.header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/scb200.h but not exported"
    assert set(_lib.SIGNATURES) | {"scb_last_error"} == set(declared)
    assert _lib.load().scb_version() >= 100


def _plan(path, nA, nB, D, grad, n_sm=148):
    jp, nsub = ctypes.c_int(0), ctypes.c_int(0)
    assert _lib.load().scb_pass_plan(path, nA, nB, D, grad, n_sm, ctypes.byref(jp), ctypes.byref(nsub)) == 0
    return jp.value, nsub.value


def test_pass_plan_matches_the_host_mirror():
    """scb_pass_plan (the planner the backend uses) against choose_jparts, plus the nsub contract."""
    assert _plan(_lib.PATH_SIMT, 1000, 1000, 512, 1)[1] == 1
    for flags in (1, 3):
        prev = _lib.load().scb_set_tc_flags(flags)
        try:
            _check_tc_plans(pair=bool(flags & 2))
        finally:
            _lib.load().scb_set_tc_flags(prev)


def _check_tc_plans(pair):
    for nA, nB, D in [(32768, 32768, 512), (4096, 4096, 512), (1000, 1000, 768), (4096, 32768, 512), (130, 130, 72),
                      (65536 // 8, 65536, 768), (128, 128, 256)]:
        n_rb, n_jb, kch = (nA + 127) // 128, (nB + 127) // 128, (D + 63) // 64
        assert _plan(_lib.PATH_TC, nA, nB, D, 0) == (scb.choose_jparts(n_rb, 1, n_jb, 148), 2)
        jp, nsub = _plan(_lib.PATH_TC, nA, nB, D, 1)
        if pair and 4 < kch <= 8:      # gradient passes on CTA pairs: equal contiguous tile spans per pair
            assert (jp, nsub) == (scb.pair_span_plan(n_rb, n_jb, 148)[2], 4)
        else:
            assert (jp, nsub) == (scb.choose_jparts(n_rb, (kch + 3) // 4, n_jb, 148), 2)


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.load()
    rc = lib.scb_row_sqnorm(None, 4, 8, 8, 7, None, None)       # bad dtype
    assert rc == -2 and b"dtype" in lib.scb_last_error()
    rc = lib.scb_lse_pass(None, 4, None, 4, 8, 8, 8, _lib.SCB_BF16, 1.0, 0, None, None, _lib.PATH_TC, None, None)
    assert rc == -1
    with pytest.raises(ValueError):
        _lib.check(rc, "lse_pass")
    # the packed-gather fold: offsets must lie inside one packed row (non-null dummy pointers; nothing is launched)
    buf = (ctypes.c_float * 4)()
    flag = (ctypes.c_int * 1)()
    p, f = ctypes.addressof(buf), ctypes.addressof(flag)
    assert lib.scb_lse2_fold_ranks(p, 2, 100, 10, 10, 25, 90, f, p, None) == -1      # sums run past the row
    assert b"offsets" in lib.scb_last_error()
    assert lib.scb_lse2_fold_ranks(p, 0, 100, 10, 10, 25, 45, f, p, None) == -1      # world = 0
    assert lib.scb_lse2_fold_ranks(p, 2, 100, 0, 0, 0, 0, f, p, None) == 0           # nothing to do
    assert lib.scb_loss_assemble(None, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0, p, p, None, None) == -1
    # the fused combine: a term's inputs must come together, the output layout must hold a row
    assert lib.scb_grad_combine(p, None, 4, 8, 8, 8, _lib.SCB_BF16, p, 1, None, None, None, 1.0, 1.0, None, 0, None, 0, 0.0,
                                None, 0.0, None, 0.0, None, p, _lib.SCB_BF16, 8, None, None, 0, 0, None, None) == -1
    assert b"anchor" in lib.scb_last_error()
    assert lib.scb_grad_combine(p, p, 4, 8, 8, 8, _lib.SCB_BF16, None, 0, None, None, None, 1.0, 1.0, None, 0, None, 0, 0.0,
                                None, 1.0, None, 0.0, None, p, _lib.SCB_BF16, 4, None, None, 0, 0, None, None) == -1
    # the fused normalise: the un-normalised rows and 1/norm come together
    assert lib.scb_grad_combine(p, p, 4, 8, 8, 8, _lib.SCB_BF16, None, 0, None, None, None, 1.0, 1.0, None, 0, None, 0, 0.0,
                                None, 1.0, None, 0.0, None, p, _lib.SCB_BF16, 8, None, p, 8, _lib.SCB_BF16, None, None) == -1
    assert b"normalise" in lib.scb_last_error()
    # the SM push of the peer gather: 16-byte granularity of the shard and of every pointer; the plan helper's arguments
    dst = (ctypes.c_void_p * 2)(p, p)
    assert lib.scb_peer_push_sm(p, 24, dst, 2, f, p, 2, f, None) == -1 and b"16-byte" in lib.scb_last_error()
    assert lib.scb_peer_push_sm(p, 32, dst, 0, f, p, 2, f, None) == -1
    assert lib.scb_peer_push_sm(None, 32, dst, 2, f, p, 2, f, None) == -1
    n_used, span, pmax = ctypes.c_int(0), ctypes.c_int64(0), ctypes.c_int(0)
    assert lib.scb_quad_plan(0, 4, 33, 0, ctypes.byref(n_used), ctypes.byref(span), ctypes.byref(pmax)) == -1


def test_no_cpu_fallback():
    x = torch.randn(8, 16)
    with pytest.raises(RuntimeError, match="CUDA"):
        scb.lunif_loss(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        scb.contrastive_loss(x, x, 0.1)


@pytest.mark.parametrize("lt,warm", YAML_LOSS_TYPES)
def test_ladder_matches_oracle(lt, warm):
    cfg = {"loss_type": lt, "only_lunif_epochs": warm, "beta_warmup_epoch": 20, "beta_decay_epoch": 50,
           "alpha_warmup_epoch": 50, "alpha_increment_epoch": 50}
    for epoch in (0, 1, 30):
        for step in (1, 150, 330, 777, 5000):
            w = scb.ladder_weights(cfg, epoch, step, 1000)
            ref = cf.ladder_terms(cfg, epoch, step, 1000)
            got = (w["anchor"], w["align"], w["unif_img"], w["unif_txt"], w["unif_cen"])
            assert got == pytest.approx(ref, abs=0), (lt, epoch, step)


def test_ladder_quirks():
    cfg = {"loss_type": "only_lunif_n_then_anchor+lalign+BETA*lunif(centroids)", "only_lunif_epochs": 0,
           "beta_warmup_epoch": 20, "beta_decay_epoch": 50}
    w = scb.ladder_weights(cfg, 0, 450, 1000)          # exp 7 AND exp 8 run the modality variant (:813 wins)
    assert w["unif_cen"] == 0.0 and w["unif_img"] == pytest.approx(0.25)
    cfg["loss_type"] += "[intended]"
    assert scb.ladder_weights(cfg, 0, 450, 1000)["unif_cen"] == pytest.approx(0.5)
    with pytest.raises(KeyError):
        scb.ladder_weights({"loss_type": "nope", "only_lunif_epochs": 0}, 0, 1, 10)
    for s in (0, 199, 200, 450, 700, 10 ** 6):
        assert scb.get_beta(s, 1000, 20, 50) == cf.get_beta(s, 1000, 20, 50)
        assert scb.get_alpha(s, 1000, 50, 50) == cf.get_alpha(s, 1000, 50, 50)
    assert scb.get_beta(5, 10) == cf.get_beta(5, 10) and scb.get_alpha(5, 10) == cf.get_alpha(5, 10)


def test_reference_yaml_strings_if_present():
    import glob
    files = glob.glob("/root/reference/experiments_configs/*.yaml") + glob.glob("/root/reference/ablatation_configs/*.yaml")
    if not files:
        pytest.skip("reference not present")
    import yaml
    seen = 0
    for f in files:
        cfg = yaml.safe_load(open(f))
        if not cfg:
            continue
        assert cfg["loss_type"] in scb.LOSS_TYPES, f
        seen += 1
    assert seen == 13


def test_choose_jparts():
    assert scb.choose_jparts(256, 2, 256, 148) == 2          # c3 gradient sweeps: 1024 items -> 7 rounds
    assert scb.choose_jparts(1, 1, 1, 148) == 1
    for n_rb, ns, n_jb in [(32, 2, 32), (8, 2, 8), (256, 1, 256), (512, 3, 512), (3, 1, 100)]:
        jp = scb.choose_jparts(n_rb, ns, n_jb, 148)
        assert 1 <= jp <= min(n_jb, 16)


def test_uniformity_signatures_run():
    from sparsify_clip_b200 import uniformity as u
    x = torch.nn.functional.normalize(torch.randn(64, 16, dtype=torch.float64), dim=-1)
    a = u.torch_uniformity1(x).item()
    assert a == pytest.approx(u.torch_uniformity_equivalent(x).item(), rel=1e-5)
    assert u.torch_uniformity(x, x).item() < 0 and u.numpy_uniformity(x, x) < 0
    assert u.uniformity10(x).item() > 0


# ----------------------------------------------------------------------------- sharded algebra (gloo, 2 ranks)
def _worker(rank, world, port, I, T, tau, out, cen=0.7, fused=True):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    backend_cuda.set_backend(FakeBackend())
    scb.set_fused(fused)
    n = I.shape[0] // world
    Il = I[rank * n:(rank + 1) * n].clone().requires_grad_(True)
    Tl = T[rank * n:(rank + 1) * n].clone().requires_grad_(True)
    tp = torch.nn.Parameter(torch.tensor(tau, dtype=torch.float64))
    w = dict(anchor=1.0, align=1.5, unif_img=0.5, unif_txt=0.25, unif_cen=cen)
    loss = scb.weighted_loss(Il, Tl, tp, w, group=True)
    (loss * 2.0).backward()
    out[rank] = (loss.item(), Il.grad.numpy() / 2.0, Tl.grad.numpy() / 2.0, tp.grad.item() / 2.0)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("cen", [0.0, 0.7])
def test_sharded_fused_composition_two_ranks(cen):
    """The fused autograd node, sharded: one-sweep LSE with the column sums folded over the ranks, the packed
    (r, c, scalars, column partials) gather, the centroid chain and scb_grad_combine, against the full-batch oracle."""
    g = torch.Generator().manual_seed(11)
    B, D, tau = 24, 16, 0.2
    I = torch.nn.functional.normalize(torch.randn(B, D, generator=g, dtype=torch.float64), dim=-1)
    T = torch.nn.functional.normalize(I + 0.5 * torch.randn(B, D, generator=g, dtype=torch.float64), dim=-1)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 30517 + os.getpid() % 1000 + int(cen * 10), I, T, tau, out, cen, True), nprocs=2, join=True)
    ref_loss, dI, dT, dtau, _ = cf.weighted_loss(I.numpy(), T.numpy(), tau, 1.0, 1.5, 0.5, 0.25, cen)
    n = B // 2
    for r in (0, 1):
        loss, gI, gT, gtau = out[r]
        assert loss == pytest.approx(ref_loss, rel=1e-6)         # scalars travel as fp32 in the packed gather
        assert np.abs(gI - dI[r * n:(r + 1) * n]).max() < 1e-6
        assert np.abs(gT - dT[r * n:(r + 1) * n]).max() < 1e-6
        assert gtau == pytest.approx(dtau, rel=1e-5)


@pytest.mark.timeout(120)
def test_sharded_two_ranks_match_full_batch_oracle():
    g = torch.Generator().manual_seed(7)
    B, D, tau = 24, 16, 0.2
    I = torch.nn.functional.normalize(torch.randn(B, D, generator=g, dtype=torch.float64), dim=-1)
    T = torch.nn.functional.normalize(I + 0.5 * torch.randn(B, D, generator=g, dtype=torch.float64), dim=-1)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29517 + os.getpid() % 1000, I, T, tau, out, 0.7, False), nprocs=2, join=True)
    ref_loss, dI, dT, dtau, _ = cf.weighted_loss(I.numpy(), T.numpy(), tau, 1.0, 1.5, 0.5, 0.25, 0.7)
    n = B // 2
    for r in (0, 1):
        loss, gI, gT, gtau = out[r]
        assert loss == pytest.approx(ref_loss, rel=1e-12)
        assert np.abs(gI - dI[r * n:(r + 1) * n]).max() < 1e-7   # grad_output travels as fp32
        assert np.abs(gT - dT[r * n:(r + 1) * n]).max() < 1e-7
        assert gtau == pytest.approx(dtau, rel=1e-6)


@pytest.mark.parametrize("fused", [False, True])
def test_single_process_fake_backend_matches_oracle(fused):
    prev = backend_cuda.set_backend(FakeBackend())
    pf = scb.set_fused(fused)
    try:
        g = torch.Generator().manual_seed(3)
        I = torch.nn.functional.normalize(torch.randn(17, 8, generator=g, dtype=torch.float64), dim=-1).requires_grad_(True)
        T = torch.nn.functional.normalize(torch.randn(17, 8, generator=g, dtype=torch.float64), dim=-1).requires_grad_(True)
        cfg = {"loss_type": "only_lunif_n_then_anchor+ALPHA*lalign+BETA*lunif(centroids)", "only_lunif_epochs": 0,
               "beta_warmup_epoch": 20, "beta_decay_epoch": 50, "alpha_warmup_epoch": 50, "alpha_increment_epoch": 50}
        loss = scb.compose_loss(cfg, I, T, 0.1, epoch=3, current_batch=600, t_total=1000)
        loss.backward()
        ref_loss, dI, dT, _, _ = cf.compose_loss(cfg, I.detach().numpy(), T.detach().numpy(), 0.1, 3, 600, 1000)
        # the fused node keeps its scalar partial sums in fp32 (they travel in the packed gather when sharded)
        assert loss.item() == pytest.approx(ref_loss, rel=1e-6 if fused else 1e-12)
        assert np.abs(I.grad.numpy() - dI).max() < 1e-6 and np.abs(T.grad.numpy() - dT).max() < 1e-6
    finally:
        scb.set_fused(pf)
        backend_cuda.set_backend(prev)


@pytest.mark.parametrize("fused,D", [(True, 8), (True, 12), (False, 8)])
def test_normalize_inside_the_composition_matches_the_explicit_chain(fused, D):
    """compose_loss(..., normalize=True) on un-normalised encoder outputs (sparsify_clip.py:768-773 inside the call) ==
    compose_loss on l2_normalize(...) of them, value and gradients w.r.t. the un-normalised rows; both routes of the
    fused node's backward (normalise backward inside the combine pass when D % 8 == 0, as a separate pass otherwise)
    and the modular path; the gradient is also checked against the oracle's normalise backward."""
    prev = backend_cuda.set_backend(FakeBackend())
    pf = scb.set_fused(fused)
    try:
        g = torch.Generator().manual_seed(5)
        E_I = (torch.randn(19, D, generator=g, dtype=torch.float64) * (0.5 + 3.0 * torch.rand(19, 1, generator=g, dtype=torch.float64)))
        E_T = E_I + 0.7 * torch.randn(19, D, generator=g, dtype=torch.float64)
        cfg = {"loss_type": "only_lunif_n_then_anchor+lalign+lunif(text)+lunif(img)", "only_lunif_epochs": 0}
        got = []
        for inside in (True, False):
            a, b = E_I.clone().requires_grad_(True), E_T.clone().requires_grad_(True)
            if inside:
                loss = scb.compose_loss(cfg, a, b, 0.1, epoch=1, current_batch=10, t_total=100, normalize=True)
            else:
                loss = scb.compose_loss(cfg, scb.l2_normalize(a), scb.l2_normalize(b), 0.1, epoch=1, current_batch=10, t_total=100)
            (loss * 3.0).backward()
            got.append((loss.item(), a.grad.numpy() / 3.0, b.grad.numpy() / 3.0))
        assert got[0][0] == pytest.approx(got[1][0], rel=1e-12)
        # (the explicit chain hands the loss node's gradient to the normalise node as fp32: 1e-8 relative)
        assert np.abs(got[0][1] - got[1][1]).max() < 1e-7 and np.abs(got[0][2] - got[1][2]).max() < 1e-7
        yI, yT = cf.l2_normalize(E_I.numpy()), cf.l2_normalize(E_T.numpy())
        ref_loss, dyI, dyT, _, _ = cf.compose_loss(cfg, yI, yT, 0.1, 1, 10, 100)
        assert got[0][0] == pytest.approx(ref_loss, rel=1e-6)
        assert np.abs(got[0][1] - cf.l2_normalize_backward(E_I.numpy(), dyI)).max() < 1e-6
        assert np.abs(got[0][2] - cf.l2_normalize_backward(E_T.numpy(), dyT)).max() < 1e-6
    finally:
        scb.set_fused(pf)
        backend_cuda.set_backend(prev)


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/scb200.h compiles as C99 and a C program linked against libscb200.so reaches the host-only entry points
    (version, launch plan, argument validation) without Python or PyTorch in the process."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "sparsify_clip_b200")
    _lib.load()                                              # builds the library if it is missing
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "scb200.h"
int main(void) {
  int jparts = 0, nsub = 0;
  if (scb_pass_plan(SCB_PATH_TC, 32768, 32768, 512, 1, 148, &jparts, &nsub) != 0) return 2;
  int rc = scb_row_sqnorm(NULL, 4, 8, 8, 7, NULL, NULL);
  printf("%d %d %d %d %d\n", scb_version() > 0, jparts, nsub, rc, strstr(scb_last_error(), "dtype") != NULL);
  return 0;
}
''')
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lscb200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    want_jp, want_nsub = _plan(_lib.PATH_TC, 32768, 32768, 512, 1)
    assert out == ["1", str(want_jp), str(want_nsub), "-2", "1"], out


def _pair_segments(n_rb, n_jb, n_sm):
    """Host mirror of the CTA-pair kernel's span walk (csrc/tc_pair.cu: SCB_PAIR_FOR_SEGMENTS / SCB_PAIR_ITEM_SETUP):
    yields (pair, row block, first tile, tiles, partial slot)."""
    pairs, span, pmax = scb.pair_span_plan(n_rb, n_jb, n_sm)
    total = n_rb * n_jb
    for pair in range(pairs):
        g, g_end = pair * span, min(pair * span + span, total)
        while g < g_end:
            rb = g // n_jb
            jb_lo = g - rb * n_jb
            nt = min(n_jb - jb_lo, g_end - g)
            jp = pair - (rb * n_jb) // span
            yield pair, rb, jb_lo, nt, jp
            g += nt
    return pmax


def test_pair_span_walk_covers_every_tile_once():
    """Every (row block, column tile) belongs to exactly one segment, the partial slots of a row block are
    0 .. k-1 with k <= the planner's slot count, and the C planner agrees with the mirror -- over many shard shapes."""
    import random
    rng = random.Random(5)
    shapes = [(256, 256, 148), (32, 256, 148), (1, 1, 148), (3, 7, 148), (64, 512, 148), (2, 300, 148), (5, 33, 4)]
    shapes += [(rng.randint(1, 80), rng.randint(1, 600), rng.choice([2, 8, 132, 148])) for _ in range(120)]
    lib = _lib.load()
    for n_rb, n_jb, n_sm in shapes:
        pairs, span, pmax = scb.pair_span_plan(n_rb, n_jb, n_sm)
        jp_c, nsub_c = ctypes.c_int(0), ctypes.c_int(0)
        prev = lib.scb_set_tc_flags(3)
        try:
            assert lib.scb_pass_plan(_lib.PATH_TC, n_rb * 128, n_jb * 128, 512, 1, n_sm, ctypes.byref(jp_c), ctypes.byref(nsub_c)) == 0
        finally:
            lib.scb_set_tc_flags(prev)
        assert (jp_c.value, nsub_c.value) == (pmax, 4), (n_rb, n_jb, n_sm)
        assert pmax <= 16 and pairs <= max(1, n_sm // 2) and pairs * span >= n_rb * n_jb
        seen = [[0] * n_jb for _ in range(n_rb)]
        slots = [[] for _ in range(n_rb)]
        for pair, rb, jb_lo, nt, jp in _pair_segments(n_rb, n_jb, n_sm):
            assert nt >= 1 and 0 <= jp < pmax, (n_rb, n_jb, n_sm, pair, rb)
            for j in range(jb_lo, jb_lo + nt):
                seen[rb][j] += 1
            slots[rb].append(jp)
        assert all(v == 1 for row in seen for v in row), (n_rb, n_jb, n_sm)
        assert all(s == list(range(len(s))) for s in slots), (n_rb, n_jb, n_sm)


def _quad_plan(n_rp, n_jb, n_cl, align):
    n_used, span, pmax = ctypes.c_int(0), ctypes.c_int64(0), ctypes.c_int(0)
    assert _lib.load().scb_quad_plan(n_rp, n_jb, n_cl, align, ctypes.byref(n_used), ctypes.byref(span), ctypes.byref(pmax)) == 0
    return n_used.value, span.value, pmax.value


def test_quad_span_walk_covers_every_tile_once():
    """Host mirror of the cluster-of-4 kernel's span walk (csrc/tc_quad.cu: SCB_QUAD_FOR_SEGMENTS / SCB_QUAD_ITEM_SETUP)
    over the C planner's output, for the equal-span split and for the row-block-aligned one (BASELINE c4's shapes
    included): every (256-row block, tile) in exactly one segment, partial slots of a row block 0 .. k-1 with k <= pmax,
    both pairs of a cluster own tiles in every segment of two or more tiles, aligned spans are whole row blocks."""
    import random
    rng = random.Random(11)
    shapes = [(128, 256, 33), (32, 512, 33), (256, 512, 33), (16, 256, 33), (1, 1, 33), (3, 7, 33), (120, 256, 33), (64, 128, 33)]
    shapes += [(rng.randint(1, 300), rng.randint(1, 600), rng.choice([1, 4, 33, 36])) for _ in range(80)]
    for n_rp, n_jb, n_cl in shapes:
        for align in (0, 1, 2):
            n_used, span, pmax = _quad_plan(n_rp, n_jb, n_cl, align)
            total = n_rp * n_jb
            assert 1 <= n_used <= n_cl and n_used * span >= total and (n_used - 1) * span < total, (n_rp, n_jb, n_cl, align)
            assert 1 <= pmax <= 16
            if align == 2:
                assert span % n_jb == 0 and pmax == 1
            if align == 1:       # accepted only within 6 % of the balanced span
                eq = _quad_plan(n_rp, n_jb, n_cl, 0)
                assert (span % n_jb == 0 and pmax == 1 and span * 100 <= max(eq[1], -(-total // min(n_cl, total))) * 106) or \
                    (n_used, span, pmax) == eq
            seen = [[0] * n_jb for _ in range(n_rp)]
            slots = [[] for _ in range(n_rp)]
            for cl in range(n_used):
                g, g_end, item = cl * span, min(cl * span + span, total), 0
                while g < g_end:
                    rp = g // n_jb
                    jb_lo = g - rp * n_jb
                    nt = min(n_jb - jb_lo, g_end - g)
                    jp = cl - (rp * n_jb) // span
                    own = [(nt - t_first + 1) // 2 if nt > t_first else 0 for t_first in ((0 ^ (item & 1)) & 1, (1 ^ (item & 1)) & 1)]
                    assert sum(own) == nt and (nt < 2 or min(own) >= 1)
                    assert 0 <= jp < pmax, (n_rp, n_jb, n_cl, align, cl, rp, jp, pmax)
                    for j in range(jb_lo, jb_lo + nt):
                        seen[rp][j] += 1
                    slots[rp].append(jp)
                    g += nt
                    item += 1
            assert all(v == 1 for row in seen for v in row), (n_rp, n_jb, n_cl, align)
            assert all(s_ == list(range(len(s_))) for s_ in slots), (n_rp, n_jb, n_cl, align)
    # BASELINE c4: the 8-GPU shard (32 row-block pairs x 512 tiles) runs 32 clusters of one row block each, the single
    # GPU (256 x 512) 32 clusters of 8; c3's 120 x 256 (row split) stays on equal spans (aligned would cost 10 %)
    assert _quad_plan(32, 512, 33, 1) == (32, 512, 1)
    assert _quad_plan(256, 512, 33, 1) == (32, 8 * 512, 1)
    assert _quad_plan(120, 256, 33, 1) == _quad_plan(120, 256, 33, 0)
