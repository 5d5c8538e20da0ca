"""Cluster-of-4 gradient kernel (tc_quad.cu) against the CTA-pair kernel (tc_pair.cu) and the fp64 oracle, then timing."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import sparsify_clip_b200 as scb
from oracle import closed_form as cf

be = scb.get_backend()
W3 = dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0)


def run(B, D, tau, flags, dtype=torch.bfloat16):
    be.lib.scb_set_tc_flags(flags)
    g = torch.Generator(device="cuda").manual_seed(B + D)
    I = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    T = torch.nn.functional.normalize(I + 0.5 * torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    Iq, Tq = I.to(dtype).float(), T.to(dtype).float()
    prev = scb.set_fp32_mode("bf16")
    try:
        Ig, Tg = Iq.clone().requires_grad_(True), Tq.clone().requires_grad_(True)
        tp = torch.nn.Parameter(torch.tensor(tau))
        loss = scb.weighted_loss(Ig, Tg, tp, W3)
        loss.backward()
        torch.cuda.synchronize()
    finally:
        scb.set_fp32_mode(prev)
    return Iq, Tq, loss.item(), Ig.grad.double().cpu().numpy(), Tg.grad.double().cpu().numpy(), tp.grad.item()


import ctypes
u = ctypes.c_int(0)
be.lib.scb_set_tc_flags(7)
print("grad kernel kind at nA=384, D=512:", be.lib.scb_grad_kernel_kind(384, 512, 148, ctypes.byref(u)), "units", u.value, flush=True)
QF = int(os.environ.get("QUAD_FLAGS", "15"))
ok = True
sizes = [(256, 512, 0.1), (384, 512, 0.1), (129, 264, 0.1), (640, 320, 0.1), (385, 384, 0.07), (1300, 448, 0.1), (2100, 456, 0.05),
         (1024, 512, 0.1), (5000, 512, 0.1), (8192, 512, 0.1), (12000, 448, 0.1), (10000, 512, 0.07), (8960, 384, 0.1)]
if len(sys.argv) > 1 and sys.argv[1] == "first":
    sizes = sizes[:2]
for (B, D, tau) in sizes:
    t0 = time.time()
    Iq, Tq, l7, dI7, dT7, dt7 = run(B, D, tau, QF)
    _, _, l3, dI3, dT3, dt3 = run(B, D, tau, 3)
    ref, dI, dT, dtau, _ = cf.weighted_loss(Iq.cpu().numpy(), Tq.cpu().numpy(), tau, 1.0, 1.0, 0.5, 0.5, 0.0)
    e_pair = max(np.linalg.norm(dI7 - dI3) / np.linalg.norm(dI3), np.linalg.norm(dT7 - dT3) / np.linalg.norm(dT3))
    e_or = max(np.linalg.norm(dI7 - dI) / np.linalg.norm(dI), np.linalg.norm(dT7 - dT) / np.linalg.norm(dT))
    e_l = abs(l7 - ref) / abs(ref)
    e_tau = abs(dt7 - dtau) / abs(dtau)
    good = e_pair <= 2e-5 and e_or <= 1e-3 and e_l <= 1e-5 and e_tau <= 1e-3
    ok &= good
    print(f"B={B} D={D}: quad vs pair {e_pair:.1e}; vs oracle grad {e_or:.1e} loss {e_l:.1e} dtau {e_tau:.1e} {'ok' if good else 'FAIL'} ({time.time() - t0:.1f}s)", flush=True)
print("QUAD_CHECK", "PASS" if ok else "FAIL", flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "first":
    sys.exit(0 if ok else 1)

# ---- timing of the sweeps at c3 and at the 8-GPU shard shape
for (nA, nB, D) in [(32768, 32768, 512), (4096, 32768, 512)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.nn.functional.normalize(torch.randn(nB, D, generator=g, device="cuda"), dim=-1).to(torch.bfloat16)
    Xr = X[:nA]
    for flags in (3, 7, 15):
        be.lib.scb_set_tc_flags(flags)
        for _ in range(3):
            be.lunif_core(Xr, X, 2.0, 0, True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        be.pass_events = []
        for _ in range(n):
            be.lunif_core(Xr, X, 2.0, 0, True)
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for _, a, b in be.pass_events) / n
        be.pass_events = None
        print(f"lunif sweep nA={nA} nB={nB} D={D} flags={flags}: {ms:.3f} ms = {4.0 * nA * nB * D / ms / 1e9:.0f} TFLOP/s", flush=True)
sys.exit(0 if ok else 1)
