"""GPU bring-up diagnostics: every pass of libscb200 against dense torch fp64 math on the same
(bf16-rounded) inputs, with error maps per 64-column chunk / 32-row quarter so that a wrong
descriptor or layout shows up as a pattern.  Each case runs in its own subprocess (a device
trap in one case must not take the others down).

usage: python tools/gpu_diag.py            # all cases
       python tools/gpu_diag.py --case tc_lse
"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["simt", "tc_lse", "tc_sums", "tc_grad_smem", "tc_grad_tmem", "e2e_tc", "e2e_simt"]


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def errmap(got, ref, tag):
    """relative error per (32-row quarter of each 128-row block) x (64-col chunk)"""
    import torch
    got, ref = got.double(), ref.double()
    n, D = ref.shape
    scale = ref.norm().item() / (n * D) ** 0.5 + 1e-300
    lines = []
    for r0 in range(0, min(n, 256), 32):
        row = []
        for c0 in range(0, D, 64):
            blk = (got[r0:r0 + 32, c0:c0 + 64] - ref[r0:r0 + 32, c0:c0 + 64])
            row.append(f"{(blk.pow(2).mean().sqrt().item() / scale):8.1e}")
        lines.append(f"    rows {r0:4d}+32: " + " ".join(row))
    print(f"  errmap[{tag}] (rms err / rms ref per 32x64 block):")
    print("\n".join(lines))


def make(B, D, dtype, seed=0, kind="corr"):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    I = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    T = torch.nn.functional.normalize(I + 0.5 * torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    if kind == "cluster":
        I[1] = I[0]
        T[3] = T[2]
    return I.to(dtype), T.to(dtype)


def dense_refs(I, T, tau, t=2.0):
    """fp64 dense references on the stored (possibly bf16) values."""
    import torch
    I64, T64 = I.double(), T.double()
    B = I.shape[0]
    S = I64 @ T64.t() / tau
    r, c = torch.logsumexp(S, 1), torch.logsumexp(S, 0)
    G = (torch.exp(S - r[:, None]) + torch.exp(S - c[None, :]))
    out = {"r": r, "c": c, "S": S}
    Gd = G.clone()
    Gd.fill_diagonal_(0)
    out["anchor_out_I"] = Gd @ T64                       # what the grad pass accumulates (no diagonal)
    out["anchor_ws"] = (G * (I64 @ T64.t())).sum()
    eye = torch.eye(B, device=I.device, dtype=torch.float64)
    Gf = (G - 2 * eye) / (2 * B)
    out["dI"], out["dT"] = Gf @ T64 / tau, Gf.t() @ I64 / tau
    out["dtau"] = -(Gf * S).sum() / tau
    out["anchor"] = ((r - S.diag()).sum() + (c - S.diag()).sum()) / (2 * B)
    for name, X in (("I", I64), ("T", T64)):
        n = (X * X).sum(1)
        d2 = (n[:, None] + n[None, :] - 2 * X @ X.t()).clamp_min(0)
        W = torch.exp(-t * d2)
        W.fill_diagonal_(0)
        out["W_" + name] = W
        out["U_" + name] = W @ X
        out["rs_" + name] = W.sum(1)
        ssum = W.sum() / 2
        out["lunif_" + name] = torch.log(ssum / (B * (B - 1) / 2))
        out["dX_" + name] = (-2 * t / ssum) * (W.sum(1)[:, None] * X - W @ X)
    return out


def case_passes(path_name, flags, shapes, dtype_name):
    import torch
    import sparsify_clip_b200 as scb
    from sparsify_clip_b200 import backend_cuda as bc
    be = scb.get_backend()
    be.lib.scb_set_tc_flags(flags)
    dtype = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[dtype_name]
    bc.force_path(bc.PATH_TC if path_name == "tc" else bc.PATH_SIMT)
    which = os.environ.get("DIAG_WHICH", "lse,sum,grad").split(",")
    for (B, D, tau) in shapes:
        I, T = make(B, D, dtype, seed=B + D, kind="cluster" if B % 2 else "corr")
        ref = dense_refs(I, T, tau)
        print(f"[{path_name} flags={flags} {dtype_name}] B={B} D={D} tau={tau}")
        if "lse" in which:
            r = be.lse(I, T, 1.0 / tau)
            c = be.lse(T, I, 1.0 / tau)
            torch.cuda.synchronize()
            print(f"  lse rows  max abs err {((r.double() - ref['r']).abs().max().item()):.3e}   "
                  f"cols {((c.double() - ref['c']).abs().max().item()):.3e}")
        if "sum" in which:
            core = be.lunif_core(I, I, 2.0, 0, False)
            torch.cuda.synchronize()
            print(f"  lunif_sum rel err {abs(core['rs_sum'].item() - ref['rs_I'].sum().item()) / ref['rs_I'].sum().item():.3e}")
            sp = be.sparsify_sum(I, I, 0).item()
            E = I.double() @ I.double().t() - (2 * torch.eye(B, device='cuda', dtype=torch.float64) - 1)
            print(f"  sparsify_sum rel err {abs(sp - (E * E).sum().item()) / (E * E).sum().item():.3e}")
        if "grad" in which:
            core = be.lunif_core(I, I, 2.0, 0, True)
            torch.cuda.synchronize()
            U = core["U"].sum(0)
            rq = core["rq"].sum(0)
            print(f"  lunif U rel err {rel(U, ref['U_I']):.3e}   rq rel err {rel(rq, ref['rs_I']):.3e}   "
                  f"rs_sum rel err {abs(core['rs_sum'].item() - ref['rs_I'].sum().item()) / ref['rs_I'].sum().item():.3e}")
            if rel(U, ref["U_I"]) > 2e-2:
                errmap(U, ref["U_I"], "lunif U")
            rl, cl = ref["r"].float().contiguous(), ref["c"].float().contiguous()
            diag = be.row_dot(I, T)
            g1 = torch.ones((), device="cuda")
            dI, ws = be.anchor_grad(I, T, T, 1.0 / tau, rl, cl, cl, diag, 0, (1.0 / tau) / (2 * B), g1, True)
            dT, _ = be.anchor_grad(T, I, I, 1.0 / tau, cl, rl, rl, diag, 0, (1.0 / tau) / (2 * B), g1, False)
            torch.cuda.synchronize()
            print(f"  anchor dI rel err {rel(dI, ref['dI']):.3e}   dT rel err {rel(dT, ref['dT']):.3e}   "
                  f"ws rel err {abs(ws.item() - ref['anchor_ws'].item()) / abs(ref['anchor_ws'].item()):.3e}")
            if rel(dI, ref["dI"]) > 2e-2:
                errmap(dI, ref["dI"], "anchor dI")


def case_e2e(path_name, dtype_name):
    """public API (autograd) vs dense refs, incl. learnable tau on CPU and grad_output != 1"""
    import torch
    import sparsify_clip_b200 as scb
    from sparsify_clip_b200 import backend_cuda as bc
    dtype = {"bf16": torch.bfloat16, "f32": torch.float32}[dtype_name]
    bc.force_path(None)
    for (B, D, tau) in [(128, 512, 0.1), (384, 512, 0.07), (1000, 768, 0.1), (130, 72, 0.5)]:
        if dtype_name == "bf16" and D % 8:
            continue
        I, T = make(B, D, dtype, seed=B)
        ref = dense_refs(I, T, tau)
        I.requires_grad_(True)
        T.requires_grad_(True)
        tp = torch.nn.Parameter(torch.tensor(tau))
        a = scb.contrastive_loss(I, T, tp)
        ui, ut = scb.lunif_loss(I), scb.lunif_loss(T)
        al = scb.lalign_loss(I, T)
        loss = a + (ui + ut) / 2 + al
        (loss * 3.0).backward()
        torch.cuda.synchronize()
        dI_ref = ref["dI"] + 0.5 * ref["dX_I"] + 2 * (I.double() - T.double()) / B
        print(f"[e2e {path_name} {dtype_name}] B={B} D={D}: anchor rel {abs(a.item() - ref['anchor'].item()) / abs(ref['anchor'].item()):.2e} "
              f"lunif_I rel {abs(ui.item() - ref['lunif_I'].item()) / abs(ref['lunif_I'].item()):.2e} "
              f"dI rel {rel(I.grad / 3.0, dI_ref):.2e} dtau rel {abs(tp.grad.item() / 3.0 - ref['dtau'].item()) / abs(ref['dtau'].item()):.2e}")


def run_case(name):
    import torch
    assert torch.cuda.is_available()
    small = [(128, 64, 0.1), (128, 512, 0.1), (257, 512, 0.07), (384, 192, 0.1), (1000, 768, 0.1)]
    if name == "simt":
        case_passes("simt", 0, [(130, 72, 0.5), (128, 512, 0.1), (257, 200, 0.07)], "f32")
    elif name == "tc_lse":
        os.environ["DIAG_WHICH"] = "lse"
        case_passes("tc", 0, small, "bf16")
        case_passes("tc", 0, [(128, 512, 0.1)], "f16")
    elif name == "tc_sums":
        os.environ["DIAG_WHICH"] = "sum"
        case_passes("tc", 0, small, "bf16")
    elif name == "tc_grad_smem":
        os.environ["DIAG_WHICH"] = "grad"
        case_passes("tc", 0, small, "bf16")
    elif name == "tc_grad_tmem":
        os.environ["DIAG_WHICH"] = "grad"
        case_passes("tc", 1, small, "bf16")
    elif name == "e2e_tc":
        case_e2e("tc", "bf16")
    elif name == "e2e_simt":
        case_e2e("simt", "f32")
    else:
        raise SystemExit(f"unknown case {name}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--timeout", type=int, default=120)
    args = ap.parse_args()
    if args.case:
        run_case(args.case)
        sys.exit(0)
    for c in CASES:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", c], capture_output=True, text=True,
                               timeout=args.timeout)
            out = (p.stdout + p.stderr).strip().splitlines()
            print(f"===== {c}: exit {p.returncode} in {time.time() - t0:.1f}s")
            print("\n".join(out[-60:]))
        except subprocess.TimeoutExpired as e:
            print(f"===== {c}: TIMEOUT after {args.timeout}s")
            print(((e.stdout or b"").decode() + (e.stderr or b"").decode())[-3000:])
        sys.stdout.flush()
