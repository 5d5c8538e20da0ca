"""TEST INFRASTRUCTURE ONLY -- mint tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The reference has no golden vectors of its own, so these are
outputs of the reference functions themselves (sparsify_clip.py:110-132, :159-164,
:186-187, :334-355 + F.normalize as at :803-805), evaluated with autograd in fp64
("truth" for the given inputs) and in fp32 (what the reference computes without
autocast).  Inputs are stored in the file, so the tests never regenerate them.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from oracle import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def make_inputs(kind, B, D, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "iid":
        I = F.normalize(torch.randn(B, D, generator=g), dim=-1)
        T = F.normalize(torch.randn(B, D, generator=g), dim=-1)
    elif kind == "corr":          # T correlated with I: diagonals dominate like a trained model
        I = F.normalize(torch.randn(B, D, generator=g), dim=-1)
        T = F.normalize(I + 0.5 * torch.randn(B, D, generator=g), dim=-1)
    elif kind == "cluster":       # 16-cluster mixture with exact duplicate rows
        k = min(16, max(1, B // 2))
        cen = F.normalize(torch.randn(k, D, generator=g), dim=-1)
        idx = torch.randint(0, k, (B,), generator=g)
        I = F.normalize(cen[idx] + 0.05 * torch.randn(B, D, generator=g), dim=-1)
        T = F.normalize(cen[idx] + 0.05 * torch.randn(B, D, generator=g), dim=-1)
        if B >= 4:                # exact duplicates
            I[1] = I[0]; T[3] = T[2]; I[B - 1] = I[B // 2]
    elif kind == "nonunit":       # rows not unit norm (pdist works on raw differences)
        I = torch.randn(B, D, generator=g) * (1.5 / D ** 0.5)
        T = I + (0.4 / D ** 0.5) * torch.randn(B, D, generator=g)
    else:
        raise ValueError(kind)
    return I, T


def bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float32)


def run_reference(ref, I32, T32, tau, dtype):
    """All terms + gradients through the reference functions and autograd."""
    out = {}
    I = I32.to(dtype).clone().requires_grad_(True)
    T = T32.to(dtype).clone().requires_grad_(True)
    tp = torch.nn.Parameter(torch.tensor(tau, dtype=dtype))

    a = ref.contrastive_loss(I, T, tp)
    gi, gt, gtau = torch.autograd.grad(a, (I, T, tp))
    out.update(anchor=a, anchor_dI=gi, anchor_dT=gt, anchor_dtau=gtau)

    al = ref.lalign_loss(I, T)
    gi, gt = torch.autograd.grad(al, (I, T))
    out.update(lalign=al, lalign_dI=gi, lalign_dT=gt)

    ui = ref.lunif_loss(I)
    (gi,) = torch.autograd.grad(ui, (I,))
    ut = ref.lunif_loss(T)
    (gt,) = torch.autograd.grad(ut, (T,))
    out.update(lunif_img=ui, lunif_img_dX=gi, lunif_txt=ut, lunif_txt_dX=gt)

    c = F.normalize(ref.compute_centroids_only(I, T), dim=-1)
    uc = ref.lunif_loss(c)
    gi, gt = torch.autograd.grad(uc, (I, T))
    out.update(lunif_cen=uc, lunif_cen_dI=gi, lunif_cen_dT=gt, centroids=c.detach())

    # exp-3 and exp-4 compositions as coded at :787-791 and :801-809
    e3 = ref.contrastive_loss(I, T, tp) + (ref.lunif_loss(I) + ref.lunif_loss(T)) / 2 + ref.lalign_loss(I, T)
    gi, gt, gtau = torch.autograd.grad(e3, (I, T, tp))
    out.update(exp3=e3, exp3_dI=gi, exp3_dT=gt, exp3_dtau=gtau)
    c = F.normalize(ref.compute_centroids_only(I, T), dim=-1)
    e4 = ref.contrastive_loss(I, T, tp) + ref.lalign_loss(I, T) + ref.lunif_loss(c)
    gi, gt, gtau = torch.autograd.grad(e4, (I, T, tp))
    out.update(exp4=e4, exp4_dI=gi, exp4_dT=gt, exp4_dtau=gtau)

    out["sparsify_img"] = ref.sparsify_loss(I)
    return {k: v.detach().double().numpy() for k, v in out.items()}


CASES = [
    # name, kind, B, D, seed, tau, bf16-rounded inputs, keep full grads
    ("b2_d8_s42", "iid", 2, 8, 42, 0.1, False, True),
    ("b3_d8_s0", "iid", 3, 8, 0, 0.07, False, True),
    ("b128_d512_iid_s0", "iid", 128, 512, 0, 0.1, False, True),
    ("b128_d512_corr_s42", "corr", 128, 512, 42, 0.1, False, False),
    ("b128_d512_cluster_s1", "cluster", 128, 512, 1, 0.1, False, False),
    ("b128_d512_corr_bf16_s2", "corr", 128, 512, 2, 0.07, True, False),
    ("b129_d64_corr_s42", "corr", 129, 64, 42, 0.1, True, True),
    ("b129_d64_cluster_s1", "cluster", 129, 64, 1, 1.0, False, True),
    ("b130_d128_nonunit_s3", "nonunit", 130, 128, 3, 0.5, True, True),
    ("b256_d512_cluster_tau001_s4", "cluster", 256, 512, 4, 0.01, True, False),
    ("b1000_d768_corr_bf16_s42", "corr", 1000, 768, 42, 0.1, True, False),
]


def sample_rows(B):
    # includes the duplicated rows planted by make_inputs("cluster")
    return sorted({0, 1, 2, 3, B // 2, B - 2, B - 1})


def main():
    ref, _ = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    for name, kind, B, D, seed, tau, rounded, full in CASES:
        I, T = make_inputs(kind, B, D, seed)
        if rounded:
            I, T = bf16_round(I), bf16_round(T)
        rows = sample_rows(B)
        r64 = run_reference(ref, I, T, tau, torch.float64)
        r32 = run_reference(ref, I, T, tau, torch.float32)
        blob = {"I": I.numpy(), "T": T.numpy(), "tau": np.float64(tau),
                "bf16_exact": np.bool_(rounded)}
        for k, v in r64.items():
            if v.ndim == 2 and not full:
                blob["f64_" + k + "_rows"] = v[rows].astype(np.float64)
                blob["f64_" + k + "_fro"] = np.float64(np.linalg.norm(v))
            elif v.ndim == 2:
                blob["f64_" + k] = v.astype(np.float32) if B > 16 else v
            else:
                blob["f64_" + k] = v
        for k, v in r32.items():
            if v.ndim == 0:
                blob["f32_" + k] = v
        if not full:
            blob["sample_rows"] = np.asarray(rows)
        if rounded:                       # exact and half the size: store the bf16 bit patterns
            blob["I"] = (I.view(torch.int32).numpy() >> 16).astype(np.uint16)
            blob["T"] = (T.view(torch.int32).numpy() >> 16).astype(np.uint16)
        blob["inputs_are_bf16_bits"] = np.bool_(rounded)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        print(f"{name}: anchor={r64['anchor']:.6f} lunif_img={r64['lunif_img']:.6f} "
              f"exp3={r64['exp3']:.6f} -> {os.path.getsize(path)/1e3:.0f} kB")


if __name__ == "__main__":
    sys.exit(main())
