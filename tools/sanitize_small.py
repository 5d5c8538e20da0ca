"""Small shapes through every tensor-core kernel (pair gradient sweeps, fused LSE sweep, single-CTA modes, combine) for
`compute-sanitizer --tool memcheck` runs: row/column/D tails, odd tile counts, every K-chunk count of the pair kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sparsify_clip_b200 as scb

for (B, D, tau) in [(385, 384, 0.07), (129, 264, 0.1), (640, 320, 0.1), (300, 512, 0.01), (257, 768, 0.1), (200, 128, 0.1)]:
    g = torch.Generator(device="cuda").manual_seed(B)
    I = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    T = torch.nn.functional.normalize(I + 0.5 * torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    for w in (dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0),
              dict(anchor=1.0, align=1.0, unif_img=0.0, unif_txt=0.0, unif_cen=1.0)):
        Ig = I.to(torch.bfloat16).requires_grad_(True)
        Tg = T.to(torch.bfloat16).requires_grad_(True)
        tp = torch.nn.Parameter(torch.tensor(tau))
        loss = scb.weighted_loss(Ig, Tg, tp, w)
        loss.backward()
        torch.cuda.synchronize()
        print(B, D, tau, "cen" if w["unif_cen"] else "exp3", f"loss {loss.item():.5f} |dI| {Ig.grad.float().norm().item():.4e} dtau {tp.grad.item():.4e}", flush=True)
print("SANITIZE_SMALL DONE")
