"""Two c3 steps (exp-3 composition, B=32768, D=512, bf16) for ncu captures: 6 k_tc_pass launches per step
in the order lse, lse, lunif, lunif, anchor_grad, anchor_grad."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sparsify_clip_b200 as scb

B = int(os.environ.get("PROF_B", "32768"))
D = int(os.environ.get("PROF_D", "512"))
steps = int(os.environ.get("PROF_STEPS", "2"))
if "PROF_TC_FLAGS" in os.environ:
    scb.get_backend().lib.scb_set_tc_flags(int(os.environ["PROF_TC_FLAGS"]))
g = torch.Generator(device="cuda").manual_seed(42)
I = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device="cuda"), dim=-1)
T = torch.nn.functional.normalize(I + 0.5 * torch.randn(B, D, generator=g, device="cuda"), dim=-1)
I = I.to(torch.bfloat16).requires_grad_(True)
T = T.to(torch.bfloat16).requires_grad_(True)
w = dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0)
for _ in range(steps):
    I.grad = T.grad = None
    loss = scb.weighted_loss(I, T, 0.1, w)
    loss.backward()
torch.cuda.synchronize()
print("loss", loss.item())
