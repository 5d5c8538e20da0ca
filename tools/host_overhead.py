"""Where the host time of one small step goes (B=4096: the GPU needs ~0.45 ms, eager launches ~0.9 ms): cProfile of the
Python side of 300 steps, plus the wall time per step with and without the profiler."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sparsify_clip_b200 as scb

B, D = int(os.environ.get("HO_B", 4096)), 512
g = torch.Generator(device="cuda").manual_seed(42)
I = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device="cuda"), dim=-1).to(torch.bfloat16).requires_grad_(True)
T = torch.nn.functional.normalize(I.detach().float() + 0.5 * torch.randn(B, D, generator=g, device="cuda"), dim=-1).to(torch.bfloat16).requires_grad_(True)
w = dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0)


def step():
    I.grad = T.grad = None
    scb.weighted_loss(I, T, 0.1, w).backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    step()
torch.cuda.synchronize()
print(f"wall per step {1e3 * (time.perf_counter() - t0) / 300:.3f} ms (eager, CPU-bound when above the GPU time)")
t0 = time.perf_counter()
for _ in range(300):
    step()
host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"host time per step (launch loop only, no final sync) {1e3 * host / 300:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
