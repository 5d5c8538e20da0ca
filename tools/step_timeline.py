"""GPU timeline of one c3 (or, PROF_CONFIG=c4 PROF_B=65536 PROF_D=768, c4) step via torch.profiler (CUPTI): kernel start/duration and idle gaps between kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import sparsify_clip_b200 as scb

B, D = int(os.environ.get("PROF_B", "32768")), int(os.environ.get("PROF_D", "512"))
g = torch.Generator(device="cuda").manual_seed(42)
I = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device="cuda"), dim=-1)
T = torch.nn.functional.normalize(I + 0.5 * torch.randn(B, D, generator=g, device="cuda"), dim=-1)
I = I.to(torch.bfloat16).requires_grad_(True)
T = T.to(torch.bfloat16).requires_grad_(True)
w = dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0)
tau = 0.1
if os.environ.get("PROF_CONFIG") == "c4":        # exp 10 mid-schedule, learnable CPU temperature (PROF_B=65536 PROF_D=768)
    w = dict(anchor=1.0, align=1.2, unif_img=0.0, unif_txt=0.0, unif_cen=0.2)
    tau = torch.nn.Parameter(torch.tensor(0.1))


def step():
    I.grad = T.grad = None
    loss = scb.weighted_loss(I, T, tau, w)
    loss.backward()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
prev_end = t0
busy = 0.0
print(f"{'start_us':>10} {'dur_us':>9} {'gap_us':>8}  name")
for e in ev:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = e.time_range.start - prev_end
    busy += d
    if d > 15 or gap > 15:
        print(f"{s:10.1f} {d:9.1f} {gap:8.1f}  {e.name[:80]}")
    prev_end = max(prev_end, e.time_range.end)
print(f"span {prev_end - t0:.1f} us, busy {busy:.1f} us, idle {prev_end - t0 - busy:.1f} us, kernels {len(ev)}")
