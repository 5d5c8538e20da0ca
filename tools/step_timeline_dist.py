"""GPU timeline of one sharded c3 step on rank 0 (torchrun): kernel start/duration and idle gaps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import sparsify_clip_b200 as scb

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, D = 32768, 512
n = B // world
g = torch.Generator(device="cuda").manual_seed(42 + rank)
I = torch.nn.functional.normalize(torch.randn(n, D, generator=g, device="cuda"), dim=-1)
T = torch.nn.functional.normalize(I + 0.5 * torch.randn(n, D, generator=g, device="cuda"), dim=-1)
I = I.to(torch.bfloat16).requires_grad_(True)
T = T.to(torch.bfloat16).requires_grad_(True)
w = dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0)


def step():
    I.grad = T.grad = None
    loss = scb.weighted_loss(I, T, 0.1, w, group=dist.group.WORLD)
    loss.backward()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
if rank == 0:
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
        if d > 8 or gap > 8:
            print(f"{s:10.1f} {d:9.1f} {gap:8.1f}  {e.name[:70]}")
        prev_end = max(prev_end, e.time_range.end)
    print(f"span {prev_end - t0:.1f} us, busy(sum) {busy:.1f} us, kernels {len(ev)}")
dist.destroy_process_group()
