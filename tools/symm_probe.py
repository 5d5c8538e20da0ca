"""Probe: all-gather over torch symmetric memory (peer-mapped buffers, copy-engine pushes, signal-pad barriers) against
NCCL's all_gather_into_tensor.  torchrun, one rank per GPU."""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, D = 32768, 512
n = B // world
x = torch.randn(n, D, device=dev).to(torch.bfloat16)
ref = torch.empty(B, D, dtype=torch.bfloat16, device=dev)
dist.all_gather_into_tensor(ref, x)
torch.cuda.synchronize()
try:
    t = symm_mem.empty((B, D), dtype=torch.bfloat16, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    bufs = [hdl.get_buffer(p, (B, D), torch.bfloat16) for p in range(world)]
    if rank == 0:
        print("symmetric memory ok: backend", hdl.get_backend() if hasattr(hdl, "get_backend") else "?", "multicast", hdl.has_multicast_support, flush=True)
except Exception as exc:
    print(f"rank {rank}: symmetric memory unavailable: {type(exc).__name__}: {exc}", flush=True)
    dist.barrier()
    os._exit(0)


def ag(src):
    hdl.barrier(channel=0)                     # every peer is done reading the previous contents
    for k in range(world):
        p = (rank + k) % world
        bufs[p][rank * n:(rank + 1) * n].copy_(src, non_blocking=True)
    hdl.barrier(channel=1)                     # every shard has landed everywhere
    return t


out = ag(x)
torch.cuda.synchronize()
print(f"rank {rank}: symm all-gather equal to NCCL: {torch.equal(out, ref)}", flush=True)
for name, fn in (("nccl", lambda: dist.all_gather_into_tensor(ref, x)), ("symm", lambda: ag(x))):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"{name}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per all-gather of {B}x{D} bf16 over {world} ranks", flush=True)
dist.barrier()
os._exit(0)
