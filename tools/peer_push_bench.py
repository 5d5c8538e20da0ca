"""Time one all-gather of a row shard through peer.py with the copy-engine pushes and with the SM push kernel.
torchrun --nproc-per-node N tools/peer_push_bench.py [shard_MiB ...]   (ranks on one node; prints on rank 0)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from sparsify_clip_b200 import peer

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
sizes = [float(a) for a in sys.argv[1:]] or [4.0, 12.0]
for mib in sizes:
    n = int(mib * (1 << 20)) // 2
    x = torch.full((n,), float(rank + 1), dtype=torch.bfloat16, device=dev)
    res = {}
    for mode, roles in (("copy engines", ()), ("sm kernel", ("B",))):
        peer.set_sm_push_roles(roles)
        ts = []
        for it in range(30):
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(1_500_000)      # ~1 ms: every launch below is queued before the first one runs (GPU-side time only)
            a.record()
            out, h = peer.all_gather_async(x, dist.group.WORLD, "B")
            h.wait()
            b.record()
            ok = bool((out.view(world, -1)[:, ::4097].float() == torch.arange(1, world + 1, device=dev)[:, None]).all())
            peer.release_all()
            torch.cuda.synchronize()
            assert ok, f"rank {rank}: gathered data wrong ({mode})"
            if it >= 5:
                ts.append(a.elapsed_time(b) * 1e3)
        t = torch.tensor(sorted(ts)[len(ts) // 2], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = t.item()
    if rank == 0:
        out_gb = (world - 1) * mib * (1 << 20) / 1e9
        print(f"world {world}, shard {mib} MiB: " + ", ".join(f"{m}: {v:.1f} us ({out_gb / (v * 1e-6):.0f} GB/s out per rank)"
                                                                 for m, v in res.items()), flush=True)
dist.destroy_process_group()
