"""Config c5 (BASELINE.json: experiment_6 full step, random-init ViT-B/32 + text tower, synthetic 224^2 images and
77-token captions, B = 4096 over 8 GPUs = 512 per GPU): time of one training step and the share of it spent in the
loss path, for the two losses experiment 6 runs (epoch 0: L_unif-only warm-up; afterwards: anchor + L_align +
L_unif(centroids)).

    python callers/bench_c5.py                      # one GPU, local batch 512 (one rank's share of c5)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 callers/bench_c5.py

Prints one JSON line per phase on rank 0.  On one GPU it also times the reference's own loss arithmetic in eager
PyTorch on the same embeddings (the stock path of sparsify_clip.py:110-132, :159-164, :186-187, :353, :804)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F

import sparsify_clip_b200 as scb
from callers.encoders import MiniCLIP
from callers.train_step import TrainStep

CFG = {"loss_type": "only_lunif_n_then_anchor+lalign+lunif(centroids)", "only_lunif_epochs": 1, "anchor_temperature": 0.1,
       "anchor_temperature_learnable": False, "learning_rate": 1e-4, "fp16": True}


def eager_reference_loss(i, t, tau, epoch):
    """The reference's op sequence for experiment 6, eager PyTorch (fp32 pdist, as under autocast)."""
    lunif = lambda x: torch.pdist(x.float()).pow(2).mul(-2).exp().mean().log()
    if epoch < 1:
        return 0.5 * (lunif(i) + lunif(t))
    logits = (i @ t.t()) / tau
    tgt = torch.arange(logits.shape[0], device=logits.device)
    anchor = 0.5 * (F.cross_entropy(logits, tgt) + F.cross_entropy(logits.t(), tgt))
    lalign = (i - t).float().norm(dim=1).pow(2).mean()
    c = F.normalize((i.float() + t.float()) / 2, dim=-1)
    return anchor + lalign + lunif(c)


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--local-batch", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--amp", choices=["bf16", "fp16"], default="bf16")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    torch.manual_seed(0)
    model = MiniCLIP().to(dev)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    n, B = a.local_batch, a.local_batch * world
    amp = torch.bfloat16 if a.amp == "bf16" else torch.float16
    ts = TrainStep(model, CFG, t_total=1000, group=group, amp_dtype=amp)
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    images = torch.randn(n, 3, 224, 224, generator=g, device=dev)
    tokens = torch.randint(0, 49408, (n, 77), generator=g, device=dev)
    for epoch, name in ((0, "epoch 0: (lunif(img) + lunif(txt)) / 2"), (1, "after warm-up: anchor + lalign + lunif(centroids)")):
        step_ms = timed(lambda: ts(images, tokens, epoch), a.steps, a.warmup)
        with torch.no_grad(), torch.autocast("cuda", dtype=amp):
            img, txt = ts.embed(images, tokens)
        img, txt = img.detach(), txt.detach()

        def loss_only():
            i, t = img.clone().requires_grad_(True), txt.clone().requires_grad_(True)
            ts.loss(i, t, epoch).backward()

        loss_ms = timed(loss_only, 4 * a.steps, a.warmup)
        tt = torch.tensor([step_ms, loss_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_ms, loss_ms = tt.tolist()
        line = {"config": f"c5: experiment_6 step, ViT-B/32 + text tower random init, synthetic data, global B={B} over {world} GPU(s), amp {a.amp}",
                "phase": name, "step_ms": step_ms, "samples_per_s": B / (step_ms * 1e-3), "loss_fwd_bwd_ms": loss_ms,
                "loss_share_of_step": loss_ms / step_ms}
        if world == 1:
            def ref_only():
                i, t = img.clone().requires_grad_(True), txt.clone().requires_grad_(True)
                eager_reference_loss(i, t, 0.1, epoch).backward()
            line["eager_reference_loss_fwd_bwd_ms"] = timed(ref_only, 4 * a.steps, a.warmup)
        if rank == 0:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
