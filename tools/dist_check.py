"""Sharded-vs-single-GPU parity on a multi-GPU box (launch with torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
Every rank holds a row shard; the loss must equal the unsharded loss and the gathered gradient slices the
unsharded gradient (computed on rank 0 with the same kernels), and both must match the fp64 oracle at small B."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import sparsify_clip_b200 as scb

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ok = True
for (B, D, tau, w) in [(1024, 512, 0.1, dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0)),
                       (2048, 768, 0.07, dict(anchor=1.0, align=1.3, unif_img=0.0, unif_txt=0.0, unif_cen=0.6)),
                       # tau = 0.02: the device-side norm bound arms the exact second LSE sweep on every rank
                       (1024, 512, 0.02, dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0)),
                       (2048, 512, 0.1, dict(anchor=1.0, align=1.0, unif_img=0.0, unif_txt=0.0, unif_cen=1.0)),
                       (8192, 512, 0.1, dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0))]:
    g = torch.Generator(device=dev).manual_seed(1234)          # same full batch on every rank
    I = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device=dev), dim=-1)
    T = torch.nn.functional.normalize(I + 0.5 * torch.randn(B, D, generator=g, device=dev), dim=-1)
    I, T = I.to(torch.bfloat16).float(), T.to(torch.bfloat16).float()
    n = B // world
    prev = scb.set_fp32_mode("bf16")
    Il = I[rank * n:(rank + 1) * n].clone().requires_grad_(True)
    Tl = T[rank * n:(rank + 1) * n].clone().requires_grad_(True)
    tp = torch.nn.Parameter(torch.tensor(tau))
    loss = scb.weighted_loss(Il, Tl, tp, w, group=True)
    loss.backward()
    gI = torch.empty(B, D, device=dev)
    gT = torch.empty(B, D, device=dev)
    dist.all_gather_into_tensor(gI, Il.grad.contiguous())
    dist.all_gather_into_tensor(gT, Tl.grad.contiguous())
    if rank == 0:
        If, Tf = I.clone().requires_grad_(True), T.clone().requires_grad_(True)
        tf_ = torch.nn.Parameter(torch.tensor(tau))
        full = scb.weighted_loss(If, Tf, tf_, w)
        full.backward()
        e_l = abs(loss.item() - full.item()) / abs(full.item())
        e_i = ((gI - If.grad).norm() / If.grad.norm()).item()
        e_t = ((gT - Tf.grad).norm() / Tf.grad.norm()).item()
        e_tau = abs(tp.grad.item() - tf_.grad.item()) / abs(tf_.grad.item())
        msg = f"B={B} D={D} world={world}: loss {loss.item():.6f} vs unsharded rel {e_l:.1e}; dI {e_i:.1e} dT {e_t:.1e} dtau {e_tau:.1e}"
        if B <= 2048:
            from oracle import closed_form as cf
            ref, dI, dT, dtau, _ = cf.weighted_loss(I.cpu().numpy(), T.cpu().numpy(), tau, w["anchor"], w["align"],
                                                    w["unif_img"], w["unif_txt"], w["unif_cen"])
            o_l = abs(loss.item() - ref) / abs(ref)
            o_i = np.linalg.norm(gI.double().cpu().numpy() - dI) / np.linalg.norm(dI)
            msg += f" | oracle: loss rel {o_l:.1e} dI rel {o_i:.1e}"
            ok &= o_l <= 1e-5 and o_i <= 1e-3
        ok &= e_l <= 2e-6 and e_i <= 2e-4 and e_t <= 2e-4 and e_tau <= 1e-4
        print(msg, flush=True)
    scb.set_fp32_mode(prev)
if rank == 0:
    print("DIST_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
