"""Step time of the BASELINE.json configurations that fit one GPU (eager launches and CUDA-graph replay)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sparsify_clip_b200 as scb

CASES = {
    "c2 exp4 anchor+lalign+lunif(centroids) B=4096 D=512": (4096, 512, dict(anchor=1.0, align=1.0, unif_img=0.0, unif_txt=0.0, unif_cen=1.0)),
    "exp3 anchor+lalign+lunif(I)+lunif(T)   B=4096 D=512": (4096, 512, dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0)),
    "c3 exp3                                B=32768 D=512": (32768, 512, dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0)),
    "exp4 (centroids)                       B=32768 D=512": (32768, 512, dict(anchor=1.0, align=1.0, unif_img=0.0, unif_txt=0.0, unif_cen=1.0)),
    "c4-like exp10 shard                    B=8192 D=768": (8192, 768, dict(anchor=1.0, align=1.2, unif_img=0.0, unif_txt=0.0, unif_cen=0.2)),
}
for name, (B, D, w) in CASES.items():
    g = torch.Generator(device="cuda").manual_seed(42)
    I = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    T = torch.nn.functional.normalize(I + 0.5 * torch.randn(B, D, generator=g, device="cuda"), dim=-1)
    I = I.to(torch.bfloat16).requires_grad_(True)
    T = T.to(torch.bfloat16).requires_grad_(True)

    def step():
        I.grad = T.grad = None
        loss = scb.weighted_loss(I, T, 0.1, w)
        loss.backward()
        return loss

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / n
    graphed = float("nan")
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        gr = torch.cuda.CUDAGraph()
        I.grad = T.grad = None
        with torch.cuda.graph(gr):
            l = scb.weighted_loss(I, T, 0.1, w)
            l.backward()
        gr.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        graphed = e0.elapsed_time(e1) / n
    except Exception as exc:
        print("  graph capture failed:", str(exc)[:100])
    flops = (6 + 4 * (1 if w["unif_cen"] else 0) + 4 * (1 if w["unif_img"] else 0) + 4 * (1 if w["unif_txt"] else 0)) * B * B * D
    print(f"{name}: eager {eager:.3f} ms ({B / eager * 1e3 / 1e6:.2f} M pairs/s), graph {graphed:.3f} ms; "
          f"algorithmic {flops / 1e12:.3f} TFLOP -> {flops / (min(eager, graphed if graphed == graphed else eager) * 1e-3) / 1e12:.0f} TFLOP/s")
