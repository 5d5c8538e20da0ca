"""Time one L_unif gradient sweep of a row shard (nA rows against 32768 columns) on one GPU: isolates the sweep's
efficiency at the shard sizes of 1/2/4/8-GPU runs from collective and multi-GPU effects."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sparsify_clip_b200 as scb

be = scb.get_backend()
B, D = 32768, 512
g = torch.Generator(device="cuda").manual_seed(42)
X = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device="cuda"), dim=-1).to(torch.bfloat16)
for nA in (32768, 16384, 8192, 4096):
    Xr = X[:nA]
    for _ in range(3):
        be.lunif_core(Xr, X, 2.0, 0, True)
    torch.cuda.synchronize()
    be.pass_events = []
    for _ in range(5):
        be.lunif_core(Xr, X, 2.0, 0, True)
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for _, a, b in be.pass_events]
    be.pass_events = None
    t = sorted(ts)[len(ts) // 2]
    print(f"nA={nA:6d}: {t:.3f} ms per sweep  ({t * 32768 / nA:.3f} ms scaled to the full batch)  plan={scb.pair_span_plan((nA + 127) // 128, 256)}")
