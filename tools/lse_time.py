"""Time the one-sweep row+column log-sum-exp (k_tc_pass M_LSE2) at given shapes: SCB_SWEEP_SHAPES=rowsxcolsxD,..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sparsify_clip_b200 as scb

be = scb.get_backend()
shapes = [tuple(int(v) for v in t.split("x")) for t in os.environ.get("SCB_SWEEP_SHAPES", "32768x32768x512").split(",")]
for (nA, nB, D) in shapes:
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.nn.functional.normalize(torch.randn(nB, D, generator=g, device="cuda"), dim=-1).to(torch.bfloat16)
    Y = torch.nn.functional.normalize(X.float() + 0.5 * torch.randn(nB, D, generator=g, device="cuda"), dim=-1).to(torch.bfloat16)
    fn = (lambda: be.lse_rows_cols(X, Y, 10.0)) if nA == nB else (lambda: be.lse(X[:nA], Y, 10.0))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = sorted(ts)[len(ts) // 2]
    print(f"[SCB_PASS_ASTAT={os.environ.get('SCB_PASS_ASTAT', 'default')}] {'lse2' if nA == nB else 'lse'} nA={nA} nB={nB} D={D}: "
          f"median {t:.3f} ms min {min(ts):.3f} ({2.0 * nA * nB * D / t / 1e9:.0f} TF/s incl. the combine kernels)", flush=True)
