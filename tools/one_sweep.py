"""A handful of launches of one gradient sweep at the c3 shape (for ncu: --launch-skip 3 --launch-count 1 -k regex:k_tc_)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sparsify_clip_b200 as scb

be = scb.get_backend()
be.lib.scb_set_tc_flags(int(sys.argv[1]) if len(sys.argv) > 1 else 7)
which = sys.argv[2] if len(sys.argv) > 2 else "lunif"
nA = nB = 32768
D = 512
if os.environ.get("SCB_SWEEP_SHAPE"):          # rows x columns x D, e.g. 8192x65536x768 (c4's 8-GPU shard)
    nA, nB, D = (int(v) for v in os.environ["SCB_SWEEP_SHAPE"].split("x"))
g = torch.Generator(device="cuda").manual_seed(1)
X = torch.nn.functional.normalize(torch.randn(nB, D, generator=g, device="cuda"), dim=-1).to(torch.bfloat16)
Y = torch.nn.functional.normalize(X.float() + 0.5 * torch.randn(nB, D, generator=g, device="cuda"), dim=-1).to(torch.bfloat16)
r = torch.full((nA,), 8.0, device="cuda")
c = torch.full((nB,), 8.0, device="cuda")
for _ in range(5):
    if which == "lunif":
        be.lunif_core(X[:nA], X, 2.0, 0, True)
    elif which == "anchor":
        be.anchor_grad_pass(X[:nA], Y, 10.0, r, c, 0, True)
    else:
        be.lse_rows_cols(X[:nA], Y, 10.0)
torch.cuda.synchronize()
print("done")
