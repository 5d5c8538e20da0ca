"""Time the B x B gradient sweeps (L_unif and anchor) of the loaded library at the c3 shape and the 8-GPU shard shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sparsify_clip_b200 as scb

be = scb.get_backend()
tag = os.environ.get("SCB_LIB_SUFFIX", "") or "default"
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 7
be.lib.scb_set_tc_flags(flags)
shapes = [(32768, 32768, 512), (4096, 32768, 512)]
if os.environ.get("SCB_SWEEP_SHAPES"):        # e.g. "8192x65536x768,4096x4096x1024" (rows x columns x D)
    shapes = [tuple(int(v) for v in t.split("x")) for t in os.environ["SCB_SWEEP_SHAPES"].split(",")]
for (nA, nB, D) in shapes:
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.nn.functional.normalize(torch.randn(nB, D, generator=g, device="cuda"), dim=-1).to(torch.bfloat16)
    Y = torch.nn.functional.normalize(X.float() + 0.5 * torch.randn(nB, D, generator=g, device="cuda"), dim=-1).to(torch.bfloat16)
    Xr = X[:nA]
    r = torch.full((nA,), 8.0, device="cuda")
    c = torch.full((nB,), 8.0, device="cuda")
    out = {}
    for name, fn in (("lunif", lambda: be.lunif_core(Xr, X, 2.0, 0, True)),
                     ("anchor", lambda: be.anchor_grad_pass(Xr, Y, 10.0, r, c, 0, True))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        n = 20
        be.pass_events = []
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for _, a, b in be.pass_events)
        be.pass_events = None
        out[name] = (ts[len(ts) // 2], ts[0])
    print(f"[{tag} flags={flags}] nA={nA} nB={nB} D={D}: lunif median {out['lunif'][0]:.3f} min {out['lunif'][1]:.3f} ms ({4.0 * nA * nB * D / out['lunif'][0] / 1e9:.0f} TF/s) | "
          f"anchor median {out['anchor'][0]:.3f} min {out['anchor'][1]:.3f} ms ({4.0 * nA * nB * D / out['anchor'][0] / 1e9:.0f} TF/s)", flush=True)
