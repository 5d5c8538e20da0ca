"""Record and print the per-role timeline of cluster 0 of one CTA-pair lunif pass (debug aid for csrc/tc_pair.cu).
usage (on the GPU box): SCB_TC_FLAGS=3 python tools/pair_trace.py [B] [D] > gpurun_out/pair_trace.log"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sparsify_clip_b200 as scb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
CAP = 4096
be = scb.get_backend()
be.lib.scb_set_tc_flags(int(os.environ.get("SCB_TC_FLAGS", "3")))
g = torch.Generator(device="cuda").manual_seed(42)
X = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device="cuda"), dim=-1).to(torch.bfloat16)
for _ in range(2):
    be.lunif_core(X, X, 2.0, 0, True)
torch.cuda.synchronize()
buf = torch.zeros(2 * 4 * CAP * 2, dtype=torch.int64, device="cuda")
be.lib.scb_debug_pair_trace(buf.data_ptr())
be.lunif_core(X, X, 2.0, 0, True)
torch.cuda.synchronize()
be.lib.scb_debug_pair_trace(None)
t = buf.cpu().view(2, 4, CAP, 2)
ROLE = ["tma", "mma", "epi", "snd"]
TAG = {120: "tma B1 chunk issue", 121: "tma V chunk issue", 10: "m1 begin", 11: "m1 S_EMPTY ok", 12: "m1 chunk FULL ok",
       20: "m2own begin", 21: "m2own G_FULL ok", 22: "m2own V pair ok", 30: "m2peer begin", 31: "m2peer W_FULL ok",
       32: "m2peer V pair ok", 23: "m2own group issued", 33: "m2peer group issued", 40: "epi wait S_FULL", 41: "epi S_FULL ok", 42: "epi G_FULL arrive",
       50: "snd G_FULL ok", 51: "snd W_EMPTY ok", 52: "snd stores issued"}
for c in range(2):
    ev = []
    for r in range(4):
        for i in range(CAP):
            w0, w1 = int(t[c, r, i, 0]), int(t[c, r, i, 1])
            if w0 == 0 and w1 == 0:
                break
            tag, tile = (w0 >> 32) & 0xffffffff, w0 & 0xffffffff
            if tag == 0:
                continue
            ev.append((w1, r, tag, tile))
    ev.sort()
    if not ev:
        continue
    t0 = ev[0][0]
    # steady-state window: skip the first 20000 cycles, print 30000 cycles
    print(f"===== CTA rank {c}: {len(ev)} events, span {ev[-1][0] - t0} cycles")
    lo = int(os.environ.get("TRACE_FROM", "30000"))
    hi = lo + int(os.environ.get("TRACE_SPAN", "26000"))
    for (clk, r, tag, tile) in ev:
        rel = clk - t0
        if lo <= rel <= hi:
            print(f"{rel:9d}  {ROLE[r]:4s} {TAG.get(tag, str(tag)):22s} {tile}")
    # summary: average interval between consecutive 'm1 begin'
    m1 = [clk for (clk, r, tag, tile) in ev if tag == 10]
    if len(m1) > 10:
        d = [b - a for a, b in zip(m1[5:-1], m1[6:])]
        print(f"mean cycles between own MMA1 starts: {sum(d) / len(d):.0f}  (ideal 4096)")
