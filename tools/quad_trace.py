"""Record the per-role timeline of cluster 0 of one cluster-of-4 L_unif sweep (debug aid for csrc/tc_quad.cu; needs a
library built with SCB_DEV=1).  Prints, per leader CTA, where the MMA issuer's time goes per step.
usage (GPU box): SCB_DEV=1 python sparsify_clip_b200/build.py --force; python tools/quad_trace.py [B [D [tc_flags]]] > gpurun_out/quad_trace.log
(512 < D <= 768 with tc_flags 31 traces the single-S-buffer variant: ideal 6144 cycles per step there)"""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sparsify_clip_b200 as scb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
CAP = 4096
be = scb.get_backend()
be.lib.scb_set_tc_flags(int(sys.argv[3]) if len(sys.argv) > 3 else 7)
g = torch.Generator(device="cuda").manual_seed(42)
X = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device="cuda"), dim=-1).to(torch.bfloat16)
for _ in range(2):
    be.lunif_core(X, X, 2.0, 0, True)
torch.cuda.synchronize()
buf = torch.zeros(4 * 4 * CAP * 2, dtype=torch.int64, device="cuda")
be.lib.scb_debug_pair_trace(buf.data_ptr())
be.lunif_core(X, X, 2.0, 0, True)
torch.cuda.synchronize()
be.lib.scb_debug_pair_trace(None)
t = buf.cpu().view(4, 4, CAP, 2)
ROLE = ["tma", "mma", "epi", "snd"]
TAG = {118: "tma pad slot", 119: "tma A chunk", 120: "tma Bhalf slot", 121: "tma V own", 122: "tma V peer", 10: "m1 begin", 11: "m1 S_EMPTY ok",
       12: "m1 kpair FULL ok", 20: "m2own begin", 21: "m2own G_MMA ok", 22: "m2own V pair ok", 23: "m2own issued", 30: "m2peer begin",
       31: "m2peer W ok", 32: "m2peer V pair ok", 33: "m2peer issued", 40: "epi wait S_FULL", 41: "epi S_FULL ok", 42: "epi G arrive",
       50: "snd G_FULL ok", 51: "snd W_EMPTY ok", 52: "snd stores issued"}
allev = {}
for c in range(4):
    ev = []
    for r in range(4):
        for i in range(CAP):
            w0, w1 = int(t[c, r, i, 0]), int(t[c, r, i, 1])
            if w0 == 0 and w1 == 0:
                break
            ev.append((w1, r, (w0 >> 32) & 0xffffffff, w0 & 0xffffffff))
    ev.sort()
    allev[c] = ev
for c in (0, 2):
    ev = allev[c]
    mma = [(clk, tag, tile) for (clk, r, tag, tile) in ev if r == 1]
    if len(mma) < 100:
        print("rank", c, "no trace (library built without SCB_DEV=1?)")
        continue
    m1 = [clk for (clk, tag, tile) in mma if tag == 10]
    d = [b - a for a, b in zip(m1[8:-2], m1[9:-1])]
    print(f"===== leader rank {c}: {len(mma)} MMA-issuer events; mean cycles between own MMA1 starts {sum(d) / len(d):.0f} (ideal 4096)")
    # attribute every interval of the issuer's timeline to the event that ENDS it (steady state: skip the first 8 steps)
    acc, cnt = defaultdict(float), defaultdict(int)
    start = m1[8]
    end = m1[-2]
    prev = None
    for (clk, tag, tile) in mma:
        if clk < start or clk > end:
            prev = (clk, tag)
            continue
        if prev is not None:
            acc[(prev[1], tag)] += clk - prev[0]
            cnt[(prev[1], tag)] += 1
        prev = (clk, tag)
    steps = len([x for x in m1 if start <= x < end])
    for (a, b), v in sorted(acc.items(), key=lambda kv: -kv[1]):
        print(f"   {TAG.get(a, a):18s} -> {TAG.get(b, b):18s}: {v / steps:8.0f} clk/step  ({cnt[(a, b)] / steps:.1f} x {v / cnt[(a, b)]:.0f})")
lo = int(os.environ.get("TRACE_FROM", "60000"))
hi = lo + int(os.environ.get("TRACE_SPAN", "14000"))
for c in range(4):
    ev = allev[c]
    if not ev:
        continue
    t0 = allev[0][0][0]
    print(f"===== CTA rank {c}: events in [{lo}, {hi}] cycles after the first event of rank 0")
    for (clk, r, tag, tile) in ev:
        rel = clk - t0
        if lo <= rel <= hi and tag not in (12,):
            print(f"{rel:9d}  r{c} {ROLE[r]:4s} {TAG.get(tag, str(tag)):20s} {tile}")
