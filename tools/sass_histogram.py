"""Opcode histogram of the tensor-core kernels in libscb200.so (cuobjdump -sass): the SASS mnemonics that prove
tcgen05 / TMEM / TMA / DSMEM use, per kernel instantiation.  usage: python tools/sass_histogram.py > profiles/<name>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sparsify_clip_b200", "libscb200.so")
KEEP = re.compile(r"^(UTCHMMA|UTCBAR|UTMALDG|UTCATOMSWS|LDTM|STTM|STAS|SYNCS|UCGABAR|MUFU\.EX2|UTMAPF|UTCCP)")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = {}
per, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur and KEEP.match(m.group(1)):
        op = m.group(1)
        op = re.sub(r"\.(SP|SHARED|CTA|MULTICAST_2X?)$", "", op)
        per[cur][op] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass sparsify_clip_b200/libscb200.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a): tensor-core / TMA /")
print("# TMEM / mbarrier opcodes per tensor-core kernel instantiation.  UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2),")
print("# UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor (TMA), LDTM/STTM = tcgen05.ld/st, STAS = st.async (DSMEM).")
print("# k_tc_quad<MODE, KCH, true> = the single-S-buffer variant (512 < D <= 768).")
tot = collections.Counter()
rows = []
for (mangled, cnt), name in zip(per.items(), names):
    m = re.search(r"(k_tc_\w+<[^>]*>)", name)
    if not m or not cnt:
        continue
    tot.update(cnt)
    rows.append((m.group(1), cnt))
fmt = lambda c: ", ".join(f"{k}={v}" for k, v in sorted(c.items()))
print("TOTAL over all k_tc_* kernels: " + fmt(tot) + "\n")
for name, cnt in sorted(rows):
    print(name + "\n    " + fmt(cnt))
