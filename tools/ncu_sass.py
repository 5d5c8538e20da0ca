"""Hot SASS of the first kernel in an `ncu --page source --csv` dump: sync/wait instructions and anything above a
sample threshold, with stall reasons.  usage: ncu_sass.py dump.csv [min_samples]"""
import csv
import sys

path = sys.argv[1]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 300
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
data = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) > 10 and r[0].startswith("0x"):
        data.append(r)
S = [int(r[ix["# Samples"]]) for r in data]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", sum(S), "instructions", len(S))
marks = [i for i, r in enumerate(data) if "USETMAXREG" in r[ix["Source"]]]
print("setmaxreg positions", marks)
b = [0] + marks + [len(S)]
for a, c in zip(b[:-1], b[1:]):
    print(f"  region #{a}-#{c}: {sum(S[a:c])} samples")
for pos, r in enumerate(data):
    src = r[ix["Source"]].strip()
    if S[pos] >= thr:
        s = {h[6:]: int(r[ix[h]]) for h in stalls if int(r[ix[h]]) > 0}
        s = dict(sorted(s.items(), key=lambda kv: -kv[1])[:2])
        print(f"#{pos:5d} {S[pos]:7d} {r[ix['Instructions Executed']]:>9} {src[:80]:80s} {s}")
