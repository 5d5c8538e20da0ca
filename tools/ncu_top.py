"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples."""
import csv
import sys

path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break          # only the first kernel of the dump
    if len(r) > 10 and r[0].startswith("0x"):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
print("stall totals:", dict(sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for pos, r in enumerate(data):
    r.append(pos)
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:n]:
    s = {h[6:]: int(r[ix[h]]) for h in stalls if int(r[ix[h]]) > 0}
    s = dict(sorted(s.items(), key=lambda kv: -kv[1])[:3])
    print(r[ix["# Samples"]].rjust(7), r[ix["Instructions Executed"]].rjust(10), f"#{r[-1]:5d}",
          r[ix["Source"]].strip()[:64].ljust(64), s)
