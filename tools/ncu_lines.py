"""Summarise `ncu --page source --csv --print-source sass,cuda`: stall samples per SOURCE line (first kernel only)."""
import csv
import sys

path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(open(path)))
fpath, hdr, out, kidx, fn_seen = None, None, [], -1, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1]
    elif r[0] == "Function Name":
        if r[1] not in fn_seen:
            fn_seen.append(r[1])
        kidx = fn_seen.index(r[1])
    elif r[0] == "Line No":
        hdr = r
    elif r[0].isdigit() and hdr and kidx == kern:
        out.append((fpath, r))
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
out = [(f, r) for f, r in out if r[ix["# Samples"]].isdigit()]
tot = sum(int(r[ix["# Samples"]]) for _, r in out)
print("kernel:", fn_seen[kern][:90], " total samples", tot)
for f, r in sorted(out, key=lambda fr: -int(fr[1][ix["# Samples"]]))[:n]:
    s = {h[6:]: int(r[ix[h]]) for h in stalls if int(r[ix[h]]) > 0}
    s = dict(sorted(s.items(), key=lambda kv: -kv[1])[:3])
    print(r[ix["# Samples"]].rjust(7), f.split("/")[-1].ljust(12), r[0].rjust(4), r[1].strip()[:70].ljust(70), s)
