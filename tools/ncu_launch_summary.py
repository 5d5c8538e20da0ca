"""Per-kernel shares of ONE step from an `ncu --metrics gpu__time_duration.sum --csv` launch list of tools/prof_step.py.
The window is one full period of the launch sequence (from just after the LSE sweep of step k-1 to the LSE sweep of
step k, inclusive), i.e. exactly one step's launches.   usage: ncu_launch_summary.py raw.csv > step.csv"""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    rd = csv.reader(l for l in f if l.startswith('"'))
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if r[ix["Metric Name"]] == "gpu__time_duration.sum":
            v = float(r[ix["Metric Value"]].replace(",", ""))
            unit = r[ix["Metric Unit"]]
            us = v / 1000.0 if unit in ("ns", "nsecond") else v * 1000.0 if unit in ("ms", "msecond") else v
            rows.append((r[ix["Kernel Name"]], us))
marks = [i for i, (k, _) in enumerate(rows) if "k_tc_pass<5" in k]
assert len(marks) >= 2, "need at least two steps in the capture"
win = rows[marks[-2] + 1:marks[-1] + 1]
agg = OrderedDict()
for k, us in win:
    a = agg.setdefault(k, [0.0, 0])
    a[0] += us
    a[1] += 1
tot = sum(a[0] for a in agg.values())
print(f"# launches per step {len(win)}, summed kernel time {tot:.0f} us (cold-cache, serialised: compare shares)")
print("us_total,count,share,kernel")
for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'{us:.1f},{n},{us / tot:.4f},"{k[:100]}"')
