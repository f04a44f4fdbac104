"""Per-kernel shares from an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`).

    python profiles/launch_shares.py profiles/r2_launches.csv > profiles/r2_launch_shares.md
"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    name = r[ki]
    short = re.sub(r"\(.*", "", name).split("::")[-1].replace("(int)", "").replace("(bool)", "")
    ns = float(r[vi].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r[ui], 1)
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += ns
step = {k: v for k, v in agg.items() if k.startswith("k_")}
total = sum(v[1] for v in step.values())
print("| kernel | launches | total us | us / launch | share of the step kernels |")
print("|---|---|---|---|---|")
for k, (n, ns) in sorted(step.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {ns / 1e3:.1f} | {ns / 1e3 / n:.1f} | {100 * ns / total:.1f} % |")
for k, (n, ns) in agg.items():
    if k.startswith("k0_"):
        print(f"| `{k}` | {n} | {ns / 1e3:.1f} | {ns / 1e3 / n:.1f} | (graph build) |")
other = [(k, v) for k, v in agg.items() if not k.startswith("k_") and not k.startswith("k0_")]
print(f"| (torch / other kernels) | {sum(v[0] for _, v in other)} | {sum(v[1] for _, v in other) / 1e3:.1f} | | |")
