"""Print the SASS instructions with the most warp-stall samples from `ncu --page source --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ci["# Samples"]]), r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print("kernel:", rows[0][1][:100])
print("total samples", tot)
for n, r in sorted(data, key=lambda t: -t[0])[:top]:
    reasons = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{100*n/tot:5.1f}%  {r[ci['Source']][:70]:70s} {reasons}")
