"""Turn `ncu --page raw --csv` output of one adapter step into a per-kernel summary (markdown + traffic.json).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    python profiles/summarize_ncu.py raw.csv arxiv profiles/r2_step_ncu_full.md
The capture must hold exactly the kernels of ONE step (8 launches in round 2, 9 in round 1).
"""
import csv
import json
import os
import sys

raw, workload, out_md = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}


def val(r, key):
    try:
        return float(r[ci[key]].replace(",", ""))
    except (KeyError, ValueError):
        return float("nan")


def scale(r, key, want):
    """value converted to `want` units (byte / us)."""
    v, u = val(r, key), units[ci[key]] if key in ci else ""
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "second": 1e6}
    return v * f.get(u, 1)


ORDER_R1 = ["project_fwd", "hop_fwd", "hop_expand_fwd", "project_bwd", "wgrad_up", "hop_bwd", "hop_expand_bwd", "wgrad_down", "finalize"]
ORDER_R2 = ["project_fwd", "hop_fwd", "hop_expand_fwd", "bwd_up", "hop_bwd", "hop_plain_bwd", "expand_wgrad_bwd", "finalize"]
ORDER = ORDER_R2 if len(data) == len(ORDER_R2) else ORDER_R1
lines = ["| # | phase | kernel | time (us) | DRAM read (MB) | DRAM write (MB) | DRAM % peak | tensor pipe % | issue active % | warps active % | regs |",
         "|---|---|---|---|---|---|---|---|---|---|---|"]
traffic = {}
for i, r in enumerate(data):
    name = r[ci["Kernel Name"]]
    short = name.split("(")[0].split("::")[-1][:40]
    phase = ORDER[i] if i < len(ORDER) else "?"
    t = scale(r, "gpu__time_duration.sum", "us")
    rd, wr = scale(r, "dram__bytes_read.sum", "byte"), scale(r, "dram__bytes_write.sum", "byte")
    traffic[phase] = int(rd + wr)
    lines.append(f"| {i} | {phase} | `{short}` | {t:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
                 f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                 f"{val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                 f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                 f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {val(r, 'launch__registers_per_thread'):.0f} |")
open(out_md, "w").write("\n".join(lines) + "\n")
tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
allt = json.load(open(tpath)) if os.path.isfile(tpath) else {}
allt[workload] = traffic
json.dump(allt, open(tpath, "w"), indent=1)
print("\n".join(lines))
