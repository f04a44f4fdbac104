"""Measurement of the shared-CSR backbone propagation (gca_propagate).
Prints one JSON line per workload: device time (CUDA events, inputs resident), achieved GB/s on the algorithmic
bytes (min: X read once + out written once + CSR; gather bound: every neighbour row fetched per edge), and the CPU
oracle (oracle.backbone_ref.gcn_conv, the reference's op sequence) timed on the host cores for the same input."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gconv_adapter_b200 import _cabi
from gconv_adapter_b200.graphs.csr import GLOBAL_GRAPH_CACHE
from gconv_adapter_b200.graphs.synthetic import make_graph
from gconv_adapter_b200.layers import propagate
from oracle import backbone_ref

PEAK = 6541.1
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

for name, heads, hd in (("arxiv", 1, 256), ("pubmed", 1, 64)):
    ei, n = make_graph(name, seed=0)
    if not bool((ei[0] == ei[1]).any()):                 # pubmed-shaped already carries its self loops
        ei = torch.cat([ei, torch.arange(n, dtype=torch.int64).repeat(2, 1)], dim=1)
    D = heads * hd
    x = torch.randn(n, heads, hd)
    xd, eid = x.cuda(), ei.cuda()
    for _ in range(3):
        y = propagate.gcn_conv(xd, eid)
    torch.cuda.synchronize()
    big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(10):
        big.zero_()                                      # flush the 126 MB L2 between iterations
        t0.record()
        y = propagate.gcn_conv(xd, eid)
        t1.record()
        torch.cuda.synchronize()
        times.append(t0.elapsed_time(t1))
    ms = sorted(times)[len(times) // 2]
    e_tot = ei.shape[1]
    a_min = 2 * n * D * 4 + 4 * e_tot + 8 * n + 4 * n
    a_gather = (e_tot + n) * D * 4 + 4 * e_tot + 8 * n + 4 * n
    torch.set_num_threads(os.cpu_count())
    backbone_ref.gcn_conv(x, ei)
    c0 = time.perf_counter()
    yr = backbone_ref.gcn_conv(x, ei)
    cpu_ms = (time.perf_counter() - c0) * 1e3
    err = (y.cpu() - yr).abs().max().item() / yr.abs().max().item()
    print(json.dumps({"op": "gcn_conv (gca_propagate)", "workload": f"{name}-shaped + self loops, D={D}", "ms": round(ms, 4),
                      "edges_per_sec": e_tot / ms * 1e3, "alg_min_MB": round(a_min / 1e6, 1), "GBs_on_min": round(a_min / ms / 1e6, 1),
                      "frac_of_measured_peak_on_min": round(a_min / ms / 1e6 / PEAK, 4), "gather_bound_MB": round(a_gather / 1e6, 1),
                      "GBs_on_gather_bound": round(a_gather / ms / 1e6, 1), "cpu_oracle_ms": round(cpu_ms, 1), "cpu_threads": os.cpu_count(),
                      "max_err_over_max_ref": err, "l2": "flushed between iterations"}))
