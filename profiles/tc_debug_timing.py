"""Experiment: per-role cycle breakdown of the tcgen05 projection kernel (library built with
GCA_EXTRA_NVCC_FLAGS=-DGCA_TC_DEBUG).  Not part of the product."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gconv_adapter_b200 import GraphStructure, _cabi
from gconv_adapter_b200.graphs.synthetic import make_graph

lib = _cabi.load()
raw = ctypes.CDLL(_cabi.LIB_PATH)
ei, n = make_graph("arxiv", seed=0)
g = GraphStructure(ei.cuda(), n, True)
d, r = 256, 16
x = torch.randn(n, d, device="cuda")
wd = torch.randn(r, d, device="cuda") * 0.1
out = torch.empty(n, r, device="cuda")
st = torch.cuda.current_stream().cuda_stream
buf = (ctypes.c_ulonglong * 16)()
for _ in range(3):
    lib.gca_fwd_project(g.handle, x.data_ptr(), d, wd.data_ptr(), out.data_ptr(), d, r, st)
raw.gca_debug_counters(buf, 1)
reps = 10
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(reps):
    lib.gca_fwd_project(g.handle, x.data_ptr(), d, wd.data_ptr(), out.data_ptr(), d, r, st)
t1.record()
torch.cuda.synchronize()
raw.gca_debug_counters(buf, 1)
ctas = 148
names = ["prod wait-empty", "prod split+STS+fence", "prod LDG issue", "mma wait-full", "mma issue+commit", "-", "epi wait-tfull", "epi total", "mma total"]
print("kernel us:", t0.elapsed_time(t1) * 1e3 / reps)
for i, nm in enumerate(names):
    print(f"{nm:24s} {buf[i] / reps / ctas:12.0f} cycles per CTA per launch")
