"""Experiment: per-role cycle breakdown of the warp-specialised hop+expand kernel (library built with
GCA_EXTRA_NVCC_FLAGS=-DGCA_WS_DEBUG, selected through GCA_LIB_PATH).  Not part of the product."""
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
z = torch.randn(n, r, device="cuda")
wu = torch.randn(d, r, device="cuda") * 0.1
bu = torch.randn(d, device="cuda") * 0.1
sc = torch.full((1,), 1.3, device="cuda")
h2 = torch.empty(n, r, device="cuda")
y = torch.empty(n, d, device="cuda")
st = torch.cuda.current_stream().cuda_stream
buf = (ctypes.c_ulonglong * 16)()

def call():
    rc = lib.gca_fwd_hop2_up(g.handle, z.data_ptr(), x.data_ptr(), d, wu.data_ptr(), bu.data_ptr(), sc.data_ptr(), 1,
                             h2.data_ptr(), y.data_ptr(), d, d, r, st)
    assert rc == 0, rc

for _ in range(3):
    call()
raw.gca_ws_debug_counters(buf, 1)
reps = 10
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(reps):
    call()
t1.record()
torch.cuda.synchronize()
raw.gca_ws_debug_counters(buf, 1)
ctas = 148
names = ["expand wait-FULL", "expand MMA issue", "expand wait residual", "expand epilogue+stores", "expand total",
         "gather wait-EMPTY", "gather total"]
if os.environ.get("GCA_HOP_EXPAND", "") == "":      # tcgen05 kernel: different slots
    names = ["epi wait-tfull", "epi wait-xfull", "epi tmem-ld wait", "epi math (LDS/FMA/STS)", "epi fence+store+wait_read+load",
             "epi total", "gather wait-hempty", "gather total", "mma wait-hfull", "mma wait-tempty", "mma total",
             "MAX epi total (x reps x ctas)", "prologue (entry -> loop start)", "CTA lifetime", "MAX CTA lifetime (x reps x ctas)"]
print("kernel us:", t0.elapsed_time(t1) * 1e3 / reps)
for i, nm in enumerate(names):
    val = buf[i] if nm.startswith("MAX") else buf[i] / reps / ctas
    print(f"{nm:36s} {val:12.0f} cycles" + ("" if nm.startswith("MAX") else " per CTA per launch"))
