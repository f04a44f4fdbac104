"""Target program for ncu: one arxiv-shaped (BASELINE.json configs[3]) adapter forward + backward
after the graph build and two warm-up steps.  Kernel launch order per step (round 2, r = 16):
project_fwd, hop_fwd, hop_expand_fwd | bwd_up, hop_bwd, hop_plain_bwd, expand_wgrad_bwd, finalize.
(round 1: ... | project_bwd, wgrad_up, hop_bwd, hop_expand_bwd, wgrad_down, finalize.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gconv_adapter_b200 import GConvAdapter
from gconv_adapter_b200.graphs.synthetic import SHAPES, make_graph, make_inputs

name = sys.argv[1] if len(sys.argv) > 1 else "arxiv"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ei, n = make_graph(name, seed=0)
s = SHAPES[name]
x, g_out, params = make_inputs(n, s.hidden, s.rank, seed=0)
m = GConvAdapter(s.hidden, s.rank, learnable_scalar=True)
sd = m.state_dict()
with torch.no_grad():
    for k, v in params.items():
        sd[k].copy_(v)
m = m.cuda()
eid, xd, gd = ei.cuda(), x.cuda().requires_grad_(True), g_out.cuda()
for _ in range(steps):
    xd.grad = None
    y = m(xd, eid)
    y.backward(gd)
torch.cuda.synchronize()
print("ok", float(y.sum()))
