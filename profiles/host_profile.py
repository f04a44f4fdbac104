"""Where the host time of a tiny-config step goes (cora-shaped adapter, eager): cProfile over 2000 fwd+bwd steps.

    python profiles/host_profile.py [workload] > profiles/r2_host_profile.txt
"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gconv_adapter_b200 import GConvAdapter
from gconv_adapter_b200.graphs.synthetic import SHAPES, make_graph, make_inputs

name = sys.argv[1] if len(sys.argv) > 1 else "cora"
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


def step():
    xd.grad = None
    y = m(xd, eid)
    y.backward(gd)


for _ in range(200):
    step()
torch.cuda.synchronize()
K = 2000
t0 = time.perf_counter()
for _ in range(K):
    step()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"{name}: host issue time {1e6 * t_issue / K:.1f} us/step, wall incl. final sync {1e6 * t_all / K:.1f} us/step")
t0 = time.perf_counter()
for _ in range(K):
    y = m(xd, eid)
t_f = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"forward only: {1e6 * t_f / K:.1f} us/call issue")
pr = cProfile.Profile()
pr.enable()
for _ in range(K):
    step()
pr.disable()
torch.cuda.synchronize()
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats("tottime").print_stats(28)
print(out.getvalue())
