"""normalization='layer_norm' on the arxiv-shaped graph: the fused LayerNorm + scalar tail (gca_layernorm.cu) against the
stock torch ops after the same fused adapter kernels.   python profiles/layernorm_bench.py > profiles/r2_layernorm.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gconv_adapter_b200 import GConvAdapter
from gconv_adapter_b200.graphs.synthetic import SHAPES, make_graph, make_inputs

name = sys.argv[1] if len(sys.argv) > 1 else "arxiv"
ei, n = make_graph(name, seed=0)
s = SHAPES[name]
x, g_out, params = make_inputs(n, s.hidden, s.rank, seed=0)
eid, xd, gd = ei.cuda(), x.cuda().requires_grad_(True), g_out.cuda()
for fuse in (True, False, None):
    m = GConvAdapter(s.hidden, s.rank, normalization="layer_norm" if fuse is not None else "none", learnable_scalar=True)
    sd = m.state_dict()
    with torch.no_grad():
        for k, v in params.items():
            sd[k].copy_(v)
    m = m.cuda()
    m.fuse_layer_norm = bool(fuse)

    def step():
        xd.grad = None
        for p in m.parameters():
            p.grad = None
        y = m(xd, eid)
        y.backward(gd)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(30):
        step()
    t1.record()
    torch.cuda.synchronize()
    print(json.dumps({"workload": name, "n": n, "d": s.hidden, "r": s.rank,
                      "tail": "none (no normalisation)" if fuse is None else ("fused LayerNorm + scalar" if fuse else "torch LayerNorm + mul"),
                      "ms_per_fwd_bwd": round(t0.elapsed_time(t1) / 30, 4)}))
