"""Generate tests/golden/*.npz with the reference's OWN GConvAdapter class.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py

The unmodified /root/reference/src/finetune/gconv_adapter.py is imported over
``oracle.pyg_shim`` (torch-geometric is not installed; its GCNConv is the restated one), run
forward + backward on CPU in fp32, and inputs / outputs / gradients are stored.  The GPU box
has no /root/reference: tests only read the committed .npz files.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import pyg_shim  # noqa: E402
from gconv_adapter_b200.graphs.synthetic import make_inputs, symmetric_random_graph  # noqa: E402

KEYS = ["scalar", "conv_down.bias", "conv_down.lin.weight", "conv_up.bias", "conv_up.lin.weight"]


def tiny_graphs():
    """Hand-checkable structures."""
    return {
        "path4": (4, [[0, 1, 1, 2, 2, 3], [1, 0, 2, 1, 3, 2]]),
        "star5": (5, [[0, 0, 0, 0, 1, 2, 3, 4], [1, 2, 3, 4, 0, 0, 0, 0]]),
        "isolated": (5, [[0, 1, 3, 4], [1, 0, 4, 3]]),                       # node 2 has no edges
        "duplicate_edge": (4, [[0, 0, 1, 1, 2, 3], [1, 1, 0, 0, 3, 2]]),     # (0,1) and (1,0) twice
        "existing_self_loops": (4, [[0, 1, 1, 2, 2, 2, 3], [1, 0, 1, 2, 2, 3, 2]]),  # loops on 1 and 2 (twice)
        "directed": (5, [[0, 1, 2, 3, 0], [1, 2, 3, 4, 4]]),                 # no reverse edges
    }


def run_case(adapter_cls, n, ei, d, r, seed, three_d=False, weight_std=0.05, **kw):
    torch.manual_seed(seed)
    x, g_out, params = make_inputs(n, d, r, seed=seed, weight_std=weight_std)
    m = adapter_cls(hidden_size=d, bottleneck_size=r, **kw)
    sd = m.state_dict()
    with torch.no_grad():
        for k in KEYS:
            if k in sd:
                sd[k].copy_(params[k])
    if isinstance(getattr(m, "normalization", None), (torch.nn.LayerNorm, torch.nn.BatchNorm1d)):
        with torch.no_grad():
            m.normalization.weight.copy_(1.0 + 0.1 * torch.randn(d, generator=torch.Generator().manual_seed(seed + 7)))
            m.normalization.bias.copy_(0.1 * torch.randn(d, generator=torch.Generator().manual_seed(seed + 8)))
    m.train()
    xin = x.unsqueeze(0) if three_d else x
    gin = g_out.unsqueeze(0) if three_d else g_out
    xin = xin.clone().requires_grad_(True)
    y = m(xin, ei)
    y.backward(gin)
    out = {"edge_index": ei.numpy(), "num_nodes": np.int64(n), "x": x.numpy(), "g_out": g_out.numpy(),
           "y": y.detach().reshape(n, d).numpy(), "g_x": xin.grad.reshape(n, d).numpy()}
    for k, p in m.named_parameters():
        out["param." + k] = p.detach().numpy()
        out["grad." + k] = p.grad.numpy() if p.grad is not None else np.zeros_like(p.detach().numpy())
    return out


def main():
    if not pyg_shim.reference_available():
        raise SystemExit("reference checkout not found; golden vectors can only be regenerated in the build container")
    cls = pyg_shim.load_reference_adapter()
    base = dict(conv_type="gcn", non_linearity="relu", normalization="none", learnable_scalar=True,
                skip_connection=True, normalize=True)
    cases = {}
    for name, (n, e) in tiny_graphs().items():
        ei = torch.tensor(e, dtype=torch.int64)
        cases["tiny_" + name] = (run_case(cls, n, ei, 8, 8, seed=11, **base), base)
    n = 300
    ei = symmetric_random_graph(n, 1200, seed=3)
    variants = {
        "mid_default": {},
        "mid_silu": {"non_linearity": "silu"},
        "mid_identity": {"non_linearity": "none"},
        "mid_noskip": {"skip_connection": False},
        "mid_noscalar": {"learnable_scalar": False},
        "mid_unnormalized": {"normalize": False},             # the shipped YAML default
        "mid_layernorm": {"normalization": "layer_norm"},
        "mid_batchnorm": {"normalization": "batch_norm"},
    }
    for name, over in variants.items():
        kw = dict(base, **over)
        cases[name] = (run_case(cls, n, ei, 32, 8, seed=5, **kw), kw)
    cases["mid_3d_input"] = (run_case(cls, n, ei, 32, 8, seed=5, three_d=True, **base), dict(base, three_d=True))
    cases["mid_r16"] = (run_case(cls, n, ei, 64, 16, seed=6, **base), base)
    cases["mid_odd_shapes"] = (run_case(cls, n, ei, 30, 5, seed=7, **base), base)       # d % 4 != 0, r not a kernel rank
    cases["mid_init_scale"] = (run_case(cls, n, ei, 32, 8, seed=5, weight_std=1e-5, **base), base)
    # BASELINE.json configs[0]: Cora-shaped, the reference's own CPU-runnable case (rows subsampled to keep it small)
    n1 = 2708
    ei1 = symmetric_random_graph(n1, 10556, seed=0)
    c1 = run_case(cls, n1, ei1, 64, 8, seed=0, **base)
    rows = np.arange(0, n1, 8)
    c1_small = {k: v for k, v in c1.items() if k not in ("x", "g_out", "y", "g_x", "edge_index")}
    c1_small["rows"] = rows
    c1_small["y_rows"] = c1["y"][rows]
    c1_small["g_x_rows"] = c1["g_x"][rows]
    c1_small["y_sum64"] = np.float64(c1["y"].astype(np.float64).sum())
    c1_small["g_x_sum64"] = np.float64(c1["g_x"].astype(np.float64).sum())
    c1_small["graph_seed"] = np.int64(0)
    c1_small["input_seed"] = np.int64(0)
    cases["cora_shaped"] = (c1_small, base)

    for name, (arrs, kw) in cases.items():
        meta = {("cfg." + k): np.array(v) for k, v in kw.items()}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs, **meta)
        print(name, {k: getattr(v, "shape", None) for k, v in arrs.items() if k in ("x", "y", "y_rows")})


if __name__ == "__main__":
    main()
