"""GPU parity of the shared-CSR backbone propagation (gca_propagate through gconv_adapter_b200.layers.propagate)
against the golden vectors of the reference's own functions and against the CPU oracle."""
import glob
import os

import numpy as np
import pytest
import torch

from gconv_adapter_b200.graphs.csr import GLOBAL_GRAPH_CACHE
from gconv_adapter_b200.graphs.synthetic import make_graph, symmetric_random_graph
from gconv_adapter_b200.layers import propagate
from oracle import backbone_ref
from util import assert_close

pytestmark = pytest.mark.gpu
GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "backbone", "*.npz"))
              if not os.path.basename(p).startswith("mol_"))     # mol_*: the molecular layers (tests/test_*molecular*.py)


def _load(path):
    z = np.load(path)
    return {k: torch.from_numpy(z[k]) if z[k].ndim else z[k].item() for k in z.files}


def _looped(ei, n):
    return torch.cat([ei, torch.arange(n, dtype=torch.int64).repeat(2, 1)], dim=1)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_golden_vectors_of_the_reference_functions(path):
    g = _load(path)
    ei = g["edge_index"].cuda()
    x = g["x"].cuda().requires_grad_(True)
    y = propagate.gcn_conv(x, ei)
    y.backward(g["g_out"].cuda())
    assert_close(y, g["gcn_y"], "gcn_conv y")                    # rtol 1e-5 (+ 1e-5 of max|ref|), fp32
    assert_close(x.grad, g["gcn_gx"], "gcn_conv g_x")
    xb = g["x"].unsqueeze(0).cuda().requires_grad_(True)
    b = g["b"].cuda().requires_grad_(True)
    yb = propagate.add_conv_relational_bias(xb, ei, b, "sigmoid")
    yb.backward(g["g_out"].unsqueeze(0).cuda())
    assert_close(yb, g["rel_y"], "relational bias y")
    assert_close(xb.grad, g["rel_gx"], "relational bias g_x")
    assert_close(b.grad, g["rel_gb"], "relational bias g_b")


@pytest.mark.parametrize("n,e,h,dd", [(19717, 88648, 1, 64), (21168, 145780, 1, 256), (5000, 30000, 3, 20), (777, 0, 2, 2)])
def test_against_the_oracle(n, e, h, dd):
    ei = _looped(symmetric_random_graph(n, e, seed=21) if e else torch.zeros(2, 0, dtype=torch.int64), n)
    gen = torch.Generator().manual_seed(5)
    x, g_out = torch.randn(n, h, dd, generator=gen), torch.randn(n, h, dd, generator=gen)
    xr = x.clone().requires_grad_(True)
    yr = backbone_ref.gcn_conv(xr, ei)
    yr.backward(g_out)
    xo = x.cuda().requires_grad_(True)
    yo = propagate.gcn_conv(xo, ei.cuda())
    yo.backward(g_out.cuda())
    assert_close(yo, yr, f"y n={n} D={h * dd}")
    assert_close(xo.grad, xr.grad, f"g_x n={n} D={h * dd}")


def test_shares_the_adapter_graph_handle_and_is_deterministic():
    from gconv_adapter_b200 import GConvAdapter
    n = 4000
    ei = _looped(symmetric_random_graph(n, 24000, seed=22), n).cuda()
    x = torch.randn(n, 64, device="cuda")
    GConvAdapter(64, 16).cuda()(x, ei)
    before = (GLOBAL_GRAPH_CACHE.hits, GLOBAL_GRAPH_CACHE.misses)
    y1 = propagate.gcn_conv(x.reshape(n, 1, 64), ei)
    y2 = propagate.gcn_conv(x.reshape(n, 1, 64), ei)
    assert GLOBAL_GRAPH_CACHE.misses == before[1] and GLOBAL_GRAPH_CACHE.hits >= before[0] + 2   # no second build
    assert torch.equal(y1, y2)


def test_rejects_graphs_without_one_self_loop_per_node():
    n = 1000
    ei = symmetric_random_graph(n, 5000, seed=23).cuda()         # no self loops at all
    with pytest.raises(ValueError, match="self loop"):
        propagate.gcn_conv(torch.randn(n, 1, 16, device="cuda"), ei)


def test_full_size_adjoint_property():
    """<A x, y> == <x, A^T y> at the arxiv-shaped size (oracle too slow there): forward and backward kernels agree."""
    ei, n = make_graph("arxiv", seed=0)
    ei = _looped(ei, n).cuda()
    x = torch.randn(n, 1, 256, device="cuda", requires_grad=True)
    yv = torch.randn(n, 1, 256, device="cuda")
    ax = propagate.gcn_conv(x, ei)
    lhs = (ax.double() * yv.double()).sum()
    ax.backward(yv)
    rhs = (x.detach().double() * x.grad.double()).sum()
    assert abs(lhs.item() - rhs.item()) <= 1e-6 * max(abs(lhs.item()), 1.0) + 1e-3
    assert torch.isfinite(ax).all()
