"""Phase-level GPU checks straight through the C ABI (ctypes): each dense kernel against a torch fp64
evaluation of the same formula, so that a tensor-core (tcgen05, 3xTF32) kernel is pinned on its own."""
import os

import pytest
import torch

from gconv_adapter_b200 import GraphStructure, _cabi
from gconv_adapter_b200.graphs.synthetic import symmetric_random_graph

pytestmark = pytest.mark.gpu


def _graph(n, seed=0):
    ei = symmetric_random_graph(n, 6 * n, seed=seed).cuda()
    return GraphStructure(ei, n, True)


@pytest.mark.parametrize("n,d,r", [(128, 64, 16), (1000, 256, 16), (4097, 256, 32), (777, 512, 16), (50000, 256, 16),
                                   (300, 64, 32), (2708, 64, 8), (500, 300, 16)])
def test_fwd_project_matches_fp64(n, d, r):
    """P' = dis * (X Wd^T).  Shapes with d % 64 == 0 and r in {16, 32} run on tcgen05."""
    lib = _cabi.load()
    g = _graph(n)
    dis = g.arrays()["dis"]
    gen = torch.Generator(device="cuda").manual_seed(n + d + r)
    x = torch.randn(n, d, device="cuda", generator=gen)
    wd = torch.randn(r, d, device="cuda", generator=gen) * 0.1
    out = torch.full((n, r), float("nan"), device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    _cabi.check(lib.gca_fwd_project(g.handle, x.data_ptr(), x.stride(0), wd.data_ptr(), out.data_ptr(), d, r, stream), "project")
    torch.cuda.synchronize()
    ref = dis.double()[:, None] * (x.double() @ wd.double().t())
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert torch.isfinite(out).all()
    print(f"project n={n} d={d} r={r}: max err / max|ref| = {err:.3e}")
    assert err < 1e-6, f"max err / max|ref| = {err:.3e}"
    # the same call twice gives the same bits
    out2 = torch.empty_like(out)
    _cabi.check(lib.gca_fwd_project(g.handle, x.data_ptr(), x.stride(0), wd.data_ptr(), out2.data_ptr(), d, r, stream), "project")
    assert torch.equal(out, out2)
