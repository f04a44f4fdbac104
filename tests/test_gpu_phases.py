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
    _cabi.check(lib.gca_fwd_project(g.handle, x.data_ptr(), x.stride(0), wd.data_ptr(), out.data_ptr(), None, d, r, stream), "project")
    torch.cuda.synchronize()
    ref = dis.double()[:, None] * (x.double() @ wd.double().t())
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert torch.isfinite(out).all()
    print(f"project n={n} d={d} r={r}: max err / max|ref| = {err:.3e}")
    assert err < 1e-6, f"max err / max|ref| = {err:.3e}"
    # the same call twice gives the same bits
    out2 = torch.empty_like(out)
    _cabi.check(lib.gca_fwd_project(g.handle, x.data_ptr(), x.stride(0), wd.data_ptr(), out2.data_ptr(), None, d, r, stream), "project")
    assert torch.equal(out, out2)


def _alloc_u8(nbytes):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device="cuda")


@pytest.mark.parametrize("n,d,r", [(50001, 256, 16), (2049, 128, 16), (4100, 64, 16), (2048, 32, 16), (8195, 256, 32),
                                   (169343, 256, 16), (1000, 256, 16), (3000, 300, 16), (2708, 64, 8)])
@pytest.mark.parametrize("with_scalar", [True, False])
def test_backward_dense_phases_match_fp64(n, d, r, with_scalar):
    """gca_bwd_up (one pass over gY: gH2' = dis * s * gY Wu together with the gWu / gbu partials) and the
    weight-gradient half of gca_bwd_hop1_down (gWd = gP^T X, <gY, X>), reduced by gca_bwd_finalize, against fp64.
    Shapes with d % 32 == 0, d <= 256, r in {16, 32} and n >= 2048 run on the TMA-fed streaming family
    (gca_stream.cu; r = 16 fuses projection and weight gradient in one launch); the others on the register-fed kernels."""
    lib = _cabi.load()
    g = _graph(n, seed=3)
    arr = g.arrays()
    dis = arr["dis"]
    gen = torch.Generator(device="cuda").manual_seed(7 * n + d + r)
    x = torch.randn(n, d, device="cuda", generator=gen)
    gy = torch.randn(n, d, device="cuda", generator=gen)
    h2 = torch.randn(n, r, device="cuda", generator=gen)
    z = torch.randn(n, r, device="cuda", generator=gen).abs()          # relu mask all ones
    wu = torch.randn(d, r, device="cuda", generator=gen) * 0.1
    wd = torch.randn(r, d, device="cuda", generator=gen) * 0.1
    bu = torch.randn(d, device="cuda", generator=gen) * 0.1
    s = torch.tensor([1.3], device="cuda") if with_scalar else None
    sp = s.data_ptr() if with_scalar else None
    gh2 = torch.full((n, r), float("nan"), device="cuda")
    gh1 = torch.empty(n, r, device="cuda")
    gp = torch.empty(n, r, device="cuda")
    gx = torch.empty(n, d, device="cuda")
    scratch = _alloc_u8(lib.gca_bwd_scratch_bytes(d, r))
    hub_bytes = lib.gca_hub_scratch_bytes(g.handle)
    hub = _alloc_u8(hub_bytes) if hub_bytes else None
    hp = hub.data_ptr() if hub is not None else None
    st = torch.cuda.current_stream().cuda_stream
    g_wd, g_wu = torch.empty(r, d, device="cuda"), torch.empty(d, r, device="cuda")
    g_bd, g_bu = torch.empty(r, device="cuda"), torch.empty(d, device="cuda")
    g_s = torch.empty(1, device="cuda") if with_scalar else None

    def run():
        _cabi.check(lib.gca_bwd_up(g.handle, gy.data_ptr(), gy.stride(0), h2.data_ptr(), wu.data_ptr(), sp, gh2.data_ptr(),
                                   scratch.data_ptr(), None, d, r, st), "bwd_up")
        _cabi.check(lib.gca_bwd_hop2(g.handle, gh2.data_ptr(), z.data_ptr(), None, 1, gh1.data_ptr(), scratch.data_ptr(), hp, None, r, st),
                    "bwd_hop2")
        _cabi.check(lib.gca_bwd_hop1_down(g.handle, gh1.data_ptr(), x.data_ptr(), x.stride(0), gy.data_ptr(), gy.stride(0),
                                          wd.data_ptr(), sp, 1, gp.data_ptr(), gx.data_ptr(), gx.stride(0), scratch.data_ptr(), hp,
                                          d, r, st), "bwd_hop1_down")
        _cabi.check(lib.gca_bwd_finalize(scratch.data_ptr(), wu.data_ptr(), bu.data_ptr(), sp, 1, g_wd.data_ptr(), g_bd.data_ptr(),
                                         g_wu.data_ptr(), g_bu.data_ptr(), g_s.data_ptr() if with_scalar else None, d, r, st),
                    "bwd_finalize")
        torch.cuda.synchronize()

    run()
    sv = 1.3 if with_scalar else 1.0
    X, GY, H2, WU, BU = x.double(), gy.double(), h2.double(), wu.double(), bu.double()

    def rel(a, b):
        return ((a.double() - b).abs().max() / b.abs().max()).item()

    errs = {
        "gH2'": rel(gh2, dis.double()[:, None] * sv * (GY @ WU)),
        "gWu": rel(g_wu, sv * (GY.t() @ H2)),
        "gbu": rel(g_bu, sv * GY.sum(0)),
        "gWd": rel(g_wd, gp.double().t() @ X),                           # gP as the kernels produced it
        "gX": rel(gx, gp.double() @ wd.double() + sv * GY),
    }
    if with_scalar:
        # <gY, X> is a sum of n*d products that cancel to ~sqrt(n*d): its fp32 error scales with sum |terms|, not |result|
        gs_ref = (GY * X).sum() + ((GY.t() @ H2) * WU).sum() + (GY.sum(0) * BU).sum()
        errs["gs"] = ((g_s.double() - gs_ref).abs() / (GY * X).abs().sum()).item()
    print(f"backward dense phases n={n} d={d} r={r}: " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    assert torch.isfinite(gh2).all()
    for k, v in errs.items():
        assert v < 2e-6, f"{k}: max err / max|ref| = {v:.3e}"
    # bitwise reproducible (fixed reduction orders, no atomics)
    keep = [t.clone() for t in (gh2, g_wd, g_wu, g_bu, gx)]
    run()
    for a, b in zip(keep, (gh2, g_wd, g_wu, g_bu, gx)):
        assert torch.equal(a, b)


@pytest.mark.parametrize("n,d,r,act", [(30001, 256, 16, 1), (3000, 64, 8, 2), (20011, 128, 32, 1)])
def test_nccl_entry_points_with_world_one_equal_the_single_gpu_calls(n, d, r, act):
    """gca_forward_nccl / gca_backward_nccl with world = 1 (no communicator needed) run the same phases as gca_forward /
    gca_backward on their own workspace layout: identical bits.  (The 2-GPU run with a raw ncclComm_t is in
    tests/test_gpu_partition.py.)"""
    lib = _cabi.load()
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(n + d + r)
    ei = symmetric_random_graph(n, 6 * n, seed=n).to(dev)
    graph = GraphStructure(ei, n, True)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dev)
    x, gy = rnd(n, d), rnd(n, d)
    wd, bd, wu, bu, sc = rnd(r, d) * 0.05, rnd(r) * 0.05, rnd(d, r) * 0.05, rnd(d) * 0.05, torch.tensor([1.3], device=dev)
    st = torch.cuda.current_stream().cuda_stream
    f32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    u8 = lambda nb: torch.empty(max(int(nb), 256), device=dev, dtype=torch.uint8)
    outs = []
    for nccl in (False, True):
        zp, h1, h2, y = f32(n, r), f32(n, r), f32(n, r), f32(n, d)
        gx, gwd, gbd, gwu, gbu, gs = f32(n, d), f32(r, d), f32(r), f32(d, r), f32(d), f32(1)
        if nccl:
            wf = u8(lib.gca_forward_nccl_workspace_bytes(graph.handle, 1, d, r))
            wb = u8(lib.gca_backward_nccl_workspace_bytes(graph.handle, 1, d, r))
            _cabi.check(lib.gca_forward_nccl(graph.handle, None, 1, 0, x.data_ptr(), d, wd.data_ptr(), bd.data_ptr(), wu.data_ptr(),
                                             bu.data_ptr(), sc.data_ptr(), act, 1, wf.data_ptr(), zp.data_ptr(), h1.data_ptr(),
                                             h2.data_ptr(), y.data_ptr(), d, d, r, st), "gca_forward_nccl")
            _cabi.check(lib.gca_backward_nccl(graph.handle, None, 1, 0, gy.data_ptr(), d, x.data_ptr(), d, zp.data_ptr(),
                                              h1.data_ptr(), h2.data_ptr(), wd.data_ptr(), wu.data_ptr(), bu.data_ptr(),
                                              sc.data_ptr(), act, 1, wb.data_ptr(), gx.data_ptr(), d, gwd.data_ptr(),
                                              gbd.data_ptr(), gwu.data_ptr(), gbu.data_ptr(), gs.data_ptr(), d, r, st),
                        "gca_backward_nccl")
        else:
            os.environ.pop("GCA_DISABLE_SMALL", None)
            wf = u8(lib.gca_forward_workspace_bytes(graph.handle, d, r))
            wb = u8(lib.gca_backward_workspace_bytes(graph.handle, d, r))
            _cabi.check(lib.gca_forward(graph.handle, x.data_ptr(), d, wd.data_ptr(), bd.data_ptr(), wu.data_ptr(), bu.data_ptr(),
                                        sc.data_ptr(), act, 1, wf.data_ptr(), zp.data_ptr(), h1.data_ptr(), h2.data_ptr(),
                                        y.data_ptr(), d, d, r, st), "gca_forward")
            _cabi.check(lib.gca_backward(graph.handle, gy.data_ptr(), d, x.data_ptr(), d, zp.data_ptr(), h1.data_ptr(), h2.data_ptr(),
                                         wd.data_ptr(), wu.data_ptr(), bu.data_ptr(), sc.data_ptr(), act, 1, wb.data_ptr(),
                                         gx.data_ptr(), d, gwd.data_ptr(), gbd.data_ptr(), gwu.data_ptr(), gbu.data_ptr(),
                                         gs.data_ptr(), d, r, st), "gca_backward")
        torch.cuda.synchronize()
        outs.append([t.clone() for t in (y, gx, gwd, gbd, gwu, gbu, gs)])
    names = ("y", "gx", "gwd", "gbd", "gwu", "gbu", "gs")
    small = n <= 4096        # gca_forward takes the fused small-graph kernels there, the nccl entry the phase kernels
    for k, a, b in zip(names, outs[0], outs[1]):
        if small:
            tol = 1e-5 * float(a.abs().max()) + (5e-7 * float((gy.abs().double() * outs[0][0].abs().double()).sum()) / 1.3 if k == "gs" else 0.0)
            assert float((a.double() - b.double()).abs().max()) <= tol, k
        else:
            assert torch.equal(a, b), k
