"""GPU tests of the row-partitioned path: (1) all ranks of a world emulated one after the other on ONE
GPU through the phase-level C ABI with partitioned graph handles (kernels that wait on each other must not
be run as separate launches on one GPU, so the emulation is sequential and shares the gathered buffers);
(2) a real 2-rank NCCL run when two GPUs are visible (gpurun --gpus 2)."""
import os
import socket

import pytest
import torch

from gconv_adapter_b200.graphs.synthetic import make_graph, make_inputs, symmetric_random_graph
from gconv_adapter_b200.partition import CudaPhases, PartitionedGConvAdapter, row_block, shard_size
from oracle.pyg_restated import GConvAdapterRef

from util import assert_close, load_module_params

pytestmark = pytest.mark.gpu


def _oracle(ei, n, d, r, x, g_out, params, mask=None):
    ref = GConvAdapterRef(d, r, learnable_scalar=True)
    load_module_params(ref, params)
    ref.relu_mask_override = mask
    xx = x.clone().requires_grad_(True)
    y = ref(xx, ei)
    y.backward(g_out)
    return y.detach(), xx.grad, {k: p.grad for k, p in ref.named_parameters()}, ref


def _emulate(world, ei, n, d, r, x, g_out, params):
    be = CudaPhases()
    dev = torch.device("cuda:0")
    eid = ei.to(dev)
    s = shard_size(n, world)
    blocks = [row_block(n, world, k) for k in range(world)]
    graphs = [be.build_graph(eid, n, True, lo, hi) for lo, hi in blocks]
    P = {k: v.to(dev).contiguous() for k, v in params.items()}
    wd, bd, wu, bu, sc = (P["conv_down.lin.weight"], P["conv_down.bias"], P["conv_up.lin.weight"], P["conv_up.bias"], P["scalar"])
    xs = [x[lo:hi].to(dev).contiguous() for lo, hi in blocks]
    gys = [g_out[lo:hi].to(dev).contiguous() for lo, hi in blocks]
    full = lambda: torch.zeros(world * s, r, device=dev)
    p_full, z_full, gh2_full, gh1_full = full(), full(), full(), full()
    h2 = [torch.empty(max(hi - lo, 1), r, device=dev) for lo, hi in blocks]
    ys = [torch.empty(hi - lo, d, device=dev) for lo, hi in blocks]
    gxs = [torch.empty(hi - lo, d, device=dev) for lo, hi in blocks]
    act = 1
    live = [k for k, (lo, hi) in enumerate(blocks) if hi > lo]
    for k in live:
        lo, hi = blocks[k]
        be.fwd_project(graphs[k], xs[k], wd, p_full[lo:hi])
    for k in live:
        lo, hi = blocks[k]
        be.fwd_hop1(graphs[k], p_full, bd, act, z_full[lo:hi], None)
    for k in live:
        be.fwd_hop2_up(graphs[k], z_full, xs[k], wu, bu, sc, True, h2[k], ys[k])
    scr = [be.bwd_scratch(d, r, dev) for _ in blocks]
    for k in live:
        lo, hi = blocks[k]
        be.bwd_up(graphs[k], gys[k], h2[k], wu, sc, gh2_full[lo:hi], scr[k])
    for k in live:
        lo, hi = blocks[k]
        be.bwd_hop2(graphs[k], gh2_full, z_full[lo:hi], None, act, gh1_full[lo:hi], scr[k])
    grads = None
    for k in live:
        lo, hi = blocks[k]
        gp = torch.empty(hi - lo, r, device=dev)
        be.bwd_hop1_down(graphs[k], gh1_full, xs[k], gys[k], wd, sc, True, gp, gxs[k], scr[k])
        g = {"conv_down.lin.weight": torch.empty_like(wd), "conv_down.bias": torch.empty_like(bd),
             "conv_up.lin.weight": torch.empty_like(wu), "conv_up.bias": torch.empty_like(bu), "scalar": torch.empty(1, device=dev)}
        be.bwd_finalize(scr[k], wu, bu, sc, True, g["conv_down.lin.weight"], g["conv_down.bias"],
                        g["conv_up.lin.weight"], g["conv_up.bias"], g["scalar"])
        grads = g if grads is None else {n_: grads[n_] + g[n_] for n_ in g}
    torch.cuda.synchronize()
    return torch.cat(ys), torch.cat(gxs), grads, z_full[:n]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_emulated_ranks_match_oracle_arxiv_quarter(world):
    ei, n = make_graph("arxiv", seed=2, scale=0.25)
    d, r = 256, 16
    x, g_out, params = make_inputs(n, d, r, seed=3)
    y, gx, grads, zp = _emulate(world, ei, n, d, r, x, g_out, params)
    yr, gxr, gr, _ = _oracle(ei, n, d, r, x, g_out, params, mask=(zp > 0).cpu())
    assert_close(y, yr, f"world={world}: y")
    assert_close(gx, gxr, f"world={world}: g_x")
    for k in gr:
        assert_close(grads[k], gr[k], f"world={world}: grad {k}")


def test_emulated_ranks_uneven_and_empty_blocks():
    n, d, r = 10, 16, 8                  # world 4 -> S = 3: blocks 3,3,3,1 ; world 8 -> S = 2: three empty ranks
    ei = symmetric_random_graph(n, 30, seed=5)
    x, g_out, params = make_inputs(n, d, r, seed=6)
    for world in (4, 8):
        y, gx, grads, zp = _emulate(world, ei, n, d, r, x, g_out, params)
        yr, gxr, gr, _ = _oracle(ei, n, d, r, x, g_out, params, mask=(zp > 0).cpu())
        assert_close(y, yr, "y")
        assert_close(gx, gxr, "g_x")
        for k in gr:
            assert_close(grads[k], gr[k], k)


def _nccl_worker(rank, world, port, n, d, r, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ei = symmetric_random_graph(n, 8 * n, seed=31)
        x, g_out, params = make_inputs(n, d, r, seed=32)
        m = PartitionedGConvAdapter(d, r, learnable_scalar=True)
        load_module_params(m, params)
        m = m.cuda()
        lo, hi = row_block(n, world, rank)
        xl = x[lo:hi].cuda().requires_grad_(True)
        y = m(xl, ei.cuda(), n)
        y.backward(g_out[lo:hi].cuda())
        torch.cuda.synchronize()
        out[rank] = {"y": y.detach().cpu(), "gx": xl.grad.cpu(), "grads": {k: p.grad.cpu() for k, p in m.named_parameters()}}
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_run_matches_oracle():
    import torch.multiprocessing as mp
    n, d, r, world = 40001, 256, 16, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_worker, args=(world, port, n, d, r, out), nprocs=world, join=True)
    ei = symmetric_random_graph(n, 8 * n, seed=31)
    x, g_out, params = make_inputs(n, d, r, seed=32)
    yr, gxr, gr, _ = _oracle(ei, n, d, r, x, g_out, params)
    y = torch.cat([out[k]["y"] for k in range(world)])
    gx = torch.cat([out[k]["gx"] for k in range(world)])
    assert_close(y, yr, "nccl: y", max_outlier_frac=1e-5)
    assert_close(gx, gxr, "nccl: g_x", max_outlier_frac=1e-4)
    for k in gr:
        for rank in range(world):
            assert_close(out[rank]["grads"][k], gr[k], f"nccl: grad {k} on rank {rank}", rtol=1e-4, atol_scale=1e-4)
