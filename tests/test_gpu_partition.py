"""GPU tests of the row-partitioned path: (1) all ranks of a world emulated one after the other on ONE
GPU through the phase-level C ABI with partitioned graph handles (kernels that wait on each other must not
be run as separate launches on one GPU, so the emulation is sequential and shares the gathered buffers);
(2) real multi-process runs (2 / 4 / 8 ranks) over NVLink peer memory and over NCCL when that many GPUs are visible
(gpurun --gpus N)."""
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


def _rank_worker(rank, world, port, n, d, r, comm, out):
    import torch.distributed as dist
    import gconv_adapter_b200.partition as part
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ei = symmetric_random_graph(n, 8 * n, seed=31)
        x, g_out, params = make_inputs(n, d, r, seed=32)
        m = PartitionedGConvAdapter(d, r, learnable_scalar=True, comm=comm)
        load_module_params(m, params)
        m = m.cuda()
        lo, hi = row_block(n, world, rank)
        eid = ei.cuda()
        gl = g_out[lo:hi].cuda()
        part.DEBUG_KEEP_SAVED = True
        res = None
        for it in range(3):                       # several steps: the barrier sequence numbers and buffer reuse are exercised
            for p_ in m.parameters():
                p_.grad = None
            xl = x[lo:hi].cuda().requires_grad_(True)
            y = m(xl, eid, n)
            mask = (part.LAST_SAVED["zp"] > 0).cpu()
            y.backward(gl)
            torch.cuda.synchronize()
            cur = {"y": y.detach().cpu(), "gx": xl.grad.cpu(), "mask": mask, "grads": {k: p_.grad.cpu() for k, p_ in m.named_parameters()}}
            if res is not None:                   # bitwise reproducible from step to step
                assert torch.equal(cur["y"], res["y"]) and torch.equal(cur["gx"], res["gx"])
                for k in cur["grads"]:
                    assert torch.equal(cur["grads"][k], res["grads"][k]), k
            res = cur
        # the same step captured in ONE CUDA graph by the module (peer exchange only) replays the same bits
        part.DEBUG_KEEP_SAVED = False
        part.LAST_SAVED.clear()
        del y, xl            # a live output keeps the eager autograd graph (and its default-stream AccumulateGrad nodes) alive
        xs = x[lo:hi].cuda().requires_grad_(True)
        held = {}

        def step():
            for p_ in m.parameters():
                p_.grad = None
            xs.grad = None
            ys = m(xs, eid, n)
            ys.backward(gl)
            held["y"] = ys.detach()

        replay = m.capture_step(step)
        assert (replay is not None) == (comm == "peer")
        if replay is not None:
            for _ in range(3):
                replay()
            torch.cuda.synchronize()
            assert torch.equal(held["y"].cpu(), res["y"]) and torch.equal(xs.grad.cpu(), res["gx"])
            for k, p_ in m.named_parameters():
                assert torch.equal(p_.grad.cpu(), res["grads"][k]), k
        out[rank] = res
        m.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,comm", [(2, "peer"), (2, "collective"), (4, "peer"), (8, "peer")])
def test_multi_gpu_run_matches_oracle(world, comm):
    """Real multi-process run (one process per GPU): NVLink peer-memory exchange (gca_push + gca_peer_barrier +
    gca_peer_allreduce) and the torch.distributed / NCCL exchange, against the full-graph CPU oracle with the tolerances
    of the single-GPU tests (the oracle is fed the ranks' own ReLU decisions, checked to differ only at ~0)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    import torch.multiprocessing as mp
    n, d, r = 40001, 256, 16
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rank_worker, args=(world, port, n, d, r, comm, out), nprocs=world, join=True)
    ei = symmetric_random_graph(n, 8 * n, seed=31)
    x, g_out, params = make_inputs(n, d, r, seed=32)
    mask = torch.cat([out[k]["mask"] for k in range(world)])
    yr, gxr, gr, ref = _oracle(ei, n, d, r, x, g_out, params, mask=mask)
    h1 = ref.last_preact
    flips = mask != (h1 > 0)
    assert flips.sum().item() <= 1e-5 * mask.numel() + 2
    if flips.any():
        assert h1[flips].abs().max().item() <= 1e-5 * h1.abs().max().item()
    y = torch.cat([out[k]["y"] for k in range(world)])
    gx = torch.cat([out[k]["gx"] for k in range(world)])
    assert_close(y, yr, f"{comm}: y")
    assert_close(gx, gxr, f"{comm}: g_x")
    for k in gr:
        floor = 5e-7 * float((g_out.abs().double() * yr.abs().double()).sum()) / 1.3 if k == "scalar" else 0.0
        for rank in range(world):
            assert_close(out[rank]["grads"][k], gr[k], f"{comm}: grad {k} on rank {rank}", noise_floor=floor)
            if comm == "peer":                    # summed in rank order on every GPU: identical bits everywhere
                assert torch.equal(out[rank]["grads"][k], out[0]["grads"][k]), k


# ---- the C-ABI entries that take the client's own ncclComm_t (include/gca.h, gca_nccl.cu) ----
def _nccl_lib():
    import ctypes as C
    lib = C.CDLL("libnccl.so.2")          # by soname: the copy torch already loaded

    class UniqueId(C.Structure):
        _fields_ = [("internal", C.c_byte * 128)]

    lib.ncclGetUniqueId.argtypes = [C.POINTER(UniqueId)]
    lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    lib.ncclCommDestroy.argtypes = [C.c_void_p]
    return lib, UniqueId


def _nccl_abi_worker(rank, world, port, n, d, r, act_name, out):
    import ctypes as C
    import torch.distributed as dist
    from gconv_adapter_b200 import GConvAdapter, _cabi
    from gconv_adapter_b200.graphs.csr import GraphCache, GraphStructure
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)       # rendezvous only: the data path is the raw communicator
    comm = C.c_void_p()
    nccl = None
    try:
        nccl, UniqueId = _nccl_lib()
        uid = UniqueId()
        if rank == 0:
            assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
        box = [bytes(uid.internal)]
        dist.broadcast_object_list(box, src=0)
        C.memmove(uid.internal, box[0], 128)
        assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0

        lib = _cabi.load()
        assert lib.gca_nccl_available() == 1
        dev = torch.device("cuda", rank)
        ei = symmetric_random_graph(n, 8 * n, seed=41)
        x, g_out, params = make_inputs(n, d, r, seed=42)
        lo, hi = row_block(n, world, rank)
        nl = hi - lo
        eid = ei.to(dev)
        graph = GraphStructure(eid, n, True, lo, hi)
        act = {"relu": 1, "silu": 2}[act_name]
        P = {k: v.to(dev).contiguous() for k, v in params.items()}
        xl, gl = x[lo:hi].to(dev).contiguous(), g_out[lo:hi].to(dev).contiguous()
        st = torch.cuda.current_stream().cuda_stream
        f32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        ws_f = torch.empty(lib.gca_forward_nccl_workspace_bytes(graph.handle, world, d, r), dtype=torch.uint8, device=dev)
        ws_b = torch.empty(lib.gca_backward_nccl_workspace_bytes(graph.handle, world, d, r), dtype=torch.uint8, device=dev)
        zp, h1, h2, y = f32(nl, r), f32(nl, r), f32(nl, r), f32(nl, d)
        gx, gwd, gbd, gwu, gbu, gs = f32(nl, d), f32(r, d), f32(r), f32(d, r), f32(d), f32(1)
        for _ in range(2):        # twice: workspaces are reusable, results repeat bit for bit
            _cabi.check(lib.gca_forward_nccl(graph.handle, comm, world, rank, xl.data_ptr(), d, P["conv_down.lin.weight"].data_ptr(),
                                             P["conv_down.bias"].data_ptr(), P["conv_up.lin.weight"].data_ptr(),
                                             P["conv_up.bias"].data_ptr(), P["scalar"].data_ptr(), act, 1, ws_f.data_ptr(),
                                             zp.data_ptr(), h1.data_ptr(), h2.data_ptr(), y.data_ptr(), d, d, r, st), "gca_forward_nccl")
            _cabi.check(lib.gca_backward_nccl(graph.handle, comm, world, rank, gl.data_ptr(), d, xl.data_ptr(), d, zp.data_ptr(),
                                              h1.data_ptr(), h2.data_ptr(), P["conv_down.lin.weight"].data_ptr(),
                                              P["conv_up.lin.weight"].data_ptr(), P["conv_up.bias"].data_ptr(), P["scalar"].data_ptr(),
                                              act, 1, ws_b.data_ptr(), gx.data_ptr(), d, gwd.data_ptr(), gbd.data_ptr(), gwu.data_ptr(),
                                              gbu.data_ptr(), gs.data_ptr(), d, r, st), "gca_backward_nccl")
            torch.cuda.synchronize()
        # the 1-GPU module on the full inputs, on this rank's GPU
        ref = GConvAdapter(d, r, non_linearity=act_name, learnable_scalar=True)
        load_module_params(ref, params)
        ref = ref.to(dev)
        ref.graph_cache = GraphCache()
        xf = x.to(dev).requires_grad_(True)
        yr = ref(xf, eid)
        yr.backward(g_out.to(dev))
        torch.cuda.synchronize()
        # same kernels, same per-row arithmetic; only the tile alignment of a shard (power-of-two box scales) may differ
        rel = lambda a, b: ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()
        assert rel(y, yr.detach()[lo:hi]) <= 2e-6, "Y rows differ from the 1-GPU module"
        assert rel(gx, xf.grad[lo:hi]) <= 2e-6, "gX rows differ from the 1-GPU module"
        got = {"conv_down.lin.weight": gwd, "conv_down.bias": gbd, "conv_up.lin.weight": gwu, "conv_up.bias": gbu, "scalar": gs}
        worst = 0.0
        for k, p_ in ref.named_parameters():
            err = (got[k].double() - p_.grad.double()).abs().max().item() / max(p_.grad.double().abs().max().item(), 1e-30)
            worst = max(worst, err if k != "scalar" else 0.0)
            if k == "scalar":       # heavily cancelling sum: fp32 noise floor as in tests/test_gpu_adapter.py::compare
                floor = 5e-7 * float((g_out.abs().double() * yr.detach().cpu().abs().double()).sum()) / abs(float(params["scalar"]))
                assert abs(float(got[k]) - float(p_.grad)) <= 1e-5 * abs(float(p_.grad)) + floor
        assert worst <= 1e-5, worst
        out[rank] = worst
        del graph
    finally:
        if nccl is not None and comm.value:
            nccl.ncclCommDestroy(comm)
        dist.destroy_process_group()


@pytest.mark.parametrize("act", ["relu", "silu"])
def test_nccl_comm_entry_points_match_single_gpu(act):
    """gca_forward_nccl / gca_backward_nccl driven by a raw ncclComm_t created through ctypes (no torch.distributed on the
    data path): Y and gX rows within 2e-6 of the 1-GPU module (same kernels), summed parameter gradients within 1e-5."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_abi_worker, args=(2, port, 30001, 256, 16, act, out), nprocs=2, join=True)
    assert len(out) == 2


def _ln_worker(rank, world, port, n, d, r, out):
    import torch.distributed as dist
    from gconv_adapter_b200 import GConvAdapter
    from gconv_adapter_b200.graphs.csr import GraphCache
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ei = symmetric_random_graph(n, 8 * n, seed=51).cuda()
        x, g_out, params = make_inputs(n, d, r, seed=52)
        lo, hi = row_block(n, world, rank)

        def build(cls):
            m = cls(d, r, normalization="layer_norm", learnable_scalar=True)
            load_module_params(m, params)
            with torch.no_grad():
                m.normalization.weight.copy_(1 + 0.2 * torch.arange(d) / d)
                m.normalization.bias.fill_(0.03)
            return m.cuda()

        part = build(PartitionedGConvAdapter)
        xl = x[lo:hi].cuda().requires_grad_(True)
        y = part(xl, ei, n)
        y.backward(g_out[lo:hi].cuda())
        ref = build(GConvAdapter)
        ref.graph_cache = GraphCache()
        xf = x.cuda().requires_grad_(True)
        yr = ref(xf, ei)
        yr.backward(g_out.cuda())
        torch.cuda.synchronize()
        rel = lambda a, b: ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()
        errs = {"y": rel(y.detach(), yr.detach()[lo:hi]), "gx": rel(xl.grad, xf.grad[lo:hi])}
        for (k, p), (_, q) in zip(sorted(part.named_parameters()), sorted(ref.named_parameters())):
            if k != "scalar":
                errs[k] = rel(p.grad, q.grad)
        out[rank] = errs
        part.close()
    finally:
        dist.destroy_process_group()


def test_partitioned_layer_norm_matches_single_gpu():
    """normalization='layer_norm' is row-local: the partitioned module (fused LayerNorm + scalar tail per shard, LayerNorm
    and scalar gradients summed over the ranks) equals the single-GPU module."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ln_worker, args=(2, port, 20001, 256, 16, out), nprocs=2, join=True)
    for rank in range(2):
        for k, v in out[rank].items():
            assert v <= (1e-5 if k not in ("y", "gx") else 2e-6), (rank, k, v)
