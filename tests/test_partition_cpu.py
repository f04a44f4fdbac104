"""world_size = 2 (and 3) over gloo on CPU: the row-partition orchestration of
gconv_adapter_b200/partition.py (shard bounds, padded gather buffers, 4 all-gathers + 1 all-reduce,
gradient hooks) with the CPU phase double, against the full-graph oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gconv_adapter_b200.graphs.synthetic import make_inputs, symmetric_random_graph
from gconv_adapter_b200.partition import PartitionedGConvAdapter, row_block, shard_size
from oracle.pyg_restated import GConvAdapterRef

from cpu_phases import CpuPhases
from util import load_module_params


def test_row_blocks_cover_all_rows_exactly_once():
    for n in (1, 7, 10, 128, 1001):
        for w in (1, 2, 3, 4, 8):
            s = shard_size(n, w)
            blocks = [row_block(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(b[1] == c[0] for b, c in zip(blocks[:-1], blocks[1:]))
            assert all(0 <= hi - lo <= s for lo, hi in blocks)
            assert all(lo == min(n, r * s) for r, (lo, hi) in enumerate(blocks))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, r, kw, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        ei = symmetric_random_graph(n, 6 * n, seed=21)
        x, g_out, params = make_inputs(n, d, r, seed=22)
        m = PartitionedGConvAdapter(d, r, phase_backend=CpuPhases(), **kw)
        load_module_params(m, params)
        lo, hi = row_block(n, world, rank)
        xl = x[lo:hi].clone().requires_grad_(True)
        y = m(xl, ei, n)
        y.backward(g_out[lo:hi])
        res = {"lo": lo, "hi": hi, "y": y.detach(), "gx": xl.grad,
               "grads": {k: p.grad.clone() for k, p in m.named_parameters()}}
        # collectives inside the step (no fused pushes): capture_step leaves the step to eager launches, on every rank,
        # and close() is safe to call more than once
        res["captured"] = m.capture_step(lambda: None)
        m.close()
        m.close()
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,kw", [
    (2, 501, dict(learnable_scalar=True)),
    (2, 501, dict(non_linearity="silu", skip_connection=False)),
    (2, 300, dict(normalization="layer_norm", learnable_scalar=True)),
    (2, 300, dict(normalize=False)),
    (3, 100, dict(learnable_scalar=True)),          # uneven shards: 34 + 34 + 32 rows
])
def test_partitioned_matches_full_graph_oracle(world, n, kw):
    d, r = 32, 8
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, d, r, kw, out), nprocs=world, join=True)
    ei = symmetric_random_graph(n, 6 * n, seed=21)
    x, g_out, params = make_inputs(n, d, r, seed=22)
    ref = GConvAdapterRef(d, r, **kw)
    load_module_params(ref, params)
    xx = x.clone().requires_grad_(True)
    yr = ref(xx, ei)
    yr.backward(g_out)
    y = torch.cat([out[k]["y"] for k in range(world)])
    gx = torch.cat([out[k]["gx"] for k in range(world)])
    assert all(out[k]["captured"] is None for k in range(world))
    assert torch.allclose(y, yr.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(gx, xx.grad, rtol=1e-4, atol=1e-5)
    for k, p in ref.named_parameters():
        for rank in range(world):       # every rank holds the full (summed) parameter gradient
            assert torch.allclose(out[rank]["grads"][k], p.grad, rtol=1e-4, atol=1e-5 * max(1.0, p.grad.abs().max().item())), (k, rank)


def test_batch_norm_is_refused():
    with pytest.raises(ValueError, match="batch_norm"):
        PartitionedGConvAdapter(16, 8, normalization="batch_norm")
