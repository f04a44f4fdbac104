"""CPU checks of the host harnesses with the oracle adapter: registration names, freeze policy, the
DIFFormer pre-adapter quirk, parallel vs sequential insertion (no CUDA needed)."""
import functools

import torch

from gconv_adapter_b200.graphs.synthetic import molecule_batch, symmetric_random_graph
from hosts import MolecularGraphPredictionHost, TransductiveHost
from oracle.pyg_restated import GConvAdapterRef


def _factory(**kw):
    return functools.partial(GConvAdapterRef, bottleneck_size=8, learnable_scalar=True, **kw)


def test_inductive_host_registers_and_freezes_like_the_reference():
    torch.manual_seed(0)
    m = MolecularGraphPredictionHost(num_layers=3, emb_dim=32)
    m.add_adapter(_factory(), ["post"], "sequential")
    names = [n for n, p in m.named_parameters() if p.requires_grad]
    assert "gnn.post_adapters.0.scalar" in names and "gnn.post_adapters.2.conv_up.lin.weight" in names
    assert "graph_pred_linear.weight" in names and "gnn.batch_norm_layers.0.weight" in names
    assert not any(n.startswith("gnn.gnn_layers") or n.startswith("node_embedding") for n in names)
    ei, batch, n = molecule_batch(batch_size=4, seed=0)
    x = torch.stack([torch.randint(0, 120, (n,)), torch.randint(0, 3, (n,))], 1)
    ea = torch.stack([torch.randint(0, 4, (ei.size(1),)), torch.randint(0, 3, (ei.size(1),))], 1)
    out = m(x, ei, ea, batch)
    assert out.shape == (4, 1)
    out.sum().backward()
    assert m.gnn.post_adapters["1"].conv_down.lin.weight.grad is not None


def test_transductive_hosts_shapes_quirk_and_parallel_mode():
    torch.manual_seed(1)
    n = 60
    ei = symmetric_random_graph(n, 240, seed=2, add_self_loops=True)
    x = torch.randn(n, 10)
    for kind in ("nodeformer", "difformer"):
        for typ in ("sequential", "parallel"):
            m = TransductiveHost(kind, 10, 16, 3)
            m.add_adapter(_factory(), ["pre", "post"], typ)
            y = m(x, [ei])
            assert y.shape == (n, 3)
            y.sum().backward()
            assert all(p.grad is not None for n_, p in m.named_parameters() if "adapter" in n_)
    # DIFFormer quirk: with alpha = 1 the residual link is unused, so a sequential PRE adapter has no effect
    m = TransductiveHost("difformer", 10, 16, 3, alpha=1.0)
    y0 = m(x, [ei])
    m.add_adapter(_factory(), ["pre"], "sequential")
    with torch.no_grad():
        for a in m.pre_adapters.values():
            a.conv_up.lin.weight.normal_(0, 1.0)
    assert torch.equal(m(x, [ei]), y0)
