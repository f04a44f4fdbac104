"""Molecular backbone layers on the adapter's graph handle (SURVEY section 8f, rank 3 - the inductive half).

Drop-ins for the reference's ``MolecularGINConv`` (/root/reference/src/layers/inductive/gin_conv.py:9-85) and
``MolecularGCNConv`` (.../gcn_conv.py:9-98): same constructors, parameter names / ``state_dict`` keys and
``forward(x, edge_index, edge_attr)``.  The reference materialises one ``[E + N, emb_dim]`` message tensor
(``x_j + edge_embedding``) per layer and scatter-adds it with atomics.  Here the message sum is split into

    sum_e x[src_e]                       one d-wide SpMM over the CSR the adapters already built for this edge_index
                                         (``gca_propagate`` with unit / degree scales; ``GLOBAL_GRAPH_CACHE``: the five
                                         layers and all adapters of a step share ONE build), and
    sum_e emb(type_e) + emb(dir_e)       = C_type @ E_type + C_dir @ E_dir with per-node (weighted) histograms of the 6 bond
                                         types and 3 directions - a [N, 6] and a [N, 3] matrix times the embedding tables,

so no per-edge d-wide tensor and no edge-order permutation of the CSR is needed.  No CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..graphs.csr import GLOBAL_GRAPH_CACHE
from .propagate import _Propagate, _out_degree_scale

NUM_BOND_TYPE = 6
NUM_BOND_DIRECTION = 3
SELF_LOOP_BOND_TYPE = 4      # gin_conv.py:54, gcn_conv.py:75


def _shared_graph(x: torch.Tensor, edge_index: torch.Tensor):
    if not x.is_cuda:
        raise RuntimeError("gconv_adapter_b200 molecular layers need CUDA tensors (there is no CPU path)")
    if x.dtype != torch.float32 or x.shape[1] % 4 != 0:
        raise RuntimeError("fp32 features with emb_dim % 4 == 0 only")
    n = x.shape[0]
    graph = GLOBAL_GRAPH_CACHE.get(edge_index, n, True)
    ok = getattr(graph, "_no_self_loops", None)
    if ok is None:
        # the handle replaces existing self loops by exactly one per node, add_self_loops ADDS one: equal iff none exist
        ok = not bool((edge_index[0] == edge_index[1]).any().item())
        graph._no_self_loops = ok
    if not ok:
        raise ValueError("the shared graph handle equals add_self_loops(edge_index) only for graphs without self loops "
                         "(molecule graphs have none)")
    return graph, n


def _histograms(edge_index, edge_attr, n, weight=None, loop_weight=None):
    """Per target node: (weighted) counts of the bond types / directions of its incoming edges, self loop included."""
    dst = edge_index[1]
    w = torch.ones(dst.numel(), dtype=torch.float32, device=dst.device) if weight is None else weight
    ct = torch.zeros(n * NUM_BOND_TYPE, dtype=torch.float32, device=dst.device)
    ct.index_add_(0, dst * NUM_BOND_TYPE + edge_attr[:, 0], w)
    cd = torch.zeros(n * NUM_BOND_DIRECTION, dtype=torch.float32, device=dst.device)
    cd.index_add_(0, dst * NUM_BOND_DIRECTION + edge_attr[:, 1], w)
    ct, cd = ct.view(n, NUM_BOND_TYPE), cd.view(n, NUM_BOND_DIRECTION)
    lw = torch.ones(n, dtype=torch.float32, device=dst.device) if loop_weight is None else loop_weight
    ct[:, SELF_LOOP_BOND_TYPE] += lw                 # the added loop: bond type 4, direction 0
    cd[:, 0] += lw
    return ct, cd


class MolecularGINConv(nn.Module):
    """Drop-in for gin_conv.py:9-85 (aggr='add')."""

    def __init__(self, emb_dim, aggr="add"):
        super().__init__()
        if aggr != "add":
            raise NotImplementedError("aggr='add' (the reference's default) only")
        self.aggr = aggr
        self.mlp = nn.Sequential(nn.Linear(emb_dim, 2 * emb_dim), nn.ReLU(), nn.Linear(2 * emb_dim, emb_dim))
        self.edge_embedding_type = nn.Embedding(NUM_BOND_TYPE, emb_dim)
        self.edge_embedding_direction = nn.Embedding(NUM_BOND_DIRECTION, emb_dim)
        torch.nn.init.xavier_uniform_(self.edge_embedding_type.weight.data)
        torch.nn.init.xavier_uniform_(self.edge_embedding_direction.weight.data)

    def aggregate(self, x, edge_index, edge_attr):
        graph, n = _shared_graph(x, edge_index)
        ones = getattr(graph, "_ones", None)
        if ones is None:
            ones = graph._ones = torch.ones(n, dtype=torch.float32, device=x.device)
        ct, cd = _histograms(edge_index, edge_attr, n)
        return (_Propagate.apply(x, graph, ones, ones)
                + ct @ self.edge_embedding_type.weight + cd @ self.edge_embedding_direction.weight)

    def forward(self, x, edge_index, edge_attr):
        return self.mlp(self.aggregate(x, edge_index, edge_attr))


class MolecularGCNConv(nn.Module):
    """Drop-in for gcn_conv.py:9-98 (aggr='add'): norm = deg^-1/2[src] * deg^-1/2[dst], degree counted over the source index."""

    def __init__(self, emb_dim, aggr="add"):
        super().__init__()
        if aggr != "add":
            raise NotImplementedError("aggr='add' (the reference's default) only")
        self.aggr, self.emb_dim = aggr, emb_dim
        self.linear = nn.Linear(emb_dim, emb_dim)
        self.edge_embedding_type = nn.Embedding(NUM_BOND_TYPE, emb_dim)
        self.edge_embedding_direction = nn.Embedding(NUM_BOND_DIRECTION, emb_dim)
        torch.nn.init.xavier_uniform_(self.edge_embedding_type.weight.data)
        torch.nn.init.xavier_uniform_(self.edge_embedding_direction.weight.data)

    def forward(self, x, edge_index, edge_attr):
        graph, n = _shared_graph(x, edge_index)
        dis = _out_degree_scale(graph)                       # (out-degree incl. the loop)^-1/2: the reference's `degree` over row
        h = self.linear(x)
        ct, cd = _histograms(edge_index, edge_attr, n, weight=dis[edge_index[0]], loop_weight=dis)
        emb = ct @ self.edge_embedding_type.weight + cd @ self.edge_embedding_direction.weight
        return _Propagate.apply(h, graph, dis, dis) + dis.unsqueeze(1) * emb
