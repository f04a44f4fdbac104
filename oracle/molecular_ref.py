"""CPU restatement of the molecular backbone layers' message passing (TEST INFRASTRUCTURE ONLY).

Follows, line by line,
* ``MolecularGINConv``  /root/reference/src/layers/inductive/gin_conv.py:9-85
* ``MolecularGCNConv``  /root/reference/src/layers/inductive/gcn_conv.py:9-98

Their third-party base class is absent from this image and restated here from its documented behaviour:
``torch_geometric.nn.MessagePassing(aggr='add')`` with the default flow source_to_target: ``propagate(edge_index, x=...,
**kw)`` calls ``message(x_j = x[edge_index[0]], **kw)``, sums the messages at ``edge_index[1]`` and passes the sums to
``update``; ``torch_geometric.utils.add_self_loops(edge_index, num_nodes=N)`` returns ``(cat([edge_index, arange(N) x 2]),
None)`` (the reference indexes the tuple with ``[0]``).  PARITY UNPINNED at that boundary; what IS pinned:
tests/golden/backbone/mol_*.npz are produced by the reference's own two classes, imported unmodified from
/root/reference over ``install_message_passing_shim()``.
"""
from __future__ import annotations

import importlib.util
import inspect
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("GCA_REFERENCE_ROOT", "/root/reference")
NUM_BOND_TYPE, NUM_BOND_DIRECTION = 6, 3          # gin_conv.py:6-7


def add_self_loops(edge_index: torch.Tensor, num_nodes: int):
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1), None


def _edge_embeddings(edge_attr, n, w_type, w_dir):
    """gin_conv.py:52-57 / gcn_conv.py:73-79: self-loop rows (bond type 4, direction 0) appended, two embedding lookups."""
    loop_attr = torch.zeros(n, 2, dtype=edge_attr.dtype, device=edge_attr.device)
    loop_attr[:, 0] = 4
    ea = torch.cat([edge_attr, loop_attr], dim=0)
    return w_type[ea[:, 0]] + w_dir[ea[:, 1]]


def gin_aggregate(x, edge_index, edge_attr, w_type, w_dir):
    """What MolecularGINConv hands to its MLP: sum over incoming edges (self loop added) of x_j + edge embedding
    (gin_conv.py:50-72)."""
    n = x.size(0)
    ei, _ = add_self_loops(edge_index, n)                                           # :50
    emb = _edge_embeddings(edge_attr, n, w_type, w_dir)                             # :52-57
    msg = x.index_select(0, ei[0]) + emb                                            # message, :72
    return torch.zeros_like(x).index_add_(0, ei[1], msg)                            # aggr = "add"


def gcn_aggregate(x_lin, edge_index, edge_attr, w_type, w_dir):
    """MolecularGCNConv after its Linear: sum of norm * (x_j + edge embedding), norm = deg^-1/2[row] * deg^-1/2[col] with the
    degree counted over the SOURCE index incl. the added loops (gcn_conv.py:36-56, :69-98)."""
    n = x_lin.size(0)
    ei, _ = add_self_loops(edge_index, n)                                           # :71
    emb = _edge_embeddings(edge_attr, n, w_type, w_dir)                             # :73-79
    row, col = ei
    w = torch.ones(row.numel(), dtype=x_lin.dtype, device=x_lin.device)            # :49
    deg = torch.zeros(n, dtype=w.dtype, device=w.device).scatter_add_(0, row, w)   # :51-52
    dis = deg.pow(-0.5)                                                             # :53
    dis[dis == float("inf")] = 0                                                    # :54
    norm = dis[row] * w * dis[col]                                                  # :56
    msg = norm.view(-1, 1) * (x_lin.index_select(0, row) + emb)                     # message, :98
    return torch.zeros_like(x_lin).index_add_(0, col, msg)


class MolecularGINConvRef(nn.Module):
    """Same parameters / state_dict keys as the reference class (gin_conv.py:18-36)."""

    def __init__(self, emb_dim, aggr="add"):
        super().__init__()
        self.aggr = aggr
        self.mlp = nn.Sequential(nn.Linear(emb_dim, 2 * emb_dim), nn.ReLU(), nn.Linear(2 * emb_dim, emb_dim))
        self.edge_embedding_type = nn.Embedding(NUM_BOND_TYPE, emb_dim)
        self.edge_embedding_direction = nn.Embedding(NUM_BOND_DIRECTION, emb_dim)
        torch.nn.init.xavier_uniform_(self.edge_embedding_type.weight.data)
        torch.nn.init.xavier_uniform_(self.edge_embedding_direction.weight.data)

    def forward(self, x, edge_index, edge_attr):
        return self.mlp(gin_aggregate(x, edge_index, edge_attr, self.edge_embedding_type.weight, self.edge_embedding_direction.weight))


class MolecularGCNConvRef(nn.Module):
    """Same parameters / state_dict keys as the reference class (gcn_conv.py:18-34)."""

    def __init__(self, emb_dim, aggr="add"):
        super().__init__()
        self.aggr, self.emb_dim = aggr, emb_dim
        self.linear = nn.Linear(emb_dim, emb_dim)
        self.edge_embedding_type = nn.Embedding(NUM_BOND_TYPE, emb_dim)
        self.edge_embedding_direction = nn.Embedding(NUM_BOND_DIRECTION, emb_dim)
        torch.nn.init.xavier_uniform_(self.edge_embedding_type.weight.data)
        torch.nn.init.xavier_uniform_(self.edge_embedding_direction.weight.data)

    def forward(self, x, edge_index, edge_attr):
        return gcn_aggregate(self.linear(x), edge_index, edge_attr, self.edge_embedding_type.weight, self.edge_embedding_direction.weight)


# ---------------------------------------------------------------------------------------------
# running the reference's own classes (this container only)
# ---------------------------------------------------------------------------------------------
class _MessagePassingShim(nn.Module):
    """The subset of torch_geometric.nn.MessagePassing the two reference layers use."""

    def __init__(self, aggr="add", **kwargs):
        super().__init__()
        if aggr != "add":
            raise NotImplementedError("shim: aggr='add' only (the reference's default)")

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        x = kwargs["x"]
        names = [p for p in inspect.signature(self.message).parameters]
        args = {}
        for name in names:
            if name == "x_j":
                args[name] = x.index_select(0, src)
            elif name == "x_i":
                args[name] = x.index_select(0, dst)
            else:
                args[name] = kwargs[name]
        msg = self.message(**args)
        out = torch.zeros(x.size(0), msg.size(1), dtype=msg.dtype, device=msg.device).index_add_(0, dst, msg)
        return self.update(out)

    def update(self, aggr_out):
        return aggr_out


def install_message_passing_shim() -> None:
    tg = sys.modules.get("torch_geometric")
    if tg is None:
        tg = types.ModuleType("torch_geometric")
        tg._gca_shim = True
        sys.modules["torch_geometric"] = tg
    nn_mod = sys.modules.get("torch_geometric.nn")
    if nn_mod is None:
        nn_mod = types.ModuleType("torch_geometric.nn")
        sys.modules["torch_geometric.nn"] = nn_mod
        tg.nn = nn_mod
    if not hasattr(nn_mod, "MessagePassing"):
        nn_mod.MessagePassing = _MessagePassingShim
    utils = sys.modules.get("torch_geometric.utils")
    if utils is None:
        utils = types.ModuleType("torch_geometric.utils")
        sys.modules["torch_geometric.utils"] = utils
        tg.utils = utils
    if not hasattr(utils, "add_self_loops"):
        utils.add_self_loops = add_self_loops


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "layers", "inductive", "gin_conv.py"))


def load_reference_classes():
    """(MolecularGINConv, MolecularGCNConv) imported unmodified from /root/reference."""
    install_message_passing_shim()
    out = []
    for fname, attr in (("gin_conv.py", "MolecularGINConv"), ("gcn_conv.py", "MolecularGCNConv")):
        path = os.path.join(REFERENCE_ROOT, "src", "layers", "inductive", fname)
        spec = importlib.util.spec_from_file_location("_reference_" + fname[:-3], path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        out.append(getattr(mod, attr))
    return tuple(out)
