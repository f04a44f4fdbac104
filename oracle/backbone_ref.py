"""CPU restatement of the backbone's normalised-adjacency propagation (TEST INFRASTRUCTURE ONLY).

Follows, line by line,
* ``gcn_conv``                  /root/reference/src/models/transductive/difformer.py:63-79
* ``add_conv_relational_bias``  /root/reference/src/models/transductive/nodeformer.py:202-224

whose third-party pieces are absent from this image and restated here from their documented behaviour:
``torch_geometric.utils.degree(index, N)`` = occurrence counts (float), ``torch_sparse.SparseTensor(row, col,
value, sparse_sizes)`` + ``torch_sparse.matmul(adj, x)`` = ``out[row[e]] += value[e] * x[col[e]]`` (sum reduce).
PARITY UNPINNED at that boundary (the reference has no tests / vectors for these functions); what IS pinned:
tests/golden/backbone_*.npz are produced by the reference's own two functions, imported unmodified from
/root/reference over ``install_backbone_shims()``.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("GCA_REFERENCE_ROOT", "/root/reference")


def degree(index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    return torch.zeros(num_nodes, dtype=torch.float32).scatter_add_(0, index, torch.ones(index.numel()))


def _spmm(row: torch.Tensor, col: torch.Tensor, value: torch.Tensor, n: int, x: torch.Tensor) -> torch.Tensor:
    """SparseTensor(row, col, value) @ x for x [n, D]."""
    return torch.zeros(n, x.shape[1], dtype=x.dtype).index_add_(0, row, value.unsqueeze(1) * x.index_select(0, col))


def gcn_conv(x: torch.Tensor, edge_index: torch.Tensor, edge_weight=None) -> torch.Tensor:
    """difformer.py:63-79.  x [N, H, D]."""
    n = x.shape[0]
    row, col = edge_index                                            # :65
    d = degree(col, n).float()                                       # :66
    d_norm_in = (1. / d[col]).sqrt()                                 # :67
    d_norm_out = (1. / d[row]).sqrt()                                # :68
    if edge_weight is None:                                          # :70-73
        value = torch.ones_like(row) * d_norm_in * d_norm_out
    else:
        value = edge_weight * d_norm_in * d_norm_out
    value = torch.nan_to_num(value, nan=0.0, posinf=0.0, neginf=0.0)  # :74
    outs = [_spmm(col, row, value, n, x[:, i]) for i in range(x.shape[1])]   # :75-77 (adj: row=col, col=row)
    return torch.stack(outs, dim=1)                                  # :78


def add_conv_relational_bias(x: torch.Tensor, edge_index: torch.Tensor, b: torch.Tensor, trans: str = "sigmoid") -> torch.Tensor:
    """nodeformer.py:202-224.  x [B, N, H, D], b [H]."""
    row, col = edge_index                                            # :207
    n = x.shape[1]
    d_in = degree(col, n).float()                                    # :208
    d_norm_in = (1. / d_in[col]).sqrt()
    d_out = degree(row, n).float()                                   # :210
    d_norm_out = (1. / d_out[row]).sqrt()
    outs = []
    for i in range(x.shape[2]):                                      # :213
        if trans == "sigmoid":
            b_i = torch.sigmoid(b[i])
        elif trans == "identity":
            b_i = b[i]
        else:
            raise NotImplementedError
        value = torch.ones_like(row) * b_i * d_norm_in * d_norm_out  # :220
        outs.append(torch.stack([_spmm(col, row, value, n, x[bb, :, i]) for bb in range(x.shape[0])], dim=0))   # :221-222
    return torch.stack(outs, dim=2)                                  # :223


# ---------------------------------------------------------------------------------------------
# running the reference's own functions (this container only)
# ---------------------------------------------------------------------------------------------
def install_backbone_shims() -> None:
    """Minimal torch_sparse / torch_geometric.utils / hydra.utils so difformer.py and nodeformer.py import."""
    if "torch_sparse" not in sys.modules:
        ts = types.ModuleType("torch_sparse")

        class SparseTensor:
            def __init__(self, row, col, value, sparse_sizes):
                self.row, self.col, self.value, self.sizes = row, col, value, sparse_sizes

        def matmul(adj, x):
            shape = x.shape
            x2 = x.reshape(shape[0], -1) if x.dim() != 2 else x
            if x.dim() == 3:                                         # [B, N, D] (nodeformer): batch over dim 0
                return torch.stack([_spmm(adj.row, adj.col, adj.value, adj.sizes[0], x[bb]) for bb in range(shape[0])], dim=0)
            return _spmm(adj.row, adj.col, adj.value, adj.sizes[0], x2).reshape(shape)

        ts.SparseTensor, ts.matmul, ts._gca_shim = SparseTensor, matmul, True
        sys.modules["torch_sparse"] = ts
    tg = sys.modules.get("torch_geometric")
    if tg is None:
        tg = types.ModuleType("torch_geometric")
        tg._gca_shim = True
        sys.modules["torch_geometric"] = tg
    if "torch_geometric.utils" not in sys.modules:
        tgu = types.ModuleType("torch_geometric.utils")
        tgu.degree = degree
        tg.utils = tgu
        sys.modules["torch_geometric.utils"] = tgu
    if "hydra" not in sys.modules:
        hy = types.ModuleType("hydra")
        hyu = types.ModuleType("hydra.utils")
        hyu.instantiate = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("hydra is not available in this image"))
        hy.utils = hyu
        sys.modules["hydra"] = hy
        sys.modules["hydra.utils"] = hyu


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "models", "transductive", "difformer.py"))


def load_reference_functions():
    """(gcn_conv, add_conv_relational_bias) imported unmodified from /root/reference."""
    install_backbone_shims()
    fns = []
    for fname, attr in (("difformer.py", "gcn_conv"), ("nodeformer.py", "add_conv_relational_bias")):
        path = os.path.join(REFERENCE_ROOT, "src", "models", "transductive", fname)
        spec = importlib.util.spec_from_file_location("_reference_" + fname[:-3], path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        fns.append(getattr(mod, attr))
    return tuple(fns)
