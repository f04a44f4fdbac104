"""numpy restatement of the graph-structure build (integer indexing + coefficients).

TEST INFRASTRUCTURE ONLY.  The reference has no CSR anywhere (PyG gathers/scatters COO,
SURVEY.md section 0 item 3), so the *contract* the CUDA build (K0) is held to is:

* integer part, bit-exact: the multiset of (target, source) pairs produced by
  ``add_remaining_self_loops`` (``oracle.pyg_restated``; normalize=True) or the input
  pairs unchanged (normalize=False), laid out as CSR by target with the sources of every
  row in ascending order, and as CSR by source (the transpose, for the backward) with
  targets ascending;
* fp32 part: ``dis = deg.pow(-0.5)`` (inf -> 0) with deg the in-degree incl. the loop, and
  per-edge ``coef = dis[row] * 1 * dis[col]`` (``gcn_norm``), compared in ULPs.
"""
from __future__ import annotations

import numpy as np
import torch

from .pyg_restated import add_remaining_self_loops, gcn_norm


def coo_after_loops(edge_index: np.ndarray, num_nodes: int, normalize: bool) -> np.ndarray:
    ei = torch.from_numpy(np.ascontiguousarray(edge_index).astype(np.int64))
    if normalize:
        ei = add_remaining_self_loops(ei, num_nodes)
    return ei.numpy()


def csr_by(keys: np.ndarray, vals: np.ndarray, num_nodes: int, lo: int = 0, hi: int | None = None):
    """CSR over rows [lo, hi) keyed on ``keys`` with ``vals`` ascending inside each row."""
    hi = num_nodes if hi is None else hi
    sel = (keys >= lo) & (keys < hi)
    k, v = keys[sel], vals[sel]
    order = np.lexsort((v, k))
    k, v = k[order], v[order]
    counts = np.bincount(k - lo, minlength=hi - lo)
    rowptr = np.zeros(hi - lo + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr.astype(np.int32), v.astype(np.int32)


def build(edge_index: np.ndarray, num_nodes: int, normalize: bool = True, lo: int = 0, hi: int | None = None):
    """Return dict(rowptr, colidx, rowptr_t, colidx_t, dis, nnz) for rows [lo, hi)."""
    ei = coo_after_loops(edge_index, num_nodes, normalize)
    src, dst = ei[0], ei[1]
    rowptr, colidx = csr_by(dst, src, num_nodes, lo, hi)        # forward: row = target, gathers sources
    rowptr_t, colidx_t = csr_by(src, dst, num_nodes, lo, hi)    # backward: row = source, gathers targets
    if normalize:
        deg = np.bincount(dst, minlength=num_nodes).astype(np.float32)
        dis = torch.from_numpy(deg).pow(-0.5)
        dis[dis == float("inf")] = 0
        dis = dis.numpy()
    else:
        dis = np.ones(num_nodes, dtype=np.float32)
    return dict(rowptr=rowptr, colidx=colidx, rowptr_t=rowptr_t, colidx_t=colidx_t,
                dis=dis, nnz=int(colidx.shape[0]), nnz_t=int(colidx_t.shape[0]))


def edge_coefficients(edge_index: np.ndarray, num_nodes: int):
    """(edge_index', coef) straight from the restated ``gcn_norm`` (fp32)."""
    ei, w = gcn_norm(torch.from_numpy(np.ascontiguousarray(edge_index).astype(np.int64)), num_nodes, torch.float32)
    return ei.numpy(), w.numpy()


def ulp_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units in the last place between two fp32 arrays (same sign assumed)."""
    ai = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    bi = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(ai - bi)
