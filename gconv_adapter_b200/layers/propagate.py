"""Backbone propagation over the adapter's graph handle.

The two graph transformers of the reference add a normalised-adjacency term next to their attention:

* DIFFormer ``gcn_conv(x, edge_index, edge_weight)``  - /root/reference/src/models/transductive/difformer.py:63-79
* NodeFormer ``add_conv_relational_bias(x, edge_index, b, trans)`` - .../nodeformer.py:202-224

Both build a ``SparseTensor`` with values ``sqrt(1/d_in[col]) * sqrt(1/d[row])`` (no self loops added) and multiply it head
by head; ``d[row]`` is the IN-degree in ``gcn_conv`` and the OUT-degree in ``add_conv_relational_bias`` - the same thing on
the undirected graphs of the default pipeline, different on directed ones (``cfg.gnn_baseline.directed`` / ogbn-proteins skip
``to_undirected``), so the NodeFormer entry passes its own source-side scale.  On graphs with exactly one self loop per node
(scripts/finetune_transductive_learning.py:112-113) the in-degree matrix is the adapter's ``D^-1/2 A' D^-1/2``, so the
CSR / CSC / dis arrays the adapters already built for this ``edge_index`` are reused (``GLOBAL_GRAPH_CACHE``) and all
heads go through ONE d-wide SpMM launch (``gca_propagate``); the backward is the same kernel on the transposed CSR.

No CPU path: tensors must live on a CUDA device, anything the shared handle cannot express raises.
"""
from __future__ import annotations

import torch

from .. import _cabi
from ..graphs.csr import GLOBAL_GRAPH_CACHE, GraphStructure


def _check_self_loops(graph: GraphStructure, edge_index: torch.Tensor) -> None:
    """The shared handle equals the reference's matrix only if every node has exactly one self loop."""
    ok = getattr(graph, "_one_self_loop_per_node", None)
    if ok is None:
        loops = edge_index[0][edge_index[0] == edge_index[1]]
        counts = torch.bincount(loops, minlength=graph.num_nodes)
        ok = bool((counts == 1).all().item())
        graph._one_self_loop_per_node = ok
    if not ok:
        raise ValueError("propagation through the shared graph handle needs exactly one self loop per node in edge_index "
                         "(as scripts/finetune_transductive_learning.py:112-113 produces); got a graph without that")


def _out_degree_scale(graph: GraphStructure) -> torch.Tensor:
    """(out-degree)^-1/2 per node from the CSR by source of the shared handle (cached on the graph object)."""
    s = getattr(graph, "_out_scale", None)
    if s is None:
        rp = graph.arrays()["rowptr_t"]
        deg = (rp[1:] - rp[:-1]).to(torch.float32)
        s = (1.0 / deg).sqrt()
        s = torch.where(torch.isfinite(s), s, torch.zeros_like(s)).contiguous()
        graph._out_scale = s
    return s


class _Propagate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d: torch.Tensor, graph: GraphStructure, src_scale=None, dst_scale=None):
        """out[i] = dst_scale[i] * sum_j src_scale[j] * x[j]  (None = the handle's in-degree^-1/2)."""
        lib = _cabi.load()
        x2d = x2d.contiguous()
        out = torch.empty_like(x2d)
        stream = torch.cuda.current_stream(x2d.device).cuda_stream
        sp = src_scale.data_ptr() if src_scale is not None else None
        dp = dst_scale.data_ptr() if dst_scale is not None else None
        _cabi.check(lib.gca_propagate(graph.handle, 0, x2d.data_ptr(), x2d.stride(0), out.data_ptr(), out.stride(0),
                                      sp, dp, x2d.shape[1], stream), "gca_propagate")
        ctx.graph, ctx.src_scale, ctx.dst_scale = graph, src_scale, dst_scale
        return out

    @staticmethod
    def backward(ctx, g_out: torch.Tensor):
        lib = _cabi.load()
        g_out = g_out.contiguous()
        g_in = torch.empty_like(g_out)
        stream = torch.cuda.current_stream(g_out.device).cuda_stream
        sp = ctx.src_scale.data_ptr() if ctx.src_scale is not None else None
        dp = ctx.dst_scale.data_ptr() if ctx.dst_scale is not None else None
        _cabi.check(lib.gca_propagate(ctx.graph.handle, 1, g_out.data_ptr(), g_out.stride(0), g_in.data_ptr(), g_in.stride(0),
                                      sp, dp, g_out.shape[1], stream), "gca_propagate")
        return g_in, None, None, None


def _propagate(x2d: torch.Tensor, edge_index: torch.Tensor, source_out_degree: bool = False) -> torch.Tensor:
    if not x2d.is_cuda:
        raise RuntimeError("gconv_adapter_b200 propagation kernels need CUDA tensors (there is no CPU path)")
    if x2d.dtype != torch.float32:
        raise TypeError("fp32 only")
    if x2d.shape[1] % 4 != 0:
        raise ValueError("heads * head_dim must be a multiple of 4")
    graph = GLOBAL_GRAPH_CACHE.get(edge_index, x2d.shape[0], True)
    _check_self_loops(graph, edge_index)
    return _Propagate.apply(x2d, graph, _out_degree_scale(graph) if source_out_degree else None)


def gcn_conv(x: torch.Tensor, edge_index: torch.Tensor, edge_weight=None) -> torch.Tensor:
    """Drop-in for difformer.py:63-79: x [N, H, D] -> [N, H, D]."""
    if edge_weight is not None:
        raise NotImplementedError("edge_weight is not supported by the shared-CSR propagation (the reference passes None)")
    n, h, dd = x.shape
    return _propagate(x.reshape(n, h * dd), edge_index).reshape(n, h, dd)


def add_conv_relational_bias(x: torch.Tensor, edge_index: torch.Tensor, b: torch.Tensor, trans: str = "sigmoid") -> torch.Tensor:
    """Drop-in for nodeformer.py:202-224: x [B=1, N, H, D], b [H] -> [1, N, H, D]."""
    if trans == "sigmoid":
        scale = torch.sigmoid(b)
    elif trans == "identity":
        scale = b
    else:
        raise NotImplementedError
    bsz, n, h, dd = x.shape
    if bsz != 1:
        raise ValueError("the reference runs the graph transformers with batch size 1 (nodeformer.py:391)")
    out = _propagate(x.reshape(n, h * dd), edge_index, source_out_degree=True).reshape(1, n, h, dd)
    return out * scale.reshape(1, 1, h, 1)
