"""CUDA-graph execution of an adapter on a STATIC graph (the transductive configs: the same ``edge_index`` and the same
input shape every step).

The small configurations (Cora- / PubMed-shaped: a few MB per call) are bound by host issue time - ten kernel launches,
the autograd bookkeeping and a dozen small allocations per forward + backward cost more than the kernels run.  Everything
libgca launches goes to the caller's stream and keeps no state outside its arguments, so a whole forward and a whole
backward can be captured once and replayed as two graph launches (``torch.cuda.make_graphed_callables``).

The graph structure is built before the capture (K0 is not part of the graph); batched molecule graphs, whose
``edge_index`` changes every step, cannot use this and stay eager.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .finetune.gconv_adapter import GConvAdapter


class _BoundAdapter(nn.Module):
    """``adapter(x, edge_index)`` as a one-argument module, so that its parameters are visible to the capture."""

    def __init__(self, adapter: GConvAdapter, edge_index: torch.Tensor):
        super().__init__()
        self.adapter = adapter
        self.edge_index = edge_index

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.adapter(x, self.edge_index)


def graphed_adapter(adapter: GConvAdapter, sample_x: torch.Tensor, edge_index: torch.Tensor, num_warmup_iters: int = 3):
    """Returns ``f(x) -> adapter(x, edge_index)`` whose forward and backward each replay one CUDA graph.

    ``sample_x`` fixes shape, dtype and ``requires_grad``; ``edge_index`` must not change afterwards (keep the returned
    callable per graph).  Gradients reach ``adapter``'s parameters exactly as in the eager call; results are bitwise
    identical to it (same kernels, same order)."""
    if not sample_x.is_cuda:
        raise RuntimeError("graphed_adapter needs CUDA tensors (there is no CPU path)")
    n = sample_x.shape[-2]
    adapter.graph_for(edge_index, n).poll(wait=True)       # structure built and validated outside the capture
    bound = _BoundAdapter(adapter, edge_index)
    sample = sample_x.detach().clone().requires_grad_(sample_x.requires_grad)
    return torch.cuda.make_graphed_callables(bound, (sample,), num_warmup_iters=num_warmup_iters)
