"""Pure-torch CPU restatement of the reference's GConv-Adapter hot path.

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see ``oracle/__init__.py``).

What is restated and where it comes from
-----------------------------------------
* ``GConvAdapterRef`` follows /root/reference/src/finetune/gconv_adapter.py:23-108 line
  for line (constructor arguments, error messages, parameter names, init, forward order:
  conv_up(act(conv_down(x))) -> ``out += x`` -> BatchNorm1d(only if out.size(0) > 1) |
  LayerNorm -> ``out * scalar``).
* ``GCNConvRef`` / ``gcn_norm`` / ``add_remaining_self_loops`` restate the *published*
  algorithm of torch-geometric's ``GCNConv`` (pinned 2.5.3 / 1.7.2; absent from
  /root/reference and from this image), as it is reached from the reference call sites
  src/finetune/gconv_adapter.py:40-41 (``ConvLayer(in_channels, out_channels,
  normalize=normalize)``), :73-78 (``.lin.weight`` [out,in] without bias, ``.bias`` [out])
  and :92 (``conv(x, edge_index)``):

      if normalize:   (add_self_loops defaults to ``normalize``)
          drop every edge with row == col, append one (i, i) loop per node, weight 1
          deg  = scatter_add(ones, col)            # in-degree at the TARGET, loop included
          dis  = deg.pow(-0.5); dis[dis == inf] = 0
          w    = dis[row] * 1 * dis[col]
      h   = x @ lin.weight.T
      msg = h.index_select(-2, row) (* w[:, None])  # flow = source_to_target
      out = zeros_like(h).scatter_add_(-2, col, msg) + bias

  ATen's CPU ``index_add_``/``scatter_add_`` sum sequentially in edge order, so the
  oracle is deterministic.  ``node_dim = -2`` makes [N, F] and [1, N, F] inputs equivalent
  (the NodeFormer host passes [1, N, H]: src/models/transductive/nodeformer.py:391,407).

Every function works in the dtype of ``x`` (fp32 for parity, fp64 as a tie-breaker).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn


# ----------------------------------------------------------------------------------
# torch_geometric.utils.add_remaining_self_loops / torch_geometric.nn.conv.gcn_conv.gcn_norm
# ----------------------------------------------------------------------------------
def add_remaining_self_loops(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """Drop existing self loops, then append exactly one (i, i) per node (fill value 1).

    With the unit edge weights the reference path uses, PyG's "keep the weight of an
    existing loop" rule degenerates to weight 1, so only the index part matters here.
    Duplicate (multi-)edges are kept.
    """
    row, col = edge_index[0], edge_index[1]
    mask = row != col
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index[:, mask], loop.unsqueeze(0).repeat(2, 1)], dim=1)


def gcn_norm(edge_index: torch.Tensor, num_nodes: int, dtype: torch.dtype = torch.float32
             ) -> Tuple[torch.Tensor, torch.Tensor]:
    """``gcn_norm(edge_index, None, N, improved=False, add_self_loops=True,
    flow='source_to_target', dtype)`` -> (edge_index', edge_weight')."""
    ei = add_remaining_self_loops(edge_index, num_nodes)
    w = torch.ones(ei.size(1), dtype=dtype, device=ei.device)
    row, col = ei[0], ei[1]
    deg = torch.zeros(num_nodes, dtype=dtype, device=ei.device).scatter_add_(0, col, w)
    dis = deg.pow_(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    w = dis[row] * w * dis[col]
    return ei, w


def glorot_(t: torch.Tensor) -> torch.Tensor:
    """torch_geometric.nn.inits.glorot: U(-a, a), a = sqrt(6 / (fan_in + fan_out))."""
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        return t.uniform_(-a, a)


class _Lin(nn.Module):
    """PyG ``Linear(in, out, bias=False, weight_initializer='glorot')``: weight [out, in]."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        glorot_(self.weight)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x @ self.weight.t()


class GCNConvRef(nn.Module):
    """Restated ``torch_geometric.nn.GCNConv(in_channels, out_channels, normalize=...)``."""

    def __init__(self, in_channels: int, out_channels: int, normalize: bool = True):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.normalize = normalize
        self.lin = _Lin(in_channels, out_channels)
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        if self.normalize:
            ei, w = gcn_norm(edge_index, x.size(-2), x.dtype)
        else:
            ei, w = edge_index, None
        h = self.lin(x)
        msg = h.index_select(-2, ei[0])
        if w is not None:
            msg = w.view(-1, 1) * msg
        out = torch.zeros_like(h).index_add_(-2, ei[1], msg)
        return out + self.bias


class _Unsupported(nn.Module):
    """'sage' / 'gat' are selectable in the reference (gconv_adapter.py:32-35) but its
    constructor cannot complete for them: SAGEConv has no ``normalize=`` semantics matching
    and neither class has ``.lin.weight`` (:73).  The oracle mirrors that as an error."""

    def __init__(self, *a, **k):
        super().__init__()
        raise AttributeError("conv has no attribute 'lin' (reference ctor fails at gconv_adapter.py:73)")


class GConvAdapterRef(nn.Module):
    """/root/reference/src/finetune/gconv_adapter.py:5-108 over ``GCNConvRef``."""

    def __init__(self, hidden_size: int, bottleneck_size: int,
                 conv_type: str = 'gcn', non_linearity: str = 'relu',
                 normalization: str = 'none', learnable_scalar: bool = False,
                 skip_connection: bool = True, normalize: bool = True):
        super().__init__()
        if conv_type == 'gcn':                                        # :30-37
            conv = GCNConvRef
        elif conv_type in ('sage', 'gat'):
            conv = _Unsupported
        else:
            raise ValueError("Invalid conv_type. Supported types: 'gcn', 'sage', 'gat'.")
        self.conv_down = conv(hidden_size, bottleneck_size, normalize=normalize)   # :40
        self.conv_up = conv(bottleneck_size, hidden_size, normalize=normalize)     # :41
        if non_linearity == 'relu':                                   # :44-51
            self.act_fn = nn.ReLU()
        elif non_linearity == 'silu':
            self.act_fn = nn.SiLU()
        elif non_linearity == 'none':
            self.act_fn = nn.Identity()
        else:
            raise ValueError("Invalid non_linearity. Supported types: 'relu', 'silu', 'none'.")
        if normalization == 'batch_norm':                             # :54-61
            self.normalization = nn.BatchNorm1d(hidden_size)
        elif normalization == 'layer_norm':
            self.normalization = nn.LayerNorm(hidden_size)
        elif normalization == 'none':
            self.normalization = None
        else:
            raise ValueError("Invalid normalization. Supported types: 'batch_norm', 'layer_norm', 'none'.")
        self.scalar = nn.Parameter(torch.ones(1)) if learnable_scalar else None   # :64-67
        self.skip_connection = skip_connection                        # :70
        # Test aid (not in the reference): when set to a bool tensor [N, r], ReLU's own decision
        # (h > 0) is replaced by this mask, so two fp32 implementations that round a pre-activation
        # to opposite sides of 0 can still be compared exactly (SURVEY.md section 7 "ReLU mask flips").
        self.relu_mask_override = None
        torch.nn.init.normal_(self.conv_down.lin.weight, mean=0.0, std=1e-5)      # :73-78
        torch.nn.init.zeros_(self.conv_down.bias)
        torch.nn.init.normal_(self.conv_up.lin.weight, mean=0.0, std=1e-5)
        torch.nn.init.zeros_(self.conv_up.bias)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor,
                edge_attr: Optional[torch.Tensor] = None) -> torch.Tensor:
        h = self.conv_down(x, edge_index)
        if self.relu_mask_override is not None and isinstance(self.act_fn, nn.ReLU):
            self.last_preact = h.detach()
            z = h * self.relu_mask_override.to(h.dtype).reshape(h.shape)
        else:
            self.last_preact = h.detach()
            z = self.act_fn(h)
        out = self.conv_up(z, edge_index)                              # :92
        if self.skip_connection:                                      # :94-95
            out += x
        if isinstance(self.normalization, nn.BatchNorm1d):            # :98-102
            if out.size(0) > 1:
                out = self.normalization(out)
        elif isinstance(self.normalization, nn.LayerNorm):
            out = self.normalization(out)
        if self.scalar is not None:                                   # :105-106
            out = out * self.scalar
        return out


# ----------------------------------------------------------------------------------
# The two normalisation formulas that DO live in the reference repo (known answers)
# ----------------------------------------------------------------------------------
def molecular_gcn_norm(edge_index: torch.Tensor, num_nodes: int, dtype=torch.float32) -> torch.Tensor:
    """/root/reference/src/layers/inductive/gcn_conv.py:36-56 (``MolecularGCNConv.norm``):
    degree by ROW over the given edge list, pow(-0.5), inf->0, dis[row]*w*dis[col]."""
    w = torch.ones(edge_index.size(1), dtype=dtype)
    row, col = edge_index[0], edge_index[1]
    deg = torch.zeros(num_nodes, dtype=dtype).scatter_add_(0, row, w)
    dis = deg.pow(-0.5)
    dis[dis == float('inf')] = 0
    return dis[row] * w * dis[col]


def dense_set_diag_normalize(edge_index: torch.Tensor, num_nodes: int, dtype=torch.float32) -> torch.Tensor:
    """/root/reference/src/dataset/transductive/data_utils.py:175-183 (``normalize``) on a
    dense stand-in for the coalesced SparseTensor: set_diag(1) -> row sums -> D^-1/2 A D^-1/2."""
    a = torch.zeros(num_nodes, num_nodes, dtype=dtype)
    a[edge_index[0], edge_index[1]] = 1
    a.fill_diagonal_(1)
    deg = a.sum(dim=1)
    dis = deg.pow(-0.5)
    dis[dis == float('inf')] = 0
    return dis.view(-1, 1) * a * dis.view(1, -1)


# ----------------------------------------------------------------------------------
# Convenience: one fwd+bwd through the oracle with explicit tensors
# ----------------------------------------------------------------------------------
def adapter_fwd_bwd(x, edge_index, params: dict, g_out, *, non_linearity='relu', normalization='none',
                    learnable_scalar=True, skip_connection=True, normalize=True, dtype=torch.float32):
    """Run GConvAdapterRef forward + backward on CPU.

    ``params`` maps state_dict keys -> tensors.  Returns (y, grads) where grads maps
    'x' and every parameter key to its gradient (all detached, in ``dtype``).
    """
    d = x.size(-1)
    r = params['conv_down.lin.weight'].size(0)
    m = GConvAdapterRef(d, r, non_linearity=non_linearity, normalization=normalization,
                        learnable_scalar=learnable_scalar, skip_connection=skip_connection,
                        normalize=normalize).to(dtype)
    with torch.no_grad():
        sd = m.state_dict()
        for k, v in params.items():
            sd[k].copy_(v.to(dtype))
    xx = x.detach().to(dtype).clone().requires_grad_(True)
    y = m(xx, edge_index)
    y.backward(g_out.to(dtype))
    grads = {'x': xx.grad.detach()}
    for k, p in m.named_parameters():
        grads[k] = p.grad.detach() if p.grad is not None else torch.zeros_like(p)
    return y.detach(), grads
