"""Host-model harnesses that reproduce the reference's adapter insertion hooks.

They are NOT the reference's backbones (those are out of scope): the graph
convolutions are small plain-torch stand-ins.  What is reproduced exactly is the boundary the adapter
sees - registration in ``nn.ModuleDict`` named ``pre_adapters`` / ``post_adapters`` with keys "0".."L-1",
the per-layer call order, sequential vs parallel insertion, the freeze policy, the argument forms:

* ``MolecularGNNHost`` / ``MolecularGraphPredictionHost``:
  /root/reference/src/models/inductive/gnn.py:63-130 and :215-284 (5-layer GIN, BatchNorm only when
  ``h.size(0) > 1``, ``adapter(h, edge_index, edge_attr)``, JK 'last', mean pooling, linear head).
* ``TransductiveHost(kind='nodeformer')``: /root/reference/src/models/transductive/nodeformer.py:390-448
  (x is unsqueezed to ``[1, N, H]``, ``adapter(z, adjs[0])``, residual is added in place: ``z += layer_[i]``).
* ``TransductiveHost(kind='difformer')``: difformer.py:216-267, including its quirk that a sequential
  pre-adapter only replaces the residual link ``layer_[i]`` while the conv still consumes ``x``.

The adapter class is a constructor argument, so the same host runs with the CUDA ``GConvAdapter`` or with
the CPU oracle's ``GConvAdapterRef`` - that is how ``tests/test_gpu_hosts.py`` proves drop-in behaviour.
"""
from __future__ import annotations

from typing import Callable, List

import torch
import torch.nn as nn
import torch.nn.functional as F

NUM_BOND_TYPE, NUM_BOND_DIRECTION = 6, 3
NUM_ATOM_TYPE, NUM_CHIRALITY_TAG = 120, 3


class _GINStandIn(nn.Module):
    """GIN-style layer with bond embeddings and an added self loop (bond type 4), sum aggregation, MLP -
    the shape of src/layers/inductive/gin_conv.py:38-85 in plain torch."""

    def __init__(self, emb_dim: int):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(emb_dim, 2 * emb_dim), nn.ReLU(), nn.Linear(2 * emb_dim, emb_dim))
        self.edge_embedding_type = nn.Embedding(NUM_BOND_TYPE, emb_dim)
        self.edge_embedding_direction = nn.Embedding(NUM_BOND_DIRECTION, emb_dim)

    def forward(self, x, edge_index, edge_attr):
        n = x.size(0)
        loop = torch.arange(n, device=x.device)
        ei = torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1)
        loop_attr = torch.zeros(n, 2, device=edge_attr.device, dtype=edge_attr.dtype)
        loop_attr[:, 0] = 4
        ea = torch.cat([edge_attr, loop_attr], dim=0)
        emb = self.edge_embedding_type(ea[:, 0]) + self.edge_embedding_direction(ea[:, 1])
        msg = x.index_select(0, ei[0]) + emb
        agg = torch.zeros_like(x).index_add_(0, ei[1], msg)
        return self.mlp(agg)


class MolecularGNNHost(nn.Module):
    def __init__(self, num_layers: int, emb_dim: int, dropout_rate: float = 0.0):
        super().__init__()
        self.num_layers, self.emb_dim, self.dropout_rate = num_layers, emb_dim, dropout_rate
        self.gnn_layers = nn.ModuleList([_GINStandIn(emb_dim) for _ in range(num_layers)])
        self.batch_norm_layers = nn.ModuleList([nn.BatchNorm1d(emb_dim) for _ in range(num_layers)])
        self.pre_adapters = nn.ModuleDict()
        self.post_adapters = nn.ModuleDict()
        self.adapter_type = None

    def add_adapter(self, adapter_factory: Callable[..., nn.Module], position: str, type: str):
        """gnn.py:63-82; ``adapter_factory(hidden_size=...)`` plays ``hydra.utils.instantiate(cfg, hidden_size=...)``."""
        self.adapter_type = type
        target = self.pre_adapters if position == 'pre' else self.post_adapters if position == 'post' else None
        if target is not None:
            for layer in range(self.num_layers):
                target[str(layer)] = adapter_factory(hidden_size=self.emb_dim)

    def forward(self, x, edge_index, edge_attr):
        hidden_states = [x]
        for layer in range(self.num_layers):
            key = str(layer)
            pre_out = None
            if key in self.pre_adapters:
                pre_out = self.pre_adapters[key](hidden_states[layer], edge_index, edge_attr)
                if self.adapter_type == "sequential":
                    hidden_states[layer] = pre_out
            h = self.gnn_layers[layer](hidden_states[layer], edge_index, edge_attr)
            if h.size(0) > 1:
                h = self.batch_norm_layers[layer](h)
            post_out = None
            if key in self.post_adapters:
                post_out = self.post_adapters[key](h, edge_index, edge_attr)
                if self.adapter_type == "sequential":
                    h = post_out
            if self.adapter_type == "parallel":
                if pre_out is not None:
                    h = h + pre_out
                if post_out is not None:
                    h = h + post_out
            if layer == self.num_layers - 1:
                h = F.dropout(h, self.dropout_rate, training=self.training)
            else:
                h = F.dropout(F.relu(h), self.dropout_rate, training=self.training)
            hidden_states.append(h)
        return hidden_states[-1]


class MolecularGraphPredictionHost(nn.Module):
    def __init__(self, num_layers: int = 5, emb_dim: int = 300, num_tasks: int = 1):
        super().__init__()
        self.node_embedding_type = nn.Embedding(NUM_ATOM_TYPE, emb_dim)
        self.node_embedding_chirality = nn.Embedding(NUM_CHIRALITY_TAG, emb_dim)
        self.gnn = MolecularGNNHost(num_layers, emb_dim)
        self.graph_pred_linear = nn.Linear(emb_dim, num_tasks)

    def add_adapter(self, adapter_factory, positions: List[str], type: str = 'sequential'):
        for position in positions:
            self.gnn.add_adapter(adapter_factory, position, type)
        for name, module in self.named_modules():                # freeze policy, gnn.py:228-241
            for _, param in module.named_parameters(recurse=False):
                param.requires_grad = ('adapter' in name or 'graph_pred_linear' in name or
                                       isinstance(module, (nn.LayerNorm, nn.BatchNorm1d)))

    def forward(self, x, edge_index, edge_attr, batch):
        h = self.node_embedding_type(x[:, 0]) + self.node_embedding_chirality(x[:, 1])
        h = self.gnn(h, edge_index, edge_attr)
        num_graphs = int(batch.max().item()) + 1
        pooled = torch.zeros(num_graphs, h.size(1), device=h.device, dtype=h.dtype).index_add_(0, batch, h)
        counts = torch.zeros(num_graphs, device=h.device, dtype=h.dtype).index_add_(0, batch, torch.ones_like(batch, dtype=h.dtype))
        return self.graph_pred_linear(pooled / counts.clamp(min=1).unsqueeze(1))


class TransductiveHost(nn.Module):
    """Insertion loop of the two graph-transformer hosts around a stand-in conv (a Linear over the hidden
    width; the real attention layers are out of scope)."""

    def __init__(self, kind: str, in_channels: int, hidden_channels: int, out_channels: int, num_layers: int = 2,
                 alpha: float = 0.5):
        super().__init__()
        if kind not in ("nodeformer", "difformer"):
            raise ValueError(kind)
        self.kind, self.num_layers, self.hidden_channels, self.alpha = kind, num_layers, hidden_channels, alpha
        self.fcs = nn.ModuleList([nn.Linear(in_channels, hidden_channels), nn.Linear(hidden_channels, out_channels)])
        self.bns = nn.ModuleList([nn.LayerNorm(hidden_channels) for _ in range(num_layers + 1)])
        self.convs = nn.ModuleList([nn.Linear(hidden_channels, hidden_channels) for _ in range(num_layers)])
        self.pre_adapters = nn.ModuleDict()
        self.post_adapters = nn.ModuleDict()
        self.adapter_type = None

    def add_adapter(self, adapter_factory, positions, type):
        self.adapter_type = type
        for position in positions:
            target = self.pre_adapters if position == 'pre' else self.post_adapters if position == 'post' else None
            if target is not None:
                for layer in range(self.num_layers):
                    target[str(layer)] = adapter_factory(hidden_size=self.hidden_channels)
        for name, module in self.named_modules():                # nodeformer.py:377-388
            for _, param in module.named_parameters(recurse=False):
                param.requires_grad = ('adapter' in name or 'fc' in name or isinstance(module, (nn.LayerNorm, nn.BatchNorm1d)))

    def forward(self, x, adjs):
        nodeformer = self.kind == "nodeformer"
        if nodeformer:
            x = x.unsqueeze(0)                                    # [1, N, H]: nodeformer.py:391
        z = F.relu(self.bns[0](self.fcs[0](x)))
        layer_ = [z]
        x = z
        for i, conv in enumerate(self.convs):
            key = str(i)
            pre_out = None
            if key in self.pre_adapters:
                pre_out = self.pre_adapters[key](layer_[i], adjs[0])
                if self.adapter_type == "sequential":
                    layer_[i] = pre_out
            if nodeformer:
                z = conv(layer_[i])
                z += layer_[i]                                    # in place, nodeformer.py:417-418
                z = F.relu(self.bns[i + 1](z))
            else:
                z = conv(x)                                       # difformer.py:241: conv consumes x, not layer_[i]
                z = self.alpha * z + (1 - self.alpha) * layer_[i]
                z = self.bns[i + 1](z)
            post_out = None
            if key in self.post_adapters:
                post_out = self.post_adapters[key](z, adjs[0])
                if self.adapter_type == "sequential":
                    z = post_out
            if self.adapter_type == "parallel":
                if pre_out is not None:
                    z = z + pre_out
                if post_out is not None:
                    z = z + post_out
            layer_.append(z)
            x = z
        out = self.fcs[-1](z)
        return out.squeeze(0) if nodeformer else out
