"""Synthetic graphs shaped like the BASELINE.json configs.

Workload generation only (numpy on the host): no dataset or network is available, so
every config is a seeded random graph with the node / edge counts of the named dataset.
``edge_index`` follows the reference's convention: int64 ``[2, E]``, row 0 = source,
row 1 = target (PyG ``flow='source_to_target'``).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


@dataclass(frozen=True)
class Shape:
    name: str
    num_nodes: int
    num_edges: int
    hidden: int
    rank: int
    self_loops_present: bool = False   # transductive callers pre-insert one loop per node


# BASELINE.json configs[0..4]
SHAPES = {
    "cora": Shape("cora", 2_708, 10_556, 64, 8),
    "molecules": Shape("molecules", 0, 0, 300, 16),                     # built by molecule_batch()
    "pubmed": Shape("pubmed", 19_717, 88_648, 64, 16, self_loops_present=True),
    "arxiv": Shape("arxiv", 169_343, 1_166_243, 256, 16),
    "products": Shape("products", 2_449_029, 61_859_140, 256, 32),
}


def symmetric_random_graph(num_nodes: int, num_edges: int, seed: int = 0, power_law: bool = False,
                           add_self_loops: bool = False) -> torch.Tensor:
    """``num_edges // 2`` unordered pairs (no self pairs, duplicates allowed) emitted in both
    directions, plus one extra directed edge when ``num_edges`` is odd.  ``power_law`` draws
    endpoints from a Zipf-like distribution (hub nodes) instead of uniformly."""
    rng = np.random.default_rng(seed)
    half = num_edges // 2

    def draw(n):
        if not power_law:
            return rng.integers(0, num_nodes, size=n, dtype=np.int64)
        # inverse-CDF of p(k) ~ (k+1)^-0.8 over node ranks, then a fixed permutation of ids
        u = rng.random(n)
        k = np.floor(num_nodes * u ** 5.0).astype(np.int64)
        return np.minimum(k, num_nodes - 1)

    a = draw(half)
    b = draw(half)
    same = a == b
    b[same] = (b[same] + 1 + rng.integers(0, num_nodes - 1, size=int(same.sum()))) % num_nodes
    if power_law:
        perm = np.random.default_rng(seed + 1).permutation(num_nodes)
        a, b = perm[a], perm[b]
    src = np.concatenate([a, b])
    dst = np.concatenate([b, a])
    if num_edges % 2:
        s = rng.integers(0, num_nodes, dtype=np.int64)
        t = (s + 1 + rng.integers(0, num_nodes - 1, dtype=np.int64)) % num_nodes
        src = np.append(src, s)
        dst = np.append(dst, t)
    if add_self_loops:
        loop = np.arange(num_nodes, dtype=np.int64)
        src = np.concatenate([src, loop])
        dst = np.concatenate([dst, loop])
    return torch.from_numpy(np.stack([src, dst]))


def molecule_batch(batch_size: int = 32, seed: int = 0, min_nodes: int = 20, max_nodes: int = 30):
    """A PyG-style block-diagonal batch of random "molecules": a random tree per graph plus
    one or two ring-closing bonds, every bond in both directions, no self loops.
    Returns (edge_index [2,E] int64, batch [N] int64, num_nodes)."""
    rng = np.random.default_rng(seed)
    srcs, dsts, batch = [], [], []
    base = 0
    for g in range(batch_size):
        n = int(rng.integers(min_nodes, max_nodes + 1))
        parent = np.array([rng.integers(max(0, i - 3), i) for i in range(1, n)], dtype=np.int64)
        child = np.arange(1, n, dtype=np.int64)
        a, b = [parent], [child]
        for _ in range(int(rng.integers(1, 3))):
            u = int(rng.integers(0, n - 5))
            a.append(np.array([u], dtype=np.int64))
            b.append(np.array([u + 5], dtype=np.int64))
        a, b = np.concatenate(a) + base, np.concatenate(b) + base
        srcs += [a, b]
        dsts += [b, a]
        batch.append(np.full(n, g, dtype=np.int64))
        base += n
    ei = np.stack([np.concatenate(srcs), np.concatenate(dsts)])
    return torch.from_numpy(ei), torch.from_numpy(np.concatenate(batch)), base


def make_graph(name: str, seed: int = 0, power_law: bool = False, scale: float = 1.0) -> tuple[torch.Tensor, int]:
    """(edge_index, num_nodes) for a named shape; ``scale`` shrinks N and E together."""
    if name == "molecules":
        ei, _, n = molecule_batch(seed=seed)
        return ei, n
    s = SHAPES[name]
    n = max(8, int(round(s.num_nodes * scale)))
    e = max(8, int(round(s.num_edges * scale)))
    return symmetric_random_graph(n, e, seed=seed, power_law=power_law,
                                  add_self_loops=s.self_loops_present), n


def make_inputs(num_nodes: int, hidden: int, rank: int, seed: int = 0, weight_std: float = 0.05):
    """Seeded fp32 inputs / parameters / upstream gradient."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(num_nodes, hidden, generator=g)
    g_out = torch.randn(num_nodes, hidden, generator=g)
    params = {
        "scalar": torch.tensor([1.3]),
        "conv_down.bias": torch.randn(rank, generator=g) * weight_std,
        "conv_down.lin.weight": torch.randn(rank, hidden, generator=g) * weight_std,
        "conv_up.bias": torch.randn(hidden, generator=g) * weight_std,
        "conv_up.lin.weight": torch.randn(hidden, rank, generator=g) * weight_std,
    }
    return x, g_out, params


def input_rows(lo: int, hi: int, hidden: int, seed: int, device, chunk: int = 65536):
    """Rows [lo, hi) of the seeded (X, gY) pair of a workload, generated on ``device`` chunk by chunk: chunk c always comes
    from its own generator seed, so every rank of a partitioned run - and rank 0's full-size reference - draws identical
    rows without anybody materialising more than it owns."""
    xs, gs = [], []
    for c in range(lo // chunk, (max(hi, lo + 1) - 1) // chunk + 1):
        g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + c)
        x = torch.randn(chunk, hidden, generator=g, device=device)
        gy = torch.randn(chunk, hidden, generator=g, device=device)
        a, b = max(lo, c * chunk) - c * chunk, min(hi, (c + 1) * chunk) - c * chunk
        xs.append(x[a:b])
        gs.append(gy[a:b])
    return torch.cat(xs).contiguous(), torch.cat(gs).contiguous()
