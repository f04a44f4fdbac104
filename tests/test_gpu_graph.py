"""GPU parity of K0 (graph-structure build) against oracle/csr_ref.py: integer arrays bit-exact,
dis bit-exact, gcn_norm edge coefficients within 1 ULP (north_star item 1)."""
import numpy as np
import pytest
import torch

from gconv_adapter_b200 import GraphCache, GraphStructure
from gconv_adapter_b200.graphs.synthetic import make_graph, molecule_batch, symmetric_random_graph
from oracle import csr_ref

pytestmark = pytest.mark.gpu


def _check(ei: torch.Tensor, n: int, normalize: bool, lo: int = 0, hi=None):
    g = GraphStructure(ei.cuda(), n, normalize, lo, hi)
    ref = csr_ref.build(ei.numpy(), n, normalize, lo, hi)
    arr = {k: v.cpu().numpy() for k, v in g.arrays().items()}
    assert g.nnz == ref["nnz"] and g.nnz_t == ref["nnz_t"]
    for k in ("rowptr", "colidx", "rowptr_t", "colidx_t"):
        assert arr[k].dtype == np.int32
        assert np.array_equal(arr[k], ref[k]), k
    hi = n if hi is None else hi
    assert np.array_equal(arr["dis"], ref["dis"][lo:hi]), "dis must equal torch's deg.pow(-0.5) bit for bit"
    return g, ref


TINY = {
    "path4": (4, [[0, 1, 1, 2, 2, 3], [1, 0, 2, 1, 3, 2]]),
    "isolated": (5, [[0, 1, 3, 4], [1, 0, 4, 3]]),
    "duplicate_edge": (4, [[0, 0, 1, 1, 2, 3], [1, 1, 0, 0, 3, 2]]),
    "existing_self_loops": (4, [[0, 1, 1, 2, 2, 2, 3], [1, 0, 1, 2, 2, 3, 2]]),
    "directed": (5, [[0, 1, 2, 3, 0], [1, 2, 3, 4, 4]]),
    "single_node": (1, [[0], [0]]),
}


@pytest.mark.parametrize("name", sorted(TINY))
@pytest.mark.parametrize("normalize", [True, False])
def test_tiny_graphs(name, normalize):
    n, e = TINY[name]
    _check(torch.tensor(e, dtype=torch.int64), n, normalize)


@pytest.mark.parametrize("normalize", [True, False])
def test_empty_edge_list(normalize):
    _check(torch.zeros(2, 0, dtype=torch.int64), 7, normalize)


@pytest.mark.parametrize("normalize", [True, False])
@pytest.mark.parametrize("n,e", [(300, 1200), (2708, 10556), (20000, 90000), (169343, 1166243)])
def test_random_graphs_small_and_multi_launch_paths(n, e, normalize):
    """(300,1200) and (2708,10556) take the single-CTA fused build, the others the multi-launch one."""
    _check(symmetric_random_graph(n, e, seed=n), n, normalize)


def test_molecule_batch_and_pubmed_shaped():
    ei, _, n = molecule_batch(seed=1)
    _check(ei, n, True)
    ei, n = make_graph("pubmed", seed=2)        # one self loop per node already present
    g, ref = _check(ei, n, True)
    assert ref["nnz"] == 88648 + n


def test_hub_rows_take_every_sort_tier():
    """star graph: the hub row has 20000 neighbours (global-memory sort), plus rows of 33..4096."""
    n = 20001
    leaves = torch.arange(1, n, dtype=torch.int64)
    hub = torch.zeros(n - 1, dtype=torch.int64)
    src = torch.cat([hub, leaves]); dst = torch.cat([leaves, hub])
    mid = torch.randint(1, n, (3000,), generator=torch.Generator().manual_seed(0))     # node 7 gets ~3000 in-edges
    src = torch.cat([src, mid]); dst = torch.cat([dst, torch.full_like(mid, 7)])
    few = torch.randint(1, n, (50,), generator=torch.Generator().manual_seed(1))        # node 9: ~50
    src = torch.cat([src, few]); dst = torch.cat([dst, torch.full_like(few, 9)])
    perm = torch.randperm(src.numel(), generator=torch.Generator().manual_seed(2))
    ei = torch.stack([src[perm], dst[perm]])
    _check(ei, n, True)
    pl, npl = make_graph("arxiv", seed=3, power_law=True, scale=0.25)
    _check(pl, npl, True)


def test_edge_coefficients_within_one_ulp_of_gcn_norm():
    n = 5000
    ei = symmetric_random_graph(n, 40000, seed=5)
    g, ref = _check(ei, n, True)
    coef = g.edge_coefficients().cpu().numpy()
    ei2, w = csr_ref.edge_coefficients(ei.numpy(), n)
    order = np.lexsort((ei2[0], ei2[1]))             # by target, then source == CSR order
    assert np.array_equal(ei2[0][order].astype(np.int32), ref["colidx"])
    ulp = csr_ref.ulp_diff(coef, w[order])
    assert ulp.max() <= 1, f"max ULP distance {ulp.max()}"
    assert ulp.max() == 0, "dis[i]*dis[j] is one IEEE multiply on both sides: expected exact"


def test_row_partition_blocks_concatenate_to_full_graph():
    n = 3001
    ei = symmetric_random_graph(n, 20000, seed=8)
    cuts = [0, 700, 1500, 1501, n]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        _check(ei, n, True, lo, hi)
    # multi-launch path with a partition
    n2, e2 = 60000, 400000
    ei2 = symmetric_random_graph(n2, e2, seed=9)
    _check(ei2, n2, True, 20000, 45000)


def test_out_of_range_ids_raise():
    ei = torch.tensor([[0, 1, 9], [1, 0, 2]], dtype=torch.int64)
    with pytest.raises(RuntimeError, match="outside"):
        GraphStructure(ei.cuda(), 5, True)
    ei = torch.tensor([[0, 1, -1], [1, 0, 2]], dtype=torch.int64)
    with pytest.raises(RuntimeError, match="outside"):
        GraphStructure(ei.cuda(), 5, True)
    with pytest.raises(RuntimeError):
        GraphStructure(torch.zeros(2, 3, dtype=torch.int64), 5, True)      # CPU tensor


def test_build_is_deterministic_and_cache_tracks_identity():
    n = 50000
    ei = symmetric_random_graph(n, 300000, seed=4).cuda()
    a = GraphStructure(ei, n, True).arrays()
    b = GraphStructure(ei, n, True).arrays()
    for k in a:
        assert torch.equal(a[k], b[k]), k
    cache = GraphCache(capacity=2)
    g1 = cache.get(ei, n, True)
    assert cache.get(ei, n, True) is g1 and cache.hits == 1
    assert cache.get(ei, n, False) is not g1                 # normalize is part of the key
    ei[0, 0] = ei[0, 0]                                      # in-place write bumps _version -> rebuild
    assert cache.get(ei, n, True) is not g1
