"""GPU parity of the molecular backbone layers on the shared graph handle (gconv_adapter_b200.layers.inductive) against the
golden vectors of the reference's own MolecularGINConv / MolecularGCNConv and against the CPU oracle on a full molecule batch
of BASELINE.json configs[1] (hidden 300)."""
import glob
import os

import numpy as np
import pytest
import torch

from gconv_adapter_b200.graphs.csr import GLOBAL_GRAPH_CACHE
from gconv_adapter_b200.graphs.synthetic import molecule_batch
from gconv_adapter_b200.layers import inductive
from oracle import molecular_ref
from util import assert_close

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "backbone", "mol_*.npz")))
OURS = {"gin": inductive.MolecularGINConv, "gcn": inductive.MolecularGCNConv}
REFS = {"gin": molecular_ref.MolecularGINConvRef, "gcn": molecular_ref.MolecularGCNConvRef}


def load(path):
    z = np.load(path)
    return {k: torch.from_numpy(z[k]) if z[k].ndim else z[k].item() for k in z.files}


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
@pytest.mark.parametrize("tag", ["gin", "gcn"])
def test_golden_vectors_of_the_reference_classes(path, tag):
    g = load(path)
    emb = g["x"].shape[1]
    torch.manual_seed(int(g[tag + "_ctor_seed"]))
    layer = OURS[tag](emb)                                        # same constructor RNG stream as the reference class
    for k, v in g.items():
        if k.startswith(tag + "_param."):
            assert torch.equal(layer.state_dict()[k[len(tag) + 7:]], v), k
    layer = layer.cuda()
    x = g["x"].cuda().requires_grad_(True)
    y = layer(x, g["edge_index"].cuda(), g["edge_attr"].cuda())
    y.backward(g["g_out"].cuda())
    assert_close(y, g[tag + "_y"], f"{tag} y")
    assert_close(x.grad, g[tag + "_gx"], f"{tag} g_x")
    for k, p in layer.named_parameters():
        if f"{tag}_grad.{k}" in g:
            assert_close(p.grad, g[f"{tag}_grad.{k}"], f"{tag} grad {k}")


@pytest.mark.parametrize("tag", ["gin", "gcn"])
def test_full_molecule_batch_against_the_oracle_and_shared_handle(tag):
    ei, _, n = molecule_batch(batch_size=32, seed=11)
    gen = torch.Generator().manual_seed(3)
    ea = torch.stack([torch.randint(0, 4, (ei.size(1),), generator=gen), torch.randint(0, 3, (ei.size(1),), generator=gen)], 1)
    x, g_out = torch.randn(n, 300, generator=gen), torch.randn(n, 300, generator=gen)
    torch.manual_seed(5)
    ref = REFS[tag](300)
    torch.manual_seed(5)
    ours = OURS[tag](300).cuda()
    xr = x.clone().requires_grad_(True)
    yr = ref(xr, ei, ea)
    yr.backward(g_out)
    eic, eac = ei.cuda(), ea.cuda()
    misses = GLOBAL_GRAPH_CACHE.misses
    xo = x.cuda().requires_grad_(True)
    yo = ours(xo, eic, eac)
    yo.backward(g_out.cuda())
    ours(xo.detach(), eic, eac)                                   # a second layer on the same batch: no second build
    assert GLOBAL_GRAPH_CACHE.misses == misses + 1
    assert_close(yo, yr, f"{tag} y")
    assert_close(xo.grad, xr.grad, f"{tag} g_x")
    for (k, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
        assert_close(p.grad, q.grad, f"{tag} grad {k}")


def test_rejects_graphs_with_self_loops():
    ei = torch.tensor([[0, 1, 2, 2], [1, 0, 2, 0]], dtype=torch.int64).cuda()
    ea = torch.zeros(4, 2, dtype=torch.int64).cuda()
    with pytest.raises(ValueError, match="self loops"):
        inductive.MolecularGINConv(8).cuda()(torch.randn(3, 8, device="cuda"), ei, ea)
