"""Generate tests/golden/backbone/mol_*.npz with the reference's OWN MolecularGINConv / MolecularGCNConv.

Run in the build container (needs /root/reference):   python tests/golden/make_golden_molecular.py

/root/reference/src/layers/inductive/{gin_conv,gcn_conv}.py are imported unmodified over
``oracle.molecular_ref.install_message_passing_shim()`` (torch_geometric is not installed), constructed from a seed, run
forward + backward on CPU in fp32 on seeded molecule batches; inputs, parameters, outputs and all gradients are stored."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import molecular_ref  # noqa: E402
from gconv_adapter_b200.graphs.synthetic import molecule_batch  # noqa: E402


def main():
    gin_cls, gcn_cls = molecular_ref.load_reference_classes()
    out_dir = os.path.join(HERE, "backbone")
    os.makedirs(out_dir, exist_ok=True)
    # (batch size, emb_dim, seed, store parameters and all gradients?)  The larger cases rebuild the parameters from the
    # seed in the test (the constructors consume the RNG identically - mol_small pins that) and keep the fixture small.
    for name, (bs, emb, seed, full) in {"mol_small": (4, 32, 1, True), "mol_batch32": (32, 64, 2, False),
                                        "mol_d300": (3, 300, 3, False)}.items():
        ei, _, n = molecule_batch(batch_size=bs, seed=seed)
        g = torch.Generator().manual_seed(50 + seed)
        ea = torch.stack([torch.randint(0, 4, (ei.size(1),), generator=g), torch.randint(0, 3, (ei.size(1),), generator=g)], 1)
        x = torch.randn(n, emb, generator=g)
        g_out = torch.randn(n, emb, generator=g)
        store = {"edge_index": ei.numpy(), "edge_attr": ea.numpy(), "num_nodes": np.int64(n), "x": x.numpy(), "g_out": g_out.numpy()}
        for tag, cls in (("gin", gin_cls), ("gcn", gcn_cls)):
            torch.manual_seed(7 + seed)
            layer = cls(emb)
            xr = x.clone().requires_grad_(True)
            y = layer(xr, ei, ea)
            y.backward(g_out)
            store[tag + "_y"] = y.detach().numpy()
            store[tag + "_gx"] = xr.grad.numpy()
            store[tag + "_ctor_seed"] = np.int64(7 + seed)
            for k, p in layer.named_parameters():
                if full:
                    store[f"{tag}_param.{k}"] = p.detach().numpy()
                if full or "edge_embedding" in k or k.endswith("bias"):
                    store[f"{tag}_grad.{k}"] = p.grad.numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **store)
        print(name, "done", n, ei.size(1))


if __name__ == "__main__":
    main()
