"""Generate tests/golden/backbone/*.npz with the reference's OWN gcn_conv / add_conv_relational_bias.

Run in the build container (needs /root/reference):   python tests/golden/make_golden_backbone.py

/root/reference/src/models/transductive/{difformer,nodeformer}.py are imported unmodified over
``oracle.backbone_ref.install_backbone_shims()`` (torch_sparse / torch_geometric / hydra are not installed), run
forward + backward on CPU in fp32 on seeded graphs with one self loop per node (the transductive pipeline's
convention, scripts/finetune_transductive_learning.py:112-113); inputs, outputs and input gradients are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import backbone_ref  # noqa: E402
from gconv_adapter_b200.graphs.synthetic import symmetric_random_graph  # noqa: E402


def looped(n, e, seed):
    ei = symmetric_random_graph(n, e, seed=seed)
    return torch.cat([ei, torch.arange(n, dtype=torch.int64).repeat(2, 1)], dim=1)


def directed_looped(n, e, seed):
    """Directed random graph (each edge in ONE direction only) + one self loop per node: in- and out-degrees differ, which
    separates gcn_conv (in-degree at both ends) from add_conv_relational_bias (out-degree at the source end)."""
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (e,), generator=g)
    dst = (src + 1 + torch.randint(0, n - 1, (e,), generator=g)) % n
    return torch.cat([torch.stack([src, dst]), torch.arange(n, dtype=torch.int64).repeat(2, 1)], dim=1)


def main():
    gcn_conv, relbias = backbone_ref.load_reference_functions()
    out_dir = os.path.join(HERE, "backbone")
    os.makedirs(out_dir, exist_ok=True)
    cases = {"small": (300, 1800, 1, 32, 1), "two_heads": (600, 4000, 2, 32, 2), "wide": (200, 1200, 4, 64, 3),
             "directed": (400, 2400, 2, 32, 4), "directed_sparse": (500, 700, 1, 64, 5)}
    for name, (n, e, h, dd, seed) in cases.items():
        ei = directed_looped(n, e, seed) if name.startswith("directed") else looped(n, e, seed)
        gen = torch.Generator().manual_seed(100 + seed)
        x = torch.randn(n, h, dd, generator=gen)
        g_out = torch.randn(n, h, dd, generator=gen)
        b = torch.randn(h, generator=gen)
        xr = x.clone().requires_grad_(True)
        y = gcn_conv(xr, ei, None)
        y.backward(g_out)
        xb = x.unsqueeze(0).clone().requires_grad_(True)
        br = b.clone().requires_grad_(True)
        yb = relbias(xb, ei, br, "sigmoid")
        yb.backward(g_out.unsqueeze(0))
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), edge_index=ei.numpy(), num_nodes=np.int64(n), x=x.numpy(),
                            g_out=g_out.numpy(), b=b.numpy(), gcn_y=y.detach().numpy(), gcn_gx=xr.grad.numpy(),
                            rel_y=yb.detach().numpy(), rel_gx=xb.grad.numpy(), rel_gb=br.grad.numpy())
        print(name, "done", tuple(y.shape))


if __name__ == "__main__":
    main()
