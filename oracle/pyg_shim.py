"""Stand-in ``torch_geometric`` so the reference's OWN glue class can be executed here.

TEST INFRASTRUCTURE ONLY.  torch-geometric is not installed in this image, so
/root/reference/src/finetune/gconv_adapter.py:3 (``from torch_geometric.nn import GCNConv,
SAGEConv, GATConv``) cannot import.  ``install()`` registers a minimal module whose
``GCNConv`` is the restated ``oracle.pyg_restated.GCNConvRef``; ``load_reference_adapter()``
then imports the *unmodified* reference file from /root/reference and returns its
``GConvAdapter`` class.  This pins the 86 lines of reference glue (constructor, init,
forward order, skip / norm / scalar) with the reference's own code; only the PyG
arithmetic underneath stays a restatement (hence "parity unpinned" at that boundary).

Used only by tests/golden/make_golden.py and by CPU tests that skip when /root/reference is
absent (it does not exist on the GPU box).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

from . import pyg_restated

REFERENCE_ROOT = os.environ.get("GCA_REFERENCE_ROOT", "/root/reference")


def install() -> None:
    if "torch_geometric" in sys.modules and not getattr(sys.modules["torch_geometric"], "_gca_shim", False):
        return  # a real PyG is present: never shadow it
    tg = types.ModuleType("torch_geometric")
    tg._gca_shim = True
    tgnn = types.ModuleType("torch_geometric.nn")

    class SAGEConv(pyg_restated._Unsupported):
        pass

    class GATConv(pyg_restated._Unsupported):
        pass

    tgnn.GCNConv = pyg_restated.GCNConvRef
    tgnn.SAGEConv = SAGEConv
    tgnn.GATConv = GATConv
    tg.nn = tgnn
    sys.modules["torch_geometric"] = tg
    sys.modules["torch_geometric.nn"] = tgnn


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "finetune", "gconv_adapter.py"))


def load_reference_adapter():
    """Import the reference's GConvAdapter class, unmodified, over the shim."""
    install()
    path = os.path.join(REFERENCE_ROOT, "src", "finetune", "gconv_adapter.py")
    spec = importlib.util.spec_from_file_location("_reference_gconv_adapter", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.GConvAdapter
