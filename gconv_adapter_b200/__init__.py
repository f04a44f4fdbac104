"""gconv_adapter_b200 - B200-native (sm_100a) implementation of the GConv-Adapter hot path.

Drop-in for ``src.finetune.gconv_adapter.GConvAdapter`` of PanPapag/GConv-Adapter
(/root/reference/src/finetune/gconv_adapter.py): same constructor, state_dict keys and call
signature; the arithmetic runs in hand-written CUDA kernels behind the C ABI of include/gca.h.
"""
from .finetune.gconv_adapter import GConvAdapter
from .graphed import graphed_adapter
from .graphs.csr import GLOBAL_GRAPH_CACHE, GraphCache, GraphStructure

__all__ = ["GConvAdapter", "GraphStructure", "GraphCache", "GLOBAL_GRAPH_CACHE", "graphed_adapter", "install_reference_alias"]
__version__ = "0.1.0"


def install_reference_alias() -> None:
    """Make ``src.finetune.gconv_adapter.GConvAdapter`` (the ``_target_`` of the reference's
    configs/finetune/gconv_adapter.yaml:2) resolve to this implementation, without touching a
    checkout of the reference: registers alias modules in ``sys.modules`` when absent."""
    import sys
    import types

    from .finetune import gconv_adapter as impl

    for name in ("src", "src.finetune"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = []          # mark as package
            sys.modules[name] = mod
    sys.modules["src.finetune.gconv_adapter"] = impl
    setattr(sys.modules["src.finetune"], "gconv_adapter", impl)
    setattr(sys.modules["src"], "finetune", sys.modules["src.finetune"])
