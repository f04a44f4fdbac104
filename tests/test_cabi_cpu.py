"""CPU tests of the boundary: libgca.so loads without a GPU, exports every symbol include/gca.h
declares, and the pure size/validation entry points behave (no compute calls here)."""
import ctypes
import os
import re

import pytest

from gconv_adapter_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gca.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gca_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_list_the_same_symbols():
    assert _declared_symbols() == sorted(_cabi.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(raw, name), name
    assert lib.gca_abi_version() == _cabi.ABI_VERSION == 2
    assert lib.gca_status_string(0) == b"ok"
    assert b"outside" in lib.gca_status_string(-4)


def test_size_queries_and_argument_validation():
    lib = _cabi.load()
    assert lib.gca_graph_workspace_bytes(1000, 100, 0, 100) > 2 * 4 * 1100
    assert lib.gca_graph_workspace_bytes(10, 100, 50, 40) == 0          # row_end < row_begin
    assert lib.gca_graph_workspace_bytes(-1, 100, 0, 100) == 0
    assert lib.gca_bwd_scratch_bytes(256, 16) % 256 == 0
    assert lib.gca_backward_workspace_bytes(None, 256, 16) == 0 and lib.gca_forward_workspace_bytes(None, 256, 16) == 0
    assert lib.gca_hub_scratch_bytes(None) == 0
    assert lib.gca_shape_is_fast(256, 16) == 1 and lib.gca_shape_is_fast(300, 16) == 1
    assert lib.gca_shape_is_fast(30, 16) == 0 and lib.gca_shape_is_fast(256, 5) == 0
    h = ctypes.c_void_p()
    # null workspace / bad ranges are rejected before any CUDA call
    assert lib.gca_graph_build(None, None, 0, 10, 0, 10, 1, None, 0, None, ctypes.byref(h)) == -2
    assert lib.gca_graph_build(None, None, 5, 10, 0, 10, 1, None, 0, None, ctypes.byref(h)) == -1
    assert lib.gca_graph_build(None, None, 0, 10, 4, 2, 1, None, 0, None, ctypes.byref(h)) == -1
    assert lib.gca_forward(None, None, 0, None, None, None, None, None, 1, 1, None, None, None, None, None, 0, 8, 8, None) == -1


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libgca.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _cabi.load()


def test_library_name_carries_the_hash_of_its_sources():
    """build() is content-addressed: the binary that is loaded was built from exactly the sources in the tree."""
    from gconv_adapter_b200 import build
    assert os.path.basename(_cabi.LIB_PATH) == f"libgca.{build.source_hash()}.so"
    assert os.path.isfile(_cabi.LIB_PATH)
