"""Graph-structure handle (CSR by target + CSR by source + D^-1/2) built on the GPU by K0.

The reference recomputes ``gcn_norm`` inside both ``GCNConv`` calls of every adapter forward
(/root/reference/src/finetune/gconv_adapter.py:92, PyG ``cached=False``) - 2 x layers x
positions times per step on a graph that, in the transductive scripts, never changes
(/root/reference/scripts/finetune_transductive_learning.py:110-119).  Here the structure is
built once per ``edge_index`` and shared by both hops, the backward, every layer and
position (``GraphCache``); the inductive path rebuilds it once per batch with one launch.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Optional

import torch

from .. import _cabi


class GraphStructure:
    """Owns the device workspace (a torch uint8 tensor) and the host descriptor of one graph
    (or of one row block ``[row_begin, row_end)`` of it, for the partitioned path)."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int, normalize: bool = True,
                 row_begin: int = 0, row_end: Optional[int] = None, validate=True):
        if edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise ValueError("edge_index must have shape [2, E]")
        if edge_index.dtype != torch.int64:
            raise ValueError("edge_index must be int64 (torch.long), as in the reference")
        if not edge_index.is_cuda:
            raise RuntimeError("gconv_adapter_b200: edge_index must live on a CUDA device (no CPU path)")
        lib = _cabi.load()
        self.device = edge_index.device
        self.num_nodes = int(num_nodes)
        self.normalize = bool(normalize)
        self.row_begin = int(row_begin)
        self.row_end = self.num_nodes if row_end is None else int(row_end)
        self.num_rows = self.row_end - self.row_begin
        self.num_edges = int(edge_index.size(1))
        ei = edge_index if edge_index.is_contiguous() else edge_index.contiguous()
        self._edge_index = ei            # keeps src/dst alive until the build has run
        nbytes = lib.gca_graph_workspace_bytes(self.num_edges, self.num_nodes, self.row_begin, self.row_end)
        if nbytes == 0:
            raise ValueError("invalid graph dimensions")
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            src_ptr = ei.data_ptr()
            dst_ptr = src_ptr + 8 * self.num_edges
            _cabi.check(lib.gca_graph_build(src_ptr, dst_ptr, self.num_edges, self.num_nodes, self.row_begin,
                                            self.row_end, int(self.normalize), self.workspace.data_ptr(), nbytes,
                                            stream, C.byref(self._handle)), "gca_graph_build")
            self.nnz = self.nnz_t = None
            self._pending = None
            if validate == "lazy":
                # no host synchronisation: the flags travel to pinned memory behind the build and are looked at by poll()
                self._flags_host = torch.zeros(8, dtype=torch.int32).pin_memory()
                _cabi.check(lib.gca_graph_validate_async(self._handle, self._flags_host.data_ptr(), stream), "gca_graph_validate_async")
                self._pending = torch.cuda.Event()
                self._pending.record(torch.cuda.current_stream(self.device))
            elif validate:
                self.validate()

    # -- lifetime ------------------------------------------------------------------
    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _cabi.load().gca_graph_destroy(h)
            except Exception:
                pass
            self._handle = C.c_void_p()

    @property
    def handle(self) -> C.c_void_p:
        return self._handle

    def poll(self, wait: bool = False) -> None:
        """Deferred validation (``validate='lazy'``): once the build has finished on the device, raise if it saw node ids
        outside [0, N) - the reference fails inside ``index_select`` / ``scatter_add_`` - and learn the hub-row counts.
        Cheap when nothing is pending; ``wait=True`` blocks until the build is done."""
        ev = self._pending
        if ev is None:
            return
        if wait:
            ev.synchronize()
        elif not ev.query():
            return
        self._pending = None
        lib = _cabi.load()
        nnz, nnz_t = C.c_int64(), C.c_int64()
        st = lib.gca_graph_validate_finish(self._handle, self._flags_host.data_ptr(), C.byref(nnz), C.byref(nnz_t))
        self.__dict__.pop("_ws_sizes", None)            # hub scratch may have become unnecessary
        if st == _cabi.GCA_ERR_INDEX_RANGE:
            raise RuntimeError(f"gconv_adapter_b200: edge_index contains node ids outside [0, {self.num_nodes})")
        _cabi.check(st, "gca_graph_validate_finish")
        self.nnz, self.nnz_t = nnz.value, nnz_t.value

    def validate(self) -> None:
        """Synchronise and raise on out-of-range node ids (the reference would fail inside
        ``index_select`` / ``scatter_add_``)."""
        lib = _cabi.load()
        nnz, nnz_t = C.c_int64(), C.c_int64()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        st = lib.gca_graph_validate(self._handle, stream, C.byref(nnz), C.byref(nnz_t))
        if st == _cabi.GCA_ERR_INDEX_RANGE:
            raise RuntimeError(f"gconv_adapter_b200: edge_index contains node ids outside [0, {self.num_nodes})")
        _cabi.check(st, "gca_graph_validate")
        self.nnz, self.nnz_t = nnz.value, nnz_t.value

    # -- inspection (tests, partitioning) -------------------------------------------
    def _view(self) -> _cabi.GraphView:
        v = _cabi.GraphView()
        _cabi.check(_cabi.load().gca_graph_get_view(self._handle, C.byref(v)), "gca_graph_get_view")
        return v

    def _slice(self, ptr: int, count: int, dtype: torch.dtype) -> torch.Tensor:
        off = ptr - self.workspace.data_ptr()
        nbytes = count * torch.empty((), dtype=dtype).element_size()
        return self.workspace[off:off + nbytes].view(dtype)

    def arrays(self) -> dict:
        """Device views of rowptr / colidx / rowptr_t / colidx_t / dis (valid while self lives)."""
        if self.nnz is None:
            self.poll(wait=True)
        if self.nnz is None:
            self.validate()
        v = self._view()
        n = self.num_rows
        return {
            "rowptr": self._slice(v.rowptr, n + 1, torch.int32),
            "colidx": self._slice(v.colidx, self.nnz, torch.int32),
            "rowptr_t": self._slice(v.rowptr_t, n + 1, torch.int32),
            "colidx_t": self._slice(v.colidx_t, self.nnz_t, torch.int32),
            "dis": self._slice(v.dis, n, torch.float32),
        }

    def edge_coefficients(self) -> torch.Tensor:
        """fp32 ``gcn_norm`` edge weights in forward-CSR order (full-graph handles only)."""
        if self.nnz is None:
            self.poll(wait=True)
        if self.nnz is None:
            self.validate()
        out = torch.empty(self.nnz, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _cabi.check(_cabi.load().gca_graph_edge_coef(self._handle, out.data_ptr(), stream), "gca_graph_edge_coef")
        return out


class GraphCache:
    """LRU of GraphStructure keyed on the identity of ``edge_index``.

    An entry keeps a reference to the tensor it was built from, so its storage cannot be
    freed and re-used while the entry lives; ``_version`` catches in-place edits.
    """

    def __init__(self, capacity: int = 8):
        self.capacity = capacity
        self._entries: "OrderedDict[tuple, tuple]" = OrderedDict()
        self.hits = 0
        self.misses = 0

    @staticmethod
    def _key(edge_index: torch.Tensor, num_nodes: int, normalize: bool, row_begin: int, row_end: int):
        try:
            version = edge_index._version
        except RuntimeError:                 # inference tensors do not track versions
            return None
        return (edge_index.data_ptr(), version, tuple(edge_index.shape), tuple(edge_index.stride()),
                edge_index.device, int(num_nodes), bool(normalize), int(row_begin), int(row_end))

    def get(self, edge_index: torch.Tensor, num_nodes: int, normalize: bool = True,
            row_begin: int = 0, row_end: Optional[int] = None, validate: bool = True) -> GraphStructure:
        row_end = num_nodes if row_end is None else row_end
        key = self._key(edge_index, num_nodes, normalize, row_begin, row_end)
        if key is None:
            # edge_index was created under torch.inference_mode(): in-place edits cannot be detected, so the structure
            # is rebuilt on every call (build the graph outside inference_mode, or under no_grad, to get it cached)
            self.misses += 1
            return GraphStructure(edge_index, num_nodes, normalize, row_begin, row_end, validate=validate)
        hit = self._entries.get(key)
        if hit is not None:
            self._entries.move_to_end(key)
            self.hits += 1
            hit[0].poll()
            return hit[0]
        self.misses += 1
        g = GraphStructure(edge_index, num_nodes, normalize, row_begin, row_end, validate=validate)
        self._entries[key] = (g, edge_index)
        while len(self._entries) > self.capacity:
            self._entries.popitem(last=False)
        return g

    def invalidate(self, edge_index: Optional[torch.Tensor] = None) -> None:
        """Drop the entry built from ``edge_index`` (or everything).  Needed only after writes that do not bump the tensor's
        version counter (``edge_index.data.copy_()``, custom / DLPack kernels, a collective receiving into the buffer):
        the cache key is (data_ptr, _version, shape, ...), so such writes would otherwise return the old structure."""
        if edge_index is None:
            self._entries.clear()
            return
        for key in [k for k, (_, t) in self._entries.items() if t is edge_index or t.data_ptr() == edge_index.data_ptr()]:
            del self._entries[key]

    def clear(self) -> None:
        self._entries.clear()


GLOBAL_GRAPH_CACHE = GraphCache()
