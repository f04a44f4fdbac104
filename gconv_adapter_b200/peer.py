"""Peer-memory group for the row-partitioned adapter: one zero-initialised device arena per rank, mapped into every
other process of the node with CUDA IPC (include/gca.h: gca_peer_*), so that kernels can store rows straight into their
peers' gathered buffers over NVLink and synchronise with a one-warp barrier kernel - no NCCL call inside a step.

``torch.distributed`` is used ONCE, at construction, to exchange the 64-byte IPC handles."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from . import _cabi


class _DeviceSpan:
    """Zero-copy view of foreign device memory for ``torch.as_tensor`` (CUDA array interface)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class PeerGroup:
    """Arena layout (identical on every rank, so an offset names the same object everywhere):
    ``[flags: world x uint32 | seq: uint32 | slots: world x slot_floats | user regions ...]``."""

    def __init__(self, region_bytes: list[int], slot_floats: int, group=None, device: Optional[torch.device] = None):
        self.lib = _cabi.load()
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _cabi.MAX_PEERS:
            raise RuntimeError(f"peer-memory groups hold at most {_cabi.MAX_PEERS} GPUs (one node)")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.slot_floats = int(slot_floats)
        off = 0
        self.off_flags = off; off += 256
        self.off_seq = off; off += 256
        self.off_slots = off; off += _align(4 * self.world * self.slot_floats)
        self.off_regions = []
        for nb in region_bytes:
            self.off_regions.append(off)
            off += _align(int(nb))
        self.nbytes = off
        # Every rank goes through the same two collectives (handle exchange, status exchange) whatever fails locally,
        # so a rank whose allocation / export / mapping fails cannot leave the others stuck in one of them.
        with torch.cuda.device(self.device):
            own = C.c_void_p()
            handle = C.create_string_buffer(_cabi.IPC_HANDLE_BYTES)
            self._own, self.ptrs, self._opened, self._closed = own, [], [], False
            err: Optional[Exception] = None
            try:
                _cabi.check(self.lib.gca_peer_alloc(self.nbytes, C.byref(own)), "gca_peer_alloc")
                _cabi.check(self.lib.gca_peer_export(own, handle), "gca_peer_export")
            except Exception as ex:      # noqa: BLE001 - reported after the collectives below
                err = ex
            handles = [None] * self.world
            dist.all_gather_object(handles, None if err else handle.raw, group=group)
            if err is None and any(h is None for h in handles):
                err = RuntimeError("a peer could not allocate / export its arena")
            if err is None:
                try:
                    for h in range(self.world):
                        if h == self.rank:
                            self.ptrs.append(own.value)
                            continue
                        p = C.c_void_p()
                        _cabi.check(self.lib.gca_peer_open(handles[h], C.byref(p)),
                                    "gca_peer_open (CUDA IPC between the GPUs of one node)")
                        self._opened.append(p)
                        self.ptrs.append(p.value)
                except Exception as ex:  # noqa: BLE001
                    err = ex
            oks = [None] * self.world    # every arena is allocated, zeroed and mapped before anyone signals through it
            dist.all_gather_object(oks, err is None, group=group)
            if err is None and not all(oks):
                err = RuntimeError(f"peer arena set-up failed on rank(s) {[k for k, o in enumerate(oks) if not o]}")
            if err is not None:
                self.close()
                raise err
        self.sync = _cabi.PeerSync()
        self.sync.world, self.sync.rank = self.world, self.rank
        for h in range(self.world):
            self.sync.flags[h] = self.ptrs[h] + self.off_flags
        self.sync.seq = self.ptrs[self.rank] + self.off_seq
        self.slots = _cabi.Push()
        self.slots.count = self.world
        for h in range(self.world):
            self.slots.dst[h] = self.ptrs[h] + self.off_slots
        self._local = torch.as_tensor(_DeviceSpan(own.value, self.nbytes), device=self.device)
        self._closed = False

    # ---- views / destinations ----
    def region(self, i: int, nbytes: int) -> torch.Tensor:
        """uint8 view of this rank's region i."""
        o = self.off_regions[i]
        return self._local[o:o + nbytes]

    def push_to_peers(self, i: int, byte_offset: int) -> _cabi.Push:
        """Destinations = the same bytes of region i in every OTHER rank's arena."""
        p = _cabi.Push()
        k = 0
        for h in range(self.world):
            if h != self.rank:
                p.dst[k] = self.ptrs[h] + self.off_regions[i] + byte_offset
                k += 1
        p.count = k
        return p

    # ---- collectives as kernels on the current stream ----
    def barrier(self) -> None:
        st = torch.cuda.current_stream(self.device).cuda_stream
        _cabi.check(self.lib.gca_peer_barrier(C.byref(self.sync), st), "gca_peer_barrier")

    def allreduce_(self, flat: torch.Tensor) -> torch.Tensor:
        """In-place sum over the ranks of a small contiguous fp32 vector (<= slot_floats), in rank order."""
        n = flat.numel()
        if n > self.slot_floats or flat.dtype != torch.float32 or not flat.is_contiguous():
            raise RuntimeError("PeerGroup.allreduce_: a contiguous fp32 vector of at most slot_floats elements")
        st = torch.cuda.current_stream(self.device).cuda_stream
        _cabi.check(self.lib.gca_peer_allreduce(C.byref(self.sync), C.byref(self.slots), flat.data_ptr(), flat.data_ptr(), n, st),
                    "gca_peer_allreduce")
        return flat

    # ---- lifetime ----
    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        try:
            torch.cuda.synchronize(self.device)
            if dist.is_initialized():
                dist.barrier(self.group)       # nobody unmaps while a peer may still store into the arena
        except Exception:
            pass
        self._local = None
        for p in self._opened:
            self.lib.gca_peer_close(p)
        self._opened = []
        self.lib.gca_peer_free(self._own)

    def __del__(self):
        # best effort without collectives (interpreter shutdown: the driver reclaims the mappings)
        if not getattr(self, "_closed", True):
            self._closed = True
