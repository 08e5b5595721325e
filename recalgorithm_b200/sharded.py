"""Row-sharded embedding tables over the GPUs of one NVSwitch box.

Host side only: partition arithmetic, buffer ownership and the peer-mapping rendezvous (torch.distributed is the
plumbing: object all-gather of CUDA-IPC handles, barriers).  The data path is two kernels of libctr_b200.so:
``ctr_embed_fm2_fwd_sharded`` (rows pulled over NVLink inside the gather) and ``ctr_sharded_grad_push`` (gradient
rows pushed into their owner's receive buffer); no NCCL collective moves embedding data.

Sharding rule: global row ``gr = field_row_offset[f] + id`` lives on rank ``gr % G`` at local row ``gr // G``.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional

import torch

from . import _lib, ops


# ------------------------------------------------------------------ partition arithmetic (pure; also used by CPU tests)
def owner_of(global_rows: torch.Tensor, G: int) -> torch.Tensor:
    return global_rows % G


def local_row_of(global_rows: torch.Tensor, G: int) -> torch.Tensor:
    return global_rows // G


def shard_rows(num_rows_total: int, G: int) -> int:
    """Rows held by every rank (the last ranks' tails may be unused)."""
    return (num_rows_total + G - 1) // G


def full_to_shard(full: torch.Tensor, rank: int, G: int) -> torch.Tensor:
    """Rank's shard of a full (V, D) table, zero padded to shard_rows."""
    n = shard_rows(full.shape[0], G)
    out = torch.zeros((n, full.shape[1]), dtype=full.dtype, device=full.device)
    part = full[rank::G]
    out[: part.shape[0]] = part
    return out


def shards_to_full(shards: List[torch.Tensor], num_rows_total: int) -> torch.Tensor:
    G = len(shards)
    full = torch.zeros((num_rows_total, shards[0].shape[1]), dtype=shards[0].dtype, device=shards[0].device)
    for r, s in enumerate(shards):
        n = full[r::G].shape[0]
        full[r::G] = s[:n]
    return full


def receive_capacity(batch: int, fields: int, G: int, slack: float = 1.25) -> int:
    """Slots per (source, owner) pair in the gradient receive buffers."""
    return int((batch * fields / G) * slack) + 1024


# ------------------------------------------------------------------ peer-mappable buffers (CUDA IPC through the C ABI)
class _Raw:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def _view(ptr: int, shape, dtype, device) -> torch.Tensor:
    ts = {torch.float32: "<f4", torch.int64: "<i8", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_Raw(ptr, shape, ts), device=device)


class PeerBuffer:
    """A cudaMalloc'ed buffer (ctr_peer_alloc) viewed as a torch tensor, with its 64-byte CUDA-IPC handle."""

    def __init__(self, shape, dtype, device):
        self.shape, self.dtype, self.device = tuple(shape), dtype, device
        nbytes = max(1, int(torch.empty((), dtype=dtype).element_size())) * max(1, int(torch.Size(shape).numel()))
        p = ctypes.c_void_p()
        _lib.check(_lib.lib().ctr_peer_alloc(nbytes, ctypes.byref(p)))
        self.ptr = int(p.value)
        self.tensor = _view(self.ptr, shape, dtype, device)
        h = ctypes.create_string_buffer(64)
        _lib.check(_lib.lib().ctr_ipc_export(ctypes.c_void_p(self.ptr), h))
        self.handle = bytes(h.raw)

    def open_peer(self, handle: bytes) -> torch.Tensor:
        q = ctypes.c_void_p()
        _lib.check(_lib.lib().ctr_ipc_import(handle, ctypes.byref(q)))
        return _view(int(q.value), self.shape, self.dtype, self.device)


def _ptr_array(ptrs: List[int]):
    return (ctypes.c_void_p * len(ptrs))(*ptrs)


class SymmBuffer:
    """Same role as PeerBuffer, backed by torch's symmetric memory (CUDA VMM allocations with 2 MB pages, handles passed
    as file descriptors): measured necessary for the table shard -- a legacy-IPC mapping of a 32 GB shard collapsed to
    7 GB/s under random 128-byte reads, while <= 5 GB shards reached 553 GB/s (peer TLB reach)."""

    def __init__(self, shape, dtype, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.shape, self.dtype, self.device = tuple(shape), dtype, device
        self.tensor = symm_mem.empty(*shape, dtype=dtype, device=device)
        self._hdl = symm_mem.rendezvous(self.tensor, group)
        self.peer_ptrs = [int(p) for p in self._hdl.buffer_ptrs]


class ShardedEmbeddingTables:
    """F per-field tables, concatenated and row-sharded over the ranks of ``group`` (one process per GPU)."""

    def __init__(self, rows_per_field, dim: int, batch_per_rank: int, group=None, device=None, init: Optional[str] = "normal",
                 seed: int = 1234, slack: float = 1.25):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.G = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.G & (self.G - 1) or self.G > 8:
            raise ValueError("world size must be a power of two <= 8")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        rows = torch.as_tensor(rows_per_field, dtype=torch.int64)
        self.num_fields, self.dim = int(rows.numel()), int(dim)
        off = torch.zeros(self.num_fields + 1, dtype=torch.int64)
        off[1:] = torch.cumsum(rows, 0)
        self.num_rows = int(off[-1])
        self.field_row_offset = off.to(self.device)
        self.local_rows = shard_rows(self.num_rows, self.G)
        self._bufs = {}
        self._symm_w = None
        if os.environ.get("CTR_PEER_BACKEND", "symm") == "symm":
            grp = group if group is not None else dist.group.WORLD
            self._symm_w = SymmBuffer((self.local_rows, self.dim), torch.float32, self.device, grp)
            self.weight = self._symm_w.tensor
        else:
            self._bufs["w"] = PeerBuffer((self.local_rows, self.dim), torch.float32, self.device)
            self.weight = self._bufs["w"].tensor
        if init == "normal":
            g = torch.Generator(device=self.device).manual_seed(seed + self.rank)
            self.weight.normal_(0, self.dim ** -0.5, generator=g)
        self.capacity = receive_capacity(batch_per_rank, self.num_fields, self.G, slack)
        self._bufs["v"] = PeerBuffer((self.G, self.capacity, self.dim), torch.float32, self.device)
        self._bufs["r"] = PeerBuffer((self.G, self.capacity), torch.int64, self.device)
        self._bufs["c"] = PeerBuffer((self.G,), torch.int64, self.device)
        self.recv_vals, self.recv_rows, self.recv_counts = (self._bufs[k].tensor for k in ("v", "r", "c"))
        self.recv_counts.zero_()
        self.counters = torch.zeros((self.G,), dtype=torch.int64, device=self.device)
        self.overflow = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self._rendezvous()

    def _rendezvous(self):
        """Exchange the CUDA-IPC handles of the shard and the receive buffers and map every peer's buffers into this
        device (the import enables NVLink peer access); the mappings live as long as the object."""
        mine = {k: buf.handle for k, buf in self._bufs.items()}
        gathered = [None] * self.G
        self.dist.all_gather_object(gathered, mine, group=self.group)
        self._peer = []
        for r in range(self.G):
            if r == self.rank:
                self._peer.append({k: buf.tensor for k, buf in self._bufs.items()})
            else:
                self._peer.append({k: self._bufs[k].open_peer(gathered[r][k]) for k in self._bufs})
        if self._symm_w is not None:
            self._w_ptrs = _ptr_array(self._symm_w.peer_ptrs)
        else:
            self._w_ptrs = _ptr_array([p["w"].data_ptr() for p in self._peer])
        self._v_ptrs = _ptr_array([p["v"].data_ptr() for p in self._peer])
        self._r_ptrs = _ptr_array([p["r"].data_ptr() for p in self._peer])
        self._c_ptrs_dev = torch.tensor([p["c"].data_ptr() for p in self._peer], dtype=torch.int64, device=self.device)
        torch.cuda.synchronize()
        self.dist.barrier(group=self.group)

    # ---- forward: pull
    def lookup_fm2(self, ids: torch.Tensor, want_tile=True, want_fm2=True, tile=None, fm2=None):
        B, F = ids.shape
        D = self.dim
        if want_tile and tile is None:
            tile = torch.empty((B, F, D), dtype=torch.float32, device=self.device)
        if want_fm2 and fm2 is None:
            fm2 = torch.empty((B, 1), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().ctr_embed_fm2_fwd_sharded(self._w_ptrs, self.G, self.field_row_offset.data_ptr(), ids.data_ptr(),
                                                        B, F, D, ops._ptr(tile), ops._ptr(fm2), ops._stream()))
        return tile, fm2

    # ---- backward: push
    def push_grads(self, ids: torch.Tensor, row_grads: torch.Tensor, barrier: bool = True):
        """Deliver (local_row, grad) of every valid (b,f) to its owner.  After the call (with barrier=True) this rank's
        ``recv_rows/recv_vals/recv_counts`` hold what all ranks sent to it."""
        B, F, D = row_grads.shape
        L = _lib.lib()
        _lib.check(L.ctr_sharded_grad_push(row_grads.data_ptr(), self.field_row_offset.data_ptr(), ids.data_ptr(), B, F, D,
                                           self.G, self.rank, self._v_ptrs, self._r_ptrs, self.capacity,
                                           self.counters.data_ptr(), self.overflow.data_ptr(), ops._stream()))
        _lib.check(L.ctr_sharded_publish_counts(self.counters.data_ptr(), self._c_ptrs_dev.data_ptr(), self.G, self.rank,
                                                ops._stream()))
        if barrier:
            torch.cuda.current_stream().synchronize()
            self.dist.barrier(group=self.group)

    def received_to_dense(self) -> torch.Tensor:
        """Densify what this rank received into a (local_rows, D) gradient shard (tests / dense consumers)."""
        if int(self.overflow.item()):
            raise RuntimeError("gradient receive buffer overflow: raise `slack`")
        dense = torch.zeros((self.local_rows, self.dim), dtype=torch.float32, device=self.device)
        L = _lib.lib()
        for src in range(self.G):
            _lib.check(L.ctr_rows_scatter_add(dense.data_ptr(), self.local_rows, self.dim, self.recv_rows[src].data_ptr(),
                                              self.recv_vals[src].data_ptr(), self.recv_counts[src:].data_ptr(), self.capacity,
                                              ops._stream()))
        return dense
