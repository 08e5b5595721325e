"""Row-sharded embedding tables over the GPUs of one NVSwitch box.

Host side only: partition arithmetic, buffer ownership and the peer-mapping rendezvous (torch.distributed is the
plumbing: object all-gather of CUDA-IPC handles, barriers).  The data path is two kernels of libctr_b200.so:
``ctr_embed_fm2_fwd_sharded`` (rows pulled over NVLink inside the gather) and ``ctr_sharded_grad_push`` (gradient
rows pushed into their owner's receive buffer); no NCCL collective moves embedding data.

Sharding rule: global row ``gr = field_row_offset[f] + id`` lives on rank ``gr % G`` at local row ``gr // G``.
"""
from __future__ import annotations

import ctypes
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


def exchange_fds(fd: int, group=None) -> List[Optional[int]]:
    """Every rank of `group` contributes one file descriptor; returns this process's duplicates of all ranks' descriptors
    (None at the own rank).  Descriptors travel over AF_UNIX sockets (SCM_RIGHTS): each rank listens on an abstract-namespace
    socket, serves its fd from a helper thread and collects the peers'.  (pidfd_getfd would be shorter but is refused between
    sibling processes under the usual ptrace restrictions.)"""
    import os
    import socket
    import threading
    import uuid
    import torch.distributed as dist
    G, rank = dist.get_world_size(group), dist.get_rank(group)
    name = "\0ctr_fd_" + uuid.uuid4().hex                   # abstract namespace: nothing to unlink
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(name)
    srv.listen(G)

    def serve():
        for _ in range(G - 1):
            conn, _ = srv.accept()
            with conn:
                socket.send_fds(conn, [b"f"], [fd])
                conn.recv(1)                                  # the peer has its duplicate before this side moves on

    th = threading.Thread(target=serve, daemon=True)
    th.start()
    names = [None] * G
    dist.all_gather_object(names, name, group=group)
    out: List[Optional[int]] = [None] * G
    for r in range(G):
        if r == rank:
            continue
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        c.connect(names[r])
        with c:
            _, fds, _, _ = socket.recv_fds(c, 1, 1)
            out[r] = int(fds[0])
            c.send(b"k")
    th.join()
    srv.close()
    return out


class VmmBuffer:
    """Same role as SymmBuffer with the allocation made by libctr_b200 itself (ctr_vmm_alloc: cuMemCreate + cuMemMap with an
    explicit size/address alignment) and the peers' mappings imported from file descriptors passed over AF_UNIX sockets."""

    def __init__(self, shape, dtype, device, group, align: int = 0):
        import os
        import torch.distributed as dist
        self.shape, self.dtype, self.device = tuple(shape), dtype, device
        nbytes = int(torch.Size(shape).numel()) * torch.empty((), dtype=dtype).element_size()
        L = _lib.lib()
        p, fd, mapped = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int64()
        _lib.check(L.ctr_vmm_alloc(max(nbytes, 1), align, ctypes.byref(p), ctypes.byref(fd), ctypes.byref(mapped)))
        self.ptr, self._fd, self.mapped_bytes, self.align = int(p.value), int(fd.value), int(mapped.value), align
        self.tensor = _view(self.ptr, shape, dtype, device)
        G, rank = dist.get_world_size(group), dist.get_rank(group)
        sizes = [None] * G
        dist.all_gather_object(sizes, self.mapped_bytes, group=group)
        fds = exchange_fds(self._fd, group)
        self.peer_ptrs, self._imported = [], []
        for r in range(G):
            if r == rank:
                self.peer_ptrs.append(self.ptr)
                continue
            q = ctypes.c_void_p()
            _lib.check(L.ctr_vmm_import(fds[r], sizes[r], align, ctypes.byref(q)))
            os.close(fds[r])
            self.peer_ptrs.append(int(q.value))
            self._imported.append(int(q.value))
        torch.cuda.synchronize()
        dist.barrier(group=group)                        # every peer has imported before anybody may close its fd
        os.close(self._fd)


class ShardedEmbeddingTables:
    """F per-field tables, concatenated and row-sharded over the ranks of ``group`` (one process per GPU).

    One training step on every rank:  ``plan(ids)`` (queue slots, any time before the backward) ->
    ``lookup_fm2(ids)`` (rows pulled over NVLink) -> ... -> ``bwd_push(tile, d_tile, d_fm2, plan)`` (gradient rows
    stored straight into their owners' queues) -> ``finish_push()`` (stream sync + barrier + overflow check) -> the owner
    consumes ``recv_rows / recv_vals / recv_counts`` (``optim.ShardedTableAdam``).  ``lookup_fm2_autograd`` wraps the
    first three for an autograd graph."""

    def __init__(self, rows_per_field, dim: int, batch_per_rank: int, group=None, device=None, init: Optional[str] = "normal",
                 seed: int = 1234, slack: float = 1.25, shard_backend: str = "symm", vmm_align: int = 0):
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
        grp = group if group is not None else dist.group.WORLD
        # the shard lives in CUDA VMM symmetric memory (2 MB pages): a legacy-IPC mapping of a 32 GB shard collapsed to
        # 7 GB/s under random peer reads (peer-TLB reach); the small receive queues keep plain IPC buffers
        if shard_backend == "vmm":
            self._symm_w = VmmBuffer((self.local_rows, self.dim), torch.float32, self.device, grp, align=vmm_align)
        elif shard_backend == "symm":
            self._symm_w = SymmBuffer((self.local_rows, self.dim), torch.float32, self.device, grp)
        else:
            raise ValueError("shard_backend must be 'symm' (torch symmetric memory) or 'vmm' (ctr_vmm_alloc)")
        self.weight = self._symm_w.tensor
        if init == "normal":
            g = torch.Generator(device=self.device).manual_seed(seed + self.rank)
            self.weight.normal_(0, self.dim ** -0.5, generator=g)
        self.capacity = receive_capacity(batch_per_rank, self.num_fields, self.G, slack)
        self._bufs["v"] = PeerBuffer((self.G, self.capacity, self.dim), torch.float32, self.device)
        self._bufs["r"] = PeerBuffer((self.G, self.capacity), torch.int64, self.device)
        self._bufs["c"] = PeerBuffer((self.G,), torch.int64, self.device)
        self.recv_vals, self.recv_rows, self.recv_counts = (self._bufs[k].tensor for k in ("v", "r", "c"))
        self.recv_counts.zero_()
        self.counters = torch.zeros((9,), dtype=torch.int64, device=self.device)      # entries per owner + a ticket
        self.overflow = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self._overflow_host = torch.zeros((1,), dtype=torch.int32).pin_memory()
        self._side, self._plan_ev = None, None
        self._rendezvous()

    def _rendezvous(self):
        """Exchange the CUDA-IPC handles of the receive buffers and map every peer's buffers into this device (the import
        enables NVLink peer access); the mappings live as long as the object."""
        mine = {k: buf.handle for k, buf in self._bufs.items()}
        gathered = [None] * self.G
        self.dist.all_gather_object(gathered, mine, group=self.group)
        self._peer = []
        for r in range(self.G):
            if r == self.rank:
                self._peer.append({k: buf.tensor for k, buf in self._bufs.items()})
            else:
                self._peer.append({k: self._bufs[k].open_peer(gathered[r][k]) for k in self._bufs})
        self._w_ptrs = _ptr_array(self._symm_w.peer_ptrs)
        self._v_ptrs = _ptr_array([p["v"].data_ptr() for p in self._peer])
        self._r_ptrs = _ptr_array([p["r"].data_ptr() for p in self._peer])
        self._c_ptrs = _ptr_array([p["c"].data_ptr() for p in self._peer])
        torch.cuda.synchronize()
        self.dist.barrier(group=self.group)

    # ---- forward: pull
    def lookup_fm2(self, ids: torch.Tensor, want_tile=True, want_fm2=True, tile=None, fm2=None, ids64_out=None):
        """ids (B,F) int64, or int32 (then ``ids64_out`` (B,F) int64, if given, receives the widened copy)."""
        B, F = ids.shape
        D = self.dim
        if want_tile and tile is None:
            tile = torch.empty((B, F, D), dtype=torch.float32, device=self.device)
        if want_fm2 and fm2 is None:
            fm2 = torch.empty((B, 1), dtype=torch.float32, device=self.device)
        L = _lib.lib()
        if ids.dtype == torch.int32:
            ops._chk(ids, torch.int32, "ids"); ops._chk(ids64_out, torch.int64, "ids64_out", (B, F))
            _lib.check(L.ctr_embed_fm2_fwd_sharded_ids32(self._w_ptrs, self.G, self.field_row_offset.data_ptr(), ops._ptr(ids),
                                                         B, F, D, ops._ptr(tile), ops._ptr(fm2), ops._ptr(ids64_out), ops._stream()))
        else:
            ops._chk(ids, torch.int64, "ids")
            _lib.check(L.ctr_embed_fm2_fwd_sharded(self._w_ptrs, self.G, self.field_row_offset.data_ptr(), ops._ptr(ids), B, F, D,
                                                   ops._ptr(tile), ops._ptr(fm2), ops._stream()))
        return tile, fm2

    # ---- backward: plan + push
    def plan(self, ids: torch.Tensor, plan: Optional[torch.Tensor] = None, side_stream: bool = False) -> torch.Tensor:
        """Queue slot of every (b,f) at its owner (int32 (B,F); -1 = invalid id / dropped) + the queues' row indices and
        counts.  The owners' queues are overwritten: the previous step's entries must have been consumed.
        side_stream=True: the plan only depends on the ids, so it runs on a second stream beside the forward that is launched
        right after it (the forward is NVLink-bound and leaves issue slots free); the push calls wait for it."""
        B, F = ids.shape
        i32 = ids.dtype == torch.int32
        ops._chk(ids, torch.int32 if i32 else torch.int64, "ids")
        if plan is None:
            plan = torch.empty((B, F), dtype=torch.int32, device=self.device)
        ops._chk(plan, torch.int32, "plan", (B, F))
        main = torch.cuda.current_stream(self.device)
        if side_stream:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            self._side.wait_stream(main)
            ids.record_stream(self._side); plan.record_stream(self._side)
        with torch.cuda.stream(self._side if side_stream else main):
            _lib.check(_lib.lib().ctr_sharded_plan(self.field_row_offset.data_ptr(), ops._ptr(ids), int(i32), B, F, self.G, self.rank,
                                                   self._r_ptrs, self._c_ptrs, self.capacity, self.counters.data_ptr(),
                                                   self.overflow.data_ptr(), ops._ptr(plan), ops._stream()))
            self._overflow_host.copy_(self.overflow, non_blocking=True)      # checked in finish_push(), after the stream sync
            if side_stream:
                self._plan_ev = torch.cuda.Event()
                self._plan_ev.record(self._side)
        return plan

    def _wait_plan(self):
        if self._plan_ev is not None:
            torch.cuda.current_stream(self.device).wait_event(self._plan_ev)
            self._plan_ev = None

    def bwd_push(self, tile, d_tile, d_fm2, plan, row_grads=None):
        """Lookup backward fused with the exchange; ``row_grads`` (optional) also keeps the values locally."""
        B, F, D = tile.shape
        ops._chk(tile, torch.float32, "tile"); ops._chk(d_tile, torch.float32, "d_tile", (B, F, D))
        if d_fm2 is not None:
            d_fm2 = d_fm2.reshape(B)
        ops._chk(d_fm2, torch.float32, "d_fm2", (B,)); ops._chk(plan, torch.int32, "plan", (B, F))
        ops._chk(row_grads, torch.float32, "row_grads", (B, F, D))
        self._wait_plan()
        _lib.check(_lib.lib().ctr_embed_fm2_bwd_push(ops._ptr(tile), ops._ptr(d_tile), ops._ptr(d_fm2), ops._ptr(plan), B, F, D,
                                                     self.G, self.rank, self._v_ptrs, self.capacity, ops._ptr(row_grads),
                                                     ops._stream()))

    def lookup_fm2_linear(self, ids: torch.Tensor, wlin: torch.Tensor, ids64_out=None):
        """Forward with the fused dense(1) head: (tile, fm2 (B,1), lin (B,1) = tile.reshape(B, F*D) @ wlin)."""
        B, F = ids.shape
        D = self.dim
        i32 = ids.dtype == torch.int32
        ops._chk(ids, torch.int32 if i32 else torch.int64, "ids"); ops._chk(ids64_out, torch.int64, "ids64_out", (B, F))
        wlin = wlin.reshape(F * D)
        ops._chk(wlin, torch.float32, "wlin", (F * D,))
        tile = torch.empty((B, F, D), dtype=torch.float32, device=self.device)
        fm2 = torch.empty((B, 1), dtype=torch.float32, device=self.device)
        lin = torch.empty((B, 1), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().ctr_embed_fm2_lin_fwd_sharded(self._w_ptrs, self.G, self.field_row_offset.data_ptr(), ops._ptr(ids), int(i32),
                                                            B, F, D, ops._ptr(wlin), ops._ptr(tile), ops._ptr(fm2), ops._ptr(lin),
                                                            ops._ptr(ids64_out), ops._stream()))
        return tile, fm2, lin

    def lin_bwd_push(self, tile, wlin, d_fm2, d_lin, plan, row_grads=None):
        """Backward of lookup_fm2_linear fused with the exchange; returns d_wlin (F*D,)."""
        B, F, D = tile.shape
        wlin = wlin.reshape(F * D)
        ops._chk(tile, torch.float32, "tile"); ops._chk(wlin, torch.float32, "wlin", (F * D,)); ops._chk(plan, torch.int32, "plan", (B, F))
        if d_fm2 is not None:
            d_fm2 = d_fm2.reshape(B)
        if d_lin is not None:
            d_lin = d_lin.reshape(B)
        ops._chk(d_fm2, torch.float32, "d_fm2", (B,)); ops._chk(d_lin, torch.float32, "d_lin", (B,))
        ops._chk(row_grads, torch.float32, "row_grads", (B, F, D))
        d_wlin = torch.empty((F * D,), dtype=torch.float32, device=self.device)
        self._wait_plan()
        _lib.check(_lib.lib().ctr_embed_fm2_lin_bwd_push(ops._ptr(tile), ops._ptr(wlin), ops._ptr(d_fm2), ops._ptr(d_lin), ops._ptr(plan),
                                                         B, F, D, self.G, self.rank, self._v_ptrs, self.capacity, ops._ptr(row_grads),
                                                         ops._ptr(d_wlin), ops._stream()))
        return d_wlin

    def finish_push(self):
        """Stream sync + cross-rank barrier: afterwards this rank's ``recv_rows/recv_vals/recv_counts`` hold what all
        ranks sent to it.  Raises if a queue overflowed (entries were dropped): raise ``slack``."""
        self._wait_plan()
        torch.cuda.current_stream().synchronize()
        over = torch.tensor([int(self._overflow_host[0])], dtype=torch.int32, device=self.device)
        self.dist.all_reduce(over, op=self.dist.ReduceOp.MAX, group=self.group)     # doubles as the barrier
        if int(over.item()):
            raise RuntimeError(f"gradient receive queue overflow (capacity {self.capacity} entries per source and owner): "
                               "entries were dropped; construct ShardedEmbeddingTables with a larger `slack`")

    def push_grads(self, ids: torch.Tensor, row_grads: torch.Tensor, barrier: bool = True, plan: Optional[torch.Tensor] = None):
        """Deliver (local_row, grad) of every valid (b,f) to its owner, for row gradients that already exist."""
        B, F, D = row_grads.shape
        ops._chk(row_grads, torch.float32, "row_grads")
        if plan is None:
            plan = self.plan(ids)
        self._wait_plan()
        _lib.check(_lib.lib().ctr_sharded_grad_push(ops._ptr(row_grads), ops._ptr(plan), B, F, D, self.G, self.rank, self._v_ptrs,
                                                    self.capacity, ops._stream()))
        if barrier:
            self.finish_push()

    def received_to_dense(self) -> torch.Tensor:
        """Densify what this rank received into a (local_rows, D) gradient shard (tests / dense consumers)."""
        dense = torch.zeros((self.local_rows, self.dim), dtype=torch.float32, device=self.device)
        L = _lib.lib()
        for src in range(self.G):
            _lib.check(L.ctr_rows_scatter_add(dense.data_ptr(), self.local_rows, self.dim, self.recv_rows[src].data_ptr(),
                                              self.recv_vals[src].data_ptr(), self.recv_counts[src:].data_ptr(), self.capacity,
                                              ops._stream()))
        return dense


class _ShardedLookupFM2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, tables: ShardedEmbeddingTables, ids: torch.Tensor):
        B, F = ids.shape
        ctx.plan = tables.plan(ids, side_stream=True)          # beside the forward (int32 or int64 ids alike)
        tile, fm2 = tables.lookup_fm2(ids)
        ctx.tables = tables
        ctx.save_for_backward(tile)
        return tile, fm2

    @staticmethod
    def backward(ctx, d_tile, d_fm2):
        (tile,) = ctx.saved_tensors
        ctx.tables.bwd_push(tile, None if d_tile is None else d_tile.contiguous(), None if d_fm2 is None else d_fm2.contiguous(),
                            ctx.plan)
        return None, None, None


class _ShardedLookupFM2Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wlin, tables: ShardedEmbeddingTables, ids: torch.Tensor):
        B, F = ids.shape
        wl = wlin.contiguous()
        ctx.plan = tables.plan(ids, side_stream=True)          # beside the forward (int32 or int64 ids alike)
        tile, fm2, lin = tables.lookup_fm2_linear(ids, wl)
        ctx.tables = tables
        ctx.save_for_backward(tile, wl)
        return fm2, lin

    @staticmethod
    def backward(ctx, d_fm2, d_lin):
        tile, wl = ctx.saved_tensors
        d_wlin = ctx.tables.lin_bwd_push(tile, wl, None if d_fm2 is None else d_fm2.contiguous(),
                                         None if d_lin is None else d_lin.contiguous(), ctx.plan)
        return d_wlin.reshape(wl.shape), None, None


def lookup_fm2_linear_autograd(tables: ShardedEmbeddingTables, ids: torch.Tensor, wlin: torch.Tensor):
    """Row-sharded lookup + FM2 + fused dense(1) head: (fm2 (B,1), lin (B,1)); the backward pushes the gradient rows to their
    owners and returns d_wlin (call ``tables.finish_push()`` before the owners' optimizer step)."""
    return _ShardedLookupFM2Linear.apply(wlin, tables, ids)


def lookup_fm2_autograd(tables: ShardedEmbeddingTables, ids: torch.Tensor, anchor: Optional[torch.Tensor] = None):
    """(B,F) ids (int64 or int32) -> (tile (B,F,D), fm2 (B,1)) inside an autograd graph; the backward pushes the gradient
    rows to their owners (call ``tables.finish_push()`` before the owners' optimizer step)."""
    if anchor is None:
        if not hasattr(tables, "_anchor"):
            tables._anchor = torch.zeros((), device=tables.device, requires_grad=True)
        anchor = tables._anchor
    return _ShardedLookupFM2.apply(anchor, tables, ids)
