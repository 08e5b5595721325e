"""CPU tests (gloo, world_size 2) of the host-side logic of the row-sharded path: partition arithmetic, shard
construction, receive-buffer sizing, and the reference exchange the CUDA push path is checked against."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from recalgorithm_b200 import sharded as S


def test_partition_roundtrip():
    full = torch.arange(23 * 4, dtype=torch.float32).reshape(23, 4)
    for G in (1, 2, 4, 8):
        shards = [S.full_to_shard(full, r, G) for r in range(G)]
        assert all(s.shape[0] == S.shard_rows(23, G) for s in shards)
        assert torch.equal(S.shards_to_full(shards, 23), full)
        gr = torch.arange(23)
        own, loc = S.owner_of(gr, G), S.local_row_of(gr, G)
        for r in range(23):
            assert torch.equal(shards[int(own[r])][int(loc[r])], full[r])
    assert S.receive_capacity(65536, 40, 8) >= 65536 * 40 // 8


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import sharded_ref as R
        g = torch.Generator().manual_seed(7)
        rows = torch.tensor([13, 1, 40, 7])
        off = torch.zeros(5, dtype=torch.int64); off[1:] = torch.cumsum(rows, 0)
        V, D, B = int(off[-1]), 8, 37
        full = torch.randn((V, D), generator=g)                               # same on every rank
        shard = S.full_to_shard(full, rank, world)
        gi = torch.Generator().manual_seed(100 + rank)                         # rank-specific batch
        ids = torch.stack([torch.randint(-1, int(r) + 2, (B,), generator=gi) for r in rows], 1)
        tile = R.lookup_reference(shard, off, ids)
        valid = (ids >= 0) & (ids < rows[None, :])
        want = full[(ids + off[:-1][None, :]).clamp(0, V - 1)] * valid[..., None]
        ok_fwd = torch.equal(tile, want)
        row_grads = torch.randn((B, 4, D), generator=gi)
        dense = R.exchange_reference(shard.shape[0], off, ids, row_grads)
        # single-process truth: gather everything, index_add into the full table, take my shard
        all_ids = [torch.empty_like(ids) for _ in range(world)]
        all_g = [torch.empty_like(row_grads) for _ in range(world)]
        dist.all_gather(all_ids, ids); dist.all_gather(all_g, row_grads)
        fullg = torch.zeros((V, D), dtype=torch.float64)
        for i_, g_ in zip(all_ids, all_g):
            v_ = ((i_ >= 0) & (i_ < rows[None, :])).reshape(-1)
            fullg.index_add_(0, (i_ + off[:-1][None, :]).reshape(-1)[v_], g_.reshape(-1, D)[v_].double())
        ok_bwd = torch.allclose(dense, S.full_to_shard(fullg, rank, world), rtol=0, atol=1e-12)
        # owner-side Adam on the shard == the shard of Adam on the whole table (oracle.layers_np.adam_sparse_apply), both variants
        from oracle import layers_np as O
        ok_adam = True
        rows_all = np.concatenate([(i_ + off[:-1][None, :]).reshape(-1)[((i_ >= 0) & (i_ < rows[None, :])).reshape(-1)].numpy() for i_ in all_ids])
        vals_all = np.concatenate([g_.reshape(-1, D)[((i_ >= 0) & (i_ < rows[None, :])).reshape(-1)].numpy() for i_, g_ in zip(all_ids, all_g)])
        touched_full = torch.zeros(V, dtype=torch.bool); touched_full[torch.as_tensor(rows_all)] = True
        touched = S.full_to_shard(touched_full[:, None].float(), rank, world)[:, 0] > 0
        for lazy in (False, True):
            w, m, v = O.adam_sparse_apply(full.numpy(), np.zeros((V, D)), np.zeros((V, D)), rows_all, vals_all, 1, 0.01, lazy=lazy)
            zs = torch.zeros_like(shard, dtype=torch.float64)
            ws, ms, vs = R.adam_reference(shard.double(), zs, zs, dense, touched, 1, 0.01, lazy)
            for got, want_full in ((ws, w), (ms, m), (vs, v)):
                ok_adam &= torch.allclose(got, S.full_to_shard(torch.as_tensor(want_full), rank, world), rtol=0, atol=1e-12)
        q.put((rank, bool(ok_fwd), bool(ok_bwd and ok_adam)))
    finally:
        dist.destroy_process_group()


def test_reference_exchange_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True, True), (1, True, True)]


def _fd_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank offers the read end of a pipe that already holds its rank byte; the peers read it through their duplicate
        r_fd, w_fd = os.pipe()
        os.write(w_fd, bytes([65 + rank]) * (world - 1))
        fds = S.exchange_fds(r_fd)
        got = []
        for r in range(world):
            if r == rank:
                assert fds[r] is None
                continue
            got.append((r, os.read(fds[r], 1)))
            os.close(fds[r])
        q.put((rank, all(b == bytes([65 + r]) for r, b in got) and len(got) == world - 1))
    finally:
        dist.destroy_process_group()


def test_fd_exchange_gloo_world3():
    """The descriptor hand-over VmmBuffer relies on (AF_UNIX + SCM_RIGHTS), without CUDA: 3 processes swap pipe ends."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fd_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(3)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True), (2, True)]
