"""Pure-torch reference of the row-sharded lookup and gradient exchange (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Runs on any backend (the CPU tests use gloo, world_size 2): every rank buckets its IndexedSlices by owner
(gr % G), the buckets are exchanged with all_gather (gloo has no all_to_all), and each owner densifies what it
received.  The CUDA path (peer-pull forward, fused push backward) is checked against this."""
from __future__ import annotations

import torch
import torch.distributed as dist


def lookup_reference(shard: torch.Tensor, field_row_offset: torch.Tensor, ids: torch.Tensor, group=None) -> torch.Tensor:
    """(B,F) ids -> (B,F,D) tile using an all_gather of the shards (fine at test sizes)."""
    G = dist.get_world_size(group)
    shards = [torch.empty_like(shard) for _ in range(G)]
    dist.all_gather(shards, shard, group=group)
    B, F = ids.shape
    rows = field_row_offset[1:] - field_row_offset[:-1]
    valid = (ids >= 0) & (ids < rows[None, :])
    gr = (ids + field_row_offset[:-1][None, :]).clamp(0, int(field_row_offset[-1]) - 1)     # invalid ids are masked below
    stacked = torch.stack(shards)                                  # (G, local_rows, D)
    out = stacked[(gr % G).reshape(-1), (gr // G).reshape(-1)].reshape(B, F, -1)
    return out * valid[..., None]


def exchange_reference(local_rows: int, field_row_offset: torch.Tensor, ids: torch.Tensor, row_grads: torch.Tensor,
                       group=None) -> torch.Tensor:
    """Dense (local_rows, D) gradient shard this rank owns, from every rank's (ids, row_grads)."""
    G, rank = dist.get_world_size(group), dist.get_rank(group)
    D = row_grads.shape[-1]
    rows = field_row_offset[1:] - field_row_offset[:-1]
    valid = ((ids >= 0) & (ids < rows[None, :])).reshape(-1)
    gr = (ids + field_row_offset[:-1][None, :]).reshape(-1)[valid]
    vals = row_grads.reshape(-1, D)[valid]
    n = torch.tensor([gr.numel()], dtype=torch.int64, device=ids.device)
    counts = [torch.zeros_like(n) for _ in range(G)]
    dist.all_gather(counts, n, group=group)
    cap = int(max(c.item() for c in counts))
    pad_r = torch.full((cap,), -1, dtype=torch.int64, device=ids.device)
    pad_v = torch.zeros((cap, D), dtype=row_grads.dtype, device=ids.device)
    pad_r[: gr.numel()] = gr
    pad_v[: gr.numel()] = vals
    all_r = [torch.empty_like(pad_r) for _ in range(G)]
    all_v = [torch.empty_like(pad_v) for _ in range(G)]
    dist.all_gather(all_r, pad_r, group=group)
    dist.all_gather(all_v, pad_v, group=group)
    dense = torch.zeros((local_rows, D), dtype=torch.float64, device=ids.device)
    for r_, v_ in zip(all_r, all_v):
        mine = (r_ >= 0) & (r_ % G == rank)
        dense.index_add_(0, r_[mine] // G, v_[mine].double())
    return dense


def adam_reference(var, m, v, g_dense, touched, t, lr, lazy, beta1=0.9, beta2=0.999, eps=1e-8):
    """float64 Adam / LazyAdam step on a shard from its densified gradient (same float32 coefficients as
    oracle.layers_np.adam_sparse_apply); `touched` marks the rows that received at least one entry."""
    import numpy as np
    om1 = float(np.float32(1.0) - np.float32(beta1)); om2 = float(np.float32(1.0) - np.float32(beta2))
    b1 = float(np.float32(beta1)); b2 = float(np.float32(beta2))
    lr_t = lr * (1.0 - beta2 ** t) ** 0.5 / (1.0 - beta1 ** t)
    m2 = b1 * m + om1 * g_dense
    v2 = b2 * v + om2 * g_dense * g_dense
    w2 = var - lr_t * m2 / (v2.sqrt() + eps)
    if lazy:
        sel = touched[:, None]
        return torch.where(sel, w2, var), torch.where(sel, m2, m), torch.where(sel, v2, v)
    return w2, m2, v2
