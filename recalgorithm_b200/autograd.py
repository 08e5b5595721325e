"""torch.autograd glue: each Function is one forward and one backward call into the C ABI.

The gradient of an embedding table is NOT a dense tensor: like TensorFlow's gather gradient
(``IndexedSlices``; the reference's optimizer consumes exactly that -- DeepFM/deepfm.py:246-250,
SURVEY A.8) it is the pair (ids, row values), attached to the table variable as ``grad_slices``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

from . import ops


@dataclass
class IndexedSlices:
    """values[b, f, :] is the gradient of row field_row_offset[f] + ids[b, f] (ids < 0 / out of range: dropped)."""
    values: torch.Tensor            # (B, F, D)
    ids: torch.Tensor               # (B, F) int64, per-field local ids
    field_row_offset: torch.Tensor  # (F+1,) int64

    def to_dense(self, num_rows: int) -> torch.Tensor:
        dense = torch.zeros((num_rows, self.values.shape[-1]), dtype=self.values.dtype, device=self.values.device)
        return ops.embed_scatter_add(dense, self.field_row_offset, self.ids, self.values)


class EmbeddingTables:
    """F per-field embedding tables stored back to back in one (V_total, D) buffer.

    Mirrors the variables ``fc.embedding_column`` creates implicitly (one ``embedding_weights`` per
    column; DeepFM/deepfm.py:83-89) but keeps them contiguous so one kernel serves all fields.
    """

    def __init__(self, rows_per_field, dim: int, device="cuda", init: Optional[str] = "truncated_normal",
                 generator: Optional[torch.Generator] = None):
        rows = torch.as_tensor(rows_per_field, dtype=torch.int64)
        self.num_fields = int(rows.numel())
        self.dim = int(dim)
        off = torch.zeros(self.num_fields + 1, dtype=torch.int64)
        off[1:] = torch.cumsum(rows, 0)
        self.num_rows = int(off[-1])
        self.field_row_offset = off.to(device)
        self.weight = torch.empty((self.num_rows, self.dim), dtype=torch.float32, device=device)
        if init == "truncated_normal":          # embedding_column default: truncated_normal(0, 1/sqrt(D))  [SURVEY A.5 / note 4]
            torch.nn.init.trunc_normal_(self.weight, mean=0.0, std=self.dim ** -0.5, a=-2 * self.dim ** -0.5,
                                        b=2 * self.dim ** -0.5, generator=generator)
        self.grad_slices: List[IndexedSlices] = []
        self._anchor = torch.zeros((), device=device, requires_grad=True)   # lets autograd reach backward()

    def zero_grad(self):
        self.grad_slices.clear()


class _LookupFM2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, tables: EmbeddingTables, ids: torch.Tensor, want_fm2: bool):
        tile, fm2 = ops.embed_fm2_fwd(tables.weight, tables.field_row_offset, ids, want_tile=True, want_fm2=want_fm2)
        ctx.tables, ctx.ids, ctx.want_fm2 = tables, ids, want_fm2
        ctx.save_for_backward(tile)
        if want_fm2:
            return tile, fm2
        return tile

    @staticmethod
    def backward(ctx, d_tile, d_fm2=None):
        (tile,) = ctx.saved_tensors
        if d_tile is not None:
            d_tile = d_tile.contiguous()
        if d_fm2 is not None:
            d_fm2 = d_fm2.contiguous()
        values = ops.embed_fm2_bwd(tile, d_tile, d_fm2 if ctx.want_fm2 else None)
        ctx.tables.grad_slices.append(IndexedSlices(values, ctx.ids, ctx.tables.field_row_offset))
        return None, None, None, None


def lookup_fm2(tables: EmbeddingTables, ids: torch.Tensor):
    """(B,F) ids -> (tile (B,F,D), fm2 logit (B,1)); differentiable w.r.t. the tables (IndexedSlices)."""
    return _LookupFM2.apply(tables._anchor, tables, ids, True)


def lookup(tables: EmbeddingTables, ids: torch.Tensor):
    """(B,F) ids -> tile (B,F,D)."""
    return _LookupFM2.apply(tables._anchor, tables, ids, False)


class _CrossStack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, xl, w, b):
        x0c = x0.contiguous()
        xlc = None if xl is None else xl.contiguous()
        wc, bc = w.contiguous(), b.contiguous()
        ctx.save_for_backward(x0c, wc, bc, *([] if xlc is None else [xlc]))
        ctx.has_xl = xlc is not None
        return ops.cross_fwd(x0c, wc, bc, xl_in=xlc)

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        x0, w, b = saved[:3]
        xl = saved[3] if ctx.has_xl else None
        dx0, dxl, dw, db = ops.cross_bwd(x0, w, b, g.contiguous(), xl_in=xl)
        return dx0, dxl, dw, db


def cross_stack(x0: torch.Tensor, w: torch.Tensor, b: torch.Tensor, xl: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All L cross layers (DCN/dcn.py:157-160) in one launch.  w, b: (L, d)."""
    return _CrossStack.apply(x0, xl, w, b)
