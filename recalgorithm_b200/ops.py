"""Thin tensor-level wrappers over the C ABI (device pointers + sizes + the current CUDA stream).

torch is plumbing here: it owns device memory and streams; every computation below is one call
into libctr_b200.so.  Functions validate devices/dtypes and raise instead of falling back.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

F32, I64 = torch.float32, torch.int64


_empty_anchor = {}


def _ptr(t: Optional[torch.Tensor]):
    """Device address of a tensor; None stays NULL (= "output not wanted").  An EMPTY tensor (batch 0) has data_ptr() == 0,
    which the C ABI would read as a missing argument: it gets the address of a small per-device anchor instead -- never
    dereferenced, every entry point returns before launching when its batch dimension is 0."""
    if t is None:
        return None
    if t.numel() == 0:
        a = _empty_anchor.get(t.device)
        if a is None:
            a = _empty_anchor[t.device] = torch.zeros(64, dtype=torch.uint8, device=t.device)
        return a.data_ptr()
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


I32 = torch.int32


def _chk(t: Optional[torch.Tensor], dtype, name: str, shape=None):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if t.device.index != torch.cuda.current_device():
        # kernels launch on the CURRENT device's current stream: a tensor of another GPU would be touched from the wrong
        # device / stream (single-process multi-GPU callers wrap their calls in `with torch.cuda.device(t.device)`)
        raise RuntimeError(f"{name}: lives on cuda:{t.device.index} but the current device is cuda:{torch.cuda.current_device()}")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")


# ------------------------------------------------------------------ Row L + FM2
def embed_fm2_fwd(table: torch.Tensor, field_row_offset: torch.Tensor, ids: torch.Tensor,
                  want_tile: bool = True, want_fm2: bool = True,
                  tile: Optional[torch.Tensor] = None, fm2: Optional[torch.Tensor] = None,
                  ids64_out: Optional[torch.Tensor] = None
                  ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Fused lookup (+ FM second-order logit).  table (V,D); field_row_offset (F+1,) i64; ids (B,F) i64 -- or i32 (half the
    bytes over PCIe; ``ids64_out`` (B,F) i64 then receives the widened copy for IndexedSlices consumers).
    Returns (tile (B,F,D) | None, fm2 (B,1) | None)."""
    B, F = ids.shape
    D = table.shape[1]
    _chk(table, F32, "table"); _chk(field_row_offset, I64, "field_row_offset", (F + 1,))
    if ids.dtype == I32:
        _chk(ids, I32, "ids"); _chk(ids64_out, I64, "ids64_out", (B, F))
        if want_tile and tile is None:
            tile = torch.empty((B, F, D), dtype=F32, device=table.device)
        if want_fm2 and fm2 is None:
            fm2 = torch.empty((B, 1), dtype=F32, device=table.device)
        _chk(tile, F32, "tile", (B, F, D)); _chk(fm2, F32, "fm2", (B, 1))
        _lib.check(_lib.lib().ctr_embed_fm2_fwd_ids32(_ptr(table), _ptr(field_row_offset), _ptr(ids), B, F, D, _ptr(tile),
                                                      _ptr(fm2), _ptr(ids64_out), _stream()))
        return tile, fm2
    _chk(ids, I64, "ids")
    if want_tile and tile is None:
        tile = torch.empty((B, F, D), dtype=F32, device=table.device)
    if want_fm2 and fm2 is None:
        fm2 = torch.empty((B, 1), dtype=F32, device=table.device)
    _chk(tile, F32, "tile", (B, F, D)); _chk(fm2, F32, "fm2", (B, 1))
    _lib.check(_lib.lib().ctr_embed_fm2_fwd(_ptr(table), _ptr(field_row_offset), _ptr(ids), B, F, D,
                                            _ptr(tile), _ptr(fm2), _stream()))
    return tile, fm2


def embed_fm2_bwd(tile: torch.Tensor, d_tile: Optional[torch.Tensor], d_fm2: Optional[torch.Tensor],
                  row_grads: Optional[torch.Tensor] = None) -> torch.Tensor:
    """IndexedSlices values (B,F,D) of the lookup gradient: d_tile + d_fm2 * (S - e)."""
    B, F, D = tile.shape
    _chk(tile, F32, "tile"); _chk(d_tile, F32, "d_tile", (B, F, D))
    if d_fm2 is not None:
        d_fm2 = d_fm2.reshape(B)
    _chk(d_fm2, F32, "d_fm2", (B,))
    if row_grads is None:
        row_grads = torch.empty_like(tile)
    _chk(row_grads, F32, "row_grads", (B, F, D))
    _lib.check(_lib.lib().ctr_embed_fm2_bwd(_ptr(tile), _ptr(d_tile), _ptr(d_fm2), B, F, D, _ptr(row_grads), _stream()))
    return row_grads


def embed_fm2_lin_fwd(table: torch.Tensor, field_row_offset: torch.Tensor, ids: torch.Tensor, wlin: torch.Tensor,
                      want_tile: bool = True, ids64_out: Optional[torch.Tensor] = None):
    """Lookup + FM2 + fused dense(1) head over the flattened tile: returns (tile | None, fm2 (B,1), lin (B,1)) with
    lin = tile.reshape(B, F*D) @ wlin.  ids int64 or int32 (``ids64_out`` then receives the widened copy)."""
    B, F = ids.shape
    D = table.shape[1]
    _chk(table, F32, "table"); _chk(field_row_offset, I64, "field_row_offset", (F + 1,))
    i32 = ids.dtype == I32
    _chk(ids, I32 if i32 else I64, "ids"); _chk(ids64_out, I64, "ids64_out", (B, F))
    wlin = wlin.reshape(F * D)
    _chk(wlin, F32, "wlin", (F * D,))
    tile = torch.empty((B, F, D), dtype=F32, device=table.device) if want_tile else None
    fm2 = torch.empty((B, 1), dtype=F32, device=table.device)
    lin = torch.empty((B, 1), dtype=F32, device=table.device)
    _lib.check(_lib.lib().ctr_embed_fm2_lin_fwd(_ptr(table), _ptr(field_row_offset), _ptr(ids), int(i32), B, F, D, _ptr(wlin),
                                                _ptr(tile), _ptr(fm2), _ptr(lin), _ptr(ids64_out), _stream()))
    return tile, fm2, lin


def embed_fm2_lin_bwd(tile: torch.Tensor, wlin: torch.Tensor, d_fm2: Optional[torch.Tensor], d_lin: Optional[torch.Tensor],
                      row_grads: Optional[torch.Tensor] = None):
    """Backward of embed_fm2_lin_fwd: (row_grads (B,F,D) = IndexedSlices values, d_wlin (F*D,))."""
    B, F, D = tile.shape
    _chk(tile, F32, "tile")
    wlin = wlin.reshape(F * D)
    _chk(wlin, F32, "wlin", (F * D,))
    if d_fm2 is not None:
        d_fm2 = d_fm2.reshape(B)
    if d_lin is not None:
        d_lin = d_lin.reshape(B)
    _chk(d_fm2, F32, "d_fm2", (B,)); _chk(d_lin, F32, "d_lin", (B,))
    if row_grads is None:
        row_grads = torch.empty_like(tile)
    _chk(row_grads, F32, "row_grads", (B, F, D))
    d_wlin = torch.empty((F * D,), dtype=F32, device=tile.device)
    _lib.check(_lib.lib().ctr_embed_fm2_lin_bwd(_ptr(tile), _ptr(wlin), _ptr(d_fm2), _ptr(d_lin), B, F, D, _ptr(row_grads),
                                                _ptr(d_wlin), _stream()))
    return row_grads, d_wlin


def embed_seq_fwd(table: torch.Tensor, ids: torch.Tensor, row_range: Optional[torch.Tensor] = None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Sequence lookup: ids (B,T) int64 all index ONE table (rows [row_range[0], row_range[1]) of `table`, default the whole
    table) -> (B,T,D); id < 0 / out of range -> zero row.  The gradient is the IndexedSlices (ids, d_out) as is."""
    B, T = ids.shape
    V, D = table.shape
    _chk(table, F32, "table"); _chk(ids, I64, "ids")
    if row_range is None:
        row_range = torch.tensor([0, V], dtype=I64, device=table.device)
    _chk(row_range, I64, "row_range", (2,))
    if out is None:
        out = torch.empty((B, T, D), dtype=F32, device=table.device)
    _chk(out, F32, "out", (B, T, D))
    _lib.check(_lib.lib().ctr_embed_seq_fwd(_ptr(table), _ptr(row_range), _ptr(ids), B, T, D, _ptr(out), _stream()))
    return out


def sigmoid_ce(logit_a: torch.Tensor, logit_b: Optional[torch.Tensor], labels: torch.Tensor, want_grad: bool = True):
    """Mean sigmoid cross-entropy of (logit_a + logit_b) vs labels and d(loss)/d(logit) in one launch: (loss (1,), d_logit (B,1))."""
    B = logit_a.numel()
    a = logit_a.reshape(B); b = None if logit_b is None else logit_b.reshape(B); y = labels.reshape(B)
    _chk(a, F32, "logit_a"); _chk(b, F32, "logit_b", (B,)); _chk(y, F32, "labels", (B,))
    loss = torch.empty((1,), dtype=F32, device=a.device)
    d = torch.empty((B, 1), dtype=F32, device=a.device) if want_grad else None
    _lib.check(_lib.lib().ctr_sigmoid_ce(_ptr(a), _ptr(b), _ptr(y), B, _ptr(loss), _ptr(d), _stream()))
    return loss, d


def embed_scatter_add(grad_table: torch.Tensor, field_row_offset: torch.Tensor, ids: torch.Tensor,
                      row_grads: torch.Tensor) -> torch.Tensor:
    B, F, D = row_grads.shape
    _chk(grad_table, F32, "grad_table"); _chk(field_row_offset, I64, "field_row_offset", (F + 1,))
    _chk(ids, I64, "ids", (B, F)); _chk(row_grads, F32, "row_grads")
    _lib.check(_lib.lib().ctr_embed_scatter_add(_ptr(grad_table), _ptr(field_row_offset), _ptr(ids), _ptr(row_grads),
                                                B, F, D, _stream()))
    return grad_table


def first_order_fwd(w: torch.Tensor, field_row_offset: torch.Tensor, ids: torch.Tensor, bias: float = 0.0) -> torch.Tensor:
    """DeepFM first-order logit (B,1): bias + sum_f w[row(b,f)].  w (V_total,) = the dense(1) kernel over the indicators."""
    B, F = ids.shape
    _chk(w, F32, "w"); _chk(field_row_offset, I64, "field_row_offset", (F + 1,)); _chk(ids, I64, "ids")
    out = torch.empty((B, 1), dtype=F32, device=w.device)
    _lib.check(_lib.lib().ctr_first_order_fwd(_ptr(w), _ptr(field_row_offset), _ptr(ids), B, F, float(bias), _ptr(out),
                                              _stream()))
    return out


def bag_lookup_fwd(table: torch.Tensor, ids: torch.Tensor, offsets: torch.Tensor,
                   out: Optional[torch.Tensor] = None, out_col: int = 0) -> torch.Tensor:
    """Multi-valued lookup, combiner='mean'.  Writes out[:, out_col:out_col+D] of a (B, stride) buffer."""
    V, D = table.shape
    B = offsets.numel() - 1
    _chk(table, F32, "table"); _chk(ids, I64, "ids"); _chk(offsets, I64, "offsets")
    if out is None:
        out = torch.empty((B, D), dtype=F32, device=table.device)
    _chk(out, F32, "out")
    stride = out.shape[1]
    if out_col + D > stride:
        raise ValueError("bag_lookup_fwd: field does not fit in the output row")
    _lib.check(_lib.lib().ctr_bag_lookup_fwd(_ptr(table), V, D, _ptr(ids), _ptr(offsets), B,
                                             out.data_ptr() + 4 * out_col, stride, _stream()))
    return out


def bag_lookup_bwd(d_out: torch.Tensor, out_col: int, V: int, D: int, ids: torch.Tensor,
                   offsets: torch.Tensor) -> torch.Tensor:
    B = offsets.numel() - 1
    _chk(d_out, F32, "d_out"); _chk(ids, I64, "ids"); _chk(offsets, I64, "offsets")
    row_grads = torch.empty((ids.numel(), D), dtype=F32, device=d_out.device)
    _lib.check(_lib.lib().ctr_bag_lookup_bwd(d_out.data_ptr() + 4 * out_col, d_out.shape[1], V, D, _ptr(ids),
                                             _ptr(offsets), B, _ptr(row_grads), _stream()))
    return row_grads


# ------------------------------------------------------------------ Row CROSS
def cross_fwd(x0: torch.Tensor, w: torch.Tensor, b: torch.Tensor, xl_in: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x_L of the cross stack.  x0 (B,d); w, b (L,d); xl_in optional start vector (default x0)."""
    B, d = x0.shape
    L = w.shape[0]
    _chk(x0, F32, "x0"); _chk(w, F32, "w", (L, d)); _chk(b, F32, "b", (L, d)); _chk(xl_in, F32, "xl_in", (B, d))
    if out is None:
        out = torch.empty_like(x0)
    _chk(out, F32, "out", (B, d))
    _lib.check(_lib.lib().ctr_cross_fwd(_ptr(x0), _ptr(xl_in), _ptr(w), _ptr(b), B, d, L, _ptr(out), _stream()))
    return out


def embed_cross_supported(F: int, D: int, L: int) -> bool:
    """Shapes the fused lookup + cross forward covers (ctr_embed_cross_fwd): D % 4 == 0, F*D <= 512, L <= 4."""
    return D % 4 == 0 and F * D <= 512 and 1 <= L <= 4


def embed_cross_fwd(table: torch.Tensor, field_row_offset: torch.Tensor, ids: torch.Tensor, w: torch.Tensor, b: torch.Tensor,
                    x0: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Lookup fused with the cross stack (DCN/dcn.py:153-160): ids (B,F) i64 | i32 -> (x0 (B,F*D), x_L (B,F*D)) in one launch.
    Shapes outside embed_cross_supported() run the lookup kernel followed by the cross kernel (same results)."""
    B, F = ids.shape
    D = table.shape[1]
    d, L = F * D, w.shape[0]
    _chk(table, F32, "table"); _chk(field_row_offset, I64, "field_row_offset", (F + 1,))
    _chk(w, F32, "w", (L, d)); _chk(b, F32, "b", (L, d))
    if ids.dtype not in (I64, I32):
        raise TypeError(f"ids must be int64 or int32, got {ids.dtype}")
    _chk(ids, ids.dtype, "ids")
    if x0 is None:
        x0 = torch.empty((B, d), dtype=F32, device=table.device)
    if out is None:
        out = torch.empty((B, d), dtype=F32, device=table.device)
    _chk(x0, F32, "x0", (B, d)); _chk(out, F32, "out", (B, d))
    if not embed_cross_supported(F, D, L):
        if ids.dtype == I32:
            ids = ids.long()
        embed_fm2_fwd(table, field_row_offset, ids, want_fm2=False, tile=x0.view(B, F, D))
        return x0, cross_fwd(x0, w, b, out=out)
    _lib.check(_lib.lib().ctr_embed_cross_fwd(_ptr(table), _ptr(field_row_offset), _ptr(ids), int(ids.dtype == I32), B, F, D,
                                              _ptr(w), _ptr(b), L, _ptr(x0), _ptr(out), _stream()))
    return x0, out


def cross_bwd(x0, w, b, g_out, xl_in=None):
    """Returns (dx0, dxl_in | None, dw, db)."""
    B, d = x0.shape
    L = w.shape[0]
    _chk(x0, F32, "x0"); _chk(w, F32, "w", (L, d)); _chk(b, F32, "b", (L, d))
    _chk(g_out, F32, "g_out", (B, d)); _chk(xl_in, F32, "xl_in", (B, d))
    dx0 = torch.empty_like(x0)
    dxl = torch.empty_like(x0) if xl_in is not None else None
    dwb = torch.empty((2,) + tuple(w.shape), dtype=F32, device=w.device)     # dw | db back to back: zeroed by one memset
    dw, db = dwb[0], dwb[1]
    _lib.check(_lib.lib().ctr_cross_bwd(_ptr(x0), _ptr(xl_in), _ptr(w), _ptr(b), _ptr(g_out), B, d, L,
                                        _ptr(dx0), _ptr(dxl), _ptr(dw), _ptr(db), _stream()))
    return dx0, dxl, dw, db


# ------------------------------------------------------------------ Row DIN-ATT
def _din_params(H, w1, b1, w2, b2, w3, b3):
    w3 = w3.reshape(32)
    b3 = b3.reshape(1)
    for t, n, s in ((w1, "w1", (4 * H, 64)), (b1, "b1", (64,)), (w2, "w2", (64, 32)), (b2, "b2", (32,)),
                    (w3, "w3", (32,)), (b3, "b3", (1,))):
        _chk(t, F32, n, s)
    return w1, b1, w2, b2, w3, b3


_din_sched = {}


def _din_sched_scratch(B: int, device, balanced: bool):
    """int32[B + 64] schedule scratch of the DIN kernels (longest-first dynamic work distribution), cached per (device, stream):
    two streams running the attention at the same time must not share the sample list."""
    if not balanced or B == 0:
        return None
    key = (device, _stream())
    t = _din_sched.get(key)
    if t is None or t.numel() < B + 64:
        t = _din_sched[key] = torch.empty((B + 64,), dtype=torch.int32, device=device)
    return t


def din_attention_fwd(query, keys, keys_length, w1, b1, w2, b2, w3, b3, is_softmax=False, want_weights=False, balanced=True):
    """DIN attention unit.  query (B,H); keys (B,T,H); keys_length (B,) int64.  Returns out (B,H) [, att_w (B,T)].
    balanced: samples are handed to the warps longest-first from a shared counter instead of round-robin."""
    B, T, H = keys.shape
    _chk(query, F32, "query", (B, H)); _chk(keys, F32, "keys"); _chk(keys_length, I64, "keys_length", (B,))
    w1, b1, w2, b2, w3, b3 = _din_params(H, w1, b1, w2, b2, w3, b3)
    out = torch.empty((B, H), dtype=F32, device=query.device)
    att = torch.empty((B, T), dtype=F32, device=query.device) if want_weights else None
    _lib.check(_lib.lib().ctr_din_attention_fwd(_ptr(query), _ptr(keys) if T > 0 else None, _ptr(keys_length), _ptr(w1),
                                                _ptr(b1), _ptr(w2), _ptr(b2), _ptr(w3), _ptr(b3), B, T, H,
                                                int(bool(is_softmax)), _ptr(out), _ptr(att),
                                                _ptr(_din_sched_scratch(B, query.device, balanced)), _stream()))
    return (out, att) if want_weights else out


def din_attention_bwd(query, keys, keys_length, w1, b1, w2, b2, w3, b3, g_out, is_softmax=False, att_w=None, balanced=True):
    """Returns (d_query, d_keys, [dw1, db1, dw2, db2, dw3, db3]).  att_w: the forward's (B,T) weights (else recomputed)."""
    B, T, H = keys.shape
    _chk(query, F32, "query", (B, H)); _chk(keys, F32, "keys"); _chk(keys_length, I64, "keys_length", (B,))
    _chk(g_out, F32, "g_out", (B, H)); _chk(att_w, F32, "att_w", (B, T))
    w3_shape, b3_shape = w3.shape, b3.shape
    w1, b1, w2, b2, w3, b3 = _din_params(H, w1, b1, w2, b2, w3, b3)
    dq = torch.empty_like(query)
    dk = torch.empty_like(keys)
    sizes = [4 * H * 64, 64, 64 * 32, 32, 32, 1]
    flat = torch.empty((sum(sizes),), dtype=F32, device=query.device)
    _lib.check(_lib.lib().ctr_din_attention_bwd(_ptr(query), _ptr(keys) if T > 0 else None, _ptr(keys_length), _ptr(w1),
                                                _ptr(b1), _ptr(w2), _ptr(b2), _ptr(w3), _ptr(b3), _ptr(g_out), _ptr(att_w), B, T, H,
                                                int(bool(is_softmax)), _ptr(dq), _ptr(dk) if T > 0 else None, _ptr(flat),
                                                _ptr(_din_sched_scratch(B, query.device, balanced)), _stream()))
    parts = list(torch.split(flat, sizes))
    shapes = [(4 * H, 64), (64,), (64, 32), (32,), tuple(w3_shape), tuple(b3_shape)]
    return dq, dk, [p.reshape(s) for p, s in zip(parts, shapes)]


# ------------------------------------------------------------------ Rows SENET / BILINEAR
def senet_fwd(x, w1, w2):
    B, F, K = x.shape
    r = w1.shape[1]
    _chk(x, F32, "x"); _chk(w1, F32, "w1", (F, r)); _chk(w2, F32, "w2", (r, F))
    out = torch.empty_like(x)
    _lib.check(_lib.lib().ctr_senet_fwd(_ptr(x), _ptr(w1), _ptr(w2), B, F, K, r, _ptr(out), _stream()))
    return out


def senet_bwd(x, w1, w2, g_out):
    B, F, K = x.shape
    r = w1.shape[1]
    _chk(x, F32, "x"); _chk(w1, F32, "w1", (F, r)); _chk(w2, F32, "w2", (r, F)); _chk(g_out, F32, "g_out", (B, F, K))
    dx, dw1, dw2 = torch.empty_like(x), torch.empty_like(w1), torch.empty_like(w2)
    _lib.check(_lib.lib().ctr_senet_bwd(_ptr(x), _ptr(w1), _ptr(w2), _ptr(g_out), B, F, K, r, _ptr(dx), _ptr(dw1),
                                        _ptr(dw2), _stream()))
    return dx, dw1, dw2


BILINEAR_TYPES = {"all": 0, "each": 1, "interaction": 2}


def _bilinear_type(type_):
    if type_ not in BILINEAR_TYPES:      # same message as FiBiNET/bilinear_interaction_layer.py:36-38
        raise ValueError(f"Bilinear Interaction type must be in ['all','each','interaction'], got '{type_}'")
    return BILINEAR_TYPES[type_]


def bilinear_w_shape(F, K, type_):
    return {"all": (K, K), "each": (F - 1, K, K), "interaction": (F * (F - 1) // 2, K, K)}[type_]


def bilinear_set_tournament(mask):
    """Tuning hook (see ctr_bilinear_set_rr in include/ctr_b200.h): bits 0..2 route type 'all' / 'each' / 'interaction' through
    the sample-batched tournament kernels (default 4: 'interaction' only; 'all' / 'each' run the staged per-sample kernels),
    bit 3 selects the round-1 kernels for 'all' / 'each'.  Returns the previous mask."""
    return int(_lib.lib().ctr_bilinear_set_rr(int(mask)))


def bilinear_fwd(x, w, type_):
    t = _bilinear_type(type_)
    B, F, K = x.shape
    _chk(x, F32, "x"); _chk(w, F32, "w", bilinear_w_shape(F, K, type_))
    P = (F - 1) * (F - 2) // 2
    out = torch.empty((B, P, K), dtype=F32, device=x.device)
    if B == 0 or P == 0:
        return out
    _lib.check(_lib.lib().ctr_bilinear_fwd(_ptr(x), _ptr(w), B, F, K, t, _ptr(out), _stream()))
    return out


def bilinear_bwd(x, w, type_, g_out):
    t = _bilinear_type(type_)
    B, F, K = x.shape
    P = (F - 1) * (F - 2) // 2
    _chk(x, F32, "x"); _chk(w, F32, "w", bilinear_w_shape(F, K, type_)); _chk(g_out, F32, "g_out", (B, P, K))
    dx, dw = torch.empty_like(x), torch.empty_like(w)
    if B == 0 or P == 0:
        return dx.zero_(), dw.zero_()
    _lib.check(_lib.lib().ctr_bilinear_bwd(_ptr(x), _ptr(w), _ptr(g_out), B, F, K, t, _ptr(dx), _ptr(dw), _stream()))
    return dx, dw


# ------------------------------------------------------------------ Row CIN
_cin_ws = {}


def _cin_workspace(nbytes: int, device) -> Optional[torch.Tensor]:
    """Per-(device, stream) cached scratch for the re-ordered / tf32-split filter (caller-owned, as the ABI requires)."""
    if nbytes == 0:
        return None
    key = (device.index, _stream())
    buf = _cin_ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty((nbytes,), dtype=torch.uint8, device=device)
        _cin_ws[key] = buf
    return buf


def cin_fwd(x0: torch.Tensor, xk: torch.Tensor, filt: torch.Tensor, want_pooled: bool = False, precision: int = 0):
    """One CIN layer.  x0 (B,m,D); xk (B,hk,D); filt (hk*m, H).  Returns out (B,H,D) [, pooled (B,H)]."""
    B, m, D = x0.shape
    hk = xk.shape[1]
    H = filt.shape[1]
    _chk(x0, F32, "x0"); _chk(xk, F32, "xk", (B, hk, D)); _chk(filt, F32, "filter", (hk * m, H))
    out = torch.empty((B, H, D), dtype=F32, device=x0.device)
    pooled = torch.empty((B, H), dtype=F32, device=x0.device) if want_pooled else None
    L = _lib.lib()
    nbytes = int(L.ctr_cin_fwd_workspace_bytes(B, m, hk, D, H))
    ws = _cin_workspace(nbytes, x0.device)
    _lib.check(L.ctr_cin_fwd(_ptr(x0), _ptr(xk), _ptr(filt), B, m, hk, D, H, _ptr(out), _ptr(pooled), int(precision),
                             _ptr(ws), nbytes, _stream()))
    return (out, pooled) if want_pooled else out


def cin_bwd(x0, xk, filt, g_out):
    """Returns (dx0, dxk, dfilter)."""
    B, m, D = x0.shape
    hk = xk.shape[1]
    H = filt.shape[1]
    _chk(x0, F32, "x0"); _chk(xk, F32, "xk", (B, hk, D)); _chk(filt, F32, "filter", (hk * m, H))
    _chk(g_out, F32, "g_out", (B, H, D))
    dx0, dxk, dw = torch.empty_like(x0), torch.empty_like(xk), torch.empty_like(filt)
    L = _lib.lib()
    nbytes = int(L.ctr_cin_bwd_workspace_bytes(B, m, hk, D, H))
    ws = _cin_workspace(nbytes, x0.device)
    _lib.check(L.ctr_cin_bwd(_ptr(x0), _ptr(xk), _ptr(filt), _ptr(g_out), B, m, hk, D, H, _ptr(dx0), _ptr(dxk), _ptr(dw),
                             _ptr(ws), nbytes, _stream()))
    return dx0, dxk, dw


# ------------------------------------------------------------------------------------ FM2 siblings
def embed_bi_fwd(table: torch.Tensor, field_row_offset: torch.Tensor, ids: torch.Tensor, want_tile: bool = True):
    """Fused lookup + NFM bi-interaction pooling.  Returns (tile (B,F,D) | None, bi (B,D))."""
    B, F = ids.shape
    D = table.shape[1]
    _chk(table, F32, "table"); _chk(field_row_offset, I64, "field_row_offset", (F + 1,)); _chk(ids, I64, "ids")
    tile = torch.empty((B, F, D), dtype=F32, device=table.device) if want_tile else None
    bi = torch.empty((B, D), dtype=F32, device=table.device)
    _lib.check(_lib.lib().ctr_embed_bi_fwd(_ptr(table), _ptr(field_row_offset), _ptr(ids), B, F, D, _ptr(tile), _ptr(bi), _stream()))
    return tile, bi


def embed_bi_bwd(tile: torch.Tensor, d_tile: Optional[torch.Tensor], d_bi: torch.Tensor) -> torch.Tensor:
    """IndexedSlices values (B,F,D): d_tile + d_bi[b,:] * (S - e)."""
    B, F, D = tile.shape
    _chk(tile, F32, "tile"); _chk(d_tile, F32, "d_tile", (B, F, D)); _chk(d_bi, F32, "d_bi", (B, D))
    row_grads = torch.empty_like(tile)
    _lib.check(_lib.lib().ctr_embed_bi_bwd(_ptr(tile), _ptr(d_tile), _ptr(d_bi), B, F, D, _ptr(row_grads), _stream()))
    return row_grads


def fwfm_fwd(tile: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
    """FwFM second-order logit (B,1).  tile (B,F,K); r (F(F-1)/2,) pair strengths in utils.index_from_upper_triangular order."""
    B, F, K = tile.shape
    _chk(tile, F32, "tile"); _chk(r, F32, "r", (F * (F - 1) // 2,))
    out = torch.empty((B, 1), dtype=F32, device=tile.device)
    _lib.check(_lib.lib().ctr_fwfm_fwd(_ptr(tile), _ptr(r), B, F, K, _ptr(out), _stream()))
    return out


def fwfm_bwd(tile: torch.Tensor, r: torch.Tensor, g: torch.Tensor):
    B, F, K = tile.shape
    g = g.reshape(B)
    _chk(tile, F32, "tile"); _chk(r, F32, "r", (F * (F - 1) // 2,)); _chk(g, F32, "g", (B,))
    d_tile, d_r = torch.empty_like(tile), torch.empty_like(r)
    _lib.check(_lib.lib().ctr_fwfm_bwd(_ptr(tile), _ptr(r), _ptr(g), B, F, K, _ptr(d_tile), _ptr(d_r), _stream()))
    return d_tile, d_r


def afm_fwd(tile: torch.Tensor, w: torch.Tensor, b: torch.Tensor, h: torch.Tensor, want_score: bool = False):
    """AFM attention pooling (B,K) (+ the (B,P) softmax scores).  w (K,T), b (T,), h (T,1) or (T,)."""
    B, F, K = tile.shape
    T = w.shape[1]
    h = h.reshape(T)
    _chk(tile, F32, "tile"); _chk(w, F32, "w", (K, T)); _chk(b, F32, "b", (T,)); _chk(h, F32, "h", (T,))
    pooled = torch.empty((B, K), dtype=F32, device=tile.device)
    score = torch.empty((B, F * (F - 1) // 2), dtype=F32, device=tile.device) if want_score else None
    _lib.check(_lib.lib().ctr_afm_fwd(_ptr(tile), _ptr(w), _ptr(b), _ptr(h), B, F, K, T, _ptr(pooled), _ptr(score), _stream()))
    return (pooled, score) if want_score else pooled


def afm_bwd(tile, w, b, h, g_pooled):
    B, F, K = tile.shape
    T = w.shape[1]
    h_shape = h.shape
    h = h.reshape(T)
    _chk(tile, F32, "tile"); _chk(w, F32, "w", (K, T)); _chk(b, F32, "b", (T,)); _chk(h, F32, "h", (T,)); _chk(g_pooled, F32, "g_pooled", (B, K))
    d_tile, d_w, d_b, d_h = torch.empty_like(tile), torch.empty_like(w), torch.empty_like(b), torch.empty_like(h)
    _lib.check(_lib.lib().ctr_afm_bwd(_ptr(tile), _ptr(w), _ptr(b), _ptr(h), _ptr(g_pooled), B, F, K, T, _ptr(d_tile), _ptr(d_w),
                                      _ptr(d_b), _ptr(d_h), _stream()))
    return d_tile, d_w, d_b, d_h.reshape(h_shape)


# ------------------------------------------------------------------------------------ BST transformer block
BST_PARAM_ORDER = ("position_embedding", "w_q", "w_k", "w_v", "w_o", "ln1_beta", "ln1_gamma", "dense_kernel", "dense_bias",
                   "ln2_beta", "ln2_gamma")


def bst_param_shapes(d: int, heads: int, max_length: int):
    return {"position_embedding": (max_length, d), "w_q": (heads, d, d), "w_k": (heads, d, d), "w_v": (heads, d, d),
            "w_o": (heads * d, d), "ln1_beta": (d,), "ln1_gamma": (d,), "dense_kernel": (d, d), "dense_bias": (d,),
            "ln2_beta": (d,), "ln2_gamma": (d,)}


def bst_pack_params(params: dict, d: int, heads: int, max_length: int) -> torch.Tensor:
    """dict of named tensors -> the packed float buffer of ctr_bst_transformer_* (include/ctr_b200.h)."""
    shapes = bst_param_shapes(d, heads, max_length)
    parts = []
    for name in BST_PARAM_ORDER:
        t = params[name]
        if tuple(t.shape) != shapes[name]:
            raise ValueError(f"{name}: expected shape {shapes[name]}, got {tuple(t.shape)}")
        parts.append(t.reshape(-1))
    packed = torch.cat(parts).contiguous()
    assert packed.numel() == _lib.lib().ctr_bst_param_count(d, heads, max_length)
    return packed


def bst_unpack_params(packed: torch.Tensor, d: int, heads: int, max_length: int) -> dict:
    out, o = {}, 0
    for name, shp in bst_param_shapes(d, heads, max_length).items():
        n = 1
        for x in shp:
            n *= x
        out[name] = packed[o:o + n].reshape(shp)
        o += n
    return out


def _bst_args(queries, keys, values, keys_length, packed, heads, max_length):
    B, T, d = queries.shape
    _chk(queries, F32, "queries"); _chk(keys, F32, "keys", (B, T, d)); _chk(values, F32, "values", (B, T, d))
    _chk(keys_length, I64, "keys_length", (B,))
    _chk(packed, F32, "params", (int(_lib.lib().ctr_bst_param_count(d, heads, max_length)),))
    return B, T, d


def bst_transformer_fwd(queries, keys, values, keys_length, packed, heads: int, max_length: int, use_position_embedding: bool = True):
    B, T, d = _bst_args(queries, keys, values, keys_length, packed, heads, max_length)
    out = torch.empty_like(queries)
    _lib.check(_lib.lib().ctr_bst_transformer_fwd(_ptr(queries), _ptr(keys), _ptr(values), _ptr(keys_length), _ptr(packed), B, T, d,
                                                  heads, max_length, int(use_position_embedding), _ptr(out), _stream()))
    return out


def bst_transformer_bwd(queries, keys, values, keys_length, packed, g_out, heads: int, max_length: int,
                        use_position_embedding: bool = True):
    B, T, d = _bst_args(queries, keys, values, keys_length, packed, heads, max_length)
    _chk(g_out, F32, "g_out", (B, T, d))
    dq, dk, dv, dp = torch.empty_like(queries), torch.empty_like(queries), torch.empty_like(queries), torch.empty_like(packed)
    _lib.check(_lib.lib().ctr_bst_transformer_bwd(_ptr(queries), _ptr(keys), _ptr(values), _ptr(keys_length), _ptr(packed),
                                                  _ptr(g_out), B, T, d, heads, max_length, int(use_position_embedding), _ptr(dq),
                                                  _ptr(dk), _ptr(dv), _ptr(dp), _stream()))
    return dq, dk, dv, dp


def ffm_fwd(tile: torch.Tensor) -> torch.Tensor:
    """FFM second-order logit (B,1) from the (B, F, F-1, K) field/slot tile (see include/ctr_b200.h)."""
    B, F, S, K = tile.shape
    if S != F - 1:
        raise ValueError(f"tile: expected shape (B, F, F-1, K), got {tuple(tile.shape)}")
    _chk(tile, F32, "tile")
    out = torch.empty((B, 1), dtype=F32, device=tile.device)
    _lib.check(_lib.lib().ctr_ffm_fwd(_ptr(tile), B, F, K, _ptr(out), _stream()))
    return out


def ffm_bwd(tile: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    B, F, S, K = tile.shape
    g = g.reshape(B)
    _chk(tile, F32, "tile"); _chk(g, F32, "g", (B,))
    d_tile = torch.empty_like(tile)
    _lib.check(_lib.lib().ctr_ffm_bwd(_ptr(tile), _ptr(g), B, F, K, _ptr(d_tile), _stream()))
    return d_tile
