"""Adam on embedding tables with IndexedSlices gradients -- the step right after the hot path.

``TableAdam(tables, lr, lazy=False)`` reproduces ``tf.train.AdamOptimizer(lr, 0.9, 0.999, 1e-8)`` applied to a table
whose gradient is IndexedSlices (DeepFM/deepfm.py:246-250): duplicates summed first, m/v decayed and the variable
updated for EVERY row each step [TF-internal, SURVEY A.8].  ``lazy=True`` is DIEN's LazyAdamOptimizer
(DIEN/dien.py:328): only the referenced rows move.  De-duplication uses torch.unique (plumbing); the row sums, the
updates and the dense sweep are kernels of libctr_b200.so.
"""
from __future__ import annotations

import math

import torch

from . import _lib, autograd, ops


class TableAdam:
    def __init__(self, tables: "autograd.EmbeddingTables", lr: float, beta1: float = 0.9, beta2: float = 0.999,
                 eps: float = 1e-8, lazy: bool = False):
        self.tables, self.lr, self.b1, self.b2, self.eps, self.lazy = tables, lr, beta1, beta2, eps, lazy
        w = tables.weight
        self.m, self.v = torch.zeros_like(w), torch.zeros_like(w)
        self.t = 0
        self._bitmap = None if lazy else torch.zeros(((tables.num_rows + 31) // 32,), dtype=torch.int32, device=w.device)

    def step(self):
        """Consume tables.grad_slices (every backward since the last zero_grad) and apply one Adam step."""
        tb = self.tables
        w = tb.weight
        V, D = w.shape
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        L = _lib.lib()
        n_unique = 0
        if tb.grad_slices:
            rows_l, vals_l = [], []
            for sl in tb.grad_slices:
                n_rows = sl.field_row_offset[1:] - sl.field_row_offset[:-1]
                valid = (sl.ids >= 0) & (sl.ids < n_rows[None, :])
                rows_l.append((sl.ids + sl.field_row_offset[:-1][None, :])[valid])
                vals_l.append(sl.values[valid])
            rows = torch.cat(rows_l); vals = torch.cat(vals_l).contiguous()
            uniq, inv = torch.unique(rows, return_inverse=True)
            n_unique = int(uniq.numel())
            if n_unique:
                summed = torch.zeros((n_unique, D), dtype=torch.float32, device=w.device)
                _lib.check(L.ctr_rows_scatter_add(summed.data_ptr(), n_unique, D, inv.contiguous().data_ptr(), vals.data_ptr(),
                                                  None, int(vals.shape[0]), ops._stream()))
                if self._bitmap is not None:
                    self._bitmap.zero_()
                _lib.check(L.ctr_adam_rows(w.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), V, D, uniq.data_ptr(),
                                           summed.data_ptr(), None, n_unique, lr_t, self.b1, self.b2, self.eps,
                                           ops._ptr(self._bitmap), ops._stream()))
        if not self.lazy:
            if self._bitmap is not None and n_unique == 0:
                self._bitmap.zero_()
            _lib.check(L.ctr_adam_dense_rest(w.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), V, D, lr_t, self.b1, self.b2,
                                             self.eps, ops._ptr(self._bitmap), ops._stream()))
        tb.zero_grad()
        return n_unique
