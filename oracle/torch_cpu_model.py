"""torch-CPU op-for-op restatement of the reference model_fn slice that bench.py times as the CPU baseline
(TEST/BENCH INFRASTRUCTURE -- see oracle/__init__.py; TF 1.14 itself cannot be installed here).

Mirrors DeepFM/deepfm.py:178-235 for the hot path: one embedding gather PER FIELD (the reference calls
fc.input_layer once per column, :187-190), ``add_n`` / ``square`` chains for the FM second-order term
(:192-200), a dense(1) consumer of the concatenated embeddings standing in for the deep part's first use of
the tile (:203-212), ``add_n`` of the logits (:214) and mean sigmoid cross-entropy (:235); backward through
autograd with sparse (IndexedSlices-like) table gradients.  No optimizer step (the GPU arm has none either).
"""
from __future__ import annotations

import torch
import torch.nn.functional as Fn


class DeepFMLookupFM2CPU:
    def __init__(self, num_fields: int, dim: int, rows_per_field: int, seed: int = 1234):
        g = torch.Generator().manual_seed(seed)
        self.F, self.D = num_fields, dim
        self.tables = []
        block_rows = min(rows_per_field, 65536)
        for _ in range(num_fields):
            # a random block tiled to the full height: memcpy-speed init of multi-GB tables (values repeat every
            # 65536 rows, which changes nothing about the gather's memory behaviour)
            block = torch.empty((block_rows, dim)).uniform_(-dim ** -0.5, dim ** -0.5, generator=g)
            t = torch.empty((rows_per_field, dim))
            for r0 in range(0, rows_per_field, block_rows):
                n = min(block_rows, rows_per_field - r0)
                t[r0:r0 + n].copy_(block[:n])
            self.tables.append(t.requires_grad_())
        self.w_deep = (torch.randn((num_fields * dim, 1), generator=g) * 0.01).requires_grad_()

    def forward(self, ids: torch.Tensor):
        fields = [Fn.embedding(ids[:, f], self.tables[f], sparse=True) for f in range(self.F)]   # F x (B, K)
        squares = [torch.square(e) for e in fields]
        s = fields[0]
        for e in fields[1:]:
            s = s + e                                                  # tf.add_n, list order
        q = squares[0]
        for e in squares[1:]:
            q = q + e
        fm2 = torch.sum(0.5 * (torch.square(s) - q), dim=1, keepdim=True)
        deep = torch.cat(fields, dim=1) @ self.w_deep
        return fm2 + deep, fm2

    def step(self, ids: torch.Tensor, labels: torch.Tensor) -> float:
        for t in self.tables:
            t.grad = None
        self.w_deep.grad = None
        logit, _ = self.forward(ids)
        loss = Fn.binary_cross_entropy_with_logits(logit, labels)
        loss.backward()
        return float(loss.detach())
