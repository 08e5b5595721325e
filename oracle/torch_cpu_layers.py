"""torch-CPU op-for-op restatements of the reference's lookup + interaction chains of BASELINE configs 2-4, timed by bench.py as
the CPU baseline of those configs (TEST/BENCH INFRASTRUCTURE -- see oracle/__init__.py; TF 1.14 itself cannot be installed).

Each class mirrors what the reference's graph does on the CPU, including what makes it slow there:
  * one embedding gather PER FIELD (`fc.input_layer` builds a sub-graph per column: DCN/dcn.py:153, xDeepFM/xdeepfm.py:158);
  * DCN/cross_layer.py:21-24 as written: `matmul(xl, wl)` -> `multiply(x0, .)` -> `add(., transpose(bl))` -> `add(., xl)`;
  * xDeepFM/cin_layer.py:21-28 as written: the (B, D, hk, m) outer product is MATERIALISED, reshaped to (B, D, hk*m) and pushed
    through a width-1 conv1d (a matmul over the last axis), then transposed; pooled sum over D (xdeepfm.py:173);
  * DIN/din_attention.py:18-41 as written: the query is tiled to (B, T, H), [q, k, q-k, q*k] is concatenated to (B, T, 4H),
    three dense layers run over EVERY position (padding included), then the mask, then `matmul(weights^T, keys)`.
A step is forward + backward (autograd; sparse table gradients = IndexedSlices) with the same upstream gradients the GPU
workloads of tools/workloads.py use; no dense tail, no optimizer step -- exactly what the GPU arm times.
"""
from __future__ import annotations

import torch
import torch.nn.functional as Fn


def _tables(num_fields, rows, dim, gen, std):
    """Per-field tables; a random 65536-row block tiled to the full height (memcpy-speed init; the gather's memory behaviour is
    that of the full table)."""
    out = []
    block_rows = min(rows, 65536)
    for _ in range(num_fields):
        block = torch.randn((block_rows, dim), generator=gen) * std
        t = torch.empty((rows, dim))
        for r0 in range(0, rows, block_rows):
            n = min(block_rows, rows - r0)
            t[r0:r0 + n].copy_(block[:n])
        out.append(t.requires_grad_())
    return out


class DCNCrossCPU:
    """Config 2: `net = input_layer(...)` (dcn.py:153) + the cross loop (dcn.py:157-160)."""

    def __init__(self, F=30, D=16, L=3, rows=1_000_000, seed=1234):
        g = torch.Generator().manual_seed(seed)
        self.F, self.D, self.L, self.rows = F, D, L, rows
        self.tables = _tables(F, rows, D, g, D ** -0.5)
        d = F * D
        self.wl = [(torch.randn((d, 1), generator=g) * 0.05).requires_grad_() for _ in range(L)]
        self.bl = [(torch.randn((d, 1), generator=g) * 0.05).requires_grad_() for _ in range(L)]

    def params(self):
        return self.tables + self.wl + self.bl

    def forward(self, ids):
        x0 = torch.cat([Fn.embedding(ids[:, f], self.tables[f], sparse=True) for f in range(self.F)], dim=1)   # (B, d)
        xl = x0
        for wl, bl in zip(self.wl, self.bl):                      # cross_layer(x0, xl, i)
            xl_wl = torch.matmul(xl, wl)                          # (B, 1)
            x0_xl_wl = torch.mul(x0, xl_wl)
            out = torch.add(x0_xl_wl, bl.t())
            xl = torch.add(out, xl)
        return xl

    def make_batch(self, B, gen):
        return (torch.randint(0, self.rows, (B, self.F), generator=gen), torch.randn((B, self.F * self.D), generator=gen))

    def step(self, ids, g_out):
        for p in self.params():
            p.grad = None
        (self.forward(ids) * g_out).sum().backward()


class XDeepFMCinCPU:
    """Config 3: lookup reshaped to (B, m, D) (xdeepfm.py:158,167) + CIN layers + the pooled sums (xdeepfm.py:166-175)."""

    def __init__(self, F=30, D=16, maps=(128, 128), rows=1_000_000, seed=1234):
        g = torch.Generator().manual_seed(seed)
        self.F, self.D, self.maps, self.rows = F, D, tuple(maps), rows
        self.tables = _tables(F, rows, D, g, D ** -0.5)
        self.filters, hk = [], F
        for h in self.maps:                                        # cin_layer_{i}_filter: (1, hk*m, hk_1)
            self.filters.append((torch.randn((1, hk * F, h), generator=g) * 0.05).requires_grad_())
            hk = h

    def params(self):
        return self.tables + self.filters

    @staticmethod
    def cin_layer(x0, xk, filt):
        B, m, D = x0.shape
        hk = xk.shape[1]
        outer = torch.einsum("bik,bjk->bkij", xk, x0)             # (B, D, hk, m), materialised like tf.einsum does
        outer = outer.reshape(B, D, hk * m)
        # tf.nn.conv1d(outer, filters (1, hk*m, hk_1), stride 1, VALID) over the D axis with width 1 == a matmul on the last axis
        xk_1 = Fn.conv1d(outer.transpose(1, 2), filt.permute(2, 1, 0)).transpose(1, 2)   # (B, D, hk_1)
        return xk_1.transpose(1, 2)                                # (B, hk_1, D)

    def forward(self, ids):
        x0 = torch.stack([Fn.embedding(ids[:, f], self.tables[f], sparse=True) for f in range(self.F)], dim=1)   # (B, m, D)
        xk, pooled = x0, []
        for filt in self.filters:
            xk = self.cin_layer(x0, xk, filt)
            pooled.append(xk.sum(-1))                              # tf.reduce_sum(x, axis=-1)
        return torch.cat(pooled, dim=1), xk                        # (B, sum maps)

    def make_batch(self, B, gen):
        return (torch.randint(0, self.rows, (B, self.F), generator=gen), torch.randn((B, sum(self.maps)), generator=gen))

    def step(self, ids, g_pooled):
        for p in self.params():
            p.grad = None
        (self.forward(ids)[0] * g_pooled).sum().backward()


class DINAttentionCPU:
    """Config 4: sequence_input_layer over the shared table (din.py:209-214) + din_attention (din.py:218), paper (non-softmax)
    weights by default."""

    def __init__(self, T=50, H=16, rows=1_000_000, seed=1234, is_softmax=False):
        g = torch.Generator().manual_seed(seed)
        self.T, self.H, self.rows, self.is_softmax = T, H, rows, is_softmax
        (self.table,) = _tables(1, rows + 1, H, g, 0.25)           # last row: the zero vector padding / OOV steps look up
        with torch.no_grad():
            self.table[rows].zero_()
        mk = lambda *s, std: (torch.randn(s, generator=g) * std).requires_grad_()   # noqa: E731
        self.w1, self.b1 = mk(4 * H, 64, std=0.2), mk(64, std=0.1)
        self.w2, self.b2 = mk(64, 32, std=0.2), mk(32, std=0.1)
        self.w3, self.b3 = mk(32, 1, std=0.3), mk(1, std=0.1)

    def params(self):
        return [self.table, self.w1, self.b1, self.w2, self.b2, self.w3, self.b3]

    def attention(self, query, keys, keys_length):
        B, T, H = keys.shape
        q = query.repeat(1, T).reshape(-1, T, H)                   # tf.tile + tf.reshape
        cross_all = torch.cat([q, keys, q - keys, q * keys], dim=-1)
        d1 = torch.relu(cross_all @ self.w1 + self.b1)
        d2 = torch.relu(d1 @ self.w2 + self.b2)
        w = d2 @ self.w3 + self.b3                                 # (B, T, 1)
        mask = (torch.arange(T)[None, :] < keys_length[:, None]).unsqueeze(-1)     # tf.sequence_mask
        if self.is_softmax:
            w = torch.where(mask, w, torch.full_like(w, float(-2 ** 32 + 1)))
            w = torch.softmax(w / (H ** 0.5), dim=1)
        else:
            w = w * mask.to(w.dtype)
        return torch.matmul(w.transpose(1, 2), keys).squeeze(1)    # (B, H)

    def forward(self, hist_ids, target_ids, keys_length):
        pad = torch.where(hist_ids >= 0, hist_ids, torch.full_like(hist_ids, self.rows))
        keys = Fn.embedding(pad, self.table, sparse=True)          # (B, T, H); padding -> the zero row
        query = Fn.embedding(target_ids[:, 0], self.table, sparse=True)
        return self.attention(query, keys, keys_length)

    def make_batch(self, B, gen):
        lens = torch.randint(0, self.T + 1, (B,), generator=gen)
        hist = torch.randint(0, self.rows, (B, self.T), generator=gen)
        hist[torch.arange(self.T)[None, :] >= lens[:, None]] = -1
        return (hist, torch.randint(0, self.rows, (B, 1), generator=gen), lens, torch.randn((B, self.H), generator=gen))

    def step(self, hist_ids, target_ids, keys_length, g_out):
        for p in self.params():
            p.grad = None
        (self.forward(hist_ids, target_ids, keys_length) * g_out).sum().backward()


BUILDERS = {"dcn_cfg2": (DCNCrossCPU, 4096), "xdeepfm_cfg3": (XDeepFMCinCPU, 8192), "din_cfg4": (DINAttentionCPU, 4096)}
