"""GPU parity: the FM2 siblings of SURVEY 8f.4 -- NFM bi-interaction (fused with the lookup), FwFM pair-weighted FM2,
AFM attention pooling -- vs the reference-executed golden vectors (inline model_fn blocks) and the oracle."""
import numpy as np
import pytest
import torch

from _util import TOL, assert_close, dev, golden, trunc_normal
from oracle import layers_np as O

pytestmark = pytest.mark.gpu


def _table_from_tile(e):
    """A table whose row (f, b) is e[b, f, :] and the ids that gather the tile back (lets golden tiles drive the fused lookup)."""
    B, F, D = e.shape
    table = np.ascontiguousarray(e.transpose(1, 0, 2).reshape(F * B, D))
    off = np.arange(F + 1, dtype=np.int64) * B
    ids = np.tile(np.arange(B, dtype=np.int64)[:, None], (1, F))
    return table, off, ids


@pytest.mark.parametrize("name", ["fm2_ref_F6_D8", "fm2_ref_F40_D32"])
def test_fm2_reference_executed_golden(name):
    from recalgorithm_b200 import ops
    g = golden(name)
    table, off, ids = _table_from_tile(g["e"])
    tile, fm2 = ops.embed_fm2_fwd(dev(table), dev(off), dev(ids))
    assert torch.equal(tile.cpu(), torch.from_numpy(g["e"]))                 # gathered rows bit-exact
    assert_close(fm2, g["out_f64"], TOL, "fm2 vs deepfm.py:184-200 executed")


@pytest.mark.parametrize("name", ["nfm_bi_F6_D8", "nfm_bi_F40_D32"])
def test_bi_interaction_golden(name):
    from recalgorithm_b200 import ops
    g = golden(name)
    table, off, ids = _table_from_tile(g["e"])
    tile, bi = ops.embed_bi_fwd(dev(table), dev(off), dev(ids))
    assert torch.equal(tile.cpu(), torch.from_numpy(g["e"]))
    assert_close(bi, g["out_f64"], TOL, "bi-interaction vs nfm.py:155-168 executed")
    _, bi2 = ops.embed_bi_fwd(dev(table), dev(off), dev(ids), want_tile=False)
    assert torch.equal(bi, bi2)


@pytest.mark.parametrize("B,F,D", [(5, 1, 4), (33, 7, 8), (64, 40, 32), (9, 70, 16), (3, 6, 128), (130, 33, 64)])
def test_bi_interaction_fwd_bwd(B, F, D):
    from recalgorithm_b200 import autograd
    rng = np.random.default_rng(B * 7 + F + D)
    rows = 11
    tables = autograd.EmbeddingTables([rows] * F, D, device="cuda")
    ids = rng.integers(-1, rows, size=(B, F)).astype(np.int64)                 # OOV ids -> zero rows
    w = tables.weight.cpu().numpy()
    off = tables.field_row_offset.cpu().numpy()
    e = O.embedding_lookup(w, ids, off)
    tile, bi = autograd.lookup_bi(tables, dev(ids))
    assert torch.equal(tile.detach().cpu(), torch.from_numpy(e))
    assert_close(bi, O.bi_interaction_fwd(e.astype(np.float64)), TOL, "bi")
    d_tile = trunc_normal(rng, (B, F, D), 1.0); d_bi = trunc_normal(rng, (B, D), 1.0)
    ((tile * dev(d_tile)).sum() + (bi * dev(d_bi)).sum()).backward()
    want = d_tile.astype(np.float64) + O.bi_interaction_bwd(e.astype(np.float64), d_bi.astype(np.float64))
    assert_close(tables.grad_slices[0].values, want, TOL, "row grads")


@pytest.mark.parametrize("name", ["fwfm_F6_D8", "fwfm_F30_D16"])
def test_fwfm_golden(name):
    from recalgorithm_b200 import ops
    g = golden(name)
    assert_close(ops.fwfm_fwd(dev(g["e"]), dev(g["r"])), g["out_f64"], TOL, "fwfm vs fwfm.py:140-158 executed")


@pytest.mark.parametrize("B,F,K", [(16, 2, 4), (37, 7, 8), (64, 40, 32), (5, 33, 3), (200, 30, 16), (3, 65, 10)])
def test_fwfm_fwd_bwd(B, F, K):
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(B + 3 * F + K)
    e = trunc_normal(rng, (B, F, K), 1.0)
    r = trunc_normal(rng, (F * (F - 1) // 2,), 0.5)
    g = trunc_normal(rng, (B,), 1.0)
    d = lambda a: a.astype(np.float64)
    assert_close(ops.fwfm_fwd(dev(e), dev(r)), O.fwfm_fwd(d(e), d(r)), TOL, "fwd")
    de, dr = ops.fwfm_bwd(dev(e), dev(r), dev(g))
    ede, edr = O.fwfm_bwd(d(e), d(r), d(g))
    assert_close(de, ede, TOL, "d_tile"); assert_close(dr, edr, TOL, "d_r")


@pytest.mark.parametrize("name", ["afm_F5_D8_t4", "afm_F30_D16_t8"])
def test_afm_golden(name):
    from recalgorithm_b200 import ops
    g = golden(name)
    pooled, score = ops.afm_fwd(dev(g["e"]), dev(g["w"]), dev(g["b"]), dev(g["h"]), want_score=True)
    assert_close(pooled, g["pooled_f64"], TOL, "pooled vs afm.py:152-186 executed")
    assert_close(score, g["score_f64"][:, :, 0], TOL, "attention_score")
    assert_close(pooled @ dev(g["p"]), g["logit_f64"], TOL, "afm_logit")


# reference defaults: embedding_dim 8, attention_factor 128, 7 fields (AFM/afm.py:37-38)
@pytest.mark.parametrize("B,F,K,T", [(1, 2, 4, 1), (70, 7, 8, 128), (33, 30, 16, 8), (5, 12, 32, 40), (9, 9, 8, 200), (17, 6, 16, 100),
                                     (4, 40, 4, 256)])
def test_afm_fwd_bwd(B, F, K, T):
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(B + F + K + T)
    e = trunc_normal(rng, (B, F, K), 1.0)
    w = trunc_normal(rng, (K, T), 0.4); b = trunc_normal(rng, (T,), 0.3); h = trunc_normal(rng, (T, 1), 0.4)
    g = trunc_normal(rng, (B, K), 1.0)
    d = lambda a: a.astype(np.float64)
    assert_close(ops.afm_fwd(dev(e), dev(w), dev(b), dev(h)), O.afm_fwd(d(e), d(w), d(b), d(h)), TOL, "pooled")
    de, dw, db, dh = ops.afm_bwd(dev(e), dev(w), dev(b), dev(h), dev(g))
    ede, edw, edb, edh = O.afm_bwd(e, w, b, h, g)
    assert_close(de, ede, TOL, "d_tile"); assert_close(db, edb, TOL, "d_b")
    # d_w sums B*P ReLU-gated products with cancellation (3120 terms at F=40): its fp32 error scales with sum|terms|
    assert_close(dw, edw, TOL, "d_w", elementwise=3.0)
    assert_close(dh, edh, TOL, "d_h", elementwise=3.0)       # same batch-reduced, softmax-weighted sum


def test_pairwise_errors_and_layers_api():
    from recalgorithm_b200 import _lib, layers as L, ops
    e = torch.zeros((2, 5, 8), device="cuda")
    with pytest.raises(ValueError):
        ops.fwfm_fwd(e, torch.zeros((9,), device="cuda"))                       # r must have F(F-1)/2 = 10 entries
    with pytest.raises(_lib.CtrError):
        ops.afm_fwd(torch.zeros((2, 5, 12), device="cuda"), torch.zeros((12, 4), device="cuda"), torch.zeros(4, device="cuda"),
                    torch.zeros(4, device="cuda"))                               # K = 12 unsupported
    store = L.set_default_store(L.VariableStore(device="cuda", seed=3))
    x = torch.randn((4, 5, 8), device="cuda", requires_grad=True)
    logit = L.fwfm_second_order(x)
    pooled = L.afm_attention(x, embedding_dim=8, attention_factor=16)
    assert logit.shape == (4, 1) and pooled.shape == (4, 8)
    assert {k: tuple(v.shape) for k, v in store.vars.items()} == {
        "fields_pair_strength/fields_pair_strength_weight": (10,), "attention_part/attention_w": (8, 16),
        "attention_part/attention_b": (16,), "attention_part/attention_h": (16, 1)}
    (logit.sum() + pooled.sum()).backward()
    assert x.grad is not None and all(v.grad is not None for v in store.vars.values())


def test_full_size_bi_interaction_properties():
    """BASELINE config 5 shape: the bi-interaction vector sums (over D) to the FM2 logit of the fused FM2 kernel, the tile is
    bit-identical between the two kernels, backward is linear in its two upstream gradients."""
    from recalgorithm_b200 import ops
    B, F, D, rows = 65536, 40, 32, 2_500_000
    gen = torch.Generator(device="cuda").manual_seed(4321)
    table = torch.empty((rows * F, D), device="cuda").normal_(0, D ** -0.5, generator=gen)
    off = torch.arange(F + 1, device="cuda") * rows
    ids = torch.randint(0, rows, (B, F), device="cuda", generator=gen)
    ids[torch.rand((B, F), device="cuda", generator=gen) < 0.05] = -1
    tile, fm2 = ops.embed_fm2_fwd(table, off, ids)
    tile_b, bi = ops.embed_bi_fwd(table, off, ids)
    assert torch.equal(tile, tile_b)
    assert_close(bi.double().sum(1, keepdim=True), fm2.double(), TOL, "sum_D bi == fm2")
    d_tile = torch.randn((B, F, D), device="cuda", generator=gen); d_bi = torch.randn((B, D), device="cuda", generator=gen)
    r1 = ops.embed_bi_bwd(tile, d_tile, d_bi); r2 = ops.embed_bi_bwd(tile, None, d_bi)
    assert_close(r1 - r2, d_tile, 1e-5, "bwd linearity")
    ones = torch.ones((B, D), device="cuda")                              # d_bi = 1 reproduces the FM2 backward with g = 1
    assert_close(ops.embed_bi_bwd(tile, None, ones), ops.embed_fm2_bwd(tile, None, torch.ones((B,), device="cuda")), 1e-6, "bi bwd == fm2 bwd at g=1")


@pytest.mark.parametrize("name", ["ffm_F4_K4", "ffm_F5_K8", "ffm_F9_K16"])
def test_ffm_golden_through_the_lookup(name):
    """FFM/ffm.py:145-160 executed on per-field (F-1, |V|, K) variables vs: reference variables -> id-major table -> fused
    lookup -> ctr_ffm_fwd.  D = (F-1)*K is 12, 32 and 128 here (12 is not a power of two: the table rows are zero padded to 16)."""
    from recalgorithm_b200 import _lib, autograd, layers as L, ops
    g = golden(name)
    B, F = g["ids"].shape
    K = g["tile"].shape[-1]
    assert_close(ops.ffm_fwd(dev(g["tile"])), g["out_f64"], TOL, "ffm on the fixture tile")
    embs = [torch.from_numpy(g[f"emb_{f}"]) for f in range(F)]
    D = L.ffm_table_dim(F, K)
    assert D >= (F - 1) * K and D & (D - 1) == 0
    tables = autograd.EmbeddingTables([e.shape[1] for e in embs], D, device="cuda", init=None)
    tables.weight.copy_(L.ffm_table_from_reference(embs))
    assert_close(L.ffm_second_order(tables, dev(g["ids"]), K), g["out_f64"], TOL, "lookup + ffm vs ffm.py executed")
    # the reference's default shape (7 fields, K = 8 -> 48 floats per row, padded to 64): forward + IndexedSlices backward
    F7, K8, V = 7, 8, 11
    gen = torch.Generator().manual_seed(3)
    embs7 = [torch.randn((F7 - 1, V, K8), generator=gen) for _ in range(F7)]
    t7 = autograd.EmbeddingTables([V] * F7, L.ffm_table_dim(F7, K8), device="cuda", init=None)
    t7.weight.copy_(L.ffm_table_from_reference(embs7))
    ids7 = torch.randint(0, V, (33, F7), generator=gen).cuda()
    out7 = L.ffm_second_order(t7, ids7, K8)
    tile7 = torch.stack([torch.stack([embs7[f][:, ids7[b, f].item(), :] for f in range(F7)]) for b in range(33)]).numpy()
    assert_close(out7, O.ffm_fwd(tile7.astype(np.float64)), TOL, "FFM at the reference's default width (48 -> 64)")
    out7.sum().backward()
    assert t7.grad_slices[0].values.shape == (33, F7, 64) and float(t7.grad_slices[0].values[:, :, 48:].abs().max()) == 0.0


@pytest.mark.parametrize("B,F,K", [(3, 2, 4), (65, 5, 8), (17, 9, 16), (8, 7, 3), (40, 12, 10)])
def test_ffm_fwd_bwd(B, F, K):
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(B + F + K)
    t = trunc_normal(rng, (B, F, F - 1, K), 1.0); g = trunc_normal(rng, (B,), 1.0)
    d = lambda a: a.astype(np.float64)
    ref = O.ffm_fwd(d(t))
    assert np.abs(ops.ffm_fwd(dev(t)).cpu().double().numpy() - ref).max() <= TOL * max(np.abs(ref).max(), float(K))
    assert_close(ops.ffm_bwd(dev(t), dev(g)), O.ffm_bwd(d(t), d(g)), TOL, "d_tile")
