"""GPU: hypothesis-driven shape/id fuzzing of the lookup + FM2 pair and the cross stack against the oracle."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from _util import TOL, assert_close, dev, trunc_normal
from oracle import layers_np as O

pytestmark = pytest.mark.gpu
COMMON = dict(deadline=None, max_examples=40, suppress_health_check=[HealthCheck.too_slow], derandomize=True)


@settings(**COMMON)
@given(B=st.integers(1, 70), F=st.integers(1, 75), logD=st.integers(2, 7), seed=st.integers(0, 2 ** 16),
       p_bad=st.sampled_from([0.0, 0.1, 0.9, 1.0]))
def test_lookup_fm2_fuzz(B, F, logD, seed, p_bad):
    from recalgorithm_b200 import ops
    D = 1 << logD
    rng = np.random.default_rng(seed)
    rows = rng.integers(1, 9, size=F)
    off = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
    table = trunc_normal(rng, (int(off[-1]), D), 0.5)
    ids = np.stack([rng.integers(0, rows[f], size=B) for f in range(F)], 1).astype(np.int64)
    bad = rng.random((B, F)) < p_bad
    ids[bad] = np.where(rng.random(int(bad.sum())) < 0.5, -1, 10_000)            # OOV and out-of-range
    clean = np.where((ids < 0) | (ids >= rows[None, :]), -1, ids)
    e = O.embedding_lookup(table, clean, off)
    tile, fm2 = ops.embed_fm2_fwd(dev(table), dev(off), dev(ids))
    assert np.array_equal(tile.cpu().numpy(), e)
    ref = O.fm2_fwd(e.astype(np.float64))
    assert np.abs(fm2.cpu().double().numpy() - ref).max() <= TOL * max(np.abs(ref).max(), 1e-6)
    g = trunc_normal(rng, (B,), 1.0); dt = trunc_normal(rng, (B, F, D), 1.0)
    want = dt.astype(np.float64) + O.fm2_bwd(e.astype(np.float64), g.astype(np.float64))
    assert_close(ops.embed_fm2_bwd(tile, dev(dt), dev(g)), want, TOL, "row grads")


@settings(**COMMON)
@given(B=st.integers(1, 40), d=st.integers(1, 700), L=st.integers(1, 8), seed=st.integers(0, 2 ** 16), xl=st.booleans())
def test_cross_fuzz(B, d, L, seed, xl):
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(seed)
    x0 = trunc_normal(rng, (B, d), 0.5)
    ws = trunc_normal(rng, (L, d), (2.0 / d) ** 0.5); bs = trunc_normal(rng, (L, d), 0.1)
    g = trunc_normal(rng, (B, d), 1.0)
    f64 = lambda a: a.astype(np.float64)
    if not xl:
        assert_close(ops.cross_fwd(dev(x0), dev(ws), dev(bs)), O.cross_stack_fwd(f64(x0), f64(ws), f64(bs))[-1], TOL, "fwd")
        dx0, _, dw, db = ops.cross_bwd(dev(x0), dev(ws), dev(bs), dev(g))
        ex0, ew, eb = O.cross_stack_bwd(f64(x0), f64(ws), f64(bs), f64(g))
        assert_close(dx0, ex0, TOL, "dx0"); assert_close(dw, ew, TOL, "dw"); assert_close(db, eb, TOL, "db")
    else:
        start = trunc_normal(rng, (B, d), 0.5)
        t = lambda a: torch.tensor(f64(a), requires_grad=True)
        x0t, st_, wt, bt = t(x0), t(start), t(ws), t(bs)
        x = st_
        for l in range(L):
            x = x0t * (x @ wt[l])[:, None] + bt[l][None, :] + x
        x.backward(torch.tensor(f64(g)))
        assert_close(ops.cross_fwd(dev(x0), dev(ws), dev(bs), xl_in=dev(start)), x.detach(), TOL, "fwd(xl)")
        dx0, dxl, dw, db = ops.cross_bwd(dev(x0), dev(ws), dev(bs), dev(g), xl_in=dev(start))
        assert_close(dx0, x0t.grad, TOL, "dx0"); assert_close(dxl, st_.grad, TOL, "dxl")
        assert_close(dw, wt.grad, TOL, "dw"); assert_close(db, bt.grad, TOL, "db")


@settings(**{**COMMON, "max_examples": 25})
@given(B=st.integers(2, 40), F=st.integers(2, 45), K=st.integers(1, 40), seed=st.integers(0, 2 ** 16))
def test_fwfm_fuzz(B, F, K, seed):
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(seed)
    e = trunc_normal(rng, (B, F, K), 1.0); r = trunc_normal(rng, (F * (F - 1) // 2,), 0.5); g = trunc_normal(rng, (B,), 1.0)
    f64 = lambda a: a.astype(np.float64)
    ref = O.fwfm_fwd(f64(e), f64(r))
    scale = np.abs(f64(r)).sum() * 4.0 / max(F * (F - 1) // 2, 1) * K      # natural size of a sum of P pair terms (cancellation-proof)
    assert np.abs(ops.fwfm_fwd(dev(e), dev(r)).cpu().double().numpy() - ref).max() <= TOL * max(np.abs(ref).max(), scale)
    de, dr = ops.fwfm_bwd(dev(e), dev(r), dev(g))
    ede, edr = O.fwfm_bwd(f64(e), f64(r), f64(g))
    assert_close(de, ede, TOL, "d_tile")
    assert np.abs(dr.cpu().double().numpy() - edr).max() <= TOL * max(np.abs(edr).max(), float(B) ** 0.5 * K * 0.5)


@settings(**{**COMMON, "max_examples": 25})
@given(B=st.integers(1, 24), F=st.integers(2, 20), K=st.sampled_from([4, 8, 16, 32]), T=st.integers(1, 64), seed=st.integers(0, 2 ** 16))
def test_afm_fuzz(B, F, K, T, seed):
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(seed)
    e = trunc_normal(rng, (B, F, K), 1.0)
    w = trunc_normal(rng, (K, T), 0.4); b = trunc_normal(rng, (T,), 0.3); h = trunc_normal(rng, (T, 1), 0.4)
    g = trunc_normal(rng, (B, K), 1.0)
    f64 = lambda a: a.astype(np.float64)
    assert_close(ops.afm_fwd(dev(e), dev(w), dev(b), dev(h)), O.afm_fwd(f64(e), f64(w), f64(b), f64(h)), TOL, "pooled")
    got = ops.afm_bwd(dev(e), dev(w), dev(b), dev(h), dev(g))
    want = O.afm_bwd(e, w, b, h, g)
    for name, a, x in zip(("d_tile", "d_w", "d_b", "d_h"), got, want):
        assert np.abs(a.cpu().double().numpy() - x).max() <= TOL * max(np.abs(x).max(), 0.05 * (B * F) ** 0.5), name


@settings(**{**COMMON, "max_examples": 20})
@given(B=st.integers(1, 6), T=st.integers(1, 40), d=st.sampled_from([4, 8, 16]), H=st.integers(1, 4), extra=st.integers(0, 3),
       pos=st.booleans(), seed=st.integers(0, 2 ** 16))
def test_bst_fuzz(B, T, d, H, extra, pos, seed):
    from oracle import bst_torch
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(seed)
    maxlen = T + extra
    q, k, v = (trunc_normal(rng, (B, T, d), 1.0) for _ in range(3))
    lengths = rng.integers(0, T + 1, size=B).astype(np.int64)
    p = {n: trunc_normal(rng, s, 0.4) for n, s in O.bst_param_shapes(d, H, maxlen).items()}
    p["ln1_gamma"] = 1 + p["ln1_gamma"]; p["ln2_gamma"] = 1 + p["ln2_gamma"]
    g = trunc_normal(rng, (B, T, d), 1.0)
    want, dq, dk, dv, dp = bst_torch.bst_transformer_bwd(q, k, v, lengths, p, H, g, use_position_embedding=pos)
    packed = ops.bst_pack_params({n: dev(a) for n, a in p.items()}, d, H, maxlen)
    assert_close(ops.bst_transformer_fwd(dev(q), dev(k), dev(v), dev(lengths), packed, H, maxlen, pos), want, TOL, "fwd")
    gq, gk, gv, gp = ops.bst_transformer_bwd(dev(q), dev(k), dev(v), dev(lengths), packed, dev(g), H, maxlen, pos)
    # gradients are compared on the scale of the largest one (a layer norm over T*d elements makes some of them tiny)
    top = max(np.abs(x).max() for x in (dq, dk, dv))
    for name, a, x in (("dq", gq, dq), ("dk", gk, dk), ("dv", gv, dv)):
        assert np.abs(a.cpu().double().numpy() - x).max() <= 2 * TOL * top, name
    got = ops.bst_unpack_params(gp, d, H, maxlen)
    ptop = max(np.abs(x).max() for x in dp.values())
    for n in O.BST_PARAM_ORDER:
        assert np.abs(got[n].cpu().double().numpy() - dp[n]).max() <= 2 * TOL * ptop, n
