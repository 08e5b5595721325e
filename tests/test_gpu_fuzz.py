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
