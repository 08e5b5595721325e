"""GPU parity: DIN attention (row DIN-ATT) through the C ABI vs golden vectors (reference source executed) and oracle."""
import numpy as np
import pytest
import torch

from _util import TOL, assert_close, dev, golden, trunc_normal
from oracle import layers_np as O

pytestmark = pytest.mark.gpu
KEYS = ("f1_att_kernel", "f1_att_bias", "f2_att_kernel", "f2_att_bias", "f3_att_kernel", "f3_att_bias")


@pytest.mark.parametrize("name", ["din_T3_smoke", "din_T1", "din_T50"])
@pytest.mark.parametrize("soft", [False, True])
def test_din_golden(name, soft):
    from recalgorithm_b200 import ops
    g = golden(name)
    params = [dev(g[k]) for k in KEYS]
    out = ops.din_attention_fwd(dev(g["query"]), dev(g["keys"]), dev(g["keys_length"]), *params, is_softmax=soft)
    ref = g[f"out_softmax{int(soft)}_f64"]
    assert torch.isfinite(out).all()
    scale = max(np.abs(ref).max(), 1e-3)
    assert np.abs(out.cpu().numpy() - ref).max() <= TOL * scale
    if not soft:     # keys_length == 0 -> exactly zero (the edge the reference's own smoke encodes)
        assert torch.all(out[dev(g["keys_length"]) == 0] == 0)


def make(rng, B, T, H, lens=None):
    q = trunc_normal(rng, (B, H), 0.5)
    k = trunc_normal(rng, (B, T, H), 0.5)
    lens = rng.integers(0, T + 1, size=B).astype(np.int64) if lens is None else np.asarray(lens, dtype=np.int64)
    ws = [trunc_normal(rng, s, sd) for s, sd in (((4 * H, 64), 0.2), ((64,), 0.1), ((64, 32), 0.2), ((32,), 0.1),
                                                  ((32, 1), 0.3), ((1,), 0.1))]
    return q, k, lens, ws


@pytest.mark.parametrize("B,T,H", [(2, 3, 4), (17, 50, 16), (64, 50, 16), (5, 1, 16), (9, 7, 8), (6, 33, 32), (3, 20, 5),
                                   (130, 12, 16)])
@pytest.mark.parametrize("soft", [False, True])
def test_din_fwd_bwd(B, T, H, soft):
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(B * 100 + T + H)
    q, k, lens, ws = make(rng, B, T, H)
    lens[0] = 0
    lens[-1] = T
    d = lambda a: a.astype(np.float64)
    out, att = ops.din_attention_fwd(dev(q), dev(k), dev(lens), *[dev(w) for w in ws], is_softmax=soft, want_weights=True)
    ref, cache = O.din_attention_fwd(d(q), d(k), lens, *[d(w) for w in ws], is_softmax=soft, return_cache=True)
    assert_close(out, ref, TOL, "out")
    assert_close(att, cache["w"][..., 0], TOL, "attention weights")
    g = trunc_normal(rng, (B, H), 1.0)
    dq, dk, dws = ops.din_attention_bwd(dev(q), dev(k), dev(lens), *[dev(w) for w in ws], dev(g), is_softmax=soft)
    gr = O.din_attention_bwd(d(q), d(k), lens, *[d(w) for w in ws], d(g), is_softmax=soft)
    assert_close(dq, gr["query"], TOL, "d_query")
    assert_close(dk, gr["keys"], TOL, "d_keys")
    for got, name in zip(dws, ("w1", "b1", "w2", "b2", "w3", "b3")):
        ref = gr[name].reshape(got.shape)
        # d/db3 is analytically 0 under softmax (shift invariance): judge it on the scale of its sibling dw3
        scale = max(np.abs(ref).max(), 1e-2, np.abs(gr["w3"]).max() if name == "b3" else 0.0)
        assert np.abs(got.cpu().double().numpy() - ref).max() <= TOL * scale, name
    # static round-robin schedule (sched_scratch = NULL): same per-sample results bit for bit
    out_s = ops.din_attention_fwd(dev(q), dev(k), dev(lens), *[dev(w) for w in ws], is_softmax=soft, balanced=False)
    dq_s, dk_s, dws_s = ops.din_attention_bwd(dev(q), dev(k), dev(lens), *[dev(w) for w in ws], dev(g), is_softmax=soft, balanced=False)
    assert torch.equal(out_s, out) and torch.equal(dq_s, dq) and torch.equal(dk_s, dk)
    assert_close(dws_s[2], dws[2], TOL, "dw2 (static vs balanced schedule)")


def test_din_config4_properties():
    """BASELINE config 4 (B=4096, T=50, H=16): masked positions are ignored; output invariant to padded keys."""
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(44)
    B, T, H = 4096, 50, 16
    q, k, lens, ws = make(rng, B, T, H)
    params = [dev(w) for w in ws]
    for soft in (False, True):
        o1 = ops.din_attention_fwd(dev(q), dev(k), dev(lens), *params, is_softmax=soft)
        k2 = k.copy()
        mask = np.arange(T)[None, :] >= lens[:, None]
        k2[mask] = 123.0                                    # garbage in the padded positions
        o2 = ops.din_attention_fwd(dev(q), dev(k2), dev(lens), *params, is_softmax=soft)
        rows = lens > 0                                     # (len == 0 under softmax averages the padded keys, like the reference)
        assert torch.equal(o1[dev(rows)], o2[dev(rows)])
        ref = O.din_attention_fwd(q.astype(np.float64), k.astype(np.float64), lens, *[w.astype(np.float64) for w in ws],
                                  is_softmax=soft)
        assert_close(o1, ref, TOL, f"config-4 forward soft={soft}")


def test_din_empty_history():
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(1)
    q, k, lens, ws = make(rng, 4, 0, 16)
    out = ops.din_attention_fwd(dev(q), dev(k), dev(lens), *[dev(w) for w in ws])
    assert torch.all(out == 0)
