"""CPU tests of the drop-in boundary's HOST side (recalgorithm_b200.layers): the reference-signature wrappers create
the reference's variables (names, shapes, scopes, initialisers) and keep its error behaviour.  The kernel-launching autograd
functions are replaced by recorders inside these tests; numerics live in the -m gpu tests."""
import math

import pytest
import torch

from recalgorithm_b200 import autograd, layers as L


@pytest.fixture()
def store(monkeypatch):
    calls = []

    def rec(name, out):
        def f(*a, **k):
            calls.append((name, a, k))
            return out(*a, **k)
        return f
    monkeypatch.setattr(autograd, "cross_stack", rec("cross_stack", lambda x0, w, b, xl=None: x0 + 0))
    monkeypatch.setattr(autograd, "cin", rec("cin", lambda x0, xk, f, want_pooled=False: torch.zeros(x0.shape[0], f.shape[1], x0.shape[2])))
    monkeypatch.setattr(autograd, "din_attention", rec("din", lambda q, k, n, *p, is_softmax=False: q))
    monkeypatch.setattr(autograd, "senet", rec("senet", lambda x, w1, w2: x))
    monkeypatch.setattr(autograd, "bilinear", rec("bilinear", lambda x, w, t: x))
    st = L.set_default_store(L.VariableStore(device="cpu", seed=0))
    st.calls = calls
    yield st
    L.set_default_store(L.VariableStore(device="cpu"))


def test_cross_layer_creates_the_references_variables(store):
    """DCN/cross_layer.py:18-19 inside `with tf.variable_scope("cross_part")` (dcn.py:156): wl_i, bl_i of shape (d,1), BOTH
    default (glorot-uniform) initialised -- the bias is not zero."""
    x0 = torch.randn(5, 12)
    with L.variable_scope("cross_part"):
        xl = x0
        for i in range(3):
            xl = L.cross_layer(x0=x0, xl=xl, index=i)
    assert sorted(store.vars) == sorted(f"cross_part/{n}_{i}" for i in range(3) for n in ("wl", "bl"))
    assert all(tuple(v.shape) == (12, 1) for v in store.vars.values())
    lim = math.sqrt(6.0 / (12 + 1))
    for v in store.vars.values():
        assert 0 < float(v.detach().abs().max()) <= lim                           # glorot-uniform over (fan_in=d, fan_out=1)
    # the first layer's xl IS x0 -> the kernel is told so (xl=None); later layers pass their own xl
    assert store.calls[0][2]["xl"] is None and store.calls[1][2]["xl"] is not None
    with L.variable_scope("cross_part"):
        with pytest.raises(ValueError, match="Trying to share variable cross_part/wl_0"):
            L.cross_layer(torch.randn(5, 7), torch.randn(5, 7), 0)               # tf.get_variable: same name, other shape
    L.cross_network(torch.randn(2, 12), 0)                                        # L = 0: no variable, no launch


def test_cin_layer_takes_string_widths_and_names_its_filter(store):
    x0, xk = torch.randn(3, 6, 4), torch.randn(3, 5, 4)
    with L.variable_scope("cin_part"):
        out = L.cin_layer(x0, xk, "7", 2)                                         # widths arrive as strings (xdeepfm.py:253)
    assert out.shape == (3, 7, 4)
    assert tuple(store.vars["cin_part/cin_layer_2_filter"].shape) == (1, 5 * 6, 7)
    assert tuple(store.calls[-1][1][2].shape) == (30, 7)                          # the kernel gets the (hk*m, H) matrix


def test_din_attention_variables_are_auto_reused(store):
    q, k, n = torch.randn(4, 8), torch.randn(4, 6, 8), torch.tensor([0, 1, 6, 3], dtype=torch.int32)
    with L.variable_scope("attention_part"):
        L.din_attention(q, k, n)
        L.din_attention(q, k, n, is_softmax=True)                                 # second call: AUTO_REUSE, nothing new
    want = {"attention_part/f1_att/kernel": (32, 64), "attention_part/f1_att/bias": (64,), "attention_part/f2_att/kernel": (64, 32),
            "attention_part/f2_att/bias": (32,), "attention_part/f3_att/kernel": (32, 1), "attention_part/f3_att/bias": (1,)}
    assert {k_: tuple(v.shape) for k_, v in store.vars.items()} == want
    assert all(float(store.vars[k_].detach().abs().max()) == 0 for k_ in want if k_.endswith("bias"))   # tf.layers.dense: zero bias
    assert store.calls[-1][2]["is_softmax"] is True and store.calls[-1][1][2].dtype == torch.int64   # lengths widened to int64


def test_fibinet_wrappers_keep_the_references_quirks_and_errors(store):
    x = torch.randn(2, 8, 16)
    L.senet(x, embedding_dim=16, reduction_ratio=2)                               # reduction from K, not from F (senet.py:18)
    assert tuple(store.vars["senet_w1"].shape) == (8, 8) and tuple(store.vars["senet_w2"].shape) == (8, 8)
    with pytest.raises(AssertionError):
        L.senet(x, embedding_dim=16, reduction_ratio=1)                           # senet.py:19
    for t, shape in (("all", (16, 16)), ("each", (7, 16, 16)), ("interaction", (28, 16, 16))):
        L.bilinear_interaction_layer(x, 16, t, "bi")
        assert tuple(store.vars[f"bi_w_{t}"].shape) == shape
    with pytest.raises(ValueError, match="Bilinear Interaction type must be in"):
        L.bilinear_interaction_layer(x, 16, "nope", "bi")                          # bilinear_interaction_layer.py:36-38


def test_assign_injects_weights_by_tf_name(store):
    with L.variable_scope("cross_part"):
        L.cross_layer(torch.randn(2, 3), torch.randn(2, 3), 0)
    store.assign({"cross_part/wl_0": [[1.0], [2.0], [3.0]], "brand/new": torch.ones(2)})
    assert store.vars["cross_part/wl_0"].flatten().tolist() == [1.0, 2.0, 3.0] and "brand/new" in store.vars
    with pytest.raises(ValueError):
        store.assign({"cross_part/wl_0": torch.zeros(4, 1)})
    assert len(store.parameters()) == 3
