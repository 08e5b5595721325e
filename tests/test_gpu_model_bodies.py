"""GPU: the five reference model_fn bodies ported onto the host API (examples/model_bodies.py) run forward + backward; gradients
reach the embedding tables as IndexedSlices and every variable of the store."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))


def _setup(F, D, rows=50, B=64, seed=0):
    from recalgorithm_b200 import autograd, layers as L
    store = L.set_default_store(L.VariableStore(device="cuda", seed=seed))
    gen = torch.Generator(device="cuda").manual_seed(seed)
    tables = autograd.EmbeddingTables([rows] * F, D, device="cuda")
    ids = torch.randint(-1, rows, (B, F), device="cuda", generator=gen)
    dense_input = torch.randn((B, 5), device="cuda", generator=gen)
    y = (torch.rand((B, 1), device="cuda", generator=gen) < 0.3).float()
    return store, tables, ids, dense_input, y


def _check(logit, y, store, tables, expect_vars):
    assert logit.shape == y.shape and torch.isfinite(logit).all()
    torch.nn.functional.binary_cross_entropy_with_logits(logit, y).backward()
    assert tables.grad_slices and all(torch.isfinite(s.values).all() for s in tables.grad_slices)
    assert any(float(s.values.abs().max()) > 0 for s in tables.grad_slices)
    names = set(store.vars)
    assert expect_vars <= names, expect_vars - names
    for n, v in store.vars.items():
        assert v.grad is not None and torch.isfinite(v.grad).all(), n


def test_deepfm_body():
    import model_bodies as M
    from recalgorithm_b200 import layers as L
    store, tables, ids, dense_input, y = _setup(F=6, D=8)
    first = torch.zeros((ids.shape[0], 1), device="cuda")
    _check(M.deepfm_logit(tables, ids, first), y, store, tables, {"fm_deep/dense/kernel", "fm_deep/deep_logit/bias"})


def test_dcn_body():
    import model_bodies as M
    from recalgorithm_b200 import autograd
    store, tables, ids, dense_input, y = _setup(F=7, D=16)
    cat = autograd.lookup(tables, ids).reshape(ids.shape[0], -1)
    _check(M.dcn_logit(dense_input, cat, num_cross_layer=3), y, store, tables,
           {"cross_part/wl_0", "cross_part/bl_2", "dnn_part/dnn_dense_1/kernel", "output_part/dense/kernel"})
    assert tuple(store.vars["cross_part/wl_0"].shape) == (5 + 7 * 16, 1)


def test_xdeepfm_body():
    import model_bodies as M
    from recalgorithm_b200 import autograd
    store, tables, ids, dense_input, y = _setup(F=8, D=16)
    x0 = autograd.lookup(tables, ids)
    _check(M.xdeepfm_logit(dense_input, x0, cin_layer_feature_maps=("16", "24")), y, store, tables,
           {"cin_part/cin_layer_1_filter", "cin_part/cin_layer_2_filter", "cin_part/dense/kernel", "linear_part/dense/bias"})
    assert tuple(store.vars["cin_part/cin_layer_2_filter"].shape) == (1, 16 * 8, 24)
    assert "cin_part/dense/bias" not in store.vars                                        # use_bias=False (xdeepfm.py:175)


@pytest.mark.parametrize("use_softmax", [False, True])
def test_din_body(use_softmax):
    import model_bodies as M
    from recalgorithm_b200 import autograd
    store, tables, ids, dense_input, y = _setup(F=4, D=16)
    B, T, H = ids.shape[0], 12, 16
    gen = torch.Generator(device="cuda").manual_seed(9)
    item = autograd.EmbeddingTables([80], H, device="cuda")                                # the shared feedid table (din.py:103-114)
    lens = torch.randint(0, T + 1, (B,), device="cuda", generator=gen)
    hist = torch.randint(0, 80, (B, T), device="cuda", generator=gen)
    hist[torch.arange(T, device="cuda")[None, :] >= lens[:, None]] = -1
    tgt = torch.randint(0, 80, (B, 1), device="cuda", generator=gen)
    keys = autograd.lookup(item, hist.reshape(B * T, 1)).reshape(B, T, H)
    target = autograd.lookup(item, tgt).reshape(B, H)
    cat = autograd.lookup(tables, ids).reshape(B, -1)
    logit, att = M.din_logit(dense_input, cat, target, keys, lens, use_softmax=use_softmax)
    assert att.shape == (B, H)
    _check(logit, y, store, tables, {"attention_part/f1_att/kernel", "attention_part/f3_att/bias", "fcn/dense/kernel"})
    assert len(item.grad_slices) == 2                                                      # history + target share one table


@pytest.mark.parametrize("typ", ["all", "each", "interaction"])
def test_fibinet_body(typ):
    import model_bodies as M
    from recalgorithm_b200 import autograd
    store, tables, ids, dense_input, y = _setup(F=6, D=8)
    x = autograd.lookup(tables, ids)
    _check(M.fibinet_logit(dense_input, x, embedding_dim=8, reduction_ratio=2, bilinear_interaction_type=typ), y, store, tables,
           {"senet_part/senet_w1", f"bilinear_interaction_part/orginal_w_{typ}", f"bilinear_interaction_part/senet_w_{typ}"})
