"""GPU: BASELINE config 1 plumbing -- DeepFM on a (synthetic) wechat_algo_data1 TFRecord, batch 256 -- through the
reference-shaped host API: TFRecord -> parse_example -> feature columns -> lookup + FM2 + first-order term, vs the oracle."""
import numpy as np
import pytest
import torch

from _util import TOL, assert_close, dev
from oracle import layers_np as O
from test_io import wechat_record

pytestmark = pytest.mark.gpu
CATS = ["userid", "feedid", "device", "authorid", "bgm_song_id", "bgm_singer_id"]      # DeepFM/deepfm.py:56-64, declaration order


def test_deepfm_config1_pipeline(tmp_path):
    from recalgorithm_b200 import autograd, feature_column as fc, io as cio, layers as L
    from recalgorithm_b200.io import tfrecord
    rng = np.random.default_rng(5)
    B, K = 256, 8
    path = str(tmp_path / "train.tfrecord")
    cio.write_records(path, [wechat_record(rng, i)[0] for i in range(B + 10)])
    vocabs = {k: cio.VocabularyFile([f"{k}_{i}".encode() for i in rng.permutation(60)]) for k in CATS}
    L.set_default_store(L.VariableStore(device="cuda", seed=3))
    cat_cols = {k: fc.categorical_column_with_vocabulary_file(k, vocabs[k]) for k in CATS}
    first_order = [fc.indicator_column(c) for c in cat_cols.values()]
    second_order = [fc.embedding_column(c, K) for c in cat_cols.values()]
    label = fc.numeric_column("read_comment", default_value=0.0)
    spec = fc.make_parse_example_spec(first_order + second_order + [label])
    batch = next(tfrecord.batches(cio.read_records(path), B))                         # batch, then parse
    features = cio.parse_example(batch, spec)
    labels = features.pop("read_comment")
    assert labels.shape == (B, 1)

    ctx = fc.LookupContext()
    with L.variable_scope("fm_first_order"):
        fm_first_order_logit = fc.indicator_dense(features, first_order, units=1, name="fm_first_order_dense")
    fields_embeddings = []
    for col in second_order:                                                           # DeepFM/deepfm.py:187-190
        fields_embeddings.append(fc.input_layer(features, [col], ctx=ctx))
    e = torch.stack(fields_embeddings, dim=1)                                          # (B, F, K)
    # the native feeder (libctr_feed.so) produces the same features -> bit-identical embeddings and first-order term
    from recalgorithm_b200.io import native
    buf, roff, rlen = native.read_tfrecord_file(path)
    feats_n = fc.parse_example_native(buf, roff[:B], rlen[:B], first_order + second_order + [label])
    assert np.array_equal(feats_n.pop("read_comment"), labels)
    e_n = torch.stack([fc.input_layer(feats_n, [col]) for col in second_order], dim=1)
    assert torch.equal(e_n, e.detach())
    with L.variable_scope("fm_first_order"):
        assert torch.equal(fc.indicator_dense(feats_n, first_order, units=1, name="fm_first_order_dense").detach(), fm_first_order_logit.detach())

    # ---- oracle on the same ids / weights
    st = L.default_store()
    ids = np.stack([vocabs[k].lookup(features[k][0]) for k in CATS], axis=1)
    assert (ids == -1).any(), "the synthetic data must exercise the OOV ('') path"
    tables = [st.vars[f"input_layer/{k}_embedding/embedding_weights"].detach().cpu().numpy() for k in CATS]
    off = np.concatenate([[0], np.cumsum([t.shape[0] for t in tables])]).astype(np.int64)
    e_ref = O.embedding_lookup(np.concatenate(tables), ids, off)
    assert np.array_equal(e.detach().cpu().numpy(), e_ref), "lookup through input_layer must be exact"
    fm2_ref = O.fm2_fwd(e_ref.astype(np.float64))
    fm2_torch = 0.5 * (e.sum(1).pow(2) - e.pow(2).sum(1)).sum(1, keepdim=True)
    assert_close(fm2_torch, fm2_ref, TOL, "fm2 from input_layer embeddings")
    # first-order term: kernel blocks are in column NAME order
    kern = st.vars["fm_first_order/fm_first_order_dense/kernel"].detach().cpu().numpy()[:, 0]
    order = sorted(CATS, key=lambda k: f"{k}_indicator")
    koff = np.concatenate([[0], np.cumsum([len(vocabs[k]) for k in order])])
    want = np.zeros((B, 1))
    for f, k in enumerate(order):
        col_ids = vocabs[k].lookup(features[k][0])
        want[:, 0] += np.where(col_ids >= 0, kern[koff[f] + np.maximum(col_ids, 0)], 0.0)
    assert_close(fm_first_order_logit, want, TOL, "first-order (indicator @ dense(1))")

    # ---- fused kernel on the same weights gives the same tile and logit
    et = autograd.EmbeddingTables([t.shape[0] for t in tables], K, device="cuda", init=None)
    et.weight.copy_(torch.from_numpy(np.concatenate(tables)))
    tile, fm2 = autograd.lookup_fm2(et, dev(ids))
    assert torch.equal(tile, e.detach())
    assert_close(fm2, fm2_ref, TOL, "fused fm2")

    # ---- backward through the column API: IndexedSlices == oracle dense gradient
    total_logit = fm_first_order_logit + fm2_torch
    loss = torch.nn.functional.binary_cross_entropy_with_logits(total_logit, dev(labels))
    loss.backward()
    g = (torch.sigmoid(total_logit.detach()) - dev(labels)).cpu().double().numpy() / B    # dL/dlogit
    dense_ref = O.embedding_lookup_bwd_dense(int(off[-1]), ids, off, O.fm2_bwd(e_ref.astype(np.float64), g[:, 0]))
    got = ctx.to_dense()
    for f, k in enumerate(CATS):
        p = st.vars[f"input_layer/{k}_embedding/embedding_weights"]
        assert_close(got[id(p)], dense_ref[off[f]:off[f + 1]], 2e-5, f"d table {k}")


def test_sequence_input_layer_and_shared_embeddings():
    from recalgorithm_b200 import feature_column as fc, io as cio, layers as L
    L.set_default_store(L.VariableStore(device="cuda", seed=4))
    vocab = cio.VocabularyFile([f"feedid_{i}".encode() for i in range(20)])
    feed = fc.categorical_column_with_vocabulary_file("feedid", vocab)
    seq = fc.sequence_categorical_column_with_vocabulary_file("his_read_comment_7d_seq", vocab)
    target_col, seq_col = fc.shared_embedding_columns([feed, seq], 16, combiner="mean")      # DIN/din.py:103-114
    features = {"feedid": ([b"feedid_3", b"feedid_99", b"feedid_0"], np.array([0, 1, 2, 3])),
                "his_read_comment_7d_seq": ([b"feedid_1", b"feedid_2", b"feedid_77", b"feedid_5"], np.array([0, 3, 3, 4]))}
    target, _ = fc.sequence_input_layer(features, [target_col])                                # (B, 1, H)
    hist, lens = fc.sequence_input_layer(features, [seq_col])                                  # (B, T, H), (B,)
    W = L.default_store().vars["input_layer/feedid_his_read_comment_7d_seq_shared_embedding/embedding_weights"]
    assert len(L.default_store().vars) == 1, "one shared table"
    assert target.shape == (3, 1, 16) and hist.shape == (3, 3, 16) and lens.tolist() == [3, 0, 1]
    assert torch.equal(target[0, 0], W[3]) and torch.all(target[1] == 0) and torch.equal(target[2, 0], W[0])
    assert torch.equal(hist[0, 0], W[1]) and torch.equal(hist[0, 1], W[2]) and torch.all(hist[0, 2] == 0)   # OOV step -> zeros
    assert torch.all(hist[1] == 0) and torch.equal(hist[2, 0], W[5]) and torch.all(hist[2, 1:] == 0)
