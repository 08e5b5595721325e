"""CPU tests of the HOST logic of recalgorithm_b200.feature_column (SURVEY 8a parity notes 1-3, Appendix A.3-A.6): column
ordering, output layout, variable names, shared tables, parse spec.  The one device call on this path (ops.bag_lookup_fwd ->
ctr_bag_lookup_fwd) is replaced by a small torch-CPU stand-in INSIDE THIS TEST ONLY, so that everything around the kernel can be
checked without a GPU; the kernel itself is covered by tests/test_gpu_embed_fm2.py / test_gpu_pipeline.py."""
import numpy as np
import pytest
import torch

from recalgorithm_b200 import feature_column as fc, io as cio, layers as L, ops


def _bag_lookup_cpu(table, ids, offsets, out=None, out_col=0):
    """What ctr_bag_lookup_fwd computes (mean over the valid ids of each bag, empty bag -> zeros), written to out[:, col:col+D]."""
    D = table.shape[1]
    B = offsets.numel() - 1
    for b in range(B):
        bag = ids[offsets[b]:offsets[b + 1]]
        bag = bag[(bag >= 0) & (bag < table.shape[0])]
        out[b, out_col:out_col + D] = table[bag].mean(0) if bag.numel() else 0.0
    return out


@pytest.fixture()
def host(monkeypatch):
    monkeypatch.setattr(ops, "bag_lookup_fwd", _bag_lookup_cpu)
    store = L.set_default_store(L.VariableStore(device="cpu", seed=1))
    yield store
    L.set_default_store(L.VariableStore(device="cpu"))


def _vocab(prefix, n):
    return cio.VocabularyFile([f"{prefix}_{i}".encode() for i in range(n)])


def test_input_layer_concatenates_columns_by_name_with_mixed_widths(host):
    """Parity notes 1 and 3: one input_layer call over columns declared in any order yields them sorted by column NAME
    (`<key>_embedding`, numeric: `<key>`), each with its own width (DCN's 16,16,2,4,... dims: dcn.py:97-103)."""
    user = fc.embedding_column(fc.categorical_column_with_vocabulary_file("userid", _vocab("userid", 9)), 4)
    dev = fc.embedding_column(fc.categorical_column_with_vocabulary_file("device", _vocab("device", 3)), 2)
    tags = fc.embedding_column(fc.categorical_column_with_vocabulary_file("manual_tag_list", _vocab("tag", 5)), 3)
    dense = fc.numeric_column("videoplayseconds")
    features = {"userid": ([b"userid_3", b"userid_8"], np.array([0, 1, 2])),
                "device": ([b"device_2"], np.array([0, 0, 1])),                       # first row: missing -> zeros
                "manual_tag_list": ([b"tag_0", b"tag_4", b"nope", b"tag_1"], np.array([0, 3, 4])),   # multi-valued, one OOV
                "videoplayseconds": np.array([[1.5], [2.5]], np.float32)}
    out = fc.input_layer(features, [user, dense, tags, dev], device="cpu")            # declaration order is NOT the layout
    assert out.shape == (2, 2 + 3 + 4 + 1)
    W = {k: v.detach() for k, v in host.vars.items()}
    assert set(W) == {"input_layer/userid_embedding/embedding_weights", "input_layer/device_embedding/embedding_weights",
                      "input_layer/manual_tag_list_embedding/embedding_weights"}
    d, t, u = (W[f"input_layer/{k}_embedding/embedding_weights"] for k in ("device", "manual_tag_list", "userid"))
    assert d.shape == (3, 2) and t.shape == (5, 3) and u.shape == (9, 4)
    # sorted names: device_embedding < manual_tag_list_embedding < userid_embedding < videoplayseconds
    assert torch.all(out[0, 0:2] == 0) and torch.equal(out[1, 0:2], d[2])
    assert torch.allclose(out[0, 2:5], (t[0] + t[4]) / 2) and torch.equal(out[1, 2:5], t[1])    # OOV pruned before the mean
    assert torch.equal(out[0, 5:9], u[3]) and torch.equal(out[1, 5:9], u[8])
    assert out[:, 9].tolist() == [1.5, 2.5]
    # embedding_column default initializer: truncated_normal(0, 1/sqrt(D)) (parity note 4)
    assert float(u.abs().max()) <= 2 * 4 ** -0.5 + 1e-6


def test_shared_embedding_columns_keep_input_order_and_share_one_table(host):
    """Parity note 2: returned in INPUT order (DIN relies on [0] = target, [1] = history, din.py:113-114); one variable named
    after the SORTED keys."""
    v = _vocab("feedid", 12)
    hist = fc.sequence_categorical_column_with_vocabulary_file("his_read_comment_7d_seq", v)
    feed = fc.categorical_column_with_vocabulary_file("feedid", v)
    a, b = fc.shared_embedding_columns([hist, feed], 4, combiner="mean")
    assert a.categorical_column is hist and b.categorical_column is feed
    assert a.variable_name == b.variable_name == "input_layer/feedid_his_read_comment_7d_seq_shared_embedding/embedding_weights"
    features = {"feedid": ([b"feedid_2"], np.array([0, 1])), "his_read_comment_7d_seq": ([b"feedid_2", b"feedid_5"], np.array([0, 2]))}
    with L.variable_scope("some_model_scope"):                                         # shared tables ignore the caller's scope
        x = fc.input_layer(features, [b], device="cpu")
        y = fc.input_layer(features, [a], device="cpu")
    assert list(host.vars) == [a.variable_name]
    Wt = host.vars[a.variable_name].detach()
    assert torch.equal(x[0], Wt[2]) and torch.allclose(y[0], (Wt[2] + Wt[5]) / 2)
    with pytest.raises(ValueError):
        fc.embedding_column(feed, 4, combiner="sum")


def test_parse_spec_and_refusals(host):
    """Appendix A.3: categorical (also wrapped) columns -> VarLenFeature(string) under the base key, once; numeric ->
    FixedLenFeature(shape, float32, default)."""
    u = fc.categorical_column_with_vocabulary_file("userid", _vocab("userid", 4))
    cols = [fc.embedding_column(u, 8), fc.indicator_column(u), fc.numeric_column("read_comment", default_value=0.0),
            fc.numeric_column("vec", shape=(3,), default_value=1.0)]
    spec = fc.make_parse_example_spec(cols)
    assert set(spec) == {"userid", "read_comment", "vec"}
    assert isinstance(spec["userid"], cio.VarLenFeature) and spec["vec"] == cio.FixedLenFeature((3,), "float", 1.0)
    with pytest.raises(ValueError, match="indicator"):
        fc.input_layer({"userid": ([b"userid_1"], np.array([0, 1]))}, [fc.indicator_column(u)], device="cpu")
    with pytest.raises(ValueError):
        fc.categorical_column_with_vocabulary_file("userid", _vocab("userid", 4), num_oov_buckets=3)
    ids = fc.single_valued_ids({"userid": ([b"userid_3", b"zzz"], np.array([0, 1, 1, 2]))}, [u])
    assert ids.tolist() == [[3], [-1], [-1]]                                           # present / missing / out of vocabulary
    with pytest.raises(ValueError, match="multi-valued"):
        fc.single_valued_ids({"userid": ([b"userid_3", b"userid_1"], np.array([0, 2]))}, [u])


def test_sequence_input_layer_layout_on_the_host(host):
    v = _vocab("feedid", 10)
    seq = fc.sequence_categorical_column_with_vocabulary_file("his_read_comment_7d_seq", v)
    (col,) = fc.shared_embedding_columns([seq], 4)
    features = {"his_read_comment_7d_seq": ([b"feedid_1", b"feedid_9", b"oov", b"feedid_0"], np.array([0, 2, 2, 4]))}
    emb, lens = fc.sequence_input_layer(features, [col], device="cpu")
    Wt = host.vars[col.variable_name].detach()
    assert emb.shape == (3, 2, 4) and lens.tolist() == [2, 0, 2]                        # T = longest row in the batch
    assert torch.equal(emb[0, 0], Wt[1]) and torch.equal(emb[0, 1], Wt[9]) and torch.all(emb[1] == 0)
    assert torch.all(emb[2, 0] == 0) and torch.equal(emb[2, 1], Wt[0])                  # an OOV step is a zero vector, not dropped
