"""Pins the Example / SequenceExample wire codec against an independent implementation: Google's protobuf
runtime, driven by the public tensorflow/core/example/{feature,example}.proto schema restated as a descriptor (TensorFlow itself
is not installable here).  Both directions and both parsers (the Python twin io/example.py and the native feeder):

  protobuf-encoded bytes  -> our parsers   == the values that were set
  our encoder's bytes     -> protobuf      == the values that were encoded
  SequenceExample bytes parsed AS Example by protobuf -> context features only, feature_lists invisible (parity note 8:
  the reference writes SequenceExamples, DataGenerator.py:429-442, and reads them with tf.parse_example, DCN/dcn.py:128-129)
"""
import numpy as np
import pytest

pb = pytest.importorskip("google.protobuf")
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory  # noqa: E402

from recalgorithm_b200 import io as cio  # noqa: E402
from recalgorithm_b200.io import example as ex, native  # noqa: E402

T = descriptor_pb2.FieldDescriptorProto


def _schema(packed: bool = True):
    """tensorflow/core/example/feature.proto + example.proto (proto3; repeated scalars are packed unless `packed` is False,
    which produces the legacy one-tag-per-value encoding old writers emit)."""
    f = descriptor_pb2.FileDescriptorProto(name=f"tf_example_{int(packed)}.proto", package=f"tfpin{int(packed)}", syntax="proto3")
    pkg = "." + f.package

    def msg(name):
        m = f.message_type.add(); m.name = name
        return m

    def field(m, name, num, typ, label=T.LABEL_OPTIONAL, type_name=None, oneof=None, pack=None):
        x = m.field.add(); x.name, x.number, x.type, x.label = name, num, typ, label
        if type_name:
            x.type_name = type_name
        if oneof is not None:
            x.oneof_index = oneof
        if pack is not None:
            x.options.packed = pack
        return x

    def map_field(m, name, num, value_type_name):
        e = m.nested_type.add(); e.name = name.title().replace("_", "") + "Entry"; e.options.map_entry = True
        field(e, "key", 1, T.TYPE_STRING); field(e, "value", 2, T.TYPE_MESSAGE, type_name=value_type_name)
        field(m, name, num, T.TYPE_MESSAGE, T.LABEL_REPEATED, type_name=f"{pkg}.{m.name}.{e.name}")

    field(msg("BytesList"), "value", 1, T.TYPE_BYTES, T.LABEL_REPEATED)
    field(msg("FloatList"), "value", 1, T.TYPE_FLOAT, T.LABEL_REPEATED, pack=packed)
    field(msg("Int64List"), "value", 1, T.TYPE_INT64, T.LABEL_REPEATED, pack=packed)
    feat = msg("Feature"); feat.oneof_decl.add().name = "kind"
    field(feat, "bytes_list", 1, T.TYPE_MESSAGE, type_name=pkg + ".BytesList", oneof=0)
    field(feat, "float_list", 2, T.TYPE_MESSAGE, type_name=pkg + ".FloatList", oneof=0)
    field(feat, "int64_list", 3, T.TYPE_MESSAGE, type_name=pkg + ".Int64List", oneof=0)
    map_field(msg("Features"), "feature", 1, pkg + ".Feature")
    field(msg("FeatureList"), "feature", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, type_name=pkg + ".Feature")
    map_field(msg("FeatureLists"), "feature_list", 1, pkg + ".FeatureList")
    field(msg("Example"), "features", 1, T.TYPE_MESSAGE, type_name=pkg + ".Features")
    se = msg("SequenceExample")
    field(se, "context", 1, T.TYPE_MESSAGE, type_name=pkg + ".Features")
    field(se, "feature_lists", 2, T.TYPE_MESSAGE, type_name=pkg + ".FeatureLists")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(f)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName(f"{f.package}.{n}"))  # noqa: E731
    return get("Example"), get("SequenceExample")


def _fill(feature, kind, values):
    if kind == "bytes":
        feature.bytes_list.value.extend(values)
    elif kind == "float":
        feature.float_list.value.extend(values)
    else:
        feature.int64_list.value.extend(values)


def _random_features(rng, n_keys=6):
    out = {}
    for i in range(n_keys):
        kind = ("bytes", "float", "int64")[int(rng.integers(0, 3))]
        n = int(rng.integers(0, 5))
        if kind == "bytes":
            vals = [bytes(rng.integers(0, 256, int(rng.integers(0, 12)), dtype=np.uint8)) for _ in range(n)]
        elif kind == "float":
            vals = [float(np.float32(x)) for x in rng.standard_normal(n)]
        else:
            vals = [int(x) for x in rng.integers(-2**62, 2**62, n)] + ([-1, 0, 2**63 - 1, -2**63][: int(rng.integers(0, 5))])
        out[f"key_{i}_{kind}"] = (kind, vals)
    return out


@pytest.mark.parametrize("packed", [True, False])
def test_protobuf_encoded_examples_parse_to_the_values_set(packed):
    Example, SequenceExample = _schema(packed)
    rng = np.random.default_rng(11)
    for trial in range(40):
        feats = _random_features(rng)
        steps = {"seq_a": [("bytes", [b"x%d" % i, b""]) for i in range(int(rng.integers(0, 4)))],
                 "seq_b": [("int64", [int(rng.integers(-9, 9))]) for _ in range(int(rng.integers(0, 3)))]}
        m = SequenceExample() if trial % 2 else Example()
        holder = m.context if trial % 2 else m.features
        for k, (kind, vals) in feats.items():
            _fill(holder.feature[k], kind, vals)
        if trial % 2:
            for k, st in steps.items():
                fl = m.feature_lists.feature_list[k]
                for kind, vals in st:
                    _fill(fl.feature.add(), kind, vals)
        blob = m.SerializeToString()
        got, got_fl = ex.parse_single(blob, read_feature_lists=True)
        for k, (kind, vals) in feats.items():
            assert k in got
            if vals:                                                  # an empty list has no kind on the wire
                assert got[k][0] == kind
            if kind == "float":
                assert np.array_equal(np.asarray(got[k][1], np.float32), np.asarray(vals, np.float32))
            else:
                assert list(got[k][1]) == list(vals)
        if trial % 2:
            for k, st in steps.items():
                if st:
                    assert [list(v) for _, v in got_fl[k]] == [list(v) for _, v in st]
        assert ex.parse_single(blob, read_feature_lists=False)[1] == {}


@pytest.mark.parametrize("packed", [True, False])
def test_native_feeder_reads_protobuf_encoded_records(packed):
    """ctr_feed_parse_examples over records written by the protobuf runtime: VarLen string keys through a vocabulary,
    FixedLen float keys (packed and one-tag-per-value float lists), missing keys, and feature_lists read only on request."""
    Example, SequenceExample = _schema(packed)
    rng = np.random.default_rng(3)
    toks = [b"tok_%d" % i for i in range(50)]
    vocab = native.Vocabulary(toks)
    recs, want_ids, want_dense, want_seq = [], [], [], []
    for b in range(200):
        m = SequenceExample()
        ids = []
        if rng.random() < 0.9:
            n = int(rng.integers(0, 4))
            vals = [toks[int(rng.integers(0, 50))] if rng.random() < 0.8 else b"oov_%d" % b for _ in range(n)]
            m.context.feature["cat"].bytes_list.value.extend(vals)
            ids = [toks.index(v) if v in toks else -1 for v in vals]
        want_ids.append(ids)
        if rng.random() < 0.7:
            x = float(np.float32(rng.standard_normal()))
            m.context.feature["dense"].float_list.value.append(x)
            want_dense.append(x)
        else:
            want_dense.append(0.5)                                    # the spec's default
        seq = [toks[int(rng.integers(0, 50))] for _ in range(int(rng.integers(0, 5)))]
        for s in seq:
            m.feature_lists.feature_list["seq"].feature.add().bytes_list.value.append(s)
        want_seq.append([toks.index(s) for s in seq])
        recs.append(m.SerializeToString())
    buf = np.frombuffer(b"".join(recs), np.uint8)
    lens = np.array([len(r) for r in recs], np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    for rfl in (False, True):
        out = native.parse_examples(buf, offs, lens, {"cat": vocab, "seq": vocab}, {"dense": (1, 0.5)}, read_feature_lists=rfl)
        ids, ro = out["cat"]
        assert [ids[ro[b]:ro[b + 1]].tolist() for b in range(200)] == want_ids
        assert np.array_equal(out["dense"][:, 0], np.asarray(want_dense, np.float32))
        sids, sro = out["seq"]
        got_seq = [sids[sro[b]:sro[b + 1]].tolist() for b in range(200)]
        assert got_seq == (want_seq if rfl else [[] for _ in range(200)])   # tf.parse_example never sees feature_lists


def test_our_encoder_is_read_back_by_protobuf_and_sequence_examples_parse_as_examples():
    Example, SequenceExample = _schema(True)
    rng = np.random.default_rng(5)
    for _ in range(30):
        feats = _random_features(rng)
        steps = {"his": [("bytes", [b"feed_%d" % int(rng.integers(0, 99))]) for _ in range(int(rng.integers(0, 6)))]}
        blob = ex.encode_sequence_example(feats, steps)
        m = SequenceExample.FromString(blob)
        assert set(m.context.feature) == set(feats)
        for k, (kind, vals) in feats.items():
            f = m.context.feature[k]
            got = {"bytes": f.bytes_list.value, "float": f.float_list.value, "int64": f.int64_list.value}[kind]
            if kind == "float":
                assert np.array_equal(np.asarray(got, np.float32), np.asarray(vals, np.float32))
            else:
                assert list(got) == list(vals)
        assert [list(x.bytes_list.value) for x in m.feature_lists.feature_list["his"].feature] == [v for _, v in steps["his"]]
        # the reference's read path: the same bytes as an Example -> field 1 is the context, field 2 is an unknown field
        e = Example.FromString(blob)
        assert set(e.features.feature) == set(feats)
        assert e.features.SerializeToString(deterministic=True) == m.context.SerializeToString(deterministic=True)
        assert not any(name.startswith("his") for name in e.features.feature)
        # and a plain Example written by us
        e2 = Example.FromString(ex.encode_example(feats))
        assert e2.features.SerializeToString(deterministic=True) == m.context.SerializeToString(deterministic=True)
    # records written through the TFRecord writer come back intact through protobuf too
    rec, ctx, seq, tags = __import__("test_io").wechat_record(rng, 7)
    m = SequenceExample.FromString(rec)
    assert m.context.feature["userid"].bytes_list.value[0] == ctx["userid"][1][0]
    assert cio.parse_single(rec)[0]["userid"] == ctx["userid"]
