"""CPU tests of the wire formats feeding the hot path: TFRecord framing, Example/SequenceExample wire
parsing incl. the reference's SequenceExample-parsed-as-Example quirk, parse spec defaults, vocabulary lookup."""
import os
import struct

import numpy as np
import pytest

from recalgorithm_b200 import feature_column as fc
from recalgorithm_b200 import io as cio
from recalgorithm_b200.io import tfrecord


def wechat_record(rng, i):
    """One record with the exact field names / kinds the reference ETL writes (DataGenerator.py:400-443)."""
    ctx = {"userid": ("bytes", [f"userid_{rng.integers(0, 50)}".encode()]),
           "feedid": ("bytes", [f"feedid_{rng.integers(0, 80)}".encode()]),
           "device": ("bytes", [f"device_{rng.integers(1, 3)}".encode()]),
           "authorid": ("bytes", [f"authorid_{rng.integers(0, 30)}".encode()]),
           "bgm_song_id": ("bytes", [b"" if rng.random() < 0.3 else f"bgm_song_id_{rng.integers(0, 20)}".encode()]),
           "bgm_singer_id": ("bytes", [b"" if rng.random() < 0.3 else f"bgm_singer_id_{rng.integers(0, 20)}".encode()]),
           "videoplayseconds": ("float", [float(np.log1p(rng.poisson(3)))]),
           "read_comment": ("float", [float(rng.random() < 0.0356)])}
    seq = [("bytes", [f"feedid_{rng.integers(0, 80)}".encode()]) for _ in range(int(rng.integers(0, 5)))]
    tags = [("bytes", [f"manual_tag_{rng.integers(0, 9)}".encode()]) for _ in range(int(rng.integers(0, 4)))]
    return cio.encode_sequence_example(ctx, {"his_read_comment_7d_seq": seq, "manual_tag_list": tags}), ctx, seq, tags


def test_crc32c_known_answers():
    assert tfrecord.crc32c(b"123456789") == 0xE3069283          # RFC 3720 check value
    assert tfrecord.crc32c(b"") == 0
    assert tfrecord.crc32c(bytes(32)) == 0x8A9136AA              # iSCSI test vector: 32 zero bytes


def test_tfrecord_roundtrip_and_corruption(tmp_path):
    recs = [b"", b"a", os.urandom(1000), b"x" * 70000]
    p = str(tmp_path / "t.tfrecord")
    assert cio.write_records(p, recs) == 4
    assert list(cio.read_records(p)) == recs
    try:                                                          # independent reader, when tensorboard is installed
        from tensorboard.compat.tensorflow_stub.pywrap_tensorflow import PyRecordReader_New
        r, got = PyRecordReader_New(p), []
        while True:
            try:
                r.GetNext(); got.append(r.record())
            except Exception:
                break
        assert got == recs
    except ImportError:
        pass
    raw = bytearray(open(p, "rb").read())
    raw[12 + 0 + 4 + 12 + 0] ^= 0xFF                              # flip a data byte of the 2nd record
    open(p, "wb").write(raw)
    with pytest.raises(IOError):
        list(cio.read_records(p))
    assert len(list(cio.read_records(p, verify=False))) == 4
    open(p, "wb").write(bytes(raw[:20]))
    with pytest.raises(IOError):
        list(cio.read_records(p, verify=False))


def test_sequence_example_parsed_as_example_drops_feature_lists():
    rng = np.random.default_rng(0)
    rec, ctx, seq, tags = wechat_record(rng, 0)
    feats, fl = cio.parse_single(rec)                             # what tf.parse_example sees (parity note 8)
    assert fl == {} and "his_read_comment_7d_seq" not in feats
    assert feats["userid"] == ctx["userid"]
    assert feats["videoplayseconds"][0] == "float" and abs(feats["videoplayseconds"][1][0] - ctx["videoplayseconds"][1][0]) < 1e-6
    feats2, fl2 = cio.parse_single(rec, read_feature_lists=True)
    assert [v for _, v in fl2["his_read_comment_7d_seq"]] == [v for _, v in seq]
    assert [v for _, v in fl2["manual_tag_list"]] == [v for _, v in tags]
    ex = cio.encode_example({"a": ("int64", [-3, 7, 2 ** 40]), "b": ("float", [1.5, -2.0]), "c": ("bytes", [])})
    f3, _ = cio.parse_single(ex)
    assert f3["a"] == ("int64", [-3, 7, 2 ** 40]) and f3["b"] == ("float", [1.5, -2.0]) and f3["c"][1] == []


def test_parse_example_batch_matches_reference_spec(tmp_path):
    rng = np.random.default_rng(1)
    recs = [wechat_record(rng, i) for i in range(17)]
    p = str(tmp_path / "train.tfrecord")
    cio.write_records(p, [r[0] for r in recs])
    vocab = {k: fc.categorical_column_with_vocabulary_file(k, cio.VocabularyFile([f"{k}_{i}".encode() for i in range(40)]))
             for k in ("userid", "feedid", "device", "authorid", "bgm_song_id", "bgm_singer_id")}
    seq_col = fc.categorical_column_with_vocabulary_file("his_read_comment_7d_seq", vocab["feedid"].vocabulary)
    cols = [fc.embedding_column(c, 8) for c in vocab.values()] + [fc.embedding_column(seq_col, 8),
                                                                   fc.numeric_column("videoplayseconds", default_value=0.0),
                                                                   fc.numeric_column("read_comment", default_value=0.0),
                                                                   fc.numeric_column("missing_dense", default_value=0.0)]
    spec = fc.make_parse_example_spec(cols)
    assert isinstance(spec["userid"], cio.VarLenFeature) and spec["read_comment"].shape == (1,)
    batch = next(tfrecord.batches(cio.read_records(p), 17))       # batch THEN parse, like utils.py:22-23
    out = cio.parse_example(batch, spec)
    assert out["read_comment"].shape == (17, 1) and out["read_comment"].dtype == np.float32
    assert np.all(out["missing_dense"] == 0.0)                    # FixedLenFeature default
    vals, offs = out["userid"]
    assert len(vals) == 17 and np.array_equal(offs, np.arange(18))
    assert out["his_read_comment_7d_seq"][0] == [] and np.all(out["his_read_comment_7d_seq"][1] == 0)   # dropped (quirk)
    out2 = cio.parse_example(batch, spec, read_feature_lists=True)
    assert out2["his_read_comment_7d_seq"][1][-1] == sum(len(r[2]) for r in recs)
    ids = vocab["bgm_song_id"].vocabulary.lookup(out["bgm_song_id"][0])
    want = [(-1 if r[1]["bgm_song_id"][1][0] == b"" else int(r[1]["bgm_song_id"][1][0].split(b"_")[-1])) for r in recs]
    assert ids.tolist() == want and ids.dtype == np.int64         # '' -> OOV -> -1


def test_vocabulary_file(tmp_path):
    p = tmp_path / "userid.txt"
    p.write_bytes(b"userid_8\nuserid_3\nuserid_11\n")
    v = cio.VocabularyFile(str(p))
    assert len(v) == 3
    assert v.lookup([b"userid_3", b"", b"userid_999", b"userid_8"]).tolist() == [1, -1, -1, 0]
    with pytest.raises(ValueError):
        fc.categorical_column_with_vocabulary_file("userid", v, num_oov_buckets=3)
    shared = fc.shared_embedding_columns([fc.categorical_column_with_vocabulary_file("feedid", v),
                                          fc.categorical_column_with_vocabulary_file("his_seq", v)], 16)
    assert [c.categorical_column.key for c in shared] == ["feedid", "his_seq"]          # input order kept
    assert shared[0].variable_name == "input_layer/feedid_his_seq_shared_embedding/embedding_weights"
