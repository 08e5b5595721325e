"""CPU tests of libctr_feed.so (include/ctr_feed.h) against the pure-Python readers and the wire-format known answers:
CRC-32C vectors, TFRecord framing and corruption handling, Example / SequenceExample parsing incl. the reference's
SequenceExample-parsed-as-Example quirk, FixedLen defaults, vocabulary lookup, multi-threaded == single-threaded."""
import ctypes
import os
import re

import numpy as np
import pytest

from recalgorithm_b200 import io as cio
from recalgorithm_b200.io import native, tfrecord
from test_io import wechat_record

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ctr_feed.h")).read()
    declared = set(re.findall(r"\b(ctr_feed_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    L = ctypes.CDLL(native.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert native.lib().ctr_feed_version() == 1


def test_crc32c_known_answers_and_agreement():
    assert native.crc32c(b"123456789") == 0xE3069283 and native.crc32c(b"") == 0 and native.crc32c(bytes(32)) == 0x8A9136AA
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 9, 63, 1000, 4097):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert native.crc32c(data) == tfrecord.crc32c(data) and native.masked_crc32c(data) == tfrecord.masked_crc32c(data)


def test_tfrecord_index_roundtrip_and_corruption(tmp_path):
    recs = [b"", b"a", os.urandom(1000), b"x" * 70000]
    p = str(tmp_path / "t.tfrecord")
    cio.write_records(p, recs)
    buf, off, ln = native.read_tfrecord_file(p)
    assert [buf[int(o): int(o + l)].tobytes() for o, l in zip(off, ln)] == recs
    mbuf, moff, mln = native.read_tfrecord_file(p, mmap=True)                       # memory-mapped variant
    assert np.array_equal(moff, off) and np.array_equal(mln, ln) and bytes(mbuf[int(off[2]): int(off[2] + ln[2])]) == recs[2]
    raw = bytearray(open(p, "rb").read())
    raw[12 + 0 + 4 + 12 + 0] ^= 0xFF                              # flip a data byte of the 2nd record
    with pytest.raises(IOError):
        native.index_tfrecord(bytes(raw))
    assert len(native.index_tfrecord(bytes(raw), verify=False)[0]) == 4
    with pytest.raises(IOError):
        native.index_tfrecord(bytes(raw[:20]), verify=False)
    bad_len = bytearray(raw); bad_len[0] ^= 1
    with pytest.raises(IOError):
        native.index_tfrecord(bytes(bad_len))
    assert len(native.index_tfrecord(b"")[0]) == 0


def test_vocabulary_matches_python(tmp_path):
    p = tmp_path / "userid.txt"
    p.write_bytes(b"userid_8\nuserid_3\r\nuserid_11\nuserid_3\n")
    for v in (native.Vocabulary(str(p)), native.Vocabulary([b"userid_8", b"userid_3", b"userid_11", b"userid_3"])):
        assert len(v) == 4
        assert v.lookup([b"userid_3", b"", b"userid_999", b"userid_8", b"userid_11"]).tolist() == [1, -1, -1, 0, 2]
    (tmp_path / "nonl.txt").write_bytes(b"a\nb")
    assert native.Vocabulary(str(tmp_path / "nonl.txt")).lookup([b"b", b"a"]).tolist() == [1, 0]
    pyv = cio.VocabularyFile(str(p))
    keys = [b"userid_%d" % i for i in range(15)] + [b""]
    assert native.Vocabulary(str(p)).lookup(keys).tolist() == pyv.lookup(keys).tolist()
    with pytest.raises(IOError):
        native.Vocabulary(str(tmp_path / "missing.txt"))


@pytest.mark.parametrize("read_fl", [False, True])
@pytest.mark.parametrize("threads", [1, 4])
def test_parse_examples_matches_python_reader(tmp_path, read_fl, threads):
    rng = np.random.default_rng(3)
    B = 300
    recs = [wechat_record(rng, i)[0] for i in range(B)]
    p = str(tmp_path / "train.tfrecord")
    cio.write_records(p, recs)
    cat_keys = ["userid", "feedid", "device", "authorid", "bgm_song_id", "bgm_singer_id", "his_read_comment_7d_seq", "manual_tag_list",
                "not_in_the_file"]
    prefix = {"his_read_comment_7d_seq": "feedid", "manual_tag_list": "manual_tag"}
    toks = {k: [f"{prefix.get(k, k)}_{i}".encode() for i in range(40)] for k in cat_keys}
    spec = {k: cio.VarLenFeature() for k in cat_keys}
    spec.update({"videoplayseconds": cio.FixedLenFeature(), "read_comment": cio.FixedLenFeature(), "missing_dense": cio.FixedLenFeature(default_value=-2.5)})
    want = cio.parse_example(list(cio.read_records(p)), spec, read_feature_lists=read_fl)
    buf, off, ln = native.read_tfrecord_file(p)
    got = native.parse_examples(buf, off, ln, {k: native.Vocabulary(toks[k]) for k in cat_keys},
                                {"videoplayseconds": (1, 0.0), "read_comment": (1, 0.0), "missing_dense": (1, -2.5)},
                                read_feature_lists=read_fl, num_threads=threads)
    for k in cat_keys:
        vals, offs = want[k]
        ids, row_offsets = got[k]
        assert np.array_equal(row_offsets, offs), k
        assert ids.tolist() == cio.VocabularyFile(toks[k]).lookup(vals).tolist(), k
    if not read_fl:
        assert got["his_read_comment_7d_seq"][1][-1] == 0          # feature_lists dropped (parity note 8)
    else:
        assert got["his_read_comment_7d_seq"][1][-1] > 0
    for k in ("videoplayseconds", "read_comment", "missing_dense"):
        assert np.array_equal(got[k], want[k]), k
    assert np.all(got["missing_dense"] == -2.5) and (got["bgm_song_id"][0] == -1).any()


def test_parse_examples_kinds_errors_and_unpacked_floats():
    ex = [cio.encode_example({"a": ("bytes", [b"x", b"y", b"x"]), "f": ("float", [1.5, -2.0]), "i": ("int64", [3])}),
          cio.encode_example({"a": ("bytes", []), "f": ("float", [])}),
          cio.encode_example({})]
    blob = b"".join(ex)
    off = np.cumsum([0] + [len(e) for e in ex[:-1]]).astype(np.uint64)
    ln = np.array([len(e) for e in ex], np.uint64)
    v = native.Vocabulary([b"x", b"y"])
    out = native.parse_examples(blob, off, ln, {"a": v}, {"f": (2, 9.0)})
    assert out["a"][0].tolist() == [0, 1, 0] and out["a"][1].tolist() == [0, 3, 3, 3]
    assert out["f"].tolist() == [[1.5, -2.0], [9.0, 9.0], [9.0, 9.0]]
    with pytest.raises(ValueError):                                  # wrong number of values for a FixedLen key
        native.parse_examples(blob, off, ln, {}, {"f": (3, 0.0)})
    with pytest.raises(ValueError):                                  # kind mismatch: int64 feature under a string key
        native.parse_examples(blob, off, ln, {"i": v}, {})
    with pytest.raises(ValueError):                                  # truncated proto
        native.parse_examples(ex[0][:-2], np.array([0], np.uint64), np.array([len(ex[0]) - 2], np.uint64), {"a": v}, {})
    # a FloatList written UNPACKED (one fixed32 per value, wire type 5) is legal proto input
    feat = b"\x12" + bytes([10]) + b"\x0d" + np.float32(1.5).tobytes() + b"\x0d" + np.float32(-2.0).tobytes()      # Feature{float_list{1:fixed32,1:fixed32}}
    entry = b"\x0a\x01f" + b"\x12" + bytes([len(feat)]) + feat
    features = b"\x0a" + bytes([len(entry)]) + entry
    rec = b"\x0a" + bytes([len(features)]) + features
    out2 = native.parse_examples(rec, np.array([0], np.uint64), np.array([len(rec)], np.uint64), {}, {"f": (2, 0.0)})
    assert out2["f"].tolist() == [[1.5, -2.0]]
    assert cio.parse_single(rec)[0]["f"] == ("float", [1.5, -2.0])


def test_parse_example_native_equals_python_features(tmp_path):
    """feature_column.parse_example_native == parse_example + per-column vocabulary lookup for the reference's column set
    (DeepFM/deepfm.py:56-93 + DIN's shared feedid / history columns)."""
    from recalgorithm_b200 import feature_column as fc
    rng = np.random.default_rng(8)
    B = 120
    p = str(tmp_path / "train.tfrecord")
    cio.write_records(p, [wechat_record(rng, i)[0] for i in range(B)])
    cats = ["userid", "feedid", "device", "authorid", "bgm_song_id", "bgm_singer_id"]
    vocabs = {k: cio.VocabularyFile([f"{k}_{i}".encode() for i in rng.permutation(60)]) for k in cats}
    cat_cols = {k: fc.categorical_column_with_vocabulary_file(k, vocabs[k]) for k in cats}
    seq = fc.categorical_column_with_vocabulary_file("his_read_comment_7d_seq", vocabs["feedid"])
    cols = [fc.indicator_column(c) for c in cat_cols.values()] + [fc.embedding_column(c, 8) for c in cat_cols.values()] + \
        list(fc.shared_embedding_columns([cat_cols["feedid"], seq], 16, combiner="mean")) + \
        [fc.numeric_column("read_comment", default_value=0.0), fc.numeric_column("videoplayseconds", default_value=0.0)]
    spec = fc.make_parse_example_spec(cols)
    buf, off, ln = native.read_tfrecord_file(p)
    for read_fl in (False, True):
        want = cio.parse_example(list(cio.read_records(p)), spec, read_feature_lists=read_fl)
        got = fc.parse_example_native(buf, off, ln, cols, read_feature_lists=read_fl)
        assert set(got) == set(want)
        for k in want:
            if isinstance(want[k], tuple):
                base = seq if k == "his_read_comment_7d_seq" else cat_cols[k]
                assert got[k][0].dtype == np.int64 and np.array_equal(got[k][1], want[k][1])
                assert got[k][0].tolist() == base.vocabulary.lookup(want[k][0]).tolist(), k
            else:
                assert got[k].shape == want[k].shape and np.array_equal(got[k], want[k]), k


# ------------------------------------------------------------------------------------------------ fuzzing
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402
from recalgorithm_b200.io.example import _enc_feature, _enc_varint, _ld  # noqa: E402

_KEYS = ["a", "b", "f", "zz"]
_entry = st.tuples(st.sampled_from(_KEYS), st.integers(0, 4), st.integers(0, 2 ** 16))


def _build_record(entries, with_unknown, with_fl, rng):
    """Example bytes with repeated map keys (last wins), unknown fields at three nesting levels, optional feature_lists."""
    feats = b""
    for key, n, seed in entries:
        r = np.random.default_rng(seed)
        if key == "f":
            kind, vals = "float", [float(x) for x in r.standard_normal(2 if n % 2 else 0).astype(np.float32)]
        else:
            kind, vals = "bytes", [b"t%d" % int(x) for x in r.integers(0, 12, n)]
        entry = _ld(1, key.encode()) + _ld(2, _enc_feature(kind, vals))
        if with_unknown:
            entry += _enc_varint((3 << 3) | 0) + _enc_varint(int(rng.integers(0, 1 << 40)))
        feats += _ld(1, entry)
        if with_unknown:
            feats += _ld(2, b"junk") + _enc_varint((5 << 3) | 5) + b"\x01\x02\x03\x04"
    rec = _ld(1, feats)
    if with_unknown:
        rec = _enc_varint((7 << 3) | 1) + b"\x00" * 8 + rec + _enc_varint((3 << 3) | 0) + _enc_varint(5)
    if with_fl:
        steps = b"".join(_ld(1, _enc_feature("bytes", [b"t%d" % int(x)])) for x in rng.integers(0, 12, int(rng.integers(0, 4))))
        rec += _ld(2, _ld(1, _ld(1, b"zz") + _ld(2, steps)))
    return rec


@settings(deadline=None, max_examples=120, suppress_health_check=[HealthCheck.too_slow], derandomize=True)
@given(records=st.lists(st.tuples(st.lists(_entry, max_size=6), st.booleans(), st.booleans()), min_size=1, max_size=12),
       read_fl=st.booleans(), seed=st.integers(0, 2 ** 16))
def test_native_parser_fuzz_against_python(records, read_fl, seed):
    rng = np.random.default_rng(seed)
    recs = [_build_record(e, u, fl, rng) for e, u, fl in records]
    blob = b"".join(recs)
    ln = np.array([len(r) for r in recs], np.uint64)
    off = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.uint64)
    toks = [b"t%d" % i for i in range(8)]                          # t8..t11 are out of vocabulary
    spec = {"a": cio.VarLenFeature(), "b": cio.VarLenFeature(), "zz": cio.VarLenFeature(), "f": cio.FixedLenFeature((2,), "float", 7.0)}
    want = cio.parse_example(recs, spec, read_feature_lists=read_fl)
    v = native.Vocabulary(toks)
    got = native.parse_examples(blob, off, ln, {"a": v, "b": v, "zz": v}, {"f": (2, 7.0)}, read_feature_lists=read_fl, num_threads=2)
    pyv = cio.VocabularyFile(toks)
    for k in ("a", "b", "zz"):
        assert np.array_equal(got[k][1], want[k][1]), k
        assert got[k][0].tolist() == pyv.lookup(want[k][0]).tolist(), k
    assert np.array_equal(got["f"], want["f"])


def test_native_parser_survives_corrupted_input():
    """Memory safety: random truncations / byte flips of valid batches must come back as a result or a ValueError."""
    rng = np.random.default_rng(4)
    recs = [wechat_record(rng, i)[0] for i in range(64)]
    v = native.Vocabulary([b"userid_%d" % i for i in range(50)])
    ok = bad = 0
    for trial in range(600):
        rec = bytearray(recs[trial % len(recs)])
        mode = trial % 3
        if mode == 0:
            rec = rec[: int(rng.integers(0, len(rec)))]
        elif mode == 1:
            for _ in range(int(rng.integers(1, 6))):
                rec[int(rng.integers(0, len(rec)))] = int(rng.integers(0, 256))
        else:
            pos = int(rng.integers(0, len(rec)))
            rec[pos:pos] = bytes(rng.integers(0, 256, int(rng.integers(1, 9)), dtype=np.uint8))
        data = bytes(rec)
        try:
            out = native.parse_examples(data, np.array([0], np.uint64), np.array([len(data)], np.uint64),
                                        {"userid": v, "his_read_comment_7d_seq": v}, {"read_comment": (1, 0.0)}, read_feature_lists=True)
            assert out["userid"][1].shape == (2,)
            ok += 1
        except ValueError:
            bad += 1
    assert ok > 0 and bad > 0


def test_parallel_crc_verification_reports_the_first_bad_record(tmp_path):
    """The payload CRCs are checked on several threads (ctr_feed_tfrecord_verify) after a sequential scan that only trusts
    lengths whose own CRC matched: same verdict as the sequential verifier, and the FIRST corrupted record is the one named."""
    recs = [bytes([i % 251]) * (40 + i % 17) for i in range(5000)]
    p = str(tmp_path / "many.tfrecord")
    cio.write_records(p, recs)
    raw = bytearray(open(p, "rb").read())
    off, ln = native.index_tfrecord(bytes(raw), num_threads=8)
    assert len(off) == 5000 and [int(x) for x in ln[:3]] == [40, 41, 42]
    bad = bytearray(raw)
    for r in (4100, 700):                                          # two corrupted payloads, in different thread chunks
        bad[int(off[r]) + 5] ^= 0x40
    for nt in (1, 3, 8):
        with pytest.raises(IOError) as e:
            native.index_tfrecord(bytes(bad), num_threads=nt)
        assert f"byte {int(off[700]) - 12} " in str(e.value)        # record 700 starts 12 bytes before its payload
    assert len(native.index_tfrecord(bytes(bad), verify=False)[0]) == 5000
    hdr = bytearray(raw); hdr[int(off[3000]) - 12] ^= 1              # a corrupted LENGTH is caught by the scan itself
    with pytest.raises(IOError):
        native.index_tfrecord(bytes(hdr), num_threads=8)


def test_plain_c_caller_of_the_feeder(tmp_path):
    """examples/feed_demo.c: the feeder's C ABI used from plain C (no Python in the process that calls it) -- index, CRC pass,
    vocabulary files, batched parse with the capacity-retry protocol -- against the Python twin."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "recalgorithm_b200", "csrc_feed")
    exe = str(tmp_path / "feed_demo")
    subprocess.run(["gcc", "-O2", "-Wall", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "feed_demo.c"), "-o", exe,
                    "-L", lib_dir, "-lctr_feed", f"-Wl,-rpath,{lib_dir}"], check=True, capture_output=True)
    rng = np.random.default_rng(4)
    n = 1500
    recs = [wechat_record(rng, i)[0] for i in range(n)]
    path = str(tmp_path / "d.tfrecord")
    cio.write_records(path, recs)
    keys = ["userid", "feedid", "bgm_song_id", "manual_tag_list"]               # the last one lives in feature_lists: parses empty
    vdir = tmp_path / "vocab"; vdir.mkdir()
    vocabs = {}
    for k in keys:
        toks = [f"{k}_{i}".encode() for i in range(40)]
        (vdir / f"{k}.txt").write_bytes(b"\n".join(toks) + b"\n")
        vocabs[k] = cio.VocabularyFile(toks)
    out = subprocess.run([exe, path, str(vdir), "128", *keys], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[0] == f"records {n} vocabulary sizes 40 40 40 40"
    spec = {k: cio.VarLenFeature("bytes") for k in keys} | {"read_comment": cio.FixedLenFeature((1,), "float", 0.0)}
    want = cio.parse_example(recs, spec)
    for k, line in zip(keys, lines[1:]):
        ids = vocabs[k].lookup(want[k][0]) if len(want[k][0]) else np.zeros(0, np.int64)
        chk = int(sum((i + 1) * (int(v) + 2) for i, v in enumerate(ids)) % (1 << 64))
        assert line == f"{k} values {len(ids)} oov {int((ids < 0).sum())} checksum {chk}", (line, k)
    assert lines[-1] == f"read_comment sum {float(want['read_comment'].sum()):.1f}"
    bad = bytearray(open(path, "rb").read()); bad[40] ^= 1
    open(path, "wb").write(bad)
    out = subprocess.run([exe, path, str(vdir), "128", *keys], capture_output=True, text=True, timeout=120)
    assert out.returncode == 1 and "crc mismatch" in out.stderr


def test_key_order_subsets_duplicates_and_unknown_keys_match_the_python_twin():
    """The native parser predicts the key of map entry j from entry j of the previous record (one memcmp instead of a hash probe);
    records whose keys come in a different order, are missing, unknown, empty, or repeated (a proto map keeps the LAST entry, two
    concatenated Example messages merge) must still parse exactly like the Python twin, in every thread split."""
    rng = np.random.default_rng(77)
    cat_keys, dense_keys = ["a", "bb", "userid", "k" * 40, ""], ["f1", "f2"]
    toks = [b"t%d" % i for i in range(30)]
    vocab_py = cio.VocabularyFile(toks); vocab_n = native.Vocabulary(toks)
    recs = []
    for _ in range(3000):
        entries = []
        for k in rng.permutation(cat_keys + dense_keys + ["unknown1", "zz"]).tolist():
            if rng.random() < 0.3:
                continue                                                       # key absent from this record
            if k in dense_keys:
                entries.append((k, ("float", [float(np.float32(rng.standard_normal()))] if rng.random() < 0.9 else [])))
            else:
                n = int(rng.integers(0, 4))
                entries.append((k, ("bytes", [toks[int(rng.integers(0, 30))] if rng.random() < 0.8 else b"oov" for _ in range(n)])))
        if entries and rng.random() < 0.3:                                     # the same key twice inside one map: the last one wins
            k, _ = entries[int(rng.integers(0, len(entries)))]
            if k not in dense_keys:
                entries.append((k, ("bytes", [toks[int(rng.integers(0, 30))]])))
        body = b"".join(cio.example._ld(1, cio.example._ld(1, k.encode()) + cio.example._ld(2, cio.example._enc_feature(kind, v)))
                        for k, (kind, v) in entries)
        rec = cio.example._ld(1, body)
        if rng.random() < 0.2:                                                 # two Example messages back to back: maps merge
            rec += cio.encode_example({"a": ("bytes", [toks[3]]), "f1": ("float", [2.5])})
        recs.append(rec)
    spec = {k: cio.VarLenFeature("bytes") for k in cat_keys} | {k: cio.FixedLenFeature((1,), "float", -1.0) for k in dense_keys}
    want = cio.parse_example(recs, spec)
    buf = np.frombuffer(b"".join(recs), np.uint8)
    lens = np.array([len(r) for r in recs], np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    for nt in (1, 2, 5, 16):
        got = native.parse_examples(buf, offs, lens, {k: vocab_n for k in cat_keys}, {k: (1, -1.0) for k in dense_keys}, num_threads=nt)
        for k in cat_keys:
            ids = vocab_py.lookup(want[k][0]) if len(want[k][0]) else np.zeros(0, np.int64)
            assert np.array_equal(got[k][0], ids) and np.array_equal(got[k][1], want[k][1]), (k, nt)
        for k in dense_keys:
            assert np.array_equal(got[k], want[k]), (k, nt)


def test_allocation_failures_come_back_as_errors_not_as_crashes():
    """No C++ exception crosses the C ABI: an allocation the process cannot satisfy is CTR_FEED_ERR_NOMEM (MemoryError here)."""
    sh = native.Shuffler(1 << 62)
    with pytest.raises(MemoryError):
        sh.emit(1 << 60, True, np.zeros(4), 4)                               # a 2^60-slot shuffle buffer
    assert native.shuffle_order(10, 3, np.linspace(0, 0.9, 10)).size == 10    # the library is still usable afterwards
