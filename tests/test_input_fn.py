"""CPU tests of recalgorithm_b200.input_fn -- the reference's train/eval input_fns (utils.py:4-47) over the native feeder."""
import threading

import numpy as np
import pytest

from recalgorithm_b200 import feature_column as fc
from recalgorithm_b200 import input_fn as I
from recalgorithm_b200 import io as cio
from test_io import wechat_record


@pytest.fixture()
def dataset(tmp_path):
    rng = np.random.default_rng(21)
    n = 103
    recs = []
    for i in range(n):                                              # userid_<i> makes every record identifiable
        ctx = {"userid": ("bytes", [f"userid_{i}".encode()]), "read_comment": ("float", [float(i % 2)]),
               "videoplayseconds": ("float", [float(i) / 8])}
        recs.append(cio.encode_example(ctx))
    p = str(tmp_path / "train.tfrecord")
    cio.write_records(p, recs)
    user = fc.categorical_column_with_vocabulary_file("userid", cio.VocabularyFile([f"userid_{i}".encode() for i in range(n)]))
    cols = [fc.embedding_column(user, 4), fc.numeric_column("videoplayseconds"), fc.numeric_column("read_comment")]
    return p, n, I.make_example_parser(cols, label_keys=["read_comment"])


def _ids(batches):
    return np.concatenate([f["userid"][0] for f, _ in batches])


def test_eval_input_fn_is_one_ordered_pass(dataset):
    p, n, parser = dataset
    out = list(I.eval_input_fn(p, parser, batch_size=16))
    assert [len(f["userid"][1]) - 1 for f, _ in out] == [16] * 6 + [7]              # last partial batch kept
    assert _ids(out).tolist() == list(range(n))
    f, l = out[0]
    assert set(f) == {"userid", "videoplayseconds"} and set(l) == {"read_comment"} and l["read_comment"].shape == (16, 1)
    assert np.allclose(f["videoplayseconds"][:, 0], np.arange(16) / 8)


def test_train_input_fn_repeat_then_batch_and_shuffle(dataset):
    p, n, parser = dataset
    out = list(I.train_input_fn(p, parser, batch_size=50, num_epochs=3, shuffle_buffer_size=0))
    ids = _ids(out)
    assert ids.tolist() == list(range(n)) * 3                                         # no shuffle: file order, epochs concatenated
    assert [len(f["userid"][1]) - 1 for f, _ in out] == [50] * 6 + [9]                # batches run across epoch borders
    sh = _ids(list(I.train_input_fn(p, parser, batch_size=32, num_epochs=2, shuffle_buffer_size=10, seed=5)))
    e1, e2 = sh[:n], sh[n:]
    assert sorted(e1.tolist()) == list(range(n)) and sorted(e2.tolist()) == list(range(n)) and e1.tolist() != e2.tolist()
    assert all(e1[i] < i + 10 for i in range(n))                                      # an element can only move up by < buffer_size
    full = _ids(list(I.train_input_fn(p, parser, batch_size=32, num_epochs=1, shuffle_buffer_size=10_000, seed=5)))
    assert sorted(full.tolist()) == list(range(n)) and full.tolist() != list(range(n))
    again = _ids(list(I.train_input_fn(p, parser, batch_size=32, num_epochs=2, shuffle_buffer_size=10, seed=5)))
    assert again.tolist() == sh.tolist()                                               # seeded


def test_shuffle_order_statistics():
    rng = np.random.default_rng(0)
    o = I.shuffle_order(10_000, 100, rng)
    assert sorted(o.tolist()) == list(range(10_000))
    disp = o - np.arange(10_000)
    assert disp.max() < 100 and disp.min() < -100                                     # late emission is unbounded, early is < buffer


def test_prefetch_thread_stops_when_consumer_leaves_and_forwards_errors(dataset):
    p, n, parser = dataset
    before = threading.active_count()
    it = I.train_input_fn(p, parser, batch_size=8, num_epochs=50, shuffle_buffer_size=0)
    next(it); next(it)
    it.close()
    import time
    time.sleep(0.5)
    assert threading.active_count() <= before + 1

    def bad_parser(batch):
        raise ValueError("boom")
    with pytest.raises(ValueError, match="boom"):
        list(I.eval_input_fn(p, bad_parser, batch_size=8))


def test_num_epochs_none_repeats_forever(dataset):
    """dataset.repeat(None): the iterator never ends by itself (Estimator callers stop it with max_steps)."""
    import itertools
    p, n, parser = dataset
    it = I.train_input_fn(p, parser, batch_size=n, num_epochs=None, shuffle_buffer_size=0)
    taken = list(itertools.islice(it, 5))                  # five full passes, no StopIteration
    it.close()
    assert len(taken) == 5


def test_mmap_reads_the_same_batches(dataset):
    """mmap=True maps the TFRecord file instead of reading it into RAM: same records, same order."""
    p, n, parser = dataset
    a = list(I.eval_input_fn(p, parser, batch_size=16))
    b = list(I.eval_input_fn(p, parser, batch_size=16, mmap=True))
    assert _ids(a).tolist() == _ids(b).tolist() == list(range(n))
    t = list(I.train_input_fn(p, parser, batch_size=50, num_epochs=1, shuffle_buffer_size=0, mmap=True))
    assert _ids(t).tolist() == list(range(n))


def test_corrupted_payload_raises_when_its_batch_is_produced(dataset, tmp_path):
    """Payload CRCs are verified per batch inside the prefetch thread (DataLossError-like timing): the batches before the
    corrupted record arrive, the one that contains it raises; a corrupted LENGTH stops the background index scan there, and the
    batches in front of it still arrive."""
    from recalgorithm_b200.io import native
    p, n, parser = dataset
    raw = bytearray(open(p, "rb").read())
    off, ln = native.index_tfrecord(bytes(raw))
    bad = bytearray(raw); bad[int(off[70]) + 2] ^= 0x10              # record 70 -> batch 4 at batch_size 16
    q = str(tmp_path / "bad.tfrecord"); open(q, "wb").write(bad)
    it = I.eval_input_fn(q, parser, batch_size=16)
    got = [next(it) for _ in range(4)]                                # records 0..63 are fine
    assert _ids(got).tolist() == list(range(64))
    with pytest.raises(IOError):
        next(it)
    hdr = bytearray(raw); hdr[int(off[5]) - 12] ^= 1
    r = str(tmp_path / "badlen.tfrecord"); open(r, "wb").write(hdr)
    it = I.eval_input_fn(r, parser, batch_size=2)
    assert _ids([next(it), next(it)]).tolist() == [0, 1, 2, 3]        # record 5 has the bad length: 0..3 fill two batches
    with pytest.raises(IOError, match="corrupted record length"):
        next(it)
    with pytest.raises(IOError):
        next(I.eval_input_fn(r, parser, batch_size=16))                 # the first batch already needs record 5
    cut = str(tmp_path / "cut.tfrecord"); open(cut, "wb").write(bytes(raw[:int(off[40]) + 3]))   # file ends inside record 40
    it = I.train_input_fn(cut, parser, batch_size=8, num_epochs=2, shuffle_buffer_size=0)
    assert _ids([next(it) for _ in range(5)]).tolist() == list(range(40))
    with pytest.raises(IOError, match="truncated"):
        next(it)


def _shuffle_order_py(n, buffer_size, draws):
    """The buffer walk of dataset.shuffle written out in Python: the readable twin ctr_feed_shuffle_order is compared with."""
    buf = list(range(min(buffer_size, n)))
    out, nxt, filled = [], len(buf), len(buf)
    for i in range(n):
        j = int(draws[i] * filled)
        out.append(buf[j])
        if nxt < n:
            buf[j] = nxt; nxt += 1
        else:
            filled -= 1; buf[j] = buf[filled]
    return out


def test_native_shuffle_order_is_the_buffer_walk():
    from recalgorithm_b200.io import native
    rng = np.random.default_rng(5)
    for n, bs in [(2, 2), (10, 3), (1000, 7), (1000, 999), (1000, 1000), (1000, 5000), (50_000, 10_000)]:
        draws = rng.random(n)
        got = native.shuffle_order(n, bs, draws)
        assert got.tolist() == _shuffle_order_py(n, bs, draws)
        assert sorted(got.tolist()) == list(range(n))                # a permutation
        pos = np.empty(n, np.int64); pos[got] = np.arange(n)
        assert (np.arange(n) - pos < bs).all()                        # element i cannot be emitted before input i - bs + 1 arrived
    assert native.shuffle_order(5, 1, np.zeros(5)).tolist() == [0, 1, 2, 3, 4]
    assert native.shuffle_order(0, 10, np.zeros(0)).size == 0
    with pytest.raises(ValueError):
        native.shuffle_order(5, 3, np.array([0.1, 0.2, 1.0, 0.3, 0.4]))   # a draw outside [0,1)
    with pytest.raises(ValueError):
        native.shuffle_order(5, 3, np.zeros(2))
    a = I.shuffle_order(20_000, 100, np.random.default_rng(3)); b = I.shuffle_order(20_000, 100, np.random.default_rng(3))
    assert a.tolist() == b.tolist() and a.tolist() != list(range(20_000))


def test_pad_ragged_is_the_row_by_row_copy():
    rng = np.random.default_rng(2)
    for B, tmax in [(0, 3), (1, 0), (5, 4), (300, 50)]:
        lens = rng.integers(0, tmax + 1, B)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        vals = rng.integers(0, 1000, int(off[-1])).astype(np.int64)
        width = max(int(lens.max()) if B else 0, 1)
        want = np.full((B, width), -1, np.int64)
        for b in range(B):
            want[b, :lens[b]] = vals[off[b]:off[b + 1]]
        assert np.array_equal(fc.pad_ragged(vals, off, width), want)
    with pytest.raises(ValueError):
        fc.pad_ragged(np.arange(3), np.array([0, 3]), 2)
    # offsets that do not start at 0 (a slice of a larger ragged array)
    assert fc.pad_ragged(np.arange(10), np.array([4, 6, 6, 9]), 3).tolist() == [[4, 5, -1], [-1, -1, -1], [6, 7, 8]]


def test_resumable_shuffle_equals_the_one_shot_walk_however_the_calls_are_cut():
    from recalgorithm_b200.io import native
    rng = np.random.default_rng(9)
    for n, bs in [(1, 5), (7, 3), (500, 50), (500, 500), (500, 9000), (20_000, 1000)]:
        draws = rng.random(n)
        want = native.shuffle_order(n, bs, draws) if n > 1 else np.zeros(1, np.int64)
        sh, got, used, avail = native.Shuffler(bs), [], 0, 0
        while True:                                                   # the input grows in random steps; emission in random bites
            avail = min(n, avail + int(rng.integers(0, 2 * bs + 3)))
            done = avail == n and rng.random() < 0.7
            out = sh.emit(avail, done, draws[used:], int(rng.integers(1, 300)))
            assert out.size == 0 or out.max() < avail                 # never names an input that has not been seen
            got.extend(out.tolist()); used += out.size
            if done and out.size == 0:
                break
        assert got == want.tolist() and used == n


def test_streaming_index_equals_the_one_shot_index(dataset, tmp_path):
    from recalgorithm_b200.io import native
    p, n, parser = dataset
    buf, off, ln = native.read_tfrecord_file(p)
    for kw in (dict(chunk_bytes=1), dict(chunk_bytes=37), dict(chunk_bytes=1 << 20), dict(mmap=True, chunk_records=1),
               dict(mmap=True, chunk_records=7)):
        ix = native.StreamingIndex(p, **kw)
        assert ix.wait_all() == n and ix.done and not ix.failed
        assert np.array_equal(ix.off[:n], off) and np.array_equal(ix.ln[:n], ln) and np.array_equal(np.asarray(ix.buf), buf)
        got, _ = ix.wait_for(10)
        assert got == n
    empty = str(tmp_path / "empty.tfrecord"); open(empty, "wb").close()
    for mm in (False, True):
        ix = native.StreamingIndex(empty, mmap=mm)
        assert ix.wait_all() == 0
    assert list(I.eval_input_fn(empty, parser, batch_size=4)) == []


def test_streamed_first_epoch_is_the_order_of_a_finished_index(dataset, monkeypatch):
    """The first epoch runs while the file is still being indexed; its order must not depend on how far the index happens to be:
    same seed -> the same batches as when the index was complete before the first element was asked for."""
    from recalgorithm_b200.io import native
    p, n, parser = dataset

    def run(slow):
        orig = native.StreamingIndex.__init__

        def init(self, path, mmap=False, **kw):
            orig(self, path, mmap=mmap, chunk_bytes=64 if slow else 1 << 25, chunk_records=3 if slow else 1 << 16)
            if not slow:
                self.wait_all()
        monkeypatch.setattr(native.StreamingIndex, "__init__", init)
        try:
            return _ids(list(I.train_input_fn(p, parser, batch_size=16, num_epochs=3, shuffle_buffer_size=10, seed=4))).tolist()
        finally:
            monkeypatch.setattr(native.StreamingIndex, "__init__", orig)
    a, b = run(True), run(False)
    assert a == b and sorted(a[:n]) == list(range(n)) and a[:n] != list(range(n))
    for mm in (False, True):
        e = _ids(list(I.eval_input_fn(p, parser, batch_size=16, mmap=mm))).tolist()
        assert e == list(range(n))
