"""The reference's input_fns (algorithm/utils.py:4-47) on top of the native feeder -- the tf.data side of the hot path.

    parser = make_example_parser(total_feature_columns + label_feature_columns, label_keys=["read_comment"])   # DCN/dcn.py:116-131
    for features, labels in train_input_fn(path, parser, batch_size=1024, num_epochs=1, shuffle_buffer_size=10000):
        ...   # features: key -> (ids int64, row_offsets) | float32 array, exactly what feature_column.input_layer takes

Pipeline order is the reference's: TFRecordDataset -> shuffle(buffer) -> repeat(num_epochs) -> batch(batch_size) ->
map(example_parser) -> prefetch(1).  Records are never copied: a batch is (file bytes, offsets[idx], lengths[idx]) and
libctr_feed.so parses it in place (multi-threaded, GIL released) while the previous batch is being consumed.
"""
from __future__ import annotations

import queue
import threading
from typing import Callable, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import feature_column as fc
from .io import native

Batch = Tuple[np.ndarray, np.ndarray, np.ndarray]            # (file bytes uint8, offsets uint64 (B,), lengths uint64 (B,))


def make_example_parser(feature_columns, label_keys: Sequence[str] = ("read_comment",), read_feature_lists: bool = False,
                        num_threads: int = 0) -> Callable[[Batch], Tuple[dict, dict]]:
    """`example_parser` of the reference's model files (e.g. DCN/dcn.py:116-131): make_parse_example_spec over the feature
    and label columns, tf.parse_example on the serialized batch, labels popped into their own dict."""
    def example_parser(serialized: Batch):
        buf, off, ln = serialized
        features = fc.parse_example_native(buf, off, ln, feature_columns, read_feature_lists=read_feature_lists, num_threads=num_threads)
        labels = {k: features.pop(k) for k in label_keys}
        return features, labels
    return example_parser


class _LoadedIndex:
    """Same face as native.StreamingIndex for records that are already indexed (several files, concatenated)."""

    def __init__(self, buf, off, ln):
        self.buf, self.off, self.ln, self.n, self.done, self.capacity = buf, off, ln, int(off.size), True, int(off.size)

    def wait_for(self, count):
        return self.n, True

    def wait_all(self):
        return self.n

    failed = False

    def check(self):
        pass

    def close(self):
        pass


def _open(filepath: Union[str, Sequence[str]], mmap: bool = False):
    """TFRecordDataset(filepath).  One file (the reference's case, utils.py:18): a native.StreamingIndex -- the file is read (or
    mapped) and indexed by a background thread while the first batches are already being parsed.  Several files are read one
    after the other (TFRecordDataset([files]) order) and concatenated, which materialises them.  Length CRCs are checked by the
    index scan, payload CRCs per batch in `_batches` (inside the prefetch thread): a corrupted record raises when the batch that
    holds it is produced -- TFRecordDataset's DataLossError timing -- and the CRC pass overlaps the consumer."""
    paths = [filepath] if isinstance(filepath, str) else list(filepath)
    if len(paths) == 1:
        return native.StreamingIndex(paths[0], mmap=mmap)
    bufs, offs, lens, base = [], [], [], 0
    for p in paths:
        b, o, l = native.read_tfrecord_file(p, verify="headers", mmap=mmap)
        bufs.append(b); offs.append(o + np.uint64(base)); lens.append(l)
        base += b.size
    return _LoadedIndex(np.concatenate(bufs), np.concatenate(offs), np.concatenate(lens))


def _load(filepath: Union[str, Sequence[str]], mmap: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(file bytes, offsets, lengths) of all records, complete (blocks until the index is)."""
    ix = _open(filepath, mmap)
    n = ix.wait_all()
    return ix.buf, ix.off[:n], ix.ln[:n]


def shuffle_order(n: int, buffer_size: int, rng: np.random.Generator) -> np.ndarray:
    """Order in which dataset.shuffle(buffer_size) emits n elements: a buffer of the next `buffer_size` elements, one of them
    drawn uniformly at each step and replaced by the next input element (tf.data semantics; buffer >= n = a uniform permutation).
    The walk is native (ctr_feed_shuffle_order: a Python loop costs ~0.7 us per record, more than parsing the record); the n
    draws come from the seeded numpy generator, so the order is reproducible per seed."""
    if buffer_size <= 1 or n <= 1:
        return np.arange(n, dtype=np.int64)
    return native.shuffle_order(n, buffer_size, rng.random(n))


class _Draws:
    """The generator's uniform draws as ONE stream, one draw per shuffled element, whatever the block sizes they are fetched in
    (numpy's Generator.random(a) followed by random(b) is random(a + b)): a streamed epoch, whose block sizes depend on timing,
    consumes exactly the draws shuffle_order would."""

    def __init__(self, rng: np.random.Generator):
        self.rng, self.buf = rng, np.empty(0, np.float64)

    def peek(self, k: int) -> np.ndarray:
        if self.buf.size < k:
            self.buf = np.concatenate([self.buf, self.rng.random(max(k - self.buf.size, 1 << 16))])
        return self.buf[:k]

    def consume(self, k: int):
        self.buf = self.buf[k:]


def _order_pieces(index, shuffle_buffer_size: int, draws: Optional[_Draws], rounds, piece: int) -> Iterator[np.ndarray]:
    """shuffle(buffer).repeat(rounds) as a stream of index arrays.  While the file is still being indexed (first epoch) the
    pieces follow the index: element i of the shuffled order only needs inputs up to i + buffer_size."""
    shuffled = shuffle_buffer_size > 1
    for _ in rounds:
        if index.done and not index.failed:                       # (a failed scan takes the streamed path: it delivers the
            n = index.wait_all()                                   # records in front of the damage, then raises)
            if shuffled and n > 1:
                order = native.shuffle_order(n, shuffle_buffer_size, draws.peek(n)); draws.consume(n)
            else:
                order = np.arange(n, dtype=np.int64)
            yield order
            continue
        emitted = 0
        sh = native.Shuffler(shuffle_buffer_size) if shuffled else None
        while True:
            want = emitted + piece + (shuffle_buffer_size if shuffled else 0)
            avail, done = index.wait_for(want)
            if shuffled:
                idx = sh.emit(avail, done, draws.peek(piece), piece)
                draws.consume(idx.size)
            else:
                idx = np.arange(emitted, min(avail, emitted + piece), dtype=np.int64)
            if idx.size:
                emitted += idx.size
                yield idx
            elif done:
                index.check()                                      # a corrupted length / truncated file ends the epoch with its error
                break


def _prefetch(gen: Iterator, depth: int = 1, on_close: Optional[Callable[[], None]] = None) -> Iterator:
    """dataset.prefetch(depth): a producer thread keeps `depth` parsed batches ahead of the consumer."""
    q: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
    stop = threading.Event()
    END = object()

    def producer():
        try:
            for item in gen:
                while not stop.is_set():
                    try:
                        q.put(item, timeout=0.1)
                        break
                    except queue.Full:
                        continue
                if stop.is_set():
                    return
            q.put(END)
        except BaseException as e:                           # surfaces in the consumer
            q.put(e)

    th = threading.Thread(target=producer, daemon=True)
    th.start()
    try:
        while True:
            item = q.get()
            if item is END:
                return
            if isinstance(item, BaseException):
                raise item
            yield item
    finally:
        stop.set()
        if on_close is not None:
            on_close()


def _batches(order_pieces: Iterator[np.ndarray], index, batch_size: int, parser) -> Iterator:
    buf, off, ln = index.buf, index.off, index.ln            # off / ln are valid for every position an order piece names
    verified = np.zeros(index.capacity, dtype=bool)          # payload CRCs are checked the first time a record is used

    def parse(idx):
        o, l = off[idx], ln[idx]
        new = ~verified[idx]
        if new.any():
            native.verify_records(buf, o[new], l[new])
            verified[idx] = True
        return parser((buf, o, l))

    pending: List[np.ndarray] = []
    have = 0
    for order in order_pieces:                               # repeat() happens BEFORE batch(): batches run across epoch borders
        pending.append(order); have += order.size
        while have >= batch_size:
            idx = np.concatenate(pending) if len(pending) > 1 else pending[0]
            take, rest = idx[:batch_size], idx[batch_size:]
            pending, have = ([rest] if rest.size else []), rest.size
            yield parse(take)
    if have:                                                 # drop_remainder=False
        idx = np.concatenate(pending) if len(pending) > 1 else pending[0]
        yield parse(idx)


def train_input_fn(filepath, example_parser, batch_size: int, num_epochs: Optional[int], shuffle_buffer_size: int,
                   seed: Optional[int] = None, mmap: bool = False) -> Iterator[Tuple[dict, dict]]:
    """utils.py:4-26.  Iterating the result is `dataset.make_one_shot_iterator()`; each epoch is reshuffled
    (tf.data's reshuffle_each_iteration default).  mmap=True maps the TFRecord file instead of reading it into RAM."""
    index = _open(filepath, mmap=mmap)
    import itertools
    # num_epochs=None repeats forever, like dataset.repeat(None) (Estimator callers bound the run with max_steps)
    rounds = itertools.count() if num_epochs is None else range(num_epochs)
    pieces = _order_pieces(index, shuffle_buffer_size, _Draws(np.random.default_rng(seed)), rounds, max(batch_size, 1 << 13))
    return _prefetch(_batches(pieces, index, batch_size, example_parser), depth=1, on_close=index.close)


def eval_input_fn(filepath, example_parser, batch_size: int, mmap: bool = False) -> Iterator[Tuple[dict, dict]]:
    """utils.py:29-47: one pass, file order, no shuffle."""
    index = _open(filepath, mmap=mmap)
    pieces = _order_pieces(index, 0, None, range(1), max(batch_size, 1 << 13))
    return _prefetch(_batches(pieces, index, batch_size, example_parser), depth=1, on_close=index.close)


class DevicePrefetcher:
    """Last stage of the feeding side: move each parsed batch to the GPU through pinned staging buffers on a side stream,
    one batch ahead of the consumer (the H2D copy of batch i+1 overlaps the kernels of batch i).

        for ids, dense, labels in DevicePrefetcher(train_input_fn(...), categorical_columns, dense_keys, label_keys):
            tile, fm2 = autograd.lookup_fm2(tables, ids)          # ids: (B, F) int64 on the device, -1 = missing / OOV

    `categorical_columns` must be single-valued (feature_column.single_valued_ids); multi-valued columns stay on the
    feature_column.input_layer path.  Yields (ids (B,F) int64, {key: float32 (B,w)}, {key: float32 (B,w)}) device tensors."""

    def __init__(self, batches, categorical_columns, dense_keys=(), label_keys=("read_comment",), device="cuda"):
        import torch
        self.torch = torch
        self.batches = iter(batches)
        self.cols, self.dense_keys, self.label_keys = list(categorical_columns), list(dense_keys), list(label_keys)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._pinned = [{}, {}]                                # two staging sets: one being copied, one being filled
        self._slot_ev = [None, None]                           # last H2D copy out of each set (host waits before refilling it)
        self._turn = 0

    def _stage(self, name: str, arr: np.ndarray):
        torch = self.torch
        slot = self._pinned[self._turn]
        t = slot.get(name)
        if t is None or t.shape != arr.shape or t.dtype != torch.from_numpy(arr).dtype:
            t = torch.empty(arr.shape, dtype=torch.from_numpy(arr).dtype).pin_memory()
            slot[name] = t
        t.copy_(torch.from_numpy(arr))
        return t.to(self.device, non_blocking=True)

    def _issue(self):
        try:
            features, labels = next(self.batches)
        except StopIteration:
            return None
        torch = self.torch
        ids = fc.single_valued_ids(features, self.cols)
        if self._slot_ev[self._turn] is not None:
            self._slot_ev[self._turn].synchronize()            # the copy that last read this staging set has finished
        with torch.cuda.stream(self.stream):
            out = (self._stage("ids", ids),
                   {k: self._stage("d:" + k, np.ascontiguousarray(features[k], np.float32)) for k in self.dense_keys},
                   {k: self._stage("l:" + k, np.ascontiguousarray(labels[k], np.float32)) for k in self.label_keys})
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._slot_ev[self._turn] = ev
        self._turn ^= 1
        return out, ev

    def __iter__(self):
        torch = self.torch
        nxt = self._issue()
        while nxt is not None:
            cur, ev = nxt
            torch.cuda.current_stream(self.device).wait_event(ev)
            for t in [cur[0], *cur[1].values(), *cur[2].values()]:
                t.record_stream(torch.cuda.current_stream(self.device))
            nxt = self._issue()                                # batch i+1 starts copying before batch i is consumed
            yield cur
