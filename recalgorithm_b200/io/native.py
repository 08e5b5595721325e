"""ctypes binding of libctr_feed.so (include/ctr_feed.h): the native twins of tfrecord.py / example.py / vocab.py.

    buf, offsets, lengths = native.read_tfrecord_file(path)                # index a whole file (CRC-32C verified)
    vocab = native.Vocabulary(path_or_tokens)                              # OOV / '' -> -1
    out = native.parse_examples(buf, offsets[:B], lengths[:B], {"userid": vocab, ...}, {"read_comment": (1, 0.0)})
    ids, row_offsets = out["userid"]                                       # ragged int64 ids, (B+1,) offsets

Same semantics as recalgorithm_b200.io.parse_example + VocabularyFile.lookup (tests/test_feed_native.py compares them),
multi-threaded, no Python in the per-record loop.  The library is built in-tree by recalgorithm_b200/build.py.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CTR_FEED_LIB") or os.path.join(os.path.dirname(_HERE), "csrc_feed", "libctr_feed.so")   # override: sanitizer builds

ERR_ARG, ERR_IO, ERR_TRUNCATED, ERR_CRC, ERR_PROTO, ERR_CAPACITY, ERR_NOMEM = -1, -2, -4, -5, -6, -7, -8


class _Cat(ctypes.Structure):
    _fields_ = [("key", ctypes.c_char_p), ("vocab", ctypes.c_void_p), ("ids", ctypes.c_void_p), ("capacity", ctypes.c_int64),
                ("row_offsets", ctypes.c_void_p), ("needed", ctypes.c_int64)]


class _Dense(ctypes.Structure):
    _fields_ = [("key", ctypes.c_char_p), ("width", ctypes.c_int64), ("default_value", ctypes.c_float), ("out", ctypes.c_void_p)]


_P, _I = ctypes.c_void_p, ctypes.c_int64
SIGNATURES = {
    "ctr_feed_last_error": (ctypes.c_char_p, []),
    "ctr_feed_version": (ctypes.c_int, []),
    "ctr_feed_crc32c": (ctypes.c_uint32, [_P, ctypes.c_uint64]),
    "ctr_feed_masked_crc32c": (ctypes.c_uint32, [_P, ctypes.c_uint64]),
    "ctr_feed_tfrecord_index": (_I, [_P, ctypes.c_uint64, ctypes.c_int, _P, _P, _I, _P]),
    "ctr_feed_tfrecord_index_from": (_I, [_P, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, _P, _P, _I, _P]),
    "ctr_feed_tfrecord_verify": (ctypes.c_int, [_P, ctypes.c_uint64, _P, _P, _I, ctypes.c_int]),
    "ctr_feed_shuffle_order": (ctypes.c_int, [_I, _I, _P, _P]),
    "ctr_feed_shuffle_create": (_P, [_I]),
    "ctr_feed_shuffle_destroy": (None, [_P]),
    "ctr_feed_shuffle_emit": (_I, [_P, _I, ctypes.c_int, _P, _I, _P]),
    "ctr_feed_vocab_create": (_P, [_P, _P, _I]),
    "ctr_feed_vocab_load": (_P, [ctypes.c_char_p]),
    "ctr_feed_vocab_size": (_I, [_P]),
    "ctr_feed_vocab_destroy": (None, [_P]),
    "ctr_feed_vocab_lookup": (ctypes.c_int, [_P, _P, _P, _I, _P]),
    "ctr_feed_parse_examples": (ctypes.c_int, [_P, _P, _P, _I, ctypes.POINTER(_Cat), _I, ctypes.POINTER(_Dense), _I, ctypes.c_int,
                                               ctypes.c_int]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m recalgorithm_b200.build` (there is no fallback inside this module; "
                               "the pure-Python readers live in recalgorithm_b200.io.tfrecord / example / vocab)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


class FeedError(RuntimeError):
    def __init__(self, code: int):
        super().__init__(f"libctr_feed error {code}: {lib().ctr_feed_last_error().decode('utf-8', 'replace')}")
        self.code = code


class FeedValueError(FeedError, ValueError):
    """Malformed wire data or a feature of the wrong kind / size (tf.parse_example raises InvalidArgumentError)."""


class FeedIOError(FeedError, IOError):
    """Truncated or corrupted TFRecord."""


def _raise(code: int):
    if code in (ERR_TRUNCATED, ERR_CRC, ERR_IO):
        raise FeedIOError(code)
    if code in (ERR_PROTO, ERR_ARG):
        raise FeedValueError(code)
    if code == ERR_NOMEM:
        raise MemoryError(f"libctr_feed: {lib().ctr_feed_last_error().decode('utf-8', 'replace')}")
    raise FeedError(code)


def _u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):                             # includes np.memmap
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        return data
    return np.frombuffer(data, dtype=np.uint8)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None or a.size == 0 else a.ctypes.data


def crc32c(data) -> int:
    a = _u8(data)
    return int(lib().ctr_feed_crc32c(_ptr(a), a.size))


def masked_crc32c(data) -> int:
    a = _u8(data)
    return int(lib().ctr_feed_masked_crc32c(_ptr(a), a.size))


def index_tfrecord(buf, verify=True, num_threads: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """(offsets, lengths) uint64 arrays of the record payloads inside a TFRecord byte buffer, found in ONE sequential scan
    (the output arrays are sized by the 16-bytes-per-record lower bound; untouched pages are never committed).
    verify=True: length CRCs during the scan, payload CRCs afterwards on `num_threads` threads (0 = all cores);
    verify="headers": length CRCs only -- the caller checks payloads later with verify_records (input_fn does it per batch, in
    the prefetch thread, so that the CRC pass overlaps the consumer); verify=False: nothing is checked."""
    a = _u8(buf)
    if verify not in (True, False, "headers"):
        raise ValueError(f"verify must be True, False or 'headers', got {verify!r}")
    bound = a.size // 16 + 1                                          # header 12 + footer 4 bytes: no record is shorter
    offsets, lengths = np.empty(bound, np.uint64), np.empty(bound, np.uint64)
    consumed = ctypes.c_uint64(0)
    n = lib().ctr_feed_tfrecord_index(_ptr(a), a.size, 2 if verify else 0, offsets.ctypes.data, lengths.ctypes.data, bound,
                                      ctypes.byref(consumed))
    if n < 0:
        _raise(int(n))
    assert consumed.value == a.size, "the 16-byte bound cannot be reached before the buffer ends"
    offsets, lengths = offsets[:n].copy(), lengths[:n].copy()         # drop the over-sized allocations
    if verify is True:
        verify_records(a, offsets, lengths, num_threads)
    return offsets, lengths


def verify_records(buf, offsets: np.ndarray, lengths: np.ndarray, num_threads: int = 0) -> None:
    """Check the payload CRC of the given records (any subset, any order) on `num_threads` threads (0 = all cores); raises
    FeedIOError naming the first corrupted one."""
    a = _u8(buf)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint64)
    if offsets.size == 0:
        return
    rc = lib().ctr_feed_tfrecord_verify(_ptr(a), a.size, offsets.ctypes.data, lengths.ctypes.data, int(offsets.size), int(num_threads))
    if rc < 0:
        _raise(int(rc))


def shuffle_order(n: int, buffer_size: int, draws: np.ndarray) -> np.ndarray:
    """Emission order of dataset.shuffle(buffer_size) over n elements for the given n uniform draws (ctr_feed_shuffle_order)."""
    draws = np.ascontiguousarray(draws, dtype=np.float64)
    if draws.size < n:
        raise ValueError(f"shuffle_order needs {n} draws, got {draws.size}")
    out = np.empty(n, np.int64)
    rc = lib().ctr_feed_shuffle_order(int(n), int(buffer_size), draws.ctypes.data, out.ctypes.data)
    if rc < 0:
        _raise(int(rc))
    return out


class Shuffler:
    """dataset.shuffle(buffer_size) over an input whose length is not known yet (ctr_feed_shuffle_create / _emit): positions
    come out as far as the input seen so far allows, one caller-supplied draw per emission, in the order
    shuffle_order(n, buffer_size, draws) gives for the same draws however the calls are cut."""

    def __init__(self, buffer_size: int):
        self._h = lib().ctr_feed_shuffle_create(int(buffer_size))
        if not self._h:
            raise FeedError(ERR_ARG)

    def emit(self, n_available: int, input_done: bool, draws: np.ndarray, max_out: int) -> np.ndarray:
        """Up to max_out positions; draws[:len(result)] are consumed."""
        draws = np.ascontiguousarray(draws, dtype=np.float64)
        max_out = min(int(max_out), int(draws.size))
        out = np.empty(max_out, np.int64)
        k = int(lib().ctr_feed_shuffle_emit(self._h, int(n_available), int(bool(input_done)), draws.ctypes.data, max_out,
                                            out.ctypes.data))
        if k < 0:
            _raise(k)
        return out[:k]

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ctr_feed_shuffle_destroy(self._h)
            self._h = None


class StreamingIndex:
    """The records of ONE TFRecord file, indexed by a background thread while the consumer already parses the front of it:
    `off[:n]` / `ln[:n]` are valid for the `n` that `wait_for` returns.  mmap=False reads the file in `chunk_bytes` pieces into
    one buffer and indexes each piece as it arrives (read, scan and parse overlap); mmap=True maps it and scans it in runs of
    `chunk_records`.  Length CRCs are checked by the scan, payload CRCs by the consumer (verify_records) -- a corrupted or
    truncated record raises when the consumer asks for records at or beyond it, like TFRecordDataset's DataLossError."""

    def __init__(self, path: str, mmap: bool = False, chunk_bytes: int = 32 << 20, chunk_records: int = 1 << 16):
        size = os.path.getsize(path)
        self.path, self.size = path, size
        self.capacity = size // 16 + 1                            # header 12 + footer 4 bytes: no record is shorter
        self.off, self.ln = np.empty(self.capacity, np.uint64), np.empty(self.capacity, np.uint64)   # pages commit when touched
        self.buf = (np.memmap(path, dtype=np.uint8, mode="r") if mmap and size > 0 else np.empty(size, np.uint8))
        self._mmap, self._chunk_bytes, self._chunk_records = bool(mmap and size > 0), int(chunk_bytes), int(chunk_records)
        self.n, self.done, self._error, self._stop = 0, False, None, False
        self._cv = threading.Condition()
        self._thread = threading.Thread(target=self._run, daemon=True, name="ctr-feed-index")
        self._thread.start()

    def _scan(self, have: int, pos: int, final: bool) -> int:
        """Index buf[pos:have); returns the new scan position."""
        L = lib()
        consumed = ctypes.c_uint64(pos)
        while pos < have and not self._stop:
            n = self.n
            k = L.ctr_feed_tfrecord_index_from(self.buf.ctypes.data, have, pos, 2, 0 if final else 1,
                                               self.off[n:].ctypes.data, self.ln[n:].ctypes.data,
                                               min(self._chunk_records, self.capacity - n), ctypes.byref(consumed))
            if k < 0:                                             # damaged record at byte consumed.value: keep what lies in front
                try:
                    _raise(int(k))
                except FeedError as e:
                    good = L.ctr_feed_tfrecord_index_from(self.buf.ctypes.data, consumed.value, pos, 2, 0,
                                                          self.off[n:].ctypes.data, self.ln[n:].ctypes.data, self.capacity - n, None)
                    with self._cv:
                        self.n = n + max(0, int(good))
                    raise e
            if k == 0 and consumed.value == pos:
                break                                             # an incomplete tail: wait for more bytes
            pos = int(consumed.value)
            with self._cv:
                self.n = n + int(k)
                self._cv.notify_all()
        return pos

    def _run(self):
        try:
            if self._mmap or self.size == 0:
                self._scan(self.size, 0, True)
            else:
                have = pos = 0
                with open(self.path, "rb", buffering=0) as f:
                    while have < self.size and not self._stop:
                        got = f.readinto(memoryview(self.buf)[have:have + self._chunk_bytes])
                        if not got:
                            raise IOError(f"{self.path}: file ended at byte {have} of {self.size} while being read")
                        have += got
                        pos = self._scan(have, pos, have >= self.size)
        except BaseException as e:                                # surfaces in the consumer (wait_for)
            self._error = e
        finally:
            with self._cv:
                self.done = True
                self._cv.notify_all()

    def wait_for(self, count: int) -> Tuple[int, bool]:
        """Block until `count` records are indexed or the scan has ended; returns (records indexed so far, ended).  Does not
        raise: the records in front of a corrupted one are still delivered -- call check() once they are used up."""
        with self._cv:
            while self.n < count and not self.done:
                self._cv.wait()
            return self.n, self.done

    @property
    def failed(self) -> bool:
        return self._error is not None

    def check(self):
        """Raise the error that ended the scan early, if any (corrupted length, truncated file)."""
        if self._error is not None:
            raise self._error

    def wait_all(self) -> int:
        n, _ = self.wait_for(self.capacity + 1)
        self.check()
        return n

    def close(self):
        self._stop = True


def read_tfrecord_file(path: str, verify=True, mmap: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Whole file -> (bytes as uint8 array, offsets, lengths).  tf.data.TFRecordDataset(path) without the iterator.
    mmap=True maps the file instead of reading it (files larger than RAM; pages are pulled in by the scan / CRC / parse passes);
    verify as in index_tfrecord."""
    buf = np.memmap(path, dtype=np.uint8, mode="r") if mmap and os.path.getsize(path) > 0 else np.fromfile(path, dtype=np.uint8)
    offsets, lengths = index_tfrecord(buf, verify)
    return buf, offsets, lengths


def _blob(items: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    offsets = np.zeros(len(items) + 1, np.uint64)
    if len(items):
        np.cumsum([len(x) for x in items], out=offsets[1:])
    return np.frombuffer(b"".join(items), dtype=np.uint8), offsets


class Vocabulary:
    """categorical_column_with_vocabulary_file semantics: id = 0-based line of the first occurrence, OOV -> -1."""

    def __init__(self, source: Union[str, Sequence[bytes]]):
        L = lib()
        if isinstance(source, str):
            self._h = L.ctr_feed_vocab_load(source.encode())
        else:
            toks = [t if isinstance(t, bytes) else str(t).encode() for t in source]
            blob, offs = _blob(toks)
            self._h = L.ctr_feed_vocab_create(_ptr(blob), offs.ctypes.data, len(toks))
        if not self._h:
            raise FeedIOError(ERR_IO)
        self.size = int(L.ctr_feed_vocab_size(self._h))

    def __len__(self) -> int:
        return self.size

    def lookup(self, keys: Sequence[bytes]) -> np.ndarray:
        keys = list(keys)
        blob, offs = _blob(keys)
        out = np.empty(len(keys), np.int64)
        rc = lib().ctr_feed_vocab_lookup(self._h, _ptr(blob), offs.ctypes.data, len(keys), _ptr(out))
        if rc:
            _raise(rc)
        return out

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ctr_feed_vocab_destroy(self._h)
            self._h = None


def parse_examples(buf, offsets: np.ndarray, lengths: np.ndarray, categorical: Dict[str, Vocabulary],
                   dense: Optional[Dict[str, Tuple[int, float]]] = None, read_feature_lists: bool = False,
                   num_threads: int = 0) -> Dict[str, object]:
    """tf.parse_example + vocabulary lookup over records ``buf[offsets[b] : offsets[b] + lengths[b]]``.

    categorical: key -> Vocabulary  => out[key] = (ids int64 (n,), row_offsets int64 (B+1,))   (missing key: empty row)
    dense: key -> (width, default)  => out[key] = float32 (B, width)"""
    a = _u8(buf)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint64)
    B = int(offsets.size)
    dense = dense or {}
    ckeys, dkeys = list(categorical), list(dense)
    cats = (_Cat * max(1, len(ckeys)))()
    dens = (_Dense * max(1, len(dkeys)))()
    cap = {k: max(16, 2 * B) for k in ckeys}
    row_off = {k: np.zeros(B + 1, np.int64) for k in ckeys}
    dense_out = {k: np.empty((B, int(dense[k][0])), np.float32) for k in dkeys}
    for i, k in enumerate(dkeys):
        dens[i] = _Dense(k.encode(), int(dense[k][0]), float(dense[k][1]), dense_out[k].ctypes.data)
    for _ in range(2):                                       # second round only if a ragged buffer was too small
        ids = {k: np.empty(cap[k], np.int64) for k in ckeys}
        for i, k in enumerate(ckeys):
            cats[i] = _Cat(k.encode(), categorical[k]._h, ids[k].ctypes.data, cap[k], row_off[k].ctypes.data, 0)
        rc = lib().ctr_feed_parse_examples(_ptr(a), _ptr(offsets), _ptr(lengths), B, cats, len(ckeys), dens, len(dkeys),
                                           int(read_feature_lists), int(num_threads))
        if rc == ERR_CAPACITY:
            for i, k in enumerate(ckeys):
                cap[k] = max(cap[k], int(cats[i].needed))
            continue
        if rc:
            _raise(rc)
        break
    else:                                                    # `needed` is exact, so the second round cannot come up short
        _raise(ERR_CAPACITY)
    out: Dict[str, object] = dict(dense_out)
    for i, k in enumerate(ckeys):
        out[k] = (ids[k][: int(cats[i].needed)], row_off[k])
    return out
