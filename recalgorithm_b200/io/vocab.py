"""Vocabulary files: one token per line, id = 0-based line number, out-of-vocabulary -> -1
(fc.categorical_column_with_vocabulary_file with num_oov_buckets=0, default_value=None -- DeepFM/deepfm.py:56-64;
files written by dataset/wechat_algo_data1/DataGenerator.py:206-210).  [TF-internal semantics, SURVEY A.4]"""
from __future__ import annotations

from typing import Iterable, Sequence, Union

import numpy as np


class VocabularyFile:
    def __init__(self, source: Union[str, Sequence[bytes]]):
        if isinstance(source, str):
            with open(source, "rb") as f:
                tokens = [ln.rstrip(b"\r\n") for ln in f]
            if tokens and tokens[-1] == b"":
                tokens.pop()
        else:
            tokens = [t if isinstance(t, bytes) else str(t).encode() for t in source]
        self.size = len(tokens)                       # vocabulary_size=None -> number of lines
        self.tokens = tokens                          # kept for the native twin (io.native.Vocabulary), built on demand
        self._native = None
        self._table = {}
        for i, t in enumerate(tokens):
            self._table.setdefault(t, i)

    def __len__(self) -> int:
        return self.size

    def native(self):
        """The same vocabulary inside libctr_feed.so (used by feature_column.parse_example_native)."""
        if self._native is None:
            from . import native as _n
            self._native = _n.Vocabulary(self.tokens)
        return self._native

    def lookup(self, keys: Iterable[bytes]) -> np.ndarray:
        get = self._table.get
        return np.fromiter((get(k, -1) for k in keys), dtype=np.int64)
