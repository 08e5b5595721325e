"""Minimal protobuf-wire codec for tf.train.Example / SequenceExample (public example.proto / feature.proto):

    Example{Features features=1}        SequenceExample{Features context=1; FeatureLists feature_lists=2}
    Features{map<string,Feature> feature=1}   Feature{oneof{BytesList=1; FloatList=2; Int64List=3}}
    BytesList{repeated bytes value=1}  FloatList{repeated float value=1 [packed]}  Int64List{repeated int64 value=1 [packed]}
    FeatureLists{map<string,FeatureList> feature_list=1}   FeatureList{repeated Feature feature=1}

Parity note (SURVEY 8a note 8): the reference's ETL writes SequenceExamples (DataGenerator.py:429-442) but every model
parses them with tf.parse_example, i.e. as Example.  `context` and `features` share field number 1, so the context features
parse; `feature_lists` (field 2) is an unknown field that is skipped -> the two sequence features come out EMPTY.
``parse_single(..., read_feature_lists=False)`` (the default) reproduces exactly that; ``True`` reads them.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple, Union

import numpy as np

Value = Union[List[bytes], List[float], List[int]]


# ------------------------------------------------------------------ wire primitives
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: bytes):
    """Iterate (field_number, wire_type, value) over one message; value is int (varint/fixed) or bytes (length-delimited)."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fnum, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fnum, wt, v


def _enc_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(fnum: int, payload: bytes) -> bytes:
    return _enc_varint((fnum << 3) | 2) + _enc_varint(len(payload)) + payload


# ------------------------------------------------------------------ decode
def _parse_feature(buf: bytes) -> Tuple[str, Value]:
    for fnum, wt, v in _fields(buf):
        if fnum == 1:       # BytesList
            return "bytes", [val for f, w, val in _fields(v) if f == 1]
        if fnum == 2:       # FloatList (packed or not)
            vals: List[float] = []
            for f, w, val in _fields(v):
                if f == 1 and w == 2:
                    vals.extend(struct.unpack(f"<{len(val) // 4}f", val))
                elif f == 1 and w == 5:
                    vals.append(struct.unpack("<f", val)[0])
            return "float", vals
        if fnum == 3:       # Int64List
            ints: List[int] = []
            for f, w, val in _fields(v):
                if f == 1 and w == 2:
                    p = 0
                    while p < len(val):
                        x, p = _varint(val, p)
                        ints.append(x - (1 << 64) if x >= 1 << 63 else x)
                elif f == 1 and w == 0:
                    ints.append(val - (1 << 64) if val >= 1 << 63 else val)
            return "int64", ints
    return "bytes", []       # empty Feature{}


def _parse_features(buf: bytes) -> Dict[str, Tuple[str, Value]]:
    out: Dict[str, Tuple[str, Value]] = {}
    for fnum, wt, entry in _fields(buf):
        if fnum != 1:
            continue
        key, feat = "", b""
        for f, w, v in _fields(entry):       # map entry: key=1, value=2
            if f == 1:
                key = v.decode("utf-8")
            elif f == 2:
                feat = v
        out[key] = _parse_feature(feat)
    return out


def parse_single(serialized: bytes, read_feature_lists: bool = False):
    """-> (features: name -> (kind, values), feature_lists: name -> list of (kind, values)).

    Field 1 is ``Example.features`` == ``SequenceExample.context``; field 2 (feature_lists) is skipped unless asked for,
    which is what parsing a SequenceExample with tf.parse_example does."""
    features: Dict[str, Tuple[str, Value]] = {}
    flists: Dict[str, List[Tuple[str, Value]]] = {}
    for fnum, wt, v in _fields(serialized):
        if fnum == 1 and wt == 2:
            features.update(_parse_features(v))
        elif fnum == 2 and wt == 2 and read_feature_lists:
            for f, w, entry in _fields(v):
                if f != 1:
                    continue
                key, fl = "", b""
                for ff, ww, vv in _fields(entry):
                    if ff == 1:
                        key = vv.decode("utf-8")
                    elif ff == 2:
                        fl = vv
                flists[key] = [_parse_feature(x) for fff, www, x in _fields(fl) if fff == 1]
    return features, flists


@dataclass(frozen=True)
class VarLenFeature:          # categorical columns -> VarLenFeature(tf.string)  [make_parse_example_spec, SURVEY A.3]
    dtype: str = "bytes"


@dataclass(frozen=True)
class FixedLenFeature:        # numeric_column(key, default_value=0.0) -> FixedLenFeature((1,), float32, [0.0])
    shape: Tuple[int, ...] = (1,)
    dtype: str = "float"
    default_value: float = 0.0


def parse_example(serialized: Sequence[bytes], spec: Dict[str, Union[VarLenFeature, FixedLenFeature]],
                  read_feature_lists: bool = False):
    """tf.parse_example on a BATCH of serialized protos (the reference batches first: utils.py:22-23).

    VarLen keys -> (values list, offsets int64 (B+1)) ragged pair (the SparseTensor's content); a missing key gives an
    empty row.  FixedLen keys -> float32 array (B, *shape), default when missing.
    With read_feature_lists=True a VarLen key is also looked up in feature_lists (each step's values concatenated)."""
    B = len(serialized)
    ragged: Dict[str, Tuple[list, np.ndarray]] = {k: ([], np.zeros(B + 1, np.int64)) for k, s in spec.items()
                                                  if isinstance(s, VarLenFeature)}
    dense: Dict[str, np.ndarray] = {k: np.full((B,) + tuple(s.shape), s.default_value, np.float32)
                                    for k, s in spec.items() if isinstance(s, FixedLenFeature)}
    for b, rec in enumerate(serialized):
        feats, flists = parse_single(rec, read_feature_lists)
        for k, (vals, offs) in ragged.items():
            got: list = []
            if k in feats:
                got = list(feats[k][1])
            elif read_feature_lists and k in flists:
                for _, step in flists[k]:
                    got.extend(step)
            vals.extend(got)
            offs[b + 1] = len(vals)
        for k, arr in dense.items():
            if k in feats and len(feats[k][1]):
                v = np.asarray(feats[k][1], np.float32)
                if v.size != arr[b].size:
                    raise ValueError(f"Key: {k}. Can't parse serialized Example: expected {arr[b].size} values, got {v.size}")
                arr[b] = v.reshape(arr[b].shape)
    out: Dict[str, object] = dict(dense)
    out.update(ragged)
    return out


# ------------------------------------------------------------------ encode (used to write synthetic fixtures)
def _enc_feature(kind: str, values) -> bytes:
    if kind == "bytes":
        return _ld(1, b"".join(_ld(1, v) for v in values))
    if kind == "float":
        return _ld(2, _ld(1, struct.pack(f"<{len(values)}f", *values)) if len(values) else b"")
    if kind == "int64":
        return _ld(3, _ld(1, b"".join(_enc_varint(int(v)) for v in values)) if len(values) else b"")
    raise ValueError(kind)


def _enc_features(features: Dict[str, Tuple[str, Value]]) -> bytes:
    return b"".join(_ld(1, _ld(1, k.encode()) + _ld(2, _enc_feature(kind, vals))) for k, (kind, vals) in features.items())


def encode_example(features: Dict[str, Tuple[str, Value]]) -> bytes:
    return _ld(1, _enc_features(features))


def encode_sequence_example(context: Dict[str, Tuple[str, Value]],
                            feature_lists: Dict[str, List[Tuple[str, Value]]]) -> bytes:
    """Same message shape as the reference ETL writes (DataGenerator.py:429-442)."""
    fl = b"".join(_ld(1, _ld(1, k.encode()) + _ld(2, b"".join(_ld(1, _enc_feature(kind, vals)) for kind, vals in steps)))
                  for k, steps in feature_lists.items())
    return _ld(1, _enc_features(context)) + _ld(2, fl)
