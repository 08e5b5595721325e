"""TFRecord framing (what tf.data.TFRecordDataset reads -- utils.py:18,41 -- and tf.io.TFRecordWriter writes --
dataset/wechat_algo_data1/DataGenerator.py:403):

    uint64_le length | uint32_le masked_crc32c(length) | data[length] | uint32_le masked_crc32c(data)
    masked(crc) = ((crc >> 15 | crc << 17) + 0xa282ead8) mod 2**32,   crc = CRC-32C (Castagnoli)
"""
from __future__ import annotations

import struct
from typing import Iterable, Iterator, List

_POLY = 0x82F63B78
_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ _POLY if _c & 1 else _c >> 1
    _TABLE.append(_c)


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    tab = _TABLE
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def read_records(path: str, verify: bool = True) -> Iterator[bytes]:
    """Yield the serialized records of one TFRecord file; raises on a truncated record or a CRC mismatch."""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise IOError(f"{path}: truncated record header")
            (length,), (len_crc,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
            if verify and masked_crc32c(head[:8]) != len_crc:
                raise IOError(f"{path}: corrupted record length (crc mismatch)")
            data = f.read(length)
            tail = f.read(4)
            if len(data) < length or len(tail) < 4:
                raise IOError(f"{path}: truncated record")
            if verify and masked_crc32c(data) != struct.unpack("<I", tail)[0]:
                raise IOError(f"{path}: corrupted record data (crc mismatch)")
            yield data


def write_records(path: str, records: Iterable[bytes]) -> int:
    n = 0
    with open(path, "wb") as f:
        for r in records:
            head = struct.pack("<Q", len(r))
            f.write(head + struct.pack("<I", masked_crc32c(head)) + r + struct.pack("<I", masked_crc32c(r)))
            n += 1
    return n


def batches(records: Iterable[bytes], batch_size: int) -> Iterator[List[bytes]]:
    """dataset.batch(batch_size) BEFORE parsing, like the reference's input_fns (utils.py:22-23)."""
    buf: List[bytes] = []
    for r in records:
        buf.append(r)
        if len(buf) == batch_size:
            yield buf
            buf = []
    if buf:
        yield buf
