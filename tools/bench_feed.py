"""Throughput of the host-side feeder (libctr_feed.so) vs the pure-Python readers on a synthetic wechat_algo_data1 TFRecord
(SequenceExamples with the reference's field names, DataGenerator.py:400-443).  CPU only.

    python tools/bench_feed.py [--records 100000] [--threads 0]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from recalgorithm_b200 import io as cio  # noqa: E402
from recalgorithm_b200.io import native  # noqa: E402
from test_io import wechat_record  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=100_000)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    base = [wechat_record(rng, i)[0] for i in range(2000)]                   # encoded in Python once, then repeated
    recs = [base[i % len(base)] for i in range(args.records)]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "train.tfrecord")
        cio.write_records(path, recs) if args.records <= 20_000 else _fast_write(path, recs)
        size = os.path.getsize(path)
        keys = ["userid", "feedid", "device", "authorid", "bgm_song_id", "bgm_singer_id"]
        toks = {k: [f"{k}_{i}".encode() for i in range(100)] for k in keys}
        t0 = time.perf_counter()
        buf, off, ln = native.read_tfrecord_file(path)
        t_index = time.perf_counter() - t0
        open_rates = {}
        for label, kw in (("read_full_verify", dict(verify=True)), ("read_length_crcs_only", dict(verify="headers")),
                          ("mmap_length_crcs_only", dict(verify="headers", mmap=True))):
            best = float("inf")
            for _ in range(3):
                t1 = time.perf_counter()
                native.read_tfrecord_file(path, **kw)
                best = min(best, time.perf_counter() - t1)
            open_rates[label] = size / 1e6 / best
        vocabs = {k: native.Vocabulary(toks[k]) for k in keys}
        dense = {"videoplayseconds": (1, 0.0), "read_comment": (1, 0.0)}
        res = {}
        for nt in sorted({1, args.threads or (os.cpu_count() or 1)}):
            best = float("inf")
            for _ in range(5):                                               # best of 5: the first call pays the page faults of the outputs
                t0 = time.perf_counter()
                out = native.parse_examples(buf, off, ln, vocabs, dense, num_threads=nt)
                best = min(best, time.perf_counter() - t0)
            res[nt] = best
        n_py = min(args.records, 3000)
        spec = {k: cio.VarLenFeature() for k in keys}
        spec.update({"videoplayseconds": cio.FixedLenFeature(), "read_comment": cio.FixedLenFeature()})
        t0 = time.perf_counter()
        batch = []
        for i, r in enumerate(cio.read_records(path)):
            batch.append(r)
            if i + 1 == n_py:
                break
        parsed = cio.parse_example(batch, spec)
        pyv = {k: cio.VocabularyFile(toks[k]) for k in keys}
        for k in keys:
            pyv[k].lookup(parsed[k][0])
        t_py = time.perf_counter() - t0
        assert np.array_equal(out["userid"][0][:n_py], pyv["userid"].lookup(parsed["userid"][0]))
        # the whole host pipeline as a model sees it (input_fn: open + shuffle + repeat + batch + parse + prefetch), one epoch
        from recalgorithm_b200 import feature_column as fc, input_fn as I
        cols = [fc.categorical_column_with_vocabulary_file(k, cio.VocabularyFile(toks[k])) for k in keys]
        parser = I.make_example_parser(cols + [fc.numeric_column("read_comment")], label_keys=["read_comment"])
        pipe = {}
        for label, make in (("eval_input_fn", lambda mm: I.eval_input_fn(path, parser, 65536, mmap=mm)),
                            ("train_input_fn_shuffle10000", lambda mm: I.train_input_fn(path, parser, 65536, 1, 10000, seed=0, mmap=mm))):
            for mm in (False, True):
                best, first = float("inf"), float("inf")
                for _ in range(5):
                    t1 = time.perf_counter()
                    n, t_first = 0, None
                    for f, _l in make(mm):
                        t_first = t_first if t_first is not None else time.perf_counter() - t1
                        n += len(f["userid"][1]) - 1
                    best, first = min(best, time.perf_counter() - t1), min(first, t_first)
                assert n == args.records
                pipe[label + ("_mmap" if mm else "_read")] = {"samples_per_s": args.records / best, "first_batch_s": first}
        print(json.dumps({"records": args.records, "file_MB": size / 1e6, "cpu_count": os.cpu_count(),
                          "input_fn": pipe,  # batch 65536, one epoch from open(): the file is read / mapped and indexed in the background
                          "native_read_index_crc_MBps": size / 1e6 / t_index,
                          "native_open_MBps": open_rates,      # input_fn opens with length CRCs only; payload CRCs run per batch in the prefetch thread
                          "native_parse_examples_per_s": {str(nt): args.records / t for nt, t in res.items()},
                          "python_read_parse_lookup_examples_per_s": n_py / t_py,
                          "what": "6 categorical keys -> vocabulary ids + 2 dense floats per record"}))


def _fast_write(path, recs):
    """TFRecord framing with the native CRC (the pure-Python CRC would dominate the set-up time)."""
    import struct
    with open(path, "wb") as f:
        cache = {}
        for r in recs:
            if r not in cache:
                head = struct.pack("<Q", len(r))
                cache[r] = head + struct.pack("<I", native.masked_crc32c(head)) + r + struct.pack("<I", native.masked_crc32c(r))
            f.write(cache[r])


if __name__ == "__main__":
    main()
