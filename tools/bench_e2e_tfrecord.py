"""End to end FROM TFRECORD BYTES: file -> native feeder (index + CRC + parse + vocabulary) -> pinned staging -> H2D ->
fused lookup + FM2 forward -> loss -> backward (IndexedSlices) on the wechat_algo_data1 schema (6 categorical fields, BASELINE
config 1's columns) at a large batch.  Prints one JSON line with samples/s of the whole chain and of its stages.

    python tools/bench_e2e_tfrecord.py [--records 2000000] [--batch 65536]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from recalgorithm_b200 import autograd, feature_column as fc, input_fn as I, io as cio  # noqa: E402
from bench_feed import _fast_write  # noqa: E402
from test_io import wechat_record  # noqa: E402

CATS = ["userid", "feedid", "device", "authorid", "bgm_song_id", "bgm_singer_id"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=2_000_000)
    ap.add_argument("--batch", type=int, default=65536)
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    base = [wechat_record(rng, i)[0] for i in range(4000)]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "train.tfrecord")
        _fast_write(path, [base[i % len(base)] for i in range(args.records)])
        cat_cols = [fc.categorical_column_with_vocabulary_file(k, cio.VocabularyFile([f"{k}_{i}".encode() for i in range(100)])) for k in CATS]
        parser = I.make_example_parser(cat_cols + [fc.numeric_column("read_comment")], label_keys=["read_comment"])
        tables = autograd.EmbeddingTables([100] * len(CATS), 8, device="cuda")
        t0 = time.perf_counter()
        n_host = sum(len(f["userid"][1]) - 1 for f, _ in I.eval_input_fn(path, parser, args.batch))
        t_host = time.perf_counter() - t0
        assert n_host == args.records

        def epoch():
            n = 0
            for ids, _, labels in I.DevicePrefetcher(I.eval_input_fn(path, parser, args.batch), cat_cols, label_keys=["read_comment"]):
                tables.zero_grad()
                tile, fm2 = autograd.lookup_fm2(tables, ids)
                loss = torch.nn.functional.binary_cross_entropy_with_logits(fm2 + tile.sum((1, 2)).unsqueeze(1) * 0.01, labels["read_comment"])
                loss.backward()
                n += ids.shape[0]
            torch.cuda.synchronize()
            return n
        epoch()
        t0 = time.perf_counter()
        n = epoch()
        t_all = time.perf_counter() - t0
        print(json.dumps({"records": args.records, "batch": args.batch, "file_MB": os.path.getsize(path) / 1e6,
                          "host_only_read_parse_samples_per_s": n_host / t_host, "end_to_end_samples_per_s": n / t_all,
                          "cpu_count": os.cpu_count(),
                          "what": "TFRecord bytes -> libctr_feed (CRC, parse, vocab) -> pinned -> H2D -> lookup+FM2 fwd -> BCE -> bwd; F=6, D=8"}))


if __name__ == "__main__":
    main()
