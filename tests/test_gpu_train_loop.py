"""GPU: the whole drop-in chain on a synthetic wechat-schema TFRecord -- native feeder -> train_input_fn -> fused lookup +
FM2 + first-order term -> sigmoid cross-entropy -> backward (IndexedSlices) -> Adam on the tables (TF-dense and lazy) -- learns
a label that is a function of the ids (loss falls well below the constant-predictor entropy)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
CATS = ["userid", "feedid", "device", "authorid", "bgm_song_id", "bgm_singer_id"]


@pytest.mark.parametrize("lazy", [False, True])
def test_deepfm_fm_part_learns_from_tfrecord(tmp_path, lazy):
    from recalgorithm_b200 import autograd, feature_column as fc, input_fn as I, io as cio, optim
    rng = np.random.default_rng(17)
    n, K, B = 4096, 8, 512
    recs = []
    for i in range(n):
        u, f = int(rng.integers(0, 40)), int(rng.integers(0, 60))
        y = float((u % 2) ^ (f % 3 == 0))                                     # label = function of (userid, feedid): needs the pair term
        ctx = {"userid": ("bytes", [f"userid_{u}".encode()]), "feedid": ("bytes", [f"feedid_{f}".encode()]),
               "device": ("bytes", [f"device_{rng.integers(1, 3)}".encode()]), "authorid": ("bytes", [f"authorid_{rng.integers(0, 30)}".encode()]),
               "bgm_song_id": ("bytes", [b"" if rng.random() < 0.3 else f"bgm_song_id_{rng.integers(0, 20)}".encode()]),
               "bgm_singer_id": ("bytes", [b"" if rng.random() < 0.3 else f"bgm_singer_id_{rng.integers(0, 20)}".encode()]),
               "read_comment": ("float", [y])}
        recs.append(cio.encode_example(ctx))
    path = str(tmp_path / "train.tfrecord")
    cio.write_records(path, recs)
    cat_cols = [fc.categorical_column_with_vocabulary_file(k, cio.VocabularyFile([f"{k}_{i}".encode() for i in range(64)])) for k in CATS]
    label = fc.numeric_column("read_comment", default_value=0.0)
    parser = I.make_example_parser(cat_cols + [label], label_keys=["read_comment"])

    torch.manual_seed(0)
    tables = autograd.EmbeddingTables([64] * len(CATS), K, device="cuda")
    first = torch.zeros((64 * len(CATS), 1), device="cuda", requires_grad=True)        # first-order weights (one per id), deepfm.py:180-181
    bias = torch.zeros((1,), device="cuda", requires_grad=True)
    off = tables.field_row_offset
    opt_t = optim.TableAdam(tables, lr=0.05, lazy=lazy)
    opt_d = torch.optim.Adam([first, bias], lr=0.05)
    losses = []
    for features, labels in I.train_input_fn(path, parser, batch_size=B, num_epochs=12, shuffle_buffer_size=2048, seed=3):
        ids = torch.from_numpy(fc.single_valued_ids(features, cat_cols)).cuda()
        y = torch.from_numpy(labels["read_comment"]).cuda()
        _, fm2 = autograd.lookup_fm2(tables, ids)
        logit = fm2 + _first_order(first, off, ids) + bias
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logit, y)
        opt_d.zero_grad()
        loss.backward()
        opt_d.step()
        opt_t.step()
        losses.append(float(loss.detach()))
    assert len(losses) == 12 * n // B
    base = float(np.mean(losses[:2]))
    assert np.mean(losses[-8:]) < 0.35 * base, (base, losses[-8:])


def _first_order(w, off, ids):
    """sum of the per-id scalar weights (differentiable in torch for this test; the fused forward is ctr_first_order_fwd)."""
    valid = ids >= 0
    rows = (ids + off[:-1][None, :]).clamp_min(0)
    return (w[rows.reshape(-1), 0].reshape(ids.shape) * valid).sum(1, keepdim=True)


def test_device_prefetcher_delivers_the_same_batches(tmp_path):
    from recalgorithm_b200 import feature_column as fc, input_fn as I, io as cio
    from test_io import wechat_record
    rng = np.random.default_rng(2)
    path = str(tmp_path / "d.tfrecord")
    cio.write_records(path, [wechat_record(rng, i)[0] for i in range(700)])
    cat_cols = [fc.categorical_column_with_vocabulary_file(k, cio.VocabularyFile([f"{k}_{i}".encode() for i in range(64)])) for k in CATS]
    cols = cat_cols + [fc.numeric_column("videoplayseconds"), fc.numeric_column("read_comment")]
    parser = I.make_example_parser(cols, label_keys=["read_comment"])
    host = list(I.eval_input_fn(path, parser, batch_size=128))
    dev_batches = list(I.DevicePrefetcher(I.eval_input_fn(path, parser, batch_size=128), cat_cols, dense_keys=["videoplayseconds"],
                                          label_keys=["read_comment"]))
    assert len(dev_batches) == len(host) == 6
    for (ids, dense, labels), (features, lab) in zip(dev_batches, host):
        assert ids.is_cuda and ids.dtype == torch.int64
        assert np.array_equal(ids.cpu().numpy(), fc.single_valued_ids(features, cat_cols))
        assert np.array_equal(dense["videoplayseconds"].cpu().numpy(), features["videoplayseconds"])
        assert np.array_equal(labels["read_comment"].cpu().numpy(), lab["read_comment"])
