/* Plain-C caller of libctr_feed.so: no Python in the process -- include/ctr_feed.h is the whole interface of the host-side
 * feeder (what tf.data.TFRecordDataset + tf.parse_example + categorical_column_with_vocabulary_file do for the reference,
 * algorithm/utils.py:18-24, DCN/dcn.py:116-131, DeepFM/deepfm.py:56-64).
 *
 *   feed_demo <file.tfrecord> <vocab dir> <batch size> <key> [<key> ...]
 *
 * Maps the file, indexes its records (length CRCs in the scan, payload CRCs afterwards on all cores), then parses it batch by
 * batch: every <key> is a single- or multi-valued categorical feature looked up in <vocab dir>/<key>.txt (id = line number,
 * OOV / '' -> -1) plus the float feature `read_comment` (default 0).  Prints, per key, the number of values, how many were
 * out of vocabulary and a checksum sum_i (i+1)*(id_i+2) over the whole file, then the sum of the labels;
 * tests/test_feed_native.py compares the lines with the Python twin.
 *
 * Build: gcc -O2 -I include examples/feed_demo.c -o feed_demo -L recalgorithm_b200/csrc_feed -lctr_feed \
 *            -Wl,-rpath,recalgorithm_b200/csrc_feed
 */
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "ctr_feed.h"

#define FAIL(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } while (0)

int main(int argc, char** argv) {
  if (argc < 5) FAIL("usage: %s file.tfrecord vocab_dir batch key [key ...]", argv[0]);
  const int64_t batch = atoll(argv[3]);
  const int n_cat = argc - 4;
  if (batch < 1) FAIL("batch must be >= 1");

  int fd = open(argv[1], O_RDONLY);
  struct stat st;
  if (fd < 0 || fstat(fd, &st)) FAIL("cannot open %s", argv[1]);
  const uint64_t size = (uint64_t)st.st_size;
  const uint8_t* buf = size ? mmap(NULL, size, PROT_READ, MAP_PRIVATE, fd, 0) : (const uint8_t*)"";
  if (buf == MAP_FAILED) FAIL("mmap failed");

  /* index: no record is shorter than its 16 bytes of framing */
  const int64_t bound = (int64_t)(size / 16 + 1);
  uint64_t* off = malloc(sizeof(uint64_t) * bound);
  uint64_t* len = malloc(sizeof(uint64_t) * bound);
  uint64_t consumed = 0;
  const int64_t n = ctr_feed_tfrecord_index(buf, size, /*verify_crc=*/2, off, len, bound, &consumed);
  if (n < 0) FAIL("index: %s", ctr_feed_last_error());
  if (ctr_feed_tfrecord_verify(buf, size, off, len, n, /*num_threads=*/0) != CTR_FEED_OK) FAIL("verify: %s", ctr_feed_last_error());

  void** vocab = calloc(n_cat, sizeof(void*));
  ctr_feed_cat_t* cats = calloc(n_cat, sizeof(ctr_feed_cat_t));
  uint64_t* checksum = calloc(n_cat, sizeof(uint64_t));
  int64_t* values = calloc(n_cat, sizeof(int64_t));
  int64_t* oov = calloc(n_cat, sizeof(int64_t));
  int64_t* cap = calloc(n_cat, sizeof(int64_t));
  for (int k = 0; k < n_cat; ++k) {
    char path[4096];
    snprintf(path, sizeof(path), "%s/%s.txt", argv[2], argv[4 + k]);
    vocab[k] = ctr_feed_vocab_load(path);
    if (!vocab[k]) FAIL("vocabulary %s: %s", path, ctr_feed_last_error());
    cap[k] = 2 * batch + 16;
    cats[k].key = argv[4 + k];
    cats[k].vocab = vocab[k];
    cats[k].ids = malloc(sizeof(int64_t) * cap[k]);
    cats[k].capacity = cap[k];
    cats[k].row_offsets = malloc(sizeof(int64_t) * (batch + 1));
  }
  float* label = malloc(sizeof(float) * batch);
  ctr_feed_dense_t dense = {"read_comment", 1, 0.0f, label};
  double label_sum = 0.0;

  for (int64_t b0 = 0; b0 < n; b0 += batch) {
    const int64_t B = n - b0 < batch ? n - b0 : batch;
    int rc = ctr_feed_parse_examples(buf, off + b0, len + b0, B, cats, n_cat, &dense, 1, /*read_feature_lists=*/0, /*threads=*/0);
    if (rc == CTR_FEED_ERR_CAPACITY) {               /* a ragged buffer was too small: `needed` says how large it has to be */
      for (int k = 0; k < n_cat; ++k)
        if (cats[k].needed > cap[k]) {
          cap[k] = cats[k].needed;
          cats[k].ids = realloc(cats[k].ids, sizeof(int64_t) * cap[k]);
          cats[k].capacity = cap[k];
        }
      rc = ctr_feed_parse_examples(buf, off + b0, len + b0, B, cats, n_cat, &dense, 1, 0, 0);
    }
    if (rc != CTR_FEED_OK) FAIL("parse (records %lld..): %s", (long long)b0, ctr_feed_last_error());
    for (int k = 0; k < n_cat; ++k)
      for (int64_t i = 0; i < cats[k].needed; ++i) {
        const int64_t id = cats[k].ids[i];
        checksum[k] += (uint64_t)(values[k] + 1) * (uint64_t)(id + 2);
        ++values[k];
        oov[k] += id < 0;
      }
    for (int64_t b = 0; b < B; ++b) label_sum += label[b];
  }
  printf("records %lld vocabulary sizes", (long long)n);
  for (int k = 0; k < n_cat; ++k) printf(" %lld", (long long)ctr_feed_vocab_size(vocab[k]));
  printf("\n");
  for (int k = 0; k < n_cat; ++k)
    printf("%s values %lld oov %lld checksum %llu\n", argv[4 + k], (long long)values[k], (long long)oov[k], (unsigned long long)checksum[k]);
  printf("read_comment sum %.1f\n", label_sum);
  for (int k = 0; k < n_cat; ++k) ctr_feed_vocab_destroy(vocab[k]);
  return 0;
}
