/* libctr_feed.so -- native host-side feeder for libctr_b200 (the step immediately BEFORE the hot path).
 *
 * Replaces, on the CPU that feeds the GPU, what the reference's tf.data pipeline does inside TensorFlow:
 *   tf.data.TFRecordDataset                    (algorithm/utils.py:18,41)        -> ctr_feed_tfrecord_index
 *   tf.parse_example(batch, parse spec)        (DCN/dcn.py:116-131)              -> ctr_feed_parse_examples
 *   categorical_column_with_vocabulary_file    (DeepFM/deepfm.py:56-64)          -> ctr_feed_vocab_*  (OOV / '' -> -1)
 * Wire formats: SURVEY.md Appendix A.1 (TFRecord framing, masked CRC-32C), A.2 (Example / SequenceExample protos; a
 * SequenceExample parsed as an Example yields its context features and drops feature_lists -- parity note 8), A.3 (parse
 * spec: VarLen string features, FixedLen float features with a default), A.4 (vocabulary: id = 0-based line, OOV = -1).
 *
 * Plain C ABI, host pointers only, no CUDA dependency; every function is thread-safe w.r.t. distinct outputs (a vocabulary
 * is immutable after creation).  Errors: negative return + ctr_feed_last_error() (thread-local text).
 * The Python binding is recalgorithm_b200/io/native.py; recalgorithm_b200/io/{tfrecord,example,vocab}.py are the readable
 * pure-Python twins the tests compare it with.
 */
#ifndef CTR_FEED_H_
#define CTR_FEED_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CTR_FEED_OK 0
#define CTR_FEED_ERR_ARG (-1)        /* null pointer / bad size */
#define CTR_FEED_ERR_IO (-2)         /* cannot read a file */
#define CTR_FEED_ERR_TRUNCATED (-4)  /* TFRecord: record cut short */
#define CTR_FEED_ERR_CRC (-5)        /* TFRecord: masked CRC-32C mismatch */
#define CTR_FEED_ERR_PROTO (-6)      /* malformed protobuf wire data, or a feature of the wrong kind / size for its spec */
#define CTR_FEED_ERR_CAPACITY (-7)   /* a ragged output buffer is too small: `needed` of the offending key(s) is set */
#define CTR_FEED_ERR_NOMEM (-8)      /* an allocation failed inside the library (no C++ exception ever crosses this ABI) */

const char* ctr_feed_last_error(void);
int ctr_feed_version(void);

/* CRC-32C (Castagnoli, reflected 0x82F63B78) and TFRecord's masked form ((crc >> 15 | crc << 17) + 0xa282ead8). */
uint32_t ctr_feed_crc32c(const uint8_t* data, uint64_t n);
uint32_t ctr_feed_masked_crc32c(const uint8_t* data, uint64_t n);

/* Index the records of a TFRecord byte buffer: offsets[i] / lengths[i] locate payload i inside buf.  Returns the number of
 * records found (<= max_records; pass max_records = 0 and null arrays to count only) or a negative error;
 * *consumed (nullable) = bytes of complete records read. */
int64_t ctr_feed_tfrecord_index(const uint8_t* buf, uint64_t n, int verify_crc, uint64_t* offsets, uint64_t* lengths,
                                int64_t max_records, uint64_t* consumed);
/* The same scan resumed at byte `start` of buf (offsets stay relative to buf, error texts name absolute byte positions): lets a
 * reader index a file chunk by chunk while the chunks behind are already being parsed.  allow_partial_tail = 1: an incomplete
 * last record is not an error -- the scan stops in front of it (buf[0, n) is a prefix of a file that is still being read).
 * On an error return *consumed is the byte at which the damaged record starts (everything in front of it is intact: scanning
 * buf[0, *consumed) again succeeds and yields the records before the damage). */
int64_t ctr_feed_tfrecord_index_from(const uint8_t* buf, uint64_t n, uint64_t start, int verify_crc, int allow_partial_tail,
                                     uint64_t* offsets, uint64_t* lengths, int64_t max_records, uint64_t* consumed);
/* verify_crc: 0 = none, 1 = length and payload CRCs inside the (sequential) scan, 2 = length CRCs only -- the scan needs those
 * to trust the lengths; the payload CRCs are then checked by ctr_feed_tfrecord_verify on `num_threads` threads (0 = all cores):
 * CTR_FEED_OK, or CTR_FEED_ERR_CRC naming the first corrupted record. */
int ctr_feed_tfrecord_verify(const uint8_t* buf, uint64_t n, const uint64_t* offsets, const uint64_t* lengths, int64_t count,
                             int num_threads);

/* dataset.shuffle(buffer_size) (algorithm/utils.py:20): the order in which a shuffle buffer emits n elements -- a buffer of the
 * next `buffer_size` inputs, one of them drawn uniformly at each step (draws[i] in [0,1) picks slot floor(draws[i] * filled))
 * and replaced by the next input, or by the buffer's last element once the input is exhausted.  The caller supplies the n
 * draws, so the order is a pure function of them (input_fn.shuffle_order feeds numpy's seeded generator).
 * buffer_size <= 1: identity; buffer_size >= n: still the buffer walk (a uniform permutation).  out: (n,) int64. */
int ctr_feed_shuffle_order(int64_t n, int64_t buffer_size, const double* draws, int64_t* out);
/* The same walk, resumable, for an input whose length is not known yet (a file that is still being indexed): emit as many
 * positions as the input seen so far allows.  n_available = inputs known to exist, input_done = 1 once that is all of them.
 * Nothing is emitted before the buffer is full (or the input done); after that an emission stops short only when it would
 * have to know whether input `next` exists and cannot.  One draw per emission, consumed in order: the result is the order
 * ctr_feed_shuffle_order gives for the same draws, however the calls are cut.  Returns the number emitted (<= max_out). */
void* ctr_feed_shuffle_create(int64_t buffer_size);
void ctr_feed_shuffle_destroy(void* shuffle);
int64_t ctr_feed_shuffle_emit(void* shuffle, int64_t n_available, int input_done, const double* draws, int64_t max_out,
                              int64_t* out);

/* Vocabulary: token i = blob[offsets[i], offsets[i+1]); id = index of the FIRST occurrence of a token. */
void* ctr_feed_vocab_create(const uint8_t* blob, const uint64_t* offsets, int64_t n_tokens);
void* ctr_feed_vocab_load(const char* path);                 /* one token per line ('\n' or '\r\n'), like the reference's files */
int64_t ctr_feed_vocab_size(const void* vocab);              /* number of lines (= vocabulary_size) */
void ctr_feed_vocab_destroy(void* vocab);
int ctr_feed_vocab_lookup(const void* vocab, const uint8_t* blob, const uint64_t* offsets, int64_t n_keys, int64_t* ids_out);

/* One categorical (VarLenFeature(string)) key of the parse spec: values are mapped through `vocab` to int64 ids and written
 * ragged: ids[row_offsets[b] .. row_offsets[b+1]) are the ids of record b (missing key -> empty row).  `capacity` = size of
 * ids; on return `needed` = total number of values (also when the call fails with CTR_FEED_ERR_CAPACITY). */
typedef struct {
  const char* key;
  const void* vocab;
  int64_t* ids;
  int64_t capacity;
  int64_t* row_offsets; /* (B + 1) */
  int64_t needed;
} ctr_feed_cat_t;

/* One dense (FixedLenFeature((width,), float32, default)) key: out is (B, width); a missing / empty feature gives the default,
 * any other length is an error (tf.parse_example raises). */
typedef struct {
  const char* key;
  int64_t width;
  float default_value;
  float* out;
} ctr_feed_dense_t;

/* tf.parse_example over a batch of B serialized Example / SequenceExample protos (record b = buf[offsets[b], +lengths[b])).
 * read_feature_lists = 0 reproduces the reference (feature_lists ignored); 1 also looks categorical keys up in
 * SequenceExample.feature_lists (values of all steps concatenated) when the context does not hold them.
 * num_threads <= 0: hardware concurrency (capped at 32 and at B / 64). */
int ctr_feed_parse_examples(const uint8_t* buf, const uint64_t* offsets, const uint64_t* lengths, int64_t B, ctr_feed_cat_t* cats,
                            int64_t n_cat, ctr_feed_dense_t* dense, int64_t n_dense, int read_feature_lists, int num_threads);

#ifdef __cplusplus
}
#endif
#endif /* CTR_FEED_H_ */
