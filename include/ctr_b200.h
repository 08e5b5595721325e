/* ctr_b200.h -- C ABI of the Blackwell-native CTR feature-interaction engine (libctr_b200.so).
 *
 * Drop-in boundary for the hot path of tangxyw/RecAlgorithm (reference paths below are relative to
 * /root/reference/algorithm).  The reference has no FFI of its own -- its boundary is Python
 * callables invoked while a TF1 graph is built -- so each entry point here
 * names the reference callable whose forward / autodiff-backward it replaces.  The Python side
 * (recalgorithm_b200/layers.py) re-exposes them under the reference's own signatures.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, int64_t sizes, `void* stream` is a cudaStream_t (NULL = default stream);
 *   - every function returns 0 on success or a negative CTR_ERR_* code; ctr_last_error() gives the text
 *     (thread-local).  Argument errors are detected before any launch;
 *   - the CALLER owns every buffer; kernels never allocate, never synchronise, and are re-entrant
 *     across streams.  Where a workspace is needed there is a *_workspace_bytes() query;
 *   - all tensors are dense row-major fp32 unless stated; ids / offsets / lengths are int64;
 *   - there is no CPU fallback: on a machine without an sm_100 GPU every compute entry point fails
 *     with CTR_ERR_CUDA.
 */
#ifndef CTR_B200_H_
#define CTR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTR_OK 0
#define CTR_ERR_INVALID_ARG (-1)
#define CTR_ERR_UNSUPPORTED (-2)
#define CTR_ERR_CUDA (-3)

/* ---- library ---------------------------------------------------------------------------------- */
const char* ctr_last_error(void);
int ctr_version(void);                          /* ABI version, currently 1 */
int ctr_device_info(int* sm_count, int* cc_major, int* cc_minor);   /* current device */
int ctr_enable_peer_access(int peer_device);     /* map `peer_device`'s memory into the current device (NVLink P2P); idempotent */
/* Peer-mappable device buffers for the row-sharded path: a plain device allocation, its 64-byte CUDA-IPC handle, and the
 * import of a peer's handle into the current device's address space (peer access over NVLink is enabled on import). */
int ctr_peer_alloc(int64_t bytes, void** ptr);
int ctr_peer_free(void* ptr);
int ctr_ipc_export(void* ptr, unsigned char* handle64);
int ctr_ipc_import(const unsigned char* handle64, void** ptr);
int ctr_ipc_close(void* ptr);
/* The same role from the CUDA virtual-memory-management API (cuMemCreate / cuMemMap), shared as POSIX file descriptors: the
 * mapping the big table shards use -- a legacy-IPC mapping of a 32 GB shard collapses under random 128-byte peer reads, and
 * with VMM the size / address alignment chosen here (align = 0: the driver's recommended granularity) sets the page size the
 * peers' TLBs see.  ctr_vmm_alloc maps `bytes` (rounded up to `align`) on the current device and returns the fd to hand to
 * the peers (e.g. through pidfd_getfd); ctr_vmm_import maps a peer's fd into the current device (read/write over NVLink). */
int ctr_vmm_granularity(int64_t* minimum, int64_t* recommended);
int ctr_vmm_alloc(int64_t bytes, int64_t align, void** ptr, int* fd, int64_t* mapped_bytes);
int ctr_vmm_import(int fd, int64_t mapped_bytes, int64_t align, void** ptr);
int ctr_vmm_free(void* ptr);
int64_t ctr_kernel_launches(void);              /* kernels launched by this library so far (process-wide) */

/* ---- Row L + FM2: fused embedding lookup + DeepFM second-order term ------------------------------
 * Replaces   fc.input_layer(features, [embedding_column])   x F   (DeepFM/deepfm.py:187-190;
 *            xDeepFM/xdeepfm.py:158,167; DCN/dcn.py:153; FiBiNET/fibinet.py:162-163)
 *   and the inline FM second-order block (DeepFM/deepfm.py:192-200).
 *
 * table            (V_total, D): the F per-field tables stored back to back; field f owns rows
 *                  [field_row_offset[f], field_row_offset[f+1]).
 * field_row_offset device int64[F+1].
 * ids              device int64 (B, F), per-field local ids.  id < 0 (the vocabulary's OOV / '' value, -1)
 *                  or id >= the field's row count gives the ZERO vector (TF: pruned id -> empty bag -> zeros).
 * tile             out (B, F, D) contiguous, or NULL (lookup+FM2 only).
 * fm2              out (B), 0.5 * sum_d[(sum_f e)^2 - sum_f e^2], or NULL (lookup only).
 * D must be a multiple of 4 and <= 128 (128-bit row chunks); other widths go through ctr_bag_lookup_*.
 */
int ctr_embed_fm2_fwd(const float* table, const int64_t* field_row_offset, const int64_t* ids,
                      int64_t B, int64_t F, int64_t D, float* tile, float* fm2, void* stream);
/* The same with int32 ids (half the id bytes over PCIe / HBM); ids64_out (may be NULL) receives the widened (B,F) int64 copy
 * for IndexedSlices consumers (optimizers, ctr_embed_scatter_add). */
int ctr_embed_fm2_fwd_ids32(const float* table, const int64_t* field_row_offset, const int32_t* ids, int64_t B, int64_t F,
                            int64_t D, float* tile, float* fm2, int64_t* ids64_out, void* stream);

/* Sequence lookup: ids (B, T) all index ONE table -- rows [row_range[0], row_range[1]) of `table` (device int64[2]) -- e.g.
 * the padded behaviour history of tf.contrib.feature_column.sequence_input_layer over a shared embedding (DIN/din.py:209-214).
 * out (B, T, D); id < 0 or out of range -> zero row (the zero padding).  One warp per sample, like ctr_embed_fm2_fwd.  Its
 * gradient is the IndexedSlices (ids, d_out) as is -- no kernel. */
int ctr_embed_seq_fwd(const float* table, const int64_t* row_range, const int64_t* ids, int64_t B, int64_t T, int64_t D,
                      float* out, void* stream);

/* Backward of the pair above = the `values` of TF's IndexedSlices gradient of the gather
 * (indices are the caller's ids):   row_grads[b,f,:] = d_tile[b,f,:] + d_fm2[b] * (S[b,:] - e[b,f,:]),
 * S = sum_f e[b,f,:].   tile = the forward output; d_tile (B,F,D) and d_fm2 (B) may each be NULL (= 0).
 * Rows whose id was invalid receive a gradient too (it is simply never applied; see ctr_embed_scatter_add). */
int ctr_embed_fm2_bwd(const float* tile, const float* d_tile, const float* d_fm2,
                      int64_t B, int64_t F, int64_t D, float* row_grads, void* stream);

/* Lookup + FM2 with a fused dense(1) consumer of the flattened tile (the first use of the tile by DeepFM's deep part,
 * DeepFM/deepfm.py:203-212, reduced to one unit): lin[b] = sum_{f,d} tile[b,f,d]*wlin[f,d].  ids: int64 (ids_are_int32 = 0)
 * or int32 (= 1; ids64_out, may be NULL, receives the widened copy).  tile may be NULL only if no backward follows. */
int ctr_embed_fm2_lin_fwd(const float* table, const int64_t* field_row_offset, const void* ids, int ids_are_int32, int64_t B,
                          int64_t F, int64_t D, const float* wlin, float* tile, float* fm2, float* lin, int64_t* ids64_out,
                          void* stream);
/* Its backward: the upstream gradient of the tile is the rank-1 product d_lin[b]*wlin[f,d] and is never materialised:
 *   row_grads[b,f,:] = d_lin[b]*wlin[f,:] + d_fm2[b]*(S[b,:] - e[b,f,:]);   d_wlin[f,:] = sum_b d_lin[b]*e[b,f,:]
 * (d_wlin (F*D) is zeroed here; fp32 atomics across CTAs).  F*D <= 1536. */
int ctr_embed_fm2_lin_bwd(const float* tile, const float* wlin, const float* d_fm2, const float* d_lin, int64_t B, int64_t F,
                          int64_t D, float* row_grads, float* d_wlin, void* stream);

/* Mean sigmoid cross-entropy of logits = logit_a (+ logit_b, may be NULL) against labels, (B) each, and its gradient in one
 * launch: *loss = mean_b[max(x,0) - x*z + log1p(exp(-|x|))] (tf.nn.sigmoid_cross_entropy_with_logits + reduce_mean,
 * DeepFM/deepfm.py:214,235), d_logit[b] = (sigmoid(x) - z)/B (may be NULL).  fp32 atomics across CTAs for the mean. */
int ctr_sigmoid_ce(const float* logit_a, const float* logit_b, const float* labels, int64_t B, float* loss, float* d_logit,
                   void* stream);

/* Densify: grad_table[field_row_offset[f] + ids[b,f], :] += row_grads[b,f,:] for valid ids (duplicates
 * summed, like the optimizer's IndexedSlices de-duplication; TF-internal).  grad_table is
 * NOT zeroed here.  fp32 red.global.add -> summation order is not deterministic. */
int ctr_embed_scatter_add(float* grad_table, const int64_t* field_row_offset, const int64_t* ids,
                          const float* row_grads, int64_t B, int64_t F, int64_t D, void* stream);

/* ---- Row (e): row-sharded tables across the GPUs of one NVSwitch box -----------------------------------------
 * Global row gr = field_row_offset[f] + id is owned by rank gr % G and stored at local row gr / G (G a power of two <= 8).
 * The reference has no multi-device path; the semantics kept are the lookup's and the IndexedSlices gradient's.
 *
 * Forward: same contract as ctr_embed_fm2_fwd, but rows are PULLED from the owners' shards inside the gather kernel.
 * shard_ptrs: HOST array of G device pointers, entry r = rank r's shard (ceil(V_total/G), D) as mapped into THIS process
 * (CUDA IPC / peer mapping for r != own rank).  No collective is involved. */
int ctr_embed_fm2_fwd_sharded(const float* const* shard_ptrs, int64_t G, const int64_t* field_row_offset,
                              const int64_t* ids, int64_t B, int64_t F, int64_t D, float* tile, float* fm2, void* stream);
/* The same with int32 ids (half the id bytes over PCIe / HBM); ids64_out (may be NULL) receives the widened (B,F) int64 copy
 * that IndexedSlices consumers downstream expect. */
int ctr_embed_fm2_fwd_sharded_ids32(const float* const* shard_ptrs, int64_t G, const int64_t* field_row_offset,
                                    const int32_t* ids, int64_t B, int64_t F, int64_t D, float* tile, float* fm2,
                                    int64_t* ids64_out, void* stream);
/* Sharded form of ctr_embed_fm2_lin_fwd (fused dense(1) head; rows pulled from the owners' shards). */
int ctr_embed_fm2_lin_fwd_sharded(const float* const* shard_ptrs, int64_t G, const int64_t* field_row_offset, const void* ids,
                                  int ids_are_int32, int64_t B, int64_t F, int64_t D, const float* wlin, float* tile, float* fm2,
                                  float* lin, int64_t* ids64_out, void* stream);
/* Gradient exchange, step 1 (independent of the forward; one pass over the ids): assigns every valid (b,f) a slot in its
 * OWNER's receive queue and writes plan[b,f] = owner << 28 | slot (-1: invalid id, or dropped because the owner's slice is
 * full -> *overflow = 1).  The queue's local-row indices are written here, as contiguous runs per owner:
 * recv_rows / recv_counts: HOST arrays of G device pointers; entry d = owner d's (G_src, capacity) int64 row queue /
 * (G_src,) int64 count vector as mapped into this process; this rank writes slice [my_rank] and, when the kernel ends,
 * recv_counts[d][my_rank] = min(entries queued at d, capacity) (recv_counts or its entries may be NULL).
 * ids: int64 (ids_are_int32 = 0) or int32 (= 1).
 * counters: device int64[9] scratch, zeroed here (ends as entries per owner + a ticket); overflow: device int, zeroed here.
 * B*F < 2^28, capacity < 2^28. */
int ctr_sharded_plan(const int64_t* field_row_offset, const void* ids, int ids_are_int32, int64_t B, int64_t F, int64_t G, int64_t my_rank,
                     int64_t* const* recv_rows, int64_t* const* recv_counts, int64_t capacity, int64_t* counters,
                     int* overflow, int32_t* plan, void* stream);
/* Gradient exchange, step 2, fused into the lookup backward: same arithmetic as ctr_embed_fm2_bwd, but every planned row of
 * d_tile + d_fm2*(S - e) is stored straight into its owner's value queue with 128-bit peer stores (recv_vals: HOST array of
 * G device pointers, entry d = owner d's (G_src, capacity, D) fp32 queue; slice [my_rank] is written).  row_grads may be
 * NULL: the IndexedSlices values then never touch local HBM.  Follow with a stream sync and a cross-rank barrier before
 * any owner reads its queues. */
int ctr_embed_fm2_bwd_push(const float* tile, const float* d_tile, const float* d_fm2, const int32_t* plan, int64_t B,
                           int64_t F, int64_t D, int64_t G, int64_t my_rank, float* const* recv_vals, int64_t capacity,
                           float* row_grads, void* stream);
/* The same for the fused dense(1) head (ctr_embed_fm2_lin_fwd_sharded): d_tile is the rank-1 product d_lin[b]*wlin[f,d] and is
 * never materialised; d_wlin (F*D, zeroed here) = sum_b d_lin[b]*tile[b].  F*D <= 1536. */
int ctr_embed_fm2_lin_bwd_push(const float* tile, const float* wlin, const float* d_fm2, const float* d_lin, const int32_t* plan,
                               int64_t B, int64_t F, int64_t D, int64_t G, int64_t my_rank, float* const* recv_vals,
                               int64_t capacity, float* row_grads, float* d_wlin, void* stream);
/* The exchange alone, for row gradients (B,F,D) produced by any other backward. */
int ctr_sharded_grad_push(const float* row_grads, const int32_t* plan, int64_t B, int64_t F, int64_t D, int64_t G,
                          int64_t my_rank, float* const* recv_vals, int64_t capacity, void* stream);
/* Owner side / generic IndexedSlices consumer: dst[rows[i], :] += vals[i, :] for i < min(*count, max_n) (count may be
 * NULL = max_n); rows outside [0, V) are ignored.  fp32 vector red.global.add. */
int ctr_rows_scatter_add(float* dst, int64_t V, int64_t D, const int64_t* rows, const float* vals, const int64_t* count,
                         int64_t max_n, void* stream);

/* ---- Adam on the IndexedSlices gradient of a table (the step right after the hot path) ------------------
 * Reference: tf.train.AdamOptimizer(lr, .9, .999, 1e-8) (DeepFM/deepfm.py:246-250); its sparse apply sums duplicate
 * indices, decays m and v of the WHOLE table and updates every row; DIEN uses LazyAdam (DIEN/dien.py:328).
 * rows (n) must be UNIQUE with grads (n, D) already summed per row (ctr_rows_scatter_add into a compact buffer does that);
 * lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t) is computed by the caller.  count: device int64 (NULL = max_n).
 * state_stride (all ctr_adam_* entry points) = floats between consecutive rows of m (and of v): D for two separate (V, D)
 * tables, 2*D for ONE interleaved (V, 2, D) buffer with v = m + D -- a row's two moments then share a DRAM page, which is
 * what the random row updates are bound by (4 instead of 6 row activations per updated row).
 * ctr_adam_rows updates m, v, var of the listed rows (and sets their bit in touched_bitmap (ceil(V/32) uint32, zeroed by
 * the caller) when given) = LazyAdam; ctr_adam_dense_rest then applies the g = 0 update to every row whose bit is clear
 * = the reference's dense semantics. */
int ctr_adam_rows(float* var, float* m, float* v, int64_t state_stride, int64_t V, int64_t D, const int64_t* rows, const float* grads,
                  const int64_t* count, int64_t max_n, float lr_t, float beta1, float beta2, float eps,
                  uint32_t* touched_bitmap, void* stream);
int ctr_adam_dense_rest(float* var, float* m, float* v, int64_t state_stride, int64_t V, int64_t D, float lr_t, float beta1, float beta2, float eps,
                        const uint32_t* touched_bitmap, void* stream);

/* Fused IndexedSlices step (no sort, no host round trip): ids (B,F) per-field local ids (out-of-range = skipped, like the
 * lookup), row_grads (B,F,D) = the IndexedSlices values of ctr_embed_fm2_bwd -- CONSUMED (duplicates of a row are summed
 * into one of its entries).  slot_of_row: int32 per table row, all -1 on entry and again on exit (persistent scratch of
 * the optimizer).  Applies the LazyAdam update to every referenced row with the SUMMED gradient (TF sums duplicates
 * first); sets the rows' bits in touched_bitmap when given (then ctr_adam_dense_rest completes tf.train.AdamOptimizer's
 * dense semantics); adds the number of distinct rows to *n_unique when given.  dup_list: int32[B*F + 1] scratch or NULL --
 * when given, the claim pass lists the entries that met an already claimed row and the merge walks that list instead of
 * re-scanning every entry.  B*F < 2^30. */
int ctr_adam_indexed_slices(float* var, float* m, float* v, int64_t state_stride, const int64_t* field_row_offset, int64_t F, int64_t D,
                            const int64_t* ids, float* row_grads, int64_t B, int32_t* slot_of_row, int32_t* dup_list,
                            float lr_t, float beta1, float beta2, float eps, uint32_t* touched_bitmap, int64_t* n_unique,
                            void* stream);

/* Lookup backward FUSED with that step (the row update without writing row-grads to HBM): computes the
 * IndexedSlices values d_tile + d_fm2*(S - e) of ctr_embed_fm2_bwd in registers and applies the LazyAdam update to every row
 * referenced ONCE in the batch on the spot; rows referenced several times park their values in dup_grads ((B,F,D) scratch,
 * written sparsely) and their entry in dup_list (int32[B*F + 1] scratch, last element = count) and are finished with the
 * SUMMED gradient by two list-driven launches.  slot_of_row / touched_bitmap / n_unique as in ctr_adam_indexed_slices.
 * F*D <= 1536, B*F < 2^30.  Same results as ctr_embed_fm2_bwd followed by ctr_adam_indexed_slices. */
int ctr_embed_fm2_bwd_adam(const float* tile, const float* d_tile, const float* d_fm2, const int64_t* field_row_offset,
                           const int64_t* ids, int64_t B, int64_t F, int64_t D, float* var, float* m, float* v,
                           int64_t state_stride, int32_t* slot_of_row, float* dup_grads, int32_t* dup_list, float lr_t, float beta1, float beta2,
                           float eps, uint32_t* touched_bitmap, int64_t* n_unique, void* stream);

/* The same step for the OWNER side of a row-sharded table: entries are the receive queues filled by ctr_sharded_grad_push --
 * rows (nseg, cap) local row ids, vals (nseg, cap, D) (consumed), counts (nseg,) filled slots per segment; duplicates of a
 * row across and inside segments are summed before the update.  dup_list: int32[nseg*cap + 1] scratch or NULL.
 * nseg*cap < 2^30. */
int ctr_adam_rows_dedup(float* var, float* m, float* v, int64_t state_stride, int64_t V, int64_t D, const int64_t* rows, float* vals,
                        const int64_t* counts, int64_t nseg, int64_t cap, int32_t* slot_of_row, int32_t* dup_list, float lr_t,
                        float beta1, float beta2, float eps, uint32_t* touched_bitmap, int64_t* n_unique, void* stream);

/* DeepFM first-order ("wide") term as a D=1 lookup.  Replaces indicator_column multi-hot (B, sum V) @
 * dense(1) (DeepFM/deepfm.py:72-80,180-181): out[b] = bias + sum_f w[field_row_offset[f] + ids[b,f]]; invalid ids add 0.
 * w (V_total) is the dense(1) kernel; its gradient is the IndexedSlices (ids, d_out[b] broadcast over F) -- no kernel
 * needed -- and d_bias = sum_b d_out[b]. */
int ctr_first_order_fwd(const float* w, const int64_t* field_row_offset, const int64_t* ids, int64_t B, int64_t F,
                        float bias, float* out, void* stream);

/* ---- Row L (general): multi-valued bag lookup, combiner='mean' ----------------------------------
 * Replaces fc.input_layer over embedding_column(col, D, combiner='mean') on a VarLen feature
 * (DCN/dcn.py:98,103; xDeepFM/xdeepfm.py:103,108) and any single-valued column whose D is not a
 * multiple of 4 (DCN/dcn.py:99-102 use D = 2 and 4).
 * ids (nnz) flat values, offsets (B+1) CSR; ids < 0 or >= V are dropped; empty bag -> zeros.
 * out[b*out_stride + 0..D) is written (lets the caller place the field inside a (B, sum_d) row). */
int ctr_bag_lookup_fwd(const float* table, int64_t V, int64_t D, const int64_t* ids, const int64_t* offsets,
                       int64_t B, float* out, int64_t out_stride, void* stream);
/* row_grads (nnz, D): d_out[b]/count_b for valid ids, 0 for dropped ids. */
int ctr_bag_lookup_bwd(const float* d_out, int64_t out_stride, int64_t V, int64_t D, const int64_t* ids,
                       const int64_t* offsets, int64_t B, float* row_grads, void* stream);

/* ---- Row CROSS: DCN cross-layer stack --------------------------------------------------------------
 * Replaces the loop `for i: cross_vec = cross_layer(x0, cross_vec, i)` (DCN/dcn.py:157-160) over
 * cross_layer (DCN/cross_layer.py:21-24):  x_{l+1} = x0 * (x_l . w_l) + b_l + x_l,  x_0 = x0.
 * x0 (B,d); w, b (L,d) = the L (d,1) variables wl_i / bl_i stacked; out (B,d) = x_L.
 * xl_in: optional (B,d) start vector (NULL = x0); lets a single layer be called as cross_layer(x0, xl, i). */
int ctr_cross_fwd(const float* x0, const float* xl_in, const float* w, const float* b,
                  int64_t B, int64_t d, int64_t L, float* out, void* stream);
/* Gradients given g_out = dL/dx_L.  dx0 (B,d) receives the gradient through every use of x0
 * (and through x_0 when xl_in == NULL); dxl_in (B,d) is written only when xl_in != NULL.
 * dw, db (L,d) are overwritten (batch-reduced with fp32 atomics across CTAs). */
int ctr_cross_bwd(const float* x0, const float* xl_in, const float* w, const float* b, const float* g_out,
                  int64_t B, int64_t d, int64_t L, float* dx0, float* dxl_in, float* dw, float* db, void* stream);
/* Lookup FUSED with the cross stack: `net = fc.input_layer(features, cols)` (DCN/dcn.py:153) followed by the loop over
 * cross_layer (DCN/dcn.py:157-160) for uniform-width embedding columns.  table / field_row_offset / ids as in
 * ctr_embed_fm2_fwd (ids int64, or int32 when ids_are_int32 != 0; invalid ids give the zero vector); d = F*D.
 * x0 (B, F*D) receives the gathered input (the backward needs it; the lookup backward of a plain gather is the view
 * dx0 -> (B,F,D), no kernel), out (B, F*D) = x_L.  Needs D % 4 == 0, F*D <= 512, L <= 4 (CTR_ERR_UNSUPPORTED otherwise: call
 * ctr_embed_fm2_fwd + ctr_cross_fwd). */
int ctr_embed_cross_fwd(const float* table, const int64_t* field_row_offset, const void* ids, int ids_are_int32,
                        int64_t B, int64_t F, int64_t D, const float* w, const float* b, int64_t L, float* x0, float* out,
                        void* stream);

/* ---- Row CIN: xDeepFM compressed-interaction layer -------------------------------------------------
 * Replaces cin_layer(x0, xk, hk_1, index) (xDeepFM/cin_layer.py:17-30):
 *   out[b,n,d] = sum_{i,j} xk[b,i,d] * x0[b,j,d] * filter[i*m + j, n]
 * x0 (B,m,D); xk (B,hk,D); filter (hk*m, H) = the conv1d filter (1, hk*m, hk_1)[0]; out (B,H,D);
 * pooled (B,H) = sum_d out (the reduce_sum of xDeepFM/xdeepfm.py:173) or NULL.
 * precision: 0 = fp32-class (3xTF32 split on the tensor cores; default, meets 1e-5),
 *            1 = single-pass TF32 (about 1e-3; for speed comparisons only).
 * workspace: ctr_cin_fwd_workspace_bytes() bytes, 128-byte aligned (holds the filter re-ordered / tf32-split for
 *            TMA); may be NULL when that query returns 0 (shapes served by the CUDA-core path). */
int64_t ctr_cin_fwd_workspace_bytes(int64_t B, int64_t m, int64_t hk, int64_t D, int64_t H);
int ctr_cin_fwd(const float* x0, const float* xk, const float* filter, int64_t B, int64_t m, int64_t hk,
                int64_t D, int64_t H, float* out, float* pooled, int precision,
                void* workspace, int64_t workspace_bytes, void* stream);
/* Gradients of ctr_cin_fwd given g_out (B,H,D); dx0 (B,m,D), dxk (B,hk,D), dfilter (hk*m,H) are overwritten. */
int ctr_cin_bwd(const float* x0, const float* xk, const float* filter, const float* g_out,
                int64_t B, int64_t m, int64_t hk, int64_t D, int64_t H,
                float* dx0, float* dxk, float* dfilter,
                void* workspace, int64_t workspace_bytes, void* stream);
int64_t ctr_cin_bwd_workspace_bytes(int64_t B, int64_t m, int64_t hk, int64_t D, int64_t H);
/* Tuning hook (process-wide, returns the previous setting): 1 (default) = the dX kernel runs as CTA pairs (tcgen05
 * cta_group::2, M = 256, each SM feeds half of every filter stage); 0 = the single-CTA form with TMA multicast.  Same results;
 * used by tools/bench_layers.py for A/B timings. */
int ctr_cin_bwd_set_dx_pair(int on);

/* ---- Row DIN-ATT: DIN attention unit -----------------------------------------------------------------
 * Replaces din_attention(query, keys, keys_length, is_softmax) (DIN/din_attention.py:17-43).
 * query (B,H); keys (B,T,H); keys_length int64 (B); dense layers f1_att (4H->64, relu), f2_att (64->32,
 * relu), f3_att (32->1): w1 (4H,64) b1 (64) w2 (64,32) b2 (32) w3 (32) b3 (1); out (B,H).
 * att_w (B,T): the final per-position weights (saved for backward), or NULL.
 * sched_scratch: device int32[B + 64] or NULL.  When given, a one-CTA pass orders the samples by descending keys_length and
 * the warps take them from a shared counter (longest-first list scheduling: a warp's cost is its samples' lengths); NULL =
 * static round-robin.  Same results either way (weight gradients are fp32-atomic sums in both). */
int ctr_din_attention_fwd(const float* query, const float* keys, const int64_t* keys_length,
                          const float* w1, const float* b1, const float* w2, const float* b2,
                          const float* w3, const float* b3, int64_t B, int64_t T, int64_t H, int is_softmax,
                          float* out, float* att_w, int32_t* sched_scratch, void* stream);
/* d_params: one flat fp32 buffer laid out [w1 | b1 | w2 | b2 | w3 | b3] (4H*64+64+64*32+32+32+1), overwritten.
 * att_w: the (B,T) weights saved by the forward, or NULL (they are then recomputed). */
int ctr_din_attention_bwd(const float* query, const float* keys, const int64_t* keys_length,
                          const float* w1, const float* b1, const float* w2, const float* b2,
                          const float* w3, const float* b3, const float* g_out, const float* att_w,
                          int64_t B, int64_t T, int64_t H, int is_softmax,
                          float* d_query, float* d_keys, float* d_params, int32_t* sched_scratch, void* stream);

/* ---- Rows SENET / BILINEAR: FiBiNET ---------------------------------------------------------------------
 * senet(input, embedding_dim, reduction_ratio) (FiBiNET/senet.py:26-34): x (B,F,K); w1 (F,r); w2 (r,F). */
int ctr_senet_fwd(const float* x, const float* w1, const float* w2, int64_t B, int64_t F, int64_t K, int64_t r,
                  float* out, void* stream);
int ctr_senet_bwd(const float* x, const float* w1, const float* w2, const float* g_out,
                  int64_t B, int64_t F, int64_t K, int64_t r, float* dx, float* dw1, float* dw2, void* stream);
/* bilinear_interaction_layer(input, embedding_dim, type, name) (FiBiNET/bilinear_interaction_layer.py:21-40).
 * type: 0 'all' w (K,K); 1 'each' w (F-1,K,K); 2 'interaction' w (F(F-1)/2,K,K).
 * Pairs are itertools.combinations(range(F-1), 2) -- fields 0..F-2 only -- so out is (B, P, K) with
 * P = (F-1)(F-2)/2, exactly like the reference. */
int ctr_bilinear_fwd(const float* x, const float* w, int64_t B, int64_t F, int64_t K, int type,
                     float* out, void* stream);
int ctr_bilinear_bwd(const float* x, const float* w, const float* g_out, int64_t B, int64_t F, int64_t K, int type,
                     float* dx, float* dw, void* stream);
/* Tuning hook (process-wide, returns the previous mask; default 4).  Bit t (t = 0,1,2): type t runs the sample-batched
 * "tournament" kernels; otherwise 'all' / 'each' run the staged per-sample kernels and 'interaction' the round-1 kernels.
 * Bit 3: 'all' / 'each' use the round-1 CTA-per-sample kernels instead of the staged ones.  Bits 4..9: samples per tile of
 * the tournament kernels; bits 10..12: their weight columns per lane (1, 2, 4); 0 = chosen automatically.  The fast kernels
 * need K in {8,16,32} and 16-byte aligned arrays; other shapes always use the round-1 kernels.  Same results in every
 * setting; used by the parity tests (all forms) and tools/bench_layers.py for A/B timings. */
int ctr_bilinear_set_rr(int mask);

/* ---- siblings of FM2 ------------------------------------------------------------------------------------------
 * NFM bi-interaction pooling (NFM/nfm.py:155-168): the fused gather of ctr_embed_fm2_fwd with a (B, D) output
 * bi[b,:] = 0.5 * ((sum_f e_f)^2 - sum_f e_f^2) instead of its sum over D; tile may be NULL.  Backward:
 * row_grads[b,f,:] = d_tile[b,f,:] (nullable) + d_bi[b,:] * (S[b,:] - e[b,f,:]). */
int ctr_embed_bi_fwd(const float* table, const int64_t* field_row_offset, const int64_t* ids, int64_t B, int64_t F, int64_t D,
                     float* tile, float* bi, void* stream);
int ctr_embed_bi_bwd(const float* tile, const float* d_tile, const float* d_bi, int64_t B, int64_t F, int64_t D,
                     float* row_grads, void* stream);

/* FwFM second-order logit (FwFM/fwfm.py:140-158): out[b] = sum_{i<j} r[pair(i,j)] * <tile[b,i,:], tile[b,j,:]>, r (F(F-1)/2,)
 * indexed like utils.py:67-82 (row-major strict upper triangle).  Backward: d_tile (B,F,K) and d_r (overwritten). */
int ctr_fwfm_fwd(const float* tile, const float* r, int64_t B, int64_t F, int64_t K, float* out, void* stream);
int ctr_fwfm_bwd(const float* tile, const float* r, const float* g, int64_t B, int64_t F, int64_t K, float* d_tile, float* d_r,
                 void* stream);

/* FFM second-order logit (FFM/ffm.py:128-160).  tile (B, F, F-1, K): the lookup of a table whose row for an id of field i is the
 * concatenation of its F-1 sub-embeddings, slot s facing field j = s+1 (s >= i) or s (s < i) -- i.e. the reference's
 * embedding_variables[i] of shape (F-1, |V_i|, K) stored id-major.  out[b] = sum_{i<j} <tile[b,i,j-1,:], tile[b,j,i,:]>.
 * Backward: d_tile[b,i,s,:] = g[b] * tile[b, partner(i,s), :] (overwritten). */
int ctr_ffm_fwd(const float* tile, int64_t B, int64_t F, int64_t K, float* out, void* stream);
int ctr_ffm_bwd(const float* tile, const float* g, int64_t B, int64_t F, int64_t K, float* d_tile, void* stream);

/* AFM attention pooling (AFM/afm.py:152-186): pairs (i<j) in the reference's order, a_p = h^T relu(W^T (e_i*e_j) + b),
 * softmax over the pair axis, pooled (B,K) = sum_p s_p (e_i*e_j).  w (K,T) row-major, b (T,), h (T,); score (B,P) optional
 * output.  K in {4,8,16,32} with K*ceil(T/32) <= 64.  Backward overwrites d_tile (B,F,K), d_w, d_b, d_h. */
int ctr_afm_fwd(const float* tile, const float* w, const float* b, const float* h, int64_t B, int64_t F, int64_t K, int64_t T,
                float* pooled, float* score, void* stream);
int ctr_afm_bwd(const float* tile, const float* w, const float* b, const float* h, const float* g_pooled, int64_t B, int64_t F,
                int64_t K, int64_t T, float* d_tile, float* d_w, float* d_b, float* d_h, void* stream);

/* BST transformer block (BST/transformer_layer.py:6-79): queries/keys/values (B,T,d), keys_length (B,) int64 (mask t >= length,
 * applied along the QUERY axis as the reference does, float32 collapse included), heads >= 1 (each projecting to d),
 * position embedding rows [0,T) of a (max_length,d) table added to queries and keys.  params packed as
 *   position_embedding (max_length,d) | w_q (H,d,d) | w_k | w_v | w_o (H*d,d) | LayerNorm beta,gamma (d,d) |
 *   dense kernel (d,d), bias (d) | LayerNorm_1 beta,gamma        = ctr_bst_param_count(d, heads, max_length) floats.
 * out (B,T,d).  Backward recomputes the forward; d_queries/d_keys/d_values (B,T,d) and d_params (same packing) are
 * overwritten.  One CTA per sample: d in {4,8,16,32,64}, T <= 128, heads <= 16 and a shared-memory footprint <= 220 KB, else -2. */
int64_t ctr_bst_param_count(int64_t d, int64_t heads, int64_t max_length);
int ctr_bst_transformer_fwd(const float* queries, const float* keys, const float* values, const int64_t* keys_length,
                            const float* params, int64_t B, int64_t T, int64_t d, int64_t heads, int64_t max_length,
                            int use_position_embedding, float* out, void* stream);
int ctr_bst_transformer_bwd(const float* queries, const float* keys, const float* values, const int64_t* keys_length,
                            const float* params, const float* g_out, int64_t B, int64_t T, int64_t d, int64_t heads,
                            int64_t max_length, int use_position_embedding, float* d_queries, float* d_keys, float* d_values,
                            float* d_params, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTR_B200_H_ */
