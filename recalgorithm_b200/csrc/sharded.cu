// Row (e) of SURVEY.md section 8: row-sharded embedding tables across the GPUs of one NVSwitch box.
//
// The reference has no multi-device code (the only trace is a commented parameter-server block,
// WideAndDeep/wide_and_deep.py:41-51).  Semantics kept: the table gradient is the IndexedSlices pair (row, value) and it
// has to reach the row's owner before the optimizer can apply it.
//
// Layout: global row gr (= field_row_offset[f] + id) is owned by rank gr % G and stored at local row gr / G.
//   forward : ctr_embed_fm2_fwd_sharded (embed_fm2.cu) PULLS rows from the owners' shards through NVLink peer mappings
//             inside the gather kernel itself -- no id exchange, no return all-to-all.
//   plan    : ctr_sharded_plan (one pass over the ids, independent of the forward): every valid (b,f) gets a slot in its
//             owner's receive queue -- CTA-level counting in shared memory, ONE global atomic per CTA chunk and owner -- and
//             the queue's row indices are staged per owner in shared memory and leave as contiguous runs (8-byte scattered
//             peer stores would cost a link packet each).  The last CTA publishes the per-owner counts.
//   backward: ctr_embed_fm2_bwd_push computes the IndexedSlices values d_tile + g*(S - e) in registers and stores each row
//             STRAIGHT into its owner's queue with 128-bit peer stores: the all-to-all(v) of the gradients is fused into the
//             backward kernel and row_grads never touches local HBM.  ctr_sharded_grad_push is the same exchange for row
//             gradients that already exist (any other interaction layer upstream).
//   After a stream sync + cross-rank barrier the owner consumes (rows, values, count) with ctr_adam_rows_dedup or
//   ctr_rows_scatter_add.
// Measured link ceilings for this access pattern (tools/peerbench.cu, all ranks active at once, 128-byte rows): pull
// 650 GB/s with L1-allocating loads (622 with .nc.L1::no_allocate), push 680-690 GB/s with plain 128-bit stores, per
// direction per rank; random rows == sequential rows, bulk-async (TMA) copies == LDG/STG, 8 GB == 32 GB shards.
#include "ctr_common.cuh"

namespace ctr {

struct PeerQueues {
  float4* vals[8];          // owner d: (G_src, capacity, D) fp32
  long long* rows[8];       // owner d: (G_src, capacity) int64 local rows
  long long* counts[8];     // owner d: (G_src,) int64 filled slots per source (published by the plan kernel)
  int G, logG, my_rank;
  long long capacity;
};

constexpr int PLAN_THREADS = 256;
constexpr int PLAN_IPT = 5;                          // ids per thread and chunk
constexpr int PLAN_CHUNK = PLAN_THREADS * PLAN_IPT;  // 1280 ids = 32 samples at F = 40
constexpr int PLAN_SLOT_BITS = 28;                   // plan word = owner << 28 | slot ; -1 = invalid id / dropped
constexpr int PLAN_SLOT_MASK = (1 << PLAN_SLOT_BITS) - 1;

__device__ __forceinline__ long long plan_load_id(const long long* p) { return ldg_stream_i64(p); }
__device__ __forceinline__ long long plan_load_id(const int* p) { return (long long)__ldg(p); }

template <typename IdT>
__global__ void __launch_bounds__(PLAN_THREADS)
sharded_plan_kernel(const long long* __restrict__ row_off, const IdT* __restrict__ ids, long long n, int F,
                    const PeerQueues q, unsigned long long* __restrict__ counters, int* __restrict__ overflow,
                    int* __restrict__ plan) {
  __shared__ unsigned int s_cnt[8], s_off[9];
  __shared__ unsigned long long s_base[8];
  __shared__ long long s_rows[PLAN_CHUNK];
  __shared__ int s_last;
  const int tid = threadIdx.x, G = q.G;
  const long long nchunks = (n + PLAN_CHUNK - 1) / PLAN_CHUNK;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    if (tid < 8) s_cnt[tid] = 0;
    __syncthreads();
    int owner[PLAN_IPT];
    unsigned int pos[PLAN_IPT];
    long long lrow[PLAN_IPT];
#pragma unroll
    for (int k = 0; k < PLAN_IPT; ++k) {
      const long long e = ch * PLAN_CHUNK + k * PLAN_THREADS + tid;
      owner[k] = -1; pos[k] = 0; lrow[k] = 0;
      if (e < n) {
        const int f = (int)(e % F);
        const long long id = plan_load_id(ids + e);
        const long long lo = __ldg(row_off + f), hi = __ldg(row_off + f + 1);
        if (id >= 0 && id < hi - lo) {
          owner[k] = (int)((lo + id) & (G - 1));
          lrow[k] = (lo + id) >> q.logG;
          pos[k] = atomicAdd(&s_cnt[owner[k]], 1u);
        }
      }
    }
    __syncthreads();
    if (tid < G) s_base[tid] = atomicAdd(&counters[tid], (unsigned long long)s_cnt[tid]);
    if (tid == 32) {
      unsigned int acc = 0;
      for (int d = 0; d < 8; ++d) { s_off[d] = acc; acc += d < G ? s_cnt[d] : 0u; }
      s_off[8] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PLAN_IPT; ++k) {
      const long long e = ch * PLAN_CHUNK + k * PLAN_THREADS + tid;
      if (e < n) {
        int word = -1;
        if (owner[k] >= 0) {
          const unsigned long long slot = s_base[owner[k]] + pos[k];
          s_rows[s_off[owner[k]] + pos[k]] = lrow[k];
          if (slot < (unsigned long long)q.capacity) word = (owner[k] << PLAN_SLOT_BITS) | (int)slot;
          else atomicOr(overflow, 1);
        }
        plan[e] = word;
      }
    }
    __syncthreads();
    // the chunk's row indices leave as one contiguous run per owner
    const unsigned int total = s_off[8];
    for (unsigned int j = tid; j < total; j += PLAN_THREADS) {
      int d = 0;
#pragma unroll
      for (int t = 1; t < 8; ++t) d += (t < G && j >= s_off[t]) ? 1 : 0;
      const unsigned long long slot = s_base[d] + (j - s_off[d]);
      if (slot < (unsigned long long)q.capacity) q.rows[d][(size_t)q.my_rank * q.capacity + slot] = s_rows[j];
    }
    __syncthreads();
  }
  // the last CTA to finish publishes how many entries this rank queued at every owner (clamped to the capacity)
  __threadfence();
  if (tid == 0) s_last = atomicAdd(&counters[8], 1ull) == (unsigned long long)gridDim.x - 1;
  __syncthreads();
  if (s_last && tid < G) {
    unsigned long long c = atomicAdd(&counters[tid], 0ull);
    if (c > (unsigned long long)q.capacity) c = (unsigned long long)q.capacity;
    if (q.counts[tid] != nullptr) q.counts[tid][q.my_rank] = (long long)c;
  }
}

__device__ __forceinline__ void stg_f4(float4* p, const float4& v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4* queue_slot(const PeerQueues& q, size_t qbase, int word, int lpr, int c) {
  return q.vals[word >> PLAN_SLOT_BITS] + (qbase + (size_t)(word & PLAN_SLOT_MASK)) * lpr + c;
}

// Lookup backward fused with the gradient exchange (arithmetic of embed_fm2_bwd_kernel): warp per sample, the sample's
// tile row is held in registers between the S pass and the gradient pass (HOLD float4 per lane; HOLD == 0 = generic
// two-pass form); every finished 128-bit piece goes to owner.vals[(my_rank*capacity + slot)*LPR + c].
// LIN (HOLD > 0 only): the upstream gradient of the tile is the rank-1 product d_lin[b]*wlin[f,d] of a fused dense(1) head
// (ctr_embed_fm2_lin_fwd): d_tile is not read, `d_tile` carries wlin (F*D) instead, and d_wlin = sum_b d_lin[b]*e[b] is
// accumulated in registers (per CTA one shared-memory reduction + one vector red.global.add per element).
template <int LPR, int HOLD, bool LIN = false>
__global__ void __launch_bounds__(256, LIN ? 2 : 1)
embed_fm2_bwd_push_kernel(const float4* __restrict__ tile, const float4* __restrict__ d_tile, const float* __restrict__ d_fm2,
                          const int* __restrict__ plan, int B, int F, const PeerQueues q, float4* __restrict__ row_grads,
                          const float* __restrict__ d_lin, float4* __restrict__ d_wlin) {
  extern __shared__ float4 s_lin[];                 // LIN: [n4] wlin, then [n4] d_wlin accumulator
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int n4 = F * LPR;
  const size_t qbase = (size_t)q.my_rank * q.capacity;
  constexpr int HA = (LIN && HOLD > 0) ? HOLD : 1;
  float4 acc[HA];
  if (LIN) {
    for (int j = threadIdx.x; j < n4; j += blockDim.x) { s_lin[j] = __ldg(d_tile + j); s_lin[n4 + j] = make_float4(0.f, 0.f, 0.f, 0.f); }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < HA; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int b = warp0; b < B; b += nwarps) {
    const float4* e_row = tile + (size_t)b * n4;
    const float4* dt_row = (d_tile && !LIN) ? d_tile + (size_t)b * n4 : nullptr;
    const int* p_row = plan + (size_t)b * F;
    float4* o_row = row_grads ? row_grads + (size_t)b * n4 : nullptr;
    const float g = d_fm2 ? __ldg(d_fm2 + b) : 0.f;
    const float gl = (LIN && d_lin) ? __ldg(d_lin + b) : 0.f;
    float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
    if (HOLD > 0) {
      constexpr int H = HOLD > 0 ? HOLD : 1;
      float4 e[H], dt[H];
      int pw[H];
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const int j = k * 32 + lane;
        e[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < n4) e[k] = ldg_stream_f4(e_row + j);
      }
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const int j = k * 32 + lane;
        dt[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        pw[k] = -1;
        if (j < n4) {
          if (LIN) { const float4 w = s_lin[j]; dt[k] = make_float4(gl * w.x, gl * w.y, gl * w.z, gl * w.w); }
          else if (dt_row != nullptr) dt[k] = ldg_stream_f4(dt_row + j);
          pw[k] = __ldg(p_row + j / LPR);
        }
      }
#pragma unroll
      for (int k = 0; k < H; ++k) { S.x += e[k].x; S.y += e[k].y; S.z += e[k].z; S.w += e[k].w; }
#pragma unroll
      for (int o = LPR; o < 32; o <<= 1) {
        S.x += __shfl_xor_sync(full, S.x, o); S.y += __shfl_xor_sync(full, S.y, o);
        S.z += __shfl_xor_sync(full, S.z, o); S.w += __shfl_xor_sync(full, S.w, o);
      }
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const int j = k * 32 + lane;
        if (j < n4) {
          float4 r;
          r.x = dt[k].x + g * (S.x - e[k].x); r.y = dt[k].y + g * (S.y - e[k].y);
          r.z = dt[k].z + g * (S.z - e[k].z); r.w = dt[k].w + g * (S.w - e[k].w);
          if (o_row != nullptr) stg_stream_f4(o_row + j, r);
          if (pw[k] >= 0) stg_f4(queue_slot(q, qbase, pw[k], LPR, j % LPR), r);
          if (LIN) { acc[k % HA].x += gl * e[k].x; acc[k % HA].y += gl * e[k].y; acc[k % HA].z += gl * e[k].z; acc[k % HA].w += gl * e[k].w; }
        }
      }
    } else {
      for (int j = lane; j < n4; j += 32) {
        const float4 v = __ldg(e_row + j);
        S.x += v.x; S.y += v.y; S.z += v.z; S.w += v.w;
      }
#pragma unroll
      for (int o = LPR; o < 32; o <<= 1) {
        S.x += __shfl_xor_sync(full, S.x, o); S.y += __shfl_xor_sync(full, S.y, o);
        S.z += __shfl_xor_sync(full, S.z, o); S.w += __shfl_xor_sync(full, S.w, o);
      }
      for (int j = lane; j < n4; j += 32) {
        const float4 v = __ldg(e_row + j);
        float4 r = dt_row ? ldg_stream_f4(dt_row + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        r.x += g * (S.x - v.x); r.y += g * (S.y - v.y); r.z += g * (S.z - v.z); r.w += g * (S.w - v.w);
        if (o_row != nullptr) stg_stream_f4(o_row + j, r);
        const int pw = __ldg(p_row + j / LPR);
        if (pw >= 0) stg_f4(queue_slot(q, qbase, pw, LPR, j % LPR), r);
      }
    }
  }
  if (LIN) {
#pragma unroll
    for (int k = 0; k < HA; ++k) {
      const int j = k * 32 + lane;
      if (j < n4) {
        float* a = reinterpret_cast<float*>(s_lin + n4 + j);
        atomicAdd(a + 0, acc[k].x); atomicAdd(a + 1, acc[k].y); atomicAdd(a + 2, acc[k].z); atomicAdd(a + 3, acc[k].w);
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n4; j += blockDim.x) atomicAdd(d_wlin + j, s_lin[n4 + j]);
  }
}

// The exchange alone: row_grads (B,F,D) already exist; every planned row is copied into its owner's queue.
template <int LPR>
__global__ void __launch_bounds__(256)
sharded_push_rows_kernel(const float4* __restrict__ row_grads, const int* __restrict__ plan, long long n_rows, const PeerQueues q) {
  const size_t total = (size_t)n_rows * LPR;
  const size_t qbase = (size_t)q.my_rank * q.capacity;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int pw = __ldg(plan + t / LPR);
    if (pw >= 0) stg_f4(queue_slot(q, qbase, pw, LPR, (int)(t % LPR)), ldg_stream_f4(row_grads + t));
  }
}

// dst[rows[i], :] += vals[i, :] for i < min(*count, max_n); rows outside [0, V) are ignored.
template <int LPR>
__global__ void __launch_bounds__(256)
rows_scatter_add_kernel(float4* __restrict__ dst, long long V, const long long* __restrict__ rows,
                        const float4* __restrict__ vals, const long long* __restrict__ count, long long max_n) {
  long long n = count ? *count : max_n;
  if (n > max_n) n = max_n;
  const size_t total = (size_t)n * LPR;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const long long row = __ldg(rows + t / LPR);
    if (row >= 0 && row < V) atomicAdd(dst + (size_t)row * LPR + t % LPR, ldg_stream_f4(vals + t));
  }
}

template <typename K>
static int resident_grid_sh(K kernel, int block, long long blocks_needed) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  long long g = (long long)per_sm * sm_count();
  if (g > blocks_needed) g = blocks_needed;
  return (int)(g < 1 ? 1 : g);
}

static int check_g(const char* fn, int64_t G, int64_t rank) {
  CTR_REQUIRE(G >= 1 && G <= 8 && (G & (G - 1)) == 0, "%s: G=%lld must be a power of two <= 8", fn, (long long)G);
  CTR_REQUIRE(rank >= 0 && rank < G, "%s: rank %lld out of range", fn, (long long)rank);
  return CTR_OK;
}

static int fill_queues(const char* fn, PeerQueues& q, int64_t G, int64_t my_rank, float* const* recv_vals,
                       int64_t* const* recv_rows, int64_t* const* recv_counts, int64_t capacity) {
  int rc = check_g(fn, G, my_rank);
  if (rc) return rc;
  CTR_REQUIRE(capacity >= 0 && capacity <= PLAN_SLOT_MASK, "%s: capacity %lld must be in [0, 2^28)", fn, (long long)capacity);
  q = PeerQueues{};
  q.G = (int)G;
  while ((1 << q.logG) < G) ++q.logG;
  q.my_rank = (int)my_rank;
  q.capacity = capacity;
  for (int r = 0; r < G; ++r) {
    if (recv_vals) {
      CTR_REQUIRE(recv_vals[r] && aligned16(recv_vals[r]), "%s: value queue %d null/unaligned", fn, r);
      q.vals[r] = reinterpret_cast<float4*>(recv_vals[r]);
    }
    if (recv_rows) {
      CTR_REQUIRE(recv_rows[r] != nullptr, "%s: row queue %d is null", fn, r);
      q.rows[r] = reinterpret_cast<long long*>(recv_rows[r]);
    }
    if (recv_counts) q.counts[r] = reinterpret_cast<long long*>(recv_counts[r]);
  }
  return CTR_OK;
}

template <int LPR, int HOLD>
static int launch_bwd_push(const float* tile, const float* d_tile, const float* d_fm2, const int32_t* plan, int64_t B, int64_t F,
                           const PeerQueues& q, float* row_grads, cudaStream_t st) {
  auto k = embed_fm2_bwd_push_kernel<LPR, HOLD, false>;
  const int grid = resident_grid_sh(k, 256, (B + 7) / 8);
  k<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(tile), reinterpret_cast<const float4*>(d_tile), d_fm2, plan, (int)B,
                          (int)F, q, reinterpret_cast<float4*>(row_grads), nullptr, nullptr);
  CTR_CHECK_LAUNCH("ctr_embed_fm2_bwd_push");
  return CTR_OK;
}

template <int LPR, int HOLD>
static int launch_lin_bwd_push(const float* tile, const float* wlin, const float* d_fm2, const float* d_lin, const int32_t* plan,
                               int64_t B, int64_t F, const PeerQueues& q, float* row_grads, float* d_wlin, cudaStream_t st) {
  auto k = embed_fm2_bwd_push_kernel<LPR, HOLD, true>;
  const size_t smem = sizeof(float4) * 2 * (size_t)F * LPR;
  if (smem > 48 * 1024) CTR_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  long long grid = (long long)per_sm * sm_count();
  if (grid > (B + 7) / 8) grid = (B + 7) / 8;
  k<<<(int)grid, 256, smem, st>>>(reinterpret_cast<const float4*>(tile), reinterpret_cast<const float4*>(wlin), d_fm2, plan, (int)B,
                                  (int)F, q, reinterpret_cast<float4*>(row_grads), d_lin, reinterpret_cast<float4*>(d_wlin));
  CTR_CHECK_LAUNCH("ctr_embed_fm2_lin_bwd_push");
  return CTR_OK;
}

template <int LPR>
static int dispatch_lin_bwd_push(const float* tile, const float* wlin, const float* d_fm2, const float* d_lin, const int32_t* plan,
                                 int64_t B, int64_t F, const PeerQueues& q, float* row_grads, float* d_wlin, cudaStream_t st) {
  const int64_t per_lane = (F * LPR + 31) / 32;
  if (per_lane <= 4) return launch_lin_bwd_push<LPR, 4>(tile, wlin, d_fm2, d_lin, plan, B, F, q, row_grads, d_wlin, st);
  if (per_lane <= 8) return launch_lin_bwd_push<LPR, 8>(tile, wlin, d_fm2, d_lin, plan, B, F, q, row_grads, d_wlin, st);
  if (per_lane <= 12) return launch_lin_bwd_push<LPR, 12>(tile, wlin, d_fm2, d_lin, plan, B, F, q, row_grads, d_wlin, st);
  set_error("ctr_embed_fm2_lin_bwd_push: F*D = %lld exceeds the register-resident limit of 1536", (long long)(F * LPR * 4));
  return CTR_ERR_UNSUPPORTED;
}

template <int LPR>
static int dispatch_bwd_push(const float* tile, const float* d_tile, const float* d_fm2, const int32_t* plan, int64_t B,
                             int64_t F, const PeerQueues& q, float* row_grads, cudaStream_t st) {
  const int64_t per_lane = (F * LPR + 31) / 32;
  if (per_lane <= 4) return launch_bwd_push<LPR, 4>(tile, d_tile, d_fm2, plan, B, F, q, row_grads, st);
  if (per_lane <= 8) return launch_bwd_push<LPR, 8>(tile, d_tile, d_fm2, plan, B, F, q, row_grads, st);
  if (per_lane <= 12) return launch_bwd_push<LPR, 12>(tile, d_tile, d_fm2, plan, B, F, q, row_grads, st);
  return launch_bwd_push<LPR, 0>(tile, d_tile, d_fm2, plan, B, F, q, row_grads, st);
}

}  // namespace ctr

using namespace ctr;

extern "C" int ctr_sharded_plan(const int64_t* field_row_offset, const void* ids, int ids_are_int32, int64_t B, int64_t F, int64_t G,
                                int64_t my_rank, int64_t* const* recv_rows, int64_t* const* recv_counts, int64_t capacity,
                                int64_t* counters, int* overflow, int32_t* plan, void* stream) {
  PeerQueues q;
  int rc = fill_queues("ctr_sharded_plan", q, G, my_rank, nullptr, recv_rows, recv_counts, capacity);
  if (rc) return rc;
  CTR_REQUIRE(field_row_offset && ids && recv_rows && counters && overflow && plan, "ctr_sharded_plan: null argument");
  CTR_REQUIRE(B >= 0 && F >= 1 && B * F <= PLAN_SLOT_MASK, "ctr_sharded_plan: bad sizes (B*F must be < 2^28)");
  cudaStream_t st = as_stream(stream);
  CTR_CUDA(cudaMemsetAsync(counters, 0, sizeof(int64_t) * 9, st));
  CTR_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), st));
  const long long n = (long long)B * F;
  const long long chunks = (n + PLAN_CHUNK - 1) / PLAN_CHUNK;
  // B == 0 still runs one CTA: it publishes the zero counts
  const int grid = (int)(chunks < 1 ? 1 : (chunks < (long long)sm_count() * 4 ? chunks : (long long)sm_count() * 4));
  if (ids_are_int32)
    sharded_plan_kernel<int><<<grid, PLAN_THREADS, 0, st>>>(reinterpret_cast<const long long*>(field_row_offset),
                                                            reinterpret_cast<const int*>(ids), n, (int)F, q,
                                                            reinterpret_cast<unsigned long long*>(counters), overflow, plan);
  else
    sharded_plan_kernel<long long><<<grid, PLAN_THREADS, 0, st>>>(reinterpret_cast<const long long*>(field_row_offset),
                                                                  reinterpret_cast<const long long*>(ids), n, (int)F, q,
                                                                  reinterpret_cast<unsigned long long*>(counters), overflow, plan);
  CTR_CHECK_LAUNCH("ctr_sharded_plan");
  return CTR_OK;
}

extern "C" int ctr_embed_fm2_bwd_push(const float* tile, const float* d_tile, const float* d_fm2, const int32_t* plan, int64_t B,
                                      int64_t F, int64_t D, int64_t G, int64_t my_rank, float* const* recv_vals,
                                      int64_t capacity, float* row_grads, void* stream) {
  PeerQueues q;
  int rc = fill_queues("ctr_embed_fm2_bwd_push", q, G, my_rank, recv_vals, nullptr, nullptr, capacity);
  if (rc) return rc;
  CTR_REQUIRE(tile && plan && recv_vals, "ctr_embed_fm2_bwd_push: null tile/plan/recv_vals");
  CTR_REQUIRE(B >= 0 && F >= 1 && B <= 0x7fffffffLL / 8 && F <= 65536, "ctr_embed_fm2_bwd_push: bad sizes");
  CTR_UNSUPPORTED(D % 4 != 0 || D > 128 || (D & (D - 1)) != 0, "ctr_embed_fm2_bwd_push: D=%lld unsupported", (long long)D);
  CTR_REQUIRE(aligned16(tile) && aligned16(d_tile) && aligned16(row_grads),
              "ctr_embed_fm2_bwd_push: tile, d_tile and row_grads must be 16-byte aligned");
  if (B == 0) return CTR_OK;
  cudaStream_t st = as_stream(stream);
  switch (D / 4) {
    case 1: return dispatch_bwd_push<1>(tile, d_tile, d_fm2, plan, B, F, q, row_grads, st);
    case 2: return dispatch_bwd_push<2>(tile, d_tile, d_fm2, plan, B, F, q, row_grads, st);
    case 4: return dispatch_bwd_push<4>(tile, d_tile, d_fm2, plan, B, F, q, row_grads, st);
    case 8: return dispatch_bwd_push<8>(tile, d_tile, d_fm2, plan, B, F, q, row_grads, st);
    case 16: return dispatch_bwd_push<16>(tile, d_tile, d_fm2, plan, B, F, q, row_grads, st);
    default: return dispatch_bwd_push<32>(tile, d_tile, d_fm2, plan, B, F, q, row_grads, st);
  }
}

extern "C" int ctr_embed_fm2_lin_bwd_push(const float* tile, const float* wlin, const float* d_fm2, const float* d_lin,
                                          const int32_t* plan, int64_t B, int64_t F, int64_t D, int64_t G, int64_t my_rank,
                                          float* const* recv_vals, int64_t capacity, float* row_grads, float* d_wlin,
                                          void* stream) {
  PeerQueues q;
  int rc = fill_queues("ctr_embed_fm2_lin_bwd_push", q, G, my_rank, recv_vals, nullptr, nullptr, capacity);
  if (rc) return rc;
  CTR_REQUIRE(tile && wlin && plan && recv_vals && d_wlin, "ctr_embed_fm2_lin_bwd_push: null tile/wlin/plan/recv_vals/d_wlin");
  CTR_REQUIRE(B >= 0 && F >= 1 && B <= 0x7fffffffLL / 8 && F <= 65536, "ctr_embed_fm2_lin_bwd_push: bad sizes");
  CTR_UNSUPPORTED(D % 4 != 0 || D > 128 || (D & (D - 1)) != 0, "ctr_embed_fm2_lin_bwd_push: D=%lld unsupported", (long long)D);
  CTR_REQUIRE(aligned16(tile) && aligned16(wlin) && aligned16(row_grads) && aligned16(d_wlin),
              "ctr_embed_fm2_lin_bwd_push: tile, wlin, row_grads and d_wlin must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  CTR_CUDA(cudaMemsetAsync(d_wlin, 0, sizeof(float) * F * D, st));
  if (B == 0) return CTR_OK;
  switch (D / 4) {
    case 1: return dispatch_lin_bwd_push<1>(tile, wlin, d_fm2, d_lin, plan, B, F, q, row_grads, d_wlin, st);
    case 2: return dispatch_lin_bwd_push<2>(tile, wlin, d_fm2, d_lin, plan, B, F, q, row_grads, d_wlin, st);
    case 4: return dispatch_lin_bwd_push<4>(tile, wlin, d_fm2, d_lin, plan, B, F, q, row_grads, d_wlin, st);
    case 8: return dispatch_lin_bwd_push<8>(tile, wlin, d_fm2, d_lin, plan, B, F, q, row_grads, d_wlin, st);
    case 16: return dispatch_lin_bwd_push<16>(tile, wlin, d_fm2, d_lin, plan, B, F, q, row_grads, d_wlin, st);
    default: return dispatch_lin_bwd_push<32>(tile, wlin, d_fm2, d_lin, plan, B, F, q, row_grads, d_wlin, st);
  }
}

extern "C" int ctr_sharded_grad_push(const float* row_grads, const int32_t* plan, int64_t B, int64_t F, int64_t D, int64_t G,
                                     int64_t my_rank, float* const* recv_vals, int64_t capacity, void* stream) {
  PeerQueues q;
  int rc = fill_queues("ctr_sharded_grad_push", q, G, my_rank, recv_vals, nullptr, nullptr, capacity);
  if (rc) return rc;
  CTR_REQUIRE(row_grads && plan && recv_vals, "ctr_sharded_grad_push: null argument");
  CTR_REQUIRE(B >= 0 && F >= 1, "ctr_sharded_grad_push: bad sizes");
  CTR_UNSUPPORTED(D % 4 != 0 || D > 128 || (D & (D - 1)) != 0, "ctr_sharded_grad_push: D=%lld unsupported", (long long)D);
  CTR_REQUIRE(aligned16(row_grads), "ctr_sharded_grad_push: row_grads must be 16-byte aligned");
  if (B == 0) return CTR_OK;
  cudaStream_t st = as_stream(stream);
  const long long n_rows = (long long)B * F, total = n_rows * (D / 4);
  const int grid = (int)((total + 255) / 256 < (long long)sm_count() * 8 ? (total + 255) / 256 : (long long)sm_count() * 8);
  auto* rg = reinterpret_cast<const float4*>(row_grads);
  switch (D / 4) {
    case 1: sharded_push_rows_kernel<1><<<grid, 256, 0, st>>>(rg, plan, n_rows, q); break;
    case 2: sharded_push_rows_kernel<2><<<grid, 256, 0, st>>>(rg, plan, n_rows, q); break;
    case 4: sharded_push_rows_kernel<4><<<grid, 256, 0, st>>>(rg, plan, n_rows, q); break;
    case 8: sharded_push_rows_kernel<8><<<grid, 256, 0, st>>>(rg, plan, n_rows, q); break;
    case 16: sharded_push_rows_kernel<16><<<grid, 256, 0, st>>>(rg, plan, n_rows, q); break;
    default: sharded_push_rows_kernel<32><<<grid, 256, 0, st>>>(rg, plan, n_rows, q); break;
  }
  CTR_CHECK_LAUNCH("ctr_sharded_grad_push");
  return CTR_OK;
}

extern "C" int ctr_rows_scatter_add(float* dst, int64_t V, int64_t D, const int64_t* rows, const float* vals,
                                    const int64_t* count, int64_t max_n, void* stream) {
  CTR_REQUIRE(dst && rows && vals, "ctr_rows_scatter_add: null argument");
  CTR_REQUIRE(V >= 0 && max_n >= 0, "ctr_rows_scatter_add: bad sizes");
  CTR_UNSUPPORTED(D % 4 != 0 || D > 128 || (D & (D - 1)) != 0, "ctr_rows_scatter_add: D=%lld unsupported", (long long)D);
  CTR_REQUIRE(aligned16(dst) && aligned16(vals), "ctr_rows_scatter_add: buffers must be 16-byte aligned");
  if (max_n == 0) return CTR_OK;
  cudaStream_t st = as_stream(stream);
  const long long total = (long long)max_n * (D / 4);
  const int grid = (int)((total + 255) / 256 < (long long)sm_count() * 16 ? (total + 255) / 256 : (long long)sm_count() * 16);
  auto* d4 = reinterpret_cast<float4*>(dst);
  auto* r = reinterpret_cast<const long long*>(rows);
  auto* v4 = reinterpret_cast<const float4*>(vals);
  auto* cn = reinterpret_cast<const long long*>(count);
  switch (D / 4) {
    case 1: rows_scatter_add_kernel<1><<<grid, 256, 0, st>>>(d4, V, r, v4, cn, max_n); break;
    case 2: rows_scatter_add_kernel<2><<<grid, 256, 0, st>>>(d4, V, r, v4, cn, max_n); break;
    case 4: rows_scatter_add_kernel<4><<<grid, 256, 0, st>>>(d4, V, r, v4, cn, max_n); break;
    case 8: rows_scatter_add_kernel<8><<<grid, 256, 0, st>>>(d4, V, r, v4, cn, max_n); break;
    case 16: rows_scatter_add_kernel<16><<<grid, 256, 0, st>>>(d4, V, r, v4, cn, max_n); break;
    default: rows_scatter_add_kernel<32><<<grid, 256, 0, st>>>(d4, V, r, v4, cn, max_n); break;
  }
  CTR_CHECK_LAUNCH("ctr_rows_scatter_add");
  return CTR_OK;
}
