// Row (e) of SURVEY.md section 8: row-sharded embedding tables across the GPUs of one NVSwitch box.
//
// The reference has no multi-device code (the only trace is a commented parameter-server block,
// WideAndDeep/wide_and_deep.py:41-51).  Semantics kept: the table gradient is the IndexedSlices pair (row, value) and it
// has to reach the row's owner before the optimizer can apply it.
//
// Layout: global row gr (= field_row_offset[f] + id) is owned by rank gr % G and stored at local row gr / G.
//   forward : ctr_embed_fm2_fwd_sharded (embed_fm2.cu) PULLS rows from the owners' shards through NVLink peer mappings
//             inside the gather kernel itself -- no id exchange, no return all-to-all.
//   backward: ctr_sharded_grad_push PUSHES every (local_row, grad row) straight into the owner's receive buffer with
//             peer stores from the kernel that reads row_grads -- the all-to-all(v) of the gradients is fused into it.
//             Slots are assigned hierarchically (warp ballots -> shared-memory counters -> one global atomic per CTA
//             iteration and owner), so there is no contended per-row atomic.  ctr_sharded_publish_counts then tells each
//             owner how many entries every source wrote; after a cross-rank barrier the owner consumes
//             (rows, values, count) with ctr_rows_scatter_add or its optimizer.
#include "ctr_common.cuh"

namespace ctr {

struct PeerRecv {
  float4* vals[8];          // owner d: (G_src, capacity, D) fp32
  long long* rows[8];       // owner d: (G_src, capacity) int64 local rows
  int G, logG, my_rank;
  long long capacity;
};

constexpr int PUSH_WARPS = 8;

template <int LPR>
__global__ void __launch_bounds__(PUSH_WARPS * 32)
sharded_grad_push_kernel(const float4* __restrict__ row_grads, const long long* __restrict__ row_off,
                         const long long* __restrict__ ids, int B, int F, const PeerRecv pr,
                         unsigned long long* __restrict__ counters, int* __restrict__ overflow) {
  constexpr int RPW = 32 / LPR;
  __shared__ unsigned int s_cnt[8];                     // rows per owner in this CTA iteration
  __shared__ unsigned int s_wbase[PUSH_WARPS][8];       // warp's first slot inside the CTA iteration, per owner
  __shared__ unsigned long long s_base[8];              // CTA iteration's first slot in the global (per-source) numbering
  const unsigned full = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR, c = lane % LPR;
  const int G = pr.G;
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int iters = (B + PUSH_WARPS - 1) / PUSH_WARPS;
  for (int it = blockIdx.x; it < iters; it += gridDim.x) {
    const int b = it * PUSH_WARPS + warp;
    const bool have = b < B;
    // ---- phase 1: count this warp's rows per owner
    unsigned int wcnt[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) wcnt[d] = 0;
    if (have) {
      for (int f0 = 0; f0 < F; f0 += 32) {
        const int f = f0 + lane;
        int dest = -1;
        if (f < F) {
          const long long id = __ldg(ids + (size_t)b * F + f);
          const long long lo = __ldg(row_off + f), hi = __ldg(row_off + f + 1);
          if (id >= 0 && id < hi - lo) dest = (int)((lo + id) & (G - 1));
        }
#pragma unroll
        for (int d = 0; d < 8; ++d)
          if (d < G) wcnt[d] += __popc(__ballot_sync(full, dest == d));
      }
    }
    if (lane < G) {
      unsigned int mine = 0;
#pragma unroll
      for (int d = 0; d < 8; ++d)
        if (lane == d) mine = wcnt[d];
      s_wbase[warp][lane] = atomicAdd(&s_cnt[lane], mine);
    }
    __syncthreads();
    if (threadIdx.x < G) s_base[threadIdx.x] = atomicAdd(&counters[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
    __syncthreads();
    // ---- phase 2: write (local row, gradient row) into the owners' receive buffers
    if (have) {
      unsigned int woff[8];
#pragma unroll
      for (int d = 0; d < 8; ++d) woff[d] = 0;
      for (int f0 = 0; f0 < F; f0 += 32) {
        const int nf = min(32, F - f0);
        const int f = f0 + lane;
        int dest = -1;
        long long lrow = 0;
        if (f < F) {
          const long long id = __ldg(ids + (size_t)b * F + f);
          const long long lo = __ldg(row_off + f), hi = __ldg(row_off + f + 1);
          if (id >= 0 && id < hi - lo) {
            dest = (int)((lo + id) & (G - 1));
            lrow = (lo + id) >> pr.logG;
          }
        }
        long long slot = -1;
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          if (d < G) {
            const unsigned mask = __ballot_sync(full, dest == d);
            if (dest == d) slot = (long long)s_base[d] + s_wbase[warp][d] + woff[d] + __popc(mask & lt);
            woff[d] += __popc(mask);
          }
        }
        if (slot >= pr.capacity) { atomicOr(overflow, 1); slot = -1; }
        if (slot >= 0) pr.rows[dest][(size_t)pr.my_rank * pr.capacity + slot] = lrow;
        // move the rows: LPR lanes per row, RPW rows per step
        for (int r0 = 0; r0 < nf; r0 += RPW) {
          const int fs = r0 + sub;
          const long long sl = __shfl_sync(full, slot, fs & 31);
          const int ds = __shfl_sync(full, dest, fs & 31);
          if (fs < nf && sl >= 0) {
            const float4 v = ldg_stream_f4(row_grads + ((size_t)b * F + f0 + fs) * LPR + c);
            pr.vals[ds][((size_t)pr.my_rank * pr.capacity + sl) * LPR + c] = v;
          }
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
    __syncthreads();
  }
}

__global__ void sharded_publish_counts_kernel(const unsigned long long* __restrict__ counters, long long* const* peer_counts,
                                              int G, int my_rank) {
  const int d = threadIdx.x;
  if (d < G) peer_counts[d][my_rank] = (long long)counters[d];
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

}  // namespace ctr

using namespace ctr;

static int check_g(const char* fn, int64_t G, int64_t rank) {
  CTR_REQUIRE(G >= 1 && G <= 8 && (G & (G - 1)) == 0, "%s: G=%lld must be a power of two <= 8", fn, (long long)G);
  CTR_REQUIRE(rank >= 0 && rank < G, "%s: rank %lld out of range", fn, (long long)rank);
  return CTR_OK;
}

extern "C" int ctr_sharded_grad_push(const float* row_grads, const int64_t* field_row_offset, const int64_t* ids, int64_t B,
                                     int64_t F, int64_t D, int64_t G, int64_t my_rank, float* const* recv_vals,
                                     int64_t* const* recv_rows, int64_t capacity, int64_t* counters, int* overflow,
                                     void* stream) {
  int rc = check_g("ctr_sharded_grad_push", G, my_rank);
  if (rc) return rc;
  CTR_REQUIRE(row_grads && field_row_offset && ids && recv_vals && recv_rows && counters && overflow,
              "ctr_sharded_grad_push: null argument");
  CTR_REQUIRE(B >= 0 && F >= 1 && capacity >= 0, "ctr_sharded_grad_push: bad sizes");
  CTR_UNSUPPORTED(D % 4 != 0 || D > 128 || (D & (D - 1)) != 0, "ctr_sharded_grad_push: D=%lld unsupported", (long long)D);
  cudaStream_t st = as_stream(stream);
  CTR_CUDA(cudaMemsetAsync(counters, 0, sizeof(int64_t) * G, st));
  if (B == 0) return CTR_OK;
  PeerRecv pr = {};
  pr.G = (int)G;
  while ((1 << pr.logG) < G) ++pr.logG;
  pr.my_rank = (int)my_rank;
  pr.capacity = capacity;
  for (int r = 0; r < G; ++r) {
    CTR_REQUIRE(recv_vals[r] && recv_rows[r] && aligned16(recv_vals[r]), "ctr_sharded_grad_push: receive buffer %d null/unaligned", r);
    pr.vals[r] = reinterpret_cast<float4*>(recv_vals[r]);
    pr.rows[r] = reinterpret_cast<long long*>(recv_rows[r]);
  }
  const long long iters = (B + PUSH_WARPS - 1) / PUSH_WARPS;
  const int grid = (int)(iters < (long long)sm_count() * 4 ? iters : (long long)sm_count() * 4);
  auto* rg = reinterpret_cast<const float4*>(row_grads);
  auto* off = reinterpret_cast<const long long*>(field_row_offset);
  auto* idp = reinterpret_cast<const long long*>(ids);
  auto* cnt = reinterpret_cast<unsigned long long*>(counters);
  switch (D / 4) {
    case 1: sharded_grad_push_kernel<1><<<grid, PUSH_WARPS * 32, 0, st>>>(rg, off, idp, (int)B, (int)F, pr, cnt, overflow); break;
    case 2: sharded_grad_push_kernel<2><<<grid, PUSH_WARPS * 32, 0, st>>>(rg, off, idp, (int)B, (int)F, pr, cnt, overflow); break;
    case 4: sharded_grad_push_kernel<4><<<grid, PUSH_WARPS * 32, 0, st>>>(rg, off, idp, (int)B, (int)F, pr, cnt, overflow); break;
    case 8: sharded_grad_push_kernel<8><<<grid, PUSH_WARPS * 32, 0, st>>>(rg, off, idp, (int)B, (int)F, pr, cnt, overflow); break;
    case 16: sharded_grad_push_kernel<16><<<grid, PUSH_WARPS * 32, 0, st>>>(rg, off, idp, (int)B, (int)F, pr, cnt, overflow); break;
    default: sharded_grad_push_kernel<32><<<grid, PUSH_WARPS * 32, 0, st>>>(rg, off, idp, (int)B, (int)F, pr, cnt, overflow); break;
  }
  CTR_CHECK_LAUNCH("ctr_sharded_grad_push");
  return CTR_OK;
}

extern "C" int ctr_sharded_publish_counts(const int64_t* counters, int64_t* const* peer_counts_dev, int64_t G,
                                          int64_t my_rank, void* stream) {
  int rc = check_g("ctr_sharded_publish_counts", G, my_rank);
  if (rc) return rc;
  CTR_REQUIRE(counters && peer_counts_dev, "ctr_sharded_publish_counts: null argument");
  sharded_publish_counts_kernel<<<1, 32, 0, as_stream(stream)>>>(reinterpret_cast<const unsigned long long*>(counters),
                                                                 reinterpret_cast<long long* const*>(peer_counts_dev), (int)G,
                                                                 (int)my_rank);
  CTR_CHECK_LAUNCH("ctr_sharded_publish_counts");
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
