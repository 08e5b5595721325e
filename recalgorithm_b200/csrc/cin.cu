// Row CIN: the xDeepFM compressed-interaction layer.
//
// Reference: cin_layer(x0, xk, hk_1, index) -- xDeepFM/cin_layer.py:17-30
//   outer[b,d,i,j] = xk[b,i,d] * x0[b,j,d]; reshape (B, D, hk*m) with flat index i*m+j; conv1d with a
//   (1, hk*m, hk_1) filter == matmul over the last axis; transpose -> (B, hk_1, D).  The reference
//   materialises the (B, D, hk*m) outer tensor (2 GB at BASELINE config 3, layer 2).
//
// B200 mapping -- the one genuinely dense contraction of the hot path, so it runs on the 5th-gen tensor cores:
//   GEMM view  C[M x N] = A[M x K] . B[K x N],  M = B*D rows r = (b,d),  K = hk*m,  N = hk_1.
//   * A (the outer product) is never written to HBM: 256 producer threads each own one row r, keep the row's
//     x0[b,:,d] (m <= 32 values) in registers and, for every K-block i, form the 32 products xk[b,i,d]*x0[b,j,d],
//     and store them straight into shared memory in the UMMA canonical K-major SWIZZLE_128B layout
//     (K is re-ordered as i*32 + j, zero padded from m to 32, so one K-block == one i == one 128-byte swizzle row);
//   * B (the filter) is pre-transposed/padded once per call into that K order (workspace) and streamed by TMA
//     (cp.async.bulk.tensor, SWIZZLE_128B) through an mbarrier pipeline;
//   * one elected thread issues tcgen05.mma (kind::tf32, M=128, N=hk_1 padded to 16, K=8); accumulators live in
//     TMEM (2 M-tiles x N columns per CTA so that every B stage is used twice); tcgen05.commit releases stages;
//   * fp32-class accuracy (north_star: 1e-5) comes from the 3xTF32 split  A.B ~= Ahi.Bhi + Alo.Bhi + Ahi.Blo
//     with fp32 accumulation in TMEM (error ~2^-22 per product); precision=1 runs a single TF32 pass (~1e-3);
//   * epilogue: tcgen05.ld -> registers -> (B, hk_1, D) stores + the pooled sum over D by warp shuffles.
//   Shapes outside the tensor path's limits (m > 32, hk_1 > 128, D not a power of two <= 32) use a plain
//   CUDA-core kernel.  The backward pass is CUDA-core in this revision (next step).
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace ctr {
namespace cin {
using namespace ctr::tc;

constexpr int TILES = 2;                 // M tiles per CTA
constexpr int KB = 32;                   // tf32 per K-block (128 B == swizzle span)
constexpr int NTHREADS = 384;            // 3 warpgroups: 8 producer/drain warps, then TMA warp + MMA warp (+2 idle)

struct FwdSmem {
  int stage_bytes, a_bytes_per_tile, b_tile_bytes, bar_off, total;
};
__host__ __device__ inline FwdSmem fwd_smem(int NP, int passes, int stages) {
  FwdSmem s;
  const int na = passes == 3 ? 2 : 1;
  s.a_bytes_per_tile = 0;                         // A lives in TMEM (TS-mode MMA)
  s.b_tile_bytes = NP * 128;
  s.stage_bytes = na * s.b_tile_bytes;
  s.bar_off = stages * s.stage_bytes;
  s.total = s.bar_off + 8 * (2 * stages + 2 * 8 + 2) + 16;
  return s;
}

// ------------------------------------------------------------------------------------------------ kernels
// filter (hk*m, H) -> Wt[2][NP][KP]: [0] = tf32-rounded value, [1] = residual; K order i*32 + j, zero padded.
__global__ void cin_split_filter_kernel(const float* __restrict__ w, float* __restrict__ wt, int m, int hk, int H, int NP) {
  const int KP = hk * KB;
  const size_t total = (size_t)NP * KP;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / KP), kk = (int)(idx % KP);
    const int i = kk / KB, j = kk % KB;
    float v = 0.f;
    if (n < H && j < m) v = __ldg(w + ((size_t)i * m + j) * H + n);
    const float hi = tf32_rna(v);
    wt[idx] = hi;
    wt[total + idx] = v - hi;
  }
}

// One halving-butterfly step of the sum over the D lanes of a sample: lanes whose bit `o` is clear keep the lower
// HALF columns, the others the upper HALF; returns the column offset this lane now owns.
template <int N, int HALF>
__device__ __forceinline__ int bfly_step(float (&acc)[N], int lane, int o) {
  const bool up = (lane & o) != 0;
#pragma unroll
  for (int q = 0; q < HALF; ++q) {
    const float keep = up ? acc[q + HALF] : acc[q];
    const float send = up ? acc[q] : acc[q + HALF];
    acc[q] = keep + __shfl_xor_sync(0xffffffffu, send, o);
  }
  return up ? HALF : 0;
}

// Accumulation note: the tensor core adds each K=8 product group into the fp32 TMEM accumulator with truncation,
// so a long accumulation chain drifts by ~0.5 ulp per MMA (measured 3.8e-5 after 1440 MMAs at K = 3840 x 3 passes).
// The K loop is therefore cut into chunks of `chunk` K-blocks: each chunk accumulates from zero in TMEM and the
// row-owning threads add the finished chunk into fp32 registers (round-to-nearest).
//
// Operand placement (second revision): the first version staged A in shared memory; ncu showed it bound by shared
// memory bandwidth (producer stores + tensor-core operand reads + TMA writes > 128 B/clk) with the proxy fence as the
// top producer stall.  A now goes registers -> TMEM (tcgen05.st; the row-owning thread IS the TMEM lane) and the MMA
// runs in TS mode (A from TMEM, B from shared memory), so shared memory only carries the filter tiles.
//   TMEM columns: [0, TILES*NP) accumulators | [256, 512) A staging: stage sa, tile t at 256 + (sa*TILES + t)*32*NA (+32 = lo)
// Work items of the persistent forward kernel.  A CTA normally processes PAIRS of 128-row M-tiles (both share every filter
// stage); with n128 tiles and G CTAs that quantises to ceil(n128 / 2G) rounds (config 3: 3.46 -> 4 rounds, 13 % idle).
// The last round therefore hands out what is left as evenly as possible: full pairs to the first CTAs, SINGLE tiles to
// the rest (the second M-tile's warps then only keep the barriers in phase), so the tail costs half a round when it can.
struct FwdItem { int tile0, nt; };                   // first 128-row tile, number of M-tiles (0: nothing for this CTA)
__device__ __forceinline__ int fwd_num_items(int n128) {
  const int G = (int)gridDim.x, R = (n128 / 2) / G;
  return R + (n128 - 2 * R * G > 0 ? 1 : 0);
}
__device__ __forceinline__ FwdItem fwd_item(int it, int n128) {
  const int G = (int)gridDim.x, c = (int)blockIdx.x, R = (n128 / 2) / G;
  if (it < R) return FwdItem{2 * (c + it * G), 2};
  const int base = 2 * R * G, rem = n128 - base;      // 0 < rem < 2G
  if (rem <= G) return FwdItem{base + c, c < rem ? 1 : 0};
  const int x = rem - G;                              // CTAs that still get a pair
  return c < x ? FwdItem{base + 2 * c, 2} : FwdItem{base + 2 * x + (c - x), 1};
}

template <int PASSES, int SB, int NPT>
__global__ void __launch_bounds__(NTHREADS, 1)
cin_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const float* __restrict__ x0,
                  const float* __restrict__ xk, float* __restrict__ out, float* __restrict__ pooled, int B, int m,
                  int hk, int logD, int H, int NP, int chunk) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned tiles: align explicitly (the launch adds 1 KB of slack)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int NA = PASSES == 3 ? 2 : 1;                 // hi (+ lo) operand copies
  constexpr int SA = 256 / (TILES * KB * NA);             // A stages that fit TMEM columns [256, 512)
  const FwdSmem L = fwd_smem(NP, PASSES, SB);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + L.bar_off;
  auto full_b = [&](int s) { return bar0 + 8 * s; };
  auto empty_b = [&](int s) { return bar0 + 8 * (SB + s); };
  auto full_a = [&](int s) { return bar0 + 8 * (2 * SB + s); };
  auto empty_a = [&](int s) { return bar0 + 8 * (2 * SB + SA + s); };
  const uint32_t acc_full = bar0 + 8 * (2 * SB + 2 * SA), acc_empty = acc_full + 8;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8 * (2 * SB + 2 * SA + 2));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = 1 << logD;
  const long long rows_total = (long long)B * D;
  const int n128 = (int)((rows_total + BM - 1) / BM);
  const int n_items = fwd_num_items(n128);
  const int nch = (hk + chunk - 1) / chunk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SB; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
    for (int s = 0; s < SA; ++s) { mbar_init(full_a(s), 8); mbar_init(empty_a(s), 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 8);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(smem_u32(tmem_ptr), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t a_stage0 = tmem_base + 256u;

  // Register re-distribution between warpgroups (registers are per SM sub-partition: 3 warps x 168 at launch):
  // the two row-owning warpgroups hold x0 (32) + the running output row (up to 128) in registers.
  if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  }
  if (warp < 8) {
    // ============================ A producers (one row each) + chunk drain + epilogue ============================
    const int t = warp >> 2;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    int sa = 0, pha = 0;
    uint32_t g = 0;                                   // chunk counter (same sequence as the MMA warp)
    float acc[NPT];
    auto drain = [&](uint32_t gi, bool on) {          // add finished chunk gi into acc, then hand TMEM back
      mbar_wait(acc_full, gi & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_sel + (uint32_t)(t * NP);
#pragma unroll
      for (int c0 = 0; c0 < NPT; c0 += 32) {            // two 16-column loads in flight per wait
        if (c0 < NP && on) {
          uint32_t v0[16], v1[16];
          tmem_ld16_nowait(taddr + c0, v0);
          if (c0 + 16 < NP) tmem_ld16_nowait(taddr + c0 + 16, v1);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 16; ++q) acc[c0 + q] += __uint_as_float(v0[q]);
          if (c0 + 16 < NP) {
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[c0 + 16 + q] += __uint_as_float(v1[q]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    };
    for (int it = 0; it < n_items; ++it) {
      const FwdItem item = fwd_item(it, n128);
      if (item.nt == 0) continue;
      const bool on = t < item.nt;                     // single-tile item: the second M-tile's warps only keep the barriers in phase
      const long long r = (long long)(item.tile0 + t) * BM + (warp & 3) * 32 + lane;
      const bool valid = on && r < rows_total;
      const int b = valid ? (int)(r >> logD) : 0;
      const int d = (int)(r & (D - 1));
      float x0v[KB];
#pragma unroll
      for (int j = 0; j < KB; ++j) x0v[j] = (valid && j < m) ? __ldg(x0 + ((size_t)b * m + j) * D + d) : 0.f;
#pragma unroll
      for (int n = 0; n < NPT; ++n) acc[n] = 0.f;
      const float* xkp = xk + (size_t)b * hk * D + d;
      float xnext = valid ? __ldg(xkp) : 0.f;
      for (int c = 0; c < nch; ++c, ++g) {
        const int i_beg = c * chunk, i_end = min(hk, (c + 1) * chunk);
        const int drain_at = min(SA, i_end - i_beg);  // previous chunk is drained once SA blocks of this one are staged
        for (int i = i_beg; i < i_end; ++i) {
          const float xi = xnext;
          if (i + 1 < hk) xnext = valid ? __ldg(xkp + (size_t)(i + 1) * D) : 0.f;
          mbar_wait(empty_a(sa), pha ^ 1);
          tc_fence_after();
          const uint32_t a_hi = a_stage0 + lane_sel + (uint32_t)((sa * TILES + t) * KB * NA);
          if (on) {
#pragma unroll
          for (int part = 0; part < KB / 8; ++part) {
            float p[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) p[q] = xi * x0v[part * 8 + q];
            if (PASSES == 3) {
              float h[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) h[q] = tf32_rna(p[q]);
              tmem_st8(a_hi + part * 8, h);
#pragma unroll
              for (int q = 0; q < 8; ++q) p[q] -= h[q];
              tmem_st8(a_hi + KB + part * 8, p);
            } else {
              tmem_st8(a_hi + part * 8, p);
            }
          }
          tmem_wait_st();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(full_a(sa));
          if (++sa == SA) { sa = 0; pha ^= 1; }
          if (c > 0 && i - i_beg + 1 == drain_at) drain(g - 1, on);
        }
      }
      drain(g - 1, on);
      // ---------------- epilogue: registers -> out (B,H,D) and pooled (B,H) ----------------
#pragma unroll
      for (int n = 0; n < NPT; ++n)
        if (n < H && valid) out[((size_t)b * H + n) * D + d] = acc[n];
      if (pooled != nullptr) {
        // sum over the D lanes of each sample: halving butterfly (after step s a lane keeps NPT >> (s+1) partial columns)
        int col0 = 0;
        if (logD > 0) col0 += bfly_step<NPT, NPT / 2>(acc, lane, D >> 1);
        if (logD > 1) col0 += bfly_step<NPT, NPT / 4>(acc, lane, D >> 2);
        if (logD > 2) col0 += bfly_step<NPT, NPT / 8>(acc, lane, D >> 3);
        if (logD > 3) col0 += bfly_step<NPT, NPT / 16>(acc, lane, D >> 4);
        if (logD > 4) col0 += bfly_step<NPT, NPT / 32>(acc, lane, D >> 5);
        const int cnt = NPT >> logD;
#pragma unroll
        for (int q = 0; q < NPT; ++q)
          if (q < cnt && valid && col0 + q < H) pooled[(size_t)b * H + col0 + q] = acc[q];
      }
    }
  } else if (warp == 8) {
    // ============================ TMA producer for the filter tiles ============================
    if (lane == 0) {
      int s = 0, ph = 0;
      for (int it = 0; it < n_items; ++it) {
        if (fwd_item(it, n128).nt == 0) continue;
        for (int i = 0; i < hk; ++i) {
          mbar_wait(empty_b(s), ph ^ 1);
          const uint32_t b_dst = sbase + s * L.stage_bytes;
          mbar_expect_tx(full_b(s), (uint32_t)(NA * L.b_tile_bytes));
          tma_load_2d(b_dst, &tmap_w, i * KB, 0, full_b(s));
          if (PASSES == 3) tma_load_2d(b_dst + L.b_tile_bytes, &tmap_w, i * KB, NP, full_b(s));
          if (++s == SB) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ============================ MMA issuer (TS mode: A from TMEM, B from shared memory) ============================
    {
      const uint32_t idesc = umma_idesc_tf32(NP);
      int sb = 0, phb = 0, sa = 0, pha = 0;
      uint32_t g = 0;
      for (int it = 0; it < n_items; ++it) {
        const int nt = fwd_item(it, n128).nt;
        if (nt == 0) continue;
        for (int c = 0; c < nch; ++c, ++g) {
          mbar_wait(acc_empty, (g & 1u) ^ 1u);                 // previous chunk has been drained out of TMEM
          tc_fence_after();
          const int i_beg = c * chunk, i_end = min(hk, (c + 1) * chunk);
          for (int i = i_beg; i < i_end; ++i) {
            mbar_wait(full_a(sa), pha);
            mbar_wait(full_b(sb), phb);
            tc_fence_after();
            if (elect_one()) {
            const uint32_t first = (i == i_beg) ? 0u : 1u;
            const uint32_t st_base = sbase + sb * L.stage_bytes;
            const uint64_t b_hi = umma_desc_sw128(st_base);
            const uint64_t b_lo = umma_desc_sw128(st_base + L.b_tile_bytes);
#pragma unroll
            for (int tt = 0; tt < TILES; ++tt) {
              if (tt >= nt) break;
              const uint32_t a_hi = a_stage0 + (uint32_t)((sa * TILES + tt) * KB * NA);
              const uint32_t a_lo = a_hi + KB;
              const uint32_t dcol = tmem_base + (uint32_t)(tt * NP);
              if (PASSES == 3) {
                // small terms first, the dominant hi*hi term last
#pragma unroll
                for (int k = 0; k < KB / 8; ++k) umma_tf32_ts(dcol, a_lo + 8 * k, b_hi + 2 * k, idesc, k > 0 ? 1u : first);
#pragma unroll
                for (int k = 0; k < KB / 8; ++k) umma_tf32_ts(dcol, a_hi + 8 * k, b_lo + 2 * k, idesc, 1u);
#pragma unroll
                for (int k = 0; k < KB / 8; ++k) umma_tf32_ts(dcol, a_hi + 8 * k, b_hi + 2 * k, idesc, 1u);
              } else {
#pragma unroll
                for (int k = 0; k < KB / 8; ++k) umma_tf32_ts(dcol, a_hi + 8 * k, b_hi + 2 * k, idesc, k > 0 ? 1u : first);
              }
            }
            umma_commit(empty_a(sa));                  // A staging columns free once these MMAs have read them
            umma_commit(empty_b(sb));                  // and the filter stage
            if (i + 1 == i_end) umma_commit(acc_full); // chunk complete -> drain
            }
            __syncwarp();
            if (++sa == SA) { sa = 0; pha ^= 1; }
            if (++sb == SB) { sb = 0; phb ^= 1; }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// ---- CUDA-core forward for shapes outside the tensor path.  One CTA per sample.
__global__ void __launch_bounds__(256)
cin_fwd_simple_kernel(const float* __restrict__ x0, const float* __restrict__ xk, const float* __restrict__ w,
                      float* __restrict__ out, float* __restrict__ pooled, int B, int m, int hk, int D, int H) {
  extern __shared__ __align__(16) float sm[];
  float* x0s = sm;
  float* xks = sm + m * D;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < m * D; i += blockDim.x) x0s[i] = __ldg(x0 + (size_t)b * m * D + i);
    for (int i = threadIdx.x; i < hk * D; i += blockDim.x) xks[i] = __ldg(xk + (size_t)b * hk * D + i);
    __syncthreads();
    for (int idx = threadIdx.x; idx < H * D; idx += blockDim.x) {
      const int n = idx / D, d = idx % D;
      float acc = 0.f;
      for (int i = 0; i < hk; ++i) {
        const float a = xks[i * D + d];
        const float* wr = w + (size_t)i * m * H + n;
        for (int j = 0; j < m; ++j) acc += (a * x0s[j * D + d]) * __ldg(wr + (size_t)j * H);
      }
      out[(size_t)b * H * D + idx] = acc;
    }
    if (pooled != nullptr) {
      __syncthreads();
      for (int n = threadIdx.x; n < H; n += blockDim.x) {
        float s = 0.f;
        for (int d = 0; d < D; ++d) s += out[((size_t)b * H + n) * D + d];
        pooled[(size_t)b * H + n] = s;
      }
    }
  }
}

// ---- CUDA-core backward, data gradients.  One CTA per sample.
//   dz[p,d] = sum_n g[n,d]*W[p,n];  dxk[i,d] += dz*x0[j,d];  dx0[j,d] += dz*xk[i,d]      (p = i*m + j)
__global__ void __launch_bounds__(256)
cin_bwd_dx_kernel(const float* __restrict__ x0, const float* __restrict__ xk, const float* __restrict__ w,
                  const float* __restrict__ g, int B, int m, int hk, int D, int H, float* __restrict__ dx0,
                  float* __restrict__ dxk) {
  extern __shared__ __align__(16) float sm[];
  float* x0s = sm;
  float* xks = x0s + m * D;
  float* gs = xks + hk * D;
  float* dx0s = gs + H * D;
  float* dxks = dx0s + m * D;
  const int d = threadIdx.x % D, pl = threadIdx.x / D, pstep = blockDim.x / D;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < m * D; i += blockDim.x) { x0s[i] = __ldg(x0 + (size_t)b * m * D + i); dx0s[i] = 0.f; }
    for (int i = threadIdx.x; i < hk * D; i += blockDim.x) { xks[i] = __ldg(xk + (size_t)b * hk * D + i); dxks[i] = 0.f; }
    for (int i = threadIdx.x; i < H * D; i += blockDim.x) gs[i] = __ldg(g + (size_t)b * H * D + i);
    __syncthreads();
    if (pl < pstep) {
      for (int p = pl; p < hk * m; p += pstep) {
        const float* wr = w + (size_t)p * H;
        float dz = 0.f;
        for (int n = 0; n < H; ++n) dz += gs[n * D + d] * __ldg(wr + n);
        const int i = p / m, j = p % m;
        atomicAdd(dxks + i * D + d, dz * x0s[j * D + d]);
        atomicAdd(dx0s + j * D + d, dz * xks[i * D + d]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m * D; i += blockDim.x) dx0[(size_t)b * m * D + i] = dx0s[i];
    for (int i = threadIdx.x; i < hk * D; i += blockDim.x) dxk[(size_t)b * hk * D + i] = dxks[i];
  }
}

// ---- CUDA-core backward, filter gradient: dW[p,n] = sum_{b,d} xk[b,i,d]*x0[b,j,d]*g[b,n,d].
// grid (ceil(hk*m / PC), nsplit); block = threads over n; each CTA streams its share of the batch.
constexpr int CIN_PC = 16;
__global__ void __launch_bounds__(256)
cin_bwd_dw_kernel(const float* __restrict__ x0, const float* __restrict__ xk, const float* __restrict__ g, int B, int m,
                  int hk, int D, int H, float* __restrict__ dw) {
  extern __shared__ __align__(16) float sm[];
  float* gs = sm;                         // (H, D+1) padded against bank conflicts
  float* zs = gs + H * (D + 1);           // (PC, D)
  const int p0 = blockIdx.x * CIN_PC;
  const int K = hk * m;
  float acc[CIN_PC];
#pragma unroll
  for (int q = 0; q < CIN_PC; ++q) acc[q] = 0.f;
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    __syncthreads();
    for (int i = threadIdx.x; i < H * D; i += blockDim.x) gs[(i / D) * (D + 1) + i % D] = __ldg(g + (size_t)b * H * D + i);
    for (int i = threadIdx.x; i < CIN_PC * D; i += blockDim.x) {
      const int q = i / D, d = i % D, p = p0 + q;
      float z = 0.f;
      if (p < K) z = __ldg(xk + ((size_t)b * hk + p / m) * D + d) * __ldg(x0 + ((size_t)b * m + p % m) * D + d);
      zs[i] = z;
    }
    __syncthreads();
    for (int n = threadIdx.x; n < H; n += blockDim.x) {     // H <= blockDim in practice: one n per thread
      for (int d = 0; d < D; ++d) {
        const float gv = gs[n * (D + 1) + d];
#pragma unroll
        for (int q = 0; q < CIN_PC; ++q) acc[q] += zs[q * D + d] * gv;
      }
    }
  }
  const int n = threadIdx.x;
  if (n < H) {
#pragma unroll
    for (int q = 0; q < CIN_PC; ++q)
      if (p0 + q < K) atomicAdd(dw + (size_t)(p0 + q) * H + n, acc[q]);
  }
}

// ------------------------------------------------------------------------------------------------ host
static bool tensor_path_ok(int64_t m, int64_t hk, int64_t D, int64_t H) {
  return m >= 1 && m <= KB && hk >= 1 && H >= 1 && H <= 128 && D >= 1 && D <= 32 && (D & (D - 1)) == 0;
}
static int64_t pad16(int64_t h) { return (h + 15) / 16 * 16; }

}  // namespace cin
}  // namespace ctr

using namespace ctr;
using namespace ctr::cin;

extern "C" int64_t ctr_cin_fwd_workspace_bytes(int64_t B, int64_t m, int64_t hk, int64_t D, int64_t H) {
  (void)B;
  if (!tensor_path_ok(m, hk, D, H)) return 0;
  return 2 * pad16(H) * hk * KB * (int64_t)sizeof(float);
}

static int check_cin(const char* fn, int64_t B, int64_t m, int64_t hk, int64_t D, int64_t H) {
  CTR_REQUIRE(B >= 0 && m >= 1 && hk >= 1 && D >= 1 && H >= 1, "%s: bad sizes B=%lld m=%lld hk=%lld D=%lld H=%lld", fn,
              (long long)B, (long long)m, (long long)hk, (long long)D, (long long)H);
  CTR_UNSUPPORTED(B * D > 0x7fffffffLL || hk * m > (1 << 24), "%s: problem too large", fn);
  return CTR_OK;
}

extern "C" int ctr_cin_fwd(const float* x0, const float* xk, const float* filter, int64_t B, int64_t m, int64_t hk,
                           int64_t D, int64_t H, float* out, float* pooled, int precision, void* workspace,
                           int64_t workspace_bytes, void* stream) {
  int rc = check_cin("ctr_cin_fwd", B, m, hk, D, H);
  if (rc) return rc;
  CTR_REQUIRE(x0 && xk && filter && out, "ctr_cin_fwd: null argument");
  CTR_REQUIRE(precision == 0 || precision == 1, "ctr_cin_fwd: precision must be 0 (3xTF32) or 1 (TF32)");
  if (B == 0) return CTR_OK;
  cudaStream_t st = as_stream(stream);
  if (!tensor_path_ok(m, hk, D, H)) {
    const size_t smem = sizeof(float) * (size_t)(m + hk) * D;
    CTR_UNSUPPORTED(smem > 200 * 1024, "ctr_cin_fwd: (m+hk)*D too large for the CUDA-core path");
    if (smem > 48 * 1024)
      CTR_CUDA(cudaFuncSetAttribute(cin_fwd_simple_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)(B < (int64_t)sm_count() * 4 ? B : (int64_t)sm_count() * 4);
    cin_fwd_simple_kernel<<<grid, 256, smem, st>>>(x0, xk, filter, out, pooled, (int)B, (int)m, (int)hk, (int)D, (int)H);
    CTR_CHECK_LAUNCH("ctr_cin_fwd(simple)");
    return CTR_OK;
  }
  const int NP = (int)pad16(H);
  const int64_t need = ctr_cin_fwd_workspace_bytes(B, m, hk, D, H);
  CTR_REQUIRE(workspace != nullptr && workspace_bytes >= need,
              "ctr_cin_fwd: workspace of %lld bytes required (ctr_cin_fwd_workspace_bytes), got %lld", (long long)need,
              (long long)workspace_bytes);
  CTR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 127) == 0, "ctr_cin_fwd: workspace must be 128-byte aligned");
  float* wt = static_cast<float*>(workspace);
  const int KP = (int)hk * KB;
  {
    const long long total = (long long)NP * KP;
    const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    cin_split_filter_kernel<<<grid, 256, 0, st>>>(filter, wt, (int)m, (int)hk, (int)H, NP);
    CTR_CHECK_LAUNCH("ctr_cin_fwd(split filter)");
  }
  EncodeTiledFn enc = encode_tiled();
  if (enc == nullptr) {
    set_error("ctr_cin_fwd: cuTensorMapEncodeTiled is not available from the driver");
    return CTR_ERR_CUDA;
  }
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)KP, (cuuint64_t)(2 * NP)};
  const cuuint64_t gstride[1] = {(cuuint64_t)KP * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)NP};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, wt, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("ctr_cin_fwd: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
    return CTR_ERR_CUDA;
  }
  int logD = 0;
  while ((1 << logD) < D) ++logD;
  const long long rows_total = (long long)B * D;
  const int num_pairs = (int)((rows_total + TILES * BM - 1) / (TILES * BM));
  const int grid = num_pairs < sm_count() ? num_pairs : sm_count();
  // K-blocks per TMEM accumulation chain of the 3xTF32 path: 12 MMAs each.  The fp32 accumulate of the tensor core truncates
  // (~2.6e-8 relative per MMA, measured 3.8e-5 after 1440 MMAs).  8 blocks = 96 MMAs.  16 blocks (192 MMAs, half the drains)
  // was measured in round 2: 2.94 -> 2.88 ms per config-3 step, but the element-wise error reaches 0.9-1.5x the 1e-5 bound
  // (profiles/r2_cin_chunk16.jsonl) -- not worth 2 %.
  constexpr int CHUNK3 = 8;
#define CIN_LAUNCH(PASSES_, SB_, NPT_, CHUNK_)                                                                        \
  {                                                                                                                   \
    const FwdSmem L = fwd_smem(NP, PASSES_, SB_);                                                                     \
    auto k = cin_fwd_tc_kernel<PASSES_, SB_, NPT_>;                                                                   \
    CTR_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total + 1024));                   \
    k<<<grid, NTHREADS, L.total + 1024, st>>>(tmap, x0, xk, out, pooled, (int)B, (int)m, (int)hk, logD, (int)H, NP,   \
                                              CHUNK_);                                                                \
  }
  if (precision == 0) {
    if (NP <= 32) CIN_LAUNCH(3, 4, 32, CHUNK3) else if (NP <= 64) CIN_LAUNCH(3, 4, 64, CHUNK3) else CIN_LAUNCH(3, 4, 128, CHUNK3)
  } else {
    if (NP <= 32) CIN_LAUNCH(1, 6, 32, 32) else if (NP <= 64) CIN_LAUNCH(1, 6, 64, 32) else CIN_LAUNCH(1, 6, 128, 32)
  }
#undef CIN_LAUNCH
  CTR_CHECK_LAUNCH("ctr_cin_fwd(tcgen05)");
  return CTR_OK;
}

extern "C" int ctr_cin_bwd_tc_supported(int64_t m, int64_t hk, int64_t D, int64_t H);                 // cin_bwd.cu
int ctr_cin_bwd_tc(const float* x0, const float* xk, const float* filter, const float* g_out, int64_t B, int64_t m,
                   int64_t hk, int64_t D, int64_t H, float* dx0, float* dxk, float* dfilter, void* workspace,
                   cudaStream_t st);                                                                          // cin_bwd.cu

extern "C" int ctr_cin_bwd(const float* x0, const float* xk, const float* filter, const float* g_out, int64_t B,
                           int64_t m, int64_t hk, int64_t D, int64_t H, float* dx0, float* dxk, float* dfilter,
                           void* workspace, int64_t workspace_bytes, void* stream) {
  int rc = check_cin("ctr_cin_bwd", B, m, hk, D, H);
  if (rc) return rc;
  CTR_REQUIRE(x0 && xk && filter && g_out && dx0 && dxk && dfilter, "ctr_cin_bwd: null argument");
  CTR_UNSUPPORTED(D > 256 || H > 256, "ctr_cin_bwd: D=%lld H=%lld too large", (long long)D, (long long)H);
  cudaStream_t st = as_stream(stream);
  CTR_CUDA(cudaMemsetAsync(dfilter, 0, sizeof(float) * hk * m * H, st));
  if (B == 0) return CTR_OK;
  if (ctr_cin_bwd_tc_supported(m, hk, D, H)) {
    const int64_t need = ctr_cin_bwd_workspace_bytes(B, m, hk, D, H);
    CTR_REQUIRE(workspace != nullptr && workspace_bytes >= need,
                "ctr_cin_bwd: workspace of %lld bytes required (ctr_cin_bwd_workspace_bytes), got %lld", (long long)need,
                (long long)workspace_bytes);
    CTR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 127) == 0, "ctr_cin_bwd: workspace must be 128-byte aligned");
    return ctr_cin_bwd_tc(x0, xk, filter, g_out, B, m, hk, D, H, dx0, dxk, dfilter, workspace, st);
  }
  {
    const size_t smem = sizeof(float) * (size_t)(2 * (m + hk) + H) * D;
    CTR_UNSUPPORTED(smem > 200 * 1024, "ctr_cin_bwd: shared memory need %zu B too large", smem);
    if (smem > 48 * 1024)
      CTR_CUDA(cudaFuncSetAttribute(cin_bwd_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)(B < (int64_t)sm_count() * 4 ? B : (int64_t)sm_count() * 4);
    cin_bwd_dx_kernel<<<grid, 256, smem, st>>>(x0, xk, filter, g_out, (int)B, (int)m, (int)hk, (int)D, (int)H, dx0, dxk);
    CTR_CHECK_LAUNCH("ctr_cin_bwd(dx)");
  }
  {
    const size_t smem = sizeof(float) * (size_t)(H * (D + 1) + CIN_PC * D);
    CTR_UNSUPPORTED(smem > 200 * 1024, "ctr_cin_bwd: shared memory need %zu B too large", smem);
    if (smem > 48 * 1024)
      CTR_CUDA(cudaFuncSetAttribute(cin_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int gx = (int)((hk * m + CIN_PC - 1) / CIN_PC);
    int gy = (int)((int64_t)sm_count() * 4 / gx);
    if (gy < 1) gy = 1;
    if (gy > B) gy = (int)B;
    cin_bwd_dw_kernel<<<dim3(gx, gy), 256, smem, st>>>(x0, xk, g_out, (int)B, (int)m, (int)hk, (int)D, (int)H, dfilter);
    CTR_CHECK_LAUNCH("ctr_cin_bwd(dw)");
  }
  return CTR_OK;
}
