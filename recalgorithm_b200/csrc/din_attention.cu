// Row DIN-ATT: the DIN attention unit over the behaviour sequence.
//
// Reference: din_attention(query, keys, keys_length, is_softmax) -- DIN/din_attention.py:17-43
//   cross = [q, k, q-k, q*k] (4H) -> dense 64 relu (f1_att) -> dense 32 relu (f2_att) -> dense 1 (f3_att);
//   mask by sequence_mask(keys_length, T); default: weights*mask; softmax variant: fill -2**32+1, THEN /sqrt(H),
//   softmax over T; out = weights^T . keys  (B,H).
//
// B200 mapping (FP32-FMA-bound, CUDA cores; north_star keeps tensor cores for CIN only):
//   * one warp per sample; the sample's keys (T x H) are staged once in shared memory;
//   * layer 1 is algebraically folded per sample:  cross@W1 = q@(W1a+W1c) + k@(W1b-W1c) + (q*k)@W1d
//       = qpart + k @ Weff,  Weff[h,c] = (W1b-W1c)[h,c] + q[h]*W1d[h,c]   (H x 64 per sample, in registers)
//     so the q side is hoisted out of the T loop and layer 1 costs H*64 instead of 4H*64 MACs per position;
//   * lane c owns output column c of each dense layer; activations of the previous layer are broadcast from
//     shared memory with 128-bit loads; W2's column lives in 64 registers;
//   * positions t >= keys_length never reach the MLP: their weight is exactly 0 (or the constant pad) in the
//     reference, so the result is identical and ~half the FLOPs disappear for uniformly distributed lengths;
//   * backward recomputes the MLP per position (nothing but the (B,T) weights is saved), produces input
//     gradients per warp, and reduces the weight gradients CTA-cooperatively with thread-owned register
//     accumulators (rank-k updates over the positions staged by all warps), merged by fp32 atomics.
#include <math.h>

#include "ctr_common.cuh"

namespace ctr {

constexpr int DIN_H1 = 64;      // f1_att units (DIN/din_attention.py:21)
constexpr int DIN_H2 = 32;      // f2_att units (:22)
constexpr int DIN_TT = 4;       // positions per warp per round
constexpr float DIN_PAD_F = -4294967295.0f;   // -2**32 + 1 (rounds to -2^32 in fp32, like the reference's fp32 tensor)

struct DinSmem {                // offsets in floats into dynamic shared memory
  int wq, wk, wd, b1, w2, b2, w3, w1t, per_warp, warp_stride, tile, tile_stride, total;
};
// offsets inside one staged position of the backward tile
__host__ __device__ inline int din_off_cross() { return DIN_H1; }
__host__ __device__ inline int din_off_dpre1(int H) { return DIN_H1 + 4 * H; }
__host__ __device__ inline int din_off_h2(int H) { return 2 * DIN_H1 + 4 * H; }
__host__ __device__ inline int din_off_dpre2(int H) { return 2 * DIN_H1 + DIN_H2 + 4 * H; }
__host__ __device__ inline int din_off_dcross(int H) { return 2 * DIN_H1 + 2 * DIN_H2 + 4 * H; }
__host__ __device__ inline int din_off_ds(int H) { return 2 * DIN_H1 + 2 * DIN_H2 + 8 * H; }

__host__ __device__ inline DinSmem din_layout(int H, int T, int warps, bool bwd) {
  DinSmem L;
  int o = 0;
  L.wq = o; o += H * DIN_H1;                      // (H,64)  W1a + W1c
  L.wk = o; o += H * DIN_H1;                      // (H,64)  W1b - W1c
  L.wd = o; o += H * DIN_H1;                      // (H,64)  W1d
  L.b1 = o; o += DIN_H1;
  L.w2 = o; o += DIN_H1 * DIN_H2;                 // (64,32) row-major (bwd reads rows; fwd loads columns to registers)
  L.b2 = o; o += DIN_H2;
  L.w3 = o; o += DIN_H2 + 4;                      // w3 (32) + b3 (1), padded
  L.w1t = o; o += bwd ? DIN_H1 * 4 * H : 0;       // (64, 4H) = W1 transposed, for dcross = dpre1 @ W1^T
  L.per_warp = o;
  // per warp: keys (T*H) | sc (T) | q (H) | go (H) | dq (H)  (+ bwd: dkeys (T*H), ds (T))
  int ws = T * H + T + 3 * H + (bwd ? T * H + T : 0);
  ws = (ws + 3) & ~3;
  L.warp_stride = ws;
  o += warps * ws;
  L.tile = o;
  // per warp, per staged position: h1 (64) | [bwd: cross (4H) | dpre1 (64) | h2 (32) | dpre2 (32) | dcross (4H) | ds (4)]
  int ts = DIN_H1 + (bwd ? 4 * H + DIN_H1 + DIN_H2 + DIN_H2 + 4 * H + 4 : 0);
  L.tile_stride = ts;
  o += warps * DIN_TT * ts;
  L.total = o;
  return L;
}

__device__ __forceinline__ void din_stage_weights(float* sm, const DinSmem& L, const float* __restrict__ w1,
                                                  const float* __restrict__ b1, const float* __restrict__ w2,
                                                  const float* __restrict__ b2, const float* __restrict__ w3,
                                                  const float* __restrict__ b3, int H, bool bwd) {
  for (int i = threadIdx.x; i < H * DIN_H1; i += blockDim.x) {
    const float a = __ldg(w1 + i), b = __ldg(w1 + H * DIN_H1 + i), c = __ldg(w1 + 2 * H * DIN_H1 + i),
                d = __ldg(w1 + 3 * H * DIN_H1 + i);
    sm[L.wq + i] = a + c;
    sm[L.wk + i] = b - c;
    sm[L.wd + i] = d;
  }
  for (int i = threadIdx.x; i < DIN_H1; i += blockDim.x) sm[L.b1 + i] = __ldg(b1 + i);
  for (int i = threadIdx.x; i < DIN_H1 * DIN_H2; i += blockDim.x) sm[L.w2 + i] = __ldg(w2 + i);
  for (int i = threadIdx.x; i < DIN_H2; i += blockDim.x) { sm[L.b2 + i] = __ldg(b2 + i); sm[L.w3 + i] = __ldg(w3 + i); }
  if (threadIdx.x == 0) sm[L.w3 + DIN_H2] = __ldg(b3);
  if (bwd)
    for (int i = threadIdx.x; i < 4 * H * DIN_H1; i += blockDim.x) {
      const int r = i / DIN_H1, c = i % DIN_H1;                     // w1[r][c] -> w1t[c][r]
      sm[L.w1t + c * 4 * H + r] = __ldg(w1 + i);
    }
}

// Per-sample preparation shared by fwd and bwd: stage keys + q, fold layer 1.
template <int HP>
__device__ __forceinline__ void din_prepare(float* sm, const DinSmem& L, float* wsm, const float* __restrict__ query,
                                            const float* __restrict__ keys, int b, int T, int H, int lane,
                                            float (&weff)[HP][2], float (&qpart)[2]) {
  float* skeys = wsm;
  float* sq = wsm + T * H + T;
  for (int i = lane; i < T * H; i += 32) skeys[i] = __ldg(keys + (size_t)b * T * H + i);
  for (int i = lane; i < H; i += 32) sq[i] = __ldg(query + (size_t)b * H + i);
  __syncwarp();
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int c = lane + 32 * u;
    float acc = sm[L.b1 + c];
#pragma unroll
    for (int h = 0; h < HP; ++h) {
      if (h < H) {
        const float qh = sq[h];
        acc += qh * sm[L.wq + h * DIN_H1 + c];
        weff[h][u] = sm[L.wk + h * DIN_H1 + c] + qh * sm[L.wd + h * DIN_H1 + c];
      } else {
        weff[h][u] = 0.f;
      }
    }
    qpart[u] = acc;
  }
}

// Forward MLP for up to DIN_TT positions t0..t0+n-1 of the staged sample.  Writes h1 to the warp's tile area and
// returns (lane-owned) h2 and the scores.
template <int HP>
__device__ __forceinline__ void din_mlp_tile(const float* sm, const DinSmem& L, const float* skeys, float* tile,
                                             int t0, int n, int H, int lane, const float (&weff)[HP][2],
                                             const float (&qpart)[2], const float (&w2col)[DIN_H1],
                                             float (&h2)[DIN_TT], float (&score)[DIN_TT]) {
#pragma unroll
  for (int tt = 0; tt < DIN_TT; ++tt) {
    if (tt < n) {
      const float* k = skeys + (t0 + tt) * H;
      float p0 = qpart[0], p1 = qpart[1];
#pragma unroll
      for (int h = 0; h < HP; ++h) {
        if (h < H) {
          const float kv = k[h];
          p0 += kv * weff[h][0];
          p1 += kv * weff[h][1];
        }
      }
      tile[tt * L.tile_stride + lane] = fmaxf(p0, 0.f);
      tile[tt * L.tile_stride + lane + 32] = fmaxf(p1, 0.f);
    }
  }
  __syncwarp();
  const float b2v = sm[L.b2 + lane], w3v = sm[L.w3 + lane], b3v = sm[L.w3 + DIN_H2];
#pragma unroll
  for (int tt = 0; tt < DIN_TT; ++tt) {
    h2[tt] = 0.f;
    score[tt] = 0.f;
    if (tt < n) {
      const float4* h1v = reinterpret_cast<const float4*>(tile + tt * L.tile_stride);
      float acc = b2v;
#pragma unroll
      for (int c4 = 0; c4 < DIN_H1 / 4; ++c4) {
        const float4 v = h1v[c4];
        acc += v.x * w2col[4 * c4 + 0];
        acc += v.y * w2col[4 * c4 + 1];
        acc += v.z * w2col[4 * c4 + 2];
        acc += v.w * w2col[4 * c4 + 3];
      }
      h2[tt] = fmaxf(acc, 0.f);
      score[tt] = warp_sum(h2[tt] * w3v) + b3v;
    }
  }
}

// scores (smem, first `len` valid) -> attention weights in place, following DIN/din_attention.py:27-38.
__device__ __forceinline__ void din_weights(float* sc, int T, int len, int H, int is_softmax, int lane) {
  if (!is_softmax) {
    for (int t = lane; t < T; t += 32) sc[t] = t < len ? sc[t] : 0.f;        // weights * mask
  } else {
    const float scale = sqrtf((float)H);
    float m = -INFINITY;
    for (int t = lane; t < T; t += 32) {
      const float v = (t < len ? sc[t] : DIN_PAD_F) / scale;                 // where(mask, w, pad) THEN / sqrt(H)
      sc[t] = v;
      m = fmaxf(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int t = lane; t < T; t += 32) {
      const float e = expf(sc[t] - m);
      sc[t] = e;
      s += e;
    }
    s = warp_sum(s);
    for (int t = lane; t < T; t += 32) sc[t] = sc[t] / s;
  }
  __syncwarp();
}

template <int HP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
din_attention_fwd_kernel(const float* __restrict__ query, const float* __restrict__ keys,
                         const long long* __restrict__ keys_length, const float* __restrict__ w1,
                         const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                         const float* __restrict__ w3, const float* __restrict__ b3, int B, int T, int H,
                         int is_softmax, float* __restrict__ out, float* __restrict__ att_w) {
  extern __shared__ __align__(16) float sm[];
  const DinSmem L = din_layout(H, T, WARPS, false);
  din_stage_weights(sm, L, w1, b1, w2, b2, w3, b3, H, false);
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float w2col[DIN_H1];
#pragma unroll
  for (int c = 0; c < DIN_H1; ++c) w2col[c] = sm[L.w2 + c * DIN_H2 + lane];
  float* wsm = sm + L.per_warp + wid * L.warp_stride;
  float* skeys = wsm;
  float* sc = wsm + T * H;
  float* tile = sm + L.tile + wid * DIN_TT * L.tile_stride;

  for (int b = blockIdx.x * WARPS + wid; b < B; b += gridDim.x * WARPS) {
    float weff[HP][2], qpart[2];
    __syncwarp();
    din_prepare<HP>(sm, L, wsm, query, keys, b, T, H, lane, weff, qpart);
    long long len64 = __ldg(keys_length + b);
    const int len = (int)(len64 < 0 ? 0 : (len64 > T ? T : len64));
    for (int t0 = 0; t0 < len; t0 += DIN_TT) {
      float h2[DIN_TT], score[DIN_TT];
      const int n = min(DIN_TT, len - t0);
      din_mlp_tile<HP>(sm, L, skeys, tile, t0, n, H, lane, weff, qpart, w2col, h2, score);
      if (lane == 0) {
#pragma unroll
        for (int tt = 0; tt < DIN_TT; ++tt)
          if (tt < n) sc[t0 + tt] = score[tt];
      }
      __syncwarp();
    }
    din_weights(sc, T, len, H, is_softmax, lane);
    for (int h = lane; h < H; h += 32) {
      float acc = 0.f;
      for (int t = 0; t < T; ++t) acc += sc[t] * skeys[t * H + h];           // matmul(w^T, keys)
      out[(size_t)b * H + h] = acc;
    }
    if (att_w != nullptr)
      for (int t = lane; t < T; t += 32) att_w[(size_t)b * T + t] = sc[t];
  }
}

// ---------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------
// d_params layout: [w1 (4H*64) | b1 (64) | w2 (64*32) | b2 (32) | w3 (32) | b3 (1)]
template <int HP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
din_attention_bwd_kernel(const float* __restrict__ query, const float* __restrict__ keys,
                         const long long* __restrict__ keys_length, const float* __restrict__ w1,
                         const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                         const float* __restrict__ w3, const float* __restrict__ b3, const float* __restrict__ g_out,
                         int B, int T, int H, int is_softmax, float* __restrict__ d_query,
                         float* __restrict__ d_keys, float* __restrict__ d_params) {
  extern __shared__ __align__(16) float sm[];
  constexpr int NT = WARPS * 32;
  const DinSmem L = din_layout(H, T, WARPS, true);
  din_stage_weights(sm, L, w1, b1, w2, b2, w3, b3, H, true);
  __shared__ int s_npos[WARPS];                 // positions staged by each warp this round
  __shared__ int s_more[2];                     // any warp still has work (double-buffered by round parity)
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, tid = threadIdx.x;
  const int H4 = 4 * H;
  float w2col[DIN_H1];
#pragma unroll
  for (int c = 0; c < DIN_H1; ++c) w2col[c] = sm[L.w2 + c * DIN_H2 + lane];
  float* wsm = sm + L.per_warp + wid * L.warp_stride;
  float* skeys = wsm;
  float* sc = wsm + T * H;                      // attention weights w[t]
  float* sq = sc + T;
  float* sgo = sq + H;
  float* sdq = sgo + H;
  float* sdk = sdq + H;                         // (T,H) dkeys accumulator
  float* sds = sdk + T * H;                     // (T) ds
  float* tile = sm + L.tile + wid * DIN_TT * L.tile_stride;

  // thread-owned weight-gradient accumulators (rank-k updates in phase B)
  //   dW1 (4H x 64): thread owns column c1 = tid % 64 and the R1 CONSECUTIVE rows starting at (tid / 64) * R1
  //   dW2 (64 x 32): thread owns column c2 = tid % 32 and the R2 consecutive rows starting at (tid / 32) * R2
  //   (consecutive rows -> the per-position row vectors are read with 128-bit shared loads)
  constexpr int R1 = (4 * HP * DIN_H1) / NT;              // rows of dW1 per thread (= HP at 256 threads; multiple of 4)
  static_assert(R1 % 4 == 0 && (4 * HP * DIN_H1) % NT == 0, "row blocks must be float4-sized");
  constexpr int R2 = (DIN_H1 * DIN_H2) / NT;               // rows of dW2 per thread
  float acc_w1[R1], acc_w2[R2], acc_b1 = 0.f, acc_b2 = 0.f, acc_w3 = 0.f, acc_b3 = 0.f;
#pragma unroll
  for (int i = 0; i < R1; ++i) acc_w1[i] = 0.f;
#pragma unroll
  for (int i = 0; i < R2; ++i) acc_w2[i] = 0.f;

  int b = blockIdx.x * WARPS + wid;             // current sample of this warp
  int t_next = 0, len = 0, round = 0;
  bool loaded = false;
  float weff[HP][2], qpart[2];

  while (true) {
    // ------------------------------------------------ phase A: each warp stages up to DIN_TT positions
    int n = 0;
    while (b < B && n == 0) {
      if (!loaded) {
        __syncwarp();
        din_prepare<HP>(sm, L, wsm, query, keys, b, T, H, lane, weff, qpart);
        for (int i = lane; i < H; i += 32) { sgo[i] = __ldg(g_out + (size_t)b * H + i); sdq[i] = 0.f; }
        long long len64 = __ldg(keys_length + b);
        len = (int)(len64 < 0 ? 0 : (len64 > T ? T : len64));
        // recompute the forward scores -> attention weights (nothing but inputs is read from HBM)
        for (int t0 = 0; t0 < len; t0 += DIN_TT) {
          float h2[DIN_TT], score[DIN_TT];
          const int nn = min(DIN_TT, len - t0);
          din_mlp_tile<HP>(sm, L, skeys, tile, t0, nn, H, lane, weff, qpart, w2col, h2, score);
          if (lane == 0) {
#pragma unroll
            for (int tt = 0; tt < DIN_TT; ++tt)
              if (tt < nn) sc[t0 + tt] = score[tt];
          }
          __syncwarp();
        }
        din_weights(sc, T, len, H, is_softmax, lane);
        // dw[t] = go . k[t];  dkeys[t,:] = w[t]*go;  ds from the mask / softmax backward
        float dot = 0.f;
        for (int t = lane; t < T; t += 32) {
          float dwv = 0.f;
          for (int h = 0; h < H; ++h) dwv += sgo[h] * skeys[t * H + h];
          sds[t] = dwv;
          dot += sc[t] * dwv;
        }
        dot = warp_sum(dot);
        const float scale = sqrtf((float)H);
        for (int t = lane; t < T; t += 32) {
          const float dwv = sds[t];
          float dsv;
          if (is_softmax) dsv = t < len ? sc[t] * (dwv - dot) / scale : 0.f;
          else dsv = t < len ? dwv : 0.f;
          sds[t] = dsv;
        }
        for (int i = lane; i < T * H; i += 32) sdk[i] = sc[i / H] * sgo[i % H];
        __syncwarp();
        loaded = true;
        t_next = 0;
      }
      if (t_next < len) {
        n = min(DIN_TT, len - t_next);
      } else {
        // sample finished: flush its input gradients
        for (int i = lane; i < T * H; i += 32) d_keys[(size_t)b * T * H + i] = sdk[i];
        for (int i = lane; i < H; i += 32) d_query[(size_t)b * H + i] = sdq[i];
        __syncwarp();
        b += gridDim.x * WARPS;
        loaded = false;
      }
    }
    if (n > 0) {
      const int t0 = t_next;
      float h2[DIN_TT], score[DIN_TT];
      din_mlp_tile<HP>(sm, L, skeys, tile, t0, n, H, lane, weff, qpart, w2col, h2, score);
      // stage cross, h2, dpre2, ds
#pragma unroll
      for (int tt = 0; tt < DIN_TT; ++tt) {
        if (tt < n) {
          float* tp = tile + tt * L.tile_stride;
          float* cross = tp + din_off_cross();
          const float* k = skeys + (t0 + tt) * H;
          for (int h = lane; h < H; h += 32) {
            const float qv = sq[h], kv = k[h];
            cross[h] = qv; cross[H + h] = kv; cross[2 * H + h] = qv - kv; cross[3 * H + h] = qv * kv;
          }
          const float dsv = sds[t0 + tt];
          tp[din_off_h2(H) + lane] = h2[tt];
          tp[din_off_dpre2(H) + lane] = h2[tt] > 0.f ? dsv * sm[L.w3 + lane] : 0.f;
          if (lane == 0) tp[din_off_ds(H)] = dsv;
        }
      }
      __syncwarp();
      // dh1[c] = sum_c2 dpre2[c2] * W2[c][c2] for c = lane, lane+32 -> dpre1 = dh1 * (h1 > 0).
      // W2 rows sit 128 B apart in smem; every lane walks the eight 16-byte chunks of its rows in a rotated order
      // so that a quarter-warp always touches eight different chunks (conflict-free LDS.128).
#pragma unroll
      for (int tt = 0; tt < DIN_TT; ++tt) {
        if (tt < n) {
          float* tp = tile + tt * L.tile_stride;
          const float4* dp2 = reinterpret_cast<const float4*>(tp + din_off_dpre2(H));
          const float4* r0 = reinterpret_cast<const float4*>(sm + L.w2 + lane * DIN_H2);
          const float4* r1 = reinterpret_cast<const float4*>(sm + L.w2 + (lane + 32) * DIN_H2);
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int c4 = 0; c4 < DIN_H2 / 4; ++c4) {
            const int ch = (c4 + lane) & (DIN_H2 / 4 - 1);
            const float4 dd = dp2[ch], x0 = r0[ch], x1 = r1[ch];
            a0 += dd.x * x0.x + dd.y * x0.y + dd.z * x0.z + dd.w * x0.w;
            a1 += dd.x * x1.x + dd.y * x1.y + dd.z * x1.z + dd.w * x1.w;
          }
          tp[din_off_dpre1(H) + lane] = tp[lane] > 0.f ? a0 : 0.f;
          tp[din_off_dpre1(H) + lane + 32] = tp[lane + 32] > 0.f ? a1 : 0.f;
        }
      }
      __syncwarp();
      // dcross[r] = sum_c dpre1[c] * W1[r][c]  (W1^T[c][r] in smem: lanes read consecutive r)
#pragma unroll
      for (int tt = 0; tt < DIN_TT; ++tt) {
        if (tt < n) {
          float* tp = tile + tt * L.tile_stride;
          const float* dpre1 = tp + din_off_dpre1(H);
          for (int r = lane; r < H4; r += 32) {
            float a = 0.f;
#pragma unroll 8
            for (int c = 0; c < DIN_H1; ++c) a += dpre1[c] * sm[L.w1t + c * H4 + r];
            tp[din_off_dcross(H) + r] = a;
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int tt = 0; tt < DIN_TT; ++tt) {
        if (tt < n) {
          const float* dcross = tile + tt * L.tile_stride + din_off_dcross(H);
          const float* k = skeys + (t0 + tt) * H;
          for (int h = lane; h < H; h += 32) {
            const float da = dcross[h], db_ = dcross[H + h], dc = dcross[2 * H + h], dd = dcross[3 * H + h];
            sdq[h] += da + dc + dd * k[h];
            sdk[(t0 + tt) * H + h] += db_ - dc + dd * sq[h];
          }
        }
      }
      t_next += n;
    }
    const int par = round & 1;
    if (lane == 0) s_npos[wid] = n;
    if (tid == 0) s_more[par] = 0;
    __syncthreads();
    if (lane == 0 && (n > 0 || b < B)) s_more[par] = 1;      // every writer stores the same value
    // ------------------------------------------------ phase B: cooperative rank-k update of the weight grads
    for (int wv = 0; wv < WARPS; ++wv) {
      const int np = s_npos[wv];
      for (int tt = 0; tt < np; ++tt) {
        const float* tp = sm + L.tile + (wv * DIN_TT + tt) * L.tile_stride;
        const float* h1 = tp;
        const float* cross = tp + din_off_cross();
        const float* dpre1 = tp + din_off_dpre1(H);
        const float* sh2 = tp + din_off_h2(H);
        const float* dpre2 = tp + din_off_dpre2(H);
        {
          const float dv = dpre1[tid % DIN_H1];
          const int r0 = (tid / DIN_H1) * R1;
#pragma unroll
          for (int i4 = 0; i4 < R1; i4 += 4) {
            if (r0 + i4 < H4) {                               // H4 is a multiple of 4: whole float4 in range
              const float4 cv = *reinterpret_cast<const float4*>(cross + r0 + i4);
              acc_w1[i4 + 0] += cv.x * dv; acc_w1[i4 + 1] += cv.y * dv; acc_w1[i4 + 2] += cv.z * dv; acc_w1[i4 + 3] += cv.w * dv;
            }
          }
          if (tid < DIN_H1) acc_b1 += dpre1[tid];
        }
        {
          const float dv = dpre2[tid % DIN_H2];
          const int c0 = (tid / DIN_H2) * R2;
#pragma unroll
          for (int i4 = 0; i4 < R2; i4 += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(h1 + c0 + i4);
            acc_w2[i4 + 0] += hv.x * dv; acc_w2[i4 + 1] += hv.y * dv; acc_w2[i4 + 2] += hv.z * dv; acc_w2[i4 + 3] += hv.w * dv;
          }
          if (tid < DIN_H2) acc_b2 += dpre2[tid];
        }
        const float dsv = tp[din_off_ds(H)];
        if (tid < DIN_H2) acc_w3 += sh2[tid] * dsv;
        else if (tid == DIN_H2) acc_b3 += dsv;
      }
    }
    __syncthreads();
    if (!s_more[par]) break;
    ++round;
  }
  // merge this CTA's partial weight gradients
  float* dW1 = d_params;
  float* dB1 = dW1 + H4 * DIN_H1;
  float* dW2 = dB1 + DIN_H1;
  float* dB2 = dW2 + DIN_H1 * DIN_H2;
  float* dW3 = dB2 + DIN_H2;
  float* dB3 = dW3 + DIN_H2;
#pragma unroll
  for (int i = 0; i < R1; ++i) {
    const int r = (tid / DIN_H1) * R1 + i;
    if (r < H4) atomicAdd(dW1 + r * DIN_H1 + tid % DIN_H1, acc_w1[i]);
  }
#pragma unroll
  for (int i = 0; i < R2; ++i) atomicAdd(dW2 + ((tid / DIN_H2) * R2 + i) * DIN_H2 + tid % DIN_H2, acc_w2[i]);
  if (tid < DIN_H1) atomicAdd(dB1 + tid, acc_b1);
  if (tid < DIN_H2) { atomicAdd(dB2 + tid, acc_b2); atomicAdd(dW3 + tid, acc_w3); }
  if (tid == DIN_H2) atomicAdd(dB3, acc_b3);
}

}  // namespace ctr

using namespace ctr;

static int check_din(const char* fn, int64_t B, int64_t T, int64_t H) {
  CTR_REQUIRE(B >= 0 && T >= 0 && H >= 1, "%s: bad sizes B=%lld T=%lld H=%lld", fn, (long long)B, (long long)T,
              (long long)H);
  CTR_UNSUPPORTED(H > 32, "%s: H=%lld > 32 unsupported", fn, (long long)H);
  return CTR_OK;
}

template <int WARPS>
static int din_fwd_launch(const float* query, const float* keys, const int64_t* len, const float* w1, const float* b1,
                          const float* w2, const float* b2, const float* w3, const float* b3, int64_t B, int64_t T,
                          int64_t H, int is_softmax, float* out, float* att_w, cudaStream_t st) {
  const DinSmem L = din_layout((int)H, (int)T, WARPS, false);
  const size_t smem = sizeof(float) * (size_t)L.total;
  CTR_UNSUPPORTED(smem > 220 * 1024, "ctr_din_attention_fwd: T=%lld H=%lld needs %zu B of shared memory", (long long)T,
                  (long long)H, smem);
  const long long need = (B + WARPS - 1) / WARPS;
#define GO(HPV)                                                                                                   \
  {                                                                                                               \
    auto k = din_attention_fwd_kernel<HPV, WARPS>;                                                                \
    if (smem > 48 * 1024) CTR_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    int per_sm = 1;                                                                                               \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, WARPS * 32, smem);                                  \
    long long grid = (long long)(per_sm < 1 ? 1 : per_sm) * sm_count();                                           \
    if (grid > need) grid = need;                                                                                 \
    k<<<(int)grid, WARPS * 32, smem, st>>>(query, keys, reinterpret_cast<const long long*>(len), w1, b1, w2, b2,  \
                                           w3, b3, (int)B, (int)T, (int)H, is_softmax, out, att_w);               \
  }
  if (H <= 4) GO(4) else if (H <= 8) GO(8) else if (H <= 16) GO(16) else GO(32)
#undef GO
  CTR_CHECK_LAUNCH("ctr_din_attention_fwd");
  return CTR_OK;
}

extern "C" int ctr_din_attention_fwd(const float* query, const float* keys, const int64_t* keys_length, const float* w1,
                                     const float* b1, const float* w2, const float* b2, const float* w3,
                                     const float* b3, int64_t B, int64_t T, int64_t H, int is_softmax, float* out,
                                     float* att_w, void* stream) {
  int rc = check_din("ctr_din_attention_fwd", B, T, H);
  if (rc) return rc;
  CTR_REQUIRE(query && keys_length && w1 && b1 && w2 && b2 && w3 && b3 && out && (keys || T == 0),
              "ctr_din_attention_fwd: null argument");
  if (B == 0) return CTR_OK;
  cudaStream_t st = as_stream(stream);
  if (T == 0) {                                           // empty history: the weighted sum over no keys is 0
    CTR_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * B * H, st));
    return CTR_OK;
  }
  return din_fwd_launch<4>(query, keys, keys_length, w1, b1, w2, b2, w3, b3, B, T, H, is_softmax, out, att_w, st);
}

extern "C" int ctr_din_attention_bwd(const float* query, const float* keys, const int64_t* keys_length, const float* w1,
                                     const float* b1, const float* w2, const float* b2, const float* w3,
                                     const float* b3, const float* g_out, int64_t B, int64_t T, int64_t H,
                                     int is_softmax, float* d_query, float* d_keys, float* d_params, void* stream) {
  int rc = check_din("ctr_din_attention_bwd", B, T, H);
  if (rc) return rc;
  CTR_REQUIRE(query && keys_length && w1 && b1 && w2 && b2 && w3 && b3 && g_out && d_query && d_params &&
                  ((keys && d_keys) || T == 0), "ctr_din_attention_bwd: null argument");
  cudaStream_t st = as_stream(stream);
  const int64_t nparams = 4 * H * DIN_H1 + DIN_H1 + DIN_H1 * DIN_H2 + DIN_H2 + DIN_H2 + 1;
  CTR_CUDA(cudaMemsetAsync(d_params, 0, sizeof(float) * nparams, st));
  if (B == 0) return CTR_OK;
  if (T == 0) {
    CTR_CUDA(cudaMemsetAsync(d_query, 0, sizeof(float) * B * H, st));
    return CTR_OK;
  }
  constexpr int WARPS = 8;
  const DinSmem L = din_layout((int)H, (int)T, WARPS, true);
  const size_t smem = sizeof(float) * (size_t)L.total;
  CTR_UNSUPPORTED(smem > 220 * 1024, "ctr_din_attention_bwd: T=%lld H=%lld needs %zu B of shared memory", (long long)T,
                  (long long)H, smem);
  const long long need = (B + WARPS - 1) / WARPS;
#define GO(HPV)                                                                                                   \
  {                                                                                                               \
    auto k = din_attention_bwd_kernel<HPV, WARPS>;                                                                \
    if (smem > 48 * 1024) CTR_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    int per_sm = 1;                                                                                               \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, WARPS * 32, smem);                                  \
    long long grid = (long long)(per_sm < 1 ? 1 : per_sm) * sm_count();                                           \
    if (grid > need) grid = need;                                                                                 \
    k<<<(int)grid, WARPS * 32, smem, st>>>(query, keys, reinterpret_cast<const long long*>(keys_length), w1, b1,  \
                                           w2, b2, w3, b3, g_out, (int)B, (int)T, (int)H, is_softmax, d_query,    \
                                           d_keys, d_params);                                                     \
  }
  if (H <= 4) GO(4) else if (H <= 8) GO(8) else if (H <= 16) GO(16) else GO(32)
#undef GO
  CTR_CHECK_LAUNCH("ctr_din_attention_bwd");
  return CTR_OK;
}
