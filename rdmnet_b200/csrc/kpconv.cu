// KPConv neighbour-gather kernels (the HBM/L2-bound part of the backbone) + strided max-pool + nearest upsample.
//
// Reference semantics: geotransformer/modules/kpconv/kpconv.py:79-122 (KPConv.forward),
// geotransformer/modules/kpconv/functional.py:6-22 (nearest_upsample), :54-67 (maxpool).
//
// KPConv is split in two: (1) kpconv_gather: A[m, k, c] = (1/cnt_m) * sum_h w[m,h,k] * F[idx[m,h], c]   (this file)
//                         (2) a dense GEMM  out = A.view(M, 15*C_in) @ W.view(15*C_in, C_out) + bias      (dense.cu)
// with w = max(0, 1 - |s[idx] - q - kp_k| / sigma) and cnt_m = max(1, #{h : sum_c F[idx[m,h], c] > 0}).
// The (M,H,K,3) / (M,H,C) temporaries the reference materialises in HBM never exist; per CTA the 15 influences of
// each neighbour are computed once into shared memory and re-used by all channel slices.
#include <stdlib.h>
#include "common.cuh"
#include "../../include/rdm_sm100.h"

#define KP_K 15

// flag[n] = (sum_c F[n,c] > 0)   (kpconv.py:113-114)
__global__ void row_positive_kernel(const float* __restrict__ f, int n, int c, unsigned char* __restrict__ flag) {
  pdl_trigger();
  pdl_wait();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  float s = 0.f;
  for (int i = lane; i < c; i += 32) s += f[(size_t)row * c + i];
  s = warp_sum(s);
  if (lane == 0) flag[row] = s > 0.f ? 1 : 0;
}

int rdm_row_positive(const float* f, int n, int c, unsigned char* flag, cudaStream_t stream) {
  if (n <= 0) return RDM_OK;
  row_positive_kernel<<<cdiv(n, 8), 256, 0, stream>>>(f, n, c, flag);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

struct KPts {
  float x[16], y[16], z[16];  // 15 kernel points + one far-away dummy (influence 0): pairs feed the packed f32x2 math
};

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---- packed fp32x2 helpers (sm_100: add / mul / fma .f32x2 take two lanes' worth of fp32 per instruction)
__device__ __forceinline__ unsigned long long pk2(float a, float b) {
  const float2 t = make_float2(a, b);
  return *reinterpret_cast<const unsigned long long*>(&t);
}
__device__ __forceinline__ float2 up2(unsigned long long v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// w[k] = max(0, 1 - |d - kp_k| / sigma), k = 0..14, w[15] = 0 (kpconv.py:98-99). Scalar on purpose: the packed form
// (add/mul/fma.rn.f32x2 over kernel-point pairs, 30 % fewer instructions) measured 7-15 % SLOWER on B200
// (gpurun s4h vs s4g: C_in = 32 layer 64.3 -> 70.4 us) - the f32x2 ops with constant-bank operands need register
// staging moves and run at half rate, and this phase is latency- not issue-bound.
__device__ __forceinline__ void influences16(float dx, float dy, float dz, const KPts& kp, float inv_sigma, float (&w)[16]) {
#pragma unroll
  for (int k = 0; k < KP_K; k++) {
    const float ex = dx - kp.x[k], ey = dy - kp.y[k], ez = dz - kp.z[k];
    const float d = sqrt_approx(fmaf(ez, ez, fmaf(ey, ey, ex * ex)));
    w[k] = fmaxf(0.f, fmaf(-d, inv_sigma, 1.f));
  }
  w[15] = 0.f;
}

// C_in == 1 (encoder1_1): 8 lanes per query (4 queries per warp), lanes stride over the neighbour list; the 15
// kernel points come by value in the constant bank; 3 shuffle rounds reduce the 15 partial sums of a query.
template <typename IdxT>
__global__ void __launch_bounds__(256) kpconv_gather_c1_kernel(const float* __restrict__ feats,
                                                               const float* __restrict__ q_pts,
                                                               const float* __restrict__ s_pts,
                                                               const IdxT* __restrict__ idx, const KPts kp,
                                                               float inv_sigma, int M, int N, int H,
                                                               const int* __restrict__ order, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, t = lane & 7;
  const int mi = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4 + (lane >> 3);
  const bool qvalid = mi < M;
  const int m = (qvalid && order != nullptr) ? order[mi] : mi;
  const int mm = qvalid ? m : M - 1;
  const float qx = q_pts[3 * (size_t)mm], qy = q_pts[3 * (size_t)mm + 1], qz = q_pts[3 * (size_t)mm + 2];
  float acc[KP_K];
#pragma unroll
  for (int k = 0; k < KP_K; k++) acc[k] = 0.f;
  int cnt = 0;
  // batches of 4 slots per lane: all index loads, then all point / feature loads, then the math - two dependent
  // memory round trips per batch instead of per slot
  for (int hb = t; hb < H; hb += 32) {
    long long jj[4];
    float px[4], py[4], pz[4], ff[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int h = hb + 8 * i;
      jj[i] = (qvalid && h < H) ? (long long)idx[(size_t)mm * H + h] : (long long)N;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const bool ok = jj[i] < N;
      const long long j = ok ? jj[i] : 0;
      px[i] = s_pts[3 * j];
      py[i] = s_pts[3 * j + 1];
      pz[i] = s_pts[3 * j + 2];
      ff[i] = ok ? feats[j] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (jj[i] >= N) continue;
      const float f = ff[i];
      cnt += f > 0.f;
      float w[16];
      influences16(px[i] - qx, py[i] - qy, pz[i] - qz, kp, inv_sigma, w);
#pragma unroll
      for (int k = 0; k < KP_K; k++) acc[k] = fmaf(w[k], f, acc[k]);
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(FULL_MASK, cnt, o);
#pragma unroll
    for (int k = 0; k < KP_K; k++) acc[k] += __shfl_xor_sync(FULL_MASK, acc[k], o);
  }
  if (!qvalid) return;
  const float inv = 1.f / (float)max(cnt, 1);
  float mine = 0.f;
#pragma unroll
  for (int k = 0; k < KP_K; k++)
    if (t == (k & 7)) {  // spread the 15 stores over the 8 lanes of the group
      if (k < 8) mine = acc[k];
    }
  if (t < 8) out[(size_t)m * KP_K + t] = mine * inv;
  float mine2 = 0.f;
#pragma unroll
  for (int k = 8; k < KP_K; k++)
    if (t == k - 8) mine2 = acc[k];
  if (t < KP_K - 8) out[(size_t)m * KP_K + 8 + t] = mine2 * inv;
}

// General case. CTA = 8 warps = QPC queries x NS channel slices (slice = 32*VEC channels).
// smem: w[QPC][H][16] floats, sidx[QPC][H] ints, nvalid[QPC], npos[QPC]
template <int VEC, typename IdxT>
__global__ void __launch_bounds__(256) kpconv_gather_kernel(const float* __restrict__ feats,
                                                            const unsigned char* __restrict__ rowpos,
                                                            const float* __restrict__ q_pts,
                                                            const float* __restrict__ s_pts,
                                                            const IdxT* __restrict__ idx, const float* __restrict__ kpts,
                                                            float sigma, int M, int N, int H, int C, int NS, int QPC,
                                                            float* __restrict__ out) {
  extern __shared__ __align__(16) float s_mem[];
  __shared__ float s_kp[KP_K * 3];
  float* s_w = s_mem;                             // QPC*H*16
  int* s_idx = (int*)(s_mem + (size_t)QPC * H * 16);  // QPC*H
  int* s_nvalid = s_idx + QPC * H;                // QPC
  int* s_npos = s_nvalid + QPC;                   // QPC
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * QPC;
  if (tid < KP_K * 3) s_kp[tid] = kpts[tid];
  if (tid < QPC) {
    s_nvalid[tid] = 0;
    s_npos[tid] = 0;
  }
  __syncthreads();
  // phase 1: influences of every (query, neighbour) of this CTA
  for (int e = tid; e < QPC * H; e += 256) {
    int qi = e / H, h = e - qi * H, m = m0 + qi;
    int j = -1;
    if (m < M) {
      long long jj = (long long)idx[(size_t)m * H + h];
      if (jj < N) j = (int)jj;
    }
    s_idx[e] = j;
    float4* wp = (float4*)(s_w + (size_t)e * 16);
    if (j >= 0) {
      float dx = s_pts[3 * (size_t)j] - q_pts[3 * m], dy = s_pts[3 * (size_t)j + 1] - q_pts[3 * m + 1],
            dz = s_pts[3 * (size_t)j + 2] - q_pts[3 * m + 2];
      float w[16];
#pragma unroll
      for (int k = 0; k < KP_K; k++) {
        float ex = dx - s_kp[3 * k], ey = dy - s_kp[3 * k + 1], ez = dz - s_kp[3 * k + 2];
        w[k] = fmaxf(0.f, 1.f - sqrtf(ex * ex + ey * ey + ez * ez) / sigma);  // kpconv.py:98-99
      }
      w[15] = 0.f;
      wp[0] = make_float4(w[0], w[1], w[2], w[3]);
      wp[1] = make_float4(w[4], w[5], w[6], w[7]);
      wp[2] = make_float4(w[8], w[9], w[10], w[11]);
      wp[3] = make_float4(w[12], w[13], w[14], w[15]);
      atomicMax(&s_nvalid[qi], h + 1);
      if (rowpos[j]) atomicAdd(&s_npos[qi], 1);
    } else {
      wp[0] = wp[1] = wp[2] = wp[3] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();
  // phase 2: one warp per (query, channel slice)
  const int qi = warp / NS, sl = warp - qi * NS, m = m0 + qi;
  if (qi >= QPC || m >= M) return;
  const int c0 = sl * 32 * VEC + lane * VEC;
  const bool active = c0 < C;  // only false on the ragged VEC==1 path
  float acc[KP_K][VEC];
#pragma unroll
  for (int k = 0; k < KP_K; k++)
#pragma unroll
    for (int v = 0; v < VEC; v++) acc[k][v] = 0.f;
  const int nv = s_nvalid[qi];
  const float* wq = s_w + (size_t)qi * H * 16;
  const int* iq = s_idx + qi * H;
#pragma unroll 4
  for (int h = 0; h < nv; h++) {
    int j = iq[h];
    float f[VEC];
    if (j >= 0 && active) {
      const float* fp = feats + (size_t)j * C + c0;
      if constexpr (VEC == 4) {
        float4 t = *(const float4*)fp;
        f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
      } else if constexpr (VEC == 2) {
        float2 t = *(const float2*)fp;
        f[0] = t.x; f[1] = t.y;
      } else {
        f[0] = *fp;
      }
    } else {
#pragma unroll
      for (int v = 0; v < VEC; v++) f[v] = 0.f;
    }
    const float4* wp = (const float4*)(wq + (size_t)h * 16);
    float4 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
    float w[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
    for (int k = 0; k < KP_K; k++)
#pragma unroll
      for (int v = 0; v < VEC; v++) acc[k][v] = fmaf(w[k], f[v], acc[k][v]);
  }
  if (!active) return;
  const float inv = 1.f / (float)max(s_npos[qi], 1);  // kpconv.py:113-116
  float* op = out + (size_t)m * KP_K * C + c0;
#pragma unroll
  for (int k = 0; k < KP_K; k++) {
    if constexpr (VEC == 4) {
      *(float4*)(op + (size_t)k * C) = make_float4(acc[k][0] * inv, acc[k][1] * inv, acc[k][2] * inv, acc[k][3] * inv);
    } else if constexpr (VEC == 2) {
      *(float2*)(op + (size_t)k * C) = make_float2(acc[k][0] * inv, acc[k][1] * inv);
    } else {
      op[(size_t)k * C] = acc[k][0] * inv;
    }
  }
}


// ---------------------------------------------------------------------------------------------- gather, main path
// Warp-autonomous mapping for C_in in {32, 64, 128k}. A warp is split into G = 32/L groups of L lanes; a lane owns
// 4 channels, so a group covers a slice of 4*L channels with 128-bit loads. Two flavours:
//   SPLIT = false : the G groups are G different queries (L = 8 -> C = 32, L = 16 -> C = 64, L = 32 -> one query and
//                   one 128-channel slice per warp); neighbour lists are walked in chunks of L slots.
//   SPLIT = true  : the warp owns ONE (query, 4L-channel slice); the G groups take different L-slot parts of each
//                   32-slot chunk and their partial sums are combined by shuffles at the end. Used when M is small,
//                   to put 2-4x more warps on the machine (the deep stages have only ~500-1100 queries).
// Per chunk: (A) lane i computes the 15 influences of its slot (fast sqrt, 1/sigma multiply, kernel points as
// constant-bank operands) into a warp-private shared tile; (B) each group streams its L rows: one LDG.128 per lane
// and 60 FFMA per row against broadcast LDS.128 reads of the influences. Rows are sorted valid-first, so (B) stops
// at the first slot that is padding for every group. No block barrier, 2.5 KB of shared memory per warp.
#define KP_WS 20  // influence row stride (floats): 80 B keeps STS.128 / LDS.128 of interleaved rows conflict free

// ---------------------------------------------------------------------------------------------- gather, v4
// Two Blackwell-specific choices, both answers to ncu captures of the register-staged predecessor
// (long-scoreboard stalls at 25 % occupancy, FMA pipe the next limiter):
//   * neighbour rows are STAGED IN SHARED MEMORY with cp.async (LDGSTS): the 32 rows of a chunk land in a per-warp
//     buffer while the previous half-chunk is being multiplied, so 16 rows per warp are in flight without holding a
//     single register (the predecessor held 4 per lane group, each pinning a float4);
//   * the 15 x 4 accumulators of a lane are 30 packed pairs updated with FFMA2 (fma.rn.f32x2, scalar-broadcast
//     influence operand): half the issue slots and register-operand traffic per row.
// Shared memory per warp: 32 rows x L x 16 B + the 32 x 20 influence tile (L = 8: 6.5 KB, 16: 10.5 KB, 32: 18.5 KB).
__device__ __forceinline__ float2 ffma2(const float w, const float2 f, const float2 c) {
  unsigned long long rw, rf, rc, rd;
  const float2 w2 = make_float2(w, w);
  rw = *reinterpret_cast<const unsigned long long*>(&w2);
  rf = *reinterpret_cast<const unsigned long long*>(&f);
  rc = *reinterpret_cast<const unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(rw), "l"(rf), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ void gather_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void gather_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void gather_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int L, bool SPLIT, typename IdxT>
__global__ void __launch_bounds__(128, 4) kpconv_gather_v4_kernel(const float* __restrict__ feats,
                                                                  const unsigned char* __restrict__ rowpos,
                                                                  const float* __restrict__ q_pts,
                                                                  const float* __restrict__ s_pts,
                                                                  const IdxT* __restrict__ idx, const KPts kp,
                                                                  float inv_sigma, int M, int N, int H, int C, int NS,
                                                                  const int* __restrict__ order,
                                                                  float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  constexpr int G = 32 / L;
  constexpr int STEP = SPLIT ? 32 : L;  // neighbour slots per chunk (per query)
  constexpr int HALF = L / 2;
  constexpr int WARP_F4 = 32 * L + 32 * KP_WS / 4;  // float4 per warp: row buffer + influence tile
  extern __shared__ float4 s_dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / L, t = lane % L;
  const int gw = blockIdx.x * 4 + warp;
  int m, slice;
  if (SPLIT || L == 32) {
    m = gw / NS;
    slice = gw - m * NS;
  } else {
    m = gw * G + g;
    slice = 0;
  }
  const bool qvalid = m < M;
  if (__all_sync(FULL_MASK, !qvalid)) return;
  if (qvalid && order != nullptr) m = order[m];
  const int mm = qvalid ? m : M - 1;
  const float qx = q_pts[3 * (size_t)mm], qy = q_pts[3 * (size_t)mm + 1], qz = q_pts[3 * (size_t)mm + 2];
  const IdxT* row = idx + (size_t)mm * H;
  const float* fbase = feats + slice * (4 * L) + 4 * t;
  float4* rowbuf = s_dyn + (size_t)warp * WARP_F4;           // [32 rows][L] float4; row (g, u) at (g * L + u) * L
  float* wtile = reinterpret_cast<float*>(rowbuf + 32 * L);  // [32][KP_WS]
  float2 acc[KP_K][2];
#pragma unroll
  for (int k = 0; k < KP_K; k++) acc[k][0] = acc[k][1] = make_float2(0.f, 0.f);
  int npos = 0;
  float4* wrow = (float4*)(wtile + (t * G + g) * KP_WS);  // rows interleaved over groups: (u, g) -> u*G + g
  const int myslot = SPLIT ? lane : t;
  auto load_j = [&](int h) -> int {
    if (qvalid && h < H) {
      long long jj = (long long)row[h];
      if (jj < N) return (int)jj;
    }
    return -1;
  };
  // asynchronous copy of rows [u0, u0 + HALF) of every group for the chunk whose slot indices are `jc`; one commit group
  auto issue_half = [&](int jc, int u0) {
#pragma unroll
    for (int u = u0; u < u0 + HALF; u++) {
      const int ju = __shfl_sync(FULL_MASK, jc, g * L + u);
      if (ju >= 0) gather_cp16(rowbuf + (g * L + u) * L + t, fbase + (size_t)ju * C);
    }
    gather_commit();
  };
  int j = load_j(myslot);
  int jn = load_j(STEP + myslot);
  issue_half(j, 0);
  issue_half(j, HALF);
  for (int h0 = 0; h0 < H; h0 += STEP) {
    if (!__any_sync(FULL_MASK, j >= 0)) break;  // rows are valid-first: nothing but padding from here on
    const int jn2 = load_j(h0 + 2 * STEP + myslot);
    // ---- (A) influences of this lane's slot
    float w[16];
    if (j >= 0) {
      influences16(s_pts[3 * (size_t)j] - qx, s_pts[3 * (size_t)j + 1] - qy, s_pts[3 * (size_t)j + 2] - qz, kp, inv_sigma, w);
      npos += rowpos[j];
    } else {
#pragma unroll
      for (int k = 0; k < 16; k++) w[k] = 0.f;
    }
    wrow[0] = make_float4(w[0], w[1], w[2], w[3]);
    wrow[1] = make_float4(w[4], w[5], w[6], w[7]);
    wrow[2] = make_float4(w[8], w[9], w[10], w[11]);
    wrow[3] = make_float4(w[12], w[13], w[14], w[15]);
    const unsigned vb = __ballot_sync(FULL_MASK, j >= 0);
    int nu = 0;
#pragma unroll
    for (int gg = 0; gg < G; gg++) {
      const unsigned mg = (L == 32) ? vb : ((vb >> (gg * L)) & ((1u << (L & 31)) - 1u));
      nu = max(nu, 32 - __clz(mg));
    }
    // ---- (B) two halves: multiply the landed rows, then refill their slots with the next chunk's rows
#pragma unroll
    for (int half = 0; half < 2; half++) {
      gather_wait<1>();  // this half's rows (the older of the two pending groups) have landed for this lane ...
      __syncwarp();      // ... and for the whole warp; also orders the influence tile stores above
      const int ue = min(nu, (half + 1) * HALF);
#pragma unroll 4
      for (int u = half * HALF; u < ue; u++) {
        const int ju = __shfl_sync(FULL_MASK, j, g * L + u);
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ju >= 0) f = rowbuf[(g * L + u) * L + t];
        const float4* wp = (const float4*)(wtile + (u * G + g) * KP_WS);
        const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
        const float ww[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
        const float2 fa = make_float2(f.x, f.y), fb = make_float2(f.z, f.w);
#pragma unroll
        for (int k = 0; k < KP_K; k++) {
          acc[k][0] = ffma2(ww[k], fa, acc[k][0]);
          acc[k][1] = ffma2(ww[k], fb, acc[k][1]);
        }
      }
      __syncwarp();  // every lane is done with this half's rows (and, after the second half, with the influence tile)
      issue_half(jn, half * HALF);
    }
    j = jn;
    jn = jn2;
  }
  gather_wait<0>();
  if (SPLIT) {
    npos = warp_sum_i(npos);
#pragma unroll
    for (int o = L; o < 32; o <<= 1)
#pragma unroll
      for (int k = 0; k < KP_K; k++) {
        acc[k][0].x += __shfl_xor_sync(FULL_MASK, acc[k][0].x, o);
        acc[k][0].y += __shfl_xor_sync(FULL_MASK, acc[k][0].y, o);
        acc[k][1].x += __shfl_xor_sync(FULL_MASK, acc[k][1].x, o);
        acc[k][1].y += __shfl_xor_sync(FULL_MASK, acc[k][1].y, o);
      }
    if (g != 0) return;
  } else {
#pragma unroll
    for (int o = L >> 1; o > 0; o >>= 1) npos += __shfl_xor_sync(FULL_MASK, npos, o);
  }
  if (!qvalid) return;
  const float inv = 1.f / (float)max(npos, 1);
  float* op = out + (size_t)m * KP_K * C + slice * (4 * L) + 4 * t;
#pragma unroll
  for (int k = 0; k < KP_K; k++)
    *(float4*)(op + (size_t)k * C) = make_float4(acc[k][0].x * inv, acc[k][0].y * inv, acc[k][1].x * inv, acc[k][1].y * inv);
}

// ---------------------------------------------------------------------------------------------- gather, sparse (default)
// The influence w[m,h,k] = max(0, 1 - |s_h - q_m - kp_k| / sigma) is SPARSE: sigma is 0.47 of the search radius, so a
// neighbour lies inside the support of only ~1.7 of the 15 kernel points (measured on the bundled KITTI pair at every
// stage: 11 % non-zero). The dense kernels above spend 89 % of their FMAs on zeros. Here a warp owns one (query, channel
// slice) and works in two phases:
//   (A) lane = neighbour slot: the 15 influences of the slot; for every kernel point a ballot compacts the non-zero
//       (row, w) pairs into that kernel point's entry list in shared memory (ascending slot order, so the summation order
//       is the dense kernels');
//   (B) kernel point by kernel point the warp walks the entry list: one broadcast LDS.64 per entry, one coalesced row
//       load per lane (LDG.32 / .64 / .128 x NV straight from L1 / L2 - a row is touched ~1.7 times), CPL*NV FMAs, and a
//       single accumulator set that is scaled and stored when the list ends. ~8 instructions per entry instead of ~50
//       per neighbour row, no 15-fold accumulator file (60 -> 4..16 registers): 2 - 5 x fewer issued instructions on the
//       C_in >= 128 layers and twice the resident warps.
// Lists are padded with (row 0, w 0) entries to a multiple of 4 so that the walk is unrolled by 4 without predicates.
template <int CPL>
struct GVec;
template <>
struct GVec<1> {
  typedef float T;
};
template <>
struct GVec<2> {
  typedef float2 T;
};
template <>
struct GVec<4> {
  typedef float4 T;
};
__device__ __forceinline__ void gv_fma(float w, const float f, float* a) { a[0] = fmaf(w, f, a[0]); }
__device__ __forceinline__ void gv_fma(float w, const float2 f, float* a) {
  a[0] = fmaf(w, f.x, a[0]);
  a[1] = fmaf(w, f.y, a[1]);
}
__device__ __forceinline__ void gv_fma(float w, const float4 f, float* a) {
  a[0] = fmaf(w, f.x, a[0]);
  a[1] = fmaf(w, f.y, a[1]);
  a[2] = fmaf(w, f.z, a[2]);
  a[3] = fmaf(w, f.w, a[3]);
}
__device__ __forceinline__ void gv_store(float* p, const float* a, float s, float) { p[0] = a[0] * s; }
__device__ __forceinline__ void gv_store(float* p, const float* a, float s, float2) { *(float2*)p = make_float2(a[0] * s, a[1] * s); }
__device__ __forceinline__ void gv_store(float* p, const float* a, float s, float4) {
  *(float4*)p = make_float4(a[0] * s, a[1] * s, a[2] * s, a[3] * s);
}

// CPL channels per lane and vector, NV vectors per lane: the warp's slice is 32 * CPL * NV channels, NS slices per row
template <int CPL, int NV, typename IdxT>
__global__ void __launch_bounds__(128) kpconv_gather_sparse_kernel(const float* __restrict__ feats,
                                                                   const unsigned char* __restrict__ rowpos,
                                                                   const float* __restrict__ q_pts, const float* __restrict__ s_pts,
                                                                   const IdxT* __restrict__ idx, const KPts kp, float inv_sigma,
                                                                   int M, int N, int H, int C, int NS, int HC,
                                                                   const int* __restrict__ order, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  typedef typename GVec<CPL>::T V;
  extern __shared__ int2 s_lists[];  // [4 warps][15][HC]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * 4 + warp;
  const int mi = (int)(gw / NS), slice = (int)(gw - (long long)mi * NS);
  if (mi >= M) return;
  const int m = order != nullptr ? order[mi] : mi;
  const float qx = q_pts[3 * (size_t)m], qy = q_pts[3 * (size_t)m + 1], qz = q_pts[3 * (size_t)m + 2];
  const IdxT* row = idx + (size_t)m * H;
  int2* lst = s_lists + (size_t)warp * KP_K * HC;
  const unsigned lt = (1u << lane) - 1u;
  int cnt[KP_K];
#pragma unroll
  for (int k = 0; k < KP_K; k++) cnt[k] = 0;
  int npos = 0;
  // ---- (A) influences -> per-kernel-point entry lists
  for (int h0 = 0; h0 < H; h0 += 32) {
    const int h = h0 + lane;
    int j = -1;
    if (h < H) {
      const long long jj = (long long)row[h];
      if (jj < N) j = (int)jj;
    }
    if (!__any_sync(FULL_MASK, j >= 0)) break;  // rows are valid-first: nothing but padding from here on
    float w[16];
    if (j >= 0) {
      influences16(s_pts[3 * (size_t)j] - qx, s_pts[3 * (size_t)j + 1] - qy, s_pts[3 * (size_t)j + 2] - qz, kp, inv_sigma, w);
      npos += rowpos[j];
    } else {
#pragma unroll
      for (int k = 0; k < 16; k++) w[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < KP_K; k++) {
      const bool nz = w[k] > 0.f;
      const unsigned mk = __ballot_sync(FULL_MASK, nz);
      if (nz) lst[k * HC + cnt[k] + __popc(mk & lt)] = make_int2(j, __float_as_int(w[k]));
      cnt[k] += __popc(mk);
    }
  }
  // pad every list with three no-op entries (row 0 exists: N >= 1; weight 0)
  if (lane < 3) {
#pragma unroll
    for (int k = 0; k < KP_K; k++) lst[k * HC + cnt[k] + lane] = make_int2(0, 0);
  }
  npos = warp_sum_i(npos);
  const float inv = 1.f / (float)max(npos, 1);  // kpconv.py:113-116
  __syncwarp();
  // ---- (B) one accumulator set, kernel point by kernel point
  const float* fbase = feats + (size_t)slice * (32 * CPL * NV) + CPL * lane;
  float* obase = out + (size_t)m * KP_K * C + (size_t)slice * (32 * CPL * NV) + CPL * lane;
#pragma unroll
  for (int k = 0; k < KP_K; k++) {
    float acc[NV][CPL];
#pragma unroll
    for (int v = 0; v < NV; v++)
#pragma unroll
      for (int c = 0; c < CPL; c++) acc[v][c] = 0.f;
    const int2* lk = lst + k * HC;
    const int n = cnt[k];
    for (int e = 0; e < n; e += 4) {
      int2 en[4];
#pragma unroll
      for (int u = 0; u < 4; u++) en[u] = lk[e + u];
      V f[4][NV];
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < NV; v++) f[u][v] = __ldg((const V*)(fbase + (size_t)en[u].x * C + v * (32 * CPL)));
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < NV; v++) gv_fma(__int_as_float(en[u].y), f[u][v], acc[v]);
    }
#pragma unroll
    for (int v = 0; v < NV; v++) gv_store(obase + (size_t)k * C + v * (32 * CPL), acc[v], inv, V());
  }
}

template <int CPL, int NV, typename IdxT>
static int launch_sparse(const float* feats, const unsigned char* rowpos, const float* q, const float* s, const IdxT* idx,
                         const KPts& kp, float inv_sigma, int M, int N, int H, int C, const int* order, float* out,
                         cudaStream_t stream) {
  const int NS = C / (32 * CPL * NV), HC = H + 3;
  const size_t smem = (size_t)4 * KP_K * HC * sizeof(int2);
  if (smem > 48 * 1024)
    RDM_CUDA(cudaFuncSetAttribute(kpconv_gather_sparse_kernel<CPL, NV, IdxT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long warps = (long long)M * NS;
  RDM_CUDA(rdm_launch_pdl(kpconv_gather_sparse_kernel<CPL, NV, IdxT>, dim3(cdiv(warps, 4)), dim3(128), smem, stream, feats, rowpos, q, s,
                          idx, kp, inv_sigma, M, N, H, C, NS, HC, order, out));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// CTA-cooperative form of the sparse gather for the deep stages (C_in % 128 == 0, H <= 256), where the queries are few
// (500 - 4000) and a warp-per-query kernel is bound by the LATENCY of its serial chain (index -> point -> influence ->
// lists -> rows), not by throughput. One CTA of 256 threads owns one query:
//   (A) thread = neighbour slot: ALL slots of the query in one step (no chunk loop); per kernel point the 8 warps ballot,
//       publish their counts, and scatter their (row, w) entries behind the exclusive prefix over the warps (slot order
//       kept: same summation order as the dense kernels);
//   (B) the 15 x (C/128) (kernel point, 128-channel slice) tasks are dealt round-robin to the 8 warps; a task issues all
//       its row loads at once (lists padded to a multiple of 8 with zero-weight entries) and stores one 512-byte row.
template <typename IdxT>
__global__ void __launch_bounds__(256) kpconv_gather_sparse_cta_kernel(const float* __restrict__ feats,
                                                                       const unsigned char* __restrict__ rowpos,
                                                                       const float* __restrict__ q_pts, const float* __restrict__ s_pts,
                                                                       const IdxT* __restrict__ idx, const KPts kp, float inv_sigma,
                                                                       int M, int N, int H, int C, int HC,
                                                                       const int* __restrict__ order, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ int2 s_lists[];  // [15][HC]
  __shared__ int s_wcnt[8][16];      // entries of warp w in list k
  __shared__ int s_npos[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = order != nullptr ? order[blockIdx.x] : blockIdx.x;
  const float qx = q_pts[3 * (size_t)m], qy = q_pts[3 * (size_t)m + 1], qz = q_pts[3 * (size_t)m + 2];
  int j = -1;
  if (tid < H) {
    const long long jj = (long long)idx[(size_t)m * H + tid];
    if (jj < N) j = (int)jj;
  }
  float w[16];
  int np = 0;
  if (j >= 0) {
    influences16(s_pts[3 * (size_t)j] - qx, s_pts[3 * (size_t)j + 1] - qy, s_pts[3 * (size_t)j + 2] - qz, kp, inv_sigma, w);
    np = rowpos[j];
  } else {
#pragma unroll
    for (int k = 0; k < 16; k++) w[k] = 0.f;
  }
  const unsigned lt = (1u << lane) - 1u;
  int mypos[KP_K];  // position of this thread's entry inside its warp's run of list k
#pragma unroll
  for (int k = 0; k < KP_K; k++) {
    const unsigned mk = __ballot_sync(FULL_MASK, w[k] > 0.f);
    mypos[k] = __popc(mk & lt);
    if (lane == 0) s_wcnt[warp][k] = __popc(mk);
  }
  np = warp_sum_i(np);
  if (lane == 0) s_npos[warp] = np;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < KP_K; k++) {
    if (w[k] > 0.f) {
      int base = 0;
      for (int ww = 0; ww < warp; ww++) base += s_wcnt[ww][k];
      s_lists[k * HC + base + mypos[k]] = make_int2(j, __float_as_int(w[k]));
    }
  }
  if (tid < KP_K * 7) {  // pad every list with seven no-op entries (row 0 exists: N >= 1; weight 0)
    const int k = tid / 7, i = tid - k * 7;
    int tot = 0;
#pragma unroll
    for (int ww = 0; ww < 8; ww++) tot += s_wcnt[ww][k];
    s_lists[k * HC + tot + i] = make_int2(0, 0);
  }
  __syncthreads();
  int npos = 0;
#pragma unroll
  for (int ww = 0; ww < 8; ww++) npos += s_npos[ww];
  const float inv = 1.f / (float)max(npos, 1);  // kpconv.py:113-116
  const int NS = C >> 7, ntask = KP_K * NS;
  for (int t = warp; t < ntask; t += 8) {
    const int k = t % KP_K, slice = t / KP_K;
    int n = 0;
#pragma unroll
    for (int ww = 0; ww < 8; ww++) n += s_wcnt[ww][k];
    const int2* lk = s_lists + k * HC;
    const float* fb = feats + (size_t)slice * 128 + 4 * lane;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int e = 0; e < n; e += 8) {
      int2 en[8];
#pragma unroll
      for (int u = 0; u < 8; u++) en[u] = lk[e + u];
      float4 f[8];
#pragma unroll
      for (int u = 0; u < 8; u++) f[u] = __ldg((const float4*)(fb + (size_t)en[u].x * C));
#pragma unroll
      for (int u = 0; u < 8; u++) gv_fma(__int_as_float(en[u].y), f[u], acc);
    }
    gv_store(out + (size_t)m * KP_K * C + (size_t)k * C + (size_t)slice * 128 + 4 * lane, acc, inv, float4());
  }
}

template <typename IdxT>
static int launch_sparse_cta(const float* feats, const unsigned char* rowpos, const float* q, const float* s, const IdxT* idx,
                             const KPts& kp, float inv_sigma, int M, int N, int H, int C, const int* order, float* out,
                             cudaStream_t stream) {
  const int HC = H + 7;
  const size_t smem = (size_t)KP_K * HC * sizeof(int2);
  if (smem > 40 * 1024)
    RDM_CUDA(cudaFuncSetAttribute(kpconv_gather_sparse_cta_kernel<IdxT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RDM_CUDA(rdm_launch_pdl(kpconv_gather_sparse_cta_kernel<IdxT>, dim3(M), dim3(256), smem, stream, feats, rowpos, q, s, idx, kp, inv_sigma,
                          M, N, H, C, HC, order, out));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

template <int L, bool SPLIT, typename IdxT>
static int launch_v4(long long warps, const float* feats, const unsigned char* rowpos, const float* q, const float* s,
                     const IdxT* idx, const KPts& kp, float inv_sigma, int M, int N, int H, int C, int NS, const int* order,
                     float* out, cudaStream_t stream) {
  const size_t smem = 4 * (size_t)(32 * L + 32 * KP_WS / 4) * sizeof(float4);
  if (smem > 48 * 1024)  // per-device attribute: set on every call (sub-microsecond) rather than cached per process
    RDM_CUDA(cudaFuncSetAttribute(kpconv_gather_v4_kernel<L, SPLIT, IdxT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RDM_CUDA(rdm_launch_pdl(kpconv_gather_v4_kernel<L, SPLIT, IdxT>, dim3(cdiv(warps, 4)), dim3(128), smem, stream, feats, rowpos, q, s,
                          idx, kp, inv_sigma, M, N, H, C, NS, order, out));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

template <typename IdxT>
static int launch_gather(const float* feats, const unsigned char* rowpos, const float* q, const float* s,
                         const IdxT* idx, const float* kpts, const float* h_kpts, float sigma, int M, int N, int H, int C,
                         const int* order, float* out, cudaStream_t stream) {
  // the 15 kernel points travel by value in the kernel-parameter constant bank: `d - kp` costs no load
  KPts kp;
  for (int k = 0; k < KP_K; k++) {
    kp.x[k] = h_kpts[3 * k];
    kp.y[k] = h_kpts[3 * k + 1];
    kp.z[k] = h_kpts[3 * k + 2];
  }
  kp.x[15] = kp.y[15] = kp.z[15] = 1.0e6f;  // dummy 16th point: |d - kp| / sigma >> 1 -> influence exactly 0
  const float inv_sigma = 1.f / sigma;
  if (C == 1) {
    RDM_CUDA(rdm_launch_pdl(kpconv_gather_c1_kernel<IdxT>, dim3(cdiv(M, 32)), dim3(256), 0, stream, feats, q, s, idx, kp, inv_sigma, M,
                            N, H, order, out));
    RDM_LAUNCH_CHECK();
    return RDM_OK;
  }
  if (C == 32 || C == 64 || (C % 128 == 0 && C <= 4096)) {
    // candidates from the cheapest mapping (largest L, groups = different queries) to the most parallel one
    // (small L, groups split one query's neighbour list); take the first that puts >= `wps` warps on each of 148 SMs.
    // RDM_GATHER_MODE (A/B knob): "auto" (default) = CTA-cooperative sparse kernel for C_in % 128 == 0, dense cp.async /
    // FFMA2 kernels for C_in = 32 / 64 (a 128-byte row is one FFMA per lane and entry: the per-entry bookkeeping of the
    // sparse forms costs more than the zero FMAs it saves - measured 1.6 vs 4.1 TB/s on G); "dense" / "sparse" / "sparsew"
    // force the dense, the CTA-cooperative sparse and the warp-per-query sparse kernels wherever they apply.
    static int mode = -1;
    if (mode < 0) {
      const char* e = getenv("RDM_GATHER_MODE");
      mode = (e == nullptr || e[0] == 'a') ? 0 : (e[0] == 'd' ? 1 : (e[6] == 'w' ? 3 : 2));
    }
    if (mode == 3 && H <= 1024) {
      const long long want_w = 148LL * 8;
#define SPARSE(CPLv, NVv) \
  return launch_sparse<CPLv, NVv, IdxT>(feats, rowpos, q, s, idx, kp, inv_sigma, M, N, H, C, order, out, stream)
      if (C == 32) SPARSE(1, 1);
      if (C == 64) SPARSE(2, 1);
      if (C % 512 == 0 && (long long)M * (C / 512) >= want_w) SPARSE(4, 4);
      if (C % 256 == 0 && (long long)M * (C / 256) >= want_w) SPARSE(4, 2);
      SPARSE(4, 1);
#undef SPARSE
    }
    // measured on the bench pyramid with the checkpoint's kernel points (profiles/README.md r2c): the CTA-cooperative sparse
    // kernel wins for C_in >= 256 (M <= ~1300 queries: 31.7 vs 39.9 us, 23.6 vs 28 us) and loses at C_in = 128, M = 3600
    // (42 vs 35.6 us), where the dense kernel already fills the machine
    if (((mode == 0 && C >= 256) || mode == 2) && C % 128 == 0 && H <= 256)
      return launch_sparse_cta<IdxT>(feats, rowpos, q, s, idx, kp, inv_sigma, M, N, H, C, order, out, stream);
    // dense mappings: candidates from the cheapest (largest L, groups = different queries) to the most parallel (small L,
    // groups split one query's neighbour list); take the first that puts >= `wps` warps on each of 148 SMs. (A wave-count
    // cost model that preferred split mappings for the strided layers measured 10-40 % SLOWER: r2c.)
    static int wps = 0;
    if (wps == 0) {
      const char* e = getenv("RDM_GATHER_WPS");  // tuning knob: warps per SM a mapping must reach before it is taken
      wps = (e && atoi(e) > 0) ? atoi(e) : 8;  // measured (profiles/r01e): 8 beats 16 on the strided / deep layers
    }
    const long long want = 148LL * wps;
#define GATHER4(Lv, SPLITv, warps, NSv) \
  return launch_v4<Lv, SPLITv, IdxT>((warps), feats, rowpos, q, s, idx, kp, inv_sigma, M, N, H, C, (NSv), order, out, stream)
      if (C == 32) {
        if (cdiv(M, 4) >= want) GATHER4(8, false, cdiv(M, 4), 1);
        GATHER4(8, true, M, 1);
      } else if (C == 64) {
        if (cdiv(M, 2) >= want) GATHER4(16, false, cdiv(M, 2), 1);
        if (M >= want) GATHER4(16, true, M, 1);
        GATHER4(8, true, 2LL * M, 2);
      } else {
        if ((long long)M * (C / 128) >= want) GATHER4(32, false, (long long)M * (C / 128), C / 128);
        if ((long long)M * (C / 64) >= want) GATHER4(16, true, (long long)M * (C / 64), C / 64);
        GATHER4(8, true, (long long)M * (C / 32), C / 32);
      }
#undef GATHER4
  }
  // other widths (not used by RDMNet): CTA-level kernel, influences staged once per query in shared memory
  int VEC = (C % 128 == 0) ? 4 : (C % 64 == 0 ? 2 : 1);
  int NS = cdiv(C, 32 * VEC);
  RDM_CHECK_ARG(NS <= 8, "rdm_kpconv_gather: C_in=%d too wide for one CTA (max 1024)", C);
  int QPC = 8 / NS;
  size_t smem = (size_t)QPC * H * 16 * 4 + (size_t)QPC * H * 4 + 2 * QPC * 4;
  int grid = cdiv(M, QPC);
#define LAUNCH(V)                                                                                                  \
  do {                                                                                                             \
    if (smem > 48 * 1024)                                                                                          \
      RDM_CUDA(cudaFuncSetAttribute(kpconv_gather_kernel<V, IdxT>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                    (int)smem));                                                                   \
    kpconv_gather_kernel<V, IdxT><<<grid, 256, smem, stream>>>(feats, rowpos, q, s, idx, kpts, sigma, M, N, H, C,  \
                                                               NS, QPC, out);                                      \
  } while (0)
  RDM_CHECK_ARG(smem <= 200 * 1024, "rdm_kpconv_gather: neighbour width H=%d too large", H);
  if (VEC == 4) LAUNCH(4);
  else if (VEC == 2) LAUNCH(2);
  else LAUNCH(1);
#undef LAUNCH
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

extern "C" int rdm_kpconv_gather(const float* s_feats, const float* q_points, const float* s_points,
                                 const void* neighbor_indices, int index_bytes, const float* kernel_points,
                                 const float* h_kernel_points, float sigma,
                                 int M, int N, int H, int C_in, const int* query_order, float* out_weighted,
                                 unsigned char* rowpos_scratch, cudaStream_t stream) {
  return rdm_kpconv_gather_impl(s_feats, q_points, s_points, neighbor_indices, index_bytes, kernel_points, h_kernel_points, sigma,
                                M, N, H, C_in, query_order, out_weighted, rowpos_scratch, 0, stream);
}

// rowpos_ready != 0: `rowpos` already holds (sum_c s_feats[n,c] > 0) for every support row (written by the producer of
// s_feats, rdm_groupnorm_apply), so the prepass is skipped.
int rdm_kpconv_gather_impl(const float* s_feats, const float* q_points, const float* s_points, const void* neighbor_indices,
                           int index_bytes, const float* kernel_points, const float* h_kernel_points, float sigma, int M, int N,
                           int H, int C_in, const int* query_order, float* out_weighted, unsigned char* rowpos_scratch,
                           int rowpos_ready, cudaStream_t stream) {
  RDM_CHECK_ARG(M >= 0 && N >= 0 && H >= 1 && C_in >= 1 && sigma > 0.f, "rdm_kpconv_gather: bad arguments");
  RDM_CHECK_ARG(index_bytes == 4 || index_bytes == 8, "rdm_kpconv_gather: index_bytes must be 4 or 8");
  RDM_CHECK_ARG(kernel_points != nullptr && h_kernel_points != nullptr, "rdm_kpconv_gather: kernel points missing");
  if (M == 0) return RDM_OK;
  if (N == 0) {  // no supports: every slot is padding
    RDM_CUDA(cudaMemsetAsync(out_weighted, 0, (size_t)M * KP_K * C_in * sizeof(float), stream));
    return RDM_OK;
  }
  const int prof = rdm_prof_begin(RDM_PROF_KPCONV_GATHER, M, N, H, C_in, stream);
  if (C_in > 1 && N > 0 && !rowpos_ready) {
    row_positive_kernel<<<cdiv(N, 8), 256, 0, stream>>>(s_feats, N, C_in, rowpos_scratch);
    RDM_LAUNCH_CHECK();
  }
  int rc;
  if (index_bytes == 8)
    rc = launch_gather<int64_t>(s_feats, rowpos_scratch, q_points, s_points, (const int64_t*)neighbor_indices,
                                kernel_points, h_kernel_points, sigma, M, N, H, C_in, query_order, out_weighted, stream);
  else
    rc = launch_gather<int>(s_feats, rowpos_scratch, q_points, s_points, (const int*)neighbor_indices, kernel_points,
                            h_kernel_points, sigma, M, N, H, C_in, query_order, out_weighted, stream);
  rdm_prof_end(prof, stream);
  return rc;
}

// ---------------------------------------------------------------------------------------------- maxpool / upsample
// out[m, c] = max_h F'[idx[m,h], c]  with F' = F plus one zero row (functional.py:64-66)
template <typename IdxT>
__global__ void maxpool_kernel(const float* __restrict__ f, const IdxT* __restrict__ idx, int M, int N, int H, int C,
                               float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  int cv = C >> 2;
  if (e >= (long long)M * cv) return;
  int m = (int)(e / cv), c4 = (int)(e - (long long)m * cv);
  float4 best = make_float4(-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f);
  const IdxT* row = idx + (size_t)m * H;
  for (int h = 0; h < H; h++) {
    long long j = (long long)row[h];
    if (j > N) continue;  // column beyond the reference's row width (rdm_mark_reference_width): not part of the tensor
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < N) v = *(const float4*)(f + (size_t)j * C + 4 * c4);
    best.x = fmaxf(best.x, v.x);
    best.y = fmaxf(best.y, v.y);
    best.z = fmaxf(best.z, v.z);
    best.w = fmaxf(best.w, v.w);
  }
  *(float4*)(out + (size_t)m * C + 4 * c4) = best;
}

extern "C" int rdm_maxpool(const float* feats, const void* neighbor_indices, int index_bytes, int M, int N, int H, int C,
                           float* out, cudaStream_t stream) {
  RDM_CHECK_ARG(C % 4 == 0 && H >= 1, "rdm_maxpool: C must be a multiple of 4");
  RDM_CHECK_ARG(index_bytes == 4 || index_bytes == 8, "rdm_maxpool: index_bytes must be 4 or 8");
  if (M == 0) return RDM_OK;
  long long total = (long long)M * (C / 4);
  if (index_bytes == 8)
    RDM_CUDA(rdm_launch_pdl(maxpool_kernel<int64_t>, dim3(cdiv(total, 256)), dim3(256), 0, stream, feats,
                            (const int64_t*)neighbor_indices, M, N, H, C, out));
  else
    RDM_CUDA(rdm_launch_pdl(maxpool_kernel<int>, dim3(cdiv(total, 256)), dim3(256), 0, stream, feats, (const int*)neighbor_indices, M, N,
                            H, C, out));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// out[m, :C1] = F'[idx[m,0]] ; out[m, C1:] = skip[m]   (functional.py:6-22 + backbone.py:129-141 concat)
template <typename IdxT>
__global__ void upsample_concat_kernel(const float* __restrict__ f, const IdxT* __restrict__ idx, int idx_stride,
                                       const float* __restrict__ skip, int M, int N, int C1, int C2,
                                       float* __restrict__ out, int ld_out) {
  pdl_trigger();
  pdl_wait();
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  int C = C1 + C2;
  if (e >= (long long)M * ld_out) return;
  int m = (int)(e / ld_out), c = (int)(e - (long long)m * ld_out);
  if (c >= C) {  // row padding (ld_out > C): defined zeros
    out[e] = 0.f;
    return;
  }
  float v;
  if (c < C1) {
    long long j = (long long)idx[(size_t)m * idx_stride];
    v = j < N ? f[(size_t)j * C1 + c] : 0.f;
  } else {
    v = skip[(size_t)m * C2 + (c - C1)];
  }
  out[e] = v;
}

extern "C" int rdm_upsample_concat(const float* feats, const void* upsample_indices, int index_bytes, int index_stride,
                                   const float* skip, int M, int N, int C1, int C2, float* out, cudaStream_t stream) {
  return rdm_upsample_concat_ld(feats, upsample_indices, index_bytes, index_stride, skip, M, N, C1, C2, out, C1 + C2, stream);
}

int rdm_upsample_concat_ld(const float* feats, const void* upsample_indices, int index_bytes, int index_stride,
                           const float* skip, int M, int N, int C1, int C2, float* out, int ld_out, cudaStream_t stream) {
  RDM_CHECK_ARG(index_bytes == 4 || index_bytes == 8, "rdm_upsample_concat: index_bytes must be 4 or 8");
  RDM_CHECK_ARG(ld_out >= C1 + C2, "rdm_upsample_concat: output row stride too small");
  if (M == 0) return RDM_OK;
  long long total = (long long)M * ld_out;
  if (index_bytes == 8)
    RDM_CUDA(rdm_launch_pdl(upsample_concat_kernel<int64_t>, dim3(cdiv(total, 256)), dim3(256), 0, stream, feats,
                            (const int64_t*)upsample_indices, index_stride, skip, M, N, C1, C2, out, ld_out));
  else
    RDM_CUDA(rdm_launch_pdl(upsample_concat_kernel<int>, dim3(cdiv(total, 256)), dim3(256), 0, stream, feats,
                            (const int*)upsample_indices, index_stride, skip, M, N, C1, C2, out, ld_out));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}
