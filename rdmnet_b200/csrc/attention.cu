// ThDRoFormer attention kernels: 3-D rotary embedding and multi-head softmax attention (flash-style, fp32).
//
// Reference semantics: rdmnet/thdroformer/thdroformer.py:56-85 (RotaryPositionalEmbedding.forward), :20-40
// (dynamic_attention, k=None branch), :108-139 (RPEMultiHeadAttention.forward) and
// geotransformer/modules/transformer/vanilla_transformer.py:31-70 (MultiHeadAttention.forward, no masks).
// Sequence lengths are <= ~450 superpoints with head_dim 32: the work is latency-bound, not a dense contraction
// worth tensor cores (0.2 GFLOP per tower), so this is a SIMT kernel with K/V tiles staged in shared memory.
#include "common.cuh"
#include "../../include/rdm_sm100.h"

// x[n, h*D + 2i]   <- x0*cos(t) - x1*sin(t)
// x[n, h*D + 2i+1] <- x1*cos(t) + x0*sin(t),   t = sigmoid(emb[n, h*D/2 + i]) * 3.14159265359 * 2
__global__ void rope_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ emb, int lde,
                            float* __restrict__ y, int ldy, int N, int C) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  int half = C >> 1;
  if (e >= (long long)N * half) return;
  int n = (int)(e / half), p = (int)(e - (long long)n * half);
  float em = emb[(size_t)n * lde + p];
  float theta = (1.f / (1.f + expf(-em))) * 3.14159265359f * 2.f;  // thdroformer.py:78
  float s, c;
  sincosf(theta, &s, &c);
  float x0 = x[(size_t)n * ldx + 2 * p], x1 = x[(size_t)n * ldx + 2 * p + 1];
  y[(size_t)n * ldy + 2 * p] = x0 * c - x1 * s;
  y[(size_t)n * ldy + 2 * p + 1] = x1 * c + x0 * s;
}

extern "C" int rdm_rope(const float* x, int ldx, const float* emb, int lde, float* y, int ldy, int N, int C,
                        cudaStream_t stream) {
  RDM_CHECK_ARG(C % 2 == 0 && N >= 0, "rdm_rope: C must be even");
  if (N == 0) return RDM_OK;
  long long total = (long long)N * (C / 2);
  rope_kernel<<<cdiv(total, 256), 256, 0, stream>>>(x, ldx, emb, lde, y, ldy, N, C);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// out[n, h*D:(h+1)*D] = softmax_j(q_n . k_j / sqrt(D)) @ V,   D <= 64.
// grid (ceil(Nq/QT), heads); 4 warps; each warp owns QT/4 queries; K/V tiles of KT keys in shared memory.
#define ATT_KT 64
#define ATT_QPW 4
#define ATT_QT (4 * ATT_QPW)

__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ Q, int ldq,
                                                        const float* __restrict__ K, int ldk,
                                                        const float* __restrict__ V, int ldv, float* __restrict__ O,
                                                        int ldo, int Nq, int Nk, int D, float scale) {
  __shared__ float Ks[ATT_KT][65];
  __shared__ float Vs[ATT_KT][64];
  __shared__ float Qs[4][ATT_QPW][64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = blockIdx.y, q0 = blockIdx.x * ATT_QT + warp * ATT_QPW;
  const int hoff = h * D;
  // stage this warp's queries (pre-scaled)
  for (int i = 0; i < ATT_QPW; i++) {
    int q = q0 + i;
    for (int d = lane; d < D; d += 32) Qs[warp][i][d] = q < Nq ? Q[(size_t)q * ldq + hoff + d] * scale : 0.f;
  }
  float m[ATT_QPW], l[ATT_QPW], acc0[ATT_QPW], acc1[ATT_QPW];
#pragma unroll
  for (int i = 0; i < ATT_QPW; i++) {
    m[i] = -3.0e38f;
    l[i] = 0.f;
    acc0[i] = acc1[i] = 0.f;
  }
  for (int k0 = 0; k0 < Nk; k0 += ATT_KT) {
    __syncthreads();
    for (int e = tid; e < ATT_KT * D; e += 128) {
      int j = e / D, d = e - j * D;
      bool ok = k0 + j < Nk;
      Ks[j][d] = ok ? K[(size_t)(k0 + j) * ldk + hoff + d] : 0.f;
      Vs[j][d] = ok ? V[(size_t)(k0 + j) * ldv + hoff + d] : 0.f;
    }
    __syncthreads();
    const int j0 = lane, j1 = lane + 32;
    const bool ok0 = k0 + j0 < Nk, ok1 = k0 + j1 < Nk;
#pragma unroll
    for (int i = 0; i < ATT_QPW; i++) {
      float s0 = 0.f, s1 = 0.f;
      for (int d = 0; d < D; d++) {
        float qd = Qs[warp][i][d];
        s0 = fmaf(qd, Ks[j0][d], s0);
        s1 = fmaf(qd, Ks[j1][d], s1);
      }
      s0 = ok0 ? s0 : -3.0e38f;
      s1 = ok1 ? s1 : -3.0e38f;
      float mx = warp_max(fmaxf(s0, s1));
      float mnew = fmaxf(m[i], mx);
      float corr = __expf(m[i] - mnew);
      float p0 = ok0 ? __expf(s0 - mnew) : 0.f, p1 = ok1 ? __expf(s1 - mnew) : 0.f;
      l[i] = l[i] * corr + warp_sum(p0 + p1);
      float a0 = acc0[i] * corr, a1 = acc1[i] * corr;
#pragma unroll 8
      for (int j = 0; j < 32; j++) {
        float pj = __shfl_sync(FULL_MASK, p0, j);
        a0 = fmaf(pj, Vs[j][lane], a0);
        if (D > 32) a1 = fmaf(pj, Vs[j][lane + 32], a1);
      }
#pragma unroll 8
      for (int j = 0; j < 32; j++) {
        float pj = __shfl_sync(FULL_MASK, p1, j);
        a0 = fmaf(pj, Vs[j + 32][lane], a0);
        if (D > 32) a1 = fmaf(pj, Vs[j + 32][lane + 32], a1);
      }
      acc0[i] = a0;
      acc1[i] = a1;
      m[i] = mnew;
    }
  }
#pragma unroll
  for (int i = 0; i < ATT_QPW; i++) {
    int q = q0 + i;
    if (q >= Nq) continue;
    float inv = 1.f / l[i];
    if (lane < D) O[(size_t)q * ldo + hoff + lane] = acc0[i] * inv;
    if (lane + 32 < D) O[(size_t)q * ldo + hoff + lane + 32] = acc1[i] * inv;
  }
}

extern "C" int rdm_attention(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, float* O,
                             int ldo, int Nq, int Nk, int heads, int head_dim, cudaStream_t stream) {
  RDM_CHECK_ARG(head_dim >= 1 && head_dim <= 64 && heads >= 1, "rdm_attention: head_dim must be <= 64");
  RDM_CHECK_ARG(Nk >= 1 || Nq == 0, "rdm_attention: empty key set");
  if (Nq == 0) return RDM_OK;
  dim3 grid(cdiv(Nq, ATT_QT), heads);
  attention_kernel<<<grid, 128, 0, stream>>>(Q, ldq, K, ldk, V, ldv, O, ldo, Nq, Nk, head_dim,
                                            1.0f / sqrtf((float)head_dim));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}
