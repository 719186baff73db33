// Training path, part 2: backward kernels of the matcher-side operators. The reference differentiates
// them with PyTorch autograd over its ATen graphs (experiments/trainval.py:43-50); the expressions differentiated here:
//   rdm_rope_bwd        RotaryPositionalEmbedding.forward            rdmnet/thdroformer/thdroformer.py:56-85
//   rdm_attention_bwd   softmax(q k^T / sqrt(d)) v per head          rdmnet/thdroformer/thdroformer.py:20-40 (k = None),
//                                                                    geotransformer/modules/transformer/vanilla_transformer.py:54-66
//   rdm_sinkhorn_bwd    LearnableLogOptimalTransport (100 unrolled   geotransformer/modules/sinkhorn/learnable_sinkhorn.py:13-66
//                       log-domain iterations)
// All fp32 and deterministic (no atomics): every gradient element is owned by one warp.
#include "common.cuh"
#include "../../include/rdm_sm100.h"

namespace {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------------ RoPE
// y[2i] = x[2i] cos t - x[2i+1] sin t, y[2i+1] = x[2i+1] cos t + x[2i] sin t, t = 2 pi sigmoid(emb[i])
__global__ void __launch_bounds__(256) rope_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ emb, int lde,
                                                       const float* __restrict__ dy, int ldy, int N, int C, float* __restrict__ dx,
                                                       float* __restrict__ demb) {
  const long long e = blockIdx.x * 256LL + threadIdx.x;
  const int half = C >> 1;
  if (e >= (long long)N * half) return;
  const int n = (int)(e / half), p = (int)(e - (long long)n * half);
  const float sg = 1.f / (1.f + expf(-emb[(size_t)n * lde + p]));
  const float theta = sg * 3.14159265359f * 2.f;
  float s, c;
  sincosf(theta, &s, &c);
  const float x0 = x[(size_t)n * ldx + 2 * p], x1 = x[(size_t)n * ldx + 2 * p + 1];
  const float g0 = dy[(size_t)n * ldy + 2 * p], g1 = dy[(size_t)n * ldy + 2 * p + 1];
  dx[(size_t)n * C + 2 * p] = g0 * c + g1 * s;
  dx[(size_t)n * C + 2 * p + 1] = g1 * c - g0 * s;
  if (demb) {
    const float dtheta = g0 * (-x0 * s - x1 * c) + g1 * (-x1 * s + x0 * c);
    demb[(size_t)n * half + p] = dtheta * 3.14159265359f * 2.f * sg * (1.f - sg);
  }
}

// ------------------------------------------------------------------------------------------------- attention
// Pass A: one warp per (query i, head h). Lanes stride the keys; two sweeps over K (row statistics, then dS) with the score
// recomputed - N <= a few hundred superpoints, the K/V rows stay in L1/L2. Writes dQ and the row statistics
// {max, 1/sum, delta = dO . O} that pass B needs.
template <int D>
__global__ void __launch_bounds__(256) attn_bwd_q_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ K, int ldk,
                                                         const float* __restrict__ V, int ldv, const float* __restrict__ O, int ldo,
                                                         const float* __restrict__ dO, int ldg, int Nq, int Nk, float scale,
                                                         float* __restrict__ dQ, int ldd, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5), h = blockIdx.y;
  if (i >= Nq) return;
  const int ho = h * D;
  float q[D], g[D];
  float delta = 0.f;
#pragma unroll
  for (int d = 0; d < D; d++) {
    q[d] = Q[(size_t)i * ldq + ho + d] * scale;
    g[d] = dO[(size_t)i * ldg + ho + d];
    delta = fmaf(g[d], O[(size_t)i * ldo + ho + d], delta);
  }
  float mx = -3.4e38f;
  for (int j = lane; j < Nk; j += 32) {
    const float4* kr = (const float4*)(K + (size_t)j * ldk + ho);
    float s = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < D / 4; d4++) {
      const float4 k4 = kr[d4];
      s = fmaf(q[4 * d4], k4.x, fmaf(q[4 * d4 + 1], k4.y, fmaf(q[4 * d4 + 2], k4.z, fmaf(q[4 * d4 + 3], k4.w, s))));
    }
    mx = fmaxf(mx, s);
  }
  mx = wmax(mx);
  float sum = 0.f;
  for (int j = lane; j < Nk; j += 32) {
    const float4* kr = (const float4*)(K + (size_t)j * ldk + ho);
    float s = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < D / 4; d4++) {
      const float4 k4 = kr[d4];
      s = fmaf(q[4 * d4], k4.x, fmaf(q[4 * d4 + 1], k4.y, fmaf(q[4 * d4 + 2], k4.z, fmaf(q[4 * d4 + 3], k4.w, s))));
    }
    sum += expf(s - mx);
  }
  const float inv = 1.f / wsum(sum);
  float acc[D];
#pragma unroll
  for (int d = 0; d < D; d++) acc[d] = 0.f;
  for (int j = lane; j < Nk; j += 32) {
    const float4* kr = (const float4*)(K + (size_t)j * ldk + ho);
    const float4* vr = (const float4*)(V + (size_t)j * ldv + ho);
    float kk[D];
    float s = 0.f, dp = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < D / 4; d4++) {
      const float4 k4 = kr[d4], v4 = vr[d4];
      kk[4 * d4] = k4.x, kk[4 * d4 + 1] = k4.y, kk[4 * d4 + 2] = k4.z, kk[4 * d4 + 3] = k4.w;
      s = fmaf(q[4 * d4], k4.x, fmaf(q[4 * d4 + 1], k4.y, fmaf(q[4 * d4 + 2], k4.z, fmaf(q[4 * d4 + 3], k4.w, s))));
      dp = fmaf(g[4 * d4], v4.x, fmaf(g[4 * d4 + 1], v4.y, fmaf(g[4 * d4 + 2], v4.z, fmaf(g[4 * d4 + 3], v4.w, dp))));
    }
    const float ds = expf(s - mx) * inv * (dp - delta) * scale;
#pragma unroll
    for (int d = 0; d < D; d++) acc[d] = fmaf(ds, kk[d], acc[d]);
  }
#pragma unroll
  for (int d = 0; d < D; d++) {
    const float t = wsum(acc[d]);
    if (lane == (d & 31)) dQ[(size_t)i * ldd + ho + d] = t;
  }
  if (lane == 0) {
    float* st = stats + ((size_t)h * Nq + i) * 3;
    st[0] = mx, st[1] = inv, st[2] = delta;
  }
}

// Pass B: one warp per (key j, head h); lanes stride the queries, using the row statistics of pass A.
template <int D>
__global__ void __launch_bounds__(256) attn_bwd_kv_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ K, int ldk,
                                                          const float* __restrict__ V, int ldv, const float* __restrict__ dO, int ldg,
                                                          int Nq, int Nk, float scale, const float* __restrict__ stats,
                                                          float* __restrict__ dK, float* __restrict__ dV, int ldd) {
  const int lane = threadIdx.x & 31, j = blockIdx.x * 8 + (threadIdx.x >> 5), h = blockIdx.y;
  if (j >= Nk) return;
  const int ho = h * D;
  float k[D], v[D], ak[D], av[D];
#pragma unroll
  for (int d = 0; d < D; d++) {
    k[d] = K[(size_t)j * ldk + ho + d];
    v[d] = V[(size_t)j * ldv + ho + d];
    ak[d] = 0.f, av[d] = 0.f;
  }
  for (int i = lane; i < Nq; i += 32) {
    const float4* qr = (const float4*)(Q + (size_t)i * ldq + ho);
    const float4* gr = (const float4*)(dO + (size_t)i * ldg + ho);
    const float* st = stats + ((size_t)h * Nq + i) * 3;
    float qq[D], gg[D];
    float s = 0.f, dp = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < D / 4; d4++) {
      const float4 q4 = qr[d4], g4 = gr[d4];
      qq[4 * d4] = q4.x, qq[4 * d4 + 1] = q4.y, qq[4 * d4 + 2] = q4.z, qq[4 * d4 + 3] = q4.w;
      gg[4 * d4] = g4.x, gg[4 * d4 + 1] = g4.y, gg[4 * d4 + 2] = g4.z, gg[4 * d4 + 3] = g4.w;
      s = fmaf(q4.x * scale, k[4 * d4], fmaf(q4.y * scale, k[4 * d4 + 1], fmaf(q4.z * scale, k[4 * d4 + 2], fmaf(q4.w * scale, k[4 * d4 + 3], s))));
      dp = fmaf(g4.x, v[4 * d4], fmaf(g4.y, v[4 * d4 + 1], fmaf(g4.z, v[4 * d4 + 2], fmaf(g4.w, v[4 * d4 + 3], dp))));
    }
    const float p = expf(s - st[0]) * st[1];
    const float ds = p * (dp - st[2]) * scale;
#pragma unroll
    for (int d = 0; d < D; d++) {
      ak[d] = fmaf(ds, qq[d], ak[d]);
      av[d] = fmaf(p, gg[d], av[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < D; d++) {
    const float a = wsum(ak[d]), b = wsum(av[d]);
    if (lane == (d & 31)) {
      dK[(size_t)j * ldd + ho + d] = a;
      dV[(size_t)j * ldd + ho + d] = b;
    }
  }
}

template <int D>
int launch_attn_bwd(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const float* O, int ldo, const float* dO,
                    int ldg, int Nq, int Nk, int heads, float* stats, float* dQ, float* dK, float* dV, int ldd, cudaStream_t stream) {
  const float scale = 1.f / sqrtf((float)D);
  attn_bwd_q_kernel<D><<<dim3(cdiv(Nq, 8), heads), 256, 0, stream>>>(Q, ldq, K, ldk, V, ldv, O, ldo, dO, ldg, Nq, Nk, scale, dQ, ldd, stats);
  RDM_LAUNCH_CHECK();
  attn_bwd_kv_kernel<D><<<dim3(cdiv(Nk, 8), heads), 256, 0, stream>>>(Q, ldq, K, ldk, V, ldv, dO, ldg, Nq, Nk, scale, stats, dK, dV, ldd);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// -------------------------------------------------------------------------------------------------- Sinkhorn
// One CTA (16 warps) per patch. The padded score matrix Z and the gradient accumulator D (= d Z) live in shared memory with an
// odd row stride (row sweeps and column sweeps are both conflict-free). Phase 1 re-runs the forward iterations with exact
// expf / logf and stores every iterate (u_t, v_t) in the workspace; phase 2 walks them backwards:
//   out = Z + u_T 1^T + 1 v_T^T - norm                      D = G, du = G 1, dv = G^T 1
//   v_t = log_nu - LSE_i(Z + u_t):   Pv = exp(Z + u_t + v_t - log_nu)   D -= Pv diag(dv),  du -= Pv dv
//   u_t = log_mu - LSE_j(Z + v_{t-1}): Pu = exp(Z + u_t + v_{t-1} - log_mu)  D -= diag(du) Pu,  dv_{t-1} = -Pu^T du
// Masked rows / columns (entries -inf: exp = 0 exactly) take no part, as in the forward kernel; the gradient arriving ON masked
// entries is ignored (the reference's losses read them as the constant 1e12: experiments/loss.py:263-271).
#define SKB_THREADS 512
#define SKB_WARPS 16
__global__ void __launch_bounds__(SKB_THREADS) sinkhorn_bwd_kernel(const float* __restrict__ scores, int R, int Cc,
                                                                  const unsigned char* __restrict__ row_masks,
                                                                  const unsigned char* __restrict__ col_masks,
                                                                  const float* __restrict__ alpha_ptr, int iters, float inf,
                                                                  const float* __restrict__ d_out, float* __restrict__ iter_ws,
                                                                  float* __restrict__ d_scores, float* __restrict__ d_alpha_part) {
  extern __shared__ float s_f[];
  const int R1 = R + 1, C1 = Cc + 1;
  const int ld = (C1 % 2 == 0) ? C1 + 1 : C1;
  float* Z = s_f;                    // R1 * ld
  float* Dm = Z + (size_t)R1 * ld;   // R1 * ld
  float* u = Dm + (size_t)R1 * ld;   // R1
  float* v = u + R1;                 // C1   (v_t)
  float* vp = v + C1;                // C1   (v_{t-1})
  float* lmu = vp + C1;              // R1
  float* lnu = lmu + R1;             // C1
  float* du = lnu + C1;              // R1
  float* dv = du + R1;               // C1
  unsigned char* rv = (unsigned char*)(dv + C1);  // R1 row valid
  unsigned char* cv = rv + R1;                    // C1 col valid
  __shared__ int s_nr, s_nc;
  __shared__ float s_red[SKB_WARPS];
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned char* rm = row_masks + (size_t)p * R;
  const unsigned char* cm = col_masks + (size_t)p * Cc;
  const float alpha = *alpha_ptr;
  if (tid == 0) s_nr = 0, s_nc = 0;
  __syncthreads();
  {
    int a = 0, b = 0;
    for (int i = tid; i < R; i += SKB_THREADS) a += rm[i] != 0;
    for (int j = tid; j < Cc; j += SKB_THREADS) b += cm[j] != 0;
    if (a) atomicAdd(&s_nr, a);
    if (b) atomicAdd(&s_nc, b);
  }
  for (int i = tid; i < R1; i += SKB_THREADS) rv[i] = i < R ? (rm[i] != 0) : 1;
  for (int j = tid; j < C1; j += SKB_THREADS) cv[j] = j < Cc ? (cm[j] != 0) : 1;
  __syncthreads();
  const float nr = (float)s_nr, nc = (float)s_nc;
  const float norm = -logf(nr + nc);
  const float* sp = scores + (size_t)p * R * Cc;
  const float* gp = d_out + (size_t)p * R1 * C1;
  for (int e = tid; e < R1 * C1; e += SKB_THREADS) {
    const int i = e / C1, j = e - i * C1;
    const bool ok = rv[i] && cv[j];
    const float z = (i < R && j < Cc) ? sp[(size_t)i * Cc + j] : alpha;
    Z[(size_t)i * ld + j] = ok ? z : -inf;
    Dm[(size_t)i * ld + j] = ok ? gp[e] : 0.f;
  }
  for (int i = tid; i < R1; i += SKB_THREADS) lmu[i] = i < R ? norm : logf(nc) + norm;
  for (int j = tid; j < C1; j += SKB_THREADS) {
    lnu[j] = j < Cc ? norm : logf(nr) + norm;
    v[j] = 0.f;
  }
  __syncthreads();
  float* ws = iter_ws + (size_t)p * iters * (R1 + C1);
  // ---- phase 1: forward iterates
  for (int it = 0; it < iters; it++) {
    for (int i = warp; i < R1; i += SKB_WARPS) {
      if (!rv[i]) {
        if (lane == 0) u[i] = 0.f;
        continue;
      }
      const float* zr = Z + (size_t)i * ld;
      float mx = -3.4e38f;
      for (int j = lane; j < C1; j += 32)
        if (cv[j]) mx = fmaxf(mx, zr[j] + v[j]);
      mx = wmax(mx);
      float s = 0.f;
      for (int j = lane; j < C1; j += 32)
        if (cv[j]) s += expf(zr[j] + v[j] - mx);
      s = wsum(s);
      if (lane == 0) u[i] = lmu[i] - (mx + logf(s));
    }
    __syncthreads();
    for (int j = warp; j < C1; j += SKB_WARPS) {
      if (!cv[j]) {
        if (lane == 0) vp[j] = 0.f;
        continue;
      }
      float mx = -3.4e38f;
      for (int i = lane; i < R1; i += 32)
        if (rv[i]) mx = fmaxf(mx, Z[(size_t)i * ld + j] + u[i]);
      mx = wmax(mx);
      float s = 0.f;
      for (int i = lane; i < R1; i += 32)
        if (rv[i]) s += expf(Z[(size_t)i * ld + j] + u[i] - mx);
      s = wsum(s);
      if (lane == 0) vp[j] = lnu[j] - (mx + logf(s));
    }
    __syncthreads();
    float* w = ws + (size_t)it * (R1 + C1);
    for (int i = tid; i < R1; i += SKB_THREADS) w[i] = u[i];
    for (int j = tid; j < C1; j += SKB_THREADS) {
      w[R1 + j] = vp[j];
      v[j] = vp[j];
    }
    __syncthreads();
  }
  // ---- phase 2: reverse sweep. du / dv start as the row / column sums of G (masked entries already zeroed in Dm)
  for (int i = warp; i < R1; i += SKB_WARPS) {
    float s = 0.f;
    for (int j = lane; j < C1; j += 32) s += Dm[(size_t)i * ld + j];
    s = wsum(s);
    if (lane == 0) du[i] = s;
  }
  for (int j = warp; j < C1; j += SKB_WARPS) {
    float s = 0.f;
    for (int i = lane; i < R1; i += 32) s += Dm[(size_t)i * ld + j];
    s = wsum(s);
    if (lane == 0) dv[j] = s;
  }
  __syncthreads();
  for (int it = iters - 1; it >= 0; it--) {
    const float* w = ws + (size_t)it * (R1 + C1);
    const float* wprev = it > 0 ? ws + (size_t)(it - 1) * (R1 + C1) : nullptr;
    for (int i = tid; i < R1; i += SKB_THREADS) u[i] = w[i];
    for (int j = tid; j < C1; j += SKB_THREADS) {
      v[j] = w[R1 + j];
      vp[j] = wprev ? wprev[R1 + j] : 0.f;
    }
    __syncthreads();
    // v_t step: rows accumulate -Pv dv
    for (int i = warp; i < R1; i += SKB_WARPS) {
      if (!rv[i]) continue;
      float* dr = Dm + (size_t)i * ld;
      const float* zr = Z + (size_t)i * ld;
      const float ui = u[i];
      float acc = 0.f;
      for (int j = lane; j < C1; j += 32) {
        if (!cv[j]) continue;
        const float g = dv[j] * expf(zr[j] + ui + v[j] - lnu[j]);
        dr[j] -= g;
        acc += g;
      }
      acc = wsum(acc);
      if (lane == 0) du[i] -= acc;
    }
    __syncthreads();
    // u_t step: columns collect -Pu^T du
    for (int j = warp; j < C1; j += SKB_WARPS) {
      if (!cv[j]) continue;
      const float vj = vp[j];
      float acc = 0.f;
      for (int i = lane; i < R1; i += 32) {
        if (!rv[i]) continue;
        const float g = du[i] * expf(Z[(size_t)i * ld + j] + vj + u[i] - lmu[i]);
        Dm[(size_t)i * ld + j] -= g;
        acc += g;
      }
      acc = wsum(acc);
      if (lane == 0) vp[j] = -acc;  // dv_{t-1}; vp is re-loaded at the top of the next round
    }
    __syncthreads();
    for (int j = tid; j < C1; j += SKB_THREADS) dv[j] = cv[j] ? vp[j] : 0.f;
    for (int i = tid; i < R1; i += SKB_THREADS) du[i] = 0.f;
    __syncthreads();
  }
  // ---- outputs: d scores = D on the live R x C block, d alpha = sum of D over the dustbin row and column
  float* dsp = d_scores + (size_t)p * R * Cc;
  float a = 0.f;
  for (int e = tid; e < R1 * C1; e += SKB_THREADS) {
    const int i = e / C1, j = e - i * C1;
    const float d = Dm[(size_t)i * ld + j];
    if (i < R && j < Cc)
      dsp[(size_t)i * Cc + j] = d;
    else
      a += d;
  }
  a = wsum(a);
  if (lane == 0) s_red[warp] = a;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w2 = 0; w2 < SKB_WARPS; w2++) t += s_red[w2];
    d_alpha_part[p] = t;
  }
}

}  // namespace

extern "C" int rdm_rope_bwd(const float* x, int ldx, const float* emb, int lde, const float* dy, int ldy, int N, int C, float* dx,
                            float* demb, cudaStream_t stream) {
  RDM_CHECK_ARG(C % 2 == 0 && N >= 0, "rdm_rope_bwd: C must be even");
  if (N == 0) return RDM_OK;
  rope_bwd_kernel<<<cdiv((long long)N * (C / 2), 256), 256, 0, stream>>>(x, ldx, emb, lde, dy, ldy, N, C, dx, demb);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

extern "C" size_t rdm_attention_bwd_workspace(int Nq, int heads) { return (size_t)(Nq > 0 ? Nq : 1) * heads * 3 * sizeof(float); }

extern "C" int rdm_attention_bwd(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const float* O, int ldo,
                                 const float* dO, int ldg, int Nq, int Nk, int heads, int head_dim, void* workspace,
                                 size_t workspace_bytes, float* dQ, float* dK, float* dV, int ldd, cudaStream_t stream) {
  RDM_CHECK_ARG(Nq >= 0 && Nk >= 1 && heads >= 1, "rdm_attention_bwd: bad shape");
  RDM_CHECK_ARG(head_dim == 16 || head_dim == 32, "rdm_attention_bwd: head_dim must be 16 or 32 (got %d)", head_dim);
  RDM_CHECK_ARG(ldk % 4 == 0 && ldv % 4 == 0 && ldq % 4 == 0 && ldg % 4 == 0, "rdm_attention_bwd: row strides must be multiples of 4 floats");
  RDM_CHECK_ARG(((uintptr_t)K | (uintptr_t)V | (uintptr_t)Q | (uintptr_t)dO) % 16 == 0, "rdm_attention_bwd: operands must be 16-byte aligned");
  if (workspace_bytes < rdm_attention_bwd_workspace(Nq, heads)) {
    rdm_set_error("rdm_attention_bwd: workspace too small");
    return RDM_ERR_WORKSPACE;
  }
  if (Nq == 0) {  // no query: the keys and values receive no gradient
    RDM_CUDA(cudaMemset2DAsync(dK, (size_t)ldd * sizeof(float), 0, (size_t)heads * head_dim * sizeof(float), Nk, stream));
    RDM_CUDA(cudaMemset2DAsync(dV, (size_t)ldd * sizeof(float), 0, (size_t)heads * head_dim * sizeof(float), Nk, stream));
    return RDM_OK;
  }
  float* stats = (float*)workspace;
  if (head_dim == 16)
    return launch_attn_bwd<16>(Q, ldq, K, ldk, V, ldv, O, ldo, dO, ldg, Nq, Nk, heads, stats, dQ, dK, dV, ldd, stream);
  return launch_attn_bwd<32>(Q, ldq, K, ldk, V, ldv, O, ldo, dO, ldg, Nq, Nk, heads, stats, dQ, dK, dV, ldd, stream);
}

extern "C" size_t rdm_sinkhorn_bwd_workspace(int num_patches, int R, int C, int num_iterations) {
  return (size_t)(num_patches > 0 ? num_patches : 1) * ((size_t)num_iterations * (R + C + 2) + 1) * sizeof(float);
}

extern "C" int rdm_sinkhorn_bwd(const float* scores, int num_patches, int R, int C, const unsigned char* row_masks,
                                const unsigned char* col_masks, const float* alpha, int num_iterations, float inf, const float* d_out,
                                void* workspace, size_t workspace_bytes, float* d_scores, float* d_alpha_partial, cudaStream_t stream) {
  RDM_CHECK_ARG(num_patches >= 0 && R >= 1 && C >= 1 && num_iterations >= 1, "rdm_sinkhorn_bwd: bad shape");
  if (num_patches == 0) return RDM_OK;
  if (workspace_bytes < rdm_sinkhorn_bwd_workspace(num_patches, R, C, num_iterations)) {
    rdm_set_error("rdm_sinkhorn_bwd: workspace too small");
    return RDM_ERR_WORKSPACE;
  }
  const int R1 = R + 1, C1 = C + 1, ld = (C1 % 2 == 0) ? C1 + 1 : C1;
  const size_t smem = ((size_t)2 * R1 * ld + 3 * R1 + 5 * C1) * sizeof(float) + align_up((size_t)R1 + C1, 16);
  RDM_CHECK_ARG(smem <= 220 * 1024, "rdm_sinkhorn_bwd: patch %d x %d too large for shared memory", R, C);
  RDM_CUDA(cudaFuncSetAttribute(sinkhorn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sinkhorn_bwd_kernel<<<num_patches, SKB_THREADS, smem, stream>>>(scores, R, C, row_masks, col_masks, alpha, num_iterations, inf, d_out,
                                                                  (float*)workspace, d_scores, d_alpha_partial);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}
