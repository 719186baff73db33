// Fused ThDRoFormer layer kernels for the RDMNet configuration (d_model = 128, 4 heads x 32, FFN 256).
//
// Reference semantics (one TransformerLayer / RPETransformerLayer):
//   rdmnet/thdroformer/thdroformer.py:108-139 (RPEMultiHeadAttention: q,k,v Linear, RoPE on q,k, softmax(qk^T/sqrt(32))v),
//   :163-172 (RPEAttentionLayer: Linear + LayerNorm(residual)), geotransformer/modules/transformer/vanilla_transformer.py:
//   31-70, 92-102 (same without RoPE), geotransformer/modules/transformer/output_layer.py:15-21 (FFN + LayerNorm).
//
// The sequences are <= ~450 superpoints: the unfused path spends its time in ~25 launches per layer. Here a layer is
// two launches, and independent problems (ref/src self-attention, the q / k,v projections of a cross layer) share one:
//   tf_project_kernel : Y = rope?(X W^T + b) for up to 6 (input, weight) jobs; CTA = 8 rows x 128 outputs.
//   tf_attend_kernel  : CTA = 8 query rows, all 4 heads: flash-style attention over K/V tiles staged in shared
//                       memory, then out-projection + residual + LayerNorm + FFN(ReLU) + residual + LayerNorm for
//                       those rows - everything after the attention is row-local.
// Weights are read from a per-layer blob of TRANSPOSED matrices (k-major), so that the 128 threads of a row block
// read 128 consecutive floats per k (coalesced, L2-resident) while the activations are broadcast from shared memory.
#include "common.cuh"
#include "../../include/rdm_sm100.h"

#define TF_D 128
#define TF_H 4
#define TF_HD 32
#define TF_F 256
#define TF_R 8        // rows per CTA
#define TF_KT 64      // keys per tile

// layer blob layout (floats), all matrices k-major ([in][out])
#define TFB_WQ 0
#define TFB_WK (TFB_WQ + TF_D * TF_D)
#define TFB_WV (TFB_WK + TF_D * TF_D)
#define TFB_WO (TFB_WV + TF_D * TF_D)
#define TFB_W1 (TFB_WO + TF_D * TF_D)          // [128][256]
#define TFB_W2 (TFB_W1 + TF_D * TF_F)          // [256][128]
#define TFB_BQ (TFB_W2 + TF_F * TF_D)
#define TFB_BK (TFB_BQ + TF_D)
#define TFB_BV (TFB_BK + TF_D)
#define TFB_BO (TFB_BV + TF_D)
#define TFB_B1 (TFB_BO + TF_D)                 // 256
#define TFB_B2 (TFB_B1 + TF_F)
#define TFB_G1 (TFB_B2 + TF_D)
#define TFB_E1 (TFB_G1 + TF_D)
#define TFB_G2 (TFB_E1 + TF_D)
#define TFB_E2 (TFB_G2 + TF_D)
#define TFB_SIZE (TFB_E2 + TF_D)

struct ProjJobs {
  rdm_tf_proj_job j[6];
};
struct AttnJobs {
  rdm_tf_attn_job j[2];
};

// ------------------------------------------------------------------------------------------------ projection
__device__ __forceinline__ void tfp_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
// grid (ceil(maxN/8), num_jobs), 256 threads: thread (o = tid & 127, rh = tid >> 7) -> output o of rows rh*4..rh*4+3.
// The job's 64 KB weight matrix is copied to shared memory with cp.async BEFORE the programmatic-dependency wait (it is
// constant data), i.e. while the previous kernel is still running; the multiply then reads it conflict-free.
__global__ void __launch_bounds__(256) tf_project_kernel(const ProjJobs jobs) {
  extern __shared__ __align__(16) float tfp_smem[];  // W [128][128] k-major, then xt [128][8]
  float* Ws = tfp_smem;
  float* xt = tfp_smem + TF_D * TF_D;
  const rdm_tf_proj_job jb = jobs.j[blockIdx.y];
  const int n0 = blockIdx.x * TF_R;
  pdl_trigger();
  if (n0 >= jb.n) return;
  const int tid = threadIdx.x, o = tid & 127, rh = tid >> 7;
#pragma unroll
  for (int i = 0; i < TF_D * TF_D / 4 / 256; i++) tfp_cp16(Ws + 4 * (tid + i * 256), jb.wt + 4 * (tid + i * 256));
  asm volatile("cp.async.commit_group;" ::: "memory");
  pdl_wait();
  for (int e = tid; e < TF_R * TF_D; e += 256) {
    int r = e >> 7, c = e & 127;
    xt[c * TF_R + r] = (n0 + r < jb.n) ? jb.x[(size_t)(n0 + r) * jb.ldx + c] : 0.f;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
  for (int k = 0; k < TF_D; k++) {
    const float w = Ws[k * TF_D + o];
    const float4 x = *(const float4*)(xt + k * TF_R + rh * 4);
    acc[0] = fmaf(w, x.x, acc[0]); acc[1] = fmaf(w, x.y, acc[1]); acc[2] = fmaf(w, x.z, acc[2]); acc[3] = fmaf(w, x.w, acc[3]);
  }
  const float b = jb.bias[o];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int n = n0 + rh * 4 + i;
    float v = acc[i] + b;
    if (jb.emb != nullptr) {  // RoPE: thdroformer.py:56-85; channel pair (2p, 2p+1) rotates by theta(emb[n][p])
      const float other = __shfl_xor_sync(FULL_MASK, v, 1);
      if (n < jb.n) {
        const float em = jb.emb[(size_t)n * jb.lde + (o >> 1)];
        const float theta = (1.f / (1.f + expf(-em))) * 3.14159265359f * 2.f;
        float s, c;
        sincosf(theta, &s, &c);
        v = (o & 1) ? fmaf(v, c, other * s) : fmaf(v, c, -other * s);
      }
    }
    if (n < jb.n) {
      if (jb.ldy_t == 0) jb.y[(size_t)n * TF_D + o] = v;      // row-major [n][128]
      else jb.y[(size_t)o * jb.ldy_t + n] = v;                // channel-major [128][ldy_t] (keys of rdm_tf_attend)
    }
  }
}

extern "C" int rdm_tf_project(const rdm_tf_proj_job* h_jobs, int num_jobs, cudaStream_t stream) {
  RDM_CHECK_ARG(h_jobs != nullptr && num_jobs >= 1 && num_jobs <= 6, "rdm_tf_project: 1..6 jobs");
  ProjJobs pj;
  int maxn = 0;
  for (int i = 0; i < num_jobs; i++) {
    pj.j[i] = h_jobs[i];
    RDM_CHECK_ARG(h_jobs[i].n >= 0 && h_jobs[i].ldx >= TF_D && (h_jobs[i].ldy_t == 0 || h_jobs[i].ldy_t >= h_jobs[i].n),
                  "rdm_tf_project: bad job %d", i);
    maxn = max(maxn, h_jobs[i].n);
  }
  if (maxn == 0) return RDM_OK;
  const size_t smem = (size_t)(TF_D * TF_D + TF_D * TF_R) * sizeof(float);
  RDM_CUDA(cudaFuncSetAttribute(tf_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  // per device
  RDM_CUDA(rdm_launch_pdl(tf_project_kernel, dim3(cdiv(maxn, TF_R), num_jobs), dim3(256), smem, stream, pj));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// ------------------------------------------------------------------------------------------------ attention + post
#define TF_WBUF 16384  // floats per weight staging buffer (64 KB)
struct AttnSmem {
  float Ks[TF_D * TF_KT];       // [channel][key]: lane-per-key reads and float4 tile stores are both conflict free
  float Vs[TF_KT * TF_D];       // [key][channel]
  float Qt[TF_D * TF_R];        // [channel][row], pre-scaled by 1/sqrt(32)
  float Ps[8][TF_KT * 4];       // per warp: [key][4 queries]
  float At[TF_D * TF_R];        // attention output / x1, transposed [channel][row]
  float Y[TF_R * TF_D];         // row-major scratch for LayerNorm
  float X1[TF_R * TF_D];        // x1 row-major (residual of the FFN)
  float Ft[TF_F * TF_R];        // FFN hidden, transposed
  float W0[TF_WBUF];            // weight staging ring (cp.async): the 320 KB of out-proj / FFN weights stream through these
  float W1[TF_WBUF];            //   two buffers in six 64 KB chunks, each chunk landing while the previous one is consumed
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// all 256 threads: asynchronous copy of one 64 KB weight chunk (16 x 16 B per thread) + commit
__device__ __forceinline__ void stage_chunk(float* dst, const float* __restrict__ src, int tid) {
#pragma unroll
  for (int i = 0; i < TF_WBUF / 4 / 256; i++) cp_async16(dst + 4 * (tid + i * 256), src + 4 * (tid + i * 256));
  cp_async_commit();
}

// acc[i] += sum_{k < K} Ws[k][o] * xt[k][r0 + i], i < 4; Ws is a staged weight chunk in shared memory (row stride ldw)
template <int K>
__device__ __forceinline__ void rowblock_gemv4_s(const float* Ws, int ldw, int o, const float* xt, int r0, float acc[4]) {
#pragma unroll 8
  for (int k = 0; k < K; k++) {
    const float w = Ws[k * ldw + o];
    const float4 x = *(const float4*)(xt + k * TF_R + r0);
    acc[0] = fmaf(w, x.x, acc[0]); acc[1] = fmaf(w, x.y, acc[1]); acc[2] = fmaf(w, x.z, acc[2]); acc[3] = fmaf(w, x.w, acc[3]);
  }
}

// LayerNorm of the 8 rows in sm.Y (one warp per row): writes row-major to dst_rm (shared or global, stride ld) and,
// if dst_t != nullptr, transposed to dst_t[c*8 + row].
__device__ __forceinline__ void layernorm_rows(const float* Y, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float* dst_rm, int ld, float* dst_t,
                                               int nvalid, int warp, int lane) {
  const int r = warp;
  const float4 v = *(const float4*)(Y + r * TF_D + 4 * lane);
  const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.f / TF_D);
  const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
  const float var = warp_sum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3) * (1.f / TF_D);
  const float rstd = rsqrtf(var + 1e-5f);
  const float4 g = *(const float4*)(gamma + 4 * lane), b = *(const float4*)(beta + 4 * lane);
  const float4 y = make_float4(d0 * rstd * g.x + b.x, d1 * rstd * g.y + b.y, d2 * rstd * g.z + b.z, d3 * rstd * g.w + b.w);
  if (r < nvalid) *(float4*)(dst_rm + (size_t)r * ld + 4 * lane) = y;
  if (dst_t != nullptr) {
    dst_t[(4 * lane + 0) * TF_R + r] = y.x;
    dst_t[(4 * lane + 1) * TF_R + r] = y.y;
    dst_t[(4 * lane + 2) * TF_R + r] = y.z;
    dst_t[(4 * lane + 3) * TF_R + r] = y.w;
  }
}

// grid (ceil(maxNq/8), num_jobs), 256 threads = 8 warps; warp w -> head w & 3, queries (w >> 2) * 4 .. + 3.
// Latency plan: K/V tile t+1 is prefetched into registers while tile t is consumed; the out-projection and the first
// half of the FFN-expand weights are copied to shared memory with cp.async during the whole attention phase, and every
// later 64 KB weight chunk lands while the previous one is being multiplied.
__global__ void __launch_bounds__(256) tf_attend_kernel(const AttnJobs jobs) {
  const rdm_tf_attn_job jb = blockIdx.y == 0 ? jobs.j[0] : jobs.j[1];
  const int n0 = blockIdx.x * TF_R;
  if (n0 >= jb.nq) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AttnSmem& sm = *reinterpret_cast<AttnSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int head = warp & 3, qh = warp >> 2;
  const int nvalid = min(TF_R, jb.nq - n0);
  const float scale = 0.17677669529663687f;  // 1/sqrt(32)
  const float* B = jb.blob;
  pdl_trigger();
  stage_chunk(sm.W0, B + TFB_WO, tid);  // group 0: out-projection [128][128]
  stage_chunk(sm.W1, B + TFB_W1, tid);  // group 1: FFN expand, k = 0..63 of [128][256]
  pdl_wait();  // the weight copies above touch constants only: they overlap the projection kernel still running
  for (int e = tid; e < TF_R * TF_D; e += 256) {
    int r = e >> 7, c = e & 127;
    sm.Qt[c * TF_R + r] = (r < nvalid) ? jb.q[(size_t)(n0 + r) * TF_D + c] * scale : 0.f;
  }
  float m[4], l[4], acc[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    m[i] = -3.0e38f;
    l[i] = 0.f;
    acc[i] = 0.f;
  }
  const int hoff = head * TF_HD;
  float* Pw = sm.Ps[warp];
  float4 kreg[8], vreg[8];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int e = tid + i * 256;
      const int c = e >> 4, j4 = (e & 15) * 4;  // K tile: channel-major in global (ld = ldk_t) and in smem
      kreg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + j4 + 3 < jb.ldk_t) kreg[i] = __ldg((const float4*)(jb.k + (size_t)c * jb.ldk_t + k0 + j4));  // cols >= nk: masked below
      const int j = e >> 5, c4 = (e & 31) * 4;
      vreg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + j < jb.nk) vreg[i] = __ldg((const float4*)(jb.v + (size_t)(k0 + j) * TF_D + c4));
    }
  };
  load_tile(0);
  for (int k0 = 0; k0 < jb.nk; k0 += TF_KT) {
    __syncthreads();  // previous tile fully consumed (and Qt visible on the first pass)
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int e = tid + i * 256;
      *(float4*)(sm.Ks + (e >> 4) * TF_KT + (e & 15) * 4) = kreg[i];
      *(float4*)(sm.Vs + (e >> 5) * TF_D + (e & 31) * 4) = vreg[i];
    }
    __syncthreads();
    if (k0 + TF_KT < jb.nk) load_tile(k0 + TF_KT);  // in flight during this tile's math
    // scores of keys (lane, lane+32) against the warp's 4 queries
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
    const float* ka = sm.Ks + hoff * TF_KT + lane;
    const float* qp = sm.Qt + hoff * TF_R + qh * 4;
#pragma unroll 8
    for (int d = 0; d < TF_HD; d++) {
      const float4 q4 = *(const float4*)(qp + d * TF_R);
      const float a = ka[d * TF_KT], b = ka[d * TF_KT + 32];
      s0[0] = fmaf(q4.x, a, s0[0]); s0[1] = fmaf(q4.y, a, s0[1]); s0[2] = fmaf(q4.z, a, s0[2]); s0[3] = fmaf(q4.w, a, s0[3]);
      s1[0] = fmaf(q4.x, b, s1[0]); s1[1] = fmaf(q4.y, b, s1[1]); s1[2] = fmaf(q4.z, b, s1[2]); s1[3] = fmaf(q4.w, b, s1[3]);
    }
    const bool ok0 = k0 + lane < jb.nk, ok1 = k0 + lane + 32 < jb.nk;
    float p0[4], p1[4], corr[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float a = ok0 ? s0[i] : -3.0e38f, b = ok1 ? s1[i] : -3.0e38f;
      const float mnew = fmaxf(m[i], warp_max(fmaxf(a, b)));
      corr[i] = __expf(m[i] - mnew);
      p0[i] = ok0 ? __expf(a - mnew) : 0.f;
      p1[i] = ok1 ? __expf(b - mnew) : 0.f;
      l[i] = l[i] * corr[i] + warp_sum(p0[i] + p1[i]);
      m[i] = mnew;
      acc[i] *= corr[i];
    }
    __syncwarp();
    *(float4*)(Pw + lane * 4) = make_float4(p0[0], p0[1], p0[2], p0[3]);
    *(float4*)(Pw + (lane + 32) * 4) = make_float4(p1[0], p1[1], p1[2], p1[3]);
    __syncwarp();
    // acc[i] (channel hoff + lane) += sum_j p[i][j] * V[j][hoff + lane]
    const float* vp = sm.Vs + hoff + lane;
#pragma unroll 8
    for (int j = 0; j < TF_KT; j++) {
      const float4 p4 = *(const float4*)(Pw + j * 4);
      const float v = vp[j * TF_D];
      acc[0] = fmaf(p4.x, v, acc[0]); acc[1] = fmaf(p4.y, v, acc[1]); acc[2] = fmaf(p4.z, v, acc[2]); acc[3] = fmaf(p4.w, v, acc[3]);
    }
  }
  // attention output, transposed: At[channel][row]
  *(float4*)(sm.At + (hoff + lane) * TF_R + qh * 4) =
      make_float4(acc[0] / l[0], acc[1] / l[1], acc[2] / l[2], acc[3] / l[3]);
  cp_async_wait<1>();  // group 0 (out-projection weights) has landed for this thread ...
  __syncthreads();     // ... and for everyone; At complete
  const int o = tid & 127, rh = tid >> 7;
  // out-projection + bias + residual(input rows) -> Y
  {
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    rowblock_gemv4_s<TF_D>(sm.W0, TF_D, o, sm.At, rh * 4, a4);
    const float b = B[TFB_BO + o];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int r = rh * 4 + i;
      const float res = (r < nvalid) ? jb.x[(size_t)(n0 + r) * jb.ldx + o] : 0.f;
      sm.Y[r * TF_D + o] = a4[i] + b + res;
    }
  }
  __syncthreads();                                        // W0 free, Y complete
  stage_chunk(sm.W0, B + TFB_W1 + TF_WBUF, tid);          // group 2: FFN expand, k = 64..127
  layernorm_rows(sm.Y, B + TFB_G1, B + TFB_E1, sm.X1, TF_D, sm.At, TF_R, warp, lane);  // x1 (row-major + transposed)
  cp_async_wait<1>();                                     // group 1
  __syncthreads();
  // FFN expand 128 -> 256, ReLU: thread tid -> hidden unit tid, all 8 rows as two groups of 4; k split over two chunks
  float a4[4] = {0.f, 0.f, 0.f, 0.f}, c4[4] = {0.f, 0.f, 0.f, 0.f};
  rowblock_gemv4_s<64>(sm.W1, TF_F, tid, sm.At, 0, a4);
  rowblock_gemv4_s<64>(sm.W1, TF_F, tid, sm.At, 4, c4);
  __syncthreads();                                        // W1 free
  stage_chunk(sm.W1, B + TFB_W2, tid);                    // group 3: FFN squeeze, k = 0..127 of [256][128]
  cp_async_wait<1>();                                     // group 2
  __syncthreads();
  rowblock_gemv4_s<64>(sm.W0, TF_F, tid, sm.At + 64 * TF_R, 0, a4);
  rowblock_gemv4_s<64>(sm.W0, TF_F, tid, sm.At + 64 * TF_R, 4, c4);
  {
    const float b = B[TFB_B1 + tid];
    *(float4*)(sm.Ft + tid * TF_R) = make_float4(fmaxf(a4[0] + b, 0.f), fmaxf(a4[1] + b, 0.f), fmaxf(a4[2] + b, 0.f), fmaxf(a4[3] + b, 0.f));
    *(float4*)(sm.Ft + tid * TF_R + 4) = make_float4(fmaxf(c4[0] + b, 0.f), fmaxf(c4[1] + b, 0.f), fmaxf(c4[2] + b, 0.f), fmaxf(c4[3] + b, 0.f));
  }
  __syncthreads();                                        // W0 free, Ft complete
  stage_chunk(sm.W0, B + TFB_W2 + TF_WBUF, tid);          // group 4: FFN squeeze, k = 128..255
  cp_async_wait<1>();                                     // group 3
  __syncthreads();
  // FFN squeeze 256 -> 128 + bias + residual(x1) -> Y
  {
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
    rowblock_gemv4_s<128>(sm.W1, TF_D, o, sm.Ft, rh * 4, s4);
    cp_async_wait<0>();                                   // group 4
    __syncthreads();
    rowblock_gemv4_s<128>(sm.W0, TF_D, o, sm.Ft + 128 * TF_R, rh * 4, s4);
    const float b = B[TFB_B2 + o];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int r = rh * 4 + i;
      sm.Y[r * TF_D + o] = s4[i] + b + sm.X1[r * TF_D + o];
    }
  }
  __syncthreads();
  layernorm_rows(sm.Y, B + TFB_G2, B + TFB_E2, jb.out + (size_t)n0 * TF_D, TF_D, nullptr, nvalid, warp, lane);
}

extern "C" size_t rdm_tf_layer_blob_floats(void) { return (size_t)TFB_SIZE; }

extern "C" int rdm_tf_attend(const rdm_tf_attn_job* h_jobs, int num_jobs, cudaStream_t stream) {
  RDM_CHECK_ARG(h_jobs != nullptr && num_jobs >= 1 && num_jobs <= 2, "rdm_tf_attend: 1..2 jobs");
  AttnJobs aj;
  int maxn = 0;
  for (int i = 0; i < num_jobs; i++) {
    aj.j[i] = h_jobs[i];
    RDM_CHECK_ARG(h_jobs[i].nq >= 0 && h_jobs[i].nk >= 1 && h_jobs[i].ldx >= TF_D && h_jobs[i].ldk_t >= h_jobs[i].nk &&
                      h_jobs[i].ldk_t % 4 == 0, "rdm_tf_attend: bad job %d", i);
    maxn = max(maxn, h_jobs[i].nq);
  }
  if (maxn == 0) return RDM_OK;
  RDM_CUDA(cudaFuncSetAttribute(tf_attend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AttnSmem)));  // per device
  RDM_CUDA(rdm_launch_pdl(tf_attend_kernel, dim3(cdiv(maxn, TF_R), num_jobs), dim3(256), sizeof(AttnSmem), stream, aj));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}
