// Dense building blocks of the backbone / transformer: fp32 GEMM (Linear, KPConv weight contraction),
// GroupNorm(+residual)(+LeakyReLU) over stacked (N, C) features, LayerNorm(+residual)(+ReLU).
//
// Reference semantics: torch.nn.Linear; geotransformer/modules/kpconv/modules.py:33-50 (GroupNorm over (1,C,N), eps 1e-5,
// statistics joint over both clouds), :78-83, :205-225 (residual + LeakyReLU(0.1)); torch.nn.LayerNorm (eps 1e-5).
#include <stdlib.h>
#include "common.cuh"
#include "../../include/rdm_sm100.h"

// ------------------------------------------------------------------------------------------------------- GEMM
// C[M,N] = A[M,K] * B + bias.  B_NK: B is [N,K] row-major (nn.Linear weight), else [K,N] row-major (KPConv weights).
// 256 threads, BK = 16, each thread owns a (TM x TN) micro-tile split in 4-wide halves to keep LDS.128 conflict-free.
#define GEMM_BK 16

template <int BM, int BN, bool B_NK>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                                    int ldb, const float* __restrict__ bias, float* __restrict__ C,
                                                    int ldc, int M, int N, int K, int k_per_split, int vecA, int vecB,
                                                    int vecC, int act) {
  pdl_trigger();
  pdl_wait();
  constexpr int TM = BM / 16, TN = BN / 16;  // 8x8 (128 tile) or 4x4 (64 tile)
  constexpr int PAD = 4;
  __shared__ __align__(16) float As[2][GEMM_BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][GEMM_BK][BN + PAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb = blockIdx.z * k_per_split, ke = min(K, kb + k_per_split);
  if (gridDim.z > 1) C += (size_t)blockIdx.z * M * N;  // split-K partial buffers (ldc == N there)

  // loader geometry for "row-major with K contiguous" operands (A always, B when B_NK): rows x 4 float4
  constexpr int A_PASSES = BM / 64, BNK_PASSES = BN / 64;
  // loader for B [K,N]: 16 rows x (BN/4) float4
  constexpr int BKN_PER_THREAD = (GEMM_BK * BN / 4) / 256;
  float4 ra[A_PASSES], rb[B_NK ? BNK_PASSES : BKN_PER_THREAD];

  auto load_rowmajor_k = [&](const float* P, int ld, int row, int rows_total, int k, int vec) -> float4 {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows_total) {
      const float* p = P + (size_t)row * ld + k;
      if (vec && k + 3 < ke) {
        v = *(const float4*)p;
      } else {
        if (k < ke) v.x = p[0];
        if (k + 1 < ke) v.y = p[1];
        if (k + 2 < ke) v.z = p[2];
        if (k + 3 < ke) v.w = p[3];
      }
    }
    return v;
  };
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int p = 0; p < A_PASSES; p++) {
      int row = (tid >> 2) + p * 64, kq = (tid & 3) * 4;
      ra[p] = load_rowmajor_k(A, lda, m0 + row, M, k0 + kq, vecA);
    }
    if constexpr (B_NK) {
#pragma unroll
      for (int p = 0; p < BNK_PASSES; p++) {
        int row = (tid >> 2) + p * 64, kq = (tid & 3) * 4;
        rb[p] = load_rowmajor_k(B, ldb, n0 + row, N, k0 + kq, vecB);
      }
    } else {
#pragma unroll
      for (int p = 0; p < BKN_PER_THREAD; p++) {
        int e = tid + p * 256;
        int kr = e / (BN / 4), nq = (e % (BN / 4)) * 4;
        int k = k0 + kr, n = n0 + nq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < ke) {
          const float* pB = B + (size_t)k * ldb + n;
          if (vecB && n + 3 < N) {
            v = *(const float4*)pB;
          } else {
            if (n < N) v.x = pB[0];
            if (n + 1 < N) v.y = pB[1];
            if (n + 2 < N) v.z = pB[2];
            if (n + 3 < N) v.w = pB[3];
          }
        }
        rb[p] = v;
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int p = 0; p < A_PASSES; p++) {
      int row = (tid >> 2) + p * 64, kq = (tid & 3) * 4;
      As[buf][kq + 0][row] = ra[p].x;
      As[buf][kq + 1][row] = ra[p].y;
      As[buf][kq + 2][row] = ra[p].z;
      As[buf][kq + 3][row] = ra[p].w;
    }
    if constexpr (B_NK) {
#pragma unroll
      for (int p = 0; p < BNK_PASSES; p++) {
        int row = (tid >> 2) + p * 64, kq = (tid & 3) * 4;
        Bs[buf][kq + 0][row] = rb[p].x;
        Bs[buf][kq + 1][row] = rb[p].y;
        Bs[buf][kq + 2][row] = rb[p].z;
        Bs[buf][kq + 3][row] = rb[p].w;
      }
    } else {
#pragma unroll
      for (int p = 0; p < BKN_PER_THREAD; p++) {
        int e = tid + p * 256;
        int kr = e / (BN / 4), nq = (e % (BN / 4)) * 4;
        *(float4*)&Bs[buf][kr][nq] = rb[p];
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  int nk = (ke - kb + GEMM_BK - 1) / GEMM_BK;
  if (nk > 0) {
    load_tiles(kb);
    store_tiles(0);
  }
  __syncthreads();
  for (int t = 0; t < nk; t++) {
    int buf = t & 1;
    if (t + 1 < nk) load_tiles(kb + (t + 1) * GEMM_BK);
#pragma unroll
    for (int k = 0; k < GEMM_BK; k++) {
      float a[TM], b[TN];
      if constexpr (TM == 8) {
        float4 a0 = *(const float4*)&As[buf][k][ty * 4], a1 = *(const float4*)&As[buf][k][BM / 2 + ty * 4];
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      } else {
        float4 a0 = *(const float4*)&As[buf][k][ty * 4];
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      }
      if constexpr (TN == 8) {
        float4 b0 = *(const float4*)&Bs[buf][k][tx * 4], b1 = *(const float4*)&Bs[buf][k][BN / 2 + tx * 4];
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
      } else {
        float4 b0 = *(const float4*)&Bs[buf][k][tx * 4];
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      }
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }
  // epilogue
  const bool add_bias = (bias != nullptr) && gridDim.z == 1;
#pragma unroll
  for (int i = 0; i < TM; i++) {
    int row = m0 + ((TM == 8 && i >= 4) ? BM / 2 + ty * 4 + (i - 4) : ty * 4 + i);
    if (row >= M) continue;
#pragma unroll
    for (int jh = 0; jh < TN / 4; jh++) {
      int col = n0 + (jh == 0 ? tx * 4 : BN / 2 + tx * 4);
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        v[j] = acc[i][jh * 4 + j];
        if (add_bias && col + j < N) v[j] += bias[col + j];
        if (gridDim.z == 1) {
          if (act == 1) v[j] = v[j] > 0.f ? v[j] : 0.1f * v[j];
          else if (act == 2) v[j] = fmaxf(v[j], 0.f);
        }
      }
      float* cp = C + (size_t)row * ldc + col;
      if (vecC && col + 3 < N) {
        *(float4*)cp = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (col + j < N) cp[j] = v[j];
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, const float* __restrict__ bias,
                                     float* __restrict__ C, int ldc, int M, int N, int act) {
  pdl_trigger();
  pdl_wait();
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)M * N) return;
  int m = (int)(e / N), n = (int)(e - (long long)m * N);
  float s = 0.f;
  for (int z = 0; z < splits; z++) s += part[(size_t)z * M * N + e];
  if (bias) s += bias[n];
  if (act == 1) s = s > 0.f ? s : 0.1f * s;
  else if (act == 2) s = fmaxf(s, 0.f);
  C[(size_t)m * ldc + n] = s;
}

// Split-K reduction with the GroupNorm statistics of the result fused in (N % 32 == 0, cpg a power of two).
// Block = 8 warps on a 64-row x 32-column panel: lane = column, so the column sums accumulate in registers with no
// shuffles; one shared-memory fold over the 8 warps, one segmented shuffle fold over the cpg columns of a group and
// one pair of double atomics per group per block.
__global__ void __launch_bounds__(256) splitk_reduce_stats_kernel(const float* __restrict__ part, int splits,
                                                                  const float* __restrict__ bias, float* __restrict__ C,
                                                                  int ldc, int M, int N, double* __restrict__ stats,
                                                                  int cpg, int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_s[8][32], s_q[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane, r0 = blockIdx.y * rows_per_block;
  const float b = bias ? bias[col] : 0.f;
  float cs = 0.f, cq = 0.f;
  for (int r = r0 + warp; r < min(M, r0 + rows_per_block); r += 8) {
    const size_t e = (size_t)r * N + col;
    float v = 0.f;
    for (int z = 0; z < splits; z++) v += part[(size_t)z * M * N + e];
    v += b;
    C[(size_t)r * ldc + col] = v;
    cs += v;
    cq = fmaf(v, v, cq);
  }
  s_s[warp][lane] = cs;
  s_q[warp][lane] = cq;
  __syncthreads();
  if (warp == 0) {
    cs = cq = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) {
      cs += s_s[w][lane];
      cq += s_q[w][lane];
    }
    const int span = cpg < 32 ? cpg : 32;
    for (int o = 1; o < span; o <<= 1) {
      cs += __shfl_xor_sync(FULL_MASK, cs, o);
      cq += __shfl_xor_sync(FULL_MASK, cq, o);
    }
    if ((lane & (span - 1)) == 0) {
      const int g = col / cpg;
      atomicAdd(&stats[2 * g], (double)cs);
      atomicAdd(&stats[2 * g + 1], (double)cq);
    }
  }
}

static inline int aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

extern "C" size_t rdm_linear_workspace(int M, int N, int K) {
  // worst case: 16 split-K partial buffers
  return (size_t)16 * M * N * sizeof(float) + 256;
}
// [K,N]-layout B on the tensor cores: room for the transposed copy in front of the split-K partials (small outputs only)
extern "C" size_t rdm_linear_kn_workspace(int M, int N, int K) {
  const size_t t = align_up((size_t)N * ((K + 3) & ~3) * sizeof(float), 256);
  return t + ((long long)M * N <= (1 << 20) ? rdm_linear_workspace(M, N, K) : 0);
}
int rdm_transpose_ld(const float* x, int rows, int cols, int ldx, float* y, int ldy, cudaStream_t stream);  // backward.cu

int rdm_linear_tc(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int M, int N, int K,
                  int act, void* workspace, size_t workspace_bytes, int* out_splits, double* gn_stats, int gn_cpg,
                  int* out_stats_fused, cudaStream_t stream);  // gemm_tc.cu

// experimental A-in-TMEM variant of the same kernel (gemm_tc_atmem.cu): -1 unless RDM_GEMM_ATMEM=1
int rdm_linear_tc_fast(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int M, int N, int K, int act,
                       void* workspace, size_t workspace_bytes, int* out_splits, double* gn_stats, int gn_cpg, int* out_stats_fused,
                       cudaStream_t stream);
int rdm_linear_tc_atmem(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int M, int N, int K,
                        int act, void* workspace, size_t workspace_bytes, int* out_splits, double* gn_stats, int gn_cpg,
                        int* out_stats_fused, const float* B_split, cudaStream_t stream);

// Registry of pre-split constant weights (weight pointer -> [2][N][ld] tf32 hi / lo copy). ONLY the module runners look
// weights up here (runtime.cu `linear`): they are entered through the host caches that re-register after any weight change
// (rdmnet_b200/modules.py cache_key); the public rdm_linear never does, so an in-place weight edit followed by a direct
// operator call cannot read a stale split.
#include <unordered_map>
static std::unordered_map<const void*, const float*> g_presplit;
extern "C" int rdm_presplit_register(const float* weight, const float* split) {
  if (split == nullptr) g_presplit.erase(weight);
  else g_presplit[weight] = split;
  return RDM_OK;
}
extern "C" void rdm_presplit_clear(void) { g_presplit.clear(); }
const float* rdm_presplit_lookup(const float* weight) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("RDM_GEMM_PRESPLIT");  // A/B knob: 0 ignores the registered splits
    on = (e && e[0] == '0') ? 0 : 1;
  }
  if (!on) return nullptr;
  auto it = g_presplit.find(weight);
  return it == g_presplit.end() ? nullptr : it->second;
}

extern "C" int rdm_linear(const float* A, int lda, const float* B, int ldb, int b_is_nk, const float* bias, float* C,
                          int ldc, int M, int N, int K, int act, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  static int use_registry = -1;  // RDM_LINEAR_USE_REGISTRY=1 (micro-benchmarks only): the operator consults the registry too
  if (use_registry < 0) {
    const char* e = getenv("RDM_LINEAR_USE_REGISTRY");
    use_registry = (e && e[0] == '1') ? 1 : 0;
  }
  return rdm_linear_gn_ps(A, lda, B, ldb, b_is_nk, bias, C, ldc, M, N, K, act, workspace, workspace_bytes, nullptr, 0, nullptr,
                          (use_registry && b_is_nk) ? rdm_presplit_lookup(B) : nullptr, stream);
}

// rdm_linear + optional fused GroupNorm statistics of the output: when the shape qualifies, {sum, sumsq} of every group
// (cpg consecutive output channels) are accumulated into gn_stats (pre-zeroed doubles) by the GEMM epilogue and
// *stats_fused = 1; otherwise *stats_fused = 0 and the caller runs the stand-alone statistics pass.
int rdm_linear_gn(const float* A, int lda, const float* B, int ldb, int b_is_nk, const float* bias, float* C, int ldc, int M,
                  int N, int K, int act, void* workspace, size_t workspace_bytes, double* gn_stats, int gn_cpg,
                  int* stats_fused, cudaStream_t stream) {
  return rdm_linear_gn_ps(A, lda, B, ldb, b_is_nk, bias, C, ldc, M, N, K, act, workspace, workspace_bytes, gn_stats, gn_cpg, stats_fused,
                          nullptr, stream);
}

int rdm_linear_gn_ps(const float* A, int lda, const float* B, int ldb, int b_is_nk, const float* bias, float* C, int ldc, int M,
                     int N, int K, int act, void* workspace, size_t workspace_bytes, double* gn_stats, int gn_cpg,
                     int* stats_fused, const float* B_split, cudaStream_t stream) {
  RDM_CHECK_ARG(M >= 0 && N >= 1 && K >= 1, "rdm_linear: bad shape M=%d N=%d K=%d", M, N, K);
  if (stats_fused) *stats_fused = 0;
  if (M == 0) return RDM_OK;
  // tensor-core path (tcgen05 kind::tf32, 3-term split, fp32-level accuracy) for nn.Linear-layout weights
  static int use_tc = -1;
  if (use_tc < 0) {
    const char* e = getenv("RDM_GEMM_TC");  // debug knob: RDM_GEMM_TC=0 forces the SIMT kernel
    use_tc = (e && e[0] == '0') ? 0 : 1;
  }
  // [K,N]-layout weights (KPConv weights in the per-operator / training path, the dx and dW products of rdm_linear_bwd): when the
  // caller's workspace has room, transpose B once into its head ([N, Kp], Kp = K rounded up to 4) and run the tensor-core
  // kernels on that; the tail of the workspace stays available for split-K partials.
  if (use_tc && !b_is_nk && M >= 64 && N >= 8 && K >= 8 && workspace != nullptr) {
    const int Kp = (K + 3) & ~3;
    const size_t need = align_up((size_t)N * Kp * sizeof(float), 256);
    if (workspace_bytes >= need && (lda % 4 == 0) && aligned16(A) && aligned16(workspace)) {
      float* Bt = (float*)workspace;
      int rc = rdm_transpose_ld(B, K, N, ldb, Bt, Kp, stream);
      if (rc != RDM_OK) return rc;
      return rdm_linear_gn_ps(A, lda, Bt, Kp, 1, bias, C, ldc, M, N, K, act, (char*)workspace + need, workspace_bytes - need, gn_stats,
                              gn_cpg, stats_fused, nullptr, stream);
    }
  }
  if (use_tc && b_is_nk && M >= 64) {
    int tc_splits = 1;
    // reduced-precision mode (rdm_set_precision(1)): one tf32 product per k-step, no operand split (gemm_tc_fast.cu)
    int rc = rdm_linear_tc_fast(A, lda, B, ldb, bias, C, ldc, M, N, K, act, workspace, workspace_bytes, &tc_splits, gn_stats, gn_cpg,
                                stats_fused, stream);
    if (rc == -1)
      rc = rdm_linear_tc_atmem(A, lda, B, ldb, bias, C, ldc, M, N, K, act, workspace, workspace_bytes, &tc_splits, gn_stats,
                               gn_cpg, stats_fused, B_split, stream);
    if (rc == -1)
      rc = rdm_linear_tc(A, lda, B, ldb, bias, C, ldc, M, N, K, act, workspace, workspace_bytes, &tc_splits, gn_stats, gn_cpg,
                         stats_fused, stream);
    if (rc == RDM_OK && tc_splits > 1) {
      const bool pow2 = gn_cpg >= 1 && (gn_cpg & (gn_cpg - 1)) == 0 && (gn_cpg <= 32 || gn_cpg % 32 == 0);
      if (gn_stats != nullptr && act == 0 && N % 32 == 0 && pow2) {
        int rpb = 64;  // rows per block: small problems get more, shorter blocks (the pass is latency-bound)
        while (rpb > 8 && (long long)(N / 32) * cdiv(M, rpb) < 296) rpb >>= 1;
        RDM_CUDA(rdm_launch_pdl(splitk_reduce_stats_kernel, dim3(N / 32, cdiv(M, rpb)), dim3(256), 0, stream, (const float*)workspace,
                                tc_splits, bias, C, ldc, M, N, gn_stats, gn_cpg, rpb));
        if (stats_fused) *stats_fused = 1;
      } else {
        RDM_CUDA(rdm_launch_pdl(splitk_reduce_kernel, dim3(cdiv((long long)M * N, 256)), dim3(256), 0, stream, (const float*)workspace,
                                tc_splits, bias, C, ldc, M, N, act));
      }
      RDM_LAUNCH_CHECK();
    }
    if (rc != -1) return rc;
  }
  int vecA = (lda % 4 == 0) && aligned16(A), vecB = (ldb % 4 == 0) && aligned16(B), vecC = (ldc % 4 == 0) && aligned16(C);
  long long t128 = (long long)cdiv(M, 128) * cdiv(N, 128), t64 = (long long)cdiv(M, 64) * cdiv(N, 64);
  bool big = t128 >= 120;
  int splits = 1;
  if (!big && t64 < 120 && workspace != nullptr) {
    splits = (int)min((long long)16, max((long long)1, (296 + t64 - 1) / t64));
    splits = min(splits, max(1, K / 256));
    while (splits > 1 && (size_t)splits * M * N * sizeof(float) > workspace_bytes) splits--;
  }
  int kps = cdiv(cdiv(K, splits), GEMM_BK) * GEMM_BK;
  splits = cdiv(K, kps);
  float* out = splits > 1 ? (float*)workspace : C;
  int ldo = splits > 1 ? N : ldc;
  int vecO = splits > 1 ? ((N % 4 == 0) && aligned16(workspace)) : vecC;
  if (big) {
    dim3 grid(cdiv(N, 128), cdiv(M, 128), 1);
    if (b_is_nk)
      RDM_CUDA(rdm_launch_pdl(sgemm_kernel<128, 128, true>, grid, dim3(256), 0, stream, A, lda, B, ldb, bias, out, ldo, M, N, K, kps, vecA, vecB,
                              vecO, act));
    else
      RDM_CUDA(rdm_launch_pdl(sgemm_kernel<128, 128, false>, grid, dim3(256), 0, stream, A, lda, B, ldb, bias, out, ldo, M, N, K, kps, vecA, vecB,
                              vecO, act));
  } else {
    dim3 grid(cdiv(N, 64), cdiv(M, 64), splits);
    if (b_is_nk)
      RDM_CUDA(rdm_launch_pdl(sgemm_kernel<64, 64, true>, grid, dim3(256), 0, stream, A, lda, B, ldb, bias, out, ldo, M, N, K, kps, vecA, vecB,
                              vecO, act));
    else
      RDM_CUDA(rdm_launch_pdl(sgemm_kernel<64, 64, false>, grid, dim3(256), 0, stream, A, lda, B, ldb, bias, out, ldo, M, N, K, kps, vecA, vecB,
                              vecO, act));
  }
  RDM_LAUNCH_CHECK();
  if (splits > 1) {
    splitk_reduce_kernel<<<cdiv((long long)M * N, 256), 256, 0, stream>>>((const float*)workspace, splits, bias, C, ldc, M, N, act);
    RDM_LAUNCH_CHECK();
  }
  return RDM_OK;
}

// ------------------------------------------------------------------------------------------------- GroupNorm
// stats[g] = {sum, sumsq} (double) over rows x (C/G) channels of group g.
__global__ void __launch_bounds__(256) groupnorm_stats_kernel(const float* __restrict__ x, int N, int C, int G,
                                                              int rows_per_cta, double* __restrict__ stats) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_part[];  // [2*C]
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
  for (int c = tid; c < C; c += 256) {
    float s = 0.f, ss = 0.f;
    for (int r = r0; r < r1; r++) {
      float v = x[(size_t)r * C + c];
      s += v;
      ss = fmaf(v, v, ss);
    }
    s_part[c] = s;
    s_part[C + c] = ss;
  }
  __syncthreads();
  const int cpg = C / G;
  for (int g = tid; g < G; g += 256) {
    double s = 0.0, ss = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; c++) {
      s += (double)s_part[c];
      ss += (double)s_part[C + c];
    }
    atomicAdd(&stats[2 * g], s);
    atomicAdd(&stats[2 * g + 1], ss);
  }
}

// y = act( (x - mean_g) * rstd_g * gamma_c + beta_c (+ res) ),  act: 0 none, 1 LeakyReLU(slope)
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const float* __restrict__ x,
                                                              const double* __restrict__ stats,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              const float* __restrict__ res, float* __restrict__ y, int N,
                                                              int C, int G, float eps, int act, float slope) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_ms[];  // mean[G], rstd[G]
  const int tid = threadIdx.x;
  const int cpg = C / G;
  for (int g = tid; g < G; g += 256) {
    double cnt = (double)N * cpg;
    double mean = stats[2 * g] / cnt;
    double var = stats[2 * g + 1] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_ms[g] = (float)mean;
    s_ms[G + g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  long long total = (long long)N * C;
  for (long long e = blockIdx.x * 256LL + tid; e < total; e += (long long)gridDim.x * 256) {
    int c = (int)(e % C), g = c / cpg;
    float v = (x[e] - s_ms[g]) * s_ms[G + g] * gamma[c] + beta[c];
    if (res) v += res[e];
    if (act == 1) v = v > 0.f ? v : v * slope;
    y[e] = v;
  }
}

// Vectorised apply: LPR lanes own one row (LPR = C/4 for C <= 128, else 32 lanes striding 128 channels), float4 loads and
// stores, per-channel {mean, rstd*gamma, beta} staged once per CTA in shared memory. The lanes of a row also have its
// channel sum for free, so the kernel can emit rowpos[r] = (sum_c y[r,c] > 0): the neighbour-count predicate of the
// KPConv that consumes y (kpconv.py:113-114) - no separate pass over the features.
template <int LPR>
__global__ void __launch_bounds__(256) groupnorm_apply_vec_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ res, float* __restrict__ y, int N, int C,
                                                                  int G, float eps, int act, float slope,
                                                                  unsigned char* __restrict__ rowpos) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_par[];  // mean[C], scale[C], beta[C]
  float *s_mean = s_par, *s_scale = s_par + C, *s_beta = s_par + 2 * C;
  const int tid = threadIdx.x, cpg = C / G;
  for (int c = tid; c < C; c += 256) {
    const int g = c / cpg;
    const double cnt = (double)N * cpg;
    const double mean = stats[2 * g] / cnt;
    double var = stats[2 * g + 1] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[c] = (float)mean;
    s_scale[c] = (float)(1.0 / sqrt(var + (double)eps)) * gamma[c];
    s_beta[c] = beta[c];
  }
  __syncthreads();
  constexpr int RPW = 32 / LPR;  // rows per warp
  const int lane = tid & 31, t = lane % LPR, sub = lane / LPR;
  const int warps = (gridDim.x * 256) >> 5, gw = (blockIdx.x * 256 + tid) >> 5;
  for (int r0 = gw * RPW; r0 < N; r0 += warps * RPW) {
    const int r = r0 + sub;
    float rs = 0.f;
    if (r < N) {
      for (int c = 4 * t; c < C; c += 4 * LPR) {
        const size_t e = (size_t)r * C + c;
        const float4 v = *(const float4*)(x + e);
        const float4 m = *(const float4*)(s_mean + c), sc = *(const float4*)(s_scale + c), b = *(const float4*)(s_beta + c);
        float4 o = make_float4((v.x - m.x) * sc.x + b.x, (v.y - m.y) * sc.y + b.y, (v.z - m.z) * sc.z + b.z,
                               (v.w - m.w) * sc.w + b.w);
        if (res != nullptr) {
          const float4 q = *(const float4*)(res + e);
          o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
        }
        if (act == 1) {
          o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
          o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
        }
        *(float4*)(y + e) = o;
        rs += (o.x + o.y) + (o.z + o.w);
      }
    }
    if (rowpos != nullptr) {
#pragma unroll
      for (int o = LPR >> 1; o > 0; o >>= 1) rs += __shfl_xor_sync(FULL_MASK, rs, o);
      if (t == 0 && r < N) rowpos[r] = rs > 0.f ? 1 : 0;
    }
  }
}

int rdm_groupnorm_stats(const float* x, int N, int C, int groups, double* stats_zeroed, cudaStream_t stream) {
  if (N == 0) return RDM_OK;
  int rows = 64;
  while (rows > 4 && cdiv(N, rows) < 296) rows >>= 1;
  RDM_CUDA(rdm_launch_pdl(groupnorm_stats_kernel, dim3(cdiv(N, rows)), dim3(256), 2 * C * sizeof(float), stream, x, N, C, groups, rows,
                          stats_zeroed));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

int rdm_groupnorm_apply(const float* x, const double* stats, const float* gamma, const float* beta, const float* residual,
                        float* y, int N, int C, int groups, float eps, int act, float slope, unsigned char* rowpos_out,
                        cudaStream_t stream) {
  if (N == 0) return RDM_OK;
  const bool al = (((uintptr_t)x | (uintptr_t)y | (uintptr_t)residual) & 15) == 0;
  const int lpr = C <= 128 ? C / 4 : 32;
  if (al && C % 4 == 0 && (lpr == 8 || lpr == 16 || lpr == 32) && (C <= 128 || C % 128 == 0) && C <= 4096) {
    const int rpw = 32 / lpr;
    const long long warps_needed = cdiv(N, rpw);
    const int grid = (int)min((long long)148 * 4, (warps_needed + 7) / 8);
    const size_t smem = 3 * (size_t)C * sizeof(float);
#define GN_APPLY(LPRv)                                                                                                     \
  RDM_CUDA(rdm_launch_pdl(groupnorm_apply_vec_kernel<LPRv>, dim3(grid), dim3(256), smem, stream, x, stats, gamma, beta, residual, y, N, \
                          C, groups, eps, act, slope, rowpos_out))
    if (lpr == 8) GN_APPLY(8);
    else if (lpr == 16) GN_APPLY(16);
    else GN_APPLY(32);
#undef GN_APPLY
    RDM_LAUNCH_CHECK();
    return RDM_OK;
  }
  RDM_CHECK_ARG(rowpos_out == nullptr, "rdm_groupnorm_apply: row-positivity output needs the vectorised path (C=%d)", C);
  long long total = (long long)N * C;
  int grid = (int)min((long long)148 * 8, (total + 255) / 256);
  groupnorm_apply_kernel<<<grid, 256, 2 * groups * sizeof(float), stream>>>(x, stats, gamma, beta, residual, y, N, C, groups, eps,
                                                                           act, slope);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

extern "C" int rdm_groupnorm(const float* x, const float* gamma, const float* beta, const float* residual, float* y,
                             int N, int C, int groups, float eps, int act, float slope, double* stats_scratch,
                             cudaStream_t stream) {
  RDM_CHECK_ARG(N >= 0 && C >= 1 && groups >= 1 && C % groups == 0, "rdm_groupnorm: C=%d not divisible by groups=%d", C, groups);
  RDM_CHECK_ARG(groups <= 1024 && C <= 6144, "rdm_groupnorm: shape too large");
  if (N == 0) return RDM_OK;
  RDM_CUDA(cudaMemsetAsync(stats_scratch, 0, sizeof(double) * 2 * groups, stream));
  if (int rc = rdm_groupnorm_stats(x, N, C, groups, stats_scratch, stream)) return rc;
  return rdm_groupnorm_apply(x, stats_scratch, gamma, beta, residual, y, N, C, groups, eps, act, slope, nullptr, stream);
}

// ------------------------------------------------------------------------------------------------- LayerNorm
// y[r] = act( LN(x[r] + res[r]) * gamma + beta ), one warp per row. act: 0 none, 2 ReLU.
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ y, int N, int C, float eps, int act) {
  pdl_trigger();
  pdl_wait();
  int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* xr = x + (size_t)row * C;
  const float* rr = res ? res + (size_t)row * C : nullptr;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c] + (rr ? rr[c] : 0.f);
  float mean = warp_sum(s) / (float)C;
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) {
    float d = xr[c] + (rr ? rr[c] : 0.f) - mean;
    ss = fmaf(d, d, ss);
  }
  float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
  for (int c = lane; c < C; c += 32) {
    float v = (xr[c] + (rr ? rr[c] : 0.f) - mean) * rstd * gamma[c] + beta[c];
    if (act == 2) v = fmaxf(v, 0.f);
    y[(size_t)row * C + c] = v;
  }
}

extern "C" int rdm_layernorm(const float* x, const float* residual, const float* gamma, const float* beta, float* y,
                             int N, int C, float eps, int act, cudaStream_t stream) {
  RDM_CHECK_ARG(N >= 0 && C >= 1, "rdm_layernorm: bad shape");
  if (N == 0) return RDM_OK;
  RDM_CUDA(rdm_launch_pdl(layernorm_kernel, dim3(cdiv(N, 8)), dim3(256), 0, stream, x, residual, gamma, beta, y, N, C, eps, act));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// ------------------------------------------------------------------------------------------------- small elementwise
// y = act(x): 1 LeakyReLU(slope), 2 ReLU, 3 sigmoid clamped to [0,1]
__global__ void activation_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int act, float slope) {
  pdl_trigger();
  pdl_wait();
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= n) return;
  float v = x[e];
  if (act == 1) v = v > 0.f ? v : v * slope;
  else if (act == 2) v = fmaxf(v, 0.f);
  else if (act == 3) v = fminf(fmaxf(1.f / (1.f + expf(-v)), 0.f), 1.f);
  y[e] = v;
}

extern "C" int rdm_activation(const float* x, float* y, int64_t n, int act, float slope, cudaStream_t stream) {
  if (n <= 0) return RDM_OK;
  activation_kernel<<<cdiv(n, 256), 256, 0, stream>>>(x, y, n, act, slope);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}
