// Training path (BASELINE config 4): backward kernels of the KPConv backbone operators. The
// reference trains through PyTorch autograd over its ~15-launch KPConv / GroupNorm / index_select graphs
// (geotransformer/modules/kpconv/kpconv.py:79-122, modules.py:33-225, functional.py:6-67); here every operator has one
// hand-written backward kernel behind the C ABI, wired into autograd by rdmnet_b200/autograd.py.
//   rdm_kpconv_gather_bwd   d s_feats  of  A[m,k,:] = (1/cnt_m) sum_h w[m,h,k] F[idx[m,h],:]   (sparse in k, scatter-add)
//   rdm_transpose           [R,C] -> [C,R]   (dW = dY^T X, dX = dY W through rdm_linear)
//   rdm_colsum              bias gradient
//   rdm_groupnorm_bwd       GroupNorm (+ residual add, + LeakyReLU) backward: dx, dgamma, dbeta, dres
//   rdm_layernorm_bwd       LayerNorm (+ residual, + ReLU) backward
//   rdm_maxpool_bwd / rdm_upsample_concat_bwd     scatter of the selected rows
//   rdm_activation_bwd      LeakyReLU / ReLU / clamp(sigmoid)
// Gradients are fp32; scatter-adds use float atomics (order-dependent in the last bits, like cuDNN / ATen index_add).
#include "common.cuh"
#include "../../include/rdm_sm100.h"

#define KP_K 15
namespace {
struct KPtsB {
  float x[16], y[16], z[16];
};

// ---- KPConv gather backward. One CTA (256 threads) per query, like the forward's CTA-cooperative sparse kernel:
// (A) thread = neighbour slot: 15 influences -> shared tile (+ the query's positive-neighbour count);
// (B) warp w takes slots w, w+8, ...: for each slot the non-zero kernel points (~1.7 of 15) contribute
//     dF[j, c] += inv_cnt * sum_k w_k dA[m, k, c]; lanes stride the channels; one atomicAdd per (slot, channel).
template <typename IdxT>
__global__ void __launch_bounds__(256) kpconv_gather_bwd_kernel(const float* __restrict__ dA, const unsigned char* __restrict__ rowpos,
                                                                const float* __restrict__ q_pts, const float* __restrict__ s_pts,
                                                                const IdxT* __restrict__ idx, const KPtsB kp, float inv_sigma, int M,
                                                                int N, int H, int C, float* __restrict__ dF) {
  extern __shared__ float s_w[];  // [H][16]
  __shared__ int s_j[256];
  __shared__ int s_npos[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, m = blockIdx.x;
  const float qx = q_pts[3 * (size_t)m], qy = q_pts[3 * (size_t)m + 1], qz = q_pts[3 * (size_t)m + 2];
  int np = 0;
  for (int h = tid; h < H; h += 256) {  // H <= 256 in practice: one pass
    int j = -1;
    const long long jj = (long long)idx[(size_t)m * H + h];
    if (jj < N) j = (int)jj;
    s_j[h] = j;
    float* w = s_w + (size_t)h * 16;
    if (j >= 0) {
      const float dx = s_pts[3 * (size_t)j] - qx, dy = s_pts[3 * (size_t)j + 1] - qy, dz = s_pts[3 * (size_t)j + 2] - qz;
#pragma unroll
      for (int k = 0; k < KP_K; k++) {
        const float ex = dx - kp.x[k], ey = dy - kp.y[k], ez = dz - kp.z[k];
        w[k] = fmaxf(0.f, 1.f - sqrtf(ex * ex + ey * ey + ez * ez) * inv_sigma);
      }
      np += rowpos[j];
    } else {
#pragma unroll
      for (int k = 0; k < KP_K; k++) w[k] = 0.f;
    }
    w[15] = 0.f;
  }
  np = warp_sum_i(np);
  if (lane == 0) s_npos[warp] = np;
  __syncthreads();
  int npos = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) npos += s_npos[i];
  const float inv = 1.f / (float)max(npos, 1);
  const float* dAm = dA + (size_t)m * KP_K * C;
  for (int h = warp; h < H; h += 8) {
    const int j = s_j[h];
    if (j < 0) continue;
    const float* w = s_w + (size_t)h * 16;
    unsigned nz = 0;
#pragma unroll
    for (int k = 0; k < KP_K; k++) nz |= (w[k] > 0.f ? 1u : 0u) << k;
    if (nz == 0) continue;
    for (int c = lane; c < C; c += 32) {
      float acc = 0.f;
      unsigned mk = nz;
      while (mk) {
        const int k = __ffs(mk) - 1;
        mk &= mk - 1;
        acc = fmaf(w[k], dAm[(size_t)k * C + c], acc);
      }
      atomicAdd(&dF[(size_t)j * C + c], acc * inv);
    }
  }
}

__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ x, int R, int C, int ldx, float* __restrict__ y,
                                                        int ldy) {
  __shared__ float t[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    if (by + i < R && bx + tx < C) t[i][tx] = x[(size_t)(by + i) * ldx + bx + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (bx + i < C && by + tx < R) y[(size_t)(bx + i) * ldy + by + tx] = t[tx][i];
}

// out[c] (+)= sum_r x[r, c]: grid (ceil(C/32), row chunks), fp32 partial sums, one atomicAdd per (CTA, column)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int R, int C, int ldx, int rows_per_cta,
                                                     float* __restrict__ out) {
  __shared__ float s[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
  float a = 0.f;
  if (c < C)
    for (int r = r0 + ty; r < r1; r += 8) a += x[(size_t)r * ldx + c];
  s[ty][threadIdx.x & 31] = a;
  __syncthreads();
  if (ty == 0 && c < C) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) v += s[i][threadIdx.x];
    atomicAdd(&out[c], v);
  }
}

// ---- GroupNorm backward. y = act(xh * gamma + beta (+ res)), xh = (x - mu_g) * rstd_g over the (N x cpg) slab of group g.
// dz = dy * act'(.) (LeakyReLU: the sign of y is the sign of the pre-activation); dres = dz;
// pass 1 (column sums): dgamma_c = sum_n dz xh, dbeta_c = sum_n dz  (doubles, atomics per CTA)
// pass 2: per group a = mean(dz gamma) , b = mean(dz gamma xh) from the column sums; dx = rstd (dz gamma - a - xh b).
__global__ void __launch_bounds__(256) gn_bwd_colsums_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                             const float* __restrict__ dy, const double* __restrict__ stats, int N, int C,
                                                             int G, float eps, int act, float slope, int rows_per_cta,
                                                             double* __restrict__ dgamma, double* __restrict__ dbeta,
                                                             float* __restrict__ dz_out) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  __shared__ double s1[8][33], s2[8][33];
  const int cpg = C / G;
  double a1 = 0.0, a2 = 0.0;
  if (c < C) {
    const int g = c / cpg;
    const double cnt = (double)N * cpg, mean = stats[2 * g] / cnt;
    double var = stats[2 * g + 1] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float mu = (float)mean, rstd = (float)(1.0 / sqrt(var + (double)eps));
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
    for (int r = r0 + ty; r < r1; r += 8) {
      const size_t e = (size_t)r * C + c;
      float dz = dy[e];
      if (act == 1 && !(y[e] > 0.f)) dz *= slope;
      if (dz_out != nullptr) dz_out[e] = dz;
      const float xh = (x[e] - mu) * rstd;
      a1 += (double)(dz * xh);
      a2 += (double)dz;
    }
  }
  s1[ty][threadIdx.x & 31] = a1;
  s2[ty][threadIdx.x & 31] = a2;
  __syncthreads();
  if (ty == 0 && c < C) {
    double v1 = 0.0, v2 = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      v1 += s1[i][threadIdx.x];
      v2 += s2[i][threadIdx.x];
    }
    atomicAdd(&dgamma[c], v1);
    atomicAdd(&dbeta[c], v2);
  }
}

__global__ void __launch_bounds__(256) gn_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ dz, const double* __restrict__ stats,
                                                        const float* __restrict__ gamma, const double* __restrict__ dgamma,
                                                        const double* __restrict__ dbeta, int N, int C, int G, float eps,
                                                        float* __restrict__ dx) {
  extern __shared__ float s_par[];  // mu[G], rstd[G], a[G], b[G]
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += 256) {
    const double cnt = (double)N * cpg, mean = stats[2 * g] / cnt;
    double var = stats[2 * g + 1] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    double a = 0.0, b = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; c++) {
      a += dbeta[c] * (double)gamma[c];   // sum dz gamma
      b += dgamma[c] * (double)gamma[c];  // sum dz gamma xh
    }
    s_par[g] = (float)mean;
    s_par[G + g] = (float)(1.0 / sqrt(var + (double)eps));
    s_par[2 * G + g] = (float)(a / cnt);
    s_par[3 * G + g] = (float)(b / cnt);
  }
  __syncthreads();
  const long long total = (long long)N * C;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int c = (int)(e % C), g = c / cpg;
    const float rstd = s_par[G + g], xh = (x[e] - s_par[g]) * rstd;
    dx[e] = rstd * (dz[e] * gamma[c] - s_par[2 * G + g] - xh * s_par[3 * G + g]);
  }
}

// ---- LayerNorm backward, one warp per row: y = act(LN(x (+ res)) * gamma + beta), act 2 = ReLU.
// dx (= d of the pre-norm sum, which is also dres) and per-column dgamma / dbeta (atomics).
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ y,
                                                     const float* __restrict__ dy, const float* __restrict__ gamma, int N, int C, float eps,
                                                     int act, float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= N) return;
  const float* xr = x + (size_t)r * C;
  const float* rr = res ? res + (size_t)r * C : nullptr;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c] + (rr ? rr[c] : 0.f);
  const float mu = warp_sum(s) / C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = xr[c] + (rr ? rr[c] : 0.f) - mu;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  float a = 0.f, b = 0.f;
  for (int c = lane; c < C; c += 32) {
    float dz = dy[(size_t)r * C + c];
    if (act == 2 && !(y[(size_t)r * C + c] > 0.f)) dz = 0.f;
    const float xh = (xr[c] + (rr ? rr[c] : 0.f) - mu) * rstd, g = dz * gamma[c];
    a += g;
    b = fmaf(g, xh, b);
    atomicAdd(&dgamma[c], dz * xh);
    atomicAdd(&dbeta[c], dz);
  }
  a = warp_sum(a) / C;
  b = warp_sum(b) / C;
  for (int c = lane; c < C; c += 32) {
    float dz = dy[(size_t)r * C + c];
    if (act == 2 && !(y[(size_t)r * C + c] > 0.f)) dz = 0.f;
    const float xh = (xr[c] + (rr ? rr[c] : 0.f) - mu) * rstd;
    dx[(size_t)r * C + c] = rstd * (dz * gamma[c] - a - xh * b);
  }
}

// ---- maxpool backward: the winning neighbour of (m, c) is recomputed (first maximum, as torch.max); padding rows are the
// zero row of functional.py:64 and receive nothing.
template <typename IdxT>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ f, const IdxT* __restrict__ idx, const float* __restrict__ dout,
                                                          int M, int N, int H, int C, float* __restrict__ df) {
  const long long e = blockIdx.x * 256LL + threadIdx.x;
  if (e >= (long long)M * C) return;
  const int m = (int)(e / C), c = (int)(e - (long long)m * C);
  float best = -3.4e38f;
  long long bj = -1;
  for (int h = 0; h < H; h++) {
    const long long j = (long long)idx[(size_t)m * H + h];
    if (j > N) continue;  // column beyond the reference's row width
    const float v = j < N ? f[(size_t)j * C + c] : 0.f;
    if (v > best) {
      best = v;
      bj = j < N ? j : -1;
    }
  }
  if (bj >= 0) atomicAdd(&df[(size_t)bj * C + c], dout[e]);
}

template <typename IdxT>
__global__ void __launch_bounds__(256) upsample_concat_bwd_kernel(const float* __restrict__ dout, const IdxT* __restrict__ idx, int idx_stride,
                                                                  int M, int N, int C1, int C2, float* __restrict__ dx,
                                                                  float* __restrict__ dskip) {
  const long long e = blockIdx.x * 256LL + threadIdx.x;
  const int C = C1 + C2;
  if (e >= (long long)M * C) return;
  const int m = (int)(e / C), c = (int)(e - (long long)m * C);
  if (c < C1) {
    const long long j = (long long)idx[(size_t)m * idx_stride];
    if (j < N) atomicAdd(&dx[(size_t)j * C1 + c], dout[e]);
  } else if (dskip != nullptr) {
    dskip[(size_t)m * C2 + (c - C1)] = dout[e];
  }
}

__global__ void __launch_bounds__(256) cast_d2f_kernel(const double* __restrict__ a, float* __restrict__ b, int n) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) b[i] = (float)a[i];
}

__global__ void __launch_bounds__(256) activation_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, long long n, int act,
                                                             float slope, float* __restrict__ dx) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i >= n) return;
  const float v = y[i], g = dy[i];
  float d;
  if (act == 1) d = v > 0.f ? g : g * slope;
  else if (act == 2) d = v > 0.f ? g : 0.f;
  else d = (v > 0.f && v < 1.f) ? g * v * (1.f - v) : 0.f;  // clamp(sigmoid(x), 0, 1): y (1 - y) inside the open interval
  dx[i] = d;
}
}  // namespace

extern "C" int rdm_kpconv_gather_bwd(const float* d_weighted, const float* s_feats, const float* q_points, const float* s_points,
                                     const void* neighbor_indices, int index_bytes, const float* h_kernel_points, float sigma, int M, int N,
                                     int H, int C_in, unsigned char* rowpos_scratch, float* d_s_feats, cudaStream_t stream) {
  RDM_CHECK_ARG(M >= 0 && N >= 1 && H >= 1 && C_in >= 1 && sigma > 0.f, "rdm_kpconv_gather_bwd: bad arguments");
  RDM_CHECK_ARG(index_bytes == 4 || index_bytes == 8, "rdm_kpconv_gather_bwd: index_bytes must be 4 or 8");
  RDM_CUDA(cudaMemsetAsync(d_s_feats, 0, (size_t)N * C_in * sizeof(float), stream));
  if (M == 0) return RDM_OK;
  // the forward's neighbour-count predicate (sum_c F[n, c] > 0; C_in == 1: the feature itself)
  int rc = rdm_row_positive(s_feats, N, C_in, rowpos_scratch, stream);
  if (rc != RDM_OK) return rc;
  KPtsB kp;
  for (int k = 0; k < KP_K; k++) {
    kp.x[k] = h_kernel_points[3 * k];
    kp.y[k] = h_kernel_points[3 * k + 1];
    kp.z[k] = h_kernel_points[3 * k + 2];
  }
  kp.x[15] = kp.y[15] = kp.z[15] = 1.0e6f;
  const size_t smem = (size_t)H * 16 * sizeof(float);
  RDM_CHECK_ARG(H <= 256 && smem <= 48 * 1024, "rdm_kpconv_gather_bwd: H = %d too wide", H);
  if (index_bytes == 8)
    kpconv_gather_bwd_kernel<int64_t><<<M, 256, smem, stream>>>(d_weighted, rowpos_scratch, q_points, s_points,
                                                               (const int64_t*)neighbor_indices, kp, 1.f / sigma, M, N, H, C_in, d_s_feats);
  else
    kpconv_gather_bwd_kernel<int><<<M, 256, smem, stream>>>(d_weighted, rowpos_scratch, q_points, s_points, (const int*)neighbor_indices, kp,
                                                           1.f / sigma, M, N, H, C_in, d_s_feats);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

int rdm_transpose_ld(const float* x, int rows, int cols, int ldx, float* y, int ldy, cudaStream_t stream) {
  RDM_CHECK_ARG(rows >= 0 && cols >= 0 && ldx >= cols && ldy >= rows, "rdm_transpose: bad shape");
  if (rows == 0 || cols == 0) return RDM_OK;
  RDM_CHECK_ARG(cdiv(rows, 32) <= 65535, "rdm_transpose: more than 2M rows");  // grid.y limit
  transpose_kernel<<<dim3(cdiv(cols, 32), cdiv(rows, 32)), 256, 0, stream>>>(x, rows, cols, ldx, y, ldy);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}
extern "C" int rdm_transpose(const float* x, int rows, int cols, int ldx, float* y, cudaStream_t stream) {
  return rdm_transpose_ld(x, rows, cols, ldx, y, rows, stream);
}

// ---- Linear backward in ONE call: y = x W^T (+ b) with W [N,K] (w_is_nk) or y = x W with W [K,N]; x [M,K], dy [M,N].
//   dx [M,K] = dy W          db [N] = column sums of dy          dW = dy^T x  ([N,K])  or  x^T dy  ([K,N])
// All three products run on the tensor-core GEMM: the operand that has the contraction index as its SLOW index is transposed
// into the workspace first (dy^T and x^T for dW: the contraction runs over the M points; W^T for dx when W is [N,K]).
extern "C" size_t rdm_linear_bwd_workspace(int M, int N, int K) {
  const size_t Mp = (size_t)((M + 3) & ~3);
  size_t t = align_up((size_t)N * Mp * 4, 256) + align_up((size_t)K * Mp * 4, 256) + align_up((size_t)N * ((K + 3) & ~3) * 4, 256) +
             align_up((size_t)K * ((N + 3) & ~3) * 4, 256);
  return t + align_up((size_t)16 * N * K * 4, 256) + ((long long)M * K <= (1 << 20) ? (size_t)16 * M * K * 4 : 0) + 1024;
}

extern "C" int rdm_linear_bwd(const float* x, const float* w, int w_is_nk, const float* dy, int M, int N, int K, float* dx, float* dw,
                              float* db_zeroed, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  RDM_CHECK_ARG(M >= 0 && N >= 1 && K >= 1, "rdm_linear_bwd: bad shape");
  if (workspace_bytes < rdm_linear_bwd_workspace(M, N, K)) {
    rdm_set_error("rdm_linear_bwd: workspace too small");
    return RDM_ERR_WORKSPACE;
  }
  if (M == 0) {
    if (dw) RDM_CUDA(cudaMemsetAsync(dw, 0, (size_t)N * K * sizeof(float), stream));
    return RDM_OK;
  }
  Workspace ws(workspace, workspace_bytes);
  const int Mp = (M + 3) & ~3, Np = (N + 3) & ~3;
  int rc;
  if (db_zeroed) {
    rc = rdm_colsum(dy, M, N, N, db_zeroed, stream);
    if (rc != RDM_OK) return rc;
  }
  if (dx) {
    // dx = dy [M,N] . B with B as [out = K, inner = N]
    const float* B = w;
    int ldb = N;
    if (w_is_nk) {  // W [N,K] -> W^T [K, Np]
      float* wt = ws.get<float>((size_t)K * Np);
      // (the padding columns N..Np are never read: the GEMM's inner extent is N, the tensor maps zero-fill beyond it)
      rc = rdm_transpose_ld(w, N, K, K, wt, Np, stream);
      if (rc != RDM_OK) return rc;
      B = wt;
      ldb = Np;
    }
    const size_t part = (long long)M * K <= (1 << 20) ? (size_t)16 * M * K * 4 : 0;
    void* p = part ? (void*)ws.get<char>(part) : nullptr;
    rc = rdm_linear_gn(dy, N, B, ldb, 1, nullptr, dx, K, M, K, N, 0, p, part, nullptr, 0, nullptr, stream);
    if (rc != RDM_OK) return rc;
  }
  if (dw) {
    float* dyt = ws.get<float>((size_t)N * Mp);  // [N, Mp]
    float* xt = ws.get<float>((size_t)K * Mp);   // [K, Mp]
    rc = rdm_transpose_ld(dy, M, N, N, dyt, Mp, stream);
    if (rc != RDM_OK) return rc;
    rc = rdm_transpose_ld(x, M, K, K, xt, Mp, stream);
    if (rc != RDM_OK) return rc;
    const size_t part = (size_t)16 * N * K * 4;
    void* p = ws.get<char>(part);
    if (w_is_nk)  // dW [N,K] = dy^T [N,M] . (x^T as [out = K, inner = M])
      rc = rdm_linear_gn(dyt, Mp, xt, Mp, 1, nullptr, dw, K, N, K, M, 0, p, part, nullptr, 0, nullptr, stream);
    else          // dW [K,N] = x^T [K,M] . (dy^T as [out = N, inner = M])
      rc = rdm_linear_gn(xt, Mp, dyt, Mp, 1, nullptr, dw, N, K, N, M, 0, p, part, nullptr, 0, nullptr, stream);
    if (rc != RDM_OK) return rc;
  }
  if (!ws.ok) {
    rdm_set_error("rdm_linear_bwd: workspace overflow");
    return RDM_ERR_WORKSPACE;
  }
  return RDM_OK;
}

extern "C" int rdm_colsum(const float* x, int rows, int cols, int ldx, float* out_zeroed_or_accum, cudaStream_t stream) {
  RDM_CHECK_ARG(rows >= 0 && cols >= 1 && ldx >= cols, "rdm_colsum: bad shape");
  if (rows == 0) return RDM_OK;
  int rpc = 256;
  while (rpc < rows && (long long)cdiv(cols, 32) * cdiv(rows, rpc) > 1184) rpc <<= 1;
  colsum_kernel<<<dim3(cdiv(cols, 32), cdiv(rows, rpc)), 256, 0, stream>>>(x, rows, cols, ldx, rpc, out_zeroed_or_accum);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

extern "C" int rdm_groupnorm_bwd(const float* x, const float* y, const float* dy, const float* gamma, int N, int C, int groups, float eps,
                                 int act, float slope, double* stats_scratch, double* dgamma_dbeta_scratch, float* dz_scratch, float* dx,
                                 float* dgamma, float* dbeta, cudaStream_t stream) {
  RDM_CHECK_ARG(N >= 0 && C >= 1 && groups >= 1 && C % groups == 0 && groups <= 1024, "rdm_groupnorm_bwd: bad shape");
  if (N == 0) return RDM_OK;
  RDM_CUDA(cudaMemsetAsync(stats_scratch, 0, sizeof(double) * 2 * groups, stream));
  RDM_CUDA(cudaMemsetAsync(dgamma_dbeta_scratch, 0, sizeof(double) * 2 * C, stream));
  int rc = rdm_groupnorm_stats(x, N, C, groups, stats_scratch, stream);
  if (rc != RDM_OK) return rc;
  double *dg = dgamma_dbeta_scratch, *db = dgamma_dbeta_scratch + C;
  int rpc = 256;
  while (rpc < N && (long long)cdiv(C, 32) * cdiv(N, rpc) > 2368) rpc <<= 1;
  gn_bwd_colsums_kernel<<<dim3(cdiv(C, 32), cdiv(N, rpc)), 256, 0, stream>>>(x, y, dy, stats_scratch, N, C, groups, eps, act, slope, rpc, dg,
                                                                             db, dz_scratch);
  RDM_LAUNCH_CHECK();
  const long long total = (long long)N * C;
  gn_bwd_dx_kernel<<<(int)min((long long)148 * 8, (total + 255) / 256), 256, 4 * groups * sizeof(float), stream>>>(
      x, dz_scratch, stats_scratch, gamma, dg, db, N, C, groups, eps, dx);
  RDM_LAUNCH_CHECK();
  // double -> float parameter gradients
  cast_d2f_kernel<<<cdiv(C, 256), 256, 0, stream>>>(dg, dgamma, C);
  RDM_LAUNCH_CHECK();
  cast_d2f_kernel<<<cdiv(C, 256), 256, 0, stream>>>(db, dbeta, C);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

extern "C" int rdm_layernorm_bwd(const float* x, const float* residual, const float* y, const float* dy, const float* gamma, int N, int C,
                                 float eps, int act, float* dx, float* dgamma_accum, float* dbeta_accum, cudaStream_t stream) {
  RDM_CHECK_ARG(N >= 0 && C >= 1, "rdm_layernorm_bwd: bad shape");
  if (N == 0) return RDM_OK;
  ln_bwd_kernel<<<cdiv(N, 8), 256, 0, stream>>>(x, residual, y, dy, gamma, N, C, eps, act, dx, dgamma_accum, dbeta_accum);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

extern "C" int rdm_maxpool_bwd(const float* feats, const void* neighbor_indices, int index_bytes, const float* d_out, int M, int N, int H,
                               int C, float* d_feats_zeroed, cudaStream_t stream) {
  RDM_CHECK_ARG(index_bytes == 4 || index_bytes == 8, "rdm_maxpool_bwd: index_bytes must be 4 or 8");
  if (M == 0) return RDM_OK;
  const long long total = (long long)M * C;
  if (index_bytes == 8)
    maxpool_bwd_kernel<int64_t><<<cdiv(total, 256), 256, 0, stream>>>(feats, (const int64_t*)neighbor_indices, d_out, M, N, H, C, d_feats_zeroed);
  else
    maxpool_bwd_kernel<int><<<cdiv(total, 256), 256, 0, stream>>>(feats, (const int*)neighbor_indices, d_out, M, N, H, C, d_feats_zeroed);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

extern "C" int rdm_upsample_concat_bwd(const float* d_out, const void* upsample_indices, int index_bytes, int index_stride, int M, int N,
                                       int C1, int C2, float* d_feats_zeroed, float* d_skip, cudaStream_t stream) {
  RDM_CHECK_ARG(index_bytes == 4 || index_bytes == 8, "rdm_upsample_concat_bwd: index_bytes must be 4 or 8");
  if (M == 0) return RDM_OK;
  const long long total = (long long)M * (C1 + C2);
  if (index_bytes == 8)
    upsample_concat_bwd_kernel<int64_t><<<cdiv(total, 256), 256, 0, stream>>>(d_out, (const int64_t*)upsample_indices, index_stride, M, N, C1,
                                                                             C2, d_feats_zeroed, d_skip);
  else
    upsample_concat_bwd_kernel<int><<<cdiv(total, 256), 256, 0, stream>>>(d_out, (const int*)upsample_indices, index_stride, M, N, C1, C2,
                                                                         d_feats_zeroed, d_skip);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

extern "C" int rdm_activation_bwd(const float* y, const float* dy, int64_t n, int act, float slope, float* dx, cudaStream_t stream) {
  RDM_CHECK_ARG(act >= 1 && act <= 3, "rdm_activation_bwd: act must be 1 (LeakyReLU), 2 (ReLU) or 3 (clamped sigmoid)");
  if (n == 0) return RDM_OK;
  activation_bwd_kernel<<<cdiv(n, 256), 256, 0, stream>>>(y, dy, (long long)n, act, slope, dx);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// out[index[i], :] += src[i, :]  (backward of index_select along dim 0, geotransformer/modules/ops/index_select.py:4-30)
namespace {
template <typename IdxT>
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const float* __restrict__ src, const IdxT* __restrict__ index, long long count,
                                                               int C, long long rows, float* __restrict__ out) {
  const long long e = blockIdx.x * 256LL + threadIdx.x;
  if (e >= count * C) return;
  const long long i = e / C;
  const long long j = (long long)index[i];
  if (j >= 0 && j < rows) atomicAdd(&out[j * C + (e - i * C)], src[e]);
}
}  // namespace

extern "C" int rdm_scatter_add_rows(const float* src, const void* index, int index_bytes, int64_t count, int row_floats, int64_t rows,
                                    float* out_accum, cudaStream_t stream) {
  RDM_CHECK_ARG(index_bytes == 4 || index_bytes == 8, "rdm_scatter_add_rows: index_bytes must be 4 or 8");
  RDM_CHECK_ARG(count >= 0 && row_floats >= 1 && rows >= 0, "rdm_scatter_add_rows: bad sizes");
  if (count == 0) return RDM_OK;
  const long long total = (long long)count * row_floats;
  if (index_bytes == 8)
    scatter_add_rows_kernel<int64_t><<<cdiv(total, 256), 256, 0, stream>>>(src, (const int64_t*)index, (long long)count, row_floats,
                                                                          (long long)rows, out_accum);
  else
    scatter_add_rows_kernel<int><<<cdiv(total, 256), 256, 0, stream>>>(src, (const int*)index, (long long)count, row_floats, (long long)rows,
                                                                      out_accum);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}
