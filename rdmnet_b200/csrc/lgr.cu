// Fine matching + pose: correspondence extraction from the Sinkhorn scores, batched weighted Procrustes (warp-level,
// in-register 3x3 SVD), local-to-global hypothesis selection and the refinement loop - all on the device, no host
// round trips (the reference does 6 GPU->CPU->GPU SVD round trips and a .tolist() sync per pair).
//
// Reference semantics:
//   LocalGlobalRegistration  geotransformer/modules/geotransformer/local_global_registration.py:49-91,138-243
//                            (configuration of experiments/config.py:152-161: k=1, mutual=False, use_dustbin=True,
//                             use_global_score=False, correspondence_limit=None)
//   weighted_procrustes      geotransformer/modules/registration/procrustes.py:6-73
#include "common.cuh"
#include "../../include/rdm_sm100.h"

// ------------------------------------------------------------------------------------------- 3x3 SVD -> rotation
// R = V diag(1,1,sign det(V U^T)) U^T for H = U S V^T (procrustes.py:53-57). One-sided Jacobi in double.
template <typename F>
__device__ void rotation_from_H_t(const F Hin[9], float R[9]) {
  const F tiny = sizeof(F) == 8 ? (F)1e-300 : (F)1e-30, conv = sizeof(F) == 8 ? (F)1e-15 : (F)2e-7,
          rtol = sizeof(F) == 8 ? (F)1e-12 : (F)1e-6;
  // A = H (columns rotated until orthogonal): A = U S, accumulated right rotations = V
  F A[9], V[9] = {(F)1, (F)0, (F)0, (F)0, (F)1, (F)0, (F)0, (F)0, (F)1};
  for (int i = 0; i < 9; i++) A[i] = Hin[i];
  for (int sweep = 0; sweep < 30; sweep++) {
    F off = (F)0;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        F alpha = (F)0, beta = (F)0, gamma = (F)0;
        for (int r = 0; r < 3; r++) {
          alpha += A[3 * r + p] * A[3 * r + p];
          beta += A[3 * r + q] * A[3 * r + q];
          gamma += A[3 * r + p] * A[3 * r + q];
        }
        off = fmax((F)off, fabs((F)gamma) / (sqrt((F)alpha * beta) + tiny));
        if (fabs((F)gamma) < tiny) continue;
        F zeta = (beta - alpha) / ((F)2 * gamma);
        F t = (zeta >= 0 ? (F)1 : (F)-1) / (fabs((F)zeta) + sqrt((F)(F)1 + zeta * zeta));
        F c = (F)1 / sqrt((F)1 + t * t), s = c * t;
        for (int r = 0; r < 3; r++) {
          F ap = A[3 * r + p], aq = A[3 * r + q];
          A[3 * r + p] = c * ap - s * aq;
          A[3 * r + q] = s * ap + c * aq;
          F vp = V[3 * r + p], vq = V[3 * r + q];
          V[3 * r + p] = c * vp - s * vq;
          V[3 * r + q] = s * vp + c * vq;
        }
      }
    if (off < conv) break;
  }
  // singular values = column norms; sort descending (LAPACK order) so that the reflection fix hits the smallest
  F sv[3];
  int ord[3] = {0, 1, 2};
  for (int j = 0; j < 3; j++) sv[j] = sqrt((F)A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
  for (int a = 0; a < 2; a++)
    for (int b = a + 1; b < 3; b++)
      if (sv[ord[b]] > sv[ord[a]]) {
        int t = ord[a];
        ord[a] = ord[b];
        ord[b] = t;
      }
  F U[9], Vs[9];
  for (int j = 0; j < 3; j++) {
    int c = ord[j];
    for (int r = 0; r < 3; r++) Vs[3 * r + j] = V[3 * r + c];
  }
  // U columns: normalised A columns; complete degenerate ones by cross products
  F tol = rtol * fmax((F)sv[ord[0]], tiny);
  int rank = 0;
  for (int j = 0; j < 3; j++) {
    int c = ord[j];
    if (sv[c] > tol) {
      for (int r = 0; r < 3; r++) U[3 * r + j] = A[3 * r + c] / sv[c];
      rank = j + 1;
    }
  }
  if (rank == 0) {
    for (int i = 0; i < 9; i++) U[i] = (i % 4 == 0) ? (F)1 : (F)0;
  } else if (rank == 1) {
    F x = U[0], y = U[3], z = U[6];
    F ax = fabs((F)x) < (F)0.9 ? (F)1 : (F)0, ay = fabs((F)x) < (F)0.9 ? (F)0 : (F)1, az = (F)0;
    F bx = y * az - z * ay, by = z * ax - x * az, bz = x * ay - y * ax;
    F nb = sqrt((F)bx * bx + by * by + bz * bz);
    U[1] = bx / nb; U[4] = by / nb; U[7] = bz / nb;
    rank = 2;
  }
  if (rank == 2) {
    U[2] = U[3] * U[7] - U[6] * U[4];
    U[5] = U[6] * U[1] - U[0] * U[7];
    U[8] = U[0] * U[4] - U[3] * U[1];
  }
  // d = sign(det(V U^T)) = sign(det V * det U)
  auto det3 = [](const F* M) -> F {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
  };
  F dd = det3(Vs) * det3(U);
  F d = dd > (F)0 ? (F)1 : (dd < (F)0 ? (F)-1 : (F)0);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      R[3 * i + j] = (float)(Vs[3 * i + 0] * U[3 * j + 0] + Vs[3 * i + 1] * U[3 * j + 1] + d * Vs[3 * i + 2] * U[3 * j + 2]);
}

__device__ void rotation_from_H(const double Hin[9], float R[9]) { rotation_from_H_t<double>(Hin, R); }

// Weighted Procrustes over n correspondences by one warp. T (row-major 4x4) written by lane 0.
__device__ void warp_procrustes(const float* __restrict__ src, const float* __restrict__ ref,
                                const float* __restrict__ w, int n, float eps, float* __restrict__ T, int lane) {
  float sw = 0.f;
  for (int i = lane; i < n; i += 32) sw += fmaxf(w[i], 0.f) * (w[i] >= 0.f);  // weights < 0 -> 0 (procrustes.py:42)
  sw = warp_sum(sw);
  float inv = 1.f / (sw + eps);
  float sc[3] = {0, 0, 0}, rc[3] = {0, 0, 0};
  for (int i = lane; i < n; i += 32) {
    float wi = (w[i] >= 0.f ? w[i] : 0.f) * inv;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      sc[d] = fmaf(src[3 * i + d], wi, sc[d]);
      rc[d] = fmaf(ref[3 * i + d], wi, rc[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
    sc[d] = warp_sum(sc[d]);
    rc[d] = warp_sum(rc[d]);
  }
  float H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = lane; i < n; i += 32) {
    float wi = (w[i] >= 0.f ? w[i] : 0.f) * inv;
    float s[3] = {src[3 * i] - sc[0], src[3 * i + 1] - sc[1], src[3 * i + 2] - sc[2]};
    float r[3] = {wi * (ref[3 * i] - rc[0]), wi * (ref[3 * i + 1] - rc[1]), wi * (ref[3 * i + 2] - rc[2])};
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++) H[3 * a + b] = fmaf(s[a], r[b], H[3 * a + b]);
  }
#pragma unroll
  for (int k = 0; k < 9; k++) H[k] = warp_sum(H[k]);
  if (lane == 0) {
    double Hd[9];
    for (int k = 0; k < 9; k++) Hd[k] = (double)H[k];
    float R[9];
    rotation_from_H(Hd, R);
    for (int a = 0; a < 3; a++) {
      float t = rc[a] - (R[3 * a] * sc[0] + R[3 * a + 1] * sc[1] + R[3 * a + 2] * sc[2]);
      T[4 * a + 0] = R[3 * a];
      T[4 * a + 1] = R[3 * a + 1];
      T[4 * a + 2] = R[3 * a + 2];
      T[4 * a + 3] = t;
    }
    T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
  }
}

extern "C" __global__ void procrustes_batch_kernel(const float* __restrict__ src, const float* __restrict__ ref,
                                                   const float* __restrict__ w, int B, int n, float eps,
                                                   float* __restrict__ T) {
  int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  warp_procrustes(src + (size_t)b * n * 3, ref + (size_t)b * n * 3, w + (size_t)b * n, n, eps, T + (size_t)b * 16, lane);
}

extern "C" int rdm_weighted_procrustes(const float* src_points, const float* ref_points, const float* weights, int batch,
                                       int n, float eps, float* out_transforms, cudaStream_t stream) {
  RDM_CHECK_ARG(batch >= 0 && n >= 0, "rdm_weighted_procrustes: bad shape");
  if (batch == 0) return RDM_OK;
  procrustes_batch_kernel<<<cdiv(batch, 4), 128, 0, stream>>>(src_points, ref_points, weights, batch, n, eps, out_transforms);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// ------------------------------------------------------------------------------------------- correspondence extraction
// One CTA per patch. s = exp(score); ref side: row argmax must beat the dustbin column; src side: column argmax must
// beat the dustbin row; OR of the two; dustbin row/col dropped; AND with the knn masks (:49-91).
// Entries are emitted in row-major (i, j) order into a per-patch staging list (<= 2K entries).
#define LGR_MAXK 128
__global__ void __launch_bounds__(256) lgr_extract_kernel(const float* __restrict__ scores, int K,
                                                          const unsigned char* __restrict__ rmask_nodes,
                                                          const unsigned char* __restrict__ cmask_nodes,
                                                          const int64_t* __restrict__ ridx, const int64_t* __restrict__ sidx,
                                                          int* __restrict__ patch_cnt, int* __restrict__ st_ij,
                                                          float* __restrict__ st_score) {
  __shared__ int s_rarg[LGR_MAXK + 1], s_carg[LGR_MAXK + 1];
  __shared__ float s_rmax[LGR_MAXK + 1], s_cmax[LGR_MAXK + 1];
  __shared__ int s_scan[33];
  const int p = blockIdx.x, tid = threadIdx.x, K1 = K + 1;
  const float* sp = scores + (size_t)p * K1 * K1;
  const unsigned char* rm = rmask_nodes + (size_t)ridx[p] * K;
  const unsigned char* cm = cmask_nodes + (size_t)sidx[p] * K;
  // row argmax over all K1 columns (first max), for the K1 rows; column argmax likewise. exp is monotone, so the
  // argmax is taken on the log scores and exp applied only for the comparisons / outputs.
  for (int i = tid; i < K1; i += 256) {
    float best = -3.4e38f;
    int bj = 0;
    for (int j = 0; j < K1; j++) {
      float v = expf(sp[(size_t)i * K1 + j]);
      if (v > best) {
        best = v;
        bj = j;
      }
    }
    s_rarg[i] = bj;
    s_rmax[i] = best;
  }
  for (int j = tid; j < K1; j += 256) {
    float best = -3.4e38f;
    int bi = 0;
    for (int i = 0; i < K1; i++) {
      float v = expf(sp[(size_t)i * K1 + j]);
      if (v > best) {
        best = v;
        bi = i;
      }
    }
    s_carg[j] = bi;
    s_cmax[j] = best;
  }
  __syncthreads();
  auto is_corr = [&](int i, int j) -> bool {
    if (!rm[i] || !cm[j]) return false;
    float dr = expf(sp[(size_t)i * K1 + K]), dc = expf(sp[(size_t)K * K1 + j]);
    bool a = (s_rarg[i] == j) && (s_rmax[i] > dr);
    bool b = (s_carg[j] == i) && (s_cmax[j] > dc);
    return a || b;
  };
  int myc = 0;
  if (tid < K) {
    for (int j = 0; j < K; j++) myc += is_corr(tid, j);
  }
  int total;
  int pre = block_exclusive_scan(tid < K ? myc : 0, s_scan, &total);
  if (tid < K && myc > 0) {
    int o = pre;
    for (int j = 0; j < K; j++)
      if (is_corr(tid, j)) {
        st_ij[(size_t)p * 2 * K + o] = tid * K + j;
        st_score[(size_t)p * 2 * K + o] = expf(sp[(size_t)tid * K1 + j]);
        o++;
      }
  }
  if (tid == 0) patch_cnt[p] = total;
}

// offsets over patches + chunk table of the patches with >= threshold correspondences (single CTA)
__global__ void __launch_bounds__(1024) lgr_offsets_kernel(const int* __restrict__ patch_cnt, int P, int threshold,
                                                           int* __restrict__ patch_off, int* __restrict__ chunk_patch,
                                                           int* __restrict__ meta /* [0]=C, [1]=num chunks, [2]=max chunk */) {
  __shared__ int s_scan[33];
  __shared__ int s_max;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) s_max = 0;
  int chunk = (P + nt - 1) / nt;
  int beg = min(P, tid * chunk), end = min(P, beg + chunk);
  int s = 0, nc = 0, mx = 0;
  for (int i = beg; i < end; i++) {
    s += patch_cnt[i];
    if (patch_cnt[i] >= threshold) {
      nc++;
      mx = max(mx, patch_cnt[i]);
    }
  }
  int total, ctotal;
  int pre = block_exclusive_scan(s, s_scan, &total);
  int cpre = block_exclusive_scan(nc, s_scan, &ctotal);
  atomicMax(&s_max, mx);
  for (int i = beg; i < end; i++) {
    patch_off[i] = pre;
    pre += patch_cnt[i];
    if (patch_cnt[i] >= threshold) chunk_patch[cpre++] = i;
  }
  __syncthreads();
  if (tid == 0) {
    patch_off[P] = total;
    meta[0] = total;
    meta[1] = ctotal;
    meta[2] = s_max;
  }
}

// scatter the staged lists into the stacked outputs (row-major order = torch.nonzero order, :147-150)
__global__ void lgr_gather_kernel(const int* __restrict__ patch_cnt, const int* __restrict__ patch_off,
                                  const int* __restrict__ st_ij, const float* __restrict__ st_score, int K,
                                  const float* __restrict__ ref_pts, const float* __restrict__ src_pts,
                                  const int64_t* __restrict__ rknn, const int64_t* __restrict__ sknn,
                                  const int64_t* __restrict__ ridx, const int64_t* __restrict__ sidx,
                                  float* __restrict__ out_ref, float* __restrict__ out_src, float* __restrict__ out_score,
                                  int* __restrict__ out_bij) {
  const int p = blockIdx.x;
  const int n = patch_cnt[p], off = patch_off[p];
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    int ij = st_ij[(size_t)p * 2 * K + e];
    int i = ij / K, j = ij - i * K;
    long long ri = rknn[(size_t)ridx[p] * K + i], si = sknn[(size_t)sidx[p] * K + j];
    int o = off + e;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      out_ref[3 * o + d] = ref_pts[3 * ri + d];
      out_src[3 * o + d] = src_pts[3 * si + d];
    }
    out_score[o] = st_score[(size_t)p * 2 * K + e];
    out_bij[3 * o] = p;
    out_bij[3 * o + 1] = i;
    out_bij[3 * o + 2] = j;
  }
}

// local hypotheses: one warp per chunk (patch with >= threshold correspondences), then its inlier count over ALL
// correspondences (:173-185)
__global__ void __launch_bounds__(128) lgr_hypothesis_kernel(const int* __restrict__ meta,
                                                             const int* __restrict__ chunk_patch,
                                                             const int* __restrict__ patch_cnt,
                                                             const int* __restrict__ patch_off,
                                                             const float* __restrict__ ref, const float* __restrict__ src,
                                                             const float* __restrict__ score, float radius, float eps,
                                                             float* __restrict__ hyp_T, int* __restrict__ hyp_inl) {
  const int c = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= meta[1]) return;
  const int p = chunk_patch[c], off = patch_off[p], n = patch_cnt[p], C = meta[0];
  float* T = hyp_T + (size_t)c * 16;
  warp_procrustes(src + 3 * (size_t)off, ref + 3 * (size_t)off, score + off, n, eps, T, lane);
  __syncwarp();
  float t[12];
#pragma unroll
  for (int k = 0; k < 12; k++) t[k] = __shfl_sync(FULL_MASK, lane == 0 ? T[k] : 0.f, 0);
  int inl = 0;
  for (int i = lane; i < C; i += 32) {
    float x = src[3 * i], y = src[3 * i + 1], z = src[3 * i + 2];
    float dx = ref[3 * i] - (t[0] * x + t[1] * y + t[2] * z + t[3]);
    float dy = ref[3 * i + 1] - (t[4] * x + t[5] * y + t[6] * z + t[7]);
    float dz = ref[3 * i + 2] - (t[8] * x + t[9] * y + t[10] * z + t[11]);
    inl += sqrtf(dx * dx + dy * dy + dz * dz) < radius;
  }
  inl = warp_sum_i(inl);
  if (lane == 0) hyp_inl[c] = inl;
}

// best hypothesis -> inlier weights -> `steps` x { weighted Procrustes over all correspondences; re-weight } (:186-202)
__global__ void __launch_bounds__(1024) lgr_refine_kernel(const int* __restrict__ meta, const float* __restrict__ hyp_T,
                                                          const int* __restrict__ hyp_inl, const float* __restrict__ ref,
                                                          const float* __restrict__ src, const float* __restrict__ score,
                                                          float radius, float eps, int steps, float* __restrict__ wbuf,
                                                          float* __restrict__ out_T) {
  __shared__ float s_T[16];
  __shared__ float s_red[32][16];
  __shared__ int s_best;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int C = meta[0], NH = meta[1];
  if (C == 0) {
    if (tid < 16) out_T[tid] = (tid % 5 == 0) ? 1.f : 0.f;
    return;
  }
  auto reweight = [&]() {  // wbuf = score * [ |ref - T src| < radius ]
    for (int i = tid; i < C; i += nt) {
      float x = src[3 * i], y = src[3 * i + 1], z = src[3 * i + 2];
      float dx = ref[3 * i] - (s_T[0] * x + s_T[1] * y + s_T[2] * z + s_T[3]);
      float dy = ref[3 * i + 1] - (s_T[4] * x + s_T[5] * y + s_T[6] * z + s_T[7]);
      float dz = ref[3 * i + 2] - (s_T[8] * x + s_T[9] * y + s_T[10] * z + s_T[11]);
      wbuf[i] = sqrtf(dx * dx + dy * dy + dz * dz) < radius ? score[i] : 0.f;
    }
  };
  // block-wide weighted Procrustes with weights wbuf -> s_T
  auto block_sum = [&](float* vals, int cnt) {  // reduce `cnt` (<=16) per-thread partials; result in s_red[0][*]
    for (int k = 0; k < cnt; k++) vals[k] = warp_sum(vals[k]);
    __syncthreads();
    if (lane == 0)
      for (int k = 0; k < cnt; k++) s_red[warp][k] = vals[k];
    __syncthreads();
    if (warp == 0) {
      for (int k = 0; k < cnt; k++) {
        float v = lane < nw ? s_red[lane][k] : 0.f;
        v = warp_sum(v);
        if (lane == 0) s_red[0][k] = v;
      }
    }
    __syncthreads();
  };
  auto solve = [&]() {
    float v[16];
    v[0] = 0.f;
    for (int i = tid; i < C; i += nt) v[0] += wbuf[i] >= 0.f ? wbuf[i] : 0.f;
    block_sum(v, 1);
    float inv = 1.f / (s_red[0][0] + eps);
    __syncthreads();
    for (int k = 0; k < 6; k++) v[k] = 0.f;
    for (int i = tid; i < C; i += nt) {
      float wi = (wbuf[i] >= 0.f ? wbuf[i] : 0.f) * inv;
      for (int d = 0; d < 3; d++) {
        v[d] = fmaf(src[3 * i + d], wi, v[d]);
        v[3 + d] = fmaf(ref[3 * i + d], wi, v[3 + d]);
      }
    }
    block_sum(v, 6);
    float sc[3] = {s_red[0][0], s_red[0][1], s_red[0][2]}, rc[3] = {s_red[0][3], s_red[0][4], s_red[0][5]};
    __syncthreads();
    for (int k = 0; k < 9; k++) v[k] = 0.f;
    for (int i = tid; i < C; i += nt) {
      float wi = (wbuf[i] >= 0.f ? wbuf[i] : 0.f) * inv;
      float s[3] = {src[3 * i] - sc[0], src[3 * i + 1] - sc[1], src[3 * i + 2] - sc[2]};
      float r[3] = {wi * (ref[3 * i] - rc[0]), wi * (ref[3 * i + 1] - rc[1]), wi * (ref[3 * i + 2] - rc[2])};
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) v[3 * a + b] = fmaf(s[a], r[b], v[3 * a + b]);
    }
    block_sum(v, 9);
    if (tid == 0) {
      // single-precision Jacobi here (the reference's SVD is LAPACK sgesdd, procrustes.py:53): this thread is the serial
      // section of the refinement loop, and B200's FP64 sqrt / divide made the double version 84 us for the 6 solves
      float Hf[9];
      for (int k = 0; k < 9; k++) Hf[k] = s_red[0][k];
      float R[9];
      rotation_from_H_t<float>(Hf, R);
      for (int a = 0; a < 3; a++) {
        s_T[4 * a] = R[3 * a];
        s_T[4 * a + 1] = R[3 * a + 1];
        s_T[4 * a + 2] = R[3 * a + 2];
        s_T[4 * a + 3] = rc[a] - (R[3 * a] * sc[0] + R[3 * a + 1] * sc[1] + R[3 * a + 2] * sc[2]);
      }
      s_T[12] = s_T[13] = s_T[14] = 0.f;
      s_T[15] = 1.f;
    }
    __syncthreads();
  };
  if (NH > 0) {
    if (tid == 0) {  // argmax (first maximum) of the inlier counts
      int best = 0, bv = hyp_inl[0];
      for (int c = 1; c < NH; c++)
        if (hyp_inl[c] > bv) {
          bv = hyp_inl[c];
          best = c;
        }
      s_best = best;
    }
    __syncthreads();
    if (tid < 16) s_T[tid] = hyp_T[(size_t)s_best * 16 + tid];
    __syncthreads();
    reweight();
  } else {  // degenerate: initialise from all correspondences (:188-193)
    for (int i = tid; i < C; i += nt) wbuf[i] = score[i];
    __syncthreads();
    solve();
    reweight();
  }
  __syncthreads();
  solve();
  for (int it = 0; it < steps - 1; it++) {
    reweight();
    __syncthreads();
    solve();
  }
  if (tid < 16) out_T[tid] = s_T[tid];
}

extern "C" size_t rdm_lgr_workspace(int num_patches, int K) {
  size_t P = (size_t)num_patches, bytes = 0;
  bytes += 4 * align_up((P + 1) * 4, 256);        // patch_cnt, patch_off, chunk_patch, hyp_inl
  bytes += 2 * align_up(P * 2 * K * 4, 256);      // staging ij + score
  bytes += align_up(P * 16 * 4, 256);             // hyp_T
  bytes += align_up(P * 2 * K * 4, 256);          // weights
  return bytes + 1024;
}

extern "C" int rdm_lgr(const float* matching_scores, int num_patches, int K, const float* ref_points_f,
                       const float* src_points_f, const int64_t* ref_knn_indices, const int64_t* src_knn_indices,
                       const unsigned char* ref_knn_masks, const unsigned char* src_knn_masks,
                       const int64_t* ref_corr_indices, const int64_t* src_corr_indices, float acceptance_radius,
                       int correspondence_threshold, int num_refinement_steps, float* out_ref_corr_points,
                       float* out_src_corr_points, float* out_corr_scores, int* out_corr_bij, float* out_transform,
                       int* out_meta, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  RDM_CHECK_ARG(K >= 1 && K <= LGR_MAXK && num_patches >= 0, "rdm_lgr: patch size must be <= 128");
  const int P = num_patches;
  Workspace ws(workspace, workspace_bytes);
  int* patch_cnt = ws.get<int>(P + 1);
  int* patch_off = ws.get<int>(P + 1);
  int* chunk_patch = ws.get<int>(P + 1);
  int* hyp_inl = ws.get<int>(P + 1);
  int* st_ij = ws.get<int>((size_t)P * 2 * K);
  float* st_score = ws.get<float>((size_t)P * 2 * K);
  float* hyp_T = ws.get<float>((size_t)P * 16);
  float* wbuf = ws.get<float>((size_t)P * 2 * K);
  if (!ws.ok) {
    rdm_set_error("rdm_lgr: workspace too small");
    return RDM_ERR_WORKSPACE;
  }
  if (P > 0) {
    lgr_extract_kernel<<<P, 256, 0, stream>>>(matching_scores, K, ref_knn_masks, src_knn_masks, ref_corr_indices,
                                             src_corr_indices, patch_cnt, st_ij, st_score);
    RDM_LAUNCH_CHECK();
  }
  lgr_offsets_kernel<<<1, 1024, 0, stream>>>(patch_cnt, P, correspondence_threshold, patch_off, chunk_patch, out_meta);
  RDM_LAUNCH_CHECK();
  if (P > 0) {
    lgr_gather_kernel<<<P, 128, 0, stream>>>(patch_cnt, patch_off, st_ij, st_score, K, ref_points_f, src_points_f,
                                            ref_knn_indices, src_knn_indices, ref_corr_indices, src_corr_indices,
                                            out_ref_corr_points, out_src_corr_points, out_corr_scores, out_corr_bij);
    RDM_LAUNCH_CHECK();
    lgr_hypothesis_kernel<<<cdiv(P, 4), 128, 0, stream>>>(out_meta, chunk_patch, patch_cnt, patch_off, out_ref_corr_points,
                                                         out_src_corr_points, out_corr_scores, acceptance_radius, 1e-5f,
                                                         hyp_T, hyp_inl);
    RDM_LAUNCH_CHECK();
  }
  lgr_refine_kernel<<<1, 1024, 0, stream>>>(out_meta, hyp_T, hyp_inl, out_ref_corr_points, out_src_corr_points,
                                           out_corr_scores, acceptance_radius, 1e-5f, num_refinement_steps, wbuf,
                                           out_transform);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// ------------------------------------------------------------------------------------------- RANSAC on correspondences
// GPU counterpart of registration_with_ransac_from_correspondences (geotransformer/utils/open3d.py:173-203, the open3d
// call experiments/infer.py:76-82 makes per pair with 50 000 iterations on the CPU). One warp per iteration: lane 0 draws
// ransac_n distinct correspondences from a counter-based generator (seed, iteration, draw) and solves the unweighted
// Kabsch problem on them (single-precision Jacobi SVD); the 32 lanes count the correspondences within
// distance_threshold; the best (inliers, lowest iteration) hypothesis wins through one 64-bit atomicMax. A second
// launch re-derives the winner from its iteration number and refits on all of its inliers.
namespace {
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// lane 0 only: hypothesis of iteration `it` -> T[12] (rows of [R|t]); false if the sample is degenerate
__device__ bool ransac_sample_fit(const float* __restrict__ src, const float* __restrict__ ref, int C, int ransac_n,
                                  unsigned long long seed, int it, float* T12) {
  int pick[8];
  unsigned long long ctr = 0;
  for (int a = 0; a < ransac_n; a++) {
    for (;;) {
      const int c = (int)(mix64(seed ^ (((unsigned long long)it << 20) + ctr++)) % (unsigned long long)C);
      bool dup = false;
      for (int b = 0; b < a; b++) dup |= pick[b] == c;
      if (!dup) {
        pick[a] = c;
        break;
      }
    }
  }
  float sc[3] = {0, 0, 0}, rc[3] = {0, 0, 0};
  for (int a = 0; a < ransac_n; a++)
    for (int d = 0; d < 3; d++) {
      sc[d] += src[3 * pick[a] + d];
      rc[d] += ref[3 * pick[a] + d];
    }
  const float inv = 1.f / (float)ransac_n;
  for (int d = 0; d < 3; d++) {
    sc[d] *= inv;
    rc[d] *= inv;
  }
  float H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int a = 0; a < ransac_n; a++) {
    const float s[3] = {src[3 * pick[a]] - sc[0], src[3 * pick[a] + 1] - sc[1], src[3 * pick[a] + 2] - sc[2]};
    const float r[3] = {ref[3 * pick[a]] - rc[0], ref[3 * pick[a] + 1] - rc[1], ref[3 * pick[a] + 2] - rc[2]};
    for (int x = 0; x < 3; x++)
      for (int y = 0; y < 3; y++) H[3 * x + y] = fmaf(s[x], r[y], H[3 * x + y]);
  }
  float R[9];
  rotation_from_H_t<float>(H, R);
  for (int a = 0; a < 3; a++) {
    T12[4 * a] = R[3 * a];
    T12[4 * a + 1] = R[3 * a + 1];
    T12[4 * a + 2] = R[3 * a + 2];
    T12[4 * a + 3] = rc[a] - (R[3 * a] * sc[0] + R[3 * a + 1] * sc[1] + R[3 * a + 2] * sc[2]);
  }
  return true;
}
__device__ __forceinline__ bool ransac_inlier(const float* T, const float* __restrict__ src, const float* __restrict__ ref, int i,
                                              float thr2) {
  const float x = src[3 * i], y = src[3 * i + 1], z = src[3 * i + 2];
  const float dx = fmaf(T[2], z, fmaf(T[1], y, T[0] * x)) + T[3] - ref[3 * i];
  const float dy = fmaf(T[6], z, fmaf(T[5], y, T[4] * x)) + T[7] - ref[3 * i + 1];
  const float dz = fmaf(T[10], z, fmaf(T[9], y, T[8] * x)) + T[11] - ref[3 * i + 2];
  return fmaf(dz, dz, fmaf(dy, dy, dx * dx)) < thr2;
}

__global__ void __launch_bounds__(128) ransac_hypothesis_kernel(const float* __restrict__ src, const float* __restrict__ ref, int C,
                                                                int ransac_n, int num_iterations, unsigned long long seed,
                                                                float thr2, unsigned long long* __restrict__ best) {
  const int it = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (it >= num_iterations) return;
  float T[12];
  if (lane == 0) ransac_sample_fit(src, ref, C, ransac_n, seed, it, T);
#pragma unroll
  for (int k = 0; k < 12; k++) T[k] = __shfl_sync(FULL_MASK, T[k], 0);
  int inl = 0;
  for (int i = lane; i < C; i += 32) inl += ransac_inlier(T, src, ref, i, thr2) ? 1 : 0;
  inl = warp_sum_i(inl);
  if (lane == 0) atomicMax(best, ((unsigned long long)inl << 32) | (unsigned long long)(0x7fffffff - it));
}

__global__ void __launch_bounds__(32) ransac_refit_kernel(const float* __restrict__ src, const float* __restrict__ ref, int C,
                                                          int ransac_n, unsigned long long seed, float thr2,
                                                          const unsigned long long* __restrict__ best, float* __restrict__ weights,
                                                          float* __restrict__ out_T, int* __restrict__ out_meta) {
  const int lane = threadIdx.x;
  const unsigned long long b = *best;
  const int it = 0x7fffffff - (int)(b & 0xffffffffull);
  float T[12];
  if (lane == 0) ransac_sample_fit(src, ref, C, ransac_n, seed, it, T);
#pragma unroll
  for (int k = 0; k < 12; k++) T[k] = __shfl_sync(FULL_MASK, T[k], 0);
  int inl = 0;
  for (int i = lane; i < C; i += 32) {
    const bool in = ransac_inlier(T, src, ref, i, thr2);
    weights[i] = in ? 1.f : 0.f;
    inl += in ? 1 : 0;
  }
  inl = warp_sum_i(inl);
  __syncwarp();
  if (inl >= 3) {
    warp_procrustes(src, ref, weights, C, 0.f, out_T, lane);
  } else if (lane == 0) {  // nothing to refit on: the sampled model itself
    for (int k = 0; k < 12; k++) out_T[k] = T[k];
    out_T[12] = out_T[13] = out_T[14] = 0.f;
    out_T[15] = 1.f;
  }
  if (lane == 0 && out_meta != nullptr) {
    out_meta[0] = inl;
    out_meta[1] = it;
  }
}
}  // namespace

extern "C" size_t rdm_ransac_workspace(int num_correspondences) { return align_up((size_t)num_correspondences * 4, 256) + 512; }

extern "C" int rdm_ransac_correspondences(const float* src_points, const float* ref_points, int num_correspondences,
                                          float distance_threshold, int ransac_n, int num_iterations, unsigned long long seed,
                                          float* out_transform, int* out_meta, void* workspace, size_t workspace_bytes,
                                          cudaStream_t stream) {
  RDM_CHECK_ARG(ransac_n >= 3 && ransac_n <= 8, "rdm_ransac_correspondences: ransac_n must be in [3, 8]");
  RDM_CHECK_ARG(num_correspondences >= ransac_n, "rdm_ransac_correspondences: fewer correspondences (%d) than ransac_n (%d)",
                num_correspondences, ransac_n);
  RDM_CHECK_ARG(num_iterations >= 1 && distance_threshold > 0.f, "rdm_ransac_correspondences: bad arguments");
  Workspace ws(workspace, workspace_bytes);
  unsigned long long* best = ws.get<unsigned long long>(1);
  float* weights = ws.get<float>(num_correspondences);
  if (!ws.ok) {
    rdm_set_error("rdm_ransac_correspondences: workspace too small");
    return RDM_ERR_WORKSPACE;
  }
  RDM_CUDA(cudaMemsetAsync(best, 0, sizeof(unsigned long long), stream));
  const float thr2 = distance_threshold * distance_threshold;
  ransac_hypothesis_kernel<<<cdiv(num_iterations, 4), 128, 0, stream>>>(src_points, ref_points, num_correspondences, ransac_n,
                                                                        num_iterations, seed, thr2, best);
  RDM_LAUNCH_CHECK();
  ransac_refit_kernel<<<1, 32, 0, stream>>>(src_points, ref_points, num_correspondences, ransac_n, seed, thr2, best, weights,
                                            out_transform, out_meta);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}
