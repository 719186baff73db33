// Training / validation side of the path: ground-truth superpoint (patch) correspondences on the GPU.
//   rdm_node_correspondences   get_node_correspondences (geotransformer/modules/registration/matching.py:252-366)
//                              and get_node_overlap (:368-436, the same overlap ratio without the sphere pre-filter)
//   rdm_node_distance_mask     get_node_correspondences_disance (:441-503)
//   rdm_compact_nonzero        torch.nonzero(mat > 0) in row-major order (+ the values), one CTA, device-side count
// Distances use the expression order of pairwise_distance (modules/ops/pairwise_distance.py:24-30: |x|^2 - 2 x.y + |y|^2,
// clamp 1e-12), like rdm_point_to_node. The reference runs these as ~25 ATen launches over dense (B,K,K) temporaries.
#include "common.cuh"
#include "../../include/rdm_sm100.h"

namespace {
__device__ __forceinline__ float sq3(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
__device__ __forceinline__ float pdist(float ax, float ay, float az, float a2, float bx, float by, float bz, float b2) {
  const float xy = fmaf(az, bz, fmaf(ay, by, __fmul_rn(ax, bx)));
  return fmaxf(__fadd_rn(__fsub_rn(a2, __fmul_rn(2.0f, xy)), b2), 1e-12f);
}
__device__ __forceinline__ float3 xform(const float* __restrict__ T, float x, float y, float z) {
  if (T == nullptr) return make_float3(x, y, z);
  return make_float3(fmaf(z, T[2], fmaf(y, T[1], x * T[0])) + T[3], fmaf(z, T[6], fmaf(y, T[5], x * T[4])) + T[7],
                     fmaf(z, T[10], fmaf(y, T[9], x * T[8])) + T[11]);
}

// one warp per node: radius of the enclosing sphere of its patch (matching.py:311-316); nodes / patch points of the src
// side are transformed on the fly (T != NULL), the transformed nodes are also written out for the pair kernel
__global__ void __launch_bounds__(256) patch_radius_kernel(const float* __restrict__ nodes, const float* __restrict__ knn_pts,
                                                           const unsigned char* __restrict__ knn_masks, const float* __restrict__ T,
                                                           int n_nodes, int K, float* __restrict__ nodes_out, float* __restrict__ radius) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= n_nodes) return;
  const float3 c = xform(T, nodes[3 * n], nodes[3 * n + 1], nodes[3 * n + 2]);
  float best = 0.f;
  for (int k = lane; k < K; k += 32) {
    if (knn_masks != nullptr && !knn_masks[(size_t)n * K + k]) continue;
    const float* p = knn_pts + ((size_t)n * K + k) * 3;
    const float3 q = xform(T, p[0], p[1], p[2]);
    const float dx = q.x - c.x, dy = q.y - c.y, dz = q.z - c.z;
    best = fmaxf(best, sqrtf(dx * dx + dy * dy + dz * dz));
  }
  best = warp_max(best);
  if (lane == 0) {
    radius[n] = best;
    nodes_out[3 * n] = c.x;
    nodes_out[3 * n + 1] = c.y;
    nodes_out[3 * n + 2] = c.z;
  }
}

// one CTA (128 threads) per (ref node m, src node n): sphere test, then the patch overlap ratio
// overlap = (|{i : exists j, d(i,j) < r^2}| / #valid_i + |{j : exists i ...}| / #valid_j) / 2      (matching.py:334-343)
template <int KMAX>
__global__ void __launch_bounds__(128) patch_overlap_kernel(const float* __restrict__ ref_nodes, const float* __restrict__ src_nodes_t,
                                                            const float* __restrict__ ref_radius, const float* __restrict__ src_radius,
                                                            const float* __restrict__ ref_knn, const float* __restrict__ src_knn,
                                                            const unsigned char* __restrict__ ref_masks, const unsigned char* __restrict__ src_masks,
                                                            const unsigned char* __restrict__ ref_knn_masks,
                                                            const unsigned char* __restrict__ src_knn_masks, const float* __restrict__ T,
                                                            int M, int N, int K, float pos_radius, int sphere_filter,
                                                            float* __restrict__ overlaps, unsigned char* __restrict__ intersect) {
  const int m = blockIdx.x / N, n = blockIdx.x - m * N, tid = threadIdx.x;
  bool cand = (ref_masks == nullptr || ref_masks[m]) && (src_masks == nullptr || src_masks[n]);
  if (cand && sphere_filter) {
    const float ax = ref_nodes[3 * m], ay = ref_nodes[3 * m + 1], az = ref_nodes[3 * m + 2];
    const float bx = src_nodes_t[3 * n], by = src_nodes_t[3 * n + 1], bz = src_nodes_t[3 * n + 2];
    const float d = sqrtf(pdist(ax, ay, az, sq3(ax, ay, az), bx, by, bz, sq3(bx, by, bz)));
    cand = (ref_radius[m] + src_radius[n] + pos_radius - d) > 0.f;
  }
  if (intersect != nullptr && tid == 0) intersect[blockIdx.x] = cand ? 1 : 0;
  if (!cand) {
    if (tid == 0) overlaps[blockIdx.x] = 0.f;
    return;
  }
  __shared__ float4 s_src[KMAX];
  __shared__ unsigned s_colhit[KMAX];
  __shared__ int s_cnt[4];
  for (int j = tid; j < K; j += 128) {
    const float* p = src_knn + ((size_t)n * K + j) * 3;
    const float3 q = xform(T, p[0], p[1], p[2]);
    const bool ok = src_knn_masks == nullptr || src_knn_masks[(size_t)n * K + j];
    s_src[j] = make_float4(q.x, q.y, q.z, ok ? sq3(q.x, q.y, q.z) : -1.f);  // w < 0 marks a masked slot
    s_colhit[j] = 0u;
  }
  if (tid < 4) s_cnt[tid] = 0;
  __syncthreads();
  const float r2 = pos_radius * pos_radius;
  int row_hits = 0, rows_valid = 0;
  for (int i = tid; i < K; i += 128) {
    if (ref_knn_masks != nullptr && !ref_knn_masks[(size_t)m * K + i]) continue;
    rows_valid++;
    const float* p = ref_knn + ((size_t)m * K + i) * 3;
    const float x = p[0], y = p[1], z = p[2], a2 = sq3(x, y, z);
    bool any = false;
    for (int j = 0; j < K; j++) {
      const float4 s = s_src[j];
      if (s.w < 0.f) continue;
      if (pdist(x, y, z, a2, s.x, s.y, s.z, s.w) < r2) {
        any = true;
        s_colhit[j] = 1u;  // benign race: every writer stores 1
      }
    }
    row_hits += any ? 1 : 0;
  }
  atomicAdd(&s_cnt[0], row_hits);
  atomicAdd(&s_cnt[1], rows_valid);
  __syncthreads();
  int col_hits = 0, cols_valid = 0;
  for (int j = tid; j < K; j += 128) {
    cols_valid += s_src[j].w >= 0.f ? 1 : 0;
    col_hits += (int)s_colhit[j];
  }
  atomicAdd(&s_cnt[2], col_hits);
  atomicAdd(&s_cnt[3], cols_valid);
  __syncthreads();
  if (tid == 0) {
    // float division by a zero count gives nan/inf in the reference as well (:341-342); nodes without points are
    // excluded by ref_masks / src_masks in every caller
    const float ro = (float)s_cnt[0] / (float)s_cnt[1], so = (float)s_cnt[2] / (float)s_cnt[3];
    overlaps[blockIdx.x] = (ro + so) / 2.f;
  }
}

// row-major compaction of the entries > 0 of a [rows, cols] matrix: (C,2) int64 indices + values, *out_count = C
__global__ void __launch_bounds__(1024) compact_nonzero_kernel(const float* __restrict__ mat, const unsigned char* __restrict__ bmat,
                                                               int rows, int cols, int64_t* __restrict__ out_idx,
                                                               float* __restrict__ out_val, int* __restrict__ out_count) {
  __shared__ int s_scan[33];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const long long total = (long long)rows * cols;
  for (long long e0 = 0; e0 < total; e0 += 1024) {
    const long long e = e0 + threadIdx.x;
    float v = 0.f;
    bool nz = false;
    if (e < total) {
      if (mat != nullptr) {
        v = mat[e];
        nz = v > 0.f;
      } else {
        nz = bmat[e] != 0;
      }
    }
    int tot;
    const int pos = block_exclusive_scan(nz ? 1 : 0, s_scan, &tot);
    const int base = s_base;
    if (nz) {
      const long long o = base + pos;
      out_idx[2 * o] = e / cols;
      out_idx[2 * o + 1] = e % cols;
      if (out_val != nullptr) out_val[o] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_count = s_base;
}

// get_node_correspondences_disance: mask[m, n] = (n is m's nearest src node and d < r) or (m is n's nearest ref node and
// d < r), AND node masks. NOTE (:484, :489) the reference compares the SQUARED distance with pos_radius itself.
__global__ void __launch_bounds__(256) nearest_mask_kernel(const float* __restrict__ a, int na, const float* __restrict__ b_raw, int nb,
                                                           const float* __restrict__ Tb, const float* __restrict__ Ta, float thr,
                                                           int a_is_row, int cols, const unsigned char* __restrict__ row_masks,
                                                           const unsigned char* __restrict__ col_masks, unsigned char* __restrict__ mask) {
  // one warp per element of `a`; b (and a) are transformed on the fly when Tb (Ta) is given
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= na) return;
  const float3 p = xform(Ta, a[3 * i], a[3 * i + 1], a[3 * i + 2]);
  const float p2 = sq3(p.x, p.y, p.z);
  float best = 3.4e38f;
  int bi = 0x7fffffff;
  for (int j = lane; j < nb; j += 32) {
    const float3 q = xform(Tb, b_raw[3 * j], b_raw[3 * j + 1], b_raw[3 * j + 2]);
    const float q2 = sq3(q.x, q.y, q.z);
    // the matrix is pairwise_distance(ref, src): x = ref (rows), y = src (cols) in either orientation
    const float d = a_is_row ? pdist(p.x, p.y, p.z, p2, q.x, q.y, q.z, q2) : pdist(q.x, q.y, q.z, q2, p.x, p.y, p.z, p2);
    if (d < best) {
      best = d;
      bi = j;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {  // torch.min returns the first minimum
    const float ob = __shfl_xor_sync(FULL_MASK, best, o);
    const int oi = __shfl_xor_sync(FULL_MASK, bi, o);
    if (ob < best || (ob == best && oi < bi)) {
      best = ob;
      bi = oi;
    }
  }
  if (lane == 0 && nb > 0 && best < thr) {
    const int r = a_is_row ? i : bi, c = a_is_row ? bi : i;
    if ((row_masks == nullptr || row_masks[r]) && (col_masks == nullptr || col_masks[c])) mask[(size_t)r * cols + c] = 1;
  }
}
}  // namespace

extern "C" size_t rdm_node_correspondences_workspace(int M, int N) {
  return align_up((size_t)M * 4, 256) + align_up((size_t)N * 4, 256) + align_up((size_t)M * 12, 256) + align_up((size_t)N * 12, 256) + 1024;
}

extern "C" int rdm_node_correspondences(const float* ref_nodes, const float* src_nodes, const float* ref_knn_points,
                                        const float* src_knn_points, const float* transform, float pos_radius, int M, int N, int K,
                                        const unsigned char* ref_masks, const unsigned char* src_masks,
                                        const unsigned char* ref_knn_masks, const unsigned char* src_knn_masks, int sphere_filter,
                                        float* out_overlaps, unsigned char* out_intersect, int64_t* out_corr_indices,
                                        float* out_corr_overlaps, int* out_count, void* workspace, size_t workspace_bytes,
                                        cudaStream_t stream) {
  RDM_CHECK_ARG(M >= 0 && N >= 0 && K >= 1 && K <= 256, "rdm_node_correspondences: K must be in [1, 256]");
  if (M == 0 || N == 0) {
    if (out_count) RDM_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int), stream));
    return RDM_OK;
  }
  Workspace ws(workspace, workspace_bytes);
  float* rr = ws.get<float>(M);
  float* sr = ws.get<float>(N);
  float* rn = ws.get<float>((size_t)M * 3);
  float* sn = ws.get<float>((size_t)N * 3);
  if (!ws.ok) {
    rdm_set_error("rdm_node_correspondences: workspace too small");
    return RDM_ERR_WORKSPACE;
  }
  patch_radius_kernel<<<cdiv(M, 8), 256, 0, stream>>>(ref_nodes, ref_knn_points, ref_knn_masks, nullptr, M, K, rn, rr);
  RDM_LAUNCH_CHECK();
  patch_radius_kernel<<<cdiv(N, 8), 256, 0, stream>>>(src_nodes, src_knn_points, src_knn_masks, transform, N, K, sn, sr);
  RDM_LAUNCH_CHECK();
  patch_overlap_kernel<256><<<M * N, 128, 0, stream>>>(rn, sn, rr, sr, ref_knn_points, src_knn_points, ref_masks, src_masks,
                                                       ref_knn_masks, src_knn_masks, transform, M, N, K, pos_radius, sphere_filter,
                                                       out_overlaps, out_intersect);
  RDM_LAUNCH_CHECK();
  if (out_corr_indices != nullptr) {
    compact_nonzero_kernel<<<1, 1024, 0, stream>>>(out_overlaps, nullptr, M, N, out_corr_indices, out_corr_overlaps, out_count);
    RDM_LAUNCH_CHECK();
  }
  return RDM_OK;
}

extern "C" int rdm_compact_nonzero(const float* mat, const unsigned char* bmat, int rows, int cols, int64_t* out_indices,
                                   float* out_values, int* out_count, cudaStream_t stream) {
  RDM_CHECK_ARG((mat != nullptr) != (bmat != nullptr), "rdm_compact_nonzero: give exactly one of mat / bmat");
  RDM_CHECK_ARG(rows >= 0 && cols >= 0, "rdm_compact_nonzero: bad sizes");
  compact_nonzero_kernel<<<1, 1024, 0, stream>>>(mat, bmat, rows, cols, out_indices, out_values, out_count);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

extern "C" int rdm_node_distance_mask(const float* ref_nodes, const float* src_nodes, const float* transform, float pos_radius, int M,
                                      int N, const unsigned char* ref_masks, const unsigned char* src_masks, unsigned char* out_mask,
                                      cudaStream_t stream) {
  RDM_CHECK_ARG(M >= 0 && N >= 0, "rdm_node_distance_mask: bad sizes");
  if (M == 0 || N == 0) return RDM_OK;
  RDM_CUDA(cudaMemsetAsync(out_mask, 0, (size_t)M * N, stream));
  nearest_mask_kernel<<<cdiv(M, 8), 256, 0, stream>>>(ref_nodes, M, src_nodes, N, transform, nullptr, pos_radius, 1, N, ref_masks,
                                                      src_masks, out_mask);
  RDM_LAUNCH_CHECK();
  nearest_mask_kernel<<<cdiv(N, 8), 256, 0, stream>>>(src_nodes, N, ref_nodes, M, nullptr, transform, pos_radius, 0, N, ref_masks,
                                                      src_masks, out_mask);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}
