// Superpoint detection / matching kernels: greedy NMS, point-to-node partition, coarse (superpoint) matching,
// patch score GEMM with row gather, and the log-domain Sinkhorn solver.
//
// Reference semantics:
//   NMS greedy loop             rdmnet/vote/vote.py:33-40
//   point_to_node_partition     geotransformer/modules/ops/pointcloud_partition.py:60-107 (+ pairwise_distance.py:4-31)
//   SuperPointMatching.forward  geotransformer/modules/geotransformer/superpoint_matching.py:14-83
//   patch gather + einsum       experiments/model.py:323-343
//   LearnableLogOptimalTransport geotransformer/modules/sinkhorn/learnable_sinkhorn.py:13-66
#include "common.cuh"
#include "../../include/rdm_sm100.h"

// ------------------------------------------------------------------------------------------------------ NMS
// sel[i] = !any(sel[j] for j in nbrs(i)), i = 0..N-1 in order (sel starts all-false; entries >= N are the sentinel).
// One warp walks the nodes sequentially; neighbour rows are prefetched PF rows at a time (they do not depend on sel).
// Parallel fixpoint of the greedy rule. sel[i] = !any(sel[j], j in nbrs(i)) evaluated for i = 0..N-1 with sel initially
// all False means: only neighbours j < i can matter, i is selected iff all of them are rejected, rejected iff one of
// them is selected. That is a DAG in index order; every round settles the nodes whose earlier neighbours are all
// settled (state only moves unknown -> final, so racing reads are harmless), and the result is exactly the sequential
// one. Rounds needed = longest dependency chain (tens for spatial data) instead of N sequential steps.
// Epilogue: the selected indices compacted in index order + counts below / at-or-above `split` (the ref | src
// boundary), so the host needs one small readback instead of two nonzero() round trips.
template <typename IdxT>
__global__ void __launch_bounds__(1024) nms_kernel(const IdxT* __restrict__ nbr, int N, int H, int split,
                                                   unsigned char* __restrict__ mask, int64_t* __restrict__ out_sel,
                                                   int* __restrict__ out_counts) {
  extern __shared__ unsigned char s_state[];  // N bytes: 0 unknown, 1 selected, 2 rejected
  __shared__ int s_scan[33];
  volatile unsigned char* st = s_state;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < N; i += nt) st[i] = 0;
  __syncthreads();
  for (int round = 0; round <= N; round++) {
    int pending = 0;
    for (int i = tid; i < N; i += nt) {
      if (st[i] != 0) continue;
      bool rej = false, wait = false;
      const IdxT* row = nbr + (size_t)i * H;
      for (int h = 0; h < H; h++) {
        const long long j = (long long)row[h];
        if (j >= i) continue;  // itself, later nodes and the padding value are False when i is evaluated
        const unsigned char sj = st[j];
        if (sj == 1) {
          rej = true;
          break;
        }
        wait |= (sj == 0);
      }
      if (rej) st[i] = 2;
      else if (!wait) st[i] = 1;
      else pending = 1;
    }
    if (!__syncthreads_or(pending)) break;
  }
  // mask + compaction in index order
  const int chunk = (N + nt - 1) / nt;
  const int beg = min(N, tid * chunk), end = min(N, beg + chunk);
  int c = 0, c_lo = 0;
  for (int i = beg; i < end; i++) {
    const int v = st[i] == 1;
    mask[i] = (unsigned char)v;
    c += v;
    c_lo += v && i < split;
  }
  int total;
  int pre = block_exclusive_scan(c, s_scan, &total);
  if (out_sel != nullptr)
    for (int i = beg; i < end; i++)
      if (st[i] == 1) out_sel[pre++] = i;
  if (out_counts != nullptr) {
    int total_lo;
    block_exclusive_scan(c_lo, s_scan, &total_lo);
    if (tid == 0) {
      out_counts[0] = total_lo;
      out_counts[1] = total - total_lo;
    }
  }
}

extern "C" int rdm_nms(const void* neighbor_indices, int index_bytes, int N, int H, int split, unsigned char* out_mask,
                       int64_t* out_selected, int* out_counts, cudaStream_t stream) {
  RDM_CHECK_ARG(index_bytes == 4 || index_bytes == 8, "rdm_nms: index_bytes must be 4 or 8");
  RDM_CHECK_ARG(H >= 1, "rdm_nms: empty table");
  RDM_CHECK_ARG(N >= 0 && N <= 200000, "rdm_nms: too many nodes");
  if (N == 0) {
    if (out_counts) RDM_CUDA(cudaMemsetAsync(out_counts, 0, 2 * sizeof(int), stream));
    return RDM_OK;
  }
  size_t smem = (size_t)N;
  if (index_bytes == 8) {
    if (smem > 40 * 1024)
      RDM_CUDA(cudaFuncSetAttribute(nms_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_kernel<int64_t><<<1, 1024, smem, stream>>>((const int64_t*)neighbor_indices, N, H, split, out_mask, out_selected, out_counts);
  } else {
    if (smem > 40 * 1024)
      RDM_CUDA(cudaFuncSetAttribute(nms_kernel<int>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_kernel<int><<<1, 1024, smem, stream>>>((const int*)neighbor_indices, N, H, split, out_mask, out_selected, out_counts);
  }
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// ---------------------------------------------------------------------------------------- point_to_node_partition
// d(node, point) = max(1e-12, (|n|^2 - 2 n.p) + |p|^2)   (pairwise_distance.py:24-30)
__device__ __forceinline__ float p2n_dist(float nx, float ny, float nz, float n2, float px, float py, float pz,
                                          float p2) {
  float xy = fmaf(nz, pz, fmaf(ny, py, __fmul_rn(nx, px)));
  float d = __fadd_rn(__fsub_rn(n2, __fmul_rn(2.0f, xy)), p2);
  return fmaxf(d, 1e-12f);
}
__device__ __forceinline__ float sq3(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

__global__ void __launch_bounds__(256) p2n_assign_kernel(const float* __restrict__ pts, int NP,
                                                         const float* __restrict__ nodes, int NN,
                                                         int* __restrict__ p2n, float* __restrict__ pdist,
                                                         int* __restrict__ node_cnt) {
  __shared__ float4 s_nodes[256];
  int i = blockIdx.x * 256 + threadIdx.x;
  float px = 0, py = 0, pz = 0, p2 = 0;
  if (i < NP) {
    px = pts[3 * i];
    py = pts[3 * i + 1];
    pz = pts[3 * i + 2];
    p2 = sq3(px, py, pz);
  }
  float best = 3.4e38f;
  int bi = 0;
  for (int n0 = 0; n0 < NN; n0 += 256) {
    __syncthreads();
    int n = n0 + threadIdx.x;
    if (n < NN) {
      float x = nodes[3 * n], y = nodes[3 * n + 1], z = nodes[3 * n + 2];
      s_nodes[threadIdx.x] = make_float4(x, y, z, sq3(x, y, z));
    }
    __syncthreads();
    int lim = min(256, NN - n0);
    for (int k = 0; k < lim; k++) {
      float4 nd = s_nodes[k];
      float d = p2n_dist(nd.x, nd.y, nd.z, nd.w, px, py, pz, p2);
      if (d < best) {  // strict: the lowest node index wins ties (torch.min on CPU)
        best = d;
        bi = n0 + k;
      }
    }
  }
  if (i < NP) {
    p2n[i] = bi;
    pdist[i] = best;
    atomicAdd(&node_cnt[bi], 1);
  }
}

__global__ void __launch_bounds__(1024) p2n_scan_kernel(int* __restrict__ node_cnt, int* __restrict__ node_off, int NN) {
  __shared__ int s_scan[33];
  int nt = blockDim.x, tid = threadIdx.x;
  int chunk = (NN + nt - 1) / nt;
  int beg = min(NN, tid * chunk), end = min(NN, beg + chunk);
  int s = 0;
  for (int i = beg; i < end; i++) s += node_cnt[i];
  int total;
  int pre = block_exclusive_scan(s, s_scan, &total);
  for (int i = beg; i < end; i++) {
    int v = node_cnt[i];
    node_off[i] = pre;
    pre += v;
    node_cnt[i] = 0;  // becomes the fill cursor
  }
  if (tid == 0) node_off[NN] = total;
}

__global__ void p2n_scatter_kernel(const int* __restrict__ p2n, const float* __restrict__ pdist, int NP,
                                   const int* __restrict__ node_off, int* __restrict__ node_cur,
                                   unsigned long long* __restrict__ keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NP) return;
  int n = p2n[i];
  int pos = node_off[n] + atomicAdd(&node_cur[n], 1);
  keys[pos] = ((unsigned long long)__float_as_uint(pdist[i]) << 32) | (unsigned int)i;
}

// one warp per node: the K owned points with the smallest (distance, index)
__global__ void __launch_bounds__(128) p2n_select_kernel(const unsigned long long* __restrict__ keys,
                                                         const int* __restrict__ node_off, int NN, int NP, int K, int KP,
                                                         int64_t* __restrict__ knn_idx, unsigned char* __restrict__ knn_mask,
                                                         unsigned char* __restrict__ node_mask) {
  extern __shared__ unsigned long long s_buf[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 4 + warp;
  if (n >= NN) return;
  unsigned long long* buf = s_buf + (size_t)warp * 2 * KP;
  const int beg = node_off[n], end = node_off[n + 1];
  int cnt = 0;
  for (int p0 = beg; p0 < end; p0 += 32) {
    int p = p0 + lane;
    if (cnt + 32 > 2 * KP) {
      for (int i = cnt + lane; i < 2 * KP; i += 32) buf[i] = ~0ULL;
      __syncwarp();
      warp_bitonic_sort_u64(buf, 2 * KP, lane);
      cnt = min(cnt, KP);
    }
    if (p < end) buf[cnt + lane] = keys[p];
    cnt += min(32, end - p0);
    __syncwarp();
  }
  int P = 32;
  while (P < cnt) P <<= 1;
  for (int i = cnt + lane; i < P; i += 32) buf[i] = ~0ULL;
  __syncwarp();
  warp_bitonic_sort_u64(buf, P, lane);
  for (int i = lane; i < K; i += 32) {
    bool ok = i < cnt;
    knn_idx[(size_t)n * K + i] = ok ? (int64_t)(unsigned int)(buf[i] & 0xffffffffULL) : (int64_t)NP;
    knn_mask[(size_t)n * K + i] = ok ? 1 : 0;
  }
  if (lane == 0) node_mask[n] = end > beg ? 1 : 0;
}

extern "C" size_t rdm_point_to_node_workspace(int num_points, int num_nodes) {
  return align_up((size_t)num_points * 4, 256) * 2 + align_up((size_t)(num_nodes + 1) * 4, 256) * 2 +
         align_up((size_t)num_points * 8, 256) + 1024;
}

extern "C" int rdm_point_to_node(const float* points, int num_points, const float* nodes, int num_nodes, int point_limit,
                                 int* out_point_to_node, unsigned char* out_node_masks, int64_t* out_knn_indices,
                                 unsigned char* out_knn_masks, void* workspace, size_t workspace_bytes,
                                 cudaStream_t stream) {
  RDM_CHECK_ARG(num_points >= 0 && num_nodes >= 1 && point_limit >= 1 && point_limit <= 2048,
                "rdm_point_to_node: bad arguments");
  Workspace ws(workspace, workspace_bytes);
  float* pdist = ws.get<float>(num_points);
  int* cnt = ws.get<int>(num_nodes + 1);
  int* off = ws.get<int>(num_nodes + 1);
  unsigned long long* keys = ws.get<unsigned long long>(num_points);
  if (!ws.ok) {
    rdm_set_error("rdm_point_to_node: workspace too small");
    return RDM_ERR_WORKSPACE;
  }
  RDM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (num_nodes + 1), stream));
  if (num_points > 0) {
    p2n_assign_kernel<<<cdiv(num_points, 256), 256, 0, stream>>>(points, num_points, nodes, num_nodes, out_point_to_node,
                                                                pdist, cnt);
    RDM_LAUNCH_CHECK();
  }
  p2n_scan_kernel<<<1, 1024, 0, stream>>>(cnt, off, num_nodes);
  RDM_LAUNCH_CHECK();
  if (num_points > 0) {
    p2n_scatter_kernel<<<cdiv(num_points, 256), 256, 0, stream>>>(out_point_to_node, pdist, num_points, off, cnt, keys);
    RDM_LAUNCH_CHECK();
  }
  int KP = 32;
  while (KP < point_limit) KP <<= 1;
  size_t smem = (size_t)4 * 2 * KP * 8;
  if (smem > 48 * 1024)
    RDM_CUDA(cudaFuncSetAttribute(p2n_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p2n_select_kernel<<<cdiv(num_nodes, 4), 128, smem, stream>>>(keys, off, num_nodes, num_points, point_limit, KP,
                                                              out_knn_indices, out_knn_masks, out_node_masks);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// ------------------------------------------------------------------------------------------- coarse matching
// xy[M,N] (= ref_feats @ src_feats^T, computed by rdm_linear) -> S = exp(-max(1e-12, 2 - 2 xy)) on valid (i,j), else 0;
// row sums (one warp per row).
__global__ void __launch_bounds__(256) cm_exp_rowsum_kernel(float* __restrict__ S, int M, int N,
                                                            const unsigned char* __restrict__ rmask,
                                                            const unsigned char* __restrict__ cmask,
                                                            float* __restrict__ rowsum) {
  int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= M) return;
  bool rv = rmask[i] != 0;
  float s = 0.f;
  for (int j = lane; j < N; j += 32) {
    float v = 0.f;
    if (rv && cmask[j]) {
      float d = fmaxf(2.0f - 2.0f * S[(size_t)i * N + j], 1e-12f);
      v = expf(-d);
    }
    S[(size_t)i * N + j] = v;
    s += v;
  }
  s = warp_sum(s);
  if (lane == 0) rowsum[i] = s;
}
__global__ void cm_colsum_kernel(const float* __restrict__ S, int M, int N, float* __restrict__ colsum) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  float s = 0.f;
  for (int i = 0; i < M; i++) s += S[(size_t)i * N + j];
  colsum[j] = s;
}
// score = (S/rowsum) * (S/colsum); invalid entries get -1 so that they never enter the top-k
__global__ void cm_normalise_kernel(float* __restrict__ S, int M, int N, const unsigned char* __restrict__ rmask,
                                    const unsigned char* __restrict__ cmask, const float* __restrict__ rowsum,
                                    const float* __restrict__ colsum, int dual) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)M * N) return;
  int i = (int)(e / N), j = (int)(e - (long long)i * N);
  float v = -1.f;
  if (rmask[i] && cmask[j]) {
    float s = S[e];
    v = dual ? (s / rowsum[i]) * (s / colsum[j]) : s;
  }
  S[e] = v;
}

// Single-CTA exact top-k (largest) over n floats >= 0 (negatives = excluded): 4-pass radix select on the float
// bits, then the selected (value desc, index asc) are sorted with a block bitonic sort. k <= 1024.
__global__ void __launch_bounds__(1024) topk_flat_kernel(const float* __restrict__ v, long long n, int k, int N,
                                                         const int* __restrict__ unused, int64_t* __restrict__ out_i,
                                                         int64_t* __restrict__ out_j, float* __restrict__ out_score,
                                                         int* __restrict__ out_count) {
  __shared__ unsigned int s_hist[256];
  __shared__ unsigned int s_prefix, s_want;
  __shared__ int s_nsel, s_neq;
  __shared__ unsigned long long s_keys[1024];
  const int tid = threadIdx.x;
  // number of candidates (>= 0)
  if (tid == 0) s_nsel = 0;
  __syncthreads();
  int local = 0;
  for (long long e = tid; e < n; e += 1024) local += v[e] >= 0.f;
  local = warp_sum_i(local);
  if ((tid & 31) == 0) atomicAdd(&s_nsel, local);
  __syncthreads();
  const int nvalid = s_nsel;
  const int kk = min(k, nvalid);
  if (tid == 0) *out_count = kk;
  if (kk == 0) return;
  // radix select: find the bit pattern T of the kk-th largest value
  unsigned int prefix = 0, want = kk;  // want = rank (1-based) among values matching the prefix so far
  for (int pass = 0; pass < 4; pass++) {
    int shift = 24 - 8 * pass;
    if (tid < 256) s_hist[tid] = 0;
    __syncthreads();
    unsigned int pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (long long e = tid; e < n; e += 1024) {
      float f = v[e];
      if (f >= 0.f) {
        unsigned int b = __float_as_uint(f);
        if ((b & pmask) == prefix) atomicAdd(&s_hist[(b >> shift) & 255], 1u);
      }
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int acc = 0;
      int d = 255;
      for (; d >= 0; d--) {
        if (acc + s_hist[d] >= want) break;
        acc += s_hist[d];
      }
      s_prefix = prefix | ((unsigned int)d << shift);
      s_want = want - acc;
    }
    __syncthreads();
    prefix = s_prefix;
    want = s_want;
    __syncthreads();
  }
  // prefix = bits of the kk-th largest value; take all strictly greater, and the `want` lowest-index equal ones
  if (tid == 0) {
    s_nsel = 0;
    s_neq = 0;
  }
  for (int i = tid; i < 1024; i += 1024) s_keys[i] = ~0ULL;
  __syncthreads();
  for (long long e0 = 0; e0 < n; e0 += 1024) {
    long long e = e0 + tid;
    bool gt = false;
    if (e < n) {
      float f = v[e];
      gt = f >= 0.f && __float_as_uint(f) > prefix;
    }
    if (gt) {
      int pos = atomicAdd(&s_nsel, 1);
      // sort key: descending value (invert bits), ascending flat index
      s_keys[pos] = ((unsigned long long)(~__float_as_uint(v[e])) << 32) | (unsigned int)e;
    }
  }
  __syncthreads();
  // equal ones in ascending index order: sequential chunks keep the order deterministic
  for (long long e0 = 0; e0 < n; e0 += 1024) {
    long long e = e0 + tid;
    bool eq = false;
    if (e < n) {
      float f = v[e];
      eq = f >= 0.f && __float_as_uint(f) == prefix;
    }
    int total;
    int pre = block_exclusive_scan(eq ? 1 : 0, (int*)s_hist, &total);
    int base = s_neq;
    if (eq && base + pre < (int)want) {
      int pos = atomicAdd(&s_nsel, 1);
      s_keys[pos] = ((unsigned long long)(~prefix) << 32) | (unsigned int)e;
    }
    __syncthreads();
    if (tid == 0) s_neq = base + total;
    __syncthreads();
    if (s_neq >= (int)want) break;
  }
  __syncthreads();
  // block bitonic sort of 1024 keys
  for (int kk2 = 2; kk2 <= 1024; kk2 <<= 1) {
    for (int j = kk2 >> 1; j > 0; j >>= 1) {
      int t = tid;
      if (t < 512) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int p = i | j;
        unsigned long long x = s_keys[i], y = s_keys[p];
        bool up = ((i & kk2) == 0);
        if ((x > y) == up) {
          s_keys[i] = y;
          s_keys[p] = x;
        }
      }
      __syncthreads();
    }
  }
  if (tid < kk) {
    unsigned long long key = s_keys[tid];
    unsigned int e = (unsigned int)(key & 0xffffffffULL);
    out_i[tid] = e / N;
    out_j[tid] = e % N;
    out_score[tid] = __uint_as_float(~(unsigned int)(key >> 32));
  }
}

extern "C" int rdm_coarse_matching(float* xy_scores, int M, int N, const unsigned char* ref_masks,
                                   const unsigned char* src_masks, int num_correspondences, int dual_normalization,
                                   int64_t* out_ref_indices, int64_t* out_src_indices, float* out_scores, int* out_count,
                                   float* sums_scratch, cudaStream_t stream) {
  RDM_CHECK_ARG(M >= 1 && N >= 1 && num_correspondences >= 1 && num_correspondences <= 1024,
                "rdm_coarse_matching: num_correspondences must be in [1,1024]");
  RDM_CHECK_ARG((long long)M * N < (1LL << 31), "rdm_coarse_matching: score matrix too large");
  float* rowsum = sums_scratch;
  float* colsum = sums_scratch + M;
  cm_exp_rowsum_kernel<<<cdiv(M, 8), 256, 0, stream>>>(xy_scores, M, N, ref_masks, src_masks, rowsum);
  RDM_LAUNCH_CHECK();
  cm_colsum_kernel<<<cdiv(N, 128), 128, 0, stream>>>(xy_scores, M, N, colsum);
  RDM_LAUNCH_CHECK();
  cm_normalise_kernel<<<cdiv((long long)M * N, 256), 256, 0, stream>>>(xy_scores, M, N, ref_masks, src_masks, rowsum,
                                                                      colsum, dual_normalization);
  RDM_LAUNCH_CHECK();
  topk_flat_kernel<<<1, 1024, 0, stream>>>(xy_scores, (long long)M * N, num_correspondences, N, nullptr,
                                          out_ref_indices, out_src_indices, out_scores, out_count);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// ------------------------------------------------------------------------------------------- patch scores
// out[p, i, j] = scale * < F_ref[rknn[ridx[p], i]], F_src[sknn[sidx[p], j]] >  ; rows with index >= N are the zero row.
// One CTA per patch; K x K tile with K = 128 (point limit), 256 threads x (8x8), BK = 16.
#define PS_K 128
__global__ void __launch_bounds__(256) patch_scores_kernel(const float* __restrict__ Fr, int Nr,
                                                           const float* __restrict__ Fs, int Ns, int C, int ld,
                                                           const int64_t* __restrict__ rknn,
                                                           const int64_t* __restrict__ sknn,
                                                           const int64_t* __restrict__ ridx,
                                                           const int64_t* __restrict__ sidx, float scale,
                                                           float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float As[2][16][PS_K + 4];
  __shared__ __align__(16) float Bs[2][16][PS_K + 4];
  __shared__ int s_ra[PS_K], s_rb[PS_K];
  const int p = blockIdx.x, tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  if (tid < PS_K) {
    long long a = rknn[(size_t)ridx[p] * PS_K + tid], b = sknn[(size_t)sidx[p] * PS_K + tid];
    s_ra[tid] = a < Nr ? (int)a : -1;
    s_rb[tid] = b < Ns ? (int)b : -1;
  }
  __syncthreads();
  float4 ra[2], rb[2];
  auto load = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 2; q++) {
      int row = (tid >> 2) + q * 64, kq = (tid & 3) * 4;
      int ia = s_ra[row], ib = s_rb[row];
      ra[q] = ia >= 0 ? *(const float4*)(Fr + (size_t)ia * ld + k0 + kq) : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[q] = ib >= 0 ? *(const float4*)(Fs + (size_t)ib * ld + k0 + kq) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 2; q++) {
      int row = (tid >> 2) + q * 64, kq = (tid & 3) * 4;
      As[buf][kq + 0][row] = ra[q].x; As[buf][kq + 1][row] = ra[q].y; As[buf][kq + 2][row] = ra[q].z; As[buf][kq + 3][row] = ra[q].w;
      Bs[buf][kq + 0][row] = rb[q].x; Bs[buf][kq + 1][row] = rb[q].y; Bs[buf][kq + 2][row] = rb[q].z; Bs[buf][kq + 3][row] = rb[q].w;
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
  const int nk = C / 16;
  load(0);
  store(0);
  __syncthreads();
  for (int t = 0; t < nk; t++) {
    int buf = t & 1;
    if (t + 1 < nk) load((t + 1) * 16);
#pragma unroll
    for (int k = 0; k < 16; k++) {
      float4 a0 = *(const float4*)&As[buf][k][ty * 4], a1 = *(const float4*)&As[buf][k][64 + ty * 4];
      float4 b0 = *(const float4*)&Bs[buf][k][tx * 4], b1 = *(const float4*)&Bs[buf][k][64 + tx * 4];
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < nk) store(buf ^ 1);
    __syncthreads();
  }
  float* op = out + (size_t)p * PS_K * PS_K;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int row = i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4);
#pragma unroll
    for (int jh = 0; jh < 2; jh++) {
      int col = jh == 0 ? tx * 4 : 64 + tx * 4;
      *(float4*)(op + (size_t)row * PS_K + col) = make_float4(acc[i][jh * 4] * scale, acc[i][jh * 4 + 1] * scale,
                                                              acc[i][jh * 4 + 2] * scale, acc[i][jh * 4 + 3] * scale);
    }
  }
}

extern "C" int rdm_patch_scores(const float* ref_feats, int Nr, const float* src_feats, int Ns, int C, int ld_feats,
                                const int64_t* ref_knn_indices, const int64_t* src_knn_indices,
                                const int64_t* ref_corr_indices, const int64_t* src_corr_indices, int num_patches,
                                int point_limit, float scale, float* out_scores, cudaStream_t stream) {
  RDM_CHECK_ARG(point_limit == PS_K, "rdm_patch_scores: num_points_in_patch must be 128");
  RDM_CHECK_ARG(C % 16 == 0 && C >= 16, "rdm_patch_scores: C must be a multiple of 16");
  RDM_CHECK_ARG(ld_feats >= C && ld_feats % 4 == 0, "rdm_patch_scores: feature row stride must be a multiple of 4 floats");
  RDM_CHECK_ARG(((uintptr_t)ref_feats & 15) == 0 && ((uintptr_t)src_feats & 15) == 0, "rdm_patch_scores: unaligned features");
  if (num_patches == 0) return RDM_OK;
  RDM_CUDA(rdm_launch_pdl(patch_scores_kernel, dim3(num_patches), dim3(256), 0, stream, ref_feats, Nr, src_feats, Ns, C, ld_feats,
                          ref_knn_indices, src_knn_indices, ref_corr_indices, src_corr_indices, scale, out_scores));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// ------------------------------------------------------------------------------------------------- Sinkhorn
// One CTA per patch pair; the padded (R+1)x(Cc+1) score matrix stays in shared memory for all iterations.
// Row phase: thread r owns row r; column phase: thread c owns column c (row stride odd -> conflict-free both ways).
__global__ void __launch_bounds__(160) sinkhorn_kernel(const float* __restrict__ scores, int R, int Cc,
                                                       const unsigned char* __restrict__ row_masks_nodes,
                                                       const unsigned char* __restrict__ col_masks_nodes,
                                                       const int64_t* __restrict__ ridx, const int64_t* __restrict__ sidx,
                                                       const float* __restrict__ alpha_ptr, int iters, float inf,
                                                       float* __restrict__ out) {
  extern __shared__ float s_f[];
  const int R1 = R + 1, C1 = Cc + 1;
  const int ld = (C1 % 2 == 0) ? C1 + 1 : C1;  // odd stride
  float* Z = s_f;                               // R1 * ld
  float* u = Z + (size_t)R1 * ld;               // R1
  float* v = u + R1;                            // C1
  float* lmu = v + C1;                          // R1
  float* lnu = lmu + R1;                        // C1
  __shared__ int s_nr, s_nc;
  const int p = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const unsigned char* rm = ridx ? row_masks_nodes + (size_t)ridx[p] * R : row_masks_nodes + (size_t)p * R;
  const unsigned char* cm = sidx ? col_masks_nodes + (size_t)sidx[p] * Cc : col_masks_nodes + (size_t)p * Cc;
  const float alpha = *alpha_ptr;
  if (tid == 0) {
    s_nr = 0;
    s_nc = 0;
  }
  __syncthreads();
  {
    int a = 0, b = 0;
    for (int i = tid; i < R; i += nt) a += rm[i] != 0;
    for (int j = tid; j < Cc; j += nt) b += cm[j] != 0;
    if (a) atomicAdd(&s_nr, a);
    if (b) atomicAdd(&s_nc, b);
  }
  __syncthreads();
  const float nr = (float)s_nr, nc = (float)s_nc;
  const float norm = -logf(nr + nc);  // learnable_sinkhorn.py:49
  const float* sp = scores + (size_t)p * R * Cc;
  for (int e = tid; e < R1 * C1; e += nt) {
    int i = e / C1, j = e - i * C1;
    bool masked = (i < R && !rm[i]) || (j < Cc && !cm[j]);
    float z = (i < R && j < Cc) ? sp[(size_t)i * Cc + j] : alpha;
    Z[(size_t)i * ld + j] = masked ? -inf : z;
  }
  for (int i = tid; i < R1; i += nt) {
    u[i] = 0.f;
    lmu[i] = i < R ? (rm[i] ? norm : -inf) : logf(nc) + norm;
  }
  for (int j = tid; j < C1; j += nt) {
    v[j] = 0.f;
    lnu[j] = j < Cc ? (cm[j] ? norm : -inf) : logf(nr) + norm;
  }
  __syncthreads();
  for (int it = 0; it < iters; it++) {
    if (tid < R1) {  // u = log_mu - logsumexp_j(Z + v)
      const float* zr = Z + (size_t)tid * ld;
      float mx = -3.4e38f;
      for (int j = 0; j < C1; j++) mx = fmaxf(mx, zr[j] + v[j]);
      float s = 0.f;
      for (int j = 0; j < C1; j++) s += __expf(zr[j] + v[j] - mx);
      u[tid] = lmu[tid] - (mx + __logf(s));
    }
    __syncthreads();
    if (tid < C1) {  // v = log_nu - logsumexp_i(Z + u)
      float mx = -3.4e38f;
      for (int i = 0; i < R1; i++) mx = fmaxf(mx, Z[(size_t)i * ld + tid] + u[i]);
      float s = 0.f;
      for (int i = 0; i < R1; i++) s += __expf(Z[(size_t)i * ld + tid] + u[i] - mx);
      v[tid] = lnu[tid] - (mx + __logf(s));
    }
    __syncthreads();
  }
  float* op = out + (size_t)p * R1 * C1;
  for (int e = tid; e < R1 * C1; e += nt) {
    int i = e / C1, j = e - i * C1;
    op[e] = Z[(size_t)i * ld + j] + u[i] + v[j] - norm;
  }
}

// ---- fast path for 128 x 128 patches (num_points_in_patch = 128, experiments/config.py:100).
// The padded score matrix is constant over the 100 iterations; only the potentials u, v change. So every thread keeps
// its share of Z in REGISTERS, twice: as a piece of a row (for u = log_mu - LSE_j(Z + v)) and as a piece of a column
// (for v = log_nu - LSE_i(Z + u)); shared memory only carries the two potential vectors. Masked rows / columns
// (-1e12 entries: exp = 0 exactly, learnable_sinkhorn.py:36-47) never influence a live entry, so the live rows and
// columns (+ the dustbin) are compacted first and the iteration runs on the compacted matrix only. Element e of a
// compacted axis belongs to quad lane e & 3, slot e >> 2: four lanes share a row / column and fold their partial
// (max, sum) pairs with two shuffles. Everything is kept in the log2 domain (one MUFU.EX2 per element, no multiply).
#define SK_N 128
#define SK_T 33   // ceil(129 / 4) slots per lane
#define SK_LD 36  // floats between the four lanes' potential slices (16-byte aligned, conflict-free LDS.128)
#define SK_THREADS 544
#define SK_NEG -1.0e30f

__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2f(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// running (m, s) <- fold in z[k0..k0+3] + p[0..3]   (log2 domain, online logsumexp)
__device__ __forceinline__ void sk_fold4(const float* z, const float4 p, float& m, float& s) {
  const float t0 = z[0] + p.x, t1 = z[1] + p.y, t2 = z[2] + p.z, t3 = z[3] + p.w;
  const float mn = fmaxf(fmaxf(m, fmaxf(t0, t1)), fmaxf(t2, t3));
  s = s * ex2f(m - mn) + ((ex2f(t0 - mn) + ex2f(t1 - mn)) + (ex2f(t2 - mn) + ex2f(t3 - mn)));
  m = mn;
}

// one half-iteration: lanes (a, q) with a < n_out produce pot_out[a] = logm[a] - LSE_k(z + pot_in); kc = slots in use
__device__ __forceinline__ void sk_phase(const float (&z)[SK_T], const float* pot_in, float* pot_out, const float* logm, int a,
                                         int q, int n_out, int kc) {
  float m = SK_NEG, s = 0.f;
  if (a < n_out) {
    const float4* p4 = (const float4*)(pot_in + q * SK_LD);
#pragma unroll
    for (int g = 0; g < 8; g++)
      if (4 * g < kc) sk_fold4(&z[4 * g], p4[g], m, s);
    if (kc > 32) {
      const float t = z[32] + pot_in[q * SK_LD + 32];
      const float mn = fmaxf(m, t);
      s = s * ex2f(m - mn) + ex2f(t - mn);
      m = mn;
    }
  }
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {  // fold the four lanes of the row / column
    const float mo = __shfl_xor_sync(FULL_MASK, m, o), so = __shfl_xor_sync(FULL_MASK, s, o);
    const float mn = fmaxf(m, mo);
    s = s * ex2f(m - mn) + so * ex2f(mo - mn);
    m = mn;
  }
  if (a < n_out && q == 0) pot_out[(a & 3) * SK_LD + (a >> 2)] = logm[a] - (m + lg2f(s));
}

__global__ void __launch_bounds__(SK_THREADS, 1) sinkhorn128_kernel(const float* __restrict__ scores,
                                                                   const unsigned char* __restrict__ row_masks_nodes,
                                                                   const unsigned char* __restrict__ col_masks_nodes,
                                                                   const int64_t* __restrict__ ridx,
                                                                   const int64_t* __restrict__ sidx,
                                                                   const float* __restrict__ alpha_ptr, int iters, float inf,
                                                                   float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float us[4 * SK_LD], vs[4 * SK_LD];
  __shared__ float lmu[SK_N + 4], lnu[SK_N + 4];
  __shared__ short liveR[SK_N + 1], liveC[SK_N + 1], posR[SK_N + 1], posC[SK_N + 1];
  __shared__ int s_n[2];
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int a = tid >> 2, q = tid & 3;
  const unsigned char* rm = ridx ? row_masks_nodes + (size_t)ridx[p] * SK_N : row_masks_nodes + (size_t)p * SK_N;
  const unsigned char* cm = sidx ? col_masks_nodes + (size_t)sidx[p] * SK_N : col_masks_nodes + (size_t)p * SK_N;
  const float alpha = *alpha_ptr;
  const float* sp = scores + (size_t)p * SK_N * SK_N;
  const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  if (warp < 2) {  // warp 0 compacts the rows, warp 1 the columns (ballot scan); the dustbin is the last live element
    const unsigned char* mk = warp == 0 ? rm : cm;
    short* live = warp == 0 ? liveR : liveC;
    short* pos = warp == 0 ? posR : posC;
    int base = 0;
    for (int c = 0; c < SK_N; c += 32) {
      const bool on = mk[c + lane] != 0;
      const unsigned b = __ballot_sync(FULL_MASK, on);
      const int at = base + __popc(b & ((1u << lane) - 1u));
      if (on) live[at] = (short)(c + lane);
      pos[c + lane] = on ? (short)at : (short)-1;
      base += __popc(b);
    }
    if (lane == 0) {
      live[base] = SK_N;
      pos[SK_N] = (short)base;
      s_n[warp] = base + 1;
    }
  }
  for (int i = tid; i < 4 * SK_LD; i += SK_THREADS) {
    us[i] = 0.f;
    vs[i] = 0.f;
  }
  __syncthreads();
  const int nR = s_n[0], nC = s_n[1];
  const float nr = (float)(nR - 1), nc = (float)(nC - 1);
  const float norm = -logf(nr + nc);  // learnable_sinkhorn.py:49
  for (int i = tid; i < nR; i += SK_THREADS) lmu[i] = (i < nR - 1 ? norm : logf(nc) + norm) * LOG2E;
  for (int j = tid; j < nC; j += SK_THREADS) lnu[j] = (j < nC - 1 ? norm : logf(nr) + norm) * LOG2E;
  // register-resident shares of the compacted matrix (log2 domain); slots beyond the live extent hold SK_NEG (exp2 -> 0)
  float zr[SK_T], zc[SK_T];
#pragma unroll
  for (int k = 0; k < SK_T; k++) {
    const int e = 4 * k + q;
    float r = SK_NEG, c = SK_NEG;
    if (a < nR && e < nC) {
      const int i = liveR[a], j = liveC[e];
      r = ((i < SK_N && j < SK_N) ? __ldg(sp + i * SK_N + j) : alpha) * LOG2E;
    }
    if (a < nC && e < nR) {
      const int i = liveR[e], j = liveC[a];
      c = ((i < SK_N && j < SK_N) ? __ldg(sp + i * SK_N + j) : alpha) * LOG2E;
    }
    zr[k] = r;
    zc[k] = c;
  }
  __syncthreads();
  const int kcC = (nC + 3) >> 2, kcR = (nR + 3) >> 2;
  for (int it = 0; it < iters; it++) {
    sk_phase(zr, vs, us, lmu, a, q, nR, kcC);  // u = log_mu - logsumexp_j(Z + v)   (learnable_sinkhorn.py:24-25)
    __syncthreads();
    sk_phase(zc, us, vs, lnu, a, q, nC, kcR);  // v = log_nu - logsumexp_i(Z + u)
    __syncthreads();
  }
  float* op = out + (size_t)p * (SK_N + 1) * (SK_N + 1);
  for (int e = tid; e < (SK_N + 1) * (SK_N + 1); e += SK_THREADS) {
    const int i = e / (SK_N + 1), j = e - i * (SK_N + 1);
    const int pr = posR[i], pc = posC[j];
    float o = -inf;  // masked rows / columns stay at -1e12 (+ finite potentials in the reference): exp() = 0 either way
    if (pr >= 0 && pc >= 0) {
      const float z = (i < SK_N && j < SK_N) ? __ldg(sp + i * SK_N + j) : alpha;
      o = z + (us[(pr & 3) * SK_LD + (pr >> 2)] + vs[(pc & 3) * SK_LD + (pc >> 2)]) * LN2 - norm;
    }
    op[e] = o;
  }
}

extern "C" int rdm_sinkhorn(const float* scores, int num_patches, int R, int C, const unsigned char* row_masks,
                            const unsigned char* col_masks, const int64_t* row_mask_gather, const int64_t* col_mask_gather,
                            const float* alpha, int num_iterations, float inf, float* out, cudaStream_t stream) {
  RDM_CHECK_ARG(R >= 1 && C >= 1 && R <= 159 && C <= 159, "rdm_sinkhorn: patch size must be <= 159");
  if (num_patches == 0) return RDM_OK;
  if (R == SK_N && C == SK_N) {
    RDM_CUDA(rdm_launch_pdl(sinkhorn128_kernel, dim3(num_patches), dim3(SK_THREADS), 0, stream, scores, row_masks, col_masks,
                            row_mask_gather, col_mask_gather, alpha, num_iterations, inf, out));
    RDM_LAUNCH_CHECK();
    return RDM_OK;
  }
  int R1 = R + 1, C1 = C + 1, ld = (C1 % 2 == 0) ? C1 + 1 : C1;
  size_t smem = ((size_t)R1 * ld + 2 * R1 + 2 * C1) * sizeof(float);
  RDM_CHECK_ARG(smem <= 200 * 1024, "rdm_sinkhorn: patch too large for shared memory");
  if (smem > 48 * 1024)
    RDM_CUDA(cudaFuncSetAttribute(sinkhorn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sinkhorn_kernel<<<num_patches, 160, smem, stream>>>(scores, R, C, row_masks, col_masks, row_mask_gather,
                                                     col_mask_gather, alpha, num_iterations, inf, out);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// ------------------------------------------------------------------------------------------- tail helpers
// Vote_layer epilogue (rdmnet/vote/vote.py:101-116): off = ctr_reg(h) is [N, 3 + C]; xyz_out = xyz + clamp(off[:, :3],
// -limit, +limit); feat_out = LayerNorm(features + off[:, 3:]). One warp per row.
__global__ void __launch_bounds__(256) vote_finish_kernel(const float* __restrict__ off, int ld_off, const float* __restrict__ xyz,
                                                          const float* __restrict__ feats, int ld_f, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float lx, float ly, float lz, float eps,
                                                          int N, int C, float* __restrict__ xyz_out, float* __restrict__ feat_out) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= N) return;
  const float* o = off + (size_t)r * ld_off;
  if (lane < 3) {
    const float lim = lane == 0 ? lx : (lane == 1 ? ly : lz);
    xyz_out[3 * r + lane] = xyz[3 * r + lane] + fminf(fmaxf(o[lane], -lim), lim);
  }
  float v[32];  // C <= 1024
  float s = 0.f;
  const int per = (C + 31) / 32;
  for (int i = 0; i < per; i++) {
    const int c = lane + 32 * i;
    v[i] = c < C ? feats[(size_t)r * ld_f + c] + o[3 + c] : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int i = 0; i < per; i++) {
    const int c = lane + 32 * i;
    const float d = c < C ? v[i] - mean : 0.f;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  for (int i = 0; i < per; i++) {
    const int c = lane + 32 * i;
    if (c < C) feat_out[(size_t)r * C + c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
  }
}

int rdm_vote_finish(const float* off, int ld_off, const float* xyz, const float* feats, int ld_f, const float* gamma,
                    const float* beta, const float* h_limit3, float eps, int N, int C, float* xyz_out, float* feat_out,
                    cudaStream_t stream) {
  RDM_CHECK_ARG(C >= 1 && C <= 1024, "rdm_vote_finish: C must be <= 1024");
  if (N == 0) return RDM_OK;
  RDM_CUDA(rdm_launch_pdl(vote_finish_kernel, dim3(cdiv(N, 8)), dim3(256), 0, stream, off, ld_off, xyz, feats, ld_f, gamma, beta,
                          h_limit3[0], h_limit3[1], h_limit3[2], eps, N, C, xyz_out, feat_out));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// Row selection after NMS (experiments/model.py:233-246): up to 4 (src, dst, C) jobs gathered by the same index list;
// `normalize` L2-normalises job 1's rows into dst_norm (F.normalize(p=2, dim=1), eps 1e-12; model.py:261-264) - used
// AFTER the second transformer, so it is a separate flag on its own launch.
struct GatherJobs {
  const float* src[4];
  float* dst[4];
  int c[4], ld[4], n;
};
__global__ void __launch_bounds__(256) gather_rows_kernel(const GatherJobs jobs, const int64_t* __restrict__ sel, int count) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= count) return;
  const long long s = sel[r];
  for (int j = 0; j < jobs.n; j++)
    for (int c = lane; c < jobs.c[j]; c += 32) jobs.dst[j][(size_t)r * jobs.c[j] + c] = jobs.src[j][(size_t)s * jobs.ld[j] + c];
}
int rdm_gather_rows(const float* const* h_src, float* const* h_dst, const int* h_c, const int* h_ld, int num_jobs,
                    const int64_t* sel, int count, cudaStream_t stream) {
  RDM_CHECK_ARG(num_jobs >= 1 && num_jobs <= 4, "rdm_gather_rows: 1..4 jobs");
  if (count == 0) return RDM_OK;
  GatherJobs g;
  g.n = num_jobs;
  for (int i = 0; i < num_jobs; i++) {
    g.src[i] = h_src[i];
    g.dst[i] = h_dst[i];
    g.c[i] = h_c[i];
    g.ld[i] = h_ld[i];
  }
  RDM_CUDA(rdm_launch_pdl(gather_rows_kernel, dim3(cdiv(count, 8)), dim3(256), 0, stream, g, sel, count));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

__global__ void __launch_bounds__(256) l2_normalize_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int C) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= N) return;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = x[(size_t)r * C + c];
    q = fmaf(v, v, q);
  }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(q)), 1e-12f);
  for (int c = lane; c < C; c += 32) y[(size_t)r * C + c] = x[(size_t)r * C + c] * inv;
}
int rdm_l2_normalize(const float* x, float* y, int N, int C, cudaStream_t stream) {
  if (N == 0) return RDM_OK;
  RDM_CUDA(rdm_launch_pdl(l2_normalize_kernel, dim3(cdiv(N, 8)), dim3(256), 0, stream, x, y, N, C));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// out[r, :C] = x[r, :C]; out[r, C] = col[r]   (torch.cat([feats_c, n2p_logit], 1), experiments/model.py:166-167)
__global__ void append_column_kernel(const float* __restrict__ x, const float* __restrict__ col, int N, int C,
                                     float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)N * (C + 1)) return;
  const int r = (int)(e / (C + 1)), c = (int)(e - (long long)r * (C + 1));
  out[e] = c < C ? x[(size_t)r * C + c] : col[r];
}
int rdm_append_column(const float* x, const float* col, int N, int C, float* out, cudaStream_t stream) {
  if (N == 0) return RDM_OK;
  RDM_CUDA(rdm_launch_pdl(append_column_kernel, dim3(cdiv((long long)N * (C + 1), 256)), dim3(256), 0, stream, x, col, N, C, out));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}
// y[r] = clamp(sigmoid(x[r * ld]), 0, 1): score head on a strided column (the p2p logit = last decoder column)
__global__ void sigmoid_column_kernel(const float* __restrict__ x, int ld, int N, float* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  const float v = 1.f / (1.f + expf(-x[(size_t)r * ld]));
  y[r] = fminf(fmaxf(v, 0.f), 1.f);
}
int rdm_sigmoid_column(const float* x, int ld, int N, float* y, cudaStream_t stream) {
  if (N == 0) return RDM_OK;
  RDM_CUDA(rdm_launch_pdl(sigmoid_column_kernel, dim3(cdiv(N, 256)), dim3(256), 0, stream, x, ld, N, y));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}
