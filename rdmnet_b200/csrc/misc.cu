// Small operators of the drop-in boundary that the unmodified experiments/{model_infer,model,loss}.py call between
// modules: index_select (geotransformer/modules/ops/index_select.py:4-30), apply_transform
// (geotransformer/modules/ops/transformation.py:7-60), and the neighbour-count histogram of
// calibrate_neighbors_stack_mode (geotransformer/utils/data.py:195-220). All HBM-bound, coalesced, no staging needed.
#include "common.cuh"
#include "../../include/rdm_sm100.h"

// out[i, :] = data[index[i], :]; rows of `c` 4-byte words (16-byte vector path when c % 4 == 0 and both bases aligned)
template <typename IdxT, typename VecT>
__global__ void __launch_bounds__(256) index_select_rows_kernel(const VecT* __restrict__ data, const IdxT* __restrict__ index,
                                                                long long count, int cv, long long rows, VecT* __restrict__ out,
                                                                int* __restrict__ err) {
  pdl_trigger();
  pdl_wait();
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= count * cv) return;
  const long long i = e / cv;
  const int c = (int)(e - i * cv);
  const long long j = (long long)index[i];
  if (j < 0 || j >= rows) {  // torch.index_select raises on an out-of-range index: flag it, the host shim raises
    if (err) atomicExch(err, 1);
    return;
  }
  out[e] = data[j * cv + c];
}

extern "C" int rdm_index_select(const void* data, int64_t rows, int row_words, const void* index, int index_bytes, int64_t count,
                                void* out, int* err_flag, cudaStream_t stream) {
  RDM_CHECK_ARG(index_bytes == 4 || index_bytes == 8, "rdm_index_select: index_bytes must be 4 or 8");
  RDM_CHECK_ARG(rows >= 0 && row_words >= 1 && count >= 0, "rdm_index_select: bad sizes");
  if (count == 0) return RDM_OK;
  const bool v4 = row_words % 4 == 0 && ((uintptr_t)data % 16 == 0) && ((uintptr_t)out % 16 == 0);
  const int cv = v4 ? row_words / 4 : row_words;
  const long long total = count * cv;
  const dim3 grid(cdiv(total, 256)), block(256);
#define GO(IdxT, VecT)                                                                                                            \
  RDM_CUDA(rdm_launch_pdl(index_select_rows_kernel<IdxT, VecT>, grid, block, 0, stream, (const VecT*)data, (const IdxT*)index, \
                          (long long)count, cv, (long long)rows, (VecT*)out, err_flag))
  if (index_bytes == 8) {
    if (v4) GO(int64_t, float4);
    else GO(int64_t, float);
  } else {
    if (v4) GO(int, float4);
    else GO(int, float);
  }
#undef GO
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// y = R p + t per point (transformation.py:42-48: points @ R^T + t); one transform per batch element, or one for all
// (t_stride = 0). Normals (optional) are rotated only (:49-50). Separate multiplies and adds in the matmul's order.
__global__ void __launch_bounds__(256) apply_transform_kernel(const float* __restrict__ pts, const float* __restrict__ T, int t_stride,
                                                              long long n_per_batch, long long total, float* __restrict__ out,
                                                              const float* __restrict__ nrm, float* __restrict__ nrm_out) {
  pdl_trigger();
  pdl_wait();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float* t = T + (size_t)(i / n_per_batch) * t_stride;
  const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const float a = t[4 * r], b = t[4 * r + 1], c = t[4 * r + 2];
    out[3 * i + r] = fmaf(z, c, fmaf(y, b, x * a)) + t[4 * r + 3];
    if (nrm != nullptr) {
      const float nx = nrm[3 * i], ny = nrm[3 * i + 1], nz = nrm[3 * i + 2];
      nrm_out[3 * i + r] = fmaf(nz, c, fmaf(ny, b, nx * a));
    }
  }
}

extern "C" int rdm_apply_transform(const float* points, const float* transforms, int batch, int64_t n_per_batch, int shared_transform,
                                   float* out, const float* normals, float* normals_out, cudaStream_t stream) {
  RDM_CHECK_ARG(batch >= 1 && n_per_batch >= 0, "rdm_apply_transform: bad sizes");
  RDM_CHECK_ARG((normals == nullptr) == (normals_out == nullptr), "rdm_apply_transform: normals and normals_out go together");
  const long long total = (long long)batch * n_per_batch;
  if (total == 0) return RDM_OK;
  RDM_CUDA(rdm_launch_pdl(apply_transform_kernel, dim3(cdiv(total, 256)), dim3(256), 0, stream, points, transforms,
                          shared_transform ? 0 : 16, (long long)n_per_batch, total, out, normals, normals_out));
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// hist[min-clipped c] += 1 for every query whose in-radius count c is < hist_n (data.py:209-211: np.bincount(...)[:hist_n]);
// per-CTA shared histogram, then one global atomic per non-empty bin.
__global__ void __launch_bounds__(256) neighbor_hist_kernel(const int* __restrict__ counts, int n, int hist_n, int* __restrict__ hist) {
  extern __shared__ int s_hist[];
  for (int i = threadIdx.x; i < hist_n; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = counts[i];
    if (c >= 0 && c < hist_n) atomicAdd(&s_hist[c], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < hist_n; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

extern "C" int rdm_neighbor_histogram(const int* counts, int n, int hist_n, int* hist_accum, cudaStream_t stream) {
  RDM_CHECK_ARG(n >= 0 && hist_n >= 1 && hist_n <= 12 * 1024, "rdm_neighbor_histogram: hist_n must be in [1, 12288]");
  if (n == 0) return RDM_OK;
  const int grid = min(cdiv(n, 256), 148 * 4);
  neighbor_hist_kernel<<<grid, 256, hist_n * sizeof(int), stream>>>(counts, n, hist_n, hist_accum);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}
