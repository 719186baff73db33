/*
 * librdm_sm100.so - C ABI of the B200-native (sm_100a) RDMNet dense-matching hot path.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name starts with h_;
 *  - every function is asynchronous on `stream` (cudaStream_t passed as void* so that the header needs no CUDA
 *    include), never synchronises, never allocates: the caller provides outputs and scratch/workspace buffers;
 *  - return value: 0 = ok, 1 = bad argument, 2 = CUDA error, 3 = workspace too small; rdm_last_error() gives the
 *    message of the last failure on the calling thread. Nothing throws. The Python host shim (rdmnet_b200/_lib.py)
 *    turns non-zero codes into RuntimeError, which is the error convention of the reference extension
 *    (TORCH_CHECK -> RuntimeError, geotransformer/extensions/common/torch_helper.h:6-35);
 *  - feature / point tensors are fp32 row-major; index tensors are int64 (index_bytes = 8, the reference dtype)
 *    or int32 (index_bytes = 4); neighbour tables are padded with the number of support rows, exactly as
 *    radius_neighbors_cpu.cpp:85 does.
 *
 * Each entry point names the reference interface (file:line under /root/reference) it replaces.
 */
#ifndef RDM_SM100_H
#define RDM_SM100_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef __CUDACC__
typedef cudaStream_t rdm_stream_t;
#else
typedef void* rdm_stream_t;
#endif

const char* rdm_last_error(void);
int rdm_version(void);

/* ---- rdmnet.ext.grid_subsampling (geotransformer/extensions/cpu/grid_subsampling/grid_subsampling.cpp:5-62,
 *      core grid_subsampling_cpu.cpp:3-75; Python wrapper geotransformer/modules/ops/grid_subsample.py:7-22).
 * points [n_total,3], lengths [batch] (device int64). out_points has capacity n_total_cap rows; the stacked result
 * occupies the first sum(out_lengths) rows, bit-exact with the reference including row order. n_total_cap must be
 * >= sum(lengths). */
size_t rdm_grid_subsample_workspace(int64_t n_total_cap, int batch);
int rdm_grid_subsample(const float* points, const int64_t* lengths, int batch, int64_t n_total_cap, float voxel_size,
                       float* out_points, int64_t* out_lengths, void* workspace, size_t workspace_bytes,
                       rdm_stream_t stream);
/* host-only: checks the embedded libstdc++ bucket-growth table against this process' std::unordered_map */
int rdm_selfcheck_bucket_table(int64_t max_elements);

/* ---- rdmnet.ext.radius_neighbors + the [:, :limit] truncation
 *      (geotransformer/extensions/cpu/radius_neighbors/radius_neighbors.cpp:5-67, core radius_neighbors_cpu.cpp:3-91,
 *       geotransformer/modules/ops/radius_search.py:7-27).
 * out_indices [nq_cap, limit]: the `limit` nearest supports with fp32 d2 < r2, ascending (d2, index), global indices,
 * padded with ns_total_pad. out_max_count (device int) = the reference's max_count (row width before truncation);
 * out_counts (optional, [nq_cap]) = per-query neighbour count. limit == 0 -> count-only pass. */
size_t rdm_radius_search_workspace(int64_t ns_cap, int batch);
int rdm_radius_search(const float* q_points, const float* s_points, const int64_t* q_lengths, const int64_t* s_lengths,
                      int batch, int64_t nq_cap, int64_t ns_cap, int64_t ns_total_pad, float radius, int limit,
                      void* out_indices, int index_bytes, int* out_counts, int* out_max_count, void* workspace,
                      size_t workspace_bytes, rdm_stream_t stream);

/* ---- KPConv.forward, gather half (geotransformer/modules/kpconv/kpconv.py:79-116):
 * out_weighted [M, 15*C_in] = (1/neighbor_num) * sum_h influence[m,h,k] * s_feats[idx[m,h], c]; follow with
 * rdm_linear(out_weighted, W.view(15*C_in, C_out), b_is_nk = 0, bias) to finish :105-120.
 * rowpos_scratch: N bytes. */
int rdm_kpconv_gather(const float* s_feats, const float* q_points, const float* s_points, const void* neighbor_indices,
                      int index_bytes, const float* kernel_points, float sigma, int M, int N, int H, int C_in,
                      float* out_weighted, unsigned char* rowpos_scratch, rdm_stream_t stream);

/* ---- maxpool (geotransformer/modules/kpconv/functional.py:54-67) */
int rdm_maxpool(const float* feats, const void* neighbor_indices, int index_bytes, int M, int N, int H, int C,
                float* out, rdm_stream_t stream);

/* ---- nearest_upsample + torch.cat([up, skip], 1) (functional.py:6-22, experiments/backbone.py:129-141).
 * index_stride = row stride (in elements) of the upsampling table; only column 0 is read. */
int rdm_upsample_concat(const float* feats, const void* upsample_indices, int index_bytes, int index_stride,
                        const float* skip, int M, int N, int C1, int C2, float* out, rdm_stream_t stream);

/* ---- nn.Linear / KPConv weight contraction: C[M,N] = A[M,K] * B (+ bias). b_is_nk = 1: B is an nn.Linear weight
 * [N,K]; 0: B is [K,N]. workspace (optional) enables deterministic split-K for small M*N. */
size_t rdm_linear_workspace(int M, int N, int K);
int rdm_linear(const float* A, int lda, const float* B, int ldb, int b_is_nk, const float* bias, float* C, int ldc,
               int M, int N, int K, void* workspace, size_t workspace_bytes, rdm_stream_t stream);

/* ---- GroupNorm over stacked features (geotransformer/modules/kpconv/modules.py:33-50) fused with the optional
 * residual add and LeakyReLU of UnaryBlock/ConvBlock/ResidualBlock (:78-83, :143-147, :222-224).
 * act: 0 none, 1 LeakyReLU(slope). stats_scratch: 2*groups doubles. */
int rdm_groupnorm(const float* x, const float* gamma, const float* beta, const float* residual, float* y, int N, int C,
                  int groups, float eps, int act, float slope, double* stats_scratch, rdm_stream_t stream);

/* ---- nn.LayerNorm(x + residual) (+ ReLU when act == 2): transformer/output_layer.py:20, vanilla_transformer.py:100,
 * rdmnet/vote/vote.py:57-60,115 */
int rdm_layernorm(const float* x, const float* residual, const float* gamma, const float* beta, float* y, int N, int C,
                  float eps, int act, rdm_stream_t stream);

/* ---- elementwise: act 1 LeakyReLU(slope), 2 ReLU, 3 clamp(sigmoid(x), 0, 1) (experiments/model.py:162-174) */
int rdm_activation(const float* x, float* y, int64_t n, int act, float slope, rdm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
