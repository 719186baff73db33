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
 *  - process model: ONE device and ONE host thread driving it per process (the reference's own arrangement: one process per
 *    GPU, geotransformer/engine/base_tester.py:70-76). The stateless operator entry points are re-entrant; the runners
 *    (rdm_build_pyramid*, rdm_encoder/decoder/backbone/match_forward) keep per-process scratch state (a pinned result
 *    buffer, a side stream with its events, the default pyramid job) bound to the device that was current at their first
 *    call, and must not be called concurrently from several threads or for several devices of one process.
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
/* number of CUDA kernels this library has launched in this process (bench.py reports the delta as gpu_launches) */
unsigned long long rdm_launch_count(void);
/* how many of those were tensor-core GEMMs (tcgen05 kind::tf32, 3-term split): lets tests assert the path taken */
unsigned long long rdm_tc_gemm_count(void);

/* ABI self-description (host only, no GPU): sizeof of every struct below in declaration order (rdm_prof_record,
 * rdm_tf_proj_job, rdm_tf_attn_job, rdm_unary_desc, rdm_block_desc, rdm_pyramid_desc, rdm_pyramid_cfg,
 * rdm_thdroformer_desc, rdm_backbone_desc, rdm_backbone_out, rdm_match_desc, rdm_match_io, rdm_match_result), then
 * offsetof(rdm_block_desc, sigma), (rdm_pyramid_desc, order), (rdm_match_desc, nms_limit), (rdm_match_io, transform),
 * (rdm_match_result, transform). Returns the number of entries. A binding checks its struct mirrors against it. */
int rdm_abi_layout(int64_t* h_out, int max_entries);

/* optional kernel timing: after rdm_prof_enable(1) the library brackets selected launches with CUDA events on the launch
 * stream; rdm_prof_read synchronises those events and returns the records. tag 1 = KPConv gather (row_positive prepass +
 * gather kernel, m/n/h/c = M, N, H, C_in), tag 2 = KPConv weight GEMM (m, n, -, c = M, K, 0, N). */
typedef struct {
  int tag;
  float ms;
  int m, n, h, c;
} rdm_prof_record;
void rdm_prof_enable(int on); /* bit mask of the tags to bracket: 1 = gather, 2 = weight GEMM, 3 = both, 0 = off */
int rdm_prof_read(rdm_prof_record* h_out, int max_records);

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
 * kernel_points [15,3] on the device and h_kernel_points = the same 45 floats in HOST memory (a module constant;
 * the kernel takes them by value in its parameter bank). rowpos_scratch: N bytes. */
int rdm_kpconv_gather(const float* s_feats, const float* q_points, const float* s_points, const void* neighbor_indices,
                      int index_bytes, const float* kernel_points, const float* h_kernel_points, float sigma, int M, int N, int H, int C_in,
                      const int* query_order /* optional permutation of 0..M-1: walk order, see rdm_pyramid_desc.order */,
                      float* out_weighted, unsigned char* rowpos_scratch, rdm_stream_t stream);

/* ---- maxpool (geotransformer/modules/kpconv/functional.py:54-67) */
int rdm_maxpool(const float* feats, const void* neighbor_indices, int index_bytes, int M, int N, int H, int C,
                float* out, rdm_stream_t stream);

/* ---- nearest_upsample + torch.cat([up, skip], 1) (functional.py:6-22, experiments/backbone.py:129-141).
 * index_stride = row stride (in elements) of the upsampling table; only column 0 is read. */
int rdm_upsample_concat(const float* feats, const void* upsample_indices, int index_bytes, int index_stride,
                        const float* skip, int M, int N, int C1, int C2, float* out, rdm_stream_t stream);

/* ---- nn.Linear / KPConv weight contraction: C[M,N] = act(A[M,K] * B (+ bias)). b_is_nk = 1: B is an nn.Linear
 * weight [N,K]; 0: B is [K,N]. act: 0 none, 1 LeakyReLU(0.1), 2 ReLU (AttentionOutput, transformer/output_layer.py:16-17).
 * workspace (optional) enables deterministic split-K for small M*N. */
size_t rdm_linear_workspace(int M, int N, int K);
int rdm_linear(const float* A, int lda, const float* B, int ldb, int b_is_nk, const float* bias, float* C, int ldc,
               int M, int N, int K, int act, void* workspace, size_t workspace_bytes, rdm_stream_t stream);

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

/* ---- RotaryPositionalEmbedding.forward (rdmnet/thdroformer/thdroformer.py:56-85): y = x*cos(t) + rot(x)*sin(t),
 * t = 2*pi*sigmoid(repeat_interleave(emb, 2)); x,y [N, C] (heads contiguous), emb [N, C/2]; ld* = row strides. */
int rdm_rope(const float* x, int ldx, const float* emb, int lde, float* y, int ldy, int N, int C, rdm_stream_t stream);

/* ---- dynamic_attention (k=None) / MultiHeadAttention core (thdroformer.py:20-40, vanilla_transformer.py:54-66):
 * O[n, h*D:(h+1)*D] = softmax_j(Q_n.K_j / sqrt(D)) V, heads laid out along the channel axis, D <= 64. */
int rdm_attention(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, float* O, int ldo, int Nq,
                  int Nk, int heads, int head_dim, rdm_stream_t stream);

/* ---- fused ThDRoFormer layers for d_model = 128, 4 heads x 32, FFN 256 (the RDMNet configuration):
 * TransformerLayer / RPETransformerLayer = MultiHeadAttention (+RoPE) + AttentionLayer + AttentionOutput
 * (rdmnet/thdroformer/thdroformer.py:88-202, transformer/vanilla_transformer.py:15-129, transformer/output_layer.py:6-21).
 * Job arrays live in HOST memory (they are copied into the kernel parameters); all pointers inside are device pointers.
 * rdm_tf_project: y[n,128] = rope?(x[n,:128] * W^T + bias) per job; wt is the TRANSPOSED weight ([in][out]); emb != NULL
 *   applies the 3-D rotary embedding with angles 2*pi*sigmoid(emb[n, c/2]) (emb row stride lde). Up to 6 jobs / launch.
 * rdm_tf_attend: out[nq,128] = LN2(x1 + FFN(x1)), x1 = LN1(x + Wo*softmax(q k^T / sqrt(32)) v + bo); q already holds
 *   the projected (and rotated) queries, k / v the projected keys / values; blob = rdm_tf_layer_blob_floats() floats:
 *   WqT WkT WvT WoT [128x128 each, k-major], W1T [128x256], W2T [256x128], bq bk bv bo [128], b1 [256], b2 [128],
 *   ln1.weight ln1.bias ln2.weight ln2.bias [128 each]. Up to 2 jobs / launch. */
typedef struct {
  const float* x;    /* [n, ldx] input rows (first 128 columns used) */
  const float* wt;   /* [128,128] transposed weight */
  const float* bias; /* [128] */
  const float* emb;  /* [n, lde] or NULL */
  float* y;          /* [n,128] row-major, or, if ldy_t != 0, channel-major [128, ldy_t] (key layout of rdm_tf_attend) */
  int n, ldx, lde, ldy_t;
} rdm_tf_proj_job;
typedef struct {
  const float* q;    /* [nq,128] */
  const float* k;    /* [128, ldk_t] channel-major, ldk_t >= nk, ldk_t % 4 == 0, 16-byte aligned */
  const float* v;    /* [nk,128] */
  const float* x;    /* [nq, ldx] layer input (residual) */
  const float* blob; /* layer blob */
  float* out;        /* [nq,128] */
  int nq, nk, ldx, ldk_t;
} rdm_tf_attn_job;
size_t rdm_tf_layer_blob_floats(void);
int rdm_tf_project(const rdm_tf_proj_job* h_jobs, int num_jobs, rdm_stream_t stream);
int rdm_tf_attend(const rdm_tf_attn_job* h_jobs, int num_jobs, rdm_stream_t stream);

/* ---- whole-module runners: one host call walks a module's kernels (the per-launch host cost of a Python loop is
 * what bounds a ~10 ms pair otherwise). Descriptor structs and pointer arrays named h_* live in HOST memory; every
 * pointer inside them is a device pointer (except h_kernel_points). `workspace` is device scratch; the matching
 * *_workspace function returns the bytes needed for the same arguments.
 *
 * rdm_encoder_forward = Encoder.forward (experiments/backbone.py:72-107): ConvBlock + 13 ResidualBlocks
 * (geotransformer/modules/kpconv/modules.py:104-225). blocks[0] is the ConvBlock (only kpconv_* and norm_conv_* set,
 * unary1/unary2/shortcut.w == NULL); stage s > 0 starts with a strided block reading subsampling[s-1].
 * out_feats[s] receives the last block of stage s ([n[s], c_out]). */
typedef struct {
  const float *w, *b;       /* nn.Linear weight [c_out, c_in], bias (w == NULL: identity / absent) */
  const float *gn_w, *gn_b; /* GroupNorm affine (NULL: no norm, LastUnaryBlock) */
  int c_in, c_out;
  int ldw;                  /* row stride of w in floats (0 = c_in); a stride padded to a multiple of 4 lets a layer with
                             * c_in % 4 != 0 (decoder4: 1281) take the TMA / tensor-core GEMM */
} rdm_unary_desc;
typedef struct {
  rdm_unary_desc unary1, unary2, shortcut;
  const float* kpconv_w;          /* [15, c_mid, c_mid'] */
  const float* kpconv_wt;         /* optional: the same weights as [c_mid', 15*c_mid] (nn.Linear layout, tensor-core GEMM) */
  const float* kpconv_b;          /* [c_mid'] or NULL */
  const float* kernel_points;     /* [15,3] device */
  const float* h_kernel_points;   /* the same 45 floats, host */
  const float *norm_conv_w, *norm_conv_b;
  int c_in, c_mid_in, c_mid_out, c_out, strided, stage;
  float sigma;
} rdm_block_desc;
typedef struct {
  const float* points[8];       /* per stage [n,3] */
  const void* neighbors[8];     /* [n[s], nb_width[s]] */
  const void* subsampling[8];   /* [n[s+1], sub_width[s]] */
  const void* upsampling[8];    /* [n[s], up_width[s]] (indices into stage s+1) */
  int n[8], nb_width[8], sub_width[8], up_width[8];
  int num_stages, index_bytes;
  const int* order[8];          /* optional per stage: a permutation of 0..n[s]-1, the order in which the KPConv gather
                                 * walks the queries of stage s (spatially coherent => neighbour rows hit in L1); results
                                 * do not depend on it. NULL = index order. */
} rdm_pyramid_desc;
/* rdm_build_pyramid = precompute_data_stack_mode (geotransformer/utils/data.py:13-77) in ONE host call: the
 * (num_stages-1) grid subsamplings chained on the device, ONE internal stream synchronisation to learn the
 * data-dependent stage sizes (the only entry point that synchronises), then every radius search at its exact size.
 * Tables are int32, fixed width = the stage's neighbour limit (a row narrower than the limit is padded: same KPConv /
 * max-pool results as the reference's max_count-wide table). skip_up0: upsampling[0] is not built (the reference
 * builds it and never reads it, experiments/backbone.py:144). up_nearest_only: upsampling tables hold only column 0,
 * the nearest coarse point - all nearest_upsample reads (geotransformer/modules/kpconv/functional.py:6-22).
 * out_buf (device, rdm_build_pyramid_bytes) receives stage lengths, points, tables and per-stage cell orders;
 * h_desc gets the pointers; h_lengths [num_stages*batch] the per-cloud stage sizes; h_d_lengths[s] the device int64
 * [batch] length vectors. points/lengths are stage 0 and are referenced, not copied. */
typedef struct {
  int num_stages, batch;
  float first_voxel;   /* voxel of the first subsampling (2 * init_voxel_size) */
  float first_radius;  /* search radius of stage 0 */
  int limits[8];
  int skip_up0, up_nearest_only;
} rdm_pyramid_cfg;
size_t rdm_build_pyramid_bytes(int64_t n0, const rdm_pyramid_cfg* h_cfg);
size_t rdm_build_pyramid_workspace(int64_t n0, const rdm_pyramid_cfg* h_cfg);
int rdm_build_pyramid(const float* points, const int64_t* lengths, int64_t n0, const rdm_pyramid_cfg* h_cfg, void* out_buf,
                      size_t out_bytes, void* workspace, size_t workspace_bytes, rdm_pyramid_desc* h_desc,
                      int64_t* h_lengths, const int64_t** h_d_lengths, rdm_stream_t stream);
/* two-phase form for pipelining pairs: _begin queues the subsampling chain + the size readback and returns at once;
 * _finish waits for that readback (host), then queues every radius search. A job handle carries the state (one
 * pyramid in flight per handle; out_buf / workspace must stay alive until _finish's work has run). */
void* rdm_pyramid_job_create(void);
void rdm_pyramid_job_destroy(void* job);
int rdm_build_pyramid_begin(void* job, const float* points, const int64_t* lengths, int64_t n0, const rdm_pyramid_cfg* h_cfg,
                            void* out_buf, size_t out_bytes, void* workspace, size_t workspace_bytes, rdm_stream_t stream);
int rdm_build_pyramid_finish(void* job, rdm_pyramid_desc* h_desc, int64_t* h_lengths, const int64_t** h_d_lengths,
                             rdm_stream_t stream);
size_t rdm_encoder_workspace(const rdm_block_desc* h_blocks, int num_blocks, const rdm_pyramid_desc* h_pyr, int groups);
int rdm_encoder_forward(const rdm_block_desc* h_blocks, int num_blocks, const rdm_pyramid_desc* h_pyr, int groups,
                        const float* in_feats, float* const* h_out_feats, void* workspace, size_t workspace_bytes,
                        rdm_stream_t stream);
/* rdm_decoder_forward = Decoder.forward (experiments/backbone.py:118-151): for i = 0..num-1 (coarse to fine),
 * x = unary_i(cat(nearest_upsample(x, upsampling[stage_i]), skip_i)); h_dec[i] = decoder4, decoder3, decoder2 (no norm).
 * coarse [n[top], c_coarse]; h_skips[i] = encoder output of stage top-1-i; out = last result [n[top-num], c_out]. */
size_t rdm_decoder_workspace(const rdm_unary_desc* h_dec, int num, const rdm_pyramid_desc* h_pyr, int top_stage, int groups);
int rdm_decoder_forward(const rdm_unary_desc* h_dec, int num, const rdm_pyramid_desc* h_pyr, int top_stage, int groups,
                        const float* coarse, int c_coarse, const float* const* h_skips, float* out, int ld_out,
                        void* workspace, size_t workspace_bytes, rdm_stream_t stream);
/* rdm_thdroformer_forward = ThDRoFormer.forward (rdmnet/thdroformer/thdroformer.py:304-347) for hidden 128 / 4 heads:
 * embedding Linear(3,64), in_proj, 2*L fused layers (h_layer_blobs[i], h_is_self[i]), out_proj. */
typedef struct {
  const float *emb_w, *emb_b;   /* [64,3], [64] */
  const float *in_w, *in_b;     /* [128, c_in], [128] */
  const float *out_w, *out_b;   /* [c_out, 128], [c_out] */
  const float* layer_blobs[32];
  int is_self[32];
  int num_layers, c_in, c_out;
} rdm_thdroformer_desc;
size_t rdm_thdroformer_workspace(int n_ref, int n_src, int c_out);
int rdm_thdroformer_forward(const rdm_thdroformer_desc* h_desc, const float* ref_points, int n_ref, const float* src_points,
                            int n_src, const float* ref_feats, int ld_ref, const float* src_feats, int ld_src,
                            float* out_ref, float* out_src, void* workspace, size_t workspace_bytes, rdm_stream_t stream);

/* ---- NMS.forward greedy loop (rdmnet/vote/vote.py:33-40) over a radius-search table [N,H]. out_mask [N] (0/1).
 * Optional: out_selected [N] receives the selected indices in ascending order, out_counts[0] / [1] how many of them
 * are < split / >= split (split = number of ref nodes: the two boolean-mask selections of model.py:233-236). */
int rdm_nms(const void* neighbor_indices, int index_bytes, int N, int H, int split, unsigned char* out_mask,
            int64_t* out_selected, int* out_counts, rdm_stream_t stream);

/* ---- point_to_node_partition (geotransformer/modules/ops/pointcloud_partition.py:60-107). */
size_t rdm_point_to_node_workspace(int num_points, int num_nodes);
int rdm_point_to_node(const float* points, int num_points, const float* nodes, int num_nodes, int point_limit,
                      int* out_point_to_node, unsigned char* out_node_masks, int64_t* out_knn_indices,
                      unsigned char* out_knn_masks, void* workspace, size_t workspace_bytes, rdm_stream_t stream);

/* ---- SuperPointMatching.forward (geotransformer/modules/geotransformer/superpoint_matching.py:14-83).
 * xy_scores [M,N] holds ref_feats @ src_feats^T on entry (rdm_linear) and is overwritten. sums_scratch: M+N floats.
 * Results sorted by descending score; out_count = min(num_correspondences, #valid pairs). */
int rdm_coarse_matching(float* xy_scores, int M, int N, const unsigned char* ref_masks, const unsigned char* src_masks,
                        int num_correspondences, int dual_normalization, int64_t* out_ref_indices,
                        int64_t* out_src_indices, float* out_scores, int* out_count, float* sums_scratch,
                        rdm_stream_t stream);

/* ---- patch gather + einsum('bnd,bmd->bnm') * scale (experiments/model.py:323-343); point_limit must be 128.
 * ld_feats = row stride of both feature tables in floats (multiple of 4; the decoder output keeps its 257th column). */
int rdm_patch_scores(const float* ref_feats, int Nr, const float* src_feats, int Ns, int C, int ld_feats,
                     const int64_t* ref_knn_indices, const int64_t* src_knn_indices, const int64_t* ref_corr_indices,
                     const int64_t* src_corr_indices, int num_patches, int point_limit, float scale, float* out_scores,
                     rdm_stream_t stream);

/* ---- LearnableLogOptimalTransport.forward (geotransformer/modules/sinkhorn/learnable_sinkhorn.py:13-66).
 * scores [P,R,C]; row_masks/col_masks are [P,R]/[P,C] or, when *_gather != NULL, node-level tables indexed through
 * the gather arrays (= knn_masks[corr_indices]); out [P,R+1,C+1]. */
int rdm_sinkhorn(const float* scores, int num_patches, int R, int C, const unsigned char* row_masks,
                 const unsigned char* col_masks, const int64_t* row_mask_gather, const int64_t* col_mask_gather,
                 const float* alpha, int num_iterations, float inf, float* out, rdm_stream_t stream);

/* ---- weighted_procrustes (geotransformer/modules/registration/procrustes.py:6-73): [B,n,3] x2 + [B,n] -> [B,4,4];
 * the SVD runs in-kernel (one warp per problem) instead of the reference's CPU round trip (:53). */
int rdm_weighted_procrustes(const float* src_points, const float* ref_points, const float* weights, int batch, int n,
                            float eps, float* out_transforms, rdm_stream_t stream);

/* ---- LocalGlobalRegistration.forward (geotransformer/modules/geotransformer/local_global_registration.py:204-243)
 * for k=1, mutual=False, use_dustbin=True, correspondence_limit=None. matching_scores [P,K+1,K+1] (log domain).
 * Outputs have capacity P*2*K rows; out_meta[0] = #correspondences, [1] = #local hypotheses, [2] = max chunk. */
size_t rdm_lgr_workspace(int num_patches, int K);
int rdm_lgr(const float* matching_scores, int num_patches, int K, const float* ref_points_f, const float* src_points_f,
            const int64_t* ref_knn_indices, const int64_t* src_knn_indices, const unsigned char* ref_knn_masks,
            const unsigned char* src_knn_masks, const int64_t* ref_corr_indices, const int64_t* src_corr_indices,
            float acceptance_radius, int correspondence_threshold, int num_refinement_steps, float* out_ref_corr_points,
            float* out_src_corr_points, float* out_corr_scores, int* out_corr_bij, float* out_transform, int* out_meta,
            void* workspace, size_t workspace_bytes, rdm_stream_t stream);

/* ---- rdm_backbone_forward: Encoder + first ThDRoFormer + n2p score head + Decoder + p2p score head
 * (experiments/model_infer.py:147-174) in ONE asynchronous host call; the encoder stage outputs live in the workspace.
 * feats_c [n[top], c]: transformer output, ref rows first; feats_f [n[top - num_dec], ld_feats_f]: decoder output (its
 * last real column is the p2p logit); n2p_scores [n[top]], p2p_scores [n[top - num_dec]]: clamp(sigmoid(logit), 0, 1). */
typedef struct {
  const rdm_block_desc* h_blocks;
  int num_blocks, groups;
  const rdm_thdroformer_desc* h_transformer1;
  const float *n2p_w, *n2p_b; /* proj_n2p_score [1, c] */
  const rdm_unary_desc* h_dec;
  int num_dec;
} rdm_backbone_desc;
typedef struct {
  float* feats_c;
  float* n2p_scores;
  float* feats_f;
  int ld_feats_f;
  float* p2p_scores;
} rdm_backbone_out;
/* Optional: a cudaEvent_t that the NEXT rdm_backbone_forward calls record on their stream right after the encoder (NULL = off).
 * Lets a caller that keeps several pairs in flight order other streams behind the KPConv gathers (rdmnet_b200.model.PairPipeline). */
int rdm_backbone_set_encoder_event(void* cuda_event);
size_t rdm_backbone_workspace(const rdm_backbone_desc* h_desc, const rdm_pyramid_desc* h_pyr, int nc_ref);
int rdm_backbone_forward(const rdm_backbone_desc* h_desc, const rdm_pyramid_desc* h_pyr, int nc_ref, const float* in_feats,
                         const rdm_backbone_out* h_out, void* workspace, size_t workspace_bytes, rdm_stream_t stream);

/* ---- rdm_match_forward: everything RDMNet.forward does after the decoder (experiments/model_infer.py:180-354) in ONE
 * host call - Vote_layer + n2n score head, NMS (radius search + greedy rule), node selection, second ThDRoFormer,
 * L2 normalisation, both point_to_node partitions, SuperPointMatching, patch scores, Sinkhorn and
 * LocalGlobalRegistration. Two internal stream synchronisations (the NMS survivor counts decide the shapes of the
 * second half; the final counts / pose are read back through pinned memory). All pointers in the descriptor are
 * device pointers except h_*; outputs are caller-allocated at the capacities given below. */
typedef struct {
  /* Vote_layer (rdmnet/vote/vote.py:43-117): two (Linear, LayerNorm, ReLU) stages, ctr_reg, out_proj LayerNorm */
  const float *v_w0, *v_b0, *v_g0, *v_e0; /* [h0, c], [h0] x3 */
  const float *v_w1, *v_b1, *v_g1, *v_e1; /* [h1, h0], [h1] x3 */
  const float *v_wr, *v_br;               /* [3 + c, h1] */
  const float *v_go, *v_eo;               /* [c] */
  float max_offset[3];
  int c, h0, h1;
  const float *n2n_w, *n2n_b;             /* proj_n2n_score [1, c] */
  const rdm_thdroformer_desc* h_transformer2;
  const float* ot_alpha;
  float nms_radius, acceptance_radius, sinkhorn_inf;
  int nms_limit, point_limit, num_correspondences, dual_normalization, sinkhorn_iterations, correspondence_threshold,
      refinement_steps;
} rdm_match_desc;
typedef struct {
  /* inputs */
  const float* points_c;        /* [nc, 3] stage-5 points, ref rows first */
  const int64_t* lengths_c;     /* [2] device */
  int nc, nc_ref;
  const float* feats_c;         /* [nc, c] first-transformer output */
  const float* n2p_scores;      /* [nc] */
  const float* points_f;        /* [nf, 3] */
  int nf, nf_ref;
  const float* feats_f;         /* [nf, c] with row stride ld_feats_f */
  int ld_feats_f;
  /* outputs, capacity in brackets */
  float* shifted_points;        /* [nc, 3] */
  float* vote_feats;            /* [nc, c] */
  float* n2n_scores;            /* [nc] */
  unsigned char* nms_mask;      /* [nc] */
  int64_t* selected;            /* [nc] */
  float* sel_points;            /* [nc, 3]   selected nodes, ref first */
  float* sel_feats_norm;        /* [nc, c]   L2-normalised second-transformer output */
  float* sel_n2p;               /* [nc] */
  float* sel_n2n;               /* [nc] */
  unsigned char* node_masks;    /* [nc] */
  int64_t* knn_indices;         /* [nc, point_limit] */
  unsigned char* knn_masks;     /* [nc, point_limit] */
  int64_t* corr_ref;            /* [num_correspondences] */
  int64_t* corr_src;            /* [num_correspondences] */
  float* corr_node_scores;      /* [num_correspondences] */
  float* matching_scores;       /* [num_correspondences, point_limit + 1, point_limit + 1] */
  float* ref_corr_points;       /* [num_correspondences * 2 * point_limit, 3] */
  float* src_corr_points;       /* same */
  float* corr_scores;           /* [num_correspondences * 2 * point_limit] */
  int* corr_bij;                /* [num_correspondences * 2 * point_limit, 3] */
  float* transform;             /* [16] */
} rdm_match_io;
typedef struct {
  int n_ref_sel, n_src_sel;     /* NMS survivors */
  int num_patches;              /* coarse correspondences actually produced (<= num_correspondences) */
  int num_corr;                 /* fine correspondences */
  float transform[16];
} rdm_match_result;
size_t rdm_match_workspace(const rdm_match_desc* h_desc, int nc, int nc_ref, int nf, int nf_ref);
int rdm_match_forward(const rdm_match_desc* h_desc, const rdm_match_io* h_io, rdm_match_result* h_result, void* workspace,
                      size_t workspace_bytes, rdm_stream_t stream);

/* The same tail as a job in three calls, for callers that keep several pairs in flight (rdmnet_b200.model.PairPipeline with
 * RDM_PIPE_OVERLAP=1): rdm_match_begin queues phase 1 (vote, NMS) and the read-back of the survivor counts and never blocks;
 * rdm_match_continue waits for those counts only, then queues everything else and the read-back of the result; rdm_match_finish waits
 * for the result (and redoes the patch stage at the exact count when fewer coarse pairs exist than requested). The job keeps copies
 * of desc / io; workspace and io buffers must stay alive until rdm_match_finish. One job per pair in flight. */
void* rdm_match_job_create(void);
void rdm_match_job_destroy(void* job);
/* Optional: a cudaEvent_t that the patch stage (patch scores, Sinkhorn, pose) of the NEXT rdm_match_continue / rdm_match_forward calls
 * waits for on its stream (NULL = off): lets the small-grid front of the tail overlap another pair's encoder while the large-grid
 * patch stage stays behind it. */
int rdm_match_set_patch_wait_event(void* cuda_event);
int rdm_match_job_reset(void* job); /* waits for the job's stream and drops what it has in flight (abandoned pipelines) */
int rdm_match_begin(void* job, const rdm_match_desc* h_desc, const rdm_match_io* h_io, void* workspace, size_t workspace_bytes,
                    rdm_stream_t stream);
int rdm_match_continue(void* job, rdm_match_result* h_result);
int rdm_match_finish(void* job, rdm_match_result* h_result);

/* ---- index_select(data, index, dim=0) (geotransformer/modules/ops/index_select.py:4-30) on rows of `row_words` 4-byte
 * words: out[i, :] = data[index[i], :], i < count. *err_flag (device int, zeroed by the caller, may be NULL) is set when
 * an index falls outside [0, rows): the host shim raises, as torch.index_select does. */
int rdm_index_select(const void* data, int64_t rows, int row_words, const void* index, int index_bytes, int64_t count,
                     void* out, int* err_flag, rdm_stream_t stream);

/* ---- apply_transform (geotransformer/modules/ops/transformation.py:7-60): out = points @ R^T + t for [batch, n, 3]
 * points and [batch, 4, 4] transforms (or one [4,4] for all batches: shared_transform = 1); normals rotate only. */
int rdm_apply_transform(const float* points, const float* transforms, int batch, int64_t n_per_batch, int shared_transform,
                        float* out, const float* normals, float* normals_out, rdm_stream_t stream);

/* ---- neighbour-count histogram of calibrate_neighbors_stack_mode (geotransformer/utils/data.py:207-211):
 * hist_accum[c] += #{q : counts[q] == c} for c < hist_n (counts = rdm_radius_search's out_counts). */
int rdm_neighbor_histogram(const int* counts, int n, int hist_n, int* hist_accum, rdm_stream_t stream);

/* ---- ground-truth superpoint correspondences, train / val path (geotransformer/modules/registration/matching.py).
 * rdm_node_correspondences = get_node_correspondences (:252-366, sphere_filter = 1) and get_node_overlap (:368-436,
 * sphere_filter = 0): src nodes / patch points are transformed by `transform` [4,4] (may be NULL) on the fly; masks are 0/1
 * bytes or NULL (= all valid). out_overlaps [M,N] (0 where the pair is filtered out); out_intersect [M,N] (may be NULL) is
 * the sphere-test matrix (return_mask = True); out_corr_indices [M*N,2] / out_corr_overlaps [M*N] / out_count receive
 * the row-major compaction of the pairs with overlap > 0 (all three may be NULL). K <= 256. */
size_t rdm_node_correspondences_workspace(int M, int N);
int rdm_node_correspondences(const float* ref_nodes, const float* src_nodes, const float* ref_knn_points,
                             const float* src_knn_points, const float* transform, float pos_radius, int M, int N, int K,
                             const unsigned char* ref_masks, const unsigned char* src_masks,
                             const unsigned char* ref_knn_masks, const unsigned char* src_knn_masks, int sphere_filter,
                             float* out_overlaps, unsigned char* out_intersect, int64_t* out_corr_indices,
                             float* out_corr_overlaps, int* out_count, void* workspace, size_t workspace_bytes,
                             rdm_stream_t stream);
/* torch.nonzero of a [rows, cols] matrix in row-major order: entries > 0 of `mat` (float) or != 0 of `bmat` (bytes),
 * exactly one of the two given; out_indices [rows*cols, 2], out_values (may be NULL), *out_count. */
int rdm_compact_nonzero(const float* mat, const unsigned char* bmat, int rows, int cols, int64_t* out_indices,
                        float* out_values, int* out_count, rdm_stream_t stream);
/* get_node_correspondences_disance (matching.py:441-503): mutual-nearest-node mask [M,N] under `transform`. */
int rdm_node_distance_mask(const float* ref_nodes, const float* src_nodes, const float* transform, float pos_radius, int M,
                           int N, const unsigned char* ref_masks, const unsigned char* src_masks, unsigned char* out_mask,
                           rdm_stream_t stream);

/* ---- registration_with_ransac_from_correspondences (geotransformer/utils/open3d.py:173-203; open3d's CPU RANSAC, called
 * per pair by experiments/infer.py:76-82 with 50 000 iterations) on the GPU: src/ref [C,3] are matched row by row.
 * Deterministic for a given seed. out_transform [4,4]; out_meta[0] = inliers of the winner, [1] = its iteration. */
size_t rdm_ransac_workspace(int num_correspondences);
int rdm_ransac_correspondences(const float* src_points, const float* ref_points, int num_correspondences,
                               float distance_threshold, int ransac_n, int num_iterations, unsigned long long seed,
                               float* out_transform, int* out_meta, void* workspace, size_t workspace_bytes,
                               rdm_stream_t stream);

/* ---- constant-weight pre-split for the tcgen05 3-term-split GEMM: out_split [2][rows][ld] = tf32 hi / lo of an nn.Linear
 * weight [rows, ld]. rdm_presplit_register(weight, split) makes the module runners (rdm_encoder/decoder/backbone/
 * thdroformer/match_forward) feed `split` to the GEMM through two TMA streams instead of splitting the weight tile in
 * every CTA and k-block; split = NULL unregisters, rdm_presplit_clear() forgets all. The caller owns the buffers and must
 * re-split after changing a weight (rdmnet_b200.modules does, keyed on the parameters' (pointer, version) fingerprint).
 * The plain operator rdm_linear never consults the registry. */
int rdm_presplit_weight(const float* weight, int rows, int ld, float* out_split, rdm_stream_t stream);
int rdm_presplit_register(const float* weight, const float* split);
void rdm_presplit_clear(void);

/* ---- ingest: voxel-barycentre downsample of a raw scan on the GPU, what the reference does offline with open3d
 * (preporcess/downsample_pcd_kitti.py:20-36, voxel_down_sample(0.3)): voxel index = floor((p - (min_bound - voxel/2)) / voxel),
 * output = mean of the points (and of the 4th column, the intensity, when stride = 4) of every occupied voxel, emitted in
 * the order of each voxel's first point in the input (open3d's own order is its hash map's, i.e. unspecified).
 * points [n, stride] with stride 3 or 4; out [n, stride] capacity; *out_count = number of voxels. Deterministic. */
size_t rdm_voxel_downsample_workspace(int n);
int rdm_voxel_downsample(const float* points, int stride, int n, float voxel, float* out, int* out_count, void* workspace,
                         size_t workspace_bytes, rdm_stream_t stream);

/* ---- training path: backward of the KPConv backbone operators (the reference differentiates its ~15-launch ATen graphs
 * per operator with autograd: geotransformer/modules/kpconv/kpconv.py:79-122, modules.py:33-225, functional.py:6-67;
 * experiments/trainval.py:43-50 -> loss.backward(), engine/epoch_based_trainer.py:104). Gradients are fp32. */
/* d s_feats [N, C_in] of rdm_kpconv_gather given d out_weighted [M, 15*C_in] (the neighbour count is a constant of the
 * backward pass, as in the reference: it comes from a comparison). d_s_feats is zeroed here; rowpos_scratch: N bytes. */
int rdm_kpconv_gather_bwd(const float* d_weighted, const float* s_feats, const float* q_points, const float* s_points,
                          const void* neighbor_indices, int index_bytes, const float* h_kernel_points, float sigma, int M, int N,
                          int H, int C_in, unsigned char* rowpos_scratch, float* d_s_feats, rdm_stream_t stream);
/* y [cols, rows] = x^T for x [rows, cols] with row stride ldx (dW = dY^T X and dX = dY W go through rdm_linear). */
int rdm_transpose(const float* x, int rows, int cols, int ldx, float* y, rdm_stream_t stream);
/* out[c] += sum_r x[r, c] (bias gradient); the caller zeroes out for a plain sum. */
int rdm_colsum(const float* x, int rows, int cols, int ldx, float* out_zeroed_or_accum, rdm_stream_t stream);
/* GroupNorm (+ LeakyReLU act = 1) backward (kpconv/modules.py:33-50, 78-83): dx [N,C], dgamma [C], dbeta [C]; dz_scratch [N,C]
 * receives dy * act'(.) = the gradient of the residual input; scratch: 2*groups and 2*C doubles. */
int rdm_groupnorm_bwd(const float* x, const float* y, const float* dy, const float* gamma, int N, int C, int groups, float eps,
                      int act, float slope, double* stats_scratch, double* dgamma_dbeta_scratch, float* dz_scratch, float* dx,
                      float* dgamma, float* dbeta, rdm_stream_t stream);
/* LayerNorm (+ residual, + ReLU act = 2) backward: dx [N,C] (= d residual); dgamma / dbeta are ACCUMULATED into. */
int rdm_layernorm_bwd(const float* x, const float* residual, const float* y, const float* dy, const float* gamma, int N, int C,
                      float eps, int act, float* dx, float* dgamma_accum, float* dbeta_accum, rdm_stream_t stream);
int rdm_maxpool_bwd(const float* feats, const void* neighbor_indices, int index_bytes, const float* d_out, int M, int N, int H,
                    int C, float* d_feats_zeroed, rdm_stream_t stream);
int rdm_upsample_concat_bwd(const float* d_out, const void* upsample_indices, int index_bytes, int index_stride, int M, int N,
                            int C1, int C2, float* d_feats_zeroed, float* d_skip, rdm_stream_t stream);
/* Linear backward in one call (x [M,K], dy [M,N], W [N,K] when w_is_nk else [K,N]): dx [M,K] = dy W, dW (W's layout) = dy^T x or
 * x^T dy, db [N] += column sums of dy (db must be zeroed by the caller); any of dx / dw / db may be NULL. The three products run on
 * the tensor-core GEMM with the contraction-slow operands transposed into the workspace. torch autograd of F.linear does the same
 * three products (the reference: loss.backward(), geotransformer/engine/epoch_based_trainer.py:104). */
size_t rdm_linear_bwd_workspace(int M, int N, int K);
int rdm_linear_bwd(const float* x, const float* w, int w_is_nk, const float* dy, int M, int N, int K, float* dx, float* dw,
                   float* db_zeroed, void* workspace, size_t workspace_bytes, rdm_stream_t stream);
/* workspace that lets rdm_linear run a [K,N]-layout B on the tensor cores (transposed copy + split-K partials) */
size_t rdm_linear_kn_workspace(int M, int N, int K);
/* out_accum[index[i], :] += src[i, :] (backward of rdm_index_select); rows of row_floats fp32. */
int rdm_scatter_add_rows(const float* src, const void* index, int index_bytes, int64_t count, int row_floats, int64_t rows,
                         float* out_accum, rdm_stream_t stream);
/* ---- precision of the dense contractions (BASELINE config 3). 0 (default): fp32-accurate 3-term tf32 split, the mode all parity
 * tests run in. 1: one kind::tf32 product per k-step (10-bit mantissa, like fp16), geometry / normalisation / Sinkhorn / pose stay
 * fp32. Process-wide; RDM_PRECISION=tf32 sets the initial value. */
int rdm_set_precision(int mode);
int rdm_get_precision(void);

/* ---- matcher-side backward kernels (csrc/train.cu). The reference: autograd over rdmnet/thdroformer/thdroformer.py:20-85 and
 * geotransformer/modules/sinkhorn/learnable_sinkhorn.py:13-66.
 * rdm_rope_bwd: dx [N,C] (dense), demb [N,C/2] (dense, may be NULL) of y = rope(x, emb).
 * rdm_attention_bwd: gradients of O = softmax(Q K^T / sqrt(d)) V per head; Q/K/V/O/dO row-strided (strides multiples of 4 floats,
 *   16-byte aligned), dQ/dK/dV with row stride ldd; head_dim 16 or 32; deterministic (no atomics).
 * rdm_sinkhorn_bwd: d scores [P,R,C] and per-patch partial sums of d alpha [P] from d out [P,R+1,C+1]; patch-level masks [P,R]/[P,C]
 *   (1 = live). Re-runs the forward iterations (exact expf/logf), keeps every iterate in the workspace, then walks them backwards.
 *   Gradient arriving on masked entries is ignored (they are the constant -inf in the forward). */
int rdm_rope_bwd(const float* x, int ldx, const float* emb, int lde, const float* dy, int ldy, int N, int C, float* dx, float* demb,
                 rdm_stream_t stream);
size_t rdm_attention_bwd_workspace(int Nq, int heads);
int rdm_attention_bwd(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const float* O, int ldo,
                      const float* dO, int ldg, int Nq, int Nk, int heads, int head_dim, void* workspace, size_t workspace_bytes,
                      float* dQ, float* dK, float* dV, int ldd, rdm_stream_t stream);
size_t rdm_sinkhorn_bwd_workspace(int num_patches, int R, int C, int num_iterations);
int rdm_sinkhorn_bwd(const float* scores, int num_patches, int R, int C, const unsigned char* row_masks,
                     const unsigned char* col_masks, const float* alpha, int num_iterations, float inf, const float* d_out,
                     void* workspace, size_t workspace_bytes, float* d_scores, float* d_alpha_partial, rdm_stream_t stream);
/* act: 1 LeakyReLU(slope), 2 ReLU, 3 clamp(sigmoid(x), 0, 1); y = the forward OUTPUT. */
int rdm_activation_bwd(const float* y, const float* dy, int64_t n, int act, float slope, float* dx, rdm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
