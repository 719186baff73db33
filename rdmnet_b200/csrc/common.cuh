// Shared helpers for librdm_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define RDM_OK 0
#define RDM_ERR_ARG 1
#define RDM_ERR_CUDA 2
#define RDM_ERR_WORKSPACE 3

void rdm_set_error(const char* fmt, ...);

#define RDM_CHECK_ARG(cond, ...)  \
  do {                            \
    if (!(cond)) {                \
      rdm_set_error(__VA_ARGS__); \
      return RDM_ERR_ARG;         \
    }                             \
  } while (0)

#define RDM_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      rdm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return RDM_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

// every kernel launch of the library goes through this macro: it also feeds rdm_launch_count()
extern unsigned long long g_rdm_launches;
#define RDM_LAUNCH_CHECK()                                      \
  do {                                                          \
    __atomic_fetch_add(&g_rdm_launches, 1ull, __ATOMIC_RELAXED); \
    RDM_CUDA(cudaGetLastError());                               \
  } while (0)

// ---- programmatic dependent launch (PDL). Kernels of the long launch chains (encoder / decoder / transformer) start
// with pdl_trigger() - "my dependents may be scheduled as soon as every CTA of mine has started" - and call pdl_wait()
// before their first access to memory a predecessor may still be using; constant data (weights) may be touched before
// it. Launched through rdm_launch_pdl (programmatic stream serialization), the next kernel's CTAs are resident and past
// their prologue when the current grid drains, which hides the ~2 us launch gap per kernel boundary. Kernels launched
// the ordinary way are unaffected (the two instructions are no-ops there). RDM_PDL=0 disables the attribute.
bool rdm_pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t rdm_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = rdm_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// optional kernel timing (api.cu); id < 0 = profiling off
int rdm_prof_begin(int tag, int m, int n, int h, int c, cudaStream_t stream);
void rdm_prof_end(int id, cudaStream_t stream);
#define RDM_PROF_KPCONV_GATHER 1
#define RDM_PROF_KPCONV_GEMM 2

// internal (not part of the C ABI): Linear with GroupNorm statistics fused into the epilogue (dense.cu / gemm_tc.cu),
// the two halves of GroupNorm (dense.cu), the KPConv gather with a ready-made row-positivity table (kpconv.cu)
int rdm_linear_gn(const float* A, int lda, const float* B, int ldb, int b_is_nk, const float* bias, float* C, int ldc, int M,
                  int N, int K, int act, void* workspace, size_t workspace_bytes, double* gn_stats, int gn_cpg,
                  int* stats_fused, cudaStream_t stream);
int rdm_linear_gn_ps(const float* A, int lda, const float* B, int ldb, int b_is_nk, const float* bias, float* C, int ldc, int M,
                     int N, int K, int act, void* workspace, size_t workspace_bytes, double* gn_stats, int gn_cpg,
                     int* stats_fused, const float* B_split, cudaStream_t stream);
const float* rdm_presplit_lookup(const float* weight);  // dense.cu: NULL unless registered (rdm_presplit_register)
int rdm_groupnorm_stats(const float* x, int N, int C, int groups, double* stats_zeroed, cudaStream_t stream);
int rdm_groupnorm_apply(const float* x, const double* stats, const float* gamma, const float* beta, const float* residual,
                        float* y, int N, int C, int groups, float eps, int act, float slope, unsigned char* rowpos_out,
                        cudaStream_t stream);

int rdm_kpconv_gather_impl(const float* s_feats, const float* q_points, const float* s_points, const void* neighbor_indices,
                           int index_bytes, const float* kernel_points, const float* h_kernel_points, float sigma, int M, int N,
                           int H, int C_in, const int* query_order, float* out_weighted, unsigned char* rowpos_scratch,
                           int rowpos_ready, cudaStream_t stream);

int rdm_mark_reference_width(int* table, long long rows, int H, int n_support, const int* d_max_count, cudaStream_t stream);
int rdm_row_positive(const float* f, int n, int c, unsigned char* flag, cudaStream_t stream);  // kpconv.cu: (sum_c f[n,c] > 0)
int rdm_order_by_load(const int* order_in, const int* neighbors, int n, int H, int n_support, int* order_out, void* scratch,
                      size_t scratch_bytes, cudaStream_t stream);
int rdm_upsample_concat_ld(const float* feats, const void* upsample_indices, int index_bytes, int index_stride,
                           const float* skip, int M, int N, int C1, int C2, float* out, int ld_out, cudaStream_t stream);

int rdm_vote_finish(const float* off, int ld_off, const float* xyz, const float* feats, int ld_f, const float* gamma,
                    const float* beta, const float* h_limit3, float eps, int N, int C, float* xyz_out, float* feat_out,
                    cudaStream_t stream);
int rdm_gather_rows(const float* const* h_src, float* const* h_dst, const int* h_c, const int* h_ld, int num_jobs,
                    const int64_t* sel, int count, cudaStream_t stream);
int rdm_l2_normalize(const float* x, float* y, int N, int C, cudaStream_t stream);

int rdm_append_column(const float* x, const float* col, int N, int C, float* out, cudaStream_t stream);
int rdm_sigmoid_column(const float* x, int ld, int N, float* y, cudaStream_t stream);

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct Workspace {
  char* base;
  size_t size, off;
  bool ok;
  __host__ Workspace(void* p, size_t n) : base((char*)p), size(n), off(0), ok(true) {}
  template <typename T>
  __host__ T* get(size_t count) {
    off = align_up(off, 256);
    T* r = (T*)(base + off);
    off += count * sizeof(T);
    if (off > size) ok = false;
    return r;
  }
};

#ifdef __CUDACC__
#define FULL_MASK 0xffffffffu

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}

// Block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32).
// `smem` must hold 33 ints. Returns exclusive prefix; *total gets the block sum.
__device__ __forceinline__ int block_exclusive_scan(int v, int* smem, int* total) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(FULL_MASK, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protect smem reuse across calls
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nw ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(FULL_MASK, winc, o);
      if (lane >= o) winc += t;
    }
    smem[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  *total = smem[32];
  return smem[warp] + inc - v;
}

// In-warp bitonic sort (ascending) of n (power of two, >= 32... or any pow2) 64-bit keys held in shared memory.
__device__ __forceinline__ void warp_bitonic_sort_u64(unsigned long long* a, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n >> 1); t += 32) {
        // element pair (i, i^j) with i having bit j clear
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int p = i | j;
        unsigned long long x = a[i], y = a[p];
        bool up = ((i & k) == 0);
        if ((x > y) == up) {
          a[i] = y;
          a[p] = x;
        }
      }
      __syncwarp();
    }
  }
}
#endif
