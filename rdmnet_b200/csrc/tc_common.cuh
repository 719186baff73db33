// Shared pieces of the tcgen05 GEMM kernels (gemm_tc.cu, gemm_tc_atmem.cu): mbarrier / TMA / UMMA wrappers, the tf32
// rounding helper, the GroupNorm-statistics epilogue and the host-side tensor-map encoder. Everything lives in an
// anonymous namespace: each translation unit gets its own copy.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;  // floats = 128 bytes
constexpr int TC_THREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  const long long t0 = clock64();
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && clock64() - t0 > 4000000000LL) __trap();  // ~2 s: a broken pipeline aborts instead of hanging
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  // K-major, SWIZZLE_128B: start>>4 | LBO=1 (unused) | SBO = 1024 B (8 rows x 128 B) | version 1 (sm_100) | layout 2
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// round-to-nearest (ties away) to tf32 precision with two integer ops on the bit pattern: same result as
// cvt.rna.tf32.f32 for finite values, but on the full-rate ALU pipe instead of the conversion unit
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

// r[j] for a runtime j without spilling the array to local memory (only the ragged-N tail path uses it)
__device__ __forceinline__ uint32_t sel32(const uint32_t (&r)[32], int j) {
  uint32_t v = r[0];
#pragma unroll
  for (int i = 1; i < 32; i++) v = (j == i) ? r[i] : v;
  return v;
}

// GroupNorm statistics fused into a GEMM epilogue. The warp holds a 32 (rows = lanes) x 32 (columns = v[0..31]) slab
// of the output; a 5-step butterfly (31 shuffles per quantity) leaves lane L with the sum over the 32 rows of column L,
// a segmented shuffle reduction folds the cpg columns of a group, and one lane per group adds {sum, sumsq} to the
// double accumulators stats[2g], stats[2g+1]. cpg is a power of two; groups never straddle a 32-column slab unless
// cpg > 32, where the whole slab belongs to one group.
__device__ __forceinline__ void gn_slab_stats(const float (&v)[32], bool row_valid, int col0, int cpg, int lane,
                                              double* __restrict__ stats) {
  float s[32], q[32];
#pragma unroll
  for (int j = 0; j < 32; j++) {
    s[j] = row_valid ? v[j] : 0.f;
    q[j] = s[j] * s[j];
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; j++) {
      const float ks = up ? s[j + off] : s[j], ss = up ? s[j] : s[j + off];
      const float kq = up ? q[j + off] : q[j], sq = up ? q[j] : q[j + off];
      s[j] = ks + __shfl_xor_sync(FULL_MASK, ss, off);
      q[j] = kq + __shfl_xor_sync(FULL_MASK, sq, off);
    }
  }
  float cs = s[0], cq = q[0];  // column col0 + lane
  const int span = cpg < 32 ? cpg : 32;
  for (int o = 1; o < span; o <<= 1) {
    cs += __shfl_xor_sync(FULL_MASK, cs, o);
    cq += __shfl_xor_sync(FULL_MASK, cq, o);
  }
  if ((lane & (span - 1)) == 0) {
    const int g = (col0 + lane) / cpg;
    atomicAdd(&stats[2 * g], (double)cs);
    atomicAdd(&stats[2 * g + 1], (double)cq);
  }
}

// ---- host side: tensor maps through the driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_encode_state = 0;  // 0 unknown, 1 ok, -1 unavailable

bool load_encoder() {
  if (g_encode_state == 0) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn != nullptr &&
        qres == cudaDriverEntryPointSuccess) {
      g_encode = (EncodeTiledFn)fn;
      g_encode_state = 1;
    } else {
      g_encode_state = -1;
    }
  }
  return g_encode_state == 1;
}

// row-major [rows, cols] fp32 with leading dimension ld (floats); box = 32 columns x box_rows, 128-byte swizzle,
// out-of-bounds elements read as zero (K and M/N tails)
bool make_map(CUtensorMap* map, const float* base, int rows, int cols, int ld, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace
