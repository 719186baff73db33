// DEFAULT tcgen05 GEMM (RDM_GEMM_ATMEM=0 / rdm_debug_gemm_variant(0) select the all-shared-memory kernel of gemm_tc.cu):
// the 3-term-split tf32 GEMM with the A operand in TENSOR MEMORY. Validated in round 2 on the whole GPU suite and the bench.
//
// Why: gemm_tc.cu is shared-memory-bandwidth bound (scripts/gemm_timeline.py: ~1100 cycles per 128 x 64 x 32 k-block for
// ~100 cycles of MMA). Per k-block it moves 24 KB (TMA) + 24 KB (converter reads) + 48 KB (converter writes of hi / lo)
// + 72 KB (operand reads of the three products) through a 128 B/clk shared memory. Here the converter threads write
// A_hi / A_lo straight into TMEM with tcgen05.st (thread = tile row = TMEM lane) and the MMAs take A from TMEM
// ("ts" form: tcgen05.mma [d], [a_tmem], b_desc), so A costs 16 KB of shared-memory reads per k-block instead of 112 KB:
// 88 KB per k-block in total. Everything else (TMA pipeline, B split in place, epilogue with bias / activation /
// GroupNorm statistics, split-K) is the gemm_tc.cu design.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/rdm_sm100.h"
#include "tc_common.cuh"

extern unsigned long long g_tc_launches_ext;

namespace {

// D[tmem] (+)= A[tmem] * B[smem desc]: A rows = TMEM lanes, K along columns
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 consecutive columns of this thread's TMEM lane <- r[0..31]
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}

template <int BN, int STAGES>
struct TcaSmem {
  // every buffer is a multiple of 1024 B: the 128-byte swizzle pattern repeats every 8 rows
  float a_raw[STAGES][TC_BM * TC_BK];  // raw fp32 A tile (TMA, SWIZZLE_128B); its hi / lo split lives in TMEM
  float b_hi[STAGES][BN * TC_BK];
  float b_lo[STAGES][BN * TC_BK];
  uint64_t raw_full[STAGES], conv_full[STAGES], empty[STAGES], accum_full;
  uint32_t tmem_base;
};

// PS = true: B arrives PRE-SPLIT (constant weights split into tf32 hi / lo once per weights epoch, rdm_presplit_weight):
// two TMA streams fill b_hi / b_lo directly and the converter warps touch A only - 48 KB less shared-memory traffic and
// half the converter instructions per k-block of a 128-wide tile.
template <int BN, int STAGES, bool PS>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tf32x3_atmem_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                   const __grid_constant__ CUtensorMap map_b,
                                                                   const __grid_constant__ CUtensorMap map_b_lo,
                                                                   const float* __restrict__ bias, float* __restrict__ C,
                                                                   int ldc, int M, int N, int K, int act, int kb_per_split,
                                                                   double* __restrict__ gn_stats, int gn_cpg,
                                                                   long long* __restrict__ dbg) {
  extern __shared__ unsigned char smem_raw[];
  using Smem = TcaSmem<BN, STAGES>;
  Smem& sm = *reinterpret_cast<Smem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  // split-K: blockIdx.z owns K blocks [kb0, kb0 + nk) and writes a raw partial tile to C + z*M*N (ldc == N there)
#define TC_STAMP(i)                                                        \
  do {                                                                     \
    if (dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) dbg[i] = clock64(); \
  } while (0)
  if (threadIdx.x == 0) TC_STAMP(0);
  const int nk_total = (K + TC_BK - 1) / TC_BK;
  const int kb0 = blockIdx.z * kb_per_split;
  const int nk = min(kb_per_split, nk_total - kb0);
  if (gridDim.z > 1) C += (size_t)blockIdx.z * M * N;

  pdl_trigger();
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");  // hide the descriptor fetch behind the set-up
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    if (PS) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&sm.raw_full[s], 1);
      mbar_init(&sm.conv_full[s], 4);  // one arrival per converter warp
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.accum_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // TMEM columns: [0, BN) accumulators, then per stage 32 columns of A_hi and 32 of A_lo (row m of the tile = lane m)
  constexpr int A_COL0 = BN;
  static_assert(BN + 64 * STAGES <= 512, "TMEM budget");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;
  if (threadIdx.x == 0) TC_STAMP(1);
  pdl_wait();  // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      constexpr uint32_t bytes = (TC_BM + (PS ? 2 : 1) * BN) * TC_BK * sizeof(float);
      for (int kb = 0; kb < nk; kb++) {
        const int s = kb % STAGES;
        mbar_wait(&sm.empty[s], ((kb / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&sm.raw_full[s], bytes);
        tma_load_2d(sm.a_raw[s], &map_a, &sm.raw_full[s], (kb0 + kb) * TC_BK, m0);
        tma_load_2d(sm.b_hi[s], &map_b, &sm.raw_full[s], (kb0 + kb) * TC_BK, n0);
        if (PS) tma_load_2d(sm.b_lo[s], &map_b_lo, &sm.raw_full[s], (kb0 + kb) * TC_BK, n0);
        if (kb == 0) TC_STAMP(2);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // kind::tf32, D = f32, A/B = tf32 K-major, N>>3 at bit 17, M>>4 at bit 24
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < nk; kb++) {
        const int s = kb % STAGES;
        mbar_wait(&sm.conv_full[s], (kb / STAGES) & 1);
        if (kb == 0) TC_STAMP(4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t dbh = umma_desc_sw128(smem_u32(sm.b_hi[s])), dbl = umma_desc_sw128(smem_u32(sm.b_lo[s]));
        const uint32_t ta_hi = tmem + (uint32_t)(A_COL0 + 64 * s), ta_lo = ta_hi + 32;
#pragma unroll
        for (int k = 0; k < TC_BK / 8; k++) {
          const uint64_t adv = (uint64_t)(k * 32 / 16);  // 8 tf32 = 32 bytes along K inside the swizzle atom (B operand)
          umma_tf32_ts(tmem, ta_lo + 8 * k, dbh + adv, idesc, (kb | k) != 0);
          umma_tf32_ts(tmem, ta_hi + 8 * k, dbl + adv, idesc, 1);
          umma_tf32_ts(tmem, ta_hi + 8 * k, dbh + adv, idesc, 1);
        }
        umma_commit(&sm.empty[s]);  // frees the stage when these MMAs have read it
      }
      umma_commit(&sm.accum_full);
    }
  } else {
    // ===== converters, then epilogue =====
    const int t = threadIdx.x - 64;  // 0..127
    for (int kb = 0; kb < nk; kb++) {
      const int s = kb % STAGES;
      mbar_wait(&sm.raw_full[s], (kb / STAGES) & 1);
      if (kb == 0 && t == 0) TC_STAMP(3);
      {
        // A: this thread owns row (q*32 + lane) of the tile = TMEM lane of the same number (a warp may only touch its
        // own lane quarter q = warp & 3). Read the row's 32 floats from the swizzled raw tile (16-byte chunk c of row r
        // sits at chunk c ^ (r & 7)), split, and store hi / lo as 32 columns each with tcgen05.st.
        const int q = warp & 3, r = q * 32 + lane;
        const float4* arow = reinterpret_cast<const float4*>(sm.a_raw[s]) + r * 8;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const float4 v = arow[c ^ (r & 7)];
          const float h0 = to_tf32(v.x), h1 = to_tf32(v.y), h2 = to_tf32(v.z), h3 = to_tf32(v.w);
          hi[4 * c] = __float_as_uint(h0); hi[4 * c + 1] = __float_as_uint(h1);
          hi[4 * c + 2] = __float_as_uint(h2); hi[4 * c + 3] = __float_as_uint(h3);
          lo[4 * c] = __float_as_uint(to_tf32(v.x - h0)); lo[4 * c + 1] = __float_as_uint(to_tf32(v.y - h1));
          lo[4 * c + 2] = __float_as_uint(to_tf32(v.z - h2)); lo[4 * c + 3] = __float_as_uint(to_tf32(v.w - h3));
        }
        const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(A_COL0 + 64 * s);
        tmem_st32(ta, hi);
        tmem_st32(ta + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      if (!PS) {
        float4* bh = reinterpret_cast<float4*>(sm.b_hi[s]);
        float4* bl = reinterpret_cast<float4*>(sm.b_lo[s]);
#pragma unroll
        for (int i = 0; i < BN * TC_BK / 4 / 128; i++) {
          const float4 v = bh[t + i * 128];
          const float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
          bh[t + i * 128] = h;
          bl[t + i * 128] = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes (B) -> visible to the UMMA reads
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // TMEM stores (A) ordered before the arrive
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.conv_full[s]);
    }
    if (t == 0) TC_STAMP(5);
    mbar_wait(&sm.accum_full, 0);
    if (t == 0) TC_STAMP(6);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = m0 + q * 32 + lane;
    float* crow = C + (size_t)row * ldc;
    const bool aligned = (ldc % 4 == 0) && (((uintptr_t)C & 15) == 0) && (((uintptr_t)bias & 15) == 0);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      const bool fast = aligned && (n0 + c0 + 32 <= N);
      uint32_t r[32];
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (fast) {  // whole 32-column slab inside N, 16-byte aligned rows: float4 bias, float4 stores, no per-element checks
        const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + c0);
        float vals[32];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                 __uint_as_float(r[4 * j + 3]));
          if (bias != nullptr) {
            const float4 b = __ldg(b4 + j);
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
          }
          if (act == 1) {
            v.x = v.x > 0.f ? v.x : 0.1f * v.x; v.y = v.y > 0.f ? v.y : 0.1f * v.y;
            v.z = v.z > 0.f ? v.z : 0.1f * v.z; v.w = v.w > 0.f ? v.w : 0.1f * v.w;
          } else if (act == 2) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
          }
          if (row < M) *reinterpret_cast<float4*>(crow + n0 + c0 + 4 * j) = v;
          vals[4 * j] = v.x; vals[4 * j + 1] = v.y; vals[4 * j + 2] = v.z; vals[4 * j + 3] = v.w;
        }
        if (gn_stats != nullptr) gn_slab_stats(vals, row < M, n0 + c0, gn_cpg, lane, gn_stats);
      } else if (row < M) {
#pragma unroll 1
        for (int j = 0; j < 32; j++) {
          const int col = n0 + c0 + j;
          if (col >= N) break;
          float v = __uint_as_float(sel32(r, j));
          if (bias != nullptr) v += __ldg(bias + col);
          if (act == 1) v = v > 0.f ? v : 0.1f * v;
          else if (act == 2) v = fmaxf(v, 0.f);
          crow[col] = v;
        }
      }
    }
  }
  if (threadIdx.x == 64) TC_STAMP(7);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) TC_STAMP(8);
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  }
}

template <int BN, int STAGES, bool PS>
int launch_tca(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mbl, const float* bias, float* C, int ldc, int M, int N,
              int K, int act, int splits, int kb_per_split, double* gn_stats, int gn_cpg, cudaStream_t stream,
              long long* dbg = nullptr) {
  const size_t smem = sizeof(TcaSmem<BN, STAGES>) + 1024;
  // per-device attribute: set on every launch (sub-microsecond), never cached per process
  RDM_CUDA(cudaFuncSetAttribute(gemm_tf32x3_atmem_kernel<BN, STAGES, PS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(N, BN), cdiv(M, TC_BM), splits);
  RDM_CUDA(rdm_launch_pdl(gemm_tf32x3_atmem_kernel<BN, STAGES, PS>, grid, dim3(TC_THREADS), smem, stream, ma, mb, mbl, bias, C, ldc, M, N,
                          K, act, kb_per_split, gn_stats, gn_cpg, dbg));
  RDM_LAUNCH_CHECK();
  __atomic_fetch_add(&g_tc_launches_ext, 1ull, __ATOMIC_RELAXED);
  return RDM_OK;
}
}  // namespace

// ---- constant-weight pre-split: hi = tf32(w), lo = tf32(w - hi), written as [2][rows][ld]
namespace {
__global__ void __launch_bounds__(256) presplit_kernel(const float* __restrict__ w, long long n, float* __restrict__ hi, float* __restrict__ lo) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i >= n) return;
  const float v = w[i], h = to_tf32(v);
  hi[i] = h;
  lo[i] = to_tf32(v - h);
}
}  // namespace

extern "C" int rdm_presplit_weight(const float* w, int rows, int ld, float* out_split, cudaStream_t stream) {
  RDM_CHECK_ARG(rows >= 1 && ld >= 1 && w != nullptr && out_split != nullptr, "rdm_presplit_weight: bad arguments");
  const long long n = (long long)rows * ld;
  presplit_kernel<<<cdiv(n, 256), 256, 0, stream>>>(w, n, out_split, out_split + n);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// Same contract as rdm_linear_tc (gemm_tc.cu). Returns -1 when the variant is off or the shape does not qualify.
int g_gemm_variant = -1;  // -1 unknown (read RDM_GEMM_ATMEM, default on), 0 off, 1 on
extern "C" void rdm_debug_gemm_variant(int v) { g_gemm_variant = v ? 1 : 0; }

// B_split (optional): the tf32 hi / lo split of B, [2][N][ldb] floats (rdm_presplit_weight): selects the PS kernels.
int rdm_linear_tc_atmem(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int M, int N, int K,
                        int act, void* workspace, size_t workspace_bytes, int* out_splits, double* gn_stats, int gn_cpg,
                        int* out_stats_fused, const float* B_split, cudaStream_t stream) {
  if (g_gemm_variant < 0) {
    const char* e = getenv("RDM_GEMM_ATMEM");
    g_gemm_variant = (e && e[0] == '0') ? 0 : 1;
  }
  if (g_gemm_variant != 1) return -1;
  *out_splits = 1;
  if (out_stats_fused) *out_stats_fused = 0;
  if (M < 1 || N < 8 || K < 8) return -1;
  if ((lda % 4) || (ldb % 4) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) return -1;
  if (!load_encoder()) return -1;
  const int nk_all = cdiv(K, TC_BK);
  const long long t128 = (long long)cdiv(M, TC_BM) * cdiv(N, 128);
  const bool can_split = workspace != nullptr && nk_all >= 16;
  const bool narrow = N <= 64 || (t128 < 100 && !(can_split && t128 * min(16, nk_all / 8) >= 100));
  const int BN = narrow ? 64 : 128;
  const long long tiles = (long long)cdiv(M, TC_BM) * cdiv(N, BN);
  const int nk = cdiv(K, TC_BK);
  int splits = 1;
  if (workspace != nullptr && tiles < 120 && nk >= 16) {
    splits = (int)min((long long)16, (296 + tiles - 1) / tiles);
    splits = min(splits, nk / 8);
    while (splits > 1 && (size_t)splits * M * N * sizeof(float) > workspace_bytes) splits--;
  }
  const int kps = cdiv(nk, splits);
  splits = cdiv(nk, kps);
  CUtensorMap ma, mb, mbl;
  if (!make_map(&ma, A, M, K, lda, TC_BM)) return -1;
  if (B_split != nullptr) {
    if (((uintptr_t)B_split & 15) || !make_map(&mb, B_split, N, K, ldb, BN) || !make_map(&mbl, B_split + (size_t)N * ldb, N, K, ldb, BN))
      return -1;
  } else {
    if (!make_map(&mb, B, N, K, ldb, BN)) return -1;
    mbl = mb;
  }
  *out_splits = splits;
  float* out = splits > 1 ? (float*)workspace : C;
  const int ldo = splits > 1 ? N : ldc;
  const float* b = splits > 1 ? nullptr : bias;
  const int a = splits > 1 ? 0 : act;
  double* st = nullptr;
  if (gn_stats != nullptr && splits == 1 && act == 0 && N % 32 == 0 && gn_cpg >= 1 && (gn_cpg & (gn_cpg - 1)) == 0 &&
      (gn_cpg <= 32 || gn_cpg % 32 == 0) && ldc % 4 == 0 && (((uintptr_t)C | (uintptr_t)bias) & 15) == 0) {
    st = gn_stats;
    if (out_stats_fused) *out_stats_fused = 1;
  }
  if (B_split != nullptr) {
    if (narrow) return launch_tca<64, 4, true>(ma, mb, mbl, b, out, ldo, M, N, K, a, splits, kps, st, gn_cpg, stream);
    return launch_tca<128, 3, true>(ma, mb, mbl, b, out, ldo, M, N, K, a, splits, kps, st, gn_cpg, stream);
  }
  if (narrow) return launch_tca<64, 4, false>(ma, mb, mbl, b, out, ldo, M, N, K, a, splits, kps, st, gn_cpg, stream);
  return launch_tca<128, 3, false>(ma, mb, mbl, b, out, ldo, M, N, K, a, splits, kps, st, gn_cpg, stream);
}
