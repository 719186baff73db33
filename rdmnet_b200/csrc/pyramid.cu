// Voxel pyramid construction on the GPU: grid_subsample and radius_search.
//
// Replaces (behind the same operator API) the reference's CPU extension `rdmnet.ext`:
//   grid_subsampling  -> geotransformer/extensions/cpu/grid_subsampling/grid_subsampling_cpu.cpp:3-75
//   radius_neighbors  -> geotransformer/extensions/cpu/radius_neighbors/radius_neighbors_cpu.cpp:3-91 (+ nanoflann)
//
// grid_subsample is bit-exact with the reference INCLUDING the output order, which in the reference is the
// iteration order of a libstdc++ std::unordered_map<size_t,...>. We reproduce that order without a hash map:
// within one "era" (fixed bucket count B) the final list order is  buckets by DESCENDING first-touch time, nodes
// inside a bucket by DESCENDING touch time, where the touch time of a node that existed before the rehash is its
// position in the previous era's list and the touch time of a node inserted later is its insertion rank. Each
// era is therefore a counting problem (min-reduce, suffix-sum, in-bucket rank), and the ~log2(M) eras are chained.
//
// All fp32 arithmetic that decides a voxel key, a barycentre or a neighbour test uses explicit round-to-nearest
// intrinsics so that nvcc cannot contract it into FMAs (the host reference build has no FMA).
#include "common.cuh"
#include "../../include/rdm_sm100.h"

// libstdc++ (g++ 13) unordered_map bucket-count sequence; rehash to BKT[k] fires when the element count
// reaches BKT[k-1]+1. Checked against the running libstdc++ by rdm_selfcheck_bucket_table().
static const long long H_BKT[] = {13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753, 42043, 85229, 172933,
                                  351061, 712697, 1447153, 2938679, 5967347, 12117689, 24607243, 49969847};
#define N_BKT 22
__constant__ long long c_bkt[N_BKT];
static bool g_bkt_uploaded = false;

static int upload_tables() {
  if (!g_bkt_uploaded) {
    RDM_CUDA(cudaMemcpyToSymbol(c_bkt, H_BKT, sizeof(H_BKT)));
    g_bkt_uploaded = true;
  }
  return RDM_OK;
}

// ------------------------------------------------------------------------------------------------ grid_subsample
struct GsWork {
  unsigned long long* hkeys;  // hash table: key+1, 0 = empty            [4*n_total + 64*batch]
  int* hfirst;                // first point index of the voxel           [same]
  int* hcnt;                  // points in the voxel                      [same]
  int* huid;                  // insertion rank of the voxel              [same]
  int* pslot;                 // hash slot of each point                  [n_total]
  unsigned long long* ukey;   // voxel key by insertion rank              [n_total]
  int* ucnt;                  // [n_total]
  int* uoff;                  // CSR offsets                              [n_total]
  int* ucur;                  // CSR fill cursors                         [n_total]
  int* ulist;                 // point indices grouped by voxel           [n_total]
  float* bary;                // barycentres by insertion rank            [3*n_total]
  int *e_t, *e_next, *e_bkt, *e_gt, *e_A;  // era state                   [n_total each]
  int *e_ft, *e_head;                      // per bucket                  [3*n_total + 32*batch each]
};

__device__ __forceinline__ unsigned int hash_u64(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  return (unsigned int)k;
}

// Chunked block-wide exclusive scan over arr[0..n) in place (int). Returns the total. All threads must call.
__device__ int block_scan_array(int* arr, int n, int* smem) {
  int nt = blockDim.x;
  int chunk = (n + nt - 1) / nt;
  int beg = min(n, (int)threadIdx.x * chunk), end = min(n, beg + chunk);
  int s = 0;
  for (int i = beg; i < end; i++) s += arr[i];
  int total;
  int pre = block_exclusive_scan(s, smem, &total);
  for (int i = beg; i < end; i++) {
    int v = arr[i];
    arr[i] = pre;
    pre += v;
  }
  __syncthreads();
  return total;
}

__global__ void __launch_bounds__(1024) gs_cloud_kernel(const float* __restrict__ points,
                                                        const int64_t* __restrict__ lengths, int batch, float voxel,
                                                        float inv_voxel, float* __restrict__ stage_out,
                                                        int64_t* __restrict__ out_lengths, GsWork w) {
  __shared__ int s_scan[33];
  __shared__ float s_red[6][32];
  __shared__ float s_org[3];
  __shared__ unsigned long long s_nxy[2];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  long long start = 0;
  for (int i = 0; i < b; i++) start += lengths[i];
  const int n = (int)lengths[b];
  if (n <= 0) {
    if (tid == 0) out_lengths[b] = 0;
    return;
  }
  const float* pts = points + 3 * start;
  // per-cloud slices of the workspace
  int T = 64;
  while (T < 2 * n) T <<= 1;
  const long long toff = 4 * start + 64LL * b, boff = 3 * start + 32LL * b;
  unsigned long long* hkeys = w.hkeys + toff;
  int *hfirst = w.hfirst + toff, *hcnt = w.hcnt + toff, *huid = w.huid + toff;
  int* pslot = w.pslot + start;
  unsigned long long* ukey = w.ukey + start;
  int *ucnt = w.ucnt + start, *uoff = w.uoff + start, *ucur = w.ucur + start, *ulist = w.ulist + start;
  float* bary = w.bary + 3 * start;
  int *e_t = w.e_t + start, *e_next = w.e_next + start, *e_bkt = w.e_bkt + start, *e_gt = w.e_gt + start,
      *e_A = w.e_A + start;
  int *e_ft = w.e_ft + boff, *e_head = w.e_head + boff;

  // P1: clear table, bounding box (cloud.cpp:5-39)
  for (int i = tid; i < T; i += nt) {
    hkeys[i] = 0ULL;
    hfirst[i] = 0x7fffffff;
    hcnt[i] = 0;
  }
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int i = tid; i < n; i += nt) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      float v = pts[3 * i + d];
      mn[d] = fminf(mn[d], v);
      mx[d] = fmaxf(mx[d], v);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(FULL_MASK, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(FULL_MASK, mx[d], o));
    }
    if ((tid & 31) == 0) {
      s_red[d][tid >> 5] = mn[d];
      s_red[3 + d][tid >> 5] = mx[d];
    }
  }
  __syncthreads();
  if (tid == 0) {
    int nw = nt >> 5;
    float o[3], m[3];
    for (int d = 0; d < 3; d++) {
      float a = s_red[d][0], c = s_red[3 + d][0];
      for (int i = 1; i < nw; i++) {
        a = fminf(a, s_red[d][i]);
        c = fmaxf(c, s_red[3 + d][i]);
      }
      // grid_subsampling_cpu.cpp:11  origin = floor(min * (float)(1./voxel)) * voxel
      o[d] = __fmul_rn(floorf(__fmul_rn(a, inv_voxel)), voxel);
      m[d] = c;
      s_org[d] = o[d];
    }
    // :13-20  sampleN = (size_t)(floor((max - origin) / voxel) + 1)
    s_nxy[0] = (unsigned long long)(long long)(floorf(__fdiv_rn(__fsub_rn(m[0], o[0]), voxel)) + 1.0f);
    s_nxy[1] = (unsigned long long)(long long)(floorf(__fdiv_rn(__fsub_rn(m[1], o[1]), voxel)) + 1.0f);
  }
  __syncthreads();
  const float ox = s_org[0], oy = s_org[1], oz = s_org[2];
  const unsigned long long nx = s_nxy[0], ny = s_nxy[1];

  // P2: voxel key per point (:32-35), insert into the hash table
  for (int i = tid; i < n; i += nt) {
    unsigned long long ix = (unsigned long long)(long long)floorf(__fdiv_rn(__fsub_rn(pts[3 * i + 0], ox), voxel));
    unsigned long long iy = (unsigned long long)(long long)floorf(__fdiv_rn(__fsub_rn(pts[3 * i + 1], oy), voxel));
    unsigned long long iz = (unsigned long long)(long long)floorf(__fdiv_rn(__fsub_rn(pts[3 * i + 2], oz), voxel));
    unsigned long long key = ix + nx * iy + nx * ny * iz;
    unsigned long long tag = key + 1ULL;  // 0 is "empty" (a key of 2^64-1 is impossible for in-range voxels)
    unsigned int h = hash_u64(key) & (T - 1);
    while (true) {
      unsigned long long prev = atomicCAS(&hkeys[h], 0ULL, tag);
      if (prev == 0ULL || prev == tag) break;
      h = (h + 1) & (T - 1);
    }
    pslot[i] = (int)h;
    atomicMin(&hfirst[h], i);
    atomicAdd(&hcnt[h], 1);
  }
  __syncthreads();

  // P3: insertion rank of every voxel = rank of its first point among all first points
  int M;
  {
    int chunk = (n + nt - 1) / nt;
    int beg = min(n, tid * chunk), end = min(n, beg + chunk);
    int s = 0;
    for (int i = beg; i < end; i++) s += (hfirst[pslot[i]] == i);
    int pre = block_exclusive_scan(s, s_scan, &M);
    for (int i = beg; i < end; i++) {
      int sl = pslot[i];
      if (hfirst[sl] == i) {
        huid[sl] = pre;
        ukey[pre] = hkeys[sl] - 1ULL;
        ucnt[pre] = hcnt[sl];
        uoff[pre] = hcnt[sl];
        ucur[pre] = 0;
        pre++;
      }
    }
  }
  __syncthreads();
  // P4: CSR of points per voxel
  block_scan_array(uoff, M, s_scan);
  for (int i = tid; i < n; i += nt) {
    int u = huid[pslot[i]];
    int p = atomicAdd(&ucur[u], 1);
    ulist[uoff[u] + p] = i;
  }
  __syncthreads();
  // P5: barycentre = (sequential fp32 sum in input order) * (float)(1.0/count)   (grid_subsampling_cpu.h:17-20, .cpp:46)
  for (int u = tid; u < M; u += nt) {
    int o = uoff[u], c = ucnt[u];
    for (int i = 1; i < c; i++) {  // insertion sort: ascending point index
      int v = ulist[o + i], j = i - 1;
      while (j >= 0 && ulist[o + j] > v) {
        ulist[o + j + 1] = ulist[o + j];
        j--;
      }
      ulist[o + j + 1] = v;
    }
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int i = 0; i < c; i++) {
      int p = ulist[o + i];
      sx = __fadd_rn(sx, pts[3 * p + 0]);
      sy = __fadd_rn(sy, pts[3 * p + 1]);
      sz = __fadd_rn(sz, pts[3 * p + 2]);
    }
    float s = (float)(1.0 / (double)c);
    bary[3 * u + 0] = __fmul_rn(sx, s);
    bary[3 * u + 1] = __fmul_rn(sy, s);
    bary[3 * u + 2] = __fmul_rn(sz, s);
  }
  // P6: emulate the unordered_map iteration order, era by era
  for (int k = 0; k < N_BKT; k++) {
    long long lo = k == 0 ? 0 : c_bkt[k - 1];
    if (lo >= M) break;
    const long long B = c_bkt[k];
    const int n_end = (int)min((long long)M, B);
    const int nB = (int)B;
    for (int x = (int)lo + tid; x < n_end; x += nt) e_t[x] = x;
    for (int i = tid; i < nB; i += nt) {
      e_ft[i] = 0x7fffffff;
      e_head[i] = -1;
    }
    __syncthreads();
    for (int x = tid; x < n_end; x += nt) {
      int bk = (int)(ukey[x] % (unsigned long long)B);
      e_bkt[x] = bk;
      atomicMin(&e_ft[bk], e_t[x]);
      e_next[x] = atomicExch(&e_head[bk], x);
    }
    __syncthreads();
    for (int x = tid; x < n_end; x += nt) {
      int bk = e_bkt[x], tx = e_t[x], cnt = 0, g = 0;
      for (int y = e_head[bk]; y >= 0; y = e_next[y]) {
        cnt++;
        g += (e_t[y] > tx);
      }
      e_gt[x] = g;
      e_A[tx] = (tx == e_ft[bk]) ? cnt : 0;
    }
    __syncthreads();
    int total = block_scan_array(e_A, n_end, s_scan);  // e_A[tau] = sum_{tau' < tau}
    // position = (#nodes in buckets first touched later) + (#nodes of my bucket touched later)
    for (int x = tid; x < n_end; x += nt) {
      int bk = e_bkt[x];
      int f = e_ft[bk];
      // nodes in buckets with first-touch > f  = total - inclusive_prefix(f) = total - (excl(f) + size(bk))
      int size_bk = (f + 1 < n_end ? e_A[f + 1] : total) - e_A[f];
      e_t[x] = total - (e_A[f] + size_bk) + e_gt[x];
    }
    __syncthreads();
  }
  for (int u = tid; u < M; u += nt) {
    long long dst = start + e_t[u];
    stage_out[3 * dst + 0] = bary[3 * u + 0];
    stage_out[3 * dst + 1] = bary[3 * u + 1];
    stage_out[3 * dst + 2] = bary[3 * u + 2];
  }
  if (tid == 0) out_lengths[b] = M;
}

// Packs the per-cloud results (each written at its cloud's input offset) into one stacked tensor.
__global__ void gs_pack_kernel(const float* __restrict__ stage, const int64_t* __restrict__ lengths,
                               const int64_t* __restrict__ out_lengths, int batch, float* __restrict__ out,
                               long long n_cap) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long src0 = 0, dst0 = 0;
  for (int b = 0; b < batch; b++) {
    long long m = out_lengths[b];
    if (i >= dst0 && i < dst0 + m) {
      long long s = src0 + (i - dst0);
      out[3 * i + 0] = stage[3 * s + 0];
      out[3 * i + 1] = stage[3 * s + 1];
      out[3 * i + 2] = stage[3 * s + 2];
      return;
    }
    src0 += lengths[b];
    dst0 += m;
  }
}

extern "C" size_t rdm_grid_subsample_workspace(int64_t n_total_cap, int batch) {
  size_t n = (size_t)n_total_cap, t = 4 * n + 64 * (size_t)batch, bk = 3 * n + 32 * (size_t)batch;
  size_t bytes = 0;
  bytes += align_up(t * 8, 256) + 3 * align_up(t * 4, 256);  // hash table
  bytes += align_up(n * 4, 256);                              // pslot
  bytes += align_up(n * 8, 256);                              // ukey
  bytes += 4 * align_up(n * 4, 256);                          // ucnt uoff ucur ulist
  bytes += align_up(3 * n * 4, 256);                          // bary
  bytes += 5 * align_up(n * 4, 256);                          // era per node
  bytes += 2 * align_up(bk * 4, 256);                         // era per bucket
  bytes += align_up(3 * n * 4, 256);                          // staging output
  return bytes + 4096;
}

extern "C" int rdm_grid_subsample(const float* points, const int64_t* lengths, int batch, int64_t n_total_cap,
                                  float voxel_size, float* out_points, int64_t* out_lengths, void* workspace,
                                  size_t workspace_bytes, cudaStream_t stream) {
  RDM_CHECK_ARG(batch >= 1 && n_total_cap >= 0 && voxel_size > 0.f, "rdm_grid_subsample: bad arguments");
  RDM_CHECK_ARG(n_total_cap < (1LL << 28), "rdm_grid_subsample: too many points");
  if (int e = upload_tables()) return e;
  if (n_total_cap == 0) {
    RDM_CUDA(cudaMemsetAsync(out_lengths, 0, sizeof(int64_t) * batch, stream));
    return RDM_OK;
  }
  size_t n = (size_t)n_total_cap, t = 4 * n + 64 * (size_t)batch, bk = 3 * n + 32 * (size_t)batch;
  Workspace ws(workspace, workspace_bytes);
  GsWork w;
  w.hkeys = ws.get<unsigned long long>(t);
  w.hfirst = ws.get<int>(t);
  w.hcnt = ws.get<int>(t);
  w.huid = ws.get<int>(t);
  w.pslot = ws.get<int>(n);
  w.ukey = ws.get<unsigned long long>(n);
  w.ucnt = ws.get<int>(n);
  w.uoff = ws.get<int>(n);
  w.ucur = ws.get<int>(n);
  w.ulist = ws.get<int>(n);
  w.bary = ws.get<float>(3 * n);
  w.e_t = ws.get<int>(n);
  w.e_next = ws.get<int>(n);
  w.e_bkt = ws.get<int>(n);
  w.e_gt = ws.get<int>(n);
  w.e_A = ws.get<int>(n);
  w.e_ft = ws.get<int>(bk);
  w.e_head = ws.get<int>(bk);
  float* stage = ws.get<float>(3 * n);
  if (!ws.ok) {
    rdm_set_error("rdm_grid_subsample: workspace too small (%zu < %zu)", workspace_bytes, ws.off);
    return RDM_ERR_WORKSPACE;
  }
  float inv_voxel = (float)(1.0 / (double)voxel_size);
  gs_cloud_kernel<<<batch, 1024, 0, stream>>>(points, lengths, batch, voxel_size, inv_voxel, stage, out_lengths, w);
  RDM_LAUNCH_CHECK();
  gs_pack_kernel<<<cdiv(n_total_cap, 256), 256, 0, stream>>>(stage, lengths, out_lengths, batch, out_points,
                                                            n_total_cap);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

// Host-side self check: the embedded bucket table equals the growth of this process' libstdc++ unordered_map.
#include <unordered_map>
extern "C" int rdm_selfcheck_bucket_table(int64_t max_elements) {
  std::unordered_map<size_t, int> m;
  size_t bc = m.bucket_count();
  int era = -1;
  for (int64_t i = 0; i < max_elements; i++) {
    m.emplace((size_t)i, 0);
    if (m.bucket_count() != bc) {
      bc = m.bucket_count();
      era++;
      long long expect_at = era == 0 ? 1 : H_BKT[era - 1] + 1;
      if (era >= N_BKT || (long long)bc != H_BKT[era] || i + 1 != expect_at) {
        rdm_set_error("bucket table mismatch at era %d: rehash at %lld to %zu", era, (long long)i + 1, bc);
        return RDM_ERR_ARG;
      }
    }
  }
  return RDM_OK;
}

// ------------------------------------------------------------------------------------------------ radius_search
#define RS_CELL_CAP (1 << 18)  // max grid cells per cloud
#define RS_DIM_CAP 4096

struct RsCloud {
  float ox, oy, oz, inv_cs;
  int nx, ny, nz, ncell;
  long long s_start, q_start;
  int ns, nq;
};

__global__ void __launch_bounds__(1024) rs_bounds_kernel(const float* __restrict__ s_pts,
                                                         const int64_t* __restrict__ q_len,
                                                         const int64_t* __restrict__ s_len, int batch, float radius,
                                                         RsCloud* __restrict__ clouds, int* __restrict__ max_count) {
  __shared__ float s_red[6][32];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  long long s0 = 0, q0 = 0;
  for (int i = 0; i < b; i++) {
    s0 += s_len[i];
    q0 += q_len[i];
  }
  const int n = (int)s_len[b];
  const float* pts = s_pts + 3 * s0;
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int i = tid; i < n; i += nt) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      float v = pts[3 * i + d];
      mn[d] = fminf(mn[d], v);
      mx[d] = fmaxf(mx[d], v);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(FULL_MASK, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(FULL_MASK, mx[d], o));
    }
    if ((tid & 31) == 0) {
      s_red[d][tid >> 5] = mn[d];
      s_red[3 + d][tid >> 5] = mx[d];
    }
  }
  __syncthreads();
  if (tid == 0) {
    if (b == 0) *max_count = 0;
    int nw = nt >> 5;
    float lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
      lo[d] = s_red[d][0];
      hi[d] = s_red[3 + d][0];
      for (int i = 1; i < nw; i++) {
        lo[d] = fminf(lo[d], s_red[d][i]);
        hi[d] = fmaxf(hi[d], s_red[3 + d][i]);
      }
    }
    RsCloud c;
    c.s_start = s0;
    c.q_start = q0;
    c.ns = n;
    c.nq = (int)q_len[b];
    if (n <= 0) {
      c.ox = c.oy = c.oz = 0.f;
      c.inv_cs = 1.f;
      c.nx = c.ny = c.nz = c.ncell = 1;
    } else {
      // cell edge slightly above the radius: a neighbour (fp32 d2 < r2) is then always within +-1 cell
      float cs = radius * 1.001f;
      const float ext = fmaxf(hi[0] - lo[0], fmaxf(hi[1] - lo[1], hi[2] - lo[2]));
      if (!(ext < 3.0e38f) || !(cs > 0.f)) {
        // non-finite extent (a NaN / Inf coordinate; the reference's KD-tree returns garbage there, it does not hang):
        // one cell holding everything, so the search degenerates to brute force over the cloud instead of never ending
        c.nx = c.ny = c.nz = 1;
        lo[0] = lo[1] = lo[2] = -3.0e38f;
        cs = 3.0e38f;
      } else
      for (int grow = 0; grow < 512; grow++) {  // 1.25^512 overflows long before: the loop always ends
        double dx = floor((double)(hi[0] - lo[0]) / cs) + 1, dy = floor((double)(hi[1] - lo[1]) / cs) + 1,
               dz = floor((double)(hi[2] - lo[2]) / cs) + 1;
        if (dx <= RS_DIM_CAP && dy <= RS_DIM_CAP && dz <= RS_DIM_CAP && dx * dy * dz <= (double)RS_CELL_CAP) {
          c.nx = (int)dx;
          c.ny = (int)dy;
          c.nz = (int)dz;
          break;
        }
        cs *= 1.25f;
        if (grow == 511) c.nx = c.ny = c.nz = 1;
      }
      c.ox = lo[0];
      c.oy = lo[1];
      c.oz = lo[2];
      c.inv_cs = 1.0f / cs;
      c.ncell = c.nx * c.ny * c.nz;
    }
    clouds[b] = c;
  }
}

__device__ __forceinline__ int rs_cell_coord(float v, float o, float inv_cs, int n) {
  float f = floorf((v - o) * inv_cs);
  f = fminf(fmaxf(f, -2.0f), (float)(n + 1));
  return (int)f;
}

__device__ __forceinline__ int rs_find_cloud_s(const RsCloud* clouds, int batch, long long j) {
  int b = 0;
  while (b + 1 < batch && j >= clouds[b + 1].s_start) b++;
  return b;
}

__global__ void rs_hist_kernel(const float* __restrict__ s_pts, const RsCloud* __restrict__ clouds, int batch,
                               int* __restrict__ cell_cnt, int* __restrict__ pcell, long long ns_cap) {
  long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j >= ns_cap) return;
  int b = rs_find_cloud_s(clouds, batch, j);
  RsCloud c = clouds[b];
  if (j >= c.s_start + c.ns) return;
  int cx = min(max(rs_cell_coord(s_pts[3 * j + 0], c.ox, c.inv_cs, c.nx), 0), c.nx - 1);
  int cy = min(max(rs_cell_coord(s_pts[3 * j + 1], c.oy, c.inv_cs, c.ny), 0), c.ny - 1);
  int cz = min(max(rs_cell_coord(s_pts[3 * j + 2], c.oz, c.inv_cs, c.nz), 0), c.nz - 1);
  int cell = (cz * c.ny + cy) * c.nx + cx;
  pcell[j] = cell;
  atomicAdd(&cell_cnt[(size_t)b * RS_CELL_CAP + cell], 1);
}

__global__ void __launch_bounds__(1024) rs_scan_kernel(const RsCloud* __restrict__ clouds, int* __restrict__ cell_cnt,
                                                       int* __restrict__ cell_start) {
  __shared__ int s_scan[33];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int ncell = clouds[b].ncell;
  int* cnt = cell_cnt + (size_t)b * RS_CELL_CAP;
  int* st = cell_start + (size_t)b * (RS_CELL_CAP + 1);
  int chunk = (ncell + nt - 1) / nt;
  int beg = min(ncell, tid * chunk), end = min(ncell, beg + chunk);
  int s = 0;
  for (int i = beg; i < end; i++) s += cnt[i];
  int total;
  int pre = block_exclusive_scan(s, s_scan, &total);
  for (int i = beg; i < end; i++) {
    int v = cnt[i];
    st[i] = pre;
    pre += v;
    cnt[i] = 0;  // becomes the scatter cursor
  }
  if (tid == 0) st[ncell] = total;
}

__global__ void rs_scatter_kernel(const float* __restrict__ s_pts, const RsCloud* __restrict__ clouds, int batch,
                                  int* __restrict__ cell_cnt, const int* __restrict__ cell_start,
                                  const int* __restrict__ pcell, float4* __restrict__ sorted, int* __restrict__ order,
                                  long long ns_cap) {
  long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j >= ns_cap) return;
  int b = rs_find_cloud_s(clouds, batch, j);
  const RsCloud& c = clouds[b];
  if (j >= c.s_start + c.ns) return;
  int cell = pcell[j];
  int pos = cell_start[(size_t)b * (RS_CELL_CAP + 1) + cell] + atomicAdd(&cell_cnt[(size_t)b * RS_CELL_CAP + cell], 1);
  sorted[c.s_start + pos] =
      make_float4(s_pts[3 * j + 0], s_pts[3 * j + 1], s_pts[3 * j + 2], __int_as_float((int)(j - c.s_start)));
  if (order != nullptr) order[c.s_start + pos] = (int)j;  // cell-sorted (spatially coherent) order of the supports
}

// One warp per query. Hits are streamed into a per-warp shared buffer of 2*KP (d2,idx) keys; when it fills up it
// is sorted and cut back to the KP best, so the result is the exact `limit` smallest under the total order
// (d2, idx) for any neighbour count.
template <typename IdxT>
__global__ void __launch_bounds__(128) rs_query_kernel(const float* __restrict__ q_pts,
                                                       const RsCloud* __restrict__ clouds, int batch,
                                                       const int* __restrict__ cell_start,
                                                       const float4* __restrict__ sorted, float radius, int limit,
                                                       int KP, IdxT* __restrict__ out, int* __restrict__ counts,
                                                       int* __restrict__ max_count, long long nq_cap,
                                                       long long ns_total_pad) {
  extern __shared__ unsigned long long s_buf[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long q = blockIdx.x * (long long)(blockDim.x >> 5) + warp;
  if (q >= nq_cap) return;
  int b = 0;
  while (b + 1 < batch && q >= clouds[b + 1].q_start) b++;
  const RsCloud c = clouds[b];
  if (q >= c.q_start + c.nq) return;  // beyond the real number of queries (capacity launch)
  unsigned long long* buf = s_buf + (size_t)warp * 2 * KP;
  const float qx = q_pts[3 * q + 0], qy = q_pts[3 * q + 1], qz = q_pts[3 * q + 2];
  const float r2 = __fmul_rn(radius, radius);  // radius_neighbors_cpu.cpp:12
  int cnt = 0, total = 0;
  if (c.ns > 0) {
    int cx = rs_cell_coord(qx, c.ox, c.inv_cs, c.nx), cy = rs_cell_coord(qy, c.oy, c.inv_cs, c.ny),
        cz = rs_cell_coord(qz, c.oz, c.inv_cs, c.nz);
    int xlo = max(cx - 1, 0), xhi = min(cx + 1, c.nx - 1);
    int ylo = max(cy - 1, 0), yhi = min(cy + 1, c.ny - 1);
    int zlo = max(cz - 1, 0), zhi = min(cz + 1, c.nz - 1);
    const int* st = cell_start + (size_t)b * (RS_CELL_CAP + 1);
    const float4* sp = sorted + c.s_start;
    if (xlo <= xhi) {
      for (int z = zlo; z <= zhi; z++) {
        for (int y = ylo; y <= yhi; y++) {
          int row = (z * c.ny + y) * c.nx;
          int beg = st[row + xlo], end = st[row + xhi + 1];
          for (int p0 = beg; p0 < end; p0 += 32) {
            int p = p0 + lane;
            bool hit = false;
            unsigned long long key = 0;
            if (p < end) {
              float4 s = sp[p];
              // nanoflann L2_Simple_Adaptor (nanoflann.hpp:432-440): ((dx*dx) + dy*dy) + dz*dz, no FMA
              float dx = __fsub_rn(qx, s.x), dy = __fsub_rn(qy, s.y), dz = __fsub_rn(qz, s.z);
              float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
              hit = d2 < r2;  // strict (nanoflann.hpp:249-252)
              key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)__float_as_int(s.w);
            }
            unsigned int m = __ballot_sync(FULL_MASK, hit);
            if (m) {
              if (cnt + 32 > 2 * KP) {  // make room: keep the KP best so far
                for (int i = cnt + lane; i < 2 * KP; i += 32) buf[i] = ~0ULL;
                __syncwarp();
                warp_bitonic_sort_u64(buf, 2 * KP, lane);
                cnt = min(cnt, KP);
              }
              if (hit) buf[cnt + __popc(m & ((1u << lane) - 1))] = key;
              cnt += __popc(m);
              total += __popc(m);
              __syncwarp();
            }
          }
        }
      }
    }
  }
  if (lane == 0) {
    if (counts) counts[q] = total;
    atomicMax(max_count, total);
  }
  if (limit <= 0) return;  // count-only pass
  int P = 32;
  while (P < cnt) P <<= 1;
  for (int i = cnt + lane; i < P; i += 32) buf[i] = ~0ULL;
  __syncwarp();
  warp_bitonic_sort_u64(buf, P, lane);
  IdxT* row = out + q * (long long)limit;
  for (int i = lane; i < limit; i += 32) {
    // radius_neighbors_cpu.cpp:83-85: global index = local + cloud offset; pad with the total number of supports
    row[i] = (i < cnt) ? (IdxT)((long long)(unsigned int)(buf[i] & 0xffffffffULL) + c.s_start) : (IdxT)ns_total_pad;
  }
}

extern "C" size_t rdm_radius_search_workspace(int64_t ns_cap, int batch) {
  size_t bytes = 0;
  bytes += align_up(sizeof(RsCloud) * batch, 256);
  bytes += align_up((size_t)batch * RS_CELL_CAP * 4, 256);
  bytes += align_up((size_t)batch * (RS_CELL_CAP + 1) * 4, 256);
  bytes += align_up((size_t)ns_cap * 4, 256);
  bytes += align_up((size_t)ns_cap * 16, 256);
  return bytes + 4096;
}

// out_order (optional, [ns_cap] ints): the supports in cell-sorted order - a spatially coherent processing order that
// the KPConv gather uses for its queries when the search is a self search (rdm_build_pyramid).
int rdm_radius_search_impl(const float* q_points, const float* s_points, const int64_t* q_lengths,
                           const int64_t* s_lengths, int batch, int64_t nq_cap, int64_t ns_cap,
                           int64_t ns_total_pad, float radius, int limit, void* out_indices, int index_bytes,
                           int* out_counts, int* out_max_count, int* out_order, void* workspace, size_t workspace_bytes,
                           cudaStream_t stream) {
  RDM_CHECK_ARG(batch >= 1 && nq_cap >= 0 && ns_cap >= 0 && radius > 0.f, "rdm_radius_search: bad arguments");
  RDM_CHECK_ARG(index_bytes == 4 || index_bytes == 8, "rdm_radius_search: index_bytes must be 4 or 8");
  RDM_CHECK_ARG(limit <= 4096, "rdm_radius_search: limit > 4096 unsupported");
  RDM_CHECK_ARG(out_max_count != nullptr, "rdm_radius_search: out_max_count is required");
  Workspace ws(workspace, workspace_bytes);
  RsCloud* clouds = ws.get<RsCloud>(batch);
  int* cell_cnt = ws.get<int>((size_t)batch * RS_CELL_CAP);
  int* cell_start = ws.get<int>((size_t)batch * (RS_CELL_CAP + 1));
  int* pcell = ws.get<int>(ns_cap);
  float4* sorted = ws.get<float4>(ns_cap);
  if (!ws.ok) {
    rdm_set_error("rdm_radius_search: workspace too small (%zu < %zu)", workspace_bytes, ws.off);
    return RDM_ERR_WORKSPACE;
  }
  rs_bounds_kernel<<<batch, 1024, 0, stream>>>(s_points, q_lengths, s_lengths, batch, radius, clouds, out_max_count);
  RDM_LAUNCH_CHECK();
  if (nq_cap == 0) return RDM_OK;
  RDM_CUDA(cudaMemsetAsync(cell_cnt, 0, (size_t)batch * RS_CELL_CAP * 4, stream));
  if (ns_cap > 0) {
    rs_hist_kernel<<<cdiv(ns_cap, 256), 256, 0, stream>>>(s_points, clouds, batch, cell_cnt, pcell, ns_cap);
    RDM_LAUNCH_CHECK();
  }
  rs_scan_kernel<<<batch, 1024, 0, stream>>>(clouds, cell_cnt, cell_start);
  RDM_LAUNCH_CHECK();
  if (ns_cap > 0) {
    rs_scatter_kernel<<<cdiv(ns_cap, 256), 256, 0, stream>>>(s_points, clouds, batch, cell_cnt, cell_start, pcell,
                                                            sorted, out_order, ns_cap);
    RDM_LAUNCH_CHECK();
  }
  int KP = 32;
  while (KP < limit) KP <<= 1;
  int warps = 4;
  size_t smem = (size_t)warps * 2 * KP * 8;
  while (smem > 200 * 1024 && warps > 1) {
    warps >>= 1;
    smem = (size_t)warps * 2 * KP * 8;
  }
  int grid = cdiv(nq_cap, warps);
  if (index_bytes == 8) {
    if (smem > 48 * 1024)
      RDM_CUDA(cudaFuncSetAttribute(rs_query_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rs_query_kernel<int64_t><<<grid, warps * 32, smem, stream>>>(q_points, clouds, batch, cell_start, sorted, radius,
                                                                limit, KP, (int64_t*)out_indices, out_counts,
                                                                out_max_count, nq_cap, ns_total_pad);
  } else {
    if (smem > 48 * 1024)
      RDM_CUDA(cudaFuncSetAttribute(rs_query_kernel<int>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rs_query_kernel<int><<<grid, warps * 32, smem, stream>>>(q_points, clouds, batch, cell_start, sorted, radius, limit,
                                                            KP, (int*)out_indices, out_counts, out_max_count, nq_cap,
                                                            ns_total_pad);
  }
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}

extern "C" int rdm_radius_search(const float* q_points, const float* s_points, const int64_t* q_lengths,
                                 const int64_t* s_lengths, int batch, int64_t nq_cap, int64_t ns_cap,
                                 int64_t ns_total_pad, float radius, int limit, void* out_indices, int index_bytes,
                                 int* out_counts, int* out_max_count, void* workspace, size_t workspace_bytes,
                                 cudaStream_t stream) {
  return rdm_radius_search_impl(q_points, s_points, q_lengths, s_lengths, batch, nq_cap, ns_cap, ns_total_pad, radius, limit,
                                out_indices, index_bytes, out_counts, out_max_count, nullptr, workspace, workspace_bytes,
                                stream);
}


// ---------------------------------------------------------------------------------------------- walk order by load
// order_out = order_in stably partitioned into 4 classes of neighbourhood fill (>= 3/4, >= 1/2, >= 1/4 of the H slots valid,
// less), heaviest class first. Rows of the table are valid-first, so a class test is one probe. The KPConv gather walks
// its queries in this order: neighbours in space stay neighbours inside a class (L1 reuse), and the partial last wave of
// the launch is made of the cheapest queries (measured: -7 % gather time, 0.57 -> 0.61 of the HBM peak on G in the bench).
__device__ __forceinline__ int load_class(const int* __restrict__ nb, int q, int H, int N) {
  const int* row = nb + (size_t)q * H;
  if (row[(3 * H) / 4 - 1 < 0 ? 0 : (3 * H) / 4 - 1] < N) return 0;
  if (row[H / 2 - 1 < 0 ? 0 : H / 2 - 1] < N) return 1;
  if (row[H / 4 - 1 < 0 ? 0 : H / 4 - 1] < N) return 2;
  return 3;
}
// three small launches: per-CTA class counts, scan of the counts (class-major), stable scatter
__global__ void __launch_bounds__(1024) obl_count_kernel(const int* __restrict__ order_in, const int* __restrict__ nb, int n, int H, int N,
                                                         unsigned char* __restrict__ cls, int* __restrict__ blk_cnt) {
  __shared__ int s_cnt[4];
  if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int i = blockIdx.x * 1024 + threadIdx.x;
  int c = -1;
  if (i < n) {
    c = load_class(nb, order_in[i], H, N);
    cls[i] = (unsigned char)c;
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const unsigned m = __ballot_sync(FULL_MASK, c == k);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt[k], __popc(m));
  }
  __syncthreads();
  if (threadIdx.x < 4) blk_cnt[threadIdx.x * gridDim.x + blockIdx.x] = s_cnt[threadIdx.x];  // class-major
}
__global__ void __launch_bounds__(1024) obl_scan_kernel(int* __restrict__ blk_cnt, int total) {
  __shared__ int s_scan[33];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < total; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    const int v = i < total ? blk_cnt[i] : 0;
    int tot;
    const int ex = block_exclusive_scan(v, s_scan, &tot);
    const int base = s_base;
    if (i < total) blk_cnt[i] = base + ex;
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + tot;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(1024) obl_scatter_kernel(const int* __restrict__ order_in, const unsigned char* __restrict__ cls, int n,
                                                           const int* __restrict__ blk_base, int* __restrict__ order_out) {
  __shared__ int s_warp[32][4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, i = blockIdx.x * 1024 + tid;
  const int c = i < n ? (int)cls[i] : -1;
  int pos = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const unsigned m = __ballot_sync(FULL_MASK, c == k);
    if (c == k) pos = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp][k] = __popc(m);
  }
  __syncthreads();
  if (c >= 0) {
    int before = 0;
    for (int w = 0; w < warp; w++) before += s_warp[w][c];
    order_out[blk_base[c * gridDim.x + blockIdx.x] + before + pos] = order_in[i];
  }
}

// scratch: n bytes (classes) + 4 * ceil(n / 1024) ints, from the pyramid builder's workspace
int rdm_order_by_load(const int* order_in, const int* neighbors, int n, int H, int n_support, int* order_out, void* scratch,
                      size_t scratch_bytes, cudaStream_t stream) {
  if (n <= 0) return RDM_OK;
  const int nblk = cdiv(n, 1024);
  const size_t need = align_up((size_t)n, 256) + (size_t)4 * nblk * sizeof(int);
  RDM_CHECK_ARG(scratch != nullptr && scratch_bytes >= need, "rdm_order_by_load: scratch too small");
  unsigned char* cls = (unsigned char*)scratch;
  int* blk = (int*)((char*)scratch + align_up((size_t)n, 256));
  obl_count_kernel<<<nblk, 1024, 0, stream>>>(order_in, neighbors, n, H, n_support, cls, blk);
  RDM_LAUNCH_CHECK();
  obl_scan_kernel<<<1, 1024, 0, stream>>>(blk, 4 * nblk);
  RDM_LAUNCH_CHECK();
  obl_scatter_kernel<<<nblk, 1024, 0, stream>>>(order_in, cls, n, blk, order_out);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}


// ---------------------------------------------------------------------------------------------- reference row width
// The reference's table has width min(max_count, limit) (radius_neighbors_cpu.cpp:59-68 + ops/radius_search.py:25-26);
// ours has the fixed width `limit`. For every consumer but one the extra columns are ordinary padding. The exception is
// the strided max-pool (kpconv/functional.py:54-67): padding reads the appended ZERO row, so a row that fills the whole
// reference width competes with 0 only if a padded column exists. Columns >= max_count therefore get the sentinel N + 1
// = "this column does not exist in the reference's tensor"; rdm_maxpool skips it, the gathers treat it as padding.
__global__ void __launch_bounds__(256) rs_mark_width_kernel(int* __restrict__ table, long long rows, int H, int N,
                                                            const int* __restrict__ max_count) {
  const int w = *max_count;
  if (w >= H) return;
  const long long e = blockIdx.x * 256LL + threadIdx.x;
  if (e >= rows * H) return;
  if ((int)(e % H) >= w) table[e] = N + 1;
}
int rdm_mark_reference_width(int* table, long long rows, int H, int n_support, const int* d_max_count, cudaStream_t stream) {
  if (rows <= 0) return RDM_OK;
  rs_mark_width_kernel<<<cdiv(rows * H, 256), 256, 0, stream>>>(table, rows, H, n_support, d_max_count);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}
