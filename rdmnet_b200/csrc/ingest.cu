// Ingest: raw sensor scan -> 0.3 m voxel-barycentre downsample on the GPU, the step the reference does offline with open3d
// (preporcess/downsample_pcd_kitti.py:20-36: pcd.voxel_down_sample(0.3) on xyz with the intensity carried as colour).
// open3d semantics: voxel index = floor((p - (min_bound - voxel/2)) / voxel); every occupied voxel yields the mean of its
// points (and of their attributes). open3d emits the voxels in its hash-map order, which is unspecified; here the order
// is canonical: voxels appear in the order of their FIRST point in the input (deterministic, sort-free).
//   K1  bounds (per-axis min)                    K2  hash insert: voxel -> slot, fixed-point sums (order independent),
//   K3  single-CTA scan over "first point" flags     first-point index by atomicMin
//   K4  emit means
#include "common.cuh"
#include "../../include/rdm_sm100.h"

namespace {
constexpr unsigned long long EMPTY = 0xffffffffffffffffull;
constexpr double FIX = 1048576.0;  // 2^20 fixed point: |coord| < 2^22 m, 1e-6 m resolution, exact integer accumulation

struct Slot {
  unsigned long long key;
  long long sum[4];
  int count, first;
};

__global__ void __launch_bounds__(1024) vd_bounds_kernel(const float* __restrict__ p, int stride, int n, float* __restrict__ mn) {
  __shared__ float s[3][32];
  float m[3] = {3.4e38f, 3.4e38f, 3.4e38f};
  for (int i = threadIdx.x; i < n; i += 1024)
    for (int a = 0; a < 3; a++) m[a] = fminf(m[a], p[(size_t)i * stride + a]);
  for (int a = 0; a < 3; a++) {
    float v = m[a];
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL_MASK, v, o));
    if ((threadIdx.x & 31) == 0) s[a][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float v = s[threadIdx.x][0];
    for (int i = 1; i < 32; i++) v = fminf(v, s[threadIdx.x][i]);
    mn[threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) vd_insert_kernel(const float* __restrict__ p, int stride, int n, float voxel,
                                                        const float* __restrict__ mn, Slot* __restrict__ table, unsigned cap_mask,
                                                        int* __restrict__ slot_of) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float* q = p + (size_t)i * stride;
  unsigned long long key = 0;
  for (int a = 0; a < 3; a++) {
    const float o = mn[a] - 0.5f * voxel;  // open3d: voxel_min_bound = min_bound - voxel_size / 2
    const long long c = (long long)floorf((q[a] - o) / voxel);
    key = (key << 21) | (unsigned long long)(c & 0x1fffff);
  }
  unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 32) & cap_mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(&table[h].key, EMPTY, key);
    if (prev == EMPTY || prev == key) break;
    h = (h + 1) & cap_mask;
  }
  slot_of[i] = (int)h;
  for (int a = 0; a < stride && a < 4; a++)
    atomicAdd((unsigned long long*)&table[h].sum[a], (unsigned long long)llrint((double)q[a] * FIX));
  atomicAdd(&table[h].count, 1);
  atomicMin(&table[h].first, i);
}

__global__ void __launch_bounds__(1024) vd_scan_kernel(const Slot* __restrict__ table, const int* __restrict__ slot_of, int n,
                                                       int* __restrict__ pos, int* __restrict__ out_count) {
  __shared__ int s_scan[33];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    const int flag = (i < n && table[slot_of[i]].first == i) ? 1 : 0;
    int tot;
    const int ex = block_exclusive_scan(flag, s_scan, &tot);
    const int base = s_base;
    if (i < n) pos[i] = flag ? base + ex : -1;
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_count = s_base;
}

__global__ void __launch_bounds__(256) vd_emit_kernel(const Slot* __restrict__ table, const int* __restrict__ slot_of,
                                                      const int* __restrict__ pos, int n, int stride, float* __restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n || pos[i] < 0) return;
  const Slot& s = table[slot_of[i]];
  for (int a = 0; a < stride && a < 4; a++) out[(size_t)pos[i] * stride + a] = (float)((double)s.sum[a] / FIX / (double)s.count);
}

__global__ void __launch_bounds__(256) vd_init_kernel(Slot* t, unsigned cap) {
  const unsigned i = blockIdx.x * 256 + threadIdx.x;
  if (i >= cap) return;
  t[i].key = EMPTY;
  t[i].sum[0] = t[i].sum[1] = t[i].sum[2] = t[i].sum[3] = 0;
  t[i].count = 0;
  t[i].first = 0x7fffffff;
}

unsigned table_cap(int n) {
  unsigned c = 1024;
  while (c < 2u * (unsigned)n) c <<= 1;
  return c;
}
}  // namespace

extern "C" size_t rdm_voxel_downsample_workspace(int n) {
  return align_up((size_t)table_cap(n) * sizeof(Slot), 256) + 2 * align_up((size_t)n * 4, 256) + 1024;
}

extern "C" int rdm_voxel_downsample(const float* points, int stride, int n, float voxel, float* out, int* out_count, void* workspace,
                                    size_t workspace_bytes, cudaStream_t stream) {
  RDM_CHECK_ARG((stride == 3 || stride == 4) && n >= 0 && voxel > 0.f, "rdm_voxel_downsample: rows of 3 (xyz) or 4 (xyzi) floats");
  if (n == 0) {
    RDM_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int), stream));
    return RDM_OK;
  }
  Workspace ws(workspace, workspace_bytes);
  const unsigned cap = table_cap(n);
  Slot* table = ws.get<Slot>(cap);
  int* slot_of = ws.get<int>(n);
  int* pos = ws.get<int>(n);
  float* mn = ws.get<float>(4);
  if (!ws.ok) {
    rdm_set_error("rdm_voxel_downsample: workspace too small");
    return RDM_ERR_WORKSPACE;
  }
  vd_init_kernel<<<cdiv(cap, 256), 256, 0, stream>>>(table, cap);
  RDM_LAUNCH_CHECK();
  vd_bounds_kernel<<<1, 1024, 0, stream>>>(points, stride, n, mn);
  RDM_LAUNCH_CHECK();
  vd_insert_kernel<<<cdiv(n, 256), 256, 0, stream>>>(points, stride, n, voxel, mn, table, cap - 1, slot_of);
  RDM_LAUNCH_CHECK();
  vd_scan_kernel<<<1, 1024, 0, stream>>>(table, slot_of, n, pos, out_count);
  RDM_LAUNCH_CHECK();
  vd_emit_kernel<<<cdiv(n, 256), 256, 0, stream>>>(table, slot_of, pos, n, stride, out);
  RDM_LAUNCH_CHECK();
  return RDM_OK;
}
