// TEST INFRASTRUCTURE ONLY (oracle). C-ABI shim around the reference's own C++ core, compiled from the
// sources where they lie under /root/reference (never copied): see oracle/Makefile. Output: oracle/_ref/libref_ext.so.
// It bypasses only the torch tensor wrappers (grid_subsampling.cpp:5-62, radius_neighbors.cpp:5-67), which do
// nothing but copy tensors into the std::vectors used below.
#include <cstdint>
#include <cstring>
#include <vector>
#include "cpu/grid_subsampling/grid_subsampling_cpu.h"
#include "cpu/radius_neighbors/radius_neighbors_cpu.h"

extern "C" {

// returns total number of subsampled points; out_points capacity must be >= n_total*3 floats
int64_t ref_grid_subsampling(const float* points, int64_t n_total, const int64_t* lengths, int64_t batch,
                             float voxel, float* out_points, int64_t* out_lengths) {
  std::vector<PointXYZ> pts(reinterpret_cast<const PointXYZ*>(points),
                            reinterpret_cast<const PointXYZ*>(points) + n_total);
  std::vector<long> len(lengths, lengths + batch);
  std::vector<PointXYZ> s_pts;
  std::vector<long> s_len;
  grid_subsampling_cpu(pts, s_pts, len, s_len, voxel);
  std::memcpy(out_points, s_pts.data(), sizeof(float) * 3 * s_pts.size());
  for (int64_t b = 0; b < batch; b++) out_lengths[b] = s_len[b];
  return (int64_t)s_pts.size();
}

// Two-call protocol: the result is cached between the "count" call (out == nullptr, returns max_count) and the
// "fetch" call (out != nullptr, copies nq*max_count int64). Single-threaded test helper.
static std::vector<long> g_last;
int64_t ref_radius_neighbors(const float* q, int64_t nq, const float* s, int64_t ns, const int64_t* q_len,
                             const int64_t* s_len, int64_t batch, float radius, int64_t* out) {
  if (out == nullptr) {
    std::vector<PointXYZ> qv(reinterpret_cast<const PointXYZ*>(q), reinterpret_cast<const PointXYZ*>(q) + nq);
    std::vector<PointXYZ> sv(reinterpret_cast<const PointXYZ*>(s), reinterpret_cast<const PointXYZ*>(s) + ns);
    std::vector<long> ql(q_len, q_len + batch), sl(s_len, s_len + batch);
    g_last.clear();
    radius_neighbors_cpu(qv, sv, ql, sl, g_last, radius);
    return nq > 0 ? (int64_t)(g_last.size() / (size_t)nq) : 0;
  }
  for (size_t i = 0; i < g_last.size(); i++) out[i] = g_last[i];
  return nq > 0 ? (int64_t)(g_last.size() / (size_t)nq) : 0;
}
}
