// Error reporting + version for librdm_sm100.so.
#include <stdarg.h>
#include <stdlib.h>
#include <vector>
#include "common.cuh"
#include "../../include/rdm_sm100.h"

static thread_local char g_err[1024] = "";

void rdm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* rdm_last_error(void) { return g_err; }
extern "C" int rdm_version(void) { return 101; }

bool rdm_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("RDM_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

unsigned long long g_rdm_launches = 0;
extern "C" unsigned long long rdm_launch_count(void) { return __atomic_load_n(&g_rdm_launches, __ATOMIC_RELAXED); }

// ---- optional in-library kernel timing (bench.py: live roofline of the KPConv gather kernel). Events are recorded on
// the launch stream right around the launches, so the measured span holds the kernels and nothing of the host loop.
namespace {
struct ProfRec {
  cudaEvent_t e0, e1;
  int tag, m, n, h, c;
};
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_prof_pool;  // events are created when profiling is switched on, never inside a timed region
int g_prof_mask = 0;  // bit 0: KPConv gather brackets, bit 1: KPConv weight-GEMM brackets
bool g_prof_on = false;
cudaEvent_t prof_take() {
  if (g_prof_pool.empty()) {
    cudaEvent_t e = nullptr;
    return cudaEventCreate(&e) == cudaSuccess ? e : nullptr;
  }
  cudaEvent_t e = g_prof_pool.back();
  g_prof_pool.pop_back();
  return e;
}
}  // namespace

int rdm_prof_begin(int tag, int m, int n, int h, int c, cudaStream_t stream) {
  if (!g_prof_on || !(g_prof_mask & (1 << (tag - 1)))) return -1;
  ProfRec r;
  r.tag = tag; r.m = m; r.n = n; r.h = h; r.c = c;
  r.e0 = prof_take();
  r.e1 = prof_take();
  if (r.e0 == nullptr || r.e1 == nullptr) return -1;
  cudaEventRecord(r.e0, stream);
  g_prof.push_back(r);
  return (int)g_prof.size() - 1;
}

void rdm_prof_end(int id, cudaStream_t stream) {
  if (id >= 0 && id < (int)g_prof.size()) cudaEventRecord(g_prof[id].e1, stream);
}

extern "C" void rdm_prof_enable(int on) {
  for (auto& r : g_prof) {  // recycle
    g_prof_pool.push_back(r.e0);
    g_prof_pool.push_back(r.e1);
  }
  g_prof.clear();
  g_prof.reserve(1 << 14);
  if (on) {
    while (g_prof_pool.size() < 8192) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreate(&e) != cudaSuccess) break;
      g_prof_pool.push_back(e);
    }
  }
  g_prof_on = on != 0;
  g_prof_mask = on;  // rdm_prof_enable(1): gather only; (3): gather + weight GEMM
}

extern "C" int rdm_prof_read(rdm_prof_record* out, int max_records) {
  int n = 0;
  for (auto& r : g_prof) {
    if (n >= max_records) break;
    float ms = 0.f;
    if (cudaEventSynchronize(r.e1) != cudaSuccess || cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
    out[n].tag = r.tag; out[n].ms = ms; out[n].m = r.m; out[n].n = r.n; out[n].h = r.h; out[n].c = r.c;
    n++;
  }
  return n;
}

// ---- ABI self-description: sizeof / selected offsetof of every struct of include/rdm_sm100.h, so that a host binding
// (rdmnet_b200/_lib.py's ctypes mirrors) can be checked against the compiled layout without a GPU (tests/test_abi_cpu.py).
#include <stddef.h>
extern "C" int rdm_abi_layout(int64_t* out, int max_entries) {
  const int64_t v[] = {
      (int64_t)sizeof(rdm_prof_record),      (int64_t)sizeof(rdm_tf_proj_job),     (int64_t)sizeof(rdm_tf_attn_job),
      (int64_t)sizeof(rdm_unary_desc),       (int64_t)sizeof(rdm_block_desc),      (int64_t)sizeof(rdm_pyramid_desc),
      (int64_t)sizeof(rdm_pyramid_cfg),      (int64_t)sizeof(rdm_thdroformer_desc), (int64_t)sizeof(rdm_backbone_desc),
      (int64_t)sizeof(rdm_backbone_out),     (int64_t)sizeof(rdm_match_desc),      (int64_t)sizeof(rdm_match_io),
      (int64_t)sizeof(rdm_match_result),
      (int64_t)offsetof(rdm_block_desc, sigma), (int64_t)offsetof(rdm_pyramid_desc, order), (int64_t)offsetof(rdm_match_desc, nms_limit),
      (int64_t)offsetof(rdm_match_io, transform), (int64_t)offsetof(rdm_match_result, transform)};
  const int n = (int)(sizeof(v) / sizeof(v[0]));
  for (int i = 0; i < n && i < max_entries; i++) out[i] = v[i];
  return n;
}
