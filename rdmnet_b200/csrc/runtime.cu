// Host-side module runners: one C call walks all kernels of Encoder / Decoder / ThDRoFormer.
// Reference wiring: experiments/backbone.py:72-107 (Encoder.forward), :118-151 (Decoder.forward),
// geotransformer/modules/kpconv/modules.py:53-225 (UnaryBlock / ConvBlock / ResidualBlock),
// rdmnet/thdroformer/thdroformer.py:229-251, 304-347 (RPEConditionalTransformer / ThDRoFormer forward).
// No kernel lives here: the functions below only sequence the launch functions of the other translation units and
// carve temporaries out of the caller's workspace (the same code runs "dry" to size that workspace).
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "../../include/rdm_sm100.h"

namespace {
struct Arena {
  char* base;
  size_t cap, off, peak;
  bool dry;
  Arena(void* p, size_t n, bool d) : base((char*)p), cap(n), off(0), peak(0), dry(d) {}
  float* f(size_t count) { return (float*)raw(count * sizeof(float)); }
  void* raw(size_t bytes) {
    off = align_up(off, 256);
    void* r = dry ? nullptr : (void*)(base + off);
    off += bytes;
    if (off > peak) peak = off;
    return r;
  }
};

#define RDM_TRY(call)          \
  do {                         \
    int _rc = (call);          \
    if (_rc != RDM_OK) return _rc; \
  } while (0)

int linear(Arena& a, const float* A, int lda, const float* W, int ldw, int nk, const float* bias, float* C, int M, int N, int K,
           cudaStream_t st, double* gn_stats = nullptr, int gn_cpg = 0, int* stats_fused = nullptr, int ldc = 0) {
  void* ws = nullptr;
  size_t wsb = 0;
  size_t mark = a.off;
  if ((long long)M * N <= (1 << 20)) {
    wsb = rdm_linear_workspace(M, N, K);
    ws = a.raw(wsb);
  }
  int rc = RDM_OK;
  if (!a.dry)
    rc = rdm_linear_gn_ps(A, lda, W, ldw, nk, bias, C, ldc > 0 ? ldc : N, M, N, K, 0, ws, wsb, gn_stats, gn_cpg, stats_fused,
                          nk ? rdm_presplit_lookup(W) : nullptr, st);
  a.off = mark;
  return rc;
}

// GroupNorm statistics slots: one pre-zeroed array of doubles per module run (a single memset instead of one per norm)
struct StatSlots {
  double* base = nullptr;
  int used = 0, cap = 0;
  double* take(int groups) {
    double* p = base ? base + (size_t)used * 2 * 64 : nullptr;
    used++;
    return p;
  }
};
constexpr int kStatSlots = 96;  // >= number of GroupNorms in one module run; 64 groups max per slot

int stat_slots_begin(Arena& a, StatSlots& ss, cudaStream_t st) {
  ss.base = (double*)a.raw(sizeof(double) * 2 * 64 * kStatSlots);
  ss.cap = kStatSlots;
  if (!a.dry) RDM_CUDA(cudaMemsetAsync(ss.base, 0, sizeof(double) * 2 * 64 * kStatSlots, st));
  return RDM_OK;
}

// out = act(GroupNorm(lin(x)) (+ residual)); the statistics ride in the GEMM epilogue when the shape allows
int linear_group_norm(Arena& a, StatSlots& ss, const float* x, int ldx, const float* W, int ldw, int nk, const float* bias,
                      const float* g, const float* b, const float* residual, float* out, int M, int N, int K, int groups, int act,
                      unsigned char* rowpos_out, cudaStream_t st) {
  RDM_CHECK_ARG(groups <= 64 && ss.used < ss.cap, "group norm: too many groups / norms in one module run");
  double* stats = ss.take(groups);
  int fused = 0;
  RDM_TRY(linear(a, x, ldx, W, ldw, nk, bias, out, M, N, K, st, stats, N / groups, &fused));
  if (a.dry) return RDM_OK;
  if (!fused) RDM_TRY(rdm_groupnorm_stats(out, M, N, groups, stats, st));
  return rdm_groupnorm_apply(out, stats, g, b, residual, out, M, N, groups, 1e-5f, act, 0.1f, rowpos_out, st);
}

// UnaryBlock (kpconv/modules.py:78-83): Linear + GroupNorm (+ residual) (+ LeakyReLU); out [M, c_out]
int unary(Arena& a, StatSlots& ss, const rdm_unary_desc& u, const float* x, int ldx, float* out, int M, int groups,
          const float* residual, int act, unsigned char* rowpos_out, cudaStream_t st, int ld_out = 0) {
  const int ldw = u.ldw > 0 ? u.ldw : u.c_in;
  if (u.gn_w != nullptr) {
    RDM_CHECK_ARG(ld_out == 0 || ld_out == u.c_out, "unary: a normalised block writes dense rows");
    return linear_group_norm(a, ss, x, ldx, u.w, ldw, 1, u.b, u.gn_w, u.gn_b, residual, out, M, u.c_out, u.c_in, groups, act,
                             rowpos_out, st);
  }
  return linear(a, x, ldx, u.w, ldw, 1, u.b, out, M, u.c_out, u.c_in, st, nullptr, 0, nullptr, ld_out);
}

// KPConv.forward (kpconv.py:79-122) + norm_conv + LeakyReLU: out [M, c_mid_out]
int kpconv_norm(Arena& a, StatSlots& ss, const rdm_block_desc& b, const float* feats, const unsigned char* rowpos_ready,
                const float* q_pts, const float* s_pts, const void* idx, int index_bytes, int M, int N, int H, const int* order,
                float* out, int groups, cudaStream_t st, cudaEvent_t after_gather = nullptr) {
  size_t mark = a.off;
  float* gathered = a.f((size_t)M * 15 * b.c_mid_in);
  unsigned char* rowpos = rowpos_ready ? const_cast<unsigned char*>(rowpos_ready) : (unsigned char*)a.raw((size_t)(N > 0 ? N : 1));
  if (!a.dry)
    RDM_TRY(rdm_kpconv_gather_impl(feats, q_pts, s_pts, idx, index_bytes, b.kernel_points, b.h_kernel_points, b.sigma, M, N, H,
                                   b.c_mid_in, order, gathered, rowpos, rowpos_ready != nullptr, st));
  if (!a.dry && after_gather != nullptr) RDM_CUDA(cudaEventRecord(after_gather, st));
  const int prof = a.dry ? -1 : rdm_prof_begin(RDM_PROF_KPCONV_GEMM, M, 15 * b.c_mid_in, 0, b.c_mid_out, st);
  const int K = 15 * b.c_mid_in;
  double* stats = ss.take(groups);
  int fused = 0;
  RDM_CHECK_ARG(groups <= 64 && ss.used <= ss.cap, "group norm: too many groups / norms in one module run");
  if (b.kpconv_wt != nullptr)
    RDM_TRY(linear(a, gathered, K, b.kpconv_wt, K, 1, b.kpconv_b, out, M, b.c_mid_out, K, st, stats, b.c_mid_out / groups, &fused));
  else
    RDM_TRY(linear(a, gathered, K, b.kpconv_w, b.c_mid_out, 0, b.kpconv_b, out, M, b.c_mid_out, K, st, stats, b.c_mid_out / groups,
                   &fused));
  rdm_prof_end(prof, st);
  if (!a.dry) {
    if (!fused) RDM_TRY(rdm_groupnorm_stats(out, M, b.c_mid_out, groups, stats, st));
    RDM_TRY(rdm_groupnorm_apply(out, stats, b.norm_conv_w, b.norm_conv_b, nullptr, out, M, b.c_mid_out, groups, 1e-5f, 1, 0.1f,
                                nullptr, st));
  }
  a.off = mark;
  return RDM_OK;
}

// The shortcut branch of a ResidualBlock (max-pool over the neighbours for strided blocks, Linear + GroupNorm when the
// width changes; modules.py:213-220) depends only on the block input. It runs on a second stream, forked right AFTER the
// block's neighbour gather (so the gather is timed alone) and joined before unary2 adds it: ~40 us of small-grid kernels
// per block that overlap the KPConv weight GEMM and its norm instead of extending the chain. RDM_DUAL_STREAM=0 disables.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork[32], join[32];
  int state = 0;  // 0 unknown, 1 ready, -1 off
  bool ready() {
    if (state == 0) {
      const char* e = getenv("RDM_DUAL_STREAM");
      state = (e && e[0] == '0') ? -1 : 1;
      if (state == 1) {
        if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) state = -1;
        for (int i = 0; i < 32 && state == 1; i++)
          if (cudaEventCreateWithFlags(&fork[i], cudaEventDisableTiming) != cudaSuccess ||
              cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming) != cudaSuccess)
            state = -1;
      }
    }
    return state == 1;
  }
};
SideStream g_side;

int encoder_run(Arena& a, const rdm_block_desc* blocks, int nb, const rdm_pyramid_desc& p, int groups, const float* in_feats,
                float* const* out_feats, cudaStream_t st) {
  const float* cur = in_feats;
  int cur_stage = 0;
  StatSlots ss;
  RDM_TRY(stat_slots_begin(a, ss, st));
  for (int i = 0; i < nb; i++) {
    const rdm_block_desc& b = blocks[i];
    const int s = b.stage;
    RDM_CHECK_ARG(s >= 0 && s < p.num_stages && (b.strided ? s == cur_stage + 1 : s == cur_stage), "rdm_encoder_forward: block %d stage", i);
    const int M = p.n[s], N = b.strided ? p.n[s - 1] : p.n[s];
    const float* q_pts = p.points[s];
    const float* s_pts = b.strided ? p.points[s - 1] : p.points[s];
    const void* idx = b.strided ? p.subsampling[s - 1] : p.neighbors[s];
    const int H = b.strided ? p.sub_width[s - 1] : p.nb_width[s];
    const bool last_of_stage = (i + 1 == nb) || (blocks[i + 1].stage != s);
    float* out = last_of_stage ? (a.dry ? nullptr : out_feats[s]) : a.f((size_t)M * b.c_out);
    size_t mark = a.off;
    if (b.unary2.w == nullptr) {  // ConvBlock (modules.py:143-147)
      RDM_TRY(kpconv_norm(a, ss, b, cur, nullptr, q_pts, s_pts, idx, p.index_bytes, M, N, H, p.order[s], out, groups, st));
    } else {  // ResidualBlock (modules.py:205-225)
      const float* x = cur;
      unsigned char* rowpos = nullptr;
      if (b.unary1.w != nullptr) {
        float* t = a.f((size_t)N * b.unary1.c_out);
        // the GroupNorm apply of unary1 also emits the KPConv neighbour-count predicate of its rows (vectorised path only)
        const int c1 = b.unary1.c_out;
        if (b.unary1.gn_w != nullptr && (c1 == 32 || c1 == 64 || c1 == 128 || (c1 % 128 == 0 && c1 <= 4096)))
          rowpos = (unsigned char*)a.raw((size_t)(N > 0 ? N : 1));
        RDM_TRY(unary(a, ss, b.unary1, cur, b.c_in, t, N, groups, nullptr, 1, rowpos, st));
        x = t;
      }
      float* c = a.f((size_t)M * b.c_mid_out);
      // shortcut buffers are carved BEFORE kpconv_norm so that they never alias its (released) gather buffer
      const bool need_mp = b.strided != 0, need_sct = b.shortcut.w != nullptr;  // (pointers are NULL in the sizing dry run)
      float* mp = need_mp ? a.f((size_t)M * b.c_in) : nullptr;
      float* sct = need_sct ? a.f((size_t)M * b.c_out) : nullptr;
      // ... and so is the split-K scratch of the shortcut GEMM: linear() carves it at the current arena offset, which after
      // kpconv_norm would be the gather buffer the main stream's weight GEMM is still reading
      size_t branch_ws = 0;
      if (need_sct && (long long)M * b.c_out <= (1 << 20)) branch_ws = rdm_linear_workspace(M, b.c_out, b.c_in) + 512;
      const size_t branch_off = a.off;
      if (branch_ws) a.raw(branch_ws);
      const bool has_branch = need_mp || need_sct;
      const bool dual = has_branch && !a.dry && i < 32 && g_side.ready();
      RDM_TRY(kpconv_norm(a, ss, b, x, rowpos, q_pts, s_pts, idx, p.index_bytes, M, N, H, p.order[s], c, groups, st,
                          dual ? g_side.fork[i] : nullptr));
      const cudaStream_t bs = dual ? g_side.stream : st;
      if (dual) RDM_CUDA(cudaStreamWaitEvent(bs, g_side.fork[i], 0));
      const float* sc = cur;
      if (b.strided) {
        if (!a.dry) RDM_TRY(rdm_maxpool(cur, idx, p.index_bytes, M, N, H, b.c_in, mp, bs));
        sc = mp;
      }
      if (b.shortcut.w != nullptr) {
        const size_t keep = a.off;
        a.off = branch_off;  // scratch inside the reserved region
        RDM_TRY(unary(a, ss, b.shortcut, sc, b.c_in, sct, M, groups, nullptr, 0, nullptr, bs));
        a.off = keep;
        sc = sct;
      }
      if (dual) {
        RDM_CUDA(cudaEventRecord(g_side.join[i], bs));
        RDM_CUDA(cudaStreamWaitEvent(st, g_side.join[i], 0));
      }
      RDM_TRY(unary(a, ss, b.unary2, c, b.c_mid_out, out, M, groups, sc, 1, nullptr, st));
    }
    a.off = mark;
    cur = out;
    cur_stage = s;
  }
  return RDM_OK;
}

int decoder_run(Arena& a, const rdm_unary_desc* dec, int num, const rdm_pyramid_desc& p, int top, int groups, const float* coarse,
                int c_coarse, const float* const* skips, float* out, int ld_out, cudaStream_t st) {
  const float* x = coarse;
  int cx = c_coarse;
  StatSlots ss;
  RDM_TRY(stat_slots_begin(a, ss, st));
  for (int i = 0; i < num; i++) {
    const int s = top - 1 - i;  // target stage
    RDM_CHECK_ARG(s >= 0, "rdm_decoder_forward: too many levels");
    const int M = p.n[s], N = p.n[s + 1];
    const int c_skip = dec[i].c_in - cx;
    RDM_CHECK_ARG(c_skip >= 0, "rdm_decoder_forward: channel mismatch at level %d", i);
    // row stride padded to a multiple of 4 floats: 16-byte rows qualify the GEMM for the TMA / tensor-core path
    const int ldc = (dec[i].c_in + 3) / 4 * 4;
    float* cat = a.f((size_t)M * ldc);
    if (!a.dry)
      RDM_TRY(rdm_upsample_concat_ld(x, p.upsampling[s], p.index_bytes, p.up_width[s], c_skip ? skips[i] : nullptr, M, N, cx,
                                     c_skip, cat, ldc, st));
    float* y = (i + 1 == num) ? out : a.f((size_t)M * dec[i].c_out);
    RDM_TRY(unary(a, ss, dec[i], cat, ldc, y, M, groups, nullptr, 1, nullptr, st, (i + 1 == num) ? ld_out : 0));
    x = y;
    cx = dec[i].c_out;
  }
  return RDM_OK;
}

int thdroformer_run(Arena& a, const rdm_thdroformer_desc& d, const float* rp, int n0, const float* sp, int n1, const float* rf,
                    int ld0, const float* sf, int ld1, float* o0, float* o1, cudaStream_t st) {
  const int D = 128;
  float* e0 = a.f((size_t)n0 * 64);
  float* e1 = a.f((size_t)n1 * 64);
  float* f0 = a.f((size_t)n0 * D);
  float* f1 = a.f((size_t)n1 * D);
  float* g0 = a.f((size_t)n0 * D);
  float* g1 = a.f((size_t)n1 * D);
  float *q0 = a.f((size_t)n0 * D), *v0 = a.f((size_t)n0 * D), *q1 = a.f((size_t)n1 * D), *v1 = a.f((size_t)n1 * D);
  const int ldk0 = (n0 + 3) / 4 * 4, ldk1 = (n1 + 3) / 4 * 4;
  float *k0 = a.f((size_t)D * ldk0), *k1 = a.f((size_t)D * ldk1);
  RDM_TRY(linear(a, rp, 3, d.emb_w, 3, 1, d.emb_b, e0, n0, 64, 3, st));
  RDM_TRY(linear(a, sp, 3, d.emb_w, 3, 1, d.emb_b, e1, n1, 64, 3, st));
  RDM_TRY(linear(a, rf, ld0, d.in_w, d.c_in, 1, d.in_b, f0, n0, D, d.c_in, st));
  RDM_TRY(linear(a, sf, ld1, d.in_w, d.c_in, 1, d.in_b, f1, n1, D, d.c_in, st));
  const size_t WQ = 0, WK = 128 * 128, WV = 2 * 128 * 128, BQ = 4 * 128 * 128 + 2 * 128 * 256, BK = BQ + 128, BV = BQ + 256;
  auto pj = [](const float* x, const float* blob, size_t w, size_t b, const float* emb, float* y, int n, int ldy_t) {
    rdm_tf_proj_job j;
    j.x = x; j.wt = blob + w; j.bias = blob + b; j.emb = emb; j.y = y; j.n = n; j.ldx = 128; j.lde = 64; j.ldy_t = ldy_t;
    return j;
  };
  auto aj = [](const float* q, const float* k, const float* v, const float* x, const float* blob, float* out, int nq, int nk, int ldk) {
    rdm_tf_attn_job j;
    j.q = q; j.k = k; j.v = v; j.x = x; j.blob = blob; j.out = out; j.nq = nq; j.nk = nk; j.ldx = 128; j.ldk_t = ldk;
    return j;
  };
  for (int l = 0; l < (a.dry ? 0 : d.num_layers); l++) {
    const float* B = d.layer_blobs[l];
    if (d.is_self[l]) {
      rdm_tf_proj_job pjs[6] = {pj(f0, B, WQ, BQ, e0, q0, n0, 0), pj(f0, B, WK, BK, e0, k0, n0, ldk0), pj(f0, B, WV, BV, nullptr, v0, n0, 0),
                                pj(f1, B, WQ, BQ, e1, q1, n1, 0), pj(f1, B, WK, BK, e1, k1, n1, ldk1), pj(f1, B, WV, BV, nullptr, v1, n1, 0)};
      RDM_TRY(rdm_tf_project(pjs, 6, st));
      rdm_tf_attn_job ajs[2] = {aj(q0, k0, v0, f0, B, g0, n0, n0, ldk0), aj(q1, k1, v1, f1, B, g1, n1, n1, ldk1)};
      RDM_TRY(rdm_tf_attend(ajs, 2, st));
      float* t = f0; f0 = g0; g0 = t;
      t = f1; f1 = g1; g1 = t;
    } else {  // sequential cross attention (thdroformer.py:244-245)
      rdm_tf_proj_job p1[3] = {pj(f0, B, WQ, BQ, nullptr, q0, n0, 0), pj(f1, B, WK, BK, nullptr, k1, n1, ldk1), pj(f1, B, WV, BV, nullptr, v1, n1, 0)};
      RDM_TRY(rdm_tf_project(p1, 3, st));
      rdm_tf_attn_job a1[1] = {aj(q0, k1, v1, f0, B, g0, n0, n1, ldk1)};
      RDM_TRY(rdm_tf_attend(a1, 1, st));
      float* t = f0; f0 = g0; g0 = t;
      rdm_tf_proj_job p2[3] = {pj(f1, B, WQ, BQ, nullptr, q1, n1, 0), pj(f0, B, WK, BK, nullptr, k0, n0, ldk0), pj(f0, B, WV, BV, nullptr, v0, n0, 0)};
      RDM_TRY(rdm_tf_project(p2, 3, st));
      rdm_tf_attn_job a2[1] = {aj(q1, k0, v0, f1, B, g1, n1, n0, ldk0)};
      RDM_TRY(rdm_tf_attend(a2, 1, st));
      t = f1; f1 = g1; g1 = t;
    }
  }
  RDM_TRY(linear(a, f0, D, d.out_w, D, 1, d.out_b, o0, n0, d.c_out, D, st));
  RDM_TRY(linear(a, f1, D, d.out_w, D, 1, d.out_b, o1, n1, d.c_out, D, st));
  return RDM_OK;
}
}  // namespace

extern "C" size_t rdm_encoder_workspace(const rdm_block_desc* h_blocks, int num_blocks, const rdm_pyramid_desc* h_pyr, int groups) {
  Arena a(nullptr, 0, true);
  float* outs[8] = {nullptr};
  if (encoder_run(a, h_blocks, num_blocks, *h_pyr, groups, nullptr, outs, 0) != RDM_OK) return 0;
  return a.peak + 4096;
}

extern "C" int rdm_encoder_forward(const rdm_block_desc* h_blocks, int num_blocks, const rdm_pyramid_desc* h_pyr, int groups,
                                   const float* in_feats, float* const* h_out_feats, void* workspace, size_t workspace_bytes,
                                   cudaStream_t stream) {
  RDM_CHECK_ARG(h_blocks && h_pyr && h_out_feats && num_blocks >= 1 && h_pyr->num_stages >= 1 && h_pyr->num_stages <= 8,
                "rdm_encoder_forward: bad arguments");
  const size_t need = rdm_encoder_workspace(h_blocks, num_blocks, h_pyr, groups);
  if (need == 0 || workspace_bytes < need) {
    rdm_set_error("rdm_encoder_forward: workspace too small (%zu needed)", need);
    return RDM_ERR_WORKSPACE;
  }
  Arena a(workspace, workspace_bytes, false);
  return encoder_run(a, h_blocks, num_blocks, *h_pyr, groups, in_feats, h_out_feats, stream);
}

extern "C" size_t rdm_decoder_workspace(const rdm_unary_desc* h_dec, int num, const rdm_pyramid_desc* h_pyr, int top_stage, int groups) {
  Arena a(nullptr, 0, true);
  const float* skips[8] = {nullptr};
  int cc = h_dec[0].c_in;  // dry run: channel split does not change the sizes
  if (decoder_run(a, h_dec, num, *h_pyr, top_stage, groups, nullptr, cc, skips, nullptr, 0, 0) != RDM_OK) return 0;
  return a.peak + 4096;
}

extern "C" int rdm_decoder_forward(const rdm_unary_desc* h_dec, int num, const rdm_pyramid_desc* h_pyr, int top_stage, int groups,
                                   const float* coarse, int c_coarse, const float* const* h_skips, float* out, int ld_out,
                                   void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  RDM_CHECK_ARG(h_dec && h_pyr && h_skips && num >= 1 && num <= 7, "rdm_decoder_forward: bad arguments");
  Arena a(workspace, workspace_bytes, false);
  // size check first: the arena hands out pointers past the end otherwise
  {
    Arena d(nullptr, 0, true);
    int rc = decoder_run(d, h_dec, num, *h_pyr, top_stage, groups, nullptr, c_coarse, h_skips, nullptr, 0, 0);
    if (rc != RDM_OK) return rc;
    if (d.peak > workspace_bytes) {
      rdm_set_error("rdm_decoder_forward: workspace too small (%zu needed)", d.peak);
      return RDM_ERR_WORKSPACE;
    }
  }
  return decoder_run(a, h_dec, num, *h_pyr, top_stage, groups, coarse, c_coarse, h_skips, out, ld_out, stream);
}

extern "C" size_t rdm_thdroformer_workspace(int n_ref, int n_src, int c_out) {
  Arena a(nullptr, 0, true);
  rdm_thdroformer_desc d = {};
  d.c_in = 128;  // the scratch does not depend on c_in
  d.c_out = c_out;
  thdroformer_run(a, d, nullptr, n_ref, nullptr, n_src, nullptr, 0, nullptr, 0, nullptr, nullptr, 0);
  return a.peak + 4096;
}

extern "C" int rdm_thdroformer_forward(const rdm_thdroformer_desc* h_desc, const float* ref_points, int n_ref,
                                       const float* src_points, int n_src, const float* ref_feats, int ld_ref,
                                       const float* src_feats, int ld_src, float* out_ref, float* out_src, void* workspace,
                                       size_t workspace_bytes, cudaStream_t stream) {
  RDM_CHECK_ARG(h_desc && n_ref >= 1 && n_src >= 1 && h_desc->num_layers >= 0 && h_desc->num_layers <= 32,
                "rdm_thdroformer_forward: bad arguments");
  if (workspace_bytes < rdm_thdroformer_workspace(n_ref, n_src, h_desc->c_out)) {
    rdm_set_error("rdm_thdroformer_forward: workspace too small");
    return RDM_ERR_WORKSPACE;
  }
  Arena a(workspace, workspace_bytes, false);
  return thdroformer_run(a, *h_desc, ref_points, n_ref, src_points, n_src, ref_feats, ld_ref, src_feats, ld_src, out_ref, out_src,
                         stream);
}

// ------------------------------------------------------------------------------------------------ pyramid builder
int rdm_radius_search_impl(const float* q_points, const float* s_points, const int64_t* q_lengths, const int64_t* s_lengths,
                           int batch, int64_t nq_cap, int64_t ns_cap, int64_t ns_total_pad, float radius, int limit,
                           void* out_indices, int index_bytes, int* out_counts, int* out_max_count, int* out_order,
                           void* workspace, size_t workspace_bytes, cudaStream_t stream);

namespace {
// worst case: every stage keeps all n0 points
size_t pyramid_bytes(int64_t n0, const rdm_pyramid_cfg& c) {
  const size_t n = (size_t)n0;
  size_t b = 4096;
  b += align_up(sizeof(int64_t) * c.num_stages * c.batch, 256);
  b += align_up(sizeof(int) * 64, 256);  // max counts
  for (int s = 1; s < c.num_stages; s++) b += align_up(n * 12, 256);
  for (int s = 0; s < c.num_stages; s++) {
    b += 2 * align_up(n * 4, 256);                   // order (cell-sorted) + order by load
    b += align_up(n * 4 * (size_t)c.limits[s], 256);  // neighbors
    if (s + 1 < c.num_stages) {
      b += align_up(n * 4 * (size_t)c.limits[s], 256);  // subsampling
      b += align_up(n * 4 * (size_t)(c.up_nearest_only ? 1 : c.limits[s + 1]), 256);
    }
  }
  return b;
}
size_t pyramid_ws(int64_t n0, const rdm_pyramid_cfg& c) {
  size_t a = rdm_grid_subsample_workspace(n0, c.batch), b = rdm_radius_search_workspace(n0, c.batch);
  return (a > b ? a : b) + 4096;
}
}  // namespace

extern "C" size_t rdm_build_pyramid_bytes(int64_t n0, const rdm_pyramid_cfg* c) { return pyramid_bytes(n0, *c); }
extern "C" size_t rdm_build_pyramid_workspace(int64_t n0, const rdm_pyramid_cfg* c) { return pyramid_ws(n0, *c); }

// Two-phase form (used by the pair pipeline: the subsampling chain of pair i+1 runs on a side stream while pair i is in
// the network): begin = chained subsamplings + asynchronous readback of the stage sizes + event; finish = wait for that
// event (host), then every radius search at its exact size. rdm_build_pyramid = begin + finish.
struct PyramidJob {
  cudaEvent_t ev = nullptr;
  int64_t* pinned = nullptr;  // [8 * 16]
  const float* points = nullptr;
  int64_t n0 = 0;
  rdm_pyramid_cfg cfg;
  void *out_buf = nullptr, *ws = nullptr;
  size_t out_bytes = 0, ws_bytes = 0;
  int64_t* d_len = nullptr;
  int* d_maxc = nullptr;
  float* pts[8];
  size_t out_off = 0;
  bool pending = false;
};

extern "C" void* rdm_pyramid_job_create(void) {
  PyramidJob* j = new PyramidJob();
  if (cudaEventCreateWithFlags(&j->ev, cudaEventDisableTiming) != cudaSuccess ||
      cudaHostAlloc((void**)&j->pinned, sizeof(int64_t) * 8 * 16, cudaHostAllocDefault) != cudaSuccess) {
    rdm_set_error("rdm_pyramid_job_create: CUDA allocation failed");
    delete j;
    return nullptr;
  }
  return j;
}

extern "C" void rdm_pyramid_job_destroy(void* job) {
  PyramidJob* j = (PyramidJob*)job;
  if (j == nullptr) return;
  if (j->ev) cudaEventDestroy(j->ev);
  if (j->pinned) cudaFreeHost(j->pinned);
  delete j;
}

extern "C" int rdm_build_pyramid_begin(void* job, const float* points, const int64_t* lengths, int64_t n0,
                                       const rdm_pyramid_cfg* h_cfg, void* out_buf, size_t out_bytes, void* workspace,
                                       size_t workspace_bytes, cudaStream_t stream) {
  PyramidJob* j = (PyramidJob*)job;
  RDM_CHECK_ARG(j && h_cfg && points && lengths, "rdm_build_pyramid_begin: null argument");
  const rdm_pyramid_cfg& c = *h_cfg;
  RDM_CHECK_ARG(c.num_stages >= 1 && c.num_stages <= 8 && c.batch >= 1 && c.batch <= 16 && n0 >= 1 && n0 < (1LL << 28),
                "rdm_build_pyramid: bad configuration");
  if (out_bytes < pyramid_bytes(n0, c) || workspace_bytes < pyramid_ws(n0, c)) {
    rdm_set_error("rdm_build_pyramid: output buffer or workspace too small");
    return RDM_ERR_WORKSPACE;
  }
  j->points = points; j->n0 = n0; j->cfg = c; j->out_buf = out_buf; j->out_bytes = out_bytes; j->ws = workspace;
  j->ws_bytes = workspace_bytes;
  Workspace out(out_buf, out_bytes);
  const int S = c.num_stages, B = c.batch;
  j->d_len = out.get<int64_t>((size_t)S * B);  // stage-major
  j->d_maxc = out.get<int>(64);
  j->pts[0] = const_cast<float*>(points);
  for (int s = 1; s < S; s++) j->pts[s] = out.get<float>((size_t)n0 * 3);
  j->out_off = out.off;
  // ---- chained subsampling: capacity launches driven by the device-side lengths, no host round trip in between
  RDM_CUDA(cudaMemcpyAsync(j->d_len, lengths, sizeof(int64_t) * B, cudaMemcpyDeviceToDevice, stream));
  float voxel = c.first_voxel;
  for (int s = 1; s < S; s++) {
    RDM_TRY(rdm_grid_subsample(j->pts[s - 1], j->d_len + (size_t)(s - 1) * B, B, n0, voxel, j->pts[s], j->d_len + (size_t)s * B,
                               workspace, workspace_bytes, stream));
    voxel *= 2.f;
  }
  RDM_CUDA(cudaMemcpyAsync(j->pinned, j->d_len, sizeof(int64_t) * S * B, cudaMemcpyDeviceToHost, stream));
  RDM_CUDA(cudaEventRecord(j->ev, stream));
  j->pending = true;
  return RDM_OK;
}

extern "C" int rdm_build_pyramid_finish(void* job, rdm_pyramid_desc* h_desc, int64_t* h_lengths, const int64_t** h_d_lengths,
                                        cudaStream_t stream) {
  PyramidJob* j = (PyramidJob*)job;
  RDM_CHECK_ARG(j && j->pending && h_desc && h_lengths && h_d_lengths, "rdm_build_pyramid_finish: no pyramid in flight");
  j->pending = false;
  RDM_CUDA(cudaEventSynchronize(j->ev));  // the one synchronisation: stage sizes decide every later shape
  const rdm_pyramid_cfg& c = j->cfg;
  const int S = c.num_stages, B = c.batch;
  const int64_t n0 = j->n0;
  Workspace out(j->out_buf, j->out_bytes);
  out.off = j->out_off;
  void* workspace = j->ws;
  const size_t workspace_bytes = j->ws_bytes;
  float** pts = j->pts;
  int64_t* d_len = j->d_len;
  int* d_maxc = j->d_maxc;
  rdm_pyramid_desc& d = *h_desc;
  memset(&d, 0, sizeof(d));
  d.num_stages = S;
  d.index_bytes = 4;
  for (int s = 0; s < S; s++) {
    int64_t tot = 0;
    for (int b = 0; b < B; b++) {
      h_lengths[s * B + b] = j->pinned[s * B + b];
      tot += j->pinned[s * B + b];
    }
    RDM_CHECK_ARG(tot >= 0 && tot <= n0, "rdm_build_pyramid: inconsistent stage size");
    d.n[s] = (int)tot;
    d.points[s] = pts[s];
    h_d_lengths[s] = d_len + (size_t)s * B;
  }
  // ---- radius searches at exact sizes (utils/data.py:35-67)
  float radius = c.first_radius;
  int nsearch = 0;
  for (int s = 0; s < S; s++) {
    const int n = d.n[s];
    int* order = out.get<int>((size_t)(n > 0 ? n : 1));
    int* nb = out.get<int>((size_t)(n > 0 ? n : 1) * c.limits[s]);
    d.order[s] = order;
    d.neighbors[s] = nb;
    d.nb_width[s] = c.limits[s];
    RDM_TRY(rdm_radius_search_impl(pts[s], pts[s], h_d_lengths[s], h_d_lengths[s], B, n, n, n, radius, c.limits[s], nb, 4, nullptr,
                                   d_maxc + nsearch++, order, workspace, workspace_bytes, stream));
    {
      // heaviest-first walk order for the KPConv gathers: the cell-sorted order stably partitioned by neighbourhood fill,
      // so that the last (partial) wave of gather CTAs holds the light queries (RDM_GATHER_HEAVY_FIRST=0 keeps cell order)
      int* order2 = out.get<int>((size_t)(n > 0 ? n : 1));
      static int heavy = -1;
      if (heavy < 0) {
        const char* e = getenv("RDM_GATHER_HEAVY_FIRST");
        heavy = (e && e[0] == '0') ? 0 : 1;
      }
      if (heavy && n > 0 && out.ok) {
        RDM_TRY(rdm_order_by_load(order, nb, n, c.limits[s], n, order2, workspace, workspace_bytes, stream));
        d.order[s] = order2;
      }
    }
    if (s + 1 < S) {
      const int m = d.n[s + 1];
      int* sub = out.get<int>((size_t)(m > 0 ? m : 1) * c.limits[s]);
      d.subsampling[s] = sub;
      d.sub_width[s] = c.limits[s];
      RDM_TRY(rdm_radius_search_impl(pts[s + 1], pts[s], h_d_lengths[s + 1], h_d_lengths[s], B, m, n, n, radius, c.limits[s], sub, 4,
                                     nullptr, d_maxc + nsearch, nullptr, workspace, workspace_bytes, stream));
      RDM_TRY(rdm_mark_reference_width(sub, m, c.limits[s], n, d_maxc + nsearch, stream));  // the max-pool's view of the width
      nsearch++;
      if (!(c.skip_up0 && s == 0)) {
        const int w = c.up_nearest_only ? 1 : c.limits[s + 1];
        int* up = out.get<int>((size_t)(n > 0 ? n : 1) * w);
        d.upsampling[s] = up;
        d.up_width[s] = w;
        RDM_TRY(rdm_radius_search_impl(pts[s], pts[s + 1], h_d_lengths[s], h_d_lengths[s + 1], B, n, m, m, radius * 2.f, w, up, 4,
                                       nullptr, d_maxc + nsearch++, nullptr, workspace, workspace_bytes, stream));
      }
    }
    radius *= 2.f;
  }
  if (!out.ok) {
    rdm_set_error("rdm_build_pyramid: output buffer too small");
    return RDM_ERR_WORKSPACE;
  }
  return RDM_OK;
}

extern "C" int rdm_build_pyramid(const float* points, const int64_t* lengths, int64_t n0, const rdm_pyramid_cfg* h_cfg,
                                 void* out_buf, size_t out_bytes, void* workspace, size_t workspace_bytes,
                                 rdm_pyramid_desc* h_desc, int64_t* h_lengths, const int64_t** h_d_lengths,
                                 cudaStream_t stream) {
  static void* job = nullptr;  // calls are serialised by the caller (GIL)
  if (job == nullptr) job = rdm_pyramid_job_create();
  if (job == nullptr) return RDM_ERR_CUDA;
  RDM_TRY(rdm_build_pyramid_begin(job, points, lengths, n0, h_cfg, out_buf, out_bytes, workspace, workspace_bytes, stream));
  return rdm_build_pyramid_finish(job, h_desc, h_lengths, h_d_lengths, stream);
}

// ------------------------------------------------------------------------------------------------ backbone runner
namespace {
// Optional marker recorded right after the encoder inside rdm_backbone_forward (rdm_backbone_set_encoder_event): the pair
// pipeline holds the PREVIOUS pair's matching tail and the NEXT pair's radius searches behind it, so that the KPConv gathers
// have the machine to themselves and the latency-bound transformer / decoder / matching kernels of two pairs share it.
cudaEvent_t g_encoder_done = nullptr;

int backbone_run(Arena& a, const rdm_backbone_desc& d, const rdm_pyramid_desc& p, int nc_ref, const float* in_feats,
                 const rdm_backbone_out& o, cudaStream_t st) {
  const int S = p.num_stages, top = S - 1;
  const int nc = p.n[top], c = d.h_transformer1->c_out;
  RDM_CHECK_ARG(nc_ref >= 1 && nc_ref < nc, "rdm_backbone_forward: both clouds need at least one coarse node");
  float* enc_out[8] = {nullptr};
  int last_c[8] = {0};
  for (int i = 0; i < d.num_blocks; i++) last_c[d.h_blocks[i].stage] = d.h_blocks[i].c_out;
  for (int s = 0; s < S; s++) enc_out[s] = a.f((size_t)p.n[s] * last_c[s]);
  RDM_TRY(encoder_run(a, d.h_blocks, d.num_blocks, p, d.groups, in_feats, enc_out, st));
  if (!a.dry && g_encoder_done != nullptr) RDM_CUDA(cudaEventRecord(g_encoder_done, st));
  // first ThDRoFormer on the coarsest stage (model.py:154-159), n2p score head (:160-167)
  const float* pc = p.points[top];
  RDM_TRY(thdroformer_run(a, *d.h_transformer1, pc, nc_ref, pc + 3 * (size_t)nc_ref, nc - nc_ref, enc_out[top], last_c[top],
                          enc_out[top] + (size_t)nc_ref * last_c[top], last_c[top], o.feats_c, o.feats_c + (size_t)nc_ref * c, st));
  float* logit = a.f((size_t)nc);
  float* cat = a.f((size_t)nc * (c + 1));
  RDM_TRY(linear(a, o.feats_c, c, d.n2p_w, c, 1, d.n2p_b, logit, nc, 1, c, st));
  if (!a.dry) {
    RDM_TRY(rdm_activation(logit, o.n2p_scores, nc, 3, 0.f, st));
    RDM_TRY(rdm_append_column(o.feats_c, logit, nc, c, cat, st));
  }
  // decoder (backbone.py:118-151) on [tf | n2p logit] and the encoder skips; last column = p2p logit (model.py:168-174)
  const float* skips[8] = {nullptr};
  for (int i = 0; i < d.num_dec; i++) skips[i] = enc_out[top - 1 - i];
  RDM_TRY(decoder_run(a, d.h_dec, d.num_dec, p, top, d.groups, cat, c + 1, skips, o.feats_f, o.ld_feats_f, st));
  if (!a.dry) {
    const int s_out = top - d.num_dec, c_out = d.h_dec[d.num_dec - 1].c_out;
    RDM_TRY(rdm_sigmoid_column(o.feats_f + (c_out - 1), o.ld_feats_f, p.n[s_out], o.p2p_scores, st));
  }
  return RDM_OK;
}
}  // namespace

extern "C" int rdm_backbone_set_encoder_event(void* cuda_event) {
  g_encoder_done = (cudaEvent_t)cuda_event;
  return RDM_OK;
}

extern "C" size_t rdm_backbone_workspace(const rdm_backbone_desc* h_desc, const rdm_pyramid_desc* h_pyr, int nc_ref) {
  Arena a(nullptr, 0, true);
  rdm_backbone_out o = {};
  if (backbone_run(a, *h_desc, *h_pyr, nc_ref, nullptr, o, 0) != RDM_OK) return 0;
  return a.peak + 4096;
}

extern "C" int rdm_backbone_forward(const rdm_backbone_desc* h_desc, const rdm_pyramid_desc* h_pyr, int nc_ref,
                                    const float* in_feats, const rdm_backbone_out* h_out, void* workspace, size_t workspace_bytes,
                                    cudaStream_t stream) {
  RDM_CHECK_ARG(h_desc && h_pyr && h_out && h_desc->h_blocks && h_desc->h_transformer1 && h_desc->h_dec && h_desc->num_dec >= 1,
                "rdm_backbone_forward: null argument");
  const size_t need = rdm_backbone_workspace(h_desc, h_pyr, nc_ref);
  if (need == 0 || workspace_bytes < need) {
    rdm_set_error("rdm_backbone_forward: workspace too small (%zu needed)", need);
    return RDM_ERR_WORKSPACE;
  }
  Arena a(workspace, workspace_bytes, false);
  return backbone_run(a, *h_desc, *h_pyr, nc_ref, in_feats, *h_out, stream);
}

// ------------------------------------------------------------------------------------------------ matching tail
namespace {
struct MatchPinned {
  int counts[2];
  int coarse_count;
  int meta[4];
  float T[16];
};


// phase 1: vote + NMS (sizes known). Returns through `io` buffers; sel counts land in pinned memory.
int match_phase1(Arena& a, const rdm_match_desc& d, const rdm_match_io& io, int* d_counts, cudaStream_t st) {
  const int nc = io.nc, c = d.c;
  size_t mark = a.off;
  float* h0 = a.f((size_t)nc * d.h0);
  float* h1 = a.f((size_t)nc * d.h1);
  float* off = a.f((size_t)nc * (3 + c));
  float* n2n_logit = a.f((size_t)nc);
  RDM_TRY(linear(a, io.feats_c, c, d.v_w0, c, 1, d.v_b0, h0, nc, d.h0, c, st));
  if (!a.dry) RDM_TRY(rdm_layernorm(h0, nullptr, d.v_g0, d.v_e0, h0, nc, d.h0, 1e-5f, 2, st));
  RDM_TRY(linear(a, h0, d.h0, d.v_w1, d.h0, 1, d.v_b1, h1, nc, d.h1, d.h0, st));
  if (!a.dry) RDM_TRY(rdm_layernorm(h1, nullptr, d.v_g1, d.v_e1, h1, nc, d.h1, 1e-5f, 2, st));
  RDM_TRY(linear(a, h1, d.h1, d.v_wr, d.h1, 1, d.v_br, off, nc, 3 + c, d.h1, st));
  if (!a.dry)
    RDM_TRY(rdm_vote_finish(off, 3 + c, io.points_c, io.feats_c, c, d.v_go, d.v_eo, d.max_offset, 1e-5f, nc, c, io.shifted_points,
                            io.vote_feats, st));
  // n2n score head (model.py:209-212): sigmoid + clamp
  RDM_TRY(linear(a, io.vote_feats, c, d.n2n_w, c, 1, d.n2n_b, n2n_logit, nc, 1, c, st));
  if (!a.dry) RDM_TRY(rdm_activation(n2n_logit, io.n2n_scores, nc, 3, 0.f, st));
  // NMS (vote.py:14-40): radius search on the shifted nodes, greedy rule, compaction
  int* table = (int*)a.raw((size_t)nc * d.nms_limit * sizeof(int));
  int* maxc = (int*)a.raw(sizeof(int));
  const size_t wsb = rdm_radius_search_workspace(nc, 2);
  void* ws = a.raw(wsb);
  if (!a.dry) {
    RDM_TRY(rdm_radius_search_impl(io.shifted_points, io.shifted_points, io.lengths_c, io.lengths_c, 2, nc, nc, nc, d.nms_radius,
                                   d.nms_limit, table, 4, nullptr, maxc, nullptr, ws, wsb, st));
    RDM_TRY(rdm_nms(table, 4, nc, d.nms_limit, io.nc_ref, io.nms_mask, io.selected, d_counts, st));
  }
  a.off = mark;
  return RDM_OK;
}

// patch scores + Sinkhorn + LGR on the first P coarse correspondences (model.py:323-361)
int match_patches(Arena& a, const rdm_match_desc& d, const rdm_match_io& io, int n0, int P, int* d_meta, cudaStream_t st) {
  const int c = d.c, K = d.point_limit;
  size_t mark = a.off;
  float* scores = a.f((size_t)P * K * K);
  const size_t wsb = rdm_lgr_workspace(P, K);
  void* ws = a.raw(wsb);
  if (!a.dry) {
    const float* ff_ref = io.feats_f;
    const float* ff_src = io.feats_f + (size_t)io.nf_ref * io.ld_feats_f;
    const int64_t* knn_src = io.knn_indices + (size_t)n0 * K;
    const unsigned char* km_src = io.knn_masks + (size_t)n0 * K;
    RDM_TRY(rdm_patch_scores(ff_ref, io.nf_ref, ff_src, io.nf - io.nf_ref, c, io.ld_feats_f, io.knn_indices, knn_src, io.corr_ref,
                             io.corr_src, P, K, 1.0f / sqrtf((float)c), scores, st));
    RDM_TRY(rdm_sinkhorn(scores, P, K, K, io.knn_masks, km_src, io.corr_ref, io.corr_src, d.ot_alpha, d.sinkhorn_iterations,
                         d.sinkhorn_inf, io.matching_scores, st));
    RDM_TRY(rdm_lgr(io.matching_scores, P, K, io.points_f, io.points_f + 3 * (size_t)io.nf_ref, io.knn_indices, knn_src,
                    io.knn_masks, km_src, io.corr_ref, io.corr_src, d.acceptance_radius, d.correspondence_threshold,
                    d.refinement_steps, io.ref_corr_points, io.src_corr_points, io.corr_scores, io.corr_bij, io.transform, d_meta, ws,
                    wsb, st));
  }
  a.off = mark;
  return RDM_OK;
}

// phase 2: everything after the node selection, for n0 ref / n1 src survivors (capacities when dry)
cudaEvent_t g_patch_wait = nullptr;
// fork / join (optional): the two point-to-node partitions only need the selected node coordinates, so they run on the library's
// side stream next to transformer 2 instead of behind it (~0.2 ms of a 1.5 ms tail)
int match_phase2(Arena& a, const rdm_match_desc& d, const rdm_match_io& io, int n0, int n1, int* d_coarse_count, int* d_meta,
                 cudaStream_t st, cudaEvent_t fork = nullptr, cudaEvent_t join = nullptr) {
  const int c = d.c, K = d.point_limit, P = d.num_correspondences;
  const int ns = n0 + n1;
  size_t mark = a.off;
  float* sel_feats = a.f((size_t)ns * c);
  float* t2 = a.f((size_t)ns * c);
  if (!a.dry) {
    const float* src[4] = {io.shifted_points, io.vote_feats, io.n2p_scores, io.n2n_scores};
    float* dst[4] = {io.sel_points, sel_feats, io.sel_n2p, io.sel_n2n};
    const int cc[4] = {3, c, 1, 1}, ld[4] = {3, c, 1, 1};
    RDM_TRY(rdm_gather_rows(src, dst, cc, ld, 4, io.selected, ns, st));
  }
  const bool dual = !a.dry && fork != nullptr && join != nullptr && g_side.ready();
  // point_to_node_partition x2 (model.py:267-272): needs sel_points only
  {
    int* p2n = (int*)a.raw(sizeof(int) * (size_t)io.nf);
    const size_t w0 = rdm_point_to_node_workspace(io.nf_ref, n0 > 0 ? n0 : 1), w1 = rdm_point_to_node_workspace(io.nf - io.nf_ref, n1 > 0 ? n1 : 1);
    void* ws0 = a.raw(w0);
    void* ws1 = a.raw(w1);
    if (!a.dry) {
      cudaStream_t ps = st;
      if (dual) {
        RDM_CUDA(cudaEventRecord(fork, st));
        RDM_CUDA(cudaStreamWaitEvent(g_side.stream, fork, 0));
        ps = g_side.stream;
      }
      RDM_TRY(rdm_point_to_node(io.points_f, io.nf_ref, io.sel_points, n0, K, p2n, io.node_masks, io.knn_indices, io.knn_masks, ws0,
                                w0, ps));
      RDM_TRY(rdm_point_to_node(io.points_f + 3 * (size_t)io.nf_ref, io.nf - io.nf_ref, io.sel_points + 3 * (size_t)n0, n1, K,
                                p2n + io.nf_ref, io.node_masks + n0, io.knn_indices + (size_t)n0 * K, io.knn_masks + (size_t)n0 * K,
                                ws1, w1, ps));
      if (dual) RDM_CUDA(cudaEventRecord(join, ps));
    }
  }
  {
    const size_t wsb = rdm_thdroformer_workspace(n0 > 0 ? n0 : 1, n1 > 0 ? n1 : 1, c);
    void* ws = a.raw(wsb);
    if (!a.dry)
      RDM_TRY(rdm_thdroformer_forward(d.h_transformer2, io.sel_points, n0, io.sel_points + 3 * (size_t)n0, n1, sel_feats, c,
                                      sel_feats + (size_t)n0 * c, c, t2, t2 + (size_t)n0 * c, ws, wsb, st));
  }
  if (!a.dry) RDM_TRY(rdm_l2_normalize(t2, io.sel_feats_norm, ns, c, st));
  // SuperPointMatching (model.py:308-311)
  {
    float* xy = a.f((size_t)(n0 > 0 ? n0 : 1) * (n1 > 0 ? n1 : 1));
    float* sums = a.f((size_t)ns + 2);
    RDM_TRY(linear(a, io.sel_feats_norm, c, io.sel_feats_norm + (size_t)n0 * c, c, 1, nullptr, xy, n0, n1, c, st));
    if (dual) RDM_CUDA(cudaStreamWaitEvent(st, join, 0));  // node masks / knn tables from the side stream
    if (!a.dry) {  // slots beyond the produced count stay at pair (0, 0): in range for the speculative patch pass below
      RDM_CUDA(cudaMemsetAsync(io.corr_ref, 0, sizeof(int64_t) * P, st));
      RDM_CUDA(cudaMemsetAsync(io.corr_src, 0, sizeof(int64_t) * P, st));
    }
    if (!a.dry)
      RDM_TRY(rdm_coarse_matching(xy, n0, n1, io.node_masks, io.node_masks + n0, P, d.dual_normalization, io.corr_ref, io.corr_src,
                                  io.corr_node_scores, d_coarse_count, sums, st));
  }
  // optional (rdm_match_set_patch_wait_event): only the patch stage - patch scores, Sinkhorn, pose: the kernels with real
  // grids - is held behind the caller's event; what precedes it (transformer 2, partitions, coarse matching: <= 108 CTAs at a
  // time) may share the machine with the next pair's encoder
  if (!a.dry && g_patch_wait != nullptr) RDM_CUDA(cudaStreamWaitEvent(st, g_patch_wait, 0));
  RDM_TRY(match_patches(a, d, io, n0, P, d_meta, st));
  a.off = mark;
  return RDM_OK;
}

size_t match_ws(const rdm_match_desc& d, int nc, int nc_ref, int nf, int nf_ref) {
  Arena a(nullptr, 0, true);
  rdm_match_io io = {};
  io.nc = nc; io.nc_ref = nc_ref; io.nf = nf; io.nf_ref = nf_ref;
  a.raw(256);  // device counters
  if (match_phase1(a, d, io, nullptr, 0) != RDM_OK) return 0;
  if (match_phase2(a, d, io, nc_ref, nc - nc_ref, nullptr, nullptr, 0) != RDM_OK) return 0;
  return a.peak + 4096;
}
}  // namespace

extern "C" size_t rdm_match_workspace(const rdm_match_desc* h_desc, int nc, int nc_ref, int nf, int nf_ref) {
  return match_ws(*h_desc, nc, nc_ref, nf, nf_ref);
}

// ---- the matching tail as a job in three host calls, for callers that keep several pairs in flight (PairPipeline):
//   rdm_match_begin     phase 1 (vote, NMS) + asynchronous read-back of the survivor counts            - never blocks
//   rdm_match_continue  waits for those counts (they fix every shape below), queues phase 2 + read-back - blocks on phase 1 only
//   rdm_match_finish    waits for the result counts + pose; the rare exact-count redo of the patch stage
// rdm_match_forward is the three in a row.
namespace {
struct MatchJob {
  MatchPinned* pinned = nullptr;
  cudaEvent_t counts_ready = nullptr, result_ready = nullptr, fork = nullptr, join = nullptr;
  rdm_match_desc d;
  rdm_match_io io;
  void* workspace = nullptr;
  size_t workspace_bytes = 0, arena_off = 0;
  int* d_cnt = nullptr;
  cudaStream_t stream = nullptr;
  int n0 = 0, n1 = 0, stage = 0;  // stage: 0 idle, 1 begun, 2 continued
};
}  // namespace

extern "C" void* rdm_match_job_create(void) {
  MatchJob* j = new MatchJob();
  if (cudaHostAlloc((void**)&j->pinned, sizeof(MatchPinned), cudaHostAllocDefault) != cudaSuccess ||
      cudaEventCreateWithFlags(&j->counts_ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&j->result_ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&j->fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&j->join, cudaEventDisableTiming) != cudaSuccess) {
    rdm_set_error("rdm_match_job_create: CUDA resource allocation failed");
    delete j;
    return nullptr;
  }
  return j;
}

extern "C" void rdm_match_job_destroy(void* job) {
  MatchJob* j = (MatchJob*)job;
  if (j == nullptr) return;
  if (j->pinned) cudaFreeHost(j->pinned);
  if (j->counts_ready) cudaEventDestroy(j->counts_ready);
  if (j->result_ready) cudaEventDestroy(j->result_ready);
  if (j->fork) cudaEventDestroy(j->fork);
  if (j->join) cudaEventDestroy(j->join);
  delete j;
}

extern "C" int rdm_match_set_patch_wait_event(void* cuda_event) {
  g_patch_wait = (cudaEvent_t)cuda_event;
  return RDM_OK;
}

// abandons whatever the job has in flight (waits for its stream first): for callers that stop consuming a pipeline half-way
extern "C" int rdm_match_job_reset(void* job) {
  MatchJob* j = (MatchJob*)job;
  RDM_CHECK_ARG(j != nullptr, "rdm_match_job_reset: null job");
  if (j->stage != 0) RDM_CUDA(cudaStreamSynchronize(j->stream));
  j->stage = 0;
  return RDM_OK;
}

extern "C" int rdm_match_begin(void* job, const rdm_match_desc* h_desc, const rdm_match_io* h_io, void* workspace, size_t workspace_bytes,
                               cudaStream_t stream) {
  MatchJob* j = (MatchJob*)job;
  RDM_CHECK_ARG(j && h_desc && h_io && h_desc->h_transformer2, "rdm_match_begin: null argument");
  RDM_CHECK_ARG(j->stage == 0, "rdm_match_begin: the job is still in flight (finish it first)");
  j->d = *h_desc;
  j->io = *h_io;
  const rdm_match_desc& d = j->d;
  const rdm_match_io& io = j->io;
  RDM_CHECK_ARG(io.nc >= 1 && io.nc_ref >= 0 && io.nc_ref <= io.nc && io.nf >= 1 && io.nf_ref >= 0 && io.nf_ref <= io.nf &&
                    d.point_limit == 128 && d.num_correspondences >= 1 && d.num_correspondences <= 1024 && d.c >= 16,
                "rdm_match_forward: bad sizes");
  if (workspace_bytes < match_ws(d, io.nc, io.nc_ref, io.nf, io.nf_ref)) {
    rdm_set_error("rdm_match_forward: workspace too small");
    return RDM_ERR_WORKSPACE;
  }
  j->workspace = workspace;
  j->workspace_bytes = workspace_bytes;
  j->stream = stream;
  Arena a(workspace, workspace_bytes, false);
  j->d_cnt = (int*)a.raw(256);  // [0,1] NMS counts, [2] coarse count, [4..7] LGR meta
  RDM_TRY(match_phase1(a, d, io, j->d_cnt, stream));
  j->arena_off = a.off;
  RDM_CUDA(cudaMemcpyAsync(j->pinned->counts, j->d_cnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
  RDM_CUDA(cudaEventRecord(j->counts_ready, stream));
  j->stage = 1;
  return RDM_OK;
}

extern "C" int rdm_match_continue(void* job, rdm_match_result* h_result) {
  MatchJob* j = (MatchJob*)job;
  RDM_CHECK_ARG(j && h_result && j->stage == 1, "rdm_match_continue: call rdm_match_begin first");
  const rdm_match_desc& d = j->d;
  const rdm_match_io& io = j->io;
  j->stage = 0;  // an error below leaves the job reusable
  RDM_CUDA(cudaEventSynchronize(j->counts_ready));  // sync 1: survivor counts = the shapes of everything below
  const int n0 = j->pinned->counts[0], n1 = j->pinned->counts[1];
  RDM_CHECK_ARG(n0 >= 0 && n1 >= 0 && n0 <= io.nc_ref && n1 <= io.nc - io.nc_ref, "rdm_match_forward: inconsistent NMS counts");
  h_result->n_ref_sel = n0;
  h_result->n_src_sel = n1;
  h_result->num_patches = 0;
  h_result->num_corr = 0;
  RDM_CHECK_ARG(n0 >= 1 && n1 >= 1, "rdm_match_forward: no superpoint survived NMS in one of the clouds");
  j->n0 = n0;
  j->n1 = n1;
  Arena a(j->workspace, j->workspace_bytes, false);
  a.off = j->arena_off;
  cudaStream_t stream = j->stream;
  RDM_TRY(match_phase2(a, d, io, n0, n1, j->d_cnt + 2, j->d_cnt + 4, stream, j->fork, j->join));
  j->arena_off = a.off;
  RDM_CUDA(cudaMemcpyAsync(&j->pinned->coarse_count, j->d_cnt + 2, sizeof(int), cudaMemcpyDeviceToHost, stream));
  RDM_CUDA(cudaMemcpyAsync(j->pinned->meta, j->d_cnt + 4, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream));
  RDM_CUDA(cudaMemcpyAsync(j->pinned->T, io.transform, 16 * sizeof(float), cudaMemcpyDeviceToHost, stream));
  RDM_CUDA(cudaEventRecord(j->result_ready, stream));
  j->stage = 2;
  return RDM_OK;
}

extern "C" int rdm_match_finish(void* job, rdm_match_result* h_result) {
  MatchJob* j = (MatchJob*)job;
  RDM_CHECK_ARG(j && h_result && j->stage == 2, "rdm_match_finish: call rdm_match_continue first");
  const rdm_match_desc& d = j->d;
  const rdm_match_io& io = j->io;
  j->stage = 0;
  RDM_CUDA(cudaEventSynchronize(j->result_ready));  // sync 2: result counts + pose
  if (j->pinned->coarse_count < d.num_correspondences) {
    // fewer valid node pairs than requested (tiny clouds): the speculative pass saw padding patches; redo it at the
    // exact count. superpoint_matching.py:52-53 takes min(num_correspondences, #pairs).
    const int P = j->pinned->coarse_count;
    RDM_CHECK_ARG(P >= 1, "rdm_match_forward: no coarse correspondence");
    Arena a(j->workspace, j->workspace_bytes, false);
    a.off = j->arena_off;
    cudaStream_t stream = j->stream;
    RDM_TRY(match_patches(a, d, io, j->n0, P, j->d_cnt + 4, stream));
    RDM_CUDA(cudaMemcpyAsync(j->pinned->meta, j->d_cnt + 4, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    RDM_CUDA(cudaMemcpyAsync(j->pinned->T, io.transform, 16 * sizeof(float), cudaMemcpyDeviceToHost, stream));
    RDM_CUDA(cudaStreamSynchronize(stream));
  }
  h_result->n_ref_sel = j->n0;
  h_result->n_src_sel = j->n1;
  h_result->num_patches = j->pinned->coarse_count;
  h_result->num_corr = j->pinned->meta[0];
  memcpy(h_result->transform, j->pinned->T, sizeof(float) * 16);
  return RDM_OK;
}

extern "C" int rdm_match_forward(const rdm_match_desc* h_desc, const rdm_match_io* h_io, rdm_match_result* h_result, void* workspace,
                                 size_t workspace_bytes, cudaStream_t stream) {
  static void* job = nullptr;  // one caller thread per runner (rdm_sm100.h)
  RDM_CHECK_ARG(h_desc && h_io && h_result && h_desc->h_transformer2, "rdm_match_forward: null argument");
  if (job == nullptr) job = rdm_match_job_create();
  if (job == nullptr) return RDM_ERR_CUDA;
  ((MatchJob*)job)->stage = 0;  // a previous call that failed half-way leaves nothing in flight that matters here
  int rc = rdm_match_begin(job, h_desc, h_io, workspace, workspace_bytes, stream);
  if (rc != RDM_OK) return rc;
  rc = rdm_match_continue(job, h_result);
  if (rc != RDM_OK) return rc;
  return rdm_match_finish(job, h_result);
}
