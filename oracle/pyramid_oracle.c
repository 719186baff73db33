/*
 * TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file's library.
 *
 * Plain-C restatement of the two native functions of the reference's `rdmnet.ext` module:
 *   oracle_grid_subsample   <- geotransformer/extensions/cpu/grid_subsampling/grid_subsampling_cpu.cpp:3-75
 *   oracle_radius_neighbors <- geotransformer/extensions/cpu/radius_neighbors/radius_neighbors_cpu.cpp:3-91
 *                              (+ nanoflann.hpp:249-252 strict `dist < radius`, :432-440 L2_Simple accumulate order,
 *                               :1287 sort by distance)
 * Parity status: PINNED - tests/test_oracle_cpu.py checks both functions against the reference's own
 * C++ sources compiled in place (oracle/_ref/libref_ext.so, see oracle/Makefile) on the bundled-scan fixtures
 * and on seeded random clouds; grid_subsample is bit-exact incl. output order, radius_neighbors is equal
 * after canonicalising the order inside runs of exactly equal d2 (the reference's order inside such runs is
 * std::sort's unspecified tie order over KD-tree traversal order; ours is ascending index).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no -march: every fp32 op must round separately, as in the
 * reference's x86-64 baseline build).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* libstdc++ (g++ 13) std::unordered_map growth: rehash to BKT[k] fires when the element count reaches
 * BKT[k-1]+1 (BKT[-1] = 0). Regenerate with oracle/probe_buckets.cpp. */
static const int64_t BKT[] = {13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753, 42043, 85229, 172933,
                              351061, 712697, 1447153, 2938679, 5967347, 12117689, 24607243, 49969847};
#define NBKT ((int)(sizeof(BKT) / sizeof(BKT[0])))

typedef struct {
  uint64_t key;
  float sx, sy, sz;
  int count;
  int64_t next; /* singly linked list, -1 = end */
} node_t;

/* Model of libstdc++'s _Hashtable with unique keys and identity hash.
 * bucket[b] holds the index of the node BEFORE the first node of bucket b (-2 = the list's before_begin
 * sentinel, -1 = empty bucket). Insertion at bucket begin; empty bucket -> node goes to list head. */
typedef struct {
  node_t* nodes;
  int64_t n, cap;
  int64_t head; /* before_begin.next */
  int64_t* bucket;
  int64_t nb;
  int era;
} table_t;

static void place(table_t* t, int64_t* bucket, int64_t nb, int64_t id) {
  int64_t b = (int64_t)(t->nodes[id].key % (uint64_t)nb);
  if (bucket[b] == -1) {
    /* empty bucket: becomes the new list head; the bucket that owned the old head now hangs off this node */
    t->nodes[id].next = t->head;
    t->head = id;
    if (t->nodes[id].next >= 0) {
      int64_t ob = (int64_t)(t->nodes[t->nodes[id].next].key % (uint64_t)nb);
      bucket[ob] = id;
    }
    bucket[b] = -2;
  } else {
    int64_t before = bucket[b];
    if (before == -2) {
      t->nodes[id].next = t->head;
      t->head = id;
    } else {
      t->nodes[id].next = t->nodes[before].next;
      t->nodes[before].next = id;
    }
  }
}

static void rehash(table_t* t, int64_t nb) {
  int64_t* nbk = (int64_t*)malloc(sizeof(int64_t) * (size_t)nb);
  for (int64_t i = 0; i < nb; i++) nbk[i] = -1;
  int64_t p = t->head;
  t->head = -1;
  while (p >= 0) { /* re-insert in current iteration order (hashtable.h _M_rehash_aux, unique keys) */
    int64_t nx = t->nodes[p].next;
    place(t, nbk, nb, p);
    p = nx;
  }
  free(t->bucket);
  t->bucket = nbk;
  t->nb = nb;
}

static int64_t find(table_t* t, uint64_t key) {
  if (t->nb == 0) return -1;
  int64_t b = (int64_t)(key % (uint64_t)t->nb);
  int64_t before = t->bucket[b];
  if (before == -1) return -1;
  int64_t p = (before == -2) ? t->head : t->nodes[before].next;
  while (p >= 0 && (int64_t)(t->nodes[p].key % (uint64_t)t->nb) == b) {
    if (t->nodes[p].key == key) return p;
    p = t->nodes[p].next;
  }
  return -1;
}

/* One cloud. Returns number of output points written to out (capacity n). grid_subsampling_cpu.cpp:3-48 */
static int64_t subsample_one(const float* pts, int64_t n, float voxel, float* out) {
  float mnx = pts[0], mny = pts[1], mnz = pts[2], mxx = mnx, mxy = mny, mxz = mnz;
  for (int64_t i = 0; i < n; i++) { /* cloud.cpp:5-39 */
    const float* p = pts + 3 * i;
    if (p[0] < mnx) mnx = p[0];
    if (p[1] < mny) mny = p[1];
    if (p[2] < mnz) mnz = p[2];
    if (p[0] > mxx) mxx = p[0];
    if (p[1] > mxy) mxy = p[1];
    if (p[2] > mxz) mxz = p[2];
  }
  /* :11  floor(minCorner * (1. / voxel_size)) * voxel_size ; the double reciprocal is narrowed to float by
   * operator*(PointXYZ, const float) (cloud.h:84) */
  float inv = (float)(1.0 / (double)voxel);
  float ox = floorf(mnx * inv) * voxel, oy = floorf(mny * inv) * voxel, oz = floorf(mnz * inv) * voxel;
  uint64_t nx = (uint64_t)(floorf((mxx - ox) / voxel) + 1.0f); /* :13-20 */
  uint64_t ny = (uint64_t)(floorf((mxy - oy) / voxel) + 1.0f);

  table_t t;
  t.cap = n > 0 ? n : 1;
  t.nodes = (node_t*)malloc(sizeof(node_t) * (size_t)t.cap);
  t.n = 0;
  t.head = -1;
  t.bucket = NULL;
  t.nb = 0;
  t.era = -1;
  for (int64_t i = 0; i < n; i++) {
    const float* p = pts + 3 * i;
    uint64_t ix = (uint64_t)floorf((p[0] - ox) / voxel); /* :32-35 */
    uint64_t iy = (uint64_t)floorf((p[1] - oy) / voxel);
    uint64_t iz = (uint64_t)floorf((p[2] - oz) / voxel);
    uint64_t key = ix + nx * iy + nx * ny * iz;
    int64_t id = find(&t, key);
    if (id < 0) {
      int64_t prev = t.era >= 0 ? BKT[t.era] : 0;
      if (t.n + 1 > prev) { /* _Prime_rehash_policy::_M_need_rehash with max_load_factor 1 */
        t.era++;
        if (t.era >= NBKT) abort();
        rehash(&t, BKT[t.era]);
      }
      id = t.n++;
      t.nodes[id].key = key;
      t.nodes[id].sx = t.nodes[id].sy = t.nodes[id].sz = 0.0f;
      t.nodes[id].count = 0;
      place(&t, t.bucket, t.nb, id);
    }
    t.nodes[id].count += 1; /* grid_subsampling_cpu.h:17-20 */
    t.nodes[id].sx += p[0];
    t.nodes[id].sy += p[1];
    t.nodes[id].sz += p[2];
  }
  int64_t m = 0;
  for (int64_t p = t.head; p >= 0; p = t.nodes[p].next, m++) { /* :45-47 */
    float s = (float)(1.0 / (double)t.nodes[p].count);
    out[3 * m + 0] = t.nodes[p].sx * s;
    out[3 * m + 1] = t.nodes[p].sy * s;
    out[3 * m + 2] = t.nodes[p].sz * s;
  }
  free(t.nodes);
  free(t.bucket);
  return m;
}

/* grid_subsampling_cpu.cpp:50-75. out_points capacity = total input points. Returns total output points. */
int64_t oracle_grid_subsample(const float* points, const int64_t* lengths, int64_t batch, float voxel,
                              float* out_points, int64_t* out_lengths) {
  int64_t start = 0, total = 0;
  for (int64_t b = 0; b < batch; b++) {
    int64_t m = subsample_one(points + 3 * start, lengths[b], voxel, out_points + 3 * total);
    out_lengths[b] = m;
    total += m;
    start += lengths[b];
  }
  return total;
}

typedef struct {
  float d2;
  int64_t idx;
} hit_t;

static int cmp_hit(const void* a, const void* b) {
  const hit_t* x = (const hit_t*)a;
  const hit_t* y = (const hit_t*)b;
  if (x->d2 < y->d2) return -1;
  if (x->d2 > y->d2) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx); /* canonical tie order: ascending index */
}

/* Pass 1 (out == NULL): returns max neighbour count over all queries. Pass 2: fills out[nq, width] with the
 * first `width` neighbours in (d2, idx) order, padded with ns_total. radius_neighbors_cpu.cpp:3-91. */
int64_t oracle_radius_neighbors(const float* q, const float* s, const int64_t* q_len, const int64_t* s_len,
                                int64_t batch, float radius, int64_t width, int64_t* out) {
  float r2 = radius * radius; /* :12 */
  int64_t nq_total = 0, ns_total = 0, max_count = 0;
  for (int64_t b = 0; b < batch; b++) {
    nq_total += q_len[b];
    ns_total += s_len[b];
  }
  int64_t q0 = 0, s0 = 0;
  hit_t* hits = NULL;
  int64_t cap = 0;
  for (int64_t b = 0; b < batch; b++) {
    if (s_len[b] > cap) {
      cap = s_len[b];
      hits = (hit_t*)realloc(hits, sizeof(hit_t) * (size_t)cap);
    }
    for (int64_t i = 0; i < q_len[b]; i++) {
      const float* qp = q + 3 * (q0 + i);
      int64_t cnt = 0;
      for (int64_t j = 0; j < s_len[b]; j++) {
        const float* sp = s + 3 * (s0 + j);
        /* nanoflann L2_Simple_Adaptor::evalMetric: result += diff*diff for d = 0,1,2 starting from 0 */
        float dx = qp[0] - sp[0], dy = qp[1] - sp[1], dz = qp[2] - sp[2];
        float d2 = dx * dx;
        d2 = d2 + dy * dy;
        d2 = d2 + dz * dz;
        if (d2 < r2) { /* RadiusResultSet::addPoint, strict */
          hits[cnt].d2 = d2;
          hits[cnt].idx = j + s0; /* :83 global index */
          cnt++;
        }
      }
      if (cnt > max_count) max_count = cnt;
      if (out) {
        qsort(hits, (size_t)cnt, sizeof(hit_t), cmp_hit);
        int64_t* row = out + (q0 + i) * width;
        for (int64_t k = 0; k < width; k++) row[k] = k < cnt ? hits[k].idx : ns_total; /* :85 pad */
      }
    }
    q0 += q_len[b];
    s0 += s_len[b];
  }
  free(hits);
  (void)nq_total;
  return max_count;
}
