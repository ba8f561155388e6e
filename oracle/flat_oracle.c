/*
 * flat_oracle.c -- CPU restatement of the reference's exact-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object, and only as the checker / CPU arm.
 *
 * The reference (Ricoledan/vectordb-from-scratch) is pure Rust and no Rust
 * toolchain exists in this image, so the reference itself cannot be run here
 * (oracle/_ref is therefore never built).  The arithmetic of the path lives
 * entirely in the reference's own sources (no third-party crate computes
 * anything), so this file restates it operation for operation:
 *
 *   Vector::norm                      src/vector.rs:35-37
 *   DistanceMetric::distance          src/distance.rs:20-33
 *   euclidean_distance                src/distance.rs:37-44
 *   cosine_distance                   src/distance.rs:47-64
 *   dot_product                       src/distance.rs:67-73
 *   FlatIndex::search                 src/flat_index.rs:52-65
 *   VectorStore::search_with_filter   src/storage.rs:249-290 (post-filter, 3x over-fetch)
 *   VectorStore::search_batch         src/storage.rs:302-310 (per-query k, fail fast)
 *
 * Arithmetic contract: IEEE-754 binary32; every sum is a left-to-right
 * sequential f32 sum with separately rounded multiply and add (Rust never
 * contracts to FMA and never re-associates); the neutral element of
 * `Iterator::sum::<f32>()` is -0.0 (Rust >= 1.83); sqrt and divide are correctly
 * rounded.  This file MUST be compiled with -ffp-contract=off and without
 * -ffast-math (the Makefile does), which makes it bit-identical to the Rust
 * build on the same inputs.
 *
 * Pinned against: every known-answer test the reference holds for this path
 * (src/distance.rs:81-143, src/vector.rs:137-140, src/flat_index.rs:81-114,
 * src/storage.rs:384-404,578-630,680-755, tests/integration_test.rs:6-47) --
 * see tests/test_oracle_kat.py -- and against an independent numpy.float32
 * scalar restatement (oracle/pyref.py).
 *
 * Tie order: the reference iterates a HashMap (random order) and stable-sorts,
 * so equal distances come out in unspecified order.  This oracle uses the
 * build's stated rule -- (distance ascending, then lower internal id) -- which
 * is one of the reference's legal outputs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_DIMENSION_MISMATCH 1 /* VectorDbError::DimensionMismatch, src/error.rs */
#define ORC_INVALID_VECTOR 2     /* VectorDbError::InvalidVector (cosine, zero norm) */
#define ORC_NAN 3                /* reference panics: partial_cmp().unwrap(), flat_index.rs:62 */

#define ORC_EUCLIDEAN 0
#define ORC_COSINE 1
#define ORC_DOT 2

/* ---- src/vector.rs:35-37 ------------------------------------------------ */
float orc_norm(const float *v, int64_t d) {
  float acc = -0.0f;
  for (int64_t i = 0; i < d; ++i) {
    float sq = v[i] * v[i];
    acc = acc + sq;
  }
  return sqrtf(acc);
}

/* ---- src/distance.rs:67-73 ---------------------------------------------- */
float orc_dot(const float *a, const float *b, int64_t d) {
  float acc = -0.0f;
  for (int64_t i = 0; i < d; ++i) {
    float p = a[i] * b[i];
    acc = acc + p;
  }
  return acc;
}

/* ---- src/distance.rs:37-44 : sqrt(sum((a-b)^2)), powi(2) == t*t ---------- */
float orc_euclidean(const float *a, const float *b, int64_t d) {
  float acc = -0.0f;
  for (int64_t i = 0; i < d; ++i) {
    float t = a[i] - b[i];
    float sq = t * t;
    acc = acc + sq;
  }
  return sqrtf(acc);
}

/* ---- src/distance.rs:47-64 ---------------------------------------------- */
int orc_cosine(const float *a, const float *b, int64_t d, float *out) {
  float n1 = orc_norm(a, d);
  float n2 = orc_norm(b, d);
  if (n1 == 0.0f || n2 == 0.0f) return ORC_INVALID_VECTOR;
  float dot = orc_dot(a, b, d);
  float denom = n1 * n2;
  float sim = dot / denom;
  /* f32::clamp(-1, 1): NaN stays NaN */
  if (sim < -1.0f) sim = -1.0f;
  else if (sim > 1.0f) sim = 1.0f;
  *out = 1.0f - sim;
  return ORC_OK;
}

/* ---- src/distance.rs:20-33 : v1 = query, v2 = stored row ------------------ */
int orc_distance(int metric, const float *q, int64_t dq, const float *x, int64_t dx,
                 float *out) {
  if (dq != dx) return ORC_DIMENSION_MISMATCH;
  switch (metric) {
    case ORC_EUCLIDEAN: *out = orc_euclidean(q, x, dq); return ORC_OK;
    case ORC_COSINE: return orc_cosine(q, x, dq, out);
    default: *out = -orc_dot(q, x, dq); return ORC_OK;
  }
}

typedef struct {
  float dist;
  uint64_t id;
} orc_pair;

static int pair_cmp(const void *pa, const void *pb) {
  const orc_pair *a = (const orc_pair *)pa, *b = (const orc_pair *)pb;
  if (a->dist < b->dist) return -1;
  if (a->dist > b->dist) return 1;
  if (a->id < b->id) return -1;
  if (a->id > b->id) return 1;
  return 0;
}

/*
 * FlatIndex::search, src/flat_index.rs:52-65: score every stored row, sort the
 * whole (id, distance) vector ascending, truncate to k.
 *   rows : n x d contiguous; ids : n internal ids (NULL => id = row number)
 *   eligible : optional bitmask over rows (bit r of word r/64); NULL = all rows.
 *              A masked search is FlatIndex::search over the eligible subset.
 *   cos_qnorm_once: the reference recomputes the query norm per row; the value
 *              is identical each time, so computing it once changes nothing.
 * Returns ORC_* ; *out_count = min(k, rows scored).
 */
int orc_flat_search(int metric, const float *rows, const uint64_t *ids, int64_t n, int64_t d,
                    const uint64_t *eligible, const float *query, int64_t dq, int64_t k,
                    uint64_t *out_ids, float *out_dist, int64_t *out_count) {
  *out_count = 0;
  if (n > 0 && dq != d) return ORC_DIMENSION_MISMATCH;
  orc_pair *res = (orc_pair *)malloc(sizeof(orc_pair) * (size_t)(n > 0 ? n : 1));
  int64_t m = 0;
  int rc = ORC_OK;
  int saw_nan = 0;
  for (int64_t r = 0; r < n; ++r) {
    if (eligible && !((eligible[r >> 6] >> (r & 63)) & 1ull)) continue;
    float dist;
    rc = orc_distance(metric, query, dq, rows + r * d, d, &dist);
    if (rc != ORC_OK) break; /* collect::<Result<_>>() stops at the first Err */
    if (dist != dist) saw_nan = 1;
    res[m].dist = dist;
    res[m].id = ids ? ids[r] : (uint64_t)r;
    ++m;
  }
  if (rc == ORC_OK && saw_nan && m > 1) rc = ORC_NAN; /* sort_by(..unwrap()) panics */
  if (rc != ORC_OK) {
    free(res);
    return rc;
  }
  qsort(res, (size_t)m, sizeof(orc_pair), pair_cmp);
  int64_t take = k < m ? k : m;
  if (take < 0) take = 0;
  for (int64_t i = 0; i < take; ++i) {
    out_ids[i] = res[i].id;
    out_dist[i] = res[i].dist;
  }
  *out_count = take;
  free(res);
  return ORC_OK;
}

/*
 * VectorStore::search_with_filter, src/storage.rs:249-290 (post-filter):
 * fetch_k = min(max(3k, k), len); FlatIndex::search(fetch_k); keep rows whose
 * metadata matches (here: bit set in `matches`); take(k).
 */
int orc_search_post_filter(int metric, const float *rows, const uint64_t *ids, int64_t n,
                           int64_t d, const uint64_t *matches, const float *query, int64_t dq,
                           int64_t k, uint64_t *out_ids, float *out_dist, int64_t *out_count) {
  *out_count = 0;
  if (n == 0) return ORC_OK; /* storage.rs:255-257: empty store => [] before any check */
  if (dq != d) return ORC_DIMENSION_MISMATCH;
  int64_t fetch_k = k * 3;
  if (fetch_k < k) fetch_k = k;
  if (fetch_k > n) fetch_k = n;
  uint64_t *fid = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(fetch_k > 0 ? fetch_k : 1));
  float *fd = (float *)malloc(sizeof(float) * (size_t)(fetch_k > 0 ? fetch_k : 1));
  int64_t got = 0;
  int rc = orc_flat_search(metric, rows, NULL, n, d, NULL, query, dq, fetch_k, fid, fd, &got);
  if (rc == ORC_OK) {
    int64_t w = 0;
    for (int64_t i = 0; i < got && w < k; ++i) {
      uint64_t r = fid[i]; /* row number */
      if ((matches[r >> 6] >> (r & 63)) & 1ull) {
        out_ids[w] = ids ? ids[r] : r;
        out_dist[w] = fd[i];
        ++w;
      }
    }
    *out_count = w;
  }
  free(fid);
  free(fd);
  return rc;
}

/*
 * VectorStore::search_batch, src/storage.rs:302-310: a sequential map of
 * single-query searches with per-query k; the first error fails the batch.
 * Output is kmax-strided.  `threads` > 1 runs queries on that many OpenMP
 * threads (the "rayon over queries" arm north_star names; results identical).
 */
int orc_search_batch(int metric, const float *rows, const uint64_t *ids, int64_t n, int64_t d,
                     const uint64_t *eligible, const float *queries, int64_t q, int64_t dq,
                     const int64_t *ks, int64_t kmax, uint64_t *out_ids, float *out_dist,
                     int64_t *out_counts, int threads) {
  int rc_all = ORC_OK;
  int64_t first_bad = q;
  (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
#endif
  for (int64_t i = 0; i < q; ++i) {
    int rc = orc_flat_search(metric, rows, ids, n, d, eligible, queries + i * dq, dq, ks[i],
                             out_ids + i * kmax, out_dist + i * kmax, out_counts + i);
    if (rc != ORC_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
      {
        if (i < first_bad) {
          first_bad = i;
          rc_all = rc;
        }
      }
    }
  }
  return rc_all;
}

/* ------------------------------------------------------------------------ *
 * Synthetic data generator: counter-based, identical in C, CUDA and numpy
 * (SURVEY.md section 8(d) M2).  element(seed,row,col) -> u32 -> f32.
 *   kind 0: U[0,1)  = (h >> 8) * 2^-24
 *   kind 1: zero-mean unit-variance "normal-like" (Irwin-Hall-4 of 16-bit
 *           uniforms from two hashes): exact integer sum, one f32 multiply.
 * ------------------------------------------------------------------------ */
static inline uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
static inline uint32_t gen_u32(uint32_t seed, uint64_t row, uint32_t col, uint32_t lane) {
  uint32_t h = mix32(seed * 0x9E3779B1u + (uint32_t)(row >> 32) + 0x7F4A7C15u * lane);
  h = mix32(h ^ (uint32_t)row);
  h = mix32(h + col * 0x85EBCA77u);
  return h;
}
float orc_gen_elem(uint32_t seed, uint64_t row, uint32_t col, int kind) {
  uint32_t h = gen_u32(seed, row, col, 0);
  if (kind == 0) return (float)(h >> 8) * (1.0f / 16777216.0f);
  uint32_t g = gen_u32(seed, row, col, 1);
  int32_t s = (int32_t)(h & 0xFFFF) + (int32_t)(h >> 16) + (int32_t)(g & 0xFFFF) + (int32_t)(g >> 16) -
              2 * 65535;
  return (float)s * (1.7320508f / 65536.0f);
}
void orc_gen_rows(uint32_t seed, uint64_t first_row, int64_t n, int64_t d, int kind, float *out) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int64_t r = 0; r < n; ++r)
    for (int64_t c = 0; c < d; ++c)
      out[r * d + c] = orc_gen_elem(seed, first_row + (uint64_t)r, (uint32_t)c, kind);
}

/*
 * FlatIndex::search over GENERATED rows, rows-parallel: the checker for full-size configurations (10M x 768 does
 * not fit a test process as an array, and one query over it takes 14 s on one core).  Row r is
 * orc_gen_elem(seed, first_row + r, *, kind) with id first_id + r, produced on the fly; every thread walks a
 * contiguous range of rows, scores ALL q queries against each row it generates (same sequential arithmetic as
 * orc_distance) and keeps, per query, its k best by (distance, id); the per-thread lists are then merged by the
 * same order.  The k smallest of a set under a total order do not depend on how the set was partitioned, so the
 * result equals score-all + sort + truncate(k) (flat_index.rs:52-65) with the stated tie rule.
 *   eligible: optional bitmask over rows (bit r of word r/64).  Output kmax-strided like orc_search_batch.
 */
int orc_search_generated(int metric, uint32_t seed, uint64_t first_row, uint64_t first_id, int64_t n, int64_t d,
                         int kind, const uint64_t *eligible, const float *queries, int64_t q, const int64_t *ks,
                         int64_t kmax, uint64_t *out_ids, float *out_dist, int64_t *out_counts, int threads) {
  if (threads < 1) threads = 1;
  int rc_all = ORC_OK;
  orc_pair *best = (orc_pair *)malloc(sizeof(orc_pair) * (size_t)threads * (size_t)q * (size_t)kmax);
  int64_t *have = (int64_t *)calloc((size_t)threads * (size_t)q, sizeof(int64_t));
  float *qnorm = (float *)malloc(sizeof(float) * (size_t)(q > 0 ? q : 1));
  for (int64_t i = 0; i < q; ++i) qnorm[i] = orc_norm(queries + i * d, d);
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
    const int t = 0, nt = 1;
#endif
    float *row = (float *)malloc(sizeof(float) * (size_t)d);
    const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
    int rc = ORC_OK;
    for (int64_t r = lo; r < hi && rc == ORC_OK; ++r) {
      if (eligible && !((eligible[r >> 6] >> (r & 63)) & 1ull)) continue;
      for (int64_t c = 0; c < d; ++c) row[c] = orc_gen_elem(seed, first_row + (uint64_t)r, (uint32_t)c, kind);
      const float xn = metric == ORC_COSINE ? orc_norm(row, d) : 1.0f;
      for (int64_t i = 0; i < q; ++i) {
        const float *qv = queries + i * d;
        float dist;
        if (metric == ORC_EUCLIDEAN) {
          dist = orc_euclidean(qv, row, d);
        } else if (metric == ORC_DOT) {
          dist = -orc_dot(qv, row, d);
        } else {
          if (qnorm[i] == 0.0f || xn == 0.0f) { rc = ORC_INVALID_VECTOR; break; }
          float sim = orc_dot(qv, row, d) / (qnorm[i] * xn);
          if (sim < -1.0f) sim = -1.0f;
          else if (sim > 1.0f) sim = 1.0f;
          dist = 1.0f - sim;
        }
        if (dist != dist) { rc = ORC_NAN; break; }
        const int64_t k = ks[i];
        if (k <= 0) continue;
        orc_pair *lst = best + ((size_t)t * (size_t)q + (size_t)i) * (size_t)kmax;
        int64_t *m = have + (size_t)t * (size_t)q + (size_t)i;
        orc_pair cand;
        cand.dist = dist;
        cand.id = first_id + (uint64_t)r;
        if (*m == k && pair_cmp(&cand, &lst[k - 1]) >= 0) continue;
        int64_t pos = *m < k ? (*m)++ : k - 1; /* insertion into the ascending list */
        while (pos > 0 && pair_cmp(&cand, &lst[pos - 1]) < 0) {
          lst[pos] = lst[pos - 1];
          --pos;
        }
        lst[pos] = cand;
      }
    }
    free(row);
    if (rc != ORC_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
      rc_all = rc;
    }
  }
  if (rc_all == ORC_OK) {
    orc_pair *all = (orc_pair *)malloc(sizeof(orc_pair) * (size_t)threads * (size_t)(kmax > 0 ? kmax : 1));
    for (int64_t i = 0; i < q; ++i) {
      int64_t m = 0;
      for (int t = 0; t < threads; ++t) {
        const orc_pair *lst = best + ((size_t)t * (size_t)q + (size_t)i) * (size_t)kmax;
        for (int64_t j = 0; j < have[(size_t)t * (size_t)q + (size_t)i]; ++j) all[m++] = lst[j];
      }
      qsort(all, (size_t)m, sizeof(orc_pair), pair_cmp);
      const int64_t take = ks[i] < m ? ks[i] : m;
      for (int64_t j = 0; j < take; ++j) {
        out_ids[i * kmax + j] = all[j].id;
        out_dist[i * kmax + j] = all[j].dist;
      }
      out_counts[i] = take;
    }
    free(all);
  }
  free(best);
  free(have);
  free(qnorm);
  return rc_all;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
