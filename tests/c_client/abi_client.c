/* abi_client.c -- a plain C client of include/gfi.h, linked against libgfi.so: catches header / ABI drift without a
 * Rust toolchain.  It makes the calls the Rust wrapper (rust/gpu-flat-index) makes for `trait Index`
 * (reference src/index.rs:11-35) and checks them against the reference's own known-answer values
 * (src/flat_index.rs:81-114, src/distance.rs:81-143).  Compiled by gcc in the CPU test-suite (compile + link only);
 * run on the GPU box by tests/test_gpu_c_client.py.  Exit code 0 = all checks passed.
 *
 * build: gcc -std=c99 -Wall -Werror -I include tests/c_client/abi_client.c -L vectordb-from-scratch_b200 -lgfi \
 *            -Wl,-rpath,$PWD/vectordb-from-scratch_b200 -lm -o tests/c_client/abi_client
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gfi.h"

static int failures = 0;
#define CHECK(cond)                                                                  \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      fprintf(stderr, "FAIL %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, gfi_last_error()); \
      ++failures;                                                                    \
    }                                                                                \
  } while (0)

static int run(const int32_t *devices, int32_t n_devices) {
  gfi_index *h = NULL;
  if (n_devices > 0)
    CHECK(gfi_create_sharded(&h, GFI_METRIC_EUCLIDEAN, 0, devices, n_devices, 0) == GFI_OK);
  else
    CHECK(gfi_create(&h, GFI_METRIC_EUCLIDEAN, 0, 0, 0) == GFI_OK);
  if (!h) return 1;
  if (n_devices > 0) CHECK(gfi_set_option(h, "shard_block", 1) == GFI_OK);

  /* flat_index.rs:81-94: three rows, the query equals row 0 */
  const float rows[3][3] = {{1.f, 2.f, 3.f}, {4.f, 5.f, 6.f}, {7.f, 8.f, 9.f}};
  for (uint64_t i = 0; i < 3; ++i) CHECK(gfi_add(h, &i, rows[i], 1, 3) == GFI_OK);
  CHECK(gfi_len(h) == 3);
  CHECK(gfi_metric(h) == GFI_METRIC_EUCLIDEAN);
  CHECK(gfi_dim(h) == 3);
  const float q[3] = {1.f, 2.f, 3.f};
  uint32_t k = 2, cnt = 0;
  uint64_t ids[4];
  float dist[4];
  CHECK(gfi_search(h, q, 1, 3, &k, NULL, 0, ids, dist, &cnt, 4) == GFI_OK);
  CHECK(cnt == 2 && ids[0] == 0 && dist[0] == 0.f && ids[1] == 1);
  CHECK(fabsf(dist[1] - sqrtf(27.f)) < 1e-6f); /* distance.rs: sqrt(3 * 3^2) */

  /* get_vector hit / miss (flat_index.rs:96-103) */
  float back[3];
  int64_t od = 0;
  CHECK(gfi_get_vector(h, 1, back, 3, &od) == GFI_OK && od == 3 && memcmp(back, rows[1], sizeof back) == 0);
  CHECK(gfi_get_vector(h, 99, back, 3, &od) == GFI_ERR_INDEX);

  /* dimension mismatch carries the payload of VectorDbError::DimensionMismatch */
  const float q2[2] = {1.f, 2.f};
  CHECK(gfi_search(h, q2, 1, 2, &k, NULL, 0, ids, dist, &cnt, 4) == GFI_ERR_DIMENSION_MISMATCH);
  int64_t e = 0, a = 0;
  gfi_last_mismatch(&e, &a);
  CHECK(e == 2 && a == 3); /* FlatIndex: expected = the query, actual = the stored row (distance.rs:21-25) */

  /* eligibility mask by internal id: only ids 1 and 2 */
  const uint64_t mask = 0x6;
  k = 3;
  CHECK(gfi_search(h, q, 1, 3, &k, &mask, 3, ids, dist, &cnt, 4) == GFI_OK);
  CHECK(cnt == 2 && ids[0] == 1 && ids[1] == 2);

  /* metadata + the reference's JSON filter form (storage.rs:44-58) */
  const char *keys[1] = {"color"};
  const char *red[1] = {"red"}, *blue[1] = {"blue"};
  CHECK(gfi_set_metadata(h, 0, 1, keys, red) == GFI_OK);
  CHECK(gfi_set_metadata(h, 1, 1, keys, blue) == GFI_OK);
  CHECK(gfi_set_metadata(h, 2, 1, keys, red) == GFI_OK);
  CHECK(gfi_search_filtered(h, q, 1, 3, &k, "{\"op\":\"eq\",\"field\":\"color\",\"value\":\"red\"}", ids, dist, &cnt, 4) == GFI_OK);
  CHECK(cnt == 2 && ids[0] == 0 && ids[1] == 2);
  const uint64_t col_ids[2] = {0, 1};
  const char *vals[2] = {"s", "m"};
  const uint32_t codes[2] = {1, 0};
  CHECK(gfi_set_metadata_column(h, "size", col_ids, 2, vals, 2, codes) == GFI_OK);
  CHECK(gfi_search_filtered(h, q, 1, 3, &k, "{\"op\":\"eq\",\"field\":\"size\",\"value\":\"m\"}", ids, dist, &cnt, 4) == GFI_OK);
  CHECK(cnt == 1 && ids[0] == 0);

  /* exact pair distances (HNSW candidate evaluation) */
  const uint64_t cand[3] = {2, 0, 77};
  float pd[3];
  uint8_t ps[3];
  CHECK(gfi_distances(h, q, 1, 3, cand, 3, pd, ps) == GFI_OK);
  CHECK(ps[0] == 0 && ps[1] == 0 && ps[2] == 1 && pd[1] == 0.f && fabsf(pd[0] - sqrtf(108.f)) < 1e-5f);

  /* remove is idempotent (flat_index.rs:106-114) */
  CHECK(gfi_remove(h, 1) == GFI_OK && gfi_remove(h, 1) == GFI_OK && gfi_len(h) == 2);
  CHECK(gfi_flush(h) == GFI_OK && gfi_compact(h) == GFI_OK);
  k = 3;
  CHECK(gfi_search(h, q, 1, 3, &k, NULL, 0, ids, dist, &cnt, 4) == GFI_OK);
  CHECK(cnt == 2 && ids[0] == 0 && ids[1] == 2);

  gfi_stats st;
  CHECK(gfi_get_stats(h, &st) == GFI_OK && st.n_live == 2 && st.shards == (n_devices > 0 ? n_devices : 1));
  CHECK(gfi_search_status(h) == GFI_OK); /* nothing pending */
  CHECK(gfi_destroy(h) == GFI_OK);
  return 0;
}

int main(void) {
  CHECK(gfi_version() >= 101);
  run(NULL, 0);
  const int32_t two_shards_one_gpu[2] = {0, 0};
  run(two_shards_one_gpu, 2);
  /* a generated bulk load through the remaining entry points */
  gfi_index *h = NULL;
  CHECK(gfi_create(&h, GFI_METRIC_COSINE, 64, 0, GFI_FLAG_NO_TENSOR) == GFI_OK);
  CHECK(gfi_reserve(h, 5000) == GFI_OK);
  CHECK(gfi_add_generated(h, 3, 0, 5000, GFI_GEN_NORMAL, 0) == GFI_OK);
  float row[64];
  int64_t od = 0;
  CHECK(gfi_get_vector(h, 4242, row, 64, &od) == GFI_OK && od == 64);
  uint32_t k = 5, cnt = 0;
  uint64_t ids[5];
  float dist[5];
  CHECK(gfi_search(h, row, 1, 64, &k, NULL, 0, ids, dist, &cnt, 5) == GFI_OK);
  CHECK(cnt == 5 && ids[0] == 4242 && dist[0] < 1e-6f);
  CHECK(gfi_destroy(h) == GFI_OK);
  if (failures == 0) printf("abi_client ok\n");
  return failures ? 1 : 0;
}
