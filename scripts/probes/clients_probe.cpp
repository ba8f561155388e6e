// Native concurrent-client probe: T host threads each issue P single-query gfi_search calls against one index
// (what the reference's HTTP handlers do under load), with and without the library's group commit.
// build: g++ -O2 -std=c++17 -pthread -I include scripts/probes/clients_probe.cpp -L vectordb-from-scratch_b200 -lgfi \
//        -Wl,-rpath,$PWD/vectordb-from-scratch_b200 -o gpurun_out/clients_probe
// usage: clients_probe <metric 0|1|2> <rows> <dim> <kind> <seed> <k> <threads> <per_thread>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "gfi.h"

int main(int argc, char** argv) {
  if (argc < 9) { fprintf(stderr, "usage: metric rows dim kind seed k threads per_thread\n"); return 2; }
  const int metric = atoi(argv[1]);
  const long long rows = atoll(argv[2]);
  const int dim = atoi(argv[3]), kind = atoi(argv[4]), seed = atoi(argv[5]), k = atoi(argv[6]);
  const int T = atoi(argv[7]), P = atoi(argv[8]);
  gfi_index* h = nullptr;
  if (gfi_create(&h, metric, dim, 0, 0) != 0) { fprintf(stderr, "create: %s\n", gfi_last_error()); return 1; }
  gfi_reserve(h, rows);
  if (gfi_add_generated(h, (uint32_t)seed, 0, rows, kind, 0) != 0) { fprintf(stderr, "add: %s\n", gfi_last_error()); return 1; }
  gfi_flush(h);
  // queries: rows of the same generator family, fetched back from a scratch index so no generator is duplicated here
  gfi_index* qh = nullptr;
  gfi_create(&qh, metric, dim, 0, 0);
  gfi_add_generated(qh, (uint32_t)seed + 1, 0, (long long)T * P, kind, 0);
  gfi_flush(qh);
  std::vector<float> queries((size_t)T * P * dim);
  for (long long i = 0; i < (long long)T * P; ++i) {
    int64_t od = 0;
    gfi_get_vector(qh, (uint64_t)i, queries.data() + (size_t)i * dim, dim, &od);
  }
  gfi_destroy(qh);
  for (int pass = 0; pass < 4; ++pass) {
    const int co = pass & 1;
    gfi_set_option(h, "coalesce", co);
    gfi_stats s0, s1;
    gfi_get_stats(h, &s0);
    std::atomic<int> errors{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t] {
        std::vector<uint64_t> ids(k);
        std::vector<float> dist(k);
        uint32_t cnt = 0, kk = (uint32_t)k;
        for (int j = 0; j < P; ++j)
          if (gfi_search(h, queries.data() + ((size_t)t * P + j) * dim, 1, dim, &kk, nullptr, 0, ids.data(), dist.data(),
                         &cnt, k) != 0 || cnt != (uint32_t)k)
            ++errors;
      });
    for (auto& x : th) x.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    gfi_get_stats(h, &s1);
    printf("{\"rows\": %lld, \"dim\": %d, \"metric\": %d, \"threads\": %d, \"per_thread\": %d, \"coalesce\": %d, \"qps\": %.1f, "
           "\"seconds\": %.4f, \"searches_run\": %lld, \"coalesced_batches\": %lld, \"tensor_queries\": %lld, "
           "\"fallback_queries\": %lld, \"errors\": %d}\n",
           rows, dim, metric, T, P, co, T * P / dt, dt, (long long)(s1.searches - s0.searches),
           (long long)(s1.coalesced_batches - s0.coalesced_batches), (long long)(s1.tensor_queries - s0.tensor_queries),
           (long long)(s1.fallback_queries - s0.fallback_queries), errors.load());
    fflush(stdout);
  }
  gfi_destroy(h);
  return 0;
}
