// queue_bench.cpp - SURVEY.md 8f.1: what the coalescing S2 queue buys.  K "encoder instances" (host threads) each submit one CU's worth
// of PUs (an 8x8 2Nx2N PU + its four 4x4 NxN PUs) per request and wait for the result before the next one - the live encoder's serial
// dependency (a PU's border is the reconstruction of its predecessors).  Prints one JSON line per K.  Built and run by bench_rows.py:
//   g++ -O2 -std=c++17 -I include profiles/ubench/queue_bench.cpp -L fast-cu-decision-hevc_b200 -lcucudecide -Wl,-rpath,... -lpthread
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include "cucudecide.h"

int main(int argc, char** argv) {
  const int nReq = argc > 1 ? atoi(argv[1]) : 2000;
  cucd_config cfg = {64, 64, 8, 64, 4, 1, 0, 1, 0};
  cucd_handle* h = nullptr;
  if (cucd_create(&cfg, &h) != CUCD_OK) { fprintf(stderr, "%s\n", cucd_last_error(nullptr)); return 1; }
  const cucd_pu_desc desc[5] = {{3, {0, 0, 0}}, {2, {0, 0, 0}}, {2, {0, 0, 0}}, {2, {0, 0, 0}}, {2, {0, 0, 0}}};
  std::vector<int16_t> org(64 + 4 * 16), brd(33 + 4 * 17);
  for (size_t i = 0; i < org.size(); i++) org[i] = (int16_t)((i * 37 + 11) & 255);
  for (size_t i = 0; i < brd.size(); i++) brd[i] = (int16_t)((i * 53 + 7) & 255);
  for (int K : {1, 2, 4, 8, 16, 32}) {
    cucd_queue* q = nullptr;
    if (cucd_queue_create(h, &q) != CUCD_OK) return 1;
    auto client = [&]() {
      std::vector<uint32_t> sad(5 * 35);
      for (int r = 0; r < nReq; r++) {
        uint64_t t = 0;
        if (cucd_queue_submit(q, 5, desc, org.data(), brd.data(), sad.data(), &t) != CUCD_OK || cucd_queue_wait(q, t) != CUCD_OK) { fprintf(stderr, "queue call failed\n"); exit(1); }
      }
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int k = 0; k < K; k++) th.emplace_back(client);
    for (auto& t : th) t.join();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long long requests = 0, pus = 0, batches = 0;
    cucd_queue_stats(q, &requests, &pus, &batches);
    cucd_queue_destroy(q);
    printf("{\"row\": \"f1_rmd_queue\", \"instances\": %d, \"requests\": %lld, \"pus_per_request\": 5, \"batches\": %lld, \"requests_per_batch\": %.2f, "
           "\"e2e\": {\"value\": %.1f, \"unit\": \"PU/s\", \"us_per_request_per_instance\": %.1f}}\n",
           K, requests, batches, (double)requests / (double)(batches ? batches : 1), (double)pus / sec, 1e6 * sec / nReq);
    fflush(stdout);
  }
  cucd_destroy(h);
  return 0;
}
