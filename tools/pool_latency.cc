// BASELINE config 5 through the session pool: N concurrent EncoderSessions, one OS thread each, every thread feeding
// 1152-sample stereo chunks through the blocking mp3b_pool_encode (the reference's encode(samples:)) as fast as the pool
// answers; prints one JSON line with the p50 / p99 latency of a call and the coalescing achieved.
//   build: g++ -O2 -std=c++17 -pthread tools/pool_latency.cc -Iinclude -Lswift-mp3_b200 -lmp3b200 -Wl,-rpath,$PWD/swift-mp3_b200 -o tools/pool_latency
//   run:   tools/pool_latency [sessions=1024] [chunks=200] [max_wait_us=300]
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "mp3b200.h"

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1024, chunks = argc > 2 ? atoi(argv[2]) : 200, wait_us = argc > 3 ? atoi(argv[3]) : 300;
  mp3b_options o; mp3b_options_default(&o);
  mp3b_pool *pool = nullptr;
  if (mp3b_pool_create(&o, n, 0, wait_us, &pool) != MP3B_OK) { fprintf(stderr, "pool create failed: %s\n", mp3b_last_error()); return 1; }
  std::vector<std::vector<double>> lat((size_t)n);
  std::atomic<int> ready{0}, failed{0};
  std::atomic<bool> go{false};
  std::vector<std::thread> th;
  for (int i = 0; i < n; ++i)
    th.emplace_back([&, i] {
      int slot = -1;
      if (mp3b_pool_open(pool, &slot) != MP3B_OK) { failed++; return; }
      std::vector<float> pcm(2304);
      for (int k = 0; k < 2304; ++k) pcm[k] = 0.5f * sinf(0.0003f * (float)(k / 2) * (float)(50 + i % 97)) + 0.01f * (float)((k * 7919 + i) % 13 - 6);
      std::vector<uint8_t> out(8192);
      ready++;
      while (!go.load()) std::this_thread::yield();
      lat[(size_t)i].reserve((size_t)chunks);
      for (int c = 0; c < chunks; ++c) {
        size_t w = 0;
        auto t0 = std::chrono::steady_clock::now();
        if (mp3b_pool_encode(pool, slot, pcm.data(), pcm.size(), out.data(), out.size(), &w) != MP3B_OK) { failed++; return; }
        lat[(size_t)i].push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
      }
      size_t w = 0;
      mp3b_pool_flush(pool, slot, out.data(), out.size(), &w);
    });
  while (ready.load() + failed.load() < n) std::this_thread::yield();
  auto t0 = std::chrono::steady_clock::now();
  go = true;
  for (auto &t : th) t.join();
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::vector<double> all;
  for (auto &v : lat) all.insert(all.end(), v.begin() + std::min<size_t>(v.size(), 20), v.end());    // skip 20 warm-up chunks
  std::sort(all.begin(), all.end());
  uint64_t steps = 0, reqs = 0;
  mp3b_pool_stats(pool, &steps, &reqs);
  auto pct = [&](double q) { return all.empty() ? 0.0 : all[std::min(all.size() - 1, (size_t)(q * (double)all.size()))]; };
  printf("{\"metric\": \"per-call latency of encode(samples:), %d concurrent sessions on %d threads (config 5, session pool)\", "
         "\"p50_ms\": %.3f, \"p99_ms\": %.3f, \"max_ms\": %.3f, \"calls\": %zu, \"failed\": %d, \"steps\": %llu, \"requests_per_step\": %.1f, "
         "\"frame_period_ms\": 26.122, \"audio_x_realtime\": %.0f, \"max_wait_us\": %d}\n",
         n, n, pct(0.50), pct(0.99), all.empty() ? 0.0 : all.back(), all.size(), failed.load(), (unsigned long long)steps,
         steps ? (double)reqs / (double)steps : 0.0, (double)n * chunks * 1152.0 / 44100.0 / wall, wait_us);
  mp3b_pool_destroy(pool);
  return failed.load() ? 2 : 0;
}
