// Host -> device copy ceiling of a box: plain pinned cudaMemcpyAsync (1-D) and the pitched 2-D form the engine's upload
// uses (engine.cc issue_h2d: cudaMemcpy2DAsync, one row per stream), at 1 / 2 / 4 / 8 concurrent GPUs, one host thread and one
// pinned buffer per GPU.  No kernels, no encode: this is the control for the end-to-end scaling of bench.py.
// build: nvcc -O2 -std=c++17 tools/h2d_bench.cu -o tools/h2d_bench     run: tools/h2d_bench [MiB per copy, default 1024] [repeats, default 6]
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Dev { int id; void *h = nullptr, *d = nullptr; cudaStream_t st; cudaEvent_t e0, e1; double gbs[3] = {0, 0, 0}; };

int main(int argc, char **argv) {
  const size_t mib = argc > 1 ? (size_t)atol(argv[1]) : 1024;
  const int reps = argc > 2 ? atoi(argv[2]) : 6;
  const size_t bytes = mib << 20;
  int ndev = 0; CK(cudaGetDeviceCount(&ndev));
  // pitched shape: 512 rows (streams) of `width` bytes, source pitch = width + 4096 (rows of a larger host arena), like a pass upload
  const size_t rows = 512, width = bytes / rows, spitch = width + 4096;
  std::vector<Dev> devs((size_t)ndev);
  for (int i = 0; i < ndev; ++i) {
    Dev &v = devs[(size_t)i]; v.id = i;
    CK(cudaSetDevice(i));
    CK(cudaHostAlloc(&v.h, rows * spitch, cudaHostAllocPortable));
    memset(v.h, 1, rows * spitch);                       // first touch by this thread
    CK(cudaMalloc(&v.d, bytes));
    CK(cudaStreamCreateWithFlags(&v.st, cudaStreamNonBlocking)); CK(cudaEventCreate(&v.e0)); CK(cudaEventCreate(&v.e1));
  }
  printf("{\"mib_per_copy\": %zu, \"repeats\": %d, \"devices\": %d, \"runs\": [\n", mib, reps, ndev);
  bool first = true;
  for (int n = 1; n <= ndev; n *= 2) {
    for (int mode = 0; mode < 3; ++mode) {               // 0: 1-D H2D, 1: pitched 2-D H2D, 2: 1-D D2H
      std::atomic<int> ready{0}; std::atomic<bool> go{false};
      std::vector<std::thread> th;
      double wall = 0;
      for (int i = 0; i < n; ++i) th.emplace_back([&, i] {
        Dev &v = devs[(size_t)i];
        CK(cudaSetDevice(v.id));
        auto copy = [&] {
          if (mode == 0) CK(cudaMemcpyAsync(v.d, v.h, bytes, cudaMemcpyHostToDevice, v.st));
          else if (mode == 1) CK(cudaMemcpy2DAsync(v.d, width, v.h, spitch, width, rows, cudaMemcpyHostToDevice, v.st));
          else CK(cudaMemcpyAsync(v.h, v.d, bytes, cudaMemcpyDeviceToHost, v.st));
        };
        copy(); CK(cudaStreamSynchronize(v.st));         // warm-up
        ready.fetch_add(1);
        while (!go.load()) std::this_thread::yield();
        CK(cudaEventRecord(v.e0, v.st));
        for (int r = 0; r < reps; ++r) copy();
        CK(cudaEventRecord(v.e1, v.st));
        CK(cudaEventSynchronize(v.e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, v.e0, v.e1));
        v.gbs[mode] = (double)bytes * reps / (ms * 1e-3) / 1e9;
      });
      while (ready.load() < n) std::this_thread::yield();
      auto t0 = std::chrono::steady_clock::now();
      go.store(true);
      for (auto &t : th) t.join();
      wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      double sum = 0, mn = 1e30;
      for (int i = 0; i < n; ++i) { sum += devs[(size_t)i].gbs[mode]; mn = std::min(mn, devs[(size_t)i].gbs[mode]); }
      printf("%s {\"gpus\": %d, \"mode\": \"%s\", \"aggregate_GBps_wall\": %.1f, \"sum_of_per_gpu_GBps\": %.1f, \"slowest_gpu_GBps\": %.1f}",
             first ? "" : ",\n", n, mode == 0 ? "h2d_1d" : mode == 1 ? "h2d_2d_pitched" : "d2h_1d", (double)bytes * reps * n / wall / 1e9, sum, mn);
      first = false;
    }
  }
  printf("\n]}\n");
  return 0;
}
