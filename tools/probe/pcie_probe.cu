// Host <-> device copy bandwidth of this box, per GPU and in aggregate: the
// ceiling of every end-to-end (host arrays in, host arrays out) number.
//
//   nvcc -O2 -o pcie_probe pcie_probe.cu -lpthread && ./pcie_probe [MiB]
//
// For n = 1, 2, 4, 8 (up to the visible devices) n host threads, one per
// device, each with its own pinned buffers, run H2D only, D2H only and both
// directions at once (two streams); one JSON line per (n, mode) with the
// aggregate GB/s and the slowest device's GB/s.  Buffers are first touched by
// the thread that uses them.
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#define CHECK(x)                                                         \
  do {                                                                   \
    cudaError_t e_ = (x);                                                \
    if (e_ != cudaSuccess) {                                             \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));           \
      exit(1);                                                           \
    }                                                                    \
  } while (0)

struct Worker {
  int device;
  size_t bytes;
  void *h_in, *h_out, *d_in, *d_out;
  cudaStream_t s_in, s_out;
  double seconds[3];
};

static pthread_barrier_t barrier;

static void run_mode(Worker* w, int mode, int reps) {
  CHECK(cudaSetDevice(w->device));
  pthread_barrier_wait(&barrier);
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < reps; ++r) {
    if (mode == 0 || mode == 2)
      CHECK(cudaMemcpyAsync(w->d_in, w->h_in, w->bytes, cudaMemcpyHostToDevice, w->s_in));
    if (mode == 1 || mode == 2)
      CHECK(cudaMemcpyAsync(w->h_out, w->d_out, w->bytes, cudaMemcpyDeviceToHost, w->s_out));
  }
  CHECK(cudaStreamSynchronize(w->s_in));
  CHECK(cudaStreamSynchronize(w->s_out));
  auto t1 = std::chrono::steady_clock::now();
  w->seconds[mode] = std::chrono::duration<double>(t1 - t0).count();
  pthread_barrier_wait(&barrier);
}

int main(int argc, char** argv) {
  size_t mib = argc > 1 ? atoi(argv[1]) : 512;
  int reps = 4;
  int devices = 0;
  CHECK(cudaGetDeviceCount(&devices));
  const char* modes[] = {"h2d", "d2h", "both"};
  std::vector<Worker> workers(devices);
  for (int d = 0; d < devices; ++d) {
    Worker& w = workers[d];
    w.device = d;
    w.bytes = mib << 20;
    CHECK(cudaSetDevice(d));
    CHECK(cudaMalloc(&w.d_in, w.bytes));
    CHECK(cudaMalloc(&w.d_out, w.bytes));
    CHECK(cudaStreamCreateWithFlags(&w.s_in, cudaStreamNonBlocking));
    CHECK(cudaStreamCreateWithFlags(&w.s_out, cudaStreamNonBlocking));
  }
  {
    std::vector<std::thread> threads;
    for (int d = 0; d < devices; ++d)
      threads.emplace_back([&, d] {
        Worker& w = workers[d];
        CHECK(cudaSetDevice(d));
        CHECK(cudaHostAlloc(&w.h_in, w.bytes, cudaHostAllocDefault));
        CHECK(cudaHostAlloc(&w.h_out, w.bytes, cudaHostAllocDefault));
        memset(w.h_in, 1, w.bytes);
        memset(w.h_out, 2, w.bytes);
      });
    for (auto& t : threads) t.join();
  }
  for (int n = 1; n <= devices; n *= 2) {
    for (int mode = 0; mode < 3; ++mode) {
      pthread_barrier_init(&barrier, nullptr, n);
      std::vector<std::thread> threads;
      for (int d = 0; d < n; ++d)
        threads.emplace_back([&, d, mode] {
          run_mode(&workers[d], mode, 1);  // warm-up
          run_mode(&workers[d], mode, reps);
        });
      for (auto& t : threads) t.join();
      pthread_barrier_destroy(&barrier);
      double slowest = 0;
      for (int d = 0; d < n; ++d)
        if (workers[d].seconds[mode] > slowest) slowest = workers[d].seconds[mode];
      const double per_dir = double(mib << 20) * reps / 1e9;
      const double dirs = mode == 2 ? 2.0 : 1.0;
      printf("{\"gpus\": %d, \"mode\": \"%s\", \"mib\": %zu, \"aggregate_gbs\": %.1f, "
             "\"per_gpu_gbs\": %.1f, \"seconds\": %.4f}\n",
             n, modes[mode], mib, per_dir * dirs * n / slowest,
             per_dir * dirs / slowest, slowest);
      fflush(stdout);
    }
  }
  return 0;
}
