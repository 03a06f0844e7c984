#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
  const size_t n = 6220800;
  uint8_t* pinned; cudaMallocHost((void**)&pinned, n); memset(pinned, 1, n);
  uint8_t* dst = (uint8_t*)malloc(n); memset(dst, 0, n);
  for (int threads : {1, 2, 4, 8}) {
    double best = 1e9;
    for (int rep = 0; rep < 20; ++rep) {
      double t0 = now();
      std::vector<std::thread> th;
      size_t chunk = n / threads;
      for (int i = 1; i < threads; ++i) th.emplace_back([=] { memcpy(dst + i * chunk, pinned + i * chunk, chunk); });
      memcpy(dst, pinned, chunk);
      for (auto& t : th) t.join();
      best = std::min(best, now() - t0);
    }
    printf("memcpy pinned->warm malloc, %d threads (spawned per call): %.0f us = %.1f GB/s\n", threads, best * 1e6, n / best / 1e9);
  }
  { double t = 0; for (int rep = 0; rep < 10; ++rep) { uint8_t* f = (uint8_t*)malloc(n); double t0 = now(); memcpy(f, pinned, n); t += now() - t0; free(f); }
    printf("memcpy pinned->fresh malloc (page faults): %.0f us\n", t / 10 * 1e6); }
  uint8_t* d; cudaMalloc((void**)&d, n); cudaStream_t s; cudaStreamCreate(&s);
  for (int k = 0; k < 2; ++k) {
    double t = 0; for (int rep = 0; rep < 20; ++rep) { double t0 = now(); cudaMemcpyAsync(dst, d, n, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); t += now() - t0; }
    printf("cudaMemcpy D2H 6.2MB to pageable: %.0f us\n", t / 20 * 1e6);
    t = 0; for (int rep = 0; rep < 20; ++rep) { double t0 = now(); cudaMemcpyAsync(pinned, d, n, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); t += now() - t0; }
    printf("cudaMemcpy D2H 6.2MB to pinned: %.0f us\n", t / 20 * 1e6);
    t = 0; for (int rep = 0; rep < 20; ++rep) { double t0 = now(); cudaMemcpyAsync(d, dst, n / 3, cudaMemcpyHostToDevice, s); cudaStreamSynchronize(s); t += now() - t0; }
    printf("cudaMemcpy H2D 2MB from pageable: %.0f us\n", t / 20 * 1e6);
  }
  return 0;
}
