// Grow-only device allocation used for the pipeline's internal buffers.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace rip {

struct DevBuf {
  void* ptr = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&ptr, n);
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
  template <typename T> T* as() const { return static_cast<T*>(ptr); }
};

}  // namespace rip
