// Grow-only device allocation used for the pipeline's internal buffers.  Owns its memory: freed by the destructor
// (movable, not copyable), so a buffer added to the pipeline cannot be forgotten in rip_destroy.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace rip {

struct DevBuf {
  void* ptr = nullptr;
  size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : ptr(o.ptr), cap(o.cap) { o.ptr = nullptr; o.cap = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); ptr = o.ptr; cap = o.cap; o.ptr = nullptr; o.cap = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
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
