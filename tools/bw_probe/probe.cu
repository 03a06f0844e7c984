// Bandwidth probe for the undistortion kernel's traffic shape (no gather): per 4 output pixels read 16 B of map and
// 16 B of source, write 12 B.  Compares a linear sweep with the CTA tile geometries of k_remap_bgrx.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o probe probe.cu && ./probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int W = 4032, H = 3040, NF = 32;

__device__ __forceinline__ void do_quad(const uint4* __restrict__ map, const uint4* __restrict__ src, uint32_t* __restrict__ dst, size_t q, size_t fq, size_t fd) {
  const uint4 m = __ldg(map + q);
  const uint4 s = src[fq + q];
  uint32_t* d = dst + fd + q * 3;
  d[0] = s.x ^ m.x; d[1] = s.y ^ m.y; d[2] = s.z ^ (m.z + m.w + s.w);
}

// linear: one quad per thread, frames in the grid's z
__global__ void k_linear(const uint4* map, const uint4* src, uint32_t* dst) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (size_t)W * H / 4) return;
  do_quad(map, src, dst, q, (size_t)blockIdx.z * W * H / 4, (size_t)blockIdx.z * W * H * 3 / 4);
}

// tiles: CTA = TW x TH pixels; warp covers 128 pixels of a row; 8 warps arranged WX across x WY down; rows step WY
template <int TW, int TH>
__global__ void k_tile(const uint4* map, const uint4* src, uint32_t* dst) {
  constexpr int WX = TW / 128, WY = 8 / WX;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = blockIdx.x * TW + (warp % WX) * 128 + lane * 4;
  if (x >= W) return;
  const int y0 = blockIdx.y * TH + warp / WX, y1 = min(blockIdx.y * TH + TH, H);
  const size_t fq = (size_t)blockIdx.z * W * H / 4, fd = (size_t)blockIdx.z * W * H * 3 / 4;
#pragma unroll 1
  for (int y = y0; y < y1; y += WY) do_quad(map, src, dst, ((size_t)y * W + x) / 4, fq, fd);
}

// source only through the map-less path (what share does the map stream cost?)
__global__ void k_linear_nomap(const uint4* src, uint32_t* dst) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (size_t)W * H / 4) return;
  const uint4 s = src[(size_t)blockIdx.z * W * H / 4 + q];
  uint32_t* d = dst + (size_t)blockIdx.z * W * H * 3 / 4 + q * 3;
  d[0] = s.x; d[1] = s.y; d[2] = s.z ^ s.w;
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e9f;
  for (int i = 0; i < reps; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
  return best;
}

int main() {
  uint4 *map, *src; uint32_t* dst;
  const size_t px = (size_t)W * H;
  CK(cudaMalloc(&map, px * 4)); CK(cudaMalloc(&src, px * 4 * NF)); CK(cudaMalloc(&dst, px * 3 * NF));
  CK(cudaMemset(map, 1, px * 4)); CK(cudaMemset(src, 2, px * 4 * NF));
  const double bytes_map = (double)px * NF * 11, bytes_nomap = (double)px * NF * 7;
  auto report = [&](const char* name, float ms, double bytes) { printf("%-28s %7.3f ms  %7.1f GB/s (algorithmic)\n", name, ms, bytes / ms * 1e-6); };
  const dim3 gl((unsigned)((px / 4 + 255) / 256), 1, NF);
  report("linear", time_ms([&] { k_linear<<<gl, 256>>>(map, src, dst); }), bytes_map);
  report("linear, no map", time_ms([&] { k_linear_nomap<<<gl, 256>>>(src, dst); }), bytes_nomap);
  report("tile 128x64", time_ms([&] { k_tile<128, 64><<<dim3((W + 127) / 128, (H + 63) / 64, NF), 256>>>(map, src, dst); }), bytes_map);
  report("tile 256x32", time_ms([&] { k_tile<256, 32><<<dim3((W + 255) / 256, (H + 31) / 32, NF), 256>>>(map, src, dst); }), bytes_map);
  report("tile 512x16", time_ms([&] { k_tile<512, 16><<<dim3((W + 511) / 512, (H + 15) / 16, NF), 256>>>(map, src, dst); }), bytes_map);
  report("tile 1024x8", time_ms([&] { k_tile<1024, 8><<<dim3((W + 1023) / 1024, (H + 7) / 8, NF), 256>>>(map, src, dst); }), bytes_map);
  report("tile 128x16", time_ms([&] { k_tile<128, 16><<<dim3((W + 127) / 128, (H + 15) / 16, NF), 256>>>(map, src, dst); }), bytes_map);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  return 0;
}
