// Stand-alone TMA probe (development aid): which tensor-map / coordinate variants the hardware accepts.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int bytes, unsigned char* out) {
  __shared__ alignas(128) unsigned char buf[8192];
  __shared__ alignas(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(buf)),
                   "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(buf)),
                   "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(c0), "r"(c1) : "memory");
  }
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  } while (!done);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no entry point\n"); return 2; }
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const int W = 640, H = 480, N = 2;
  std::vector<unsigned char> h((size_t)W * H * N);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(i * 7 + (i >> 8));
  unsigned char *d, *dout; cudaMalloc(&d, h.size()); cudaMalloc(&dout, 8192);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  CUtensorMap m; memset(&m, 0, sizeof m);
  int rank = 3, c0 = 0, c1 = 0, c2 = 0; cuuint32_t b0 = 144, b1 = 34;
  CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  switch (variant) {
    case 0: rank = 2; b0 = 128; b1 = 32; break;                 // plain 2D, aligned
    case 1: rank = 2; b0 = 144; b1 = 34; break;                 // 2D, 144-wide box
    case 2: rank = 2; c0 = -4; c1 = -1; break;                  // 2D negative coords
    case 3: rank = 3; break;                                    // 3D at origin
    case 4: rank = 3; c0 = 124; c1 = 31; c2 = 1; break;         // 3D interior, unaligned x
    case 5: rank = 3; c0 = -4; c1 = -1; break;                  // 3D negative
    case 6: rank = 3; c0 = 128; c1 = 32; c2 = 1; b0 = 128; b1 = 32; break;
    case 7: rank = 3; c0 = 0; c1 = -1; break;
    case 8: rank = 3; c0 = -16; c1 = -1; b0 = 160; break;
    case 9: rank = 3; c0 = 112; c1 = 31; c2 = 1; b0 = 160; break;
    case 10: rank = 3; c0 = 624; c1 = 479; c2 = 1; b0 = 160; break;   // runs off the right / bottom edge
  }
  cuuint64_t dims[3] = {W, H, N}; cuuint64_t strides[2] = {W, (cuuint64_t)W * H}; cuuint32_t box[3] = {b0, b1, 1}; cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d encode rc=%d\n", variant, (int)r);
  const int bytes = b0 * b1;
  if (rank == 3) k<3><<<1, 128>>>(m, c0, c1, c2, bytes, dout); else k<2><<<1, 128>>>(m, c0, c1, c2, bytes, dout);
  cudaError_t e = cudaDeviceSynchronize();
  printf("variant %d run: %s\n", variant, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<unsigned char> o(bytes); cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < (int)b1; ++y) for (int x = 0; x < (int)b0; ++x) {
      int gx = c0 + x, gy = c1 + y; unsigned char want = 0;
      if (gx >= 0 && gx < W && gy >= 0 && gy < H) want = h[(size_t)c2 * W * H + (size_t)gy * W + gx];
      bad += o[y * b0 + x] != want;
    }
    printf("variant %d mismatches=%d\n", variant, bad);
  }
  return 0;
}
