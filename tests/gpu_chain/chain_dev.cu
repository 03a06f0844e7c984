// TEST INFRASTRUCTURE ONLY (never linked into the product library): runs the strip kernel's four-pixel chain
// (chain_quad.cuh) and the generic kernels' chain_pixel() (pixel_math.cuh) ON THE DEVICE over the same pixels, tables in
// shared memory like the kernels keep them, and reports where they differ.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#include "../../raw_image_pipeline_b200/csrc/chain_tables.hpp"

using namespace rip;

struct DevArgs {
  const uint8_t* in; const float* mask; uint8_t* out_quad; uint8_t* out_ref; long n;
  const uint8_t* blob; const uint8_t* sblob; const float* wbf; ChainConsts k; int wbg;
};

template <uint32_t STAGES>
__global__ void k_chain_dev(const __grid_constant__ DevArgs A) {
  extern __shared__ uint8_t dyn_raw[];  // the strip tables must sit on a 4096-byte boundary of the shared window (rip_strip.cuh)
  uint8_t* const s_tab = dyn_raw + ((0u - (uint32_t)__cvta_generic_to_shared(dyn_raw)) & 4095u);
  __shared__ alignas(16) uint8_t s_blob[TABLE_BYTES];
  __shared__ float s_wbf[768];
  for (int i = threadIdx.x; i < STRIP_BLOB_BYTES; i += blockDim.x) s_tab[i] = A.sblob[i];
  for (int i = threadIdx.x; i < TABLE_BYTES; i += blockDim.x) s_blob[i] = A.blob[i];
  __syncthreads();
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {  // per-frame tables, as k_fused_strip builds them
    s_wbf[i] = A.wbf[i]; s_tab[SOFF_WB + i] = (uint8_t)__float2int_rz(A.wbf[i]);
    reinterpret_cast<float*>(s_tab + SOFF_WBF)[i] = A.wbf[i];
  }
  __syncthreads();
  const StripTables T = strip_tables_at(taddr_of_shared(s_tab));
  const ChainTables C = chain_tables_from_blob(s_blob, s_wbf);
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < A.n; q += (long)gridDim.x * blockDim.x) {
    uint32_t Bw = 0, Gw = 0, Rw = 0; float m[4];
    for (int j = 0; j < 4; ++j) {
      const long i = q * 4 + j;
      Bw |= (uint32_t)A.in[3 * i] << (8 * j); Gw |= (uint32_t)A.in[3 * i + 1] << (8 * j); Rw |= (uint32_t)A.in[3 * i + 2] << (8 * j);
      m[j] = A.mask ? A.mask[i] : 1.0f;
    }
    uint32_t px[4];
    if (A.wbg) chain_quad<STAGES, true>(Bw, Gw, Rw, m, A.k, T, px);
    else chain_quad<STAGES, false>(Bw, Gw, Rw, m, A.k, T, px);
    for (int j = 0; j < 4; ++j) {
      const long i = q * 4 + j;
      A.out_quad[3 * i] = px[j]; A.out_quad[3 * i + 1] = px[j] >> 8; A.out_quad[3 * i + 2] = px[j] >> 16;
      const uint32_t r = chain_pixel<STAGES>(A.in[3 * i], A.in[3 * i + 1], A.in[3 * i + 2], m[j], false, A.k, C);
      A.out_ref[3 * i] = r; A.out_ref[3 * i + 1] = r >> 8; A.out_ref[3 * i + 2] = r >> 16;
    }
  }
}

template <uint32_t S>
static int run_stage(uint32_t stages, const DevArgs& a) {
  if (stages == S) {
    const int smem = STRIP_TABLE_BYTES + 4096;
    if (cudaFuncSetAttribute(k_chain_dev<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -4;
    k_chain_dev<S><<<148 * 2, 256, smem>>>(a);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
  }
  if constexpr (S < 31) return run_stage<S + 1>(stages, a);
  return -3;
}

// in: n x 3 u8 (n % 4 == 0); out_quad / out_ref: n x 3 u8 (host).  Returns 0 on success.
extern "C" int chain_dev_run(unsigned stages, long n, const uint8_t* in, const float* mask, const float* cc, const double* enh,
                             const uint8_t* wb, const uint8_t* gamma, uint8_t* out_quad, uint8_t* out_ref) {
  ChainTableParams q;
  q.enh_gain[0] = enh[0]; q.enh_gain[1] = enh[1]; q.enh_gain[2] = enh[2];
  std::vector<uint8_t> blob(TABLE_BYTES), sblob(STRIP_BLOB_BYTES);
  build_chain_blob(q, blob.data());
  if (stages & ST_GAMMA) {
    memcpy(blob.data() + OFF_GAMMA, gamma, 256);
    uint16_t* g2 = reinterpret_cast<uint16_t*>(blob.data() + OFF_G2);
    for (int i = 0; i < 256; ++i) g2[i] = kSrgbGammaTab[gamma[i]];
  }
  build_strip_blob(blob.data(), sblob.data());
  float wbf[768]; bool gid = true;
  for (int i = 0; i < 768; ++i) wbf[i] = (float)wb[i];
  for (int i = 0; i < 256; ++i) gid = gid && wb[256 + i] == i;
  DevArgs a{};
  memcpy(a.k.cc, cc, sizeof a.k.cc); a.k.cc_bias[0] = a.k.cc_bias[1] = a.k.cc_bias[2] = 0.0f;
  a.k.wb_g_identity = gid ? 1 : 0; a.wbg = gid ? 0 : 1; a.n = n;
  chain_consts_finish(a.k);
  uint8_t *d_in, *d_q, *d_r, *d_blob, *d_sblob; float *d_mask = nullptr, *d_wbf;
  if (cudaMalloc(&d_in, 3 * n) || cudaMalloc(&d_q, 3 * n) || cudaMalloc(&d_r, 3 * n) || cudaMalloc(&d_blob, TABLE_BYTES) ||
      cudaMalloc(&d_sblob, STRIP_BLOB_BYTES) || cudaMalloc(&d_wbf, sizeof wbf)) return -1;
  if (mask) { if (cudaMalloc(&d_mask, 4 * n)) return -1; cudaMemcpy(d_mask, mask, 4 * n, cudaMemcpyHostToDevice); }
  cudaMemcpy(d_in, in, 3 * n, cudaMemcpyHostToDevice);
  cudaMemcpy(d_blob, blob.data(), TABLE_BYTES, cudaMemcpyHostToDevice);
  cudaMemcpy(d_sblob, sblob.data(), STRIP_BLOB_BYTES, cudaMemcpyHostToDevice);
  cudaMemcpy(d_wbf, wbf, sizeof wbf, cudaMemcpyHostToDevice);
  a.in = d_in; a.mask = d_mask; a.out_quad = d_q; a.out_ref = d_r; a.blob = d_blob; a.sblob = d_sblob; a.wbf = d_wbf;
  int rc = run_stage<0>(stages & 31u, a);
  cudaMemcpy(out_quad, d_q, 3 * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(out_ref, d_r, 3 * n, cudaMemcpyDeviceToHost);
  cudaFree(d_in); cudaFree(d_q); cudaFree(d_r); cudaFree(d_blob); cudaFree(d_sblob); cudaFree(d_wbf); if (d_mask) cudaFree(d_mask);
  return rc;
}
