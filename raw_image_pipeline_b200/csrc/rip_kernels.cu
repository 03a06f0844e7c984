// sm_100a kernels of the RAW chain: the generic family (any shape, rotation, alignment, 1- and 3-channel inputs),
// the white-balance LUT builders and the undistortion gathers.  The TMA-fed fast family for camera-shaped Bayer frames
// lives in rip_fast.cu; both run the same per-pixel code (pixel_math.cuh) and produce identical bytes.
//
//   k_fused<STAGES>   debayer -> flip -> [WB LUT] -> [colour calibration] -> [gamma LUT] ->
//                     [Lab vignetting] -> [HSV enhancer] -> BGR8, one pass, everything after the
//                     Bayer tile load in registers                (raw_image_pipeline.hpp:143-166)
//   k_pca_stats       whole-frame sums / maxima of the debayered frame   (white_balance.cpp:89-102)
//   k_pca_lut         2x2 solve + per-frame 256-entry LUTs               (white_balance.cpp:105-127)
//   k_gain_lut        per-frame gain LUTs (ccc)                          (ccc.cpp:383-386)
//   k_remap<CH>       cv::remap fixed-point bilinear gather               (undistortion.cpp:240-245)
//   k_remap_bgrx      the same gather from the fast path's 4-byte intermediate, float or packed fixed-point map
//   k_mono            1-channel non-Bayer passthrough (flip + gamma)
//
// Work decomposition of k_fused / k_pca_stats: a frame is cut into 128x32-pixel tiles; a persistent grid of 256-thread
// CTAs walks the tile list of the whole batch (frame-major).  Per tile: stage the Bayer tile
// (+1 pixel halo, 144 B x 34 rows) in shared memory, each thread demosaics 4 horizontally
// adjacent pixels per row from three packed 32-bit words per row, runs the chain on them in
// registers, and the BGR8 result is staged in shared memory so that global stores are full
// 16-byte vectors regardless of the 3-byte pixel size.
#include <cuda_runtime.h>
#include <stdint.h>

#include "frame_math.cuh"
#include "kernels.hpp"

namespace rip {

constexpr int TW = 128;             // tile width  (pixels) = 32 lanes x 4 px
constexpr int TH = 32;              // tile height (pixels) = 8 warps x 4 rows
constexpr int NTHREADS = 256;
constexpr int IN_WORDS = 36;        // 144 B per staged Bayer row: x0-4 .. x0+139
constexpr int IN_ROWS = TH + 2;     // y0-1 .. y0+TH
constexpr int OUT_ROW_BYTES = TW * 3;

struct __align__(16) SmemFused {
  uint8_t tables[TABLE_BYTES];
  float wbf[768];
  uint32_t in[IN_ROWS * IN_WORDS];
  uint8_t out[TH * OUT_ROW_BYTES];
};

// destination (post-flip) coordinate of input pixel (iy, ix): inverse of flip_source()
__device__ __forceinline__ void flip_dest(int angle, int rows, int cols, int iy, int ix, int& oy, int& ox) {
  if (angle == 90) { oy = ix; ox = rows - 1 - iy; }
  else if (angle == 180) { oy = rows - 1 - iy; ox = cols - 1 - ix; }
  else if (angle == 270) { oy = cols - 1 - ix; ox = iy; }
  else { oy = iy; ox = ix; }
}

// ---- stage the Bayer tile -------------------------------------------------------------------
__device__ __forceinline__ void stage_bayer_tile(uint32_t* s_in, const uint8_t* fin, int pitch, int rows, int cols,
                                                 int y0, int x0, bool aligned4) {
  for (int i = threadIdx.x; i < IN_ROWS * IN_WORDS; i += NTHREADS) {
    const int rr = i / IN_WORDS, wi = i - rr * IN_WORDS;
    const int y = y0 - 1 + rr, x = x0 - 4 + 4 * wi;
    uint32_t v = 0;
    if (y >= 0 && y < rows && x + 3 >= 0 && x < cols) {
      const uint8_t* p = fin + (size_t)y * pitch + x;
      if (aligned4 && x >= 0 && x + 3 < cols) {
        v = __ldg(reinterpret_cast<const uint32_t*>(p));
      } else {
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (x + b >= 0 && x + b < cols) v |= (uint32_t)__ldg(p + b) << (8 * b);
      }
    }
    s_in[i] = v;
  }
}

// ---- fetch the four (b,g,r) triples a thread owns: pixels (y, x..x+3) of the input frame ------
template <int SRC>
__device__ __forceinline__ void fetch_quad(const FrameParams& P, const uint8_t* fin, const uint32_t* s_in, int y0, int y,
                                           int x, int lane, bool aligned4, int b[4], int g[4], int r[4]) {
  if (SRC == SRC_BAYER) {
    const int yc = y < 1 ? 1 : (y > P.rows - 2 ? P.rows - 2 : y);
    const int tr = yc - y0 + 1;  // staged row index of yc
    if (x >= 4 && x + 4 <= P.cols - 1 && tr >= 1 && tr <= TH) {
      uint32_t w[3][3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int j = 0; j < 3; ++j) w[a][j] = s_in[(tr - 1 + a) * IN_WORDS + lane + j];
      const bool row_has_r = ((yc & 1) == ((P.cfa >> 1) & 1));
      const int cpar = row_has_r ? (P.cfa & 1) : ((P.cfa & 1) ^ 1);
      demosaic_quad(w, row_has_r, cpar, b, g, r);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        b[k] = g[k] = r[k] = 0;
        if (x + k < P.cols) demosaic_at(fin, P.rows, P.cols, (size_t)P.in_pitch, y, x + k, P.cfa, b[k], g[k], r[k]);
      }
    }
  } else {
    const uint8_t* p = fin + (size_t)y * P.in_pitch + 3 * x;
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    if (aligned4 && x + 3 < P.cols) {
      const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
      w0 = __ldg(q); w1 = __ldg(q + 1); w2 = __ldg(q + 2);
    } else {
      const int nb = 3 * min(4, P.cols - x);
      for (int i = 0; i < nb; ++i) {
        const uint32_t v = __ldg(p + i);
        if (i < 4) w0 |= v << (8 * i);
        else if (i < 8) w1 |= v << (8 * (i - 4));
        else w2 |= v << (8 * (i - 8));
      }
    }
    const int c0[4] = {(int)(w0 & 255), (int)(w0 >> 24), (int)((w1 >> 16) & 255), (int)((w2 >> 8) & 255)};
    const int c1[4] = {(int)((w0 >> 8) & 255), (int)(w1 & 255), (int)(w1 >> 24), (int)((w2 >> 16) & 255)};
    const int c2[4] = {(int)((w0 >> 16) & 255), (int)((w1 >> 8) & 255), (int)(w2 & 255), (int)(w2 >> 24)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      g[k] = c1[k];
      if (SRC == SRC_RGB) { r[k] = c0[k]; b[k] = c2[k]; }  // debayer.cpp:72-73 COLOR_RGB2BGR
      else { b[k] = c0[k]; r[k] = c2[k]; }
    }
  }
}

// ---- copy one staged output row segment to global with the widest congruent vector -----------
__device__ __forceinline__ void copy_row(uint8_t* dst, const uint8_t* src, int nbytes, int lane) {
  const uintptr_t d = reinterpret_cast<uintptr_t>(dst);
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(src);
  if (((d ^ s) & 15) == 0) {
    int head = (int)((16 - (d & 15)) & 15);
    head = head < nbytes ? head : nbytes;
    if (lane < head) dst[lane] = src[lane];
    const int n16 = (nbytes - head) >> 4;
    const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    for (int i = lane; i < n16; i += 32) d4[i] = s4[i];
    const int done = head + (n16 << 4);
    if (lane < nbytes - done) dst[done + lane] = src[done + lane];
  } else if (((d ^ s) & 3) == 0) {
    int head = (int)((4 - (d & 3)) & 3);
    head = head < nbytes ? head : nbytes;
    if (lane < head) dst[lane] = src[lane];
    const int n4 = (nbytes - head) >> 2;
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src + head);
    uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + head);
    for (int i = lane; i < n4; i += 32) d4[i] = s4[i];
    const int done = head + (n4 << 2);
    if (lane < nbytes - done) dst[done + lane] = src[done + lane];
  } else {
    for (int i = lane; i < nbytes; i += 32) dst[i] = src[i];
  }
}

// vignetting-mask values and cv2 row-tail flags of the four pixels (y, x..x+3) a thread owns
template <uint32_t STAGES>
__device__ __forceinline__ void quad_position_inputs(const FrameParams& P, int y, int x, float m[4], bool tail[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) { m[k] = 1.0f; tail[k] = false; }
  if (STAGES & ST_VIG) {  // the mask table is stored in input-frame coordinates
    const float* mrow = P.vig + (size_t)y * P.vig_pitch + x;
    if (x + 3 < P.cols && ((P.vig_pitch | x) & 3) == 0) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(mrow));
      m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (x + k < P.cols) m[k] = __ldg(mrow + k);
    }
  }
  if (STAGES & ST_ENH) {
    const int tail_start = P.ocols & ~31;  // cv2's scalar row tail in HSV2BGR (pixel_math.cuh)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int oy, ox;
      flip_dest(P.angle, P.rows, P.cols, y, min(x + k, P.cols - 1), oy, ox);
      tail[k] = ox >= tail_start;
    }
  }
}

// =============================================================================================
// fused kernel
// =============================================================================================
template <uint32_t STAGES, int SRC>
__global__ void __launch_bounds__(NTHREADS) k_fused(const __grid_constant__ FrameParams P) {
  __shared__ SmemFused sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_x = (P.cols + TW - 1) / TW, tiles_y = (P.rows + TH - 1) / TH;
  const long long tiles_per_frame = (long long)tiles_x * tiles_y;
  const long long total = tiles_per_frame * P.n_frames;

  if (STAGES & (ST_GAMMA | ST_VIG | ST_ENH)) {
    const uint4* src = reinterpret_cast<const uint4*>(P.tables);
    uint4* dst = reinterpret_cast<uint4*>(sm.tables);
    for (int i = threadIdx.x; i < TABLE_BYTES / 16; i += NTHREADS) dst[i] = __ldg(src + i);
  }
  const ChainTables T = chain_tables_from_blob(sm.tables, sm.wbf);
  const bool in_aligned4 = ((reinterpret_cast<uintptr_t>(P.in) | (uintptr_t)P.in_pitch | (uintptr_t)P.in_frame_stride) & 3) == 0;
  const bool staged_out = (P.angle == 0 || P.angle == 180);
  int cur_frame = -1;

  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int frame = (int)(t / tiles_per_frame);
    const int rem = (int)(t - (long long)frame * tiles_per_frame);
    const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
    const int x0 = tx * TW, y0 = ty * TH;
    const uint8_t* fin = P.in + (long long)frame * P.in_frame_stride;
    uint8_t* fout = P.out + (long long)frame * P.out_frame_stride;

    __syncthreads();  // previous tile fully copied out / tables visible
    if ((STAGES & ST_WB) && frame != cur_frame) {
      const float* src = P.wbf + (size_t)frame * 768;
      for (int i = threadIdx.x; i < 768; i += NTHREADS) sm.wbf[i] = src[i];  // plain load: written by a prior kernel
      cur_frame = frame;
    }
    if (SRC == SRC_BAYER) stage_bayer_tile(sm.in, fin, P.in_pitch, P.rows, P.cols, y0, x0, in_aligned4);
    __syncthreads();

    const int x = x0 + 4 * lane;
#pragma unroll 1
    for (int rr = 0; rr < TH / 8; ++rr) {
      const int r_in_tile = warp + 8 * rr;
      const int y = y0 + r_in_tile;
      if (y >= P.rows || x >= P.cols) continue;
      int b[4], g[4], r[4];
      fetch_quad<SRC>(P, fin, sm.in, y0, y, x, lane, in_aligned4, b, g, r);
      float m[4];
      bool tail[4];
      quad_position_inputs<STAGES>(P, y, x, m, tail);
      uint32_t px[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) px[k] = chain_pixel<STAGES>(b[k], g[k], r[k], m[k], tail[k], P.k, T);
      if (staged_out) {
        if (P.angle == 0) {
          uint32_t* o = reinterpret_cast<uint32_t*>(sm.out + r_in_tile * OUT_ROW_BYTES + 12 * lane);
          o[0] = prmt(px[0], px[1], 0x4210); o[1] = prmt(px[1], px[2], 0x5421); o[2] = prmt(px[2], px[3], 0x6542);
        } else {  // 180: mirrored inside the tile, pixel order reversed
          uint32_t* o = reinterpret_cast<uint32_t*>(sm.out + (TH - 1 - r_in_tile) * OUT_ROW_BYTES + 12 * (31 - lane));
          o[0] = prmt(px[3], px[2], 0x4210); o[1] = prmt(px[2], px[1], 0x5421); o[2] = prmt(px[1], px[0], 0x6542);
        }
      } else {  // 90 / 270: scattered byte stores (rare mode; correctness path)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (x + k >= P.cols) break;
          int oy, ox;
          flip_dest(P.angle, P.rows, P.cols, y, x + k, oy, ox);
          uint8_t* o = fout + (size_t)oy * P.out_pitch + 3 * ox;
          o[0] = (uint8_t)px[k]; o[1] = (uint8_t)(px[k] >> 8); o[2] = (uint8_t)(px[k] >> 16);
        }
      }
    }

    if (staged_out) {
      __syncthreads();
      const int oy0 = (P.angle == 0) ? y0 : P.rows - y0 - TH;
      const int ox0 = (P.angle == 0) ? x0 : P.cols - x0 - TW;
      const int cbeg = ox0 < 0 ? -ox0 : 0;
      const int cend = min(TW, P.ocols - ox0);
      for (int rr = warp; rr < TH; rr += NTHREADS / 32) {
        const int gy = oy0 + rr;
        if (gy < 0 || gy >= P.orows) continue;
        copy_row(fout + (size_t)gy * P.out_pitch + 3 * (ox0 + cbeg), sm.out + rr * OUT_ROW_BYTES + 3 * cbeg,
                 3 * (cend - cbeg), lane);
      }
    }
  }
}

// =============================================================================================
// PCA white-balance statistics over the debayered frame (flip does not change sums / maxima)
// =============================================================================================
template <int SRC>
__global__ void __launch_bounds__(NTHREADS) k_pca_stats(const __grid_constant__ FrameParams P) {
  __shared__ __align__(16) uint32_t s_in[IN_ROWS * IN_WORDS];
  __shared__ unsigned long long s_acc[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_x = (P.cols + TW - 1) / TW, tiles_y = (P.rows + TH - 1) / TH;
  const long long tiles_per_frame = (long long)tiles_x * tiles_y;
  const long long total = tiles_per_frame * P.n_frames;
  const bool in_aligned4 = ((reinterpret_cast<uintptr_t>(P.in) | (uintptr_t)P.in_pitch | (uintptr_t)P.in_frame_stride) & 3) == 0;
  if (threadIdx.x < 8) s_acc[threadIdx.x] = 0;
  int cur_frame = -1;

  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int frame = (int)(t / tiles_per_frame);
    const int rem = (int)(t - (long long)frame * tiles_per_frame);
    const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
    const int x0 = tx * TW, y0 = ty * TH;
    const uint8_t* fin = P.in + (long long)frame * P.in_frame_stride;
    __syncthreads();
    if (frame != cur_frame) {
      if (cur_frame >= 0 && threadIdx.x < 8) {
        unsigned long long* dst = P.stats + (size_t)cur_frame * 8 + threadIdx.x;
        if (threadIdx.x < 5) atomicAdd(dst, s_acc[threadIdx.x]); else atomicMax(dst, s_acc[threadIdx.x]);
        s_acc[threadIdx.x] = 0;
      }
      cur_frame = frame;
    }
    if (SRC == SRC_BAYER) stage_bayer_tile(s_in, fin, P.in_pitch, P.rows, P.cols, y0, x0, in_aligned4);
    __syncthreads();
    const int x = x0 + 4 * lane;
    unsigned sb = 0, sb2 = 0, sr = 0, sr2 = 0, sg = 0, mb = 0, mg = 0, mr = 0;
#pragma unroll 1
    for (int rr = 0; rr < TH / 8; ++rr) {
      const int y = y0 + warp + 8 * rr;
      if (y >= P.rows || x >= P.cols) continue;
      int b[4], g[4], r[4];
      fetch_quad<SRC>(P, fin, s_in, y0, y, x, lane, in_aligned4, b, g, r);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (x + k < P.cols) {
          sb += b[k]; sb2 += b[k] * b[k]; sr += r[k]; sr2 += r[k] * r[k]; sg += g[k];
          mb = max(mb, (unsigned)b[k]); mg = max(mg, (unsigned)g[k]); mr = max(mr, (unsigned)r[k]);
        }
      }
    }
    // per-warp-tile totals fit 32 bits: 512 px * 65025 < 2^26
    sb = __reduce_add_sync(0xffffffffu, sb); sb2 = __reduce_add_sync(0xffffffffu, sb2);
    sr = __reduce_add_sync(0xffffffffu, sr); sr2 = __reduce_add_sync(0xffffffffu, sr2);
    sg = __reduce_add_sync(0xffffffffu, sg);
    mb = __reduce_max_sync(0xffffffffu, mb); mg = __reduce_max_sync(0xffffffffu, mg); mr = __reduce_max_sync(0xffffffffu, mr);
    if (lane == 0) {
      atomicAdd(&s_acc[0], (unsigned long long)sb); atomicAdd(&s_acc[1], (unsigned long long)sb2);
      atomicAdd(&s_acc[2], (unsigned long long)sr); atomicAdd(&s_acc[3], (unsigned long long)sr2);
      atomicAdd(&s_acc[4], (unsigned long long)sg);
      atomicMax(&s_acc[5], (unsigned long long)mb); atomicMax(&s_acc[6], (unsigned long long)mg);
      atomicMax(&s_acc[7], (unsigned long long)mr);
    }
  }
  __syncthreads();
  if (cur_frame >= 0 && threadIdx.x < 8) {
    unsigned long long* dst = P.stats + (size_t)cur_frame * 8 + threadIdx.x;
    if (threadIdx.x < 5) atomicAdd(dst, s_acc[threadIdx.x]); else atomicMax(dst, s_acc[threadIdx.x]);
  }
}

// one CTA per frame: thread x builds LUT entry x for B and R (G stays identity); entries are stored as
// floats (exact integers) because their consumer is the fp32 colour-calibration mix
__global__ void __launch_bounds__(256) k_pca_lut(const unsigned long long* __restrict__ stats, float* __restrict__ wbf,
                                                float* __restrict__ coeff_out) {
  const int frame = blockIdx.x, x = threadIdx.x;
  const PcaCoeff c = pca_coefficients(stats + (size_t)frame * 8);
  float* lut = wbf + (size_t)frame * 768;
  lut[x] = (float)pca_lut_entry(x, c.alpha_b, c.beta_b);
  lut[256 + x] = (float)x;
  lut[512 + x] = (float)pca_lut_entry(x, c.alpha_r, c.beta_r);
  if (coeff_out && x == 0) {
    float* o = coeff_out + (size_t)frame * 4;
    o[0] = c.alpha_b; o[1] = c.beta_b; o[2] = c.alpha_r; o[3] = c.beta_r;
  }
}

__global__ void __launch_bounds__(256) k_gain_lut(const float* __restrict__ gains_bgr, float* __restrict__ wbf) {
  const int frame = blockIdx.x, x = threadIdx.x;
  float* lut = wbf + (size_t)frame * 768;
#pragma unroll
  for (int c = 0; c < 3; ++c) lut[256 * c + x] = (float)gain_lut_entry(x, gains_bgr[(size_t)frame * 3 + c]);
}

// =============================================================================================
// undistortion: fixed-point bilinear gather, 4 output pixels per thread
// =============================================================================================
template <int CH>
__global__ void __launch_bounds__(256) k_remap(const __grid_constant__ RemapParams P) {
  const int groups_x = (P.ocols + 3) >> 2;
  const long long per_frame = (long long)groups_x * P.orows;
  const long long total = per_frame * P.n_frames;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int frame = (int)(i / per_frame);
    const long long rem = i - (long long)frame * per_frame;
    const int y = (int)(rem / groups_x), x = ((int)(rem - (long long)y * groups_x)) << 2;
    const uint8_t* src = P.src + (long long)frame * P.src_frame_stride;
    uint8_t* dst = P.dst + (long long)frame * P.dst_frame_stride + (size_t)y * P.dpitch + (size_t)x * CH;
    const float2* mp = P.map + (size_t)y * P.ocols + x;
    const int nv = min(4, P.ocols - x);
    int o[4][CH];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < nv) {
        const float2 m = __ldg(mp + k);
        remap_pixel<CH>(src, P.rows, P.cols, (size_t)P.pitch, m.x, m.y, o[k]);
      } else {
#pragma unroll
        for (int c = 0; c < CH; ++c) o[k][c] = 0;
      }
    }
    if (nv == 4 && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) {
      if (CH == 3) {
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        d[0] = (uint32_t)o[0][0] | ((uint32_t)o[0][1] << 8) | ((uint32_t)o[0][2] << 16) | ((uint32_t)o[1][0] << 24);
        d[1] = (uint32_t)o[1][1] | ((uint32_t)o[1][2] << 8) | ((uint32_t)o[2][0] << 16) | ((uint32_t)o[2][1] << 24);
        d[2] = (uint32_t)o[2][2] | ((uint32_t)o[3][0] << 8) | ((uint32_t)o[3][1] << 16) | ((uint32_t)o[3][2] << 24);
      } else {
        *reinterpret_cast<uint32_t*>(dst) =
            (uint32_t)o[0][0] | ((uint32_t)o[1][0] << 8) | ((uint32_t)o[2][0] << 16) | ((uint32_t)o[3][0] << 24);
      }
    } else {
      for (int k = 0; k < nv; ++k)
#pragma unroll
        for (int c = 0; c < CH; ++c) dst[k * CH + c] = (uint8_t)o[k][c];
    }
  }
}

// undistortion from the fused kernel's 4-byte intermediate: 4 consecutive output pixels per thread and row
// (neighbouring pixels share taps, which the L1 serves), one 32-bit load per tap.  A lane-adjacent pixel assignment with a shared-
// memory exchange for the stores was measured slower (3.05 vs 2.37 ms per 64 x 12 MP).
template <bool PACKED>
__global__ void __launch_bounds__(256) k_remap_bgrx(const __grid_constant__ RemapParams P) {
  // a CTA covers 128 x 64 output pixels: warp w walks rows w, w+8, ... of the tile (the per-thread set-up is
  // amortised over 32 pixels; vertically adjacent outputs of one CTA share source rows in L1)
  const int x = (blockIdx.x * 32 + (threadIdx.x & 31)) << 2, frame = blockIdx.z;
  if (x >= P.ocols) return;
  const int pitch_px = P.pitch >> 2;
  const bool vec_ok = (P.ocols & 3) == 0 &&
                      ((reinterpret_cast<uintptr_t>(P.map) | reinterpret_cast<uintptr_t>(P.pmap) | reinterpret_cast<uintptr_t>(P.dst)) & 15) == 0 &&
                      (P.dst_frame_stride & 3) == 0;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(P.src + (long long)frame * P.src_frame_stride);
  const int y_begin = blockIdx.y * 64 + (threadIdx.x >> 5);
  const int y_end = min(blockIdx.y * 64 + 64, P.orows);
  uint8_t* dst = P.dst + (long long)frame * P.dst_frame_stride + (size_t)y_begin * P.dpitch + (size_t)x * 3;
  size_t moff = (size_t)y_begin * P.ocols + x;
#pragma unroll 1
  for (int y = y_begin; y < y_end; y += 8, dst += (size_t)8 * P.dpitch, moff += (size_t)8 * P.ocols) {
    if (vec_ok) {
      uint32_t p0, p1, p2, p3;
      if (PACKED) {
        const uint4 m = __ldg(reinterpret_cast<const uint4*>(P.pmap + moff));
        p0 = remap_pixel_bgrx_packed(src, P.rows, P.cols, pitch_px, m.x, x, y);
        p1 = remap_pixel_bgrx_packed(src, P.rows, P.cols, pitch_px, m.y, x + 1, y);
        p2 = remap_pixel_bgrx_packed(src, P.rows, P.cols, pitch_px, m.z, x + 2, y);
        p3 = remap_pixel_bgrx_packed(src, P.rows, P.cols, pitch_px, m.w, x + 3, y);
      } else {
        const float4* mp = reinterpret_cast<const float4*>(P.map + moff);
        const float4 m01 = __ldg(mp), m23 = __ldg(mp + 1);
        p0 = remap_pixel_bgrx(src, P.rows, P.cols, pitch_px, m01.x, m01.y);
        p1 = remap_pixel_bgrx(src, P.rows, P.cols, pitch_px, m01.z, m01.w);
        p2 = remap_pixel_bgrx(src, P.rows, P.cols, pitch_px, m23.x, m23.y);
        p3 = remap_pixel_bgrx(src, P.rows, P.cols, pitch_px, m23.z, m23.w);
      }
      uint32_t* d = reinterpret_cast<uint32_t*>(dst);
      d[0] = prmt(p0, p1, 0x4210); d[1] = prmt(p1, p2, 0x5421); d[2] = prmt(p2, p3, 0x6542);
    } else {
      const int nv = min(4, P.ocols - x);
      for (int k = 0; k < nv; ++k) {
        uint32_t px;
        if (PACKED) {
          px = remap_pixel_bgrx_packed(src, P.rows, P.cols, pitch_px, __ldg(P.pmap + moff + k), x + k, y);
        } else {
          const float2 m = __ldg(P.map + moff + k);
          px = remap_pixel_bgrx(src, P.rows, P.cols, pitch_px, m.x, m.y);
        }
        dst[3 * k] = (uint8_t)px; dst[3 * k + 1] = (uint8_t)(px >> 8); dst[3 * k + 2] = (uint8_t)(px >> 16);
      }
    }
  }
}

// 1-channel passthrough: out(oy, ox) = gamma[in(flip_source(oy, ox))]
__global__ void __launch_bounds__(256) k_mono(const __grid_constant__ FrameParams P, int use_gamma) {
  const long long per_frame = (long long)P.orows * P.ocols;
  const long long total = per_frame * P.n_frames;
  const uint8_t* gamma = P.tables + OFF_GAMMA;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int frame = (int)(i / per_frame);
    const long long rem = i - (long long)frame * per_frame;
    const int oy = (int)(rem / P.ocols), ox = (int)(rem - (long long)oy * P.ocols);
    int iy, ix;
    flip_source(P.angle, P.rows, P.cols, oy, ox, iy, ix);
    int v = P.in[(long long)frame * P.in_frame_stride + (size_t)iy * P.in_pitch + ix];
    if (use_gamma) v = __ldg(gamma + v);
    P.out[(long long)frame * P.out_frame_stride + (size_t)oy * P.out_pitch + ox] = (uint8_t)v;
  }
}

// =============================================================================================
// launchers
// =============================================================================================
static int grid_for(const FrameParams& p, int sm_count, int ctas_per_sm) {
  const long long tiles = (long long)((p.cols + TW - 1) / TW) * ((p.rows + TH - 1) / TH) * p.n_frames;
  const long long cap = (long long)sm_count * ctas_per_sm;
  return (int)(tiles < cap ? tiles : cap);
}

template <uint32_t STAGES>
static cudaError_t launch_fused_src(const FrameParams& p, int sm_count, cudaStream_t stream) {
  int occ = 0;
  const void* fn = p.src == SRC_BAYER ? (const void*)k_fused<STAGES, SRC_BAYER>
                 : p.src == SRC_BGR ? (const void*)k_fused<STAGES, SRC_BGR> : (const void*)k_fused<STAGES, SRC_RGB>;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, NTHREADS, 0);
  if (e != cudaSuccess) return e;
  if (occ < 1) occ = 1;
  const int grid = grid_for(p, sm_count, occ);
  if (grid <= 0) return cudaSuccess;
  if (p.src == SRC_BAYER) k_fused<STAGES, SRC_BAYER><<<grid, NTHREADS, 0, stream>>>(p);
  else if (p.src == SRC_BGR) k_fused<STAGES, SRC_BGR><<<grid, NTHREADS, 0, stream>>>(p);
  else k_fused<STAGES, SRC_RGB><<<grid, NTHREADS, 0, stream>>>(p);
  return cudaGetLastError();
}

template <uint32_t S>
static cudaError_t dispatch_stages(uint32_t stages, const FrameParams& p, int sm_count, cudaStream_t stream) {
  if (stages == S) return launch_fused_src<S>(p, sm_count, stream);
  if constexpr (S < ST_ALL) return dispatch_stages<S + 1>(stages, p, sm_count, stream);
  return cudaErrorInvalidValue;
}

cudaError_t launch_fused(uint32_t stages, const FrameParams& p, int sm_count, cudaStream_t stream, int* launches) {
  if (launches) ++*launches;
  return dispatch_stages<0>(stages & ST_ALL, p, sm_count, stream);
}

cudaError_t launch_pca_stats(const FrameParams& p, int sm_count, cudaStream_t stream, int* launches) {
  cudaError_t e = cudaMemsetAsync(p.stats, 0, sizeof(unsigned long long) * 8 * p.n_frames, stream);
  if (e != cudaSuccess) return e;
  const int grid = grid_for(p, sm_count, 4);
  if (grid <= 0) return cudaSuccess;
  if (launches) ++*launches;
  if (p.src == SRC_BAYER) k_pca_stats<SRC_BAYER><<<grid, NTHREADS, 0, stream>>>(p);
  else if (p.src == SRC_BGR) k_pca_stats<SRC_BGR><<<grid, NTHREADS, 0, stream>>>(p);
  else k_pca_stats<SRC_RGB><<<grid, NTHREADS, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_pca_lut(const unsigned long long* stats, float* wbf, float* coeff_out, int n_frames,
                           cudaStream_t stream, int* launches) {
  if (n_frames <= 0) return cudaSuccess;
  if (launches) ++*launches;
  k_pca_lut<<<n_frames, 256, 0, stream>>>(stats, wbf, coeff_out);
  return cudaGetLastError();
}

cudaError_t launch_gain_lut(const float* gains_bgr, float* wbf, int n_frames, cudaStream_t stream, int* launches) {
  if (n_frames <= 0) return cudaSuccess;
  if (launches) ++*launches;
  k_gain_lut<<<n_frames, 256, 0, stream>>>(gains_bgr, wbf);
  return cudaGetLastError();
}

cudaError_t launch_mono(const FrameParams& p, bool gamma, cudaStream_t stream, int* launches) {
  const long long total = (long long)p.orows * p.ocols * p.n_frames;
  if (total <= 0) return cudaSuccess;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  if (launches) ++*launches;
  k_mono<<<(int)blocks, 256, 0, stream>>>(p, gamma ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launch_remap(int channels, const RemapParams& p, cudaStream_t stream, int* launches) {
  const long long total = (long long)((p.ocols + 3) >> 2) * p.orows * p.n_frames;
  if (total <= 0) return cudaSuccess;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  if (launches) ++*launches;
  if (channels == 3) k_remap<3><<<(int)blocks, 256, 0, stream>>>(p);
  else if (channels == 1) k_remap<1><<<(int)blocks, 256, 0, stream>>>(p);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// EXTENSION: 16-bit Bayer -> BGR8 (frame_math.cuh demosaic_at16).  One thread per pixel, three byte stores; a pre-pass, not
// a tuned kernel -- the fast paths are for the 8-bit encodings the reference supports.
__global__ void __launch_bounds__(256) k_bayer16_to_bgr8(const uint8_t* __restrict__ in, long long in_frame_stride, int in_pitch, int rows,
                                                         int cols, int n_frames, int cfa, uint8_t* __restrict__ out) {
  const long long per_frame = (long long)rows * cols, total = per_frame * n_frames;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / per_frame);
    const long long rem = i - (long long)f * per_frame;
    const int y = (int)(rem / cols), x = (int)(rem - (long long)y * cols);
    int b, g, r;
    demosaic_at16(reinterpret_cast<const uint16_t*>(in + (long long)f * in_frame_stride), rows, cols, (size_t)in_pitch >> 1, y, x, cfa, b, g, r);
    uint8_t* o = out + 3 * i;
    o[0] = (uint8_t)b; o[1] = (uint8_t)g; o[2] = (uint8_t)r;
  }
}

cudaError_t launch_bayer16_to_bgr8(const uint8_t* in, long long in_frame_stride, int in_pitch, int rows, int cols, int n_frames, int cfa,
                                   uint8_t* out, cudaStream_t stream, int* launches) {
  const long long total = (long long)rows * cols * n_frames;
  if (total <= 0) return cudaSuccess;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  if (launches) ++*launches;
  k_bayer16_to_bgr8<<<(int)blocks, 256, 0, stream>>>(in, in_frame_stride, in_pitch, rows, cols, n_frames, cfa, out);
  return cudaGetLastError();
}

// Validity mask of the rectified image (the `rect_mask_` the reference declares but never fills, undistortion.hpp:136):
// 255 where all four bilinear taps of cv::remap lie inside the source image -- the pixel is an interpolation of real
// pixels only --, 0 where the constant border contributes.  Depends on the map and the source size only.
__global__ void __launch_bounds__(256) k_rect_mask(const float2* __restrict__ map, int orows, int ocols, int rows, int cols,
                                                   uint8_t* __restrict__ mask) {
  const long long total = (long long)orows * ocols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float2 m = __ldg(map + i);
    const int ix = remap_fix(m.x) >> 5, iy = remap_fix(m.y) >> 5;  // the integers cv::remap derives (frame_math.cuh remap_pixel)
    mask[i] = (ix >= 0 && ix + 1 < cols && iy >= 0 && iy + 1 < rows) ? 255 : 0;
  }
}

cudaError_t launch_rect_mask(const float2* map, int orows, int ocols, int rows, int cols, uint8_t* mask, cudaStream_t stream, int* launches) {
  const long long total = (long long)orows * ocols;
  if (total <= 0) return cudaSuccess;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  if (launches) ++*launches;
  k_rect_mask<<<(int)blocks, 256, 0, stream>>>(map, orows, ocols, rows, cols, mask);
  return cudaGetLastError();
}

cudaError_t launch_remap_bgrx(const RemapParams& p, int sm_count, cudaStream_t stream, int* launches) {
  (void)sm_count;
  if (p.ocols <= 0 || p.orows <= 0 || p.n_frames <= 0) return cudaSuccess;
  if (p.n_frames > 65535) return cudaErrorInvalidValue;
  const dim3 grid((unsigned)((p.ocols + 127) / 128), (unsigned)((p.orows + 63) / 64), (unsigned)p.n_frames);
  if (launches) ++*launches;
  if (p.pmap) k_remap_bgrx<true><<<grid, 256, 0, stream>>>(p);
  else k_remap_bgrx<false><<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace rip
