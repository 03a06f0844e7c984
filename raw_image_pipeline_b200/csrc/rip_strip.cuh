// Fused chain, strip kernel: the fast path's debayer -> flip -> WB -> colour calibration -> gamma -> vignetting -> enhancer
// (raw_image_pipeline.hpp:143-166) for camera-shaped Bayer frames, organised so that no warp ever waits for another:
//
//   * a WARP owns a vertical strip of the frame, 128 pixels wide (32 lanes x 4 adjacent pixels) and `seg_h` rows tall,
//     and walks down its rows.  Every Bayer row is read from shared memory exactly once (three 32-bit words per lane)
//     into a sliding three-row window held in registers (bayer_window.cuh), so the demosaic costs ~9 instructions per
//     pixel instead of the ~25 of a per-row 3x3 fetch;
//   * each warp feeds itself: lane 0 has the TMA unit copy chunks of 4 rows x 160 bytes (strip + 16-byte halo columns,
//     zero fill outside the frame) into the warp's private ring of three chunks, two chunks ahead of the arithmetic,
//     completion signalled on the warp's own mbarriers.  There is no __syncthreads() in the steady state (only when the
//     CTA moves on to another frame and swaps the per-frame white-balance table);
//   * the 4-byte intermediate (B,G,R,0 -- what the undistortion gather reads) leaves the registers directly: one
//     16-byte store per lane and row, 512 contiguous bytes per warp.  BGR8 output (12 bytes per lane) is assembled in a
//     per-warp staging buffer and written by the warp's own TMA stores, two rows at a time;
//   * a CTA is 8 warps = 8 adjacent strips; CTA units (frame, row segment, strip group) are dealt round-robin over a
//     persistent grid, so the grid works on a narrow band of one or two frames at a time (L2 locality of the vignetting
//     mask and of the per-frame tables).
//
// Per-pixel arithmetic: pixel_math.cuh (bit-exact against the cv2 oracle, tests/test_pixel_math_host.py).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "bayer_window.cuh"
#include "chain_quad.cuh"
#include "frame_math.cuh"
#include "kernels.hpp"
#include "tma.cuh"

namespace rip {

namespace {

constexpr int SW = 128;             // strip width in pixels
constexpr int NW = 8, NT = NW * 32;  // warps (= adjacent strips) per CTA
constexpr int CH = 4;               // Bayer rows per TMA chunk
constexpr int NS = 3;               // chunks in a warp's ring
constexpr int ROW_B = 160, ROW_W = ROW_B / 4;  // staged Bayer row: columns x0-16 .. x0+143 (the TMA needs 16-byte aligned x)
constexpr int X_WORD0 = 3;          // word holding columns x0-4 .. x0-1
constexpr int CHUNK_B = ROW_B * CH;
constexpr int OR_ROWS = 2;          // rows per output TMA store (BGR8 only)
constexpr int OUT_ROW_B = SW * 3;

template <uint32_t STAGES, bool BGRX>
struct StripSmem {
  // tables: a verbatim copy of the strip blob (chain_tables.hpp SOFF_*), truncated to what the stage set reads
  static constexpr int TBL = (STAGES & ST_ENH) ? STRIP_TABLE_BYTES : (STAGES & ST_VIG) ? SOFF_SV : (STAGES & ST_GAMMA) ? SOFF_G2 : (STAGES & ST_WB) ? SOFF_GAMMA : 16;
  static constexpr int COPY_LO = (STAGES & ST_VIG) ? 0 : SOFF_GAMMA;  // first blob byte a stage set without vignetting needs
  static constexpr int OUTB = BGRX ? 128 : NW * 2 * OR_ROWS * OUT_ROW_B;
  alignas(4096) uint8_t tables[(TBL + 127) / 128 * 128];
  alignas(128) uint8_t in[NW][NS][CHUNK_B];
  alignas(128) uint8_t out[OUTB];  // [warp][buffer][row][384]
  alignas(8) unsigned long long mbar[NW][NS];
};

// the enhancer's row-tail pixels (cv2's scalar loop rounds where the vector loop truncates, pixel_math.cuh): rare
// (only frames whose width is not a multiple of 32), so out of line and compiled once; the pixel sits in byte 0.
// (the G table holds the identity under pca, so the variant with a G lookup serves both white-balance methods)
template <uint32_t STAGES>
__device__ __noinline__ uint32_t chain_px_tail(uint32_t Bw, uint32_t Gw, uint32_t Rw, float m, const ChainConsts& k, const StripTables t) {
  return chain_px<STAGES, 0, true, true, false>(Bw, Gw, Rw, m, k, t);
}

// KEY = stage bits | KEY_WBG (the G channel has a white-balance table: ccc).  Colour calibration with a non-zero bias is
// not handled here (launch_fused_strip's caller routes it to the tile kernel).
constexpr uint32_t KEY_WBG = 32u;
template <uint32_t KEY, bool BGRX, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_fused_strip(const __grid_constant__ FrameParams P, const __grid_constant__ StripGeom G,
                                                    const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                                                    const __grid_constant__ CUtensorMap out_map1) {
  constexpr uint32_t STAGES = KEY & ST_ALL;
  constexpr bool WBG = (KEY & KEY_WBG) != 0;
  __shared__ StripSmem<STAGES, BGRX> sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long* mbar = sm.mbar[warp];
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(&mbar[s], 1);
    fence_mbar_init();
  }
  if (STAGES & (ST_GAMMA | ST_VIG | ST_ENH)) {
    constexpr int LO = StripSmem<STAGES, BGRX>::COPY_LO / 16, HI = StripSmem<STAGES, BGRX>::TBL / 16;
    const uint4* src = reinterpret_cast<const uint4*>(P.strip_tables);
    uint4* dst = reinterpret_cast<uint4*>(sm.tables);
    for (int i = LO + tid; i < HI; i += NT) dst[i] = __ldg(src + i);
  }
  // chain_quad.cuh merges indices into table addresses with OR / byte permutes: needs this alignment of the shared address
  if ((smem_u32(sm.tables) & 4095u) != 0u) __trap();
  __syncthreads();
  const StripTables T = strip_tables_at(smem_u32(sm.tables));
  const bool rev = P.angle == 180;
  const int tail_start = P.ocols & ~31;
  int cur_frame = -1;
  uint32_t slot = 0, ph = 0;  // this warp's ring: next slot to consume, per-slot mbarrier parities
  uint32_t obuf = 0;          // BGR8 staging: buffer being filled

  for (long long u = blockIdx.x; u < G.total_units; u += gridDim.x) {
    const int frame = (int)(u / G.units_per_frame);
    const int rem = (int)(u - (long long)frame * G.units_per_frame);
    const int seg = rem / G.ngroups, grp = rem - seg * G.ngroups;
    if ((STAGES & ST_WB) && frame != cur_frame) {  // uniform over the CTA
      __syncthreads();                             // nobody reads the previous frame's table any more
      const float* src = P.wbf + (size_t)frame * 768;
      for (int i = tid; i < 768; i += NT) sm.tables[SOFF_WB + i] = (uint8_t)__float2int_rz(src[i]);  // plain load: written by a prior kernel
      __syncthreads();
      cur_frame = frame;
    }
    const int strip = grp * NW + warp;
    if (strip >= G.nstrips) continue;

    // strips and segments are anchored in the OUTPUT frame (TMA stores reject negative coordinates, loads zero-fill)
    const int oya = seg * G.seg_h, oyb = min(oya + G.seg_h, P.orows);
    const int ox0 = strip * SW;
    const int x0 = rev ? P.cols - SW - ox0 : ox0;  // may be negative for the last strip of a rotated frame
    const int iya = rev ? P.rows - oyb : oya, iyb = rev ? P.rows - oya : oyb;  // input rows [iya, iyb)
    const int x = x0 + 4 * lane;
    const bool active = x >= 0 && x < P.cols;
    // OpenCV's border rule: output row y is the interior formula at row clamp(y, 1, H-2) (frame_math.cuh demosaic_at)
    const int c_first = min(max(iya, 1), P.rows - 2), c_last = min(max(iyb - 1, 1), P.rows - 2);
    const int r0 = c_first - 1;               // first Bayer row this unit reads
    const int nrows = c_last + 2 - r0;        // rows r0 .. c_last + 1
    const int nchunks = (nrows + CH - 1) / CH;
    int issued = nchunks < NS ? nchunks : NS;
    if (lane == 0) {
      uint32_t s = slot;
      for (int c = 0; c < issued; ++c) {
        mbar_expect_tx(&mbar[s], CHUNK_B);
        tma_load_3d(sm.in[warp][s], &in_map, &mbar[s], x0 - 16, r0 + c * CH, frame);
        s = s + 1 == NS ? 0 : s + 1;
      }
    }
    // The first chunk (rows r0 .. r0 + 3) holds the first two rows of the window.
    mbar_wait(&mbar[slot], (ph >> slot) & 1u);
    ph ^= 1u << slot;
    const uint32_t* rowp = reinterpret_cast<const uint32_t*>(sm.in[warp][slot]) + X_WORD0 + lane;
    BayerRow rn = load_bayer_row(rowp, r0, P.cfa);
    BayerRow rm = load_bayer_row(rowp + ROW_W, r0 + 1, P.cfa);
    const uint32_t colfix = x == 0 ? 0x3211u : (x + 4 == P.cols ? 0x2210u : 0x3210u);  // column 0 <- 1, W-1 <- W-2
    const int oxb = rev ? P.cols - 4 - x : x;  // output column of the quad's lowest-address pixel
    const bool tail_quad = (STAGES & ST_ENH) && oxb >= tail_start;
    uint8_t* const outf = P.out + (long long)frame * P.out_frame_stride;
    const float* vig = (STAGES & ST_VIG) ? P.vig + (size_t)iya * P.vig_pitch + x : nullptr;

    for (int j = 2; j < nrows; ++j) {  // j: index, counted from r0, of the window's bottom row; centre row c = r0 + j - 1
      if ((j & (CH - 1)) == 0) {       // the bottom row enters the next chunk: the previous one is consumed, refill its slot
        __syncwarp();
        if (issued < nchunks) {
          if (lane == 0) {
            mbar_expect_tx(&mbar[slot], CHUNK_B);
            tma_load_3d(sm.in[warp][slot], &in_map, &mbar[slot], x0 - 16, r0 + issued * CH, frame);
          }
          ++issued;
        }
        slot = slot + 1 == NS ? 0 : slot + 1;
        mbar_wait(&mbar[slot], (ph >> slot) & 1u);
        ph ^= 1u << slot;
        rowp = reinterpret_cast<const uint32_t*>(sm.in[warp][slot]) + X_WORD0 + lane;
      }
      const int c = r0 + j - 1;
      const BayerRow rs = load_bayer_row(rowp + (j & (CH - 1)) * ROW_W, c + 1, P.cfa);
      uint32_t Bw, Gw, Rw;
      demosaic_window(rn, rm, rs, c, P.cfa, Bw, Gw, Rw);
      if (colfix != 0x3210u) { Bw = prmt(Bw, 0u, colfix); Gw = prmt(Gw, 0u, colfix); Rw = prmt(Rw, 0u, colfix); }
      rn = rm; rm = rs;
      // OpenCV's border rule: centre row 1 also serves output row 0, centre row H-2 also row H-1
      const int ylo = c == 1 ? iya : c, yhi = c == P.rows - 2 ? iyb - 1 : c;
      for (int y = ylo; y <= yhi; ++y) {
        const int oy = rev ? P.rows - 1 - y : y;
        uint32_t px[4] = {0u, 0u, 0u, 0u};
        if (active) {
          float m[4] = {1.0f, 1.0f, 1.0f, 1.0f};
          if (STAGES & ST_VIG) {  // mask stored in input-frame coordinates
            const float4 v = __ldg(reinterpret_cast<const float4*>(vig));
            m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w;
          }
          if (!tail_quad) {
            chain_quad<STAGES, WBG, false>(Bw, Gw, Rw, m, P.k, T, px);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) px[k] = chain_px_tail<STAGES>(Bw >> (8 * k), Gw >> (8 * k), Rw >> (8 * k), m[k], P.k, T);
          }
        }
        if (STAGES & ST_VIG) vig += P.vig_pitch;
        if (BGRX) {
          if (active) {
            const uint4 v = rev ? make_uint4(px[3], px[2], px[1], px[0]) : make_uint4(px[0], px[1], px[2], px[3]);
            *reinterpret_cast<uint4*>(outf + (size_t)oy * P.out_pitch + (size_t)oxb * 4) = v;
          }
        } else {
          const int jo = (y - iya) & (OR_ROWS - 1);
          if (jo == 0) {  // the buffer about to be filled was handed to the TMA two groups ago
            if (lane == 0) tma_wait_read<1>();
            __syncwarp();
          }
          uint8_t* grp_base = sm.out + (size_t)((warp * 2 + obuf) * OR_ROWS) * OUT_ROW_B;
          if (active) {
            uint32_t* o = reinterpret_cast<uint32_t*>(grp_base + (rev ? OR_ROWS - 1 - jo : jo) * OUT_ROW_B + 12 * (rev ? 31 - lane : lane));
            if (!rev) { o[0] = prmt(px[0], px[1], 0x4210); o[1] = prmt(px[1], px[2], 0x5421); o[2] = prmt(px[2], px[3], 0x6542); }
            else { o[0] = prmt(px[3], px[2], 0x4210); o[1] = prmt(px[2], px[1], 0x5421); o[2] = prmt(px[1], px[0], 0x6542); }
          }
          if (jo == OR_ROWS - 1) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {  // 4-byte elements; the TMA unit clips columns beyond the frame
              tma_store_3d(&out_map, grp_base, ox0 * 3 / 4, rev ? oy : oy - (OR_ROWS - 1), frame);
              tma_commit();
            }
            obuf ^= 1u;
          }
        }
      }
    }
    if (!BGRX) {  // rows left over when the unit's height is not a multiple of OR_ROWS: one-row stores
      const int left = (iyb - iya) & (OR_ROWS - 1);
      if (left) {
        uint8_t* grp_base = sm.out + (size_t)((warp * 2 + obuf) * OR_ROWS) * OUT_ROW_B;
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          for (int jo = 0; jo < left; ++jo) {
            const int y = iyb - left + jo, oy = rev ? P.rows - 1 - y : y;
            tma_store_3d(&out_map1, grp_base + (rev ? OR_ROWS - 1 - jo : jo) * OUT_ROW_B, ox0 * 3 / 4, oy, frame);
          }
          tma_commit();
        }
        obuf ^= 1u;
      }
    }
    slot = slot + 1 == NS ? 0 : slot + 1;  // the unit's last chunk is consumed
  }
  if (!BGRX && lane == 0) tma_wait_read<0>();  // shared memory must stay valid until the last store has read it
}

template <uint32_t KEY, bool BGRX, int MINB>
cudaError_t launch_strip_instance(const FrameParams& p, const StripGeom& g, const CUtensorMap& im, const CUtensorMap& om, const CUtensorMap& om1,
                                  int sm_count, cudaStream_t stream) {
  static_assert(sizeof(StripSmem<(KEY & ST_ALL), BGRX>) <= 48 * 1024, "k_fused_strip keeps its shared memory static");
  static int occ_of_device[64] = {0};  // per instantiation and device (benign race: every writer stores the same value)
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  int occ = occ_of_device[dev & 63];
  if (occ == 0) {
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fused_strip<KEY, BGRX, MINB>, NT, 0);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    occ_of_device[dev & 63] = occ;
  }
  const long long cap = (long long)sm_count * occ;
  const int grid = (int)(g.total_units < cap ? g.total_units : cap);
  k_fused_strip<KEY, BGRX, MINB><<<grid, NT, 0, stream>>>(p, g, im, om, om1);
  return cudaGetLastError();
}

template <uint32_t K, bool BGRX>
cudaError_t dispatch_strip(uint32_t key, const FrameParams& p, const StripGeom& g, const CUtensorMap& im, const CUtensorMap& om,
                           const CUtensorMap& om1, int sm_count, cudaStream_t stream) {
  if constexpr ((K & KEY_WBG) == 0 || (K & ST_WB) != 0) {  // a G table only exists with white balance
    if (key == K) return launch_strip_instance<K, BGRX, 4>(p, g, im, om, om1, sm_count, stream);
  }
  if constexpr (K < 63) return dispatch_strip<K + 1, BGRX>(key, p, g, im, om, om1, sm_count, stream);
  return cudaErrorInvalidValue;
}

}  // namespace

}  // namespace rip
