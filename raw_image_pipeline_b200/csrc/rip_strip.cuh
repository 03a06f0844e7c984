// Fused chain, strip kernel: the fast path's debayer -> flip -> WB -> colour calibration -> gamma -> vignetting -> enhancer
// (raw_image_pipeline.hpp:143-166) for camera-shaped Bayer frames, organised so that no warp ever waits for another and
// the per-row code is one straight line:
//
//   * a WARP owns a vertical strip of the frame, 128 pixels wide (32 lanes x 4 adjacent pixels) and up to `seg_h` rows
//     tall, and walks down its rows.  Every Bayer row is read from shared memory exactly once (three 32-bit words per
//     lane) into a sliding three-row window held in registers (bayer_window.cuh);
//   * each warp feeds itself: lane 0 has the TMA unit copy chunks of 4 rows x 160 bytes (strip + 16-byte halo columns,
//     zero fill outside the frame) into the warp's private ring of three chunks, two chunks ahead of the arithmetic,
//     completion signalled on the warp's own mbarriers.  A chunk is exactly four output rows: for the light stage sets
//     the loop body is the straight-line code of those 4 rows with the CFA-phase selectors of even and odd rows hoisted
//     out of it; for the heavy ones (Lab / HSV stages) it is a rolled row loop that flips the phase per row, to stay
//     inside the instruction cache.  There is no __syncthreads() in the steady state (only when the CTA moves on to
//     another frame and swaps the per-frame white-balance table);
//   * the warp index is taken through a shuffle, so the compiler knows that ring slot, mbarrier addresses, TMA
//     coordinates and staging bases are warp-uniform and keeps them in uniform registers (otherwise it re-broadcasts
//     them in front of every TMA / mbarrier instruction: -22 integer-pipe instructions per row);
//   * stage sets that are table lookups on bytes only (no colour calibration, Lab or HSV step) use ByteChain below:
//     per-lane extraction selectors carry the frame-edge column replication and the 180-degree pixel order;
//   * OpenCV's border rule (output row 0 = the interior formula at row 1, row H-1 = at row H-2) is served by two extra
//     one-row units per strip and frame instead of special cases in the row loop;
//   * the 4-byte intermediate (B,G,R,0 -- what the undistortion gather reads) leaves the registers directly: one
//     16-byte store per lane and row, 512 contiguous bytes per warp.  BGR8 output (12 bytes per lane) is assembled in a
//     per-warp staging buffer and written by the warp's own TMA stores, four rows at a time;
//   * a CTA is 8 warps = 8 adjacent strips; every CTA of the persistent grid takes one contiguous run of the CTA units
//     (frame, row segment, strip group), so it swaps the per-frame white-balance tables at most twice.
//
// Per-pixel arithmetic: chain_quad.cuh (bit-exact against pixel_math.cuh, which is bit-exact against the cv2 oracle).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "bayer_window.cuh"
#include "chain_quad.cuh"
#include "frame_math.cuh"
#include "kernels.hpp"
#include "tma.cuh"

namespace rip {

namespace {

constexpr int SW = 128;             // strip width in pixels
constexpr int NW = 8, NT = NW * 32;  // warps (= adjacent strips) per CTA
constexpr int CH = 4;               // Bayer rows per TMA chunk = output rows per pass of the loop body
constexpr int NS = 3;               // chunks in a warp's ring
constexpr int ROW_B = 160, ROW_W = ROW_B / 4;  // staged Bayer row: columns x0-16 .. x0+143 (the TMA needs 16-byte aligned x)
constexpr int X_WORD0 = 3;          // word holding columns x0-4 .. x0-1
constexpr int CHUNK_B = ROW_B * CH;
constexpr int OUT_ROW_B = SW * 3;   // BGR8 staging row
constexpr int GROUP_B = CH * OUT_ROW_B;

// Shared memory of a CTA (dynamic): [padding][tables][warp rings][BGR8 staging][mbarriers].  chain_quad.cuh merges table
// indices into table addresses with OR / byte permutes, which needs the table block on a 4096-byte boundary of the SHARED
// WINDOW ADDRESS; static __shared__ alignment cannot give that (the window starts with 1 KB of system-reserved memory,
// alignas() counts from the end of it), so the block is aligned at run time and the launch requests 4 KB of slack.
template <uint32_t STAGES, bool BGRX>
struct StripSmem {
  // tables: the strip blob's layout (chain_tables.hpp SOFF_*), truncated to the prefix the stage set reads
  static constexpr int TBL_RAW = ((STAGES & ST_CC) && (STAGES & ST_WB)) ? STRIP_TABLE_BYTES : (STAGES & ST_VIG) ? STRIP_VIG_END
                                 : (STAGES & ST_ENH) ? STRIP_ENH_END : (STAGES & ST_WB) ? SOFF_SF : (STAGES & ST_GAMMA) ? SOFF_WB : 128;
  static constexpr int TBL = (TBL_RAW + 127) / 128 * 128;
  static constexpr int OFF_IN = TBL;                                   // [warp][slot][CHUNK_B]
  static constexpr int OFF_OUT = OFF_IN + NW * NS * CHUNK_B;           // [warp][buffer][row][384]
  static constexpr int OFF_MBAR = OFF_OUT + (BGRX ? 0 : NW * 2 * GROUP_B);
  static constexpr int BYTES = OFF_MBAR + NW * NS * 8;
  static constexpr int REQUEST = BYTES + 4096 - 16;  // dynamic shared memory is 16-byte aligned: at most 4080 bytes of padding
};

extern __shared__ uint8_t strip_smem_raw[];

// the enhancer's row-tail pixels (cv2's scalar loop rounds where the vector loop truncates, pixel_math.cuh): rare
// (only frames whose width is not a multiple of 32), so out of line and compiled once; the pixel sits in byte 0.
// (the G table holds the identity under pca, so the variant with a G lookup serves both white-balance methods)
template <uint32_t STAGES>
__device__ __noinline__ uint32_t chain_px_tail(uint32_t Bw, uint32_t Gw, uint32_t Rw, float m, const ChainConsts& k, const StripTables t) {
  return chain_px<STAGES, 0, true, true>(Bw, Gw, Rw, m, k, t);
}

__device__ __forceinline__ void copy_table_range(uint8_t* sm, const uint8_t* blob, int lo, int hi, int tid) {
  const uint4* src = reinterpret_cast<const uint4*>(blob);
  uint4* dst = reinterpret_cast<uint4*>(sm);
  for (int i = lo / 16 + tid; i < hi / 16; i += NT) dst[i] = __ldg(src + i);
}

// three staging words of one lane and row (12 bytes, lane stride 3 words: conflict-free), by shared-window address
__device__ __forceinline__ void sts_row12(uint32_t a, uint32_t w0, uint32_t w1, uint32_t w2) {
  asm volatile("st.shared.u32 [%0], %1;\n\tst.shared.u32 [%0+4], %2;\n\tst.shared.u32 [%0+8], %3;" ::"r"(a), "r"(w0), "r"(w1), "r"(w2) : "memory");
}
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {  // IMAD: the multiply-add pipe, not the integer ALU
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// Stage sets whose chain is table lookups on bytes only (none / white balance / gamma / both; no colour calibration, Lab
// or HSV step).  The four pixels of a quad are taken out of the packed channel words with per-lane byte selectors that
// already contain the frame-edge column replication and, for a rotated frame, the reversed pixel order -- neither costs
// an instruction per row -- and the output words are assembled by multiply-adds.  out[j] = byte j of the selectors' order.
template <uint32_t STAGES, bool WBG>
struct ByteChain {
  uint32_t sel[4];  // 0x7650 + source byte of output pixel j
  __device__ __forceinline__ void init(uint32_t colfix, bool rev) {
#pragma unroll
    for (int j = 0; j < 4; ++j) sel[j] = 0x7650u + ((colfix >> (4 * (rev ? 3 - j : j))) & 15u);
  }
  __device__ __forceinline__ uint32_t channel(uint32_t w, int j, taddr wb, const StripTables& t, bool has_wb) const {
    uint32_t v;
    if ((STAGES & ST_WB) && has_wb) {
      v = lds_u8(taddr_byte_sel(wb, w, sel[j]));  // the kernel stores gamma(white balance(v)) in the per-frame table when gamma is on
    } else if (STAGES & ST_GAMMA) {
      v = lds_u8(taddr_byte_sel(t.gamma, w, sel[j]));
    } else {
      v = prmt(w, 0u, sel[j]);
    }
    return v;
  }
  __device__ __forceinline__ void lookup(uint32_t Bw, uint32_t Gw, uint32_t Rw, const StripTables& t, uint32_t b[4], uint32_t g[4], uint32_t r[4]) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      b[j] = channel(Bw, j, t.wb_b, t, true);
      g[j] = channel(Gw, j, t.wb_g, t, WBG);
      r[j] = channel(Rw, j, t.wb_r, t, true);
    }
  }
};

// KEY = stage bits | KEY_WBG (the G channel has a white-balance table: ccc).  Colour calibration with a non-zero bias is
// not handled here (launch_fused_strip's caller routes it to the tile kernel).
constexpr uint32_t KEY_WBG = 32u;
template <uint32_t KEY, bool BGRX, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_fused_strip(const __grid_constant__ FrameParams P, const __grid_constant__ StripGeom G,
                                                       const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                                                       const __grid_constant__ CUtensorMap out_map1) {
  constexpr uint32_t STAGES = KEY & ST_ALL;
  constexpr bool WBG = (KEY & KEY_WBG) != 0;
  // rows of straight-line code: the whole chunk for the light stage sets; one row (rolled loop, CFA phase flipped per row)
  // where the Lab / HSV chain of four pixels is already ~10 KB of code -- unrolled bodies thrash the instruction cache of
  // warps that run out of step (measured: 7.2 vs 5.5 ms per 64 x 12 MP)
  constexpr int UNR = (STAGES & (ST_VIG | ST_ENH)) ? 1 : 4;
  constexpr bool BYTE_ONLY = (STAGES & (ST_CC | ST_VIG | ST_ENH)) == 0;
  using L = StripSmem<STAGES, BGRX>;
  const int tid = threadIdx.x, lane = tid & 31;
  // the warp index through a shuffle: tells the compiler it is warp-uniform, so everything derived from it (ring, staging
  // and mbarrier addresses, TMA coordinates) lives in uniform registers instead of being re-broadcast before every TMA op
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  uint8_t* const sm = strip_smem_raw + ((0u - smem_u32(strip_smem_raw)) & 4095u);  // 4096-byte aligned shared-window address
  uint8_t* const sm_in = sm + L::OFF_IN + warp * (NS * CHUNK_B);   // this warp's ring
  uint8_t* const sm_out = sm + L::OFF_OUT + warp * (2 * GROUP_B);  // this warp's two staging groups (BGR8 only)
  const uint32_t sm_out_a = smem_u32(sm_out);
  unsigned long long* const mbar = reinterpret_cast<unsigned long long*>(sm + L::OFF_MBAR) + warp * NS;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(&mbar[s], 1);
    fence_mbar_init();
  }
  if ((STAGES & ST_GAMMA) && !(STAGES & ST_VIG)) copy_table_range(sm, P.strip_tables, SOFF_GAMMA, SOFF_WB, tid);
  if (STAGES & ST_ENH) copy_table_range(sm, P.strip_tables, SOFF_SF, STRIP_ENH_END, tid);
  if (STAGES & ST_VIG) copy_table_range(sm, P.strip_tables, SOFF_LABC, STRIP_VIG_END, tid);
  __syncthreads();
  const StripTables T = strip_tables_at(taddr_of_shared(sm));
  const bool rev = P.angle == 180;
  const int tail_start = P.ocols & ~31;
  int cur_frame = -1;
  uint32_t slot = 0, ph = 0;  // this warp's ring: slot of the next chunk to consume, per-slot mbarrier parities
  uint32_t obuf = 0;          // BGR8 staging: group being filled

  // Unit list: frame-major, then row segment, then strip group.  With per-frame white-balance tables every CTA takes one
  // contiguous run of it: it changes frame -- and reloads the tables behind two barriers -- at most twice (dealt
  // round-robin, every unit of a CTA lies in a different frame: the reloads cost 28 % of the kernel).  Without them the
  // units are dealt round-robin: the grid then works on one band of one frame at a time (measured 0.84 vs 0.88 ms).
  constexpr bool PER_FRAME_TABLES = (STAGES & ST_WB) != 0;
  const long long u_first = PER_FRAME_TABLES ? G.total_units * blockIdx.x / gridDim.x : (long long)blockIdx.x;
  const long long u_end = PER_FRAME_TABLES ? G.total_units * (blockIdx.x + 1) / gridDim.x : G.total_units;
  const long long u_step = PER_FRAME_TABLES ? 1 : (long long)gridDim.x;
  for (long long u = u_first; u < u_end; u += u_step) {
    const int frame = (int)(u / G.units_per_frame);
    const int rem = (int)(u - (long long)frame * G.units_per_frame);
    const int seg = rem / G.ngroups, grp = rem - seg * G.ngroups;
    // per-frame white-balance tables (uniform over the CTA): floats where the colour calibration consumes them, bytes otherwise
    if ((STAGES & ST_WB) && frame != cur_frame) {
      __syncthreads();  // nobody reads the previous frame's tables any more
      const float* src = P.wbf + (size_t)frame * 768;  // plain loads: written by a prior kernel
      for (int i = tid; i < 768; i += NT) {
        if (STAGES & ST_CC) reinterpret_cast<float*>(sm + SOFF_WBF)[i] = src[i];
        else if (BYTE_ONLY && (STAGES & ST_GAMMA)) sm[SOFF_WB + i] = sm[SOFF_GAMMA + __float2int_rz(src[i])];  // ByteChain: gamma folded into the per-frame table, one lookup per channel
        else sm[SOFF_WB + i] = (uint8_t)__float2int_rz(src[i]);
      }
      __syncthreads();
      cur_frame = frame;
    }
    const int strip = grp * NW + warp;
    if (strip >= G.nstrips) continue;

    // Rows of the unit: `n` centre rows ca .. ca + n - 1 (the rows the demosaic formula is evaluated at), emitted as the
    // input-frame rows ye0 .. ye0 + n - 1.  Segments cover the interior rows 1 .. H-2; the two border units re-evaluate
    // rows 1 and H-2 for the frame's first and last row (OpenCV's border rule, frame_math.cuh demosaic_at).
    int ca, n, ye0;
    if (seg < G.nseg) { ca = 1 + seg * G.seg_h; n = min(G.seg_h, P.rows - 1 - ca); ye0 = ca; }
    else if (seg == G.nseg) { ca = 1; n = 1; ye0 = 0; }
    else { ca = P.rows - 2; n = 1; ye0 = P.rows - 1; }

    // strips are anchored in the OUTPUT frame (TMA stores reject negative coordinates, loads zero-fill)
    const int ox0 = strip * SW;
    const int x0 = rev ? P.cols - SW - ox0 : ox0;  // may be negative for the last strip of a rotated frame
    const int x = x0 + 4 * lane;
    const bool active = x >= 0 && x < P.cols;
    const int xc = min(max(x, 0), P.cols - 4);     // inactive lanes compute on a valid address and store nothing
    const int oxb = rev ? P.cols - 4 - x : x;      // output column of the quad's lowest-address pixel
    const bool tail_quad = (STAGES & ST_ENH) && oxb >= tail_start;
    const uint32_t colfix = x == 0 ? 0x3211u : (x + 4 == P.cols ? 0x2210u : 0x3210u);  // column 0 <- 1, W-1 <- W-2
    ByteChain<STAGES, WBG> bytes;
    if (BYTE_ONLY) bytes.init(colfix, rev);
    const BayerPhase phase = bayer_phase(ca, P.cfa);  // "even" rows: ca, ca + 2, ...
    BayerPhaseRT phase_rt = bayer_phase_rt(ca, P.cfa);  // rolled row loop only: phase of the current centre row

    // chunk k of the unit holds Bayer rows ca - 3 + 4k .. ca + 4k: chunk 0 only supplies the two rows above the first
    // centre's bottom row, chunk k >= 1 the bottom rows of centres ca + 4(k-1) .. ca + 4(k-1) + 3
    const int nq = (n + CH - 1) / CH;
    int issued = nq + 1 < NS ? nq + 1 : NS;
    if (lane == 0) {
      uint32_t s = slot;
      for (int c = 0; c < issued; ++c) {
        mbar_expect_tx(&mbar[s], CHUNK_B);
        tma_load_3d(sm_in + s * CHUNK_B, &in_map, &mbar[s], x0 - 16, ca - 3 + c * CH, frame);
        s = s + 1 == NS ? 0 : s + 1;
      }
    }
    auto refill_and_advance = [&]() {  // the chunk in `slot` is consumed: load the unit's next unissued chunk into it
      __syncwarp();
      if (issued <= nq) {
        if (lane == 0) {
          mbar_expect_tx(&mbar[slot], CHUNK_B);
          tma_load_3d(sm_in + slot * CHUNK_B, &in_map, &mbar[slot], x0 - 16, ca - 3 + issued * CH, frame);
        }
        ++issued;
      }
      slot = slot + 1 == NS ? 0 : slot + 1;
    };
    mbar_wait(&mbar[slot], (ph >> slot) & 1u);
    ph ^= 1u << slot;
    const uint32_t* rowp = reinterpret_cast<const uint32_t*>(sm_in + slot * CHUNK_B) + X_WORD0 + lane;
    BayerRow rn = load_bayer_row<true>(rowp + 2 * ROW_W, phase);   // row ca - 1
    BayerRow rm = load_bayer_row<false>(rowp + 3 * ROW_W, phase);  // row ca
    refill_and_advance();

    const int oy0 = rev ? P.rows - 1 - ye0 : ye0;  // output row of the first emitted row; the following ones go down (up when rotated)
    // per-lane positions as 32-bit offsets from uniform bases (a frame of the 4-byte intermediate is < 2^31 bytes)
    uint8_t* const outf = P.out + (long long)frame * P.out_frame_stride;
    const int ostep = rev ? -P.out_pitch : P.out_pitch;
    int ooff = oy0 * P.out_pitch + oxb * 4;       // BGRX only
    int voff = ye0 * P.vig_pitch + xc;            // mask in input-frame coordinates, floats
    const int lane_pos = 12 * (rev ? 31 - lane : lane);
    const int sstep = rev ? -OUT_ROW_B : OUT_ROW_B;
    int done = 0;

    for (int k = 1; k <= nq; ++k) {
      mbar_wait(&mbar[slot], (ph >> slot) & 1u);
      ph ^= 1u << slot;
      rowp = reinterpret_cast<const uint32_t*>(sm_in + slot * CHUNK_B) + X_WORD0 + lane;
      const int nv = min(CH, n - done);  // rows of this chunk that exist (4 except at the end of a ragged unit)
      uint32_t sp = 0;                   // BGR8: this lane's place in the staging row being filled (shared-window address)
      if (!BGRX) {
        if (lane == 0) tma_wait_read<1>();  // the group about to be filled was handed to the TMA two groups ago
        __syncwarp();
        sp = sm_out_a + obuf * GROUP_B + (rev ? (CH - 1) * OUT_ROW_B : 0) + lane_pos;
      }
      // one output row: row `tt` of the chunk.  `tc`: the row's parity when it is a compile-time constant (unrolled
      // bodies), -1 when the phase is tracked at run time (rolled loop).
      auto row_step = [&](auto tc, int tt) {
        constexpr int t = decltype(tc)::value;
        uint32_t Bw, Gw, Rw;
        if constexpr (t < 0) {
          const BayerRow rs = load_bayer_row_below(rowp + tt * ROW_W, phase_rt);
          demosaic_window(rn, rm, rs, phase_rt, Bw, Gw, Rw);
          phase_rt.flip();
          rn = rm; rm = rs;
        } else {
          const BayerRow rs = load_bayer_row<(t & 1) == 0>(rowp + tt * ROW_W, phase);  // the bottom row has the other parity
          demosaic_window<(t & 1) != 0>(rn, rm, rs, phase, Bw, Gw, Rw);
          rn = rm; rm = rs;
        }
        if constexpr (BYTE_ONLY) {
          uint32_t b[4], g[4], r[4];
          bytes.lookup(Bw, Gw, Rw, T, b, g, r);
          if (BGRX) {
            if (active && tt < nv) {
              uint4 v;
              v.x = mad_u32(r[0], 65536u, mad_u32(g[0], 256u, b[0])); v.y = mad_u32(r[1], 65536u, mad_u32(g[1], 256u, b[1]));
              v.z = mad_u32(r[2], 65536u, mad_u32(g[2], 256u, b[2])); v.w = mad_u32(r[3], 65536u, mad_u32(g[3], 256u, b[3]));
              *reinterpret_cast<uint4*>(outf + ooff) = v;
            }
            ooff += ostep;
          } else {
            sts_row12(sp, mad_u32(b[1], 16777216u, mad_u32(r[0], 65536u, mad_u32(g[0], 256u, b[0]))),
                      mad_u32(g[2], 16777216u, mad_u32(b[2], 65536u, mad_u32(r[1], 256u, g[1]))),
                      mad_u32(r[3], 16777216u, mad_u32(g[3], 65536u, mad_u32(b[3], 256u, r[2]))));
            sp += sstep;
          }
        } else {
          if (colfix != 0x3210u) { Bw = prmt(Bw, 0u, colfix); Gw = prmt(Gw, 0u, colfix); Rw = prmt(Rw, 0u, colfix); }
          float m[4] = {1.0f, 1.0f, 1.0f, 1.0f};
          if (STAGES & ST_VIG) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(P.vig + voff));
            m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w;
            voff += P.vig_pitch;
          }
          uint32_t px[4];
          if (!tail_quad) {
            chain_quad<STAGES, WBG>(Bw, Gw, Rw, m, P.k, T, px);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) px[j] = chain_px_tail<STAGES>(Bw >> (8 * j), Gw >> (8 * j), Rw >> (8 * j), m[j], P.k, T);
          }
          if (BGRX) {
            if (active && tt < nv) {
              if (rev) *reinterpret_cast<uint4*>(outf + ooff) = make_uint4(px[3], px[2], px[1], px[0]);
              else *reinterpret_cast<uint4*>(outf + ooff) = make_uint4(px[0], px[1], px[2], px[3]);
            }
            ooff += ostep;
          } else {  // lanes beyond the frame edge write staging bytes the TMA store clips
            if (rev) sts_row12(sp, prmt(px[3], px[2], 0x4210), prmt(px[2], px[1], 0x5421), prmt(px[1], px[0], 0x6542));
            else sts_row12(sp, prmt(px[0], px[1], 0x4210), prmt(px[1], px[2], 0x5421), prmt(px[2], px[3], 0x6542));
            sp += sstep;
          }
        }
      };
      if constexpr (UNR == 1) {
#pragma unroll 1
        for (int tt = 0; tt < CH; ++tt) row_step(std::integral_constant<int, -1>{}, tt);
      } else {
        row_step(std::integral_constant<int, 0>{}, 0);
        row_step(std::integral_constant<int, 1>{}, 1);
        row_step(std::integral_constant<int, 2>{}, 2);
        row_step(std::integral_constant<int, 3>{}, 3);
      }
      if (!BGRX) {  // hand the group to the TMA: one 4-row store, or row by row at the end of a ragged unit
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          const uint8_t* gb = sm_out + obuf * GROUP_B;
          const int oyg = oy0 + (rev ? -done : done);  // output row of the group's first processed row
          if (nv == CH) {
            tma_store_3d(&out_map, gb, ox0 * 3 / 4, rev ? oyg - (CH - 1) : oyg, frame);  // 4-byte elements; clipped at the frame edge
          } else {
            for (int j = 0; j < nv; ++j)
              tma_store_3d(&out_map1, gb + (rev ? CH - 1 - j : j) * OUT_ROW_B, ox0 * 3 / 4, rev ? oyg - j : oyg + j, frame);
          }
          tma_commit();
        }
        obuf ^= 1u;
      }
      done += CH;
      refill_and_advance();
    }
  }
  if (!BGRX && lane == 0) tma_wait_read<0>();  // shared memory must stay valid until the last store has read it
}

template <uint32_t KEY, bool BGRX, int MINB>
cudaError_t launch_strip_instance(const FrameParams& p, const StripGeom& g, const CUtensorMap& im, const CUtensorMap& om, const CUtensorMap& om1,
                                  int sm_count, cudaStream_t stream) {
  constexpr int smem = StripSmem<(KEY & ST_ALL), BGRX>::REQUEST;
  static int occ_of_device[64] = {0};  // per instantiation and device (benign race: every writer stores the same value)
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  int occ = occ_of_device[dev & 63];
  if (occ == 0) {
    e = cudaFuncSetAttribute(k_fused_strip<KEY, BGRX, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fused_strip<KEY, BGRX, MINB>, NT, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    occ_of_device[dev & 63] = occ;
  }
  const long long cap = (long long)sm_count * occ;
  const int grid = (int)(g.total_units < cap ? g.total_units : cap);
  k_fused_strip<KEY, BGRX, MINB><<<grid, NT, smem, stream>>>(p, g, im, om, om1);
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
