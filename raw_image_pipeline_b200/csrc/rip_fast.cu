// Fast path of the fused chain and of the PCA statistics for the shapes cameras actually deliver
// (8-bit Bayer, width a multiple of 16, no 90/270 rotation, 16-byte aligned buffers):
//
//   * Bayer tiles (128 x 32 px + halo, 160 B x 34 rows) are brought into shared memory by the TMA
//     unit (cp.async.bulk.tensor, zero fill outside the frame), double buffered behind mbarriers so the
//     load of tile i+1 runs under the arithmetic of tile i;
//   * the demosaic stencil runs on packed bytes (4 px per 32-bit word, frame_math.cuh demosaic_quad_swar);
//   * the BGR8 tile is assembled in shared memory and leaves through one TMA store per tile
//     (cp.async.bulk.tensor ... bulk_group), clipped to the frame by the hardware;
//   * persistent grid: resident CTAs walk the tile list of the whole batch frame-major;
//   * undistortion (k_remap_tile): a producer warp TMA-copies each output tile's source box of the 4-byte intermediate
//     into shared memory, eight consumer warps gather and blend from there (see the kernel's header comment).
//
// Everything else (ragged widths, unaligned buffers, 90/270 rotations, 3-channel inputs) takes the
// generic kernels in rip_kernels.cu; both produce identical bytes.
#include <cuda.h>
#include <type_traits>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "bayer_window.cuh"
#include "frame_math.cuh"
#include "kernels.hpp"
#include "tma.cuh"

namespace rip {

namespace {

constexpr int TW = 128, NT = 256;
constexpr int TH = 32;        // tile height of the fused kernel (occupancy: 3-5 CTAs of ~45-60 KB per SM)
constexpr int TH_STATS = 128;  // the statistics kernel has no output tile: taller tiles amortise the per-tile work
// The TMA unit needs the innermost coordinate 16-byte aligned (probed: tools/tma_probe, unaligned x traps with
// "illegal instruction"), so the 1-pixel halo is fetched as a 16-byte column on each side.
constexpr int IN_PITCH = 160;           // staged Bayer row: columns x0-16 .. x0+143
constexpr int IN_WORDS = IN_PITCH / 4;  // 40
constexpr int IN_X_WORD0 = 3;           // word holding columns x0-4 .. x0-1
constexpr int IN_BYTES = IN_PITCH * (TH + 2);                  // rows y0-1 .. y0+TH: what one TMA load delivers
constexpr int IN_BUF = (IN_BYTES + 127) / 128 * 128;
constexpr int IN_BYTES_STATS = IN_PITCH * (TH_STATS + 2), IN_BUF_STATS = (IN_BYTES_STATS + 127) / 128 * 128;
// output tile: BGR8 (3 B/px, what the caller receives) or BGRX (4 B/px, the intermediate the undistortion
// gather reads with one 32-bit load per tap)
template <bool BGRX> struct OutFmt { static constexpr int PITCH = TW * (BGRX ? 4 : 3), BUF = PITCH * TH; };

template <bool BGRX>
struct FastSmem {
  alignas(128) uint8_t in[2][IN_BUF];
  alignas(128) uint8_t out[OutFmt<BGRX>::BUF];  // single buffer: the TMA store of tile i has drained long before tile i+1 is assembled
  alignas(16) uint8_t tables[TABLE_BYTES];
  alignas(16) float wbf[768];
  alignas(8) unsigned long long mbar[2];
};

// Tiles are enumerated on a grid anchored at the origin of the OUTPUT frame (TMA stores do not take negative
// coordinates; loads do, with zero fill).  (x0, y0) is the tile's origin in the INPUT frame: equal to the output
// origin without rotation, mirrored -- and possibly negative for the partial edge tiles -- for 180 degrees.
// Every CTA walks one contiguous run of the batch's tile list; the position is advanced incrementally (the index
// divisions would otherwise cost more instructions per tile than the demosaic of a row).
struct TileCoord {
  int frame, x0, y0, ox0, oy0;
};
struct TileIter {
  int frame, ty, tx;
  __device__ __forceinline__ void init(long long t, int tiles_x, long long tiles_per_frame) {
    frame = (int)(t / tiles_per_frame);
    const int rem = (int)(t - (long long)frame * tiles_per_frame);
    ty = rem / tiles_x;
    tx = rem - ty * tiles_x;
  }
  __device__ __forceinline__ void advance(int tiles_x, int tiles_y) {
    if (++tx == tiles_x) {
      tx = 0;
      if (++ty == tiles_y) { ty = 0; ++frame; }
    }
  }
  __device__ __forceinline__ TileCoord coord(bool rev, int rows, int cols, int th = TH) const {
    TileCoord c;
    c.frame = frame;
    c.ox0 = tx * TW;
    c.oy0 = ty * th;
    c.x0 = rev ? cols - c.ox0 - TW : c.ox0;
    c.y0 = rev ? rows - c.oy0 - th : c.oy0;
    return c;
  }
};

// advance a TileIter by a fixed number of tiles (the grid size) without divisions
struct TileStride {
  int gx, gy, gf;
  __device__ __forceinline__ void init(int g, int tiles_x, int tiles_y) {
    gx = g % tiles_x;
    const int r = g / tiles_x;
    gy = r % tiles_y;
    gf = r / tiles_y;
  }
  __device__ __forceinline__ void advance(TileIter& q, int tiles_x, int tiles_y) const {
    q.tx += gx;
    if (q.tx >= tiles_x) { q.tx -= tiles_x; ++q.ty; }
    q.ty += gy;
    if (q.ty >= tiles_y) { q.ty -= tiles_y; ++q.frame; }
    q.frame += gf;
  }
};

// packed B / G / R words of the four pixels (y, x .. x+3); `s_in` is the staged tile whose row 0 is y0 - 1
__device__ __forceinline__ void quad_bgr_words(const uint32_t* s_in, int rows, int cols, int cfa, int y0, int y, int x, int lane,
                                               uint32_t& Bw, uint32_t& Gw, uint32_t& Rw) {
  const int yc = y < 1 ? 1 : (y > rows - 2 ? rows - 2 : y);  // OpenCV's border rule (frame_math.cuh demosaic_at)
  const uint32_t* p = s_in + (yc - y0) * IN_WORDS + IN_X_WORD0 + lane;  // staged row of yc - 1, word of columns x-4..x-1
  uint32_t w[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < 3; ++j) w[a][j] = p[a * IN_WORDS + j];
  const bool row_has_r = (((yc ^ (cfa >> 1)) & 1) == 0);
  const int cpar = (cfa ^ (cfa >> 1) ^ yc) & 1;  // == row_has_r ? (cfa & 1) : (cfa & 1) ^ 1
  demosaic_quad_swar(w, row_has_r, cpar, Bw, Gw, Rw);
  if (x == 0 || x + 4 == cols) {  // frame border columns (two quads per row): column 0 <- column 1, column W-1 <- column W-2
    const uint32_t fix = x == 0 ? 0x3211u : 0x2210u;
    Bw = prmt(Bw, 0u, fix); Gw = prmt(Gw, 0u, fix); Rw = prmt(Rw, 0u, fix);
  }
}

// the same quad straight from global memory (edge tiles only, see k_fused_fast)
__device__ __noinline__ uint3 quad_bgr_words_global(const FrameParams& P, int frame, int y, int x) {
  const uint8_t* fin = P.in + (long long)frame * P.in_frame_stride;
  uint32_t Bw = 0, Gw = 0, Rw = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int b, g, r;
    demosaic_at(fin, P.rows, P.cols, (size_t)P.in_pitch, y, x + k, P.cfa, b, g, r);
    Bw |= (uint32_t)b << (8 * k); Gw |= (uint32_t)g << (8 * k); Rw |= (uint32_t)r << (8 * k);
  }
  return make_uint3(Bw, Gw, Rw);
}

// =============================================================================================
// fused kernel, fast path
// =============================================================================================
template <uint32_t STAGES, bool BGRX>
__global__ void __launch_bounds__(NT, 4) k_fused_fast(const __grid_constant__ FrameParams P, const __grid_constant__ CUtensorMap in_map,
                                                   const __grid_constant__ CUtensorMap out_map) {
  // static shared memory (< 48 KB): table addresses are link-time constants, so lookups are `LDS [index + constant]`
  __shared__ FastSmem<BGRX> sm;
  constexpr int OUT_PITCH = OutFmt<BGRX>::PITCH;
  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps what derives from it (row number,
  // CFA phase selectors, row predicates) in uniform registers
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int tiles_x = (P.cols + TW - 1) / TW, tiles_y = (P.rows + TH - 1) / TH;
  const long long tiles_per_frame = (long long)tiles_x * tiles_y;
  const long long total = tiles_per_frame * P.n_frames;
  const bool rev = P.angle == 180;

  if (tid == 0) {
    mbar_init(&sm.mbar[0], 1);
    mbar_init(&sm.mbar[1], 1);
    fence_mbar_init();
  }
  if (STAGES & (ST_GAMMA | ST_VIG | ST_ENH)) {
    const uint4* src = reinterpret_cast<const uint4*>(P.tables);
    uint4* dst = reinterpret_cast<uint4*>(sm.tables);
    for (int i = tid; i < TABLE_BYTES / 16; i += NT) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const ChainTables T = chain_tables_from_blob(sm.tables, sm.wbf);

  long long t = total * blockIdx.x / gridDim.x;
  const long long t_end = total * (blockIdx.x + 1) / gridDim.x;
  TileIter ti;
  ti.init(t, tiles_x, tiles_per_frame);
  if (t < t_end && tid == 0) {
    const TileCoord c = ti.coord(rev, P.rows, P.cols);
    mbar_expect_tx(&sm.mbar[0], IN_BYTES);
    tma_load_3d(sm.in[0], &in_map, &sm.mbar[0], c.x0 - 16, c.y0 - 1, c.frame);
  }
  int cur_frame = -1;
  const int tail_start = P.ocols & ~31;  // cv2's scalar row tail in HSV2BGR (pixel_math.cuh)

  for (int it = 0; t < t_end; ++t, ++it) {
    const int buf = it & 1;
    const TileCoord c = ti.coord(rev, P.rows, P.cols);
    ti.advance(tiles_x, tiles_y);
    if (tid == 0) {
      if (t + 1 < t_end) {  // in[buf ^ 1] was last read before the barrier that ended iteration it - 1
        const TileCoord cn = ti.coord(rev, P.rows, P.cols);
        mbar_expect_tx(&sm.mbar[buf ^ 1], IN_BYTES);
        tma_load_3d(sm.in[buf ^ 1], &in_map, &sm.mbar[buf ^ 1], cn.x0 - 16, cn.y0 - 1, cn.frame);
      }
      tma_wait_read<0>();  // the store of the previous tile has finished reading `out`
    }
    if ((STAGES & ST_WB) && c.frame != cur_frame) {
      const float* src = P.wbf + (size_t)c.frame * 768;
      for (int i = tid; i < 768; i += NT) sm.wbf[i] = src[i];  // plain load: written by a prior kernel
      cur_frame = c.frame;
    }
    __syncthreads();  // `out` reusable, wbf visible
    mbar_wait(&sm.mbar[buf], (uint32_t)(it >> 1) & 1u);

    const uint32_t* s_in = reinterpret_cast<const uint32_t*>(sm.in[buf]);
    uint8_t* s_out = sm.out;
    const int x = c.x0 + 4 * lane;
    const bool x_in = x >= 0 && x < P.cols;
    // The frame's last row opens a tile (rows % TH == 1), or its first row closes one (same, rotated by 180): the border
    // rule needs a Bayer row two above / below the tile, outside the staged halo.  One single-row tile row per frame.
    const bool edge_tile = c.y0 == P.rows - 1 || c.y0 == 1 - TH;
#pragma unroll 1
    for (int rr = 0; rr < TH / 8 && x_in; ++rr) {
      const int r_in_tile = warp + 8 * rr;
      const int y = c.y0 + r_in_tile;
      if ((unsigned)y >= (unsigned)P.rows) continue;
      uint32_t Bw, Gw, Rw;
      if (!edge_tile) {
        quad_bgr_words(s_in, P.rows, P.cols, P.cfa, c.y0, y, x, lane, Bw, Gw, Rw);
      } else {
        const uint3 w = quad_bgr_words_global(P, c.frame, y, x);
        Bw = w.x; Gw = w.y; Rw = w.z;
      }
      const int oxb = rev ? P.cols - 4 - x : x;  // output column of the quad's lowest-address pixel
      float m[4] = {1.0f, 1.0f, 1.0f, 1.0f};
      if (STAGES & ST_VIG) {  // mask table is stored in input-frame coordinates
        const float4 v = __ldg(reinterpret_cast<const float4*>(P.vig + (size_t)y * P.vig_pitch + x));
        m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w;
      }
      uint32_t px[4];
      if ((STAGES & ST_ENH) && oxb >= tail_start) {  // whole quad lies in cv2's scalar row tail (rare: width % 32 != 0)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          px[k] = chain_pixel<STAGES>((int)prmt(Bw, 0u, 0x4440u + k), (int)prmt(Gw, 0u, 0x4440u + k), (int)prmt(Rw, 0u, 0x4440u + k), m[k], true, P.k, T);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          px[k] = chain_pixel<STAGES>((int)prmt(Bw, 0u, 0x4440u + k), (int)prmt(Gw, 0u, 0x4440u + k), (int)prmt(Rw, 0u, 0x4440u + k), m[k], false, P.k, T);
      }
      if (BGRX) {
        if (!rev) *reinterpret_cast<uint4*>(s_out + r_in_tile * OUT_PITCH + 16 * lane) = make_uint4(px[0], px[1], px[2], px[3]);
        else *reinterpret_cast<uint4*>(s_out + (TH - 1 - r_in_tile) * OUT_PITCH + 16 * (31 - lane)) = make_uint4(px[3], px[2], px[1], px[0]);
      } else if (!rev) {
        uint32_t* o = reinterpret_cast<uint32_t*>(s_out + r_in_tile * OUT_PITCH + 12 * lane);
        o[0] = prmt(px[0], px[1], 0x4210); o[1] = prmt(px[1], px[2], 0x5421); o[2] = prmt(px[2], px[3], 0x6542);
      } else {  // 180: mirrored inside the tile, pixel order reversed
        uint32_t* o = reinterpret_cast<uint32_t*>(s_out + (TH - 1 - r_in_tile) * OUT_PITCH + 12 * (31 - lane));
        o[0] = prmt(px[3], px[2], 0x4210); o[1] = prmt(px[2], px[1], 0x5421); o[2] = prmt(px[1], px[0], 0x6542);
      }
    }
    fence_async_smem();  // make this thread's shared-memory writes visible to the TMA unit
    __syncthreads();
    if (tid == 0) {
      // the TMA unit clips rows/columns beyond the tensor; 4-byte elements
      tma_store_3d(&out_map, s_out, BGRX ? c.ox0 : (c.ox0 * 3) / 4, c.oy0, c.frame);
      tma_commit();
    }
  }
  if (tid == 0) tma_wait_read<0>();  // shared memory must stay valid until the last store has read it
}

// =============================================================================================
// PCA white-balance statistics, fast path (white_balance.cpp:89-102)
// =============================================================================================
__device__ __forceinline__ unsigned dp4a_u(uint32_t a, uint32_t b, unsigned c) { return __dp4a(a, b, c); }

__global__ void __launch_bounds__(NT) k_pca_stats_fast(const __grid_constant__ FrameParams P, const __grid_constant__ CUtensorMap in_map) {
  __shared__ alignas(128) uint8_t s_inb[2][IN_BUF_STATS];
  __shared__ alignas(8) unsigned long long mbar[2];
  __shared__ unsigned long long s_acc[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = (P.cols + TW - 1) / TW, tiles_y = (P.rows + TH_STATS - 1) / TH_STATS;
  const long long tiles_per_frame = (long long)tiles_x * tiles_y;
  const long long total = tiles_per_frame * P.n_frames;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
  }
  if (tid < 8) s_acc[tid] = 0;
  __syncthreads();
  const bool rev = false;  // sums and maxima do not depend on the rotation
  // each CTA takes one contiguous run of tiles (at most two frame changes -> at most three flushes)
  long long t = total * blockIdx.x / gridDim.x;
  const long long t_end = total * (blockIdx.x + 1) / gridDim.x;
  TileIter ti;
  ti.init(t, tiles_x, tiles_per_frame);
  if (t < t_end && tid == 0) {
    const TileCoord c = ti.coord(rev, P.rows, P.cols, TH_STATS);
    mbar_expect_tx(&mbar[0], IN_BYTES_STATS);
    tma_load_3d(s_inb[0], &in_map, &mbar[0], c.x0 - 16, c.y0 - 1, c.frame);
  }
  int cur_frame = -1;
  // per-thread partial results of the current frame
  unsigned sb = 0, sr = 0, sg = 0;             // sums of <= 2^? values: flushed per tile (see below)
  unsigned long long sb2 = 0, sr2 = 0;
  uint32_t mx_b = 0, mx_g = 0, mx_r = 0;       // running maxima in two 16-bit lanes (VIMNMX3.U16x2; byte-wise max is emulated)

  auto flush = [&](int frame) {
    // fold the packed maxima to scalars, reduce over the warp, one shared atomic per warp, then one global per CTA
    unsigned mb = max(mx_b & 0xffffu, mx_b >> 16), mg = max(mx_g & 0xffffu, mx_g >> 16), mr = max(mx_r & 0xffffu, mx_r >> 16);
    unsigned long long vb = sb, vr = sr, vg = sg, vb2 = sb2, vr2 = sr2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      vb += __shfl_xor_sync(0xffffffffu, vb, o); vr += __shfl_xor_sync(0xffffffffu, vr, o); vg += __shfl_xor_sync(0xffffffffu, vg, o);
      vb2 += __shfl_xor_sync(0xffffffffu, vb2, o); vr2 += __shfl_xor_sync(0xffffffffu, vr2, o);
    }
    mb = __reduce_max_sync(0xffffffffu, mb); mg = __reduce_max_sync(0xffffffffu, mg); mr = __reduce_max_sync(0xffffffffu, mr);
    if (lane == 0) {
      atomicAdd(&s_acc[0], vb); atomicAdd(&s_acc[1], vb2); atomicAdd(&s_acc[2], vr); atomicAdd(&s_acc[3], vr2); atomicAdd(&s_acc[4], vg);
      atomicMax(&s_acc[5], (unsigned long long)mb); atomicMax(&s_acc[6], (unsigned long long)mg); atomicMax(&s_acc[7], (unsigned long long)mr);
    }
    __syncthreads();
    if (tid < 8) {
      unsigned long long* dst = P.stats + (size_t)frame * 8 + tid;
      if (tid < 5) atomicAdd(dst, s_acc[tid]); else atomicMax(dst, s_acc[tid]);
      s_acc[tid] = 0;
    }
    __syncthreads();
    sb = sr = sg = 0; sb2 = sr2 = 0; mx_b = mx_g = mx_r = 0;
  };

  for (int it = 0; t < t_end; ++t, ++it) {
    const int buf = it & 1;
    const TileCoord c = ti.coord(rev, P.rows, P.cols, TH_STATS);
    ti.advance(tiles_x, tiles_y);
    if (cur_frame >= 0 && c.frame != cur_frame) flush(cur_frame);  // uniform over the CTA
    cur_frame = c.frame;
    __syncthreads();  // everyone is done reading in[buf ^ 1]
    if (tid == 0) {
      if (t + 1 < t_end) {
        const TileCoord cn = ti.coord(rev, P.rows, P.cols, TH_STATS);
        mbar_expect_tx(&mbar[buf ^ 1], IN_BYTES_STATS);
        tma_load_3d(s_inb[buf ^ 1], &in_map, &mbar[buf ^ 1], cn.x0 - 16, cn.y0 - 1, cn.frame);
      }
    }
    mbar_wait(&mbar[buf], (uint32_t)(it >> 1) & 1u);
    const uint32_t* s_in = reinterpret_cast<const uint32_t*>(s_inb[buf]);
    const int x = c.x0 + 4 * lane;
    unsigned tb2 = 0, tr2 = 0;  // RPW rows x 4 px x 255^2 x 3 < 2^23
    // Warp w owns RPW consecutive tile rows.  Frame rows 0 and H-1 are copies of rows 1 and H-2 (OpenCV's border rule),
    // so only rows 1 .. H-2 are demosaiced and those two count twice (three times when H == 3).
    constexpr int RPW = TH_STATS / 8;
    const int ya = max(c.y0 + RPW * warp, 1), yb = min(c.y0 + RPW * warp + RPW - 1, P.rows - 2);
    if (x < P.cols && ya <= yb) {
      // CFA phase of the rows ya, ya + 2, ... computed once per tile; the rows in between use it mirrored (bayer_window.cuh)
      const BayerPhase ph = bayer_phase(ya, P.cfa);
      const uint32_t* const row0 = s_in + (1 - c.y0) * IN_WORDS + IN_X_WORD0 + lane;  // staged row of frame row 0
      BayerRow rn = load_bayer_row<true>(row0 + (ya - 1) * IN_WORDS, ph);
      BayerRow rm = load_bayer_row<false>(row0 + ya * IN_WORDS, ph);
      const bool edge_col = x == 0 || x + 4 == P.cols;  // frame border columns: column 0 <- column 1, column W-1 <- column W-2
      const uint32_t fix = x == 0 ? 0x3211u : 0x2210u;
      auto row_step = [&](auto odd_tag, int y) {
        constexpr bool ODD = decltype(odd_tag)::value;
        if (y > yb) return;
        const BayerRow rs = load_bayer_row<!ODD>(row0 + (y + 1) * IN_WORDS, ph);
        uint32_t Bw, Gw, Rw;
        demosaic_window<ODD>(rn, rm, rs, ph, Bw, Gw, Rw);
        if (edge_col) { Bw = prmt(Bw, 0u, fix); Gw = prmt(Gw, 0u, fix); Rw = prmt(Rw, 0u, fix); }
        sb = dp4a_u(Bw, 0x01010101u, sb); sr = dp4a_u(Rw, 0x01010101u, sr); sg = dp4a_u(Gw, 0x01010101u, sg);
        tb2 = dp4a_u(Bw, Bw, tb2); tr2 = dp4a_u(Rw, Rw, tr2);
        if (y == 1 || y == P.rows - 2) {  // these rows also stand for the frame's first / last row (both when H == 3)
          const uint32_t extra = (uint32_t)(y == 1) + (uint32_t)(y == P.rows - 2);
          const uint32_t ones = 0x01010101u * extra;
          sb = dp4a_u(Bw, ones, sb); sr = dp4a_u(Rw, ones, sr); sg = dp4a_u(Gw, ones, sg);
          tb2 += extra * dp4a_u(Bw, Bw, 0u); tr2 += extra * dp4a_u(Rw, Rw, 0u);
        }
        mx_b = __vimax3_u16x2(mx_b, lanes16(Bw, 0), lanes16(Bw, 1));
        mx_g = __vimax3_u16x2(mx_g, lanes16(Gw, 0), lanes16(Gw, 1));
        mx_r = __vimax3_u16x2(mx_r, lanes16(Rw, 0), lanes16(Rw, 1));
        rn = rm; rm = rs;
      };
#pragma unroll 2
      for (int k = 0; k < RPW; k += 2) {
        row_step(std::false_type{}, ya + k);
        row_step(std::true_type{}, ya + k + 1);
      }
    }
    sb2 += tb2; sr2 += tr2;
    // sb/sr/sg grow by <= 16 * 255 per tile: a CTA walks < 2^20 tiles of one frame, no overflow before the flush
  }
  if (cur_frame >= 0) flush(cur_frame);
}

// =============================================================================================
// undistortion, tile path (undistortion.cpp:214-245 -> cv::remap INTER_LINEAR, BORDER_CONSTANT 0)
// =============================================================================================
// An output tile of RT_W x RT_H pixels gathers from a compact region of the 4-byte intermediate.  The host knows that
// region for every tile (remap_tile_table, built once per map): a producer warp has the TMA unit copy a fixed
// BOX_W x BOX_H-pixel box at the footprint's origin into shared memory, two tiles ahead of the consumers; rows and
// columns of the box outside the image arrive as zeros, which is cv::remap's constant border.  Eight consumer warps
// gather the four taps of each pixel from shared memory (lane <-> adjacent pixels: conflict-free), blend, and exchange
// the packed results through a per-warp row so that every lane stores 12 contiguous bytes.  Tiles whose footprint fits
// the box and that hold no "far" entry (the common case) run without any per-pixel test; the others test every pixel
// and take the global-memory gather of the generic kernel where the box does not reach.
constexpr int RT_W = REMAP_TILE_W, RT_H = REMAP_TILE_H, BOX_W = REMAP_BOX_W, BOX_H = REMAP_BOX_H;
constexpr int RT_STAGES = 2;
constexpr int RT_CONSUMERS = 8, RT_THREADS = (RT_CONSUMERS + 1) * 32;
constexpr int BOX_BYTES = BOX_W * BOX_H * 4;
constexpr int RT_ROWS_WARP = RT_H / RT_CONSUMERS;  // consecutive tile rows owned by a warp, one per iteration
constexpr int RT_FRAME_GROUP = 8;                  // frames that share one load of a tile's map entries
static_assert(RT_W == 128 && RT_H % RT_CONSUMERS == 0, "a warp covers one 128-pixel tile row per iteration");

struct RemapTileSmem {
  alignas(128) uint32_t box[RT_STAGES][BOX_W * BOX_H];
  alignas(16) uint32_t px[RT_CONSUMERS][RT_W];
  int origin[RT_STAGES][4];  // bx0, by0, flags
  alignas(8) unsigned long long full[RT_STAGES], empty[RT_STAGES];
};


// one pixel of a tile without the FAST flag: shared-memory taps when they are inside the box, else the global gather
__device__ __noinline__ uint32_t remap_tile_pixel_slow(const RemapParams& P, const uint32_t* box, int bx0, int by0, const uint32_t* src,
                                                       int sx, int sy) {
  const int rx = (sx >> 5) - bx0, ry = (sy >> 5) - by0;
  if ((unsigned)rx < (unsigned)(BOX_W - 1) && (unsigned)ry < (unsigned)(BOX_H - 1)) {
    const uint32_t* q = box + ry * BOX_W + rx;
    return remap_blend(q[0], q[1], q[BOX_W], q[BOX_W + 1], sx, sy);
  }
  return remap_pixel_bgrx_fix(src, P.rows, P.cols, P.pitch >> 2, sx, sy);
}

__global__ void __launch_bounds__(RT_THREADS, 3) k_remap_tile(const __grid_constant__ RemapParams P, const __grid_constant__ CUtensorMap src_map) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  RemapTileSmem& sm = *reinterpret_cast<RemapTileSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;  // (a shuffle-derived, provably uniform warp index makes this kernel slower: 1.63 vs 1.47 ms)
  const int tiles_x = (P.ocols + RT_W - 1) / RT_W, tiles_y = (P.orows + RT_H - 1) / RT_H;
  const long long tiles_per_frame = (long long)tiles_x * tiles_y;
  if (tid == 0) {
    for (int i = 0; i < RT_STAGES; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], RT_CONSUMERS); }
    fence_mbar_init();
  }
  __syncthreads();
  // A unit of work is one tile for one GROUP of up to RT_FRAME_GROUP consecutive frames: the map is the same for every
  // frame, so after the group's first frame a tile's 12 KB of map rows come from L2 / L1 instead of DRAM (the map was
  // 40 % of this kernel's DRAM reads when every frame streamed it in again).
  // Units are dealt round-robin, group-major: CTA c takes units c, c + G, c + 2G, ...  The CTAs then work on one band of
  // adjacent tile rows of the same frames at a time, so the rows that vertically adjacent boxes share are still in L2
  // when the next tile row asks for them.
  const int n_groups = (P.n_frames + RT_FRAME_GROUP - 1) / RT_FRAME_GROUP;
  const long long total = tiles_per_frame * n_groups;
  long long t = blockIdx.x;
  const long long t_end = total;
  TileStride ts;
  ts.init((int)gridDim.x, tiles_x, tiles_y);
  TileIter ti;  // ti.frame counts frame groups here
  ti.init(t, tiles_x, tiles_per_frame);

  if (warp == RT_CONSUMERS) {
    // ---- producer warp ----
    int buf = 0; uint32_t round = 0;
    for (; t < t_end; t += gridDim.x, ts.advance(ti, tiles_x, tiles_y)) {
      const int4 info = __ldg(P.tiles + (ti.ty * tiles_x + ti.tx));
      // L2 prefetch of the tile's map rows (the consumers read them when they reach this unit)
      const int x0 = ti.tx * RT_W, y0 = ti.ty * RT_H;
      if (lane < RT_H)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.tmap + (size_t)(y0 + lane) * P.tmap_pitch + x0), "r"(RT_W * 4) : "memory");
      const int f0 = ti.frame * RT_FRAME_GROUP, f1 = min(f0 + RT_FRAME_GROUP, P.n_frames);
      for (int f = f0; f < f1; ++f) {
        if (round > 0) mbar_wait(&sm.empty[buf], (round - 1) & 1u);
        if (lane == 0) {
          sm.origin[buf][0] = info.x; sm.origin[buf][1] = info.y; sm.origin[buf][2] = info.z;
          mbar_expect_tx(&sm.full[buf], BOX_BYTES);
          tma_load_3d(sm.box[buf], &src_map, &sm.full[buf], info.x, info.y, f);
        }
        __syncwarp();
        if (++buf == RT_STAGES) { buf = 0; ++round; }
      }
    }
    return;
  }

  // ---- consumer warps ----
  // lane l, slot k: column x0 + l + 32 k of the iteration's row.  The four map entries of an iteration are loaded one
  // iteration ahead (across frame and unit boundaries too) from the tile-padded copy of the map (remap_tile_table):
  // entries beyond the image edge point at the edge pixel's source position, so they need no special case (and are never
  // stored).  Within a unit the same 12 KB of map rows are read once per frame: after the first frame they are L2 / L1 hits.
  auto map_load = [&](const TileIter& q, int i, uint32_t mm[4]) {
    const uint32_t* row = P.tmap + (size_t)(q.ty * RT_H + warp * RT_ROWS_WARP + i) * P.tmap_pitch + (q.tx * RT_W + lane);
#pragma unroll
    for (int k = 0; k < 4; ++k) mm[k] = __ldg(row + 32 * k);
  };
  int buf = 0; uint32_t round = 0;
  uint32_t mc[4] = {0u, 0u, 0u, 0u};
  if (t < t_end) map_load(ti, 0, mc);
  TileIter tn = ti;
  for (; t < t_end; t += gridDim.x, ts.advance(ti, tiles_x, tiles_y)) {
    const int x0 = ti.tx * RT_W, yw = ti.ty * RT_H + warp * RT_ROWS_WARP;
    ts.advance(tn, tiles_x, tiles_y);
    const int f0 = ti.frame * RT_FRAME_GROUP, f1 = min(f0 + RT_FRAME_GROUP, P.n_frames);
#pragma unroll 1
    for (int f = f0; f < f1; ++f) {
      mbar_wait(&sm.full[buf], round & 1u);
      const int bx0 = sm.origin[buf][0], by0 = sm.origin[buf][1];
      const bool fast = (sm.origin[buf][2] & REMAP_TILE_FAST) != 0;
      const uint32_t* box = sm.box[buf];
      uint8_t* drow = P.dst + (long long)f * P.dst_frame_stride + (size_t)yw * P.dpitch + (size_t)(x0 + 4 * lane) * 3;
#pragma unroll
      for (int i = 0; i < RT_ROWS_WARP; ++i, drow += P.dpitch) {
        uint32_t mn[4] = {0u, 0u, 0u, 0u};
        if (i + 1 < RT_ROWS_WARP) map_load(ti, i + 1, mn);
        else if (f + 1 < f1) map_load(ti, 0, mn);                  // same tile, next frame of the group
        else if (t + gridDim.x < t_end) map_load(tn, 0, mn);       // next unit
        const int ya = yw + i;
        uint32_t p[4];
        if (fast) {
          // every tap of every pixel of this tile lies inside the box, and no entry is "far"
          const uint32_t* rowbase = box + ((ya - by0) * BOX_W + (x0 + lane - bx0));
          uint32_t tp[4][4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t* q = rowbase + 32 * k + (remap_packed_dyi(mc[k]) * BOX_W + remap_packed_dxi(mc[k]));
            tp[k][0] = q[0]; tp[k][1] = q[1]; tp[k][2] = q[BOX_W]; tp[k][3] = q[BOX_W + 1];
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) p[k] = remap_blend_w(tp[k][0], tp[k][1], tp[k][2], tp[k][3], mc[k] & 31u, (mc[k] >> 10) & 0x7c0u);
        } else {
          const uint32_t* src = reinterpret_cast<const uint32_t*>(P.src + (long long)f * P.src_frame_stride);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            int sx, sy;
            remap_unpack_entry(mc[k], x0 + lane + 32 * k, ya, sx, sy);
            p[k] = remap_tile_pixel_slow(P, box, bx0, by0, src, sx, sy);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) sm.px[warp][lane + 32 * k] = p[k];
        __syncwarp();
        const uint4 q = *reinterpret_cast<const uint4*>(&sm.px[warp][4 * lane]);  // 4 consecutive pixels
        __syncwarp();
        if (x0 + 4 * lane < P.ocols && ya < P.orows) {  // ocols % 4 == 0 (launcher): the quad is complete
          uint32_t* d = reinterpret_cast<uint32_t*>(drow);
          d[0] = prmt(q.x, q.y, 0x4210); d[1] = prmt(q.y, q.z, 0x5421); d[2] = prmt(q.z, q.w, 0x6542);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) mc[k] = mn[k];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[buf]);  // this warp is done reading box[buf]
      if (++buf == RT_STAGES) { buf = 0; ++round; }
    }
  }
}

// ---- tensor maps ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  // function-local static: initialised once, thread-safe (pipelines on different GPUs may run on different host threads)
  static const EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(p);
    return static_cast<EncodeTiledFn>(nullptr);
  }();
  return fn;
}

bool make_map(CUtensorMap* map, CUtensorMapDataType type, const void* base, cuuint64_t d0, cuuint64_t d1, cuuint64_t d2, cuuint64_t stride1,
              cuuint64_t stride2, cuuint32_t b0, cuuint32_t b1) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {stride1, stride2};
  const cuuint32_t box[3] = {b0, b1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, type, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool in_map_for(const FrameParams& p, CUtensorMap* map, int th) {
  // a single frame still needs a legal stride for the (unused) frame dimension
  const cuuint64_t fstride = p.n_frames > 1 ? (cuuint64_t)p.in_frame_stride : (cuuint64_t)p.in_pitch * p.rows;
  return make_map(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, p.in, (cuuint64_t)p.cols, (cuuint64_t)p.rows, (cuuint64_t)p.n_frames,
                  (cuuint64_t)p.in_pitch, fstride, IN_PITCH, th + 2);
}

long long total_tiles(const FrameParams& p, int th) {
  return (long long)((p.cols + TW - 1) / TW) * ((p.rows + th - 1) / th) * p.n_frames;
}

template <uint32_t S, bool BGRX>
cudaError_t dispatch_fast(uint32_t stages, const FrameParams& p, const CUtensorMap& im, const CUtensorMap& om, int sm_count,
                          cudaStream_t stream) {
  if (stages == S) {
    static int occ_of_device[64] = {0};  // per instantiation and device
    static_assert(sizeof(FastSmem<BGRX>) <= 48 * 1024, "k_fused_fast keeps its shared memory static");
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int& occ = occ_of_device[dev & 63];
    if (occ == 0) {
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fused_fast<S, BGRX>, NT, 0);
      if (e != cudaSuccess) return e;
      if (occ < 1) occ = 1;
    }
    const long long tiles = total_tiles(p, TH), cap = (long long)sm_count * occ;
    const int grid = (int)(tiles < cap ? tiles : cap);
    k_fused_fast<S, BGRX><<<grid, NT, 0, stream>>>(p, im, om);
    return cudaGetLastError();
  }
  if constexpr (S < ST_ALL) return dispatch_fast<S + 1, BGRX>(stages, p, im, om, sm_count, stream);
  return cudaErrorInvalidValue;
}

}  // namespace

bool make_tensor_map_3d(CUtensorMap* map, CUtensorMapDataType type, const void* base, cuuint64_t d0, cuuint64_t d1, cuuint64_t d2,
                        cuuint64_t stride1, cuuint64_t stride2, cuuint32_t b0, cuuint32_t b1) {
  return make_map(map, type, base, d0, d1, d2, stride1, stride2, b0, b1);
}
bool tensor_maps_available() { return encode_tiled_fn() != nullptr; }

bool fast_path_ok(const FrameParams& p) {
  if (p.src != SRC_BAYER || !(p.angle == 0 || p.angle == 180)) return false;
  if (p.cols % 16 != 0 || p.cols < 16 || p.rows < 3 || p.n_frames < 1) return false;
  if (p.in_pitch % 16 != 0 || (reinterpret_cast<uintptr_t>(p.in) & 15) != 0) return false;
  if (p.n_frames > 1 && (p.in_frame_stride % 16 != 0 || p.in_frame_stride <= 0)) return false;
  return encode_tiled_fn() != nullptr;
}

bool fast_out_ok(const FrameParams& p, bool bgrx) {
  if (p.out_pitch != p.ocols * (bgrx ? 4 : 3) || (reinterpret_cast<uintptr_t>(p.out) & 15) != 0) return false;
  if (p.n_frames > 1 && (p.out_frame_stride % 16 != 0 || p.out_frame_stride <= 0)) return false;
  return true;
}

cudaError_t launch_fused_fast(uint32_t stages, const FrameParams& p, bool bgrx, int sm_count, cudaStream_t stream, int* launches) {
  CUtensorMap im, om;
  if (!in_map_for(p, &im, TH)) return cudaErrorInvalidValue;
  const cuuint64_t ofs = p.n_frames > 1 ? (cuuint64_t)p.out_frame_stride : (cuuint64_t)p.out_pitch * p.orows;
  // the output is described in 4-byte elements: a BGR8 row of `ocols` pixels is ocols * 3 / 4 of them
  const int pitch_elems = bgrx ? TW : TW * 3 / 4;
  if (!make_map(&om, CU_TENSOR_MAP_DATA_TYPE_UINT32, p.out, (cuuint64_t)(bgrx ? p.ocols : p.ocols * 3 / 4), (cuuint64_t)p.orows,
                (cuuint64_t)p.n_frames, (cuuint64_t)p.out_pitch, ofs, pitch_elems, TH))
    return cudaErrorInvalidValue;
  if (launches) ++*launches;
  return bgrx ? dispatch_fast<0, true>(stages & ST_ALL, p, im, om, sm_count, stream)
              : dispatch_fast<0, false>(stages & ST_ALL, p, im, om, sm_count, stream);
}

cudaError_t launch_pca_stats_fast(const FrameParams& p, int sm_count, cudaStream_t stream, int* launches) {
  CUtensorMap im;
  if (!in_map_for(p, &im, TH_STATS)) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(p.stats, 0, sizeof(unsigned long long) * 8 * p.n_frames, stream);
  if (e != cudaSuccess) return e;
  static int occ_of_device[64] = {0};  // per device (benign race: every writer stores the same value)
  int dev = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  int occ = occ_of_device[dev & 63];
  if (occ == 0) {
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pca_stats_fast, NT, 0);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    occ_of_device[dev & 63] = occ;
  }
  const long long tiles = total_tiles(p, TH_STATS), cap = (long long)sm_count * occ;
  const int grid = (int)(tiles < cap ? tiles : cap);
  if (launches) ++*launches;
  k_pca_stats_fast<<<grid, NT, 0, stream>>>(p, im);
  return cudaGetLastError();
}

bool remap_tile_ok(const RemapParams& p) {
  if (!p.tmap || !p.tiles || (p.ocols & 3) != 0 || p.ocols < 4 || p.orows < 1 || p.n_frames < 1) return false;
  if ((reinterpret_cast<uintptr_t>(p.src) & 15) != 0 || (p.pitch & 15) != 0 || p.pitch != p.cols * 4) return false;
  if ((reinterpret_cast<uintptr_t>(p.dst) & 3) != 0 || (p.dpitch & 3) != 0 || (p.dst_frame_stride & 3) != 0) return false;
  if (p.n_frames > 1 && (p.src_frame_stride % 16 != 0 || p.src_frame_stride <= 0)) return false;
  return encode_tiled_fn() != nullptr;
}

void remap_tile_table(const uint32_t* packed, int orows, int ocols, int* table, uint32_t* padded) {
  const int tiles_x = (ocols + RT_W - 1) / RT_W, tiles_y = (orows + RT_H - 1) / RT_H;
  const int pw = tiles_x * RT_W, ph = tiles_y * RT_H;
  const uint32_t far_entry = 0x80008000u;  // REMAP_FAR in both halves
  // the map padded to whole tiles: a padding entry at (x, y) carries the edge entry's displacement re-based to (x, y)
  for (int y = 0; y < ph; ++y)
    for (int x = 0; x < pw; ++x) {
      const int xc = x < ocols ? x : ocols - 1, yc = y < orows ? y : orows - 1;
      uint32_t e = packed[(size_t)yc * ocols + xc];
      if ((x != xc || y != yc) && (e & 0xffffu) != 0x8000u) {
        const int dx = (int)(int16_t)(e & 0xffffu) - 32 * (x - xc), dy = (int)(int16_t)(e >> 16) - 32 * (y - yc);
        e = (dx > REMAP_FAR && dy > REMAP_FAR) ? ((uint32_t)(uint16_t)dx | ((uint32_t)(uint16_t)dy << 16)) : far_entry;
      }
      padded[(size_t)y * pw + x] = e;
    }
  for (int ty = 0; ty < tiles_y; ++ty)
    for (int tx = 0; tx < tiles_x; ++tx) {
      int ix_lo = INT32_MAX, ix_hi = INT32_MIN, iy_lo = INT32_MAX, iy_hi = INT32_MIN;
      bool far = false;
      for (int y = ty * RT_H; y < ty * RT_H + RT_H; ++y)
        for (int x = tx * RT_W; x < tx * RT_W + RT_W; ++x) {
          const uint32_t e = padded[(size_t)y * pw + x];
          if ((e & 0xffffu) == 0x8000u) { far = true; continue; }
          const int ix = x + remap_packed_dxi(e), iy = y + remap_packed_dyi(e);
          ix_lo = ix < ix_lo ? ix : ix_lo; ix_hi = ix > ix_hi ? ix : ix_hi;
          iy_lo = iy < iy_lo ? iy : iy_lo; iy_hi = iy > iy_hi ? iy : iy_hi;
        }
      int* t = table + 4 * ((size_t)ty * tiles_x + tx);
      if (ix_lo == INT32_MAX) { t[0] = 0; t[1] = 0; t[2] = 0; t[3] = 0; continue; }  // nothing maps into the source
      const int bx0 = ix_lo & ~3;  // floor to a multiple of 4 pixels: 16-byte aligned TMA coordinate
      const bool fits = ix_hi + 1 - bx0 <= BOX_W - 1 && iy_hi + 1 - iy_lo <= BOX_H - 1;
      t[0] = bx0; t[1] = iy_lo; t[2] = (fits && !far) ? REMAP_TILE_FAST : 0; t[3] = 0;
    }
}

cudaError_t launch_remap_tile(const RemapParams& p, int sm_count, cudaStream_t stream, int* launches) {
  CUtensorMap sm;
  const cuuint64_t fstride = p.n_frames > 1 ? (cuuint64_t)p.src_frame_stride : (cuuint64_t)p.pitch * p.rows;
  if (!make_map(&sm, CU_TENSOR_MAP_DATA_TYPE_UINT32, p.src, (cuuint64_t)p.cols, (cuuint64_t)p.rows, (cuuint64_t)p.n_frames,
                (cuuint64_t)p.pitch, fstride, BOX_W, BOX_H))
    return cudaErrorInvalidValue;
  static int occ_of_device[64] = {0};  // per device: the shared-memory opt-in is a per-device attribute
  constexpr size_t smem = sizeof(RemapTileSmem);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  int& occ = occ_of_device[dev & 63];
  if (occ == 0) {
    e = cudaFuncSetAttribute(k_remap_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_remap_tile, RT_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
  }
  const long long units = (long long)((p.ocols + RT_W - 1) / RT_W) * ((p.orows + RT_H - 1) / RT_H) *
                          ((p.n_frames + RT_FRAME_GROUP - 1) / RT_FRAME_GROUP);
  const long long cap = (long long)sm_count * occ;
  const int grid = (int)(units < cap ? units : cap);
  if (launches) ++*launches;
  k_remap_tile<<<grid, RT_THREADS, smem, stream>>>(p, sm);
  return cudaGetLastError();
}

}  // namespace rip
