// Launch interface between the host pipeline (rip_api.cu) and the sm_100a kernels
// (rip_kernels.cu).  Plain structs; everything device-side is a raw pointer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "chain_tables.hpp"

namespace rip {

enum : int { SRC_BAYER = 0, SRC_BGR = 1, SRC_RGB = 2, SRC_MONO = 3, SRC_BAYER16 = 4 /* host side only: converted to SRC_BGR by launch_bayer16_to_bgr8 */ };

struct FrameParams {
  const uint8_t* in;          // n_frames x rows x in_pitch
  uint8_t* out;               // n_frames x orows x out_pitch (BGR8)
  long long in_frame_stride;  // bytes
  long long out_frame_stride; // bytes
  int in_pitch, out_pitch;    // bytes per row
  int rows, cols;             // input frame
  int orows, ocols;           // after flip
  int n_frames;
  int cfa;                    // CFA_* (frame_math.cuh)
  int angle;                  // 0 / 90 / 180 / 270
  int src;                    // SRC_*
  const uint8_t* tables;      // static blob (TABLE_BYTES)
  const uint8_t* strip_tables;  // the same tables in the strip kernel's layout (STRIP_TABLE_BYTES, chain_tables.hpp)
  const float* wbf;           // n_frames x 3 x 256 per-frame white-balance LUTs (B,G,R) as floats, or null
  const float* vig;           // vignetting mask in INPUT-frame coordinates: rows x cols, entry (y, x) = mask at flip_dest(y, x); + 4 rows of padding
  int vig_pitch;              // floats per mask row (== cols)
  ChainConsts k;
  unsigned long long* stats;  // n_frames x 8 (stats kernel only)
};

struct RemapParams {
  const uint8_t* src;  // n_frames x rows x pitch (CH channels)
  uint8_t* dst;        // n_frames x orows x dpitch
  long long src_frame_stride, dst_frame_stride;
  int rows, cols, pitch;
  int orows, ocols, dpitch;
  int n_frames;
  const float2* map;   // orows x ocols (x, y)
  const uint32_t* pmap;  // optional packed fixed-point map (frame_math.cuh remap_pack_entry), used by launch_remap_bgrx when set
  const int4* tiles;     // optional tile table of the packed map (remap_tile_table), used by launch_remap_tile
  const uint32_t* tmap;  // ... and the packed map padded to whole tiles, `tmap_pitch` entries per row
  int tmap_pitch;
};

// Tile geometry of launch_remap_tile and its host-side table: per REMAP_TILE_W x REMAP_TILE_H output tile
// {bx0, by0, flags, 0} = origin of the source box (bx0 % 4 == 0) and REMAP_TILE_FAST when every tap of every pixel of
// the tile lies inside the REMAP_BOX_W x REMAP_BOX_H box at that origin and no entry is "far".
constexpr int REMAP_TILE_W = 128, REMAP_TILE_H = 24, REMAP_BOX_W = 176, REMAP_BOX_H = 48;
constexpr int REMAP_TILE_FAST = 1;
// `padded`: the packed map padded to whole tiles (ceil(ocols / W) * W entries per row, ceil(orows / H) * H rows)
void remap_tile_table(const uint32_t* packed, int orows, int ocols, int* table /* 4 ints per tile, row-major tiles */, uint32_t* padded);

// All launchers enqueue on `stream`, return the CUDA error of the launch, and add the number of
// kernels launched to *launches.
cudaError_t launch_fused(uint32_t stages, const FrameParams& p, int sm_count, cudaStream_t stream, int* launches);
cudaError_t launch_pca_stats(const FrameParams& p, int sm_count, cudaStream_t stream, int* launches);
cudaError_t launch_pca_lut(const unsigned long long* stats, float* wbf, float* coeff_out, int n_frames,
                           cudaStream_t stream, int* launches);
cudaError_t launch_gain_lut(const float* gains_bgr, float* wbf, int n_frames, cudaStream_t stream, int* launches);
// Fast path (rip_fast.cu): TMA-staged tiles, packed-byte demosaic.  fast_path_ok() looks at the input side
// (Bayer, width % 16 == 0, no 90/270 rotation, 16-byte aligned), fast_out_ok() at the output buffer.
bool fast_path_ok(const FrameParams& p);
bool fast_out_ok(const FrameParams& p, bool bgrx);
// bgrx: write 4-byte B,G,R,0 pixels (out_pitch = ocols * 4) -- the intermediate format of launch_remap_bgrx
cudaError_t launch_fused_fast(uint32_t stages, const FrameParams& p, bool bgrx, int sm_count, cudaStream_t stream, int* launches);
cudaError_t launch_pca_stats_fast(const FrameParams& p, int sm_count, cudaStream_t stream, int* launches);
// Strip kernel (rip_strip.cu): the same fast path with warp-private TMA rings and a sliding-window demosaic.  Applies
// wherever fast_path_ok() && fast_out_ok() hold; needs p.strip_tables.
struct StripGeom {
  int nstrips, ngroups;   // 128-pixel strips per output row; groups of 8 adjacent strips (one CTA unit each)
  int nseg, seg_h;        // row segments of the interior rows 1 .. H-2 per frame and their height; + 2 one-row border units
  int units_per_frame;
  long long total_units;
};
StripGeom strip_geometry(const FrameParams& p);
// `wb_has_g_table`: the G channel's white-balance table is not the identity (ccc); `variant`: experiment switch
bool strip_kernel_ok(uint32_t stages, const FrameParams& p);
bool strip_kernel_preferred(uint32_t stages);  // the kernel family the default configuration picks for this stage set
cudaError_t launch_fused_strip(uint32_t stages, bool wb_has_g_table, const FrameParams& p, bool bgrx, int variant, int sm_count,
                               cudaStream_t stream, int* launches);
// 1-channel non-Bayer input (mono8 ...): only flip and the gamma LUT apply (the colour modules skip images that do not
// have 3 channels: white_balance.hpp:50-52, color_calibration.hpp:47-49, color_enhancer.hpp:38-40)
cudaError_t launch_mono(const FrameParams& p, bool gamma, cudaStream_t stream, int* launches);
cudaError_t launch_remap(int channels, const RemapParams& p, cudaStream_t stream, int* launches);
// EXTENSION (16-bit Bayer, frame_math.cuh demosaic_at16): n frames of rows x cols u16 (`in_pitch` / `in_frame_stride` in
// bytes) -> tightly packed BGR8 frames; the chain then runs on those like on a bgr8 input.
cudaError_t launch_bayer16_to_bgr8(const uint8_t* in, long long in_frame_stride, int in_pitch, int rows, int cols, int n_frames, int cfa,
                                   uint8_t* out, cudaStream_t stream, int* launches);
// u8 validity mask of the rectified image: 255 where all four taps of the remap lie inside the rows x cols source
cudaError_t launch_rect_mask(const float2* map, int orows, int ocols, int rows, int cols, uint8_t* mask, cudaStream_t stream, int* launches);
// same remap from a 4-byte-per-pixel B,G,R,0 source (p.pitch = cols * 4) to BGR8
cudaError_t launch_remap_bgrx(const RemapParams& p, int sm_count, cudaStream_t stream, int* launches);
// tile path of the same (rip_fast.cu): TMA-staged source boxes, needs the packed map and ocols % 4 == 0
bool remap_tile_ok(const RemapParams& p);
cudaError_t launch_remap_tile(const RemapParams& p, int sm_count, cudaStream_t stream, int* launches);

}  // namespace rip
