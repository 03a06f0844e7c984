// Host side of the strip kernel (rip_strip.cuh): geometry, tensor maps, and the instantiations that write the 4-byte
// intermediate.  The BGR8 instantiations are compiled in rip_strip_bgr8.cu (parallel build).
#include "rip_strip.cuh"

namespace rip {

cudaError_t launch_fused_strip_bgr8(uint32_t key, const FrameParams& p, const StripGeom& g, const CUtensorMap& im, const CUtensorMap& om,
                                    const CUtensorMap& om1, int sm_count, cudaStream_t stream);

StripGeom strip_geometry(const FrameParams& p) {
  StripGeom g{};
  g.nstrips = (p.ocols + SW - 1) / SW;
  g.ngroups = (g.nstrips + NW - 1) / NW;
  // row segments cover the interior rows 1 .. H-2 (the demosaic formula's domain); rows 0 and H-1 are two extra one-row
  // units per strip.  Segment height: a multiple of the chunk height close to 80 rows.
  const int interior = p.rows - 2;
  int nseg = (interior + 40) / 80;
  if (nseg < 1) nseg = 1;
  int h = (interior + nseg - 1) / nseg;
  h = (h + CH - 1) / CH * CH;
  g.seg_h = h;
  g.nseg = (interior + h - 1) / h;
  g.units_per_frame = (g.nseg + 2) * g.ngroups;
  g.total_units = (long long)g.units_per_frame * p.n_frames;
  return g;
}

bool strip_kernel_ok(uint32_t stages, const FrameParams& p) {
  // a colour calibration with a bias stays with the tile kernel (no bias add is compiled into the strip instantiations)
  if ((stages & ST_CC) && p.k.has_bias) return false;
  // per-lane positions inside a frame are 32-bit byte offsets (output rows, vignetting-mask rows)
  if ((long long)p.orows * p.out_pitch >= (1ll << 31) || (long long)p.rows * p.vig_pitch >= (1ll << 31)) return false;
  return !(stages & (ST_GAMMA | ST_VIG | ST_ENH)) || p.strip_tables != nullptr;
}

bool strip_kernel_preferred(uint32_t stages) {
  // Measured (profiles/r2_strip_kernel.md): the strip kernel wins where the per-pixel arithmetic is light (debayer + gamma:
  // 0.67 vs 1.02 ms per 64 x 12 MP); with the Lab / HSV stages the chain dominates and the tile kernel's synchronised warps
  // use the instruction cache better.
  return (stages & (ST_VIG | ST_ENH)) == 0;
}

cudaError_t launch_fused_strip(uint32_t stages, bool wb_has_g_table, const FrameParams& p, bool bgrx, int variant, int sm_count,
                               cudaStream_t stream, int* launches) {
  CUtensorMap im, om, om1;
  const cuuint64_t ifs = p.n_frames > 1 ? (cuuint64_t)p.in_frame_stride : (cuuint64_t)p.in_pitch * p.rows;
  if (!make_tensor_map_3d(&im, CU_TENSOR_MAP_DATA_TYPE_UINT8, p.in, (cuuint64_t)p.cols, (cuuint64_t)p.rows, (cuuint64_t)p.n_frames,
                          (cuuint64_t)p.in_pitch, ifs, ROW_B, CH))
    return cudaErrorInvalidValue;
  om = im; om1 = im;  // the 4-byte intermediate is stored from registers
  if (!bgrx) {
    const cuuint64_t ofs = p.n_frames > 1 ? (cuuint64_t)p.out_frame_stride : (cuuint64_t)p.out_pitch * p.orows;
    // BGR8 rows described in 4-byte elements: ocols * 3 / 4 of them (ocols % 16 == 0 on this path)
    if (!make_tensor_map_3d(&om, CU_TENSOR_MAP_DATA_TYPE_UINT32, p.out, (cuuint64_t)(p.ocols * 3 / 4), (cuuint64_t)p.orows,
                            (cuuint64_t)p.n_frames, (cuuint64_t)p.out_pitch, ofs, SW * 3 / 4, CH) ||
        !make_tensor_map_3d(&om1, CU_TENSOR_MAP_DATA_TYPE_UINT32, p.out, (cuuint64_t)(p.ocols * 3 / 4), (cuuint64_t)p.orows,
                            (cuuint64_t)p.n_frames, (cuuint64_t)p.out_pitch, ofs, SW * 3 / 4, 1))
      return cudaErrorInvalidValue;
  }
  const StripGeom g = strip_geometry(p);
  const uint32_t key = (stages & ST_ALL) | (((stages & ST_WB) && wb_has_g_table) ? KEY_WBG : 0u);
  if (launches) ++*launches;
  if (!bgrx) return launch_fused_strip_bgr8(key, p, g, im, om, om1, sm_count, stream);
  // experiment switch ("debug/fused_kernel" = 2 / 3): the benchmarked instantiation at 3 / 2 CTAs per SM (80 / 128 registers)
  if (variant == 3 && key == ST_ALL) return launch_strip_instance<ST_ALL, true, 3>(p, g, im, om, om1, sm_count, stream);
  return dispatch_strip<0, true>(key, p, g, im, om, om1, sm_count, stream);
}

}  // namespace rip
