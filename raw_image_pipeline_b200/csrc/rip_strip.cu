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
  int nseg = (p.orows + 40) / 80;
  if (nseg < 1) nseg = 1;
  int h = (p.orows + nseg - 1) / nseg;
  h = (h + OR_ROWS - 1) / OR_ROWS * OR_ROWS;
  g.seg_h = h;
  g.nseg = (p.orows + h - 1) / h;
  g.units_per_frame = g.nseg * g.ngroups;
  g.total_units = (long long)g.units_per_frame * p.n_frames;
  return g;
}

bool strip_kernel_ok(uint32_t stages, const FrameParams& p) {
  // a colour calibration with a bias stays with the tile kernel (no bias add is compiled into the strip instantiations)
  return !((stages & ST_CC) && p.k.has_bias) && p.strip_tables != nullptr;
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
                            (cuuint64_t)p.n_frames, (cuuint64_t)p.out_pitch, ofs, SW * 3 / 4, OR_ROWS) ||
        !make_tensor_map_3d(&om1, CU_TENSOR_MAP_DATA_TYPE_UINT32, p.out, (cuuint64_t)(p.ocols * 3 / 4), (cuuint64_t)p.orows,
                            (cuuint64_t)p.n_frames, (cuuint64_t)p.out_pitch, ofs, SW * 3 / 4, 1))
      return cudaErrorInvalidValue;
  }
  const StripGeom g = strip_geometry(p);
  const uint32_t key = (stages & ST_ALL) | (((stages & ST_WB) && wb_has_g_table) ? KEY_WBG : 0u);
  if (launches) ++*launches;
  if (!bgrx) return launch_fused_strip_bgr8(key, p, g, im, om, om1, sm_count, stream);
  // experiment switch ("debug/fused_kernel" = 2): the benchmarked instantiation at 3 CTAs per SM (85 registers)
  if (variant == 2 && key == ST_ALL) return launch_strip_instance<ST_ALL, true, 3>(p, g, im, om, om1, sm_count, stream);
  return dispatch_strip<0, true>(key, p, g, im, om, om1, sm_count, stream);
}

}  // namespace rip
