// BGR8-output instantiations of the strip kernel (rip_strip.cuh), in their own translation unit so that they compile in
// parallel with the 4-byte-intermediate ones (rip_strip.cu).
#include "rip_strip.cuh"

namespace rip {

cudaError_t launch_fused_strip_bgr8(uint32_t key, const FrameParams& p, const StripGeom& g, const CUtensorMap& im, const CUtensorMap& om,
                                    const CUtensorMap& om1, int sm_count, cudaStream_t stream) {
  return dispatch_strip<0, false>(key, p, g, im, om, om1, sm_count, stream);
}

}  // namespace rip
