// Arithmetic of the convolutional-colour-constancy white balance
// (raw_image_pipeline_white_balance/src/.../convolutional_color_constancy.cpp:91-386), written
// once for the sm_100a kernels (ccc.cu) and the CPU-only test harness (tests/hostsim), see
// pixel_math.cuh for the host/device convention.
#pragma once
#include <math.h>

#include "frame_math.cuh"

namespace rip {

constexpr int CCC_SMALL_W = 360;  // small_size_ (ccc.cpp:22)
constexpr int CCC_SMALL_H = 270;
constexpr int CCC_BINS = 256;     // model width = height = 256 (model/default.bin)

// One destination index of cv::resize(INTER_LINEAR) on 8-bit data: source index and the two
// 11-bit fixed-point weights (imgproc/src/resize.cpp, fixpt branch of resizeGeneric_).
struct CccAxisCoef {
  int s;         // source index of the first tap
  short a0, a1;  // weights, INTER_RESIZE_COEF_SCALE = 2048
};

// Host only (double arithmetic like OpenCV's set-up loop).  `vertical`: OpenCV zeroes the
// fraction at the image border only for the horizontal axis; vertically the two row pointers
// are clamped instead and the weights stay.
inline CccAxisCoef ccc_axis_coef(int src, int dst, int d) {
  const double inv_scale = (double)dst / src;
  const double scale = 1. / inv_scale;
  float f = (float)((d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= s;
  CccAxisCoef c;
  c.s = s;
  auto sat_short = [](float v) { long r = lrintf(v); return (short)(r < -32768 ? -32768 : (r > 32767 ? 32767 : r)); };
  c.a0 = sat_short((1.f - f) * 2048);
  c.a1 = sat_short(f * 2048);
  return c;
}
inline CccAxisCoef ccc_axis_coef_horizontal(int src, int dst, int d) {
  CccAxisCoef c = ccc_axis_coef(src, dst, d);
  if (c.s < 0) { c.s = 0; c.a0 = 2048; c.a1 = 0; }
  if (c.s >= src - 1) { c.s = src - 1; c.a0 = 2048; c.a1 = 0; }
  return c;
}

// horizontal pass: int32 S = p[s]*a0 + p[s+1]*a1   (HResizeLinear, ONE = 2048)
RIP_HD int ccc_hresize(int p0, int p1, int a0, int a1) { return p0 * a0 + p1 * a1; }

// vertical pass (VResizeLinearVec_32s8u, the SIMD form every row takes):
//   ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2, saturated to u8
RIP_HD int ccc_vresize(int S0, int S1, int b0, int b1) {
  const int x0 = S0 >> 4, x1 = S1 >> 4;
  const int v = (((b0 * x0) >> 16) + ((b1 * x1) >> 16) + 2) >> 2;
  return clamp_u8(v);
}

// Histogram bin of one pixel of the small image (calculateHistogramFeature, ccc.cpp:210-271).
// Returns false when the pixel is skipped (masked out or log of zero).
//   gray: cv::cvtColor(COLOR_BGR2GRAY) on CV_32F as this OpenCV build computes it (IPP):
//         fma(R, 0.299f, fma(B, 0.114f, G * 0.587f))          [probed exhaustively, 2^24 triples]
//   mask: !(gray > 255*bright) && (gray > 255*dark)   (THRESH_BINARY_INV & THRESH_BINARY)
//   log : cv::log on CV_32F, tabulated for the 256 possible inputs (kCvLogTab, from cv2)
RIP_HD bool ccc_bin(int b, int g, int r, float thr_hi, float thr_lo, const float* log_tab, float uv0, float bin_size,
                    int& u, int& v) {
  if (b == 0 || g == 0 || r == 0) return false;  // log(0) = -inf -> !isfinite
  const float gray = RIP_FMA((float)r, 0.299f, RIP_FMA((float)b, 0.114f, RIP_FMUL((float)g, 0.587f)));
  if (gray > thr_hi) return false;
  if (!(gray > thr_lo)) return false;
  const float lb = log_tab[b], lg = log_tab[g], lr = log_tab[r];
  const float fu = RIP_FSUB(RIP_FSUB(lg, lr), uv0) / bin_size;  // IEEE division on host and device
  const float fv = RIP_FSUB(RIP_FSUB(lg, lb), uv0) / bin_size;
  int iu = (int)roundf(fu), iv = (int)roundf(fv);               // round(): half away from zero
  u = iu < 0 ? 0 : (iu > 255 ? 255 : iu);
  v = iv < 0 ? 0 : (iv > 255 ? 255 : iv);
  return true;
}

// computeGains (ccc.cpp:342-381) from the arg-max position; exp_tab[k] = 1.0f / expf(-(k*bin + uv0))
// evaluated on the host by the same libm the reference would call.  Output order B, G, R.
RIP_HD void ccc_gains(int uv_x, int uv_y, const float* exp_tab, float gains_bgr[3]) {
  const float gr = exp_tab[uv_x], gg = 1.0f, gb = exp_tab[uv_y];
  float fac = gr < gg ? gr : gg;
  fac = fac < gb ? fac : gb;
  gains_bgr[0] = gb / fac; gains_bgr[1] = gg / fac; gains_bgr[2] = gr / fac;
}

}  // namespace rip
