// Neighbourhood / per-frame arithmetic shared by the sm_100a kernels and the CPU-only test
// harness (tests/hostsim): bilinear demosaic, cv::remap fixed-point bilinear, and the PCA
// white-balance solve.  See pixel_math.cuh for the host/device convention.
#pragma once
#include "pixel_math.cuh"

namespace rip {

// CFA phase: bit0 = column parity of the R sites, bit1 = row parity of the R sites.
//   bayer_rggb8 -> 0 (R at (0,0));  bayer_grbg8 -> 1 (R at (0,1));
//   bayer_gbrg8 -> 2 (R at (1,0));  bayer_bggr8 -> 3 (R at (1,1)).      
enum : int { CFA_RGGB = 0, CFA_GRBG = 1, CFA_GBRG = 2, CFA_BGGR = 3 };

// ---- bilinear demosaic at one site: debayer.cpp:45-79 == cv::demosaicing + R/B swap ------
// out(y,x) = interior formula evaluated at (clamp(y,1,H-2), clamp(x,1,W-2)) -- reproduces
// OpenCV's "copy the neighbouring interior pixel into the border" rule without copy passes.
RIP_HD void demosaic_at(const uint8_t* raw, int rows, int cols, size_t pitch, int y, int x, int cfa,
                        int& b, int& g, int& r) {
  y = y < 1 ? 1 : (y > rows - 2 ? rows - 2 : y);
  x = x < 1 ? 1 : (x > cols - 2 ? cols - 2 : x);
  const uint8_t* p = raw + (size_t)y * pitch + x;
  const uint8_t* pn = p - pitch;
  const uint8_t* ps = p + pitch;
  const int c = p[0];
  const bool row_has_r = ((y & 1) == ((cfa >> 1) & 1));
  const bool col_is_r = ((x & 1) == (cfa & 1));
  // colour (non-green) site <=> (row_has_r && col_is_r) || (!row_has_r && !col_is_r)
  if (row_has_r == col_is_r) {
    const int cross = (pn[0] + ps[0] + p[-1] + p[1] + 2) >> 2;
    const int diag = (pn[-1] + pn[1] + ps[-1] + ps[1] + 2) >> 2;
    g = cross;
    if (row_has_r) { r = c; b = diag; } else { b = c; r = diag; }
  } else {
    const int horiz = (p[-1] + p[1] + 1) >> 1;
    const int vert = (pn[0] + ps[0] + 1) >> 1;
    g = c;
    if (row_has_r) { r = horiz; b = vert; } else { b = horiz; r = vert; }
  }
}

// ---- EXTENSION: 16-bit Bayer (the reference throws for bayer_*16, debayer.cpp:76-78) -------------------
// cv::demosaicing on a CV_16UC1 frame uses the same integer formulas and the same border rule one depth up; the result is
// reduced to 8 bits like cv::Mat::convertTo(CV_8U, 1 / 257.f) == saturate_cast<uchar>(cvRound(v / 257.f)) -- which equals
// (v + 128) / 257 for every 16-bit v (no ties: 257 is odd; checked over all 65536 values, tests/test_pixel_math_host.py).
RIP_HD int reduce16to8(int v) { return (v + 128) / 257; }
RIP_HD void demosaic_at16(const uint16_t* raw, int rows, int cols, size_t pitch_elems, int y, int x, int cfa, int& b, int& g, int& r) {
  y = y < 1 ? 1 : (y > rows - 2 ? rows - 2 : y);
  x = x < 1 ? 1 : (x > cols - 2 ? cols - 2 : x);
  const uint16_t* p = raw + (size_t)y * pitch_elems + x;
  const uint16_t* pn = p - pitch_elems;
  const uint16_t* ps = p + pitch_elems;
  const int c = p[0];
  const bool row_has_r = ((y & 1) == ((cfa >> 1) & 1));
  const bool col_is_r = ((x & 1) == (cfa & 1));
  if (row_has_r == col_is_r) {
    const int cross = (pn[0] + ps[0] + p[-1] + p[1] + 2) >> 2;
    const int diag = (pn[-1] + pn[1] + ps[-1] + ps[1] + 2) >> 2;
    g = cross;
    if (row_has_r) { r = c; b = diag; } else { b = c; r = diag; }
  } else {
    const int horiz = (p[-1] + p[1] + 1) >> 1;
    const int vert = (pn[0] + ps[0] + 1) >> 1;
    g = c;
    if (row_has_r) { r = horiz; b = vert; } else { b = horiz; r = vert; }
  }
  b = reduce16to8(b); g = reduce16to8(g); r = reduce16to8(r);
}

// ---- four horizontally adjacent sites x..x+3 (x % 4 == 0) from three rows of packed words --
// w[row][0] = bytes x-4..x-1, w[row][1] = bytes x..x+3, w[row][2] = bytes x+4..x+7; rows are
// y-1, y, y+1.  Valid only when all four sites are interior columns (1 <= x, x+3 <= W-2) and the
// row has already been clamped by the caller.  `cpar` = column parity of the colour sites in
// this row, `row_has_r` as above.  Results are bit-identical to demosaic_at().
RIP_HD void demosaic_quad(const uint32_t w[3][3], bool row_has_r, int cpar, int b[4], int g[4], int r[4]) {
  int n[6], c[6], s[6];  // columns x-1 .. x+4
  n[0] = w[0][0] >> 24; c[0] = w[1][0] >> 24; s[0] = w[2][0] >> 24;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    n[k + 1] = (w[0][1] >> (8 * k)) & 255;
    c[k + 1] = (w[1][1] >> (8 * k)) & 255;
    s[k + 1] = (w[2][1] >> (8 * k)) & 255;
  }
  n[5] = w[0][2] & 255; c[5] = w[1][2] & 255; s[5] = w[2][2] & 255;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = k + 1;
    if ((k & 1) == cpar) {
      const int cross = (n[i] + s[i] + c[i - 1] + c[i + 1] + 2) >> 2;
      const int diag = (n[i - 1] + n[i + 1] + s[i - 1] + s[i + 1] + 2) >> 2;
      g[k] = cross;
      if (row_has_r) { r[k] = c[i]; b[k] = diag; } else { b[k] = c[i]; r[k] = diag; }
    } else {
      const int horiz = (c[i - 1] + c[i + 1] + 1) >> 1;
      const int vert = (n[i] + s[i] + 1) >> 1;
      g[k] = c[i];
      if (row_has_r) { r[k] = horiz; b[k] = vert; } else { b[k] = horiz; r[k] = vert; }
    }
  }
}

// ---- the same four sites with packed-byte (SWAR) arithmetic: the kernels' fast path -------------
// Inputs as for demosaic_quad (three rows of three words).  Outputs are packed words whose byte k is the
// B / G / R value of column x + k.  ~45 integer operations per quad instead of ~100.
RIP_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int shift) {  // low 32 bits of ((hi:lo) >> shift), 0 < shift < 32
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, shift);
#else
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> shift);
#endif
}
// per-byte (a + b + 1) >> 1
RIP_HD uint32_t avg_round_u8x4(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) & 0xfefefefeu) >> 1); }
// bytes (p, p + 2) of w as two 16-bit lanes
RIP_HD uint32_t lanes16(uint32_t w, int p) { return prmt(w, 0u, 0x4240u + 0x0101u * (uint32_t)p); }

RIP_HD void demosaic_quad_swar(const uint32_t w[3][3], bool row_has_r, int cpar, uint32_t& Bw, uint32_t& Gw, uint32_t& Rw) {
  const uint32_t Nc = w[0][1], Mc = w[1][1], Sc = w[2][1];
  const uint32_t Nl = funnel_r(w[0][0], Nc, 24), Nr = funnel_r(Nc, w[0][2], 8);  // columns x-1..x+2 / x+1..x+4
  const uint32_t Ml = funnel_r(w[1][0], Mc, 24), Mr = funnel_r(Mc, w[1][2], 8);
  const uint32_t Sl = funnel_r(w[2][0], Sc, 24), Sr = funnel_r(Sc, w[2][2], 8);
  const uint32_t H = avg_round_u8x4(Ml, Mr);  // (W + E + 1) >> 1 at every column
  const uint32_t V = avg_round_u8x4(Nc, Sc);  // (N + S + 1) >> 1
  // colour sites (column parity cpar): 4-neighbour sums in 16-bit lanes
  const uint32_t lsel = 0x4240u + 0x0101u * (uint32_t)cpar;  // PRMT selector of lanes16(., cpar), computed once
  const uint32_t X = ((prmt(Nc, 0u, lsel) + prmt(Sc, 0u, lsel) + prmt(Ml, 0u, lsel) + prmt(Mr, 0u, lsel) + 0x00020002u) >> 2) & 0x00ff00ffu;
  const uint32_t D = ((prmt(Nl, 0u, lsel) + prmt(Nr, 0u, lsel) + prmt(Sl, 0u, lsel) + prmt(Sr, 0u, lsel) + 0x00020002u) >> 2) & 0x00ff00ffu;
  // merge: at colour sites G = cross, native = raw, opposite = diagonal; at green sites G = raw,
  // the row's colour = horizontal average, the other colour = vertical average
  const uint32_t sel = 0x7250u - 0x4c4cu * (uint32_t)cpar;   // 0x7250 / 0x2604: lane values at the colour sites, second operand elsewhere
  const uint32_t site = 0x00ff00ffu << (8 * cpar);           // byte mask of the colour sites
  Gw = prmt(X, Mc, sel);
  const uint32_t row_colour = (Mc & site) | (H & ~site);
  const uint32_t other_colour = prmt(D, V, sel);
  Rw = row_has_r ? row_colour : other_colour;
  Bw = row_has_r ? other_colour : row_colour;
}

// ---- flip.cpp:37-58 as an index map: source coordinate of output pixel (oy, ox) ----------
// angle 90: out(y,x)=in(H-1-x, y); 180: in(H-1-y, W-1-x); 270: in(x, W-1-y)   
RIP_HD void flip_source(int angle, int rows, int cols, int oy, int ox, int& iy, int& ix) {
  if (angle == 90) { iy = rows - 1 - ox; ix = oy; }
  else if (angle == 180) { iy = rows - 1 - oy; ix = cols - 1 - ox; }
  else if (angle == 270) { iy = ox; ix = cols - 1 - oy; }
  else { iy = oy; ix = ox; }
}

// ---- cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) on 8UC{1,3}: undistortion.cpp:240-245 ----
template <int CH>
RIP_HD void remap_pixel(const uint8_t* src, int rows, int cols, size_t pitch, float mx, float my, int out[CH]) {
  const int sx = remap_fix(mx), sy = remap_fix(my);
  const int ix = sx >> 5, iy = sy >> 5;
  const int ax = sx & 31, ay = sy & 31;
  const int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32;
  const int w10 = ay * (32 - ax) * 32, w11 = ay * ax * 32;
  const bool x0 = (unsigned)ix < (unsigned)cols, x1 = (unsigned)(ix + 1) < (unsigned)cols;
  const bool y0 = (unsigned)iy < (unsigned)rows, y1 = (unsigned)(iy + 1) < (unsigned)rows;
  const uint8_t* p0 = src + (long long)iy * (long long)pitch + (long long)ix * CH;
  const uint8_t* p1 = p0 + pitch;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    int acc = 16384;
    if (y0 && x0) acc += w00 * p0[c];
    if (y0 && x1) acc += w01 * p0[CH + c];
    if (y1 && x0) acc += w10 * p1[c];
    if (y1 && x1) acc += w11 * p1[CH + c];
    out[c] = clamp_u8(acc >> 15);
  }
}

// ---- the same remap from a 4-byte B,G,R,0 source (the fused kernel's intermediate format) ------------
// One 32-bit load per tap; the bilinear weights factor exactly in integers,
//   sum w*p = 32 * [ (32-ay) * ((32-ax) p00 + ax p01) + ay * ((32-ax) p10 + ax p11) ],
// so the horizontal step is a byte dot product (dp4a) and (acc + 2^14) >> 15 == (V + 512) >> 10.
RIP_HD uint32_t dot4_u8(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __dp4a(a, b, 0u);
#else
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r += ((a >> (8 * i)) & 255u) * ((b >> (8 * i)) & 255u);
  return r;
#endif
}
// returns b | g << 8 | r << 16; `pitch_px` = source row pitch in pixels; (sx, sy) = cvRound(map * 32)
RIP_HD uint32_t remap_pixel_bgrx_fix(const uint32_t* src, int rows, int cols, int pitch_px, int sx, int sy) {
  const int ix = sx >> 5, iy = sy >> 5;
  const uint32_t ax = (uint32_t)(sx & 31), ay = (uint32_t)(sy & 31);
  uint32_t t00 = 0, t01 = 0, t10 = 0, t11 = 0;
  if ((unsigned)ix < (unsigned)(cols - 1) && (unsigned)iy < (unsigned)(rows - 1)) {  // all four taps inside
    const uint32_t* p = src + (iy * pitch_px + ix);  // 32-bit pixel index: a frame holds < 2^31 pixels
    const uint32_t* q = p + pitch_px;
    t00 = p[0]; t01 = p[1]; t10 = q[0]; t11 = q[1];
  } else {  // BORDER_CONSTANT 0: taps outside the image contribute nothing
    const bool x0 = (unsigned)ix < (unsigned)cols, x1 = (unsigned)(ix + 1) < (unsigned)cols;
    const bool y0 = (unsigned)iy < (unsigned)rows, y1 = (unsigned)(iy + 1) < (unsigned)rows;
    const long long o = (long long)iy * (long long)pitch_px + ix;
    if (y0 && x0) t00 = src[o];
    if (y0 && x1) t01 = src[o + 1];
    if (y1 && x0) t10 = src[o + (long long)pitch_px];
    if (y1 && x1) t11 = src[o + (long long)pitch_px + 1];
  }
  const uint32_t wA = ax * 255u + 32u;  // bytes (32 - ax, ax, 0, 0)
  const uint32_t wB = wA << 16;         // bytes (0, 0, 32 - ax, ax)
  const uint32_t bg0 = prmt(t00, t01, 0x5140u), bg1 = prmt(t10, t11, 0x5140u);  // B00 B01 G00 G01
  const uint32_t r0 = prmt(t00, t01, 0x3362u), r1 = prmt(t10, t11, 0x3362u);    // R00 R01 0 0 (byte 3 of a pixel is 0)
  const uint32_t by = 32u - ay;
  const uint32_t vb = by * dot4_u8(bg0, wA) + ay * dot4_u8(bg1, wA) + 512u;
  const uint32_t vg = by * dot4_u8(bg0, wB) + ay * dot4_u8(bg1, wB) + 512u;
  const uint32_t vr = by * dot4_u8(r0, wA) + ay * dot4_u8(r1, wA) + 512u;
  return (vb >> 10) | ((vg >> 10) << 8) | ((vr >> 10) << 16);
}
RIP_HD uint32_t remap_pixel_bgrx(const uint32_t* src, int rows, int cols, int pitch_px, float mx, float my) {
  return remap_pixel_bgrx_fix(src, rows, cols, pitch_px, remap_fix(mx), remap_fix(my));
}

// Packed fixed-point map entry: the 1/32-pixel source coordinate relative to the destination pixel,
// two int16: lo = sx - 32 * x, hi = sy - 32 * y.  Exactly the integers cv::remap derives from the float maps, at
// half the bytes.  REMAP_FAR marks "every tap is outside the image" (also used for NaN / infinite map values).
constexpr int REMAP_FAR = -32768;
RIP_HD bool remap_pack_entry(float mx, float my, int x, int y, int rows, int cols, uint32_t& packed) {
  const int sx = remap_fix(mx), sy = remap_fix(my);
  const int ix = sx >> 5, iy = sy >> 5;
  const bool any_tap = ix >= -1 && ix < cols && iy >= -1 && iy < rows;  // otherwise the result is 0 whatever the value
  if (!any_tap) { packed = (uint32_t)(uint16_t)REMAP_FAR | ((uint32_t)(uint16_t)REMAP_FAR << 16); return true; }
  const int dx = sx - 32 * x, dy = sy - 32 * y;
  if (dx <= REMAP_FAR || dx > 32767 || dy <= REMAP_FAR || dy > 32767) return false;  // does not fit: keep the float map
  packed = (uint32_t)(uint16_t)dx | ((uint32_t)(uint16_t)dy << 16);
  return true;
}
RIP_HD void remap_unpack_entry(uint32_t packed, int x, int y, int& sx, int& sy) {
  const int dx = (int)(int16_t)(packed & 0xffffu), dy = (int)(int16_t)(packed >> 16);
  // REMAP_FAR -> a coordinate no image contains: every tap fails the bounds tests and the result is 0
  sx = dx == REMAP_FAR ? INT32_MIN : 32 * x + dx;
  sy = 32 * y + dy;
}
RIP_HD uint32_t remap_pixel_bgrx_packed(const uint32_t* src, int rows, int cols, int pitch_px, uint32_t packed, int x, int y) {
  int sx, sy;
  remap_unpack_entry(packed, x, y, sx, sy);
  return remap_pixel_bgrx_fix(src, rows, cols, pitch_px, sx, sy);
}
// integer parts of a packed entry's displacement (floor(d / 32)); the low 5 bits of each half are the 1/32 fractions
RIP_HD int remap_packed_dxi(uint32_t packed) { return (int)(packed << 16) >> 21; }
RIP_HD int remap_packed_dyi(uint32_t packed) { return (int)packed >> 21; }
// The blend of remap_pixel_bgrx_fix for taps that are already in registers (tile kernel, rip_fast.cu).  The vertical
// weights carry a factor 64, which puts (V + 512) >> 10 into byte 2 of each sum (V * 64 + 2^15 < 2^24), so the three
// channels are packed by two byte permutes.  Returns b | g << 8 | r << 16.
RIP_HD uint32_t remap_blend_w(uint32_t t00, uint32_t t01, uint32_t t10, uint32_t t11, uint32_t ax, uint32_t ay6) {
  const uint32_t by6 = 2048u - ay6;
  const uint32_t wA = ax * 255u + 32u;  // bytes (32 - ax, ax, 0, 0)
  const uint32_t wB = wA << 16;         // bytes (0, 0, 32 - ax, ax)
  const uint32_t bg0 = prmt(t00, t01, 0x5140u), bg1 = prmt(t10, t11, 0x5140u);  // B00 B01 G00 G01
  const uint32_t r0 = prmt(t00, t01, 0x3362u), r1 = prmt(t10, t11, 0x3362u);    // R00 R01 0 0 (byte 3 of a pixel is 0)
  const uint32_t vb = by6 * dot4_u8(bg0, wA) + ay6 * dot4_u8(bg1, wA) + 0x8000u;
  const uint32_t vg = by6 * dot4_u8(bg0, wB) + ay6 * dot4_u8(bg1, wB) + 0x8000u;
  const uint32_t vr = by6 * dot4_u8(r0, wA) + ay6 * dot4_u8(r1, wA) + 0x8000u;
  return prmt(prmt(vb, vg, 0x3362u), vr, 0x7610u);  // bytes: vb.2, vg.2, vr.2, vr.3 (= 0)
}
RIP_HD uint32_t remap_blend(uint32_t t00, uint32_t t01, uint32_t t10, uint32_t t11, int sx, int sy) {
  return remap_blend_w(t00, t01, t10, t11, (uint32_t)(sx & 31), ((uint32_t)sy & 31u) << 6);
}

// ---- PCA white balance: white_balance.cpp:73-136 ----------------------------
// stats = { sum_b, sum_b2, sum_r, sum_r2, sum_g, max_b, max_g, max_r } as exact integers.
// Eigen::Matrix2f inverse (fixed-size closed form, fp32, no FMA) then the per-pixel
// addWeighted -> threshold(TRUNC 255) -> convertTo(CV_8U), tabulated for all 256 inputs.
struct PcaCoeff { float alpha_b, beta_b, alpha_r, beta_r; };

RIP_HD void pca_solve2(float s2, float s1, float m2, float m1, float v0, float v1, float& alpha, float& beta) {
  const float det = RIP_FSUB(RIP_FMUL(s2, m1), RIP_FMUL(m2, s1));
  const float invdet = 1.0f / det;  // IEEE division on both host and device
  const float i00 = RIP_FMUL(m1, invdet), i01 = RIP_FMUL(-s1, invdet);
  const float i10 = RIP_FMUL(-m2, invdet), i11 = RIP_FMUL(s2, invdet);
  alpha = RIP_FADD(RIP_FMUL(i00, v0), RIP_FMUL(i01, v1));
  beta = RIP_FADD(RIP_FMUL(i10, v0), RIP_FMUL(i11, v1));
}

RIP_HD PcaCoeff pca_coefficients(const unsigned long long* st) {
  PcaCoeff c;
  // cv::sum / cv::minMaxLoc return doubles holding exact integers; Eigen's `<<` narrows to float
  const float sum_b = (float)(double)st[0], sum_b2 = (float)(double)st[1];
  const float sum_r = (float)(double)st[2], sum_r2 = (float)(double)st[3];
  const float sum_g = (float)(double)st[4];
  const float max_b = (float)st[5], max_g = (float)st[6], max_r = (float)st[7];
  const float max_b2 = (float)(st[5] * st[5]), max_r2 = (float)(st[7] * st[7]);
  pca_solve2(sum_b2, sum_b, max_b2, max_b, sum_g, max_g, c.alpha_b, c.beta_b);
  pca_solve2(sum_r2, sum_r, max_r2, max_r, sum_g, max_g, c.alpha_r, c.beta_r);
  return c;
}

// cv::addWeighted(x^2, alpha, x, beta, 0) on CV_32F accumulates in double and rounds once;
// x^2*alpha and x*beta are exact in double (16+24 bits), so a plain double mul/add suffices.
RIP_HD int pca_lut_entry(int x, float alpha, float beta) {
#if defined(__CUDA_ARCH__)
  const double acc = __dadd_rn(__dmul_rn((double)(x * x), (double)alpha), __dmul_rn((double)x, (double)beta));
#else
  const double acc = (double)(x * x) * (double)alpha + (double)x * (double)beta;
#endif
  float y = (float)acc;
  if (y > 255.0f) y = 255.0f;  // THRESH_TRUNC (NaN stays NaN -> 0)
  return sat_u8_rint(y);
}

// ccc.cpp:383-386: cv::multiply(u8 image, Scalar(gain_b_, gain_g_, gain_r_)) with float gains
// widened to double -> saturate_cast<uchar>((double)x * (double)gain)  (see enhance())
RIP_HD int gain_lut_entry(int x, float gain) { return enh_gain_lut_entry(x, (double)gain); }

}  // namespace rip
