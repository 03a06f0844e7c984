// Per-pixel arithmetic of the RAW chain, written once for the sm_100a kernels
// (rip_kernels.cu).  It also compiles as plain host C++ so the CPU-only test harness
// (tests/hostsim) can check every formula exhaustively against cv2 *before* a kernel ever
// runs on a GPU.  The host build is test infrastructure; the product never calls it.
//
// Every function reproduces, bit for bit, what the reference's CPU path computes through
// OpenCV; the reference call site is cited on each.
//
// Instruction-count notes (the full chain is issue-bound on B200): integer
// <-> float conversions stay off the quarter-rate XU pipe (I2FP / F2IP / magic-number adds),
// 8-bit results are produced as the low byte of a word and assembled with byte permutes (PRMT),
// everything that is a function of one 8-bit value (gains, x * 1/255f, hue sector and fraction,
// gamma folded into sRGBGammaTab) is tabulated on the host (chain_tables.hpp).
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define RIP_HD __device__ __forceinline__
#define RIP_FMUL(a, b) __fmul_rn((a), (b))
#define RIP_FADD(a, b) __fadd_rn((a), (b))
#define RIP_FSUB(a, b) __fsub_rn((a), (b))
#define RIP_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define RIP_RINT_I(x) __float2int_rn(x)
#define RIP_TRUNC_I(x) __float2int_rz(x)
#define RIP_FLOORF(x) floorf(x)
#else
#include <cmath>
#include <cstring>
#define RIP_HD static inline
// host build is compiled with -ffp-contract=off, so these stay separately rounded
#define RIP_FMUL(a, b) ((a) * (b))
#define RIP_FADD(a, b) ((a) + (b))
#define RIP_FSUB(a, b) ((a) - (b))
#define RIP_FMA(a, b, c) fmaf((a), (b), (c))
#define RIP_RINT_I(x) ((int)lrintf(x))
#define RIP_TRUNC_I(x) ((int)(x))
#define RIP_FLOORF(x) floorf(x)
#endif

namespace rip {

// stage bits of the fused kernel (order = raw_image_pipeline.hpp:143-172)
enum : uint32_t {
  ST_WB = 1u << 0,     // per-frame white-balance LUT on B and R (pca) or B,G,R (ccc)
  ST_CC = 1u << 1,     // 3x3 colour calibration
  ST_GAMMA = 1u << 2,  // gamma LUT
  ST_VIG = 1u << 3,    // Lab-L vignetting
  ST_ENH = 1u << 4,    // HSV enhancer
  ST_ALL = 31u
};

constexpr int HUE_BIAS = 32;  // bgr_to_hsv leaves negative hues (-30 .. -1) un-wrapped; the table wraps them

// hue entry of the enhancer: everything HSV2BGR derives from the (gained) 8-bit hue
struct alignas(8) HueEntry {
  float f;       // fractional part of the sector coordinate
  uint32_t sel;  // PRMT selector putting (b, g, r) of that sector into bytes 0..2
};

// Lookup tables the chain needs (pointers into shared or global memory); see chain_tables.hpp.
struct ChainTables {
  const float* wbf;        // [3][256] per-frame B,G,R white-balance LUTs as floats (exact integers)
  const uint8_t* gamma;    // [256]   gamma LUT (identity when gamma is off)
  const uint16_t* g2;      // [256]   sRGBGammaTab_b[gamma[x]]  (gamma folded in when enabled)
  const uint16_t* lab_c;   // [2041]  LabCbrtTab_b
  const uint32_t* lab_yf;  // [256]   (ify << 16) | y
  const uint8_t* inv_g;    // [4096]  sRGBInvGammaTab_b
  const int32_t* sdiv;     // [256]
  const int32_t* hdiv;     // [256]
  const HueEntry* hue;     // [288]   indexed by HUE_BIAS + the un-gained, un-wrapped hue (-32 .. 255)
  const float* sf;         // [256]   float(sat_gain_lut[s]) * (1/255f)
  const float* vf;         // [256]   float(val_gain_lut[v]) * (1/255f)
};

struct ChainConsts {
  float cc[9];      // row-major, channel order B,G,R (Matx33f, color_calibration.cpp:78-79)
  float cc_bias[3];
  int wb_g_identity;  // 1: the G white-balance LUT is the identity (pca): used by the path without colour calibration
  int has_bias;     // 0: every cc_bias is +-0 and the add is skipped (chain_quad.cuh)
};
// fills the derived members from cc / cc_bias
RIP_HD void chain_consts_finish(ChainConsts& k) {
  k.has_bias = (k.cc_bias[0] != 0.0f || k.cc_bias[1] != 0.0f || k.cc_bias[2] != 0.0f) ? 1 : 0;  // NaN counts as a bias
}

RIP_HD int clamp_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// saturate_cast<uchar>(float): cvRound (half to even) then clamp.  Device: F2IP.U8.F32 (fast pipe).
RIP_HD int sat_u8_rint(float y) {
#if defined(__CUDA_ARCH__)
  unsigned r;
  asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(y));  // NaN -> 0 like x86's cvRound(NaN)=INT_MIN
  return (int)r;
#else
  if (!(y == y)) return 0;
  if (y <= -1.0f) return 0;
  if (y >= 256.0f) return 255;
  return clamp_u8((int)lrintf(y));
#endif
}

// exact (float)x for a small non-negative int without touching the conversion pipe
RIP_HD float u8_to_float(int x) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(__uint_as_float(0x4B000000u | (unsigned)x), 8388608.0f);
#else
  return (float)x;
#endif
}

// word whose LOW BYTE is trunc(x) for 0 <= x < 256 (upper bytes unspecified)
RIP_HD uint32_t trunc_u8_word(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(__fadd_rz(x, 8388608.0f));
#else
  return (uint32_t)((int)x & 255);
#endif
}

// byte permute: result byte i = byte (sel >> 4i & 7) of the 8-byte value {b, a}
RIP_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));  // selectors here never set a nibble's msb
  return r;
#else
  const uint64_t v = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 255) << (8 * i);
  return r;
#endif
}

RIP_HD uint32_t pack_bgr(int b, int g, int r) { return (uint32_t)b | ((uint32_t)g << 8) | ((uint32_t)r << 16); }

RIP_HD int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// ---- colour calibration: color_calibration.cpp:91-104 ---------------------
// cv::gemm on N x 3 fp32 == separately rounded products, summed left to right, then the
// bias add (cv::add with a Scalar), then convertTo(CV_8U).
RIP_HD void color_calibrate_f(float fb, float fg, float fr, const ChainConsts& k, int& b, int& g, int& r) {
  int o[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float t0 = RIP_FMUL(fb, k.cc[3 * j + 0]);
    float t1 = RIP_FMUL(fg, k.cc[3 * j + 1]);
    float t2 = RIP_FMUL(fr, k.cc[3 * j + 2]);
    // cv::add with the bias Scalar; a +0.0f bias changes nothing that survives the u8 conversion (-0 -> +0, NaN -> 0)
    const float y = RIP_FADD(RIP_FADD(RIP_FADD(t0, t1), t2), k.cc_bias[j]);
    o[j] = sat_u8_rint(y);
  }
  b = o[0]; g = o[1]; r = o[2];
}

// ---- 8-bit BGR -> Lab: cv::cvtColor(COLOR_BGR2Lab), vignetting_correction.cpp:73 (A.7) --
// `gt` is sRGBGammaTab_b, optionally with the gamma LUT folded in.  L, a, b are returned WITHOUT the saturate_cast:
// for every one of the 2^24 8-bit inputs they already lie in [0, 255] (checked exhaustively in
// tests/test_pixel_math_host.py).  (An fp32 evaluation of the dot products -- exact, all integers < 2^24 -- with the
// index taken by F2I.FLOOR was measured: it trades 8 integer-pipe operations for 5 on the conversion unit and is
// 0.6 % slower.)
RIP_HD void bgr_to_lab(int b, int g, int r, const uint16_t* gt, const uint16_t* lab_c, int& L, int& A, int& B) {
  const int R_ = gt[r], G_ = gt[g], B_ = gt[b];
  const int fX = lab_c[(R_ * 1777 + G_ * 1541 + B_ * 778 + 2048) >> 12];
  const int fY = lab_c[(R_ * 871 + G_ * 2929 + B_ * 296 + 2048) >> 12];
  const int fZ = lab_c[(R_ * 73 + G_ * 448 + B_ * 3575 + 2048) >> 12];
  L = (296 * fY - 1336934 + 16384) >> 15;
  A = (500 * (fX - fY) + 128 * 32768 + 16384) >> 15;
  B = (200 * (fY - fZ) + 128 * 32768 + 16384) >> 15;
}

// abToXZ_b, linear branch (v <= 3390): (v * 108) / 841 - 290 with C integer division (truncation
// toward zero).  |v * 108| < 900000, where n / 841 == (n * 5106977) >> 32 exactly.
RIP_HD int lab_xz_linear(int v) {
  const int n = v * 108;
  const unsigned an = (unsigned)(n < 0 ? -n : n);
#if defined(__CUDA_ARCH__)
  const int q = (int)__umulhi(an, 5106977u);
#else
  const int q = (int)(((uint64_t)an * 5106977u) >> 32);
#endif
  return (n < 0 ? -q : q) - 290;
}
// cubic branch (v > 3390 > 0): ((v * v) / 16384) * v / 16384, all operands positive
RIP_HD int lab_xz_cubic(int v) { return (int)(((unsigned)(((unsigned)(v * v) >> 14) * v)) >> 14); }

// ---- 8-bit Lab -> BGR: cv::cvtColor(COLOR_Lab2BGR), vignetting_correction.cpp:92 (A.8) --
RIP_HD void lab_to_bgr(int L, int A, int B, const ChainTables& t, int& b, int& g, int& r) {
  const uint32_t yf = t.lab_yf[L];
  const int y = (int)(yf & 0xffffu), ify = (int)(yf >> 16);
  const int fx = ify + ((A * 268435 + 128) >> 13) - 4194;        // 5 * 53687 = 268435
  const int fz = ify - ((B * 41943 + 16) >> 9) + 10484;          // -(bdiv): -(... - 10485 + 1)
  int x = lab_xz_cubic(fx), z = lab_xz_cubic(fz);
  if ((fx < fz ? fx : fz) <= 3390) {  // dark pixels only: whole warps skip this in ordinary image regions
    if (fx <= 3390) x = lab_xz_linear(fx);
    if (fz <= 3390) z = lab_xz_linear(fz);
  }
  int ro = (12615 * x - 6296 * y - 2223 * z + 8192) >> 14;
  int go = (-3773 * x + 7684 * y + 185 * z + 8192) >> 14;
  int bo = (217 * x - 836 * y + 4715 * z + 8192) >> 14;
  ro = ro < 0 ? 0 : (ro > 4095 ? 4095 : ro);
  go = go < 0 ? 0 : (go > 4095 ? 4095 : go);
  bo = bo < 0 ? 0 : (bo > 4095 ? 4095 : bo);
  r = t.inv_g[ro]; g = t.inv_g[go]; b = t.inv_g[bo];
}

// ---- vignetting: vignetting_correction.cpp:68-93 (A.6) ---------------------------------
// L' = saturate_cast<uchar>((float)L * mask)  (cv::multiply fp32, then convertTo CV_8U)
RIP_HD void vignetting(int& b, int& g, int& r, float mask, const ChainTables& t) {
  int L, A, B;
  bgr_to_lab(b, g, r, t.g2, t.lab_c, L, A, B);
  L = sat_u8_rint(RIP_FMUL(u8_to_float(L), mask));
  lab_to_bgr(L, A, B, t, b, g, r);
}

// ---- 8-bit BGR -> HSV (H in [0,180)): color_enhancer.cpp:40 (A.9) -----------------------
RIP_HD void bgr_to_hsv(int b, int g, int r, const ChainTables& t, int& h, int& s, int& v) {
  int vmax = b > g ? b : g; vmax = vmax > r ? vmax : r;
  int vmin = b < g ? b : g; vmin = vmin < r ? vmin : r;
  const int d = vmax - vmin;
  s = (d * t.sdiv[vmax] + 2048) >> 12;
  int hh = r - g + 4 * d;            // v == b
  if (vmax == g) hh = b - r + 2 * d;
  if (vmax == r) hh = g - b;         // tested first in OpenCV: highest priority
  h = (hh * t.hdiv[d] + 2048) >> 12;  // -30 .. 150; OpenCV adds 180 to negative values: done by the consumer
  v = vmax;
}

// ---- enhancer gains + 8-bit HSV -> BGR: color_enhancer.cpp:42-46 (A.9; fp32 with FMA) ------
// cv::multiply(u8 hsv, Scalar(hue_gain_, saturation_gain_, value_gain_)) works in double
// (saturate_cast<uchar>((double)c * gain)); its three 256-entry results are folded, together
// with everything HSV2RGB_b derives from a single channel, into t.hue / t.sf / t.vf:
//   hh = h' * (6/180f); if (hh >= 6) hh -= 6; sector = floor(hh); f = hh - sector
//   s = s' * (1/255f);  v = v' * (1/255f)
// OpenCV converts each image row in vector chunks of 32 pixels whose results are TRUNCATED to
// u8, and finishes the remaining (width % 32) pixels with scalar code that ROUNDS
// (saturate_cast) [probed exhaustively, cv2 4.13.0: both paths use the fused form 1 - s*f].
// `row_tail` = this pixel's column >= (width & ~31).  Returns b | g << 8 | r << 16.
RIP_HD uint32_t hsv_gain_to_bgr(int h, int s, int v, bool row_tail, const ChainTables& t) {
  const HueEntry he = t.hue[h + HUE_BIAS];  // h may be -30 .. -1 (un-wrapped)
  const float f = he.f, sf = t.sf[s], vf = t.vf[v];
  const float t1 = RIP_FMUL(vf, RIP_FSUB(1.0f, sf));
  const float t2 = RIP_FMUL(vf, RIP_FMA(-sf, f, 1.0f));
  const float t3 = RIP_FMUL(vf, RIP_FMA(-sf, RIP_FSUB(1.0f, f), 1.0f));
  const float c0 = RIP_FMUL(vf, 255.0f), c1 = RIP_FMUL(t1, 255.0f), c2 = RIP_FMUL(t2, 255.0f), c3 = RIP_FMUL(t3, 255.0f);
  uint32_t m0, m1, m2, m3;
  if (row_tail) {
    m0 = (uint32_t)sat_u8_rint(c0); m1 = (uint32_t)sat_u8_rint(c1); m2 = (uint32_t)sat_u8_rint(c2); m3 = (uint32_t)sat_u8_rint(c3);
  } else {
    m0 = trunc_u8_word(c0); m1 = trunc_u8_word(c1); m2 = trunc_u8_word(c2); m3 = trunc_u8_word(c3);
  }
  const uint32_t tab = prmt(prmt(m0, m1, 0x0040), prmt(m2, m3, 0x0040), 0x5410);  // bytes t0, t1, t2, t3
  return prmt(tab, 0u, he.sel);  // selector nibble 3 picks a zero byte
}

// ---- double-precision gain LUT entry: saturate_cast<uchar>((double)x * gain) -------------
// OpenCV picks the working depth of cv::multiply(u8, Scalar) from the scalar (arithm.cpp
// actualScalarDepth): any non-integer gain makes it CV_64F; integer gains give the same numbers
// in fp32.  [probed: 55*1.1 -> 61 (double), fp32 would give 60.]
RIP_HD int enh_gain_lut_entry(int x, double gain) {
  const double y = (double)x * gain;
  if (!(y == y)) return 0;
  if (y <= -1.0) return 0;
  if (y >= 256.0) return 255;
#if defined(__CUDA_ARCH__)
  return clamp_u8(__double2int_rn(y));
#else
  return clamp_u8((int)lrint(y));
#endif
}

// ---- the chain after debayer+flip: raw_image_pipeline.hpp:151-166 -----------------------
// in: (b, g, r) of the debayered pixel; returns the packed BGR8 result b | g << 8 | r << 16.
template <uint32_t STAGES>
RIP_HD uint32_t chain_pixel(int b, int g, int r, float mask, bool row_tail, const ChainConsts& k, const ChainTables& t) {
  if (STAGES & ST_CC) {
    float fb, fg, fr;
    if (STAGES & ST_WB) {  // white_balance.cpp:117-127 (pca) / ccc.cpp:383-386: per-frame LUTs, kept as floats
      fb = t.wbf[b]; fg = t.wbf[256 + g]; fr = t.wbf[512 + r];  // pca: the G table holds the identity
    } else {
      fb = u8_to_float(b); fg = u8_to_float(g); fr = u8_to_float(r);
    }
    color_calibrate_f(fb, fg, fr, k, b, g, r);
  } else if (STAGES & ST_WB) {
    b = RIP_TRUNC_I(t.wbf[b]); r = RIP_TRUNC_I(t.wbf[512 + r]);
    if (!k.wb_g_identity) g = RIP_TRUNC_I(t.wbf[256 + g]);
  }
  if (STAGES & ST_VIG) {
    vignetting(b, g, r, mask, t);  // gamma (if enabled) is folded into t.g2
  } else if (STAGES & ST_GAMMA) {  // gamma_correction.cpp:54-56 cv::LUT
    b = t.gamma[b]; g = t.gamma[g]; r = t.gamma[r];
  }
  if (STAGES & ST_ENH) {
    int h, s, v;
    bgr_to_hsv(b, g, r, t, h, s, v);
    return hsv_gain_to_bgr(h, s, v, row_tail, t);
  }
  return pack_bgr(b, g, r);
}

// ---- cv::remap INTER_LINEAR fixed point: undistortion.cpp:240-245 (A.10) ----------------
// sx = cvRound(mapx * 32); ix = sx >> 5; ax = sx & 31; weights (32-ay)(32-ax)*32 ... sum 2^15;
// out = (sum w*p + 2^14) >> 15, taps outside the image contribute 0 (BORDER_CONSTANT 0).
RIP_HD int remap_fix(float m) {
#if defined(__CUDA_ARCH__)
  return __float2int_rn(RIP_FMUL(m, 32.0f));
#else
  float v = m * 32.0f;
  if (!(v == v) || v >= 2147483648.0f || v < -2147483648.0f) return INT32_MIN;
  return (int)lrintf(v);
#endif
}

}  // namespace rip
