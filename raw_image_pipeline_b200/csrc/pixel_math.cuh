// Per-pixel arithmetic of the RAW chain, written once for the sm_100a kernels
// (rip_kernels.cu).  It also compiles as plain host C++ so the CPU-only test harness
// (tests/hostsim) can check every formula exhaustively against cv2 *before* a kernel ever
// runs on a GPU.  The host build is test infrastructure; the product never calls it.
//
// Every function reproduces, bit for bit, what the reference's CPU path computes through
// OpenCV; the reference call site is cited on each.
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

// Lookup tables the chain needs (pointers into shared or global memory).
struct ChainTables {
  const uint8_t* wb;        // [3][256]  per-frame B,G,R white-balance LUTs
  const uint8_t* gamma;     // [256]
  const uint16_t* srgb_g;   // [256]   sRGBGammaTab_b
  const uint16_t* lab_c;    // [2041]  LabCbrtTab_b
  const uint32_t* lab_yf;   // [256]   (ify << 16) | y
  const uint8_t* inv_g;     // [4096]  sRGBInvGammaTab_b
  const int32_t* sdiv;      // [256]
  const int32_t* hdiv;      // [256]
  const uint8_t* enh;       // [3][256] enhancer gain LUTs for H,S,V (see enhance())
};

struct ChainConsts {
  float cc[9];      // row-major, channel order B,G,R (Matx33f, color_calibration.cpp:78-79)
  float cc_bias[3];
};

RIP_HD int clamp_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// saturate_cast<uchar>(float): cvRound (half to even) then clamp.
RIP_HD int sat_u8_rint(float y) {
#if defined(__CUDA_ARCH__)
  // clamp first so the conversion cannot overflow; NaN -> 0 like x86's cvRound(NaN)=INT_MIN
  unsigned r;
  asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(y));
  return (int)r;
#else
  if (!(y == y)) return 0;
  if (y <= -1.0f) return 0;
  if (y >= 256.0f) return 255;
  return clamp_u8((int)lrintf(y));
#endif
}

RIP_HD int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// ---- colour calibration: color_calibration.cpp:91-104 ---------------------
// cv::gemm on N x 3 fp32 == separately rounded products, summed left to right, then the
// bias add (cv::add with a Scalar), then convertTo(CV_8U).
RIP_HD void color_calibrate(int& b, int& g, int& r, const ChainConsts& k) {
  const float fb = (float)b, fg = (float)g, fr = (float)r;
  int o[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float t0 = RIP_FMUL(fb, k.cc[3 * j + 0]);
    float t1 = RIP_FMUL(fg, k.cc[3 * j + 1]);
    float t2 = RIP_FMUL(fr, k.cc[3 * j + 2]);
    float y = RIP_FADD(RIP_FADD(RIP_FADD(t0, t1), t2), k.cc_bias[j]);
    o[j] = sat_u8_rint(y);
  }
  b = o[0]; g = o[1]; r = o[2];
}

// ---- 8-bit BGR -> Lab: cv::cvtColor(COLOR_BGR2Lab), vignetting_correction.cpp:73 (A.7) --
RIP_HD void bgr_to_lab(int b, int g, int r, const ChainTables& t, int& L, int& A, int& B) {
  const int R_ = t.srgb_g[r], G_ = t.srgb_g[g], B_ = t.srgb_g[b];
  const int fX = t.lab_c[descale(R_ * 1777 + G_ * 1541 + B_ * 778, 12)];
  const int fY = t.lab_c[descale(R_ * 871 + G_ * 2929 + B_ * 296, 12)];
  const int fZ = t.lab_c[descale(R_ * 73 + G_ * 448 + B_ * 3575, 12)];
  L = clamp_u8(descale(296 * fY - 1336934, 15));
  A = clamp_u8(descale(500 * (fX - fY) + 128 * 32768, 15));
  B = clamp_u8(descale(200 * (fY - fZ) + 128 * 32768, 15));
}

// abToXZ_b evaluated arithmetically (C integer division semantics)
RIP_HD int lab_ab_to_xz(int v) {
  if (v <= 3390) return (v * 108) / 841 - 290;
  return (((v * v) / 16384) * v) / 16384;
}

// ---- 8-bit Lab -> BGR: cv::cvtColor(COLOR_Lab2BGR), vignetting_correction.cpp:92 (A.8) --
RIP_HD void lab_to_bgr(int L, int A, int B, const ChainTables& t, int& b, int& g, int& r) {
  const uint32_t yf = t.lab_yf[L];
  const int y = (int)(yf & 0xffffu), ify = (int)(yf >> 16);
  const int adiv = ((5 * A * 53687 + 128) >> 13) - 4194;
  const int bdiv = ((B * 41943 + 16) >> 9) - 10485 + 1;
  const int x = lab_ab_to_xz(ify + adiv);
  const int z = lab_ab_to_xz(ify - bdiv);
  int ro = descale(12615 * x - 6296 * y - 2223 * z, 14);
  int go = descale(-3773 * x + 7684 * y + 185 * z, 14);
  int bo = descale(217 * x - 836 * y + 4715 * z, 14);
  ro = ro < 0 ? 0 : (ro > 4095 ? 4095 : ro);
  go = go < 0 ? 0 : (go > 4095 ? 4095 : go);
  bo = bo < 0 ? 0 : (bo > 4095 ? 4095 : bo);
  r = t.inv_g[ro]; g = t.inv_g[go]; b = t.inv_g[bo];
}

// ---- vignetting: vignetting_correction.cpp:68-93 (A.6) ---------------------------------
// L' = saturate_cast<uchar>((float)L * mask)  (cv::multiply fp32, then convertTo CV_8U)
RIP_HD void vignetting(int& b, int& g, int& r, float mask, const ChainTables& t) {
  int L, A, B;
  bgr_to_lab(b, g, r, t, L, A, B);
  L = sat_u8_rint(RIP_FMUL((float)L, mask));
  lab_to_bgr(L, A, B, t, b, g, r);
}

// ---- 8-bit BGR -> HSV (H in [0,180)): color_enhancer.cpp:40 (A.9) -----------------------
RIP_HD void bgr_to_hsv(int b, int g, int r, const ChainTables& t, int& h, int& s, int& v) {
  int vmax = b > g ? b : g; vmax = vmax > r ? vmax : r;
  int vmin = b < g ? b : g; vmin = vmin < r ? vmin : r;
  const int d = vmax - vmin;
  s = (d * t.sdiv[vmax] + 2048) >> 12;
  int hh = (vmax == r) ? (g - b) : ((vmax == g) ? (b - r + 2 * d) : (r - g + 4 * d));
  hh = (hh * t.hdiv[d] + 2048) >> 12;
  if (hh < 0) hh += 180;
  h = clamp_u8(hh);
  v = vmax;
}

// ---- 8-bit HSV -> BGR: color_enhancer.cpp:46 (A.9; fp32 with FMA) ------------------------
// OpenCV's HSV2RGB_b converts each image row in vector chunks of 32 pixels (AVX2 dispatch: 4 x 8
// lanes) whose results are TRUNCATED to u8, and finishes the remaining (width % 32) pixels of
// the row with scalar code that ROUNDS (saturate_cast) [probed exhaustively, cv2 4.13.0: both
// paths use the fused form 1 - s*f].  `row_tail` = this pixel's column >= (width & ~31).
RIP_HD void hsv_to_bgr(int h, int s, int v, bool row_tail, int& b, int& g, int& r) {
  float hh = RIP_FMUL((float)h, 6.0f / 180.0f);
  if (hh >= 6.0f) hh = RIP_FSUB(hh, 6.0f);
  const float secf = RIP_FLOORF(hh);
  const float f = RIP_FSUB(hh, secf);
  int sec = (int)secf;
  sec = sec < 0 ? 0 : (sec > 5 ? 5 : sec);
  const float sf = RIP_FMUL((float)s, 1.0f / 255.0f);
  const float vf = RIP_FMUL((float)v, 1.0f / 255.0f);
  const float t0 = vf;
  const float t1 = RIP_FMUL(vf, RIP_FSUB(1.0f, sf));
  const float t2 = RIP_FMUL(vf, RIP_FMA(-sf, f, 1.0f));
  const float t3 = RIP_FMUL(vf, RIP_FMA(-sf, RIP_FSUB(1.0f, f), 1.0f));
  // sector table (b,g,r): {1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0}
  float fb, fg, fr;
  switch (sec) {
    case 0: fb = t1; fg = t3; fr = t0; break;
    case 1: fb = t1; fg = t0; fr = t2; break;
    case 2: fb = t3; fg = t0; fr = t1; break;
    case 3: fb = t0; fg = t2; fr = t1; break;
    case 4: fb = t0; fg = t1; fr = t3; break;
    default: fb = t2; fg = t1; fr = t0; break;
  }
  fb = RIP_FMUL(fb, 255.0f); fg = RIP_FMUL(fg, 255.0f); fr = RIP_FMUL(fr, 255.0f);
  if (row_tail) {
    b = sat_u8_rint(fb); g = sat_u8_rint(fg); r = sat_u8_rint(fr);
  } else {
    b = RIP_TRUNC_I(fb) & 255; g = RIP_TRUNC_I(fg) & 255; r = RIP_TRUNC_I(fr) & 255;
  }
}

// ---- enhancer: color_enhancer.cpp:38-47 -------------------------------------------------
// cv::multiply(u8 image, Scalar(hue_gain_, saturation_gain_, value_gain_)): OpenCV picks the
// working depth from the scalar (arithm.cpp actualScalarDepth): any non-integer gain makes it
// CV_64F, i.e. saturate_cast<uchar>((double)c * gain) with the gain kept in double; integer
// gains give the same numbers in fp32.  [probed: 55*1.1 -> 61 (double), fp32 would give 60.]
// A double multiply per channel per pixel is poor use of the SM, so the three 256-entry
// products are tabulated on the host (enh_gain_lut_entry) -- exact by construction.
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

RIP_HD void enhance(int& b, int& g, int& r, bool row_tail, const ChainTables& t) {
  int h, s, v;
  bgr_to_hsv(b, g, r, t, h, s, v);
  h = t.enh[h]; s = t.enh[256 + s]; v = t.enh[512 + v];
  hsv_to_bgr(h, s, v, row_tail, b, g, r);
}

// ---- the chain after debayer+flip: raw_image_pipeline.hpp:151-166 -----------------------
template <uint32_t STAGES>
RIP_HD void chain_pixel(int& b, int& g, int& r, float mask, bool row_tail, const ChainConsts& k, const ChainTables& t) {
  if (STAGES & ST_WB) {  // white_balance.cpp:117-127 (pca) / ccc.cpp:383-386: per-frame LUTs
    b = t.wb[b]; g = t.wb[256 + g]; r = t.wb[512 + r];
  }
  if (STAGES & ST_CC) color_calibrate(b, g, r, k);
  if (STAGES & ST_GAMMA) {  // gamma_correction.cpp:54-56 cv::LUT
    b = t.gamma[b]; g = t.gamma[g]; r = t.gamma[r];
  }
  if (STAGES & ST_VIG) vignetting(b, g, r, mask, t);
  if (STAGES & ST_ENH) enhance(b, g, r, row_tail, t);
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
