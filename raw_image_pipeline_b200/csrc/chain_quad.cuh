// The chain after debayer+flip (raw_image_pipeline.hpp:151-166) for the four pixels a lane of the strip kernel owns, in
// the form with the fewest issued instructions.  Same results as pixel_math.cuh chain_pixel() -- which is what the generic
// kernels run and what tests/test_pixel_math_host.py checks exhaustively against cv2; tests/test_chain_quad_host.py checks
// this file against chain_pixel() over the whole 2^24 colour cube for every stage set.
//
// The full chain is bound by instruction issue; what this form saves over chain_pixel():
//   * inputs stay packed (byte k of Bw/Gw/Rw = pixel k); a channel that needs no table (no white balance, or the G
//     channel under pca) becomes the float 2^23 + v with ONE byte permute (then one subtraction: no conversion pipe);
//   * table addresses are explicit 32-bit shared-window addresses (see `taddr`): indices are merged into aligned table
//     bases by the instruction that masks / extracts them, so no lookup pays a separate base addition;
//   * BGR -> Lab: the sRGB table holds 2 G + 1; the rows of the XYZ matrix sum to 4096, which makes the rounding constant
//     of the descale implicit; float(L) is built as (v >> 15) | 0x4B000000 by one funnel shift;
//   * Lab -> BGR: the LabToYF entry of L' (one 128-bit load) carries y already multiplied by the three matrix coefficients
//     (+ the rounding constant) and both abToXZ offsets;
//   * sdiv[v] and the value-gain float share one 64-bit entry; 3-input min/max; HSV2BGR's four candidates are gathered
//     with three byte permutes; a zero bias is not added.
// Measured and rejected (profiles/r2_strip_kernel.md): 128-bit per-channel PRODUCT tables for the colour calibration and
// for the XYZ dot products cut 30 instructions per pixel but triple the shared-memory wavefronts (random 16-byte entries
// conflict 2-3 ways): the kernel becomes shared-memory bound, 9.3 vs 5.5 ms per 64 x 12 MP.
#pragma once
#include "pixel_math.cuh"

namespace rip {

struct alignas(8) Pair32 { uint32_t x, y; };          // one 64-bit shared-memory load
struct alignas(16) Quad32 { uint32_t x, y, z, w; };    // one 128-bit shared-memory load

// Table addresses.  On the device they are 32-bit shared-window addresses and every lookup is an explicit ld.shared, so
// that the address arithmetic is exactly what is written here (the compiler otherwise adds the window base with a
// separate instruction per lookup): a byte index is merged into a 256-byte aligned base by the byte permute that extracts
// it, a masked index is OR-ed into an aligned base by the same LOP3 that masks it.  On the host (test build) they are
// byte pointers.
#if defined(__CUDA_ARCH__)
typedef uint32_t taddr;
RIP_HD uint32_t lds_u8(taddr a) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
RIP_HD uint32_t lds_u16(taddr a) { uint32_t v; asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
RIP_HD uint32_t lds_u32(taddr a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
RIP_HD Pair32 lds_u64(taddr a) { Pair32 v; asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
RIP_HD Quad32 lds_u128(taddr a) {
  Quad32 v;
  asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
// base + byte K of w; base % 256 == 0
template <int K> RIP_HD taddr taddr_byte(taddr base, uint32_t w) { return prmt(w, base, 0x7650u + K); }
// base + (v & mask); base % (mask + 1 rounded up to a power of two) == 0
RIP_HD taddr taddr_masked(taddr base, uint32_t v, uint32_t mask) { return (v & mask) | base; }
// the same with the byte chosen at run time: sel = 0x7650 + byte number
RIP_HD taddr taddr_byte_sel(taddr base, uint32_t w, uint32_t sel) { return prmt(w, base, sel); }
#else
typedef const uint8_t* taddr;
RIP_HD uint32_t lds_u8(taddr a) { return *a; }
RIP_HD uint32_t lds_u16(taddr a) { uint16_t v; memcpy(&v, a, 2); return v; }
RIP_HD uint32_t lds_u32(taddr a) { uint32_t v; memcpy(&v, a, 4); return v; }
RIP_HD Pair32 lds_u64(taddr a) { Pair32 v; memcpy(&v, a, 8); return v; }
RIP_HD Quad32 lds_u128(taddr a) { Quad32 v; memcpy(&v, a, 16); return v; }
template <int K> RIP_HD taddr taddr_byte(taddr base, uint32_t w) { return base + ((w >> (8 * K)) & 255u); }
RIP_HD taddr taddr_masked(taddr base, uint32_t v, uint32_t mask) { return base + (v & mask); }
RIP_HD taddr taddr_byte_sel(taddr base, uint32_t w, uint32_t sel) { return base + ((w >> (8 * (sel & 3u))) & 255u); }
#endif
// base + 4 * byte K of w (a table of 32-bit entries); base % 1024 == 0
template <int K> RIP_HD taddr taddr_byte4(taddr base, uint32_t w) {
  return taddr_masked(base, K == 0 ? (w << 2) : (w >> (8 * K - 2)), 0x3fcu);
}

#if defined(__CUDACC__)
// table address of a shared-memory pointer (device code; the host branch only serves nvcc's host pass, which parses kernels)
__device__ __forceinline__ taddr taddr_of_shared(const void* p) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__cvta_generic_to_shared(p);
#else
  return static_cast<const uint8_t*>(p);
#endif
}
#endif

// tables of the strip kernel (chain_tables.hpp lays them out: SOFF_*)
struct StripTables {
  taddr gamma;             // u8[256]
  taddr wb_b, wb_g, wb_r;  // u8[256] each, 256-byte aligned: per-frame white-balance LUTs (stage sets without colour calibration)
  taddr wbf_b, wbf_g, wbf_r;  // f32[256] each, 1024-byte aligned: the same as floats (stage sets with colour calibration)
  taddr g2;                // u16[256]  2 * sRGBGammaTab_b[gamma[x]] + 1
  taddr sf;                // f32[256], 1024-byte aligned
  taddr hdiv;              // i32[256]
  taddr sv;                // {u32,u32}[256]  x = sdiv[v], y = float bits of (value gain lut)[v] * (1/255f)
  taddr hue;               // HueEntry[288], selector for prmt({t0 t1 0 0}, {t2 t3 0 0})
  taddr lab_c;             // u16[2048] LabCbrtTab_b, 4096-byte aligned
  taddr inv_g;             // u8[4096]  sRGBInvGammaTab_b
  taddr yf4;               // {i32 x4}[256]: -6296 y + 8192, 7684 y + 8192, -836 y + 8192, (ify - 4194) << 16 | (ify + 10484)
};

RIP_HD float bits_to_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
RIP_HD int max3i(int a, int b, int c) {
#if defined(__CUDA_ARCH__)
  return __vimax3_s32(a, b, c);
#else
  const int m = a > b ? a : b; return m > c ? m : c;
#endif
}
RIP_HD int min3i(int a, int b, int c) {
#if defined(__CUDA_ARCH__)
  return __vimin3_s32(a, b, c);
#else
  const int m = a < b ? a : b; return m < c ? m : c;
#endif
}
RIP_HD uint32_t funnel_shift_r(uint32_t lo, uint32_t hi, int shift) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, shift);
#else
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> shift);
#endif
}

// float 2^23 + byte K of `w` (exact): bytes {w.K, 00, 00, 4B}
template <int K>
RIP_HD float biased_float_of_byte(uint32_t w) { return bits_to_float(prmt(w, 0x4B000000u, 0x7650u + K)); }

// TAIL: the pixel lies in cv2's scalar row tail of HSV2BGR (columns >= width & ~31), which rounds where the vector loop
// truncates (pixel_math.cuh hsv_gain_to_bgr).  WBG: the G channel has a white-balance table too (ccc; pca leaves G
// untouched).  A colour calibration with a non-zero bias is not handled here (the strip kernel leaves such
// configurations to the tile kernel, so no instantiation carries a bias add).
template <uint32_t STAGES, int K, bool TAIL = false, bool WBG = true>
RIP_HD uint32_t chain_px(uint32_t Bw, uint32_t Gw, uint32_t Rw, float mask, const ChainConsts& k, const StripTables& t) {
  int b = 0, g = 0, r = 0;
  if (STAGES & ST_CC) {
    float fb, fg, fr;
    if (STAGES & ST_WB) {  // white_balance.cpp:117-127 (pca) / ccc.cpp:383-386: per-frame LUTs, kept as floats
      fb = bits_to_float(lds_u32(taddr_byte4<K>(t.wbf_b, Bw)));
      fr = bits_to_float(lds_u32(taddr_byte4<K>(t.wbf_r, Rw)));
      fg = WBG ? bits_to_float(lds_u32(taddr_byte4<K>(t.wbf_g, Gw))) : RIP_FSUB(biased_float_of_byte<K>(Gw), 8388608.0f);
    } else {  // (float)v = (2^23 + v) - 2^23 exactly
      fb = RIP_FSUB(biased_float_of_byte<K>(Bw), 8388608.0f); fg = RIP_FSUB(biased_float_of_byte<K>(Gw), 8388608.0f);
      fr = RIP_FSUB(biased_float_of_byte<K>(Rw), 8388608.0f);
    }
    // cv::gemm on N x 3 fp32 == separately rounded products, summed left to right; the nine matrix entries
    // are constant-bank operands of the multiplies
    float y[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      y[j] = RIP_FADD(RIP_FADD(RIP_FMUL(fb, k.cc[3 * j + 0]), RIP_FMUL(fg, k.cc[3 * j + 1])), RIP_FMUL(fr, k.cc[3 * j + 2]));
    b = sat_u8_rint(y[0]); g = sat_u8_rint(y[1]); r = sat_u8_rint(y[2]);
  } else if (STAGES & ST_WB) {
    b = (int)lds_u8(taddr_byte<K>(t.wb_b, Bw)); r = (int)lds_u8(taddr_byte<K>(t.wb_r, Rw));
    g = WBG ? (int)lds_u8(taddr_byte<K>(t.wb_g, Gw)) : (int)prmt(Gw, 0u, 0x4440u + K);
  } else {
    b = (int)prmt(Bw, 0u, 0x4440u + K); g = (int)prmt(Gw, 0u, 0x4440u + K); r = (int)prmt(Rw, 0u, 0x4440u + K);
  }
  if (STAGES & ST_VIG) {  // vignetting_correction.cpp:68-93; gamma (if enabled) is folded into the forward tables
    // t.g2 holds 2 G + 1 (G = sRGBGammaTab_b entry): every row of the XYZ matrix sums to 4096, so
    // sum c_i (2 G_i + 1) = 2 (sum c_i G_i + 2048) -- the rounding constant of the descale comes for free, and
    // LabCbrtTab_b[(dot + 2048) >> 12] sits at byte offset 2 * index = (sum >> 12) & 0x1ffe (sum < 2^25)
    const int R_ = (int)lds_u16(t.g2 + r + r), G_ = (int)lds_u16(t.g2 + g + g), B_ = (int)lds_u16(t.g2 + b + b);
    const int fX = (int)lds_u16(taddr_masked(t.lab_c, (uint32_t)(R_ * 1777 + G_ * 1541 + B_ * 778) >> 12, 0x1ffeu));
    const int fY = (int)lds_u16(taddr_masked(t.lab_c, (uint32_t)(R_ * 871 + G_ * 2929 + B_ * 296) >> 12, 0x1ffeu));
    const int fZ = (int)lds_u16(taddr_masked(t.lab_c, (uint32_t)(R_ * 73 + G_ * 448 + B_ * 3575) >> 12, 0x1ffeu));
    // L = (296 fY - 1336934 + 16384) >> 15 lies in 0..255 for every 8-bit input (tests/test_pixel_math_host.py), so
    // (v >> 15) | 0x4B000000 is the float 2^23 + L
    const float xL = bits_to_float(funnel_shift_r((uint32_t)(296 * fY - 1320550), 0x2580u, 15));
    const int Lp = sat_u8_rint(RIP_FMUL(RIP_FSUB(xL, 8388608.0f), mask));
    const int A = (500 * (fX - fY) + 4210688) >> 15;   // 128 * 32768 + 16384
    const int B = (200 * (fY - fZ) + 4210688) >> 15;
    const Quad32 yf = lds_u128(t.yf4 + 16 * Lp);
    const int fx = ((int)yf.w >> 16) + ((A * 268435 + 128) >> 13);
    const int fz = (int)(yf.w & 0xffffu) - ((B * 41943 + 16) >> 9);
    int x = lab_xz_cubic(fx), z = lab_xz_cubic(fz);
    if ((fx < fz ? fx : fz) <= 3390) {  // dark pixels only
      if (fx <= 3390) x = lab_xz_linear(fx);
      if (fz <= 3390) z = lab_xz_linear(fz);
    }
    int ro = (12615 * x + (int)yf.x - 2223 * z) >> 14;
    int go = (-3773 * x + (int)yf.y + 185 * z) >> 14;
    int bo = (217 * x + (int)yf.z + 4715 * z) >> 14;
    ro = ro < 0 ? 0 : (ro > 4095 ? 4095 : ro);
    go = go < 0 ? 0 : (go > 4095 ? 4095 : go);
    bo = bo < 0 ? 0 : (bo > 4095 ? 4095 : bo);
    r = (int)lds_u8(t.inv_g + ro); g = (int)lds_u8(t.inv_g + go); b = (int)lds_u8(t.inv_g + bo);
  } else if (STAGES & ST_GAMMA) {  // gamma_correction.cpp:54-56 cv::LUT
    b = (int)lds_u8(t.gamma + b); g = (int)lds_u8(t.gamma + g); r = (int)lds_u8(t.gamma + r);
  }
  if (STAGES & ST_ENH) {  // color_enhancer.cpp:38-47
    const int vmax = max3i(b, g, r), vmin = min3i(b, g, r);
    const int d = vmax - vmin;
    const Pair32 sv = lds_u64(t.sv + 8 * vmax);
    int hh = r - g + 4 * d;
    if (vmax == g) hh = b - r + 2 * d;
    if (vmax == r) hh = g - b;
    const int h = (hh * (int)lds_u32(t.hdiv + 4 * d) + 2048) >> 12;  // -30 .. 150, wrapped by the table
    const Pair32 he = lds_u64(t.hue + 8 * (h + HUE_BIAS));             // HueEntry {f, sel}
    // s = (d * sdiv[v] + 2048) >> 12 in 0..255: the float table entry sits at byte offset 4 s = ((..) >> 10) & 0x3fc
    const float sf = bits_to_float(lds_u32(taddr_masked(t.sf, (uint32_t)(d * (int)sv.x + 2048) >> 10, 0x3fcu)));
    const float f = bits_to_float(he.x), vf = bits_to_float(sv.y);
    const float t1 = RIP_FMUL(vf, RIP_FSUB(1.0f, sf));
    const float t2 = RIP_FMUL(vf, RIP_FMA(-sf, f, 1.0f));
    const float t3 = RIP_FMUL(vf, RIP_FMA(-sf, RIP_FSUB(1.0f, f), 1.0f));
    const float c0 = RIP_FMUL(vf, 255.0f), c1 = RIP_FMUL(t1, 255.0f), c2 = RIP_FMUL(t2, 255.0f), c3 = RIP_FMUL(t3, 255.0f);
    uint32_t m0, m1, m2, m3;
    if (TAIL) {
      m0 = (uint32_t)sat_u8_rint(c0); m1 = (uint32_t)sat_u8_rint(c1); m2 = (uint32_t)sat_u8_rint(c2); m3 = (uint32_t)sat_u8_rint(c3);
    } else {
      m0 = trunc_u8_word(c0); m1 = trunc_u8_word(c1); m2 = trunc_u8_word(c2); m3 = trunc_u8_word(c3);
    }
    // truncated words are 0x4B0000vv on the device (vv on the host), rounded ones are vv: byte 1 is a zero either way
    return prmt(prmt(m0, m1, 0x1140u), prmt(m2, m3, 0x1140u), he.y);
  }
  return pack_bgr(b, g, r);
}

template <uint32_t STAGES, bool WBG = true>
RIP_HD void chain_quad(uint32_t Bw, uint32_t Gw, uint32_t Rw, const float m[4], const ChainConsts& k, const StripTables& t, uint32_t px[4]) {
  px[0] = chain_px<STAGES, 0, false, WBG>(Bw, Gw, Rw, m[0], k, t);
  px[1] = chain_px<STAGES, 1, false, WBG>(Bw, Gw, Rw, m[1], k, t);
  px[2] = chain_px<STAGES, 2, false, WBG>(Bw, Gw, Rw, m[2], k, t);
  px[3] = chain_px<STAGES, 3, false, WBG>(Bw, Gw, Rw, m[3], k, t);
}

}  // namespace rip
