// Host-side construction of the static lookup-table blob the fused kernel copies into shared
// memory (layout below), from OpenCV's colour-conversion constants (cv_tables.inc) and the
// pipeline parameters.  Plain C++ (no CUDA types): shared by the library (rip_api.cu) and the
// CPU-only test harness (tests/hostsim).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "cv_tables.inc"
#include "pixel_math.cuh"

namespace rip {

// byte offsets inside the blob
enum : int {
  OFF_GAMMA = 0,      // u8[256]       gamma LUT (identity if gamma is off)     gamma_correction.cpp:35-42
  OFF_INVG = 256,     // u8[4096]      sRGBInvGammaTab_b
  OFF_G2 = 4352,      // u16[256]      sRGBGammaTab_b[gamma[x]]
  OFF_LABC = 4864,    // u16[2048]     LabCbrtTab_b (2041 used)
  OFF_YF = 8960,      // u32[256]      LabToYF_b packed (ify << 16) | y
  OFF_SDIV = 9984,    // i32[256]
  OFF_HDIV = 11008,   // i32[256]
  OFF_HUE = 12032,    // HueEntry[288] enhancer hue gain + HSV2BGR sector/fraction, entry j <-> hue j - HUE_BIAS (-32 .. 255)
  OFF_SF = 14336,     // f32[256]      enhancer saturation gain, * 1/255f
  OFF_VF = 15360,     // f32[256]      enhancer value gain, * 1/255f
  TABLE_BYTES = 16384
};

struct ChainTableParams {
  bool gamma_enabled = false;
  double gamma_k = 1.0;
  double enh_gain[3] = {1.0, 1.0, 1.0};  // (hue_gain_, saturation_gain_, value_gain_) member values
  bool operator==(const ChainTableParams& o) const {
    return gamma_enabled == o.gamma_enabled && gamma_k == o.gamma_k && memcmp(enh_gain, o.enh_gain, sizeof enh_gain) == 0;
  }
};

// gamma_correction.cpp:38-41:  float f = i / 255.0; f = pow(f, k_); lut = saturate_cast<uchar>(f * 255.0);
inline void build_gamma_lut(double k, uint8_t lut[256]) {
  for (int i = 0; i < 256; ++i) {
    float f = (float)(i / 255.0);
    f = (float)pow((double)f, k);
    const double v = (double)f * 255.0;
    int iv;
    if (!(v == v)) iv = 0;
    else if (v <= -1.0) iv = 0;
    else if (v >= 256.0) iv = 255;
    else iv = (int)lrint(v);
    lut[i] = (uint8_t)(iv < 0 ? 0 : (iv > 255 ? 255 : iv));
  }
}

// PRMT selector for HSV2RGB_b's sector table {1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0} (b, g, r -> index
// into (t0, t1, t2, t3)); byte 3 of the result is don't-care.
inline uint32_t hsv_sector_selector(int sector) {
  static const uint32_t k[6] = {0x4031, 0x4201, 0x4103, 0x4120, 0x4310, 0x4012};  // b | g << 4 | r << 8 | (zero byte) << 12
  return k[sector < 0 ? 0 : (sector > 5 ? 5 : sector)];
}

inline void build_chain_blob(const ChainTableParams& q, uint8_t* blob) {
  memset(blob, 0, TABLE_BYTES);
  uint8_t* gamma = blob + OFF_GAMMA;
  if (q.gamma_enabled) build_gamma_lut(q.gamma_k, gamma);
  else for (int i = 0; i < 256; ++i) gamma[i] = (uint8_t)i;
  memcpy(blob + OFF_INVG, kSrgbInvGammaTab, sizeof kSrgbInvGammaTab);
  uint16_t* g2 = reinterpret_cast<uint16_t*>(blob + OFF_G2);
  for (int i = 0; i < 256; ++i) g2[i] = kSrgbGammaTab[gamma[i]];
  memcpy(blob + OFF_LABC, kLabCbrtTab, sizeof kLabCbrtTab);
  memcpy(blob + OFF_YF, kLabToYF, sizeof kLabToYF);
  memcpy(blob + OFF_SDIV, kHsvSdiv, sizeof kHsvSdiv);
  memcpy(blob + OFF_HDIV, kHsvHdiv, sizeof kHsvHdiv);
  HueEntry* hue = reinterpret_cast<HueEntry*>(blob + OFF_HUE);
  float* sf = reinterpret_cast<float*>(blob + OFF_SF);
  float* vf = reinterpret_cast<float*>(blob + OFF_VF);
  for (int i = 0; i < 256 + HUE_BIAS; ++i) {
    // hue entry i serves the raw hue i - HUE_BIAS; BGR2HSV wraps negative hues by +180 before the u8 store
    int h8 = i - HUE_BIAS;
    if (h8 < 0) h8 += 180;
    // color_enhancer.cpp:42 cv::multiply(hsv, Scalar(hue_gain_, saturation_gain_, value_gain_))
    const int h = enh_gain_lut_entry(h8, q.enh_gain[0]);
    // HSV2RGB_b: h * (6/180f), wrap, sector, fraction (fp32, each operation rounded on its own)
    volatile float hh = (float)h * (6.0f / 180.0f);
    if (hh >= 6.0f) hh = hh - 6.0f;
    const float secf = floorf(hh);
    volatile float f = hh - secf;
    hue[i].f = f;
    hue[i].sel = hsv_sector_selector((int)secf);
  }
  for (int i = 0; i < 256; ++i) {
    const int s = enh_gain_lut_entry(i, q.enh_gain[1]), v = enh_gain_lut_entry(i, q.enh_gain[2]);
    volatile float s1 = (float)s * (1.0f / 255.0f), v1 = (float)v * (1.0f / 255.0f);
    sf[i] = s1; vf[i] = v1;
  }
}

// ---- tables of the strip kernel (chain_quad.cuh StripTables) ---------------------------------------------------
// Byte offsets; the kernel copies the static parts of the blob verbatim into 4096-byte aligned shared memory, ordered so
// that a stage set needs a prefix: gamma only -> 256 B, + white balance -> 1 KB, + enhancer -> 7.25 KB, + vignetting ->
// 21 KB, + colour calibration after white balance -> 24 KB.  chain_quad.cuh relies on the alignments noted.  SOFF_WB and
// SOFF_WBF are holes in the blob: per-frame data the kernel fills.
enum : int {
  SOFF_GAMMA = 0,       // u8[256]
  SOFF_WB = 256,        // u8[3][256]       per-frame white-balance LUTs B, G, R (256-byte aligned)
  SOFF_SF = 1024,       // f32[256]         (1024-byte aligned)
  SOFF_HDIV = 2048,     // i32[256]
  SOFF_SV = 3072,       // {u32, u32}[256]  x = sdiv[v], y = float bits of the value-gain entry
  SOFF_HUE = 5120,      // HueEntry[288], selector over {t0 t1 0 0 | t2 t3 0 0}
  SOFF_LABC = 8192,     // u16[2048]        (4096-byte aligned)
  SOFF_INVG = 12288,    // u8[4096]
  SOFF_YF4 = 16384,     // {i32 x4}[256]    -6296 y + 8192, 7684 y + 8192, -836 y + 8192, (ify - 4194) << 16 | (ify + 10484)
  SOFF_G2 = 20480,      // u16[256]         2 * sRGBGammaTab_b[gamma[x]] + 1
  SOFF_WBF = 21504,     // f32[3][256]      per-frame white-balance LUTs as floats (1024-byte aligned each)
  STRIP_TABLE_BYTES = 24576,
  STRIP_BLOB_BYTES = 21504,  // the blob ends where the per-frame float tables start
  STRIP_ENH_END = 7424,
  STRIP_VIG_END = 20992
};

// `blob`: the legacy blob of the same parameters (build_chain_blob)
inline void build_strip_blob(const uint8_t* blob, uint8_t* sblob) {
  memset(sblob, 0, STRIP_BLOB_BYTES);
  memcpy(sblob + SOFF_GAMMA, blob + OFF_GAMMA, 256);
  memcpy(sblob + SOFF_SF, blob + OFF_SF, 1024);
  memcpy(sblob + SOFF_HDIV, blob + OFF_HDIV, 1024);
  const int32_t* sdiv = reinterpret_cast<const int32_t*>(blob + OFF_SDIV);
  const uint32_t* vf = reinterpret_cast<const uint32_t*>(blob + OFF_VF);
  uint32_t* sv = reinterpret_cast<uint32_t*>(sblob + SOFF_SV);
  for (int i = 0; i < 256; ++i) { sv[2 * i] = (uint32_t)sdiv[i]; sv[2 * i + 1] = vf[i]; }
  const HueEntry* hue = reinterpret_cast<const HueEntry*>(blob + OFF_HUE);
  HueEntry* hue2 = reinterpret_cast<HueEntry*>(sblob + SOFF_HUE);
  for (int i = 0; i < 256 + HUE_BIAS; ++i) {
    // legacy selector nibbles index bytes {t0 t1 t2 t3} (4 = a zero byte); here the candidates sit in {t0 t1 0 0 | t2 t3 0 0}
    static const uint32_t remap[8] = {0, 1, 4, 5, 2, 2, 2, 2};
    uint32_t sel = 0;
    for (int n = 0; n < 4; ++n) sel |= remap[(hue[i].sel >> (4 * n)) & 7u] << (4 * n);
    hue2[i].f = hue[i].f; hue2[i].sel = sel;
  }
  memcpy(sblob + SOFF_LABC, blob + OFF_LABC, 4096);
  memcpy(sblob + SOFF_INVG, blob + OFF_INVG, 4096);
  const uint32_t* yf = reinterpret_cast<const uint32_t*>(blob + OFF_YF);
  int32_t* yf4 = reinterpret_cast<int32_t*>(sblob + SOFF_YF4);
  for (int i = 0; i < 256; ++i) {
    const int y = (int)(yf[i] & 0xffffu), ify = (int)(yf[i] >> 16);
    // Lab2RGB: (c_x x + c_y y + c_z z + 8192) >> 14 with c_y = -6296 / 7684 / -836 for R / G / B
    yf4[4 * i + 0] = -6296 * y + 8192; yf4[4 * i + 1] = 7684 * y + 8192; yf4[4 * i + 2] = -836 * y + 8192;
    // fx = ify + adiv, adiv = (..) - 4194;  fz = ify - bdiv, bdiv = (..) - 10485 + 1
    yf4[4 * i + 3] = (int32_t)(((uint32_t)(uint16_t)(int16_t)(ify - 4194) << 16) | (uint32_t)(ify + 10484));
  }
  const uint16_t* g2 = reinterpret_cast<const uint16_t*>(blob + OFF_G2);
  uint16_t* g2s = reinterpret_cast<uint16_t*>(sblob + SOFF_G2);
  for (int i = 0; i < 256; ++i) g2s[i] = (uint16_t)(2 * g2[i] + 1);  // <= 4081
}

// pointers into a blob (host memory or shared memory); wbf is set by the caller
inline
#if defined(__CUDACC__)
__host__ __device__
#endif
ChainTables chain_tables_from_blob(const uint8_t* t, const float* wbf) {
  ChainTables c;
  c.wbf = wbf;
  c.gamma = t + OFF_GAMMA;
  c.inv_g = t + OFF_INVG;
  c.g2 = reinterpret_cast<const uint16_t*>(t + OFF_G2);
  c.lab_c = reinterpret_cast<const uint16_t*>(t + OFF_LABC);
  c.lab_yf = reinterpret_cast<const uint32_t*>(t + OFF_YF);
  c.sdiv = reinterpret_cast<const int32_t*>(t + OFF_SDIV);
  c.hdiv = reinterpret_cast<const int32_t*>(t + OFF_HDIV);
  c.hue = reinterpret_cast<const HueEntry*>(t + OFF_HUE);
  c.sf = reinterpret_cast<const float*>(t + OFF_SF);
  c.vf = reinterpret_cast<const float*>(t + OFF_VF);
  return c;
}

}  // namespace rip

#include "chain_quad.cuh"

namespace rip {

inline
#if defined(__CUDACC__)
__host__ __device__
#endif
StripTables strip_tables_at(taddr t) {  // t: address of the table block (device: shared-window address, 4096-byte aligned)
  StripTables c;
  c.gamma = t + SOFF_GAMMA;
  c.wb_b = t + SOFF_WB; c.wb_g = t + (SOFF_WB + 256); c.wb_r = t + (SOFF_WB + 512);
  c.sf = t + SOFF_SF; c.hdiv = t + SOFF_HDIV; c.sv = t + SOFF_SV; c.hue = t + SOFF_HUE;
  c.lab_c = t + SOFF_LABC; c.inv_g = t + SOFF_INVG; c.yf4 = t + SOFF_YF4;
  c.g2 = t + SOFF_G2;
  c.wbf_b = t + SOFF_WBF; c.wbf_g = t + (SOFF_WBF + 1024); c.wbf_r = t + (SOFF_WBF + 2048);
  return c;
}

}  // namespace rip
