// TEST INFRASTRUCTURE ONLY.  Compiles the kernels' per-pixel arithmetic (pixel_math.cuh,
// frame_math.cuh) as plain host C++ so the CPU-only test-suite can compare every formula
// exhaustively with cv2 before the CUDA build ever meets a GPU.  Never linked into, loaded
// by, or shipped with the product library (raw_image_pipeline_b200/librip_b200.so).
#include <cstring>
#include <vector>
#include "../../raw_image_pipeline_b200/csrc/ccc_math.cuh"
#include "../../raw_image_pipeline_b200/csrc/chain_tables.hpp"

using namespace rip;

// blob with the enhancer gains folded in; `gamma` (optional) replaces the gamma LUT and is folded into g2
static std::vector<uint8_t> make_blob(const double* enh, const uint8_t* gamma) {
  ChainTableParams q;
  if (enh) { q.enh_gain[0] = enh[0]; q.enh_gain[1] = enh[1]; q.enh_gain[2] = enh[2]; }
  std::vector<uint8_t> blob(TABLE_BYTES);
  build_chain_blob(q, blob.data());
  if (gamma) {
    memcpy(blob.data() + OFF_GAMMA, gamma, 256);
    uint16_t* g2 = reinterpret_cast<uint16_t*>(blob.data() + OFF_G2);
    for (int i = 0; i < 256; ++i) g2[i] = kSrgbGammaTab[gamma[i]];
  }
  return blob;
}

extern "C" {

// raw (unclamped) L, a, b as int32 triples: the kernels rely on them being 0..255 already
void hs_bgr2lab_raw(long n, const uint8_t* in, int* out) {
  const std::vector<uint8_t> blob = make_blob(nullptr, nullptr);
  const ChainTables t = chain_tables_from_blob(blob.data(), nullptr);
  for (long i = 0; i < n; ++i) bgr_to_lab(in[3 * i], in[3 * i + 1], in[3 * i + 2], t.g2, t.lab_c, out[3 * i], out[3 * i + 1], out[3 * i + 2]);
}
void hs_bgr2lab(long n, const uint8_t* in, uint8_t* out) {
  const std::vector<uint8_t> blob = make_blob(nullptr, nullptr);
  const ChainTables t = chain_tables_from_blob(blob.data(), nullptr);
  for (long i = 0; i < n; ++i) {
    int L, A, B; bgr_to_lab(in[3 * i], in[3 * i + 1], in[3 * i + 2], t.g2, t.lab_c, L, A, B);
    out[3 * i] = L; out[3 * i + 1] = A; out[3 * i + 2] = B;
  }
}
void hs_lab2bgr(long n, const uint8_t* in, uint8_t* out) {
  const std::vector<uint8_t> blob = make_blob(nullptr, nullptr);
  const ChainTables t = chain_tables_from_blob(blob.data(), nullptr);
  for (long i = 0; i < n; ++i) {
    int b, g, r; lab_to_bgr(in[3 * i], in[3 * i + 1], in[3 * i + 2], t, b, g, r);
    out[3 * i] = b; out[3 * i + 1] = g; out[3 * i + 2] = r;
  }
}
void hs_bgr2hsv(long n, const uint8_t* in, uint8_t* out) {
  const std::vector<uint8_t> blob = make_blob(nullptr, nullptr);
  const ChainTables t = chain_tables_from_blob(blob.data(), nullptr);
  for (long i = 0; i < n; ++i) {
    int h, s, v; bgr_to_hsv(in[3 * i], in[3 * i + 1], in[3 * i + 2], t, h, s, v);
    out[3 * i] = h < 0 ? h + 180 : h; out[3 * i + 1] = s; out[3 * i + 2] = v;  // the wrap the hue table applies
  }
}
// `width`: row length of the image the n pixels form (selects OpenCV's scalar row tail)
void hs_hsv2bgr(long n, int width, const uint8_t* in, uint8_t* out) {
  const std::vector<uint8_t> blob = make_blob(nullptr, nullptr);  // unit gains
  const ChainTables t = chain_tables_from_blob(blob.data(), nullptr);
  for (long i = 0; i < n; ++i) {
    const bool tail = (int)(i % width) >= (width & ~31);
    const uint32_t p = hsv_gain_to_bgr(in[3 * i], in[3 * i + 1], in[3 * i + 2], tail, t);
    out[3 * i] = p & 255; out[3 * i + 1] = (p >> 8) & 255; out[3 * i + 2] = (p >> 16) & 255;
  }
}

// full per-pixel chain exactly as the fused kernel runs it; mask may be null when ST_VIG is off
void hs_chain(unsigned stages, long n, int width, const uint8_t* in, const float* mask, const float* cc, const float* bias,
              const double* enh, const uint8_t* wb, const uint8_t* gamma, uint8_t* out) {
  const std::vector<uint8_t> blob = make_blob(enh, (stages & ST_GAMMA) ? gamma : nullptr);
  float wbf[768];
  bool g_identity = true;
  for (int i = 0; i < 768; ++i) wbf[i] = (float)wb[i];
  for (int i = 0; i < 256; ++i) g_identity = g_identity && wb[256 + i] == i;
  const ChainTables t = chain_tables_from_blob(blob.data(), wbf);
  ChainConsts k;
  memcpy(k.cc, cc, sizeof k.cc); memcpy(k.cc_bias, bias, sizeof k.cc_bias);
  k.wb_g_identity = g_identity ? 1 : 0;
  for (long i = 0; i < n; ++i) {
    const int b = in[3 * i], g = in[3 * i + 1], r = in[3 * i + 2];
    const float m = mask ? mask[i] : 1.0f;
    const bool tail = (int)(i % width) >= (width & ~31);
    uint32_t p;
    switch (stages & ST_ALL) {
#define RIP_CASE(S) case S: p = chain_pixel<S>(b, g, r, m, tail, k, t); break;
      RIP_CASE(0) RIP_CASE(1) RIP_CASE(2) RIP_CASE(3) RIP_CASE(4) RIP_CASE(5) RIP_CASE(6) RIP_CASE(7)
      RIP_CASE(8) RIP_CASE(9) RIP_CASE(10) RIP_CASE(11) RIP_CASE(12) RIP_CASE(13) RIP_CASE(14) RIP_CASE(15)
      RIP_CASE(16) RIP_CASE(17) RIP_CASE(18) RIP_CASE(19) RIP_CASE(20) RIP_CASE(21) RIP_CASE(22) RIP_CASE(23)
      RIP_CASE(24) RIP_CASE(25) RIP_CASE(26) RIP_CASE(27) RIP_CASE(28) RIP_CASE(29) RIP_CASE(30) RIP_CASE(31)
#undef RIP_CASE
      default: p = 0;
    }
    out[3 * i] = p & 255; out[3 * i + 1] = (p >> 8) & 255; out[3 * i + 2] = (p >> 16) & 255;
  }
}

// the same chain through the strip kernel's four-pixel form (chain_quad.cuh) and its table layout; n % 4 == 0.
// `tail` != 0: every pixel is treated as a row-tail pixel of HSV2BGR (the strip kernel's out-of-line variant)
void hs_chain_quad(unsigned stages, long n, int tail, const uint8_t* in, const float* mask, const float* cc, const float* bias,
                   const double* enh, const uint8_t* wb, const uint8_t* gamma, uint8_t* out) {
  (void)bias;  // the strip kernel leaves configurations with a bias to the tile kernel
  const std::vector<uint8_t> blob = make_blob(enh, (stages & ST_GAMMA) ? gamma : nullptr);
  std::vector<uint8_t> tab(STRIP_TABLE_BYTES);
  build_strip_blob(blob.data(), tab.data());
  bool g_identity = true;
  for (int i = 0; i < 256; ++i) g_identity = g_identity && wb[256 + i] == i;
  memcpy(tab.data() + SOFF_WB, wb, 768);
  for (int i = 0; i < 768; ++i) { const float f = (float)wb[i]; memcpy(tab.data() + SOFF_WBF + 4 * i, &f, 4); }  // as the kernel fills them
  const StripTables t = strip_tables_at(tab.data());
  ChainConsts k;
  memcpy(k.cc, cc, sizeof k.cc); k.cc_bias[0] = k.cc_bias[1] = k.cc_bias[2] = 0.0f; k.wb_g_identity = g_identity ? 1 : 0;
  chain_consts_finish(k);
  for (long i = 0; i < n; i += 4) {
    uint32_t Bw = 0, Gw = 0, Rw = 0;
    float m[4];
    for (int q = 0; q < 4; ++q) {
      Bw |= (uint32_t)in[3 * (i + q)] << (8 * q); Gw |= (uint32_t)in[3 * (i + q) + 1] << (8 * q); Rw |= (uint32_t)in[3 * (i + q) + 2] << (8 * q);
      m[q] = mask ? mask[i + q] : 1.0f;
    }
    uint32_t px[4];
    switch (stages & ST_ALL) {
    // the strip kernel's instantiations: no G table under pca (identity)
#define RIP_CASE(S) case S: \
      if (tail) for (int q = 0; q < 4; ++q) px[q] = chain_px<S, 0, true, true>(Bw >> (8 * q), Gw >> (8 * q), Rw >> (8 * q), m[q], k, t); \
      else if (g_identity) chain_quad<S, false>(Bw, Gw, Rw, m, k, t, px); \
      else chain_quad<S, true>(Bw, Gw, Rw, m, k, t, px); \
      break;
      RIP_CASE(0) RIP_CASE(1) RIP_CASE(2) RIP_CASE(3) RIP_CASE(4) RIP_CASE(5) RIP_CASE(6) RIP_CASE(7)
      RIP_CASE(8) RIP_CASE(9) RIP_CASE(10) RIP_CASE(11) RIP_CASE(12) RIP_CASE(13) RIP_CASE(14) RIP_CASE(15)
      RIP_CASE(16) RIP_CASE(17) RIP_CASE(18) RIP_CASE(19) RIP_CASE(20) RIP_CASE(21) RIP_CASE(22) RIP_CASE(23)
      RIP_CASE(24) RIP_CASE(25) RIP_CASE(26) RIP_CASE(27) RIP_CASE(28) RIP_CASE(29) RIP_CASE(30) RIP_CASE(31)
#undef RIP_CASE
      default: px[0] = px[1] = px[2] = px[3] = 0;
    }
    for (int q = 0; q < 4; ++q) {
      out[3 * (i + q)] = px[q] & 255; out[3 * (i + q) + 1] = (px[q] >> 8) & 255; out[3 * (i + q) + 2] = (px[q] >> 16) & 255;
    }
  }
}

// mode 0: demosaic_at everywhere; mode 1: demosaic_quad where legal; mode 2: demosaic_quad_swar where legal
void hs_demosaic(const uint8_t* raw, int rows, int cols, int cfa, int angle, int mode, uint8_t* out) {
  const int orows = (angle == 90 || angle == 270) ? cols : rows;
  const int ocols = (angle == 90 || angle == 270) ? rows : cols;
  for (int oy = 0; oy < orows; ++oy)
    for (int ox = 0; ox < ocols; ++ox) {
      int iy, ix; flip_source(angle, rows, cols, oy, ox, iy, ix);
      int b, g, r;
      const int x4 = ix & ~3;
      if (mode >= 1 && x4 >= 4 && x4 + 7 < cols && x4 + 3 <= cols - 2) {
        int yc = iy < 1 ? 1 : (iy > rows - 2 ? rows - 2 : iy);
        uint32_t w[3][3];
        for (int rr = 0; rr < 3; ++rr)
          for (int j = 0; j < 3; ++j) memcpy(&w[rr][j], raw + (size_t)(yc - 1 + rr) * cols + x4 - 4 + 4 * j, 4);
        const bool row_has_r = ((yc & 1) == ((cfa >> 1) & 1));
        const int cpar = row_has_r ? (cfa & 1) : ((cfa & 1) ^ 1);
        if (mode == 2) {
          uint32_t Bw, Gw, Rw;
          demosaic_quad_swar(w, row_has_r, cpar, Bw, Gw, Rw);
          b = (Bw >> (8 * (ix & 3))) & 255; g = (Gw >> (8 * (ix & 3))) & 255; r = (Rw >> (8 * (ix & 3))) & 255;
        } else {
          int bb[4], gg[4], rr4[4];
          demosaic_quad(w, row_has_r, cpar, bb, gg, rr4);
          b = bb[ix & 3]; g = gg[ix & 3]; r = rr4[ix & 3];
        }
      } else {
        demosaic_at(raw, rows, cols, cols, iy, ix, cfa, b, g, r);
      }
      uint8_t* o = out + ((size_t)oy * ocols + ox) * 3;
      o[0] = b; o[1] = g; o[2] = r;
    }
}

// 16-bit Bayer extension: demosaic at 16 bits + reduction to BGR8 (frame_math.cuh demosaic_at16)
void hs_demosaic16(const uint16_t* raw, int rows, int cols, int cfa, uint8_t* out) {
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      int b, g, r;
      demosaic_at16(raw, rows, cols, (size_t)cols, y, x, cfa, b, g, r);
      uint8_t* o = out + ((size_t)y * cols + x) * 3;
      o[0] = (uint8_t)b; o[1] = (uint8_t)g; o[2] = (uint8_t)r;
    }
}
int hs_reduce16to8(int v) { return reduce16to8(v); }

void hs_remap(const uint8_t* src, int rows, int cols, int ch, const float* mx, const float* my, int orows, int ocols,
              uint8_t* out) {
  for (long i = 0; i < (long)orows * ocols; ++i) {
    int o[3];
    if (ch == 3) remap_pixel<3>(src, rows, cols, (size_t)cols * 3, mx[i], my[i], o);
    else remap_pixel<1>(src, rows, cols, (size_t)cols, mx[i], my[i], o);
    for (int c = 0; c < ch; ++c) out[i * ch + c] = o[c];
  }
}

// src: rows x cols B,G,R,0 pixels (uint32)
void hs_remap_bgrx(const uint32_t* src, int rows, int cols, const float* mx, const float* my, int orows, int ocols, uint8_t* out) {
  for (long i = 0; i < (long)orows * ocols; ++i) {
    const uint32_t p = remap_pixel_bgrx(src, rows, cols, cols, mx[i], my[i]);
    out[3 * i] = p & 255; out[3 * i + 1] = (p >> 8) & 255; out[3 * i + 2] = (p >> 16) & 255;
  }
}

// packed fixed-point map path; returns 0 when a displacement does not fit (caller falls back to the float map)
int hs_remap_bgrx_packed(const uint32_t* src, int rows, int cols, const float* mx, const float* my, int orows, int ocols, uint8_t* out) {
  std::vector<uint32_t> packed((size_t)orows * ocols);
  for (int y = 0; y < orows; ++y)
    for (int x = 0; x < ocols; ++x)
      if (!remap_pack_entry(mx[(size_t)y * ocols + x], my[(size_t)y * ocols + x], x, y, rows, cols, packed[(size_t)y * ocols + x])) return 0;
  for (int y = 0; y < orows; ++y)
    for (int x = 0; x < ocols; ++x) {
      const long i = (long)y * ocols + x;
      const uint32_t p = remap_pixel_bgrx_packed(src, rows, cols, cols, packed[i], x, y);
      out[3 * i] = p & 255; out[3 * i + 1] = (p >> 8) & 255; out[3 * i + 2] = (p >> 16) & 255;
    }
  return 1;
}

// The tile kernel's per-pixel arithmetic (rip_fast.cu k_remap_tile, test-free path) on a host copy of its shared-memory
// box: taps addressed as row base + dyi * box_w + dxi straight from the packed entry, blended by remap_blend_w.
// `box` holds box_h x box_w pixels whose upper-left corner is source pixel (bx0, by0); zeros outside the image.
uint32_t hs_remap_tile_pixel(const uint32_t* box, int box_w, int bx0, int by0, uint32_t packed, int x, int y) {
  const uint32_t* q = box + ((y - by0) * box_w + (x - bx0)) + (remap_packed_dyi(packed) * box_w + remap_packed_dxi(packed));
  return remap_blend_w(q[0], q[1], q[box_w], q[box_w + 1], packed & 31u, (packed >> 10) & 0x7c0u);
}
uint32_t hs_remap_pack(float mx, float my, int x, int y, int rows, int cols, int* ok) {
  uint32_t e = 0;
  *ok = remap_pack_entry(mx, my, x, y, rows, cols, e) ? 1 : 0;
  return e;
}
uint32_t hs_remap_pixel_packed(const uint32_t* src, int rows, int cols, uint32_t packed, int x, int y) {
  return remap_pixel_bgrx_packed(src, rows, cols, cols, packed, x, y);
}

void hs_pca_lut(const unsigned long long* stats, uint8_t* lut_b, uint8_t* lut_r, float* coeff) {
  PcaCoeff c = pca_coefficients(stats);
  coeff[0] = c.alpha_b; coeff[1] = c.beta_b; coeff[2] = c.alpha_r; coeff[3] = c.beta_r;
  for (int x = 0; x < 256; ++x) {
    lut_b[x] = pca_lut_entry(x, c.alpha_b, c.beta_b);
    lut_r[x] = pca_lut_entry(x, c.alpha_r, c.beta_r);
  }
}

void hs_gain_lut(float gain, uint8_t* lut) {
  for (int x = 0; x < 256; ++x) lut[x] = gain_lut_entry(x, gain);
}

// ---- CCC white balance pieces (ccc_math.cuh) ---------------------------------------------------
// cv::resize(img, 360x270, INTER_LINEAR) of a rows x cols BGR8 image
void hs_ccc_small(const uint8_t* img, int rows, int cols, uint8_t* out) {
  for (int dy = 0; dy < CCC_SMALL_H; ++dy) {
    const CccAxisCoef cy = ccc_axis_coef(rows, CCC_SMALL_H, dy);
    const int y0 = cy.s < 0 ? 0 : (cy.s > rows - 1 ? rows - 1 : cy.s);
    const int y1 = cy.s + 1 < 0 ? 0 : (cy.s + 1 > rows - 1 ? rows - 1 : cy.s + 1);
    for (int dx = 0; dx < CCC_SMALL_W; ++dx) {
      const CccAxisCoef cx = ccc_axis_coef_horizontal(cols, CCC_SMALL_W, dx);
      const int x0 = cx.s, x1 = cx.s + 1 > cols - 1 ? cols - 1 : cx.s + 1;
      for (int c = 0; c < 3; ++c) {
        const int S0 = ccc_hresize(img[((size_t)y0 * cols + x0) * 3 + c], img[((size_t)y0 * cols + x1) * 3 + c], cx.a0, cx.a1);
        const int S1 = ccc_hresize(img[((size_t)y1 * cols + x0) * 3 + c], img[((size_t)y1 * cols + x1) * 3 + c], cx.a0, cx.a1);
        out[((size_t)dy * CCC_SMALL_W + dx) * 3 + c] = (uint8_t)ccc_vresize(S0, S1, cy.a0, cy.a1);
      }
    }
  }
}

// histogram counts of the small image; returns the number of samples that landed in a bin
long hs_ccc_counts(const uint8_t* small_bgr, long n, float thr_hi, float thr_lo, float uv0, float bin_size, unsigned* counts) {
  float log_tab[256];
  memcpy(log_tab, kCvLogTabBits, sizeof log_tab);
  long used = 0;
  for (long i = 0; i < n; ++i) {
    int u, v;
    if (ccc_bin(small_bgr[3 * i], small_bgr[3 * i + 1], small_bgr[3 * i + 2], thr_hi, thr_lo, log_tab, uv0, bin_size, u, v)) {
      counts[u * CCC_BINS + v] += 1;
      ++used;
    }
  }
  return used;
}

void hs_ccc_gains(int uv_x, int uv_y, const float* exp_tab, float* gains_bgr) { ccc_gains(uv_x, uv_y, exp_tab, gains_bgr); }

}  // extern "C"
