// TEST INFRASTRUCTURE ONLY.  Compiles the kernels' per-pixel arithmetic (pixel_math.cuh,
// frame_math.cuh) as plain host C++ so the CPU-only test-suite can compare every formula
// exhaustively with cv2 before the CUDA build ever meets a GPU.  Never linked into, loaded
// by, or shipped with the product library (raw_image_pipeline_b200/librip_b200.so).
#include <cstring>
#include "../../raw_image_pipeline_b200/csrc/ccc_math.cuh"
#include "../../raw_image_pipeline_b200/csrc/cv_tables.inc"

using namespace rip;

static ChainTables make_tables(const uint8_t* wb, const uint8_t* gamma) {
  ChainTables t;
  t.wb = wb; t.gamma = gamma;
  t.srgb_g = kSrgbGammaTab; t.lab_c = kLabCbrtTab; t.lab_yf = kLabToYF; t.inv_g = kSrgbInvGammaTab;
  t.sdiv = kHsvSdiv; t.hdiv = kHsvHdiv; t.enh = nullptr;
  return t;
}

extern "C" {

void hs_bgr2lab(long n, const uint8_t* in, uint8_t* out) {
  ChainTables t = make_tables(nullptr, nullptr);
  for (long i = 0; i < n; ++i) {
    int L, A, B; bgr_to_lab(in[3 * i], in[3 * i + 1], in[3 * i + 2], t, L, A, B);
    out[3 * i] = L; out[3 * i + 1] = A; out[3 * i + 2] = B;
  }
}
void hs_lab2bgr(long n, const uint8_t* in, uint8_t* out) {
  ChainTables t = make_tables(nullptr, nullptr);
  for (long i = 0; i < n; ++i) {
    int b, g, r; lab_to_bgr(in[3 * i], in[3 * i + 1], in[3 * i + 2], t, b, g, r);
    out[3 * i] = b; out[3 * i + 1] = g; out[3 * i + 2] = r;
  }
}
void hs_bgr2hsv(long n, const uint8_t* in, uint8_t* out) {
  ChainTables t = make_tables(nullptr, nullptr);
  for (long i = 0; i < n; ++i) {
    int h, s, v; bgr_to_hsv(in[3 * i], in[3 * i + 1], in[3 * i + 2], t, h, s, v);
    out[3 * i] = h; out[3 * i + 1] = s; out[3 * i + 2] = v;
  }
}
// `width`: row length of the image the n pixels form (selects OpenCV's scalar row tail)
void hs_hsv2bgr(long n, int width, const uint8_t* in, uint8_t* out) {
  for (long i = 0; i < n; ++i) {
    const bool tail = (int)(i % width) >= (width & ~31);
    int b, g, r; hsv_to_bgr(in[3 * i], in[3 * i + 1], in[3 * i + 2], tail, b, g, r);
    out[3 * i] = b; out[3 * i + 1] = g; out[3 * i + 2] = r;
  }
}

// full per-pixel chain; mask may be null when ST_VIG is off
void hs_chain(unsigned stages, long n, int width, const uint8_t* in, const float* mask, const float* cc, const float* bias,
              const double* enh, const uint8_t* wb, const uint8_t* gamma, uint8_t* out) {
  ChainTables t = make_tables(wb, gamma);
  ChainConsts k;
  memcpy(k.cc, cc, sizeof k.cc); memcpy(k.cc_bias, bias, sizeof k.cc_bias);
  uint8_t enh_lut[768];
  for (int c = 0; c < 3; ++c)
    for (int x = 0; x < 256; ++x) enh_lut[256 * c + x] = enh_gain_lut_entry(x, enh[c]);
  t.enh = enh_lut;
  for (long i = 0; i < n; ++i) {
    int b = in[3 * i], g = in[3 * i + 1], r = in[3 * i + 2];
    const float m = mask ? mask[i] : 1.0f;
    if (stages & ST_WB) { b = t.wb[b]; g = t.wb[256 + g]; r = t.wb[512 + r]; }
    if (stages & ST_CC) color_calibrate(b, g, r, k);
    if (stages & ST_GAMMA) { b = t.gamma[b]; g = t.gamma[g]; r = t.gamma[r]; }
    if (stages & ST_VIG) vignetting(b, g, r, m, t);
    if (stages & ST_ENH) enhance(b, g, r, (int)(i % width) >= (width & ~31), t);
    out[3 * i] = b; out[3 * i + 1] = g; out[3 * i + 2] = r;
  }
}

// mode 0: demosaic_at everywhere; mode 1: demosaic_quad where legal (the kernels' fast path)
void hs_demosaic(const uint8_t* raw, int rows, int cols, int cfa, int angle, int mode, uint8_t* out) {
  const int orows = (angle == 90 || angle == 270) ? cols : rows;
  const int ocols = (angle == 90 || angle == 270) ? rows : cols;
  for (int oy = 0; oy < orows; ++oy)
    for (int ox = 0; ox < ocols; ++ox) {
      int iy, ix; flip_source(angle, rows, cols, oy, ox, iy, ix);
      int b, g, r;
      const int x4 = ix & ~3;
      if (mode == 1 && x4 >= 4 && x4 + 7 < cols && x4 + 3 <= cols - 2) {
        int yc = iy < 1 ? 1 : (iy > rows - 2 ? rows - 2 : iy);
        uint32_t w[3][3];
        for (int rr = 0; rr < 3; ++rr)
          for (int j = 0; j < 3; ++j) memcpy(&w[rr][j], raw + (size_t)(yc - 1 + rr) * cols + x4 - 4 + 4 * j, 4);
        const bool row_has_r = ((yc & 1) == ((cfa >> 1) & 1));
        const int cpar = row_has_r ? (cfa & 1) : ((cfa & 1) ^ 1);
        int bb[4], gg[4], rr4[4];
        demosaic_quad(w, row_has_r, cpar, bb, gg, rr4);
        b = bb[ix & 3]; g = gg[ix & 3]; r = rr4[ix & 3];
      } else {
        demosaic_at(raw, rows, cols, cols, iy, ix, cfa, b, g, r);
      }
      uint8_t* o = out + ((size_t)oy * ocols + ox) * 3;
      o[0] = b; o[1] = g; o[2] = r;
    }
}

void hs_remap(const uint8_t* src, int rows, int cols, int ch, const float* mx, const float* my, int orows, int ocols,
              uint8_t* out) {
  for (long i = 0; i < (long)orows * ocols; ++i) {
    int o[3];
    if (ch == 3) remap_pixel<3>(src, rows, cols, (size_t)cols * 3, mx[i], my[i], o);
    else remap_pixel<1>(src, rows, cols, (size_t)cols, mx[i], my[i], o);
    for (int c = 0; c < ch; ++c) out[i * ch + c] = o[c];
  }
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
