// Convolutional colour constancy white balance on the device
// (raw_image_pipeline_white_balance/src/.../convolutional_color_constancy.cpp:91-386).
//
//   k_ccc_hist      360x270 bilinear resample of the debayered+flipped frame (only the 4 source
//                   pixels each sample touches are demosaiced) -> log-chroma bin -> integer counts
//   k_ccc_weights   counts -> the fp32 value `count` sequential additions of 1/97200 produce
//   k_ccc_fft       256-point radix-2 FFTs of eight rows per CTA in shared memory (fp64), output transposed;
//                   four passes give  IDFT2( DFT2(hist) * DFT2(filter) )   (ccc.cpp:273-298)
//   k_ccc_argmax    + bias, first maximum in row-major order (cv::minMaxLoc)
//   k_ccc_gains     [Kalman tracker ->] gains (ccc.cpp:300-381)
//
// The reference evaluates the circular convolution with fp32 FFTs; fp64 here is simply the
// accurate value of the same sum (the arg-max margin on real frames is ~1e-2 relative, fp32 FFT
// noise ~2e-7).
#include "ccc.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/rip_b200.h"
#include "ccc_math.cuh"
#include "cv_tables.inc"

namespace rip {

namespace {

constexpr int N_SMALL = CCC_SMALL_W * CCC_SMALL_H;  // 97200
constexpr int N_BINS2 = CCC_BINS * CCC_BINS;        // 65536
constexpr int CHUNK = 64;                           // frames per FFT batch (2 x 64 MB of spectra)

struct KfState {
  float x[2];
  int steps;  // correct() calls so far (indexes the gain table)
  int first;  // first_frame_
};

// (b, g, r) of pixel (oy, ox) of the debayered + flipped frame
template <int SRC>
__device__ __forceinline__ void pixel_post_flip(const FrameParams& P, const uint8_t* fin, int oy, int ox, int& b, int& g, int& r) {
  int iy, ix;
  flip_source(P.angle, P.rows, P.cols, oy, ox, iy, ix);
  if (SRC == SRC_BAYER) {
    demosaic_at(fin, P.rows, P.cols, (size_t)P.in_pitch, iy, ix, P.cfa, b, g, r);
  } else {
    const uint8_t* p = fin + (size_t)iy * P.in_pitch + 3 * ix;
    g = p[1];
    if (SRC == SRC_RGB) { r = p[0]; b = p[2]; } else { b = p[0]; r = p[2]; }
  }
}

template <int SRC>
__global__ void __launch_bounds__(256) k_ccc_hist(const __grid_constant__ FrameParams P, int frame0, const CccAxisCoef* __restrict__ xc,
                                                  const CccAxisCoef* __restrict__ yc, const float* __restrict__ log_tab, float thr_hi,
                                                  float thr_lo, float uv0, float bin_size, unsigned* __restrict__ counts) {
  __shared__ float s_log[256];
  s_log[threadIdx.x] = log_tab[threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= N_SMALL) return;
  const int dy = i / CCC_SMALL_W, dx = i - dy * CCC_SMALL_W;
  const uint8_t* fin = P.in + (long long)(frame0 + blockIdx.y) * P.in_frame_stride;
  const CccAxisCoef cy = yc[dy], cx = xc[dx];
  const int y0 = min(max(cy.s, 0), P.orows - 1), y1 = min(max(cy.s + 1, 0), P.orows - 1);
  const int x0 = cx.s, x1 = min(cx.s + 1, P.ocols - 1);
  int p00[3], p01[3], p10[3], p11[3];
  pixel_post_flip<SRC>(P, fin, y0, x0, p00[0], p00[1], p00[2]);
  pixel_post_flip<SRC>(P, fin, y0, x1, p01[0], p01[1], p01[2]);
  pixel_post_flip<SRC>(P, fin, y1, x0, p10[0], p10[1], p10[2]);
  pixel_post_flip<SRC>(P, fin, y1, x1, p11[0], p11[1], p11[2]);
  int s[3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
    s[c] = ccc_vresize(ccc_hresize(p00[c], p01[c], cx.a0, cx.a1), ccc_hresize(p10[c], p11[c], cx.a0, cx.a1), cy.a0, cy.a1);
  int u, v;
  if (ccc_bin(s[0], s[1], s[2], thr_hi, thr_lo, s_log, uv0, bin_size, u, v))
    atomicAdd(counts + (size_t)blockIdx.y * N_BINS2 + u * CCC_BINS + v, 1u);
}

__global__ void __launch_bounds__(256) k_ccc_weights(const unsigned* __restrict__ counts, const float* __restrict__ weight, float* __restrict__ hist,
                                                    long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) hist[i] = weight[min(counts[i], (unsigned)N_SMALL)];
}

// Eight 256-point FFTs (eight consecutive rows of one frame) per CTA of 256 threads, radix 2 in shared memory (fp64),
// output transposed: with eight rows in flight every transposed output row receives 128 contiguous bytes per CTA (one row
// per CTA wrote 16-byte pieces 4 KB apart), and a stage's twiddle factor is loaded once per thread for its four butterflies.
//   MODE 0: real fp32 input, forward        MODE 1: complex input, forward
//   MODE 2: complex input multiplied by `mul` (the filter spectrum), inverse (unnormalised)
//   MODE 3: complex input, inverse (unnormalised)
constexpr int FFT_ROWS = 8;     // rows per CTA
constexpr int FFT_PITCH = 257;  // double2 per shared row: the odd pitch makes the transposed read-out conflict-free
// MODE 0 packs TWO frames into one complex transform: the histogram of frame 2j becomes the real part and that of frame
// 2j + 1 (when it exists: `n_real` frames in all) the imaginary part of batch entry j.  The filter is real, so
// IDFT2(DFT2(a + i b) * DFT2(f)) = conv(a, f) + i conv(b, f): after the four passes the real part of entry j is the
// response of frame 2j and the imaginary part that of frame 2j + 1 -- half the transforms for the same sums.
template <int MODE>
__global__ void __launch_bounds__(256) k_ccc_fft(const void* __restrict__ in, double2* __restrict__ out, const double2* __restrict__ tw,
                                                 const double2* __restrict__ mul, int n_real) {
  __shared__ double2 s[FFT_ROWS * FFT_PITCH];
  __shared__ double2 s_tw[128];
  const int row0 = blockIdx.x * FFT_ROWS, t = threadIdx.x;
  const size_t base = (size_t)blockIdx.y * N_BINS2;
  if (t < 128) {
    double2 w = tw[t];
    if (MODE >= 2) w.y = -w.y;
    s_tw[t] = w;
  }
#pragma unroll
  for (int r = 0; r < FFT_ROWS; ++r) {  // row r, element t: coalesced
    const size_t idx = (size_t)(row0 + r) * 256 + t;
    double2 v;
    if (MODE == 0) {
      const float* h = static_cast<const float*>(in) + 2 * base + idx;  // frames 2j and 2j + 1
      v.x = (double)h[0];
      v.y = 2 * (int)blockIdx.y + 1 < n_real ? (double)h[N_BINS2] : 0.0;
    } else {
      v = static_cast<const double2*>(in)[base + idx];
      if (MODE == 2) {
        const double2 m = mul[idx];
        v = make_double2(v.x * m.x - v.y * m.y, v.x * m.y + v.y * m.x);
      }
    }
    s[r * FFT_PITCH + (__brev((unsigned)t) >> 24)] = v;
  }
  __syncthreads();
  // butterfly b of rows (t >> 7) + {0, 2, 4, 6}
  const int b = t & 127;
  double2* const sr = s + (t >> 7) * FFT_PITCH;
#pragma unroll
  for (int half = 1; half < 256; half <<= 1) {
    const int pos = b & (half - 1);
    const int i = ((b - pos) << 1) + pos, j = i + half;
    const double2 w = s_tw[pos * (128 / half)];
#pragma unroll
    for (int k = 0; k < FFT_ROWS / 2; ++k) {
      double2* const q = sr + 2 * k * FFT_PITCH;
      const double2 a = q[i], c = q[j];
      const double2 cw = make_double2(c.x * w.x - c.y * w.y, c.x * w.y + c.y * w.x);
      q[i] = make_double2(a.x + cw.x, a.y + cw.y);
      q[j] = make_double2(a.x - cw.x, a.y - cw.y);
    }
    __syncthreads();
  }
  // transposed: output row e receives this CTA's eight values at columns row0 .. row0 + 7
  const int r = t & (FFT_ROWS - 1);
#pragma unroll
  for (int k = 0; k < 256 / (256 / FFT_ROWS); ++k) {
    const int e = (t >> 3) + (256 / FFT_ROWS) * k;
    out[base + (size_t)e * 256 + row0 + r] = s[r * FFT_PITCH + e];
  }
}

// first maximum (row-major) of  response / 65536 + bias ; one CTA per frame
__global__ void __launch_bounds__(1024) k_ccc_argmax(const double2* __restrict__ resp, const float* __restrict__ bias, int2* __restrict__ uv) {
  __shared__ double s_val[1024];
  __shared__ int s_idx[1024];
  // frame f: entry f / 2 of the batch, real (even f) or imaginary (odd f) part -- see k_ccc_fft
  const double* r = reinterpret_cast<const double*>(resp + (size_t)(blockIdx.x >> 1) * N_BINS2) + (blockIdx.x & 1);
  double best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int i = threadIdx.x; i < N_BINS2; i += 1024) {
    const double v = r[2 * i] * (1.0 / 65536.0) + (double)bias[i];
    if (v > best) { best = v; best_i = i; }  // increasing i per thread: strict > keeps the first
  }
  s_val[threadIdx.x] = best; s_idx[threadIdx.x] = best_i;
  __syncthreads();
  for (int step = 512; step > 0; step >>= 1) {
    if (threadIdx.x < step) {
      const double v = s_val[threadIdx.x + step];
      const int i = s_idx[threadIdx.x + step];
      if (v > s_val[threadIdx.x] || (v == s_val[threadIdx.x] && i < s_idx[threadIdx.x])) { s_val[threadIdx.x] = v; s_idx[threadIdx.x] = i; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int i = s_idx[0] == 0x7fffffff ? 0 : s_idx[0];
    uv[blockIdx.x] = make_int2(i & 255, i >> 8);  // cv::Point(x = column, y = row)
  }
}

// kalmanFiltering (ccc.cpp:300-340) + computeGains (:342-381).  With temporal consistency the n
// frames are consecutive frames of one stream: a single thread runs the recurrence.
__global__ void k_ccc_gains(const int2* __restrict__ uv_in, int n, int temporal, int reset_first, KfState* kf,
                            const float* __restrict__ kgain, const float* __restrict__ exp_tab, float* __restrict__ gains,
                            int2* __restrict__ uv_out) {
  if (temporal) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    KfState st = *kf;
    if (reset_first) st.first = 1;
    for (int f = 0; f < n; ++f) {
      int2 uv = uv_in[f];
      if (st.first) {
        st.first = 0;
        st.x[0] = (float)uv.x; st.x[1] = (float)uv.y;
      } else {
        // cv::KalmanFilter predict()+correct() with A = H = I: x += K * (z - x), K tabulated (cv_tables.inc)
        const float K = kgain[min(st.steps, 39)];
        st.steps += 1;
        st.x[0] = __fadd_rn(st.x[0], __fmul_rn(K, __fsub_rn((float)uv.x, st.x[0])));
        st.x[1] = __fadd_rn(st.x[1], __fmul_rn(K, __fsub_rn((float)uv.y, st.x[1])));
        uv = make_int2((int)st.x[0], (int)st.x[1]);  // float -> int truncation (ccc.cpp:336-337)
      }
      uv.x = min(max(uv.x, 0), 255); uv.y = min(max(uv.y, 0), 255);
      uv_out[f] = uv;
      ccc_gains(uv.x, uv.y, exp_tab, gains + 3 * f);
    }
    *kf = st;
  } else {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const int2 uv = uv_in[f];
    uv_out[f] = uv;
    ccc_gains(uv.x, uv.y, exp_tab, gains + 3 * f);
  }
}

#define CCC_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(_e); return RIP_ERR_CUDA; } \
  } while (0)

int ccc_device_init(CccState& c, cudaStream_t stream, int* launches, std::string& err) {
  if (c.device_ready) return RIP_OK;
  if (c.w != CCC_BINS || c.h != CCC_BINS) { err = "CCC model must be 256 x 256"; return RIP_ERR_UNSUPPORTED; }
  // tables: cv::log | 1/expf(-(k*bin + uv0)) | Kalman gains
  std::vector<float> tabs(256 + 256 + 40);
  memcpy(tabs.data(), kCvLogTabBits, 256 * sizeof(float));
  const float bin_size = 1.0f / 64.0f, uv0 = -1.421875f;
  for (int k = 0; k < 256; ++k) {
    const float L = (float)k * bin_size + uv0;   // two roundings (ccc.cpp:359-360; -ffp-contract=off)
    tabs[256 + k] = 1.0f / std::exp(-L);         // std::exp(float) == expf, this host's libm like the reference
  }
  memcpy(tabs.data() + 512, kCccKalmanGainBits, 40 * sizeof(float));
  std::vector<float> weight(N_SMALL + 1);
  {
    const float num_pixels = (float)N_SMALL;
    const float w = 1.0f / num_pixels;
    weight[0] = 0.f;
    for (int k = 1; k <= N_SMALL; ++k) weight[k] = weight[k - 1] + w;  // hist(u,v) += pixel_weight
  }
  std::vector<double> tw(256);
  for (int k = 0; k < 128; ++k) {
    const double a = -2.0 * 3.14159265358979323846 * k / 256.0;
    tw[2 * k] = std::cos(a); tw[2 * k + 1] = std::sin(a);
  }
  KfState kf{{128.f, 128.f}, 0, 1};
  CCC_CUDA(c.d_tabs.reserve(tabs.size() * sizeof(float)));
  CCC_CUDA(c.d_weight.reserve(weight.size() * sizeof(float)));
  CCC_CUDA(c.d_twiddle.reserve(tw.size() * sizeof(double)));
  CCC_CUDA(c.d_bias.reserve(N_BINS2 * sizeof(float)));
  CCC_CUDA(c.d_kf.reserve(sizeof(KfState)));
  CCC_CUDA(c.d_filter_fft.reserve(N_BINS2 * sizeof(double2)));
  DevBuf tmp_f, tmp_c;
  CCC_CUDA(tmp_f.reserve(N_BINS2 * sizeof(float)));
  CCC_CUDA(tmp_c.reserve(N_BINS2 * sizeof(double2)));
  CCC_CUDA(cudaMemcpyAsync(c.d_tabs.ptr, tabs.data(), tabs.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
  CCC_CUDA(cudaMemcpyAsync(c.d_weight.ptr, weight.data(), weight.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
  CCC_CUDA(cudaMemcpyAsync(c.d_twiddle.ptr, tw.data(), tw.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
  CCC_CUDA(cudaMemcpyAsync(c.d_bias.ptr, c.bias.data(), N_BINS2 * sizeof(float), cudaMemcpyHostToDevice, stream));
  CCC_CUDA(cudaMemcpyAsync(c.d_kf.ptr, &kf, sizeof kf, cudaMemcpyHostToDevice, stream));
  CCC_CUDA(cudaMemcpyAsync(tmp_f.ptr, c.filter.data(), N_BINS2 * sizeof(float), cudaMemcpyHostToDevice, stream));
  k_ccc_fft<0><<<dim3(256 / FFT_ROWS, 1), 256, 0, stream>>>(tmp_f.ptr, tmp_c.as<double2>(), c.d_twiddle.as<double2>(), nullptr, 1);
  k_ccc_fft<1><<<dim3(256 / FFT_ROWS, 1), 256, 0, stream>>>(tmp_c.ptr, c.d_filter_fft.as<double2>(), c.d_twiddle.as<double2>(), nullptr, 1);
  CCC_CUDA(cudaGetLastError());
  if (launches) *launches += 2;
  CCC_CUDA(cudaStreamSynchronize(stream));  // host vectors and temporaries go out of scope
  tmp_f.release(); tmp_c.release();
  c.device_ready = true;
  return RIP_OK;
}

}  // namespace

// File layout (tools/convert_ccc_model.py): "RIPCCC1\0", int32 width, int32 height, then the
// filter and the bias as height x width fp32, already transposed like ccc.cpp:131-132 does.
bool ccc_load_model(CccState& c, const std::string& path, std::string& err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { err = "CCC model file " + path + " not found"; c.model_error = err; return false; }
  char magic[8];
  int wh[2] = {0, 0};
  bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "RIPCCC1\0", 8) == 0 && fread(wh, sizeof(int), 2, f) == 2 &&
            wh[0] > 0 && wh[1] > 0 && wh[0] <= 4096 && wh[1] <= 4096;
  if (ok) {
    c.w = wh[0]; c.h = wh[1];
    const size_t n = (size_t)c.w * c.h;
    c.filter.resize(n); c.bias.resize(n);
    ok = fread(c.filter.data(), sizeof(float), n, f) == n && fread(c.bias.data(), sizeof(float), n, f) == n;
  }
  fclose(f);
  if (!ok) { err = "CCC model file " + path + " is malformed"; c.model_error = err; return false; }
  c.model_loaded = true;
  c.uv_x = c.h / 2; c.uv_y = c.w / 2;  // ccc.cpp:172
  return true;
}

void ccc_release(CccState& c) {
  c.d_filter_fft.release(); c.d_bias.release(); c.d_twiddle.release(); c.d_weight.release(); c.d_tabs.release();
  c.d_kf.release(); c.d_coef.release();
  c.device_ready = false; c.coef_rows = c.coef_cols = -1;
  c.d_last_uv = c.d_last_response = nullptr; c.last_n = 0;
}

int ccc_white_balance(CccState& c, const Params& q, const FrameParams& fp, DevBuf& work, DevBuf& gains, int sm_count,
                      cudaStream_t stream, int* launches, std::string& err) {
  (void)sm_count;
  if (!c.model_loaded) { err = c.model_error.empty() ? "CCC model not loaded" : c.model_error; return RIP_ERR_IO; }
  int rc = ccc_device_init(c, stream, launches, err);
  if (rc != RIP_OK) return rc;
  const int n = fp.n_frames;
  // resize coefficients for this (post-flip) frame size
  if (c.coef_rows != fp.orows || c.coef_cols != fp.ocols) {
    std::vector<CccAxisCoef> coef(CCC_SMALL_W + CCC_SMALL_H);
    for (int d = 0; d < CCC_SMALL_W; ++d) coef[d] = ccc_axis_coef_horizontal(fp.ocols, CCC_SMALL_W, d);
    for (int d = 0; d < CCC_SMALL_H; ++d) coef[CCC_SMALL_W + d] = ccc_axis_coef(fp.orows, CCC_SMALL_H, d);
    CCC_CUDA(cudaStreamSynchronize(stream));
    CCC_CUDA(c.d_coef.reserve(coef.size() * sizeof(CccAxisCoef)));
    CCC_CUDA(cudaMemcpyAsync(c.d_coef.ptr, coef.data(), coef.size() * sizeof(CccAxisCoef), cudaMemcpyHostToDevice, stream));
    CCC_CUDA(cudaStreamSynchronize(stream));
    c.coef_rows = fp.orows; c.coef_cols = fp.ocols;
  }
  const int chunk = n < CHUNK ? n : CHUNK;
  // work: counts | hist (fp32) | spectrum A | spectrum B | arg-max uv (n) | final uv (n)
  const size_t off_hist = (size_t)chunk * N_BINS2 * sizeof(unsigned);
  const size_t off_a = off_hist + (size_t)chunk * N_BINS2 * sizeof(float);
  const size_t off_b = off_a + (size_t)chunk * N_BINS2 * sizeof(double2);
  const size_t off_uv = off_b + (size_t)chunk * N_BINS2 * sizeof(double2);
  const size_t total = off_uv + 2 * (size_t)n * sizeof(int2);
  CCC_CUDA(work.reserve(total));
  CCC_CUDA(gains.reserve((size_t)n * 3 * sizeof(float)));
  uint8_t* wbase = work.as<uint8_t>();
  unsigned* counts = reinterpret_cast<unsigned*>(wbase);
  float* hist = reinterpret_cast<float*>(wbase + off_hist);
  double2* sa = reinterpret_cast<double2*>(wbase + off_a);
  double2* sb = reinterpret_cast<double2*>(wbase + off_b);
  int2* uv_arg = reinterpret_cast<int2*>(wbase + off_uv);
  int2* uv_fin = uv_arg + n;
  const float* log_tab = c.d_tabs.as<float>();
  const float* exp_tab = log_tab + 256;
  const float* kgain = log_tab + 512;
  const CccAxisCoef* xc = c.d_coef.as<CccAxisCoef>();
  const CccAxisCoef* yc = xc + CCC_SMALL_W;
  const double2* tw = c.d_twiddle.as<double2>();
  // WhiteBalanceModule forwards its thresholds on every frame (white_balance.hpp:71); `255 * thr` is
  // int * float in the reference (ccc.cpp:215-218)
  const float thr_hi = 255 * (float)q.wb_bright_thr, thr_lo = 255 * (float)q.wb_dark_thr;
  const float uv0 = -1.421875f, bin_size = 1.0f / 64.0f;

  for (int f0 = 0; f0 < n; f0 += chunk) {
    const int m = (n - f0) < chunk ? (n - f0) : chunk;
    CCC_CUDA(cudaMemsetAsync(counts, 0, (size_t)m * N_BINS2 * sizeof(unsigned), stream));
    const dim3 grid_h((N_SMALL + 255) / 256, m);
    if (fp.src == SRC_BAYER) k_ccc_hist<SRC_BAYER><<<grid_h, 256, 0, stream>>>(fp, f0, xc, yc, log_tab, thr_hi, thr_lo, uv0, bin_size, counts);
    else if (fp.src == SRC_BGR) k_ccc_hist<SRC_BGR><<<grid_h, 256, 0, stream>>>(fp, f0, xc, yc, log_tab, thr_hi, thr_lo, uv0, bin_size, counts);
    else k_ccc_hist<SRC_RGB><<<grid_h, 256, 0, stream>>>(fp, f0, xc, yc, log_tab, thr_hi, thr_lo, uv0, bin_size, counts);
    const long long nb = (long long)m * N_BINS2;
    k_ccc_weights<<<(unsigned)((nb + 255) / 256), 256, 0, stream>>>(counts, c.d_weight.as<float>(), hist, nb);
    const dim3 grid_f(256 / FFT_ROWS, (m + 1) / 2);  // two frames per complex transform
    k_ccc_fft<0><<<grid_f, 256, 0, stream>>>(hist, sa, tw, nullptr, m);
    k_ccc_fft<1><<<grid_f, 256, 0, stream>>>(sa, sb, tw, nullptr, m);
    k_ccc_fft<2><<<grid_f, 256, 0, stream>>>(sb, sa, tw, c.d_filter_fft.as<double2>(), m);
    k_ccc_fft<3><<<grid_f, 256, 0, stream>>>(sa, sb, tw, nullptr, m);
    k_ccc_argmax<<<m, 1024, 0, stream>>>(sb, c.d_bias.as<float>(), uv_arg + f0);
    CCC_CUDA(cudaGetLastError());
    if (launches) *launches += 7;
    c.d_last_response = sb + (size_t)((m - 1) >> 1) * N_BINS2;
    c.last_response_part = (m - 1) & 1;
  }
  const int temporal = q.wb_temporal_consistency ? 1 : 0;
  k_ccc_gains<<<temporal ? 1 : (n + 127) / 128, temporal ? 1 : 128, 0, stream>>>(uv_arg, n, temporal, c.pending_reset ? 1 : 0,
                                                                                 static_cast<KfState*>(c.d_kf.ptr), kgain, exp_tab,
                                                                                 gains.as<float>(), uv_fin);
  CCC_CUDA(cudaGetLastError());
  if (launches) *launches += 1;
  if (temporal) c.pending_reset = false;
  c.d_last_uv = uv_fin;
  c.last_n = n;
  return RIP_OK;
}

int ccc_fetch_last(CccState& c, const DevBuf& gains, cudaStream_t stream, std::string& err) {
  if (!c.d_last_uv || c.last_n <= 0) return RIP_OK;
  int uv[2];
  float g[3];
  CCC_CUDA(cudaMemcpyAsync(uv, static_cast<const int2*>(c.d_last_uv) + (c.last_n - 1), sizeof uv, cudaMemcpyDeviceToHost, stream));
  CCC_CUDA(cudaMemcpyAsync(g, gains.as<float>() + 3 * (size_t)(c.last_n - 1), sizeof g, cudaMemcpyDeviceToHost, stream));
  CCC_CUDA(cudaStreamSynchronize(stream));
  c.uv_x = uv[0]; c.uv_y = uv[1];
  c.gain_b = g[0]; c.gain_g = g[1]; c.gain_r = g[2];
  return RIP_OK;
}

}  // namespace rip
