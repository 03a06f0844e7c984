// C ABI (include/rip_b200.h) + the host pipeline object: owns the parameters (HostState), the
// device-resident tables, and orchestrates  stats -> LUT -> fused -> remap  per batch of frames.
// There is no CPU pixel path in this library: without a CUDA device every pixel call fails.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only; ranges cost nothing unless a profiler is attached

#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <atomic>
#include <thread>
#include <vector>

#include "../../include/rip_b200.h"
#include "ccc.hpp"
#include "cv_tables.inc"
#include "devbuf.hpp"
#include "frame_math.cuh"
#include "host_state.hpp"
#include "kernels.hpp"

#ifndef RIP_DEFAULT_CONFIG_DIR
#define RIP_DEFAULT_CONFIG_DIR ""
#endif

using namespace rip;

namespace {

thread_local std::string g_create_error;

// per-stream scratch for one in-flight batch
struct Scratch {
  DevBuf color;   // pre-undistortion images when the caller does not provide a buffer
  DevBuf bgr8;    // 16-bit Bayer extension: the demosaiced frames reduced to BGR8
  DevBuf wb;      // n x 3 x 256 float
  DevBuf stats;   // n x 8 u64
  DevBuf coeff;   // n x 4 float (pca coefficients, kept for inspection)
  DevBuf gains;   // n x 3 float (ccc)
  DevBuf ccc;     // ccc working memory
  void release() { bgr8.release(); color.release(); wb.release(); stats.release(); coeff.release(); gains.release(); ccc.release(); }
};

struct Slot {  // host<->device streaming slot for rip_apply_batch_host
  cudaStream_t stream = nullptr;
  DevBuf in, out;
  Scratch scratch;
};

struct FrameGeom {
  int rows, cols, channels;  // input
  int bytes_per_sample = 1;  // 2 for the 16-bit Bayer extension
  int src, cfa;              // SRC_*, CFA_*
  int angle;                 // effective flip angle
  int frows, fcols;          // after flip
  bool color;                // 3 channels after debayer
  bool undistort;
  int orows, ocols, ochannels;  // final
  std::string out_encoding;
};

// memcpy that leaves the lines of the PINNED side (`pinned`: dst or src) in the last-level cache only: every 64-byte line is
// demoted (CLDEMOTE, a hint: a no-op where unsupported) after use.  The DMA engine of the GPU reads / overwrites that
// buffer next, and snooping lines out of eight cores' private L2s slowed the device-side copies 4-5x (measured: 1014 vs
// 213 us for H2D + kernels + D2H of a 1080p frame, profiles/r2_apply_latency.md).
inline void demote_range(const uint8_t* p, size_t n) {
#if defined(__x86_64__)
  for (const uint8_t* q = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)63); q < p + n; q += 64)
    asm volatile(".byte 0x0f, 0x1c, 0x07" ::"D"(q) : "memory");  // cldemote (%rdi)
#else
  (void)p; (void)n;
#endif
}
inline void copy_and_demote(uint8_t* dst, const uint8_t* src, size_t n, int pinned_side /* 0: dst, 1: src */) {
  const size_t kBlock = 16 << 10;
  for (size_t o = 0; o < n; o += kBlock) {
    const size_t m = std::min(kBlock, n - o);
    memcpy(dst + o, src + o, m);
    demote_range(pinned_side ? src + o : dst + o, m);
  }
}

// A few persistent host threads that copy between the caller's (pageable) images and the pipeline's pinned staging
// buffers: one thread moves ~10 GB/s, a 1080p BGR8 frame is 6.2 MB, and the copies would otherwise dominate rip_apply.
class CopyPool {
 public:
  static CopyPool& get() { static CopyPool pool; return pool; }
  // dst[0..n) = src[0..n), split over the pool and the calling thread; returns when done
  // `pinned_side`: which of the two buffers the GPU's DMA engine touches next (0: dst, 1: src), see copy_and_demote
  void copy(uint8_t* dst, const uint8_t* src, size_t n, int pinned_side) {
    const size_t kMin = (size_t)512 << 10;  // not worth a hand-over below this
    const int parts = (int)std::min<size_t>(workers_.size() + 1, (n + kMin - 1) / kMin);
    if (parts <= 1) { copy_and_demote(dst, src, n, pinned_side); return; }
    const size_t chunk = ((n + parts - 1) / parts + 63) & ~(size_t)63;
    {
      std::lock_guard<std::mutex> lk(m_);
      for (int i = 1; i < parts; ++i) {
        const size_t off = chunk * i;
        if (off >= n) break;
        jobs_.push_back({dst + off, src + off, std::min(chunk, n - off), pinned_side});
        ++pending_;
      }
    }
    cv_.notify_all();
    copy_and_demote(dst, src, std::min(chunk, n), pinned_side);
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [&] { return pending_ == 0; });
  }

 private:
  struct Job { uint8_t* d; const uint8_t* s; size_t n; int pinned_side; };
  CopyPool() {
    unsigned hw = std::thread::hardware_concurrency();
    int n = hw >= 16 ? 7 : (hw >= 8 ? 3 : (hw >= 4 ? 1 : 0));
    if (const char* v = getenv("RIP_B200_COPY_THREADS")) n = std::max(0, atoi(v) - 1);
    for (int i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
  }
  ~CopyPool() {
    { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
    cv_.notify_all();
    for (std::thread& t : workers_) t.join();
  }
  void run() {
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || !jobs_.empty(); });
        if (stop_ && jobs_.empty()) return;
        j = jobs_.front(); jobs_.pop_front();
      }
      copy_and_demote(j.d, j.s, j.n, j.pinned_side);
      { std::lock_guard<std::mutex> lk(m_); if (--pending_ == 0) done_.notify_all(); }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  std::deque<Job> jobs_;
  int pending_ = 0;
  bool stop_ = false;
};

}  // namespace

struct rip_pipeline {
  HostState hs;
  std::string last_error;
  int device = -1;
  bool cuda_ready = false;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  long long kernel_launches = 0;

  // optional per-kernel timing with CUDA events on the launching stream ("profile/kernel_events")
  enum { SPAN_STATS = 0, SPAN_LUT = 1, SPAN_FUSED = 2, SPAN_REMAP = 3, SPAN_KINDS = 4 };
  struct Span { int kind; cudaEvent_t a, b; };
  bool profile = false;
  bool force_generic = false;  // "debug/force_generic_kernels": tests run both kernel families
  bool force_float_map = false;  // "debug/force_float_map": undistortion reads the fp32 map even where the packed one exists
  bool emit_rect_mask = false;  // "undistortion/rect_mask": getRectMask() returns a real validity mask (the reference's is always empty)
  int fused_kernel = 0;  // "debug/fused_kernel": 0 = measured choice per stage set, 1 = tile kernel (rip_fast.cu), >= 2 = strip kernel (rip_strip.cuh)
  bool force_gather_remap = false;  // "debug/force_gather_remap": undistortion gathers from global memory even where the tile kernel applies
  std::vector<Span> spans;
  cudaError_t span_begin(int kind, cudaStream_t s) {
    static const char* const kNames[SPAN_KINDS] = {"rip:pca_stats", "rip:wb_lut", "rip:fused_chain", "rip:undistort"};
    nvtxRangePushA(kNames[kind]);  // NVTX range around the launch (tracing)
    if (!profile) return cudaSuccess;
    Span sp{kind, nullptr, nullptr};
    cudaError_t e = cudaEventCreate(&sp.a);
    if (e == cudaSuccess) e = cudaEventCreate(&sp.b);
    if (e == cudaSuccess) e = cudaEventRecord(sp.a, s);
    spans.push_back(sp);
    return e;
  }
  cudaError_t span_end(cudaStream_t s) {
    nvtxRangePop();
    if (!profile || spans.empty()) return cudaSuccess;
    return cudaEventRecord(spans.back().b, s);
  }

  // device-resident parameters
  DevBuf d_tables, d_strip_tables; bool tables_valid = false; ChainTableParams tables_key;
  DevBuf d_vig; int vig_rows = -1, vig_cols = -1, vig_angle = -1, vig_pitch = 0; double vig_par[3] = {0, 0, 0};
  DevBuf d_map; uint64_t map_epoch = 0; int map_w = 0, map_h = 0;
  DevBuf d_tiles, d_tmap; int tmap_pitch = 0;  // tile table and tile-padded copy of the packed map (kernels.hpp remap_tile_table)
  DevBuf d_pmap; bool pmap_ok = false; int pmap_src_rows = -1, pmap_src_cols = -1; uint64_t pmap_epoch = 0;  // packed fixed-point map
  std::vector<float> h_map;  // host copy (debug / tests)
  CccState ccc;

  // single-frame path (rip_apply) state, kept for the getters
  DevBuf d_in, d_out, d_tmp;
  Scratch scratch;
  bool have_frame = false;
  FrameGeom last_geom{};
  std::string last_in_encoding;
  float last_pca[4] = {0, 0, 0, 0};

  std::vector<Slot> slots;
  // rip_apply: pinned staging buffers and the captured CUDA graph of the last (shape, configuration)
  uint8_t* h_stage_in = nullptr; size_t h_stage_in_cap = 0;
  uint8_t* h_stage_out = nullptr; size_t h_stage_out_cap = 0;
  cudaGraphExec_t graph_exec = nullptr;
  // what the captured graph is valid for: frame shape / encoding, configuration, and the buffers baked into its nodes
  struct GraphKey {
    int rows = 0, cols = 0, channels = 0; std::string encoding; uint64_t epoch = 0;
    const void *d_in = nullptr, *d_out = nullptr, *h_in = nullptr, *h_out = nullptr, *wb = nullptr, *color = nullptr, *stats = nullptr;
    bool operator==(const GraphKey& o) const {
      return rows == o.rows && cols == o.cols && channels == o.channels && encoding == o.encoding && epoch == o.epoch && d_in == o.d_in &&
             d_out == o.d_out && h_in == o.h_in && h_out == o.h_out && wb == o.wb && color == o.color && stats == o.stats;
    }
  } graph_key;
  int graph_kernel_nodes = 0;  // kernels one replay launches (for stats/kernel_launches)
  // the scratch buffers process_device uses are part of the key: a reallocation invalidates the graph
  void remember_graph_buffers() { graph_key.wb = scratch.wb.ptr; graph_key.color = scratch.color.ptr; graph_key.stats = scratch.stats.ptr; }
  GraphKey graph_key_partial() const { GraphKey k = graph_key; k.wb = k.color = k.stats = nullptr; return k; }
  GraphKey graph_key_full() const { GraphKey k = graph_key_partial(); return (graph_key.wb == scratch.wb.ptr && graph_key.color == scratch.color.ptr && graph_key.stats == scratch.stats.ptr) ? k : GraphKey(); }
  int graph_warm = 0;          // direct (un-captured) runs with the current key: lazy one-time initialisation happens there
  bool use_graph = true;       // "apply/cuda_graph"
  // "apply/register_caller_buffers" (opt-in): rip_apply page-locks the caller's image / output buffers the first time it
  // sees them (cudaHostRegister) and remembers them, so a caller that cycles through a fixed set of ordinary buffers (a
  // camera ring, a reused cv::Mat) gets direct copies without allocating anything specially.  The caller promises that a
  // registered buffer stays mapped while the pipeline lives (a freed-and-reused address would be DMA'd from stale pages).
  bool register_caller_buffers = false;
  struct Registered { const uint8_t* base; size_t bytes; };
  std::vector<Registered> registered;
  uint64_t config_epoch = 1;   // bumped by every setter / loader
  long long graph_replays = 0;
  double apply_us[5] = {0, 0, 0, 0, 0};  // last rip_apply: copy-in, enqueue (launch), wait for the device, copy-out ("stats/apply_us")
  // rip_apply_batch_device: one scratch set per caller stream, so that calls in flight on different streams never share
  // white-balance tables / statistics / intermediates (calls on ONE stream are ordered by the stream itself)
  std::map<cudaStream_t, Scratch> dev_scratch;
  // batch entry points leave the CCC estimate of their last frame on the device; the stats getters fetch it on demand
  const DevBuf* ccc_pending_gains = nullptr;
  cudaStream_t ccc_pending_stream = nullptr;

  int fail(int code, const std::string& msg) { last_error = msg; return code; }
  int cuda_fail(cudaError_t e, const char* what) {
    return fail(RIP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  }
};

namespace {

#define RIP_CUDA(p, expr)                                   \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) return (p)->cuda_fail(_e, #expr); \
  } while (0)

// Host-built tables (colour tables, vignetting mask, undistortion maps) go to the device on the pipeline's own stream and
// the call waits for it.  A plain cudaMemcpy from pageable memory returns once the data sits in the driver's staging
// buffer -- the tail of the DMA may still be in flight, and the pipeline's streams are non-blocking, i.e. not ordered
// behind the legacy stream such a copy runs on: the first kernel after a table build could read the tail of a table
// before it arrived (seen as wrong bottom rows in the first frame after a rebuild).
static cudaError_t upload_tables(rip_pipeline* p, void* dst, const void* src, size_t bytes) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, p->stream);
  return e != cudaSuccess ? e : cudaStreamSynchronize(p->stream);
}

int ensure_cuda(rip_pipeline* p) {
  if (p->cuda_ready) {
    RIP_CUDA(p, cudaSetDevice(p->device));
    return RIP_OK;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return p->fail(RIP_ERR_CUDA, std::string("no CUDA device available (this library has no CPU fallback): ") +
                                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  if (p->device < 0) {
    int cur = 0;
    RIP_CUDA(p, cudaGetDevice(&cur));
    p->device = cur;
  }
  RIP_CUDA(p, cudaSetDevice(p->device));
  cudaDeviceProp prop;
  RIP_CUDA(p, cudaGetDeviceProperties(&prop, p->device));
  if (prop.major < 10)
    return p->fail(RIP_ERR_CUDA, "device compute capability " + std::to_string(prop.major) + "." + std::to_string(prop.minor) +
                                     " < 10.0: this library is built for sm_100a (B200) only");
  p->sm_count = prop.multiProcessorCount;
  RIP_CUDA(p, cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  p->cuda_ready = true;
  return RIP_OK;
}

// debayer.cpp:45-79 + debayer.hpp:74-81: classify the encoding
int classify_encoding(rip_pipeline* p, const std::string& enc, int channels, int& src, int& cfa, std::string& out_enc) {
  src = -1; cfa = 0; out_enc = enc;
  if (enc == "bayer_rggb8") { src = SRC_BAYER; cfa = CFA_RGGB; }
  else if (enc == "bayer_grbg8") { src = SRC_BAYER; cfa = CFA_GRBG; }
  else if (enc == "bayer_gbrg8") { src = SRC_BAYER; cfa = CFA_GBRG; }
  else if (enc == "bayer_bggr8") { src = SRC_BAYER; cfa = CFA_BGGR; }
  if (src == SRC_BAYER) {
    if (channels != 1) return p->fail(RIP_ERR_INVALID_ARGUMENT, "Encoding [" + enc + "] needs a 1-channel image");
    out_enc = "bgr8";
    return RIP_OK;
  }
  if (enc == "rgb8") {
    if (channels != 3) return p->fail(RIP_ERR_INVALID_ARGUMENT, "Encoding [rgb8] needs a 3-channel image");
    src = SRC_RGB;  // CPU branch swaps channels but keeps the encoding string (debayer.cpp:72-73)
    return RIP_OK;
  }
  // EXTENSION (opt-in): 16-bit Bayer is demosaiced at 16 bits and reduced to 8 (frame_math.cuh demosaic_at16)
  if (p->hs.p.debayer_allow_16bit) {
    int c16 = -1;
    if (enc == "bayer_rggb16") c16 = CFA_RGGB; else if (enc == "bayer_grbg16") c16 = CFA_GRBG;
    else if (enc == "bayer_gbrg16") c16 = CFA_GBRG; else if (enc == "bayer_bggr16") c16 = CFA_BGGR;
    if (c16 >= 0) {
      if (channels != 1) return p->fail(RIP_ERR_INVALID_ARGUMENT, "Encoding [" + enc + "] needs a 1-channel image");
      src = SRC_BAYER16; cfa = c16; out_enc = "bgr8";
      return RIP_OK;
    }
  }
  // BAYER_TYPES with the reference's missing comma (debayer.hpp:77-78): these names throw
  static const char* kListed[] = {"bayer_rggb8bayer_bggr16", "bayer_gbrg16", "bayer_grbg16", "bayer_rggb16"};
  for (const char* s : kListed)
    if (enc == s) return p->fail(RIP_ERR_INVALID_ARGUMENT, "Encoding [" + enc + "] is a valid pattern but is not supported!");
  if (channels == 3) { src = SRC_BGR; return RIP_OK; }
  if (channels == 1) { src = SRC_MONO; return RIP_OK; }  // passes through debayer untouched (debayer.cpp:45-79)
  return p->fail(RIP_ERR_INVALID_ARGUMENT, "images must have 1 or 3 channels");
}

int frame_geometry(rip_pipeline* p, int rows, int cols, int channels, const std::string& enc, FrameGeom& g) {
  const Params& q = p->hs.p;
  if (rows < 3 || cols < 3) return p->fail(RIP_ERR_INVALID_ARGUMENT, "image must be at least 3x3");
  g.rows = rows; g.cols = cols; g.channels = channels;
  int rc = classify_encoding(p, enc, channels, g.src, g.cfa, g.out_encoding);
  if (rc != RIP_OK) return rc;
  g.color = g.src != SRC_MONO;
  g.bytes_per_sample = g.src == SRC_BAYER16 ? 2 : 1;
  g.angle = 0;
  if (q.flip_enabled && (q.flip_angle == 90 || q.flip_angle == 180 || q.flip_angle == 270)) g.angle = q.flip_angle;
  const bool swap = (g.angle == 90 || g.angle == 270);
  g.frows = swap ? cols : rows;
  g.fcols = swap ? rows : cols;
  g.undistort = q.und_enabled && q.und_available && q.dist_model != "none";
  g.orows = g.frows; g.ocols = g.fcols; g.ochannels = g.color ? 3 : 1;
  if (g.undistort) {
    if (q.dist_w <= 0 || q.dist_h <= 0) return p->fail(RIP_ERR_INVALID_ARGUMENT, "undistortion enabled without an image size");
    g.orows = q.dist_h; g.ocols = q.dist_w;  // cv::remap output has the map's size (undistortion.cpp:216,241)
  }
  return RIP_OK;
}

int stage_mask(rip_pipeline* p, const FrameGeom& g, uint32_t& stages, int& wb_kind) {
  const Params& q = p->hs.p;
  stages = 0; wb_kind = 0;
  if (q.wb_enabled && g.color) {  // white_balance.hpp:45-86
    if (q.wb_method == "pca") wb_kind = 1;
    else if (q.wb_method == "ccc") wb_kind = 2;
    else if (q.wb_method == "simple" || q.wb_method == "gray_world" || q.wb_method == "grey_world" || q.wb_method == "learned")
      return p->fail(RIP_ERR_UNSUPPORTED, "White Balance method [" + q.wb_method +
                                              "] relies on cv::xphoto and is outside this library's scope; use 'ccc' or 'pca'");
    else
      return p->fail(RIP_ERR_INVALID_ARGUMENT, "White Balance method [" + q.wb_method +
                                                   "] not supported. Supported algorithms: 'simple', 'gray_world', 'learned', 'ccc', 'pca'");
    stages |= ST_WB;
  }
  if (q.cc_enabled && g.color && q.cc_available) stages |= ST_CC;  // color_calibration.hpp:42-56
  if (q.gamma_enabled) stages |= ST_GAMMA;                          // gamma_correction.hpp:32-43
  if (q.vig_enabled) {                                              // vignetting_correction.hpp:26-33
    if (!g.color)  // the reference's cv::cvtColor(BGR2Lab) throws for a 1-channel image (vignetting_correction.cpp:73)
      return p->fail(RIP_ERR_INVALID_ARGUMENT, "vignetting correction needs a 3-channel image (cv::cvtColor BGR2Lab rejects 1 channel)");
    stages |= ST_VIG;
  }
  if (q.enh_enabled && g.color) stages |= ST_ENH;                   // color_enhancer.hpp:33-43
  return RIP_OK;
}

int ensure_tables(rip_pipeline* p) {
  const Params& q = p->hs.p;
  ChainTableParams key;
  key.gamma_enabled = q.gamma_enabled; key.gamma_k = q.gamma_k;
  key.enh_gain[0] = q.enh_hue_gain; key.enh_gain[1] = q.enh_saturation_gain; key.enh_gain[2] = q.enh_value_gain;
  if (p->tables_valid && key == p->tables_key) return RIP_OK;
  std::vector<uint8_t> blob(TABLE_BYTES, 0);
  build_chain_blob(key, blob.data());
  RIP_CUDA(p, cudaDeviceSynchronize());  // nothing in flight may still read the old tables
  RIP_CUDA(p, p->d_tables.reserve(TABLE_BYTES));
  RIP_CUDA(p, upload_tables(p, p->d_tables.ptr, blob.data(), TABLE_BYTES));
  std::vector<uint8_t> sblob(STRIP_BLOB_BYTES, 0);
  build_strip_blob(blob.data(), sblob.data());
  RIP_CUDA(p, p->d_strip_tables.reserve(STRIP_BLOB_BYTES));
  RIP_CUDA(p, upload_tables(p, p->d_strip_tables.ptr, sblob.data(), STRIP_BLOB_BYTES));
  p->tables_valid = true; p->tables_key = key;
  return RIP_OK;
}

// `rows` x `cols`: the INPUT frame; `angle`: the flip applied after the debayer.  The reference builds the mask for the
// flipped image; the device copy is stored in input-frame coordinates (entry (y, x) = mask at the output position of
// input pixel (y, x)), so a thread reads the masks of its four pixels with one aligned 16-byte load whatever the rotation.
int ensure_vignetting(rip_pipeline* p, int rows, int cols, int angle) {
  const Params& q = p->hs.p;
  const double par[3] = {q.vig_scale, q.vig_a2, q.vig_a4};
  if (p->vig_rows == rows && p->vig_cols == cols && p->vig_angle == angle && memcmp(par, p->vig_par, sizeof par) == 0) return RIP_OK;
  const bool swap = angle == 90 || angle == 270;
  const int orows = swap ? cols : rows, ocols = swap ? rows : cols;
  std::vector<float> quad;
  int qr = 0, qc = 0;
  build_vignetting_quadrant(orows, ocols, q.vig_scale, q.vig_a2, q.vig_a4, quad, qr, qc);
  std::vector<float> full((size_t)rows * cols);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      int oy = y, ox = x;  // flip.cpp:37-58 as destination coordinates (inverse of frame_math.cuh flip_source)
      if (angle == 90) { oy = x; ox = rows - 1 - y; }
      else if (angle == 180) { oy = rows - 1 - y; ox = cols - 1 - x; }
      else if (angle == 270) { oy = cols - 1 - x; ox = y; }
      full[(size_t)y * cols + x] = quad[(size_t)(std::abs(2 * oy - orows) >> 1) * qc + (std::abs(2 * ox - ocols) >> 1)];
    }
  RIP_CUDA(p, cudaDeviceSynchronize());
  // four rows of padding (1.0f): the strip kernel reads the mask rows of a whole 4-row chunk even where the frame ends inside it
  full.resize(full.size() + (size_t)4 * cols, 1.0f);
  RIP_CUDA(p, p->d_vig.reserve(full.size() * sizeof(float)));
  RIP_CUDA(p, upload_tables(p, p->d_vig.ptr, full.data(), full.size() * sizeof(float)));
  p->vig_rows = rows; p->vig_cols = cols; p->vig_angle = angle; p->vig_pitch = cols; memcpy(p->vig_par, par, sizeof par);
  return RIP_OK;
}

void build_host_map(rip_pipeline* p) {
  const Params& q = p->hs.p;
  if (p->map_epoch == p->hs.und_epoch && !p->h_map.empty()) return;
  fisheye_rectify_map(q.dist_K, q.dist_D, q.dist_R, q.rect_K, q.dist_w, q.dist_h, p->h_map);
  for (float& v : p->h_map)
    if (v != v) v = -1e9f;  // NaN would round to 0 on the device; make it out-of-range like cvRound(NaN)
  p->map_w = q.dist_w; p->map_h = q.dist_h;
}

int ensure_map(rip_pipeline* p) {
  if (p->map_epoch == p->hs.und_epoch && p->d_map.ptr) return RIP_OK;
  build_host_map(p);
  RIP_CUDA(p, cudaDeviceSynchronize());
  RIP_CUDA(p, p->d_map.reserve(p->h_map.size() * sizeof(float)));
  RIP_CUDA(p, upload_tables(p, p->d_map.ptr, p->h_map.data(), p->h_map.size() * sizeof(float)));
  p->map_epoch = p->hs.und_epoch;
  return RIP_OK;
}

// Packed fixed-point version of the map (4 B/px instead of 8) for a source image of rows x cols; falls back to the
// float map when a displacement does not fit 16 bits.
// Host side of the packed map: the entries, and for the tile kernel the footprint table and the tile-padded copy.
// Returns false when a displacement does not fit 16 bits (the float map stays in use).
bool build_host_packed_map(rip_pipeline* p, int src_rows, int src_cols, std::vector<uint32_t>& packed, std::vector<int>& table,
                           std::vector<uint32_t>& padded, int& tmap_pitch) {
  build_host_map(p);
  const int w = p->map_w, h = p->map_h;
  packed.resize((size_t)w * h);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const float* m = &p->h_map[((size_t)y * w + x) * 2];
      if (!remap_pack_entry(m[0], m[1], x, y, src_rows, src_cols, packed[(size_t)y * w + x])) return false;
    }
  const int tiles_x = (w + REMAP_TILE_W - 1) / REMAP_TILE_W, tiles_y = (h + REMAP_TILE_H - 1) / REMAP_TILE_H;
  table.resize((size_t)4 * tiles_x * tiles_y);
  padded.resize((size_t)tiles_x * REMAP_TILE_W * tiles_y * REMAP_TILE_H);
  remap_tile_table(packed.data(), h, w, table.data(), padded.data());
  tmap_pitch = tiles_x * REMAP_TILE_W;
  return true;
}

int ensure_packed_map(rip_pipeline* p, int src_rows, int src_cols) {
  if (p->pmap_epoch == p->hs.und_epoch && p->pmap_src_rows == src_rows && p->pmap_src_cols == src_cols) return RIP_OK;
  std::vector<uint32_t> packed, padded;
  std::vector<int> table;
  int pitch = 0;
  p->pmap_ok = build_host_packed_map(p, src_rows, src_cols, packed, table, padded, pitch);
  if (p->pmap_ok) {
    RIP_CUDA(p, cudaDeviceSynchronize());
    RIP_CUDA(p, p->d_pmap.reserve(packed.size() * sizeof(uint32_t)));
    RIP_CUDA(p, upload_tables(p, p->d_pmap.ptr, packed.data(), packed.size() * sizeof(uint32_t)));
    RIP_CUDA(p, p->d_tiles.reserve(table.size() * sizeof(int)));
    RIP_CUDA(p, upload_tables(p, p->d_tiles.ptr, table.data(), table.size() * sizeof(int)));
    RIP_CUDA(p, p->d_tmap.reserve(padded.size() * sizeof(uint32_t)));
    RIP_CUDA(p, upload_tables(p, p->d_tmap.ptr, padded.data(), padded.size() * sizeof(uint32_t)));
    p->tmap_pitch = pitch;
  }
  p->pmap_epoch = p->hs.und_epoch; p->pmap_src_rows = src_rows; p->pmap_src_cols = src_cols;
  return RIP_OK;
}

// The pipeline proper, on device memory (raw_image_pipeline.hpp:143-172).
// `keep_bgr_color`: the caller may later ask for getDistColorImage(), so the pre-undistortion image must exist as BGR8;
// otherwise (batch entry points without a dist_color buffer) it is kept in the 4-byte format the gather prefers.
// `no_undistort`: stop after the colour chain (the image getDistColorImage() returns); `reuse_wb`: the white-balance tables
// in `sc` are those of this very frame (left there by the pass that produced the rectified image) -- do not recompute them
// (the CCC Kalman tracker must not advance twice for one frame).
int process_device(rip_pipeline* p, Scratch& sc, const FrameGeom& g, const uint8_t* d_in, size_t in_pitch,
                   size_t in_frame_stride, int n, uint8_t* d_out, size_t out_frame_stride, uint8_t* d_color_user,
                   uint32_t stages_override, bool use_override, cudaStream_t stream, bool keep_bgr_color = true,
                   bool no_undistort = false, bool reuse_wb = false) {
  const Params& q = p->hs.p;
  uint32_t stages = 0;
  int wb_kind = 0;
  if (use_override) stages = stages_override;
  else {
    int rc = stage_mask(p, g, stages, wb_kind);
    if (rc != RIP_OK) return rc;
  }
  const bool undistort = g.undistort && !use_override && !no_undistort;
  int launches = 0;
  if (stages & (ST_GAMMA | ST_VIG | ST_ENH)) { int rc = ensure_tables(p); if (rc != RIP_OK) return rc; }
  if (stages & ST_VIG) { int rc = ensure_vignetting(p, g.rows, g.cols, g.angle); if (rc != RIP_OK) return rc; }
  if (undistort) { int rc = ensure_map(p); if (rc != RIP_OK) return rc; }

  FrameParams fp{};
  fp.in = d_in; fp.in_frame_stride = (long long)in_frame_stride; fp.in_pitch = (int)in_pitch;
  fp.rows = g.rows; fp.cols = g.cols; fp.orows = g.frows; fp.ocols = g.fcols; fp.n_frames = n;
  fp.cfa = g.cfa; fp.angle = g.angle; fp.src = g.src;
  if (g.src == SRC_BAYER16) {  // extension: demosaic at 16 bits, reduce to BGR8, then the chain as for a bgr8 input
    const size_t frame8 = (size_t)g.rows * g.cols * 3;
    RIP_CUDA(p, sc.bgr8.reserve(frame8 * n));
    RIP_CUDA(p, launch_bayer16_to_bgr8(d_in, (long long)in_frame_stride, (int)in_pitch, g.rows, g.cols, n, g.cfa, sc.bgr8.as<uint8_t>(), stream, &launches));
    fp.in = sc.bgr8.as<uint8_t>(); fp.in_frame_stride = (long long)frame8; fp.in_pitch = g.cols * 3; fp.src = SRC_BGR;
  }
  fp.tables = p->d_tables.as<uint8_t>();
  fp.strip_tables = p->d_strip_tables.as<uint8_t>();
  fp.vig = p->d_vig.as<float>(); fp.vig_pitch = p->vig_pitch;
  for (int i = 0; i < 9; ++i) fp.k.cc[i] = q.cc_matrix[i];
  for (int i = 0; i < 3; ++i) fp.k.cc_bias[i] = (float)q.cc_bias[i];  // Scalar double -> fp32 on cv::add
  fp.k.wb_g_identity = 0;
  chain_consts_finish(fp.k);
  const bool fast_in = !p->force_generic && fast_path_ok(fp);
  const bool bgrx = undistort && !d_color_user && !keep_bgr_color && fast_in;
  const int och = g.color ? 3 : 1;
  const size_t color_frame = (size_t)g.frows * g.fcols * (bgrx ? 4 : och);
  if (undistort) {
    if (d_color_user) { fp.out = d_color_user; }
    else { RIP_CUDA(p, sc.color.reserve(color_frame * n)); fp.out = sc.color.as<uint8_t>(); }
    fp.out_frame_stride = (long long)color_frame;
  } else {
    fp.out = d_out; fp.out_frame_stride = (long long)out_frame_stride;
  }
  fp.out_pitch = g.fcols * (bgrx ? 4 : och);

  if (stages & ST_WB) {
    RIP_CUDA(p, sc.wb.reserve((size_t)n * 768 * sizeof(float)));
    fp.wbf = sc.wb.as<float>();
    fp.k.wb_g_identity = wb_kind == 1 ? 1 : 0;  // pca leaves G untouched (white_balance.cpp:117-127)
    if (reuse_wb) {
      // tables of this frame are already in sc.wb
    } else if (wb_kind == 1) {
      RIP_CUDA(p, sc.stats.reserve((size_t)n * 8 * sizeof(unsigned long long)));
      RIP_CUDA(p, sc.coeff.reserve((size_t)n * 4 * sizeof(float)));
      fp.stats = sc.stats.as<unsigned long long>();
      RIP_CUDA(p, p->span_begin(rip_pipeline::SPAN_STATS, stream));
      if (fast_in) RIP_CUDA(p, launch_pca_stats_fast(fp, p->sm_count, stream, &launches));
      else RIP_CUDA(p, launch_pca_stats(fp, p->sm_count, stream, &launches));
      RIP_CUDA(p, p->span_end(stream));
      RIP_CUDA(p, p->span_begin(rip_pipeline::SPAN_LUT, stream));
      RIP_CUDA(p, launch_pca_lut(fp.stats, sc.wb.as<float>(), sc.coeff.as<float>(), n, stream, &launches));
      RIP_CUDA(p, p->span_end(stream));
    } else {
      int rc = ccc_white_balance(p->ccc, q, fp, sc.ccc, sc.gains, p->sm_count, stream, &launches, p->last_error);
      if (rc != RIP_OK) return rc;
      RIP_CUDA(p, launch_gain_lut(static_cast<const float*>(sc.gains.ptr), sc.wb.as<float>(), n, stream, &launches));
    }
  }
  RIP_CUDA(p, p->span_begin(rip_pipeline::SPAN_FUSED, stream));
  if (!g.color) RIP_CUDA(p, launch_mono(fp, (stages & ST_GAMMA) != 0, stream, &launches));
  else if (fast_in && fast_out_ok(fp, bgrx)) {
    // "debug/fused_kernel": 0 = the measured choice per stage set, 1 = always the tile kernel, >= 2 = always the strip kernel
    const bool strip = p->fused_kernel == 0 ? strip_kernel_preferred(stages) : p->fused_kernel != 1;
    if (!strip || !strip_kernel_ok(stages, fp))
      RIP_CUDA(p, launch_fused_fast(stages, fp, bgrx, p->sm_count, stream, &launches));  // tile kernel (round 1)
    else
      RIP_CUDA(p, launch_fused_strip(stages, wb_kind == 2, fp, bgrx, p->fused_kernel, p->sm_count, stream, &launches));
  }
  else if (bgrx) return p->fail(RIP_ERR_CUDA, "internal: 4-byte intermediate needs the fast path");
  else RIP_CUDA(p, launch_fused(stages, fp, p->sm_count, stream, &launches));
  RIP_CUDA(p, p->span_end(stream));
  if (undistort) {
    RemapParams rp{};
    rp.src = fp.out; rp.src_frame_stride = fp.out_frame_stride;
    rp.rows = g.frows; rp.cols = g.fcols; rp.pitch = g.fcols * (bgrx ? 4 : och);
    rp.dst = d_out; rp.dst_frame_stride = (long long)out_frame_stride;
    rp.orows = g.orows; rp.ocols = g.ocols; rp.dpitch = g.ocols * och;
    rp.n_frames = n;
    rp.map = p->d_map.as<float2>();
    rp.pmap = nullptr;
    if (bgrx && !p->force_float_map) {
      int rc = ensure_packed_map(p, g.frows, g.fcols);
      if (rc != RIP_OK) return rc;
      if (p->pmap_ok) { rp.pmap = p->d_pmap.as<uint32_t>(); rp.tiles = p->d_tiles.as<int4>(); rp.tmap = p->d_tmap.as<uint32_t>(); rp.tmap_pitch = p->tmap_pitch; }
    }
    RIP_CUDA(p, p->span_begin(rip_pipeline::SPAN_REMAP, stream));
    if (bgrx && !p->force_gather_remap && remap_tile_ok(rp)) RIP_CUDA(p, launch_remap_tile(rp, p->sm_count, stream, &launches));
    else if (bgrx) RIP_CUDA(p, launch_remap_bgrx(rp, p->sm_count, stream, &launches));
    else RIP_CUDA(p, launch_remap(och, rp, stream, &launches));
    RIP_CUDA(p, p->span_end(stream));
  }
  p->kernel_launches += launches;
  return RIP_OK;
}

// The reference bakes its default YAML paths from __FILE__ (raw_image_pipeline.cpp:8-12); here
// they live in `config/` next to the shared library (override: RIP_B200_CONFIG_DIR).
std::string default_config_dir() {
  const char* env = getenv("RIP_B200_CONFIG_DIR");
  if (env && *env) return env;
  Dl_info info;
  if (dladdr(reinterpret_cast<const void*>(&rip_device_count), &info) && info.dli_fname) {
    std::string so = info.dli_fname;
    const size_t slash = so.rfind('/');
    return (slash == std::string::npos ? std::string(".") : so.substr(0, slash)) + "/config";
  }
  return RIP_DEFAULT_CONFIG_DIR;
}

int create_common(int use_gpu, rip_pipeline** out, rip_pipeline*& p) {
  if (!out) { g_create_error = "out is NULL"; return RIP_ERR_INVALID_ARGUMENT; }
  p = new rip_pipeline();
  p->hs.p.use_gpu = use_gpu != 0;
  p->hs.config_dir = default_config_dir();
  if (const char* v = getenv("RIP_B200_FUSED_KERNEL")) p->fused_kernel = atoi(v);  // experiment switch, see "debug/fused_kernel"
  std::string err;
  if (!ccc_load_model(p->ccc, p->hs.config_dir + "/ccc_model.bin", err)) p->hs.log += "Warning: " + err + "\n";
  *out = p;
  return RIP_OK;
}

bool key_is(const char* key, const char* name) { return strcmp(key, name) == 0; }

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int rip_create(int use_gpu, const char* params_path, const char* calibration_path, const char* color_calibration_path,
               rip_pipeline** out) {
  rip_pipeline* p = nullptr;
  int rc = create_common(use_gpu, out, p);
  if (rc != RIP_OK) return rc;
  const std::string dir = p->hs.config_dir;
  // raw_image_pipeline.cpp:23-40
  p->hs.load_params((params_path && *params_path) ? params_path : dir + "/pipeline_params_example.yaml");
  if (calibration_path && *calibration_path) p->hs.load_camera_calibration(calibration_path);
  p->hs.load_color_calibration((color_calibration_path && *color_calibration_path) ? color_calibration_path
                                                                                   : dir + "/alphasense_color_calib_example.yaml");
  return RIP_OK;
}

int rip_create_default(int use_gpu, rip_pipeline** out) {
  rip_pipeline* p = nullptr;
  int rc = create_common(use_gpu, out, p);
  if (rc != RIP_OK) return rc;
  const std::string dir = p->hs.config_dir;
  // raw_image_pipeline.cpp:16-21
  p->hs.load_params(dir + "/pipeline_params_example.yaml");
  p->hs.load_camera_calibration(dir + "/alphasense_calib_example.yaml");
  p->hs.load_color_calibration(dir + "/alphasense_color_calib_example.yaml");
  return RIP_OK;
}

void rip_destroy(rip_pipeline* p) {
  if (!p) return;
  if (p->cuda_ready) {
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    for (auto& sp : p->spans) {  // profiling events nobody asked for
      if (sp.a) cudaEventDestroy(sp.a);
      if (sp.b) cudaEventDestroy(sp.b);
    }
    if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
    for (const auto& r : p->registered) cudaHostUnregister(const_cast<uint8_t*>(r.base));
    if (p->h_stage_in) cudaFreeHost(p->h_stage_in);
    if (p->h_stage_out) cudaFreeHost(p->h_stage_out);
    ccc_release(p->ccc);
    for (Slot& s : p->slots)
      if (s.stream) cudaStreamDestroy(s.stream);
    if (p->stream) cudaStreamDestroy(p->stream);
  }
  delete p;  // every DevBuf frees its memory in its destructor
}

const char* rip_last_error(const rip_pipeline* p) { return p ? p->last_error.c_str() : g_create_error.c_str(); }

int rip_load_params(rip_pipeline* p, const char* path) { ++p->config_epoch; p->hs.load_params(path ? path : ""); return RIP_OK; }
int rip_load_camera_calibration(rip_pipeline* p, const char* path) { ++p->config_epoch; p->hs.load_camera_calibration(path ? path : ""); return RIP_OK; }
int rip_load_color_calibration(rip_pipeline* p, const char* path) { ++p->config_epoch; p->hs.load_color_calibration(path ? path : ""); return RIP_OK; }
int rip_init_undistortion(rip_pipeline* p) { ++p->config_epoch; p->hs.init_undistortion(); return RIP_OK; }
int rip_reset_white_balance_temporal_consistency(rip_pipeline* p) {
  ++p->config_epoch;
  if (p->hs.p.wb_method == "ccc") p->ccc.pending_reset = true;  // white_balance.cpp:42-47 -> first_frame_ = true
  return RIP_OK;
}

// ---- setters --------------------------------------------------------------------------------
int rip_set_bool(rip_pipeline* p, const char* key, int value) {
  Params& q = p->hs.p;
  ++p->config_epoch;  // a captured rip_apply graph is tied to the configuration it was captured under
  const bool v = value != 0;
  if (key_is(key, "gpu")) q.use_gpu = v;
  else if (key_is(key, "debug")) q.debug = v;
  else if (key_is(key, "profile/kernel_events")) p->profile = v;
  else if (key_is(key, "debug/force_generic_kernels")) p->force_generic = v;
  else if (key_is(key, "debug/force_float_map")) p->force_float_map = v;
  else if (key_is(key, "debug/force_gather_remap")) p->force_gather_remap = v;
  else if (key_is(key, "undistortion/rect_mask")) p->emit_rect_mask = v;
  else if (key_is(key, "apply/cuda_graph")) p->use_graph = v;
  else if (key_is(key, "apply/register_caller_buffers")) p->register_caller_buffers = v;
  else if (key_is(key, "debayer/enabled")) q.debayer_enabled = v;
  else if (key_is(key, "debayer/allow_16bit")) q.debayer_allow_16bit = v;
  else if (key_is(key, "flip/enabled")) q.flip_enabled = v;
  else if (key_is(key, "white_balance/enabled")) q.wb_enabled = v;
  else if (key_is(key, "white_balance/temporal_consistency")) q.wb_temporal_consistency = v;
  else if (key_is(key, "color_calibration/enabled")) q.cc_enabled = v;
  else if (key_is(key, "gamma_correction/enabled")) q.gamma_enabled = v;
  else if (key_is(key, "vignetting_correction/enabled")) q.vig_enabled = v;
  else if (key_is(key, "color_enhancer/enabled")) q.enh_enabled = v;
  else if (key_is(key, "undistortion/enabled")) q.und_enabled = v;
  else return p->fail(RIP_ERR_UNKNOWN_KEY, std::string("unknown bool key: ") + key);
  return RIP_OK;
}

int rip_set_int(rip_pipeline* p, const char* key, int value) {
  ++p->config_epoch;  // a captured rip_apply graph is tied to the configuration it was captured under
  if (key_is(key, "flip/angle")) p->hs.p.flip_angle = value;
  else if (key_is(key, "debug/fused_kernel")) p->fused_kernel = value;
  else return p->fail(RIP_ERR_UNKNOWN_KEY, std::string("unknown int key: ") + key);
  return RIP_OK;
}

int rip_set_double(rip_pipeline* p, const char* key, double value) {
  Params& q = p->hs.p;
  ++p->config_epoch;  // a captured rip_apply graph is tied to the configuration it was captured under
  if (key_is(key, "white_balance/clipping_percentile")) q.wb_clipping_percentile = value;
  else if (key_is(key, "gamma_correction/k")) q.gamma_k = value;
  else if (key_is(key, "color_enhancer/hue_gain")) p->hs.set_hue_gain(value);
  else if (key_is(key, "color_enhancer/saturation_gain")) p->hs.set_saturation_gain(value);
  else if (key_is(key, "color_enhancer/value_gain")) p->hs.set_value_gain(value);
  else if (key_is(key, "undistortion/balance")) { q.und_balance = value; p->hs.init_undistortion(); }
  else if (key_is(key, "undistortion/fov_scale")) { q.und_fov_scale = value; p->hs.init_undistortion(); }
  else return p->fail(RIP_ERR_UNKNOWN_KEY, std::string("unknown double key: ") + key);
  return RIP_OK;
}

int rip_set_string(rip_pipeline* p, const char* key, const char* value) {
  Params& q = p->hs.p;
  ++p->config_epoch;  // a captured rip_apply graph is tied to the configuration it was captured under
  const std::string v = value ? value : "";
  if (key_is(key, "debayer/encoding")) q.debayer_encoding = v;
  else if (key_is(key, "white_balance/method")) q.wb_method = v;
  else if (key_is(key, "gamma_correction/method")) q.gamma_method = v;
  else if (key_is(key, "undistortion/distortion_model")) p->hs.set_distortion_model(v);
  else return p->fail(RIP_ERR_UNKNOWN_KEY, std::string("unknown string key: ") + key);
  return RIP_OK;
}

int rip_set_doubles(rip_pipeline* p, const char* key, const double* v, int n) {
  Params& q = p->hs.p;
  ++p->config_epoch;  // a captured rip_apply graph is tied to the configuration it was captured under
  auto need = [&](int want) -> int {
    if (n < want || !v)
      return p->fail(RIP_ERR_INVALID_ARGUMENT, std::string(key) + " needs " + std::to_string(want) + " values, got " + std::to_string(n));
    return RIP_OK;
  };
  int rc;
  if (key_is(key, "white_balance/saturation_threshold")) { if ((rc = need(2))) return rc; q.wb_bright_thr = v[0]; q.wb_dark_thr = v[1]; }
  else if (key_is(key, "color_calibration/matrix")) { if ((rc = need(9))) return rc; for (int i = 0; i < 9; ++i) q.cc_matrix[i] = (float)v[i]; }
  else if (key_is(key, "color_calibration/bias")) { if ((rc = need(3))) return rc; q.cc_bias[0] = v[0]; q.cc_bias[1] = v[1]; q.cc_bias[2] = v[2]; q.cc_bias[3] = 0; }
  else if (key_is(key, "vignetting_correction/parameters")) { if ((rc = need(3))) return rc; q.vig_scale = v[0]; q.vig_a2 = v[1]; q.vig_a4 = v[2]; }
  else if (key_is(key, "undistortion/image_size")) { if ((rc = need(2))) return rc; p->hs.set_image_size((int)v[0], (int)v[1]); }
  else if (key_is(key, "undistortion/new_image_size")) { if ((rc = need(2))) return rc; p->hs.set_new_image_size((int)v[0], (int)v[1]); }
  else if (key_is(key, "undistortion/camera_matrix")) { if ((rc = need(9))) return rc; p->hs.set_camera_matrix(v); }
  else if (key_is(key, "undistortion/distortion_coefficients")) { if ((rc = need(4))) return rc; p->hs.set_distortion_coefficients(v); }
  else if (key_is(key, "undistortion/rectification_matrix")) { if ((rc = need(9))) return rc; p->hs.set_rectification_matrix(v); }
  else if (key_is(key, "undistortion/projection_matrix")) { if ((rc = need(12))) return rc; p->hs.set_projection_matrix(v); }
  else return p->fail(RIP_ERR_UNKNOWN_KEY, std::string("unknown doubles key: ") + key);
  return RIP_OK;
}

// ---- getters --------------------------------------------------------------------------------
int rip_get_bool(rip_pipeline* p, const char* key, int* value) {
  const Params& q = p->hs.p;
  bool v;
  if (key_is(key, "gpu")) v = q.use_gpu;
  else if (key_is(key, "debug")) v = q.debug;
  else if (key_is(key, "debayer/enabled")) v = q.debayer_enabled;
  else if (key_is(key, "debayer/allow_16bit")) v = q.debayer_allow_16bit;
  else if (key_is(key, "flip/enabled")) v = q.flip_enabled;
  else if (key_is(key, "white_balance/enabled")) v = q.wb_enabled;
  else if (key_is(key, "white_balance/temporal_consistency")) v = q.wb_temporal_consistency;
  else if (key_is(key, "color_calibration/enabled")) v = q.cc_enabled;
  else if (key_is(key, "color_calibration/available")) v = q.cc_available;
  else if (key_is(key, "gamma_correction/enabled")) v = q.gamma_enabled;
  else if (key_is(key, "vignetting_correction/enabled")) v = q.vig_enabled;
  else if (key_is(key, "color_enhancer/enabled")) v = q.enh_enabled;
  else if (key_is(key, "undistortion/enabled")) v = q.und_enabled;
  else if (key_is(key, "undistortion/available")) v = q.und_available;
  else return p->fail(RIP_ERR_UNKNOWN_KEY, std::string("unknown bool key: ") + key);
  *value = v ? 1 : 0;
  return RIP_OK;
}

// the CCC estimate a device-batch call left behind (synchronises that call's stream)
static int ccc_fetch_pending(rip_pipeline* p) {
  if (!p->ccc_pending_gains) return RIP_OK;
  int rc = ensure_cuda(p);
  if (rc != RIP_OK) return rc;
  rc = ccc_fetch_last(p->ccc, *p->ccc_pending_gains, p->ccc_pending_stream, p->last_error);
  p->ccc_pending_gains = nullptr;
  return rc;
}

int rip_get_int(rip_pipeline* p, const char* key, int* value) {
  const Params& q = p->hs.p;
  if (key_is(key, "stats/ccc_u") || key_is(key, "stats/ccc_v")) { int rc = ccc_fetch_pending(p); if (rc != RIP_OK) return rc; }
  if (key_is(key, "flip/angle")) *value = q.flip_angle;
  else if (key_is(key, "dist/image_height")) *value = q.dist_h;
  else if (key_is(key, "dist/image_width")) *value = q.dist_w;
  else if (key_is(key, "rect/image_height")) *value = q.rect_h;
  else if (key_is(key, "rect/image_width")) *value = q.rect_w;
  else if (key_is(key, "stats/kernel_launches")) *value = (int)p->kernel_launches;
  else if (key_is(key, "stats/graph_replays")) *value = (int)p->graph_replays;
  else if (key_is(key, "stats/ccc_u")) *value = p->ccc.uv_x;
  else if (key_is(key, "stats/ccc_v")) *value = p->ccc.uv_y;
  else return p->fail(RIP_ERR_UNKNOWN_KEY, std::string("unknown int key: ") + key);
  return RIP_OK;
}

int rip_get_double(rip_pipeline* p, const char* key, double* value) {
  const Params& q = p->hs.p;
  if (key_is(key, "white_balance/clipping_percentile")) *value = q.wb_clipping_percentile;
  else if (key_is(key, "gamma_correction/k")) *value = q.gamma_k;
  // *member* values (after the reference's cross-wired setters)
  else if (key_is(key, "color_enhancer/hue_gain_member")) *value = q.enh_hue_gain;
  else if (key_is(key, "color_enhancer/saturation_gain_member")) *value = q.enh_saturation_gain;
  else if (key_is(key, "color_enhancer/value_gain_member")) *value = q.enh_value_gain;
  else if (key_is(key, "undistortion/balance")) *value = q.und_balance;
  else if (key_is(key, "undistortion/fov_scale")) *value = q.und_fov_scale;
  else return p->fail(RIP_ERR_UNKNOWN_KEY, std::string("unknown double key: ") + key);
  return RIP_OK;
}

int rip_get_string(rip_pipeline* p, const char* key, char* value, size_t capacity) {
  const Params& q = p->hs.p;
  std::string v;
  if (key_is(key, "debayer/encoding")) v = q.debayer_encoding;
  else if (key_is(key, "white_balance/method")) v = q.wb_method;
  else if (key_is(key, "gamma_correction/method")) v = q.gamma_method;
  else if (key_is(key, "dist/distortion_model")) v = p->hs.dist_distortion_model();
  else if (key_is(key, "rect/distortion_model")) v = p->hs.rect_distortion_model();
  else if (key_is(key, "log")) v = p->hs.log;
  else return p->fail(RIP_ERR_UNKNOWN_KEY, std::string("unknown string key: ") + key);
  if (!value || capacity < v.size() + 1) return p->fail(RIP_ERR_BUFFER_TOO_SMALL, "string buffer too small");
  memcpy(value, v.c_str(), v.size() + 1);
  return RIP_OK;
}

int rip_get_doubles(rip_pipeline* p, const char* key, double* values, int capacity, int* n) {
  const Params& q = p->hs.p;
  std::vector<double> v;
  auto from = [&](const double* s, int c) { v.assign(s, s + c); };
  if (key_is(key, "color_calibration/matrix")) { for (int i = 0; i < 9; ++i) v.push_back((double)q.cc_matrix[i]); }
  else if (key_is(key, "color_calibration/bias")) from(q.cc_bias, 4);  // cv::Mat(cv::Scalar) is 4x1
  else if (key_is(key, "white_balance/saturation_threshold")) { v = {q.wb_bright_thr, q.wb_dark_thr}; }
  else if (key_is(key, "vignetting_correction/parameters")) { v = {q.vig_scale, q.vig_a2, q.vig_a4}; }
  else if (key_is(key, "dist/camera_matrix")) from(q.dist_K, 9);
  else if (key_is(key, "dist/distortion_coefficients")) from(q.dist_D, 4);
  else if (key_is(key, "dist/rectification_matrix")) from(q.dist_R, 9);
  else if (key_is(key, "dist/projection_matrix")) from(q.dist_P, 12);
  else if (key_is(key, "rect/camera_matrix")) from(q.rect_K, 9);
  else if (key_is(key, "rect/distortion_coefficients")) from(q.rect_D, 4);
  else if (key_is(key, "rect/rectification_matrix")) from(q.rect_R, 9);
  else if (key_is(key, "rect/projection_matrix")) from(q.rect_P, 12);
  else if (key_is(key, "stats/pca_coefficients")) { for (float f : p->last_pca) v.push_back((double)f); }
  else if (key_is(key, "stats/kernel_ms")) {
    if (p->cuda_ready) { int rc = ensure_cuda(p); if (rc != RIP_OK) return rc; }
    // {stats, lut, fused, remap} total milliseconds and span counts since the last query; the
    // caller must have synchronised the stream(s) the work was enqueued on
    v.assign(2 * rip_pipeline::SPAN_KINDS, 0.0);
    for (auto& sp : p->spans) {
      float ms = 0.f;
      if (sp.a && sp.b && cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
        v[sp.kind] += ms; v[rip_pipeline::SPAN_KINDS + sp.kind] += 1.0;
      }
      if (sp.a) cudaEventDestroy(sp.a);
      if (sp.b) cudaEventDestroy(sp.b);
    }
    p->spans.clear();
  }
  else if (key_is(key, "stats/apply_us")) { v.assign(p->apply_us, p->apply_us + 5); }
  else if (key_is(key, "stats/ccc_gains")) { int rc = ccc_fetch_pending(p); if (rc != RIP_OK) return rc; v = {(double)p->ccc.gain_b, (double)p->ccc.gain_g, (double)p->ccc.gain_r}; }
  else return p->fail(RIP_ERR_UNKNOWN_KEY, std::string("unknown doubles key: ") + key);
  if (n) *n = (int)v.size();
  if (!values || capacity < (int)v.size()) return p->fail(RIP_ERR_BUFFER_TOO_SMALL, "doubles buffer too small");
  for (size_t i = 0; i < v.size(); ++i) values[i] = v[i];
  return RIP_OK;
}

// host tables for inspection / CPU-side tests (no GPU needed)
int rip_debug_table(rip_pipeline* p, const char* name, int rows, int cols, void* out, size_t capacity, size_t* bytes) {
  std::vector<uint8_t> data;
  if (key_is(name, "gamma_lut")) { data.resize(256); build_gamma_lut(p->hs.p.gamma_k, data.data()); }
  else if (key_is(name, "enhancer_luts")) { data.resize(768); build_enhancer_luts(p->hs.p, data.data()); }
  else if (key_is(name, "vignetting_mask")) {  // expanded to the full rows x cols mask
    std::vector<float> q; int qr, qc;
    build_vignetting_quadrant(rows, cols, p->hs.p.vig_scale, p->hs.p.vig_a2, p->hs.p.vig_a4, q, qr, qc);
    data.resize((size_t)rows * cols * 4);
    float* m = reinterpret_cast<float*>(data.data());
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < cols; ++j) m[(size_t)i * cols + j] = q[(size_t)(std::abs(2 * i - rows) >> 1) * qc + (std::abs(2 * j - cols) >> 1)];
  } else if (key_is(name, "undistortion_map")) {  // interleaved (x, y), dist_h x dist_w
    build_host_map(p);
    data.resize(p->h_map.size() * 4);
    memcpy(data.data(), p->h_map.data(), data.size());
  } else if (key_is(name, "undistortion_packed_map") || key_is(name, "undistortion_tile_table") || key_is(name, "undistortion_tile_map")) {
    // host-side products for a rows x cols source image: packed entries (u32, dist_h x dist_w), the tile kernel's
    // footprint table (4 x i32 per tile) and its tile-padded copy of the packed map
    std::vector<uint32_t> packed, padded;
    std::vector<int> table;
    int pitch = 0;
    if (!build_host_packed_map(p, rows, cols, packed, table, padded, pitch))
      return p->fail(RIP_ERR_UNSUPPORTED, "the undistortion map has displacements that do not fit the packed form");
    const void* src = key_is(name, "undistortion_packed_map") ? (const void*)packed.data()
                      : key_is(name, "undistortion_tile_table") ? (const void*)table.data() : (const void*)padded.data();
    data.resize(key_is(name, "undistortion_packed_map") ? packed.size() * 4
                : key_is(name, "undistortion_tile_table") ? table.size() * 4 : padded.size() * 4);
    memcpy(data.data(), src, data.size());
  } else if (key_is(name, "ccc_response")) {  // last frame of the last call: 256 x 256 fp64, == cv2 response / 65536 - bias
    if (!p->ccc.d_last_response) return p->fail(RIP_ERR_INVALID_ARGUMENT, "no CCC frame processed yet");
    { int rc = ensure_cuda(p); if (rc != RIP_OK) return rc; }
    std::vector<double> cplx(2 * 65536);
    RIP_CUDA(p, cudaDeviceSynchronize());
    RIP_CUDA(p, cudaMemcpy(cplx.data(), p->ccc.d_last_response, cplx.size() * sizeof(double), cudaMemcpyDeviceToHost));
    data.resize(65536 * sizeof(double));
    double* o = reinterpret_cast<double*>(data.data());
    for (int i = 0; i < 65536; ++i) o[i] = cplx[2 * i + p->ccc.last_response_part] * (1.0 / 65536.0);
  } else return p->fail(RIP_ERR_UNKNOWN_KEY, std::string("unknown table: ") + name);
  if (bytes) *bytes = data.size();
  if (!out || capacity < data.size()) return p->fail(RIP_ERR_BUFFER_TOO_SMALL, "table buffer too small");
  memcpy(out, data.data(), data.size());
  return RIP_OK;
}

// ---- frames ---------------------------------------------------------------------------------
int rip_output_shape(rip_pipeline* p, int rows, int cols, int channels, const char* encoding, int* out_rows, int* out_cols,
                     int* out_channels) {
  FrameGeom g;
  int rc = frame_geometry(p, rows, cols, channels, encoding ? encoding : "", g);
  if (rc != RIP_OK) return rc;
  if (out_rows) *out_rows = g.orows;
  if (out_cols) *out_cols = g.ocols;
  if (out_channels) *out_channels = g.ochannels;
  return RIP_OK;
}

namespace {
cudaError_t reserve_pinned(uint8_t*& ptr, size_t& cap, size_t n) {
  if (n <= cap) return cudaSuccess;
  if (ptr) cudaFreeHost(ptr);
  ptr = nullptr; cap = 0;
  cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&ptr), n);
  if (e == cudaSuccess) cap = n;
  return e;
}
}  // namespace

// RawImagePipeline::apply (raw_image_pipeline.cpp:190-205).  One frame, host to host: the caller's image is copied into a
// pinned staging buffer by a few host threads (CopyPool), the H2D copy, the kernels of the chain (4-byte intermediate +
// TMA-staged tile undistortion, like the batch entry points) and the D2H copy run as ONE CUDA-graph launch from the second
// frame of a (shape, configuration) on, and the result is copied out of pinned memory the same way.  The pre-undistortion
// colour image is no longer produced on the way: getDistColorImage() recomputes it on demand from the retained input.
}  // extern "C"

// page-locked host memory the copy engines can address directly (cudaHostAlloc / cudaHostRegister / rip_pinned_alloc)
static bool host_pointer_is_pinned(const void* ptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// "apply/register_caller_buffers": true when [ptr, ptr + bytes) is (now) page-locked
static bool register_caller_buffer(rip_pipeline* p, const uint8_t* ptr, size_t bytes) {
  for (const auto& r : p->registered)
    if (ptr >= r.base && ptr + bytes <= r.base + r.bytes) return true;
  // a bounded set and no eviction: this is for callers that cycle through a few buffers; a caller that keeps bringing new
  // ones (registering costs far more than one staged copy) simply gets the staged path for everything past the 64th
  if (p->registered.size() >= 64) return false;
  if (cudaHostRegister(const_cast<uint8_t*>(ptr), bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return false; }
  p->registered.push_back({ptr, bytes});
  return true;
}

extern "C" {

int rip_pinned_alloc(size_t bytes, void** ptr) {
  if (!ptr || bytes == 0) return RIP_ERR_INVALID_ARGUMENT;
  *ptr = nullptr;
  return cudaHostAlloc(ptr, bytes, cudaHostAllocPortable) == cudaSuccess ? RIP_OK : RIP_ERR_CUDA;
}

int rip_pinned_free(void* ptr) {
  if (!ptr) return RIP_OK;
  return cudaFreeHost(ptr) == cudaSuccess ? RIP_OK : RIP_ERR_CUDA;
}

int rip_apply(rip_pipeline* p, const uint8_t* data, int rows, int cols, int channels, size_t step, char* encoding,
              size_t encoding_capacity, uint8_t* out, size_t out_capacity, int* out_rows, int* out_cols, int* out_channels) {
  if (!data || !out || !encoding) return p->fail(RIP_ERR_INVALID_ARGUMENT, "null argument");
  FrameGeom g;
  int rc = frame_geometry(p, rows, cols, channels, encoding, g);
  if (rc != RIP_OK) return rc;
  uint32_t st; int wbk;
  if ((rc = stage_mask(p, g, st, wbk)) != RIP_OK) return rc;
  const size_t out_bytes = (size_t)g.orows * g.ocols * g.ochannels;
  if (out_capacity < out_bytes) return p->fail(RIP_ERR_BUFFER_TOO_SMALL, "output buffer too small");
  if (g.out_encoding.size() + 1 > encoding_capacity) return p->fail(RIP_ERR_BUFFER_TOO_SMALL, "encoding buffer too small");
  if ((rc = ensure_cuda(p)) != RIP_OK) return rc;
  const size_t row_bytes = (size_t)cols * channels * g.bytes_per_sample;
  const size_t pitch = (row_bytes + 15) & ~(size_t)15;
  if (!step) step = row_bytes;
  RIP_CUDA(p, p->d_in.reserve(pitch * rows));
  RIP_CUDA(p, p->d_out.reserve(out_bytes));
  // Page-locked caller buffers (rip_pinned_alloc, cudaHostAlloc, cudaHostRegister) are read / written by the copy engines
  // directly; pageable ones go through the pipeline's own pinned staging buffers.
  bool direct_in = step == pitch && host_pointer_is_pinned(data);
  bool direct_out = host_pointer_is_pinned(out);
  if (p->register_caller_buffers) {
    if (!direct_in && step == pitch) direct_in = register_caller_buffer(p, data, pitch * rows);
    if (!direct_out) direct_out = register_caller_buffer(p, out, out_bytes);
  }
  if (!direct_in) RIP_CUDA(p, reserve_pinned(p->h_stage_in, p->h_stage_in_cap, pitch * rows));
  if (!direct_out) RIP_CUDA(p, reserve_pinned(p->h_stage_out, p->h_stage_out_cap, out_bytes));
  auto now_us = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  // whatever happens below, no copy engine may still be reading `data` or writing `out` when this call returns
  struct DrainOnExit {
    cudaStream_t s; bool armed = true;
    ~DrainOnExit() { if (armed) cudaStreamSynchronize(s); }
  } drain{p->stream};
  const double t0 = now_us();
  // caller's image -> pinned staging (rows re-pitched to a multiple of 16 bytes: the TMA fast path)
  if (direct_in) RIP_CUDA(p, cudaMemcpyAsync(p->d_in.ptr, data, pitch * rows, cudaMemcpyHostToDevice, p->stream));
  else if (step == pitch) CopyPool::get().copy(p->h_stage_in, data, pitch * rows, /*pinned_side=*/0);
  else { for (int y = 0; y < rows; ++y) memcpy(p->h_stage_in + (size_t)y * pitch, data + (size_t)y * step, row_bytes); demote_range(p->h_stage_in, pitch * rows); }

  auto enqueue = [&]() -> int {  // everything between the two host copies, on p->stream (copies of caller-owned pinned buffers stay outside: their addresses change per call)
    if (!direct_in) RIP_CUDA(p, cudaMemcpyAsync(p->d_in.ptr, p->h_stage_in, pitch * rows, cudaMemcpyHostToDevice, p->stream));
    int r = process_device(p, p->scratch, g, p->d_in.as<uint8_t>(), pitch, pitch * rows, 1, p->d_out.as<uint8_t>(), out_bytes, nullptr, 0,
                           false, p->stream, /*keep_bgr_color=*/false);
    if (r != RIP_OK) return r;
    if (!direct_out) RIP_CUDA(p, cudaMemcpyAsync(p->h_stage_out, p->d_out.ptr, out_bytes, cudaMemcpyDeviceToHost, p->stream));
    return RIP_OK;
  };
  // CCC keeps host-side state per frame (tracker reset flag, result pointers), profiling records events: both stay un-captured
  const bool graph_ok = p->use_graph && !p->profile && wbk != 2;
  rip_pipeline::GraphKey key;
  key.rows = rows; key.cols = cols; key.channels = channels; key.encoding = encoding; key.epoch = p->config_epoch;
  key.d_in = p->d_in.ptr; key.d_out = p->d_out.ptr; key.h_in = direct_in ? nullptr : p->h_stage_in; key.h_out = direct_out ? nullptr : p->h_stage_out;
  bool launched = false;
  const double t1 = now_us();
  if (graph_ok && p->graph_exec && key == p->graph_key_full()) {
    RIP_CUDA(p, cudaGraphLaunch(p->graph_exec, p->stream));
    p->kernel_launches += p->graph_kernel_nodes; ++p->graph_replays;
    launched = true;
  }
  if (!launched) {
    const bool same = graph_ok && key == p->graph_key_partial();
    if (same && p->graph_warm >= 1) {  // second frame of this (shape, configuration): every lazy initialisation has happened
      if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
      const long long launches_before = p->kernel_launches;
      cudaGraph_t graph = nullptr;
      RIP_CUDA(p, cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
      rc = enqueue();
      cudaError_t ce = cudaStreamEndCapture(p->stream, &graph);
      if (rc != RIP_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      RIP_CUDA(p, ce);
      ce = cudaGraphInstantiate(&p->graph_exec, graph, 0);
      cudaGraphDestroy(graph);
      RIP_CUDA(p, ce);
      p->graph_kernel_nodes = (int)(p->kernel_launches - launches_before);
      p->graph_key = key; p->remember_graph_buffers();
      RIP_CUDA(p, cudaGraphLaunch(p->graph_exec, p->stream));
      ++p->graph_replays;
    } else {
      if ((rc = enqueue()) != RIP_OK) return rc;
      if (same) ++p->graph_warm;
      else { p->graph_key = key; p->remember_graph_buffers(); p->graph_warm = graph_ok ? 1 : 0;
             if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; } }
    }
  }
  if (direct_out) RIP_CUDA(p, cudaMemcpyAsync(out, p->d_out.ptr, out_bytes, cudaMemcpyDeviceToHost, p->stream));
  const double t1b = now_us();
  if (wbk == 1) RIP_CUDA(p, cudaMemcpyAsync(p->last_pca, p->scratch.coeff.ptr, sizeof p->last_pca, cudaMemcpyDeviceToHost, p->stream));
  if (wbk == 2 && (rc = ccc_fetch_last(p->ccc, p->scratch.gains, p->stream, p->last_error)) != RIP_OK) return rc;
  p->ccc_pending_gains = nullptr;
  const double t2 = now_us();
  RIP_CUDA(p, cudaStreamSynchronize(p->stream));
  drain.armed = false;
  const double t3 = now_us();
  if (!direct_out) CopyPool::get().copy(out, p->h_stage_out, out_bytes, /*pinned_side=*/1);
  const double t4 = now_us();
  p->apply_us[0] = t1 - t0; p->apply_us[1] = t1b - t1; p->apply_us[2] = t2 - t1b; p->apply_us[3] = t3 - t2; p->apply_us[4] = t4 - t3;
  p->have_frame = true; p->last_geom = g; p->last_in_encoding = encoding;
  memcpy(encoding, g.out_encoding.c_str(), g.out_encoding.size() + 1);
  if (out_rows) *out_rows = g.orows;
  if (out_cols) *out_cols = g.ocols;
  if (out_channels) *out_channels = g.ochannels;
  return RIP_OK;
}

int rip_get_image(rip_pipeline* p, int which, uint8_t* out, size_t out_capacity, int* rows, int* cols, int* channels) {
  auto shape = [&](int r, int c, int ch) { if (rows) *rows = r; if (cols) *cols = c; if (channels) *channels = ch; };
  // undistortion.hpp:136: the reference never writes rect_mask_, getRectMask() is empty.  That stays the default; with
  // "undistortion/rect_mask" set the mask exists: 255 where the rectified pixel interpolates real pixels only.
  if (!p->have_frame || (which == RIP_IMAGE_RECT_MASK && !(p->emit_rect_mask && p->last_geom.undistort))) { shape(0, 0, 0); return RIP_OK; }
  const FrameGeom& g = p->last_geom;
  int rc = ensure_cuda(p);
  if (rc != RIP_OK) return rc;
  if (which == RIP_IMAGE_RECT_MASK) {
    shape(g.orows, g.ocols, 1);
    const size_t bytes = (size_t)g.orows * g.ocols;
    if (!out || out_capacity < bytes) return p->fail(RIP_ERR_BUFFER_TOO_SMALL, "image buffer too small");
    if ((rc = ensure_map(p)) != RIP_OK) return rc;
    RIP_CUDA(p, p->d_tmp.reserve(bytes));
    int launches = 0;
    RIP_CUDA(p, launch_rect_mask(p->d_map.as<float2>(), g.orows, g.ocols, g.frows, g.fcols, p->d_tmp.as<uint8_t>(), p->stream, &launches));
    p->kernel_launches += launches;
    RIP_CUDA(p, cudaMemcpyAsync(out, p->d_tmp.ptr, bytes, cudaMemcpyDeviceToHost, p->stream));
    RIP_CUDA(p, cudaStreamSynchronize(p->stream));
    return RIP_OK;
  }
  const uint8_t* src = nullptr;
  int r = 0, c = 0;
  if (which == RIP_IMAGE_PROCESSED) { src = p->d_out.as<uint8_t>(); r = g.orows; c = g.ocols; }
  else if (which == RIP_IMAGE_DIST_COLOR) {
    r = g.frows; c = g.fcols;
    if (!g.undistort) src = p->d_out.as<uint8_t>();
    else {
      // UndistortionModule's snapshot of its input (undistortion.hpp:68-71): recomputed on demand from the retained input
      // with the white-balance tables the frame was processed with (apply() keeps it only in the 4-byte intermediate)
      const size_t bytes = (size_t)r * c * g.ochannels;
      RIP_CUDA(p, p->d_tmp.reserve(bytes));
      const size_t pitch = (((size_t)g.cols * g.channels * g.bytes_per_sample) + 15) & ~(size_t)15;
      rc = process_device(p, p->scratch, g, p->d_in.as<uint8_t>(), pitch, pitch * g.rows, 1, p->d_tmp.as<uint8_t>(), bytes, nullptr, 0, false,
                          p->stream, /*keep_bgr_color=*/true, /*no_undistort=*/true, /*reuse_wb=*/true);
      if (rc != RIP_OK) return rc;
      src = p->d_tmp.as<uint8_t>();
    }
  } else if (which == RIP_IMAGE_DIST_DEBAYERED) {
    // FlipModule's snapshot (flip.hpp:36-45): recomputed on demand from the retained input
    r = g.frows; c = g.fcols;
    const size_t bytes = (size_t)r * c * g.ochannels;
    RIP_CUDA(p, p->d_tmp.reserve(bytes));
    const size_t pitch = (((size_t)g.cols * g.channels * g.bytes_per_sample) + 15) & ~(size_t)15;
    Scratch dummy;
    rc = process_device(p, dummy, g, p->d_in.as<uint8_t>(), pitch, pitch * g.rows, 1, p->d_tmp.as<uint8_t>(), bytes, nullptr, 0, true,
                        p->stream);
    if (rc != RIP_OK) return rc;
    src = p->d_tmp.as<uint8_t>();
  } else return p->fail(RIP_ERR_INVALID_ARGUMENT, "unknown image id");
  shape(r, c, g.ochannels);
  const size_t bytes = (size_t)r * c * g.ochannels;
  if (!out || out_capacity < bytes) return p->fail(RIP_ERR_BUFFER_TOO_SMALL, "image buffer too small");
  RIP_CUDA(p, cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, p->stream));
  RIP_CUDA(p, cudaStreamSynchronize(p->stream));
  return RIP_OK;
}

int rip_apply_batch_device(rip_pipeline* p, const uint8_t* d_in, size_t in_frame_stride, int n_frames, int rows, int cols,
                           int channels, const char* encoding, uint8_t* d_out, size_t out_frame_stride, uint8_t* d_dist_color,
                           void* cuda_stream) {
  if (!d_in || !d_out || n_frames <= 0) return p->fail(RIP_ERR_INVALID_ARGUMENT, "null argument or empty batch");
  FrameGeom g;
  int rc = frame_geometry(p, rows, cols, channels, encoding ? encoding : "", g);
  if (rc != RIP_OK) return rc;
  if ((rc = ensure_cuda(p)) != RIP_OK) return rc;
  const size_t pitch = (size_t)cols * channels * g.bytes_per_sample;
  if (in_frame_stride < pitch * rows) return p->fail(RIP_ERR_INVALID_ARGUMENT, "in_frame_stride smaller than a frame");
  if (out_frame_stride < (size_t)g.orows * g.ocols * g.ochannels) return p->fail(RIP_ERR_INVALID_ARGUMENT, "out_frame_stride smaller than a frame");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  Scratch& sc = p->dev_scratch[st];
  rc = process_device(p, sc, g, d_in, pitch, in_frame_stride, n_frames, d_out, out_frame_stride, d_dist_color, 0, false, st,
                      /*keep_bgr_color=*/false);
  if (rc == RIP_OK && p->hs.p.wb_enabled && p->hs.p.wb_method == "ccc") { p->ccc_pending_gains = &sc.gains; p->ccc_pending_stream = st; }
  return rc;
}

}  // extern "C"

// The body of the host batch entry points: frames are taken chunk by chunk from `next_chunk` (which returns the first frame
// of the next chunk and its length, or a length of 0 when the batch is exhausted) and pushed through three slots so that the
// copies of chunk i+1 / i-1 overlap the kernels of chunk i.
template <class NextChunk>
static int apply_host_chunks(rip_pipeline* p, const FrameGeom& g, const uint8_t* in, size_t in_frame_stride, int rows, int cols, int channels,
                             uint8_t* out, size_t out_frame_stride, NextChunk next_chunk) {
  const size_t in_frame = (size_t)rows * cols * channels * g.bytes_per_sample, out_frame = (size_t)g.orows * g.ocols * g.ochannels;
  const int kSlots = 3;
  if (p->slots.size() < (size_t)kSlots) {
    const size_t old = p->slots.size();
    p->slots.resize(kSlots);
    for (size_t i = old; i < p->slots.size(); ++i) RIP_CUDA(p, cudaStreamCreateWithFlags(&p->slots[i].stream, cudaStreamNonBlocking));
  }
  // The CCC Kalman tracker is a recurrence over the frames of one camera stream (ccc.cpp:300-340): with temporal
  // consistency on, all chunks go through one slot so that they execute in order on one CUDA stream.
  const Params& q = p->hs.p;
  const int n_slots = (q.wb_enabled && q.wb_method == "ccc" && q.wb_temporal_consistency) ? 1 : kSlots;
  int last_slot = -1;
  // the chunk loop proper; whatever it returns, no copy into or out of the caller's buffers may still be in flight when
  // the entry point returns (the caller is free to release them), so every exit goes through the drain below
  auto run = [&]() -> int {
    for (int slot_i = 0;; slot_i = (slot_i + 1) % n_slots) {
      Slot& s = p->slots[slot_i];
      RIP_CUDA(p, cudaStreamSynchronize(s.stream));  // slot buffers free again; only then is the next chunk claimed
      int f0 = 0, n = 0;
      next_chunk(&f0, &n);
      if (n <= 0) return RIP_OK;
      last_slot = slot_i;
      RIP_CUDA(p, s.in.reserve(in_frame * n));
      RIP_CUDA(p, s.out.reserve(out_frame * n));
      RIP_CUDA(p, cudaMemcpy2DAsync(s.in.ptr, in_frame, in + (size_t)f0 * in_frame_stride, in_frame_stride, in_frame, n,
                                    cudaMemcpyHostToDevice, s.stream));
      const int rc = process_device(p, s.scratch, g, s.in.as<uint8_t>(), (size_t)cols * channels * g.bytes_per_sample, in_frame, n,
                                    s.out.as<uint8_t>(), out_frame, nullptr, 0, false, s.stream, /*keep_bgr_color=*/false);
      if (rc != RIP_OK) return rc;
      RIP_CUDA(p, cudaMemcpy2DAsync(out + (size_t)f0 * out_frame_stride, out_frame_stride, s.out.ptr, out_frame, out_frame, n,
                                    cudaMemcpyDeviceToHost, s.stream));
    }
  };
  int rc = run();
  for (Slot& s : p->slots) {
    const cudaError_t e = cudaStreamSynchronize(s.stream);
    if (e != cudaSuccess && rc == RIP_OK) rc = p->cuda_fail(e, "cudaStreamSynchronize(slot stream)");
  }
  if (rc != RIP_OK) return rc;
  if (last_slot >= 0 && q.wb_enabled && q.wb_method == "ccc") {  // the estimate of the last frame this pipeline saw, like after apply()
    Slot& s = p->slots[last_slot];
    if ((rc = ccc_fetch_last(p->ccc, s.scratch.gains, s.stream, p->last_error)) != RIP_OK) return rc;
    p->ccc_pending_gains = nullptr;
  }
  return RIP_OK;
}

static int host_chunk_frames(size_t in_frame, size_t out_frame) {
  const int chunk = (int)((48u << 20) / (in_frame + out_frame));
  return chunk < 1 ? 1 : (chunk > 16 ? 16 : chunk);
}

extern "C" {

int rip_apply_batch_host(rip_pipeline* p, const uint8_t* in, size_t in_frame_stride, int n_frames, int rows, int cols,
                         int channels, const char* encoding, uint8_t* out, size_t out_frame_stride) {
  if (!in || !out || n_frames <= 0) return p->fail(RIP_ERR_INVALID_ARGUMENT, "null argument or empty batch");
  FrameGeom g;
  int rc = frame_geometry(p, rows, cols, channels, encoding ? encoding : "", g);
  if (rc != RIP_OK) return rc;
  if ((rc = ensure_cuda(p)) != RIP_OK) return rc;
  const size_t in_frame = (size_t)rows * cols * channels * g.bytes_per_sample, out_frame = (size_t)g.orows * g.ocols * g.ochannels;
  if (in_frame_stride < in_frame || out_frame_stride < out_frame) return p->fail(RIP_ERR_INVALID_ARGUMENT, "frame stride smaller than a frame");
  const int chunk = host_chunk_frames(in_frame, out_frame);
  int next = 0;
  return apply_host_chunks(p, g, in, in_frame_stride, rows, cols, channels, out, out_frame_stride, [&](int* f0, int* n) {
    *f0 = next;
    *n = (n_frames - next) < chunk ? (n_frames - next) : chunk;
    next += *n;
  });
}

int rip_apply_batch_host_multi(rip_pipeline* const* handles, int n_handles, const uint8_t* in, size_t in_frame_stride, int n_frames,
                               int rows, int cols, int channels, const char* encoding, uint8_t* out, size_t out_frame_stride) {
  if (!handles || n_handles <= 0 || !handles[0]) return RIP_ERR_INVALID_ARGUMENT;
  rip_pipeline* p0 = handles[0];
  if (!in || !out || n_frames <= 0) return p0->fail(RIP_ERR_INVALID_ARGUMENT, "null argument or empty batch");
  for (int i = 0; i < n_handles; ++i) {
    if (!handles[i]) return p0->fail(RIP_ERR_INVALID_ARGUMENT, "null pipeline handle");
    const Params& q = handles[i]->hs.p;
    if (q.wb_enabled && q.wb_method == "ccc" && q.wb_temporal_consistency && n_handles > 1)
      return p0->fail(RIP_ERR_INVALID_ARGUMENT, "CCC temporal consistency tracks one camera stream: its frames cannot be sharded over pipelines");
  }
  std::vector<FrameGeom> geom(n_handles);
  for (int i = 0; i < n_handles; ++i) {
    const int rc = frame_geometry(handles[i], rows, cols, channels, encoding ? encoding : "", geom[i]);
    if (rc != RIP_OK) {
      if (i != 0) p0->last_error = "pipeline " + std::to_string(i) + ": " + handles[i]->last_error;
      return rc;
    }
  }
  const FrameGeom& g = geom[0];
  const size_t in_frame = (size_t)rows * cols * channels * g.bytes_per_sample, out_frame = (size_t)g.orows * g.ocols * g.ochannels;
  if (in_frame_stride < in_frame || out_frame_stride < out_frame) return p0->fail(RIP_ERR_INVALID_ARGUMENT, "frame stride smaller than a frame");
  // Frames are independent, so the chunks are dealt on demand rather than as fixed shards: the GPUs of one box do
  // not see the same host bandwidth (profiles/pcie_ceiling.json: 7.4 against 15.9 GB/s per GPU with eight busy), and with
  // fixed shards the whole batch waits for the slowest link.  A pipeline claims its next chunk when one of its three slots
  // is free again.  Every frame gets the same bytes whichever pipeline takes it (all pipelines carry the same configuration).
  const int chunk = host_chunk_frames(in_frame, out_frame);
  std::atomic<int> next{0};
  auto claim = [&](int* f0, int* n) {
    const int b = next.fetch_add(chunk, std::memory_order_relaxed);
    *f0 = b;
    *n = b >= n_frames ? 0 : ((n_frames - b) < chunk ? (n_frames - b) : chunk);
  };
  std::vector<int> status(n_handles, RIP_OK);
  std::vector<std::thread> workers;
  for (int i = 0; i < n_handles; ++i)
    workers.emplace_back([&, i] {
      if ((status[i] = ensure_cuda(handles[i])) != RIP_OK) return;  // also makes the pipeline's device current on this thread
      status[i] = apply_host_chunks(handles[i], geom[i], in, in_frame_stride, rows, cols, channels, out, out_frame_stride, claim);
    });
  for (std::thread& t : workers) t.join();
  for (int i = 0; i < n_handles; ++i)
    if (status[i] != RIP_OK) {
      if (i != 0) p0->last_error = "pipeline " + std::to_string(i) + ": " + handles[i]->last_error;
      return status[i];
    }
  return RIP_OK;
}

int rip_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return -1;
  return n;
}

int rip_set_device(rip_pipeline* p, int device) {
  if (p->cuda_ready) return p->fail(RIP_ERR_INVALID_ARGUMENT, "device already selected");
  p->device = device;
  return RIP_OK;
}

}  // extern "C"
