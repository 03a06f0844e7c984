// Convolutional colour constancy white balance (raw_image_pipeline_white_balance/.../
// convolutional_color_constancy.cpp:91-386): host state + device orchestration.
#pragma once
#include <string>
#include <vector>

#include "devbuf.hpp"
#include "host_state.hpp"
#include "kernels.hpp"

namespace rip {

struct CccState {
  // model (ccc.cpp:116-133): filter and bias, already transposed like the reference does on load
  int w = 0, h = 0;
  std::vector<float> filter, bias;
  bool model_loaded = false;
  std::string model_error;
  // outputs of the last frame of the last call (host copies, filled by ccc_fetch_last)
  int uv_x = 128, uv_y = 128;  // cv::Point uv_pos_ (x = column, y = row of the response arg-max)
  float gain_b = 1.f, gain_g = 1.f, gain_r = 1.f;
  // temporal consistency (ccc.cpp:300-340): the Kalman state lives on the device (d_kf); a reset
  // requested through resetTemporalConsistency() is applied by the next launch
  bool pending_reset = false;
  // device-resident constants
  DevBuf d_filter_fft;  // 256 x 256 double2: 2-D DFT of the filter
  DevBuf d_bias;        // 256 x 256 float
  DevBuf d_twiddle;     // 128 double2: exp(-2 pi i k / 256)
  DevBuf d_weight;      // 97201 float: histogram value after k sequential `+= 1/97200`
  DevBuf d_tabs;        // 256 float cv::log table | 256 float gain table | 40 float Kalman gains
  DevBuf d_kf;          // KfState
  DevBuf d_coef;        // 360 + 270 CccAxisCoef for the current source size
  int coef_rows = -1, coef_cols = -1;
  bool device_ready = false;
  // where the last call left its per-frame results (inside the caller's work buffer)
  const void* d_last_uv = nullptr;        // int2 (x, y) after temporal filtering
  const void* d_last_response = nullptr;  // 256 x 256 double2 holding 65536 * (conv) of the last frame in its real (last_response_part == 0) or imaginary part
  int last_response_part = 0;
  int last_n = 0;
};

bool ccc_load_model(CccState& c, const std::string& path, std::string& err);
void ccc_release(CccState& c);

// Computes per-frame gains (B,G,R) for the n frames described by `fp` into `gains` (device,
// n x 3 floats).  Returns a RIP_* status; on failure `err` holds the message.
int ccc_white_balance(CccState& c, const Params& q, const FrameParams& fp, DevBuf& work, DevBuf& gains, int sm_count,
                      cudaStream_t stream, int* launches, std::string& err);
// copies uv / gains of the last frame of the last call to the host fields (synchronises `stream`)
int ccc_fetch_last(CccState& c, const DevBuf& gains, cudaStream_t stream, std::string& err);

}  // namespace rip
