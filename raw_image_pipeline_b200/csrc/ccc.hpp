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
  // outputs of the last frame
  int uv_x = 128, uv_y = 128;  // cv::Point uv_pos_ (x = column, y = row of the response arg-max)
  float gain_b = 1.f, gain_g = 1.f, gain_r = 1.f;
  // temporal consistency (ccc.cpp:300-340)
  bool first_frame = true;
  float kf_state[2] = {128.f, 128.f};
  float kf_cov[4] = {0.f, 0.f, 0.f, 0.f};
  // device copies
  DevBuf d_filter, d_bias, d_repeat_tab;
  bool device_ready = false;
};

bool ccc_load_model(CccState& c, const std::string& path, std::string& err);
void ccc_release(CccState& c);

// Computes per-frame gains (B,G,R) for the n frames described by `fp` into `gains` (device,
// n x 3 floats).  Returns a RIP_* status; on failure `err` holds the message.
int ccc_white_balance(CccState& c, const Params& q, const FrameParams& fp, DevBuf& work, DevBuf& gains, int sm_count,
                      cudaStream_t stream, int* launches, std::string& err);

}  // namespace rip
