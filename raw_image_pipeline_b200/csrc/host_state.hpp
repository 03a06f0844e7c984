// Host-side state of one pipeline instance: the parameters of the reference's eight module
// classes plus the host-computed tables the kernels consume.  Plain C++ (no CUDA types) so it
// can be exercised without a GPU.
//
// Mirrors: raw_image_pipeline.cpp:44-165 (loadParams + defaults), modules/*.cpp setters,
// utils.hpp:61-74 (get-with-default), undistortion.cpp:157-238, color_calibration.cpp:52-89,
// gamma_correction.cpp:29-49, vignetting_correction.cpp:32-63.
#pragma once
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

namespace rip {

// ---- tiny YAML subset (block maps, scalars, flow sequences, comments) --------------------
// All four reference config files use nothing else.  Result: "a/b/c" -> raw scalar text.
struct YamlDoc {
  std::map<std::string, std::string> kv;
  bool has(const std::string& k) const { return kv.count(k) != 0; }
  bool get_bool(const std::string& k, bool def) const;
  int get_int(const std::string& k, int def) const;
  double get_double(const std::string& k, double def) const;
  std::string get_string(const std::string& k, const std::string& def) const;
  std::vector<double> get_doubles(const std::string& k) const;
};
bool yaml_load_file(const std::string& path, YamlDoc& doc, std::string& err);
bool file_exists(const std::string& path);

struct Mat33 { double v[9]; };

struct Params {
  bool use_gpu = false, debug = false;
  // debayer.cpp / debayer.hpp
  bool debayer_enabled = true;
  std::string debayer_encoding = "auto";
  bool debayer_allow_16bit = false;  // EXTENSION: accept bayer_*16 (the reference throws for them); "debayer/allow_16bit"
  // flip.cpp
  bool flip_enabled = false;
  int flip_angle = 0;
  // white_balance.cpp
  bool wb_enabled = false;
  std::string wb_method = "ccc";
  double wb_clipping_percentile = 20.0, wb_bright_thr = 0.8, wb_dark_thr = 0.1;
  bool wb_temporal_consistency = true;
  // color_calibration.cpp
  bool cc_enabled = false, cc_available = false;
  float cc_matrix[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // cv::Matx33f
  double cc_bias[4] = {0, 0, 0, 0};                    // cv::Scalar
  // gamma_correction.cpp
  bool gamma_enabled = false;
  std::string gamma_method = "custom";
  double gamma_k = 0.8;
  // vignetting_correction.cpp
  bool vig_enabled = false;
  double vig_scale = 1.5, vig_a2 = 1e-3, vig_a4 = 1e-6;
  // color_enhancer.cpp -- *member* values (uninitialised in the reference; 1.0 here, App. B-5)
  bool enh_enabled = false;
  double enh_hue_gain = 1.0, enh_saturation_gain = 1.0, enh_value_gain = 1.0;
  // undistortion.cpp
  bool und_enabled = false, und_available = false;
  std::string dist_model = "none", rect_model = "none";
  int dist_w = 0, dist_h = 0, rect_w = 0, rect_h = 0;
  double dist_K[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, rect_K[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double dist_D[4] = {0, 0, 0, 0}, rect_D[4] = {0, 0, 0, 0};
  double dist_R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, rect_R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double dist_P[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}, rect_P[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  double und_balance = 0.0, und_fov_scale = 1.0;
};

// Host-computed tables.
// build_gamma_lut (gamma_correction.cpp:35-42) lives in chain_tables.hpp
void build_enhancer_luts(const Params& p, uint8_t lut[768]);             // color_enhancer.cpp:42
// vignetting_correction.cpp:32-63 for a rows x cols image, stored as the (rows/2+1) x (cols/2+1)
// quadrant indexed by (|2i - rows| >> 1, |2j - cols| >> 1): the mask depends only on |i - rows/2|
// and |j - cols/2| (pow(x, 2) is even), so this is lossless.
void build_vignetting_quadrant(int rows, int cols, double scale, double a2, double a4, std::vector<float>& q,
                               int& qrows, int& qcols);
// cv::fisheye::estimateNewCameraMatrixForUndistortRectify (undistortion.cpp:199-208)
void fisheye_new_camera_matrix(const double K[9], const double D[4], int w, int h, const double R[9], double balance,
                               int new_w, int new_h, double fov_scale, double newK[9]);
// cv::fisheye::initUndistortRectifyMap(K, D, R, P, size, CV_32F) (undistortion.cpp:212-220);
// output interleaved (x, y) float pairs, rows x cols.
void fisheye_rectify_map(const double K[9], const double D[4], const double R[9], const double P[9], int w, int h,
                         std::vector<float>& map_xy);

// Pipeline-level host state + the reference's setter semantics.
struct HostState {
  Params p;
  std::string config_dir;  // where the default YAML files live
  uint64_t und_epoch = 1;  // bumped whenever the rectify map must be rebuilt
  std::string log;         // what the reference would have printed to std::cout

  void load_params(const std::string& path);              // raw_image_pipeline.cpp:44-165
  void load_camera_calibration(const std::string& path);  // undistortion.cpp:157-195
  void load_color_calibration(const std::string& path);   // color_calibration.cpp:52-76
  void init_undistortion();                                // undistortion.cpp:197-238 (new K; maps are lazy)

  // enhancer setters are cross-wired in the reference (color_enhancer.cpp:23-33)
  void set_hue_gain(double g) { p.enh_value_gain = g; }
  void set_saturation_gain(double g) { p.enh_saturation_gain = g; }
  void set_value_gain(double g) { p.enh_hue_gain = g; }

  void set_image_size(int w, int h);
  void set_new_image_size(int w, int h);
  void set_camera_matrix(const double* v);
  void set_distortion_coefficients(const double* v);
  void set_distortion_model(const std::string& m);
  void set_rectification_matrix(const double* v);
  void set_projection_matrix(const double* v);

  std::string rect_distortion_model() const;  // undistortion.cpp:94-104
  std::string dist_distortion_model() const;  // undistortion.cpp:106-112
};

}  // namespace rip
