// Drop-in replacement of the reference's `raw_image_pipeline::RawImagePipeline`
// (raw_image_pipeline/include/raw_image_pipeline/raw_image_pipeline.hpp:36-137), header-only, over
// the C ABI of librip_b200.so (include/rip_b200.h).  Same class name, namespace, constructors,
// method names, argument meaning and exception types, so raw_image_pipeline_ros.cpp and
// raw_image_pipeline_python.cpp compile against it unchanged.
//
// Image type: where OpenCV's headers are available the interface uses cv::Mat exactly like the
// reference; otherwise (e.g. this repository's own CI image, which has no OpenCV C++ headers) a
// minimal value type with the same `rows / cols / channels() / data / clone()` surface is used.
#pragma once

#include <rip_b200.h>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#if defined(RIP_B200_FORCE_PLAIN_IMAGE)
#define RIP_B200_HAS_OPENCV 0
#elif defined(__has_include)
#if __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>
#define RIP_B200_HAS_OPENCV 1
#else
#define RIP_B200_HAS_OPENCV 0
#endif
#else
#define RIP_B200_HAS_OPENCV 0
#endif

namespace raw_image_pipeline {

#if RIP_B200_HAS_OPENCV
using Image = cv::Mat;
using Matrix = cv::Mat;
namespace detail {
inline Image make_image(int rows, int cols, int channels) { return cv::Mat(rows, cols, CV_8UC(channels)); }
inline Matrix make_matrix(int rows, int cols, const double* v) { return cv::Mat(rows, cols, CV_64F, const_cast<double*>(v)).clone(); }
inline Matrix make_matrix_f(int rows, int cols, const double* v) { cv::Mat m; cv::Mat(rows, cols, CV_64F, const_cast<double*>(v)).convertTo(m, CV_32F); return m; }
inline const uint8_t* image_data(const Image& m) { return m.data; }
inline uint8_t* image_data(Image& m) { return m.data; }
inline size_t image_step(const Image& m) { return m.step[0]; }
}  // namespace detail
#else
// rows x cols x channels, 8-bit, densely packed
struct Image {
  int rows = 0, cols = 0;
  std::vector<uint8_t> storage;
  uint8_t* data = nullptr;
  Image() = default;
  Image(int r, int c, int ch) : rows(r), cols(c), storage((size_t)r * c * ch), data(storage.data()), ch_(ch) {}
  Image(const Image& o) : rows(o.rows), cols(o.cols), storage(o.storage), data(storage.data()), ch_(o.ch_) {}
  Image& operator=(const Image& o) { rows = o.rows; cols = o.cols; storage = o.storage; data = storage.data(); ch_ = o.ch_; return *this; }
  int channels() const { return ch_; }
  bool empty() const { return rows == 0 || cols == 0; }
  Image clone() const { return *this; }
 private:
  int ch_ = 1;
};
// small row-major matrix of doubles (what the reference returns as CV_64F cv::Mat)
struct Matrix {
  int rows = 0, cols = 0;
  std::vector<double> v;
  double at(int r, int c) const { return v[(size_t)r * cols + c]; }
};
namespace detail {
inline Image make_image(int rows, int cols, int channels) { return Image(rows, cols, channels); }
inline Matrix make_matrix(int rows, int cols, const double* v) { Matrix m; m.rows = rows; m.cols = cols; m.v.assign(v, v + rows * cols); return m; }
inline Matrix make_matrix_f(int rows, int cols, const double* v) { return make_matrix(rows, cols, v); }
inline const uint8_t* image_data(const Image& m) { return m.data; }
inline uint8_t* image_data(Image& m) { return m.data; }
inline size_t image_step(const Image& m) { return (size_t)m.cols * m.channels(); }
}  // namespace detail
#endif

class RawImagePipeline {
 public:
  // raw_image_pipeline.cpp:16-21
  explicit RawImagePipeline(bool use_gpu) { check_create(rip_create_default(use_gpu ? 1 : 0, &h_)); }
  // raw_image_pipeline.cpp:23-40
  RawImagePipeline(bool use_gpu, const std::string& params_path, const std::string& calibration_path,
                   const std::string& color_calibration_path) {
    check_create(rip_create(use_gpu ? 1 : 0, params_path.c_str(), calibration_path.c_str(), color_calibration_path.c_str(), &h_));
  }
  ~RawImagePipeline() { rip_destroy(h_); }
  RawImagePipeline(const RawImagePipeline&) = delete;
  RawImagePipeline& operator=(const RawImagePipeline&) = delete;

  //-----------------------------------------------------------------------------
  // Main interfaces
  //-----------------------------------------------------------------------------
  // raw_image_pipeline.cpp:190-205: in place; the image may change type/size (1ch -> 3ch, flip 90/270)
  bool apply(Image& image, std::string& encoding) {
    Image out = run(image, encoding);
    image = out;
    return true;
  }
  // raw_image_pipeline.cpp:182-188
  Image process(const Image& image, std::string& encoding) { return run(image, encoding); }

  // Loaders
  void loadParams(const std::string& file_path) { check(rip_load_params(h_, file_path.c_str())); }
  void loadCameraCalibration(const std::string& file_path) { check(rip_load_camera_calibration(h_, file_path.c_str())); }
  void loadColorCalibration(const std::string& file_path) { check(rip_load_color_calibration(h_, file_path.c_str())); }
  void initUndistortion() { check(rip_init_undistortion(h_)); }

  // Other interfaces
  void resetWhiteBalanceTemporalConsistency() { check(rip_reset_white_balance_temporal_consistency(h_)); }
  void setGpu(bool use_gpu) { set_bool("gpu", use_gpu); }
  void setDebug(bool debug) { set_bool("debug", debug); }

  //-----------------------------------------------------------------------------
  // Setters
  //-----------------------------------------------------------------------------
  void setDebayer(bool enabled) { set_bool("debayer/enabled", enabled); }
  void setDebayerEncoding(const std::string& encoding) { set_string("debayer/encoding", encoding); }
  void setFlip(bool enabled) { set_bool("flip/enabled", enabled); }
  void setFlipAngle(int angle) { check(rip_set_int(h_, "flip/angle", angle)); }
  void setWhiteBalance(bool enabled) { set_bool("white_balance/enabled", enabled); }
  void setWhiteBalanceMethod(const std::string& method) { set_string("white_balance/method", method); }
  void setWhiteBalancePercentile(const double& percentile) { set_double("white_balance/clipping_percentile", percentile); }
  void setWhiteBalanceSaturationThreshold(const double& bright_thr, const double& dark_thr) {
    set_doubles("white_balance/saturation_threshold", {bright_thr, dark_thr});
  }
  void setWhiteBalanceTemporalConsistency(bool enabled) { set_bool("white_balance/temporal_consistency", enabled); }
  void setColorCalibration(bool enabled) { set_bool("color_calibration/enabled", enabled); }
  void setColorCalibrationMatrix(const std::vector<double>& m) { set_doubles("color_calibration/matrix", m); }
  void setColorCalibrationBias(const std::vector<double>& b) { set_doubles("color_calibration/bias", b); }
  Matrix getColorCalibrationMatrix() const { double v[16]; get_doubles("color_calibration/matrix", v); return detail::make_matrix_f(3, 3, v); }
  Matrix getColorCalibrationBias() const { double v[16]; get_doubles("color_calibration/bias", v); return detail::make_matrix(4, 1, v); }
  void setGammaCorrection(bool enabled) { set_bool("gamma_correction/enabled", enabled); }
  void setGammaCorrectionMethod(const std::string& method) { set_string("gamma_correction/method", method); }
  void setGammaCorrectionK(const double& k) { set_double("gamma_correction/k", k); }
  void setVignettingCorrection(bool enabled) { set_bool("vignetting_correction/enabled", enabled); }
  void setVignettingCorrectionParameters(const double& scale, const double& a2, const double& a4) {
    set_doubles("vignetting_correction/parameters", {scale, a2, a4});
  }
  void setColorEnhancer(bool enabled) { set_bool("color_enhancer/enabled", enabled); }
  void setColorEnhancerHueGain(const double& gain) { set_double("color_enhancer/hue_gain", gain); }
  void setColorEnhancerSaturationGain(const double& gain) { set_double("color_enhancer/saturation_gain", gain); }
  void setColorEnhancerValueGain(const double& gain) { set_double("color_enhancer/value_gain", gain); }
  void setUndistortion(bool enabled) { set_bool("undistortion/enabled", enabled); }
  void setUndistortionImageSize(int width, int height) { set_doubles("undistortion/image_size", {(double)width, (double)height}); }
  void setUndistortionNewImageSize(int width, int height) { set_doubles("undistortion/new_image_size", {(double)width, (double)height}); }
  void setUndistortionBalance(double balance) { set_double("undistortion/balance", balance); }
  void setUndistortionFovScale(double fov_scale) { set_double("undistortion/fov_scale", fov_scale); }
  void setUndistortionCameraMatrix(const std::vector<double>& m) { set_doubles("undistortion/camera_matrix", m); }
  void setUndistortionDistortionCoefficients(const std::vector<double>& c) { set_doubles("undistortion/distortion_coefficients", c); }
  void setUndistortionDistortionModel(const std::string& model) { set_string("undistortion/distortion_model", model); }
  void setUndistortionRectificationMatrix(const std::vector<double>& m) { set_doubles("undistortion/rectification_matrix", m); }
  void setUndistortionProjectionMatrix(const std::vector<double>& m) { set_doubles("undistortion/projection_matrix", m); }

  //-----------------------------------------------------------------------------
  // Getters
  //-----------------------------------------------------------------------------
  bool isDebayerEnabled() const { return get_bool("debayer/enabled"); }
  bool isFlipEnabled() const { return get_bool("flip/enabled"); }
  bool isWhiteBalanceEnabled() const { return get_bool("white_balance/enabled"); }
  bool isColorCalibrationEnabled() const { return get_bool("color_calibration/enabled"); }
  bool isGammaCorrectionEnabled() const { return get_bool("gamma_correction/enabled"); }
  bool isVignettingCorrectionEnabled() const { return get_bool("vignetting_correction/enabled"); }
  bool isColorEnhancerEnabled() const { return get_bool("color_enhancer/enabled"); }
  bool isUndistortionEnabled() const { return get_bool("undistortion/enabled"); }

  int getDistImageHeight() const { return get_int("dist/image_height"); }
  int getDistImageWidth() const { return get_int("dist/image_width"); }
  std::string getDistDistortionModel() const { return get_string("dist/distortion_model"); }
  Matrix getDistCameraMatrix() const { return get_matrix("dist/camera_matrix", 3, 3); }
  Matrix getDistDistortionCoefficients() const { return get_matrix("dist/distortion_coefficients", 1, 4); }
  Matrix getDistRectificationMatrix() const { return get_matrix("dist/rectification_matrix", 3, 3); }
  Matrix getDistProjectionMatrix() const { return get_matrix("dist/projection_matrix", 3, 4); }
  int getRectImageHeight() const { return get_int("rect/image_height"); }
  int getRectImageWidth() const { return get_int("rect/image_width"); }
  std::string getRectDistortionModel() const { return get_string("rect/distortion_model"); }
  Matrix getRectCameraMatrix() const { return get_matrix("rect/camera_matrix", 3, 3); }
  Matrix getRectDistortionCoefficients() const { return get_matrix("rect/distortion_coefficients", 1, 4); }
  Matrix getRectRectificationMatrix() const { return get_matrix("rect/rectification_matrix", 3, 3); }
  Matrix getRectProjectionMatrix() const { return get_matrix("rect/projection_matrix", 3, 4); }

  Image getDistDebayeredImage() const { return get_image(RIP_IMAGE_DIST_DEBAYERED); }
  Image getDistColorImage() const { return get_image(RIP_IMAGE_DIST_COLOR); }
  Image getRectMask() const { return get_image(RIP_IMAGE_RECT_MASK); }
  Image getProcessedImage() const { return get_image(RIP_IMAGE_PROCESSED); }

  // not in the reference: the underlying C handle (batch / device entry points of rip_b200.h)
  rip_pipeline* handle() const { return h_; }

 private:
  rip_pipeline* h_ = nullptr;

  [[noreturn]] static void raise(int code, const char* msg) {
    const std::string m = msg ? msg : "raw_image_pipeline error";
    if (code == RIP_ERR_INVALID_ARGUMENT) throw std::invalid_argument(m);  // what the reference throws
    throw std::runtime_error(m);
  }
  void check_create(int rc) {
    if (rc != RIP_OK) raise(rc, rip_last_error(nullptr));
  }
  void check(int rc) const {
    if (rc != RIP_OK) raise(rc, rip_last_error(h_));
  }
  void set_bool(const char* k, bool v) { check(rip_set_bool(h_, k, v ? 1 : 0)); }
  void set_double(const char* k, double v) { check(rip_set_double(h_, k, v)); }
  void set_string(const char* k, const std::string& v) { check(rip_set_string(h_, k, v.c_str())); }
  void set_doubles(const char* k, const std::vector<double>& v) { check(rip_set_doubles(h_, k, v.data(), (int)v.size())); }
  bool get_bool(const char* k) const { int v = 0; check(rip_get_bool(h_, k, &v)); return v != 0; }
  int get_int(const char* k) const { int v = 0; check(rip_get_int(h_, k, &v)); return v; }
  std::string get_string(const char* k) const { char buf[256]; check(rip_get_string(h_, k, buf, sizeof buf)); return buf; }
  int get_doubles(const char* k, double v[16]) const { int n = 0; check(rip_get_doubles(h_, k, v, 16, &n)); return n; }
  Matrix get_matrix(const char* k, int rows, int cols) const { double v[16]; get_doubles(k, v); return detail::make_matrix(rows, cols, v); }

  Image run(const Image& image, std::string& encoding) {
    int orows = 0, ocols = 0, och = 0;
    check(rip_output_shape(h_, image.rows, image.cols, image.channels(), encoding.c_str(), &orows, &ocols, &och));
    Image out = detail::make_image(orows, ocols, och);
    char enc[64];
    std::strncpy(enc, encoding.c_str(), sizeof enc - 1);
    enc[sizeof enc - 1] = 0;
    check(rip_apply(h_, detail::image_data(image), image.rows, image.cols, image.channels(), detail::image_step(image), enc, sizeof enc,
                    detail::image_data(out), (size_t)orows * ocols * och, &orows, &ocols, &och));
    encoding = enc;
    return out;
  }
  Image get_image(int which) const {
    int r = 0, c = 0, ch = 0;
    int rc = rip_get_image(h_, which, nullptr, 0, &r, &c, &ch);
    if (rc != RIP_OK && rc != RIP_ERR_BUFFER_TOO_SMALL) check(rc);
    if (r == 0 || c == 0) return Image();
    Image out = detail::make_image(r, c, ch);
    check(rip_get_image(h_, which, detail::image_data(out), (size_t)r * c * ch, &r, &c, &ch));
    return out;
  }
};

}  // namespace raw_image_pipeline
