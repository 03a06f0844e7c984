// Compiles the cv::Mat branch of the drop-in header against a minimal stand-in for <opencv2/core.hpp>
// (tests/cpp/fake_opencv) and drives it the way raw_image_pipeline_ros.cpp does: cv::Mat in/out, std::vector<double>
// setters, cv::Mat matrix getters.  `host` mode needs no GPU.
#include <raw_image_pipeline/raw_image_pipeline.hpp>

#include <cstdio>
#include <string>

static_assert(RIP_B200_HAS_OPENCV == 1, "this test must see the (fake) OpenCV header");

int main(int argc, char** argv) {
  raw_image_pipeline::RawImagePipeline p(false);
  p.setUndistortionCameraMatrix({1, 0, 2, 0, 3, 4, 0, 0, 1});
  const cv::Mat K = p.getDistCameraMatrix();
  if (K.rows != 3 || K.cols != 3 || K.depth() != CV_64F) return 1;
  if (reinterpret_cast<const double*>(K.data)[5] != 4.0) return 2;
  const cv::Mat M = p.getColorCalibrationMatrix();
  if (M.depth() != CV_32F || M.rows != 3) return 3;
  if (!p.getRectMask().empty() || !p.getProcessedImage().empty()) return 4;
  if (argc >= 2 && std::string(argv[1]) == "gpu") {
    cv::Mat img(64, 96, CV_8UC1);
    for (int i = 0; i < 64 * 96; ++i) img.data[i] = (unsigned char)(i * 7);
    std::string enc = "bayer_rggb8";
    p.setWhiteBalance(false); p.setUndistortion(false); p.setGammaCorrection(true);
    if (!p.apply(img, enc) || enc != "bgr8" || img.channels() != 3 || img.rows != 64 || img.cols != 96) return 5;
    const cv::Mat out = p.process(img, enc);
    if (out.channels() != 3) return 6;
  }
  std::puts("OK");
  return 0;
}
