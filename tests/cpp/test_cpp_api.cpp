// Exercises the header-only C++ class (include/raw_image_pipeline/raw_image_pipeline.hpp) the way
// raw_image_pipeline_ros.cpp drives the reference: construct, configure through setters, apply().
//   test_cpp_api host                      -> host-only checks (no GPU needed), prints "OK"
//   test_cpp_api run <in.raw> <rows> <cols> <encoding> <out.raw> <calib.yaml>
//                                          -> full chain on the GPU, writes the BGR8 result
#include <raw_image_pipeline/raw_image_pipeline.hpp>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

using raw_image_pipeline::Image;
using raw_image_pipeline::RawImagePipeline;

static int host_checks() {
  RawImagePipeline p(false);
  if (!p.isDebayerEnabled() || p.isFlipEnabled() || !p.isWhiteBalanceEnabled() || !p.isUndistortionEnabled()) return 1;
  if (p.getDistImageWidth() != 720 || p.getDistImageHeight() != 540) return 2;
  if (p.getDistDistortionModel() != "equidistant" || p.getRectDistortionModel() != "none") return 3;
  p.setFlip(true); p.setFlipAngle(90);
  if (!p.isFlipEnabled()) return 4;
  p.setWhiteBalanceMethod("not_a_method");
  bool thrown = false;
  try {
    Image img(16, 16, 3);
    std::string enc = "bgr8";
    p.apply(img, enc);
  } catch (const std::invalid_argument&) { thrown = true; }   // white_balance.hpp:81-85
  catch (const std::runtime_error&) { thrown = true; }        // no GPU on this machine: RIP_ERR_CUDA comes first
  if (!thrown) return 5;
  const auto K = p.getDistCameraMatrix();
  if (K.rows != 3 || K.cols != 3) return 6;
  std::puts("OK");
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && std::string(argv[1]) == "host") return host_checks();
  if (argc < 8 || std::string(argv[1]) != "run") { std::fprintf(stderr, "usage\n"); return 64; }
  const int rows = std::atoi(argv[3]), cols = std::atoi(argv[4]);
  std::string enc = argv[5];
  Image img(rows, cols, 1);
  std::ifstream(argv[2], std::ios::binary).read(reinterpret_cast<char*>(img.data), (std::streamsize)rows * cols);
  RawImagePipeline p(false, "", argv[7], "");
  const double sx = cols / 720.0, sy = rows / 540.0;
  p.setFlip(true); p.setFlipAngle(180);
  p.setWhiteBalance(true); p.setWhiteBalanceMethod("pca");
  p.setColorCalibration(true);
  p.setColorCalibrationMatrix({2.4276948, 0.21479778, -0.30818, 0.09277014, 1.1962607, -0.09772757, -0.24436986, -0.22239459, 2.099912});
  p.setGammaCorrection(true); p.setGammaCorrectionMethod("custom"); p.setGammaCorrectionK(0.8);
  p.setVignettingCorrection(true); p.setVignettingCorrectionParameters(1.5, 1e-3, 1e-6);
  p.setColorEnhancer(true); p.setColorEnhancerSaturationGain(1.2);
  p.setUndistortionImageSize(cols, rows);
  p.setUndistortionCameraMatrix({347.548139773951 * sx, 0.0, 342.454373227748 * sx, 0.0, 347.434712422309 * sy, 271.368057185649 * sy, 0.0, 0.0, 1.0});
  p.setUndistortionDistortionCoefficients({-0.0396482888762527, -0.00367688950406141, 0.00391742438164282, -0.00178738156007817});
  p.setUndistortionBalance(0.0); p.setUndistortionFovScale(0.8);
  p.setUndistortion(true);
  if (!p.apply(img, enc)) return 2;
  if (enc != "bgr8" || img.channels() != 3 || img.rows != rows || img.cols != cols) return 3;
  const Image color = p.getDistColorImage();
  if (color.rows != rows || color.cols != cols || !p.getRectMask().empty()) return 4;
  std::ofstream(argv[6], std::ios::binary).write(reinterpret_cast<const char*>(img.data), (std::streamsize)rows * cols * 3);
  std::puts("OK");
  return 0;
}
