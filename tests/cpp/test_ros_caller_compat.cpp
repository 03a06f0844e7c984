// Source compatibility with the reference's on-robot caller: every RawImagePipeline method that
// raw_image_pipeline_ros/src/raw_image_pipeline_ros.cpp:52-291 invokes, called here with the argument and result types
// that file uses (bool / int / double / std::string / std::vector<double> parameters read from the ROS parameter
// server; cv::Mat images, masks and calibration matrices handed to the CameraInfo publishers).  Built against the
// stand-in <opencv2/core.hpp> of tests/cpp/fake_opencv; runs without a GPU (no pixel entry point is reached unless the
// first argument is "gpu").
#include <raw_image_pipeline/raw_image_pipeline.hpp>

#include <cstdio>
#include <memory>
#include <string>
#include <vector>

static_assert(RIP_B200_HAS_OPENCV == 1, "this test must see the (fake) OpenCV header");

// what the wrapper's publishers take (raw_image_pipeline_ros.cpp:240-285)
static int camera_info(const cv::Mat& mask, int height, int width, const std::string& model, const cv::Mat& D, const cv::Mat& K,
                       const cv::Mat& R, const cv::Mat& P) {
  (void)mask;
  return height > 0 && width > 0 && !model.empty() && D.depth() == CV_64F && K.rows == 3 && K.cols == 3 && R.rows == 3 && P.rows == 3 &&
                 P.cols == 4
             ? 0
             : 1;
}

int main(int argc, char** argv) {
  const bool gpu = argc >= 2 && std::string(argv[1]) == "gpu";
  const std::string cfg = argc >= 3 ? argv[2] : "";
  // the wrapper holds the pipeline in a std::unique_ptr built from the use_gpu parameter (raw_image_pipeline_ros.cpp:22)
  std::unique_ptr<raw_image_pipeline::RawImagePipeline> pipeline = std::make_unique<raw_image_pipeline::RawImagePipeline>(false);

  bool flag = true;
  int angle = 180, width = 720, height = 540;
  double value = 0.8;
  std::string text;
  std::vector<double> m9 = {347.5, 0, 342.4, 0, 347.4, 271.3, 0, 0, 1}, d4 = {-0.0396, -0.0037, 0.0039, -0.0018};
  std::vector<double> eye9 = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p12 = {347.5, 0, 342.4, 0, 0, 347.4, 271.3, 0, 0, 0, 1, 0};
  std::vector<double> cc9 = {2.43, 0.21, -0.31, 0.09, 1.20, -0.10, -0.24, -0.22, 2.10}, bias3 = {0, 0, 0};

  pipeline->setDebug(false);
  pipeline->setDebayer(flag);
  pipeline->setDebayerEncoding(std::string("auto"));
  pipeline->setFlip(flag);
  pipeline->setFlipAngle(angle);
  pipeline->setWhiteBalance(flag);
  pipeline->setWhiteBalanceMethod(std::string("pca"));
  pipeline->setWhiteBalancePercentile(10.0);
  pipeline->setWhiteBalanceSaturationThreshold(0.8, 0.2);
  pipeline->setWhiteBalanceTemporalConsistency(false);
  pipeline->setColorCalibration(flag);
  if (!cfg.empty()) pipeline->loadColorCalibration(cfg + "/alphasense_color_calib_example.yaml");
  pipeline->setColorCalibrationMatrix(cc9);
  pipeline->setColorCalibrationBias(bias3);
  pipeline->setGammaCorrection(flag);
  pipeline->setGammaCorrectionMethod(std::string("custom"));
  pipeline->setGammaCorrectionK(value);
  pipeline->setVignettingCorrection(flag);
  pipeline->setVignettingCorrectionParameters(1.5, 1e-3, 1e-6);
  pipeline->setColorEnhancer(flag);
  pipeline->setColorEnhancerHueGain(1.0);
  pipeline->setColorEnhancerSaturationGain(1.2);
  pipeline->setColorEnhancerValueGain(1.0);
  pipeline->setUndistortion(flag);
  pipeline->setUndistortionBalance(0.0);
  pipeline->setUndistortionFovScale(value);
  if (!cfg.empty()) pipeline->loadCameraCalibration(cfg + "/alphasense_calib_example.yaml");
  pipeline->setUndistortionImageSize(width, height);
  pipeline->setUndistortionCameraMatrix(m9);
  pipeline->setUndistortionDistortionCoefficients(d4);
  pipeline->setUndistortionDistortionModel(std::string("equidistant"));
  pipeline->setUndistortionRectificationMatrix(eye9);
  pipeline->setUndistortionProjectionMatrix(p12);
  pipeline->initUndistortion();

  if (!pipeline->isDebayerEnabled()) return 2;
  // the one-argument constructor loads the default calibration (raw_image_pipeline.cpp:8-12, 33-36), so it is "available"
  if (!pipeline->isUndistortionEnabled()) return 3;

  if (camera_info(pipeline->getRectMask(), pipeline->getRectImageHeight(), pipeline->getRectImageWidth(), pipeline->getRectDistortionModel(),
                  pipeline->getRectDistortionCoefficients(), pipeline->getRectCameraMatrix(), pipeline->getRectRectificationMatrix(),
                  pipeline->getRectProjectionMatrix()))
    return 4;
  if (camera_info(cv::Mat(), pipeline->getDistImageHeight(), pipeline->getDistImageWidth(), pipeline->getDistDistortionModel(),
                  pipeline->getDistDistortionCoefficients(), pipeline->getDistCameraMatrix(), pipeline->getDistRectificationMatrix(),
                  pipeline->getDistProjectionMatrix()))
    return 5;

  if (gpu) {  // the image callback (raw_image_pipeline_ros.cpp:230-285)
    cv::Mat image(height, width, CV_8UC1);
    for (int i = 0; i < height * width; ++i) image.data[i] = (unsigned char)((i * 13) ^ (i >> 7));
    std::string encoding = "bayer_rggb8";
    if (!pipeline->apply(image, encoding) || encoding != "bgr8" || image.channels() != 3) return 6;
    const cv::Mat debayered = pipeline->getDistDebayeredImage(), color = pipeline->getDistColorImage(), processed = pipeline->getProcessedImage();
    if (debayered.channels() != 3 || color.rows != height || processed.cols != width) return 7;
  }
  pipeline->resetWhiteBalanceTemporalConsistency();
  std::puts("OK");
  return 0;
}
