// TEST INFRASTRUCTURE ONLY: the smallest stand-in for <opencv2/core.hpp> that lets the cv::Mat branch of
// include/raw_image_pipeline/raw_image_pipeline.hpp compile in an image without OpenCV's C++ headers.
// It models only what that header touches (8-bit / fp32 / fp64 dense matrices, clone, convertTo).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC(n) CV_MAKETYPE(CV_8U, (n))
#define CV_8UC1 CV_8UC(1)
#define CV_8UC3 CV_8UC(3)

namespace cv {
class Mat {
 public:
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  struct Step { size_t v[2] = {0, 0}; size_t operator[](int i) const { return v[i]; } } step;
  Mat() = default;
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* ext) : rows(r), cols(c), data(static_cast<unsigned char*>(ext)), type_(type) { step.v[0] = (size_t)c * elemSize(); step.v[1] = elemSize(); }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t elemSize() const { static const size_t d[] = {1, 1, 2, 2, 4, 4, 8}; return d[depth()] * channels(); }
  bool empty() const { return rows == 0 || cols == 0 || !data; }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step[0], data + (size_t)r * step[0], (size_t)cols * elemSize());
    return m;
  }
  void convertTo(Mat& dst, int depth) const {
    Mat m(rows, cols, CV_MAKETYPE(depth, channels()));
    const size_t n = (size_t)rows * cols * channels();
    for (size_t i = 0; i < n; ++i) {
      const double v = this->depth() == CV_64F ? reinterpret_cast<const double*>(data)[i]
                     : this->depth() == CV_32F ? reinterpret_cast<const float*>(data)[i] : data[i];
      if (depth == CV_64F) reinterpret_cast<double*>(m.data)[i] = v;
      else if (depth == CV_32F) reinterpret_cast<float*>(m.data)[i] = (float)v;
      else m.data[i] = (unsigned char)v;
    }
    dst = m;
  }
  template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step[0] + (size_t)c * sizeof(T)); }
 private:
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    step.v[0] = (size_t)c * elemSize(); step.v[1] = elemSize();
    store_ = std::make_shared<std::vector<unsigned char>>((size_t)r * step.v[0]);
    data = store_->data();
  }
  int type_ = 0;
  std::shared_ptr<std::vector<unsigned char>> store_;
};
}  // namespace cv
