/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked into the product library.
 *
 * C restatement of the reference's vignetting-mask construction
 * (raw_image_pipeline/src/raw_image_pipeline/modules/vignetting_correction.cpp:32-63,
 * called as precomputeVignettingMask(image.cols, image.rows) at :69).
 *
 * For an image of `rows` x `cols` the reference ends up with a rows x cols CV_32F mask
 * whose entry (i, j) is built from r = sqrt(pow(j - cols/2.0, 2) + pow(i - rows/2.0, 2)),
 * k = pow(r,2)*a2 + pow(r,4)*a4 stored as float, followed by three separate OpenCV fp32
 * passes:  m = k * (float)(1.0 / max)   (MatExpr `mask / max` -> convertTo with alpha=1/max,
 *                                         skipped when max <= 0)
 *          m = m * (float)scale          (MatExpr `mask * scale` -> convertTo alpha=scale)
 *          m = m + 1.0f                  (cv::add with Scalar(1.0))
 * The double expression is evaluated with this box's libm, exactly as the C++ would be.
 */
#include <math.h>

int oracle_vignetting_mask(int rows, int cols, double scale, double a2, double a4, float* out) {
  const double half_c = cols / 2.0;
  const double half_r = rows / 2.0;
  float kmax = -INFINITY;
  for (int j = 0; j < cols; ++j) {
    for (int i = 0; i < rows; ++i) {
      double r = sqrt(pow(j - half_c, 2) + pow(i - half_r, 2));
      double k = pow(r, 2) * a2 + pow(r, 4) * a4;
      float kf = (float)k;
      out[(long)i * cols + j] = kf;
      if (kf > kmax) kmax = kf;
    }
  }
  const long n = (long)rows * cols;
  if ((double)kmax > 0) {
    const float inv = (float)(1.0 / (double)kmax);
    for (long t = 0; t < n; ++t) out[t] = out[t] * inv;
  }
  const float s = (float)scale;
  for (long t = 0; t < n; ++t) out[t] = out[t] * s;
  for (long t = 0; t < n; ++t) out[t] = out[t] + 1.0f;
  return 0;
}
