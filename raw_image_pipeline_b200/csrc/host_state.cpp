#include "host_state.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>
#include <thread>

#include "chain_tables.hpp"

namespace rip {

// ---------------------------------------------------------------------------------------------
// YAML subset
// ---------------------------------------------------------------------------------------------
static std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) ++a;
  while (b > a && isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

static std::string strip_comment(const std::string& s) {
  bool sq = false, dq = false;
  for (size_t i = 0; i < s.size(); ++i) {
    char c = s[i];
    if (c == '\'' && !dq) sq = !sq;
    else if (c == '"' && !sq) dq = !dq;
    else if (c == '#' && !sq && !dq && (i == 0 || isspace((unsigned char)s[i - 1]))) return s.substr(0, i);
  }
  return s;
}

static std::string unquote(const std::string& s) {
  if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\'')))
    return s.substr(1, s.size() - 2);
  return s;
}

bool file_exists(const std::string& path) {
  std::ifstream f(path.c_str());
  return f.good();
}

bool yaml_load_file(const std::string& path, YamlDoc& doc, std::string& err) {
  std::ifstream f(path.c_str());
  if (!f.good()) { err = "cannot open " + path; return false; }
  std::vector<std::pair<int, std::string>> stack;  // (indent, key)
  std::string line;
  std::string pending_key;  // key whose flow sequence continues on following lines
  std::string pending_val;
  int lineno = 0;
  while (std::getline(f, line)) {
    ++lineno;
    std::string body = strip_comment(line);
    if (!pending_key.empty()) {  // multi-line flow sequence
      pending_val += " " + trim(body);
      if (pending_val.find(']') != std::string::npos) { doc.kv[pending_key] = pending_val; pending_key.clear(); }
      continue;
    }
    if (trim(body).empty()) continue;
    int indent = 0;
    while (indent < (int)body.size() && body[indent] == ' ') ++indent;
    std::string t = trim(body);
    if (t == "---" || t == "...") continue;
    size_t colon = std::string::npos;
    {
      bool sq = false, dq = false;
      for (size_t i = 0; i < t.size(); ++i) {
        if (t[i] == '\'' && !dq) sq = !sq;
        else if (t[i] == '"' && !sq) dq = !dq;
        else if (t[i] == ':' && !sq && !dq && (i + 1 == t.size() || isspace((unsigned char)t[i + 1]))) { colon = i; break; }
      }
    }
    if (colon == std::string::npos) {
      std::ostringstream o; o << path << ":" << lineno << ": unsupported YAML construct";
      err = o.str();
      return false;
    }
    std::string key = unquote(trim(t.substr(0, colon)));
    std::string val = trim(t.substr(colon + 1));
    while (!stack.empty() && stack.back().first >= indent) stack.pop_back();
    std::string full;
    for (auto& s : stack) full += s.second + "/";
    full += key;
    if (val.empty()) {
      stack.push_back({indent, key});
      doc.kv[full] = "";  // marks a map node
    } else if (val[0] == '[' && val.find(']') == std::string::npos) {
      pending_key = full; pending_val = val;
    } else {
      doc.kv[full] = val;
    }
  }
  return true;
}

bool YamlDoc::get_bool(const std::string& k, bool def) const {
  auto it = kv.find(k);
  if (it == kv.end()) return def;
  std::string v = unquote(it->second);
  std::transform(v.begin(), v.end(), v.begin(), ::tolower);
  if (v == "true" || v == "yes" || v == "on" || v == "y") return true;
  if (v == "false" || v == "no" || v == "off" || v == "n") return false;
  return def;
}
double YamlDoc::get_double(const std::string& k, double def) const {
  auto it = kv.find(k);
  if (it == kv.end()) return def;
  std::string v = unquote(it->second);
  char* end = nullptr;
  double d = strtod(v.c_str(), &end);
  if (end == v.c_str() || *end != 0) return def;
  return d;
}
int YamlDoc::get_int(const std::string& k, int def) const {
  auto it = kv.find(k);
  if (it == kv.end()) return def;
  std::string v = unquote(it->second);
  char* end = nullptr;
  long d = strtol(v.c_str(), &end, 10);
  if (end == v.c_str() || *end != 0) return def;
  return (int)d;
}
std::string YamlDoc::get_string(const std::string& k, const std::string& def) const {
  auto it = kv.find(k);
  if (it == kv.end() || it->second.empty()) return def;
  return unquote(it->second);
}
std::vector<double> YamlDoc::get_doubles(const std::string& k) const {
  std::vector<double> out;
  auto it = kv.find(k);
  if (it == kv.end()) return out;
  std::string v = it->second;
  size_t a = v.find('['), b = v.rfind(']');
  if (a == std::string::npos || b == std::string::npos || b < a) return out;
  std::stringstream ss(v.substr(a + 1, b - a - 1));
  std::string item;
  while (std::getline(ss, item, ',')) {
    item = trim(item);
    if (item.empty()) continue;
    out.push_back(strtod(item.c_str(), nullptr));
  }
  return out;
}

// ---------------------------------------------------------------------------------------------
// host tables
// ---------------------------------------------------------------------------------------------
void build_enhancer_luts(const Params& p, uint8_t lut[768]) {
  const double g[3] = {p.enh_hue_gain, p.enh_saturation_gain, p.enh_value_gain};
  for (int c = 0; c < 3; ++c)
    for (int x = 0; x < 256; ++x) lut[256 * c + x] = (uint8_t)enh_gain_lut_entry(x, g[c]);
}

void build_vignetting_quadrant(int rows, int cols, double scale, double a2, double a4, std::vector<float>& q, int& qrows,
                               int& qcols) {
  // Entry (qi, qj) <-> every pixel (i, j) with |2i - rows| >> 1 == qi and |2j - cols| >> 1 == qj.
  // Even size: |i - rows/2.0| == qi (qi = rows/2 only at i = 0); odd size: |i - rows/2.0| == qi + 0.5.
  // Every quadrant entry corresponds to at least one real pixel, so max(quadrant) == max(mask).
  // pow(x, 2) is even in x, so evaluating at the non-negative offset reproduces the reference's
  // value for both signs bit for bit.
  qrows = rows / 2 + 1;
  qcols = cols / 2 + 1;
  q.assign((size_t)qrows * qcols, 0.0f);
  const double half_c = cols / 2.0, half_r = rows / 2.0;
  const int i0 = (rows + 1) / 2, j0 = (cols + 1) / 2;  // (i0 + qi) - rows/2.0 == qi (+0.5 if odd)
  auto worker = [&](int r_begin, int r_end, float* local_max) {
    float m = -std::numeric_limits<float>::infinity();
    for (int qi = r_begin; qi < r_end; ++qi) {
      const double dy = (double)(i0 + qi) - half_r;
      for (int qj = 0; qj < qcols; ++qj) {
        const double dx = (double)(j0 + qj) - half_c;
        // vignetting_correction.cpp:43-44 (x <-> column, y <-> row after the (cols, rows) call at :69)
        const double r = std::sqrt(std::pow(dx, 2) + std::pow(dy, 2));
        const double k = std::pow(r, 2) * a2 + std::pow(r, 4) * a4;
        const float kf = (float)k;
        q[(size_t)qi * qcols + qj] = kf;
        if (kf > m) m = kf;
      }
    }
    *local_max = m;
  };
  unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if ((size_t)qrows * qcols < (1u << 16)) nt = 1;
  std::vector<std::thread> th;
  std::vector<float> maxes(nt, -std::numeric_limits<float>::infinity());
  for (unsigned t = 0; t < nt; ++t) {
    int b = (int)((long)qrows * t / nt), e = (int)((long)qrows * (t + 1) / nt);
    if (nt == 1) worker(b, e, &maxes[t]);
    else th.emplace_back(worker, b, e, &maxes[t]);
  }
  for (auto& t : th) t.join();
  float kmax = -std::numeric_limits<float>::infinity();
  for (float m : maxes) kmax = std::max(kmax, m);
  const size_t n = q.size();
  if ((double)kmax > 0) {
    const float inv = (float)(1.0 / (double)kmax);  // MatExpr mask / max -> convertTo(alpha = 1/max)
    for (size_t t = 0; t < n; ++t) q[t] = q[t] * inv;
  }
  const float s = (float)scale;                      // MatExpr mask * scale -> convertTo(alpha = scale)
  for (size_t t = 0; t < n; ++t) q[t] = q[t] * s;
  for (size_t t = 0; t < n; ++t) q[t] = q[t] + 1.0f;  // mask += 1.0
}

// ---- fisheye -------------------------------------------------------------------------------
static void mat33_mul(const double* a, const double* b, double* c) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}

static void mat33_inv(const double* m, double* o) {
  const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  const double det = a * A + b * B + c * C;
  const double id = 1.0 / det;
  o[0] = A * id; o[1] = -(b * i - c * h) * id; o[2] = (b * f - c * e) * id;
  o[3] = B * id; o[4] = (a * i - c * g) * id;  o[5] = -(a * f - c * d) * id;
  o[6] = C * id; o[7] = -(a * h - b * g) * id; o[8] = (a * e - b * d) * id;
}

// cv::fisheye::undistortPoints for one point (TermCriteria COUNT+EPS, 10, 1e-8), R applied, no P.
static void fisheye_undistort_point(double px, double py, const double K[9], const double D[4], const double R[9],
                                    double& ox, double& oy) {
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  const double pwx = (px - cx) / fx, pwy = (py - cy) / fy;
  double theta_d = std::sqrt(pwx * pwx + pwy * pwy);
  const double half_pi = 3.1415926535897932384626433832795 / 2.;
  theta_d = std::min(std::max(-half_pi, theta_d), half_pi);
  bool converged = false;
  double theta = theta_d, scale = 0.0;
  const double eps = 1e-8;
  if (std::fabs(theta_d) > eps) {
    for (int j = 0; j < 10; ++j) {
      const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
      const double k0 = D[0] * t2, k1 = D[1] * t4, k2 = D[2] * t6, k3 = D[3] * t8;
      const double fix = (theta * (1 + k0 + k1 + k2 + k3) - theta_d) / (1 + 3 * k0 + 5 * k1 + 7 * k2 + 9 * k3);
      theta = theta - fix;
      if (std::fabs(fix) < eps) { converged = true; break; }
    }
    scale = std::tan(theta) / theta_d;
  } else {
    converged = true;
  }
  const bool flipped = ((theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0));
  if (converged && !flipped) {
    const double ux = pwx * scale, uy = pwy * scale;
    const double rx = R[0] * ux + R[1] * uy + R[2], ry = R[3] * ux + R[4] * uy + R[5], rz = R[6] * ux + R[7] * uy + R[8];
    ox = rx / rz; oy = ry / rz;
  } else {
    ox = -1000000.0; oy = -1000000.0;
  }
}

void fisheye_new_camera_matrix(const double K[9], const double D[4], int w, int h, const double R[9], double balance,
                               int new_w, int new_h, double fov_scale, double newK[9]) {
  balance = std::min(std::max(balance, 0.0), 1.0);
  double pts[4][2] = {{(double)(w / 2), 0.0}, {(double)w, (double)(h / 2)}, {(double)(w / 2), (double)h}, {0.0, (double)(h / 2)}};
  for (auto& p : pts) fisheye_undistort_point(p[0], p[1], K, D, R, p[0], p[1]);
  double cn0 = (pts[0][0] + pts[1][0] + pts[2][0] + pts[3][0]) * (1.0 / 4), cn1 = (pts[0][1] + pts[1][1] + pts[2][1] + pts[3][1]) * (1.0 / 4);
  const double aspect = K[0] / K[4];
  cn1 *= aspect;
  for (auto& p : pts) p[1] *= aspect;
  double minx = DBL_MAX, miny = DBL_MAX, maxx = -DBL_MAX, maxy = -DBL_MAX;
  for (auto& p : pts) {
    miny = std::min(miny, p[1]); maxy = std::max(maxy, p[1]);
    minx = std::min(minx, p[0]); maxx = std::max(maxx, p[0]);
  }
  const double f1 = w * 0.5 / (cn0 - minx), f2 = w * 0.5 / (maxx - cn0);
  const double f3 = h * 0.5 * aspect / (cn1 - miny), f4 = h * 0.5 * aspect / (maxy - cn1);
  const double fmin = std::min(f1, std::min(f2, std::min(f3, f4)));
  const double fmax = std::max(f1, std::max(f2, std::max(f3, f4)));
  double f = balance * fmin + (1.0 - balance) * fmax;
  f *= fov_scale > 0 ? 1.0 / fov_scale : 1.0;
  double nf0 = f, nf1 = f;
  double nc0 = -cn0 * f + w * 0.5, nc1 = -cn1 * f + (h * aspect) * 0.5;
  nf1 /= aspect; nc1 /= aspect;
  if (new_w > 0 && new_h > 0) {
    const double rx = new_w / (double)w, ry = new_h / (double)h;
    nf0 *= rx; nf1 *= ry; nc0 *= rx; nc1 *= ry;
  }
  const double out[9] = {nf0, 0, nc0, 0, nf1, nc1, 0, 0, 1};
  memcpy(newK, out, sizeof out);
}

void fisheye_rectify_map(const double K[9], const double D[4], const double R[9], const double P[9], int w, int h,
                         std::vector<float>& map_xy) {
  map_xy.assign((size_t)w * h * 2, 0.f);
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  double PR[9], iR[9];
  mat33_mul(P, R, PR);
  mat33_inv(PR, iR);
  const double inf = std::numeric_limits<double>::infinity();
  auto worker = [&](int r0, int r1) {
    for (int i = r0; i < r1; ++i) {
      float* m = &map_xy[(size_t)i * w * 2];
      double _x = i * iR[1] + iR[2], _y = i * iR[4] + iR[5], _w = i * iR[7] + iR[8];
      for (int j = 0; j < w; ++j) {
        double u, v;
        if (_w <= 0) {
          u = (_x > 0) ? -inf : inf;
          v = (_y > 0) ? -inf : inf;
        } else {
          const double x = _x / _w, y = _y / _w;
          const double r = std::sqrt(x * x + y * y);
          const double theta = std::atan(r);
          const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
          const double theta_d = theta * (1 + D[0] * t2 + D[1] * t4 + D[2] * t6 + D[3] * t8);
          const double scale = (r == 0) ? 1.0 : theta_d / r;
          u = fx * x * scale + cx;
          v = fy * y * scale + cy;
        }
        m[2 * j] = (float)u;
        m[2 * j + 1] = (float)v;
        _x += iR[0]; _y += iR[3]; _w += iR[6];
      }
    }
  };
  unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if ((size_t)w * h < (1u << 16)) nt = 1;
  if (nt == 1) { worker(0, h); return; }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) th.emplace_back(worker, (int)((long)h * t / nt), (int)((long)h * (t + 1) / nt));
  for (auto& t : th) t.join();
}

// ---------------------------------------------------------------------------------------------
// HostState
// ---------------------------------------------------------------------------------------------
void HostState::load_params(const std::string& path) {
  log += "Loading raw_image_pipeline params from file " + path + "\n";
  if (!file_exists(path)) {
    // The reference leaves every module pointer null here (raw_image_pipeline.cpp:162-164) and
    // crashes on first use; we keep the YAML defaults instead.
    log += "Warning: parameters file doesn't exist\n";
    return;
  }
  YamlDoc y;
  std::string err;
  if (!yaml_load_file(path, y, err)) { log += "Warning: " + err + "\n"; return; }
  Params d;  // defaults of raw_image_pipeline.cpp:58-153
  p.debayer_enabled = y.get_bool("debayer/enabled", true);
  p.debayer_encoding = y.get_string("debayer/encoding", "auto");
  p.flip_enabled = y.get_bool("flip/enabled", false);
  p.flip_angle = y.get_int("flip/angle", 0);
  p.wb_enabled = y.get_bool("white_balance/enabled", false);
  p.wb_method = y.get_string("white_balance/method", "ccc");
  p.wb_clipping_percentile = y.get_double("white_balance/clipping_percentile", 20.0);
  p.wb_bright_thr = y.get_double("white_balance/saturation_bright_thr", 0.8);
  p.wb_dark_thr = y.get_double("white_balance/saturation_dark_thr", 0.1);
  p.wb_temporal_consistency = y.get_bool("white_balance/temporal_consistency", true);
  p.cc_enabled = y.get_bool("color_calibration/enabled", false);
  p.gamma_enabled = y.get_bool("gamma_correction/enabled", false);
  p.gamma_method = y.get_string("gamma_correction/method", "custom");
  p.gamma_k = y.get_double("gamma_correction/k", 0.8);
  p.vig_enabled = y.get_bool("vignetting_correction/enabled", false);
  p.vig_scale = y.get_double("vignetting_correction/scale", 1.5);
  p.vig_a2 = y.get_double("vignetting_correction/a2", 1e-3);
  p.vig_a4 = y.get_double("vignetting_correction/a4", 1e-6);
  // raw_image_pipeline.cpp:137-145: flag read from `run_color_enhancer`, setHueGain called 3x
  p.enh_enabled = y.get_bool("color_enhancer/run_color_enhancer", false);
  set_hue_gain(y.get_double("color_enhancer/hue_gain", 1.0));
  set_hue_gain(y.get_double("color_enhancer/saturation_gain", 1.0));
  set_hue_gain(y.get_double("color_enhancer/value_gain", 1.0));
  p.und_enabled = y.get_bool("undistortion/enabled", false);
  p.und_balance = y.get_double("undistortion/balance", 0.0);
  p.und_fov_scale = y.get_double("undistortion/fov_scale", 1.0);
  init_undistortion();
  (void)d;
}

void HostState::load_camera_calibration(const std::string& path) {
  log += "Loading camera calibration from file " + path + "\n";
  YamlDoc y;
  std::string err;
  if (file_exists(path) && yaml_load_file(path, y, err)) {
    set_image_size(y.get_int("image_width", 320), y.get_int("image_height", 240));
    std::vector<double> k = y.get_doubles("camera_matrix/data");
    std::vector<double> d = y.get_doubles("distortion_coefficients/data");
    std::vector<double> r = y.get_doubles("rectification_matrix/data");
    std::vector<double> pm = y.get_doubles("projection_matrix/data");
    k.resize(9, 0.0); d.resize(4, 0.0); r.resize(9, 0.0); pm.resize(12, 0.0);
    set_camera_matrix(k.data());
    set_distortion_coefficients(d.data());
    set_distortion_model(y.get_string("distortion_model", "none"));
    set_rectification_matrix(r.data());
    set_projection_matrix(pm.data());
    init_undistortion();
    p.und_available = true;
  } else {
    log += "Warning: Calibration file doesn't exist\n";
    p.und_available = false;
    set_image_size(320, 240);
    // undistortion.cpp:182-185 feeds a 16-element list into Matx33d: the first nine are taken
    const double k16[9] = {1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0};
    const double d4[4] = {0, 0, 0, 0};
    const double r9[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double p12[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    set_camera_matrix(k16);
    set_distortion_coefficients(d4);
    set_distortion_model("none");
    set_rectification_matrix(r9);
    set_projection_matrix(p12);
  }
}

void HostState::load_color_calibration(const std::string& path) {
  log += "Loading color calibration from file " + path + "\n";
  YamlDoc y;
  std::string err;
  if (file_exists(path) && yaml_load_file(path, y, err)) {
    std::vector<double> m = y.get_doubles("matrix/data");
    if (m.size() != 9) m = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; ++i) p.cc_matrix[i] = (float)m[i];
    std::vector<double> b = y.get_doubles("bias/data");
    if (b.size() != 3) b = {0, 0, 0};
    p.cc_bias[0] = b[0]; p.cc_bias[1] = b[1]; p.cc_bias[2] = b[2]; p.cc_bias[3] = 0;
    p.cc_available = true;
  } else {
    p.cc_available = false;
    log += "Warning: Color calibration file doesn't exist\n";
  }
}

void HostState::init_undistortion() {
  if (p.dist_w > 0 && p.dist_h > 0) {
    double nk[9];
    fisheye_new_camera_matrix(p.dist_K, p.dist_D, p.dist_w, p.dist_h, p.dist_R, p.und_balance, p.rect_w, p.rect_h,
                              p.und_fov_scale, nk);
    memcpy(p.rect_K, nk, sizeof nk);
  }
  for (int i = 0; i < 4; ++i) p.rect_D[i] = 0;
  const double eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  memcpy(p.rect_R, eye, sizeof eye);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) p.rect_P[4 * i + j] = p.rect_K[3 * i + j];
  ++und_epoch;
}

void HostState::set_image_size(int w, int h) { p.dist_w = w; p.dist_h = h; p.rect_w = w; p.rect_h = h; init_undistortion(); }
void HostState::set_new_image_size(int w, int h) { p.rect_w = w; p.rect_h = h; init_undistortion(); }
void HostState::set_camera_matrix(const double* v) { memcpy(p.dist_K, v, 72); memcpy(p.rect_K, v, 72); init_undistortion(); }
void HostState::set_distortion_coefficients(const double* v) { memcpy(p.dist_D, v, 32); memcpy(p.rect_D, v, 32); init_undistortion(); }
void HostState::set_distortion_model(const std::string& m) { p.dist_model = m; p.rect_model = m; init_undistortion(); }
void HostState::set_rectification_matrix(const double* v) { memcpy(p.dist_R, v, 72); memcpy(p.rect_R, v, 72); init_undistortion(); }
void HostState::set_projection_matrix(const double* v) { memcpy(p.dist_P, v, 96); memcpy(p.rect_P, v, 96); init_undistortion(); }

std::string HostState::rect_distortion_model() const {
  if (p.und_available) return p.und_enabled ? "none" : p.rect_model;
  return "none";
}
std::string HostState::dist_distortion_model() const { return p.und_available ? p.dist_model : "none"; }

}  // namespace rip
