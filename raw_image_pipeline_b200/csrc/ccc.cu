#include "ccc.hpp"

#include <cstdio>
#include <cstring>

#include "../../include/rip_b200.h"

namespace rip {

// File layout (tools/convert_ccc_model.py): "RIPCCC1\0", int32 width, int32 height, then the
// filter and the bias as height x width fp32, already transposed like ccc.cpp:131-132 does.
bool ccc_load_model(CccState& c, const std::string& path, std::string& err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { err = "CCC model file " + path + " not found"; c.model_error = err; return false; }
  char magic[8];
  int wh[2] = {0, 0};
  bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "RIPCCC1\0", 8) == 0 && fread(wh, sizeof(int), 2, f) == 2 &&
            wh[0] > 0 && wh[1] > 0 && wh[0] <= 4096 && wh[1] <= 4096;
  if (ok) {
    c.w = wh[0]; c.h = wh[1];
    const size_t n = (size_t)c.w * c.h;
    c.filter.resize(n); c.bias.resize(n);
    ok = fread(c.filter.data(), sizeof(float), n, f) == n && fread(c.bias.data(), sizeof(float), n, f) == n;
  }
  fclose(f);
  if (!ok) { err = "CCC model file " + path + " is malformed"; c.model_error = err; return false; }
  c.model_loaded = true;
  c.uv_x = c.h / 2; c.uv_y = c.w / 2;  // ccc.cpp:172
  return true;
}

void ccc_release(CccState& c) {
  c.d_filter.release(); c.d_bias.release(); c.d_repeat_tab.release();
  c.device_ready = false;
}

int ccc_white_balance(CccState& c, const Params&, const FrameParams&, DevBuf&, DevBuf&, int, cudaStream_t, int*, std::string& err) {
  (void)c;
  err = "White Balance method [ccc] is not implemented yet in this build";
  return RIP_ERR_UNSUPPORTED;
}

}  // namespace rip
