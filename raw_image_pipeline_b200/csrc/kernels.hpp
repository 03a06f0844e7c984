// Launch interface between the host pipeline (rip_api.cu) and the sm_100a kernels
// (rip_kernels.cu).  Plain structs; everything device-side is a raw pointer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pixel_math.cuh"

namespace rip {

// byte offsets inside the static table blob (device global, copied to shared memory per CTA)
enum : int {
  OFF_GAMMA = 0,      // u8[256]    gamma LUT                       gamma_correction.cpp:35-42
  OFF_ENH = 256,      // u8[3][256] enhancer H,S,V gain LUTs        color_enhancer.cpp:42
  OFF_INVG = 1024,    // u8[4096]   sRGBInvGammaTab_b
  OFF_SRGBG = 5120,   // u16[256]   sRGBGammaTab_b
  OFF_LABC = 5632,    // u16[2048]  LabCbrtTab_b (2041 used)
  OFF_YF = 9728,      // u32[256]   LabToYF_b packed (ify << 16) | y
  OFF_SDIV = 10752,   // i32[256]
  OFF_HDIV = 11776,   // i32[256]
  TABLE_BYTES = 12800
};

enum : int { SRC_BAYER = 0, SRC_BGR = 1, SRC_RGB = 2 };

struct FrameParams {
  const uint8_t* in;          // n_frames x rows x in_pitch
  uint8_t* out;               // n_frames x orows x out_pitch (BGR8)
  long long in_frame_stride;  // bytes
  long long out_frame_stride; // bytes
  int in_pitch, out_pitch;    // bytes per row
  int rows, cols;             // input frame
  int orows, ocols;           // after flip
  int n_frames;
  int cfa;                    // CFA_* (frame_math.cuh)
  int angle;                  // 0 / 90 / 180 / 270
  int src;                    // SRC_*
  const uint8_t* tables;      // static blob (TABLE_BYTES)
  const uint8_t* wb;          // n_frames x 768 per-frame white-balance LUTs (B,G,R) or null
  const float* vig;           // vignetting quadrant or null
  int vig_pitch;              // floats per quadrant row
  ChainConsts k;
  unsigned long long* stats;  // n_frames x 8 (stats kernel only)
};

struct RemapParams {
  const uint8_t* src;  // n_frames x rows x pitch (CH channels)
  uint8_t* dst;        // n_frames x orows x dpitch
  long long src_frame_stride, dst_frame_stride;
  int rows, cols, pitch;
  int orows, ocols, dpitch;
  int n_frames;
  const float2* map;   // orows x ocols (x, y)
};

// All launchers enqueue on `stream`, return the CUDA error of the launch, and add the number of
// kernels launched to *launches.
cudaError_t launch_fused(uint32_t stages, const FrameParams& p, int sm_count, cudaStream_t stream, int* launches);
cudaError_t launch_pca_stats(const FrameParams& p, int sm_count, cudaStream_t stream, int* launches);
cudaError_t launch_pca_lut(const unsigned long long* stats, uint8_t* wb, float* coeff_out, int n_frames,
                           cudaStream_t stream, int* launches);
cudaError_t launch_gain_lut(const float* gains_bgr, uint8_t* wb, int n_frames, cudaStream_t stream, int* launches);
cudaError_t launch_remap(int channels, const RemapParams& p, cudaStream_t stream, int* launches);

}  // namespace rip
