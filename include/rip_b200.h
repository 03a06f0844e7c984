/* rip_b200.h -- C ABI of the B200-native RAW image pipeline (librip_b200.so).
 *
 * Drop-in boundary for the hot path of leggedrobotics/raw_image_pipeline: everything the
 * reference's `raw_image_pipeline::RawImagePipeline` class
 * (raw_image_pipeline/include/raw_image_pipeline/raw_image_pipeline.hpp:36-137) does per frame
 * -- Debayer -> Flip -> WhiteBalance -> ColorCalibration -> Gamma -> Vignetting ->
 * ColorEnhancer -> Undistortion (hpp:143-172) -- computed by hand-written sm_100a CUDA kernels,
 * with results bit-identical to the reference's CPU/OpenCV path.
 *
 * Plain C: opaque handle, pointers and sizes, int status returns, no exceptions, no C++ or
 * torch types.  The header-only C++ class in include/raw_image_pipeline/raw_image_pipeline.hpp
 * and the Python class raw_image_pipeline_b200.RawImagePipeline are thin wrappers over it.
 *
 * Not thread-safe per handle (like the reference: one instance per camera stream).
 * There is NO CPU fallback: every call that touches pixels needs a CUDA device and fails
 * with RIP_ERR_CUDA otherwise.
 */
#ifndef RIP_B200_H_
#define RIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RIP_API __attribute__((visibility("default")))
#else
#define RIP_API
#endif

typedef struct rip_pipeline rip_pipeline;

enum {
  RIP_OK = 0,
  RIP_ERR_INVALID_ARGUMENT = 1, /* what the reference throws std::invalid_argument for */
  RIP_ERR_CUDA = 2,             /* no device / CUDA runtime failure */
  RIP_ERR_IO = 3,               /* file missing or unparsable where the reference would crash */
  RIP_ERR_BUFFER_TOO_SMALL = 4,
  RIP_ERR_UNKNOWN_KEY = 5,
  RIP_ERR_UNSUPPORTED = 6       /* in the reference API but outside this library's scope (xphoto WB) */
};

/* which cached image rip_get_image returns (raw_image_pipeline.cpp:222-236) */
enum {
  RIP_IMAGE_DIST_DEBAYERED = 0, /* getDistDebayeredImage(): after debayer+flip          */
  RIP_IMAGE_DIST_COLOR = 1,     /* getDistColorImage(): input of the undistortion stage */
  RIP_IMAGE_PROCESSED = 2,      /* getProcessedImage(): final output                    */
  RIP_IMAGE_RECT_MASK = 3       /* getRectMask(): never written by the reference -> empty */
};

/* ---- life cycle ------------------------------------------------------------------------
 * rip_create         == RawImagePipeline(bool, params, calib, color_calib)  (raw_image_pipeline.cpp:23-40)
 *                       NULL or "" selects the reference's default for that argument
 *                       (empty calibration path = no camera calibration loaded).
 * rip_create_default == RawImagePipeline(bool use_gpu)                      (raw_image_pipeline.cpp:16-21)
 * `use_gpu` is kept for source compatibility: this library always computes on the GPU and
 * always reproduces the reference's CPU-path results. */
RIP_API int rip_create(int use_gpu, const char* params_path, const char* calibration_path,
               const char* color_calibration_path, rip_pipeline** out);
RIP_API int rip_create_default(int use_gpu, rip_pipeline** out);
RIP_API void rip_destroy(rip_pipeline* p);
/* message of the last failed call on `p` (or of the last failed rip_create* when p == NULL) */
RIP_API const char* rip_last_error(const rip_pipeline* p);

/* ---- loaders (raw_image_pipeline.hpp:53-56, 59) ------------------------------------------ */
RIP_API int rip_load_params(rip_pipeline* p, const char* path);             /* loadParams            */
RIP_API int rip_load_camera_calibration(rip_pipeline* p, const char* path); /* loadCameraCalibration */
RIP_API int rip_load_color_calibration(rip_pipeline* p, const char* path);  /* loadColorCalibration  */
RIP_API int rip_init_undistortion(rip_pipeline* p);                         /* initUndistortion      */
RIP_API int rip_reset_white_balance_temporal_consistency(rip_pipeline* p);  /* resetWhiteBalanceTemporalConsistency */

/* ---- keyed setters / getters --------------------------------------------------------------
 * One key per reference setter/getter (raw_image_pipeline.hpp:61-132); keys follow the YAML
 * sections of raw_image_pipeline.cpp:58-159.  See INTEGRATION.md for the full table.
 * Input kinds (debayer.cpp:45-79): bayer_{rggb,grbg,gbrg,bggr}8 (1 channel -> BGR8), rgb8 / any other 3-channel
 * encoding (bgr8 ...), and 1-channel non-Bayer images (mono8 ...: only flip, gamma and undistortion apply, like the
 * reference's modules that skip images without 3 channels; vignetting rejects them).
 *   bool    gpu, debug, <module>/enabled with module in {debayer, flip, white_balance,
 *           color_calibration, gamma_correction, vignetting_correction, color_enhancer,
 *           undistortion}, white_balance/temporal_consistency
 *   int     flip/angle; (get) dist/image_height, dist/image_width, rect/image_height,
 *           rect/image_width, stats/kernel_launches, stats/ccc_u, stats/ccc_v (CCC arg-max of the last rip_apply)
 *   double  white_balance/clipping_percentile, gamma_correction/k, color_enhancer/hue_gain,
 *           color_enhancer/saturation_gain, color_enhancer/value_gain (the reference's
 *           cross-wired setters are reproduced), undistortion/balance, undistortion/fov_scale
 *   string  debayer/encoding, white_balance/method, gamma_correction/method,
 *           undistortion/distortion_model; (get) dist/distortion_model, rect/distortion_model
 *   doubles white_balance/saturation_threshold[2] (bright, dark),
 *           color_calibration/matrix[9], color_calibration/bias[3],
 *           vignetting_correction/parameters[3] (scale, a2, a4),
 *           undistortion/image_size[2] (width, height), undistortion/new_image_size[2],
 *           undistortion/camera_matrix[9], undistortion/distortion_coefficients[4],
 *           undistortion/rectification_matrix[9], undistortion/projection_matrix[12];
 *           (get) dist|rect/camera_matrix[9], /distortion_coefficients[4],
 *           /rectification_matrix[9], /projection_matrix[12]; stats/pca_coefficients[4], stats/ccc_gains[3] (last
 *           rip_apply), stats/kernel_ms[8] (per-kernel CUDA-event totals and counts since the last query)
 * Development switches (bool): profile/kernel_events (CUDA events around every kernel launch),
 *   debug/force_generic_kernels (skip the TMA fast path), debug/force_float_map (undistortion reads the fp32 map
 *   instead of the packed fixed-point one), debug/force_gather_remap (undistortion gathers from global memory instead
 *   of the TMA-staged tile kernel); (int) debug/fused_kernel: 0 = the measured choice per stage set, 1 = tile kernel,
 *   2 = strip kernel.  None of them changes a single output byte.
 * Extension (bool, default off): apply/register_caller_buffers -- rip_apply page-locks (cudaHostRegister) the image and
 *   output buffers it is handed, once per buffer, and then copies straight from / into them; for callers that cycle through
 *   a fixed set of ordinary buffers.  The caller promises that such a buffer stays mapped while the pipeline lives.
 * Extension (bool): undistortion/rect_mask -- getRectMask() returns a real validity mask (u8, 255 where all four taps of
 *   the bilinear remap lie inside the source image) instead of the reference's never-written empty image.        */
RIP_API int rip_set_bool(rip_pipeline* p, const char* key, int value);
RIP_API int rip_set_int(rip_pipeline* p, const char* key, int value);
RIP_API int rip_set_double(rip_pipeline* p, const char* key, double value);
RIP_API int rip_set_string(rip_pipeline* p, const char* key, const char* value);
RIP_API int rip_set_doubles(rip_pipeline* p, const char* key, const double* values, int n);
RIP_API int rip_get_bool(rip_pipeline* p, const char* key, int* value);
RIP_API int rip_get_int(rip_pipeline* p, const char* key, int* value);
RIP_API int rip_get_double(rip_pipeline* p, const char* key, double* value);
RIP_API int rip_get_string(rip_pipeline* p, const char* key, char* value, size_t capacity);
RIP_API int rip_get_doubles(rip_pipeline* p, const char* key, double* values, int capacity, int* n);

/* ---- per-frame hot path -------------------------------------------------------------------
 * rip_apply == RawImagePipeline::apply(cv::Mat&, std::string&)  (raw_image_pipeline.cpp:190-205).
 * `data` is a host image (rows x cols x channels u8, `step` bytes per row); `encoding` is
 * in/out like the reference's std::string& ("bayer_rggb8" -> "bgr8").  The result is written
 * to `out` (host, tightly packed); its shape is returned through out_rows/out_cols/out_channels
 * (1ch -> 3ch on debayer, rows/cols swap for flip 90/270).  Call rip_output_shape first to size
 * `out`.  `out` may alias `data` when the shapes allow it (the copy-out happens last).        */
RIP_API int rip_output_shape(rip_pipeline* p, int rows, int cols, int channels, const char* encoding,
                     int* out_rows, int* out_cols, int* out_channels);
RIP_API int rip_apply(rip_pipeline* p, const uint8_t* data, int rows, int cols, int channels, size_t step,
              char* encoding, size_t encoding_capacity, uint8_t* out, size_t out_capacity,
              int* out_rows, int* out_cols, int* out_channels);
/* getDistDebayeredImage / getDistColorImage / getProcessedImage / getRectMask: deep copies of
 * the images cached by the last rip_apply.  rows = cols = 0 when empty.                       */
RIP_API int rip_get_image(rip_pipeline* p, int which, uint8_t* out, size_t out_capacity, int* rows,
                  int* cols, int* channels);

/* ---- batch entry points (no reference counterpart: how a B200 is kept busy) ----------------
 * Device-resident: `d_in` holds n frames (rows x cols x channels u8, tightly packed rows,
 * `in_frame_stride` bytes apart), `d_out` receives n output frames `out_frame_stride` bytes
 * apart, both in the memory of the pipeline's device.  `d_dist_color` (optional, may be NULL)
 * receives the pre-undistortion images when undistortion is enabled.  All work is enqueued on
 * `cuda_stream` (a cudaStream_t; NULL = legacy default stream) and the call returns without
 * synchronising.  Frames are independent (one camera frame each); white balance statistics
 * are per frame.  With CCC temporal consistency the frames are consecutive frames of ONE
 * stream.                                                                                    */
RIP_API int rip_apply_batch_device(rip_pipeline* p, const uint8_t* d_in, size_t in_frame_stride, int n_frames,
                           int rows, int cols, int channels, const char* encoding, uint8_t* d_out,
                           size_t out_frame_stride, uint8_t* d_dist_color, void* cuda_stream);
/* Host-to-host: same, from/to host memory (pinned memory gives full PCIe rate); copies and
 * kernels of consecutive chunks are overlapped on internal streams; returns when `out` is
 * complete.                                                                                  */
RIP_API int rip_apply_batch_host(rip_pipeline* p, const uint8_t* in, size_t in_frame_stride, int n_frames,
                         int rows, int cols, int channels, const char* encoding, uint8_t* out,
                         size_t out_frame_stride);

/* Host-to-host over several GPUs of one box (frames are independent, so they shard with no collective):
 * `handles[i]` is a pipeline bound to its own device (rip_set_device) and configured like the others.  Every pipeline
 * runs on its own host thread and claims the next chunk of <= 16 frames whenever one of its three copy/compute slots
 * is free, so GPUs behind a slower host link take fewer frames instead of holding the call back (the GPUs of one box
 * do not see the same host bandwidth: profiles/pcie_ceiling.json); returns when `out` is complete.  Every frame gets
 * the same bytes whichever pipeline takes it.  With CCC temporal consistency a batch is ONE camera stream and must go
 * to one pipeline: the call then fails with RIP_ERR_INVALID_ARGUMENT.  Returns the first failing pipeline's status (its
 * message is in that handle's rip_last_error).  The reference has no multi-GPU path (raw_image_pipeline.cpp:193-196
 * uploads to the one current device).                                                                             */
RIP_API int rip_apply_batch_host_multi(rip_pipeline* const* handles, int n_handles, const uint8_t* in, size_t in_frame_stride,
                               int n_frames, int rows, int cols, int channels, const char* encoding, uint8_t* out,
                               size_t out_frame_stride);

/* Page-locked host memory.  rip_apply copies a pageable caller image through the pipeline's own pinned staging
 * buffers (two host copies per frame); an image / output buffer that is page-locked -- from these two calls,
 * cudaHostAlloc or cudaHostRegister -- is read / written by the GPU's copy engines directly (the reference's
 * cv::cuda::GpuMat::upload / download, raw_image_pipeline.cpp:193-203, behave the same way for cv::cuda::HostMem).    */
RIP_API int rip_pinned_alloc(size_t bytes, void** ptr);
RIP_API int rip_pinned_free(void* ptr);

/* ---- inspection ---------------------------------------------------------------------------
 * Host-computed tables exactly as the kernels consume them (no GPU needed): "gamma_lut" (256 B,
 * gamma_correction.cpp:35-42), "enhancer_luts" (768 B), "vignetting_mask" (rows x cols fp32,
 * vignetting_correction.cpp:32-63), "undistortion_map" (dist_h x dist_w interleaved (x, y) fp32,
 * undistortion.cpp:212-220), "undistortion_packed_map" (u32 per output pixel for a rows x cols source image: two
 * int16 = cvRound(map * 32) relative to the pixel), "undistortion_tile_table" (4 x i32 per 128 x 24 output tile:
 * footprint origin x, y, flags, 0) and "undistortion_tile_map" (the packed map padded to whole tiles) -- the host-side
 * inputs of the tile undistortion kernel --, "ccc_response" (256 x 256 fp64, the convolution part of the last CCC
 * response; needs a processed frame).  `rows`/`cols`: the mask size for "vignetting_mask", the source image size for
 * the three packed-map tables.                                                                                 */
RIP_API int rip_debug_table(rip_pipeline* p, const char* name, int rows, int cols, void* out, size_t capacity,
                            size_t* bytes);

/* ---- device selection ---------------------------------------------------------------------- */
RIP_API int rip_device_count(void);                     /* < 0 on CUDA error */
RIP_API int rip_set_device(rip_pipeline* p, int device); /* before the first frame; default = current device */

#ifdef __cplusplus
}
#endif
#endif /* RIP_B200_H_ */
