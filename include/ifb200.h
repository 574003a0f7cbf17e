/*
 * ifb200.h -- C ABI of the B200-native resample engine (libifb200.so).
 *
 * This is the drop-in boundary for ONE hot path of imazen/imageflow: the BGRA8
 * resample / composite / colour-matrix path behind
 *     imageflow_core::graphics::scaling::scale_and_render        (graphics/scaling.rs:19-90)
 *     imageflow_core::graphics::color_matrix::window_bgra32_apply_color_matrix (graphics/color_matrix.rs:5-28)
 * which the flow-graph executors DrawImageDef::render (flow/nodes/scale_render.rs:304-313) and
 * ColorMatrixSrgbMutDef::mutate (flow/nodes/color.rs:26-28) call.  (All paths relative to the
 * reference checkout; INTEGRATION.md shows the Rust `extern "C"` block that binds these symbols.)
 *
 * Plain pointers and sizes only; no C++/torch types.  Every function is thread-safe, never throws
 * or aborts across the boundary, never retains caller pointers after it returns (batch jobs: after
 * ifb200_batch_sync returns).  There is NO CPU fallback: without a usable CUDA device every compute
 * entry point fails with IFB200_ERR_NO_DEVICE.
 */
#ifndef IFB200_H
#define IFB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IFB200_ABI_VERSION_MAJOR 1
#define IFB200_ABI_VERSION_MINOR 2

/* Error codes.  1..3 map onto the imageflow ErrorKind values raised by scale_and_render
 * (scaling.rs:24-48,145,191,202,240); 10..12 onto WeightsError (weights.rs:494-504). */
enum ifb200_status {
    IFB200_OK = 0,
    IFB200_ERR_INVALID_ARGUMENT = 1,      /* ErrorKind::InvalidArgument  (scaling.rs:24-29,38-40) */
    IFB200_ERR_NOT_IMPLEMENTED = 2,       /* ErrorKind::MethodNotImplemented (scaling.rs:43-48)   */
    IFB200_ERR_INVALID_STATE = 3,         /* ErrorKind::InvalidState (scaling.rs:145,191,202,240) */
    IFB200_ERR_TOTAL_WEIGHT_ZERO = 10,    /* WeightsError::TotalWeightZero (weights.rs:755-757)   */
    IFB200_ERR_SOURCE_COUNT_TOO_LARGE = 11,
    IFB200_ERR_NO_PIXEL_INPUTS = 12,
    IFB200_ERR_BAD_FILTER = 13,
    IFB200_ERR_CAPACITY = 14,
    IFB200_ERR_NO_DEVICE = 20,            /* no CUDA device / driver: there is no CPU fallback    */
    IFB200_ERR_CUDA = 21,                 /* a CUDA runtime call failed; message has the detail   */
    IFB200_ERR_OUT_OF_MEMORY = 22
};

/* weights.rs:45-78 -- `Filter` repr(C) discriminants, used verbatim. */
enum ifb200_filter {
    IFB200_FILTER_ROBIDOUX_FAST = 1, IFB200_FILTER_ROBIDOUX = 2, IFB200_FILTER_ROBIDOUX_SHARP = 3,
    IFB200_FILTER_GINSENG = 4, IFB200_FILTER_GINSENG_SHARP = 5, IFB200_FILTER_LANCZOS = 6,
    IFB200_FILTER_LANCZOS_SHARP = 7, IFB200_FILTER_LANCZOS2 = 8, IFB200_FILTER_LANCZOS2_SHARP = 9,
    IFB200_FILTER_CUBIC_FAST = 10, IFB200_FILTER_CUBIC = 11, IFB200_FILTER_CUBIC_SHARP = 12,
    IFB200_FILTER_CATMULL_ROM = 13, IFB200_FILTER_MITCHELL = 14, IFB200_FILTER_CUBIC_BSPLINE = 15,
    IFB200_FILTER_HERMITE = 16, IFB200_FILTER_JINC = 17, IFB200_FILTER_RAW_LANCZOS3 = 18,
    IFB200_FILTER_RAW_LANCZOS3_SHARP = 19, IFB200_FILTER_RAW_LANCZOS2 = 20, IFB200_FILTER_RAW_LANCZOS2_SHARP = 21,
    IFB200_FILTER_TRIANGLE = 22, IFB200_FILTER_LINEAR = 23, IFB200_FILTER_BOX = 24,
    IFB200_FILTER_CATMULL_ROM_FAST = 25, IFB200_FILTER_CATMULL_ROM_FAST_SHARP = 26, IFB200_FILTER_FASTEST = 27,
    IFB200_FILTER_MITCHELL_FAST = 28, IFB200_FILTER_NCUBIC = 29, IFB200_FILTER_NCUBIC_SHARP = 30,
    IFB200_FILTER_LEGACY_IDCT = 31
};

/* graphics/bitmaps.rs:156-160 BitmapCompositing */
enum ifb200_compose { IFB200_REPLACE_SELF = 0, IFB200_BLEND_WITH_SELF = 1, IFB200_BLEND_WITH_MATTE = 2 };
/* weights.rs:16-23 LobeRatio */
enum ifb200_lobe { IFB200_LOBE_NATURAL = 0, IFB200_LOBE_EXACT = 1, IFB200_LOBE_SHARPEN_PERCENT = 2 };

/* One scale_and_render call: (input window, canvas window, ScaleAndRenderParams) -- scaling.rs:9-23. */
typedef struct ifb200_resample_desc {
    const uint8_t* in;  uint32_t in_w, in_h, in_stride;   /* BGRA8 sRGB, straight alpha; stride in bytes (bitmaps.rs:712-740) */
    uint8_t* canvas;    uint32_t cv_w, cv_h, cv_stride;   /* the canvas BEFORE cropping to (x,y,w,h) (scaling.rs:30-41)      */
    uint32_t x, y, w, h;                                   /* ScaleAndRenderParams.x/y/w/h                                    */
    int32_t  filter;                                       /* enum ifb200_filter (ScaleAndRenderParams.interpolation_filter)   */
    float    sharpen_percent;                              /* sharpen_percent_goal; <= 0 means off (scaling.rs:104-106)       */
    int32_t  linear;                                       /* 1 = WorkingFloatspace::LinearRGB, 0 = StandardRGB (scaling.rs:52) */
    int32_t  alpha_meaningful;                             /* input.info().alpha_meaningful() (scaling.rs:51)                 */
    int32_t  compose;                                      /* enum ifb200_compose = canvas.info().compose() (scaling.rs:50)   */
    uint8_t  matte_bgra[4];                                /* BlendWithMatte colour as to_bgra8() (scaling.rs:67)            */
    const float* color_matrix;                             /* optional fused ColorMatrixSrgb (row-major [5][5]) applied to the
                                                              destination rect after the render; NULL = none (color_matrix.rs) */
} ifb200_resample_desc;

/* ---- library / device ---------------------------------------------------------------------- */
uint32_t    ifb200_abi_version(void);                      /* (major << 16) | minor */
/* Device of the host-buffer (drop-in) calls made by threads that have not made one yet (each calling thread keeps its own context on
 * the device that was selected when it made its first call).  Default: the IFB200_DEVICE environment variable, else device 0. */
int         ifb200_set_dropin_device(int device);
const char* ifb200_status_name(int status);
int         ifb200_device_count(void);                     /* 0 when no CUDA device / driver is usable */

/* ---- host-side specification helpers (no GPU needed) ----------------------------------------
 * populate_weights (weights.rs:681-788) for one axis. left/right hold out_size entries, offsets
 * out_size+1 (prefix offsets into weights[]).  kernel_width_scale = set_kernel_width_scale factor
 * (weights.rs:156-158), lobe_* = LobeRatio (weights.rs:16-40). */
int ifb200_weights(int filter, double kernel_width_scale, int lobe_mode, float lobe_value,
                   uint32_t out_size, uint32_t in_size,
                   uint32_t* left, uint32_t* right, uint32_t* offsets,
                   float* weights, size_t weights_cap);
/* ColorContext::byte_to_float (color.rs:23-48) and LINEAR_TO_SRGB_LUT (lut.rs:14) as uploaded to the device */
void ifb200_byte_to_float_table(int linear, float out[256]);
void ifb200_linear_to_srgb_table(uint8_t out[16384]);
/* ColorFilterSrgb presets (flow/nodes/color.rs:86-225): 0 sepia, 1 grayscale_ntsc, 2 grayscale_flat,
 * 3 grayscale_bt709, 4 grayscale_ry, 5 invert, 6 alpha(p), 7 contrast(p), 8 brightness(p), 9 saturation(p) */
int  ifb200_color_filter_matrix(int which, float p, float out[25]);
/* detect_content with a HOST bitmap (drop-in for graphics/whitespace.rs:284); and its host half alone: the window walk over a
 * w*h code map (layout: imageflow_b200/csrc/ifb_whitespace.h), no CUDA call -- `centres` (optional) = pixels visited. */
int  ifb200_detect_content_bgra8(const uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, int alpha_meaningful, uint32_t threshold,
                                 uint32_t rect[4], char* err, size_t err_cap);
int  ifb200_detect_content_from_codes(const uint8_t* codes, uint32_t w, uint32_t h, uint32_t rect[4], uint64_t* centres);
/* Host-side cost of preparing kernel tables (benchmarks of mixed workloads, where every geometry is new): builds the plans of
   descs[0..n) -- only in_w, in_h, w, h, filter, sharpen_percent are read -- on `threads` host threads and discards them.
   No CUDA call.  *table_hash (optional) = a hash of every table built, independent of the thread count. */
int  ifb200_plan_probe(const ifb200_resample_desc* descs, size_t n, int threads, double* seconds, uint64_t* table_bytes,
                       uint64_t* table_hash, char* err, size_t err_cap);

/* The streaming ring kernel's host tables for one geometry (only in_w, in_h, w, h, filter, sharpen_percent are read), as they
   would be uploaded (alpha_meaningful selects the 3- or 4-channel kernel variant's strip width): call with buf = NULL for the size.  info->ok == 0: the geometry does not run on the ring kernel.
   Layout: imageflow_b200/csrc/ifb_hv_kernel.cuh (HvStripDev, HvBandDev, HvPlanDev).  No CUDA call. */
typedef struct ifb200_hv_plan_info {
    int32_t ok, av, n_strips, n_bands, cap_px /* weight records per strip in hw; hdone has cap_px + 64 bytes per strip */, avp;
    uint64_t o_strips, o_hw, o_hdone, o_vw, o_vdone, o_bands, total;
} ifb200_hv_plan_info;
int  ifb200_hv_plan_tables(const ifb200_resample_desc* d, int strip_cols, int n_band_pairs, ifb200_hv_plan_info* info, uint8_t* buf, size_t cap,
                           char* err, size_t err_cap);

/* ---- drop-in calls: HOST buffers, synchronous (what the Rust adapter calls) ------------------
 * Replaces the bodies of scaling.rs:93-251 (resize_to_canvas / resize_with_matte /
 * resize_and_composite).  err (may be NULL) receives a NUL-terminated message on failure. */
int ifb200_scale_and_render(const ifb200_resample_desc* desc, char* err, size_t err_cap);
/* The same for n independent calls (e.g. the DrawImageExact nodes of one graph, or export_4_sizes): identical
 * results, but uploads, kernels and downloads of consecutive jobs are pipelined on several CUDA streams.
 * Jobs must not alias each other's canvases.  Argument errors are reported before any job runs. */
int ifb200_scale_and_render_many(const ifb200_resample_desc* descs, size_t n, char* err, size_t err_cap);
/* Replaces color_matrix.rs:5-28 (in place). */
int ifb200_color_matrix_bgra8(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, const float m[25],
                              char* err, size_t err_cap);

/* Replaces Bitmap::apply_matte (graphics/blend.rs:6-59), the encoder-side flatten over a solid colour (in place,
 * linear light).  No-op unless alpha is meaningful (blend.rs:10-13).  [SURVEY.md section 8(f), item 2] */
int ifb200_apply_matte_bgra8(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, const uint8_t matte_bgra[4],
                             int alpha_meaningful, char* err, size_t err_cap);

/* Replaces graphics::transpose::bitmap_window_transpose (graphics/transpose.rs:95-121; checks of :46-79):
 * to[x][y] = from[y][x] for a w x h BGRA8 window; `to` is h pixels wide and w rows tall.  Strides in bytes.
 * [SURVEY.md section 8(f), item 3] */
int ifb200_transpose_bgra8(const uint8_t* from, uint32_t from_stride, uint32_t w, uint32_t h, uint8_t* to, uint32_t to_stride,
                           char* err, size_t err_cap);
/* Replace flow_bitmap_bgra_flip_vertical_safe / flow_bitmap_bgra_flip_horizontal_safe (graphics/flip.rs:10-22, 25-39), in place. */
int ifb200_flip_vertical_bgra8(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, char* err, size_t err_cap);
int ifb200_flip_horizontal_bgra8(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, char* err, size_t err_cap);

/* Replaces the work of WhiteBalanceSrgbMutDef::mutate (flow/nodes/white_balance.rs:93-121): per-channel histograms
 * (graphics/histogram.rs:7-20), area thresholds and byte maps (white_balance.rs:14-48), in-place remap of B, G, R (:50-67).
 * threshold < 0 = None (0.006).  [SURVEY.md section 8(f), item 4] */
int ifb200_white_balance_srgb_bgra8(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, float threshold, char* err, size_t err_cap);

/* Replaces flow_scale_spatial[_srgb]_{n}x{n} (c_components/lib/codecs_jpeg_idct_fast.c:302-4428; libjpeg hook codec_jpeg_wrapper.c:232-340,
 * enabled by JpegDownscaleHints, mozjpeg_decoder.rs:593-611) for a whole plane of samples: every 8x8 block of `in` (blocks_x x blocks_y
 * blocks, row stride in_stride bytes) becomes the n x n block at (bx*n, by*n) of `out`; n = 1..7; srgb != 0 = the linear-light variants.
 * Bit-identical to the reference functions (integer arithmetic, the reference's own weight and table literals).
 * [SURVEY.md section 8(f), item 1] */
int ifb200_block_scale_u8(const uint8_t* in, uint32_t in_stride, uint32_t blocks_x, uint32_t blocks_y, uint8_t* out, uint32_t out_stride,
                          int n, int srgb, char* err, size_t err_cap);

/* ---- device-resident batch API (the metric path; not in the reference) -----------------------
 * descs[i].in / .canvas are DEVICE pointers on the batch's device; color_matrix stays a HOST pointer.
 * enqueue is asynchronous on `cuda_stream`, a cudaStream_t with the usual CUDA meaning (NULL = the legacy
 * default stream); pass IFB200_STREAM_OWN to use the batch's private non-blocking stream, which is the
 * stream ifb200_batch_sync waits on.  The jobs of ONE call must be independent of each other (none may read what
 * another writes): they may run concurrently.  Calls on the same stream are ordered as usual. */
#define IFB200_STREAM_OWN ((void*)(intptr_t)-1)
typedef struct ifb200_batch ifb200_batch;
int  ifb200_batch_create(int device, ifb200_batch** out, char* err, size_t err_cap);
int  ifb200_batch_enqueue(ifb200_batch* b, const ifb200_resample_desc* descs, size_t n, void* cuda_stream,
                          char* err, size_t err_cap);
int  ifb200_batch_color_matrix(ifb200_batch* b, uint8_t* dev_px, uint32_t w, uint32_t h, uint32_t stride,
                               const float m[25], void* cuda_stream, char* err, size_t err_cap);
int  ifb200_batch_apply_matte(ifb200_batch* b, uint8_t* dev_px, uint32_t w, uint32_t h, uint32_t stride,
                              const uint8_t matte_bgra[4], int alpha_meaningful, void* cuda_stream, char* err, size_t err_cap);
int  ifb200_batch_transpose(ifb200_batch* b, const uint8_t* dev_from, uint32_t from_stride, uint32_t w, uint32_t h,
                            uint8_t* dev_to, uint32_t to_stride, void* cuda_stream, char* err, size_t err_cap);
int  ifb200_batch_flip_vertical(ifb200_batch* b, uint8_t* dev_px, uint32_t w, uint32_t h, uint32_t stride,
                                void* cuda_stream, char* err, size_t err_cap);
int  ifb200_batch_flip_horizontal(ifb200_batch* b, uint8_t* dev_px, uint32_t w, uint32_t h, uint32_t stride,
                                  void* cuda_stream, char* err, size_t err_cap);
int  ifb200_batch_block_scale(ifb200_batch* b, const uint8_t* dev_in, uint32_t in_stride, uint32_t blocks_x, uint32_t blocks_y,
                              uint8_t* dev_out, uint32_t out_stride, int n, int srgb, void* cuda_stream, char* err, size_t err_cap);
int  ifb200_batch_white_balance(ifb200_batch* b, uint8_t* dev_px, uint32_t w, uint32_t h, uint32_t stride, float threshold,
                                void* cuda_stream, char* err, size_t err_cap);
/* graphics/whitespace.rs:284-331 detect_content(&BitmapWindowMut<u8>, threshold) -> RectCorners{x1, y1, x2, y2}: the rectangle the
 * reference's whitespace scan finds (trim / `trim.threshold`).  The per-pixel arithmetic (approximate_grayscale :426-523, Scharr
 * and the local edge box :525-613) runs on the GPU for all pixels; the reference's order-dependent window walk (:333-421) is
 * replayed on the host over that one-byte-per-pixel map, so the result is the reference's, not the bounding box of all edges.
 * alpha_meaningful selects the Bgra32 (1) or Bgr32 (0) grayscale.  Synchronous (a rectangle comes back). */
int  ifb200_batch_detect_content(ifb200_batch* b, const uint8_t* dev_px, uint32_t w, uint32_t h, uint32_t stride, int alpha_meaningful,
                                 uint32_t threshold, uint32_t rect[4], void* cuda_stream, char* err, size_t err_cap);
/* The GPU half of detect_content alone: the one-byte-per-pixel code map (layout: imageflow_b200/csrc/ifb_whitespace.h; what
 * sobel_scharr_detect :525-613 derives per 3x3 neighbourhood) of a DEVICE bitmap into a DEVICE buffer of w*h bytes, row pitch w.
 * Asynchronous on `cuda_stream`; ifb200_detect_content_from_codes() walks a host copy of it. */
int  ifb200_batch_whitespace_codes(ifb200_batch* b, const uint8_t* dev_px, uint32_t w, uint32_t h, uint32_t stride, int alpha_meaningful,
                                   uint32_t threshold, uint8_t* dev_codes, void* cuda_stream, char* err, size_t err_cap);
int  ifb200_batch_sync(ifb200_batch* b, char* err, size_t err_cap);
void ifb200_batch_destroy(ifb200_batch* b);
/* knobs / introspection (benchmarks, tests) */
enum ifb200_option {
    IFB200_OPT_FORCE_GENERIC = 1,      /* 1: always use the two-kernel generic path (parity cross-check)  */
    IFB200_OPT_STRIP_COLUMNS = 2,      /* ring kernel: widest strip of output columns one warp works on: a multiple of 16 up to 128
                                          (default 64; clipped to what the kernel variant holds: 64, or 50 / 40 at ring depth 6) */
    IFB200_OPT_MIN_ITEMS = 3           /* ring kernel: split images into row bands until a launch has at least this many warp work
                                          items.  0 (default): as many as the device has warps, and launches that still cannot
                                          fill the device get narrower strips; > 0: exactly the strip width asked for */
};
int      ifb200_batch_set_option(ifb200_batch* b, int option, int64_t value);
uint64_t ifb200_batch_kernel_launches(const ifb200_batch* b);   /* total kernels launched so far   */
/* Where the calling thread has spent its enqueue calls so far, in seconds (inclusive: [1..4] are parts of [0]; [2] and [3] parts of [4]):
   out[0] whole enqueue calls, [1] building plans, [2] pinned staging slots (search + cudaMallocHost), [3] memcpy of tables into
   them, [4] table uploads (allocation + [2] + [3] + cudaMemcpyAsync + events); then counts: [5] table uploads, [6] bytes staged,
   [7] cudaMallocHost calls.  Fills min(n, 8) entries, returns 8. */
int      ifb200_batch_host_profile(const ifb200_batch* b, double* out, int n);
uint64_t ifb200_batch_fused_jobs(const ifb200_batch* b);        /* jobs that took the streaming ring kernel */
uint64_t ifb200_batch_generic_jobs(const ifb200_batch* b);      /* jobs that took the generic pair */
uint64_t ifb200_batch_tile_jobs(const ifb200_batch* b);         /* jobs that took the tile kernel (up-scales, 1:1) */
/* 1 when the streaming ring kernel (the fast path for down-scales) can run on this device/driver; 0 when jobs that would take it
 * fall back to the tile kernel or the generic pair -- `why` (may be NULL) then says what is missing. */
int      ifb200_batch_ring_status(const ifb200_batch* b, char* why, size_t why_cap);

#ifdef __cplusplus
}
#endif
#endif /* IFB200_H */
