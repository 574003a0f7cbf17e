/*
 * ifb_oracle.h -- CPU ORACLE for the imageflow BGRA resample hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it.
 * The product library (imageflow_b200/csrc) never links, imports or calls it.
 *
 * It restates, in plain C, the algorithm of the reference path
 * (paths relative to /root/reference):
 *   - imageflow_core/src/graphics/weights.rs      (filters, populate_weights)   PINNED by golden tables
 *   - imageflow_core/src/graphics/color.rs        (ColorContext, uchar_clamp_ff) PINNED by KATs
 *   - imageflow_core/src/graphics/lut.rs          (linear_to_srgb_lut)          PINNED (16384/16384)
 *   - imageflow_core/src/graphics/scaling.rs      (dispatch, composite :254-287, A=255 fix-up :227-232)
 *   - imageflow_core/src/graphics/color_matrix.rs (5x5 matrix on sRGB bytes)
 *   - imageflow_core/src/graphics/blend.rs        (apply_matte)
 *   - imageflow_core/src/graphics/transpose.rs, flip.rs (data movement only: exact by definition)
 *
 * PARITY STATUS.  Weights, transfer functions, colour matrix, canvas composite
 * and apply_matte are pinned against the reference's own golden vectors / KATs
 * (tests/test_oracle_golden.py).  The fp32 tap arithmetic of the separable
 * filter itself lives in the un-vendored crate `zenresize` 0.3.1
 * (Cargo.lock:4103-4106) whose source is not available: for that part the
 * oracle is a first-principles statement (documented in DESIGN.md) and is
 * "PARITY UNPINNED" -- the reference's own acceptance band for it is
 * Tolerance::off_by_one() (tests/integration/visuals/scaling.rs:18).
 */
#ifndef IFB_ORACLE_H
#define IFB_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes (mirror WeightsError, weights.rs:494-504, and ErrorKind used in scaling.rs) */
enum {
    IFO_OK = 0,
    IFO_ERR_INVALID_ARGUMENT = 1,      /* scaling.rs:24-29,38-40 */
    IFO_ERR_NOT_IMPLEMENTED = 2,       /* scaling.rs:43-48 */
    IFO_ERR_INVALID_STATE = 3,         /* scaling.rs:145,191,202,240 */
    IFO_ERR_TOTAL_WEIGHT_ZERO = 10,    /* weights.rs:755-757 */
    IFO_ERR_SOURCE_COUNT_TOO_LARGE = 11, /* weights.rs:719-722 */
    IFO_ERR_NO_PIXEL_INPUTS = 12,      /* weights.rs:613-615 */
    IFO_ERR_BAD_FILTER = 13,
    IFO_ERR_CAPACITY = 14
};

enum { IFO_LOBE_NATURAL = 0, IFO_LOBE_EXACT = 1, IFO_LOBE_SHARPEN_PERCENT = 2 }; /* weights.rs:16-23 */
enum { IFO_COMPOSE_REPLACE_SELF = 0, IFO_COMPOSE_BLEND_WITH_SELF = 1, IFO_COMPOSE_BLEND_WITH_MATTE = 2 };

/* Same field order/meaning as ifb200_resample_desc (include/ifb200.h); declared
 * independently so the oracle has no dependency on the product headers. */
typedef struct {
    const uint8_t* in;  uint32_t in_w, in_h, in_stride;
    uint8_t* canvas;    uint32_t cv_w, cv_h, cv_stride;
    uint32_t x, y, w, h;
    int32_t  filter;
    float    sharpen_percent;
    int32_t  linear;
    int32_t  alpha_meaningful;
    int32_t  compose;
    uint8_t  matte_bgra[4];
    const float* color_matrix;
} ifo_desc;

/* weights.rs:681-788.  left/right: out_size entries.  offsets: out_size+1 entries
 * (prefix offsets into weights[]).  Returns IFO_OK or an IFO_ERR_*. */
int ifo_weights(int filter, double kernel_width_scale, int lobe_mode, float lobe_value,
                uint32_t out_size, uint32_t in_size,
                uint32_t* left, uint32_t* right, uint32_t* offsets,
                float* weights, size_t weights_cap);
/* weights.rs:333-350 */
double ifo_percent_negative_weight(int filter, double kernel_width_scale);
/* raw filter function value (weights.rs:352-458) */
double ifo_filter_eval(int filter, double kernel_width_scale, double x);

/* color.rs:23-48,85-91: byte -> working float space. linear!=0 => LinearRGB else StandardRGB */
void  ifo_byte_to_float_table(int linear, float out[256]);
/* lut.rs:4-8 + the f64 generator in tests/integration/color_conversion.rs:381-388 */
void  ifo_linear_to_srgb_table(uint8_t out[16384]);
uint8_t ifo_floatspace_to_srgb(int linear, float v);   /* color.rs:59-69 */
uint8_t ifo_uchar_clamp_ff(float v);                   /* color.rs:101-108 */

/* scaling.rs:19-90 (+ optional colour matrix applied to the destination rect afterwards) */
int ifo_scale_and_render(const ifo_desc* d);
/* same, many images, OpenMP over images (CPU baseline harness). returns first error */
int ifo_scale_and_render_batch(const ifo_desc* d, size_t n, int threads);
/* the two intermediate stages, for parity debugging: H-pass result (in_h x out_w x 4 floats; rows no output window
 * touches stay as the caller left them) and the final premultiplied float pixel rows (out_h x out_w x 4 floats).
 * either may be NULL */
int ifo_resample_stages(const ifo_desc* d, float* hpass, float* final_rows);

/* color_matrix.rs:5-28; m is row-major [5][5] */
void ifo_color_matrix(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, const float* m);
/* flow/nodes/color.rs:86-225 presets. which: 0 sepia,1 grayscale_ntsc,2 grayscale_flat,3 grayscale_bt709,
 * 4 grayscale_ry,5 invert,6 alpha(p),7 contrast(p),8 brightness(p),9 saturation(p) */
int  ifo_color_filter_matrix(int which, float p, float out[25]);
/* blend.rs:6-59 */
void ifo_apply_matte(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, const uint8_t matte_bgra[4], int alpha_meaningful);

/* graphics/transpose.rs:95-121: to[x][y] = from[y][x], w x h BGRA8 words; strides in bytes */
void ifo_transpose(const uint8_t* from, uint32_t from_stride, uint32_t w, uint32_t h, uint8_t* to, uint32_t to_stride);
/* graphics/flip.rs:10-22 and :25-39, in place */
void ifo_flip_vertical(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride);
void ifo_flip_horizontal(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride);

/* flow/nodes/white_balance.rs:14-121 + graphics/histogram.rs:7-20, in place; threshold < 0 = None (0.006f32).
 * maps_out (may be NULL) receives the three 256-entry byte maps in R, G, B order. */
void ifo_white_balance(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, float threshold, uint8_t* maps_out);

/* graphics/whitespace.rs:284-331 detect_content, with everything under it (get_search_rect :220-281, check_region :333-421,
 * approximate_grayscale :426-523 for Bgra32 (alpha meaningful) / Bgr32, sobel_scharr_detect :525-634), replayed in the
 * reference's order: which windows are evaluated depends on the bounding box found so far.
 * rect_out = {x1, y1, x2, y2}.  Returns 0, or IFO_ERR_INVALID_ARGUMENT for empty bitmaps.
 * centres_out (optional): number of window-interior pixels the scan evaluated. */
int ifo_detect_content(const uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, int alpha_meaningful, uint32_t threshold,
                       uint32_t rect_out[4], uint64_t* centres_out);
/* The window-independent part of sobel_scharr_detect: one code per pixel as the centre of its 3x3 neighbourhood.
 * 0xFF: border pixel, or Scharr value <= threshold.  Otherwise bits 1:0 = local_min_x (0..2), 3:2 = local_max_x - 1,
 * 5:4 = local_min_y, 7:6 = local_max_y - 1 (whitespace.rs:567-613), relative to (x - 1, y - 1).  map: w * h bytes. */
void ifo_whitespace_codes(const uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, int alpha_meaningful, uint32_t threshold, uint8_t* map);
/* detect_content with the window walk replayed over such a map (the threshold is already in the codes): same result,
 * same number of evaluated centres as ifo_detect_content on the bitmap the map was made from. */
int ifo_detect_content_from_codes(const uint8_t* codes, uint32_t w, uint32_t h, uint32_t rect_out[4], uint64_t* centres_out);

/* imageflow_b200/synth.py noise frames, generated at memory speed (bench.py's CPU arm) */
void ifo_synth_noise(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, uint32_t seed, int alpha_mixed);

int ifo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
