// ifb_weights.h -- host-side contribution windows (product code).
//
// Restates imageflow_core/src/graphics/weights.rs of the reference:
//   filter presets  weights.rs:176-331      filter functions  weights.rs:352-458
//   bessj1          weights.rs:460-492      negative-lobe area weights.rs:333-350
//   LobeRatio       weights.rs:16-40        populate_weights  weights.rs:681-788
// All math in f64, weights stored as f32 exactly where the reference casts (`as f32`).
#pragma once
#include <cstdint>
#include <vector>

namespace ifb {

enum class Lobe : int { Natural = 0, Exact = 1, SharpenPercent = 2 };

struct AxisWeights {
    uint32_t in_size = 0, out_size = 0;
    std::vector<uint32_t> left, right;   // per output sample, inclusive source range (zero-trimmed)
    std::vector<uint32_t> offset;        // out_size+1 prefix offsets into w
    std::vector<float> w;
    uint32_t max_taps = 0;
};

// returns an ifb200_status value (0 = ok)
int compute_axis_weights(int filter_id, double kernel_width_scale, Lobe lobe, float lobe_value,
                         uint32_t out_size, uint32_t in_size, AxisWeights& out);

// ColorContext::new (color.rs:23-48): byte -> working space
void byte_to_float_table(bool linear, float out[256]);
// LINEAR_TO_SRGB_LUT generator (tests/integration/color_conversion.rs:381-388 == lut.rs:14)
void linear_to_srgb_table(uint8_t out[16384]);
// ColorFilterSrgb presets, flow/nodes/color.rs:86-225
int color_filter_matrix(int which, float p, float out[25]);

}  // namespace ifb
