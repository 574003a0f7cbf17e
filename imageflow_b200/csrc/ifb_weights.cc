// ifb_weights.cc -- see ifb_weights.h.  Product code: must not include or link anything from oracle/.
#include "ifb_weights.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "../../include/ifb200.h"

namespace ifb {
namespace {

constexpr double kPi = 3.14159265358979323846264338327950288;

enum class Shape { Cubic, CubicFast, Sinc, SincWindowed, Box, Triangle, Jinc, Ginseng };

// InterpolationDetails (weights.rs:109-125)
struct Kernel {
    Shape shape = Shape::Box;
    double window = 2.0, blur = 1.0;
    double p1 = 0, p2 = 1, p3 = 1, q1 = 0, q2 = 1, q3 = 1, q4 = 1;

    static Kernel simple(Shape s, double window, double blur) {
        Kernel k; k.shape = s; k.window = window; k.blur = blur; return k;
    }
    // weights.rs:159-174
    static Kernel bc(double window, double blur, double b, double c) {
        Kernel k; k.shape = Shape::Cubic; k.window = window; k.blur = blur;
        const double bx2 = b + b;
        k.p1 = 1.0 - (1.0 / 3.0) * b;
        k.p2 = -3.0 + bx2 + c;
        k.p3 = 2.0 - 1.5 * b - c;
        k.q1 = (4.0 / 3.0) * b + 4.0 * c;
        k.q2 = -8.0 * c - bx2;
        k.q3 = b + 5.0 * c;
        k.q4 = (-1.0 / 6.0) * b - c;
        return k;
    }

    static double j1(double x) {  // weights.rs:460-492
        const double ax = std::fabs(x);
        double v;
        if (ax < 8.0) {
            const double y = x * x;
            const double num = x * (72362614232.0 + y * (-7895059235.0 + y * (242396853.1 + y * (-2972611.439 + y * (15704.48260 + y * (-30.16036606))))));
            const double den = 144725228442.0 + y * (2300535178.0 + y * (18583304.74 + y * (99447.43394 + y * (376.9991397 + y * 1.0))));
            v = num / den;
        } else {
            const double z = 8.0 / ax, y = z * z, xx = ax - 2.356194491;
            const double a1 = 1.0 + y * (0.183105e-2 + y * (-0.3516396496e-4 + y * (0.2457520174e-5 + y * (-0.240337019e-6))));
            const double a2 = 0.04687499995 + y * (-0.2002690873e-3 + y * (0.8449199096e-5 + y * (-0.88228987e-6 + y * 0.105787412e-6)));
            v = std::sqrt((2.0 / kPi) / ax) * (std::cos(xx) * a1 - z * std::sin(xx) * a2);
        }
        return x < 0.0 ? -v : v;
    }

    double eval(double x) const {
        switch (shape) {
        case Shape::Cubic: {                       // weights.rs:352-361
            const double t = std::fabs(x) / blur;
            if (t < 1.0) return p1 + t * (t * (p2 + t * p3));
            if (t < 2.0) return q1 + t * (q2 + t * (q3 + t * q4));
            return 0.0;
        }
        case Shape::CubicFast: {                   // weights.rs:363-373
            const double a = std::fabs(x) / blur, sq = a * a;
            if (a < 1.0) return 1.0 - 2.0 * sq + sq * a;
            if (a < 2.0) return 4.0 - 8.0 * a + 5.0 * sq - sq * a;
            return 0.0;
        }
        case Shape::Sinc: {                        // weights.rs:375-386
            double a = std::fabs(x) / blur;
            if (a == 0.0) return 1.0;
            if (a > window) return 0.0;
            a *= kPi;
            return std::sin(a) / a;
        }
        case Shape::SincWindowed: {                // weights.rs:404-416
            const double v = x / blur, a = std::fabs(v);
            if (a == 0.0) return 1.0;
            if (a > window) return 0.0;
            return window * std::sin(kPi * v / window) * std::sin(v * kPi) / (kPi * kPi * v * v);
        }
        case Shape::Box: {                         // weights.rs:387-394
            const double v = x / blur;
            return (v >= -window && v < window) ? 1.0 : 0.0;
        }
        case Shape::Triangle: {                    // weights.rs:395-402
            const double v = std::fabs(x) / blur;
            return v < 1.0 ? 1.0 - v : 0.0;
        }
        case Shape::Jinc: {                        // weights.rs:418-427
            const double v = std::fabs(x) / blur;
            return v == 0.0 ? 0.5 * kPi : j1(kPi * v) / v;
        }
        case Shape::Ginseng: {                     // weights.rs:444-458
            const double a = std::fabs(x) / blur, tpi = a * kPi;
            if (a == 0.0) return 1.0;
            if (a > 3.0) return 0.0;
            const double ji = 1.2196698912665046 * tpi / window;
            return (j1(ji) / (ji * 0.5)) * std::sin(tpi) / tpi;
        }
        }
        return 0.0;
    }

    // weights.rs:333-350
    double negative_lobe_ratio() const {
        constexpr int samples = 50;
        const double step = window / samples;
        double last = eval(-step), pos = 0.0, neg = 0.0;
        for (int i = 0; i < samples + 3; ++i) {
            const double h = eval(i * step);
            const double area = (h + last) / 2.0 * step;
            last = h;
            if (area > 0.0) pos += area; else neg -= area;
        }
        return neg / pos;
    }
};

// InterpolationDetails::create (weights.rs:176-331); ids = weights.rs:45-78
bool make_kernel(int id, Kernel& k) {
    constexpr double rb = 0.3782157550939987, rc = 0.3108921224530007;
    constexpr double sb = 0.2620145123990142, sc = 0.3689927438004929;
    constexpr double blur3 = 0.9812505644269356, blur2 = 0.9549963639785485;
    switch (id) {
    case IFB200_FILTER_ROBIDOUX_FAST:       k = Kernel::bc(1.05, 1.0, rb, rc); return true;
    case IFB200_FILTER_ROBIDOUX:            k = Kernel::bc(2.0, 1.0, rb, rc); return true;
    case IFB200_FILTER_ROBIDOUX_SHARP:      k = Kernel::bc(2.0, 1.0, sb, sc); return true;
    case IFB200_FILTER_GINSENG:             k = Kernel::simple(Shape::Ginseng, 3.0, 1.0); return true;
    case IFB200_FILTER_GINSENG_SHARP:       k = Kernel::simple(Shape::Ginseng, 3.0, blur3); return true;
    case IFB200_FILTER_LANCZOS:             k = Kernel::simple(Shape::SincWindowed, 3.0, 1.0); return true;
    case IFB200_FILTER_LANCZOS_SHARP:       k = Kernel::simple(Shape::SincWindowed, 3.0, blur3); return true;
    case IFB200_FILTER_LANCZOS2:            k = Kernel::simple(Shape::SincWindowed, 2.0, 1.0); return true;
    case IFB200_FILTER_LANCZOS2_SHARP:      k = Kernel::simple(Shape::SincWindowed, 2.0, blur2); return true;
    case IFB200_FILTER_CUBIC_FAST:          k = Kernel::simple(Shape::CubicFast, 2.0, 1.0); return true;
    case IFB200_FILTER_CUBIC:               k = Kernel::bc(2.0, 1.0, 0.0, 1.0); return true;
    case IFB200_FILTER_CUBIC_SHARP:         k = Kernel::bc(2.0, blur2, 0.0, 1.0); return true;
    case IFB200_FILTER_CATMULL_ROM:         k = Kernel::bc(2.0, 1.0, 0.0, 0.5); return true;
    case IFB200_FILTER_MITCHELL:            k = Kernel::bc(2.0, 1.0, 1.0 / 3.0, 1.0 / 3.0); return true;
    case IFB200_FILTER_CUBIC_BSPLINE:       k = Kernel::bc(2.0, 1.0, 1.0, 0.0); return true;
    case IFB200_FILTER_HERMITE:             k = Kernel::bc(1.0, 1.0, 0.0, 0.0); return true;
    case IFB200_FILTER_JINC:                k = Kernel::simple(Shape::Jinc, 6.0, 1.0); return true;
    case IFB200_FILTER_RAW_LANCZOS3:        k = Kernel::simple(Shape::Sinc, 3.0, 1.0); return true;
    case IFB200_FILTER_RAW_LANCZOS3_SHARP:  k = Kernel::simple(Shape::Sinc, 3.0, blur3); return true;
    case IFB200_FILTER_RAW_LANCZOS2:        k = Kernel::simple(Shape::Sinc, 2.0, 1.0); return true;
    case IFB200_FILTER_RAW_LANCZOS2_SHARP:  k = Kernel::simple(Shape::Sinc, 2.0, blur2); return true;
    case IFB200_FILTER_TRIANGLE:
    case IFB200_FILTER_LINEAR:              k = Kernel::simple(Shape::Triangle, 1.0, 1.0); return true;
    case IFB200_FILTER_BOX:                 k = Kernel::simple(Shape::Box, 0.5, 1.0); return true;
    case IFB200_FILTER_CATMULL_ROM_FAST:    k = Kernel::bc(1.0, 1.0, 0.0, 0.5); return true;
    case IFB200_FILTER_CATMULL_ROM_FAST_SHARP: k = Kernel::bc(1.0, 13.0 / 16.0, 0.0, 0.5); return true;
    case IFB200_FILTER_FASTEST:             k = Kernel::bc(0.74, 0.74, rb, rc); return true;
    case IFB200_FILTER_MITCHELL_FAST:       k = Kernel::bc(1.0, 1.0, 1.0 / 3.0, 1.0 / 3.0); return true;
    case IFB200_FILTER_NCUBIC:              k = Kernel::bc(2.5, 1.0 / 1.1685777620836933, rb, rc); return true;
    case IFB200_FILTER_NCUBIC_SHARP:        k = Kernel::bc(2.5, 1.0 / 1.105822933719019, sb, sc); return true;
    case IFB200_FILTER_LEGACY_IDCT:         k = Kernel::bc(2.0, 1.0 / 1.1685777620836932, rb, rc); return true;
    default: return false;
    }
}

}  // namespace

int compute_axis_weights(int filter_id, double kernel_width_scale, Lobe lobe, float lobe_value,
                         uint32_t out_size, uint32_t in_size, AxisWeights& out) {
    Kernel k;
    if (!make_kernel(filter_id, k)) return IFB200_ERR_BAD_FILTER;
    if (out_size == 0 || in_size == 0) return IFB200_ERR_INVALID_ARGUMENT;
    k.blur *= kernel_width_scale;                                   // weights.rs:156-158

    const double natural = k.negative_lobe_ratio();
    double desired = natural;                                       // LobeRatio::resolve, weights.rs:33-39
    if (lobe == Lobe::Exact) desired = std::min(1.0, std::max(0.0, static_cast<double>(lobe_value)));
    else if (lobe == Lobe::SharpenPercent) desired = std::min(1.0, std::max(natural, static_cast<double>(lobe_value) / 100.0));
    const bool relobe = std::fabs(desired - natural) > 1e-10;

    const double scale = static_cast<double>(out_size) / static_cast<double>(in_size);
    const double down = std::min(1.0, scale);
    const double reach = k.window / down;
    const uint32_t max_window = static_cast<uint32_t>(static_cast<int32_t>(std::ceil(2.0 * ((k.window + 0.5) / down - 0.00001))) + 1);

    out = AxisWeights();
    out.in_size = in_size; out.out_size = out_size;
    out.left.resize(out_size); out.right.resize(out_size); out.offset.resize(static_cast<size_t>(out_size) + 1);
    // Windows are written straight into out.w (sized for the worst case, cut back at the end); the filter's case analysis
    // is hoisted out of the tap loop.  With blur == 1 the division x / blur is the identity and is skipped.
    const size_t cap = static_cast<size_t>(out_size) * std::min<uint32_t>(max_window, in_size);
    out.w.resize(cap);
    float* const wbase = out.w.data();
    size_t used = 0;
    const bool unit_blur = k.blur == 1.0;
    const bool plain_cubic = k.shape == Shape::Cubic && unit_blur;

    for (uint32_t u = 0; u < out_size; ++u) {
        const double center = (u + 0.5) / scale - 0.5;
        const int32_t lo = static_cast<int32_t>(std::ceil(center - reach - 0.0001));
        const int32_t hi = static_cast<int32_t>(std::floor(center + reach + 0.0001));
        uint32_t first = static_cast<uint32_t>(std::max(0, lo));
        uint32_t last = static_cast<uint32_t>(std::min(hi, static_cast<int32_t>(in_size) - 1));
        const uint32_t count = last - first + 1u;
        if (count > max_window) return IFB200_ERR_SOURCE_COUNT_TOO_LARGE;
        if (used + count > cap) return IFB200_ERR_SOURCE_COUNT_TOO_LARGE;

        float* const win = wbase + used;
        double sum = 0.0, sum_neg = 0.0, sum_pos = 0.0;
        for (uint32_t i = 0; i < count; ++i) {
            const double x = down * (static_cast<double>(first + i) - center);
            double v;
            if (plain_cubic) {                                      // Kernel::eval, Shape::Cubic with t = |x| / 1.0
                const double t = std::fabs(x);
                v = t < 1.0 ? k.p1 + t * (t * (k.p2 + t * k.p3)) : (t < 2.0 ? k.q1 + t * (k.q2 + t * (k.q3 + t * k.q4)) : 0.0);
            } else {
                v = k.eval(x);
            }
            if (std::fabs(v) <= 2e-8) v = 0.0;                      // weights.rs:728-730
            win[i] = static_cast<float>(v);
            sum += v;
            sum_neg += std::min(v, 0.0);
            sum_pos += std::max(v, 0.0);
        }
        float scale_neg = static_cast<float>(1.0 / sum), scale_pos = scale_neg;
        if (sum <= 0.0 || relobe) {                                 // weights.rs:743-759
            if (sum_neg < 0.0) {
                if (desired < 1.0) {
                    const double want_pos = 1.0 / (1.0 - desired);
                    const double want_neg = desired * -want_pos;
                    scale_pos = static_cast<float>(want_pos / sum_pos);
                    scale_neg = static_cast<float>(want_neg / sum_neg);
                }
            } else if (sum == 0.0) {
                return IFB200_ERR_TOTAL_WEIGHT_ZERO;
            }
        }
        for (uint32_t i = 0; i < count; ++i) win[i] *= (win[i] < 0.0f) ? scale_neg : scale_pos;

        size_t b = 0, e = count;                                    // zero-trim, weights.rs:771-782
        while (e > b && win[e - 1] == 0.0f) { --e; --last; }
        while (b < e && win[b] == 0.0f) { ++b; ++first; }
        if (b == e) return IFB200_ERR_NO_PIXEL_INPUTS;
        if (b) std::memmove(win, win + b, (e - b) * sizeof(float));
        out.left[u] = first; out.right[u] = last; out.offset[u] = static_cast<uint32_t>(used);
        used += e - b;
        out.max_taps = std::max<uint32_t>(out.max_taps, static_cast<uint32_t>(e - b));
    }
    out.w.resize(used);
    out.offset[out_size] = static_cast<uint32_t>(out.w.size());
    return IFB200_OK;
}

void byte_to_float_table(bool linear, float out[256]) {
    for (int n = 0; n < 256; ++n) {
        const float s = static_cast<float>(n) * (1.0f / 255.0f);    // color.rs:38
        if (!linear) { out[n] = s; continue; }
        out[n] = (s <= 0.04045f) ? s / 12.92f                        // color.rs:85-91
                                 : std::pow((s + 0.055f) / (1.0f + 0.055f), 2.4f);  // float overload == powf
    }
}

void linear_to_srgb_table(uint8_t out[16384]) {
    for (int i = 0; i < 16384; ++i) {
        const double lin = i / 16383.0;
        const double s = (lin <= 0.0031308) ? 12.92 * lin : 1.055 * std::pow(lin, 1.0 / 2.4) - 0.055;
        out[i] = static_cast<uint8_t>(std::min(255.0, std::max(0.0, s * 255.0 + 0.5)));
    }
}

int color_filter_matrix(int which, float p, float o[25]) {
    auto set = [&](std::initializer_list<float> v) { std::copy(v.begin(), v.end(), o); };
    auto gray = [&](float r, float g, float b) {
        set({r, r, r, 0, 0, g, g, g, 0, 0, b, b, b, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1});
    };
    switch (which) {
    case 0: set({0.393f, 0.349f, 0.272f, 0, 0, 0.769f, 0.686f, 0.534f, 0, 0, 0.189f, 0.168f, 0.131f, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0}); return 0;
    case 1: gray(0.229f, 0.587f, 0.114f); return 0;
    case 2: gray(0.5f, 0.5f, 0.5f); return 0;
    case 3: gray(0.2125f, 0.7154f, 0.0721f); return 0;
    case 4: gray(0.5f, 0.419f, 0.081f); return 0;
    case 5: set({-1, 0, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, 1, 0, 1, 1, 1, 0, 1}); return 0;
    case 6: set({1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, p, 0, 0, 0, 0, 0, 1}); return 0;
    case 7: { const float c = p + 1.0f, t = 0.5f * (1.0f - c);
              set({c, 0, 0, 0, 0, 0, c, 0, 0, 0, 0, 0, c, 0, 0, 0, 0, 0, 1, 0, t, t, t, 0, 1}); return 0; }
    case 8: set({1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, p, p, p, 0, 1}); return 0;
    case 9: { const float s = std::max(p + 1.0f, 0.0f), c = 1.0f - s;
              const float cr = 0.3086f * c, cg = 0.6094f * c, cb = 0.0820f * c;
              set({cr + s, cr, cr, 0, 0, cg, cg + s, cg, 0, 0, cb, cb, cb + s, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1}); return 0; }
    default: return IFB200_ERR_INVALID_ARGUMENT;
    }
}

}  // namespace ifb
