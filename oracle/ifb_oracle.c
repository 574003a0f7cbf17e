/*
 * ifb_oracle.c -- CPU ORACLE (test infrastructure; see ifb_oracle.h header comment).
 *
 * Plain-C restatement of the reference algorithm.  Build: see oracle/Makefile
 * (-O2 -ffp-contract=off: every fused multiply-add in this file is an explicit fmaf()).
 *
 * Resample arithmetic ("PARITY UNPINNED" part, stated from first principles because the
 * reference delegates it to the un-vendored crate zenresize 0.3.1, scaling.rs:4-6):
 *
 *   load    p[j][x] = (T[b]*af, T[g]*af, T[r]*af, af), af = a*(1/255f)    when alpha is meaningful
 *                     (T[b], T[g], T[r], 0)                                otherwise (scaling.rs:294-302)
 *   H pass  H[j][X][c] = fmaf-chain over the taps k = left_x..right_x of output column X, ascending, starting from +0
 *                        (every source row is filtered horizontally as it streams in, like zenresize's push_row)
 *   V pass  F[y][X][c] = fmaf-chain over the H-filtered rows j = left_y..right_y, ascending, starting from +0
 *   store   un-premultiply (only if a > 0), encode, compose (scaling.rs:55-88)
 *
 * Round 2 changed the order from V-then-H (with the H sum grouped by four source columns) to this plain H-then-V statement:
 * it is the order of a streaming resizer, and both chains are strictly sequential (no grouping); DESIGN.md section 3.
 */
#include "ifb_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ filters (weights.rs:109-492) */

typedef struct filt filt;
typedef double (*filt_fn)(const filt*, double);
struct filt {
    double window, blur;
    double p1, p2, p3, q1, q2, q3, q4;
    filt_fn fn;
};

static const double PI_ = 3.14159265358979323846264338327950288;

/* Numerical-Recipes style J1, restated from weights.rs:460-492 */
static double bessel_j1(double x) {
    double ax = fabs(x), r;
    if (ax < 8.0) {
        double y = x * x;
        double n = x * (72362614232.0 + y * (-7895059235.0 + y * (242396853.1 + y * (-2972611.439 + y * (15704.48260 + y * (-30.16036606))))));
        double d = 144725228442.0 + y * (2300535178.0 + y * (18583304.74 + y * (99447.43394 + y * (376.9991397 + y * 1.0))));
        r = n / d;
    } else {
        double z = 8.0 / ax, y = z * z, xx = ax - 2.356194491;
        double a = 1.0 + y * (0.183105e-2 + y * (-0.3516396496e-4 + y * (0.2457520174e-5 + y * (-0.240337019e-6))));
        double b = 0.04687499995 + y * (-0.2002690873e-3 + y * (0.8449199096e-5 + y * (-0.88228987e-6 + y * 0.105787412e-6)));
        r = sqrt(0.63661977236758134308 / ax) * (cos(xx) * a - z * sin(xx) * b);
    }
    return x < 0.0 ? -r : r;
}

static double f_cubic(const filt* d, double x) {          /* weights.rs:352-361 */
    double t = fabs(x) / d->blur;
    if (t < 1.0) return d->p1 + t * (t * (d->p2 + t * d->p3));
    if (t < 2.0) return d->q1 + t * (d->q2 + t * (d->q3 + t * d->q4));
    return 0.0;
}
static double f_cubic_fast(const filt* d, double x) {     /* weights.rs:363-373 */
    double a = fabs(x) / d->blur, s = a * a;
    if (a < 1.0) return 1.0 - 2.0 * s + s * a;
    if (a < 2.0) return 4.0 - 8.0 * a + 5.0 * s - s * a;
    return 0.0;
}
static double f_sinc(const filt* d, double x) {           /* weights.rs:375-386 */
    double a = fabs(x) / d->blur;
    if (a == 0.0) return 1.0;
    if (a > d->window) return 0.0;
    a *= PI_;
    return sin(a) / a;
}
static double f_box(const filt* d, double x) {            /* weights.rs:387-394 */
    double v = x / d->blur;
    return (v >= -d->window && v < d->window) ? 1.0 : 0.0;
}
static double f_triangle(const filt* d, double x) {       /* weights.rs:395-402 */
    double v = fabs(x) / d->blur;
    return v < 1.0 ? 1.0 - v : 0.0;
}
static double f_sinc_windowed(const filt* d, double x) {  /* weights.rs:404-416 */
    double v = x / d->blur, a = fabs(v);
    if (a == 0.0) return 1.0;
    if (a > d->window) return 0.0;
    return d->window * sin(PI_ * v / d->window) * sin(v * PI_) / (PI_ * PI_ * v * v);
}
static double f_jinc(const filt* d, double x) {           /* weights.rs:418-427 */
    double v = fabs(x) / d->blur;
    if (v == 0.0) return 0.5 * PI_;
    return bessel_j1(PI_ * v) / v;
}
static double f_ginseng(const filt* d, double x) {        /* weights.rs:444-458 */
    double a = fabs(x) / d->blur, tp = a * PI_;
    if (a == 0.0) return 1.0;
    if (a > 3.0) return 0.0;
    double ji = 1.2196698912665046 * tp / d->window;
    double jo = bessel_j1(ji) / (ji * 0.5);
    return jo * sin(tp) / tp;
}

static filt mk(double window, double blur, filt_fn fn) {  /* Default + overrides, weights.rs:126-142 */
    filt f; f.window = window; f.blur = blur; f.fn = fn;
    f.p1 = 0.0; f.p2 = 1.0; f.p3 = 1.0; f.q1 = 0.0; f.q2 = 1.0; f.q3 = 1.0; f.q4 = 1.0;
    return f;
}
static filt mk_bc(double window, double blur, double b, double c) { /* weights.rs:159-174 */
    filt f = mk(window, blur, f_cubic);
    double b2 = b + b;
    f.p1 = 1.0 - (1.0 / 3.0) * b;
    f.p2 = -3.0 + b2 + c;
    f.p3 = 2.0 - 1.5 * b - c;
    f.q1 = (4.0 / 3.0) * b + 4.0 * c;
    f.q2 = -8.0 * c - b2;
    f.q3 = b + 5.0 * c;
    f.q4 = (-1.0 / 6.0) * b - c;
    return f;
}

/* Filter ids are the repr(C) values of weights.rs:45-78 */
static int make_filter(int id, filt* out) {
    const double RB = 0.3782157550939987, RC = 0.3108921224530007;   /* Robidoux */
    const double SB = 0.2620145123990142, SC = 0.3689927438004929;   /* RobidouxSharp */
    const double L3S = 0.9812505644269356, L2S = 0.9549963639785485;
    switch (id) {
    case 1:  *out = mk_bc(1.05, 1.0, RB, RC); break;                 /* RobidouxFast */
    case 2:  *out = mk_bc(2.0, 1.0, RB, RC); break;                  /* Robidoux */
    case 3:  *out = mk_bc(2.0, 1.0, SB, SC); break;                  /* RobidouxSharp */
    case 4:  *out = mk(3.0, 1.0, f_ginseng); break;
    case 5:  *out = mk(3.0, L3S, f_ginseng); break;
    case 6:  *out = mk(3.0, 1.0, f_sinc_windowed); break;            /* Lanczos */
    case 7:  *out = mk(3.0, L3S, f_sinc_windowed); break;
    case 8:  *out = mk(2.0, 1.0, f_sinc_windowed); break;
    case 9:  *out = mk(2.0, L2S, f_sinc_windowed); break;
    case 10: *out = mk(2.0, 1.0, f_cubic_fast); break;
    case 11: *out = mk_bc(2.0, 1.0, 0.0, 1.0); break;                /* Cubic */
    case 12: *out = mk_bc(2.0, L2S, 0.0, 1.0); break;
    case 13: *out = mk_bc(2.0, 1.0, 0.0, 0.5); break;                /* CatmullRom */
    case 14: *out = mk_bc(2.0, 1.0, 1.0 / 3.0, 1.0 / 3.0); break;    /* Mitchell */
    case 15: *out = mk_bc(2.0, 1.0, 1.0, 0.0); break;                /* CubicBSpline */
    case 16: *out = mk_bc(1.0, 1.0, 0.0, 0.0); break;                /* Hermite */
    case 17: *out = mk(6.0, 1.0, f_jinc); break;
    case 18: *out = mk(3.0, 1.0, f_sinc); break;
    case 19: *out = mk(3.0, L3S, f_sinc); break;
    case 20: *out = mk(2.0, 1.0, f_sinc); break;
    case 21: *out = mk(2.0, L2S, f_sinc); break;
    case 22: case 23: *out = mk(1.0, 1.0, f_triangle); break;
    case 24: *out = mk(0.5, 1.0, f_box); break;
    case 25: *out = mk_bc(1.0, 1.0, 0.0, 0.5); break;
    case 26: *out = mk_bc(1.0, 13.0 / 16.0, 0.0, 0.5); break;
    case 27: *out = mk_bc(0.74, 0.74, RB, RC); break;                /* Fastest */
    case 28: *out = mk_bc(1.0, 1.0, 1.0 / 3.0, 1.0 / 3.0); break;
    case 29: *out = mk_bc(2.5, 1.0 / 1.1685777620836933, RB, RC); break;
    case 30: *out = mk_bc(2.5, 1.0 / 1.105822933719019, SB, SC); break;
    case 31: *out = mk_bc(2.0, 1.0 / 1.1685777620836932, RB, RC); break;
    default: return IFO_ERR_BAD_FILTER;
    }
    return IFO_OK;
}

static double percent_negative(const filt* d) {           /* weights.rs:333-350 */
    const int samples = 50;
    double step = d->window / (double)samples;
    double last = d->fn(d, -step), pos = 0.0, neg = 0.0;
    for (int i = 0; i < samples + 3; i++) {
        double h = d->fn(d, (double)i * step);
        double area = (h + last) / 2.0 * step;
        last = h;
        if (area > 0.0) pos += area; else neg -= area;
    }
    return neg / pos;
}

double ifo_percent_negative_weight(int filter, double kws) {
    filt d; if (make_filter(filter, &d)) return NAN;
    d.blur *= kws;
    return percent_negative(&d);
}
double ifo_filter_eval(int filter, double kws, double x) {
    filt d; if (make_filter(filter, &d)) return NAN;
    d.blur *= kws;
    return d.fn(&d, x);
}

static double resolve_lobe(int mode, float v, double natural) {   /* weights.rs:33-39 */
    if (mode == IFO_LOBE_EXACT) { double r = (double)v; return r < 0.0 ? 0.0 : (r > 1.0 ? 1.0 : r); }
    if (mode == IFO_LOBE_SHARPEN_PERCENT) { double g = (double)v / 100.0; double m = natural > g ? natural : g; return m < 1.0 ? m : 1.0; }
    return natural;
}

int ifo_weights(int filter, double kws, int lobe_mode, float lobe_value,
                uint32_t out_size, uint32_t in_size,
                uint32_t* left, uint32_t* right, uint32_t* offsets,
                float* weights, size_t cap)
{
    filt d; int e = make_filter(filter, &d); if (e) return e;
    d.blur *= kws;                                                   /* weights.rs:156-158 */
    double natural = percent_negative(&d);
    double desired = resolve_lobe(lobe_mode, lobe_value, natural);
    double scale = (double)out_size / (double)in_size;
    double ds = scale < 1.0 ? scale : 1.0;
    double half = (d.window + 0.5) / ds;
    uint32_t alloc = (uint32_t)((int32_t)ceil(2.0 * (half - 0.00001)) + 1);
    float* tmp = (float*)malloc(sizeof(float) * ((size_t)alloc + 2));
    if (!tmp) return IFO_ERR_CAPACITY;
    size_t used = 0;
    for (uint32_t u = 0; u < out_size; u++) {
        double center = ((double)u + 0.5) / scale - 0.5;
        int32_t le = (int32_t)ceil(center - d.window / ds - 0.0001);
        int32_t re = (int32_t)floor(center + d.window / ds + 0.0001);
        uint32_t l = (uint32_t)(le > 0 ? le : 0);
        uint32_t r = (uint32_t)(re < (int32_t)in_size - 1 ? re : (int32_t)in_size - 1);
        uint32_t n = r - l + 1u;                                     /* wrapping, as weights.rs:717-718 */
        if (n > alloc) { free(tmp); return IFO_ERR_SOURCE_COUNT_TOO_LARGE; }
        double tot = 0.0, tneg = 0.0, tpos = 0.0;
        for (uint32_t ix = l, t = 0; t < n; ix++, t++) {
            double a = d.fn(&d, ds * ((double)ix - center));
            if (fabs(a) <= 2e-8) a = 0.0;
            tmp[t] = (float)a;
            tot += a;
            tneg += a < 0.0 ? a : 0.0;
            tpos += a > 0.0 ? a : 0.0;
        }
        float nf = (float)(1.0 / tot), pf = nf;
        if (tot <= 0.0 || fabs(desired - natural) > 1e-10) {
            if (tneg < 0.0) {
                if (desired < 1.0) {
                    double tp = 1.0 / (1.0 - desired);
                    double tn = desired * -tp;
                    pf = (float)(tp / tpos);
                    nf = (float)(tn / tneg);
                }
            } else if (tot == 0.0) {
                free(tmp); return IFO_ERR_TOTAL_WEIGHT_ZERO;
            }
        }
        for (uint32_t t = 0; t < n; t++) tmp[t] = tmp[t] < 0.0f ? tmp[t] * nf : tmp[t] * pf;
        /* zero-trim both ends (weights.rs:771-782) */
        uint32_t a0 = 0, a1 = n;
        while (a1 > a0 && tmp[a1 - 1] == 0.0f) { a1--; r--; }
        while (a0 < a1 && tmp[a0] == 0.0f) { a0++; l++; }
        if (a1 == a0) { free(tmp); return IFO_ERR_NO_PIXEL_INPUTS; } /* weights.rs:613-615 */
        if (used + (a1 - a0) > cap) { free(tmp); return IFO_ERR_CAPACITY; }
        left[u] = l; right[u] = r; offsets[u] = (uint32_t)used;
        memcpy(weights + used, tmp + a0, sizeof(float) * (a1 - a0));
        used += a1 - a0;
    }
    offsets[out_size] = (uint32_t)used;
    free(tmp);
    return IFO_OK;
}

/* ------------------------------------------------------------------ colour (color.rs, lut.rs) */

static float srgb_to_linear_f(float s) {                   /* color.rs:85-91 */
    if (s <= 0.04045f) return s / 12.92f;
    return powf((s + 0.055f) / (1.0f + 0.055f), 2.4f);
}
void ifo_byte_to_float_table(int linear, float out[256]) { /* color.rs:23-48 */
    for (int n = 0; n < 256; n++) {
        float v = (float)n * (1.0f / 255.0f);
        out[n] = linear ? srgb_to_linear_f(v) : v;
    }
}
void ifo_linear_to_srgb_table(uint8_t out[16384]) {         /* color_conversion.rs:381-388 (generator of lut.rs:14) */
    for (int i = 0; i < 16384; i++) {
        double lin = (double)i / 16383.0;
        double s = lin <= 0.0031308 ? 12.92 * lin : 1.055 * pow(lin, 1.0 / 2.4) - 0.055;
        double v = s * 255.0 + 0.5;
        if (v < 0.0) v = 0.0;
        if (v > 255.0) v = 255.0;
        out[i] = (uint8_t)v;
    }
}
static uint8_t g_lut16k[16384];
static float g_t_lin[256], g_t_srgb[256];
static int g_tables_ready = 0;
static void ensure_tables(void) {
    if (g_tables_ready) return;
#ifdef _OPENMP
#pragma omp critical(ifo_tables)
#endif
    {
        if (!g_tables_ready) {
            ifo_linear_to_srgb_table(g_lut16k);
            ifo_byte_to_float_table(1, g_t_lin);
            ifo_byte_to_float_table(0, g_t_srgb);
            g_tables_ready = 1;
        }
    }
}
uint8_t ifo_uchar_clamp_ff(float clr) {                      /* color.rs:101-108 */
    if (clr != clr) return 0;                                /* NaN as i16 == 0 */
    double v = (double)clr + 0.5;
    int32_t i;
    if (v >= 32767.0) i = 32767; else if (v <= -32768.0) i = -32768; else i = (int32_t)v; /* saturating `as i16` */
    uint16_t r = (uint16_t)(int16_t)i;
    if (r > 255) r = clr < 0.0f ? 0 : 255;
    return (uint8_t)r;
}
static inline uint8_t lut16k(float lin) {                    /* lut.rs:4-8 */
    float s = lin * 16383.0f;
    if (!(s > 0.0f)) s = 0.0f;                               /* clamp; NaN -> 0 like `NaN as usize` */
    if (s > 16383.0f) s = 16383.0f;
    return g_lut16k[(int)s];
}
static inline uint8_t encode(int linear, float v) {          /* color.rs:59-69 (gamma mode unused on this path) */
    return linear ? lut16k(v) : ifo_uchar_clamp_ff(255.0f * v);
}
uint8_t ifo_floatspace_to_srgb(int linear, float v) { ensure_tables(); return encode(linear, v); }

/* ------------------------------------------------------------------ colour matrix, matte */

void ifo_color_matrix(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, const float* m) { /* color_matrix.rs:5-28 */
    float m40 = m[4 * 5 + 0] * 255.0f, m41 = m[4 * 5 + 1] * 255.0f, m42 = m[4 * 5 + 2] * 255.0f, m43 = m[4 * 5 + 3] * 255.0f;
    for (uint32_t y = 0; y < h; y++) {
        uint8_t* p = px + (size_t)y * stride;
        for (uint32_t x = 0; x < w; x++, p += 4) {
            float b = (float)p[0], g = (float)p[1], r = (float)p[2], a = (float)p[3];
            uint8_t nr = ifo_uchar_clamp_ff(m[0 * 5 + 0] * r + m[1 * 5 + 0] * g + m[2 * 5 + 0] * b + m[3 * 5 + 0] * a + m40);
            uint8_t ng = ifo_uchar_clamp_ff(m[0 * 5 + 1] * r + m[1 * 5 + 1] * g + m[2 * 5 + 1] * b + m[3 * 5 + 1] * a + m41);
            uint8_t nb = ifo_uchar_clamp_ff(m[0 * 5 + 2] * r + m[1 * 5 + 2] * g + m[2 * 5 + 2] * b + m[3 * 5 + 2] * a + m42);
            uint8_t na = ifo_uchar_clamp_ff(m[0 * 5 + 3] * r + m[1 * 5 + 3] * g + m[2 * 5 + 3] * b + m[3 * 5 + 3] * a + m43);
            p[0] = nb; p[1] = ng; p[2] = nr; p[3] = na;
        }
    }
}

static void mat_gray(float r, float g, float b, float* o) {   /* flow/nodes/color.rs:95-103 */
    float m[25] = { r, r, r, 0, 0,  g, g, g, 0, 0,  b, b, b, 0, 0,  0, 0, 0, 1, 0,  0, 0, 0, 0, 1 };
    memcpy(o, m, sizeof m);
}
int ifo_color_filter_matrix(int which, float p, float* o) {   /* flow/nodes/color.rs:86-225 */
    switch (which) {
    case 0: { float m[25] = { 0.393f, 0.349f, 0.272f, 0, 0,  0.769f, 0.686f, 0.534f, 0, 0,  0.189f, 0.168f, 0.131f, 0, 0,  0, 0, 0, 1, 0,  0, 0, 0, 0, 0 };
              memcpy(o, m, sizeof m); return 0; }
    case 1: mat_gray(0.229f, 0.587f, 0.114f, o); return 0;
    case 2: mat_gray(0.5f, 0.5f, 0.5f, o); return 0;
    case 3: mat_gray(0.2125f, 0.7154f, 0.0721f, o); return 0;
    case 4: mat_gray(0.5f, 0.419f, 0.081f, o); return 0;
    case 5: { float m[25] = { -1, 0, 0, 0, 0,  0, -1, 0, 0, 0,  0, 0, -1, 0, 0,  0, 0, 0, 1, 0,  1, 1, 1, 0, 1 };
              memcpy(o, m, sizeof m); return 0; }
    case 6: { float m[25] = { 1, 0, 0, 0, 0,  0, 1, 0, 0, 0,  0, 0, 1, 0, 0,  0, 0, 0, p, 0,  0, 0, 0, 0, 1 };
              memcpy(o, m, sizeof m); return 0; }
    case 7: { float c = p + 1.0f, t = 0.5f * (1.0f - c);
              float m[25] = { c, 0, 0, 0, 0,  0, c, 0, 0, 0,  0, 0, c, 0, 0,  0, 0, 0, 1, 0,  t, t, t, 0, 1 };
              memcpy(o, m, sizeof m); return 0; }
    case 8: { float m[25] = { 1, 0, 0, 0, 0,  0, 1, 0, 0, 0,  0, 0, 1, 0, 0,  0, 0, 0, 1, 0,  p, p, p, 0, 1 };
              memcpy(o, m, sizeof m); return 0; }
    case 9: { float s = p + 1.0f; if (s < 0.0f) s = 0.0f;
              float c = 1.0f - s, cr = 0.3086f * c, cg = 0.6094f * c, cb = 0.0820f * c;
              float m[25] = { cr + s, cr, cr, 0, 0,  cg, cg + s, cg, 0, 0,  cb, cb, cb + s, 0, 0,  0, 0, 0, 1, 0,  0, 0, 0, 0, 1 };
              memcpy(o, m, sizeof m); return 0; }
    default: return IFO_ERR_INVALID_ARGUMENT;
    }
}

void ifo_apply_matte(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, const uint8_t matte[4], int alpha_meaningful) { /* blend.rs:6-59 */
    if (!alpha_meaningful) return;
    ensure_tables();
    const float a2f = 1.0f / 255.0f;
    float ma = (float)matte[3] * a2f;
    float mb = g_t_lin[matte[0]], mg = g_t_lin[matte[1]], mr = g_t_lin[matte[2]];
    for (uint32_t y = 0; y < h; y++) {
        uint8_t* p = px + (size_t)y * stride;
        for (uint32_t x = 0; x < w; x++, p += 4) {
            uint8_t pa = p[3];
            if (pa == 0) { p[0] = matte[0]; p[1] = matte[1]; p[2] = matte[2]; p[3] = matte[3]; }
            else if (pa != 255) {
                float paf = (float)(int32_t)pa * a2f;
                float m_a = (1.0f - paf) * ma;
                float fa = m_a + paf;
                uint8_t nb = lut16k((g_t_lin[p[0]] * paf + mb * m_a) / fa);
                uint8_t ng = lut16k((g_t_lin[p[1]] * paf + mg * m_a) / fa);
                uint8_t nr = lut16k((g_t_lin[p[2]] * paf + mr * m_a) / fa);
                p[0] = nb; p[1] = ng; p[2] = nr; p[3] = ifo_uchar_clamp_ff(255.0f * fa);
            }
        }
    }
}

/* ------------------------------------------------------------------ the resample itself */

typedef struct { uint32_t n; uint32_t *left, *right, *off; float* w; uint32_t max_taps; } wtab;

static void wtab_free(wtab* t) { free(t->left); free(t->right); free(t->off); free(t->w); memset(t, 0, sizeof *t); }
static int wtab_make(wtab* t, int filter, float sharpen, uint32_t out_size, uint32_t in_size) {
    memset(t, 0, sizeof *t);
    filt d; int e = make_filter(filter, &d); if (e) return e;
    double scale = (double)out_size / (double)in_size, ds = scale < 1.0 ? scale : 1.0;
    size_t per = (size_t)ceil(2.0 * ((d.window + 0.5) / ds)) + 2;
    size_t cap = per * out_size;
    t->n = out_size;
    t->left = (uint32_t*)malloc(sizeof(uint32_t) * out_size);
    t->right = (uint32_t*)malloc(sizeof(uint32_t) * out_size);
    t->off = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)out_size + 1));
    t->w = (float*)malloc(sizeof(float) * cap);
    if (!t->left || !t->right || !t->off || !t->w) { wtab_free(t); return IFO_ERR_CAPACITY; }
    /* scaling.rs:104-106: resize_sharpen only when the goal is > 0 */
    e = ifo_weights(filter, 1.0, sharpen > 0.0f ? IFO_LOBE_SHARPEN_PERCENT : IFO_LOBE_NATURAL, sharpen,
                    out_size, in_size, t->left, t->right, t->off, t->w, cap);
    if (e) { wtab_free(t); return e; }
    for (uint32_t i = 0; i < out_size; i++) { uint32_t k = t->off[i + 1] - t->off[i]; if (k > t->max_taps) t->max_taps = k; }
    return IFO_OK;
}

/* convert one input row to working floats (4 per pixel) */
static void load_row(const uint8_t* src, uint32_t w, const float* T, int am, float* dst) {
    const float a2f = 1.0f / 255.0f;
    if (am) {
        for (uint32_t x = 0; x < w; x++) {
            float af = (float)src[4 * x + 3] * a2f;
            dst[4 * x + 0] = T[src[4 * x + 0]] * af;
            dst[4 * x + 1] = T[src[4 * x + 1]] * af;
            dst[4 * x + 2] = T[src[4 * x + 2]] * af;
            dst[4 * x + 3] = af;
        }
    } else {
        for (uint32_t x = 0; x < w; x++) {
            dst[4 * x + 0] = T[src[4 * x + 0]];
            dst[4 * x + 1] = T[src[4 * x + 1]];
            dst[4 * x + 2] = T[src[4 * x + 2]];
            dst[4 * x + 3] = 0.0f;
        }
    }
}

/* scaling.rs:254-287 */
static void composite_over_canvas(const float* src, uint8_t* cv, uint32_t w, int linear, int am, const float* T) {
    float dac = am ? 1.0f / 255.0f : 0.0f;
    float dao = am ? 0.0f : 1.0f;
    for (uint32_t x = 0; x < w; x++, src += 4, cv += 4) {
        float sa = src[3];
        if (sa > 0.994f || !am) {
            cv[0] = encode(linear, src[0]); cv[1] = encode(linear, src[1]); cv[2] = encode(linear, src[2]); cv[3] = 255;
        } else {
            uint8_t da = cv[3];
            float dc = (1.0f - sa) * (dac * (float)(int32_t)da + dao);
            float fa = sa + dc;
            uint8_t b = encode(linear, (src[0] + dc * T[cv[0]]) / fa);
            uint8_t g = encode(linear, (src[1] + dc * T[cv[1]]) / fa);
            uint8_t r = encode(linear, (src[2] + dc * T[cv[2]]) / fa);
            cv[0] = b; cv[1] = g; cv[2] = r; cv[3] = ifo_uchar_clamp_ff(fa * 255.0f);
        }
    }
}

static int resample_core(const ifo_desc* d, float* dbg_v, float* dbg_h) {
    ensure_tables();
    if (!d || !d->in || !d->canvas) return IFO_ERR_INVALID_ARGUMENT;
    /* scaling.rs:24-29 */
    if ((uint64_t)d->h + d->y > d->cv_h || (uint64_t)d->w + d->x > d->cv_w) return IFO_ERR_INVALID_ARGUMENT;
    if (d->w == 0 || d->h == 0 || d->in_w == 0 || d->in_h == 0) return IFO_ERR_INVALID_ARGUMENT;
    if (d->compose < 0 || d->compose > 2) return IFO_ERR_INVALID_ARGUMENT;
    const int linear = d->linear != 0, am = d->alpha_meaningful != 0;
    const float* T = linear ? g_t_lin : g_t_srgb;
    const uint32_t iw = d->in_w, ih = d->in_h, ow = d->w, oh = d->h;

    wtab wv, wh; int e;
    if ((e = wtab_make(&wv, d->filter, d->sharpen_percent, oh, ih))) return e;
    if ((e = wtab_make(&wh, d->filter, d->sharpen_percent, ow, iw))) { wtab_free(&wv); return e; }

    /* Streaming order (the reference pushes rows through zenresize::StreamingResize, scaling.rs:220-249: each pushed row is
     * filtered horizontally, output rows are emitted as soon as their vertical window is complete):
     *   H pass  hrow[j][X][c] = fmaf chain over the taps k = left_x .. right_x, ascending, from +0
     *   V pass  frow[X][c]    = fmaf chain over the rows j = left_y .. right_y, ascending, from +0
     * ring cache of H-filtered rows: enough for the widest V window */
    uint32_t ring = wv.max_taps + 1;
    float* cache = (float*)malloc(sizeof(float) * 4 * (size_t)ow * ring);
    int64_t* cached_row = (int64_t*)malloc(sizeof(int64_t) * ring);
    float* vrow = (float*)malloc(sizeof(float) * 4 * (size_t)iw);       /* one converted source row */
    float* frow = (float*)malloc(sizeof(float) * 4 * (size_t)ow);
    if (!cache || !cached_row || !vrow || !frow) { free(cache); free(cached_row); free(vrow); free(frow); wtab_free(&wv); wtab_free(&wh); return IFO_ERR_CAPACITY; }
    for (uint32_t i = 0; i < ring; i++) cached_row[i] = -1;

    /* matte in working space (scaling.rs:141-143: B,G,R,A positional) */
    float mt[4] = { 0, 0, 0, 0 };
    if (d->compose == IFO_COMPOSE_BLEND_WITH_MATTE && am) {
        float ma = (float)d->matte_bgra[3] * (1.0f / 255.0f);
        mt[0] = T[d->matte_bgra[0]] * ma; mt[1] = T[d->matte_bgra[1]] * ma; mt[2] = T[d->matte_bgra[2]] * ma; mt[3] = ma;
    }

    for (uint32_t y = 0; y < oh; y++) {
        const uint32_t l = wv.left[y], r = wv.right[y];
        const float* wy = wv.w + wv.off[y];
        memset(frow, 0, sizeof(float) * 4 * (size_t)ow);
        for (uint32_t j = l; j <= r; j++) {
            uint32_t slot = j % ring;
            float* cr = cache + (size_t)slot * ow * 4;
            if (cached_row[slot] != (int64_t)j) {
                /* ---- H pass of source row j: sequential fmaf chain per output column */
                load_row(d->in + (size_t)j * d->in_stride, iw, T, am, vrow);
                for (uint32_t X = 0; X < ow; X++) {
                    const uint32_t hl = wh.left[X], hr = wh.right[X];
                    const float* wx = wh.w + wh.off[X];
                    float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
                    for (uint32_t k = hl; k <= hr; k++) {
                        const float wgt = wx[k - hl];
                        const float* v = vrow + (size_t)k * 4;
                        p0 = fmaf(wgt, v[0], p0); p1 = fmaf(wgt, v[1], p1); p2 = fmaf(wgt, v[2], p2); p3 = fmaf(wgt, v[3], p3);
                    }
                    cr[4 * X + 0] = p0; cr[4 * X + 1] = p1; cr[4 * X + 2] = p2; cr[4 * X + 3] = p3;
                }
                cached_row[slot] = j;
                if (dbg_v) memcpy(dbg_v + (size_t)j * ow * 4, cr, sizeof(float) * 4 * (size_t)ow);
            }
            /* ---- V pass: fmaf chain over rows, ascending */
            const float wgt = wy[j - l];
            for (size_t i = 0; i < (size_t)ow * 4; i++) frow[i] = fmaf(wgt, cr[i], frow[i]);
        }
        if (dbg_h) memcpy(dbg_h + (size_t)y * ow * 4, frow, sizeof(float) * 4 * (size_t)ow);
        /* ---- store */
        uint8_t* dst = d->canvas + (size_t)(d->y + y) * d->cv_stride + (size_t)d->x * 4;
        if (d->compose == IFO_COMPOSE_BLEND_WITH_SELF) {
            composite_over_canvas(frow, dst, ow, linear, am, T);
        } else {
            for (uint32_t X = 0; X < ow; X++) {
                float b = frow[4 * X], g = frow[4 * X + 1], rr = frow[4 * X + 2], a = frow[4 * X + 3];
                if (!am) {                                            /* scaling.rs:227-232 */
                    dst[4 * X] = encode(linear, b); dst[4 * X + 1] = encode(linear, g); dst[4 * X + 2] = encode(linear, rr); dst[4 * X + 3] = 255;
                    continue;
                }
                if (d->compose == IFO_COMPOSE_BLEND_WITH_MATTE) {
                    float t = 1.0f - a;
                    b = b + t * mt[0]; g = g + t * mt[1]; rr = rr + t * mt[2]; a = a + t * mt[3];
                }
                if (a > 0.0f) { b = b / a; g = g / a; rr = rr / a; }
                dst[4 * X] = encode(linear, b); dst[4 * X + 1] = encode(linear, g); dst[4 * X + 2] = encode(linear, rr);
                dst[4 * X + 3] = ifo_uchar_clamp_ff(a * 255.0f);
            }
        }
    }
    free(cache); free(cached_row); free(vrow); free(frow);
    wtab_free(&wv); wtab_free(&wh);
    if (d->color_matrix)
        ifo_color_matrix(d->canvas + (size_t)d->y * d->cv_stride + (size_t)d->x * 4, ow, oh, d->cv_stride, d->color_matrix);
    return IFO_OK;
}

int ifo_scale_and_render(const ifo_desc* d) { return resample_core(d, NULL, NULL); }
int ifo_resample_stages(const ifo_desc* d, float* v, float* h) { return resample_core(d, v, h); }

int ifo_scale_and_render_batch(const ifo_desc* d, size_t n, int threads) {
    int err = 0;
    ensure_tables();
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (long i = 0; i < (long)n; i++) {
        int e = resample_core(&d[i], NULL, NULL);
        if (e) {
#ifdef _OPENMP
#pragma omp critical(ifo_err)
#endif
            { if (!err) err = e; }
        }
    }
    (void)threads;
    return err;
}

/* imageflow_b200/synth.py noise_np, byte for byte (bench.py's CPU arm needs thousands of frames: numpy is too slow for that) */
static uint32_t synth_mix(uint32_t v) { v ^= v >> 16; v *= 0x7FEB352Du; v ^= v >> 15; v *= 0x846CA68Bu; v ^= v >> 16; return v; }
void ifo_synth_noise(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, uint32_t seed, int alpha_mixed) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long y = 0; y < (long)h; y++) {
        uint8_t* r = px + (size_t)y * stride;
        for (uint32_t x = 0; x < w; x++) {
            uint32_t base = synth_mix((x * 0x9E3779B1u) ^ ((uint32_t)y * 0x85EBCA77u) ^ (0x1F2E3D4Cu + seed));
            uint8_t a = 255;
            if (alpha_mixed) {
                uint32_t h2 = synth_mix(base ^ 0xA5A5A5A5u), sel = (h2 >> 24) & 3u;
                a = sel == 0 ? 0 : sel == 1 ? 255 : (uint8_t)((h2 >> 8) & 0xFF);
            }
            r[4 * x] = (uint8_t)base; r[4 * x + 1] = (uint8_t)(base >> 8); r[4 * x + 2] = (uint8_t)(base >> 16); r[4 * x + 3] = a;
        }
    }
}

int ifo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* graphics/transpose.rs:95-121 (bitmap_window_transpose): the u32 pixel at (x, y) moves to (y, x) */
void ifo_transpose(const uint8_t* from, uint32_t from_stride, uint32_t w, uint32_t h, uint8_t* to, uint32_t to_stride) {
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x)
            memcpy(to + (size_t)x * to_stride + (size_t)y * 4, from + (size_t)y * from_stride + (size_t)x * 4, 4);
}
/* graphics/flip.rs:10-22: top and bottom rows swap; the middle row of an odd height stays */
void ifo_flip_vertical(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride) {
    for (uint32_t y = 0; y < h / 2; ++y) {
        uint8_t* a = px + (size_t)y * stride; uint8_t* b = px + (size_t)(h - 1 - y) * stride;
        for (uint32_t i = 0; i < w * 4; ++i) { const uint8_t t = a[i]; a[i] = b[i]; b[i] = t; }
    }
}
/* graphics/flip.rs:25-39: every row reversed pixel-wise */
void ifo_flip_horizontal(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride) {
    for (uint32_t y = 0; y < h; ++y) {
        uint32_t* r = (uint32_t*)(px + (size_t)y * stride);
        for (uint32_t x = 0; x < w / 2; ++x) { const uint32_t t = r[x]; r[x] = r[w - 1 - x]; r[w - 1 - x] = t; }
    }
}

/* white_balance.rs:14-41 (area_threshold): note that BOTH scans compare against low_threshold, as the reference does */
static void ifo_area_threshold(const uint64_t* hist, uint64_t total, double low_threshold, uint64_t* low_out, uint64_t* high_out) {
    uint64_t low = 0, high = 255, area = 0;
    const double pixel_count = (double)total;
    for (int ix = 0; ix < 256; ++ix) { area += hist[ix]; if ((double)area / pixel_count > low_threshold) { low = (uint64_t)ix; break; } }
    area = 0;
    for (int ix = 255; ix >= 0; --ix) { area += hist[ix]; if ((double)area / pixel_count > low_threshold) { high = (uint64_t)ix; break; } }
    *low_out = low; *high_out = high;
}
void ifo_white_balance(uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, float threshold, uint8_t* maps_out) {
    uint64_t hist[3][256];
    memset(hist, 0, sizeof hist);
    for (uint32_t y = 0; y < h; ++y) {                               /* histogram.rs:7-20, order R, G, B */
        const uint8_t* r = px + (size_t)y * stride;
        for (uint32_t x = 0; x < w; ++x) { hist[0][r[x * 4 + 2]]++; hist[1][r[x * 4 + 1]]++; hist[2][r[x * 4 + 0]]++; }
    }
    const double low_threshold = (double)(threshold < 0.0f ? 0.006f : threshold);   /* white_balance.rs:76-77 */
    uint8_t maps[3][256];
    for (int c = 0; c < 3; ++c) {
        uint64_t low, high;
        ifo_area_threshold(hist[c], (uint64_t)w * h, low_threshold, &low, &high);
        const double scale = 255.0 / (double)(high - low);           /* white_balance.rs:44-48; usize subtraction wraps in release builds */
        for (uint64_t v = 0; v < 256; ++v) {
            const uint64_t d = v > low ? v - low : 0;                /* saturating_sub */
            double m = round((double)d * scale);
            m = fmin(m, 255.0); m = fmax(m, 0.0);                    /* f64::min / f64::max ignore a NaN operand */
            maps[c][v] = (uint8_t)m;
        }
    }
    for (uint32_t y = 0; y < h; ++y) {                               /* white_balance.rs:50-67 */
        uint8_t* r = px + (size_t)y * stride;
        for (uint32_t x = 0; x < w; ++x) { r[x * 4 + 2] = maps[0][r[x * 4 + 2]]; r[x * 4 + 1] = maps[1][r[x * 4 + 1]]; r[x * 4 + 0] = maps[2][r[x * 4 + 0]]; }
    }
    if (maps_out) memcpy(maps_out, maps, sizeof maps);
}

/* ------------------------------------------------------------------------------------------------
 * Whitespace detection (graphics/whitespace.rs), SURVEY.md section 8(f) item 4. */
typedef enum { WS_TOP, WS_RIGHT, WS_BOTTOM, WS_LEFT, WS_NONDIR } ws_edge;
typedef struct { ws_edge edge; float x1p, y1p, x2p, y2p; } ws_region;
/* whitespace.rs:30-131 (twelve thin strips), :133-158 (four inward scans), :159-165 (everything at once) */
static const ws_region WS_QUICK[12] = {
    {WS_LEFT, 0.0f, 0.5f, 0.5f, 0.5f},     {WS_RIGHT, 0.5f, 0.5f, 1.0f, 0.5f},
    {WS_LEFT, 0.0f, 0.677f, 0.5f, 0.677f}, {WS_RIGHT, 0.5f, 0.677f, 1.0f, 0.677f},
    {WS_LEFT, 0.0f, 0.333f, 0.5f, 0.333f}, {WS_RIGHT, 0.5f, 0.333f, 1.0f, 0.333f},
    {WS_TOP, 0.5f, 0.0f, 0.5f, 0.5f},      {WS_TOP, 0.677f, 0.0f, 0.677f, 0.5f},   {WS_TOP, 0.333f, 0.0f, 0.333f, 0.5f},
    {WS_BOTTOM, 0.5f, 0.5f, 0.5f, 1.0f},   {WS_BOTTOM, 0.677f, 0.5f, 0.677f, 1.0f}, {WS_BOTTOM, 0.333f, 0.5f, 0.333f, 1.0f},
};
static const ws_region WS_INWARD[4] = {
    {WS_TOP, 0.0f, 0.0f, 1.0f, 1.0f}, {WS_RIGHT, 0.0f, 0.0f, 1.0f, 1.0f}, {WS_BOTTOM, 0.0f, 0.0f, 1.0f, 1.0f}, {WS_LEFT, 0.0f, 0.0f, 1.0f, 1.0f},
};
static const ws_region WS_FULL = {WS_NONDIR, 0.0f, 0.0f, 1.0f, 1.0f};

typedef struct { uint32_t w, h, threshold, min_x, max_x, min_y, max_y; uint64_t centres; } ws_search;
static uint32_t ws_min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static uint32_t ws_max(uint32_t a, uint32_t b) { return a > b ? a : b; }

/* whitespace.rs:220-281; returns 0 when nothing is left to search */
static int ws_search_rect(const ws_search* s, const ws_region* r, uint32_t* ox1, uint32_t* oy1, uint32_t* ox2, uint32_t* oy2) {
    uint32_t x1 = ws_min(s->w, (uint32_t)floorf(r->x1p * (float)(s->w - 1)));
    uint32_t x2 = ws_min(s->w, (uint32_t)floorf(r->x2p * (float)(s->w - 1)));
    uint32_t y1 = ws_min(s->h, (uint32_t)floorf(r->y1p * (float)(s->h - 1)));
    uint32_t y2 = ws_min(s->h, (uint32_t)floorf(r->y2p * (float)(s->h - 1)));
    switch (r->edge) {
    case WS_LEFT:   x1 = 0; x2 = ws_min(x2, s->min_x); break;
    case WS_RIGHT:  x1 = ws_max(x1, s->max_x); x2 = s->w; break;
    case WS_TOP:    y1 = 0; y2 = ws_min(y2, s->min_y); break;
    case WS_BOTTOM: y1 = ws_max(y1, s->max_y); y2 = s->h; break;
    default: break;
    }
    if (x1 == x2 || y1 == y2) return 0;                     /* :256-258 -- note: every thin strip of WS_QUICK ends here */
    const uint32_t min_w = (r->edge == WS_RIGHT || r->edge == WS_LEFT) ? 3u : 7u;
    const uint32_t min_h = (r->edge == WS_TOP || r->edge == WS_BOTTOM) ? 3u : 7u;
    while (y2 - y1 < min_h && (y1 > 0 || y2 < s->h)) { y1 = y1 > 0 ? y1 - 1 : 0; y2 = ws_min(s->h, y2 + 1); }
    while (x2 - x1 < min_w && (x1 > 0 || x2 < s->w)) { x1 = x1 > 0 ? x1 - 1 : 0; x2 = ws_min(s->w, x2 + 1); }
    *ox1 = x1; *oy1 = y1; *ox2 = x2; *oy2 = y2;
    return 1;
}

/* whitespace.rs:465-505: Bgra32 (alpha meaningful) weighs by alpha and rounds up; Bgr32 ignores the fourth byte */
static uint8_t ws_gray(const uint8_t* p, int alpha_meaningful) {
    const uint32_t lum = 233u * p[0] + 1197u * p[1] + 610u * p[2];
    if (!alpha_meaningful) return (uint8_t)(lum / 2048u);
    const uint32_t v = lum * p[3];
    const uint32_t g = v / 524288u + (v % 524288u ? 1u : 0u);           /* div_ceil, then `as u16`, then clamp */
    return g > 255u ? 255u : (uint8_t)g;
}

/* whitespace.rs:540-613 for one 3x3 neighbourhood m (row-major): 0xFF, or the packed local box (see ifb_oracle.h) */
static uint8_t ws_code(const uint8_t m[9], int32_t threshold) {
    const int32_t gx = 3 * m[0] + 10 * m[3] + 3 * m[6] - 3 * m[2] - 10 * m[5] - 3 * m[8];
    const int32_t gy = 3 * m[0] + 10 * m[1] + 3 * m[2] - 3 * m[6] - 10 * m[7] - 3 * m[8];
    if (abs(gx) + abs(gy) <= threshold) return 0xFF;
    uint32_t lminx = 2, lminy = 2, lmaxx = 1, lmaxy = 1;
    for (uint32_t my = 0; my < 3; ++my) {
        int found = 0;
        if (abs((int32_t)m[my * 3] - (int32_t)m[my * 3 + 1]) > threshold) { lminx = ws_min(lminx, 1); lmaxx = ws_max(lmaxx, 1); found = 1; }
        if (abs((int32_t)m[my * 3 + 1] - (int32_t)m[my * 3 + 2]) > threshold) { lminx = ws_min(lminx, 2); lmaxx = ws_max(lmaxx, 2); found = 1; }
        if (found) { lminy = ws_min(lminy, my); lmaxy = ws_max(lmaxy, my + 1); }
    }
    for (uint32_t mx = 0; mx < 3; ++mx) {
        int found = 0;
        if (abs((int32_t)m[mx] - (int32_t)m[mx + 3]) > threshold) { lminy = ws_min(lminy, 1); lmaxy = ws_max(lmaxy, 1); found = 1; }
        if (abs((int32_t)m[mx + 3] - (int32_t)m[mx + 6]) > threshold) { lminy = ws_min(lminy, 2); lmaxy = ws_max(lmaxy, 2); found = 1; }
        if (found) { lminx = ws_min(lminx, mx); lmaxx = ws_max(lmaxx, mx + 1); }
    }
    return (uint8_t)(lminx | ((lmaxx - 1) << 2) | (lminy << 4) | ((lmaxy - 1) << 6));
}

/* whitespace.rs:333-421 */
/* codes != NULL: take the per-pixel codes from this w*h map (ifo_whitespace_codes) instead of computing them from px */
static void ws_check_region(ws_search* s, const uint8_t* px, uint32_t stride, int am, const ws_region* region, const uint8_t* codes) {
    uint32_t x1, y1, x2, y2;
    if (!ws_search_rect(s, region, &x1, &y1, &x2, &y2)) return;
    const uint32_t w = x2 - x1, h = y2 - y1, buf_size = 2048u;
    const uint32_t window_width = ws_min(w, region->edge == WS_NONDIR ? buf_size / 7u : (uint32_t)ceilf(sqrtf((float)buf_size)));
    const uint32_t window_height = ws_min(h, buf_size / window_width);
    if (window_width <= 2 || window_height <= 2) return;    /* the reference would divide by zero here; not reachable from detect_content */
    const uint32_t vertical_windows = (uint32_t)ceilf((float)h / (float)(window_height - 2));
    const uint32_t horizontal_windows = (uint32_t)ceilf((float)w / (float)(window_width - 2));
    uint8_t gray[2048];
    for (uint32_t wr = 0; wr < vertical_windows; ++wr) {
        for (uint32_t wc = 0; wc < horizontal_windows; ++wc) {
            uint32_t bx = x1 + (window_width - 2) * wc, by = y1 + (window_height - 2) * wr;
            uint32_t bw = ws_min(ws_max(3, x2 - bx), window_width), bh = ws_min(ws_max(3, y2 - by), window_height);
            const uint32_t bx2 = bx + bw, by2 = by + bh;
            const int excluded_x = s->min_x < bx && s->max_x > bx2;
            const int excluded_y = s->min_y < by && s->max_y > by2;
            if (excluded_x && excluded_y) continue;
            if (excluded_y && s->min_x < bx2 && bx2 < s->max_x) bw = ws_max(3, s->min_x - bx);
            else if (excluded_y && s->max_x > bx && bx > s->min_x) { bx = ws_min(bx2 - 3, s->max_x); bw = bx2 - bx; }
            if (excluded_x && s->min_y < by2 && by2 < s->max_y) bh = ws_max(3, s->min_y - by);
            else if (excluded_x && s->max_y > by && by > s->min_y) { by = ws_min(by2 - 3, s->max_y); bh = by2 - by; }
            if (by + bh > s->h) { if (bh <= s->h) by = s->h - bh; else { by = 0; bh = s->h; } }
            if (bx + bw > s->w) { if (bw <= s->w) bx = s->w - bw; else { bx = 0; bw = s->w; } }
            if ((size_t)bw * bh > sizeof gray) return;       /* cannot happen: bw <= window_width, bh <= window_height */
            if (!codes)
                for (uint32_t y = 0; y < bh; ++y)
                    for (uint32_t x = 0; x < bw; ++x) gray[y * bw + x] = ws_gray(px + (size_t)(by + y) * stride + (size_t)(bx + x) * 4, am);
            for (uint32_t y = 1; y + 1 < bh; ++y) {
                for (uint32_t x = 1; x + 1 < bw; ++x) {
                    uint8_t c;
                    if (codes) c = codes[(size_t)(by + y) * s->w + (bx + x)];
                    else {
                        uint8_t m[9];
                        for (int k = 0; k < 9; ++k) m[k] = gray[(y - 1 + k / 3) * bw + (x - 1 + k % 3)];
                        c = ws_code(m, (int32_t)s->threshold);
                    }
                    ++s->centres;
                    if (c == 0xFF) continue;
                    const uint32_t lminx = (c & 3u) + bx + x - 1, lmaxx = ((c >> 2) & 3u) + 1 + bx + x - 1;
                    const uint32_t lminy = ((c >> 4) & 3u) + by + y - 1, lmaxy = ((c >> 6) & 3u) + 1 + by + y - 1;
                    if (lminx < s->min_x) s->min_x = lminx;
                    if (lmaxx > s->max_x) s->max_x = lmaxx;
                    if (lminy < s->min_y) s->min_y = lminy;
                    if (lmaxy > s->max_y) s->max_y = lmaxy;
                }
            }
        }
    }
}

static int ws_detect(const uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, int alpha_meaningful, uint32_t threshold,
                     uint32_t rect_out[4], uint64_t* centres_out, const uint8_t* codes) {                /* whitespace.rs:284-331 */
    if ((!px && !codes) || !rect_out || w == 0 || h == 0 || w > 0x7fffffffu || h > 0x7fffffffu) return IFO_ERR_INVALID_ARGUMENT;
    if (centres_out) *centres_out = 0;
    if (w < 3 || h < 3) { rect_out[0] = 0; rect_out[1] = 0; rect_out[2] = w; rect_out[3] = h; return 0; }
    ws_search s = {w, h, threshold, w, 0, h, 0, 0};
    for (int i = 0; i < 12; ++i) ws_check_region(&s, px, stride, alpha_meaningful, &WS_QUICK[i], codes);
    const int64_t separately = (int64_t)s.min_x * s.h + (int64_t)s.min_y * s.w + ((int64_t)s.w - s.max_x) * s.h + ((int64_t)s.h - s.max_y) * s.w;
    if (separately > (int64_t)s.h * s.w) ws_check_region(&s, px, stride, alpha_meaningful, &WS_FULL, codes);
    else for (int i = 0; i < 4; ++i) ws_check_region(&s, px, stride, alpha_meaningful, &WS_INWARD[i], codes);
    if (s.min_x == w && s.max_x == 0 && s.min_y == h && s.max_y == 0) { rect_out[0] = 0; rect_out[1] = 0; rect_out[2] = w; rect_out[3] = h; }
    else { rect_out[0] = s.min_x; rect_out[1] = s.min_y; rect_out[2] = s.max_x; rect_out[3] = s.max_y; }
    if (centres_out) *centres_out = s.centres;
    return 0;
}

int ifo_detect_content(const uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, int alpha_meaningful, uint32_t threshold,
                       uint32_t rect_out[4], uint64_t* centres_out) {
    return ws_detect(px, w, h, stride, alpha_meaningful, threshold, rect_out, centres_out, NULL);
}
int ifo_detect_content_from_codes(const uint8_t* codes, uint32_t w, uint32_t h, uint32_t rect_out[4], uint64_t* centres_out) {
    return ws_detect(NULL, w, h, 0, 0, 0, rect_out, centres_out, codes);
}

void ifo_whitespace_codes(const uint8_t* px, uint32_t w, uint32_t h, uint32_t stride, int alpha_meaningful, uint32_t threshold, uint8_t* map) {
    for (uint32_t y = 0; y < h; ++y) {
        for (uint32_t x = 0; x < w; ++x) {
            if (x == 0 || y == 0 || x + 1 >= w || y + 1 >= h) { map[(size_t)y * w + x] = 0xFF; continue; }
            uint8_t m[9];
            for (int k = 0; k < 9; ++k) m[k] = ws_gray(px + (size_t)(y - 1 + k / 3) * stride + (size_t)(x - 1 + k % 3) * 4, alpha_meaningful);
            map[(size_t)y * w + x] = ws_code(m, (int32_t)threshold);
        }
    }
}
