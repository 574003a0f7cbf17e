/* Exhaustive check (all 2^32 float bit patterns) that the three-instruction clamp used by the second tile kernel,
 *     min(cvt.rzi.u32.f32(add.rz.f32(x, 0.5f)), 255)
 * equals the reference's uchar_clamp_ff (graphics/color.rs:101-108):
 *     let v = (x as f64 + 0.5) as i16 as u16;  if v > 255 { if x < 0.0 { 0 } else { 255 } } else { v as u8 }
 * (Rust float->int `as` casts saturate and send NaN to 0).
 * build: gcc -O2 -fopenmp -frounding-math tools/check_clamp_rz.c -o /tmp/check_clamp_rz -lm ; run: /tmp/check_clamp_rz */
#include <fenv.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static uint32_t reference(float x) {
    const double d = (double)x + 0.5;                                   /* exact: a float plus 0.5 fits in a double */
    int16_t i;
    if (d != d) i = 0; else if (d >= 32767.0) i = 32767; else if (d <= -32768.0) i = -32768; else i = (int16_t)d;   /* trunc */
    const uint16_t v = (uint16_t)i;
    if (v > 255) return x < 0.0f ? 0u : 255u;
    return v;
}
static uint32_t cvt_rzi_u32(float s) {                                  /* PTX cvt.rzi.u32.f32: NaN -> 0, saturating */
    if (s != s) return 0u;
    if (s <= 0.0f) return 0u;
    if (s >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)s;                                                 /* truncation */
}
int main(void) {
    unsigned long long bad = 0;
#pragma omp parallel reduction(+ : bad)
    {
        fesetround(FE_TOWARDZERO);                                      /* add.rz */
#pragma omp for schedule(static)
        for (long long b = 0; b < (1ll << 32); ++b) {
            const uint32_t bits = (uint32_t)b;
            float x; memcpy(&x, &bits, 4);
            volatile float s = x + 0.5f;                                /* rounded toward zero */
            uint32_t got = cvt_rzi_u32(s);
            if (got > 255u) got = 255u;
            if (got != reference(x)) { if (bad < 5) printf("mismatch at %a: %u vs %u\n", x, got, reference(x)); ++bad; }
        }
    }
    printf("%llu mismatches over all 2^32 floats\n", bad);
    return bad != 0;
}
