/* check_shared_reciprocal.c -- brute-force check of the division used by fused_tile2_kernel's BlendWithSelf composite
 * (imageflow_b200/csrc/ifb_tile2_kernel.cuh, t2_div3): three numerators share one divisor d,
 *     y = RN(1 / d);  q0 = RN(x * y);  q = RN(q0 + RN_exact(x - d * q0) * y)         (two fused multiply-adds)
 * against the IEEE quotient RN(x / d) that the reference computes (scaling.rs:254-287).  Guards replicated from the kernel:
 * d a positive normal in [2^-31, 2^33) whose significand is not all ones (other divisors take the library division in the
 * kernel).  Numerators are tested from 2^-100 up (bounded above by construction: sums of bytes times filter weights); below
 * that the residual x - d * q0 may be inexact, the kernel does not test for it, and does not need to: the quotient is then
 * below 2^-69 on either path and every encoding (x 16383 or x 255, truncated) sends it to the same byte as 0.  Trials: uniformly random bit patterns inside the guards, and "hard" cases built next to rounding
 * midpoints of the quotient (x = RN((q + ulp/2) * d) and its two neighbours).  Build: gcc -O2 -ffp-contract=off -o crs
 * tools/check_shared_reciprocal.c -lm ; run: ./crs [millions of trials per class, default 200].  Prints the mismatch count. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s[2] = {0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull};
static inline uint64_t rnd(void) { uint64_t a = s[0], b = s[1]; s[0] = b; a ^= a << 23; s[1] = a ^ b ^ (a >> 17) ^ (b >> 26); return s[1] + b; }
static inline float f_of(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t u_of(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static inline int guards(float d, float x) {
    const uint32_t db = u_of(d), xb = u_of(x) & 0x7fffffffu;
    if (db - 0x30000000u >= 0x20000000u) return 0;            /* 2^-31 <= d < 2^33 */
    if ((db & 0x7fffffu) == 0x7fffffu) return 0;              /* Markstein's exception */
    if (xb != 0 && xb < 0x0d800000u) return 0;                /* |x| < 2^-100 */
    if (xb >= 0x5f800000u) return 0;                          /* |x| >= 2^64: never produced */
    return 1;
}
static inline float fast_div(float x, float d, float y) {
    const float q0 = x * y;
    const float r = fmaf(-d, q0, x);
    return fmaf(r, y, q0);
}

int main(int argc, char** argv) {
    const long long n = (argc > 1 ? atoll(argv[1]) : 200) * 1000000ll;
    long long bad = 0, tested = 0;
    for (long long i = 0; i < n; ++i) {                       /* class 1: random operands */
        const uint64_t a = rnd();
        const float d = f_of(0x30000000u + (uint32_t)(a % 0x20000000u));
        const float x = f_of((uint32_t)(a >> 32) % 0x5f800000u | ((a >> 31) & 1u ? 0x80000000u : 0u));
        if (!guards(d, x)) continue;
        ++tested;
        const float y = 1.0f / d;
        if (u_of(fast_div(x, d, y)) != u_of(x / d)) { if (bad++ < 10) printf("random: x=%a d=%a fast=%a ieee=%a\n", x, d, fast_div(x, d, y), x / d); }
    }
    for (long long i = 0; i < n; ++i) {                       /* class 2: quotients next to rounding midpoints */
        const uint64_t a = rnd();
        const float d = f_of(0x3a000000u + (uint32_t)(a % 0x06000000u));            /* the composite's range: 2^-11 .. 2 */
        const float q = f_of(0x38000000u + (uint32_t)((a >> 29) % 0x08000000u));    /* quotients 2^-15 .. 2 */
        const double mid = (double)q + 0.5 * ((double)nextafterf(q, 4.0f) - (double)q);
        const float xc = (float)(mid * (double)d);
        const float y = 1.0f / d;
        float xs[3] = {nextafterf(xc, 0.0f), xc, nextafterf(xc, 1e30f)};
        for (int k = 0; k < 3; ++k) {
            const float x = (a >> 63) ? -xs[k] : xs[k];
            if (!guards(d, x)) continue;
            ++tested;
            if (u_of(fast_div(x, d, y)) != u_of(x / d)) { if (bad++ < 10) printf("midpoint: x=%a d=%a fast=%a ieee=%a\n", x, d, fast_div(x, d, y), x / d); }
        }
    }
    printf("tested %lld, mismatches %lld\n", tested, bad);
    return bad != 0;
}
