// ifb_kernels.cuh -- sm_100a device code of the resample hot path (product code).
//
// Stands in for the arithmetic the reference delegates to zenresize 0.3.1 behind
// graphics/scaling.rs:93-251 (resize_to_canvas / resize_with_matte / resize_and_composite),
// plus composite_premul_f32_over_srgb_u8 (scaling.rs:254-287) and
// window_bgra32_apply_color_matrix (color_matrix.rs:5-28) as store epilogues.
//
// Arithmetic contract (identical in every kernel here and in oracle/ifb_oracle.c):
//   load   p = (T[b]*af, T[g]*af, T[r]*af, af), af = A8[a]          alpha meaningful   (CH = 4)
//          p = (T[b], T[g], T[r])                                    otherwise          (CH = 3)
//   V pass fmaf chain over source rows, ascending, from +0
//   H pass per aligned group of 4 source columns an fmaf chain from +0, group partials added ascending
//   store  un-premultiply (a > 0), encode, compose, optional colour matrix
// Compiled with -fmad=false: the only fused multiply-adds are the explicit __fmaf_rn below.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ifbk {

// ---------------------------------------------------------------- device-side descriptors
struct JobDev {                 // one scale_and_render call
    const uint8_t* in;          // input window origin
    uint8_t* out;               // canvas origin already offset to (x, y)
    uint32_t in_stride, out_stride;
    uint32_t flags;             // bit0 linear, bit1 alpha_meaningful, bits2-3 compose, bit4 has colour matrix
    float matte[4];             // premultiplied working-space matte (B,G,R,A positional; scaling.rs:141-143)
    float cm[20];               // cm[c*5 + k]: output channel c (0=r,1=g,2=b,3=a) = sum_k cm[c*5+k]*{r,g,b,a,1} (bias already *255)
};
enum : uint32_t { JF_LINEAR = 1u, JF_ALPHA = 2u, JF_COMPOSE_SHIFT = 2, JF_CM = 16u };

struct Tables {                 // per-device constant tables
    const float* t_lin;         // ColorContext::byte_to_float, LinearRGB (color.rs:23-48)
    const float* t_srgb;        // same, StandardRGB (== v * (1/255f)); also the alpha table
    const uint8_t* lut16k;      // LINEAR_TO_SRGB_LUT (lut.rs:14)
};

struct AxisDev {                // CSR contribution windows of one axis (weights.rs PixelRowWeights)
    const uint32_t* left; const uint32_t* right; const uint32_t* off; const float* w;
};

struct StripDev { int X0, X1, k0, pad; };       // output columns [X0,X1) read source columns from k0 (multiple of 4)
struct BandDev  { int Y0, Y1, j0, j1; };        // output rows [Y0,Y1) read source rows j0..j1 inclusive

struct FusedPlanDev {
    uint32_t in_w, in_h, out_w, out_h;
    int n_strips, n_bands;
    const float* vw;            // [in_h][AV] ring-slot weights per source row
    const uint32_t* vdone;      // [in_h]  (first completed y << 8) | count
    const StripDev* strips;
    const BandDev* bands;
    const float* hw;            // [strip][SH][4][NT]
    const int* hxa;             // [strip][NT] first output column touched by thread t's 4 columns
    const uint32_t* hrd;        // [strip][NT] reader u: (first contributing thread) | (count << 16)
};

// ---------------------------------------------------------------- scalar helpers
// color.rs:101-108 uchar_clamp_ff: trunc(x + 0.5) computed exactly, saturated to [0,255], NaN -> 0
__device__ __forceinline__ uint32_t uchar_clamp_ff(float x) {
    if (!(x > 0.0f)) return 0u;
    if (x >= 255.0f) return 255u;
    const float fl = floorf(x);
    const float fr = x - fl;                 // exact
    return (uint32_t)fl + (fr >= 0.5f ? 1u : 0u);
}
// lut.rs:4-8
__device__ __forceinline__ uint32_t lut_encode(const uint8_t* __restrict__ lut16k, float lin) {
    float s = __fmul_rn(lin, 16383.0f);
    s = fminf(fmaxf(s, 0.0f), 16383.0f);     // NaN -> 0, as `NaN as usize`
    return (uint32_t)__ldg(lut16k + (int)s);
}
__device__ __forceinline__ uint32_t encode(bool linear, const uint8_t* __restrict__ lut16k, float v) {
    return linear ? lut_encode(lut16k, v) : uchar_clamp_ff(__fmul_rn(255.0f, v));   // color.rs:59-69
}

// Store epilogue for one destination pixel. F = premultiplied working-space (b,g,r,a).
// `dst` is read only for BlendWithSelf.  Returns packed BGRA8.
__device__ __forceinline__ uint32_t finish_pixel(float b, float g, float r, float a, const JobDev& job,
                                                 const Tables& tb, const uint8_t* dst) {
    const bool linear = job.flags & JF_LINEAR;
    const bool am = job.flags & JF_ALPHA;
    const uint32_t compose = (job.flags >> JF_COMPOSE_SHIFT) & 3u;
    uint32_t ob, og, orr, oa;
    if (compose == 1u) {                                   // BlendWithSelf: scaling.rs:254-287
        if (a > 0.994f || !am) {
            ob = encode(linear, tb.lut16k, b); og = encode(linear, tb.lut16k, g); orr = encode(linear, tb.lut16k, r); oa = 255u;
        } else {
            const uint32_t d = *reinterpret_cast<const uint32_t*>(dst);
            const float* T = linear ? tb.t_lin : tb.t_srgb;
            const float da = (float)(int)(d >> 24);
            const float dc = __fmul_rn(__fsub_rn(1.0f, a), __fadd_rn(__fmul_rn(1.0f / 255.0f, da), 0.0f));
            const float fa = __fadd_rn(a, dc);
            ob = encode(linear, tb.lut16k, __fdiv_rn(__fadd_rn(b, __fmul_rn(dc, __ldg(T + (d & 0xffu)))), fa));
            og = encode(linear, tb.lut16k, __fdiv_rn(__fadd_rn(g, __fmul_rn(dc, __ldg(T + ((d >> 8) & 0xffu)))), fa));
            orr = encode(linear, tb.lut16k, __fdiv_rn(__fadd_rn(r, __fmul_rn(dc, __ldg(T + ((d >> 16) & 0xffu)))), fa));
            oa = uchar_clamp_ff(__fmul_rn(fa, 255.0f));
        }
    } else if (!am) {                                      // scaling.rs:227-232
        ob = encode(linear, tb.lut16k, b); og = encode(linear, tb.lut16k, g); orr = encode(linear, tb.lut16k, r); oa = 255u;
    } else {
        if (compose == 2u) {                               // BlendWithMatte (scaling.rs:119-148)
            const float t = __fsub_rn(1.0f, a);
            b = __fadd_rn(b, __fmul_rn(t, job.matte[0]));
            g = __fadd_rn(g, __fmul_rn(t, job.matte[1]));
            r = __fadd_rn(r, __fmul_rn(t, job.matte[2]));
            a = __fadd_rn(a, __fmul_rn(t, job.matte[3]));
        }
        if (a > 0.0f) { b = __fdiv_rn(b, a); g = __fdiv_rn(g, a); r = __fdiv_rn(r, a); }
        ob = encode(linear, tb.lut16k, b); og = encode(linear, tb.lut16k, g); orr = encode(linear, tb.lut16k, r);
        oa = uchar_clamp_ff(__fmul_rn(a, 255.0f));
    }
    if (job.flags & JF_CM) {                               // color_matrix.rs:5-28, on sRGB bytes
        const float fr = (float)orr, fg = (float)og, fb = (float)ob, fa = (float)oa;
        const float* m = job.cm;
        auto row = [&](int c) {
            float s = __fmul_rn(m[c * 5 + 0], fr);
            s = __fadd_rn(s, __fmul_rn(m[c * 5 + 1], fg));
            s = __fadd_rn(s, __fmul_rn(m[c * 5 + 2], fb));
            s = __fadd_rn(s, __fmul_rn(m[c * 5 + 3], fa));
            return uchar_clamp_ff(__fadd_rn(s, m[c * 5 + 4]));
        };
        orr = row(0); og = row(1); ob = row(2); oa = row(3);
    }
    return ob | (og << 8) | (orr << 16) | (oa << 24);
}

// ---------------------------------------------------------------- generic two-kernel path
// Any geometry / filter.  V pass writes a float4 intermediate [out_h][in_w]; H pass reads it.
__global__ void __launch_bounds__(128) vpass_generic_kernel(const JobDev* __restrict__ jobs, Tables tb, AxisDev av,
                                                            uint32_t in_w, uint32_t out_h, float4* __restrict__ inter) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y;
    const JobDev& job = jobs[blockIdx.z];
    if (x >= in_w) return;
    const bool am = job.flags & JF_ALPHA;
    const float* __restrict__ T = (job.flags & JF_LINEAR) ? tb.t_lin : tb.t_srgb;
    const uint32_t l = av.left[y], r = av.right[y];
    const float* __restrict__ w = av.w + av.off[y];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const uint8_t* __restrict__ src = job.in + (size_t)x * 4;
    for (uint32_t j = l; j <= r; ++j) {
        const uint32_t px = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)j * job.in_stride));
        const float wt = __ldg(w + (j - l));
        float pb = __ldg(T + (px & 0xffu)), pg = __ldg(T + ((px >> 8) & 0xffu)), pr = __ldg(T + ((px >> 16) & 0xffu)), pa = 0.0f;
        if (am) {
            pa = __ldg(tb.t_srgb + (px >> 24));
            pb = __fmul_rn(pb, pa); pg = __fmul_rn(pg, pa); pr = __fmul_rn(pr, pa);
        }
        a0 = __fmaf_rn(wt, pb, a0); a1 = __fmaf_rn(wt, pg, a1); a2 = __fmaf_rn(wt, pr, a2); a3 = __fmaf_rn(wt, pa, a3);
    }
    inter[((size_t)blockIdx.z * out_h + y) * in_w + x] = make_float4(a0, a1, a2, a3);
}

__global__ void __launch_bounds__(128) hpass_generic_kernel(const JobDev* __restrict__ jobs, Tables tb, AxisDev ah,
                                                            uint32_t in_w, uint32_t out_w, uint32_t out_h,
                                                            const float4* __restrict__ inter) {
    const uint32_t X = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y;
    const JobDev& job = jobs[blockIdx.z];
    if (X >= out_w) return;
    const uint32_t l = ah.left[X], r = ah.right[X];
    const float* __restrict__ w = ah.w + ah.off[X];
    const float4* __restrict__ row = inter + ((size_t)blockIdx.z * out_h + y) * in_w;
    float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
    for (uint32_t g = l >> 2; g <= (r >> 2); ++g) {
        const uint32_t k0 = max(g * 4u, l), k1 = min(g * 4u + 3u, r);
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
        for (uint32_t k = k0; k <= k1; ++k) {
            const float wt = __ldg(w + (k - l));
            const float4 v = row[k];
            p0 = __fmaf_rn(wt, v.x, p0); p1 = __fmaf_rn(wt, v.y, p1); p2 = __fmaf_rn(wt, v.z, p2); p3 = __fmaf_rn(wt, v.w, p3);
        }
        f0 = __fadd_rn(f0, p0); f1 = __fadd_rn(f1, p1); f2 = __fadd_rn(f2, p2); f3 = __fadd_rn(f3, p3);
    }
    uint8_t* dst = job.out + (size_t)y * job.out_stride + (size_t)X * 4;
    *reinterpret_cast<uint32_t*>(dst) = finish_pixel(f0, f1, f2, f3, job, tb, dst);
}

// ---------------------------------------------------------------- standalone colour matrix (color_matrix.rs:5-28)
__global__ void __launch_bounds__(256) color_matrix_kernel(uint8_t* __restrict__ px, uint32_t w, uint32_t h, uint32_t stride,
                                                           const float* __restrict__ m20) {
    __shared__ float m[20];
    if (threadIdx.x < 20) m[threadIdx.x] = m20[threadIdx.x];
    __syncthreads();
    const uint32_t total = w * h;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t y = i / w, x = i - y * w;
        uint32_t* p = reinterpret_cast<uint32_t*>(px + (size_t)y * stride) + x;
        const uint32_t v = *p;
        const float fb = (float)(v & 0xffu), fg = (float)((v >> 8) & 0xffu), fr = (float)((v >> 16) & 0xffu), fa = (float)(v >> 24);
        uint32_t o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float s = __fmul_rn(m[c * 5 + 0], fr);
            s = __fadd_rn(s, __fmul_rn(m[c * 5 + 1], fg));
            s = __fadd_rn(s, __fmul_rn(m[c * 5 + 2], fb));
            s = __fadd_rn(s, __fmul_rn(m[c * 5 + 3], fa));
            o[c] = uchar_clamp_ff(__fadd_rn(s, m[c * 5 + 4]));
        }
        *p = o[2] | (o[1] << 8) | (o[0] << 16) | (o[3] << 24);
    }
}

// ---------------------------------------------------------------- fused down-scale kernel
// One CTA = (job, strip of output columns, band of output rows).  Thread t owns source columns
// k0+4t .. k0+4t+3 for the whole band:
//   pass 1 (V): streams source rows top to bottom straight from HBM into registers (one 16-byte load
//               per row), converts through the shared-memory LUT once, and accumulates into a ring of
//               AV register accumulators (one per output row whose window covers the current row).
//   pass 2 (H): when an output row completes, each thread multiplies its 4 V values by its
//               register-resident H weights into <= SH per-output partial sums, parks them in shared
//               memory; after one __syncthreads thread u sums the partials of output column X0+u in
//               ascending order, runs the store epilogue and writes one coalesced BGRA8 row segment.
// Every source pixel is read from HBM once (plus strip/band halos), converted once, and the
// V-filtered intermediate never leaves the SM.
template <int AV, int SH, int CH, int PF>
__global__ void __launch_bounds__(256, (AV * 4 * CH + SH * 4 <= 84) ? 2 : 1) fused_down_kernel(const JobDev* __restrict__ jobs, Tables tb, FusedPlanDev pl) {
    extern __shared__ float smem[];
    const int t = threadIdx.x;
    const int NT = blockDim.x;
    float* sT = smem;                              // 256: colour transfer table
    float* sA = smem + 256;                        // 256: alpha table (only CH == 4)
    float* sPart = smem + 512;                     // 2 x CH x SH x NT partial sums

    const JobDev& job = jobs[blockIdx.y];
    const int strip = blockIdx.x % pl.n_strips;
    const int band = blockIdx.x / pl.n_strips;
    const StripDev sd = pl.strips[strip];
    const BandDev bd = pl.bands[band];

    {
        const float* __restrict__ T = (job.flags & JF_LINEAR) ? tb.t_lin : tb.t_srgb;
        for (int i = t; i < 256; i += NT) { sT[i] = __ldg(T + i); sA[i] = __ldg(tb.t_srgb + i); }
    }
    float hw[SH][4];
#pragma unroll
    for (int q = 0; q < SH; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) hw[q][i] = __ldg(pl.hw + ((size_t)(strip * SH + q) * 4 + i) * NT + t);
    int plane0 = __ldg(pl.hxa + strip * NT + t) % SH;
    const uint32_t rd = __ldg(pl.hrd + strip * NT + t);
    const int NX = sd.X1 - sd.X0;
    __syncthreads();

    int col = sd.k0 + 4 * t;
    if (col > (int)pl.in_w - 4) col = (int)pl.in_w - 4;        // threads past the edge re-read the last group; their H weights are 0
    const uint8_t* __restrict__ src = job.in + (size_t)col * 4;
    const size_t stride = job.in_stride;

    float acc[AV][4][CH];
#pragma unroll
    for (int s = 0; s < AV; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < CH; ++c) acc[s][i][c] = 0.0f;

    uint4 pf[PF];
#pragma unroll
    for (int d = 0; d < PF; ++d) {
        const int j = min(bd.j0 + d, bd.j1);
        pf[d] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)j * stride));
    }
    int nrow = 0;

    for (int jb = bd.j0; jb <= bd.j1; jb += PF) {
#pragma unroll
        for (int d = 0; d < PF; ++d) {
            const int j = jb + d;
            if (j <= bd.j1) {
                const uint4 raw = pf[d];
                {
                    const int jn = min(j + PF, bd.j1);
                    pf[d] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)jn * stride));
                }
                float wv[AV];
#pragma unroll
                for (int s = 0; s < AV; ++s) wv[s] = __ldg(pl.vw + (size_t)j * AV + s);
                const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float p[CH];
                    p[0] = sT[w4[i] & 0xffu];
                    p[1] = sT[(w4[i] >> 8) & 0xffu];
                    p[2] = sT[(w4[i] >> 16) & 0xffu];
                    if (CH == 4) {
                        const float af = sA[w4[i] >> 24];
                        p[0] = __fmul_rn(p[0], af); p[1] = __fmul_rn(p[1], af); p[2] = __fmul_rn(p[2], af);
                        p[CH - 1] = af;
                    }
#pragma unroll
                    for (int s = 0; s < AV; ++s)
#pragma unroll
                        for (int c = 0; c < CH; ++c) acc[s][i][c] = __fmaf_rn(wv[s], p[c], acc[s][i][c]);
                }
                const uint32_t dn = __ldg(pl.vdone + j);
                const int ndone = dn & 0xffu;
                for (int e = 0; e < ndone; ++e) {
                    const int y = (int)(dn >> 8) + e;
                    const int slot = y % AV;
                    float v[4][CH];
#pragma unroll
                    for (int s = 0; s < AV; ++s) {
                        if (s == slot) {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
#pragma unroll
                                for (int c = 0; c < CH; ++c) { v[i][c] = acc[s][i][c]; acc[s][i][c] = 0.0f; }
                        }
                    }
                    if (y < bd.Y0 || y >= bd.Y1) continue;            // halo rows of a neighbouring band
                    float* pb = sPart + (size_t)(nrow & 1) * (CH * SH) * NT;
#pragma unroll
                    for (int q = 0; q < SH; ++q) {
                        int plane = plane0 + q; if (plane >= SH) plane -= SH;
#pragma unroll
                        for (int c = 0; c < CH; ++c) {
                            float ps = __fmaf_rn(hw[q][0], v[0][c], 0.0f);
                            ps = __fmaf_rn(hw[q][1], v[1][c], ps);
                            ps = __fmaf_rn(hw[q][2], v[2][c], ps);
                            ps = __fmaf_rn(hw[q][3], v[3][c], ps);
                            pb[(c * SH + plane) * NT + t] = ps;
                        }
                    }
                    __syncthreads();
                    if (t < NX) {
                        const int X = sd.X0 + t;
                        const int plane = X % SH;
                        const int tg0 = rd & 0xffffu, ng = rd >> 16;
                        float F[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                        for (int g = 0; g < ng; ++g) {
#pragma unroll
                            for (int c = 0; c < CH; ++c) F[c] = __fadd_rn(F[c], pb[(c * SH + plane) * NT + tg0 + g]);
                        }
                        uint8_t* dst = job.out + (size_t)y * job.out_stride + (size_t)X * 4;
                        *reinterpret_cast<uint32_t*>(dst) = finish_pixel(F[0], F[1], F[2], CH == 4 ? F[3] : 0.0f, job, tb, dst);
                    }
                    ++nrow;
                }
            }
        }
    }
}

}  // namespace ifbk
