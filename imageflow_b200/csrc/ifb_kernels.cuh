// ifb_kernels.cuh -- sm_100a device code of the resample hot path (product code).
//
// Stands in for the arithmetic the reference delegates to zenresize 0.3.1 behind
// graphics/scaling.rs:93-251 (resize_to_canvas / resize_with_matte / resize_and_composite),
// plus composite_premul_f32_over_srgb_u8 (scaling.rs:254-287) and
// window_bgra32_apply_color_matrix (color_matrix.rs:5-28) as store epilogues.
//
// Arithmetic contract (identical in every kernel here and in oracle/ifb_oracle.c):
//   load   p = (T[b]*af, T[g]*af, T[r]*af, af), af = a*(1/255f)     alpha meaningful   (CH = 4)
//          p = (T[b], T[g], T[r])                                    otherwise          (CH = 3)
//   V pass fmaf chain over source rows, ascending, from +0
//   H pass per aligned group of 4 source columns an fmaf chain from +0, group partials added ascending
//   store  un-premultiply (a > 0), encode, compose, optional colour matrix
// Compiled with -fmad=false: the only fused multiply-adds are the explicit __fmaf_rn below.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace ifbk {

// ---------------------------------------------------------------- device-side descriptors (their own file: tests/cpu_emu includes it too)
#include "ifb_types.cuh"

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
// SIMPLE = true: the caller guarantees compose == ReplaceSelf and no colour matrix (the common thumbnail case);
// the composite / matte / matrix code is then not even compiled into the kernel.
template <bool SIMPLE = false>
__device__ __forceinline__ uint32_t finish_pixel(float b, float g, float r, float a, const uint32_t flags, const JobDev& job,
                                                 const Tables& tb, const uint8_t* dst) {
    const bool linear = flags & JF_LINEAR;
    const bool am = flags & JF_ALPHA;
    const uint32_t compose = SIMPLE ? 0u : (flags >> JF_COMPOSE_SHIFT) & 3u;
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
    if (!SIMPLE && (flags & JF_CM)) {                  // color_matrix.rs:5-28, on sRGB bytes
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
template <bool SIMPLE = false>
__device__ __forceinline__ uint32_t finish_pixel(float b, float g, float r, float a, const JobDev& job, const Tables& tb, const uint8_t* dst) {
    return finish_pixel<SIMPLE>(b, g, r, a, job.flags, job, tb, dst);
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

// ---------------------------------------------------------------- tile kernel (up-scales, 1:1, mild down-scales)
// One CTA = (job, TOW x TOH tile of output pixels).  The few source pixels the tile needs are converted once into a
// shared-memory float4 tile, the V pass writes a second shared-memory tile [TOH][source columns], the H pass reads
// it; nothing but the source pixels and the destination pixels touches HBM.  Same arithmetic, same bits as the
// other kernels (V chain ascending; H per aligned group of 4 source columns, partials added ascending).

__global__ void __launch_bounds__(256) fused_tile_kernel(const JobDev* __restrict__ jobs, Tables tb, AxisDev av, AxisDev ah, TilePlanDev pl) {
    extern __shared__ __align__(16) float4 tsm[];
    float4* sIn = tsm;                                   // [max_ir][max_ic]
    float4* sV = tsm + (size_t)pl.max_ir * pl.max_ic;    // [toh][max_ic]
    const JobDev& job = jobs[blockIdx.y];
    const int tx = blockIdx.x % pl.tiles_x, ty = blockIdx.x / pl.tiles_x;
    const int X0 = tx * pl.tow, X1 = min(X0 + pl.tow, (int)pl.out_w);
    const int Y0 = ty * pl.toh, Y1 = min(Y0 + pl.toh, (int)pl.out_h);
    const int c0 = (int)__ldg(ah.left + X0), c1 = (int)__ldg(ah.right + (X1 - 1));
    const int r0 = (int)__ldg(av.left + Y0), r1 = (int)__ldg(av.right + (Y1 - 1));
    const int ic = c1 - c0 + 1, ir = r1 - r0 + 1, pitch = pl.max_ic;
    const bool am = job.flags & JF_ALPHA;
    const float* __restrict__ T = (job.flags & JF_LINEAR) ? tb.t_lin : tb.t_srgb;
    // ---- A: source tile -> working floats
    for (int i = threadIdx.x; i < ir * ic; i += blockDim.x) {
        const int r = i / ic, c = i - r * ic;
        const uint32_t px = __ldg(reinterpret_cast<const uint32_t*>(job.in + (size_t)(r0 + r) * job.in_stride) + (c0 + c));
        float pb = __ldg(T + (px & 0xffu)), pg = __ldg(T + ((px >> 8) & 0xffu)), pr = __ldg(T + ((px >> 16) & 0xffu)), pa = 0.0f;
        if (am) {
            pa = __fmul_rn(__uint2float_rn(px >> 24), 1.0f / 255.0f);
            pb = __fmul_rn(pb, pa); pg = __fmul_rn(pg, pa); pr = __fmul_rn(pr, pa);
        }
        sIn[r * pitch + c] = make_float4(pb, pg, pr, pa);
    }
    __syncthreads();
    // ---- B: V pass for the tile's output rows over its source columns
    const int nrows = Y1 - Y0;
    for (int i = threadIdx.x; i < nrows * ic; i += blockDim.x) {
        const int yl = i / ic, c = i - yl * ic;
        const uint32_t l = __ldg(av.left + Y0 + yl), r = __ldg(av.right + Y0 + yl);
        const float* __restrict__ w = av.w + __ldg(av.off + Y0 + yl);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (uint32_t j = l; j <= r; ++j) {
            const float wt = __ldg(w + (j - l));
            const float4 v = sIn[(int)(j - r0) * pitch + c];
            a0 = __fmaf_rn(wt, v.x, a0); a1 = __fmaf_rn(wt, v.y, a1); a2 = __fmaf_rn(wt, v.z, a2); a3 = __fmaf_rn(wt, v.w, a3);
        }
        sV[yl * pitch + c] = make_float4(a0, a1, a2, a3);
    }
    __syncthreads();
    // ---- C: H pass + store epilogue
    const int ncols = X1 - X0;
    for (int i = threadIdx.x; i < nrows * ncols; i += blockDim.x) {
        const int yl = i / ncols, xl = i - yl * ncols;
        const int X = X0 + xl;
        const uint32_t l = __ldg(ah.left + X), r = __ldg(ah.right + X);
        const float* __restrict__ w = ah.w + __ldg(ah.off + X);
        const float4* __restrict__ row = sV + yl * pitch - c0;
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
        uint8_t* dst = job.out + (size_t)(Y0 + yl) * job.out_stride + (size_t)X * 4;
        *reinterpret_cast<uint32_t*>(dst) = finish_pixel(f0, f1, f2, f3, job, tb, dst);
    }
}

// ---------------------------------------------------------------- tile kernel, second form (its own file: tests/cpu_emu runs this source on the CPU)
#ifndef IFB_DYNAMIC_SMEM                                 // (tests/cpu_emu defines it as a pointer to an exactly-sized heap block)
#define IFB_DYNAMIC_SMEM(name_) extern __shared__ __align__(16) unsigned char name_[]
#endif
#include "ifb_tile2_kernel.cuh"

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

// ---------------------------------------------------------------- apply_matte (graphics/blend.rs:6-59)
// Encoder-side flatten over a solid colour, in place, linear light.
__global__ void __launch_bounds__(256) apply_matte_kernel(uint8_t* __restrict__ px, uint32_t w, uint32_t h, uint32_t stride,
                                                          uint32_t matte_bgra, Tables tb) {
    const float a2f = 1.0f / 255.0f;
    const float ma = __fmul_rn((float)(matte_bgra >> 24), a2f);
    const float mb = __ldg(tb.t_lin + (matte_bgra & 0xffu)), mg = __ldg(tb.t_lin + ((matte_bgra >> 8) & 0xffu)), mr = __ldg(tb.t_lin + ((matte_bgra >> 16) & 0xffu));
    const uint32_t total = w * h;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t y = i / w, x = i - y * w;
        uint32_t* p = reinterpret_cast<uint32_t*>(px + (size_t)y * stride) + x;
        const uint32_t v = *p;
        const uint32_t pa = v >> 24;
        if (pa == 0u) { *p = matte_bgra; continue; }
        if (pa == 255u) continue;
        const float paf = __fmul_rn((float)(int)pa, a2f);
        const float m_a = __fmul_rn(__fsub_rn(1.0f, paf), ma);
        const float fa = __fadd_rn(m_a, paf);
        auto ch = [&](uint32_t byte, float m) {
            const float lin = __fdiv_rn(__fadd_rn(__fmul_rn(__ldg(tb.t_lin + byte), paf), __fmul_rn(m, m_a)), fa);
            return lut_encode(tb.lut16k, lin);
        };
        const uint32_t nb = ch(v & 0xffu, mb), ng = ch((v >> 8) & 0xffu, mg), nr = ch((v >> 16) & 0xffu, mr);
        *p = nb | (ng << 8) | (nr << 16) | (uchar_clamp_ff(__fmul_rn(255.0f, fa)) << 24);
    }
}

// ---------------------------------------------------------------- transpose / flips
// Pure data movement on BGRA8 words, bound by HBM: 4 bytes read + 4 bytes written per pixel.
// graphics/transpose.rs:95-121 (bitmap_window_transpose): to[x][y] = from[y][x].  32x32-word tiles through shared memory
// (padded to 33 columns: conflict-free), 32x8 threads, both the reads and the writes are full 128-byte rows.
__global__ void __launch_bounds__(256) transpose_bgra8_kernel(const uint8_t* __restrict__ from, uint32_t from_stride, uint32_t w, uint32_t h,
                                                              uint8_t* __restrict__ to, uint32_t to_stride) {
    __shared__ uint32_t tile[32][33];
    const uint32_t x0 = blockIdx.x * 32u, y0 = blockIdx.y * 32u;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const uint32_t x = x0 + threadIdx.x, y = y0 + threadIdx.y + j;
        if (x < w && y < h) tile[threadIdx.y + j][threadIdx.x] = __ldcs(reinterpret_cast<const uint32_t*>(from + (size_t)y * from_stride) + x);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const uint32_t oy = x0 + threadIdx.y + j, ox = y0 + threadIdx.x;      // destination row = source column
        if (oy < w && ox < h) __stcs(reinterpret_cast<uint32_t*>(to + (size_t)oy * to_stride) + ox, tile[threadIdx.x][threadIdx.y + j]);
    }
}
// graphics/flip.rs:10-22 (flow_bitmap_bgra_flip_vertical_safe): rows y and h-1-y swap, in place; the middle row of an odd
// height stays.  One thread per element of the top half; T = uint4 (four pixels) when rows are 16-byte aligned and the
// width is a multiple of 4, else uint32_t.
template <class T>
__global__ void __launch_bounds__(256) flip_vertical_bgra8_kernel(uint8_t* __restrict__ px, uint32_t w_elems, uint32_t h, uint32_t stride) {
    const uint32_t half = h / 2u;
    const uint64_t total = (uint64_t)half * w_elems;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(i / w_elems), x = (uint32_t)(i - (uint64_t)y * w_elems);
        T* a = reinterpret_cast<T*>(px + (size_t)y * stride) + x;
        T* b = reinterpret_cast<T*>(px + (size_t)(h - 1u - y) * stride) + x;
        const T va = *a, vb = *b;
        *a = vb; *b = va;
    }
}
// graphics/flip.rs:25-39 (flow_bitmap_bgra_flip_horizontal_safe): every row reversed pixel-wise, in place.
__global__ void __launch_bounds__(256) flip_horizontal_bgra8_kernel(uint8_t* __restrict__ px, uint32_t w, uint32_t h, uint32_t stride) {
    const uint32_t half = w / 2u;
    const uint64_t total = (uint64_t)half * h;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(i / half), x = (uint32_t)(i - (uint64_t)y * half);
        uint32_t* row = reinterpret_cast<uint32_t*>(px + (size_t)y * stride);
        const uint32_t va = row[x], vb = row[w - 1u - x];
        row[x] = vb; row[w - 1u - x] = va;
    }
}
// the same on groups of four pixels (rows 16-byte aligned, width a multiple of 4): group g swaps with group n-1-g, each
// reversed inside; the middle group of an odd count is reversed in place
__global__ void __launch_bounds__(256) flip_horizontal_bgra8_v4_kernel(uint8_t* __restrict__ px, uint32_t w4, uint32_t h, uint32_t stride) {
    const uint32_t per_row = (w4 + 1u) / 2u;
    const uint64_t total = (uint64_t)per_row * h;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(i / per_row), g = (uint32_t)(i - (uint64_t)y * per_row);
        uint4* row = reinterpret_cast<uint4*>(px + (size_t)y * stride);
        const uint4 a = row[g], b = row[w4 - 1u - g];
        row[g] = make_uint4(b.w, b.z, b.y, b.x);
        row[w4 - 1u - g] = make_uint4(a.w, a.z, a.y, a.x);      // g == w4-1-g (middle group): both stores write the same value
    }
}

// ---------------------------------------------------------------- white balance (first half)
// flow/nodes/white_balance.rs:93-121 = three passes: per-channel histograms (graphics/histogram.rs:7-20), the area-threshold
// byte mappings (white_balance.rs:14-48, f64 arithmetic, a handful of operations), and the in-place remap (:50-67).
// hist[0..255] = R, [256..511] = G, [512..767] = B (histogram.rs: "histogram order is RGB").
// One private histogram per warp (8 x 768 counters = 24 KB of shared memory): lanes of a warp still collide on popular bins,
// warps do not.  V4: rows are 16-byte aligned and the width is a multiple of 4 -> four pixels per load.
template <bool V4>
__global__ void __launch_bounds__(256) histogram_bgra8_kernel(const uint8_t* __restrict__ px, uint32_t w, uint32_t h, uint32_t stride,
                                                              unsigned long long* __restrict__ hist) {
    __shared__ uint32_t sh[8][768];
    for (int i = threadIdx.x; i < 8 * 768; i += 256) (&sh[0][0])[i] = 0u;
    __syncthreads();
    uint32_t* mine = sh[threadIdx.x >> 5];
    auto count = [&](uint32_t v) {
        atomicAdd(&mine[(v >> 16) & 0xffu], 1u);
        atomicAdd(&mine[256u + ((v >> 8) & 0xffu)], 1u);
        atomicAdd(&mine[512u + (v & 0xffu)], 1u);
    };
    const uint32_t we = V4 ? w / 4u : w;
    const uint64_t total = (uint64_t)we * h;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(i / we), x = (uint32_t)(i - (uint64_t)y * we);
        if (V4) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4*>(px + (size_t)y * stride) + x);
            count(v.x); count(v.y); count(v.z); count(v.w);
        } else {
            count(__ldcs(reinterpret_cast<const uint32_t*>(px + (size_t)y * stride) + x));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 768; i += 256) {
        uint32_t c = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) c += sh[k][i];
        if (c) atomicAdd(&hist[i], (unsigned long long)c);
    }
}
// white_balance.rs:14-48.  One block of 256 threads: thread c < 3 scans channel c's histogram from both ends (both scans
// compare against low_threshold, as the reference does), then thread v writes maps[c * 256 + v] for the three channels.
// `high - low` is a usize subtraction in the reference: it wraps when the thresholds cross (release build semantics).
__global__ void __launch_bounds__(256) white_balance_maps_kernel(const unsigned long long* __restrict__ hist, unsigned long long total_pixels,
                                                                 double low_threshold, uint8_t* __restrict__ maps) {
    __shared__ unsigned long long lo[3], hi[3];
    const int t = threadIdx.x;
    if (t < 3) {
        const unsigned long long* hc = hist + t * 256;
        const double pixel_count = (double)total_pixels;
        unsigned long long low = 0, high = 255, area = 0;
        for (int ix = 0; ix < 256; ++ix) { area += hc[ix]; if (__ddiv_rn((double)area, pixel_count) > low_threshold) { low = ix; break; } }
        area = 0;
        for (int ix = 255; ix >= 0; --ix) { area += hc[ix]; if (__ddiv_rn((double)area, pixel_count) > low_threshold) { high = ix; break; } }
        lo[t] = low; hi[t] = high;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double scale = __ddiv_rn(255.0, (double)(hi[c] - lo[c]));                       // u64 wrap-around like usize
        const unsigned long long d = (unsigned long long)t > lo[c] ? (unsigned long long)t - lo[c] : 0ull;   // saturating_sub
        double m = round(__dmul_rn((double)d, scale));
        m = fmax(fmin(m, 255.0), 0.0);                                                         // NaN -> 255 -> 255, as f64::min/max
        maps[c * 256 + t] = (uint8_t)m;
    }
}
// white_balance.rs:50-67: r, g, b through their byte maps, alpha untouched, in place
template <bool V4>
__global__ void __launch_bounds__(256) apply_byte_maps_bgra8_kernel(uint8_t* __restrict__ px, uint32_t w, uint32_t h, uint32_t stride,
                                                                    const uint8_t* __restrict__ maps) {
    __shared__ uint8_t sm[768];
    for (int i = threadIdx.x; i < 768; i += 256) sm[i] = maps[i];
    __syncthreads();
    auto remap = [&](uint32_t v) {
        return (v & 0xff000000u) | ((uint32_t)sm[(v >> 16) & 0xffu] << 16) | ((uint32_t)sm[256u + ((v >> 8) & 0xffu)] << 8) | (uint32_t)sm[512u + (v & 0xffu)];
    };
    const uint32_t we = V4 ? w / 4u : w;
    const uint64_t total = (uint64_t)we * h;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(i / we), x = (uint32_t)(i - (uint64_t)y * we);
        if (V4) {
            uint4* p = reinterpret_cast<uint4*>(px + (size_t)y * stride) + x;
            const uint4 v = *p;
            *p = make_uint4(remap(v.x), remap(v.y), remap(v.z), remap(v.w));
        } else {
            uint32_t* p = reinterpret_cast<uint32_t*>(px + (size_t)y * stride) + x;
            *p = remap(*p);
        }
    }
}

// ---------------------------------------------------------------- whitespace detection (second half)
// In its own file so that tests/ can also execute this very source under a CPU emulation of the CUDA built-ins it uses.
#include "ifb_whitespace_kernel.cuh"

// ---------------------------------------------------------------- fused down-scale kernel
// One CTA = (job, strip of output columns, band of output rows).  Thread t owns source columns
// k0+4t .. k0+4t+3 for the whole band:
//   pass 1 (V): streams source rows top to bottom, 16 bytes per thread per row -- through a private shared-memory FIFO
//               filled by cp.async where its stages fit, else into two register sets -- converts them through a
//               bank-conflict-free (lane-replicated) shared-memory LUT once, and accumulates into a ring of AV register
//               accumulators: output row y lives in slot y mod AV for as long as its window is open.  The
//               per-row "program" (slot weights + which output rows complete) is streamed through a
//               double-buffered shared-memory chunk with cp.async.
//   pass 2 (H): when an output row completes, each thread multiplies its 4 V values by its H weights into <= SH
//               per-output partial sums, parks them in shared memory and arrives on an mbarrier; one completion later
//               thread u waits for that mbarrier (normally already complete), sums the partials of output column X0+u in
//               ascending order, runs the store epilogue and writes one coalesced BGRA8 row segment.  Partials are
//               double-buffered and consecutive rows are finished by alternating halves of the CTA; the CTA never
//               rendezvous per output row.
// Every source pixel is read from HBM once (plus strip/band halos), converted once, and the
// V-filtered intermediate never leaves the SM.
//
// Shared memory is addressed through explicit ld/st.shared with 32-bit window addresses held in registers: every
// hot-loop access is `[register + immediate]`, nothing is recomputed per row.
constexpr int kProgChunk = 32;                       // source rows per program chunk
constexpr int kLutBytes = 256 * 256;                 // LUT region: 256 rows of 256 B (see below)
// Offset of a CTA's dynamic shared memory inside the shared window (1 KB is reserved by the system on sm_90+).  The LUT
// gather folds it into the immediate field of the load; the engine checks it with smem_base_probe_kernel before it
// ever launches a fused kernel, and uses the other kernels if a driver should lay shared memory out differently.
constexpr uint32_t kSmemWindowBase = 0x400u;

template <int AV> struct ProgLayout {
    static constexpr int kW = AV;                    // weight words, by ring slot (FFMA2 takes the weight as a broadcast scalar operand)
    static constexpr int kDone = kW;                 // index of the completion word: (first completed y << 8) | (its slot << 4) | count
    static constexpr int kWords = (kW + 1 + 3) / 4 * 4;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async16_to(uint32_t smem_addr, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
// TMA-style L2 prefetch of a contiguous global range (16-byte aligned, size a multiple of 16): no register, no scoreboard
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// A value the compiler must keep in a register (it cannot re-derive it, so it cannot rematerialise it per use).
__device__ __forceinline__ uint32_t pinned_reg(uint32_t x) { uint32_t y; asm volatile("mov.b32 %0, %1;" : "=r"(y) : "r"(x)); return y; }
// read-only-after-setup table gather (may be scheduled freely: its operands depend on the pixel just loaded)
__device__ __forceinline__ float lds_table(uint32_t a) { float v; asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float2 lds_f32x2(uint32_t a) {
    float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ uint4 lds_u32x4(uint32_t a) {
    uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

// Split barrier on a shared-memory mbarrier: a warp announces "my partials of this row are parked" without waiting
// (arrive has release semantics), and only waits -- one completion later -- for all warps to have done so.
__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n"
                 "IFB_MBAR_WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@!p bra IFB_MBAR_WAIT_%=;\n\t}" ::"r"(addr), "r"(parity) : "memory");
}

__global__ void smem_base_probe_kernel(uint32_t* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if (threadIdx.x == 0) *out = (uint32_t)__cvta_generic_to_shared(smem_raw);
}

// Shared-memory map of the fused kernel.
//   [0, 64 KB)   row v (256 B): bytes 0..127 = T[v] replicated for the 32 lanes, so the byte address of a lookup is
//                (v << 8) | (lane << 2): ONE PRMT builds it from the packed pixel, and the gather is bank-conflict
//                free for any image content.  Bytes 128..255 of the rows ("holes") hold the strip's H weights:
//                word w of the H-weight block (see FusedPlanDev::hw) lives at ((w >> 5) << 8) + 128 + ((w & 31) << 2),
//                which puts pair k of thread t at hole_base(t) + k * 16 * NT.
//   then         row-program double buffer, partial sums (2 x CH x SH x NT floats), two mbarriers.
//   then         row stages: ST x NT x 16 B, each thread's private FIFO of source rows in flight (cp.async), when it fits.
constexpr int kSmemPerCtaFor2 = 113 * 1024;          // 228 KB per SM, 1 KB reserved per CTA: two CTAs of this size fit
template <int AV, int SH, int CH, int NT> struct FusedSmem {
    static constexpr int kProgBytes = 2 * kProgChunk * ProgLayout<AV>::kWords * 4;
    static constexpr int kPartBuf = CH * SH * NT * 4;           // one partial-sum buffer
    static constexpr int kProgOff = kLutBytes;
    static constexpr int kPartOff = kProgOff + kProgBytes;
    static constexpr int kBarOff = kPartOff + 2 * kPartBuf;      // two 8-byte mbarriers (one per partial buffer)
    static constexpr int kStageOff = kBarOff + 16;
    static constexpr int kRowStage = NT * 16;
    // Stages that fit next to a second CTA (at least 3, at most 6); if fewer fit the kernel prefetches into registers
    // instead (0 stages) -- unless a second CTA does not fit anyway, then the single CTA takes 6 stages.
    static constexpr int kFit = (kSmemPerCtaFor2 - kStageOff) / kRowStage;
    static constexpr int kStages = kFit >= 3 ? (kFit > 6 ? 6 : kFit) : (kStageOff <= kSmemPerCtaFor2 ? 0 : 6);
    static constexpr int kTotal = kStageOff + kStages * kRowStage;
    static constexpr int kHwPairs = (SH / 2) * 4 + (SH & 1) * 2; // float2 pairs of H weights per thread
    static_assert(kHwPairs * 2 * NT / 32 <= 256, "H weights must fit in the LUT holes");
};

// GA ("gather ahead", staged rows with an even stage count only): the table lookups of source row i+1 are issued before the
// multiply-adds of row i, into a second set of working registers, so that a warp covers its own shared-memory latency
// instead of relying on the three other warps of its scheduler.  Same operations on the same values: bit-identical.
template <int AV, int SH, int CH, int PF, int NT, bool SIMPLE, bool GA = false>
__global__ void __launch_bounds__(NT, 512 / NT) fused_down_kernel(const JobDev* __restrict__ jobs, Tables tb, FusedPlanDev pl) {
    using PL = ProgLayout<AV>;
    using SM = FusedSmem<AV, SH, CH, NT>;
    constexpr int NV = 4 * CH;                      // working floats per thread per row, channel-planar: [c][pixel]
    constexpr int kRec = PL::kWords * 4;            // bytes per program record
    constexpr int ST = SM::kStages;                 // > 0: source rows are staged through shared memory, PF is not used
    constexpr int RING = ST > 0 ? ST : 2 * PF;      // unrolled copies of the row code
    static_assert(RING <= 12, "ring positions");
    static_assert(!GA || (ST > 0 && ST % 2 == 0), "gather-ahead needs an even number of row stages");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int t = threadIdx.x;
    unsigned char* sLut = smem_raw;
    uint32_t* sProg = reinterpret_cast<uint32_t*>(smem_raw + SM::kProgOff);
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem_raw);
    if ((sb & 0xffffu) != kSmemWindowBase) __trap();      // never taken: the engine probes the layout before using this kernel

    const JobDev& job = jobs[blockIdx.y];
    const int strip = blockIdx.x % pl.n_strips;
    const int band = blockIdx.x / pl.n_strips;
    const StripDev sd = pl.strips[strip];
    const BandDev bd = pl.bands[band];

    {
        const float* __restrict__ T = (job.flags & JF_LINEAR) ? tb.t_lin : tb.t_srgb;
        for (int i = t; i < 256 * 32; i += NT)
            *reinterpret_cast<float*>(sLut + ((i >> 5) << 8) + ((i & 31) << 2)) = __ldg(T + (i >> 5));
        for (int w = t; w < SM::kHwPairs * 2 * NT; w += NT)
            *reinterpret_cast<float*>(sLut + ((w >> 5) << 8) + 128 + ((w & 31) << 2)) = __ldg(pl.hw + (size_t)strip * SH * 4 * NT + w);
    }
    // ---- per-thread constants
    // outputs of this strip are finished by alternating halves of the CTA when they fit in one half
    const int NX = sd.X1 - sd.X0;
    const bool alternate = NX <= NT / 2;
    const int my_half = t / (NT / 2);
    const int u = alternate ? t - my_half * (NT / 2) : t;             // output column of the strip this thread finishes
    // bit p of fin: this thread finishes the rows emitted into partial buffer p
    const uint32_t fin = u < NX ? (alternate ? (1u << my_half) : 3u) : 0u;
    const uint32_t rmeta = __ldg(pl.hrd + strip * NT + (u < NX ? u : 0));
    // partial (c, plane, thread) is the float at kPartOff + buffer * kPartBuf + ((c * SH + plane) * NT + thread) * 4
    const uint32_t rd_base = sb + SM::kPartOff + ((rmeta >> 28) * NT + (rmeta & 0xfffu)) * 4u;
    const uint32_t rd_groups = (rmeta >> 12) & 0xfffu;
    const uint32_t wr_base = sb + SM::kPartOff + (uint32_t)t * 4u;
    const uint32_t hw_base = sb + (((uint32_t)t >> 4) << 8) + 128u + (((uint32_t)t & 15u) << 3);
    const uint32_t lane4 = pinned_reg(((uint32_t)(t & 31) * 4u) | ((sb >> 16) << 8));   // PRMT operand: low byte + window bits 16..31
    const uint32_t flags = job.flags;
    uint8_t* const out_col = job.out + (size_t)(sd.X0 + u) * 4;
    const uint32_t out_stride = job.out_stride;

    if (t == 0) { mbar_init(sb + SM::kBarOff, NT / 32); mbar_init(sb + SM::kBarOff + 8, NT / 32); }   // one arrival per warp
    // program chunk 0
    const uint32_t* __restrict__ gprog = pl.vprog + (size_t)bd.j0 * PL::kWords;
    const int total_rows = bd.j1 - bd.j0 + 1;
    {
        const int n16 = min(kProgChunk, total_rows) * PL::kWords / 4;
        for (int i = t; i < n16; i += NT) cp_async16(sProg + i * 4, gprog + i * 4);
        cp_async_wait_all();
    }
    __syncthreads();

    int col = sd.k0 + 4 * t;
    // threads past the edge re-read the last aligned group (its tail may be row padding); their H weights are 0
    if (col > (int)((pl.in_w - 1) & ~3u)) col = (int)((pl.in_w - 1) & ~3u);
    const size_t stride = job.in_stride;
    const uint8_t* __restrict__ pnext = job.in + (size_t)col * 4 + (size_t)bd.j0 * stride;
    int left = total_rows;                          // source rows not requested yet, counting the one pnext points at (>= 1)

    // acc[s]: output row y with y mod AV == s, while its window is open
    float acc[AV][NV];
#pragma unroll
    for (int s = 0; s < AV; ++s)
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[s][i] = 0.0f;

    // bytes of a source row this CTA reads (thread 0's pointer is the start of the segment); rows are padded to 16 bytes
    const uint32_t seg_bytes = min((uint32_t)NT * 16u, ((pl.in_w * 4u + 15u) & ~15u) - (uint32_t)sd.k0 * 4u);
    // Source rows in flight.
    // ST > 0: each thread owns a FIFO of ST 16-byte slots in shared memory; row i+ST-1 is requested (cp.async, L1 bypass)
    //   when row i is consumed, and cp.async groups are counted, so ST-1 rows stay in flight per thread at no register
    //   cost.  One thread also asks L2 for the CTA's rows ST..2ST-1 further ahead (bulk prefetch), so that the cp.async
    //   requests are L2 hits.
    // ST == 0 (the stages do not fit next to a second CTA): two register sets of PF rows.  While set A is consumed the PF
    //   loads of set B are outstanding (and vice versa).  All loads share one hardware scoreboard and a scoreboard wait
    //   drains every load issued before it, so the next set is only requested after the current one has landed.
    // Past the last row of the band the last row is requested again (never consumed).
    uint4 pf[2][ST > 0 ? 1 : PF];
    const uint32_t st_base = sb + SM::kStageOff + (uint32_t)t * 16u;
    auto request_row = [&](const int stage) {
        cp_async16_to(st_base + stage * SM::kRowStage, pnext);
        cp_async_commit();
        if (left > 1) { pnext += stride; --left; }
    };
    auto request_set = [&](uint4 (&dst)[ST > 0 ? 1 : PF]) {
        if (left > PF) {
#pragma unroll
            for (int i = 0; i < PF; ++i) { dst[i] = __ldcs(reinterpret_cast<const uint4*>(pnext)); pnext += stride; }
            left -= PF;
            if (t == 0 && left >= PF) {                   // the set after this one: on its way into L2
#pragma unroll
                for (int i = 0; i < PF; ++i) l2_prefetch_bulk(pnext + (size_t)i * stride, seg_bytes);
            }
        } else {
#pragma unroll
            for (int i = 0; i < PF; ++i) {
                dst[i] = __ldcs(reinterpret_cast<const uint4*>(pnext));
                if (left > 1) { pnext += stride; --left; }
            }
        }
    };
    if (ST > 0) {
#pragma unroll
        for (int i = 0; i + 1 < ST; ++i) request_row(i);
    } else {
        request_set(pf[0]);
    }
    // sRGB bytes -> working floats: window address = (byte << 8) | (lane << 2), one PRMT per lookup
    auto gather = [&](const uint4& raw, float (&p)[NV]) {
        const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t v = w4[i];
            p[0 * 4 + i] = lds_table(__byte_perm(v, lane4, 0x6504) + kSmemWindowBase);
            p[1 * 4 + i] = lds_table(__byte_perm(v, lane4, 0x6514) + kSmemWindowBase);
            p[2 * 4 + i] = lds_table(__byte_perm(v, lane4, 0x6524) + kSmemWindowBase);
            if (CH == 4) {
                // alpha table entry == a * (1/255f) (color.rs:38): computed, not gathered
                const float af = __fmul_rn(__uint2float_rn(v >> 24), 1.0f / 255.0f);
                p[0 * 4 + i] = __fmul_rn(p[0 * 4 + i], af);
                p[1 * 4 + i] = __fmul_rn(p[1 * 4 + i], af);
                p[2 * 4 + i] = __fmul_rn(p[2 * 4 + i], af);
                p[(CH - 1) * 4 + i] = af;
            }
        }
    };
    float pq[GA ? 2 : 1][NV];                       // GA: working floats of the row being accumulated and of the next one
    if (GA) {                                       // row 0 has landed once all but the newest ST-2 requests have
        cp_async_wait_group<(ST > 1 ? ST - 2 : 0)>();
        gather(lds_u32x4(st_base), pq[0]);
    }
    int ring_pos = 0;
    uint32_t buf = 0;                               // partial buffer (and mbarrier) of the next emitted row == nrow & 1
    uint32_t nrow = 0;                              // rows emitted so far
    int yprev = 0;                                  // the emitted row that is not finished yet (nrow > 0)
    const uint32_t bar_base = sb + SM::kBarOff;

    // H pass, second half: thread u adds the partials of output column X0+u in ascending group order, then the store epilogue
    auto finish = [&](const int y, const uint32_t b) {
        uint32_t a = rd_base + b * SM::kPartBuf;
        float F[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        int n = (int)rd_groups;
        for (; n >= 4; n -= 4, a += 16) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int c = 0; c < CH; ++c) F[c] = __fadd_rn(F[c], lds_f32(a + (c * SH * NT + g) * 4));
        }
        if (n & 2) {
#pragma unroll
            for (int g = 0; g < 2; ++g)
#pragma unroll
                for (int c = 0; c < CH; ++c) F[c] = __fadd_rn(F[c], lds_f32(a + (c * SH * NT + g) * 4));
            a += 8;
        }
        if (n & 1) {
#pragma unroll
            for (int c = 0; c < CH; ++c) F[c] = __fadd_rn(F[c], lds_f32(a + (c * SH * NT) * 4));
        }
        uint8_t* dst = out_col + (size_t)y * out_stride;
        *reinterpret_cast<uint32_t*>(dst) = finish_pixel<SIMPLE>(F[0], F[1], F[2], CH == 4 ? F[3] : 0.0f, flags, job, tb, dst);
    };

    for (int c0 = 0; c0 < total_rows; c0 += kProgChunk) {
        const int chunk = c0 / kProgChunk;
        {   // stream the next program chunk while this one is consumed
            const int rows_next = min(kProgChunk, total_rows - (c0 + kProgChunk));
            if (rows_next > 0) {
                uint32_t* dst = sProg + ((chunk + 1) & 1) * kProgChunk * PL::kWords;
                const uint32_t* srcp = gprog + (size_t)(c0 + kProgChunk) * PL::kWords;
                for (int i = t; i < rows_next * PL::kWords / 4; i += NT) cp_async16(dst + i * 4, srcp + i * 4);
            }
        }
        uint32_t pa = sb + SM::kProgOff + (uint32_t)(chunk & 1) * (kProgChunk * kRec);      // next program record
        const uint32_t pa_end = pa + (uint32_t)min(kProgChunk, total_rows - c0) * kRec;
        // One source row (ring position D of RING): fetch, LUT-convert, accumulate.
        // Returns the completion word of the row (0 = no output row completes here).
        auto do_row = [&](auto dtag) -> uint32_t {
            constexpr int D = decltype(dtag)::value;
            uint4 raw;
            if (ST > 0) {
                if (D == 0 && t == 0 && left > 2 * ST) {              // L2 prefetch of the ST rows after the ST next requests
#pragma unroll
                    for (int i = 0; i < ST; ++i) l2_prefetch_bulk(pnext + (size_t)(ST + i) * stride, seg_bytes);
                }
                request_row((D + ST - 1) % (ST > 0 ? ST : 1));        // row i+ST-1 into the slot consumed one row ago
                if (GA) {
                    cp_async_wait_group<(ST > 1 ? ST - 2 : 0)>();     // all but the newest ST-2 requests have landed: row i+1 is here
                    raw = lds_u32x4(st_base + ((D + 1) % (ST > 0 ? ST : 1)) * SM::kRowStage);
                } else {
                    cp_async_wait_group<(ST > 0 ? ST - 1 : 0)>();     // all but the newest ST-1 requests have landed: row i is here
                    raw = lds_u32x4(st_base + D * SM::kRowStage);
                }
            } else {
                constexpr int SET = D / PF, IDX = D % PF;
                raw = pf[SET][IDX];
                if (IDX == 0) {                   // set SET has landed: request the other set
                    // pl.zero is 0 at run time; tying the address to the data just consumed keeps the compiler from
                    // issuing these loads ahead of the scoreboard wait for the current set (see the comment at pf[]).
                    pnext = pnext + (raw.x & pl.zero);
                    request_set(pf[SET ^ 1]);
                }
            }
            // ---- program record: AV slot weights, then the completion word
            uint32_t rec[PL::kWords];
            {
                const uint4 q = lds_u32x4(pa);
                rec[0] = q.x; rec[1] = q.y; rec[2] = q.z; rec[3] = q.w;
                if (PL::kWords > 4) {
                    if (AV + 1 - 4 == 1) rec[4] = lds_u32(pa + 16);
                    else { const uint4 q2 = lds_u32x4(pa + 16); rec[4] = q2.x; rec[5] = q2.y; rec[6] = q2.z; rec[7] = q2.w; }
                }
                pa += kRec;
            }
            // ---- sRGB bytes -> working floats (GA: of the NEXT row; this row's were gathered one row ago)
            float (&p)[NV] = pq[GA ? (D & 1) : 0];
            gather(raw, pq[GA ? ((D + 1) & 1) : 0]);
            // ---- ring accumulate (packed fp32 FMA: two IEEE fmaf per instruction)
#pragma unroll
            for (int s = 0; s < AV; ++s) {
                const float ws = __uint_as_float(rec[s]);
                const float2 w2 = make_float2(ws, ws);          // becomes a scalar (.F32) operand of FFMA2
#pragma unroll
                for (int k = 0; k < NV / 2; ++k) {
                    const float2 r2 = __ffma2_rn(w2, make_float2(p[2 * k], p[2 * k + 1]), make_float2(acc[s][2 * k], acc[s][2 * k + 1]));
                    acc[s][2 * k] = r2.x; acc[s][2 * k + 1] = r2.y;
                }
            }
            return rec[PL::kDone];
        };
        // The RING ring positions are copies of do_row; the completion code below exists once: a row that completes
        // output rows breaks out of the switch, and the loop re-enters at the next ring position.
        while (pa != pa_end) {
            uint32_t dn = 0;
            switch (ring_pos) {
#define IFB_ROW_CASE(D_) \
            case D_: if (D_ < RING) { dn = do_row(std::integral_constant<int, (D_) % RING>{}); ring_pos = ((D_) + 1) % RING; if (dn || pa == pa_end) break; }
            IFB_ROW_CASE(0) IFB_ROW_CASE(1) IFB_ROW_CASE(2) IFB_ROW_CASE(3) IFB_ROW_CASE(4) IFB_ROW_CASE(5) IFB_ROW_CASE(6) IFB_ROW_CASE(7)
            IFB_ROW_CASE(8) IFB_ROW_CASE(9) IFB_ROW_CASE(10) IFB_ROW_CASE(11)
#undef IFB_ROW_CASE
            default: ring_pos = 0; break;
            }
            // ---- completed output rows: consecutive y, consecutive slots
            const int ndone = dn & 0xfu;
            int slot = (dn >> 4) & 0xfu;
            // H pass, first half: the finished V row (ring slot `v`) times this thread's H weights, by partial plane (output
            // column mod SH): planes (2k, 2k+1) as float2 pairs -- FFMA2 takes the V value as a broadcast scalar operand --
            // and an odd last plane as two float2.  One copy of this code per ring slot: the row is consumed where it lies.
            auto emit = [&](float (&v)[NV]) {
                const uint32_t wr = wr_base + buf * SM::kPartBuf;
#pragma unroll
                for (int qp = 0; qp < SH / 2; ++qp) {
                    float2 h[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) h[i] = lds_f32x2(hw_base + (qp * 4 + i) * 16 * NT);
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        float2 ps = make_float2(0.0f, 0.0f);
#pragma unroll
                        for (int i = 0; i < 4; ++i) { const float x = v[c * 4 + i]; ps = __ffma2_rn(h[i], make_float2(x, x), ps); }
                        sts_f32(wr + ((c * SH + 2 * qp) * NT) * 4, ps.x);
                        sts_f32(wr + ((c * SH + 2 * qp + 1) * NT) * 4, ps.y);
                    }
                }
                if (SH & 1) {
                    const float2 h01 = lds_f32x2(hw_base + ((SH / 2) * 4 + 0) * 16 * NT);
                    const float2 h23 = lds_f32x2(hw_base + ((SH / 2) * 4 + 1) * 16 * NT);
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        float ps = __fmaf_rn(h01.x, v[c * 4 + 0], 0.0f);
                        ps = __fmaf_rn(h01.y, v[c * 4 + 1], ps);
                        ps = __fmaf_rn(h23.x, v[c * 4 + 2], ps);
                        ps = __fmaf_rn(h23.y, v[c * 4 + 3], ps);
                        sts_f32(wr + ((c * SH + SH - 1) * NT) * 4, ps);
                    }
                }
#pragma unroll
                for (int k = 0; k < NV; ++k) v[k] = 0.0f;             // the slot is free for output row y + AV
            };
            for (int e = 0; e < ndone; ++e) {
                const int y = (int)(dn >> 8) + e;
                const int cur = slot;
                slot = slot + 1 == AV ? 0 : slot + 1;
                if (y >= bd.Y0 && y < bd.Y1) {
                    // Finish the row emitted at the previous completion: every warp parked its partials a row ago, so the
                    // wait rarely blocks, and warps only need to stay within one output row of each other.
                    if (nrow) {
                        mbar_wait(bar_base + (buf ^ 1u) * 8u, ((nrow - 1u) >> 1) & 1u);
                        if ((fin >> (buf ^ 1u)) & 1u) finish(yprev, buf ^ 1u);
                    }
                    switch (cur) {
#define IFB_SLOT_CASE(S_) case S_: if (S_ < AV) emit(acc[(S_) % AV]); break;
                    IFB_SLOT_CASE(0) IFB_SLOT_CASE(1) IFB_SLOT_CASE(2) IFB_SLOT_CASE(3) IFB_SLOT_CASE(4) IFB_SLOT_CASE(5)
#undef IFB_SLOT_CASE
                    default: break;
                    }
                    __syncwarp();
                    if ((t & 31) == 0) mbar_arrive(bar_base + buf * 8u);
                    yprev = y;
                    ++nrow;
                    buf ^= 1u;
                } else {                                              // halo row of a neighbouring band: only free the slot
                    switch (cur) {
#define IFB_SLOT_CASE(S_) case S_: if (S_ < AV) { _Pragma("unroll") for (int k = 0; k < NV; ++k) acc[(S_) % AV][k] = 0.0f; } break;
                    IFB_SLOT_CASE(0) IFB_SLOT_CASE(1) IFB_SLOT_CASE(2) IFB_SLOT_CASE(3) IFB_SLOT_CASE(4) IFB_SLOT_CASE(5)
#undef IFB_SLOT_CASE
                    default: break;
                    }
                }
            }
        }
        // the next program chunk was requested a whole chunk of rows ago: with staged rows its cp.async group is long
        // complete (groups retire in order, wait_group<ST-1> ran every row); only its visibility to the other threads is needed
        if (ST == 0 || c0 + kProgChunk >= total_rows) cp_async_wait_all();
        __syncthreads();
    }
    if (nrow) {                                                       // the last emitted row of the band
        mbar_wait(bar_base + (buf ^ 1u) * 8u, ((nrow - 1u) >> 1) & 1u);
        if ((fin >> (buf ^ 1u)) & 1u) finish(yprev, buf ^ 1u);
    }
}

}  // namespace ifbk
