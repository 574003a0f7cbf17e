// Micro-benchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fma ubench_fma.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NACC = 24;     // independent chains per thread
constexpr int ITER = 4096;

__global__ void k_ffma(float* out, float w0, float w1) {
    float a[NACC];
    for (int i = 0; i < NACC; ++i) a[i] = threadIdx.x * 1e-3f + i;
    float x = w0, y = w1;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) a[i] = __fmaf_rn(x, a[i], y);
    }
    float s = 0; for (int i = 0; i < NACC; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float w0, float w1) {
    float2 a[NACC / 2];
    for (int i = 0; i < NACC / 2; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float2 x = make_float2(w0, w0), y = make_float2(w1, w1);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NACC / 2; ++i) a[i] = __ffma2_rn(x, a[i], y);
    }
    float s = 0; for (int i = 0; i < NACC / 2; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// accumulate form: acc = w * p + acc with distinct p per chain (like the resample inner loop)
__global__ void k_ffma_acc(float* out, const float* __restrict__ in, float w0) {
    float a[NACC], p[NACC];
    for (int i = 0; i < NACC; ++i) { a[i] = 0.f; p[i] = in[threadIdx.x + i]; }
    float w[4] = {w0, w0 * 0.5f, w0 * 0.25f, w0 * 0.125f};
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int i = 0; i < NACC / 4; ++i) a[s * (NACC / 4) + i] = __fmaf_rn(w[s], p[i], a[s * (NACC / 4) + i]);
    }
    float s = 0; for (int i = 0; i < NACC; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2_acc(float* out, const float* __restrict__ in, float w0) {
    float2 a[NACC / 2], p[NACC / 8];
    for (int i = 0; i < NACC / 2; ++i) a[i] = make_float2(0.f, 0.f);
    for (int i = 0; i < NACC / 8; ++i) p[i] = make_float2(in[threadIdx.x + 2 * i], in[threadIdx.x + 2 * i + 1]);
    float2 w[4] = {make_float2(w0, w0), make_float2(w0 * .5f, w0 * .5f), make_float2(w0 * .25f, w0 * .25f), make_float2(w0 * .125f, w0 * .125f)};
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int i = 0; i < NACC / 8; ++i) a[s * (NACC / 8) + i] = __ffma2_rn(w[s], p[i], a[s * (NACC / 8) + i]);
    }
    float s = 0; for (int i = 0; i < NACC / 2; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}
int main() {
    float *out, *in; cudaMalloc(&out, 148 * 8 * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4);
    for (int warps : {4, 8, 16, 32}) {
        dim3 g(148 * 2), b(warps * 16);   // 2 CTAs/SM
        double fl = 2.0 * 148 * 2 * warps * 16 * (double)NACC * ITER;
        float t1 = timeit([&] { k_ffma<<<g, b>>>(out, 0.999f, 0.001f); });
        float t2 = timeit([&] { k_ffma2<<<g, b>>>(out, 0.999f, 0.001f); });
        float t3 = timeit([&] { k_ffma_acc<<<g, b>>>(out, in, 0.5f); });
        float t4 = timeit([&] { k_ffma2_acc<<<g, b>>>(out, in, 0.5f); });
        printf("warps/SM %2d: FFMA %.3f ms %.1f TF | FFMA2 %.3f ms %.1f TF | FFMA-acc %.3f ms %.1f TF | FFMA2-acc %.3f ms %.1f TF\n", warps,
               t1, fl / t1 * 1e-9, t2, fl / t2 * 1e-9, t3, fl / t3 * 1e-9, t4, fl / t4 * 1e-9);
    }
    return 0;
}
