// ifb_whitespace_kernel.cuh -- included by ifb_kernels.cuh inside namespace ifbk (product code, sm_100a).
// Uses only: threadIdx, blockIdx, __shared__, __syncthreads, __ldg, min, max, abs -- tests/test_whitespace_product.py compiles
// this file with g++ under a sequential emulation of those and checks the codes against the CPU oracle.
// graphics/whitespace.rs: the part of detect_content that is the same for every window -- approximate_grayscale (:426-523,
// Bgra32 when alpha is meaningful, else Bgr32) and the per-centre half of sobel_scharr_detect (:525-613) -- for every pixel at
// once.  One byte per pixel (layout in ifb_whitespace.h); the reference's order-dependent window walk is replayed over this
// map on the host (ifb_whitespace.cc).  32 x 8 pixels per CTA through a 34 x 10 grayscale tile; 4 bytes read + 1 written per pixel.
__device__ __forceinline__ uint32_t ws_gray(uint32_t bgra, bool alpha_meaningful) {
    const uint32_t lum = 233u * (bgra & 0xffu) + 1197u * ((bgra >> 8) & 0xffu) + 610u * ((bgra >> 16) & 0xffu);
    if (!alpha_meaningful) return lum / 2048u;                                    // Bgr32
    const uint32_t v = lum * (bgra >> 24);                                        // Bgra32: weighed by alpha, rounded up, clamped
    return min((v + 524287u) / 524288u, 255u);
}
__global__ void __launch_bounds__(256) whitespace_codes_kernel(const uint8_t* __restrict__ px, uint32_t w, uint32_t h, uint32_t stride,
                                                               uint32_t alpha_meaningful, int threshold, uint8_t* __restrict__ codes, uint32_t row0) {
    // row0: first image row of this launch (grid.y is bounded, bitmaps are not: tall ones take several launches)
    __shared__ uint8_t g[10][36];
    const int x0 = (int)blockIdx.x * 32 - 1, y0 = (int)(row0 + blockIdx.y * 8) - 1;       // image position of g[0][0]
    for (int i = threadIdx.y * 32 + threadIdx.x; i < 10 * 34; i += 256) {
        const int ty = i / 34, tx = i - ty * 34, x = x0 + tx, y = y0 + ty;
        uint32_t v = 0;
        if (x >= 0 && y >= 0 && x < (int)w && y < (int)h) v = ws_gray(__ldg(reinterpret_cast<const uint32_t*>(px + (size_t)y * stride) + x), alpha_meaningful != 0);
        g[ty][tx] = (uint8_t)v;
    }
    __syncthreads();
    const int x = (int)blockIdx.x * 32 + threadIdx.x, y = (int)(row0 + blockIdx.y * 8) + threadIdx.y;
    if (x >= (int)w || y >= (int)h) return;
    uint32_t code = 0xFFu;
    if (x > 0 && y > 0 && x + 1 < (int)w && y + 1 < (int)h) {
        int m[3][3];
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 3; ++i) m[j][i] = g[threadIdx.y + j][threadIdx.x + i];
        const int gx = 3 * m[0][0] + 10 * m[1][0] + 3 * m[2][0] - 3 * m[0][2] - 10 * m[1][2] - 3 * m[2][2];
        const int gy = 3 * m[0][0] + 10 * m[0][1] + 3 * m[0][2] - 3 * m[2][0] - 10 * m[2][1] - 3 * m[2][2];
        if (abs(gx) + abs(gy) > threshold) {
            int lo_x = 2, lo_y = 2, hi_x = 1, hi_y = 1;
#pragma unroll
            for (int j = 0; j < 3; ++j) {                                         // differences along x in row j: :576-591
                const bool e01 = abs(m[j][0] - m[j][1]) > threshold, e12 = abs(m[j][1] - m[j][2]) > threshold;
                if (e01) { lo_x = min(lo_x, 1); hi_x = max(hi_x, 1); }
                if (e12) { lo_x = min(lo_x, 2); hi_x = max(hi_x, 2); }
                if (e01 || e12) { lo_y = min(lo_y, j); hi_y = max(hi_y, j + 1); }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {                                         // differences along y in column i: :592-608
                const bool e01 = abs(m[0][i] - m[1][i]) > threshold, e12 = abs(m[1][i] - m[2][i]) > threshold;
                if (e01) { lo_y = min(lo_y, 1); hi_y = max(hi_y, 1); }
                if (e12) { lo_y = min(lo_y, 2); hi_y = max(hi_y, 2); }
                if (e01 || e12) { lo_x = min(lo_x, i); hi_x = max(hi_x, i + 1); }
            }
            code = (uint32_t)lo_x | ((uint32_t)(hi_x - 1) << 2) | ((uint32_t)lo_y << 4) | ((uint32_t)(hi_y - 1) << 6);
        }
    }
    codes[(size_t)y * w + x] = (uint8_t)code;
}

