// Evaluates the product's block-scaler tables (imageflow_b200/csrc/ifb_idct_tables.inc, the data the GPU kernel uses) on the CPU with the
// dense formula of idct_block_scale_kernel, so that the extraction of the reference's literals can be checked without a GPU
// against outputs of the reference itself (tests/golden/idct_golden.npz, oracle/_ref).  TEST INFRASTRUCTURE.
#include <cstddef>
#include <cstdint>
#include "../../imageflow_b200/csrc/ifb_idct_tables.inc"
extern "C" void idct_tables_eval(const uint8_t* in, uint32_t in_stride, uint32_t blocks_x, uint32_t blocks_y, uint8_t* out, uint32_t out_stride, int n, int srgb) {
    const IdctScaler& sc = kIdctScalers[srgb ? 1 : 0][n - 1];
    for (uint32_t by = 0; by < blocks_y; ++by)
        for (uint32_t bx = 0; bx < blocks_x; ++bx)
            for (int y = 0; y < n; ++y)
                for (int x = 0; x < n; ++x) {
                    int32_t sum = sc.bias[y][x];
                    for (int i = 0; i < 8; ++i) {
                        int32_t v = 0;
                        for (int j = 0; j < 8; ++j) {
                            const uint8_t s = in[(size_t)(by * 8 + j) * in_stride + bx * 8 + i];
                            v += sc.wv[y][j] * (srgb ? (int32_t)kIdct_lut_srgb_to_linear[s] : (int32_t)s);
                        }
                        sum += sc.wh[x][i] * v;
                    }
                    const int32_t q = sum >> sc.shift[y][x];
                    out[(size_t)(by * n + y) * out_stride + bx * n + x] =
                        sum < 0 ? (uint8_t)0 : (sum >= sc.sat[y][x] ? (uint8_t)255 : (srgb ? kIdct_lut_linear_to_srgb[q] : (uint8_t)q));
                }
}
