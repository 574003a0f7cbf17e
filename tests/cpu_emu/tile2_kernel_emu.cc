// Runs imageflow_b200/csrc/ifb_tile2_kernel.cuh -- the product's CUDA source of fused_tile2_kernel, unmodified -- on the CPU:
// one OS thread per CUDA thread of a block, std::barrier for __syncthreads, a per-warp exchange for __reduce_max_sync, the CUDA
// arithmetic intrinsics mapped onto IEEE host operations (built with -ffp-contract=off -frounding-math).  Blocks of the grid
// run one after the other.  Dynamic shared memory is an exactly-sized heap block and every global buffer is whatever the
// caller allocated, so building this file with -fsanitize=address turns a run into a memory check of the kernel.
// TEST INFRASTRUCTURE: nothing in the product calls this; the GPU remains the place where the kernel is run and timed.
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

namespace emu {
struct Idx { unsigned x, y, z; };
static thread_local Idx threadIdx;
static Idx blockIdx, gridDim;
static unsigned char* smem;
static std::barrier<>* block_barrier;
static std::barrier<>* warp_barrier[8];
static unsigned warp_scratch[8][32];
}  // namespace emu
using emu::blockIdx;
using emu::gridDim;
#define threadIdx emu::threadIdx
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define IFB_DYNAMIC_SMEM(name_) unsigned char* const name_ = emu::smem
#define __syncthreads() emu::block_barrier->arrive_and_wait()
using std::max;
using std::min;
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct alignas(8) float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t v = ((uint64_t)y << 32) | x; uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __uint2float_rn(unsigned v) { return (float)v; }
static inline float __fadd_rz(float a, float b) {          // add.rz.f32 without touching the thread's rounding mode (the compiler
    const double s = (double)a + (double)b;                  // may move arithmetic across fesetround): the double sum is exact
    float f = (float)s;                                      // for the operands met here; then round it toward zero
    if (std::fabs((double)f) > std::fabs(s)) f = std::nextafterf(f, 0.0f);
    return f;
}
static inline unsigned __float2uint_rz(float s) {           // cvt.rzi.u32.f32: NaN -> 0, saturating
    if (!(s > 0.0f)) return 0u;
    if (s >= 4294967296.0f) return 0xffffffffu;
    return (unsigned)s;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    emu::warp_scratch[warp][lane] = v;
    emu::warp_barrier[warp]->arrive_and_wait();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m = std::max(m, emu::warp_scratch[warp][i]);
    emu::warp_barrier[warp]->arrive_and_wait();
    return m;
}

namespace ifbk {
#include "../../imageflow_b200/csrc/ifb_types.cuh"
#ifndef IFB_TILE2_MINB
#define IFB_TILE2_MINB 4
#endif
#include "../../imageflow_b200/csrc/ifb_tile2_kernel.cuh"
}  // namespace ifbk
#undef threadIdx

using namespace ifbk;
using KernelFn = void (*)(const JobDev*, uint32_t, Tables, AxisDev, AxisDev, TilePlanDev);
static KernelFn pick(int ch, int linear, int compose, int cm) {
#define IFB_PICK(CH_, LIN_, CO_, CM_) if (ch == CH_ && linear == LIN_ && compose == CO_ && cm == CM_) return fused_tile2_kernel<CH_, LIN_ != 0, CO_, CM_ != 0>;
#define IFB_PICK_CM(CH_, LIN_, CO_) IFB_PICK(CH_, LIN_, CO_, 0) IFB_PICK(CH_, LIN_, CO_, 1)
#define IFB_PICK_LIN(CH_, CO_) IFB_PICK_CM(CH_, 0, CO_) IFB_PICK_CM(CH_, 1, CO_)
    IFB_PICK_LIN(4, 0) IFB_PICK_LIN(4, 1) IFB_PICK_LIN(4, 2) IFB_PICK_LIN(3, 0)
    return nullptr;
}

extern "C" uint32_t emu_tile2_smem_bytes(int max_ir, int max_ic, int linear) { return Tile2Smem::make(max_ir, max_ic, linear != 0).total; }
extern "C" int emu_tile2_tile_h(void) { return kTile2H; }
extern "C" uint32_t emu_tile2_sizeof_jobdev(void) { return (uint32_t)sizeof(JobDev); }

// one launch of fused_tile2_kernel<ch, linear, compose, cm> with `grid` blocks of 256 threads
extern "C" int emu_tile2_launch(int ch, int linear, int compose, int cm, unsigned grid, const void* jobs, uint32_t n_jobs,
                                const float* t_lin, const float* t_srgb, const uint8_t* lut16k,
                                const uint32_t* v_left, const uint32_t* v_right, const uint32_t* v_off, const float* v_w,
                                const uint32_t* h_left, const uint32_t* h_right, const uint32_t* h_off, const float* h_w,
                                const int32_t* plan10, const float* tile_vw, const uint32_t* tile_vq, const float* tile_hw) {
    KernelFn fn = pick(ch, linear, ch == 4 ? compose : 0, cm);
    if (!fn) return 1;
    TilePlanDev pl{};
    pl.in_w = (uint32_t)plan10[0]; pl.in_h = (uint32_t)plan10[1]; pl.out_w = (uint32_t)plan10[2]; pl.out_h = (uint32_t)plan10[3];
    pl.tow = plan10[4]; pl.toh = plan10[5]; pl.tiles_x = plan10[6]; pl.tiles_y = plan10[7]; pl.max_ic = plan10[8]; pl.max_ir = plan10[9];
    pl.h4 = plan10[10]; pl.vw = tile_vw; pl.vq = tile_vq; pl.hw = tile_hw;
    const Tables tb{t_lin, t_srgb, lut16k};
    const AxisDev av{v_left, v_right, v_off, v_w}, ah{h_left, h_right, h_off, h_w};
    const size_t smem_bytes = Tile2Smem::make(pl.max_ir, pl.max_ic, linear != 0).total;
    emu::gridDim = {grid, 1, 1};
    for (unsigned b = 0; b < grid; ++b) {
        emu::blockIdx = {b, 0, 0};
        // exactly sized (Tile2Smem::total is a multiple of 16) and filled with NaN patterns: an access past the end is an
        // AddressSanitizer report, and uninitialised shared memory can never pass for a result
        std::unique_ptr<unsigned char, decltype(&std::free)> sm(static_cast<unsigned char*>(std::aligned_alloc(16, smem_bytes)), &std::free);
        std::memset(sm.get(), 0xFF, smem_bytes);
        emu::smem = sm.get();
        std::barrier<> bb(256);
        emu::block_barrier = &bb;
        std::vector<std::unique_ptr<std::barrier<>>> wb;
        for (int w = 0; w < 8; ++w) { wb.emplace_back(new std::barrier<>(32)); emu::warp_barrier[w] = wb.back().get(); }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < 256; ++t)
            th.emplace_back([=] {
                emu::threadIdx = {t, 0, 0};
                fn(static_cast<const JobDev*>(jobs), n_jobs, tb, av, ah, pl);
            });
        for (auto& x : th) x.join();
    }
    return 0;
}
