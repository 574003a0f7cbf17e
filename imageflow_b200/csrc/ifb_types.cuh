// ifb_types.cuh -- descriptors the kernels read; included by ifb_kernels.cuh inside namespace ifbk (product code).
// Plain structs only, so that tests/cpu_emu can include this file with a host compiler.
// ---------------------------------------------------------------- device-side descriptors
struct JobDev {                 // one scale_and_render call
    const uint8_t* in;          // input window origin
    uint8_t* out;               // canvas origin already offset to (x, y)
    uint32_t in_stride, out_stride;
    uint32_t flags;             // bit0 linear, bit1 alpha_meaningful, bits2-3 compose, bit4 has colour matrix
    uint32_t in_xoff;           // ring kernel: pixels between the 16-byte aligned base of its TMA descriptor and `in` (0..3)
    float matte[4];             // premultiplied working-space matte (B,G,R,A positional; scaling.rs:141-143)
    float cm[20];               // cm[c*5 + k]: output channel c (0=r,1=g,2=b,3=a) = sum_k cm[c*5+k]*{r,g,b,a,1} (bias already *255)
};
enum : uint32_t { JF_LINEAR = 1u, JF_ALPHA = 2u, JF_COMPOSE_SHIFT = 2, JF_CM = 16u };

// ---------------------------------------------------------------- scalar helper shared by every store epilogue
// color.rs:101-108 uchar_clamp_ff: trunc(x + 0.5) computed exactly, saturated to [0,255], NaN -> 0
__device__ __forceinline__ uint32_t uchar_clamp_ff(float x) {
    if (!(x > 0.0f)) return 0u;
    if (x >= 255.0f) return 255u;
    const float fl = floorf(x);
    const float fr = x - fl;                 // exact
    return (uint32_t)fl + (fr >= 0.5f ? 1u : 0u);
}

struct Tables {                 // per-device constant tables
    const float* t_lin;         // ColorContext::byte_to_float, LinearRGB (color.rs:23-48)
    const float* t_srgb;        // same, StandardRGB (== v * (1/255f)); also the alpha table
    const uint8_t* lut16k;      // LINEAR_TO_SRGB_LUT (lut.rs:14)
};

struct AxisDev {                // CSR contribution windows of one axis (weights.rs PixelRowWeights)
    const uint32_t* left; const uint32_t* right; const uint32_t* off; const float* w;
};

// tile kernels: one CTA works on tow x toh output pixels at a time
struct TilePlanDev {
    uint32_t in_w, in_h, out_w, out_h;
    int tow, toh;               // tile size in output pixels
    int tiles_x, tiles_y;
    int max_ic, max_ir;         // largest source extent of any tile (shared-memory tile dimensions)
    int h4;                     // every H window has at most four taps: hw holds them padded to four per output column
    const float* vw;            // [tiles_y * toh][8]: V window of an output row padded with zeros to the six source rows of its quad
    const uint32_t* vq;         // [tiles_y * toh / 4]: first of the quad's six source rows + 1, or 0 if its windows do not fit
    const float* hw;            // [tiles_x * tow][4] (h4 only)
};
