// ifb_types.cuh -- descriptors the kernels read; included by ifb_kernels.cuh inside namespace ifbk (product code).
// Plain structs only, so that tests/cpu_emu can include this file with a host compiler.
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
    uint32_t zero;              // always 0 (run-time constant used to order loads after a scoreboard wait)
    const uint32_t* vprog;      // [in_h][ProgLayout::kWords]: weight (float bits) of the output row in each ring slot (row y lives in
                                //   slot y mod AV), then ((first completed y << 8) | (its slot << 4) | count)
    const StripDev* strips;
    const BandDev* bands;
    const float* hw;            // [strip][SH*4*NT]: H weights of thread t by partial plane p = output column mod SH and own column i:
                                //   p < 2*(SH/2): word 2*(((p/2)*4+i)*NT+t) + (p&1); odd last plane: word 2*(((SH/2)*4+i/2)*NT+t) + (i&1)
    const uint32_t* hrd;        // [strip][NT] reader u: (first contributing thread) | (count << 12) | ((X mod SH) << 28)
};

// tile kernels: one CTA works on tow x toh output pixels at a time
struct TilePlanDev {
    uint32_t in_w, in_h, out_w, out_h;
    int tow, toh;               // tile size in output pixels
    int tiles_x, tiles_y;
    int max_ic, max_ir;         // largest source extent of any tile (shared-memory tile dimensions)
};
