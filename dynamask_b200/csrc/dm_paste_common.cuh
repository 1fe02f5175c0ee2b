// Helpers shared by the paste kernels (dm_paste.cu) and the fused paste -> RLE kernels (dm_rle.cu):
// window of an instance, source coordinates, bilinear axis terms, sigmoid, per-pixel evaluation.
#pragma once
#include "dm_common.cuh"

namespace dm {

constexpr int kPasteThreads = 256;
#ifndef DM_BAND_ROWS
#define DM_BAND_ROWS 16
#endif
#ifndef DM_BAND_CTAS
#define DM_BAND_CTAS 8
#endif
constexpr int kBandRows = DM_BAND_ROWS;       // window rows per band of the paste kernels (C4 shape, fused launch: 32 rows 172 us,
                                              // 16 rows 167 us, 8 rows 181 us; 16 CTAs per instance 180 / 167 / 180 us)
constexpr int kRleBandRows = 32;              // canvas rows per staging band of the paste -> RLE kernels
constexpr int kBandCtas = DM_BAND_CTAS;        // CTAs per instance (grid.x); CTA b takes bands b, b + 8, ...
constexpr int kColTab = 1024;       // window columns whose x terms are staged in shared memory (8 KB)
constexpr int kVPairs = 128;        // (value, slope) pairs of one warp's y-interpolated mask row: S + 3 <= 128
constexpr int kMaskStage = 6144;    // floats of sigmoid(mask) window staged per CTA (24 KB)

struct PasteParams {
    const float* masks;
    long long stride_n, stride_c;
    const int64_t* labels;
    int N, sh, sw;
    int apply_sigmoid;
    const float* boxes;
    int img_h, img_w;
    int x_lo, y_lo, rw, rh;  // region origin and size
    long long total;         // N * rh * rw output elements
    float thr;
    void* out;
    // optional per-instance selection (switch-driven inference): instance n is pasted only when
    // select[n] == select_value; NULL pastes every instance
    const int32_t* select = nullptr;
    int select_value = 0;
};

__device__ __forceinline__ void window_1d(float lo_c, float hi_c, int S, int size, int& a, int& b) {
    const float w = hi_c - lo_c;
    // degenerate / non-finite extents: the reference's inf->0 patch makes every pixel sample the
    // mask centre, so nothing can be skipped.
    if (!(fabsf(w) >= 1e-3f) || !(fabsf(w) < 1e30f) || !(fabsf(lo_c) < 1e30f)) {
        a = 0;
        b = size;
        return;
    }
    const float lo = fminf(lo_c, hi_c), hi = fmaxf(lo_c, hi_c);
    const float margin = fabsf(w) / (2.0f * (float)S);
    const float fa = floorf(lo - margin - 0.5f) - 1.0f;
    const float fb = ceilf(hi + margin - 0.5f) + 2.0f;
    a = (int)fminf(fmaxf(fa, 0.0f), (float)size);
    b = (int)fminf(fmaxf(fb, 0.0f), (float)size);
}

// normalised -> mask-pixel coordinate of canvas pixel centre `pc` along one axis
__device__ __forceinline__ float src_coord(int pc, float c0, float c1, int S) {
    float g = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(__fadd_rn((float)pc, 0.5f), c0),
                                             __fsub_rn(c1, c0)), 2.0f), 1.0f);
    if (isinf(g)) g = 0.0f;
    return __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.0f), (float)S), 1.0f), 2.0f);
}

// One axis of the bilinear tap pair at source coordinate `i`:
//   state 0: the pixel is exactly zero (sample outside (-1, S)), 1: live, 2: NaN coordinate.
struct AxisTerm {
    int lo;        // first tap index, in [-1, S-1]
    float wl, wh;  // weights of tap lo and tap lo+1
    int state;
};

__device__ __forceinline__ AxisTerm axis_term(float i, int S) {
    AxisTerm t;
    t.lo = 0; t.wl = 0.0f; t.wh = 0.0f;
    if (i != i) { t.state = 2; return t; }
    if (!(i > -1.0f && i < (float)S)) { t.state = 0; return t; }
    const float f = floorf(i);
    t.lo = (int)f;
    t.wh = __fsub_rn(i, f);
    t.wl = __fsub_rn(__fadd_rn(f, 1.0f), i);
    t.state = 1;
    return t;
}

// staged path: ex2.approx + approximate reciprocal (~1e-7 from the exact form)
__device__ __forceinline__ float sigmoidf_fast(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }

__device__ __forceinline__ float sigmoidf_exact(float v) {
    return __frcp_rn(__fadd_rn(1.0f, expf(-v)));  // == 1 / (1 + e^-v), correctly rounded
}

// the reference's four-tap sum, in grid_sample's order
__device__ __forceinline__ float bilerp(float nw, float ne, float sw, float se, const AxisTerm& cx,
                                        const AxisTerm& ry) {
    float acc = nw * (cx.wl * ry.wl);
    acc += ne * (cx.wh * ry.wl);
    acc += sw * (cx.wl * ry.wh);
    acc += se * (cx.wh * ry.wh);
    return acc;
}

template <int MODE>
__device__ __forceinline__ uint32_t encode(float v, float thr) {
    if (MODE == DM_PASTE_BOOL) return v >= thr ? 1u : 0u;
    // reference: (val * 255).to(uint8); values are in [0,1] so the cast never saturates
    const float s = v * 255.0f;
    return (s != s) ? 0u : (uint32_t)(unsigned char)(int)s;
}

// One instance as the per-pixel path sees it: taps straight from global memory.
struct Instance {
    const float* m;
    float x0, y0, x1, y1;
    int xa, xb, ya, yb;  // conservative non-zero window, canvas coordinates
    int sh, sw, apply_sigmoid;

    __device__ __forceinline__ void load(const PasteParams& p, long long n) {
        const long long cls = p.labels ? p.labels[n] : 0;
        const float4 bx = *reinterpret_cast<const float4*>(p.boxes + 4 * (size_t)n);
        x0 = bx.x; y0 = bx.y; x1 = bx.z; y1 = bx.w;
        window_1d(x0, x1, p.sw, p.img_w, xa, xb);
        window_1d(y0, y1, p.sh, p.img_h, ya, yb);
        m = p.masks + n * p.stride_n + cls * p.stride_c;
        sh = p.sh; sw = p.sw; apply_sigmoid = p.apply_sigmoid;
    }
    __device__ __forceinline__ float tap(int y, int x) const {
        if (y < 0 || y >= sh || x < 0 || x >= sw) return 0.0f;
        const float v = __ldg(m + y * sw + x);
        return apply_sigmoid ? sigmoidf_exact(v) : v;
    }
    // canvas pixel (px, py) -> interpolated value
    __device__ float eval(int px, int py) const {
        if (px < xa || px >= xb || py < ya || py >= yb) return 0.0f;
        const AxisTerm cx = axis_term(src_coord(px, x0, x1, sw), sw);
        const AxisTerm ry = axis_term(src_coord(py, y0, y1, sh), sh);
        if (cx.state == 0 || ry.state == 0) return 0.0f;
        if (cx.state == 2 || ry.state == 2) return __int_as_float(0x7fc00000);
        return bilerp(tap(ry.lo, cx.lo), tap(ry.lo, cx.lo + 1), tap(ry.lo + 1, cx.lo), tap(ry.lo + 1, cx.lo + 1), cx, ry);
    }
};

}  // namespace dm
