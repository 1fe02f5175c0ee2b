// Shared helpers for the sm_100a kernels of libdynamask_sm100.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dynamask_sm100.h"

namespace dm {

// ---- error plumbing (thread-local text only; no other global state) ------------------------
void set_cuda_error(cudaError_t e, const char* where);
void count_launch(int n = 1);

#define DM_CUDA_CHECK(expr, where)                      \
    do {                                                \
        cudaError_t _e = (expr);                        \
        if (_e != cudaSuccess) {                        \
            ::dm::set_cuda_error(_e, where);            \
            return DM_ECUDA;                            \
        }                                               \
    } while (0)

#define DM_LAUNCH_CHECK(where)                          \
    do {                                                \
        ::dm::count_launch();                           \
        cudaError_t _e = cudaGetLastError();            \
        if (_e != cudaSuccess) {                        \
            ::dm::set_cuda_error(_e, where);            \
            return DM_ECUDA;                            \
        }                                               \
    } while (0)

int sm_count();  // cached per device, read-only after first query

// ---- RoI geometry shared by forward / backward ----------------------------------------------
// Mirrors the setup of the aligned avg-pool RoIAlign.
// mode 0: RoIAlign (sampling grid, clamped border).  mode 1 / 2: SimpleRoIAlign point sampling
// (one sample per bin, zero padding; 1 = align_corners False, 2 = align_corners True), in which
// case (rsw, rsh) is the RoI corner and (bw, bh) its extent in input-image pixels.
struct RoiGeom {
    float rsw, rsh, bw, bh;
    int gw, gh;
    int mode;
    float scale;
};

__device__ __forceinline__ RoiGeom roi_geom(const float* __restrict__ roi5, float scale, int ph,
                                            int pw, int sampling_ratio, int aligned) {
    RoiGeom g;
    g.mode = 0;
    g.scale = scale;
    const float off = aligned ? 0.5f : 0.0f;
    g.rsw = __fsub_rn(__fmul_rn(roi5[1], scale), off);
    g.rsh = __fsub_rn(__fmul_rn(roi5[2], scale), off);
    const float rew = __fsub_rn(__fmul_rn(roi5[3], scale), off);
    const float reh = __fsub_rn(__fmul_rn(roi5[4], scale), off);
    float rw = __fsub_rn(rew, g.rsw), rh = __fsub_rn(reh, g.rsh);
    if (!aligned) {
        rw = fmaxf(rw, 1.0f);
        rh = fmaxf(rh, 1.0f);
    }
    g.bh = __fdiv_rn(rh, (float)ph);
    g.bw = __fdiv_rn(rw, (float)pw);
    g.gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)ph));
    g.gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)pw));
    return g;
}

// sample coordinate of bin p, sub-sample i:  (start + p*bin) + ((i+.5)*bin)/grid
__device__ __forceinline__ float sample_coord(float start, float bin, int grid, int p, int i) {
    return __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)),
                     __fdiv_rn(__fmul_rn(__fadd_rn((float)i, 0.5f), bin), (float)grid));
}

// One axis of a bilinear tap: returns false when the coordinate is rejected (< -1 or > size).
__device__ __forceinline__ bool axis_tap(float v, int size, int& lo, int& hi, float& l, float& h) {
    if (v < -1.0f || v > (float)size) return false;
    if (v <= 0.0f) v = 0.0f;
    lo = (int)v;
    if (lo >= size - 1) {
        hi = lo = size - 1;
        v = (float)lo;
    } else {
        hi = lo + 1;
    }
    l = __fsub_rn(v, (float)lo);
    h = __fsub_rn(1.0f, l);
    return true;
}

// ---- SimpleRoIAlign (mmcv.ops.point_sample.SimpleRoIAlign as used by SFMStage,
// mmdet/models/roi_heads/mask_heads/dynamask_head.py:74,104-105) ---------------------------------
// One grid_sample point per bin.  The coordinate follows the fp32 op sequence of the mmcv helpers
// (generate_grid -> rel_roi_point_to_rel_img_point -> point_sample -> grid_sample unnormalise).
__device__ __forceinline__ RoiGeom point_geom(const float* __restrict__ roi5, float scale, int aligned) {
    RoiGeom g;
    g.mode = aligned ? 1 : 2;
    g.scale = scale;
    g.rsw = roi5[1];
    g.rsh = roi5[2];
    g.bw = __fsub_rn(roi5[3], roi5[1]);
    g.bh = __fsub_rn(roi5[4], roi5[2]);
    g.gw = g.gh = 1;
    return g;
}

// pixel coordinate of bin p's point on an axis of `size` pixels, P bins
__device__ __forceinline__ float point_coord(float corner, float extent, float scale, int mode, int P, int size, int p) {
    // generate_grid: F.affine_grid's base grid is linspace(-1, 1, P) * (P - 1) / P (evaluated from
    // whichever end is nearer, as ATen's linspace does), then normalised to [0, 1]
    float lin = -1.0f;
    if (P > 1) {
        const float step = __fdiv_rn(2.0f, (float)(P - 1));
        lin = p < P / 2 ? __fadd_rn(-1.0f, __fmul_rn(step, (float)p))
                        : __fsub_rn(1.0f, __fmul_rn(step, (float)(P - p - 1)));
    }
    const float rel = __fdiv_rn(__fadd_rn(__fdiv_rn(__fmul_rn(lin, (float)(P - 1)), (float)P), 1.0f), 2.0f);
    const float ab = __fadd_rn(__fmul_rn(rel, extent), corner);                    // rel -> abs image
    const float q = __fmul_rn(__fdiv_rn(ab, (float)size), scale);                  // abs -> rel image
    const float gc = __fsub_rn(__fmul_rn(q, 2.0f), 1.0f);                          // denormalize
    if (mode == 1) return __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gc, 1.0f), (float)size), 1.0f), 2.0f);
    return __fmul_rn(__fdiv_rn(__fadd_rn(gc, 1.0f), 2.0f), (float)(size - 1));
}

// bilinear taps with zero padding: an out-of-range tap keeps weight 0 and a clamped index
__device__ __forceinline__ bool point_tap(float v, int size, int& lo, int& hi, float& l, float& h) {
    if (!(v > -1.0f && v < (float)size)) return false;
    const float fl = floorf(v);
    lo = (int)fl;
    hi = lo + 1;
    l = __fsub_rn(v, fl);
    h = __fsub_rn(1.0f, l);
    if (lo < 0) { lo = 0; h = 0.0f; }
    if (hi > size - 1) { hi = size - 1; l = 0.0f; }
    return true;
}

// One axis of sample (bin p, sub-sample i) under either mode.  axis 0 = x, 1 = y.
__device__ __forceinline__ bool geom_tap(const RoiGeom& g, int axis, int P, int size, int p, int i,
                                         int& lo, int& hi, float& l, float& h) {
    const float start = axis ? g.rsh : g.rsw, bin = axis ? g.bh : g.bw;
    if (g.mode == 0) return axis_tap(sample_coord(start, bin, axis ? g.gh : g.gw, p, i), size, lo, hi, l, h);
    return point_tap(point_coord(start, bin, g.scale, g.mode, P, size, p), size, lo, hi, l, h);
}

// Fire-and-forget reduction into GLOBAL memory (SASS: RED.E.ADD.F32).  `atomicAdd` on a pointer whose
// address space the compiler cannot prove (anything that crossed a __noinline__ call) compiles to a
// generic ATOM that returns a predicate plus shared / local fall-back code the warp has to wait for.
__device__ __forceinline__ void red_add(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "f"(v) : "memory");
}

// Same, predicated inside the asm block: no branch (BSSY / BRA / BSYNC) around the reduction.
__device__ __forceinline__ void red_add_if(float* p, float v, bool on) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.s32 q, %2, 0;\n@q red.global.add.f32 [%0], %1;\n}"
                 ::"l"(__cvta_generic_to_global(p)), "f"(v), "r"((int)on) : "memory");
}

// Same, the target given as a global base address plus a 32-bit float offset: the address is one wide
// multiply-add inside the asm block (no generic-to-global conversion, no 64-bit pointer to carry).
__device__ __forceinline__ void red_add_off(unsigned long long gbase, int off, float v) {
    asm volatile("{\n.reg .u64 a;\nmad.wide.s32 a, %1, 4, %0;\nred.global.add.f32 [a], %2;\n}"
                 ::"l"(gbase), "r"(off), "f"(v) : "memory");
}

// ... and predicated inside the asm block (no branch around the reduction)
__device__ __forceinline__ void red_add_off_if(unsigned long long gbase, int off, float v, int on) {
    asm volatile("{\n.reg .pred q;\n.reg .u64 a;\nsetp.ne.s32 q, %3, 0;\nmad.wide.s32 a, %1, 4, %0;\n@q red.global.add.f32 [a], %2;\n}"
                 ::"l"(gbase), "r"(off), "f"(v), "r"(on) : "memory");
}

// Four reductions under ONE predicate: elements off, off + sc, off + 2 sc, off + 3 sc (float offsets)
// from a global base address -- the same pixel of four consecutive channels.
__device__ __forceinline__ void red4_off_if(unsigned long long gbase, int off, int sc, float v0, float v1, float v2,
                                            float v3, int on) {
    asm volatile(
        "{\n.reg .pred q;\n.reg .u64 a0, a1, a2, a3;\n.reg .s32 o1, o2, o3;\n"
        "setp.ne.s32 q, %8, 0;\n"
        "add.s32 o1, %1, %2;\nadd.s32 o2, o1, %2;\nadd.s32 o3, o2, %2;\n"
        "mad.wide.s32 a0, %1, 4, %0;\nmad.wide.s32 a1, o1, 4, %0;\nmad.wide.s32 a2, o2, 4, %0;\nmad.wide.s32 a3, o3, 4, %0;\n"
        "@!q bra RED4_SKIP;\n"
        "red.global.add.f32 [a0], %3;\nred.global.add.f32 [a1], %4;\n"
        "red.global.add.f32 [a2], %5;\nred.global.add.f32 [a3], %6;\n"
        "RED4_SKIP:\n}"
        ::"l"(gbase), "r"(off), "r"(sc), "f"(v0), "f"(v1), "f"(v2), "f"(v3), "r"(0), "r"(on) : "memory");
}

__device__ __forceinline__ void st_stream(float4* p, const float4& v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float2* p, const float2& v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float* p, const float& v) { __stcs(p, v); }

}  // namespace dm
