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
struct RoiGeom {
    float rsw, rsh, bw, bh;
    int gw, gh;
};

__device__ __forceinline__ RoiGeom roi_geom(const float* __restrict__ roi5, float scale, int ph,
                                            int pw, int sampling_ratio, int aligned) {
    RoiGeom g;
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

__device__ __forceinline__ void st_stream(float4* p, const float4& v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float2* p, const float2& v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float* p, const float& v) { __stcs(p, v); }

}  // namespace dm
