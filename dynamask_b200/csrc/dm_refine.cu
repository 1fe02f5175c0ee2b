// Next row: the inference-time stage-to-stage refinement of
// DynaMaskRoIHead.simple_test_mask (mmdet/models/roi_heads/dynamask_roi_head.py:137-149), fused.
//
// Reference, per chunk of <= 100 detections and for stages s = 28 -> 56 -> 112:
//     m   = sigmoid(pred[s]) >= 0.5
//     nb  = generate_block_target(m, boundary_width=1) != 1          (losses/cross_entropy_loss.py:123-154)
//     nbu = interpolate(nb.float(), size[s+1], bilinear, align_corners=True) >= 0.5
//     pre = interpolate(pred[s], size[s+1], bilinear, align_corners=True)
//     pred[s+1][nbu] = pre[nbu]                                       (in place; cascades)
// i.e. ~12 elementwise / conv / interpolate launches per stage pair over [N,1,S,S] tensors.
//
// With boundary_width = 1 the block target reduces to a 3x3 test (derived from the two Laplacian
// convolutions, whose zero padding becomes ONE in the inverted image):
//     boundary(p) =  m(p) and some cell of the zero-padded 3x3 window is 0
//                 or !m(p) and some in-bounds cell of the 3x3 window is 1
// Here: one CTA per instance keeps every non-final stage in shared memory, applies the stages in
// order and streams the final stage through registers: each prediction is read once and each
// refined prediction written once.  Compiled without FMA contraction so that the >= 0.5 decisions
// on the interpolated mask follow the reference's separate multiply / add roundings.
#include "dm_common.cuh"

namespace dm {

constexpr int kRefineMaxStages = 4;
constexpr int kRefineThreads = 256;

struct RefineParams {
    const float* in[kRefineMaxStages];
    float* out[kRefineMaxStages];
    int h[kRefineMaxStages], w[kRefineMaxStages];
    int n_stages, N;
};

// ATen's bilinear source index for align_corners=True: src = dst * (in - 1) / (out - 1)
struct Lerp {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ Lerp lerp_of(int dst, float scale, int in_size) {
    Lerp r;
    const float s = __fmul_rn(scale, (float)dst);
    r.i0 = (int)s;
    r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
    r.l1 = __fsub_rn(s, (float)r.i0);
    r.l0 = __fsub_rn(1.0f, r.l1);
    return r;
}
__device__ __forceinline__ float bilerp(const Lerp& y, const Lerp& x, float v00, float v01, float v10, float v11) {
    return __fadd_rn(__fmul_rn(y.l0, __fadd_rn(__fmul_rn(x.l0, v00), __fmul_rn(x.l1, v01))),
                     __fmul_rn(y.l1, __fadd_rn(__fmul_rn(x.l0, v10), __fmul_rn(x.l1, v11))));
}
__device__ __forceinline__ float area_scale(int in_size, int out_size) {
    return out_size > 1 ? __fdiv_rn((float)(in_size - 1), (float)(out_size - 1)) : 0.0f;
}

// nb[p] = 1 when pixel p of the h x w prediction `pred` (shared memory) is NOT a boundary pixel
__device__ void non_boundary(const float* pred, int h, int w, uint8_t* m, uint8_t* nb) {
    const int n = h * w;
    for (int p = threadIdx.x; p < n; p += kRefineThreads)
        m[p] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-pred[p]))) >= 0.5f ? 1 : 0;
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += kRefineThreads) {
        const int y = p / w, x = p - y * w;
        int cells = 0, ones = 0;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int yy = y + dy, xx = x + dx;
                if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
                    ++cells;
                    ones += m[yy * w + xx];
                }
            }
        const bool boundary = m[p] ? (ones < 9) : (ones > 0);
        (void)cells;
        nb[p] = boundary ? 0 : 1;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kRefineThreads) refine_kernel(const __grid_constant__ RefineParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = blockIdx.x;
    // layout: two float planes (previous / current non-final stage), then the m and nb byte planes
    int max_px = 0;
    for (int s = 0; s + 1 < p.n_stages; ++s) max_px = max(max_px, p.h[s] * p.w[s]);
    const int plane = (max_px + 3) & ~3;
    float* bufA = reinterpret_cast<float*>(smem_raw);
    float* bufB = bufA + plane;
    uint8_t* m = reinterpret_cast<uint8_t*>(bufB + plane);
    uint8_t* nb = m + plane;

    float* prev = bufA;
    float* cur = bufB;
    {
        const int px = p.h[0] * p.w[0];
        const float* src = p.in[0] + (size_t)n * px;
        for (int i = threadIdx.x; i < px; i += kRefineThreads) prev[i] = src[i];
        if (p.out[0] && p.out[0] != p.in[0])
            for (int i = threadIdx.x; i < px; i += kRefineThreads) p.out[0][(size_t)n * px + i] = src[i];
    }
    __syncthreads();
    for (int s = 0; s + 1 < p.n_stages; ++s) {
        const int h0 = p.h[s], w0 = p.w[s], h1 = p.h[s + 1], w1 = p.w[s + 1];
        non_boundary(prev, h0, w0, m, nb);
        const float sy = area_scale(h0, h1), sx = area_scale(w0, w1);
        const bool last = s + 2 == p.n_stages;
        const int px = h1 * w1;
        const float* src = p.in[s + 1] + (size_t)n * px;
        float* dst = p.out[s + 1] ? p.out[s + 1] + (size_t)n * px : nullptr;
        for (int i = threadIdx.x; i < px; i += kRefineThreads) {
            const int y = i / w1, x = i - y * w1;
            const Lerp ly = lerp_of(y, sy, h0), lx = lerp_of(x, sx, w0);
            const int a = ly.i0 * w0 + lx.i0, b = ly.i0 * w0 + lx.i1, c = ly.i1 * w0 + lx.i0, d = ly.i1 * w0 + lx.i1;
            const float nbu = bilerp(ly, lx, (float)nb[a], (float)nb[b], (float)nb[c], (float)nb[d]);
            float v = src[i];
            if (nbu >= 0.5f) v = bilerp(ly, lx, prev[a], prev[b], prev[c], prev[d]);
            if (!last) cur[i] = v;
            if (dst) dst[i] = v;
        }
        __syncthreads();
        float* t = prev;
        prev = cur;
        cur = t;
    }
}

}  // namespace dm

extern "C" int dm_refine_stages(const float* const* stage_ptrs, const int32_t* sizes_hw, int n_stages,
                                int N, float* const* out_ptrs, dm_stream_t stream) {
    if (!stage_ptrs || !sizes_hw || !out_ptrs || n_stages < 2 || n_stages > dm::kRefineMaxStages || N < 0)
        return DM_EINVAL;
    dm::RefineParams p;
    int max_px = 0;
    for (int s = 0; s < n_stages; ++s) {
        p.in[s] = stage_ptrs[s];
        p.out[s] = out_ptrs[s];
        p.h[s] = sizes_hw[2 * s];
        p.w[s] = sizes_hw[2 * s + 1];
        if (!p.in[s] || p.h[s] < 1 || p.w[s] < 1) return DM_EINVAL;
        if (s + 1 < n_stages && p.h[s] * p.w[s] > max_px) max_px = p.h[s] * p.w[s];
    }
    if (!p.out[n_stages - 1]) return DM_EINVAL;
    p.n_stages = n_stages;
    p.N = N;
    if (N == 0) return DM_OK;
    const int plane = (max_px + 3) & ~3;
    const size_t smem = (size_t)plane * (2 * sizeof(float) + 2);
    if (smem > 200 * 1024) return DM_EUNSUPPORTED;
    DM_CUDA_CHECK(cudaFuncSetAttribute(dm::refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                  "dm_refine_stages");
    dm::refine_kernel<<<N, dm::kRefineThreads, smem, (cudaStream_t)stream>>>(p);
    DM_LAUNCH_CHECK("dm_refine_stages");
    return DM_OK;
}
