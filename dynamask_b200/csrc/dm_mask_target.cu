// Stage 4: training mask targets from ground-truth bitmaps, all sizes and all images in one launch.
//
// Replaces BitmapMasks.crop_and_resize (mmdet/core/mask/structures.py:256-286), the clip of
// mask_target_single (mmdet/core/mask/mask_target.py:49-51) and the per-image x per-size loop of
// DynaMaskHead.get_targets (mmdet/models/roi_heads/mask_heads/dynamask_head.py:246-271).
//
// The reference uploads every image's masks once per size, blows the K selected masks up to a
// [K,H,W] fp32 tensor and runs RoIAlign on it.  Here the uint8 bitmaps are read in place (the
// uint8->fp32 conversion is exact), one thread per output pixel.
//
// Bit-exactness: the result is (avg >= 0.5) of fp32 bilinear samples of a binary image and exact
// ties do occur, so this TU is compiled with -fmad=false and each output accumulates its samples
// serially, iy outer / ix inner, in the operation order of SURVEY.md Appendix A.1.  No shuffle
// or tree reduction is allowed here.
#include "dm_common.cuh"

namespace dm {

struct MaskTargetParams {
    const uint8_t* blob;
    const int64_t* img_offsets;
    const int32_t* img_ghw;
    int B;
    const float* boxes;
    const int64_t* inds;
    const int32_t* roi_img;
    int K;
    int clip;
    int n_sizes;
    int sh[DM_MAX_BUCKETS], sw[DM_MAX_BUCKETS];
    int off[DM_MAX_BUCKETS + 1];  // prefix of sh*sw: pixels per RoI before size s
    float* out[DM_MAX_BUCKETS];
};

constexpr int kMtThreads = 256;

__global__ void __launch_bounds__(kMtThreads)
mask_target_kernel(const __grid_constant__ MaskTargetParams p) {
    const int per_roi = p.off[p.n_sizes];
    const long long total = (long long)p.K * per_roi;
    for (long long idx = (long long)blockIdx.x * kMtThreads + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * kMtThreads) {
        const int k = (int)(idx / per_roi);
        const int r = (int)(idx - (long long)k * per_roi);
        int s = 0;
        while (s + 1 < p.n_sizes && r >= p.off[s + 1]) ++s;
        const int e = r - p.off[s];
        const int sh = p.sh[s], sw = p.sw[s];
        const int ph = e / sw, pw = e - ph * sw;
        float* out = p.out[s] + ((size_t)k * sh + ph) * sw + pw;

        const int img = p.roi_img ? p.roi_img[k] : 0;
        if (img < 0 || img >= p.B) { *out = 0.0f; continue; }
        const int G = p.img_ghw[img * 3 + 0], H = p.img_ghw[img * 3 + 1], W = p.img_ghw[img * 3 + 2];
        const long long gi = p.inds[k];
        if (gi < 0 || gi >= G) { *out = 0.0f; continue; }
        const uint8_t* __restrict__ m = p.blob + p.img_offsets[img] + (size_t)gi * H * W;

        float r5[5];
        r5[0] = 0.0f;
        r5[1] = p.boxes[4 * (size_t)k + 0];
        r5[2] = p.boxes[4 * (size_t)k + 1];
        r5[3] = p.boxes[4 * (size_t)k + 2];
        r5[4] = p.boxes[4 * (size_t)k + 3];
        if (p.clip) {
            r5[1] = fminf(fmaxf(r5[1], 0.0f), (float)W);
            r5[3] = fminf(fmaxf(r5[3], 0.0f), (float)W);
            r5[2] = fminf(fmaxf(r5[2], 0.0f), (float)H);
            r5[4] = fminf(fmaxf(r5[4], 0.0f), (float)H);
        }
        const RoiGeom g = roi_geom(r5, 1.0f, sh, sw, 0, 1);
        const int cnt = g.gh * g.gw;
        const float count = (float)(cnt > 1 ? cnt : 1);
        float acc = 0.0f;
        for (int iy = 0; iy < g.gh; ++iy) {
            const float y = sample_coord(g.rsh, g.bh, g.gh, ph, iy);
            int yl, yh;
            float ly, hy;
            if (!axis_tap(y, H, yl, yh, ly, hy)) continue;
            const uint8_t* __restrict__ row_l = m + (size_t)yl * W;
            const uint8_t* __restrict__ row_h = m + (size_t)yh * W;
            for (int ix = 0; ix < g.gw; ++ix) {
                const float x = sample_coord(g.rsw, g.bw, g.gw, pw, ix);
                int xl, xh;
                float lx, hx;
                if (!axis_tap(x, W, xl, xh, lx, hx)) continue;
                const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx);
                const float w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
                const float v1 = (float)__ldg(row_l + xl), v2 = (float)__ldg(row_l + xh);
                const float v3 = (float)__ldg(row_h + xl), v4 = (float)__ldg(row_h + xh);
                const float v = __fadd_rn(
                    __fadd_rn(__fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2)), __fmul_rn(w3, v3)),
                    __fmul_rn(w4, v4));
                acc = __fadd_rn(acc, v);
            }
        }
        *out = __fdiv_rn(acc, count) >= 0.5f ? 1.0f : 0.0f;
    }
}

}  // namespace dm

extern "C" int dm_mask_target(const uint8_t* gt_blob, const int64_t* img_offsets,
                              const int32_t* img_ghw, int B, const float* boxes,
                              const int64_t* inds, const int32_t* roi_img, int K, int clip,
                              const int32_t* sizes_hw, int n_sizes, float* const* out_ptrs,
                              dm_stream_t stream) {
    if (K < 0 || B < 0 || n_sizes < 1 || n_sizes > DM_MAX_BUCKETS) return DM_EINVAL;
    if (!sizes_hw || !out_ptrs) return DM_EINVAL;
    if (K == 0) return DM_OK;
    if (!gt_blob || !img_offsets || !img_ghw || !boxes || !inds || B < 1) return DM_EINVAL;
    dm::MaskTargetParams p;
    p.blob = gt_blob;
    p.img_offsets = img_offsets;
    p.img_ghw = img_ghw;
    p.B = B;
    p.boxes = boxes;
    p.inds = inds;
    p.roi_img = roi_img;
    p.K = K;
    p.clip = clip;
    p.n_sizes = n_sizes;
    p.off[0] = 0;
    for (int s = 0; s < n_sizes; ++s) {
        p.sh[s] = sizes_hw[2 * s];
        p.sw[s] = sizes_hw[2 * s + 1];
        if (p.sh[s] < 1 || p.sw[s] < 1 || !out_ptrs[s]) return DM_EINVAL;
        p.off[s + 1] = p.off[s] + p.sh[s] * p.sw[s];
        p.out[s] = out_ptrs[s];
    }
    const long long total = (long long)K * p.off[n_sizes];
    long long blocks = (total + dm::kMtThreads - 1) / dm::kMtThreads;
    const long long cap = (long long)dm::sm_count() * 32;
    if (blocks > cap) blocks = cap;
    dm::mask_target_kernel<<<(unsigned)blocks, dm::kMtThreads, 0, (cudaStream_t)stream>>>(p);
    DM_LAUNCH_CHECK("dm_mask_target");
    return DM_OK;
}
