// Stage 4: training mask targets from ground-truth bitmaps, all sizes and all images in one launch.
//
// Replaces BitmapMasks.crop_and_resize (mmdet/core/mask/structures.py:256-286), the clip of
// mask_target_single (mmdet/core/mask/mask_target.py:49-51) and the per-image x per-size loop of
// DynaMaskHead.get_targets (mmdet/models/roi_heads/mask_heads/dynamask_head.py:246-271).
//
// The reference uploads every image's masks once per size, blows the K selected masks up to a
// [K,H,W] fp32 tensor and runs RoIAlign on it.  Here the uint8 bitmaps are read in place (the
// uint8->fp32 conversion is exact), one thread per output pixel, the per-axis tap arithmetic of a
// (RoI, size) shared through shared memory.
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
    int band_off[DM_MAX_BUCKETS + 1];  // prefix of the 256-output bands per size
    int bands_per_roi;
    float* out[DM_MAX_BUCKETS];
};

constexpr int kMtThreads = 256;
constexpr int kMtXTab = 1536;   // x tap entries staged per CTA: S_w bins x g_w samples (< box width + S_w)
constexpr int kMtYTab = 1024;   // y tap entries: the band's bin rows x g_h samples

// Geometry of RoI k at size s: false when the output is identically zero (bad image / mask index).
struct MtRoi {
    RoiGeom g;
    const uint8_t* m;
    int H, W;
};

__device__ __forceinline__ bool mt_roi(const MaskTargetParams& p, int k, int s, MtRoi& r) {
    const int img = p.roi_img ? p.roi_img[k] : 0;
    if (img < 0 || img >= p.B) return false;
    const int G = p.img_ghw[img * 3 + 0];
    r.H = p.img_ghw[img * 3 + 1];
    r.W = p.img_ghw[img * 3 + 2];
    const long long gi = p.inds[k];
    if (gi < 0 || gi >= G) return false;
    r.m = p.blob + p.img_offsets[img] + (size_t)gi * r.H * r.W;
    float r5[5];
    r5[0] = 0.0f;
    r5[1] = p.boxes[4 * (size_t)k + 0];
    r5[2] = p.boxes[4 * (size_t)k + 1];
    r5[3] = p.boxes[4 * (size_t)k + 2];
    r5[4] = p.boxes[4 * (size_t)k + 3];
    if (p.clip) {
        r5[1] = fminf(fmaxf(r5[1], 0.0f), (float)r.W);
        r5[3] = fminf(fmaxf(r5[3], 0.0f), (float)r.W);
        r5[2] = fminf(fmaxf(r5[2], 0.0f), (float)r.H);
        r5[4] = fminf(fmaxf(r5[4], 0.0f), (float)r.H);
    }
    r.g = roi_geom(r5, 1.0f, p.sh[s], p.sw[s], 0, 1);
    return true;
}

// One axis tap of one sample, as staged in shared memory: {lo, hi, l, h}, exactly what sample_coord /
// axis_tap give per sample.  A sample the reference skips is staged as taps on pixel 0 with both weights
// zero: it then adds +0.0f to the running sum, which leaves that (never negative) sum unchanged bit for
// bit -- and the sample loops need no branch, so the loads of several samples can be in flight at once
// (the only serial dependency left is the one add per sample).
__device__ __forceinline__ float4 mt_tap(float start, float bin, int grid, int p, int i, int size) {
    int lo, hi;
    float l, h;
    if (!axis_tap(sample_coord(start, bin, grid, p, i), size, lo, hi, l, h))
        return make_float4(__int_as_float(0), __int_as_float(0), 0.0f, 0.0f);
    return make_float4(__int_as_float(lo), __int_as_float(hi), l, h);
}

// One CTA per (RoI, size, band of 256 outputs).  The tap of a sample depends on one axis only: the
// CTA evaluates the x taps of all S_w x g_w sample columns and the y taps of its bin rows once (the
// two exact divisions and the clamping per tap), stages them in shared memory, and every output then
// walks its g_h x g_w samples with one 16-byte table read, four byte loads, eight multiplies and four
// adds each -- in the reference's order (iy outer, ix inner, serial sum), bit for bit.
// Boxes too large for the tables (wider than ~1400 px at the coarse sizes) take the per-sample form.
__global__ void __launch_bounds__(kMtThreads)
mask_target_kernel(const __grid_constant__ MaskTargetParams p) {
    __shared__ float4 s_xt[kMtXTab];   // [ix][pw]: lanes of a warp (consecutive pw) read consecutive entries
    __shared__ float4 s_yt[kMtYTab];   // [ph - ph_lo][iy]
    const int k = blockIdx.x / p.bands_per_roi;
    const int rb = blockIdx.x - k * p.bands_per_roi;
    int s = 0;
    while (s + 1 < p.n_sizes && rb >= p.band_off[s + 1]) ++s;
    const int band = rb - p.band_off[s];
    const int sh = p.sh[s], sw = p.sw[s];
    const int e = band * kMtThreads + threadIdx.x;
    const bool live = e < sh * sw;
    const int ph = live ? e / sw : 0, pw = live ? e - (e / sw) * sw : 0;
    float* out = p.out[s] + ((size_t)k * sh + ph) * sw + pw;
    MtRoi roi;
    if (!mt_roi(p, k, s, roi) || roi.g.gh <= 0 || roi.g.gw <= 0) {   // CTA-uniform
        if (live) *out = 0.0f;
        return;
    }
    const RoiGeom& g = roi.g;
    const int gh = g.gh, gw = g.gw;
    const float count = (float)(gh * gw);
    const int ph_lo = (band * kMtThreads) / sw;
    const int ph_hi = min(sh - 1, (band * kMtThreads + kMtThreads - 1) / sw);
    const int nx = sw * gw, ny = (ph_hi - ph_lo + 1) * gh;
    float acc = 0.0f;
    if (nx <= kMtXTab && ny <= kMtYTab) {   // CTA-uniform
        for (int i = threadIdx.x; i < nx; i += kMtThreads) {
            const int ix = i / sw, c = i - ix * sw;
            s_xt[i] = mt_tap(g.rsw, g.bw, gw, c, ix, roi.W);
        }
        for (int i = threadIdx.x; i < ny; i += kMtThreads) {
            const int r = i / gh, iy = i - r * gh;
            s_yt[i] = mt_tap(g.rsh, g.bh, gh, ph_lo + r, iy, roi.H);
        }
        __syncthreads();
        if (!live) return;
        const float4* yt = s_yt + (ph - ph_lo) * gh;
        const float4* xt = s_xt + pw;
        const int W = roi.W;
        for (int iy = 0; iy < gh; ++iy) {
            const float4 ty = yt[iy];
            const float ly = ty.z, hy = ty.w;
            const uint8_t* __restrict__ row_l = roi.m + (size_t)__float_as_int(ty.x) * W;
            const uint8_t* __restrict__ row_h = roi.m + (size_t)__float_as_int(ty.y) * W;
#pragma unroll 4
            for (int ix = 0; ix < gw; ++ix) {
                const float4 tx = xt[ix * sw];
                const int xl = __float_as_int(tx.x), xh = __float_as_int(tx.y);
                const float lx = tx.z, hx = tx.w;
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
    } else {
        if (!live) return;
        for (int iy = 0; iy < gh; ++iy) {
            const float y = sample_coord(g.rsh, g.bh, gh, ph, iy);
            int yl, yh;
            float ly, hy;
            if (!axis_tap(y, roi.H, yl, yh, ly, hy)) continue;
            const uint8_t* __restrict__ row_l = roi.m + (size_t)yl * roi.W;
            const uint8_t* __restrict__ row_h = roi.m + (size_t)yh * roi.W;
            for (int ix = 0; ix < gw; ++ix) {
                const float x = sample_coord(g.rsw, g.bw, gw, pw, ix);
                int xl, xh;
                float lx, hx;
                if (!axis_tap(x, roi.W, xl, xh, lx, hx)) continue;
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
    }
    *out = __fdiv_rn(acc, count) >= 0.5f ? 1.0f : 0.0f;
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
    p.band_off[0] = 0;
    for (int s = 0; s < n_sizes; ++s)
        p.band_off[s + 1] = p.band_off[s] + (p.sh[s] * p.sw[s] + dm::kMtThreads - 1) / dm::kMtThreads;
    p.bands_per_roi = p.band_off[n_sizes];
    const long long blocks = (long long)K * p.bands_per_roi;
    if (blocks >= (1ll << 31)) return DM_EUNSUPPORTED;
    dm::mask_target_kernel<<<(unsigned)blocks, dm::kMtThreads, 0, (cudaStream_t)stream>>>(p);
    DM_LAUNCH_CHECK("dm_mask_target");
    return DM_OK;
}
