// Stage 3: paste instance masks into image canvases -- sigmoid + bilinear resample + threshold in
// one pass, one 16-byte store per 16 output pixels.
//
// Replaces _do_paste_mask (mmdet/models/roi_heads/mask_heads/fcn_mask_head.py:240-308) and the
// sigmoid / class select / threshold / index_put around it in DynaMaskHead.get_seg_masks
// (mmdet/models/roi_heads/mask_heads/dynamask_head.py:279-342).  The reference materialises an
// [N,H,W,2] fp32 sampling grid (8 B per output pixel), an fp32 result, a thresholded copy and an
// index_put; here the only HBM traffic is the N*H*W output bytes plus the N tiny masks.
//
// Arithmetic follows the reference expression order:
//   gx = ((px + .5) - x0) / (x1 - x0) * 2 - 1 ; +-inf -> 0        (fcn_mask_head.py:284-296)
//   ix = ((gx + 1) * S_w - 1) / 2                                  (grid_sample, align_corners=False)
//   bilinear, zero padding, four taps weighted (east-ix)*(south-iy) ...
// Pixels whose sample falls outside (-1, S) are exactly zero; a conservative per-instance window
// lets whole 16-byte chunks be zero-filled without evaluating the expression.
#include "dm_common.cuh"

namespace dm {

struct PasteParams {
    const float* masks;
    long long stride_n, stride_c;
    const int64_t* labels;
    int N, sh, sw;
    int apply_sigmoid;
    const float* boxes;
    int img_h, img_w;
    int x_lo, y_lo, rw, rh;  // region origin and size
    float thr;
    void* out;
    long long total;  // N * rh * rw
};

struct PasteInst {
    const float* m;
    float x0, y0, x1, y1;
    int xa, xb, ya, yb;  // conservative non-zero window in canvas pixels, [a, b)
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

__device__ __forceinline__ PasteInst load_inst(const PasteParams& p, int n) {
    PasteInst it;
    const long long c = p.labels ? p.labels[n] : 0;
    it.m = p.masks + (long long)n * p.stride_n + c * p.stride_c;
    const float4 b = *reinterpret_cast<const float4*>(p.boxes + 4 * (size_t)n);
    it.x0 = b.x; it.y0 = b.y; it.x1 = b.z; it.y1 = b.w;
    window_1d(it.x0, it.x1, p.sw, p.img_w, it.xa, it.xb);
    window_1d(it.y0, it.y1, p.sh, p.img_h, it.ya, it.yb);
    return it;
}

// normalised -> mask-pixel coordinate of canvas pixel centre `pc` along one axis
__device__ __forceinline__ float src_coord(int pc, float c0, float c1, int S) {
    float g = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(__fadd_rn((float)pc, 0.5f), c0),
                                             __fsub_rn(c1, c0)), 2.0f), 1.0f);
    if (isinf(g)) g = 0.0f;
    return __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.0f), (float)S), 1.0f), 2.0f);
}

__device__ __forceinline__ float mask_val(const PasteInst& it, const PasteParams& p, int y, int x) {
    if (y < 0 || y >= p.sh || x < 0 || x >= p.sw) return 0.0f;
    float v = __ldg(it.m + (size_t)y * p.sw + x);
    if (p.apply_sigmoid) v = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));
    return v;
}

struct RowTerm {
    int yn;
    float wn, ws;  // weights of the north / south mask rows
    bool live;
};

__device__ __forceinline__ RowTerm row_term(const PasteInst& it, const PasteParams& p, int py) {
    RowTerm r;
    r.live = (py >= it.ya && py < it.yb);
    r.yn = 0; r.wn = 0.0f; r.ws = 0.0f;
    if (!r.live) return r;
    const float iy = src_coord(py, it.y0, it.y1, p.sh);
    if (iy != iy) {  // NaN row (0/0): every pixel of the row is NaN in the reference
        r.yn = 0; r.wn = iy; r.ws = iy;
        return r;
    }
    if (!(iy > -1.0f && iy < (float)p.sh)) { r.live = false; return r; }
    const float fy = floorf(iy);
    r.yn = (int)fy;
    r.wn = __fsub_rn(__fadd_rn(fy, 1.0f), iy);
    r.ws = __fsub_rn(iy, fy);
    return r;
}

__device__ __forceinline__ float paste_value(const PasteInst& it, const PasteParams& p,
                                             const RowTerm& rt, int px) {
    if (!rt.live || px < it.xa || px >= it.xb) return 0.0f;
    const float ix = src_coord(px, it.x0, it.x1, p.sw);
    if (ix != ix || rt.wn != rt.wn) return __int_as_float(0x7fc00000);
    if (!(ix > -1.0f && ix < (float)p.sw)) return 0.0f;
    const float fx = floorf(ix);
    const int xw = (int)fx;
    const float we = __fsub_rn(ix, fx), ww = __fsub_rn(__fadd_rn(fx, 1.0f), ix);
    float acc = mask_val(it, p, rt.yn, xw) * (ww * rt.wn);
    acc += mask_val(it, p, rt.yn, xw + 1) * (we * rt.wn);
    acc += mask_val(it, p, rt.yn + 1, xw) * (ww * rt.ws);
    acc += mask_val(it, p, rt.yn + 1, xw + 1) * (we * rt.ws);
    return acc;
}

template <int MODE>
__device__ __forceinline__ uint32_t encode(float v, float thr) {
    if (MODE == DM_PASTE_BOOL) return v >= thr ? 1u : 0u;
    // reference: (val * 255).to(uint8); values are in [0,1] so the cast never saturates
    const float s = v * 255.0f;
    return (s != s) ? 0u : (uint32_t)(unsigned char)(int)s;
}

constexpr int kPasteThreads = 256;

template <int MODE>
__global__ void __launch_bounds__(kPasteThreads)
paste_kernel(const __grid_constant__ PasteParams p) {
    constexpr int V = (MODE == DM_PASTE_F32) ? 4 : 16;
    const long long plane = (long long)p.rh * p.rw;
    const long long nchunks = (p.total + V - 1) / V;
    for (long long chunk = (long long)blockIdx.x * kPasteThreads + threadIdx.x; chunk < nchunks;
         chunk += (long long)gridDim.x * kPasteThreads) {
        const long long e0 = chunk * V;
        int n = (int)(e0 / plane);
        const int rem = (int)(e0 - (long long)n * plane);
        int row = rem / p.rw;
        int col = rem - row * p.rw;
        PasteInst it = load_inst(p, n);
        const bool full = (e0 + V <= p.total);
        // fast path: chunk inside one row and entirely outside the instance's window
        if (full && col + V <= p.rw) {
            const int py = p.y_lo + row, px0 = p.x_lo + col;
            if (py < it.ya || py >= it.yb || px0 + V <= it.xa || px0 >= it.xb) {
                if (MODE == DM_PASTE_F32)
                    __stcs(reinterpret_cast<float4*>(p.out) + chunk, make_float4(0.f, 0.f, 0.f, 0.f));
                else
                    __stcs(reinterpret_cast<uint4*>(p.out) + chunk, make_uint4(0u, 0u, 0u, 0u));
                continue;
            }
        }
        RowTerm rt = row_term(it, p, p.y_lo + row);
        float fv[4];
        uint32_t packed[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float v = 0.0f;
            const bool in_range = full || (e0 + j < p.total);
            if (in_range) v = paste_value(it, p, rt, p.x_lo + col);
            if (MODE == DM_PASTE_F32) fv[j & 3] = v;
            else packed[j >> 2] |= encode<MODE>(v, p.thr) << (8 * (j & 3));
            if (!full && in_range) {
                if (MODE == DM_PASTE_F32) reinterpret_cast<float*>(p.out)[e0 + j] = v;
                else reinterpret_cast<uint8_t*>(p.out)[e0 + j] = (uint8_t)encode<MODE>(v, p.thr);
            }
            // advance to the next output element (may wrap to the next row / instance)
            if (++col == p.rw) {
                col = 0;
                if (++row == p.rh) {
                    row = 0;
                    ++n;
                    if (n < p.N) it = load_inst(p, n);
                }
                if (n < p.N) rt = row_term(it, p, p.y_lo + row);
            }
        }
        if (full) {
            if (MODE == DM_PASTE_F32)
                __stcs(reinterpret_cast<float4*>(p.out) + chunk, make_float4(fv[0], fv[1], fv[2], fv[3]));
            else
                __stcs(reinterpret_cast<uint4*>(p.out) + chunk,
                       make_uint4(packed[0], packed[1], packed[2], packed[3]));
        }
    }
}

}  // namespace dm

extern "C" int dm_paste_masks(const float* masks, int64_t mask_stride_n, int64_t mask_stride_c,
                              const int64_t* labels, int N, int S_h, int S_w, int apply_sigmoid,
                              const float* boxes, int img_h, int img_w, int x_lo, int y_lo,
                              int x_hi, int y_hi, float thr, int out_mode, void* out,
                              dm_stream_t stream) {
    if (N < 0 || S_h < 1 || S_w < 1 || img_h < 0 || img_w < 0) return DM_EINVAL;
    if (x_lo < 0 || y_lo < 0 || x_hi > img_w || y_hi > img_h || x_hi < x_lo || y_hi < y_lo)
        return DM_EINVAL;
    if (out_mode != DM_PASTE_BOOL && out_mode != DM_PASTE_U8 && out_mode != DM_PASTE_F32)
        return DM_EINVAL;
    dm::PasteParams p;
    p.rw = x_hi - x_lo;
    p.rh = y_hi - y_lo;
    p.total = (long long)N * p.rh * p.rw;
    if (p.total == 0) return DM_OK;
    if (!masks || !boxes || !out) return DM_EINVAL;
    if ((reinterpret_cast<uintptr_t>(out) & 15u) || (reinterpret_cast<uintptr_t>(boxes) & 15u))
        return DM_EINVAL;
    p.masks = masks;
    p.stride_n = mask_stride_n;
    p.stride_c = mask_stride_c;
    p.labels = labels;
    p.N = N;
    p.sh = S_h;
    p.sw = S_w;
    p.apply_sigmoid = apply_sigmoid;
    p.boxes = boxes;
    p.img_h = img_h;
    p.img_w = img_w;
    p.x_lo = x_lo;
    p.y_lo = y_lo;
    p.thr = thr;
    p.out = out;
    const int V = out_mode == DM_PASTE_F32 ? 4 : 16;
    const long long nchunks = (p.total + V - 1) / V;
    long long blocks = (nchunks + dm::kPasteThreads - 1) / dm::kPasteThreads;
    const long long cap = (long long)dm::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    switch (out_mode) {
        case DM_PASTE_BOOL:
            dm::paste_kernel<DM_PASTE_BOOL><<<(unsigned)blocks, dm::kPasteThreads, 0, st>>>(p);
            break;
        case DM_PASTE_U8:
            dm::paste_kernel<DM_PASTE_U8><<<(unsigned)blocks, dm::kPasteThreads, 0, st>>>(p);
            break;
        default:
            dm::paste_kernel<DM_PASTE_F32><<<(unsigned)blocks, dm::kPasteThreads, 0, st>>>(p);
            break;
    }
    DM_LAUNCH_CHECK("dm_paste_masks");
    return DM_OK;
}
