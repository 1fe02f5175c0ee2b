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
// bounds the pixels that have to be evaluated at all.
//
// Work layout.  The [N, rh, rw] output is treated as one flat array cut into 16 KB tiles, one CTA
// per tile (a write-only stream on this part runs fastest as many small CTAs, profiles/
// r01_membench.md).  >90 % of the tiles miss their instance's window: they are four 16-byte
// streaming zero stores per thread and nothing else.  A tile that meets the window
//   1. zeroes a 16 KB shared image of itself,
//   2. stages the source x coordinate of every window column (the x terms depend on the column
//      only) and sigmoid(mask) of the mask rows its canvas rows can reach, with a zero border,
//   3. evaluates the window pixels one per lane (a warp per canvas row) into the shared image,
//   4. streams the image out with the same four 16-byte stores per thread.
// Tiles that straddle two instances (or the end of the output) take a per-pixel path.
#include "dm_common.cuh"

namespace dm {

constexpr int kPasteThreads = 256;
constexpr int kTileBytes = 16384;   // output bytes per CTA: 4 x 16 B per thread
constexpr int kIxTab = 2048;        // window columns whose x coordinate is staged in shared memory
constexpr int kMaskStage = 3072;    // floats of sigmoid(mask) window staged per tile (12 KB)

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
    float inv_T;             // 1 / (rh * rw), first guess of the instance of a flat element index
    float thr;
    void* out;
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

__device__ __forceinline__ float sigmoidf_exact(float v) {
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));
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

template <int MODE>
__global__ void __launch_bounds__(kPasteThreads, 5)
paste_kernel(const __grid_constant__ PasteParams p) {
    constexpr int ES = (MODE == DM_PASTE_F32) ? 4 : 1;  // bytes per output element
    constexpr int V = 16 / ES;                          // elements per 16-byte chunk
    constexpr int TE = kTileBytes / ES;                 // elements per tile
    __shared__ __align__(16) unsigned char s_tile[kTileBytes];
    __shared__ __align__(16) float s_mask[kMaskStage];
    __shared__ float s_ix[kIxTab];

    const long long E0 = (long long)blockIdx.x * TE;  // first flat element of the tile
    const int T = p.rh * p.rw;                        // elements per instance (< 2^30, host-checked)
    long long n = (long long)((float)E0 * p.inv_T);
    n = n < 0 ? 0 : (n >= p.N ? p.N - 1 : n);
    while (n * T > E0) --n;
    while ((n + 1) * T <= E0) ++n;
    uint4* const out16 = reinterpret_cast<uint4*>(p.out) + (long long)blockIdx.x * (kTileBytes / 16);

    if (E0 + TE > (n + 1) * T) {
        // ---- per-pixel path: the tile straddles instances or is the last, partial one ------------
        Instance in;
        long long cur = -1;
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            const int ck = threadIdx.x + k * kPasteThreads;
            const long long Ec = E0 + (long long)ck * V;
            if (Ec >= p.total) break;
            long long nc = Ec / T;
            int e = (int)(Ec - nc * T);
            int row = e / p.rw, col = e - row * p.rw;
            uint32_t w[4] = {0u, 0u, 0u, 0u};
            const int cnt = (int)min((long long)V, p.total - Ec);
            for (int j = 0; j < cnt; ++j) {
                if (nc != cur) { in.load(p, nc); cur = nc; }
                const float v = in.eval(p.x_lo + col, p.y_lo + row);
                if (MODE == DM_PASTE_F32) w[j & 3] = __float_as_uint(v);
                else w[j >> 2] |= encode<MODE>(v, p.thr) << (8 * (j & 3));
                if (++col == p.rw) { col = 0; if (++row == p.rh) { row = 0; ++nc; } }
            }
            if (cnt == V) {
                __stcs(out16 + ck, make_uint4(w[0], w[1], w[2], w[3]));
            } else {
                for (int j = 0; j < cnt; ++j) {
                    if (MODE == DM_PASTE_F32) reinterpret_cast<float*>(p.out)[Ec + j] = __uint_as_float(w[j & 3]);
                    else reinterpret_cast<uint8_t*>(p.out)[Ec + j] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
                }
            }
        }
        return;
    }

    // ---- the tile lies inside instance n: rows [row_b, row_e] of its region --------------------
    const int e_lo = (int)(E0 - n * T);
    const int row_b = e_lo / p.rw;
    const int row_e = (e_lo + TE - 1) / p.rw;
    const float4 bx = *reinterpret_cast<const float4*>(p.boxes + 4 * (size_t)n);
    int xa, xb, ya, yb;
    window_1d(bx.y, bx.w, p.sh, p.img_h, ya, yb);
    window_1d(bx.x, bx.z, p.sw, p.img_w, xa, xb);
    // window in region coordinates, clipped to the region and to the tile's rows
    const int wxa = max(xa - p.x_lo, 0), wxb = min(xb - p.x_lo, p.rw);
    const int ra = max(max(ya - p.y_lo, 0), row_b), rb = min(min(yb - p.y_lo, p.rh) - 1, row_e);  // inclusive
    if (ra > rb || wxa >= wxb) {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) __stcs(out16 + threadIdx.x + k * kPasteThreads, z);
        return;
    }

    // ---- live tile ---------------------------------------------------------------------------------
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) reinterpret_cast<uint4*>(s_tile)[threadIdx.x + k * kPasteThreads] = z;
    }
    const int ww = wxb - wxa;
    const bool use_tab = ww <= kIxTab;
    if (use_tab)
        for (int c = threadIdx.x; c < ww; c += kPasteThreads) s_ix[c] = src_coord(p.x_lo + wxa + c, bx.x, bx.z, p.sw);
    const long long cls = p.labels ? p.labels[n] : 0;
    const float* __restrict__ m = p.masks + n * p.stride_n + cls * p.stride_c;
    const int st_w = p.sw + 2;
    int st_lo = 0;
    bool staged = false;
    {
        // mask rows the live canvas rows can reach (the source coordinate is monotone in py)
        const float ia = src_coord(p.y_lo + ra, bx.y, bx.w, p.sh);
        const float ib = src_coord(p.y_lo + rb, bx.y, bx.w, p.sh);
        if (ia == ia && ib == ib) {
            const int lo = (int)fmaxf(floorf(fminf(ia, ib)), -1.0f);
            const int hi = (int)fminf(floorf(fmaxf(ia, ib)) + 1.0f, (float)p.sh);
            const int nrow = hi - lo + 1;
            if (nrow >= 1 && nrow * st_w <= kMaskStage) {
                for (int q = threadIdx.x; q < nrow * st_w; q += kPasteThreads) {
                    const int yy = q / st_w;
                    const int y = lo + yy, x = q - yy * st_w - 1;
                    float v = 0.0f;
                    if (y >= 0 && y < p.sh && x >= 0 && x < p.sw) {
                        v = __ldg(m + y * p.sw + x);
                        if (p.apply_sigmoid) v = sigmoidf_exact(v);
                    }
                    s_mask[q] = v;
                }
                staged = true;
                st_lo = lo;
            }
        }
    }
    __syncthreads();

    // ---- evaluate the window pixels of the tile: a warp per canvas row, a lane per pixel ----------
    Instance in;
    if (!staged) in.load(p, n);
    for (int r = ra + warp; r <= rb; r += kPasteThreads / 32) {
        const AxisTerm ry = axis_term(src_coord(p.y_lo + r, bx.y, bx.w, p.sh), p.sh);
        if (ry.state == 0) continue;
        const int e_row = r * p.rw - e_lo;  // tile offset of the row's column 0
        const int ca = max(wxa, -e_row), cb = min(wxb, TE - e_row);
        const float* mrow = s_mask + (ry.lo - st_lo) * st_w + 1;
        for (int c = ca + lane; c < cb; c += 32) {
            const float ix = use_tab ? s_ix[c - wxa] : src_coord(p.x_lo + c, bx.x, bx.z, p.sw);
            const AxisTerm cx = axis_term(ix, p.sw);
            if (cx.state == 0) continue;
            float v;
            if (cx.state == 2 || ry.state == 2) {
                v = __int_as_float(0x7fc00000);
            } else if (staged) {
                const float* q = mrow + cx.lo;
                v = bilerp(q[0], q[1], q[st_w], q[st_w + 1], cx, ry);
            } else {
                v = bilerp(in.tap(ry.lo, cx.lo), in.tap(ry.lo, cx.lo + 1), in.tap(ry.lo + 1, cx.lo),
                           in.tap(ry.lo + 1, cx.lo + 1), cx, ry);
            }
            if (MODE == DM_PASTE_F32) reinterpret_cast<float*>(s_tile)[e_row + c] = v;
            else s_tile[e_row + c] = (unsigned char)encode<MODE>(v, p.thr);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k)
        __stcs(out16 + threadIdx.x + k * kPasteThreads, reinterpret_cast<const uint4*>(s_tile)[threadIdx.x + k * kPasteThreads]);
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
    const long long per_inst = (long long)p.rh * p.rw;
    if (per_inst == 0 || N == 0) return DM_OK;
    if (per_inst >= (1ll << 30) || (long long)S_h * S_w >= (1ll << 30)) return DM_EUNSUPPORTED;
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
    p.total = per_inst * N;
    p.inv_T = 1.0f / (float)per_inst;
    const int ES = out_mode == DM_PASTE_F32 ? 4 : 1;
    const long long tiles = (p.total * ES + dm::kTileBytes - 1) / dm::kTileBytes;
    if (tiles >= (1ll << 31)) return DM_EUNSUPPORTED;
    dim3 grid((unsigned)tiles);
    cudaStream_t st = (cudaStream_t)stream;
    switch (out_mode) {
        case DM_PASTE_BOOL:
            dm::paste_kernel<DM_PASTE_BOOL><<<grid, dm::kPasteThreads, 0, st>>>(p);
            break;
        case DM_PASTE_U8:
            dm::paste_kernel<DM_PASTE_U8><<<grid, dm::kPasteThreads, 0, st>>>(p);
            break;
        default:
            dm::paste_kernel<DM_PASTE_F32><<<grid, dm::kPasteThreads, 0, st>>>(p);
            break;
    }
    DM_LAUNCH_CHECK("dm_paste_masks");
    return DM_OK;
}
