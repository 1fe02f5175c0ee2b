// Stage 3: paste instance masks into image canvases -- sigmoid + bilinear resample + threshold.
//
// Replaces _do_paste_mask (mmdet/models/roi_heads/mask_heads/fcn_mask_head.py:240-308) and the
// sigmoid / class select / threshold / index_put around it in DynaMaskHead.get_seg_masks
// (mmdet/models/roi_heads/mask_heads/dynamask_head.py:279-342).  The reference materialises an
// [N,H,W,2] fp32 sampling grid (8 B per output pixel), an fp32 result, a thresholded copy and an
// index_put; here the only HBM traffic is the N*H*W output bytes plus the N tiny masks.
//
// Arithmetic follows the reference:
//   gx = ((px + .5) - x0) / (x1 - x0) * 2 - 1 ; +-inf -> 0        (fcn_mask_head.py:284-296)
//   ix = ((gx + 1) * S_w - 1) / 2                                  (grid_sample, align_corners=False)
//   bilinear, zero padding.
// Pixels whose sample falls outside (-1, S) are exactly zero; a conservative per-instance window
// bounds the pixels that have to be evaluated at all (~2-3 % of a COCO-shaped canvas).
//
// Window kernel (paste_window_body): grid = (32-row bands of a window, instances).  A CTA stages
// sigmoid(mask) for the mask rows its canvas rows can reach and the x terms of the window's columns
// (both depend on one axis only), then a warp per canvas row interpolates along y into a private row
// buffer and along x straight into global memory, one pixel per lane.
// The zero background comes in one of two ways:
//   * fused (default for canvases of >= 64 KB per instance): ONE launch, paste_fused_kernel.  The window
//     CTAs own every element of their instances' window ROWS (zeros included), the rest of the output
//     is zeroed in 64 KB tiles dealt to the same CTAs, which issue their tiles first -- fire-and-forget
//     stores that drain to HBM while the CTA goes on to its window bands.  183 us for the C4 shape.
//   * two launches (small canvases, and DM_PASTE_FUSED=0): paste_fill_kernel zeroes the whole output, 16
//     bytes per thread, one shot (7.4 TB/s, profiles/r01_membench.md), then paste_window_kernel writes
//     the non-zero pixels only.  227 us for the C4 shape (115 us + 83 us + launch gap).
// An earlier single-launch form (every 16 KB tile decides "zero or evaluate") spent more time in
// the latency chains of its sparse live tiles than in the fill itself.
//
// The four-tap sum is evaluated separably, which regroups the reference's
// nw*(wl*wn) + ne*(wh*wn) + sw*(wl*ws) + se*(wh*ws) as wl*(wn*nw + ws*sw) + wh*(wn*ne + ws*se):
// a few 1e-8 apart, far inside the 1e-5 / 99.99 % bars (tests/test_gpu_parity.py).
#include <cstdlib>

#include "dm_paste_common.cuh"

namespace dm {

template <int MODE>
__device__ __forceinline__ void put(const PasteParams& p, long long e, float v) {
    if (MODE == DM_PASTE_F32) reinterpret_cast<float*>(p.out)[e] = v;
    else reinterpret_cast<uint8_t*>(p.out)[e] = (uint8_t)encode<MODE>(v, p.thr);
}

// 1. zero fill: one 16-byte streaming store per thread (the last, partial chunk element-wise)
__global__ void __launch_bounds__(256) paste_fill_kernel(uint4* __restrict__ out, long long n16, long long nbytes) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < n16) {
        __stcs(out + i, make_uint4(0u, 0u, 0u, 0u));
    } else if (i == n16) {
        unsigned char* tail = reinterpret_cast<unsigned char*>(out + n16);
        for (long long k = 0; k < nbytes - n16 * 16; ++k) tail[k] = 0;
    }
}

// 2. the instances' windows
//
// Per canvas row r the warp builds VD[i] = (V[x], V[x+1] - V[x]) for mask column x = i - 1, where
// V[x] = wn * sigmoid(mask)[yn][x] + ws * sigmoid(mask)[yn+1][x] (zero outside the mask); a pixel
// is then V[lo] + wh * (V[lo+1] - V[lo]) -- one 8-byte table read, one 8-byte VD read, one FMA.
// Columns whose sample misses (-1, S) point at a (0, 0) entry, NaN columns at a (NaN, NaN) entry,
// so the pixel loop has no branches.
// The rows an instance's window kernel CTAs own in FULL mode (they write every element of these rows,
// the fill role none of them): region rows [wya, wyb), empty when the instance is not pasted at all.
// Both roles of paste_fused_kernel call this, so they agree by construction.
__device__ __forceinline__ bool window_rows(const PasteParams& p, int n, int& wya, int& wyb, int& wxa, int& wxb,
                                            float4& bx) {
    if (p.select && p.select[n] != p.select_value) return false;
    bx = *reinterpret_cast<const float4*>(p.boxes + 4 * (size_t)n);
    int xa, xb, ya, yb;
    window_1d(bx.y, bx.w, p.sh, p.img_h, ya, yb);
    window_1d(bx.x, bx.z, p.sw, p.img_w, xa, xb);
    // window in region coordinates, clipped to the region
    wya = max(ya - p.y_lo, 0); wyb = min(yb - p.y_lo, p.rh);
    wxa = max(xa - p.x_lo, 0); wxb = min(xb - p.x_lo, p.rw);
    return wya < wyb && wxa < wxb;
}

constexpr int kZeroPage = 2048;   // bytes of zeros in shared memory, the source of the TMA fill

// shared -> global bulk copy (TMA, SASS: UBLKCP.G.S), tracked by the issuing thread's bulk group
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src_sa, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// zero `nbytes` bytes at `q` (any alignment) with one warp: 16-byte stores for the aligned middle, or --
// with the CTA's zero page at shared address `zero_sa` -- bulk copies of up to kZeroPage bytes issued by
// the first lanes (the TMA unit writes them; nothing waits in the SM's store queue)
__device__ __forceinline__ void warp_zero(unsigned char* q, int nbytes, int lane, unsigned zero_sa = 0u) {
    if (nbytes <= 0) return;
    const int head = min(nbytes, (int)((16u - (unsigned)(reinterpret_cast<uintptr_t>(q) & 15u)) & 15u));
    if (lane < head) q[lane] = 0;
    q += head;
    nbytes -= head;
    const int nv = nbytes >> 4;
    if (zero_sa) {
        const int mid = nv << 4;
        for (int o = lane * kZeroPage; o < mid; o += 32 * kZeroPage) bulk_s2g(q + o, zero_sa, (unsigned)min(kZeroPage, mid - o));
    } else {
        for (int i = lane; i < nv; i += 32) __stcs(reinterpret_cast<uint4*>(q) + i, make_uint4(0u, 0u, 0u, 0u));
    }
    if (lane < (nbytes & 15)) q[(nv << 4) + lane] = 0;
}

// FULL = false: only non-zero pixels of the window are written (the canvas was zeroed before).
// FULL = true: every element of the window's ROWS is written -- zeros outside the window's columns and
// where the sample misses the mask -- so the rows need no zero fill and the fill can run beside it.
template <int MODE, bool FULL>
__device__ __forceinline__ void paste_window_body(const PasteParams& p, int bxi, int nbx, int byi, int nby, unsigned zero_sa = 0u) {
    __shared__ __align__(16) float s_mask[kMaskStage];          // sigmoid(mask) rows, one-pixel zero border
    __shared__ __align__(8) float2 s_col[kColTab];              // per window column {VD index, wh}
    __shared__ __align__(8) float2 s_vd[kPasteThreads / 32][kVPairs];
    __shared__ __align__(16) float4 s_row[kBandRows];           // per band row {lo, wl, wh, state}
    constexpr int ES = MODE == DM_PASTE_F32 ? 4 : 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int st_w = p.sw + 2;
    const int zero_i = p.sw + 1, nan_i = p.sw + 2;  // VD entries for "exactly zero" / "NaN" columns
    const long long T = (long long)p.rh * p.rw;
    for (int n = byi; n < p.N; n += nby) {
        float4 bx;
        int wya, wyb, wxa, wxb;
        if (!window_rows(p, n, wya, wyb, wxa, wxb, bx)) continue;  // CTA-uniform
        if (wya + bxi * kBandRows >= wyb) continue;                // CTA-uniform: no band for this CTA
        // this CTA takes bands bxi, bxi + nbx, ... of the window's rows
      for (int r0 = wya + bxi * kBandRows; r0 < wyb; r0 += nbx * kBandRows) {
        const int r1 = min(r0 + kBandRows, wyb);      // exclusive
        const long long cls = p.labels ? p.labels[n] : 0;
        const float* __restrict__ m = p.masks + (long long)n * p.stride_n + cls * p.stride_c;
        // mask rows the band's canvas rows can reach (the source coordinate is monotone in py)
        bool staged = false;
        int mlo = 0, mtot = 0;
        {
            const float ia = src_coord(p.y_lo + r0, bx.y, bx.w, p.sh);
            const float ib = src_coord(p.y_lo + r1 - 1, bx.y, bx.w, p.sh);
            if (ia == ia && ib == ib && p.sw + 3 <= kVPairs) {
                mlo = (int)fmaxf(floorf(fminf(ia, ib)), -1.0f);
                const int hi = (int)fminf(floorf(fmaxf(ia, ib)) + 1.0f, (float)p.sh);
                const int mrows = hi - mlo + 1;
                if (mrows >= 1 && mrows * st_w <= kMaskStage) { staged = true; mtot = mrows * st_w; }
            }
        }
        const long long obase = (long long)n * T;
        if (!staged) {
            // direct path (small boxes: the band reaches more mask rows than the scratch holds; NaN
            // geometry): taps from global memory
            Instance in;
            in.load(p, n);
            for (int r = r0 + warp; r < r1; r += kPasteThreads / 32) {
                if (FULL) {
                    unsigned char* row8 = reinterpret_cast<unsigned char*>(p.out) + (obase + (long long)r * p.rw) * ES;
                    if (r == wya) warp_zero(row8, wxa * ES, lane, zero_sa);
                    warp_zero(row8 + (long long)wxb * ES, (p.rw - wxb + (r + 1 < wyb ? wxa : 0)) * ES, lane, zero_sa);
                }
                for (int c = wxa + lane; c < wxb; c += 32) {
                    const float v = in.eval(p.x_lo + c, p.y_lo + r);
                    if (FULL) put<MODE>(p, obase + (long long)r * p.rw + c, v != 0.0f ? v : 0.0f);
                    else if (v != 0.0f) put<MODE>(p, obase + (long long)r * p.rw + c, v);
                }
            }
            continue;
        }
        __syncthreads();  // the scratch of the previous instance is no longer read
        // ---- mask window: a warp per mask row, every load of the row issued before its first use ----
        {
            const int mrows = mtot / st_w;
            for (int yy = warp; yy < mrows; yy += kPasteThreads / 32) {
                const int y = mlo + yy;
                const bool yin = y >= 0 && y < p.sh;
                const float* __restrict__ mr = m + y * p.sw - 1;
                float* sr = s_mask + yy * st_w;
                float mv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int xx = lane + 32 * k;  // st_w <= kVPairs = 128
                    mv[k] = 0.0f;
                    if (yin && xx >= 1 && xx <= p.sw) mv[k] = __ldg(mr + xx);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int xx = lane + 32 * k;
                    if (xx < st_w) {
                        float v = mv[k];
                        if (p.apply_sigmoid) v = sigmoidf_fast(v);
                        sr[xx] = (yin && xx >= 1 && xx <= p.sw) ? v : 0.0f;
                    }
                }
            }
        }
        const int ww = wxb - wxa;
        const bool use_tab = ww <= kColTab;
        auto col_entry = [&](int c) -> float2 {  // c in region coordinates
            const AxisTerm a = axis_term(src_coord(p.x_lo + c, bx.x, bx.z, p.sw), p.sw);
            const int i = a.state == 1 ? a.lo + 1 : (a.state == 0 ? zero_i : nan_i);
            return make_float2(__int_as_float(i), a.wh);
        };
        if (use_tab)
            for (int c = threadIdx.x; c < ww; c += kPasteThreads) s_col[c] = col_entry(wxa + c);
        if (threadIdx.x < r1 - r0) {
            const AxisTerm a = axis_term(src_coord(p.y_lo + r0 + threadIdx.x, bx.y, bx.w, p.sh), p.sh);
            s_row[threadIdx.x] = make_float4(__int_as_float(a.lo), a.wl, a.wh, __int_as_float(a.state));
        }
        __syncthreads();
        // ---- a warp per canvas row: y pass into the warp's VD buffer, x pass to global memory -------
        float2* vd = s_vd[warp];
        for (int r = r0 + warp; r < r1; r += kPasteThreads / 32) {
            const float4 rt = s_row[r - r0];
            const int rstate = __float_as_int(rt.w);
            if (FULL) {
                // The zeros right of this row's window and left of the next row's are one contiguous run:
                // a row's warp writes [wxa, rw) of its row and [0, wxa) of the next window row (the first
                // window row also its own left part), so each row costs one zero run instead of two.
                unsigned char* row8 = reinterpret_cast<unsigned char*>(p.out) + (obase + (long long)r * p.rw) * ES;
                if (r == wya) warp_zero(row8, wxa * ES, lane, zero_sa);
                const int tail = p.rw - wxb + (r + 1 < wyb ? wxa : 0);
                if (rstate == 0) {  // the row's samples miss the mask: all zeros
                    warp_zero(row8 + (long long)wxa * ES, (wxb - wxa + tail) * ES, lane, zero_sa);
                    continue;
                }
                warp_zero(row8 + (long long)wxb * ES, tail * ES, lane, zero_sa);
            }
            if (rstate == 0) continue;  // warp-uniform
            __syncwarp();
            if (rstate == 1) {
                const float* m0 = s_mask + (__float_as_int(rt.x) - mlo) * st_w;
                for (int i = lane; i < st_w; i += 32) vd[i].x = rt.y * m0[i] + rt.z * m0[i + st_w];
            } else {
                for (int i = lane; i < st_w; i += 32) vd[i].x = __int_as_float(0x7fc00000);
            }
            __syncwarp();
            for (int i = lane; i <= p.sw; i += 32) vd[i].y = vd[i + 1].x - vd[i].x;
            if (lane == 0) {
                vd[zero_i] = make_float2(0.0f, 0.0f);
                vd[nan_i] = make_float2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
            }
            __syncwarp();
            // region rows are far below 2^31 elements: 32-bit offsets from the row's first element
            unsigned char* const orow8 = reinterpret_cast<unsigned char*>(p.out) + (obase + (long long)r * p.rw) * (MODE == DM_PASTE_F32 ? 4 : 1);
            auto emit = [&](int c, const float2 e) {
                const float2 pr = vd[__float_as_int(e.x)];
                const float v = fmaf(e.y, pr.y, pr.x);
                if (MODE == DM_PASTE_F32) {
                    if (FULL) reinterpret_cast<float*>(orow8)[c] = v != 0.0f ? v : 0.0f;
                    else if (v != 0.0f) reinterpret_cast<float*>(orow8)[c] = v;
                } else {
                    const uint32_t b = encode<MODE>(v, p.thr);
                    if (FULL || b) orow8[c] = (unsigned char)b;
                }
            };
            if (use_tab) {
                const float2* ct = s_col - wxa;
                int c = wxa + lane;
                for (; c + 96 < wxb; c += 128) {
                    const float2 e0 = ct[c], e1 = ct[c + 32], e2 = ct[c + 64], e3 = ct[c + 96];
                    emit(c, e0); emit(c + 32, e1); emit(c + 64, e2); emit(c + 96, e3);
                }
                for (; c < wxb; c += 32) emit(c, ct[c]);
            } else {
                for (int c = wxa + lane; c < wxb; c += 32) emit(c, col_entry(c));
            }
        }
      }  // bands
    }
}

template <int MODE>
__global__ void __launch_bounds__(kPasteThreads, 5)
paste_window_kernel(const __grid_constant__ PasteParams p) {
    paste_window_body<MODE, false>(p, (int)blockIdx.x, (int)gridDim.x, (int)blockIdx.y, (int)gridDim.y);
}

// Zero fill and windows in ONE launch over disjoint rows.  Window CTAs (8 per instance slot, as in the
// two-launch form) own every element of their instances' window rows; the zero fill of everything
// else is dealt to the same CTAs in 64 KB tiles, which each CTA issues FIRST: the stores are
// fire-and-forget, so they drain to HBM while the CTA goes on to its instruction-bound window work.
// (A split into fill CTAs and window CTAs shares the 5 CTA slots of an SM between the two roles and
// slows the window role down by what the fill role occupies: 183 us vs 227 us for two launches.)
constexpr int kFillVecs = 16;                              // 16-byte stores per fill thread
constexpr int kFillTile = kPasteThreads * kFillVecs * 16;   // bytes per fill tile (64 KB)

// The byte range of instance n's window rows (what the fill skips), absolute in the output; empty
// (0, 0) when the instance is not pasted or lies behind the output.
template <int ES>
__device__ __forceinline__ void skip_range(const PasteParams& p, long long n, long long t0, bool second,
                                           long long& lo, long long& hi) {
    const long long TB = (long long)p.rh * p.rw * ES;   // bytes per instance (>= kFillTile on this path)
    lo = hi = 0;
    float4 bx;
    int wya, wyb, wxa, wxb;
    if (n >= p.N || (second && n * TB >= t0 + kFillTile)) return;
    if (window_rows(p, (int)n, wya, wyb, wxa, wxb, bx)) {
        lo = n * TB + (long long)wya * p.rw * ES;
        hi = n * TB + (long long)wyb * p.rw * ES;
    }
}

// One 64 KB tile of the zero fill.  The tile touches at most two instances; [a_lo, a_hi) and
// [b_lo, b_hi) are their skipped byte ranges (window rows).
template <int ES>
__device__ __forceinline__ void fill_tile(const PasteParams& p, long long tile, long long nbytes, long long a_lo,
                                          long long a_hi, long long b_lo, long long b_hi, unsigned zero_sa) {
    const long long t0 = tile * kFillTile;
    unsigned char* const out8 = reinterpret_cast<unsigned char*>(p.out);
    // CTA-uniform fast paths: a tile wholly inside window rows has nothing to do, a whole tile clear
    // of them is 16 plain streaming stores per thread; only tiles on a boundary test every vector
    const long long t1 = min(t0 + kFillTile, nbytes);
    if ((t0 >= a_lo && t1 <= a_hi) || (t0 >= b_lo && t1 <= b_hi)) return;
    if (t1 - t0 == kFillTile && (t1 <= a_lo || t0 >= a_hi) && (t1 <= b_lo || t0 >= b_hi)) {
        if (zero_sa) {
            // TMA fill: the tile leaves as 32 bulk copies of the CTA's zero page (shared -> global), one per
            // lane of warp 0.  The copies are carried out by the TMA unit, not by the LSU: the SM's
            // load / store queues stay free for the window role that follows, where 16 x 8 warp-wide
            // streaming stores per tile used to sit in them until HBM had taken the bytes.
            if (threadIdx.x < kFillTile / kZeroPage)
                bulk_s2g(out8 + t0 + (long long)threadIdx.x * kZeroPage, zero_sa, kZeroPage);
            return;
        }
        uint4* q = reinterpret_cast<uint4*>(out8 + t0) + threadIdx.x;
#pragma unroll
        for (int k = 0; k < kFillVecs; ++k) __stcs(q + k * kPasteThreads, make_uint4(0u, 0u, 0u, 0u));
        return;
    }
#pragma unroll 4
    for (int k = 0; k < kFillVecs; ++k) {
        const long long o = t0 + ((long long)k * kPasteThreads + threadIdx.x) * 16;
        if (o >= nbytes) break;
        const long long e = min(o + 16, nbytes);
        if ((o >= a_lo && e <= a_hi) || (o >= b_lo && e <= b_hi)) continue;          // inside window rows
        if (e - o == 16 && (e <= a_lo || o >= a_hi) && (e <= b_lo || o >= b_hi)) {   // clear of them
            __stcs(reinterpret_cast<uint4*>(out8 + o), make_uint4(0u, 0u, 0u, 0u));
            continue;
        }
        for (long long q = o; q < e; ++q)
            if (!(q >= a_lo && q < a_hi) && !(q >= b_lo && q < b_hi)) out8[q] = 0;
    }
}

// Persistent: the launch holds as many CTAs as the GPU keeps resident (5 per SM) and each walks the
// "virtual CTAs" v = blockIdx.x, blockIdx.x + gridDim.x, ... of the 8-per-instance-slot decomposition:
// the fill tiles of v first, then v's window bands.  (One CTA per virtual CTA was 6400 CTAs of 43 KB of
// shared memory for the C4 shape: 47 us of the launch went into starting and retiring CTAs that found
// nothing to paste -- measured with both roles switched off, DM_PASTE_DIAG=3.)
template <int MODE>
__global__ void __launch_bounds__(kPasteThreads, 5)
paste_fused_kernel(const __grid_constant__ PasteParams p, int win_y, int n_virtual, long long n_fill, long long nbytes,
                   int tma_fill, int diag) {
    constexpr int ES = MODE == DM_PASTE_F32 ? 4 : 1;
    // The skip ranges of a virtual CTA's tiles (two box loads and two window computations per tile) are
    // worked out ONCE by two threads per tile and handed to the others through shared memory: with every
    // thread deriving them the fill role spent ~20 % of the kernel's instructions on geometry.
    constexpr int kSlots = 16;
    __shared__ long long s_rng[kSlots][4];
    __shared__ __align__(128) unsigned char s_zero[kZeroPage];
    const long long TB = (long long)p.rh * p.rw * ES;
    unsigned zero_sa = 0u;
    if (tma_fill) {
        if (threadIdx.x < kZeroPage / 16) reinterpret_cast<uint4*>(s_zero)[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async-proxy reads
        zero_sa = (unsigned)__cvta_generic_to_shared(s_zero);
    }
    for (int v = blockIdx.x; v < n_virtual; v += gridDim.x) {
        __syncthreads();   // s_rng of the previous round is no longer read
        {
            const int k = threadIdx.x >> 1, which = threadIdx.x & 1;
            const long long tile = v + (long long)k * n_virtual;
            if (k < kSlots && tile < n_fill) {
                const long long t0 = tile * kFillTile;
                long long lo, hi;
                skip_range<ES>(p, t0 / TB + which, t0, which != 0, lo, hi);
                s_rng[k][2 * which] = lo;
                s_rng[k][2 * which + 1] = hi;
            }
        }
        __syncthreads();
        int k = 0;
        for (long long tile = v; tile < n_fill && !(diag & 2); tile += n_virtual, ++k) {   // (diag: measurement only)
            long long a_lo, a_hi, b_lo, b_hi;
            if (k < kSlots) {
                a_lo = s_rng[k][0]; a_hi = s_rng[k][1]; b_lo = s_rng[k][2]; b_hi = s_rng[k][3];
            } else {
                const long long t0 = tile * kFillTile;
                skip_range<ES>(p, t0 / TB, t0, false, a_lo, a_hi);
                skip_range<ES>(p, t0 / TB + 1, t0, true, b_lo, b_hi);
            }
            fill_tile<ES>(p, tile, nbytes, a_lo, a_hi, b_lo, b_hi, zero_sa);
        }
        if (!(diag & 1)) paste_window_body<MODE, true>(p, v % kBandCtas, kBandCtas, v / kBandCtas, win_y, (diag & 4) ? 0u : zero_sa);
    }
    // the zero page must outlive the copies that read it
    if (tma_fill) { bulk_commit(); bulk_wait_all(); }
}

}  // namespace dm

static int paste_impl(const float* masks, int64_t mask_stride_n, int64_t mask_stride_c,
                      const int64_t* labels, int N, int S_h, int S_w, int apply_sigmoid,
                      const float* boxes, int img_h, int img_w, int x_lo, int y_lo, int x_hi, int y_hi,
                      float thr, int out_mode, const int32_t* select, int select_value, int zero_fill,
                      void* out, dm_stream_t stream) {
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
    p.select = select;
    p.select_value = select_value;
    p.total = per_inst * N;
    const int ES = out_mode == DM_PASTE_F32 ? 4 : 1;
    cudaStream_t st = (cudaStream_t)stream;
    const long long nbytes_all = p.total * ES;
    // A/B knobs, read once: DM_PASTE_FUSED=0 takes the two-launch form, DM_PASTE_TMA=0 fills with plain stores
    static const bool fused_on = [] { const char* e = getenv("DM_PASTE_FUSED"); return !(e && *e == '0'); }();
    static const int tma_fill = [] { const char* e = getenv("DM_PASTE_TMA"); return (e && *e == '0') ? 0 : 1; }();
    static const int diag = [] { const char* e = getenv("DM_PASTE_DIAG"); return e ? atoi(e) : 0; }();   // 1: no windows, 2: no fill
    const bool fused = zero_fill && per_inst * ES >= dm::kFillTile && fused_on;
    if (fused) {
        // one launch: every CTA issues its share of the zero fill, then pastes its window bands
        const int win_y = N < 65535 ? N : 65535;
        const int n_virtual = dm::kBandCtas * win_y;
        const long long n_fill = (nbytes_all + dm::kFillTile - 1) / dm::kFillTile;
        int dev = 0, sms = 0;
        DM_CUDA_CHECK(cudaGetDevice(&dev), "dm_paste_masks/fused");
        DM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "dm_paste_masks/fused");
        static const int persist = [] { const char* e = getenv("DM_PASTE_PERSIST"); return e ? atoi(e) : 0; }();   // > 0: persistent launch with this many CTAs per SM (measured slower: static round robin loses the hardware scheduler's load balancing, 221 vs 176 us)
        const unsigned grid = persist > 0 && (long long)sms * persist < n_virtual ? (unsigned)(sms * persist) : (unsigned)n_virtual;
        switch (out_mode) {
            case DM_PASTE_BOOL: dm::paste_fused_kernel<DM_PASTE_BOOL><<<grid, dm::kPasteThreads, 0, st>>>(p, win_y, n_virtual, n_fill, nbytes_all, tma_fill, diag); break;
            case DM_PASTE_U8: dm::paste_fused_kernel<DM_PASTE_U8><<<grid, dm::kPasteThreads, 0, st>>>(p, win_y, n_virtual, n_fill, nbytes_all, tma_fill, diag); break;
            default: dm::paste_fused_kernel<DM_PASTE_F32><<<grid, dm::kPasteThreads, 0, st>>>(p, win_y, n_virtual, n_fill, nbytes_all, tma_fill, diag); break;
        }
        DM_LAUNCH_CHECK("dm_paste_masks/fused");
        return DM_OK;
    }
    if (zero_fill) {
        const long long nbytes = p.total * ES, n16 = nbytes / 16;
        const long long blocks = (n16 + 1 + 255) / 256;
        if (blocks >= (1ll << 31)) return DM_EUNSUPPORTED;
        dm::paste_fill_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<uint4*>(out), n16, nbytes);
        DM_LAUNCH_CHECK("dm_paste_masks/fill");
    }
    // 8 CTAs per instance, each taking every 8th band of the window: few CTAs find nothing to do
    const int bands = (p.rh + dm::kBandRows - 1) / dm::kBandRows;
    dim3 grid((unsigned)(bands < dm::kBandCtas ? bands : dm::kBandCtas), (unsigned)(N < 65535 ? N : 65535));
    switch (out_mode) {
        case DM_PASTE_BOOL: dm::paste_window_kernel<DM_PASTE_BOOL><<<grid, dm::kPasteThreads, 0, st>>>(p); break;
        case DM_PASTE_U8: dm::paste_window_kernel<DM_PASTE_U8><<<grid, dm::kPasteThreads, 0, st>>>(p); break;
        default: dm::paste_window_kernel<DM_PASTE_F32><<<grid, dm::kPasteThreads, 0, st>>>(p); break;
    }
    DM_LAUNCH_CHECK("dm_paste_masks");
    return DM_OK;
}

extern "C" int dm_paste_masks(const float* masks, int64_t mask_stride_n, int64_t mask_stride_c,
                              const int64_t* labels, int N, int S_h, int S_w, int apply_sigmoid,
                              const float* boxes, int img_h, int img_w, int x_lo, int y_lo,
                              int x_hi, int y_hi, float thr, int out_mode, void* out,
                              dm_stream_t stream) {
    return paste_impl(masks, mask_stride_n, mask_stride_c, labels, N, S_h, S_w, apply_sigmoid, boxes, img_h,
                      img_w, x_lo, y_lo, x_hi, y_hi, thr, out_mode, nullptr, 0, 1, out, stream);
}

extern "C" int dm_paste_masks_select(const float* masks, int64_t mask_stride_n, int64_t mask_stride_c,
                                     const int64_t* labels, int N, int S_h, int S_w, int apply_sigmoid,
                                     const float* boxes, int img_h, int img_w, int x_lo, int y_lo,
                                     int x_hi, int y_hi, float thr, int out_mode, const int32_t* select,
                                     int select_value, int zero_fill, void* out, dm_stream_t stream) {
    if (!select) return DM_EINVAL;
    return paste_impl(masks, mask_stride_n, mask_stride_c, labels, N, S_h, S_w, apply_sigmoid, boxes, img_h,
                      img_w, x_lo, y_lo, x_hi, y_hi, thr, out_mode, select, select_value, zero_fill, out, stream);
}
