// Next row of the path: run-length encoding of the pasted instance
// masks on the device, so the [N,H,W] canvases never cross PCIe -- or, fused with the paste, never
// exist at all.
//
// Replaces the host loop  encode_mask_results(get_seg_masks(...))  of the reference:
//   DynaMaskHead.get_seg_masks -> N numpy bool [H,W]   (mmdet/models/roi_heads/mask_heads/dynamask_head.py:341)
//   encode_mask_results -> pycocotools mask.encode     (mmdet/core/mask/utils.py:36-63, mmdet/apis/test.py:54-57)
// pycocotools (third party, not in the reference tree) encodes a mask in COLUMN-major order as
// alternating run lengths starting with a run of zeros.  A run list is fully described by the
// sorted flat indices (x * H + y) at which the value changes, so the kernels emit those
// "transitions"; turning them into counts and into pycocotools' compressed string is a few
// microseconds of host work per instance (dm_rle_compress_host below).
//
// Two passes over the same pixels (count, then write at exclusive-scan offsets):
//   dm_paste_rle        pixels are evaluated from the mask logits exactly as paste_window_kernel does
//                       (same tables, same FMA), one thread per canvas column walking down the
//                       instance's window; the canvas is never written.
//   dm_rle_from_canvas  pixels are read from an existing [N,H,W] uint8/bool canvas.
// Every column is encoded on its own, starting from 0 and returning to 0 after its last window row;
// where a run really continues into the next column the two coinciding transitions cancel on the
// host.
#include <cstdlib>
#include <cstring>

#include "dm_paste_common.cuh"

namespace dm {

struct RleParams {
    PasteParams p;
    int32_t* col_counts;          // [N][nseg][rw] transitions per (row segment, column); window columns only are defined
    int nseg, seg_rows;           // row segments of the region (grid.z): a column's rows are shared out over nseg CTAs
    int32_t* inst_totals;         // [N] transitions per instance (pass 1 adds into it: zero it first)
    const int64_t* inst_offsets;  // [N] exclusive scan of inst_totals (pass 2)
    int32_t* trans;               // [sum] transitions, instance after instance (pass 2)
    const uint8_t* canvas;        // canvas source, or null
    int32_t* slots;               // optional [N][rw][kRleSlots]: pass 1 also RECORDS a column's first kRleSlots transitions, so
                                  // that pass 3 (a copy) replaces the second evaluation for every instance whose columns all fit
    int32_t* inst_over;           // [N][nseg][gridDim.x] set by pass 1 when a column of the CTA's 256 has more transitions than slots
    const int32_t* status;        // optional device word: non-zero = the transitions do not fit the caller's buffers,
                                  // pass 2 writes nothing (dm_paste_rle_strings)
};

constexpr int kRleThreads = 256;
constexpr int kRleSlots = 32;  // transitions per column recorded in pass 1 (128 bytes per column; a clean mask has 2-4 per
                               // column, the benchmark's noisy synthetic logits ~16-20)

// exclusive scan of one int per thread over the CTA; returns the thread's prefix, `total` = CTA sum
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();  // s_warp may still be read from a previous call
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int base = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < kRleThreads / 32; ++w) {
        const int t = s_warp[w];
        if (w < warp) base += t;
        total += t;
    }
    return base + inc - v;
}

template <int PASS>
__global__ void __launch_bounds__(kRleThreads, 4)
paste_rle_kernel(const __grid_constant__ RleParams q) {
    const PasteParams& p = q.p;
    __shared__ __align__(16) float s_mask[kMaskStage];
    __shared__ __align__(8) float2 s_vd[kRleThreads / 32][kVPairs];
    __shared__ int s_warp[kRleThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int st_w = p.sw + 2;
    const int zero_i = p.sw + 1, nan_i = p.sw + 2;
    if (PASS >= 2 && q.status && *q.status) return;   // CTA-uniform
    // Row segments: the chain of a column -- one thread walking down the window, 8 rows per barrier pair -- is what
    // a launch waits for (a 500-row window: ~100 us with ~100 live CTAs on 148 SMs), so the region's rows are cut
    // into nseg segments (grid.z) and every (column block, segment) is a CTA of its own.  A segment that does not
    // start the window evaluates the row above it first to know the state it continues from; only the segment
    // holding the window's last row closes the column.
    const int seg = blockIdx.z, nseg = q.nseg;
    auto cc = [&](int n_, int s_, int x_) -> int32_t& { return q.col_counts[((size_t)n_ * nseg + s_) * p.rw + x_]; };
    for (int n = blockIdx.y; n < p.N; n += gridDim.y) {
        // with recorded slots: pass 3 copies the column blocks whose columns all fitted, pass 2 re-evaluates the others
        if (PASS >= 2 && q.inst_over &&
            (q.inst_over[((size_t)n * nseg + seg) * gridDim.x + blockIdx.x] != 0) != (PASS == 2)) continue;   // CTA-uniform
        const float4 bx = *reinterpret_cast<const float4*>(p.boxes + 4 * (size_t)n);
        int xa, xb, ya, yb;
        window_1d(bx.x, bx.z, p.sw, p.img_w, xa, xb);
        const int wxa = max(xa - p.x_lo, 0), wxb = min(xb - p.x_lo, p.rw);
        const int c0 = wxa + blockIdx.x * kRleThreads;  // this CTA's first column
        if (c0 >= wxb) continue;                         // CTA-uniform
        window_1d(bx.y, bx.w, p.sh, p.img_h, ya, yb);
        const int wya = max(ya - p.y_lo, 0), wyb = min(yb - p.y_lo, p.rh);
        // this CTA's rows [sa, sb) of the window
        const int sa = max(wya, seg * q.seg_rows), sb = min(wyb, (seg + 1) * q.seg_rows);
        if (sa >= sb) continue;                          // CTA-uniform (pass 1 leaves the zeroed count in place)
        const int x = c0 + threadIdx.x;
        const bool live = x < wxb;
        // this thread's column: VD index and weight, constant down the column
        int ci = zero_i;
        float cw = 0.0f;
        if (live) {
            const AxisTerm a = axis_term(src_coord(p.x_lo + x, bx.x, bx.z, p.sw), p.sw);
            ci = a.state == 1 ? a.lo + 1 : (a.state == 0 ? zero_i : nan_i);
            cw = a.wh;
        }
        int32_t* out = nullptr;
        if (PASS >= 2) {
            // transitions of the instance's earlier columns: those of earlier CTAs + a scan inside this one
            int part = 0;
            for (int c = wxa + threadIdx.x; c < c0; c += kRleThreads)
                for (int s2 = 0; s2 < nseg; ++s2) part += cc(n, s2, c);
            int before = 0, dummy = 0;
            block_exclusive_scan(part, s_warp, before);
            int col_total = 0, above = 0;                // the column's transitions in all segments / in the segments above
            if (live)
                for (int s2 = 0; s2 < nseg; ++s2) {
                    const int v = cc(n, s2, x);
                    col_total += v;
                    if (s2 < seg) above += v;
                }
            const int mine = live ? cc(n, seg, x) : 0;
            const int pre = block_exclusive_scan(col_total, s_warp, dummy);
            out = q.trans + q.inst_offsets[n] + before + pre + above;
            if (PASS == 3) {   // the column's transitions were recorded by pass 1: copy them into place
                if (live) {
                    const int32_t* sl = q.slots + (((size_t)n * nseg + seg) * p.rw + x) * kRleSlots;
                    for (int k = 0; k < mine; ++k) out[k] = sl[k];
                }
                continue;
            }
        }
        const long long cls = p.labels ? p.labels[n] : 0;
        const float* __restrict__ m = p.masks + (long long)n * p.stride_n + cls * p.stride_c;
        // band height: as many canvas rows as keep the reachable mask rows inside the scratch
        int hb = kRleBandRows;
        {
            const float ratio = fabsf((float)p.sh / (bx.w - bx.y));  // mask rows per canvas row
            const int cap = kMaskStage / st_w - 3;
            if (ratio == ratio && ratio * (float)hb > (float)cap) hb = max(1, (int)((float)cap / ratio));
        }
        Instance in;
        in.load(p, n);
        int prev = 0, cnt = 0;
        const int col_base = x * p.rh;
        int32_t* const rec = (PASS == 1 && q.slots && live) ? q.slots + (((size_t)n * nseg + seg) * p.rw + x) * kRleSlots : nullptr;
        const int r_first = sa > wya ? sa - 1 : sa;      // a continuing segment looks at the row above it first
        bool peek = r_first < sa;
        auto step = [&](int row, bool b) {
            if (peek) {                                  // (the row above the segment: state only, nothing emitted)
                prev = (int)b;
                peek = false;
                return;
            }
            if ((int)b != prev) {
                if (PASS == 1) {
                    if (rec && cnt < kRleSlots) rec[cnt] = col_base + row;
                    ++cnt;
                } else {
                    *out++ = col_base + row;
                }
                prev = (int)b;
            }
        };
        for (int r0 = r_first; r0 < sb; r0 += hb) {
            const int r1 = min(r0 + hb, sb);
            bool staged = false;
            int mlo = 0, mrows = 0;
            {
                const float ia = src_coord(p.y_lo + r0, bx.y, bx.w, p.sh);
                const float ib = src_coord(p.y_lo + r1 - 1, bx.y, bx.w, p.sh);
                if (ia == ia && ib == ib && p.sw + 3 <= kVPairs) {
                    mlo = (int)fmaxf(floorf(fminf(ia, ib)), -1.0f);
                    const int hi = (int)fminf(floorf(fmaxf(ia, ib)) + 1.0f, (float)p.sh);
                    mrows = hi - mlo + 1;
                    staged = mrows >= 1 && mrows * st_w <= kMaskStage;
                }
            }
            if (!staged) {  // CTA-uniform: per-pixel evaluation from global memory
                for (int row = r0; row < r1; ++row) step(row, live && in.eval(p.x_lo + x, p.y_lo + row) >= p.thr);
                continue;
            }
            __syncthreads();  // scratch of the previous band is no longer read
            for (int yy = warp; yy < mrows; yy += kRleThreads / 32) {
                const int y = mlo + yy;
                const bool yin = y >= 0 && y < p.sh;
                const float* __restrict__ mr = m + y * p.sw - 1;
                float* sr = s_mask + yy * st_w;
                float mv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int xx = lane + 32 * k;
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
            __syncthreads();
            for (int s0 = r0; s0 < r1; s0 += kRleThreads / 32) {
                // warp w prepares the (value, slope) row of canvas row s0 + w
                const int row = s0 + warp;
                float2* vd = s_vd[warp];
                if (row < r1) {
                    const AxisTerm ry = axis_term(src_coord(p.y_lo + row, bx.y, bx.w, p.sh), p.sh);
                    if (ry.state == 1) {
                        const float* m0 = s_mask + (ry.lo - mlo) * st_w;
                        for (int i = lane; i < st_w; i += 32) vd[i].x = ry.wl * m0[i] + ry.wh * m0[i + st_w];
                    } else {
                        const float f = ry.state == 0 ? 0.0f : __int_as_float(0x7fc00000);
                        for (int i = lane; i < st_w; i += 32) vd[i].x = f;
                    }
                    __syncwarp();
                    for (int i = lane; i <= p.sw; i += 32) vd[i].y = vd[i + 1].x - vd[i].x;
                    if (lane == 0) {
                        vd[zero_i] = make_float2(0.0f, 0.0f);
                        vd[nan_i] = make_float2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
                    }
                }
                __syncthreads();
                const int nk = min(kRleThreads / 32, r1 - s0);
                for (int k = 0; k < nk; ++k) {
                    const float2 pr = s_vd[k][ci];
                    const float v = fmaf(cw, pr.y, pr.x);
                    step(s0 + k, live && v >= p.thr);
                }
                __syncthreads();
            }
        }
        if (prev && sb == wyb) step(wyb, false);  // every column returns to 0 after its last window row
        if (PASS == 1) {
            if (live) cc(n, seg, x) = cnt;
            if (q.inst_over && cnt > kRleSlots) q.inst_over[((size_t)n * nseg + seg) * gridDim.x + blockIdx.x] = 1;
            int total = 0;
            block_exclusive_scan(cnt, s_warp, total);
            if (threadIdx.x == 0 && total) atomicAdd(q.inst_totals + n, total);
        }
    }
}

// Same encoding from an existing canvas [N][H][W] (one byte per pixel, non-zero = foreground).
template <int PASS>
__global__ void __launch_bounds__(kRleThreads)
canvas_rle_kernel(const __grid_constant__ RleParams q) {
    const PasteParams& p = q.p;  // only N, rh, rw are used
    __shared__ int s_warp[kRleThreads / 32];
    const int c0 = blockIdx.x * kRleThreads;
    for (int n = blockIdx.y; n < p.N; n += gridDim.y) {
        const int x = c0 + threadIdx.x;
        const bool live = x < p.rw;
        int32_t* out = nullptr;
        if (PASS == 2) {
            int part = 0;
            for (int c = threadIdx.x; c < c0; c += kRleThreads) part += q.col_counts[(size_t)n * p.rw + c];
            int before = 0, dummy = 0;
            block_exclusive_scan(part, s_warp, before);
            const int mine = live ? q.col_counts[(size_t)n * p.rw + x] : 0;
            const int pre = block_exclusive_scan(mine, s_warp, dummy);
            out = q.trans + q.inst_offsets[n] + before + pre;
        }
        const uint8_t* __restrict__ src = q.canvas + (size_t)n * p.rh * p.rw + x;
        int prev = 0, cnt = 0;
        const int col_base = x * p.rh;
        if (live) {
            for (int y = 0; y < p.rh; ++y) {
                const int b = __ldcs(src + (size_t)y * p.rw) != 0;
                if (b != prev) {
                    if (PASS == 1) ++cnt;
                    else *out++ = col_base + y;
                    prev = b;
                }
            }
            if (prev) {
                if (PASS == 1) ++cnt;
                else *out++ = col_base + p.rh;
            }
        }
        if (PASS == 1) {
            if (live) q.col_counts[(size_t)n * p.rw + x] = cnt;
            int total = 0;
            block_exclusive_scan(cnt, s_warp, total);
            if (threadIdx.x == 0 && total) atomicAdd(q.inst_totals + n, total);
        }
    }
}

}  // namespace dm

static int rle_fill(dm::RleParams& q, int N, int rh, int rw, int pass, int32_t* col_counts, int32_t* inst_totals,
                    const int64_t* inst_offsets, int32_t* transitions) {
    if (pass < 1 || pass > 3) return DM_EINVAL;   // (3: internal, dm_paste_rle_strings)
    if (!col_counts || !inst_totals) return DM_EINVAL;
    if (pass >= 2 && (!inst_offsets || !transitions)) return DM_EINVAL;
    if ((long long)rh * rw >= (1ll << 30)) return DM_EUNSUPPORTED;
    q.p.N = N;
    q.p.rh = rh;
    q.p.rw = rw;
    q.col_counts = col_counts;
    q.nseg = 1;
    q.seg_rows = rh > 0 ? rh : 1;
    q.inst_totals = inst_totals;
    q.inst_offsets = inst_offsets;
    q.trans = transitions;
    q.canvas = nullptr;
    return DM_OK;
}

static int paste_rle_impl(const float* masks, int64_t mask_stride_n, int64_t mask_stride_c,
                          const int64_t* labels, int N, int S_h, int S_w, int apply_sigmoid,
                          const float* boxes, int img_h, int img_w, int x_lo, int y_lo, int x_hi,
                          int y_hi, float thr, int pass, int32_t* col_counts, int32_t* inst_totals,
                          const int64_t* inst_offsets, int32_t* transitions, const int32_t* status, int32_t* slots,
                          int32_t* inst_over, int nseg, int seg_rows, dm_stream_t stream) {
    if (N < 0 || S_h < 1 || S_w < 1 || img_h < 0 || img_w < 0) return DM_EINVAL;
    if (x_lo < 0 || y_lo < 0 || x_hi > img_w || y_hi > img_h || x_hi < x_lo || y_hi < y_lo) return DM_EINVAL;
    dm::RleParams q;
    memset(&q, 0, sizeof(q));
    const int rc = rle_fill(q, N, y_hi - y_lo, x_hi - x_lo, pass, col_counts, inst_totals, inst_offsets, transitions);
    if (rc != DM_OK) return rc;
    if (N == 0 || q.p.rh == 0 || q.p.rw == 0) return DM_OK;
    if (!masks || !boxes || (reinterpret_cast<uintptr_t>(boxes) & 15u)) return DM_EINVAL;
    q.p.masks = masks;
    q.p.stride_n = mask_stride_n;
    q.p.stride_c = mask_stride_c;
    q.p.labels = labels;
    q.p.sh = S_h;
    q.p.sw = S_w;
    q.p.apply_sigmoid = apply_sigmoid;
    q.p.boxes = boxes;
    q.p.img_h = img_h;
    q.p.img_w = img_w;
    q.p.x_lo = x_lo;
    q.p.y_lo = y_lo;
    q.p.thr = thr;
    q.status = status;
    q.slots = slots;
    q.inst_over = inst_over;
    if (nseg > 1) { q.nseg = nseg; q.seg_rows = seg_rows; }
    q.p.total = (long long)q.p.rh * q.p.rw * N;
    dim3 grid((unsigned)((q.p.rw + dm::kRleThreads - 1) / dm::kRleThreads), (unsigned)(N < 65535 ? N : 65535), (unsigned)q.nseg);
    cudaStream_t st = (cudaStream_t)stream;
    if (pass == 1) dm::paste_rle_kernel<1><<<grid, dm::kRleThreads, 0, st>>>(q);
    else if (pass == 3) dm::paste_rle_kernel<3><<<grid, dm::kRleThreads, 0, st>>>(q);
    else dm::paste_rle_kernel<2><<<grid, dm::kRleThreads, 0, st>>>(q);
    DM_LAUNCH_CHECK("dm_paste_rle");
    return DM_OK;
}

extern "C" int dm_paste_rle(const float* masks, int64_t mask_stride_n, int64_t mask_stride_c,
                            const int64_t* labels, int N, int S_h, int S_w, int apply_sigmoid,
                            const float* boxes, int img_h, int img_w, int x_lo, int y_lo, int x_hi,
                            int y_hi, float thr, int pass, int32_t* col_counts, int32_t* inst_totals,
                            const int64_t* inst_offsets, int32_t* transitions, dm_stream_t stream) {
    if (pass != 1 && pass != 2) return DM_EINVAL;
    return paste_rle_impl(masks, mask_stride_n, mask_stride_c, labels, N, S_h, S_w, apply_sigmoid, boxes, img_h, img_w,
                          x_lo, y_lo, x_hi, y_hi, thr, pass, col_counts, inst_totals, inst_offsets, transitions, nullptr,
                          nullptr, nullptr, 1, 0, stream);
}

extern "C" int dm_rle_from_canvas(const uint8_t* canvas, int N, int H, int W, int pass, int32_t* col_counts,
                                  int32_t* inst_totals, const int64_t* inst_offsets, int32_t* transitions,
                                  dm_stream_t stream) {
    if (N < 0 || H < 0 || W < 0) return DM_EINVAL;
    dm::RleParams q;
    memset(&q, 0, sizeof(q));
    const int rc = rle_fill(q, N, H, W, pass, col_counts, inst_totals, inst_offsets, transitions);
    if (rc != DM_OK) return rc;
    if (N == 0 || H == 0 || W == 0) return DM_OK;
    if (!canvas) return DM_EINVAL;
    q.canvas = canvas;
    dim3 grid((unsigned)((W + dm::kRleThreads - 1) / dm::kRleThreads), (unsigned)(N < 65535 ? N : 65535));
    cudaStream_t st = (cudaStream_t)stream;
    if (pass == 1) dm::canvas_rle_kernel<1><<<grid, dm::kRleThreads, 0, st>>>(q);
    else dm::canvas_rle_kernel<2><<<grid, dm::kRleThreads, 0, st>>>(q);
    DM_LAUNCH_CHECK("dm_rle_from_canvas");
    return DM_OK;
}

// ---------------------------------------------------------------------------------------------
// Device side: transitions -> pycocotools' compressed "counts" strings (rleToString), so that only
// the strings cross PCIe and the host does no per-run work.  One CTA per instance, three launches:
//   compact   drops the coinciding transition pairs (a run continuing across a column boundary;
//             they only ever come in pairs, so "has an equal neighbour" is the host's
//             skip-both rule), writes the survivors to `compact` and the string length to str_len
//   scan      exclusive scan of the N string lengths (one CTA) -> str_offsets[N+1]
//   write     count k = t'[k] - t'[k-1] (and the closing run); from the fourth count on the
//             difference to the count two before is stored; base-32 varint, 6 bits per char + 48
// ---------------------------------------------------------------------------------------------
namespace dm {

// chars of rleToString's code for x (signed base-32 varint, continuation bit 0x20)
__device__ __forceinline__ int rle_code_len(long long x) {
    int n = 0;
    bool more = true;
    while (more) {
        const int c = (int)(x & 0x1f);
        x >>= 5;
        more = (c & 0x10) ? x != -1 : x != 0;
        ++n;
    }
    return n;
}

__device__ __forceinline__ void rle_code_write(long long x, char* out) {
    bool more = true;
    while (more) {
        int c = (int)(x & 0x1f);
        x >>= 5;
        more = (c & 0x10) ? x != -1 : x != 0;
        if (more) c |= 0x20;
        *out++ = (char)(c + 48);
    }
}

// Value stored for count k of an instance whose surviving transitions are tp[0..m): with
// t(i) = 0 for i < 0, tp[i] for i < m, total_pixels beyond, count(k) = t(k) - t(k-1) (k == m is the
// closing run); the string holds count(k) for k < 3, else count(k) - count(k-2).

// Four consecutive items per thread and step: the kernels are chains of block scans (a few thousand transitions
// per instance, one CTA per instance), so fewer, fatter steps is what shortens them (50 + 43 us -> see profiles/).
constexpr int kRleItems = 4;

template <int PASS>   // 1: compact + lengths, 2: write strings
__global__ void __launch_bounds__(kRleThreads)
rle_string_kernel(const int32_t* __restrict__ trans, const int64_t* __restrict__ inst_offsets, long long total_pixels,
                  int32_t* __restrict__ compact, int32_t* __restrict__ kept, int32_t* __restrict__ str_len,
                  const int64_t* __restrict__ str_offsets, char* __restrict__ out, const int32_t* __restrict__ status) {
    __shared__ int s_warp[kRleThreads / 32];
    if (status && *status) return;   // CTA-uniform: nothing was written by pass 2
    const int n = blockIdx.x;
    const long long o0 = inst_offsets[n];
    const int cnt = (int)(inst_offsets[n + 1] - o0);
    const int32_t* t = trans + o0;
    int32_t* tp = compact + o0;
    constexpr int STEP = kRleThreads * kRleItems;
    if (PASS == 1) {
        int m = 0;
        for (int i0 = 0; i0 < cnt; i0 += STEP) {
            const int ib = i0 + threadIdx.x * kRleItems;
            // the thread's items and their two neighbours
            int v[kRleItems + 2];
#pragma unroll
            for (int j = 0; j < kRleItems + 2; ++j) {
                const int i = ib + j - 1;
                v[j] = (i >= 0 && i < cnt) ? t[i] : -1 - j;   // (out of range: differs from every real neighbour)
            }
            int keep[kRleItems], mine = 0;
#pragma unroll
            for (int j = 0; j < kRleItems; ++j) {
                const int i = ib + j;
                keep[j] = i < cnt && (i == 0 || v[j] != v[j + 1]) && (i + 1 >= cnt || v[j + 2] != v[j + 1]);
                mine += keep[j];
            }
            int tot;
            int pos = m + block_exclusive_scan(mine, s_warp, tot);
#pragma unroll
            for (int j = 0; j < kRleItems; ++j)
                if (keep[j]) tp[pos++] = v[j + 1];
            m += tot;
        }
        __syncthreads();
        // counts 0 .. m-1 from the transitions, and the closing run when the last one ends early
        const long long last = m > 0 ? (long long)tp[m - 1] : 0ll;
        const int nc = m + ((last < total_pixels || m == 0) ? 1 : 0);
        int len = 0;
        for (int k0 = threadIdx.x * kRleItems; k0 < nc; k0 += STEP) {
            // counts k0-2 .. k0+3 from the transitions k0-3 .. k0+3
            long long tv[kRleItems + 3];
#pragma unroll
            for (int j = 0; j < kRleItems + 3; ++j) {
                const int i = k0 + j - 3;
                tv[j] = i < 0 ? 0ll : (i < m ? (long long)tp[i] : total_pixels);
            }
#pragma unroll
            for (int j = 0; j < kRleItems; ++j) {
                const int k = k0 + j;
                if (k < nc) {
                    const long long c = tv[j + 3] - tv[j + 2];
                    len += rle_code_len(k > 2 ? c - (tv[j + 1] - tv[j]) : c);
                }
            }
        }
        int tot;
        block_exclusive_scan(len, s_warp, tot);
        if (threadIdx.x == 0) {
            kept[n] = m;
            str_len[n] = tot;
        }
    } else {
        const int m = kept[n];
        const long long last = m > 0 ? (long long)tp[m - 1] : 0ll;
        const int nc = m + ((last < total_pixels || m == 0) ? 1 : 0);
        char* dst = out + str_offsets[n];
        int base = 0;
        for (int kb = 0; kb < nc; kb += STEP) {
            const int k0 = kb + threadIdx.x * kRleItems;
            long long tv[kRleItems + 3];
#pragma unroll
            for (int j = 0; j < kRleItems + 3; ++j) {
                const int i = k0 + j - 3;
                tv[j] = i < 0 ? 0ll : (i < m ? (long long)tp[i] : total_pixels);
            }
            long long x[kRleItems];
            int len[kRleItems], mine = 0;
#pragma unroll
            for (int j = 0; j < kRleItems; ++j) {
                const int k = k0 + j;
                x[j] = 0;
                len[j] = 0;
                if (k < nc) {
                    const long long c = tv[j + 3] - tv[j + 2];
                    x[j] = k > 2 ? c - (tv[j + 1] - tv[j]) : c;
                    len[j] = rle_code_len(x[j]);
                }
                mine += len[j];
            }
            int tot;
            int pos = base + block_exclusive_scan(mine, s_warp, tot);
#pragma unroll
            for (int j = 0; j < kRleItems; ++j) {
                if (len[j]) rle_code_write(x[j], dst + pos);
                pos += len[j];
            }
            base += tot;
        }
    }
}

// exclusive scan of N string lengths -> str_offsets[N+1] (one CTA; N is a few hundred)
__global__ void __launch_bounds__(kRleThreads) rle_offsets_kernel(const int32_t* __restrict__ str_len, int N,
                                                                  int64_t* __restrict__ str_offsets) {
    __shared__ int s_warp[kRleThreads / 32];
    long long base = 0;
    for (int i0 = 0; i0 < N; i0 += kRleThreads) {
        const int i = i0 + threadIdx.x;
        const int v = i < N ? str_len[i] : 0;
        int tot;
        const int pos = block_exclusive_scan(v, s_warp, tot);
        if (i < N) str_offsets[i] = base + pos;
        base += tot;
    }
    if (threadIdx.x == 0) str_offsets[N] = base;
}

// exclusive scan of the N per-instance transition counts -> inst_offsets[N+1]; header[0] = status (bit 0: the
// total exceeds `capacity`, the later launches of the same call then write nothing; bits 8..: column blocks whose
// recorded slots overflowed, i.e. that are evaluated a second time), header[1] = total
__global__ void __launch_bounds__(kRleThreads) rle_totals_scan_kernel(const int32_t* __restrict__ totals, int N,
                                                                      long long capacity, int64_t* __restrict__ inst_offsets,
                                                                      int32_t* __restrict__ status, int64_t* __restrict__ header,
                                                                      const int32_t* __restrict__ inst_over, int n_flags) {
    __shared__ int s_warp[kRleThreads / 32];
    long long base = 0;
    for (int i0 = 0; i0 < N; i0 += kRleThreads) {
        const int i = i0 + threadIdx.x;
        const int v = i < N ? totals[i] : 0;
        int tot;
        const int pos = block_exclusive_scan(v, s_warp, tot);
        if (i < N) inst_offsets[i] = base + pos;
        base += tot;
    }
    int over_blocks = 0;
    if (inst_over) {
        int mine = 0;
        for (int i = threadIdx.x; i < n_flags; i += kRleThreads) mine += inst_over[i] != 0;
        block_exclusive_scan(mine, s_warp, over_blocks);
    }
    if (threadIdx.x == 0) {
        inst_offsets[N] = base;
        const int over = base > capacity ? 1 : 0;
        *status = over;
        header[0] = (long long)over | ((long long)over_blocks << 8);
        header[1] = base;
    }
}

// an overflowing call still hands the host a well-formed (all-zero) offset table
__global__ void __launch_bounds__(kRleThreads) rle_clear_offsets_kernel(int64_t* __restrict__ str_offsets, int N,
                                                                        const int32_t* __restrict__ status) {
    if (!*status) return;
    for (int i = threadIdx.x; i <= N; i += kRleThreads) str_offsets[i] = 0;
}

}  // namespace dm

extern "C" int dm_rle_strings(const int32_t* transitions, const int64_t* inst_offsets, int N, int64_t total_pixels,
                              int32_t* compact, int32_t* kept, int32_t* str_len, int64_t* str_offsets, char* out,
                              dm_stream_t stream) {
    if (N < 0 || total_pixels < 0) return DM_EINVAL;
    if (N == 0) return DM_OK;
    if (!transitions || !inst_offsets || !compact || !kept || !str_len || !str_offsets || !out) return DM_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    dm::rle_string_kernel<1><<<N, dm::kRleThreads, 0, st>>>(transitions, inst_offsets, (long long)total_pixels, compact,
                                                           kept, str_len, nullptr, nullptr, nullptr);
    DM_LAUNCH_CHECK("dm_rle_strings/compact");
    dm::rle_offsets_kernel<<<1, dm::kRleThreads, 0, st>>>(str_len, N, str_offsets);
    DM_LAUNCH_CHECK("dm_rle_strings/scan");
    dm::rle_string_kernel<2><<<N, dm::kRleThreads, 0, st>>>(transitions, inst_offsets, (long long)total_pixels, compact,
                                                           kept, str_len, str_offsets, out, nullptr);
    DM_LAUNCH_CHECK("dm_rle_strings/write");
    return DM_OK;
}

// Whole paste -> RLE-string pipeline in ONE call with no host round trip in the middle: count, scan
// the counts on the device, write the transitions, build the strings.  The caller sizes the buffers
// from a transition CAPACITY instead of the exact total; when the masks need more, the call writes
// status = 1 into header[0] and nothing else (the caller repeats with header[1] transitions).
//   workspace  device scratch of dm_paste_rle_strings_workspace(N, x_hi - x_lo, capacity) bytes, 16-byte aligned
//   header     device int64 [2 + N + 1]: status, total transitions, string offsets (N + 1)
//   out        device chars, at least 6 * capacity + 8 * N + 8 bytes
namespace {
// layout of the caller's workspace (all offsets 16-byte aligned)
struct RleWorkspace {
    int64_t col_counts, totals, kept, str_len, status, inst_over, inst_offsets, slots, trans, compact, bytes;
};
// rows per segment / segments of a region of rh rows: 128-row segments, at most 16 of them
constexpr int kRleSegRows = 128, kRleMaxSeg = 16;
void rle_segments(int rh, int& nseg, int& seg_rows) {
    seg_rows = kRleSegRows;
    if ((rh + seg_rows - 1) / seg_rows > kRleMaxSeg) seg_rows = (rh + kRleMaxSeg - 1) / kRleMaxSeg;
    nseg = rh > 0 ? (rh + seg_rows - 1) / seg_rows : 1;
}
// The workspace is sized for the segmented form without recorded slots AND for the one-segment form with them
// (record_slots picks one per call): the slots of a segmented call would be nseg times as many.
RleWorkspace rle_workspace(int N, int rw, int rh_max, int64_t capacity) {
    const int64_t n = N, w = rw > 0 ? rw : 1, cap = capacity > 0 ? capacity : 1;
    int nseg, seg_rows;
    rle_segments(rh_max, nseg, seg_rows);
    auto up = [](int64_t v) { return (v + 15) & ~15ll; };
    RleWorkspace ws;
    int64_t o = 0;
    // col_counts | totals | kept | str_len | status (4 words) | inst_over: one memset clears them
    ws.col_counts = o; o = up(o + 4 * n * nseg * w);
    ws.totals = o; o += 4 * n;
    ws.kept = o; o += 4 * n;
    ws.str_len = o; o += 4 * n;
    ws.status = o; o += 16;
    ws.inst_over = o; o = up(o + 4 * n * nseg * ((w + dm::kRleThreads - 1) / dm::kRleThreads));
    ws.inst_offsets = o; o = up(o + 8 * (n + 1));
    ws.slots = o; o = up(o + 4 * n * w * dm::kRleSlots);
    ws.trans = o; o = up(o + 4 * cap);
    ws.compact = o; o = up(o + 4 * cap);
    ws.bytes = o;
    return ws;
}
}  // namespace

extern "C" int64_t dm_paste_rle_strings_workspace(int N, int rw, int rh, int64_t capacity) {
    if (N < 0 || rw < 0 || rh < 0 || capacity < 0) return -1;
    return rle_workspace(N, rw, rh, capacity).bytes;
}

extern "C" int dm_paste_rle_strings(const float* masks, int64_t mask_stride_n, int64_t mask_stride_c,
                                    const int64_t* labels, int N, int S_h, int S_w, int apply_sigmoid,
                                    const float* boxes, int img_h, int img_w, int x_lo, int y_lo, int x_hi,
                                    int y_hi, float thr, int record_slots, void* workspace, int64_t capacity,
                                    int64_t* header, char* out, dm_stream_t stream) {
    if (N < 0 || capacity < 0) return DM_EINVAL;
    if (N == 0) return DM_OK;
    if (!workspace || !header || !out || (reinterpret_cast<uintptr_t>(workspace) & 15u)) return DM_EINVAL;
    if (x_hi < x_lo || y_hi < y_lo) return DM_EINVAL;
    const RleWorkspace ws = rle_workspace(N, x_hi - x_lo, y_hi - y_lo, capacity);
    int nseg = 1, seg_rows = y_hi - y_lo;
    if (!record_slots) rle_segments(y_hi - y_lo, nseg, seg_rows);   // (recorded slots: one segment, as sized)
    char* base = static_cast<char*>(workspace);
    int32_t* col_counts = reinterpret_cast<int32_t*>(base + ws.col_counts);
    int32_t* totals = reinterpret_cast<int32_t*>(base + ws.totals);
    int32_t* kept = reinterpret_cast<int32_t*>(base + ws.kept);
    int32_t* str_len = reinterpret_cast<int32_t*>(base + ws.str_len);
    int32_t* status = reinterpret_cast<int32_t*>(base + ws.status);
    int32_t* inst_over = reinterpret_cast<int32_t*>(base + ws.inst_over);
    int64_t* inst_offsets = reinterpret_cast<int64_t*>(base + ws.inst_offsets);
    int32_t* slots = reinterpret_cast<int32_t*>(base + ws.slots);
    int32_t* trans = reinterpret_cast<int32_t*>(base + ws.trans);
    int32_t* compact = reinterpret_cast<int32_t*>(base + ws.compact);
    cudaStream_t st = (cudaStream_t)stream;
    DM_CUDA_CHECK(cudaMemsetAsync(col_counts, 0, (size_t)(ws.inst_offsets - ws.col_counts), st), "dm_paste_rle_strings/memset");
    if (!record_slots) slots = inst_over = nullptr;   // plain two-pass evaluation (noisy masks: most columns would overflow)
    auto pass = [&](int no) {
        return paste_rle_impl(masks, mask_stride_n, mask_stride_c, labels, N, S_h, S_w, apply_sigmoid, boxes, img_h, img_w,
                              x_lo, y_lo, x_hi, y_hi, thr, no, col_counts, totals, inst_offsets, trans, status, slots,
                              inst_over, nseg, seg_rows, stream);
    };
    const int n_flags = (int)((ws.inst_offsets - ws.inst_over) / 4);   // (zeroed beyond what a call uses)
    // 1: count, and record up to kRleSlots transitions per column
    int rc = pass(1);
    if (rc != DM_OK) return rc;
    dm::rle_totals_scan_kernel<<<1, dm::kRleThreads, 0, st>>>(totals, N, (long long)capacity, inst_offsets, status, header,
                                                              inst_over, n_flags);
    DM_LAUNCH_CHECK("dm_paste_rle_strings/scan");
    // 3: column blocks whose columns all fitted their slots are copied into place; 2: the others are evaluated again
    if (record_slots) {
        rc = pass(3);
        if (rc != DM_OK) return rc;
    }
    rc = pass(2);
    if (rc != DM_OK) return rc;
    const long long total_pixels = (long long)(y_hi - y_lo) * (x_hi - x_lo);
    int64_t* str_offsets = header + 2;
    dm::rle_string_kernel<1><<<N, dm::kRleThreads, 0, st>>>(trans, inst_offsets, total_pixels, compact, kept, str_len,
                                                           nullptr, nullptr, status);
    DM_LAUNCH_CHECK("dm_paste_rle_strings/compact");
    dm::rle_offsets_kernel<<<1, dm::kRleThreads, 0, st>>>(str_len, N, str_offsets);
    DM_LAUNCH_CHECK("dm_paste_rle_strings/offsets");
    dm::rle_clear_offsets_kernel<<<1, dm::kRleThreads, 0, st>>>(str_offsets, N, status);
    DM_LAUNCH_CHECK("dm_paste_rle_strings/clear");
    dm::rle_string_kernel<2><<<N, dm::kRleThreads, 0, st>>>(trans, inst_offsets, total_pixels, compact, kept, str_len,
                                                           str_offsets, out, status);
    DM_LAUNCH_CHECK("dm_paste_rle_strings/write");
    return DM_OK;
}

// Host side: transitions of ONE instance -> pycocotools' compressed counts string.
// counts = run lengths alternating from a run of zeros; the string is rleToString's base-32 varint
// code with each count beyond the second stored as a difference from the count two before.
// Returns the string length (no terminator), or -1 when `cap` is too small.
extern "C" int64_t dm_rle_compress_host(const int32_t* transitions, int64_t n, int64_t total_pixels, char* out,
                                        int64_t cap) {
    int64_t len = 0, k = 0;          // k = counts emitted so far
    int64_t c1 = 0, c2 = 0;          // counts one and two before the current one
    int64_t last = 0;                // flat index where the current run starts
    auto emit = [&](int64_t cnt) -> bool {
        int64_t x = cnt;
        if (k > 2) x -= c2;
        bool more = true;
        while (more) {
            char c = (char)(x & 0x1f);
            x >>= 5;
            more = (c & 0x10) ? x != -1 : x != 0;
            if (more) c |= 0x20;
            c += 48;
            if (len >= cap) return false;
            out[len++] = c;
        }
        c2 = c1;
        c1 = cnt;
        ++k;
        return true;
    };
    int64_t i = 0;
    while (i < n) {
        // two coinciding transitions (a run continuing across a column boundary) cancel
        if (i + 1 < n && transitions[i] == transitions[i + 1]) { i += 2; continue; }
        const int64_t t = transitions[i++];
        if (!emit(t - last)) return -1;
        last = t;
    }
    if (last < total_pixels || k == 0) {
        if (!emit(total_pixels - last)) return -1;
    }
    return len;
}

// Host side, batch form: the transitions of N instances (instance n owns
// transitions[offsets[n] .. offsets[n+1])) -> N compressed strings written back to back into
// `out`; str_offsets[n] .. str_offsets[n+1] delimits string n.  Returns the total length or -1 when
// `cap` is too small (6 bytes per transition + 8 per instance always suffice).
extern "C" int64_t dm_rle_compress_batch_host(const int32_t* transitions, const int64_t* offsets, int64_t N,
                                              int64_t total_pixels, char* out, int64_t cap,
                                              int64_t* str_offsets) {
    int64_t pos = 0;
    for (int64_t n = 0; n < N; ++n) {
        str_offsets[n] = pos;
        const int64_t len = dm_rle_compress_host(transitions + offsets[n], offsets[n + 1] - offsets[n],
                                                 total_pixels, out + pos, cap - pos);
        if (len < 0) return -1;
        pos += len;
    }
    str_offsets[N] = pos;
    return pos;
}
