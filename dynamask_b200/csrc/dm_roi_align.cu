// Stage 2: aligned avg-pool multi-level RoIAlign, forward and backward, for sm_100a.
//
// Replaces mmcv._ext.roi_align_forward / roi_align_backward as reached from
// mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:52-54 together with the per-level
// select / align / scatter loop of SingleRoIExtractor.forward
// (mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:53-81).
//
// Formulation.  For one RoI and one channel, bilinear sampling followed by the sampling-grid
// average is a separable linear map
//        out[ph][pw] = sum_y sum_x  Ay[ph][y] * Ax[pw][x] * feat[y][x]
// where Ax / Ay are banded: bin pw touches the pixels xs[pw] .. xs[pw]+JX-1 only.  The kernel
//   1. folds the g_w (g_h) sample weights of every bin into the banded tables Ax, Ay and stages
//      them in shared memory once per work unit (they are shared by all channels),
//   2. stages the RoI's feature patch (rows Y0..Y1, columns X0..X1) for a group of channels,
//   3. lets every WARP own whole channels: the warp stages its channel's patch in a private slice
//      of shared memory and its lanes are that channel's column strips (VEC adjacent pooled
//      columns each).  A lane walks its strip top to bottom; the X-interpolated patch rows it
//      currently needs live in a register window (V[y][pw] = sum_j Ax[pw][j] * patch[y][xs[pw]+j],
//      recomputed from shared memory only when the window slides), each pooled row is
//      out = sum_j Ay[ph][j] * window[j] and leaves as one 16-byte streaming store.
// HBM sees each patch element once and each output element once.  Backward is the transpose: the
// same strip walk reads grad_out once, straight into registers (two batches of rows in flight),
// accumulates the band rows in the register window, and when a band row is complete the warp
// drops it into a 512-byte private row buffer, gathers the patch-gradient row from it with the
// transposed X table and issues one global reduction (RED.ADD.F32) per touched feature pixel --
// i.e. the atomics are aggregated per warp before they reach L2.
// After the per-unit table build no CTA-wide barrier is needed: warps run independently, so one
// warp's load latency is hidden by the others' stores.
// Geometries whose bands are wider than 8 pixels (or pooled rows wider than 32 strips) fall back
// to a CTA-wide two-pass shared-memory form, and those that do not fit shared memory at all to
// direct sample-by-sample evaluation.
//
// Scheduling.  All resolution buckets run in ONE persistent launch: grid = SMs x resident CTAs,
// work units = (RoI, channel slab) enumerated bucket by bucket from the largest output size to the
// smallest, static round-robin over CTAs.  Bucket membership comes from device memory
// (dm_assign's perm / seg_offsets), so no host synchronisation is needed.
//
// Summation order differs from the sample-by-sample order of the reference; the contract for
// feature RoIAlign is 1e-5 relative (forward) / 1e-4 (backward), see tests/test_roi_align_gpu.py.
#include <atomic>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include <cuda.h>   // CUtensorMap and the cuTensorMapEncodeTiled prototype only: the entry point is
                    // resolved through cudaGetDriverEntryPoint, the library does not link libcuda

#include "dm_common.cuh"

namespace dm {

// Paths off the steady state (table build, fallbacks).  Measured both ways on C2: keeping them out
// of line (-DDM_COLD=__noinline__) costs the forward 2 % (7.74 vs 7.57 ms) and leaves the backward
// unchanged, so they stay inlined.
#ifndef DM_COLD
#define DM_COLD
#endif

// CTA size differs per direction (register budget): device code reads it from blockDim
#ifndef DM_BWD_THREADS
#define DM_BWD_THREADS 256
#endif
#ifndef DM_BWD_REGS
#define DM_BWD_REGS 128
#endif
#ifndef DM_BWD_DYNAMIC
#define DM_BWD_DYNAMIC 1   // dynamic unit scheduling in the backward: 0 never, 1 except single-bucket launches of tiny
                           // pooled sizes, 2 always.  Round 1 measured it level with static ownership (8.32 vs 8.25 ms on
                           // C2); since the scheduling thread moved off the table-building warps and fetches the next
                           // unit's RoI with its ticket it wins at every size but 7x7 (kernel-only, 2048 RoIs per size:
                           // 14x14 609 -> 547 us, 28x28 849 -> 798, 56x56 2019 -> 1782, 112x112 4702 -> 4510,
                           // 56x56 single-level x 256 RoIs 2169 -> 1415; 7x7 x 1024 RoIs 290 -> 327: stays static)
#endif
constexpr int kBwdThreads = DM_BWD_THREADS;   // backward: 2 CTAs/SM x 8 warps at 128 registers
#ifndef DM_FWD_THREADS
#define DM_FWD_THREADS 256
#endif
#ifndef DM_FWD_HOIST
#define DM_FWD_HOIST 1   // X strip constants (first column, weights) of a lane in registers instead of re-read per patch row
#endif
#ifndef DM_FWD_REGS
#define DM_FWD_REGS (DM_FWD_HOIST ? 128 : 96)
#endif
constexpr int kFwdThreads = DM_FWD_THREADS;   // forward: 2 CTAs/SM x 8 warps at 128 registers
#define RA_THREADS ((int)blockDim.x)
#define RA_WARPS ((int)(blockDim.x >> 5))

struct LevelDesc {
    float* ptr;  // const for forward, accumulated into for backward
    int N, C, H, W;
    long long sN, sC, sH, sW;
    float scale;
    int cw;  // floats per asynchronous copy when staging patch rows: 4, 2 or 1 (alignment of rows)
};

struct BucketDesc {
    float* ptr;  // written by forward, read by backward
    int ph, pw;
    long long sN, sC, sH, sW;
    int cg;     // channels per work unit
    int nslab;  // ceil(C / cg)
    int vec;    // vector width of the pooled-row accesses (4, 2 or 1)
    int bvec;   // same for the backward walk, which moves grad_out with bulk copies when > 1:
                // dense rows, 16-byte aligned planes
};

// Patch loads by TMA (cp.async.bulk.tensor.3d): one tensor map per (pyramid level, channels per
// box, box-width class).  A box is 4 << set channels x BH patch rows x BW columns of one image: BW the
// smallest class that covers the patch from its first column rounded down to a 16-byte boundary,
// the channel count what one warp pass consumes (a lane always produces four pooled values per
// patch row: VEC columns of 4 / VEC channels), BH such that a box is at most 4 KB.
constexpr int kTmaSets = 3;         // boxes of 4, 8, 16 channels
constexpr int kTmaLevels = 5;       // pyramid levels that can have maps (kernel-parameter space)
constexpr int kNumBW = 10;
__host__ __device__ constexpr int tma_bw(int cls) {
    return cls == 0 ? 8 : cls == 1 ? 12 : cls == 2 ? 16 : cls == 3 ? 20 : cls == 4 ? 24 : cls == 5 ? 28
         : cls == 6 ? 32 : cls == 7 ? 40 : cls == 8 ? 48 : 64;
}
__host__ __device__ constexpr int tma_bh(int set, int cls) { return (cls <= 6 ? 8 : 4) >> set; }
constexpr int kTmaMaxBW = 64;
constexpr int kTmaMaxSlots = 4;     // chunk slots per warp (ring depth)
struct alignas(64) TmaMaps {
    CUtensorMap m[kTmaLevels][kTmaSets][kNumBW];
};

struct RaParams {
    LevelDesc lv[DM_MAX_LEVELS];
    BucketDesc bk[DM_MAX_BUCKETS];
    int order[DM_MAX_BUCKETS];
    int L, nb, K, C;
    const float* rois;
    const int32_t* lvl;
    const int32_t* perm;
    const int32_t* seg;
    int sampling_ratio, aligned;
    int mode;  // 0 RoIAlign, 1 SimpleRoIAlign (one zero-padded grid_sample point per bin)
    int interleave;  // walk the buckets at the same fractional pace (1) or one after the other (0)
    int smem_floats;
    // dynamic scheduling: one ticket counter per bucket (zeroed before the launch), NULL = static
    // round-robin ownership
    unsigned* tickets;
    float bias;      // how much later (per bucket rank) the smaller buckets are paced
    // TMA patch loads: bit c of tma_mask[l][s] = level l has a tensor map for channel set s, width class c;
    // tma_rowmajor = 1: boxes land as [row][channel][BW] (tensor dims W, C*N, H), 0: [channel][row][BW]
    unsigned tma_mask[kTmaLevels][kTmaSets];
    int tma_rowmajor;
    int tma_slots;   // ring depth wanted (2 .. kTmaMaxSlots)
    int diag;        // measurement only (DM_RA_DIAG)
    int l2_prefetch; // forward: L2 prefetch of a unit's patch while its tables are built
    int bwd_x;       // backward: X-first walk for small pooled sizes
    int bwd_groups;  // ... with at most this many lane groups (1, 2, 4)
};

// exact n / d whenever n * d < 2^32 (indices here are far below that)
struct FastDiv {
    unsigned m, d;
    __device__ __forceinline__ void init(unsigned dd) {
        d = dd;
        m = dd > 1 ? 0xFFFFFFFFu / dd + 1u : 0u;
    }
    __device__ __forceinline__ void set(unsigned dd, unsigned mm) { d = dd; m = mm; }
    __device__ __forceinline__ unsigned div(unsigned n) const { return d > 1 ? __umulhi(n, m) : n; }
};

// fraction of thread slots doing work when n items are dealt round-robin to the CTA
__device__ __forceinline__ float cta_util(int n) {
    return (float)n / (float)(((n + RA_THREADS - 1) / RA_THREADS) * RA_THREADS);
}

__device__ __forceinline__ int pow2_shift_ge(int n) {  // smallest s with (1 << s) >= n
    return n <= 1 ? 0 : 32 - __clz(n - 1);
}

template <int VEC>
__device__ __forceinline__ void ld_vec(const float* p, float (&a)[VEC]) {
    if (VEC == 4) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        a[0] = v.x; a[1 % VEC] = v.y; a[2 % VEC] = v.z; a[3 % VEC] = v.w;
    } else if (VEC == 2) {
        const float2 v = *reinterpret_cast<const float2*>(p);
        a[0] = v.x; a[1 % VEC] = v.y;
    } else {
        a[0] = *p;
    }
}

template <int VEC>
__device__ __forceinline__ void ld_vec_i(const int* p, int (&a)[VEC]) {
    if (VEC == 4) {
        const int4 v = *reinterpret_cast<const int4*>(p);
        a[0] = v.x; a[1 % VEC] = v.y; a[2 % VEC] = v.z; a[3 % VEC] = v.w;
    } else if (VEC == 2) {
        const int2 v = *reinterpret_cast<const int2*>(p);
        a[0] = v.x; a[1 % VEC] = v.y;
    } else {
        a[0] = *p;
    }
}

template <int VEC>
__device__ __forceinline__ void ldg_stream_vec(const float* p, float (&a)[VEC]) {
    if (VEC == 4) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
        a[0] = v.x; a[1 % VEC] = v.y; a[2 % VEC] = v.z; a[3 % VEC] = v.w;
    } else if (VEC == 2) {
        const float2 v = __ldcs(reinterpret_cast<const float2*>(p));
        a[0] = v.x; a[1 % VEC] = v.y;
    } else {
        a[0] = __ldcs(p);
    }
}

template <int VEC>
__device__ __forceinline__ void st_stream_vec(float* p, const float (&a)[VEC]) {
    if (VEC == 4) st_stream(reinterpret_cast<float4*>(p), make_float4(a[0], a[1 % VEC], a[2 % VEC], a[3 % VEC]));
    else if (VEC == 2) st_stream(reinterpret_cast<float2*>(p), make_float2(a[0], a[1 % VEC]));
    else st_stream(p, a[0]);
}

// Pooled-row store of the TMA forward path: streaming (evict-first) like every other pooled store.  Plain
// write-back stores were measured and make no difference (14x14: 351.7 vs 349.2 us, 7x7: 201.8 vs 199.3).
#ifndef DM_TMA_ROTATE
#define DM_TMA_ROTATE 0   // forward TMA walk: rotating register window (unrolled JW deep) instead of shifting it.
                          // Measured SLOWER (kernel-only, ncu): 14x14 349 -> 370 us, 28x28 557 -> 624-642, 7x7 x 1024 RoIs 199 -> 241:
                          // the walk is fetch-bound (no_inst is its top stall) and the unrolled body is four times the code
#endif
#ifndef DM_TMA_STREAM_STORES
#define DM_TMA_STREAM_STORES 1
#endif
template <int VEC>
__device__ __forceinline__ void st_out_vec(float* p, const float (&a)[VEC]) {
#if DM_TMA_STREAM_STORES
    st_stream_vec<VEC>(p, a);
#else
    if (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1 % VEC], a[2 % VEC], a[3 % VEC]);
    else if (VEC == 2) *reinterpret_cast<float2*>(p) = make_float2(a[0], a[1 % VEC]);
    else *p = a[0];
#endif
}

// Packed FP32 pairs (FFMA2 / FMUL2 on sm_100a): two FMAs per issued instruction.  The hot loops are
// issue-bound, not FP-bound, so halving the FP instruction count is worth the inline PTX.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(r)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(r)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&r);
}
// acc = w * x  /  acc += w * x  over VEC lanes' worth of columns, w broadcast as the pair (w, w)
template <int VEC>
__device__ __forceinline__ void vmul(float (&acc)[VEC], float2 ww, const float (&x)[VEC]) {
    if (VEC == 1) {
        acc[0] = ww.x * x[0];
    } else {
#pragma unroll
        for (int e = 0; e + 1 < VEC; e += 2) {
            const float2 r = fmul2(ww, make_float2(x[e], x[e + 1]));
            acc[e] = r.x;
            acc[e + 1] = r.y;
        }
    }
}
template <int VEC>
__device__ __forceinline__ void vfma(float (&acc)[VEC], float2 ww, const float (&x)[VEC]) {
    if (VEC == 1) {
        acc[0] += ww.x * x[0];
    } else {
#pragma unroll
        for (int e = 0; e + 1 < VEC; e += 2) {
            const float2 r = ffma2(ww, make_float2(x[e], x[e + 1]), make_float2(acc[e], acc[e + 1]));
            acc[e] = r.x;
            acc[e + 1] = r.y;
        }
    }
}

// Ampere-style asynchronous global->shared copies (LDGSTS): used as a register-free prefetch ring.
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
    else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cp_async_sa(unsigned sa, const void* gsrc) {
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
    else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}
// Unpredicated form: `live` false copies nothing and zero-fills the destination (src-size 0), so the
// instruction, its address arithmetic and its memory descriptor stay outside any per-lane branch.
// The source is `base + 4 * off` with a 32-bit float offset: the address is formed by one wide
// multiply-add inside the asm block, so no 64-bit pointer is carried (and shuffled) around the loop.
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill_sa(unsigned sa, const float* base, int off, bool live) {
    const int n = live ? BYTES : 0;
    if (BYTES == 16)
        // (.cg: the L1-allocating .ca form of the same copy was measured 5 % slower in the backward)
        asm volatile("{\n.reg .u64 a;\nmad.wide.s32 a, %3, 4, %1;\ncp.async.cg.shared.global [%0], [a], 16, %2;\n}"
                     ::"r"(sa), "l"(base), "r"(n), "r"(off) : "memory");
    else if (BYTES == 8)
        asm volatile("{\n.reg .u64 a;\nmad.wide.s32 a, %3, 4, %1;\ncp.async.ca.shared.global [%0], [a], 8, %2;\n}"
                     ::"r"(sa), "l"(base), "r"(n), "r"(off) : "memory");
    else
        asm volatile("{\n.reg .u64 a;\nmad.wide.s32 a, %3, 4, %1;\ncp.async.ca.shared.global [%0], [a], 4, %2;\n}"
                     ::"r"(sa), "l"(base), "r"(n), "r"(off) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Bulk asynchronous copies (TMA, 1-D): one elected lane moves whole runs of pooled rows
// global -> shared and the data's arrival is signalled on an mbarrier (SASS: UBLKCP + SYNCS).
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// Tiled TMA load of one 3-D box global -> shared (SASS: UTMALDG.3D), completion counted in bytes
// on an mbarrier.  `map` points at a CUtensorMap in kernel-parameter space (__grid_constant__).
__device__ __forceinline__ void tma_load_3d(unsigned dst, const void* map, int c0, int c1, int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
// L2 prefetch of one box (SASS: UTMAPF): no shared-memory destination, no completion to wait for
__device__ __forceinline__ void tma_prefetch_3d(const void* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Banded weight tables of one RoI, staged in shared memory.
// ---------------------------------------------------------------------------------------------
enum { ST_JX = 0, ST_X0, ST_X1, ST_PA, ST_PB, ST_JY, ST_Y0, ST_Y1, ST_QA, ST_QB, ST_MR, ST_TW, ST_NEED, ST_N };

struct Tables {
    int* xs;    // [pw]  first feature column touched by bin pw
    int* ys;    // [ph]
    float* wx;  // [JX][pw] folded weights (already divided by g_w)
    float* wy;  // [JY][ph]
    int JX, JY, X0, X1, Y0, Y1;
    int JXa, JYa;  // rows allocated for wx / wy (>= JX / JY, zero padded up to the window class)
    unsigned mR;   // FastDiv magic of R = Y1 - Y0 + 1
    float* ytab;   // packed per-pooled-row records {ys - Y0, wy[0..JYa)} when JYa is a window class
    int ystride;   // floats per record (2 * JYa: the weights as broadcast pairs), 0 when there is no packed table
    int* rcnt;     // [R] pooled rows whose band starts at patch row r (only with a packed table)
    int floats;    // shared-memory floats consumed (multiple of 4)
};

// register-window class for a band width: 2, 4, 8, or 0 when wider than 8
__device__ __forceinline__ int window_class(int j) { return j <= 2 ? 2 : (j <= 4 ? 4 : (j <= 8 ? 8 : 0)); }

// One warp scans one axis: lane p owns bin p (and p + 32, ...).  Writes the first pixel of every
// bin's band to s_start (INT_MAX when the bin has no valid sample) and the axis statistics
// {widest band, first pixel, last pixel, first valid bin, last valid bin} to stat[0..5).
__device__ __forceinline__ void axis_scan_warp(const RoiGeom& g, int axis, int P, int size, int grid,
                                               int* s_start, int* stat) {
    const int lane = threadIdx.x & 31;
    int jmax = 0, pmin = INT_MAX, pmax = -1, bmin = INT_MAX, bmax = -1;
    for (int p = lane; p < P; p += 32) {
        int first = INT_MAX, last = -1;
        for (int i = 0; i < grid; ++i) {
            int lo, hi;
            float l, h;
            if (geom_tap(g, axis, P, size, p, i, lo, hi, l, h)) {
                first = min(first, lo);
                last = max(last, hi);
            }
        }
        s_start[p] = first;
        if (last >= 0) {
            jmax = max(jmax, last - first + 1);
            pmin = min(pmin, first);
            pmax = max(pmax, last);
            bmin = min(bmin, p);
            bmax = max(bmax, p);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        jmax = max(jmax, __shfl_xor_sync(0xffffffffu, jmax, o));
        pmin = min(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
        pmax = max(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
        bmin = min(bmin, __shfl_xor_sync(0xffffffffu, bmin, o));
        bmax = max(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
    }
    if (lane == 0) {
        stat[0] = jmax; stat[1] = pmin; stat[2] = pmax; stat[3] = bmin; stat[4] = bmax;
    }
}

// One warp fills one axis' folded weights w[Ja][P]: the lane that owns bin p fixes up its start
// (bins without a valid sample sit at the two ends; they get a start that keeps the starts
// monotone, their weights stay zero), clears its column and accumulates its samples -- no other
// thread touches column p, so no atomics and no barrier in between.
__device__ __forceinline__ void axis_fill_warp(const RoiGeom& g, int axis, int P, int size, int grid, int* s_start,
                                               float* w, int Ja, int first_pix, int bin_a, int start_last) {
    const int lane = threadIdx.x & 31;
    const float inv = 1.0f / (float)grid;
    for (int p = lane; p < P; p += 32) {
        int st = s_start[p];
        if (st == INT_MAX) {
            st = p < bin_a ? first_pix : start_last;
            s_start[p] = st;
        }
        for (int j = 0; j < Ja; ++j) w[j * P + p] = 0.0f;
        for (int i = 0; i < grid; ++i) {
            int lo, hi;
            float l, h;
            if (geom_tap(g, axis, P, size, p, i, lo, hi, l, h)) {
                w[(lo - st) * P + p] += h * inv;
                w[(hi - st) * P + p] += l * inv;
            }
        }
    }
}

// Returns false (uniformly) when the RoI has no valid sample at all -> output is all zeros.
// `fits` is set false when the tables alone exceed the shared-memory budget.
// Warp 0 builds the X axis, warp 1 the Y axis (incl. the packed Y records), the other warps only
// meet them at the two barriers.
__device__ DM_COLD bool build_tables(const RoiGeom& g, int Ph, int Pw, int H, int W, float* smem,
                             int smem_floats, int* stat, Tables& t, bool& fits) {
    fits = true;
    t.xs = reinterpret_cast<int*>(smem);
    t.ys = t.xs + Pw;
    if (g.gw <= 0 || g.gh <= 0) return false;
    if (Pw + Ph > smem_floats) { fits = false; return true; }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy_warp = RA_WARPS > 1 ? 1 : 0;
    if (warp == 0) axis_scan_warp(g, 0, Pw, W, g.gw, t.xs, stat + ST_JX);
    if (warp == wy_warp) axis_scan_warp(g, 1, Ph, H, g.gh, t.ys, stat + ST_JY);
    __syncthreads();
    t.JX = stat[ST_JX]; t.X0 = stat[ST_X0]; t.X1 = stat[ST_X1];
    t.JY = stat[ST_JY]; t.Y0 = stat[ST_Y0]; t.Y1 = stat[ST_Y1];
    if (t.JX <= 0 || t.JY <= 0) return false;
    const int pa = stat[ST_PA], pb = stat[ST_PB], qa = stat[ST_QA], qb = stat[ST_QB];
    const int base = (Pw + Ph + 3) & ~3;
    // both tables padded to the common window class when both bands are <= 8 wide (forward walk),
    // otherwise each to its own class (the backward walk only needs the Y table)
    const int wc = window_class(max(t.JX, t.JY));
    t.JXa = wc ? wc : (window_class(t.JX) ? window_class(t.JX) : t.JX);
    t.JYa = wc ? wc : (window_class(t.JY) ? window_class(t.JY) : t.JY);
    const int wfloats = (t.JXa * Pw + t.JYa * Ph + 3) & ~3;
    t.ystride = (t.JYa == 2 || t.JYa == 4 || t.JYa == 8) ? 2 * t.JYa : 0;
    const int Rr = t.Y1 - t.Y0 + 1;
    const int rfloats = t.ystride ? ((Rr + 3) & ~3) : 0;
    t.floats = base + wfloats + t.ystride * Ph + rfloats;
    {
        const unsigned R = (unsigned)Rr;
        t.mR = R > 1 ? 0xFFFFFFFFu / R + 1u : 0u;
    }
    if (t.floats > smem_floats) { fits = false; return true; }
    t.wx = smem + base;
    t.wy = t.wx + t.JXa * Pw;
    t.ytab = smem + base + wfloats;
    t.rcnt = reinterpret_cast<int*>(t.ytab + t.ystride * Ph);
    // (the lanes owning bins pb / qb hold valid starts, which the fill never rewrites)
    const int xs_last = t.xs[pb], ys_last = t.ys[qb];
    if (warp == 0) axis_fill_warp(g, 0, Pw, W, g.gw, t.xs, t.wx, t.JXa, t.X0, pa, xs_last);
    if (warp == wy_warp) {
        for (int i = lane; i < rfloats; i += 32) t.rcnt[i] = 0;
        axis_fill_warp(g, 1, Ph, H, g.gh, t.ys, t.wy, t.JYa, t.Y0, qa, ys_last);
        __syncwarp();
        if (t.ystride) {
            // packed record of pooled row ph: its JYa row weights as broadcast pairs (written by the
            // lane that just produced them), and the count of pooled rows per first band row
            for (int ph = lane; ph < Ph; ph += 32) {
                float* rec = t.ytab + ph * t.ystride;
                for (int j = 0; j < t.JYa; ++j) {
                    const float v = t.wy[j * Ph + ph];
                    rec[2 * j] = v;
                    rec[2 * j + 1] = v;
                }
                atomicAdd(&t.rcnt[t.ys[ph] - t.Y0], 1);
            }
        }
    }
    __syncthreads();
    return true;
}

// one packed Y record: the JW row weights of a pooled row as broadcast pairs (w_j, w_j)
template <int JW>
__device__ __forceinline__ void load_yrec(const float* __restrict__ rec, float2 (&w)[JW]) {
    const float4* r = reinterpret_cast<const float4*>(rec);
#pragma unroll
    for (int j = 0; j < JW; j += 2) {
        const float4 a = r[j >> 1];
        w[j] = make_float2(a.x, a.y);
        w[j + 1] = make_float2(a.z, a.w);
    }
}

// ---------------------------------------------------------------------------------------------
// Unit bookkeeping
// ---------------------------------------------------------------------------------------------
struct Unit {
    int b;      // bucket
    int i;      // position inside the bucket (row of the bucket's output tensor)
    int slab;   // channel slab
    // the RoI's record and pyramid level when the scheduler already fetched them (dynamic scheduling:
    // the thread that takes the next ticket also loads that unit's RoI, so the two dependent global
    // loads perm -> rois / lvl are off the unit's critical path)
    int have;
    int lv;
    float r[5];
};

// the unit's RoI record (5 floats) and level: prefetched by the scheduler, else read here
__device__ __forceinline__ void unit_roi(const RaParams& p, const Unit& un, const int* s_seg, float (&rr)[5], int& lv) {
    if (un.have) {
#pragma unroll
        for (int e = 0; e < 5; ++e) rr[e] = un.r[e];
        lv = un.lv;
    } else {
        const int pos = s_seg[un.b] + un.i;
        const int k = p.perm ? p.perm[pos] : pos;
        const float* roi = p.rois + 5 * (size_t)k;
#pragma unroll
        for (int e = 0; e < 5; ++e) rr[e] = roi[e];
        lv = p.lvl ? p.lvl[k] : 0;
    }
}

__device__ __forceinline__ long long total_units(const RaParams& p, const int* s_seg) {
    long long tot = 0;
    for (int j = 0; j < p.nb; ++j) {
        const int b = p.order[j];
        tot += (long long)(s_seg[b + 1] - s_seg[b]) * p.bk[b].nslab;
    }
    return tot;
}

__device__ __forceinline__ Unit decode_unit(const RaParams& p, const int* s_seg, long long u) {
    Unit un;
    un.b = p.order[p.nb - 1];
    un.i = 0;
    un.slab = 0;
    un.have = 0;
    for (int j = 0; j < p.nb; ++j) {
        const int b = p.order[j];
        const long long n = (long long)(s_seg[b + 1] - s_seg[b]) * p.bk[b].nslab;
        if (u < n) {
            un.b = b;
            // slab-major inside a RoI so consecutive CTAs share one RoI's patch in L2
            un.i = (int)(u / p.bk[b].nslab);
            un.slab = (int)(u - (long long)un.i * p.bk[b].nslab);
            return un;
        }
        u -= n;
    }
    return un;
}

// ---------------------------------------------------------------------------------------------
// Fallbacks: zero fill, and direct (sample-by-sample) evaluation for geometries whose tables or
// tiles do not fit in shared memory (e.g. a whole 800x1344 map pooled to 14x14).
// ---------------------------------------------------------------------------------------------
__device__ DM_COLD void zero_unit(const BucketDesc& B, int i, int c0, int c1) {
    const int per_c = B.ph * B.pw;
    const int n = (c1 - c0) * per_c;
    for (int e = threadIdx.x; e < n; e += RA_THREADS) {
        const int c = e / per_c, r = e - c * per_c;
        const int ph = r / B.pw, pw = r - ph * B.pw;
        B.ptr[(long long)i * B.sN + (long long)(c0 + c) * B.sC + (long long)ph * B.sH + (long long)pw * B.sW] = 0.0f;
    }
}

template <bool BWD>
__device__ DM_COLD void direct_unit(const LevelDesc& Lv, const BucketDesc& B, const RoiGeom& g, int batch,
                            int i, int c0, int c1) {
    const int per_c = B.ph * B.pw;
    const int n = (c1 - c0) * per_c;
    const int cnt = g.gh * g.gw;
    const float inv_count = 1.0f / (float)(cnt > 1 ? cnt : 1);
    for (int e = threadIdx.x; e < n; e += RA_THREADS) {
        const int c = e / per_c, r = e - c * per_c;
        const int ph = r / B.pw, pw = r - ph * B.pw;
        float* o = B.ptr + (long long)i * B.sN + (long long)(c0 + c) * B.sC + (long long)ph * B.sH + (long long)pw * B.sW;
        float* f = Lv.ptr + (long long)batch * Lv.sN + (long long)(c0 + c) * Lv.sC;
        float acc = 0.0f;
        const float gv = BWD ? *o * inv_count : 0.0f;
        for (int iy = 0; iy < g.gh; ++iy) {
            int yl, yh;
            float ly, hy;
            if (!geom_tap(g, 1, B.ph, Lv.H, ph, iy, yl, yh, ly, hy)) continue;
            for (int ix = 0; ix < g.gw; ++ix) {
                int xl, xh;
                float lx, hx;
                if (!geom_tap(g, 0, B.pw, Lv.W, pw, ix, xl, xh, lx, hx)) continue;
                float* p1 = f + (long long)yl * Lv.sH + (long long)xl * Lv.sW;
                float* p2 = f + (long long)yl * Lv.sH + (long long)xh * Lv.sW;
                float* p3 = f + (long long)yh * Lv.sH + (long long)xl * Lv.sW;
                float* p4 = f + (long long)yh * Lv.sH + (long long)xh * Lv.sW;
                if (BWD) {
                    red_add(p1, gv * hy * hx);
                    red_add(p2, gv * hy * lx);
                    red_add(p3, gv * ly * hx);
                    red_add(p4, gv * ly * lx);
                } else {
                    acc += hy * hx * __ldg(p1) + hy * lx * __ldg(p2) + ly * hx * __ldg(p3) + ly * lx * __ldg(p4);
                }
            }
        }
        if (!BWD) *o = acc * inv_count;
    }
}

// SimpleRoIAlign on a patch too large for the streaming walk (a big RoI on a fine map: the points
// are sparse in the patch, so staging whole patch rows would mostly move pixels nobody samples).
// Every output is four taps; the two X taps / weights of a pooled column and the two Y taps of a
// pooled row come from the banded tables (bands are at most 2 wide in point mode).  A warp owns
// whole pooled rows of a channel, lanes run along the row: coalesced stores (forward) / loads
// (backward), gathers served by L1 / L2, one RED per tap in the backward (the points are at least
// a pixel apart here, so there is no contention to aggregate).
template <bool BWD>
__device__ __noinline__ void point_sparse_unit(const LevelDesc& Lv, const BucketDesc& B, const Tables& t, int batch,
                                  int i, int c0, int c1) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Pw = B.pw, Ph = B.ph;
    const int rows = (c1 - c0) * Ph;
    // not inlined: the tables live in shared memory (LDS, not generic LD)
    __builtin_assume(__isShared(t.xs));
    __builtin_assume(__isShared(t.ys));
    __builtin_assume(__isShared(t.wx));
    __builtin_assume(__isShared(t.wy));
    float* fbase = Lv.ptr + (long long)batch * Lv.sN + (long long)c0 * Lv.sC;
    float* obase = B.ptr + (long long)i * B.sN + (long long)c0 * B.sC;
    for (int row = warp; row < rows; row += RA_WARPS) {
        const int c = row / Ph, ph = row - c * Ph;
        const int y0 = t.ys[ph], y1 = min(y0 + 1, Lv.H - 1);
        const float wy0 = t.wy[ph], wy1 = t.JYa > 1 ? t.wy[Ph + ph] : 0.0f;
        float* f0 = fbase + (long long)c * Lv.sC + (long long)y0 * Lv.sH;
        float* f1 = fbase + (long long)c * Lv.sC + (long long)y1 * Lv.sH;
        float* o = obase + (long long)c * B.sC + (long long)ph * B.sH;
        for (int pw = lane; pw < Pw; pw += 32) {
            const int x0 = t.xs[pw], x1 = min(x0 + 1, Lv.W - 1);
            const float wx0 = t.wx[pw], wx1 = t.JXa > 1 ? t.wx[Pw + pw] : 0.0f;
            if (BWD) {
                const float g = __ldcs(o + (long long)pw * B.sW);
                const float a = g * wy0, b = g * wy1;
                if (a * wx0 != 0.0f) red_add(f0 + (long long)x0 * Lv.sW, a * wx0);
                if (a * wx1 != 0.0f) red_add(f0 + (long long)x1 * Lv.sW, a * wx1);
                if (b * wx0 != 0.0f) red_add(f1 + (long long)x0 * Lv.sW, b * wx0);
                if (b * wx1 != 0.0f) red_add(f1 + (long long)x1 * Lv.sW, b * wx1);
            } else {
                const float v = wy0 * (wx0 * __ldg(f0 + (long long)x0 * Lv.sW) + wx1 * __ldg(f0 + (long long)x1 * Lv.sW)) +
                                wy1 * (wx0 * __ldg(f1 + (long long)x0 * Lv.sW) + wx1 * __ldg(f1 + (long long)x1 * Lv.sW));
                __stcs(o + (long long)pw * B.sW, v);
            }
        }
    }
}

// Rows of the feature map needed by pooled rows [p0, p1): ys[p0] .. min(ys[p1-1]+JY-1, Y1)
__device__ __forceinline__ int band_rows(const Tables& t, int p0, int p1) {
    return min(t.ys[p1 - 1] + t.JY - 1, t.Y1) - t.ys[p0] + 1;
}

// ---------------------------------------------------------------------------------------------
// Patch staging: rows [Yt0, Yt0+R) x columns [X0, X0+fw) of `cs` channels -> shared memory
// [cs][R][fws] (fws >= fw; columns >= fw are zero so banded reads never need a clamp).
// One warp per (channel, row), PF rows in flight per warp.
// ---------------------------------------------------------------------------------------------
__device__ DM_COLD void stage_patch(const LevelDesc& Lv, float* patch, int batch, int c0, int cs, int Yt0,
                            int R, int X0, int fw, int fws) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* __restrict__ src = Lv.ptr + (long long)batch * Lv.sN + (long long)c0 * Lv.sC + (long long)Yt0 * Lv.sH + (long long)X0 * Lv.sW;
    const long long sC = Lv.sC, sH = Lv.sH, sW = Lv.sW;
    const int nrows = cs * R;
    FastDiv fdR;
    fdR.init(R);
    constexpr int PF = 8;
    for (int x = lane; x < fws; x += 32) {
        const float* __restrict__ sx = src + (long long)x * sW;
        const bool live = x < fw;
        for (int r0 = warp; r0 < nrows; r0 += PF * RA_WARPS) {
            float v[PF];
#pragma unroll
            for (int q = 0; q < PF; ++q) {
                const int row = r0 + q * RA_WARPS;
                v[q] = 0.0f;
                if (live && row < nrows) {
                    const int c = fdR.div(row), r = row - c * R;
                    v[q] = __ldg(sx + c * sC + r * sH);
                }
            }
#pragma unroll
            for (int q = 0; q < PF; ++q) {
                const int row = r0 + q * RA_WARPS;
                if (row < nrows) patch[row * fws + x] = v[q];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Forward, fast path: warp-autonomous strip walk.  Requires JX, JY <= JW (tables are zero padded
// to JW rows), Pw / VEC <= 32 strips per channel and a channel patch that fits the warp's slice.
// `cpw` channels are processed per warp pass (lanes = cpw x PwV strips).
// ---------------------------------------------------------------------------------------------
// The warp's patches arrive as one continuous stream of patch rows (all rows of its first channel
// batch, then of its next batch, ...) through a small ring of shared-memory row slots filled by
// cp.async: kFwdPF rows are always in flight, so neither the start of a channel nor the slide of
// the window ever waits for a cold load, and a warp needs ~1-3 KB of shared memory instead of a
// whole patch.
constexpr int kFwdPF = 12;            // patch rows in flight per warp (~2 us of lead at full load)
constexpr int kFwdRing = kFwdPF + 2;  // + the row being read + the row other lanes may still read

__device__ __forceinline__ void cp_async_w(int cw, float* dst, const float* src) {
    if (cw == 4) cp_async<16>(dst, src);
    else if (cw == 2) cp_async<8>(dst, src);
    else cp_async<4>(dst, src);
}

// shared-memory floats of one ring slot: `cpw` channels x row stride
__device__ __forceinline__ int fwd_row_stride(int X0, int X1, int cw, int jw, int& X0a, int& fwp) {
    X0a = X0 & ~(cw - 1);              // patch rows start on a copy boundary
    fwp = (X1 - X0a + cw) & ~(cw - 1);  // floats copied per row
    return (fwp + jw - 1 + 3) & ~3;     // + zero pad for the padded taps
}

// Register budget matters here (no spills in the row loop): offsets are 32-bit (one
// level's map and one RoI's pooled block are far below 2^31 floats), every lane owns at most one
// copy per patch row (the caller keeps cpw * copies-per-row <= 32), and only scalars cross the
// call boundary.
struct FwdWarpArgs {
    const float* src0;   // feature element (batch, c0, Y0, X0a)
    float* obase;        // pooled element (i, c0, 0, 0)
    const float* ytab;   // packed Y records (shared)
    const int* rcnt;     // pooled rows per patch row (shared)
    const int* xs;       // first feature column of every pooled column (shared)
    const float* wx;     // folded X weights [JW][Pw] (shared)
    float* ring;         // this warp's row slots (shared)
    int sC, sH;          // feature strides in floats
    int osC, osH;        // pooled strides in floats
    int Pw, Ph, R, X0a, fws, cpr, cw, cpw, nc;
    int nring;           // ring slots actually used (<= kFwdRing): wide patch rows get a shallower ring
};

template <int VEC, int JW>
__device__ __noinline__ void fwd_warp_core(const float* src0_, float* obase_, const float* ytab_, const int* rcnt_, const int* xs_, const float* wx_, float* ring_, int sC_, int sH_, int osC_, int osH_, int Pw_, int Ph_, int R_, int X0a_, int fws_, int cpr_, int cw_, int cpw_, int nc_, int nring_) {
    FwdWarpArgs a;   // (scalars across the call, see fwd_warp_tma_core)
    a.src0 = src0_;
    a.obase = obase_;
    a.ytab = ytab_;
    a.rcnt = rcnt_;
    a.xs = xs_;
    a.wx = wx_;
    a.ring = ring_;
    a.sC = sC_;
    a.sH = sH_;
    a.osC = osC_;
    a.osH = osH_;
    a.Pw = Pw_;
    a.Ph = Ph_;
    a.R = R_;
    a.X0a = X0a_;
    a.fws = fws_;
    a.cpr = cpr_;
    a.cw = cw_;
    a.cpw = cpw_;
    a.nc = nc_;
    a.nring = nring_;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int PwV = a.Pw / VEC;
    const int R = a.R, cpw = a.cpw, nc = a.nc;
    const int rowf = cpw * a.fws;  // floats per ring slot
    const int sub = lane / PwV, pv = lane - sub * PwV;
    const bool lane_on = sub < cpw;
    const int subc = lane_on ? sub : 0;
    const float* __restrict__ ytab = a.ytab;
    const int* __restrict__ rcnt = a.rcnt;
    float* const ring = a.ring;
    // the function is not inlined: tell the compiler these all live in shared memory (LDS, not LD)
    __builtin_assume(__isShared(ytab));
    __builtin_assume(__isShared(rcnt));
    __builtin_assume(__isShared(ring));
    __builtin_assume(__isShared(a.xs));
    __builtin_assume(__isShared(a.wx));
    // strip constants (first patch column and X weights of this lane's VEC pooled columns) stay in
    // shared memory and are re-read once per patch row: registers are the scarcer resource here
    const int* const xsp = a.xs + (lane_on ? pv * VEC : 0);
    const float* const wxp = a.wx + (lane_on ? pv * VEC : 0);
    const int nring = a.nring, pf = nring - 2;
    // the pad columns are never written by the copies: they must hold finite values
    for (int q = lane * 4; q < nring * rowf; q += 128) *reinterpret_cast<float4*>(ring + q) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    // ---- producer side: this lane's copies of every patch row --------------------------------------
    // copy q of a ring row is channel q / cpr, floats (q % cpr) * cw ..; a lane owns q = lane (and
    // lane + 32, ... when the rows are wider than a warp's worth of copies)
    const int step = RA_WARPS * cpw;
    const int ncopy = cpw * a.cpr;
    const int cc_c = lane / a.cpr;                     // channel of this lane's first copy inside the batch
    const int cc_x = (lane - cc_c * a.cpr) * a.cw;     // first float of that copy inside the row
    const float* i_src = a.src0 + (warp * cpw + cc_c) * a.sC + cc_x;  // this lane's source, next row to issue
    float* const i_dst = ring + cc_c * a.fws + cc_x;
    int i_cb = warp * cpw, i_r = 0, i_slot = 0;
    auto issue = [&]() {
        if (i_cb < nc) {
#if !defined(DM_DIAG_NO_READS)
            if (cc_c < min(cpw, nc - i_cb)) cp_async_w(a.cw, i_dst + i_slot * rowf, i_src);
            if (ncopy > 32) {
                for (int q = lane + 32; q < ncopy; q += 32) {
                    const int c = q / a.cpr, x = (q - c * a.cpr) * a.cw;
                    if (c < min(cpw, nc - i_cb))
                        cp_async_w(a.cw, ring + i_slot * rowf + c * a.fws + x, i_src + (c - cc_c) * a.sC + (x - cc_x));
                }
            }
#endif
            i_src += a.sH;
            if (++i_r == R) {
                i_r = 0;
                i_cb += step;
                i_src += step * a.sC - R * a.sH;
            }
            i_slot = i_slot + 1 == nring ? 0 : i_slot + 1;
        }
        cp_async_commit();
    };
    // ---- consumer side: next patch row of the stream -> this lane's X-interpolated strip values ----
    const float* pr = ring + subc * a.fws - a.X0a;  // this lane's channel in the slot being read next
    int r_slot = 0;
#if DM_FWD_HOIST
    // strip constants in registers (the kernel is built for 128 registers: 2 CTAs x 8 warps fill the file)
    int xo_r[VEC];
    float wx_r[JW <= 4 ? JW : 1][VEC];
    if (JW <= 4) {
        ld_vec_i<VEC>(xsp, xo_r);
#pragma unroll
        for (int j = 0; j < (JW <= 4 ? JW : 1); ++j) ld_vec<VEC>(wxp + j * a.Pw, wx_r[j]);
    }
#endif
    auto consume = [&](float (&v)[VEC]) {
        issue();
        if (pf == kFwdPF) cp_async_wait<kFwdPF>();
        else if (pf >= 6) cp_async_wait<6>();
        else cp_async_wait<3>();
        __syncwarp();
        int xo[VEC];
#if DM_FWD_HOIST
        if (JW <= 4) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) xo[e] = xo_r[e];
        } else
#endif
        ld_vec_i<VEC>(xsp, xo);
#pragma unroll
        for (int j = 0; j < JW; ++j) {
            float wj[VEC], pj[VEC];
#if DM_FWD_HOIST
            if (JW <= 4) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) wj[e] = wx_r[JW <= 4 ? j : 0][e];
            } else
#endif
            ld_vec<VEC>(wxp + j * a.Pw, wj);
#pragma unroll
            for (int e = 0; e < VEC; ++e) pj[e] = pr[xo[e] + j];
            if (VEC == 1) {
                v[0] = j == 0 ? wj[0] * pj[0] : v[0] + wj[0] * pj[0];
            } else {
#pragma unroll
                for (int e = 0; e + 1 < VEC; e += 2) {
                    const float2 r = j == 0 ? fmul2(make_float2(wj[e], wj[e + 1]), make_float2(pj[e], pj[e + 1]))
                                            : ffma2(make_float2(wj[e], wj[e + 1]), make_float2(pj[e], pj[e + 1]), make_float2(v[e], v[e + 1]));
                    v[e] = r.x;
                    v[e + 1] = r.y;
                }
            }
        }
        if (++r_slot == nring) { r_slot = 0; pr -= (nring - 1) * rowf; } else { pr += rowf; }
    };
    for (int d = 0; d < pf; ++d) issue();

    constexpr int YS = 2 * JW;
    float* o_cb = a.obase + (warp * cpw + subc) * a.osC + pv * VEC;
    for (int cb = warp * cpw; cb < nc; cb += step, o_cb += step * a.osC) {
        const bool on = lane_on && sub < nc - cb;
        float win[JW][VEC];
#pragma unroll
        for (int j = 0; j < JW; ++j) {
            if (j < R) {
                consume(win[j]);
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) win[j][e] = 0.0f;
            }
        }
        float* o = o_cb;
        const float* yrec = ytab;
        int left = a.Ph, base = 0;
        // patch-row major: all pooled rows whose band starts at `base` share one window
        for (;; ++base) {
            const int n = rcnt[base];
#pragma unroll 2
            for (int k = 0; k < n; ++k) {
                float2 w[JW];
                load_yrec<JW>(yrec, w);
                yrec += YS;
                float acc[VEC];
                vmul<VEC>(acc, w[0], win[0]);
#pragma unroll
                for (int j = 1; j < JW; ++j) vfma<VEC>(acc, w[j], win[j]);
#if defined(DM_DIAG_NO_STORES)
                if (on && acc[0] == 123.456f) st_stream_vec<VEC>(o, acc);
#elif defined(DM_DIAG_PLAIN_STORES)
                if (on) { if (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]); else st_stream_vec<VEC>(o, acc); }
#else
                if (on) st_stream_vec<VEC>(o, acc);
#endif
                o += a.osH;
            }
            left -= n;
            if (left <= 0) break;
#pragma unroll
            for (int j = 0; j + 1 < JW; ++j)
#pragma unroll
                for (int e = 0; e < VEC; ++e) win[j][e] = win[j + 1][e];
            if (base + JW < R) {
                consume(win[JW - 1]);
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) win[JW - 1][e] = 0.0f;
            }
        }
        // keep the stream aligned: rows of this batch the walk did not need
        for (int r = base + JW; r < R; ++r) {
            float dummy[VEC];
            consume(dummy);
        }
    }
    cp_async_wait<0>();
}

template <int VEC, int JW>
__device__ __forceinline__ void fwd_warp(const FwdWarpArgs& a) {
    fwd_warp_core<VEC, JW>(a.src0, a.obase, a.ytab, a.rcnt, a.xs, a.wx, a.ring, a.sC, a.sH, a.osC, a.osH, a.Pw, a.Ph, a.R, a.X0a, a.fws, a.cpr, a.cw, a.cpw, a.nc, a.nring);
}

// ---------------------------------------------------------------------------------------------
// Forward, TMA path: the same warp-autonomous strip walk, but the warp's patches arrive as TMA box
// loads -- one cp.async.bulk.tensor.3d per chunk of BH patch rows x 16 / VEC channels x BW columns,
// issued by lane 0 into a ring of chunk slots and signalled on one mbarrier per slot -- instead of
// one 16-byte cp.async per lane and patch row.  Lanes = 4 channels x Pw / VEC strips; a lane
// carries 4 / VEC channels (4 channels apart), so it always produces four pooled values per patch
// row and the control flow of the walk is shared by twice / four times the arithmetic.  The steady state has no copy bookkeeping in the
// row loop: per chunk one parity wait, one warp barrier and (lane 0) one expect_tx + one TMA.
// The last chunk of a channel batch is shifted up so that it ends on the patch's last row (the rows
// it repeats come from L2 and are skipped), so nothing below the patch is fetched.
// Requires Pw / VEC <= 8, patch rows of <= 64 floats from the aligned first column, JX, JY <= JW.
// ---------------------------------------------------------------------------------------------
struct FwdTmaArgs {
    const void* map;     // tensor map of (level, width class), kernel-parameter space
    float* obase;        // pooled element (i, c0, 0, 0)
    const float* ytab;   // packed Y records (shared)
    const int* rcnt;     // pooled rows per patch row (shared)
    const int* xs;       // first feature column of every pooled column (shared)
    const float* wx;     // folded X weights [JW][Pw] (shared)
    float* ring;         // this warp's chunk slots (shared, 128-byte aligned) followed by a zeroed pad
    unsigned bar_sa;     // shared address of this warp's first mbarrier (one per slot)
    unsigned* phase;     // this warp's parity bits, carried from unit to unit (shared)
    int osC, osH;        // pooled strides in floats
    int Pw, Ph, R, X0, Y0, BW, BH, nslot, nc;   // X0: first column of the boxes (multiple of 4)
    int cidx0;           // tensor coordinate of channel c0 of this RoI's image: batch * C + c0
    int chs, rws;        // floats between channels / between rows inside a slot
    int diag;            // measurement only (DM_RA_DIAG): 1 = no loads (compute on stale slots), 2 = loads only
};

// channels per lane on the TMA path: a lane keeps JW window rows x CH channels x VEC columns in
// registers -- 32 values for JW <= 4 (eight independent chains per patch row), 32 for JW = 8
__host__ __device__ constexpr int tma_ch(int vec, int jw) { return vec == 1 ? 4 : (jw <= 4 ? 8 / vec : 4 / vec); }

// (The arguments cross the call as scalars: a struct passed by value to a __noinline__ function goes
// through local memory -- ~30 STL / LDL / generic LD per warp and unit in front of the first TMA request.)
template <int VEC, int JW>
__device__ __noinline__ void fwd_warp_tma_core(const void* map_, float* obase_, const float* ytab_, const int* rcnt_, const int* xs_, const float* wx_, float* ring_, unsigned bar_sa_, unsigned* phase_, int osC_, int osH_, int Pw_, int Ph_, int R_, int X0_, int Y0_, int BW_, int BH_, int nslot_, int nc_, int cidx0_, int chs_, int rws_, int diag_) {
    FwdTmaArgs a;
    a.map = map_;
    a.obase = obase_;
    a.ytab = ytab_;
    a.rcnt = rcnt_;
    a.xs = xs_;
    a.wx = wx_;
    a.ring = ring_;
    a.bar_sa = bar_sa_;
    a.phase = phase_;
    a.osC = osC_;
    a.osH = osH_;
    a.Pw = Pw_;
    a.Ph = Ph_;
    a.R = R_;
    a.X0 = X0_;
    a.Y0 = Y0_;
    a.BW = BW_;
    a.BH = BH_;
    a.nslot = nslot_;
    a.nc = nc_;
    a.cidx0 = cidx0_;
    a.chs = chs_;
    a.rws = rws_;
    a.diag = diag_;
    constexpr int CH = tma_ch(VEC, JW);  // channels per lane
    constexpr int BC = 4 * CH;           // channels per box = per warp pass
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int PwV = a.Pw / VEC;
    const int R = a.R, BH = a.BH, nc = a.nc, nslot = a.nslot;
    const int sub = lane / PwV, pv = lane - sub * PwV;
    const bool lane_on = sub < 4;
    const int subc = lane_on ? sub : 0;
    const float* __restrict__ ytab = a.ytab;
    const int* __restrict__ rcnt = a.rcnt;
    float* const ring = a.ring;
    __builtin_assume(__isShared(ytab));
    __builtin_assume(__isShared(rcnt));
    __builtin_assume(__isShared(ring));
    __builtin_assume(__isShared(a.xs));
    __builtin_assume(__isShared(a.wx));
    __builtin_assume(__isShared(a.phase));
    const int step = RA_WARPS * BC;
    const int rws = a.rws;
    const int qs = 4 * a.chs;                              // floats between a lane's channels in a slot
    const int slotf = BC * a.BW * BH;                      // floats per slot
    const unsigned slot_bytes = (unsigned)slotf * 4u;
    const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring);
    // the padded taps of the last channel's last row run past the ring: that pad must be finite
    if (lane < 8) ring[nslot * slotf + lane] = 0.0f;
    const int nchunk = (R + BH - 1) / BH;                  // chunks per channel batch
    const int last_shift = R >= BH ? nchunk * BH - R : 0;  // rows the last chunk repeats
    const int nbatch = warp * BC < nc ? (nc - warp * BC + step - 1) / step : 0;
    const int q_total = nbatch * nchunk;
    unsigned par = *a.phase;
    __syncwarp();

    // ---- producer side (warp-uniform bookkeeping, lane 0 issues) -------------------------------
    int q_issue = 0, i_k = 0, i_slot = 0;
    int i_c = a.cidx0 + warp * BC;                         // channel coordinate of the next chunk
    auto issue = [&]() {
        if (q_issue < q_total) {
            if (lane == 0 && !(a.diag & 1)) {
                const unsigned bar = a.bar_sa + 8u * (unsigned)i_slot;
                const int y = a.Y0 + i_k * BH - (i_k == nchunk - 1 ? last_shift : 0);
                mbar_expect_tx(bar, slot_bytes);
                const bool rowmajor = a.rws > a.chs;   // tensor dims (x, channel, y) or (x, y, channel)
                tma_load_3d(ring_sa + (unsigned)i_slot * slot_bytes, a.map, a.X0, rowmajor ? i_c : y, rowmajor ? y : i_c, bar);
            }
            ++q_issue;
            if (++i_k == nchunk) { i_k = 0; i_c += step; }
            i_slot = i_slot + 1 == nslot ? 0 : i_slot + 1;
        }
    };
    // ---- consumer side ------------------------------------------------------------------------
    int q_cons = 0, c_k = 0, c_slot = 0, rows_left = 0;
    const float* pr = ring;   // first float of the current patch row's channel 0 (warp-uniform)
    auto acquire = [&]() {    // the next chunk of the stream becomes current
        if (q_cons > 0) {     // the slot just drained takes the next request
            __syncwarp();
            issue();
        }
        if (!(a.diag & 1)) {
            mbar_wait(a.bar_sa + 8u * (unsigned)c_slot, (par >> c_slot) & 1u);
            par ^= 1u << c_slot;
        }
        const int skip = c_k == nchunk - 1 ? last_shift : 0;
        rows_left = min(BH, R) - skip;
        pr = ring + c_slot * slotf + skip * rws;
        c_slot = c_slot + 1 == nslot ? 0 : c_slot + 1;
        ++q_cons;
        if (++c_k == nchunk) c_k = 0;
    };
    // strip constants: first patch column (as an offset inside a slot row) and X weights of this
    // lane's VEC pooled columns
    int xo_r[VEC];
    {
        int xs_l[VEC];
        ld_vec_i<VEC>(a.xs + (lane_on ? pv * VEC : 0), xs_l);
#pragma unroll
        for (int e = 0; e < VEC; ++e) xo_r[e] = xs_l[e] - a.X0 + subc * a.chs;
    }
    const float* const wxp = a.wx + (lane_on ? pv * VEC : 0);
    float wx_r[JW <= 4 ? JW : 1][VEC];
    if (JW <= 4) {
#pragma unroll
        for (int j = 0; j < (JW <= 4 ? JW : 1); ++j) ld_vec<VEC>(wxp + j * a.Pw, wx_r[j]);
    }
    // X-interpolates the next patch row of the stream: v[q][e] = channel sub + 4 q, pooled column e
    auto consume = [&](float (&v)[CH][VEC]) {
        if (rows_left == 0) acquire();
        --rows_left;
#pragma unroll
        for (int j = 0; j < JW; ++j) {
            float wj[VEC];
            if (JW <= 4) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) wj[e] = wx_r[JW <= 4 ? j : 0][e];
            } else {
                ld_vec<VEC>(wxp + j * a.Pw, wj);
            }
#pragma unroll
            for (int q = 0; q < CH; ++q) {
                float pj[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) pj[e] = pr[xo_r[e] + q * qs + j];
                if (VEC == 1) {
                    v[q][0] = j == 0 ? wj[0] * pj[0] : v[q][0] + wj[0] * pj[0];
                } else {
#pragma unroll
                    for (int e = 0; e + 1 < VEC; e += 2) {
                        const float2 r = j == 0 ? fmul2(make_float2(wj[e], wj[e + 1]), make_float2(pj[e], pj[e + 1]))
                                                : ffma2(make_float2(wj[e], wj[e + 1]), make_float2(pj[e], pj[e + 1]),
                                                        make_float2(v[q][e], v[q][e + 1]));
                        v[q][e] = r.x;
                        v[q][e + 1] = r.y;
                    }
                }
            }
        }
        pr += rws;
    };
    auto skip_rows = [&](int n) {   // rows of this batch the walk does not need
        while (n > 0) {
            if (rows_left == 0) acquire();
            const int m = min(n, rows_left);
            rows_left -= m;
            pr += m * rws;
            n -= m;
        }
    };
    for (int d = 0; d < nslot; ++d) issue();

    constexpr int YS = 2 * JW;
    const int oq = 4 * a.osC;   // floats between a lane's channels in the pooled block
    float* o_cb = a.obase + (warp * BC + subc) * a.osC + pv * VEC;
    for (int cb = warp * BC; cb < nc; cb += step, o_cb += step * a.osC) {
        if (a.diag & 2) { skip_rows(R); continue; }
        bool on[CH];
#pragma unroll
        for (int q = 0; q < CH; ++q) on[q] = lane_on && sub + 4 * q < nc - cb;
        float win[JW][CH][VEC];
#pragma unroll
        for (int j = 0; j < JW; ++j)
#pragma unroll
            for (int q = 0; q < CH; ++q)
#pragma unroll
                for (int e = 0; e < VEC; ++e) win[j][q][e] = 0.0f;
        int used = 0;   // patch rows of this batch consumed so far
        float* o = o_cb;
        const float* yrec = ytab;
        int left = a.Ph;
        // patch-row major: one patch row enters the window per step (the first JW - 1 steps only fill
        // it); all pooled rows whose band starts at `base` share one window.  The window ROTATES: patch
        // row r lives in win[r % JW], and the walk is unrolled JW deep so that the slot of every logical
        // band row is a compile-time index -- no register moves when the window advances.
#if DM_TMA_ROTATE
        bool more = true;
        for (int r0 = 0; more; r0 += JW) {
#pragma unroll
            for (int u = 0; u < JW; ++u) {
                const int r = r0 + u;
                if (r < R) {
                    consume(win[u]);
                    ++used;
                } else {
#pragma unroll
                    for (int q = 0; q < CH; ++q)
#pragma unroll
                        for (int e = 0; e < VEC; ++e) win[u][q][e] = 0.0f;
                }
                const int base = r - (JW - 1);
                if (base < 0) continue;
                const int n = rcnt[base];
                for (int k = 0; k < n; ++k) {
                    float2 w[JW];
                    load_yrec<JW>(yrec, w);
                    yrec += YS;
#pragma unroll
                    for (int q = 0; q < CH; ++q) {
                        float acc[VEC];
                        vmul<VEC>(acc, w[0], win[(u + 1) % JW][q]);
#pragma unroll
                        for (int j = 1; j < JW; ++j) vfma<VEC>(acc, w[j], win[(u + 1 + j) % JW][q]);
                        if (on[q]) st_out_vec<VEC>(o + q * oq, acc);
                    }
                    o += a.osH;
                }
                left -= n;
                if (left <= 0) { more = false; break; }
            }
        }
#else
        for (int base = 1 - JW;; ++base) {
#pragma unroll
            for (int j = 0; j + 1 < JW; ++j)
#pragma unroll
                for (int q = 0; q < CH; ++q)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) win[j][q][e] = win[j + 1][q][e];
            if (base + JW - 1 < R) {
                consume(win[JW - 1]);
                ++used;
            } else {
#pragma unroll
                for (int q = 0; q < CH; ++q)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) win[JW - 1][q][e] = 0.0f;
            }
            if (base < 0) continue;
            const int n = rcnt[base];
            for (int k = 0; k < n; ++k) {
                float2 w[JW];
                load_yrec<JW>(yrec, w);
                yrec += YS;
#pragma unroll
                for (int q = 0; q < CH; ++q) {
                    float acc[VEC];
                    vmul<VEC>(acc, w[0], win[0][q]);
#pragma unroll
                    for (int j = 1; j < JW; ++j) vfma<VEC>(acc, w[j], win[j][q]);
                    if (on[q]) st_out_vec<VEC>(o + q * oq, acc);
                }
                o += a.osH;
            }
            left -= n;
            if (left <= 0) break;
        }
#endif
        skip_rows(R - used);   // keep the stream aligned: every batch spans exactly R rows
    }
    // every requested chunk has been waited for: nothing is in flight when the slots are reused
    __syncwarp();
    if (lane == 0) *a.phase = par;
}

template <int VEC, int JW>
__device__ __forceinline__ void fwd_warp_tma(const FwdTmaArgs& a) {
    fwd_warp_tma_core<VEC, JW>(a.map, a.obase, a.ytab, a.rcnt, a.xs, a.wx, a.ring, a.bar_sa, a.phase, a.osC, a.osH, a.Pw, a.Ph, a.R, a.X0, a.Y0, a.BW, a.BH, a.nslot, a.nc, a.cidx0, a.chs, a.rws, a.diag);
}

// ---------------------------------------------------------------------------------------------
// Forward, generic path (bands wider than 8, or bands too tall for one tile): two passes through
// shared memory, pooled rows processed in tiles [p0, p1).
// ---------------------------------------------------------------------------------------------
template <int VEC>
__device__ DM_COLD void fwd_tile(const LevelDesc& Lv, const BucketDesc& B, const Tables t, float* tile,
                         int batch, int i, int c0, int cs, int p0, int p1) {
    const int Pw = B.pw;
    const int fw = t.X1 - t.X0 + 1;
    const int Yt0 = t.ys[p0];
    const int R = band_rows(t, p0, p1);
    float* V = tile;                    // [cs][R][Pw]
    float* patch = tile + cs * R * Pw;  // [cs][R][fw]
    stage_patch(Lv, patch, batch, c0, cs, Yt0, R, t.X0, fw, fw);
    __syncthreads();
    {   // X pass: V[c][r][pw] = sum_j wx[j][pw] * patch[c][r][xs[pw]-X0+j]
        const int sh = pow2_shift_ge(Pw);
        const int items = (cs * R) << sh;
        const int JX = t.JX;
        for (int vi = threadIdx.x; vi < items; vi += RA_THREADS) {
            const int row = vi >> sh, pw = vi & ((1 << sh) - 1);
            if (pw >= Pw) continue;
            const float* prow = patch + row * fw;
            const int x0 = t.xs[pw] - t.X0;
            float acc = 0.0f;
            for (int j = 0; j < JX; ++j) acc += t.wx[j * Pw + pw] * prow[min(x0 + j, fw - 1)];
            V[row * Pw + pw] = acc;
        }
    }
    __syncthreads();
    {   // Y pass: out[c][ph][pw..pw+VEC) = sum_j wy[j][ph] * V[c][ys[ph]-Yt0+j][pw..]
        const int PwV = Pw / VEC;
        const int sh = pow2_shift_ge(PwV);
        const int nph = p1 - p0;
        const int items = (cs * nph) << sh;
        const int JY = t.JY;
        FastDiv fdP;
        fdP.init(nph);
        float* obase = B.ptr + (long long)i * B.sN + (long long)c0 * B.sC;
        for (int vi = threadIdx.x; vi < items; vi += RA_THREADS) {
            const int row = vi >> sh, pv = vi & ((1 << sh) - 1);
            if (pv >= PwV) continue;
            const int c = fdP.div(row), ph = p0 + (row - c * nph);
            const int r0 = t.ys[ph] - Yt0;
            const float* vc = V + c * R * Pw + pv * VEC;
            float acc[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] = 0.0f;
            for (int j = 0; j < JY; ++j) {
                const float w = t.wy[j * B.ph + ph];
                float v[VEC];
                ld_vec<VEC>(vc + min(r0 + j, R - 1) * Pw, v);
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc[e] += w * v[e];
            }
            st_stream_vec<VEC>(obase + (long long)c * B.sC + (long long)ph * B.sH + (long long)(pv * VEC) * B.sW, acc);
        }
    }
}

// per-CTA state of the TMA path: one mbarrier per (warp, slot) and the warps' parity bits
struct TmaShared {
    unsigned long long* bars;   // [warps][kTmaMaxSlots]
    unsigned* phase;            // [warps]
};

// (Experiment, off by default -- DM_RA_L2_PREFETCH=1.)  While the first two warps build the unit's tables, the
// others ask the TMA unit to pull the unit's patch into L2 (cp.async.bulk.prefetch.tensor): the walk's box loads then find their data in L2, so
// neither the cold start of a unit's ring nor its steady state waits on DRAM latency.  The bounds come
// straight from the first and last sample coordinate of each axis (what the table scan will find,
// give or take rejected samples -- this is a hint, a column too many or too few costs nothing but
// a few bytes); the last chunk of a batch is shifted up like the loader's, so no row below the patch
// is touched.
template <int VEC>
__device__ __forceinline__ void prefetch_patch(const RaParams& p, const TmaMaps& tm, const LevelDesc& Lv,
                                               const RoiGeom& g, int lv, int Pw, int Ph, int cidx0, int nc) {
    const int set = VEC == 4 ? 1 : 2;                // the boxes of the JW <= 4 walk: 8 / 16 channels
    const int BC = 4 << set;
    const unsigned mask = p.tma_mask[lv][set];
    if (!mask || g.gw <= 0 || g.gh <= 0) return;
    const float vx0 = sample_coord(g.rsw, g.bw, g.gw, 0, 0), vx1 = sample_coord(g.rsw, g.bw, g.gw, Pw - 1, g.gw - 1);
    const float vy0 = sample_coord(g.rsh, g.bh, g.gh, 0, 0), vy1 = sample_coord(g.rsh, g.bh, g.gh, Ph - 1, g.gh - 1);
    if (!(vx1 >= -1.0f && vx0 <= (float)Lv.W && vy1 >= -1.0f && vy0 <= (float)Lv.H)) return;
    const int X0 = min(max((int)floorf(fmaxf(vx0, 0.0f)), 0), Lv.W - 1);
    const int X1 = min(max((int)floorf(fmaxf(vx1, 0.0f)) + 1, 0), Lv.W - 1);
    const int Y0 = min(max((int)floorf(fmaxf(vy0, 0.0f)), 0), Lv.H - 1);
    const int Y1 = min(max((int)floorf(fmaxf(vy1, 0.0f)) + 1, 0), Lv.H - 1);
    const int X0a = X0 & ~3, fwa = X1 - X0a + 1, R = Y1 - Y0 + 1;
    if (fwa > kTmaMaxBW) return;
    int cls = 0;
    while (cls < kNumBW - 1 && tma_bw(cls) < fwa) ++cls;
    if (!((mask >> cls) & 1u)) return;
    const int BH = tma_bh(set, cls);
    const int nchunk = (R + BH - 1) / BH, nbatch = (nc + BC - 1) / BC;
    const void* map = &tm.m[lv][set][cls];
    const int stride = (RA_WARPS - 2) * 32;
    for (int q = (int)threadIdx.x - 64; q < nbatch * nchunk; q += stride) {
        const int b = q / nchunk, k = q - b * nchunk;
        const int y = (k == nchunk - 1 && R >= BH) ? Y1 + 1 - BH : Y0 + k * BH;
        const int c = cidx0 + b * BC;
        if (p.tma_rowmajor) tma_prefetch_3d(map, X0a, c, y);
        else tma_prefetch_3d(map, X0a, y, c);
    }
}

template <int VEC>
__device__ void fwd_unit(const RaParams& p, const TmaMaps& tm, const TmaShared ts, const Unit& un,
                         const int* s_seg, float* smem, int* stat) {
    const BucketDesc& B = p.bk[un.b];
    const int c0 = un.slab * B.cg, c1 = min(c0 + B.cg, p.C);
    float roi[5];
    int lv;
    unit_roi(p, un, s_seg, roi, lv);
    const int batch = (int)roi[0];
    if (lv < 0 || lv >= p.L || batch < 0 || batch >= p.lv[lv < 0 || lv >= p.L ? 0 : lv].N) {
        zero_unit(B, un.i, c0, c1);
        return;
    }
    const LevelDesc& Lv = p.lv[lv];
    const RoiGeom g = p.mode ? point_geom(roi, Lv.scale, p.aligned) : roi_geom(roi, Lv.scale, B.ph, B.pw, p.sampling_ratio, p.aligned);
    if (g.mode && !(g.bw >= 0.0f && g.bh >= 0.0f)) {
        // SimpleRoIAlign on a box with x2 < x1 or y2 < y1: the points run backwards, so the bands
        // are not monotone -- evaluate sample by sample
        direct_unit<false>(Lv, B, g, batch, un.i, c0, c1);
        return;
    }
    if (p.l2_prefetch && g.mode == 0 && B.pw / VEC <= 8 && lv < kTmaLevels && RA_WARPS > 2 && (threadIdx.x >> 5) >= 2)
        prefetch_patch<VEC>(p, tm, Lv, g, lv, B.pw, B.ph, batch * p.C + c0, c1 - c0);
    Tables t;
    bool fits;
    if (!build_tables(g, B.ph, B.pw, Lv.H, Lv.W, smem, p.smem_floats, stat, t, fits)) {
        zero_unit(B, un.i, c0, c1);
        return;
    }
    const int fw = fits ? t.X1 - t.X0 + 1 : 0;
    const int avail = p.smem_floats - (fits ? t.floats : 0);
    // smallest possible generic tile: one channel, one pooled row
    if (!fits || (long long)t.JY * (fw + B.pw) > avail) {
        direct_unit<false>(Lv, B, g, batch, un.i, c0, c1);
        return;
    }
    float* tile = smem + t.floats;
    const int Rfull = t.Y1 - t.Y0 + 1;
    const int wc = window_class(max(t.JX, t.JY));
    const int PwV = B.pw / VEC;
    if (wc && PwV <= 8 && fw <= kTmaMaxBW - 3 && lv < kTmaLevels && B.sW == 1 &&
        B.sC < (1 << 24) && B.sH < (1 << 24)) {
        // TMA path: the warps' patches arrive as box loads of 16 / VEC channels x BH rows x BW columns.
        // A box must start on a 16-byte boundary of global memory (an unaligned first column is an
        // illegal instruction: tools/tma_probe.cu), so it starts at X0 rounded down to 4 columns
        const int BC = 4 * tma_ch(VEC, wc);
        const int set = BC == 4 ? 0 : (BC == 8 ? 1 : 2);
        const int X0a = t.X0 & ~3;
        const int fwa = t.X1 - X0a + 1;
        int cls = 0;
        while (cls < kNumBW - 1 && tma_bw(cls) < fwa) ++cls;
        const int BW = tma_bw(cls), BH = tma_bh(set, cls);
        const int slot_bytes = BC * BW * BH * 4;
        // the slots must be 128-byte aligned; every warp's ring ends in a 128-byte pad
        const unsigned tile_sa = (unsigned)__cvta_generic_to_shared(tile);
        const int lead = (int)((128u - (tile_sa & 127u)) & 127u);
        const int per_warp = ((avail * 4 - lead) / RA_WARPS) & ~127;
        int nslot = (per_warp - 128) / slot_bytes;
        nslot = nslot > p.tma_slots ? p.tma_slots : nslot;
        if (((p.tma_mask[lv][set] >> cls) & 1u) && nslot >= 2 && fwa <= BW) {
            const int warp = threadIdx.x >> 5;
            FwdTmaArgs a;
            a.map = &tm.m[lv][set][cls];
            a.obase = B.ptr + (long long)un.i * B.sN + (long long)c0 * B.sC;
            a.ytab = t.ytab; a.rcnt = t.rcnt; a.xs = t.xs; a.wx = t.wx;
            a.ring = reinterpret_cast<float*>(reinterpret_cast<char*>(tile) + lead + (size_t)warp * per_warp);
            a.bar_sa = (unsigned)__cvta_generic_to_shared(ts.bars + warp * kTmaMaxSlots);
            a.phase = ts.phase + warp;
            a.osC = (int)B.sC; a.osH = (int)B.sH;
            a.Pw = B.pw; a.Ph = B.ph; a.R = Rfull; a.X0 = X0a; a.Y0 = t.Y0; a.BW = BW; a.BH = BH;
            a.nslot = nslot; a.nc = c1 - c0;
            a.cidx0 = batch * p.C + c0;
            a.chs = p.tma_rowmajor ? BW : BH * BW;
            a.rws = p.tma_rowmajor ? BC * BW : BW;
            a.diag = p.diag;
            if (p.diag & 8) return;   // measurement only: tables and unit bookkeeping alone
            if (wc == 2) fwd_warp_tma<VEC, 2>(a);
            else if (wc == 4) fwd_warp_tma<VEC, 4>(a);
            else fwd_warp_tma<VEC, 8>(a);
            return;
        }
    }
    if (p.diag & 4) return;   // measurement only: units that do not take the TMA path are skipped
    if (wc && PwV <= 32 && B.sW == 1 && Lv.sW == 1 && Lv.sC < (1 << 24) && Lv.sH < (1 << 24) &&
        B.sC < (1 << 24) && B.sH < (1 << 24)) {
        // fast path: every warp streams its channels' patch rows through a private ring of row slots
        const int slice = (avail / RA_WARPS) & ~3;
        FwdWarpArgs a;
        int fwp;
        a.cw = Lv.cw;
        a.fws = fwd_row_stride(t.X0, t.X1, a.cw, wc, a.X0a, fwp);
        a.cpr = fwp / a.cw;
        // channels per warp pass and ring depth: the full ring (12 rows in flight) when it fits;
        // patch rows wider than that get a shallower ring (6 or 3 rows in flight -- the bytes in
        // flight stay about the same) and, beyond 32 copies per row, several copies per lane
        a.cpw = min(min(32 / PwV, slice / (kFwdRing * a.fws)), 32 / a.cpr);
        a.nring = kFwdRing;
        if (a.cpw < 32 / PwV) {
            // keep the lanes busy first (as many channels per pass as the strips allow), then give
            // the ring what is left: 8 slots (6 rows in flight) or 5 (3 in flight)
            a.nring = 0;
            for (int c = 32 / PwV; c >= 1 && a.nring == 0; --c) {
                const int nr = slice / (c * a.fws);
                if (nr >= 5) { a.cpw = c; a.nring = nr >= kFwdRing ? kFwdRing : (nr >= 8 ? 8 : 5); }
            }
        }
        if (a.cpw >= 1 && a.nring > 0) {
            a.src0 = Lv.ptr + (long long)batch * Lv.sN + (long long)c0 * Lv.sC + (long long)t.Y0 * Lv.sH + a.X0a;
            a.obase = B.ptr + (long long)un.i * B.sN + (long long)c0 * B.sC;
            a.ytab = t.ytab; a.rcnt = t.rcnt; a.xs = t.xs; a.wx = t.wx;
            a.ring = tile + (threadIdx.x >> 5) * slice;
            a.sC = (int)Lv.sC; a.sH = (int)Lv.sH; a.osC = (int)B.sC; a.osH = (int)B.sH;
            a.Pw = B.pw; a.Ph = B.ph; a.R = Rfull; a.nc = c1 - c0;
            if (wc == 2) fwd_warp<VEC, 2>(a);
            else if (wc == 4) fwd_warp<VEC, 4>(a);
            else fwd_warp<VEC, 8>(a);
            return;
        }
    }
    if (p.mode && t.JX <= 2 && t.JY <= 2) {
        point_sparse_unit<false>(Lv, B, t, batch, un.i, c0, c1);
        return;
    }
    const int per_row = fw + B.pw;
    if ((long long)Rfull * per_row <= avail) {
        const int cs_max = min(c1 - c0, avail / (Rfull * per_row));
        for (int c = c0; c < c1; c += cs_max) {
            fwd_tile<VEC>(Lv, B, t, tile, batch, un.i, c, min(cs_max, c1 - c), 0, B.ph);
            __syncthreads();
        }
    } else {
        // tall RoI: one channel at a time, pooled rows in tiles whose band fits
        for (int c = c0; c < c1; ++c) {
            int q0 = 0;
            while (q0 < B.ph) {
                int q1 = q0 + 1;
                while (q1 < B.ph && (long long)band_rows(t, q0, q1 + 1) * per_row <= avail) ++q1;
                fwd_tile<VEC>(Lv, B, t, tile, batch, un.i, c, 1, q0, q1);
                __syncthreads();
                q0 = q1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Backward unit
// ---------------------------------------------------------------------------------------------
// Fast path: warp-autonomous strip walk.  Lanes = cpw channels x PwV column strips; every lane
// walks all pooled rows top to bottom with the JW band rows it is accumulating in registers.
// grad_out rows arrive through a per-lane cp.async ring (kRing-1 rows in flight per lane, no
// registers tied up).  All lanes of a warp see the same pooled-row sequence, so a band row
// completes for the whole warp at once: it is dropped into the warp's row buffer, the
// patch-gradient row is gathered from it (lane = feature column, its transposed X weights held in
// registers) and reduced into the gradient map with one RED per element.
#ifndef DM_BWD_RING
#define DM_BWD_RING 8
#endif
constexpr int kRing = DM_BWD_RING;   // ring depth of the grad_out prefetch, in pooled rows (VEC == 1 path)
// VEC > 1 (dense, 16-byte aligned grad_out): the warp's grad_out rows arrive as bulk copies of
// kBulkRows pooled rows x cpw channels per chunk through a ring of kBulkSlots chunk slots
#ifndef DM_BULK_ROWS
#define DM_BULK_ROWS 4
#endif
#ifndef DM_BULK_SLOTS
#define DM_BULK_SLOTS 4
#endif
#ifndef DM_BULK_ENABLE
#define DM_BULK_ENABLE 0   // measured slower than the per-lane cp.async ring: 12.2 / 11.0 ms vs 9.9 ms
#endif
#ifndef DM_BWD_SMEM_KB
#define DM_BWD_SMEM_KB 108
#endif
constexpr int kBulkRows = DM_BULK_ROWS;
constexpr int kBulkSlots = DM_BULK_SLOTS;
constexpr int kBulkSlotFloats = 128 * kBulkRows;  // >= cpw * kBulkRows * Pw  (cpw * Pw <= 32 * VEC = 128)
// floats of a warp's grad_out ring: chunk slots + one 8-byte mbarrier per slot, or the cp.async ring
__host__ __device__ constexpr int kBwdRingFloats(int vec) {
    return (vec > 1 && DM_BULK_ENABLE) ? kBulkSlots * kBulkSlotFloats + ((2 * kBulkSlots + 3) & ~3) : kRing * 32 * vec;
}
constexpr int kBwdRowBuf = 128;  // floats of the warp's row buffer: cpw rows of Pw rounded up to 4
constexpr int kBwdPad = 84;      // zeros after the row buffer: the aligned tap windows may overrun it

// Taps the register windows of the fast retire do not cover: feature columns beyond the lanes'
// reach and pooled columns beyond a lane's window (rare geometries).  `cover` = pooled columns a
// lane group's windows span, `cols` = feature columns the lanes own.
__device__ __noinline__ void bwd_retire_wide(const float* rowbuf, int Pws, int nact, int fw, int cols, int cover,
                                             const int* plo, const int* pcnt, const float* wxT, int TW,
                                             float* drow0, int dsC, int lane) {
    __builtin_assume(__isShared(rowbuf));
    __builtin_assume(__isShared(plo));
    __builtin_assume(__isShared(pcnt));
    __builtin_assume(__isShared(wxT));
    for (int s2 = 0; s2 < nact; ++s2) {
        const float* ur = rowbuf + s2 * Pws;
        float* drow = drow0 + s2 * dsC;   // points at column 0 of the patch row
        for (int x = lane; x < fw; x += 32) {
            const int lo = plo[x], n = pcnt[x];
            const float* up = ur + lo;
            const float* wp = wxT + x * TW;
            float a = 0.0f;
            for (int q = (x < cols ? min(n, (lo & ~3) + cover - lo) : 0); q < n; ++q) a += wp[q] * up[q];
            if (a != 0.0f) red_add(drow + x, a);
        }
    }
}

struct BwdWarpArgs {
    const float* gbase;  // grad_out element (i, c0, 0, 0)
    float* dbase;        // gradient-map element (batch, c0, Y0, X0)
    const float* ytab;   // packed Y records (shared)
    const int* rcnt;     // pooled rows per band row (shared)
    const int* plo;      // first pooled column touching feature column x (shared)
    const int* pcnt;     // pooled columns touching feature column x (shared)
    const float* wxT;    // transposed X weights [fw][TW] (shared)
    float* wsm;          // this warp's scratch (shared)
    int gsC, gsH, dsC, dsH;
    int Ph, Pw, R, fw, TW, cpw, nc;
    int split;           // lanes per feature column (1, 2 or 4): taps of one column split across lanes
    int wide;            // some taps are outside the register windows
};

// The hot loop lives in a __noinline__ function taking plain scalars so that every instantiation
// gets its own register allocation and nothing is re-derived from the kernel-parameter structs
// inside the loop.  Requirements (checked by the caller): unit inner stride of grad_out and of the
// gradient map, Pw / VEC <= 32, JY <= JW.
// Retire: lane group `part` of feature column x holds the X weights of the 4 * NV pooled columns
// starting at (plo[x] & ~3) + 4 * NV * part in registers, reads them from the row buffer with NV
// 16-byte loads, and the parts are summed with shuffles: one RED per touched feature pixel.
template <int VEC, int JW, int NV>
__device__ __noinline__ void bwd_warp_core_s(const float* gbase_, float* dbase_, const float* ytab_, const int* rcnt_, const int* plo_, const int* pcnt_, const float* wxT_, float* wsm_, int gsC_, int gsH_, int dsC_, int dsH_, int Ph_, int Pw_, int R_, int fw_, int TW_, int cpw_, int nc_, int split_, int wide_) {
    BwdWarpArgs a;   // (scalars across the call, see fwd_warp_tma_core)
    a.gbase = gbase_;
    a.dbase = dbase_;
    a.ytab = ytab_;
    a.rcnt = rcnt_;
    a.plo = plo_;
    a.pcnt = pcnt_;
    a.wxT = wxT_;
    a.wsm = wsm_;
    a.gsC = gsC_;
    a.gsH = gsH_;
    a.dsC = dsC_;
    a.dsH = dsH_;
    a.Ph = Ph_;
    a.Pw = Pw_;
    a.R = R_;
    a.fw = fw_;
    a.TW = TW_;
    a.cpw = cpw_;
    a.nc = nc_;
    a.split = split_;
    a.wide = wide_;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Pw = a.Pw, Ph = a.Ph, R = a.R, cpw = a.cpw, nc = a.nc;
    const int PwV = Pw / VEC;
    const int Pws = (Pw + 3) & ~3;
    const int sub = lane / PwV, pv = lane - sub * PwV;
    const bool lane_on = sub < cpw;
    const float* __restrict__ ytab = a.ytab;
    const int* __restrict__ rcnt = a.rcnt;
    __builtin_assume(__isShared(ytab));
    __builtin_assume(__isShared(rcnt));
    __builtin_assume(__isShared(a.wsm));
    __builtin_assume(__isShared(a.plo));
    __builtin_assume(__isShared(a.pcnt));
    __builtin_assume(__isShared(a.wxT));
    // warp-private shared memory: the grad_out ring ([kRing][32 lanes] floats for VEC == 1, bulk
    // chunk slots + their mbarriers otherwise), then [cpw][Pws] row buffer followed by kBwdPad zeros
    constexpr bool BULK = VEC > 1 && DM_BULK_ENABLE;
    float* const ring = a.wsm + (BULK ? 0 : lane * VEC);
    const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring);
    const unsigned bar_sa = (unsigned)__cvta_generic_to_shared(a.wsm + kBulkSlots * kBulkSlotFloats);
    float* const rowbuf = a.wsm + kBwdRingFloats(VEC);
    float* const myrow = rowbuf + (lane_on ? sub : 0) * Pws + pv * VEC;
    // zero weights meet whatever lies beyond a window's taps: it must be finite
    for (int q = lane; q < kBwdRowBuf + kBwdPad; q += 32) rowbuf[q] = 0.0f;
    // this lane as owner of a feature column: its tap window, weights in registers
    const int cols = 32 / a.split;
    const int part = lane / cols, xl = lane - part * cols;
    const bool xon = xl < a.fw;
    const int xlo = xon ? a.plo[xl] : 0, xn = xon ? a.pcnt[xl] : 0;
    const int pa = xon ? (xlo & ~3) + 4 * NV * part : 0;
    const float* const upx = rowbuf + pa;
    float wq[4 * NV];
#pragma unroll
    for (int q = 0; q < 4 * NV; ++q) {
        const int t = pa + q - xlo;
        wq[q] = (t >= 0 && t < xn) ? a.wxT[xl * a.TW + t] : 0.0f;
    }
    const bool red_on = xon && part == 0;
    const int step = RA_WARPS * cpw;
    // ---- bulk ring state: the chunk stream runs over (channel batch, rows) without a break --------
    const int nch = (Ph + kBulkRows - 1) / kBulkRows;   // chunks per channel batch
    int p_cb = warp * cpw, p_ch = 0;                    // next chunk to request
    unsigned p_slot = 0;
    auto bulk_issue = [&]() {                           // warp-uniform bookkeeping, lane 0 issues
        if (p_cb < nc) {
            if (lane == 0) {
                const int rows = min(kBulkRows, Ph - p_ch * kBulkRows);
                const int np = min(cpw, nc - p_cb);
                const unsigned bytes = (unsigned)(rows * Pw * 4);
                const unsigned bar = bar_sa + 8 * p_slot;
                mbar_expect_tx(bar, bytes * np);
                const float* src = a.gbase + p_cb * a.gsC + p_ch * kBulkRows * Pw;
                unsigned dst = ring_sa + p_slot * (kBulkSlotFloats * 4);
                for (int s2 = 0; s2 < np; ++s2) {
                    bulk_g2s(dst, src, bytes, bar);
                    src += a.gsC;
                    dst += kBulkRows * Pw * 4;
                }
            }
            if (++p_ch == nch) { p_ch = 0; p_cb += step; }
            p_slot = p_slot + 1 == kBulkSlots ? 0 : p_slot + 1;
        }
    };
    unsigned c_slot = 0, c_par = 0;                     // chunk being consumed
    int c_left = 0;                                     // its rows not yet consumed
    const float* gp = ring;                             // this lane's next row in it
    if (BULK) {
        if (lane == 0) {
            for (int q = 0; q < kBulkSlots; ++q) mbar_init(bar_sa + 8 * q, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < kBulkSlots; ++q) bulk_issue();
    }
    __syncwarp();

    const float* gwarp = a.gbase + pv * VEC;
    constexpr int YS = 2 * JW;
    // ring slots by byte offset: the ring is a power of two in size
    constexpr unsigned kSlotB = 32 * VEC * 4, kRingMask = kRing * kSlotB - 1;
    // ---- cp.async ring state: this lane's grad_out rows form ONE stream over (channel batch, pooled
    // row), so the first rows of the next batch are already in flight while the current one retires
    // and no batch starts on a cold load.  The next row to request is p_base + 4 * p_off (a 32-bit
    // offset: one wide multiply-add per address instead of a 64-bit pointer carried around the loop).
    const float* const p_base = gwarp + (warp * cpw + (lane_on ? sub : 0)) * a.gsC;
    int p_off = 0, p_rows = Ph;   // float offset of the next row, rows of its batch not yet requested
    // a lane is live for its first batches only (the last batch of a slab may be partial): count rows
    int p_cnt = lane_on && nc - warp * cpw > sub ? Ph * ((nc - warp * cpw - sub + step - 1) / step) : 0;
    const int p_jump = step * a.gsC - Ph * a.gsH;   // from the end of a batch to the start of the next
    auto request = [&](unsigned slot_sa) {
        cp_async_zfill_sa<VEC * 4>(slot_sa, p_base, p_off, p_cnt > 0);
        --p_cnt;
        p_off += a.gsH;
        if (--p_rows == 0) {
            p_rows = Ph;
            p_off += p_jump;
        }
        cp_async_commit();
    };
    if (!BULK && warp * cpw < nc) {
#pragma unroll
        for (int d = 0; d < kRing - 1; ++d) request(ring_sa + d * kSlotB);
    }
    unsigned off_w = (kRing - 1) * kSlotB;  // ring slot the next prefetch lands in
    unsigned off_r = 0;                     // ring slot holding the next pooled row
    for (int cb = warp * cpw; cb < nc; cb += step) {
        const int nact = min(cpw, nc - cb);
        const bool on = lane_on && sub < nact;
        float* drow = a.dbase + cb * a.dsC + xl;                           // gradient-map row being retired
        float acc[JW][VEC];
#pragma unroll
        for (int j = 0; j < JW; ++j)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[j][e] = 0.0f;
        int rows_left = Ph;                     // pooled rows of this channel batch not yet consumed
        const float* yrec = ytab;
        // band-row major: once the pooled rows whose band starts at `base` are in, band row `base`
        // is complete for every lane of the warp
        for (int base = 0; base < R; ++base) {
            const int n = rcnt[base];
            for (int k = 0; k < n; ++k) {
                float gv[VEC];
                float2 w[JW];
                if (BULK) {
                    if (c_left == 0) {  // first row of a chunk: wait for its bytes
                        mbar_wait(bar_sa + 8 * c_slot, c_par);
                        c_left = min(kBulkRows, rows_left);
                        gp = ring + c_slot * kBulkSlotFloats + (lane_on ? sub : 0) * (kBulkRows * Pw) + pv * VEC;
                    }
                    load_yrec<JW>(yrec, w);
                    yrec += YS;
                    ld_vec<VEC>(gp, gv);
                    gp += Pw;
                    --rows_left;
                    if (--c_left == 0) {  // chunk consumed by every lane: its slot takes the next request
                        __syncwarp();
                        bulk_issue();
                        c_slot = c_slot + 1 == kBulkSlots ? 0 : c_slot + 1;
                        c_par ^= (c_slot == 0);
                    }
                } else {
                    // the prefetch lands in the slot the previous pooled row was read from
                    request(ring_sa + off_w);
                    load_yrec<JW>(yrec, w);
                    yrec += YS;
                    cp_async_wait<kRing - 1>();  // this lane's copy of the pooled row has landed
                    ld_vec<VEC>(reinterpret_cast<const float*>(reinterpret_cast<const char*>(ring) + off_r), gv);
                    off_w = off_r;
                    off_r = (off_r + kSlotB) & kRingMask;
                }
#pragma unroll
                for (int j = 0; j < JW; ++j) vfma<VEC>(acc[j], w[j], gv);
            }
            // ---- retire band row `base` -----------------------------------------------------------
            if (on) {
                if (VEC == 1) myrow[0] = acc[0][0];
                else if (VEC == 2) *reinterpret_cast<float2*>(myrow) = make_float2(acc[0][0], acc[0][1 % VEC]);
                else *reinterpret_cast<float4*>(myrow) = make_float4(acc[0][0], acc[0][1 % VEC], acc[0][2 % VEC], acc[0][3 % VEC]);
            }
            __syncwarp();
            {
                const float* up = upx;
                float* dp = drow;
                auto taps = [&](const float* u4) {
                    float2 r2 = make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const float4 u = *reinterpret_cast<const float4*>(u4 + 4 * v);
                        r2 = ffma2(make_float2(wq[4 * v], wq[4 * v + 1]), make_float2(u.x, u.y), r2);
                        r2 = ffma2(make_float2(wq[4 * v + 2], wq[4 * v + 3]), make_float2(u.z, u.w), r2);
                    }
                    return r2.x + r2.y;
                };
                if (a.split == 1) {
                    for (int s2 = 0; s2 < nact; ++s2) {
                        const float r = taps(up);
                        red_add_if(dp, r, xon);   // every column of the patch is touched: a zero test would only cost
                        up += Pws;
                        dp += a.dsC;
                    }
                } else {
                    for (int s2 = 0; s2 < nact; ++s2) {
                        float r = taps(up);
                        if (a.split >= 4) r += __shfl_down_sync(0xffffffffu, r, 16);
                        r += __shfl_down_sync(0xffffffffu, r, a.split >= 4 ? 8 : 16);
                        red_add_if(dp, r, red_on);
                        up += Pws;
                        dp += a.dsC;
                    }
                }
            }
            if (a.wide) bwd_retire_wide(rowbuf, Pws, nact, a.fw, cols, 4 * NV * a.split, a.plo, a.pcnt, a.wxT, a.TW, drow - xl, a.dsC, lane);
            __syncwarp();
#pragma unroll
            for (int j = 0; j + 1 < JW; ++j)
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc[j][e] = acc[j + 1][e];
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[JW - 1][e] = 0.0f;
            drow += a.dsH;
        }
        __syncwarp();
    }
    if (!BULK) cp_async_wait<0>();   // the trailing zero-fill requests have landed before the ring is reused
    if (BULK) {
        // every requested chunk has been consumed; the barriers' storage is reused by the next unit
        __syncwarp();
        if (lane == 0)
            for (int q = 0; q < kBulkSlots; ++q) mbar_inval(bar_sa + 8 * q);
    }
}

// ---------------------------------------------------------------------------------------------
// Backward, X-first walk for small pooled sizes.  Lanes are the patch's FEATURE COLUMNS from the
// start: a lane gathers its column's share of a pooled grad_out row with the transposed X taps
// (NT scalar shared-memory loads straight from the landed grad_out block), and accumulates the JW
// band rows that pooled row touches in registers; a completed band row leaves as one RED per lane
// and channel.  No transposition through a row buffer, no warp barrier per band row, and the lane
// already owns the address it reduces into.  Four channels per lane (channel pairs as packed FP32).
// grad_out arrives as bulk copies (cp.async.bulk, completion on an mbarrier ring): the planes of a
// 4-channel batch are contiguous, so a batch is ONE copy when it fits a slot (P <= 14), else one
// copy of 8 pooled rows per channel.
// Requires a dense NCHW grad_out bucket with 16-byte aligned planes, JY <= JW, at most NT pooled
// columns over any feature column, unit inner stride of the gradient map.
// ---------------------------------------------------------------------------------------------
struct BwdXArgs {
    const float* gbase;  // grad_out element (i, c0, 0, 0)
    float* dbase;        // gradient-map element (batch, c0, Y0, X0)
    const float* ytab;   // packed Y records (shared)
    const int* rcnt;     // pooled rows per band row (shared)
    const int* plo;      // first pooled column touching feature column x (shared)
    const int* pcnt;     // pooled columns touching feature column x (shared)
    const float* wxT;    // transposed X weights [fw][TW] (shared)
    float* slots;        // this warp's grad_out slots (shared, 16-byte aligned), a pad behind them
    unsigned bar_sa;     // shared address of this warp's first mbarrier
    unsigned* phase;     // this warp's parity bits (shared)
    int dsC, dsH;        // gradient-map strides in floats
    int Ph, Pw, R, fw, TW, nc, nslot;
    int rb;              // pooled rows per chunk (== Ph: the whole 4-channel batch is one copy)
    int slotf;           // floats per slot
    int diag;            // measurement only (DM_RA_DIAG): 16 = no reductions, 32 = no grad_out loads
    int groups;          // lane groups (1, 2, 4): narrow patches give every group its own four channels
};

template <int JW, int NT>
__device__ __forceinline__ void bwd_warp_x_body(const BwdXArgs& a) {
    constexpr int CH = 4;   // channels per lane
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Ph = a.Ph, Pw = a.Pw, R = a.R, nc = a.nc, nslot = a.nslot, rb = a.rb;
    const float* __restrict__ ytab = a.ytab;
    const int* __restrict__ rcnt = a.rcnt;
    __builtin_assume(__isShared(ytab));
    __builtin_assume(__isShared(rcnt));
    __builtin_assume(__isShared(a.slots));
    __builtin_assume(__isShared(a.plo));
    __builtin_assume(__isShared(a.pcnt));
    __builtin_assume(__isShared(a.wxT));
    __builtin_assume(__isShared(a.phase));
    // narrow patches: the warp's lanes split into G groups of LG lanes, each group a different set of
    // four channels, so a warp pass covers 4 G channels
    const int G = a.groups, LG = 32 / G;
    const int grp = lane / LG, xl = lane - grp * LG;
    const int BC = CH * G;                                 // channels per warp pass
    const int plane = Ph * Pw;
    const bool whole = rb == Ph;
    const int chs = whole ? plane : rb * Pw;               // floats between channels inside a slot
    const int nchunk = (Ph + rb - 1) / rb;
    const int step = RA_WARPS * BC;
    const int npass = (a.fw + LG - 1) / LG;                // column passes (patches wider than a group)
    const int nbatch = warp * BC < nc ? (nc - warp * BC + step - 1) / step : 0;
    const int q_total = nbatch * npass * nchunk;
    const unsigned slot_bytes = (unsigned)a.slotf * 4u;
    const unsigned slots_sa = (unsigned)__cvta_generic_to_shared(a.slots);
    unsigned par = *a.phase;
    // stale bytes behind a short chunk and the pad behind the last slot meet zero weights: finite
    for (int q = lane * 4; q < nslot * a.slotf + 32; q += 128)
        *reinterpret_cast<float4*>(a.slots + q) = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();

    // ---- producer side: the chunk stream runs over (channel batch, column pass, chunk) -------------
    int q_issue = 0, i_k = 0, i_p = 0, i_slot = 0, i_cb = warp * BC;
    auto issue = [&]() {
        if (q_issue < q_total) {
            if (lane == 0 && !(a.diag & 32)) {
                const unsigned bar = a.bar_sa + 8u * (unsigned)i_slot;
                const unsigned dst = slots_sa + (unsigned)i_slot * slot_bytes;
                const int nch = min(BC, nc - i_cb);
                const int rows = min(rb, Ph - i_k * rb);
                const float* src = a.gbase + (long long)i_cb * plane + i_k * rb * Pw;
                if (whole) {
                    const unsigned bytes = (unsigned)(nch * plane * 4);
                    mbar_expect_tx(bar, bytes);
                    bulk_g2s(dst, src, bytes, bar);
                } else {
                    const unsigned bytes = (unsigned)(rows * Pw * 4);
                    mbar_expect_tx(bar, bytes * (unsigned)nch);
                    for (int c = 0; c < nch; ++c) bulk_g2s(dst + (unsigned)(c * chs * 4), src + (long long)c * plane, bytes, bar);
                }
            }
            ++q_issue;
            if (++i_k == nchunk) {
                i_k = 0;
                if (++i_p == npass) { i_p = 0; i_cb += step; }
            }
            i_slot = i_slot + 1 == nslot ? 0 : i_slot + 1;
        }
    };
    int q_cons = 0, c_slot = 0;
    const float* chunk = a.slots;   // the chunk being read
    auto acquire = [&]() {
        if (q_cons > 0) {           // the slot just drained takes the next request
            __syncwarp();
            issue();
        }
        if (!(a.diag & 32)) {
            mbar_wait(a.bar_sa + 8u * (unsigned)c_slot, (par >> c_slot) & 1u);
            par ^= 1u << c_slot;
        }
        chunk = a.slots + c_slot * a.slotf;
        c_slot = c_slot + 1 == nslot ? 0 : c_slot + 1;
        ++q_cons;
    };
    for (int d = 0; d < nslot; ++d) issue();

    constexpr int YS = 2 * JW;
    for (int cb = warp * BC; cb < nc; cb += step) {
        const int nact = max(0, min(CH, nc - cb - CH * grp));   // live channels of this lane's group
        for (int cp = 0; cp < npass; ++cp) {
            // this lane as owner of feature column x: its taps, weights as broadcast pairs
            const int x = cp * LG + xl;
            const bool xon = x < a.fw;
            const int lo = xon ? a.plo[x] : 0, n = xon ? a.pcnt[x] : 0;
            float2 wq[NT];
#pragma unroll
            for (int q = 0; q < NT; ++q) {
                const float w = q < n ? a.wxT[x * a.TW + q] : 0.0f;
                wq[q] = make_float2(w, w);
            }
            float2 acc[JW][CH / 2];
#pragma unroll
            for (int j = 0; j < JW; ++j)
#pragma unroll
                for (int c = 0; c < CH / 2; ++c) acc[j][c] = make_float2(0.0f, 0.0f);
            // this lane's element of band row 0, first channel of its group: a global address and a
            // running float offset
            const unsigned long long dg =
                (unsigned long long)__cvta_generic_to_global(a.dbase + (long long)(cb + CH * grp) * a.dsC + x);
            int doff = 0;
            // (the caller only sends slabs of whole channel groups here: a group's four channels are all live)
            const int red_on = xon && nact > 0 && !(a.diag & 16);
            const int lane_off = lo + CH * grp * chs;   // this lane's first tap inside a chunk row, its first channel
            const float* yrec = ytab;
            int rows_left = 0;                 // pooled rows of the current chunk not yet read
            const float* gp = chunk;           // this lane's first tap of the next pooled row
            // The band rows rotate through the accumulators: in step u of a group of JW band rows the
            // logical band row j lives in acc[(j + u) % JW] (compile-time after unrolling), so a retired
            // row's accumulator is simply cleared and reused -- no register moves.
            for (int base0 = 0; base0 < R; base0 += JW) {
#pragma unroll
                for (int u = 0; u < JW; ++u) {
                    const int base = base0 + u;
                    if (base < R) {
                        const int nrow = rcnt[base];
                        for (int k = 0; k < nrow; ++k) {
                            if (rows_left == 0) {
                                acquire();
                                rows_left = rb;
                                gp = chunk + lane_off;
                            }
                            __builtin_assume(__isShared(gp));   // (LDS, not generic LD)
                            --rows_left;
                            float2 w[JW];
                            load_yrec<JW>(yrec, w);
                            yrec += YS;
                            float2 t[CH / 2];
#pragma unroll
                            for (int q = 0; q < NT; ++q) {
#pragma unroll
                                for (int c = 0; c < CH / 2; ++c) {
                                    const float2 g2 = make_float2(gp[(2 * c) * chs + q], gp[(2 * c + 1) * chs + q]);
                                    t[c] = q == 0 ? fmul2(wq[0], g2) : ffma2(wq[q], g2, t[c]);
                                }
                            }
                            gp += Pw;
#pragma unroll
                            for (int j = 0; j < JW; ++j)
#pragma unroll
                                for (int c = 0; c < CH / 2; ++c) acc[(j + u) % JW][c] = ffma2(w[j], t[c], acc[(j + u) % JW][c]);
                        }
                        // band row `base` is complete: one RED per channel from the lane that owns the pixel
                        red4_off_if(dg, doff, a.dsC, acc[u][0].x, acc[u][0].y, acc[u][1].x, acc[u][1].y, red_on);
                        doff += a.dsH;
#pragma unroll
                        for (int c = 0; c < CH / 2; ++c) acc[u][c] = make_float2(0.0f, 0.0f);
                    }
                }
            }
            // (every pooled row starts a band inside the patch, so the pass has read all its chunks)
        }
    }
    __syncwarp();
    if (lane == 0) *a.phase = par;
}

template <int VEC, int JW, int NV>
__device__ __forceinline__ void bwd_warp_core(const BwdWarpArgs& a) {
    bwd_warp_core_s<VEC, JW, NV>(a.gbase, a.dbase, a.ytab, a.rcnt, a.plo, a.pcnt, a.wxT, a.wsm, a.gsC, a.gsH, a.dsC, a.dsH, a.Ph, a.Pw, a.R, a.fw, a.TW, a.cpw, a.nc, a.split, a.wide);
}

// (By value, unlike the other three hot functions: the scalar-argument shell of this one -- 23 arguments --
// produced illegal / misaligned addresses on the device with nvcc 12.9 whatever the argument order;
// the struct goes through local memory once per warp and unit.)
template <int JW, int NT>
__device__ __noinline__ void bwd_warp_x_byval(const BwdXArgs a) {
    bwd_warp_x_body<JW, NT>(a);
}

template <int JW, int NT>
__device__ __forceinline__ void bwd_warp_x(const BwdXArgs& a) {
    bwd_warp_x_byval<JW, NT>(a);
}

template <int VEC, int JW>
__device__ __forceinline__ void bwd_warp(const BwdWarpArgs& a, int need) {
    // tap windows: 8 pooled columns per lane when that covers every feature column, else 20
    if (need <= 8 * a.split) bwd_warp_core<VEC, JW, 2>(a);
    else bwd_warp_core<VEC, JW, 5>(a);
}

// shared-memory floats a warp needs on the fast path: ring + row buffer + zero pad
__host__ __device__ constexpr int kBwdWarpFloats(int vec) { return kBwdRingFloats(vec) + kBwdRowBuf + kBwdPad + 4; }

// Column pass, generic path for bands taller than 8 rows: shared-memory reductions into a zeroed U.
template <int VEC>
__device__ DM_COLD void bwd_column_pass_generic(const BucketDesc& B, const Tables t, float* U, int i, int c0, int cs) {
    const int Pw = B.pw, Ph = B.ph;
    const int PwV = Pw / VEC;
    const int R = t.Y1 - t.Y0 + 1;
    const int per_c = Ph * PwV;
    const int items = cs * per_c;
    for (int e = threadIdx.x; e < cs * R * Pw; e += RA_THREADS) U[e] = 0.0f;
    __syncthreads();
    const float* gbase = B.ptr + (long long)i * B.sN + (long long)c0 * B.sC;
    for (int it = threadIdx.x; it < items; it += RA_THREADS) {
        const int c = it / per_c, rem = it - c * per_c;
        const int ph = rem / PwV, pv = rem - ph * PwV;
        float gv[VEC];
        ldg_stream_vec<VEC>(gbase + (long long)c * B.sC + (long long)ph * B.sH + (long long)(pv * VEC) * B.sW, gv);
        float* uc = U + c * R * Pw + pv * VEC;
        const int r0 = t.ys[ph] - t.Y0;
        for (int j = 0; j < t.JY; ++j) {
            const float w = t.wy[j * Ph + ph];
            const int r = r0 + j;
            if (w != 0.0f && r < R) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) atomicAdd(uc + r * Pw + e, w * gv[e]);
            }
        }
    }
}

// Row pass + flush: grad_patch[c][r][x] = sum_{pw in [plo[x], phi[x]]} wxT[x][pw-plo[x]] * U[c][r][pw],
// one global reduction per touched feature element.
__device__ DM_COLD void bwd_row_pass(const LevelDesc& Lv, const BucketDesc& B, const Tables t, const float* U,
                             const int* plo, const int* pcnt, const float* wxT, int TW, int batch,
                             int c0, int cs) {
    const int Pw = B.pw;
    const int fw = t.X1 - t.X0 + 1;
    const int R = t.Y1 - t.Y0 + 1;
    const int sh = pow2_shift_ge(fw);
    const int items = (cs * R) << sh;
    FastDiv fdR;
    fdR.init(R);
    float* dst = Lv.ptr + (long long)batch * Lv.sN + (long long)c0 * Lv.sC + (long long)t.Y0 * Lv.sH + (long long)t.X0 * Lv.sW;
    for (int vi = threadIdx.x; vi < items; vi += RA_THREADS) {
        const int row = vi >> sh, x = vi & ((1 << sh) - 1);
        if (x >= fw) continue;
        const float* up = U + row * Pw + plo[x];
        const float* wp = wxT + x * TW;
        const int n = pcnt[x];
        float acc = 0.0f;
        for (int q = 0; q < n; ++q) acc += wp[q] * up[q];
        if (acc != 0.0f) {
            const int c = fdR.div(row), r = row - c * R;
            red_add(dst + (long long)c * Lv.sC + (long long)r * Lv.sH + (long long)x * Lv.sW, acc);
        }
    }
}

template <int VEC>
__device__ void bwd_unit(const RaParams& p, const TmaShared ts, const Unit& un, const int* s_seg, float* smem,
                         int* stat) {
    const BucketDesc& B = p.bk[un.b];
    const int c0 = un.slab * B.cg, c1 = min(c0 + B.cg, p.C);
    float roi[5];
    int lv;
    unit_roi(p, un, s_seg, roi, lv);
    const int batch = (int)roi[0];
    if (lv < 0 || lv >= p.L) return;
    const LevelDesc& Lv = p.lv[lv];
    if (batch < 0 || batch >= Lv.N) return;
    const RoiGeom g = p.mode ? point_geom(roi, Lv.scale, p.aligned) : roi_geom(roi, Lv.scale, B.ph, B.pw, p.sampling_ratio, p.aligned);
    if (g.mode && !(g.bw >= 0.0f && g.bh >= 0.0f)) {
        direct_unit<true>(Lv, B, g, batch, un.i, c0, c1);
        return;
    }
    Tables t;
    bool fits;
    if (!build_tables(g, B.ph, B.pw, Lv.H, Lv.W, smem, p.smem_floats, stat, t, fits)) return;
    const int fw = fits ? t.X1 - t.X0 + 1 : 0;
    const int R = fits ? t.Y1 - t.Y0 + 1 : 0;
    if (p.mode && fits && t.JX <= 2 && t.JY <= 2 && fw > 32) {
        // SimpleRoIAlign, patch wider than a warp: points at least ~a pixel apart, nothing to aggregate
        point_sparse_unit<true>(Lv, B, t, batch, un.i, c0, c1);
        return;
    }
    // transposed X tables: for feature column x the pooled columns [plo, plo+pcnt) whose band covers
    // it, and their weights wxT[x][q]
    int* plo = reinterpret_cast<int*>(smem + (fits ? t.floats : 0));
    int* pcnt = plo + fw;
    int* s_tw = stat + ST_TW;
    if (fits && 2 * fw <= p.smem_floats - t.floats) {
        if (threadIdx.x == 0) { *s_tw = 0; stat[ST_NEED] = 0; }
        __syncthreads();
        for (int x = threadIdx.x; x < fw; x += RA_THREADS) {
            const int xa = x + t.X0;
            // xs is monotone: first bin whose band reaches xa, last bin whose band starts at or before xa
            int lo = 0, hi = B.pw;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (t.xs[mid] + t.JX - 1 < xa) lo = mid + 1; else hi = mid;
            }
            int a = lo;
            lo = 0; hi = B.pw;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (t.xs[mid] <= xa) lo = mid + 1; else hi = mid;
            }
            const int n = max(lo - a, 0);
            plo[x] = a;
            pcnt[x] = n;
            atomicMax(s_tw, n);
            atomicMax(&stat[ST_NEED], n + (a & 3));
        }
        __syncthreads();
    }
    const int TW = fits ? (*s_tw | 1) : 0;  // odd stride: conflict-free column-wise reads
    const int extra = (2 * fw + fw * TW + 3) & ~3;
    const long long avail = (long long)p.smem_floats - (fits ? t.floats : 0) - extra;
    if (!fits || avail < (long long)RA_WARPS * kBwdWarpFloats(VEC)) {
        direct_unit<true>(Lv, B, g, batch, un.i, c0, c1);
        return;
    }
    float* wxT = reinterpret_cast<float*>(pcnt + fw);
    for (int e = threadIdx.x; e < fw * TW; e += RA_THREADS) {
        const int x = e / TW, q = e - x * TW;
        float w = 0.0f;
        if (q < pcnt[x]) {
            const int pw = plo[x] + q;
            w = t.wx[(x + t.X0 - t.xs[pw]) * B.pw + pw];
        }
        wxT[e] = w;
    }
    const int wc = (t.JYa == 2 || t.JYa == 4 || t.JYa == 8) ? t.JYa : 0;  // rows the Y table really has
    if (p.diag & 64) { __syncthreads(); return; }   // measurement only: tables and unit bookkeeping alone
    if (wc && p.bwd_x && *s_tw <= 8 && B.sW == 1 && B.sH == B.pw && B.sC == (long long)B.ph * B.pw && Lv.sW == 1 &&
        ((c1 - c0) & 3) == 0 && (c0 & 3) == 0 && Lv.sC < (1 << 24) && Lv.sH < (1 << 24)) {
        // X-first walk (small pooled sizes): lanes own feature columns, grad_out lands by bulk copies
        const float* gbase = B.ptr + (long long)un.i * B.sN + (long long)c0 * B.sC;
        const int plane = B.ph * B.pw;
        const int per_warp = (int)((avail / RA_WARPS) & ~3ll);   // floats
        // lane groups for narrow patches (each group four channels), as far as two slots fit
        int G = fw <= 8 ? 4 : (fw <= 16 ? 2 : 1);
        G = G > p.bwd_groups ? p.bwd_groups : G;
        while (G > 1 && (((c1 - c0) % (4 * G)) != 0)) G >>= 1;
        int rb = 0, slotf = 0, nslot = 0;
        for (; G >= 1 && nslot < 2; G >>= 1) {
            const int cap = (per_warp - 32) / 2;   // floats a slot may take
            rb = 0;
            if (4 * G * plane <= cap) { rb = B.ph; slotf = 4 * G * plane; }
            else if ((B.pw & 3) == 0 && B.pw <= 32) {
                for (int r = 8; r >= 2 && !rb; r >>= 1)
                    if (4 * G * r * B.pw <= cap && r < B.ph) { rb = r; slotf = 4 * G * r * B.pw; }
            }
            nslot = rb ? (per_warp - 32) / slotf : 0;
            nslot = nslot > 3 ? 3 : nslot;
            if (nslot >= 2) break;
        }
        if (rb && nslot >= 2 && (reinterpret_cast<uintptr_t>(gbase) & 15u) == 0) {
            __syncthreads();
            const int warp = threadIdx.x >> 5;
            BwdXArgs a;
            a.gbase = gbase;
            a.dbase = Lv.ptr + (long long)batch * Lv.sN + (long long)c0 * Lv.sC + (long long)t.Y0 * Lv.sH + t.X0;
            a.ytab = t.ytab; a.rcnt = t.rcnt; a.plo = plo; a.pcnt = pcnt; a.wxT = wxT;
            a.slots = smem + t.floats + extra + (size_t)warp * per_warp;
            a.bar_sa = (unsigned)__cvta_generic_to_shared(ts.bars + warp * kTmaMaxSlots);
            a.phase = ts.phase + warp;
            a.dsC = (int)Lv.sC; a.dsH = (int)Lv.sH;
            a.Ph = B.ph; a.Pw = B.pw; a.R = R; a.fw = fw; a.TW = TW; a.nc = c1 - c0; a.nslot = nslot;
            a.rb = rb; a.slotf = slotf; a.diag = p.diag; a.groups = G;
            const bool nt4 = *s_tw <= 4;
            if (wc == 2) { if (nt4) bwd_warp_x<2, 4>(a); else bwd_warp_x<2, 8>(a); }
            else if (wc == 4) { if (nt4) bwd_warp_x<4, 4>(a); else bwd_warp_x<4, 8>(a); }
            else { if (nt4) bwd_warp_x<8, 4>(a); else bwd_warp_x<8, 8>(a); }
            return;
        }
    }
    if (p.diag & 128) { __syncthreads(); return; }   // measurement only: units off the X-first walk are skipped
    {
        const int PwV = B.pw / VEC;
        if (wc && PwV <= 32 && B.sW == 1 && Lv.sW == 1 && Lv.sC < (1 << 24) && Lv.sH < (1 << 24) &&
            B.sC < (1 << 24) && B.sH < (1 << 24) && (long long)(c1 - c0 + RA_WARPS * 32) * B.sC < (1ll << 31)) {
            // fast path: warp-private ring + row buffer, no CTA-wide barrier after this point
            __syncthreads();
            BwdWarpArgs a;
            a.gbase = B.ptr + (long long)un.i * B.sN + (long long)c0 * B.sC;
            a.dbase = Lv.ptr + (long long)batch * Lv.sN + (long long)c0 * Lv.sC + (long long)t.Y0 * Lv.sH + t.X0;
            a.ytab = t.ytab; a.rcnt = t.rcnt; a.plo = plo; a.pcnt = pcnt; a.wxT = wxT;
            a.wsm = smem + t.floats + extra + (threadIdx.x >> 5) * kBwdWarpFloats(VEC);
            a.gsC = (int)B.sC; a.gsH = (int)B.sH; a.dsC = (int)Lv.sC; a.dsH = (int)Lv.sH;
            a.Ph = B.ph; a.Pw = B.pw; a.R = R; a.fw = fw; a.TW = TW; a.cpw = 32 / PwV; a.nc = c1 - c0;
            // lanes per feature column: split the taps of narrow patches over 2 or 4 lanes
            const int need = stat[ST_NEED];
            a.split = 1;
            if (need > 20 && fw <= 16) a.split = 2;
            if (need > 40 && fw <= 8) a.split = 4;
            const int cover = (need <= 8 * a.split ? 8 : 20) * a.split;
            a.wide = (fw > 32 / a.split || need > cover) ? 1 : 0;
            if (wc == 2) bwd_warp<VEC, 2>(a, need);
            else if (wc == 4) bwd_warp<VEC, 4>(a, need);
            else bwd_warp<VEC, 8>(a, need);
            return;
        }
    }
    float* U = smem + t.floats + extra;
    const int per_c = R * B.pw;
    if (per_c > avail) {
        __syncthreads();
        direct_unit<true>(Lv, B, g, batch, un.i, c0, c1);
        return;
    }
    const int cs_max = min(c1 - c0, (int)(avail / per_c));
    __syncthreads();
    for (int c = c0; c < c1; c += cs_max) {
        const int cs = min(cs_max, c1 - c);
        bwd_column_pass_generic<VEC>(B, t, U, un.i, c, cs);
        __syncthreads();
        bwd_row_pass(Lv, B, t, U, plo, pcnt, wxT, TW, batch, c, cs);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent kernel
// ---------------------------------------------------------------------------------------------
template <bool BWD, bool DYN>
__global__ void __maxnreg__(BWD ? DM_BWD_REGS : DM_FWD_REGS) ra_kernel(const __grid_constant__ RaParams p,
                                                                          const __grid_constant__ TmaMaps tm) {
    extern __shared__ __align__(128) float smem[];
    __shared__ int s_seg[DM_MAX_BUCKETS + 1];
    __shared__ int s_stat[ST_N];
    __shared__ __align__(8) unsigned long long s_bars[32 * kTmaMaxSlots];
    __shared__ unsigned s_phase[32];
    TmaShared ts;
    ts.bars = s_bars;
    ts.phase = s_phase;
    if (threadIdx.x <= p.nb) s_seg[threadIdx.x] = p.seg ? p.seg[threadIdx.x] : (threadIdx.x == 0 ? 0 : p.K);
    {
        // one mbarrier per (warp, chunk slot), initialised once per CTA; parities carry over from unit to unit
        if (threadIdx.x < RA_WARPS * kTmaMaxSlots)
            mbar_init((unsigned)__cvta_generic_to_shared(s_bars + threadIdx.x), 1);
        if (threadIdx.x < 32) s_phase[threadIdx.x] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (DYN) {
        // Dynamic scheduling.  The per-unit work differs by more than an order of magnitude (a RoI's
        // patch is 0.1 ... 1 MB whatever it is pooled to), and with static ownership the SMs were
        // active for only 88 % (forward) / 92 % (backward) of the launch (ncu sm__cycles_active
        // avg / elapsed).  A CTA takes its next unit from the bucket that is globally least advanced
        // (same fractional pace for every bucket, as in the static interleaved walk; holding the smaller
        // buckets back so that the launch ends on small units was measured and does not pay).
        __shared__ int s_next[2];
        __shared__ float s_roi[5];   // the next unit's RoI record and level, loaded by the scheduling thread
        __shared__ int s_lv;
        // the scheduling thread is lane 0 of the LAST warp: the first two warps build the unit's tables,
        // and must not start late because one of their lanes waits for a ticket to come back from L2
        const bool sched_thread = threadIdx.x == (unsigned)(RA_THREADS - 32);
        // thread 0's view: units this CTA has taken per bucket, buckets found exhausted.  Pacing by the
        // CTA's own counts (every CTA ends up with ~1/grid of each bucket) costs one L2 round trip per
        // unit -- the ticket -- instead of two (reading the global counters first).
        __shared__ int taken[DM_MAX_BUCKETS];   // touched by thread 0 only (kept out of its registers)
        __shared__ unsigned gone;
        if (sched_thread) {
            gone = 0u;
            for (int j = 0; j < DM_MAX_BUCKETS; ++j) taken[j] = 0;
        }
        auto fetch = [&]() {   // the scheduling thread
            for (;;) {
                int jsel = -1;
                float best = 0.0f;
                for (int j = 0; j < p.nb; ++j) {
                    if ((gone >> j) & 1u) continue;
                    const int b = p.order[j];
                    const unsigned n = (unsigned)(s_seg[b + 1] - s_seg[b]) * (unsigned)p.bk[b].nslab;
                    if (n == 0u) { gone |= 1u << j; continue; }
                    const float key = p.interleave ? ((float)taken[j] + 0.5f) / (float)n * (1.0f + p.bias * (float)j) : (float)j;
                    if (jsel < 0 || key < best) { jsel = j; best = key; }
                }
                if (jsel < 0) { s_next[0] = -1; return; }
                const int b = p.order[jsel];
                const unsigned n = (unsigned)(s_seg[b + 1] - s_seg[b]) * (unsigned)p.bk[b].nslab;
                const unsigned t = atomicAdd(p.tickets + jsel, 1u);
                if (t < n) {
                    ++taken[jsel];
                    s_next[0] = jsel;
                    s_next[1] = (int)t;
                    const int pos = s_seg[b] + (int)(t / (unsigned)p.bk[b].nslab);
                    const int k = p.perm ? p.perm[pos] : pos;
                    const float* roi = p.rois + 5 * (size_t)k;
#pragma unroll
                    for (int e = 0; e < 5; ++e) s_roi[e] = roi[e];
                    s_lv = p.lvl ? p.lvl[k] : 0;
                    return;
                }
                gone |= 1u << jsel;
            }
        };
        if (sched_thread) fetch();
        __syncthreads();
        for (;;) {
            const int jsel = s_next[0];
            const unsigned u = (unsigned)s_next[1];
            Unit un;
            un.have = 1;
            un.lv = s_lv;
#pragma unroll
            for (int e = 0; e < 5; ++e) un.r[e] = s_roi[e];
            __syncthreads();   // everyone holds the unit: the scheduling thread may fetch the next one
            if (jsel < 0) break;
            if (sched_thread) fetch();
            un.b = p.order[jsel];
            // slab-major inside a RoI so concurrent CTAs share one RoI's patch in L2
            un.i = (int)(u / (unsigned)p.bk[un.b].nslab);
            un.slab = (int)(u - (unsigned)un.i * (unsigned)p.bk[un.b].nslab);
            const int vec = BWD ? p.bk[un.b].bvec : p.bk[un.b].vec;
            if (BWD) {
                if (vec == 4) bwd_unit<4>(p, ts, un, s_seg, smem, s_stat);
                else if (vec == 2) bwd_unit<2>(p, ts, un, s_seg, smem, s_stat);
                else bwd_unit<1>(p, ts, un, s_seg, smem, s_stat);
            } else {
                if (vec == 4) fwd_unit<4>(p, tm, ts, un, s_seg, smem, s_stat);
                else if (vec == 2) fwd_unit<2>(p, tm, ts, un, s_seg, smem, s_stat);
                else fwd_unit<1>(p, tm, ts, un, s_seg, smem, s_stat);
            }
            __syncthreads();
        }
        return;
    }
    // Ownership is a static round-robin over the units enumerated bucket by bucket (largest pooled
    // size first).  The WALK is interleaved: a CTA visits its units of every bucket at the same
    // fractional pace, so the latency-bound small-resolution units are spread over the whole launch
    // and hide under the bandwidth-bound 112x112 stream instead of piling up in the tail.
    long long first[DM_MAX_BUCKETS];   // this CTA's first unit inside bucket order[j]
    int mine[DM_MAX_BUCKETS], done[DM_MAX_BUCKETS];
    float inv_mine[DM_MAX_BUCKETS];
    {
        long long base = 0;
        for (int j = 0; j < p.nb; ++j) {
            const int b = p.order[j];
            const long long n = (long long)(s_seg[b + 1] - s_seg[b]) * p.bk[b].nslab;
            const long long G = gridDim.x;
            const long long f = (((long long)blockIdx.x - base) % G + G) % G;
            first[j] = f;
            mine[j] = f < n ? (int)((n - f + G - 1) / G) : 0;
            done[j] = 0;
            inv_mine[j] = mine[j] > 0 ? 1.0f / (float)mine[j] : 0.0f;
            base += n;
        }
    }
    for (;;) {
        int jsel = -1;
        float best = 0.0f;
        for (int j = 0; j < p.nb; ++j) {
            if (done[j] >= mine[j]) continue;
            const float key = p.interleave ? ((float)done[j] + 0.5f) * inv_mine[j] : (float)j;
            if (jsel < 0 || key < best) { jsel = j; best = key; }
        }
        if (jsel < 0) break;
        Unit un;
        un.have = 0;
        {
            const long long u = first[jsel] + (long long)done[jsel] * gridDim.x;
            ++done[jsel];
            un.b = p.order[jsel];
            // slab-major inside a RoI so consecutive CTAs share one RoI's patch in L2
            un.i = (int)(u / p.bk[un.b].nslab);
            un.slab = (int)(u - (long long)un.i * p.bk[un.b].nslab);
        }
        const int vec = BWD ? p.bk[un.b].bvec : p.bk[un.b].vec;
        if (BWD) {
            if (vec == 4) bwd_unit<4>(p, ts, un, s_seg, smem, s_stat);
            else if (vec == 2) bwd_unit<2>(p, ts, un, s_seg, smem, s_stat);
            else bwd_unit<1>(p, ts, un, s_seg, smem, s_stat);
        } else {
            if (vec == 4) fwd_unit<4>(p, tm, ts, un, s_seg, smem, s_stat);
            else if (vec == 2) fwd_unit<2>(p, tm, ts, un, s_seg, smem, s_stat);
            else fwd_unit<1>(p, tm, ts, un, s_seg, smem, s_stat);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// Experiment knobs, read from the environment ONCE per process (they select between measured
// variants; the defaults are the shipped configuration).
struct RaConfig {
    int want, cg, interleave, fwd_smem_kb, bwd_smem_kb, fwd_dynamic, bwd_dynamic, bias;
    int fwd_threads, tma, tma_rowmajor, tma_slots, tma_l2, diag, bwd_x, bwd_groups, l2_prefetch;
};
static const RaConfig& config() {
    static const RaConfig c = [] {
        RaConfig r;
        r.want = env_int("DM_RA_WANT", 0);   // work units per SM wanted from a single-bucket launch; 0: by pooled size
        r.cg = env_int("DM_RA_CG", 0);
        r.interleave = env_int("DM_RA_INTERLEAVE", 1);
        r.fwd_smem_kb = env_int("DM_RA_FWD_SMEM_KB", 100);
        r.bwd_smem_kb = env_int("DM_RA_BWD_SMEM_KB", DM_BWD_SMEM_KB);
        r.fwd_dynamic = env_int("DM_RA_FWD_DYNAMIC", 1);
        r.bwd_dynamic = env_int("DM_RA_BWD_DYNAMIC", DM_BWD_DYNAMIC);
        r.bias = env_int("DM_RA_BIAS", 0);
        r.fwd_threads = env_int("DM_RA_FWD_THREADS", kFwdThreads);
        r.tma = env_int("DM_RA_TMA", 1);
        r.tma_rowmajor = env_int("DM_RA_TMA_ROWMAJOR", 1);
        r.tma_slots = env_int("DM_RA_TMA_SLOTS", 3);
        r.tma_l2 = env_int("DM_RA_TMA_L2", 0);
        r.diag = env_int("DM_RA_DIAG", 0);
        r.bwd_x = env_int("DM_RA_BWD_X", 1);
        r.bwd_groups = env_int("DM_RA_BWD_GROUPS", 4);
        // off: measured slower (kernel-only: 14x14 352 -> 396 us, 28x28 557 -> 650, 7x7 x 1024 RoIs 202 -> 218);
        // the walk waits on the TMA unit's turnaround, not on DRAM, and the prefetches queue in front of its loads
        r.l2_prefetch = env_int("DM_RA_L2_PREFETCH", 0);
        return r;
    }();
    return c;
}

// ---- tensor maps ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static const EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// The maps of one level: pure host arithmetic, ~0.05 us per map (tools/tma_probe.cu), so they are
// simply encoded per launch -- no cache, no state.  Returns the bit mask of the usable width classes.
static unsigned level_maps(const LevelDesc& d, int set, int rowmajor, int l2, CUtensorMap* out) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return 0u;
    // TMA needs unit inner stride, a 16-byte aligned base, strides that are multiples of 16 bytes,
    // and (here) images stacked at C * sC so that (image, channel) is one tensor dimension
    if (d.sW != 1 || (reinterpret_cast<uintptr_t>(d.ptr) & 15u) || (d.sH & 3) || (d.sC & 3) ||
        d.sN != (long long)d.C * d.sC || d.sH < d.W || d.sC < 1)
        return 0u;
    unsigned mask = 0u;
    const cuuint64_t nc = (cuuint64_t)d.N * (cuuint64_t)d.C;
    const cuuint32_t bc = 4u << set;
    for (int c = 0; c < kNumBW; ++c) {
        cuuint64_t dims[3];
        cuuint64_t strides[2];
        cuuint32_t box[3];
        const cuuint32_t ones[3] = {1, 1, 1};
        dims[0] = (cuuint64_t)d.W;
        box[0] = (cuuint32_t)tma_bw(c);
        if (rowmajor) {   // (x, channel, y): a box lands as [row][channel][BW]
            dims[1] = nc; dims[2] = (cuuint64_t)d.H;
            strides[0] = (cuuint64_t)d.sC * 4; strides[1] = (cuuint64_t)d.sH * 4;
            box[1] = bc; box[2] = (cuuint32_t)tma_bh(set, c);
        } else {          // (x, y, channel): [channel][row][BW]
            dims[1] = (cuuint64_t)d.H; dims[2] = nc;
            strides[0] = (cuuint64_t)d.sH * 4; strides[1] = (cuuint64_t)d.sC * 4;
            box[1] = (cuuint32_t)tma_bh(set, c); box[2] = bc;
        }
        const CUtensorMapL2promotion prom = l2 == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                          : l2 == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                          : l2 == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                    : CU_TENSOR_MAP_L2_PROMOTION_NONE;
        const CUresult r = enc(&out[c], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(d.ptr), dims, strides,
                               box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, prom,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) mask |= 1u << c;
        else memset(&out[c], 0, sizeof(CUtensorMap));
    }
    return mask;
}

static int fill_params(RaParams& p, float* const* feat_ptrs, const int32_t* feat_shapes,
                       const int64_t* feat_strides, const float* spatial_scales, int L,
                       const float* rois, int K, const int32_t* lvl, const int32_t* perm,
                       const int32_t* seg, int nb, const int32_t* out_hw, float* const* out_ptrs,
                       const int64_t* out_strides, int sampling_ratio, int aligned, bool fwd_tma) {
    if (L < 1 || L > DM_MAX_LEVELS || nb < 1 || nb > DM_MAX_BUCKETS || K < 0) return DM_EINVAL;
    if (!feat_ptrs || !feat_shapes || !feat_strides || !spatial_scales || !out_hw || !out_ptrs || !out_strides)
        return DM_EINVAL;
    if ((perm == nullptr) != (seg == nullptr)) return DM_EINVAL;
    if (nb > 1 && !seg) return DM_EINVAL;
    if (K > 0 && !rois) return DM_EINVAL;
    if (L > 1 && !lvl) return DM_EINVAL;
    if (sampling_ratio < 0) return DM_EINVAL;
    p.L = L;
    p.nb = nb;
    p.K = K;
    p.C = feat_shapes[1];
    for (int l = 0; l < L; ++l) {
        LevelDesc& d = p.lv[l];
        d.ptr = feat_ptrs[l];
        d.N = feat_shapes[4 * l + 0];
        d.C = feat_shapes[4 * l + 1];
        d.H = feat_shapes[4 * l + 2];
        d.W = feat_shapes[4 * l + 3];
        d.sN = feat_strides[4 * l + 0];
        d.sC = feat_strides[4 * l + 1];
        d.sH = feat_strides[4 * l + 2];
        d.sW = feat_strides[4 * l + 3];
        d.scale = spatial_scales[l];
        if (!d.ptr || d.N < 1 || d.C != p.C || d.H < 1 || d.W < 1) return DM_EINVAL;
        d.cw = 1;
        if (d.sW == 1) {
            const uintptr_t a = reinterpret_cast<uintptr_t>(d.ptr);
            if (d.W % 4 == 0 && d.sH % 4 == 0 && d.sC % 4 == 0 && d.sN % 4 == 0 && a % 16 == 0) d.cw = 4;
            else if (d.W % 2 == 0 && d.sH % 2 == 0 && d.sC % 2 == 0 && d.sN % 2 == 0 && a % 8 == 0) d.cw = 2;
        }
    }
    if (p.C < 1) return DM_EINVAL;
    for (int b = 0; b < nb; ++b) {
        BucketDesc& d = p.bk[b];
        d.ptr = out_ptrs[b];
        d.ph = out_hw[2 * b];
        d.pw = out_hw[2 * b + 1];
        d.sN = out_strides[4 * b + 0];
        d.sC = out_strides[4 * b + 1];
        d.sH = out_strides[4 * b + 2];
        d.sW = out_strides[4 * b + 3];
        if (d.ph < 1 || d.pw < 1 || d.ph >= (1 << 15) || d.pw >= (1 << 15)) return DM_EINVAL;
        // a few MB of pooled output per work unit: the banded tables are rebuilt per unit and the
        // warps of a CTA only re-synchronise at unit boundaries
        int cg = 1572864 / (d.ph * d.pw);
        int pw2 = 1;
        while (pw2 * 2 <= cg) pw2 *= 2;
        cg = cg < 1 ? 1 : pw2;
        cg = cg < 16 ? 16 : (cg > 256 ? 256 : cg);
        if (nb == 1 && K > 0) {
            // a single small bucket (one extractor call of a training step, ~100 detections at
            // inference): the RoIs' patches differ by orders of magnitude in area, so the biggest
            // units set the launch time -- split channels until the persistent grid has ~24 units per
            // SM, but never below what one pass of a CTA's warps covers (8 warps x channels per warp).
            // Measured on the C3 extractor calls (tools/c3_breakdown.py): 7x7 x 1024 RoIs bwd 0.89 ->
            // 0.54 ms, 56x56 single-level x 256 RoIs fwd 1.33 -> 0.78 ms.
            const int vec_guess = (d.pw % 4 == 0) ? 4 : ((d.pw % 2 == 0) ? 2 : 1);
            const int pwv_guess = d.pw / vec_guess;
            const int cpw_guess = pwv_guess >= 32 ? 1 : 32 / pwv_guess;
            // (on the TMA path a warp pass covers 4 * tma_ch channels)
            const int cg_min = 8 * (pwv_guess <= 8 && fwd_tma ? 4 * tma_ch(vec_guess, 4) : (cpw_guess > 4 ? 4 : cpw_guess));
            // the BACKWARD of pooled sizes up to 14x14 prefers whole-RoI units (a unit's tables and its cold ring start
            // weigh more than the balance; kernel-only, 6 vs 24 units per SM: 14x14 x 256 RoIs bwd 139 -> 109 us,
            // 7x7 x 1024 RoIs bwd 300 -> 286); the forward does not (14x14 x 2132 RoIs 338 -> 366 us), and 56x56 x 256
            // RoIs needs the fine units in both directions (fwd 641 -> 1354 us).  (fwd_tma is false for the backward.)
            const int want_sm = config().want > 0 ? config().want : (!fwd_tma && d.ph * d.pw <= 196 ? 6 : 24);
            const long long want = (long long)want_sm * sm_count();
            int small = 256;
            while (small > cg_min && (long long)K * ((p.C + small - 1) / small) < want) small >>= 1;
            if (small < cg) cg = small;
        }
        cg = config().cg > 0 ? config().cg : cg;
        if (cg > p.C) cg = p.C;
        d.cg = cg;
        d.nslab = (p.C + cg - 1) / cg;
        const uintptr_t a = reinterpret_cast<uintptr_t>(d.ptr);
        int vec = 1;
        if (d.sW == 1) {
            if (d.pw % 4 == 0 && d.sH % 4 == 0 && d.sC % 4 == 0 && d.sN % 4 == 0 && a % 16 == 0) vec = 4;
            else if (d.pw % 2 == 0 && d.sH % 2 == 0 && d.sC % 2 == 0 && d.sN % 2 == 0 && a % 8 == 0) vec = 2;
        }
        d.vec = vec;
        const bool dense = d.sW == 1 && d.sH == d.pw && (d.ph * d.pw) % 4 == 0 && d.sC % 4 == 0 &&
                           d.sN % 4 == 0 && a % 16 == 0;
        d.bvec = dense ? vec : 1;
    }
    // largest outputs first
    for (int b = 0; b < nb; ++b) p.order[b] = b;
    for (int a = 0; a < nb; ++a)
        for (int b = a + 1; b < nb; ++b)
            if (p.bk[p.order[b]].ph * p.bk[p.order[b]].pw > p.bk[p.order[a]].ph * p.bk[p.order[a]].pw) {
                const int tmp = p.order[a];
                p.order[a] = p.order[b];
                p.order[b] = tmp;
            }
    p.rois = rois;
    p.lvl = lvl;
    p.perm = perm;
    p.seg = seg;
    p.sampling_ratio = sampling_ratio;
    p.aligned = aligned ? 1 : 0;
    p.mode = 0;
    p.interleave = config().interleave;
    return DM_OK;
}

// occupancy and the dynamic shared-memory opt-in of one kernel variant, resolved once per device
struct KernelSetup {
    std::atomic<int> grid{0};
};
template <bool BWD>
static int kernel_grid(bool dyn, int threads, int smem_bytes, const char* where, int& grid) {
    static KernelSetup setup[64][2];
    static std::mutex mu;
    int dev = 0;
    DM_CUDA_CHECK(cudaGetDevice(&dev), where);
    if (dev < 0 || dev >= 64) return DM_EUNSUPPORTED;
    grid = setup[dev][dyn].grid.load(std::memory_order_acquire);
    if (grid > 0) return DM_OK;
    std::lock_guard<std::mutex> lock(mu);
    auto kern = dyn ? ra_kernel<BWD, true> : ra_kernel<BWD, false>;
    DM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes), where);
    int occ = 0;
    DM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem_bytes), where);
    if (occ < 1) return DM_EUNSUPPORTED;
    grid = sm_count() * occ;
    setup[dev][dyn].grid.store(grid, std::memory_order_release);
    return DM_OK;
}

// `sched` = caller-owned device scratch of DM_SCHED_SCRATCH_BYTES (work-ticket counters, zeroed here
// on the caller's stream) or NULL: static round-robin ownership of the work units.
template <bool BWD>
static int launch(RaParams& p, cudaStream_t st, unsigned* sched, const char* where) {
    const RaConfig& cf = config();
    const int smem_bytes = (BWD ? cf.bwd_smem_kb : cf.fwd_smem_kb) * 1024;
    const int threads = BWD ? kBwdThreads : cf.fwd_threads;
    p.smem_floats = smem_bytes / 4;
    // static and dynamic scheduling are separate instantiations (each with its own register allocation)
    bool dyn = sched != nullptr && (BWD ? cf.bwd_dynamic : cf.fwd_dynamic) != 0;
    if (BWD && cf.bwd_dynamic == 1 && p.nb == 1 && p.bk[0].ph * p.bk[0].pw <= 100) dyn = false;
    int grid = 0;
    const int rc = kernel_grid<BWD>(dyn, threads, smem_bytes, where, grid);
    if (rc != DM_OK) return rc;
    p.tickets = nullptr;
    p.bias = 0.001f * (float)cf.bias;
    if (dyn) {
        p.tickets = sched;
        DM_CUDA_CHECK(cudaMemsetAsync(p.tickets, 0, sizeof(unsigned) * DM_MAX_BUCKETS, st), where);
    }
    // tensor maps of the levels (forward patch loads)
    static_assert(sizeof(TmaMaps) + sizeof(RaParams) <= 32000, "kernel parameter space");
    TmaMaps tm;
    p.tma_rowmajor = cf.tma_rowmajor ? 1 : 0;
    p.diag = cf.diag;
    p.l2_prefetch = (!BWD && cf.tma && cf.l2_prefetch) ? 1 : 0;
    p.bwd_x = (cf.bwd_x && p.mode == 0) ? 1 : 0;
    p.bwd_groups = cf.bwd_groups >= 4 ? 4 : (cf.bwd_groups >= 2 ? 2 : 1);
    p.tma_slots = cf.tma_slots < 2 ? 2 : (cf.tma_slots > kTmaMaxSlots ? kTmaMaxSlots : cf.tma_slots);
    for (int l = 0; l < kTmaLevels; ++l)
        for (int s2 = 0; s2 < kTmaSets; ++s2) p.tma_mask[l][s2] = 0u;
    if (!BWD && cf.tma && p.mode == 0) {
        // only the channel sets this launch's buckets can use (buckets with <= 8 strips per pooled row)
        for (int b = 0; b < p.nb; ++b) {
            const BucketDesc& d = p.bk[b];
            if (d.pw / d.vec > 8 || d.sW != 1) continue;
            for (int jw = 4; jw <= 8; jw += 4) {
                const int bc = 4 * tma_ch(d.vec, jw);
                const int set = bc == 4 ? 0 : (bc == 8 ? 1 : 2);
                for (int l = 0; l < p.L && l < kTmaLevels; ++l)
                    if (!p.tma_mask[l][set]) p.tma_mask[l][set] = level_maps(p.lv[l], set, p.tma_rowmajor, cf.tma_l2, tm.m[l][set]);
            }
        }
    }
    auto kern = dyn ? ra_kernel<BWD, true> : ra_kernel<BWD, false>;
    kern<<<grid, threads, smem_bytes, st>>>(p, tm);
    DM_LAUNCH_CHECK(where);
    return DM_OK;
}

}  // namespace dm

extern "C" int dm_roi_align_fwd(const float* const* feat_ptrs, const int32_t* feat_shapes,
                                const int64_t* feat_strides, const float* spatial_scales,
                                int num_levels, const float* rois, int K, const int32_t* lvl,
                                const int32_t* perm, const int32_t* seg_offsets, int num_buckets,
                                const int32_t* out_hw, float* const* out_ptrs,
                                const int64_t* out_strides, int sampling_ratio, int aligned,
                                void* sched_scratch, dm_stream_t stream) {
    dm::RaParams p;
    const int rc = dm::fill_params(p, const_cast<float* const*>(feat_ptrs), feat_shapes, feat_strides,
                                   spatial_scales, num_levels, rois, K, lvl, perm, seg_offsets,
                                   num_buckets, out_hw, out_ptrs, out_strides, sampling_ratio, aligned,
                                   dm::config().tma != 0);
    if (rc != DM_OK) return rc;
    if (K == 0) return DM_OK;
    for (int b = 0; b < num_buckets; ++b)
        if (!out_ptrs[b] && !seg_offsets) return DM_EINVAL;
    return dm::launch<false>(p, (cudaStream_t)stream, static_cast<unsigned*>(sched_scratch), "dm_roi_align_fwd");
}

extern "C" int dm_roi_align_bwd(float* const* grad_feat_ptrs, const int32_t* feat_shapes,
                                const int64_t* feat_strides, const float* spatial_scales,
                                int num_levels, const float* rois, int K, const int32_t* lvl,
                                const int32_t* perm, const int32_t* seg_offsets, int num_buckets,
                                const int32_t* out_hw, const float* const* grad_out_ptrs,
                                const int64_t* grad_out_strides, int sampling_ratio, int aligned,
                                int zero_init, void* sched_scratch, dm_stream_t stream) {
    dm::RaParams p;
    const int rc = dm::fill_params(p, grad_feat_ptrs, feat_shapes, feat_strides, spatial_scales,
                                   num_levels, rois, K, lvl, perm, seg_offsets, num_buckets, out_hw,
                                   const_cast<float* const*>(grad_out_ptrs), grad_out_strides,
                                   sampling_ratio, aligned, false);
    if (rc != DM_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (zero_init) {
        for (int l = 0; l < num_levels; ++l) {
            const dm::LevelDesc& d = p.lv[l];
            // dense map: the four strides address exactly N*C*H*W distinct elements
            const long long n = (long long)d.N * d.C * d.H * d.W;
            const long long span = (d.N - 1) * d.sN + (d.C - 1) * d.sC + (d.H - 1) * d.sH + (d.W - 1) * d.sW + 1;
            if (span != n) return DM_EINVAL;
            DM_CUDA_CHECK(cudaMemsetAsync(d.ptr, 0, sizeof(float) * (size_t)n, st), "dm_roi_align_bwd/memset");
        }
    }
    if (K == 0) return DM_OK;
    return dm::launch<true>(p, st, static_cast<unsigned*>(sched_scratch), "dm_roi_align_bwd");
}

// ---------------------------------------------------------------------------------------------
// SimpleRoIAlign: one zero-padded grid_sample point per bin.  Same separable
// banded map as RoIAlign with a one-sample grid, so it runs through the same persistent kernels
// with the tables built by point_geom / point_tap.
// ---------------------------------------------------------------------------------------------
extern "C" int dm_simple_roi_align_fwd(const float* feat, const int32_t* feat_shape,
                                       const int64_t* feat_strides, float spatial_scale,
                                       const float* rois, int K, int out_h, int out_w, float* out,
                                       const int64_t* out_strides, int aligned, void* sched_scratch,
                                       dm_stream_t stream) {
    if (!feat || !feat_shape || !feat_strides || !out_strides || (K > 0 && !out)) return DM_EINVAL;
    dm::RaParams p;
    float* fp = const_cast<float*>(feat);
    const int32_t hw[2] = {out_h, out_w};
    const int rc = dm::fill_params(p, &fp, feat_shape, feat_strides, &spatial_scale, 1, rois, K, nullptr,
                                   nullptr, nullptr, 1, hw, &out, out_strides, 0, aligned, false);
    if (rc != DM_OK) return rc;
    if (K == 0) return DM_OK;
    p.mode = 1;
    return dm::launch<false>(p, (cudaStream_t)stream, static_cast<unsigned*>(sched_scratch), "dm_simple_roi_align_fwd");
}

extern "C" int dm_simple_roi_align_bwd(float* grad_feat, const int32_t* feat_shape,
                                       const int64_t* feat_strides, float spatial_scale,
                                       const float* rois, int K, int out_h, int out_w,
                                       const float* grad_out, const int64_t* grad_out_strides,
                                       int aligned, int zero_init, void* sched_scratch, dm_stream_t stream) {
    if (!grad_feat || !feat_shape || !feat_strides || !grad_out_strides || (K > 0 && !grad_out)) return DM_EINVAL;
    dm::RaParams p;
    float* go = const_cast<float*>(grad_out);
    const int32_t hw[2] = {out_h, out_w};
    const int rc = dm::fill_params(p, &grad_feat, feat_shape, feat_strides, &spatial_scale, 1, rois, K,
                                   nullptr, nullptr, nullptr, 1, hw, &go, grad_out_strides, 0, aligned, false);
    if (rc != DM_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (zero_init) {
        const dm::LevelDesc& d = p.lv[0];
        const long long n = (long long)d.N * d.C * d.H * d.W;
        const long long span = (d.N - 1) * d.sN + (d.C - 1) * d.sC + (d.H - 1) * d.sH + (d.W - 1) * d.sW + 1;
        if (span != n) return DM_EINVAL;
        DM_CUDA_CHECK(cudaMemsetAsync(d.ptr, 0, sizeof(float) * (size_t)n, st), "dm_simple_roi_align_bwd/memset");
    }
    if (K == 0) return DM_OK;
    p.mode = 1;
    return dm::launch<true>(p, st, static_cast<unsigned*>(sched_scratch), "dm_simple_roi_align_bwd");
}
