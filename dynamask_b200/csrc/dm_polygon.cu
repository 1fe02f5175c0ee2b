// Next row: training mask targets from POLYGON ground truth.
//
// Replaces, for every positive RoI and every target size,
//   PolygonMasks.crop_and_resize   mmdet/core/mask/structures.py:465-499  (shift by the box corner,
//                                  scale by out / max(box extent, 1) -- python loops on the host)
//   PolygonMasks.to_ndarray        structures.py:541-550 -> polygon_to_bitmap :561-575
//                                  (pycocotools frPyObjects -> merge -> decode on the host)
// and the clip / float / upload around them (mask_target.py:49-58, dynamask_head.py:248-261).
//
// The rasterisation rule is pycocotools' rleFrPoly (common/maskApi.c, restated in
// oracle/dm_oracle.c): vertices are scaled by 5 and rounded, each edge is walked one up-sampled
// pixel at a time along its major axis, and wherever the up-sampled x changes onto the centre of
// an output column the pair (column x, first row y at or below the edge) is a run end of the
// column-major run-length code.  A pixel is foreground when an odd number of run ends lie at or
// before its column-major index; polygons of one object are OR-ed.
//
// One CTA per (RoI, size, band of columns).  Warps own edges, lanes own the up-sampled steps of an
// edge; run ends are counted into a shared-memory histogram over the band's column-major
// positions, a column-parallel parity scan turns it into the bitmap.  A target that fits shared
// memory is one band; a large one (PolygonMasks.to_ndarray / to_tensor rasterise at IMAGE size,
// mmdet/core/mask/structures.py:541-558) is cut into bands of columns: every band walks all edges,
// keeps the run ends that fall inside it and carries in the parity of those that fall before it.  All coordinate arithmetic is the
// reference's float32 / float64 / int sequence with FMA contraction off, so targets are bit-exact.
#include "dm_common.cuh"

namespace dm {

constexpr int kPolyThreads = 128;
constexpr int kPolyMaxSizes = 8;

struct PolyParams {
    const double* xy;         // interleaved vertices of every polygon
    const int64_t* voff;      // [P+1] vertex offsets per polygon
    const int32_t* ooff;      // [G+1] polygon offsets per object
    const int32_t* img_meta;  // [B*3] (first object, H, W) per image
    const float* boxes;       // [K,4]
    const int64_t* inds;      // [K]
    const int32_t* roi_img;   // [K] or null
    float* out[kPolyMaxSizes];
    int sh[kPolyMaxSizes], sw[kPolyMaxSizes];
    int wb[kPolyMaxSizes];    // columns per band
    int B, K, clip, n_sizes, G;
};

struct Edge {
    int xs, ys, dx, dy, flip, major_x;
    double s;
    int n;  // points on the edge
};

__device__ __forceinline__ int up5(double v) { return (int)__dadd_rn(__dmul_rn(5.0, v), 0.5); }

// vertex j of the polygon after crop + resize: (p - corner) * scale, in float64 like numpy
__device__ __forceinline__ void vertex(const double* __restrict__ p, int j, double cx, double cy, double fx,
                                       double fy, int& x, int& y) {
    x = up5(__dmul_rn(__dsub_rn(p[2 * j], cx), fx));
    y = up5(__dmul_rn(__dsub_rn(p[2 * j + 1], cy), fy));
}

__device__ __forceinline__ Edge make_edge(int xs, int ys, int xe, int ye) {
    Edge e;
    e.dx = abs(xe - xs);
    e.dy = abs(ys - ye);
    e.major_x = e.dx >= e.dy;
    e.flip = (e.major_x && xs > xe) || (!e.major_x && ys > ye);
    if (e.flip) { int t = xs; xs = xe; xe = t; t = ys; ys = ye; ye = t; }
    e.xs = xs;
    e.ys = ys;
    // a zero-length edge divides 0 by 0 in maskApi.c; its single point is the vertex either way
    e.s = e.major_x ? (e.dx ? __ddiv_rn((double)(ye - ys), (double)e.dx) : 0.0)
                    : __ddiv_rn((double)(xe - xs), (double)e.dy);
    e.n = (e.major_x ? e.dx : e.dy) + 1;
    return e;
}

__device__ __forceinline__ void edge_point(const Edge& e, int d, int& u, int& v) {
    if (e.major_x) {
        const int t = e.flip ? e.dx - d : d;
        u = t + e.xs;
        v = (int)__dadd_rn(__dadd_rn((double)e.ys, __dmul_rn(e.s, (double)t)), 0.5);
    } else {
        const int t = e.flip ? e.dy - d : d;
        v = t + e.ys;
        u = (int)__dadd_rn(__dadd_rn((double)e.xs, __dmul_rn(e.s, (double)t)), 0.5);
    }
}

// run end produced by the step (u0,v0) -> (u1,v1), or -1
__device__ __forceinline__ int crossing(int u0, int v0, int u1, int v1, int h, int w) {
    if (u1 == u0) return -1;
    double xd = (double)(u1 < u0 ? u1 : u1 - 1);
    xd = __dsub_rn(__ddiv_rn(__dadd_rn(xd, 0.5), 5.0), 0.5);
    if (floor(xd) != xd || xd < 0.0 || xd > (double)(w - 1)) return -1;
    double yd = (double)(v1 < v0 ? v1 : v0);
    yd = __dsub_rn(__ddiv_rn(__dadd_rn(yd, 0.5), 5.0), 0.5);
    if (yd < 0.0) yd = 0.0; else if (yd > (double)h) yd = (double)h;
    yd = ceil(yd);
    return (int)xd * h + (int)yd;
}

__global__ void __launch_bounds__(kPolyThreads) polygon_target_kernel(const __grid_constant__ PolyParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_carry[1024];
    __shared__ int s_cin;                                             // run ends before the band
    const int k = blockIdx.x, si = blockIdx.y;
    const int h = p.sh[si], w = p.sw[si];
    const int x0 = blockIdx.z * p.wb[si];
    if (x0 >= w) return;
    const int bw_ = min(p.wb[si], w - x0);                            // columns of this band
    const int npx = h * bw_;
    const int a0 = x0 * h;                                            // column-major position of the band's first pixel
    int* cnt = reinterpret_cast<int*>(smem_raw);                      // [npx + 1]
    uint8_t* bitmap = reinterpret_cast<uint8_t*>(cnt + ((npx + 4) & ~3));  // [npx] row-major within the band
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = kPolyThreads >> 5;

    for (int i = threadIdx.x; i < npx; i += kPolyThreads) bitmap[i] = 0;

    const int img = p.roi_img ? p.roi_img[k] : 0;
    const long long gi = p.inds[k];
    bool valid = img >= 0 && img < p.B;
    int obj = 0;
    float x1 = p.boxes[4 * k], y1 = p.boxes[4 * k + 1], x2 = p.boxes[4 * k + 2], y2 = p.boxes[4 * k + 3];
    if (valid) {
        const int base = p.img_meta[3 * img], H = p.img_meta[3 * img + 1], W = p.img_meta[3 * img + 2];
        const int next = img + 1 < p.B ? p.img_meta[3 * (img + 1)] : p.G;
        valid = gi >= 0 && base + gi < next;
        obj = base + (int)gi;
        if (p.clip) {
            x1 = fminf(fmaxf(x1, 0.0f), (float)W);
            x2 = fminf(fmaxf(x2, 0.0f), (float)W);
            y1 = fminf(fmaxf(y1, 0.0f), (float)H);
            y2 = fminf(fmaxf(y2, 0.0f), (float)H);
        }
    }
    // structures.py:483-487: w = max(x2 - x1, 1) in float32; scale = out / w in float32
    const float bw = fmaxf(__fsub_rn(x2, x1), 1.0f), bh = fmaxf(__fsub_rn(y2, y1), 1.0f);
    const double fx = (double)__fdiv_rn((float)w, bw), fy = (double)__fdiv_rn((float)h, bh);
    const double cx = (double)x1, cy = (double)y1;

    const int q0 = valid ? p.ooff[obj] : 0, q1 = valid ? p.ooff[obj + 1] : 0;
    for (int q = q0; q < q1; ++q) {
        const long long v0 = p.voff[q];
        const int kv = (int)(p.voff[q + 1] - v0);
        if (kv <= 0) continue;
        const double* __restrict__ poly = p.xy + 2 * v0;
        __syncthreads();
        for (int i = threadIdx.x; i <= npx; i += kPolyThreads) cnt[i] = 0;
        if (threadIdx.x == 0) s_cin = 0;
        __syncthreads();
        for (int j = warp; j < kv; j += nwarp) {
            int xa, ya, xb, yb;
            vertex(poly, j, cx, cy, fx, fy, xa, ya);
            vertex(poly, j + 1 == kv ? 0 : j + 1, cx, cy, fx, fy, xb, yb);
            const Edge e = make_edge(xa, ya, xb, yb);
            for (int d = lane; d < e.n; d += 32) {
                int u1, v1, u0, v0p;
                edge_point(e, d, u1, v1);
                if (d > 0) {
                    edge_point(e, d - 1, u0, v0p);
                } else {
                    if (j == 0) continue;  // the very first point has no predecessor
                    int xp, yp;
                    vertex(poly, j - 1, cx, cy, fx, fy, xp, yp);
                    const Edge pe = make_edge(xp, yp, xa, ya);
                    edge_point(pe, pe.n - 1, u0, v0p);
                }
                const int a = crossing(u0, v0p, u1, v1, h, w);
                if (a >= 0 && a < a0) atomicAdd(&s_cin, 1);
                else if (a >= a0 && a < a0 + npx) atomicAdd(&cnt[a - a0], 1);
            }
        }
        __syncthreads();
        // parity scan in column-major order: column totals, exclusive prefix, then each column
        for (int x = threadIdx.x; x < bw_; x += kPolyThreads) {
            int t = 0;
            for (int y = 0; y < h; ++y) t ^= cnt[x * h + y];
            s_carry[x] = t & 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int c = s_cin & 1;
            for (int x = 0; x < bw_; ++x) {
                const int t = s_carry[x];
                s_carry[x] = c;
                c ^= t;
            }
        }
        __syncthreads();
        for (int x = threadIdx.x; x < bw_; x += kPolyThreads) {
            int par = s_carry[x];
            for (int y = 0; y < h; ++y) {
                par ^= cnt[x * h + y] & 1;
                if (par) bitmap[y * bw_ + x] = 1;
            }
        }
    }
    __syncthreads();
    float* o = p.out[si] + (size_t)k * h * w + x0;
    for (int i = threadIdx.x; i < npx; i += kPolyThreads) {
        const int y = i / bw_, x = i - y * bw_;
        o[(size_t)y * w + x] = bitmap[i] ? 1.0f : 0.0f;
    }
}

}  // namespace dm

extern "C" int dm_polygon_target(const double* poly_xy, const int64_t* vert_offsets,
                                 const int32_t* obj_poly_offsets, int G, const int32_t* img_meta, int B,
                                 const float* boxes, const int64_t* inds, const int32_t* roi_img, int K,
                                 int clip, const int32_t* sizes_hw, int n_sizes, float* const* out_ptrs,
                                 dm_stream_t stream) {
    if (K < 0 || B < 1 || G < 0 || n_sizes < 1 || n_sizes > dm::kPolyMaxSizes || !sizes_hw || !out_ptrs)
        return DM_EINVAL;
    if (K == 0) return DM_OK;
    if (!vert_offsets || !obj_poly_offsets || !img_meta || !boxes || !inds) return DM_EINVAL;
    dm::PolyParams p;
    p.xy = poly_xy;
    p.voff = vert_offsets;
    p.ooff = obj_poly_offsets;
    p.img_meta = img_meta;
    p.boxes = boxes;
    p.inds = inds;
    p.roi_img = roi_img;
    p.B = B;
    p.K = K;
    p.G = G;
    p.clip = clip ? 1 : 0;
    p.n_sizes = n_sizes;
    // a band = as many columns as fit ~96 KB of shared memory (4 bytes of histogram + 1 of bitmap per pixel),
    // at most 1024 (the per-column carry array); one band for the usual target sizes
    constexpr int kBandPx = 19000;
    int max_px = 0, max_bands = 1;
    for (int s = 0; s < n_sizes; ++s) {
        p.sh[s] = sizes_hw[2 * s];
        p.sw[s] = sizes_hw[2 * s + 1];
        p.out[s] = out_ptrs[s];
        if (p.sh[s] < 1 || p.sw[s] < 1 || !p.out[s]) return DM_EINVAL;
        if (p.sh[s] > kBandPx) return DM_EUNSUPPORTED;   // a single column would not fit
        int wb = kBandPx / p.sh[s];
        wb = wb > 1024 ? 1024 : wb;
        wb = wb > p.sw[s] ? p.sw[s] : wb;
        p.wb[s] = wb;
        const int bands = (p.sw[s] + wb - 1) / wb;
        if (bands > max_bands) max_bands = bands;
        if (p.sh[s] * wb > max_px) max_px = p.sh[s] * wb;
    }
    if (max_bands > 65535) return DM_EUNSUPPORTED;
    const size_t smem = (size_t)((max_px + 4) & ~3) * 4 + (size_t)max_px;
    DM_CUDA_CHECK(cudaFuncSetAttribute(dm::polygon_target_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem), "dm_polygon_target");
    dm::polygon_target_kernel<<<dim3(K, n_sizes, max_bands), dm::kPolyThreads, smem, (cudaStream_t)stream>>>(p);
    DM_LAUNCH_CHECK("dm_polygon_target");
    return DM_OK;
}
