// Stage 1: FPN level + resolution bucket per RoI, and a stable grouping of the RoIs by bucket.
//
// Replaces SingleRoIExtractor.map_roi_levels
// (mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:32-51) and the argmax over
// the mask-switch one-hot (mmdet/models/roi_heads/dynamask_roi_head.py:84-114, :197-203).
//
// This translation unit is compiled with -fmad=false: the level is an integer result of fp32
// arithmetic and must be bit-exact.  The op sequence is the one torch runs on CUDA tensors for
// the reference expression: sub, sub, mul, sqrt, mul-by-reciprocal (ATen's true-divide by a
// python scalar multiplies by 1/b on CUDA), add 1e-6, log2f, floor, clamp.
//
// One CTA does the whole job: K is a few thousand at most on this path (8192 in the microbench),
// the work is ~40 bytes per RoI, and a single CTA gives a deterministic, stable order without any
// inter-CTA protocol or host synchronisation.
#include "dm_common.cuh"

namespace dm {

constexpr int kAssignThreads = 1024;

__global__ void __launch_bounds__(kAssignThreads, 1)
assign_kernel(const float* __restrict__ rois, int K, const float* __restrict__ onehot, int nb,
              int num_levels, float recip_finest, int32_t* __restrict__ lvl_out,
              int32_t* __restrict__ bucket_out, int32_t* __restrict__ perm,
              int32_t* __restrict__ seg_offsets) {
    __shared__ int s_count[DM_MAX_BUCKETS];       // RoIs per bucket
    __shared__ int s_base[DM_MAX_BUCKETS];        // running write cursor per bucket
    __shared__ int s_warp[DM_MAX_BUCKETS][32];    // per-warp counts of the current chunk
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < DM_MAX_BUCKETS) s_count[tid] = 0;
    __syncthreads();

    // pass 1: level, bucket, histogram
    for (int k = tid; k < K; k += kAssignThreads) {
        const float* r = rois + 5 * (size_t)k;
        const float w = __fsub_rn(r[3], r[1]);
        const float h = __fsub_rn(r[4], r[2]);
        const float s = __fsqrt_rn(__fmul_rn(w, h));
        const float t = floorf(log2f(__fadd_rn(__fmul_rn(s, recip_finest), 1e-6f)));
        int lv;
        if (t != t) lv = -1;
        else lv = (int)fminf(fmaxf(t, 0.0f), (float)(num_levels - 1));
        if (lvl_out) lvl_out[k] = lv;
        int b = 0;
        if (onehot) {
            const float* o = onehot + (size_t)k * nb;
            float best = o[0];
            for (int j = 1; j < nb; ++j) {
                const float v = o[j];
                if (v > best) { best = v; b = j; }
            }
        }
        if (bucket_out) bucket_out[k] = b;
        atomicAdd(&s_count[b], 1);
    }
    __syncthreads();
    if (!perm || !seg_offsets) return;
    if (tid == 0) {
        int acc = 0;
        for (int b = 0; b < nb; ++b) {
            seg_offsets[b] = acc;
            s_base[b] = acc;
            acc += s_count[b];
        }
        seg_offsets[nb] = acc;
    }
    __syncthreads();

    // pass 2: stable scatter, one chunk of 1024 RoIs at a time, in order
    for (int k0 = 0; k0 < K; k0 += kAssignThreads) {
        const int k = k0 + tid;
        int b = -1;
        if (k < K) {
            b = 0;
            if (onehot) {
                const float* o = onehot + (size_t)k * nb;
                float best = o[0];
                for (int j = 1; j < nb; ++j) {
                    const float v = o[j];
                    if (v > best) { best = v; b = j; }
                }
            }
        }
        int my_rank = 0;  // number of earlier RoIs of this chunk's warp with the same bucket
        for (int v = 0; v < nb; ++v) {
            const unsigned m = __ballot_sync(0xffffffffu, b == v);
            if (b == v) my_rank = __popc(m & ((1u << lane) - 1u));
            if (lane == 0) s_warp[v][warp] = __popc(m);
        }
        __syncthreads();
        if (k < K) {
            int before = 0;
            for (int w2 = 0; w2 < warp; ++w2) before += s_warp[b][w2];
            perm[s_base[b] + before + my_rank] = k;
        }
        __syncthreads();
        if (tid < nb) {
            int tot = 0;
            for (int w2 = 0; w2 < 32; ++w2) tot += s_warp[tid][w2];
            s_base[tid] += tot;
        }
        __syncthreads();
    }
}

}  // namespace dm

extern "C" int dm_assign(const float* rois, int K, const float* onehot, int num_buckets,
                         int num_levels, float finest_scale, int32_t* lvl, int32_t* bucket,
                         int32_t* perm, int32_t* seg_offsets, dm_stream_t stream) {
    if (K < 0 || num_levels < 1 || num_levels > DM_MAX_LEVELS) return DM_EINVAL;
    if (num_buckets < 1 || num_buckets > DM_MAX_BUCKETS) return DM_EINVAL;
    if ((perm == nullptr) != (seg_offsets == nullptr)) return DM_EINVAL;
    if (K > 0 && rois == nullptr) return DM_EINVAL;
    if (!(finest_scale > 0.0f)) return DM_EINVAL;
    const float recip = 1.0f / finest_scale;
    dm::assign_kernel<<<1, dm::kAssignThreads, 0, (cudaStream_t)stream>>>(
        rois, K, onehot, num_buckets, num_levels, recip, lvl, bucket, perm, seg_offsets);
    DM_LAUNCH_CHECK("dm_assign");
    return DM_OK;
}
