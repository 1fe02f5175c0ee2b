/*
 * dynamask_sm100.h -- C ABI of libdynamask_sm100.so, the B200 (sm_100a) implementation of the
 * DynaMask per-instance mask hot path.
 *
 * This is the drop-in boundary.  Every entry point takes plain device
 * pointers, sizes and a CUDA stream; no torch / ATen types cross it.  The library allocates no
 * memory: every buffer (inputs, outputs, scratch) belongs to the caller, every call is
 * asynchronous on `stream` and re-entrant across streams (and across CUDA-graph replays: a launch
 * keeps its scheduling counters in the caller's `sched_scratch`, not in the library).  The library
 * holds no mutable device state; on the host it keeps the thread-local error text, a launch
 * counter and per-device constants resolved once (SM count, kernel occupancy).
 *
 * Reference interfaces replaced (paths relative to the reference tree, lslrh/DynaMask):
 *   dm_assign          <- SingleRoIExtractor.map_roi_levels
 *                           mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:32-51
 *                         + argmax of the mask-switch one-hot,
 *                           mmdet/models/roi_heads/dynamask_roi_head.py:84-114 (use at :197-203)
 *   dm_roi_align_fwd   <- mmcv._ext.roi_align_forward as reached from
 *                           mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:52-54 and the
 *                           per-level select/scatter loop single_level_roi_extractor.py:53-81
 *   dm_roi_align_bwd   <- mmcv._ext.roi_align_backward (autograd of the above)
 *   dm_paste_masks     <- _do_paste_mask, mmdet/models/roi_heads/mask_heads/fcn_mask_head.py:240-308
 *                         + sigmoid / class select / threshold of DynaMaskHead.get_seg_masks,
 *                           mmdet/models/roi_heads/mask_heads/dynamask_head.py:279-342
 *   dm_mask_target     <- BitmapMasks.crop_and_resize, mmdet/core/mask/structures.py:256-286
 *                         + clip of mask_target_single, mmdet/core/mask/mask_target.py:49-51
 *                         + the four-size loop of DynaMaskHead.get_targets, dynamask_head.py:246-271
 *
 * All functions return DM_OK (0) or a negative DM_E* code and never throw.
 */
#ifndef DYNAMASK_SM100_H_
#define DYNAMASK_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DM_MAX_LEVELS 8
#define DM_MAX_BUCKETS 8
/* bytes of caller-owned device scratch one RoIAlign launch may be given (see sched_scratch) */
#define DM_SCHED_SCRATCH_BYTES 64

#define DM_OK 0
#define DM_EINVAL (-1)       /* bad argument (null pointer, size, stride) */
#define DM_ECUDA (-2)        /* a CUDA runtime call or launch failed; see dm_last_cuda_error() */
#define DM_EUNSUPPORTED (-3) /* valid request this build does not implement */

/* output modes of dm_paste_masks */
#define DM_PASTE_BOOL 0  /* uint8 {0,1}: (value >= thr) */
#define DM_PASTE_U8 1    /* uint8: (uint8)(value * 255)  (reference thr < 0 branch) */
#define DM_PASTE_F32 2   /* float: raw interpolated value (the _do_paste_mask return) */

typedef void* dm_stream_t; /* cudaStream_t */

int dm_version(void);
const char* dm_error_string(int code);
/* Text of the last CUDA error seen by the calling thread ("" if none). */
const char* dm_last_cuda_error(void);
/* Number of kernels launched by this library from the calling process so far. */
int64_t dm_launch_count(void);

/*
 * Stage 1: per-RoI FPN level + resolution bucket + stable grouping by bucket.
 *   rois        device [K,5] fp32 (batch_idx, x1, y1, x2, y2), row stride 5
 *   onehot      device [K,num_buckets] fp32 mask-switch output, or NULL (every RoI -> bucket 0)
 *   lvl         device [K] int32 out: clamp(floor(log2(sqrt(w*h)/finest_scale + 1e-6)), 0, L-1);
 *               -1 when the expression is NaN (negative area: the reference matches no level)
 *   bucket      device [K] int32 out: first argmax of the one-hot row
 *   perm        device [K] int32 out: RoI indices grouped by bucket, original order inside a bucket
 *   seg_offsets device [num_buckets+1] int32 out: bucket b owns perm[seg[b] .. seg[b+1])
 * Any of lvl / bucket / perm / seg_offsets may be NULL when not wanted (perm and seg_offsets
 * must be given together).  One single-CTA launch, no host synchronisation.
 */
int dm_assign(const float* rois, int K, const float* onehot, int num_buckets, int num_levels,
              float finest_scale, int32_t* lvl, int32_t* bucket, int32_t* perm,
              int32_t* seg_offsets, dm_stream_t stream);

/*
 * Stage 2 forward: aligned / avg-pool multi-level RoIAlign, all buckets in one persistent launch.
 *   feat_ptrs      host array [L] of device pointers, one fp32 map per FPN level
 *   feat_shapes    host [L*4]  (N, C, H, W) per level (C equal on all levels)
 *   feat_strides   host [L*4]  element strides (n, c, h, w) per level; any layout
 *   spatial_scales host [L]    1 / stride
 *   rois           device [K,5]
 *   lvl            device [K] int32 from dm_assign, or NULL (every RoI reads level 0)
 *   perm, seg_offsets  device, from dm_assign, or both NULL (one bucket holding RoIs 0..K-1)
 *   out_hw         host [num_buckets*2] pooled (h, w) of each bucket
 *   out_ptrs       host [num_buckets] device pointers, bucket b is [seg[b+1]-seg[b], C, h, w]
 *   out_strides    host [num_buckets*4] element strides (n, c, h, w) of each bucket's output
 *   sampling_ratio 0 = adaptive ceil(roi_size / out_size);  aligned 1 = half-pixel shift
 *   sched_scratch  device scratch of DM_SCHED_SCRATCH_BYTES owned by the caller and not touched by
 *                  anyone else until the launch has finished (the work-ticket counters of the
 *                  dynamic unit scheduler; cleared by the library on `stream`), or NULL: the work
 *                  units are then dealt statically (slower on mixed-resolution launches)
 * Every output element of every listed RoI is written (zeros where the reference yields zeros).
 */
int dm_roi_align_fwd(const float* const* feat_ptrs, const int32_t* feat_shapes,
                     const int64_t* feat_strides, const float* spatial_scales, int num_levels,
                     const float* rois, int K, const int32_t* lvl, const int32_t* perm,
                     const int32_t* seg_offsets, int num_buckets, const int32_t* out_hw,
                     float* const* out_ptrs, const int64_t* out_strides, int sampling_ratio,
                     int aligned, void* sched_scratch, dm_stream_t stream);

/*
 * Stage 2 backward: scatters grad_out of every bucket into the per-level gradient maps.
 * Arguments mirror dm_roi_align_fwd.  zero_init != 0 clears every grad map first (the maps must
 * then be dense: each level's storage is the N*C*H*W elements its strides span).
 */
int dm_roi_align_bwd(float* const* grad_feat_ptrs, const int32_t* feat_shapes,
                     const int64_t* feat_strides, const float* spatial_scales, int num_levels,
                     const float* rois, int K, const int32_t* lvl, const int32_t* perm,
                     const int32_t* seg_offsets, int num_buckets, const int32_t* out_hw,
                     const float* const* grad_out_ptrs, const int64_t* grad_out_strides,
                     int sampling_ratio, int aligned, int zero_init, void* sched_scratch,
                     dm_stream_t stream);

/*
 * Stage 3: paste N instance masks into image canvases, fused sigmoid + bilinear + threshold.
 *   masks          device fp32; instance n, class c plane at masks + n*stride_n + c*stride_c, [S_h,S_w] dense
 *   labels         device [N] int64 class per instance, or NULL (class 0)
 *   apply_sigmoid  1: masks hold logits;  0: masks hold probabilities
 *   boxes          device [N,4] fp32 xyxy in canvas pixels
 *   region         paste window [y_lo,y_hi) x [x_lo,x_hi) of the img_h x img_w canvas
 *   out            device [N, y_hi-y_lo, x_hi-x_lo] dense; element type per out_mode
 */
int dm_paste_masks(const float* masks, int64_t mask_stride_n, int64_t mask_stride_c,
                   const int64_t* labels, int N, int S_h, int S_w, int apply_sigmoid,
                   const float* boxes, int img_h, int img_w, int x_lo, int y_lo, int x_hi,
                   int y_hi, float thr, int out_mode, void* out, dm_stream_t stream);

/*
 * Stage 4: training mask targets for every positive RoI of a batch at several sizes, one launch.
 *   gt_blob      device uint8: the ground-truth bitmaps of all images, image b at gt_blob + img_offsets[b],
 *                laid out [G_b, H_b, W_b] dense
 *   img_offsets  device [B] int64 byte offsets;  img_ghw device [B*3] int32 (G_b, H_b, W_b)
 *   boxes        device [K,4] fp32 xyxy in image pixels; inds device [K] int64 assigned gt index
 *   roi_img      device [K] int32 image of each RoI, or NULL (all image 0)
 *   clip         1: clip x to [0,W_b], y to [0,H_b] first (mask_target_single); 0: as given
 *   sizes_hw     host [n_sizes*2]; out_ptrs host [n_sizes] device pointers, each [K, h, w] fp32 in {0,1}
 * Bit-exact with the reference's fp32 operation order (no FMA contraction, serial accumulation).
 */
int dm_mask_target(const uint8_t* gt_blob, const int64_t* img_offsets, const int32_t* img_ghw,
                   int B, const float* boxes, const int64_t* inds, const int32_t* roi_img, int K,
                   int clip, const int32_t* sizes_hw, int n_sizes, float* const* out_ptrs,
                   dm_stream_t stream);

/*
 * Next row: COCO run-length encoding of the pasted masks on the device.
 * Replaces encode_mask_results(get_seg_masks(...)) -- mmdet/core/mask/utils.py:36-63 applied to the
 * N numpy canvases of mmdet/models/roi_heads/mask_heads/dynamask_head.py:341 (pycocotools
 * mask.encode on the host after an N*H*W-byte device->host copy).
 *
 * A column-major run list is described by the sorted flat indices x*H + y at which the value
 * changes ("transitions").  Two passes over the same pixels:
 *   pass 1  writes col_counts [N, rw] int32 (transitions per column; only an instance's window
 *           columns are defined) and ADDS into inst_totals [N] int32 (zero it first);
 *   pass 2  given inst_offsets [N] int64 (exclusive scan of inst_totals) writes the transitions of
 *           instance n at transitions + inst_offsets[n], sorted.
 * dm_paste_rle evaluates the pixels from the mask logits exactly like dm_paste_masks in
 * DM_PASTE_BOOL mode (same arguments), never writing a canvas; dm_rle_from_canvas reads an existing
 * [N,H,W] one-byte-per-pixel canvas.
 */
int dm_paste_rle(const float* masks, int64_t mask_stride_n, int64_t mask_stride_c,
                 const int64_t* labels, int N, int S_h, int S_w, int apply_sigmoid,
                 const float* boxes, int img_h, int img_w, int x_lo, int y_lo, int x_hi, int y_hi,
                 float thr, int pass, int32_t* col_counts, int32_t* inst_totals,
                 const int64_t* inst_offsets, int32_t* transitions, dm_stream_t stream);
int dm_rle_from_canvas(const uint8_t* canvas, int N, int H, int W, int pass, int32_t* col_counts,
                       int32_t* inst_totals, const int64_t* inst_offsets, int32_t* transitions,
                       dm_stream_t stream);
/*
 * Device side of rleToString: the transitions of N instances (instance n owns
 * transitions[inst_offsets[n] .. inst_offsets[n+1]); inst_offsets has N+1 entries, device) become
 * pycocotools' compressed "counts" strings, written back to back into `out` with string n at
 * out[str_offsets[n] .. str_offsets[n+1]) -- only the strings need to cross PCIe.  Caller-owned
 * device scratch: compact [sum of transitions] int32, kept [N] int32, str_len [N] int32;
 * str_offsets [N+1] int64 and out (6 bytes per transition + 8 per instance always suffice) are
 * outputs.  Three small launches (compact + lengths, scan, write), no host synchronisation.
 * Replaces the string building of pycocotools mask.encode reached from
 * mmdet/core/mask/utils.py:36-63.
 */
int dm_rle_strings(const int32_t* transitions, const int64_t* inst_offsets, int N, int64_t total_pixels,
                   int32_t* compact, int32_t* kept, int32_t* str_len, int64_t* str_offsets, char* out,
                   dm_stream_t stream);
/*
 * The whole paste -> RLE-string pipeline of one image in ONE call, with no host round trip in the
 * middle: count (dm_paste_rle pass 1), scan the per-instance counts on the device, write the
 * transitions (pass 2), build the strings (dm_rle_strings).  The caller sizes the buffers from a
 * transition CAPACITY instead of the exact total, so the call can be enqueued without waiting for
 * anything; results are read after one synchronisation (or an event) chosen by the caller, which
 * lets an inference loop enqueue image i+1 while image i's strings travel to the host
 * (mmdet/apis/test.py:24-57 handles one image after the other).
 *   workspace  device scratch of dm_paste_rle_strings_workspace(N, x_hi - x_lo, y_hi - y_lo, capacity) bytes,
 *              16-byte aligned, owned by the caller until the call has finished on `stream`
 *   capacity   transitions the buffers hold (all instances together)
 *   record_slots  0 (default form): count + write passes, the region's rows cut into segments of 128 rows so
 *              that a tall window is walked by several CTAs per column block; != 0: one segment, the
 *              counting pass also records up to 32 transitions per canvas column and column blocks
 *              whose columns all fit are copied into place instead of being evaluated a second time
 *   header     device int64 [2 + N + 1]: header[0] = status: bit 0 set = the masks have more
 *              transitions than `capacity` (nothing else was written, repeat with capacity >=
 *              header[1]), bits 8.. = column blocks that overflowed their slots and were evaluated
 *              twice; header[1] = total transitions, header[2 .. 2+N] = string offsets
 *   out        device chars, at least 6 * capacity + 8 * N + 8 bytes
 * Other arguments as dm_paste_rle.  Replaces get_seg_masks + encode_mask_results
 * (mmdet/models/roi_heads/mask_heads/dynamask_head.py:279-342, mmdet/core/mask/utils.py:36-63).
 */
int64_t dm_paste_rle_strings_workspace(int N, int rw, int rh, int64_t capacity);
int dm_paste_rle_strings(const float* masks, int64_t mask_stride_n, int64_t mask_stride_c,
                         const int64_t* labels, int N, int S_h, int S_w, int apply_sigmoid,
                         const float* boxes, int img_h, int img_w, int x_lo, int y_lo, int x_hi,
                         int y_hi, float thr, int record_slots, void* workspace, int64_t capacity,
                         int64_t* header, char* out, dm_stream_t stream);
/*
 * HOST function: the transitions of one instance (host memory) -> pycocotools' compressed "counts"
 * string (rleToString).  Coinciding transition pairs cancel.  Returns the length written to `out`
 * (no terminator) or -1 if `cap` is too small.
 */
int64_t dm_rle_compress_host(const int32_t* transitions, int64_t n, int64_t total_pixels, char* out,
                             int64_t cap);
/*
 * HOST function, batch form: instance n owns transitions[offsets[n] .. offsets[n+1]); the N strings
 * are written back to back into `out`, string n at out[str_offsets[n] .. str_offsets[n+1]).
 * Returns the total length or -1 if `cap` is too small (6 * transitions + 8 * N always suffices).
 */
int64_t dm_rle_compress_batch_host(const int32_t* transitions, const int64_t* offsets, int64_t N,
                                   int64_t total_pixels, char* out, int64_t cap, int64_t* str_offsets);

/*
 * Next row: SimpleRoIAlign, the per-RoI semantic-feature gather of SFMStage
 * (mmdet/models/roi_heads/mask_heads/dynamask_head.py:74 construction, :104-105 call; the three
 * stages sample strides 16/8/4 maps with spatial_scale 1/4, :185-194,228).  Replaces
 * mmcv.ops.SimpleRoIAlign = generate_grid + rel_roi_point_to_rel_img_point + point_sample
 * (F.grid_sample, bilinear, zero padding, align_corners = !aligned): one sample per output bin at
 *   x = x1 + (pw + 0.5) / out_w * (x2 - x1),  pixel = x / W * spatial_scale * W - 0.5   (aligned)
 * taps outside the map contribute zero (no border clamp, unlike RoIAlign).
 *   feat        device [N,C,H,W] fp32, feat_shape host [4], feat_strides host [4] (elements)
 *   rois        device [K,5] (batch_idx, x1, y1, x2, y2) in input-image pixels
 *   out         device [K,C,out_h,out_w], out_strides host [4]; row k belongs to RoI k
 * The backward adds into grad_feat (cleared first when zero_init != 0).  sched_scratch as in
 * dm_roi_align_fwd.
 */
int dm_simple_roi_align_fwd(const float* feat, const int32_t* feat_shape,
                            const int64_t* feat_strides, float spatial_scale, const float* rois,
                            int K, int out_h, int out_w, float* out, const int64_t* out_strides,
                            int aligned, void* sched_scratch, dm_stream_t stream);
int dm_simple_roi_align_bwd(float* grad_feat, const int32_t* feat_shape,
                            const int64_t* feat_strides, float spatial_scale, const float* rois,
                            int K, int out_h, int out_w, const float* grad_out,
                            const int64_t* grad_out_strides, int aligned, int zero_init,
                            void* sched_scratch, dm_stream_t stream);

/*
 * Next row: inference-time stage-to-stage refinement, fused.  Replaces the
 * loop of DynaMaskRoIHead.simple_test_mask, mmdet/models/roi_heads/dynamask_roi_head.py:137-149
 * (sigmoid >= 0.5, generate_block_target(boundary_width=1) of
 * mmdet/models/losses/cross_entropy_loss.py:123-154, two bilinear align_corners=True
 * interpolations and a masked overwrite per stage pair).
 *   stage_ptrs  host [n_stages] device pointers, stage s is [N, h_s, w_s] fp32 logits, dense
 *   sizes_hw    host [n_stages*2]
 *   out_ptrs    host [n_stages] device pointers for the refined stages; out_ptrs[s] may alias
 *               stage_ptrs[s] (the reference refines in place) and may be NULL for every stage but
 *               the last (stage 0 is never changed)
 * For s = 0 .. n_stages-2: pixels of stage s+1 whose up-sampled non-boundary mask of (refined)
 * stage s is >= 0.5 take the up-sampled stage-s logit.  One CTA per instance, one launch.
 */
int dm_refine_stages(const float* const* stage_ptrs, const int32_t* sizes_hw, int n_stages, int N,
                     float* const* out_ptrs, dm_stream_t stream);

/*
 * Next row: mask targets from POLYGON ground truth, all RoIs and all
 * sizes in one launch.  Replaces PolygonMasks.crop_and_resize (mmdet/core/mask/structures.py:465-499)
 * + PolygonMasks.to_ndarray / polygon_to_bitmap (structures.py:541-575: pycocotools frPyObjects ->
 * merge -> decode) + the clip / float / upload of mask_target_single (mmdet/core/mask/mask_target.py:49-58)
 * and of DynaMaskHead.get_targets (mmdet/models/roi_heads/mask_heads/dynamask_head.py:248-261).
 *   poly_xy          device float64, interleaved (x, y) vertices of every polygon of the batch
 *   vert_offsets     device [P+1] int64: polygon q owns vertices [vert_offsets[q], vert_offsets[q+1])
 *   obj_poly_offsets device [G+1] int32: object g owns polygons [obj_poly_offsets[g], obj_poly_offsets[g+1])
 *   img_meta         device [B*3] int32: (first object of the image, H, W) -- H, W bound the clip
 *   boxes / inds / roi_img / clip / sizes_hw / out_ptrs   as in dm_mask_target
 * Per RoI: vertex' = (vertex - box corner) * (out / max(box extent, 1)) in float64 with a float32
 * scale, then pycocotools' rleFrPoly rule (x5 up-sampling, run ends on column centres), union over
 * the object's polygons.  Integer / float64 arithmetic without contraction: bit-exact.
 */
int dm_polygon_target(const double* poly_xy, const int64_t* vert_offsets,
                      const int32_t* obj_poly_offsets, int G, const int32_t* img_meta, int B,
                      const float* boxes, const int64_t* inds, const int32_t* roi_img, int K,
                      int clip, const int32_t* sizes_hw, int n_sizes, float* const* out_ptrs,
                      dm_stream_t stream);

/*
 * Next row: switch-driven inference.  The reference sketches it in comments
 * (mmdet/models/roi_heads/dynamask_roi_head.py:176-203): paste ALL four stage predictions of every
 * detection (four get_seg_masks passes), then keep chunk_segm_result[mask_labels[j]][j].  Here each
 * detection is pasted once, from the stage its mask-switch label selects: call once per stage with
 * that stage's [N,C,S,S] masks; instance n is written only when select[n] == select_value
 * (select = the bucket array of dm_assign).  zero_fill != 0 clears the whole [N,h,w] output first
 * (first call of a group).  Other arguments as dm_paste_masks.
 */
int dm_paste_masks_select(const float* masks, int64_t mask_stride_n, int64_t mask_stride_c,
                          const int64_t* labels, int N, int S_h, int S_w, int apply_sigmoid,
                          const float* boxes, int img_h, int img_w, int x_lo, int y_lo, int x_hi,
                          int y_hi, float thr, int out_mode, const int32_t* select, int select_value,
                          int zero_fill, void* out, dm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNAMASK_SM100_H_ */
