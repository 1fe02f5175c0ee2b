"""Config-shaped workloads of BASELINE.json (configs[2..4]) for bench.py: every rank runs them, the
step times are reduced with MAX over ranks and the units summed, so each N of the scaling run
carries img/s for C3 (training step), C4 (inference, 64 images sharded by image) and C5 (dense
1024x2048 inference).  Reference flows matched:
  C3  mmdet/models/roi_heads/dynamask_roi_head.py:48-73 (bbox / mask / 56x56 switch extractors) and
      mmdet/models/roi_heads/mask_heads/dynamask_head.py:246-271 (get_targets, fresh ground truth per step)
  C4  mmdet/apis/test.py:24-57 + dynamask_roi_head.py:117-158 (per image: 14x14 extractor -> [head,
      PyTorch, not timed] -> stage refinement -> paste -> RLE results on the host)
  C5  the same tail at 300 detections on a 1024x2048 image (configs/refinemask/lvis, cityscapes)
Nothing here touches oracle/; GPU only.
"""
import time

import numpy as np
import torch

import synth

STRIDES = [4, 8, 16, 32]


class _Cfg:
    mask_thr_binary = 0.5


def _reduce_max(ms, dev, world):
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def _footprint_bytes(rois, lvl, shapes, strides, C):
    """4 C sum_r fp_r: feature pixels in every RoI's bilinear footprint at its level."""
    r = rois.double().cpu()
    lvl = lvl.cpu().long()
    fp = torch.zeros(r.size(0), dtype=torch.float64)
    for l, (h, w) in enumerate(shapes):
        s = 1.0 / strides[l]
        m = lvl == l
        if not bool(m.any()):
            continue
        x_lo = torch.clamp(torch.floor(r[m, 1] * s - 0.5), min=0)
        x_hi = torch.clamp(torch.floor(r[m, 3] * s - 0.5) + 1, max=w - 1)
        y_lo = torch.clamp(torch.floor(r[m, 2] * s - 0.5), min=0)
        y_hi = torch.clamp(torch.floor(r[m, 4] * s - 0.5) + 1, max=h - 1)
        fp[m] = torch.clamp(x_hi - x_lo + 1, min=0) * torch.clamp(y_hi - y_lo + 1, min=0)
    return 4.0 * C * float(fp.sum())


# ----------------------------------------------------------------------------------------------
# C3: training step, 2 images per GPU (weak scaling)
# ----------------------------------------------------------------------------------------------
def run_c3(dm, ops, dev, rank, world, peak, steps=10, warmup=3, img_hw=(800, 1344), channels=256):
    H, W = img_hw
    g = torch.Generator().manual_seed(300 + rank)
    rng = np.random.default_rng(300 + rank)
    # objects per image: a fixed spread over 1..20 (mean 10.5), the same on every rank (weak scaling compares equal work)
    n_obj = iter([3, 18, 7, 14, 10, 5, 16, 11])
    shapes = synth.pyramid_shapes(H, W)
    feats = [torch.randn(2, channels, h, w, device=dev) for (h, w) in shapes]
    rois = synth.make_rois(2, 512, H, W, g).to(dev)                       # bbox head: 512 samples per image
    # new ground truth for every step: a pool of 4 batches of mask objects built beforehand (the data loader's job:
    # the reference's constructor copies the arrays, np.stack); a multi-image batch is packed and uploaded by every
    # mask_target call, nothing is cached on the device
    pool_b, pool_p = [], []
    for _ in range(4):
        imgs_b, imgs_p = [], []
        for _ in range(2):
            g_img = next(n_obj)
            m = synth.make_gt_masks(g_img, H, W, rng)
            pb, pi = synth.jitter_boxes_from_masks(m, 128, rng)
            imgs_b.append((m, torch.from_numpy(pb).to(dev), torch.from_numpy(pi).to(dev), dm.BitmapMasks(m, H, W)))
            objs = synth.make_polygons(g_img, H, W, rng)
            qb, qi = synth.jitter_boxes_from_polygons(objs, 128, rng)
            imgs_p.append((objs, torch.from_numpy(qb).to(dev), torch.from_numpy(qi).to(dev), dm.PolygonMasks(objs, H, W)))
        pool_b.append(imgs_b)
        pool_p.append(imgs_p)
    ext7 = dm.SingleRoIExtractor(dict(type='RoIAlign', output_size=7, sampling_ratio=0), channels, STRIDES)
    ext14 = dm.SingleRoIExtractor(dict(type='RoIAlign', output_size=14, sampling_ratio=0), channels, STRIDES)
    ext56 = dm.SingleRoIExtractor(dict(type='RoIAlign', output_size=56, sampling_ratio=0), channels, [4])
    fr = [f.requires_grad_() for f in feats]

    def step(i, polygons):
        for f in fr:
            f.grad = None
        imgs = (pool_p if polygons else pool_b)[i % 4]
        pos = [torch.cat([torch.full((128, 1), float(b), device=dev), imgs[b][1][:, :4]], 1) for b in range(2)]
        r_mask = torch.cat(pos)
        o7 = ext7(fr, rois)
        o14 = ext14(fr, r_mask)
        o56 = ext56([fr[0].detach()], r_mask)                                 # dynamask_roi_head.py:59 (detached)
        gts = [imgs[b][3] for b in range(2)]      # the batch's masks are packed and uploaded by every call (no cache)
        tg = dm.multi_size_mask_targets([imgs[b][1] for b in range(2)], [imgs[b][2] for b in range(2)], gts)
        torch.autograd.backward([o7, o14], [o7.detach(), o14.detach()])       # stand-in for the heads' gradients
        return o56, tg

    out = {}
    for polygons in (False, True):
        for i in range(warmup):
            step(i, polygons)
        _barrier(world)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for i in range(steps):
            step(i, polygons)
        b.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / steps * 1e3
        dev_ms = a.elapsed_time(b) / steps
        ms = _reduce_max(max(wall, dev_ms), dev, world)
        key = 'polygon_gt' if polygons else 'bitmap_gt'
        out[key] = {'ms_per_step': ms, 'img_per_s': 2 * world / ms * 1e3,
                    'rois_per_s': (1024 + 256 + 256) * world / ms * 1e3,
                    'gt_upload': 'ground truth of the step packed and uploaded inside the step (%s)' % (
                        'polygon vertices, KBs' if polygons else 'uint8 bitmaps, ~11 MB per image, one pinned staging copy')}
    # roofline of the extractor + target kernels of one step (algorithmic bytes)
    lvl7 = ops.assign(rois, None, 4, 56.0, 1)[0]
    imgs = pool_b[0]
    r_mask = torch.cat([torch.cat([torch.full((128, 1), float(b), device=dev), imgs[b][1][:, :4]], 1) for b in range(2)])
    lvl14 = ops.assign(r_mask, None, 4, 56.0, 1)[0]
    fp7 = _footprint_bytes(rois, lvl7, shapes, STRIDES, channels)
    fp14 = _footprint_bytes(r_mask, lvl14, shapes, STRIDES, channels)
    fp56 = _footprint_bytes(r_mask, torch.zeros_like(lvl14), shapes[:1], [4], channels)
    pyramid = 4.0 * 2 * channels * sum(h * w for h, w in shapes)
    by = (4.0 * channels * 1024 * 49 * 2 + 3 * fp7 + pyramid            # 7x7 fwd + bwd (+ zero-init)
          + 4.0 * channels * 256 * 196 * 2 + 3 * fp14 + pyramid         # 14x14 fwd + bwd
          + 4.0 * channels * 256 * 3136 + fp56                          # 56x56 forward
          + 256 * 4.0 * (14 * 14 + 28 * 28 + 56 * 56 + 112 * 112))      # targets written (reads: the boxes' pixels)
    ms = out['bitmap_gt']['ms_per_step']
    out['workload'] = ('C3 training step per GPU: 2 images (800x1344), bbox extractor 7x7 x 1024 RoIs fwd+bwd, mask '
                       'extractor 14x14 x 256 positives fwd+bwd, 56x56 single-level switch input x 256 fwd, mask targets '
                       'at 14/28/56/112 from ground truth that is new every step; heads / losses excluded (PyTorch)')
    out['roofline'] = {'bound': 'hbm', 'algorithmic_bytes': by, 'achieved': by / ms / 1e6, 'peak': peak, 'unit': 'GB/s',
                       'frac': by / ms / 1e6 / peak,
                       'note': 'whole step by the host clock (launch gaps of ~10 small calls included)'}
    return out


# ----------------------------------------------------------------------------------------------
# C4 / C5: inference tail per image, results on the host as RLE
# ----------------------------------------------------------------------------------------------
def _tail_setup(dm, dev, seed, n_img, dets, img_hw, channels, small_frac):
    H, W = img_hw
    g = torch.Generator().manual_seed(seed)
    shapes = synth.pyramid_shapes(H, W)
    images = []
    for _ in range(n_img):
        boxes = synth.make_boxes(dets, H, W, g, small_frac=small_frac)
        rois = torch.cat([torch.zeros(dets, 1), boxes], 1).to(dev)
        det = torch.cat([boxes, torch.ones(dets, 1)], 1).to(dev)
        images.append((rois, det))
    # one pyramid and one set of stage logits stand for every image's (their content does not change the work)
    feats = [torch.randn(1, channels, h, w, device=dev) for (h, w) in shapes]
    # every stage predicts the same object: radial blob + N(0,1) noise per pixel at each stage's resolution (round 2's
    # earlier runs fed pure noise to the 28 / 56 stages: ~50 sign changes per canvas column, 237 KB of RLE per image)
    stages = [synth.make_mask_logits(dets, s, g).to(dev) for s in (28, 56, 112)]
    labels = torch.zeros(dets, dtype=torch.long, device=dev)
    ext = dm.SingleRoIExtractor(dict(type='RoIAlign', output_size=14, sampling_ratio=0), channels, STRIDES)
    return feats, stages, labels, ext, images


def _tail_image(dm, ext, feats, stages, labels, rois, det, ori_shape, wait=True):
    ins = ext(feats, rois)                                                # 14x14 instance features -> the head
    final = dm.refine_stage_instance_preds([t.clone() for t in stages])   # dynamask_roi_head.py:136-148, fused
    rles = dm.get_seg_masks_rle(final, det, labels, _Cfg, ori_shape, 1.0, False, wait=wait)
    return ins, rles


def run_tail(dm, dev, rank, world, peak, total_images, dets, img_hw, ori_hw, strong, seed, small_frac=0.0,
             channels=256, reps=2):
    """`strong`: total_images are shared out over the ranks (C4: 64 images over N GPUs); otherwise every
    rank runs total_images of its own (C5 weak scaling)."""
    mine = len(range(rank, total_images, world)) if strong else total_images
    feats, stages, labels, ext, images = _tail_setup(dm, dev, seed + rank, max(mine, 1), dets, img_hw, channels,
                                                     small_frac)
    ori_shape = (ori_hw[0], ori_hw[1], 3)
    # The loop of mmdet/apis/test.py:24-57, software-pipelined one image deep: image i+1 is enqueued before the
    # strings of image i are collected (get_seg_masks_rle(wait=False) enqueues without a host synchronisation),
    # so the host's launch work overlaps the device's.  Every image's strings are on the host, as Python bytes,
    # before the clock stops.
    def one_pass(n_img):
        nbytes = 0
        pending = None
        for i in range(n_img):
            _, nxt = _tail_image(dm, ext, feats, stages, labels, images[i][0], images[i][1], ori_shape, wait=False)
            if pending is not None:
                nbytes += sum(len(r['counts']) for r in pending.result())
            pending = nxt
        if pending is not None:
            nbytes += sum(len(r['counts']) for r in pending.result())
        return nbytes

    # warm-up through the SAME pipelined loop: its two sets of buffers in flight are first allocated here (one
    # cudaHostAlloc of a new size cost 69 ms on a 2-GPU box, tools/gpu/r03_tail.py)
    one_pass(mine)
    _barrier(world)
    # enough passes for ~32 images, each timed on its own; the MEDIAN pass is reported (a single cudaHostAlloc or a
    # neighbour's burst on the box is 1-70 ms against a 2-30 ms pass)
    reps = max(reps, min(8, -(-32 // max(mine, 1))))
    rle_bytes = 0
    passes = []
    with torch.no_grad():   # mmdet/apis/test.py:24-57 runs the model under no_grad
        for _ in range(reps):
            t0 = time.perf_counter()
            rle_bytes += one_pass(mine)
            torch.cuda.synchronize()
            passes.append((time.perf_counter() - t0) * 1e3)
    wall_ms = sorted(passes)[len(passes) // 2]
    # the same images one at a time (collect before the next image is enqueued): the latency view
    t1 = time.perf_counter()
    for i in range(mine):
        _tail_image(dm, ext, feats, stages, labels, images[i][0], images[i][1], ori_shape)
    serial_ms = (time.perf_counter() - t1) * 1e3 / max(mine, 1)
    ms = _reduce_max(wall_ms, dev, world)
    n_total = total_images if strong else total_images * world
    # device time of the same calls (CUDA events around one pass), for the share of host work
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lvl = None
    by = 0.0
    shapes = synth.pyramid_shapes(*img_hw)
    from dynamask_b200 import ops
    for i in range(mine):
        lvl = ops.assign(images[i][0], None, 4, 56.0, 1)[0]
        by += 4.0 * channels * dets * 196 + _footprint_bytes(images[i][0], lvl, shapes, STRIDES, channels)   # extractor
        by += dets * 4.0 * (28 * 28 + 56 * 56 + 112 * 112) * 2                                        # refinement r/w
        by += dets * 4.0 * 112 * 112                                                                  # paste reads the logits
    return {'images': n_total, 'images_this_rank': mine, 'ms_per_pass': ms, 'img_per_s': n_total / ms * 1e3,
            'instances_per_s': n_total * dets / ms * 1e3, 'ms_per_image_this_rank': wall_ms / max(mine, 1),
            'ms_per_image_unpipelined': serial_ms, 'passes_ms': [round(v, 3) for v in passes], 'pipeline': 'one image deep (image i+1 enqueued before image i is collected)',
            'rle_bytes_to_host_per_image': rle_bytes / max(reps * mine, 1),
            'roofline': {'bound': 'latency (per-image calls of ~0.1 ms kernels; results leave as ~100 KB of RLE)',
                         'algorithmic_bytes_this_rank': by, 'achieved': by / wall_ms / 1e6, 'peak': peak, 'unit': 'GB/s',
                         'frac': by / wall_ms / 1e6 / peak},
            'e2e': 'results on the host inside the timed region (RLE strings); inputs device resident (backbone output)'}


def run_all(dm, ops, dev, rank, world, peak):
    res = {}
    res['c3'] = run_c3(dm, ops, dev, rank, world, peak)
    c4 = run_tail(dm, dev, rank, world, peak, 64, 100, (800, 1344), (800, 1333), True, 400)
    c4['workload'] = ('C4 inference, 64 images (800x1333, 100 detections each) sharded by image over the ranks, per image: '
                      '14x14 mask extractor (256 ch) -> [head convolutions: PyTorch, not timed] -> fused stage refinement '
                      '28/56/112 -> fused paste->RLE, RLE strings on the host; stage logits: blob + N(0,1) noise at every stage')
    c4['scaling'] = 'strong'
    res['c4'] = c4
    c5 = run_tail(dm, dev, rank, world, peak, 4, 300, (1024, 2048), (1024, 2048), False, 500, small_frac=0.8)
    c5['workload'] = ('C5 dense case, 4 images per GPU of 1024x2048 with 300 detections each (80 % small), same per-image '
                      'tail as C4')
    c5['scaling'] = 'weak'
    res['c5'] = c5
    return res
