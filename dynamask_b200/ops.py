"""torch custom ops over the C ABI of ``libdynamask_sm100.so``.

Each op allocates its outputs with the PyTorch caching allocator, passes raw device pointers,
sizes and the current CUDA stream across ``extern "C"`` and returns immediately (no host
synchronisation).  Ops are registered for CUDA tensors only: a CPU tensor raises
``NotImplementedError`` from the dispatcher -- there is no CPU or eager-PyTorch fallback.
"""
import ctypes
from typing import List, Optional, Sequence

import torch
from torch import Tensor

from . import _lib

__all__ = ['assign', 'roi_align_forward', 'roi_align_backward', 'paste_masks', 'mask_target',
           'multilevel_roi_align', 'simple_roi_align_forward', 'simple_roi_align_backward', 'refine_stages_', 'polygon_target', 'paste_masks_switched']

PASTE_BOOL, PASTE_U8, PASTE_F32 = 0, 1, 2


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _arr(ctype, vals):
    return (ctype * max(len(vals), 1))(*vals)


def _sched_scratch(dev):
    """Caller-owned scratch of one RoIAlign launch (DM_SCHED_SCRATCH_BYTES): the work-ticket counters
    of the dynamic unit scheduler.  A fresh block of the caching allocator per launch, so launches on
    different streams -- or replays of a captured graph -- never share counters; the allocator only
    hands the block out again in stream order."""
    return torch.empty(16, dtype=torch.int32, device=dev)


def _f32c(t, name):
    if t.dtype != torch.float32:
        raise TypeError('%s must be float32, got %s' % (name, t.dtype))
    return t


def call(op, *args, grad_inputs=()):
    """Call a ``dynamask::`` custom op.  In plain eager execution with nothing to differentiate (inference under
    ``torch.no_grad()``, index-only ops) the op's Python body runs directly: the dispatcher round trip of
    ``torch.library.custom_op`` costs ~90 us per call on this box, a third of the host time of the per-image
    inference tail (``tools/gpu/r03_tail.py``).  Under autograd, ``torch.compile`` tracing or any tensor subclass /
    mode the registered op is used, so those see exactly what they saw before."""
    body = getattr(op, '_init_fn', None)
    if body is None or torch.compiler.is_compiling():
        return op(*args)
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in grad_inputs):
        return op(*args)
    for a in args:
        for t in (a if isinstance(a, (list, tuple)) else (a,)):
            if isinstance(t, Tensor) and (type(t) is not Tensor or not t.is_cuda):
                return op(*args)   # subclasses (fake / functional tensors) and CPU tensors (which must raise) go through
    try:   # any active __torch_dispatch__ / __torch_function__ mode must see the op
        if torch._C._len_torch_dispatch_stack() > 0 or torch._C._len_torch_function_stack() > 0:
            return op(*args)
    except AttributeError:   # another torch version: stay on the registered op
        return op(*args)
    return body(*args)


# --------------------------------------------------------------------------------------------
# dm_assign
# --------------------------------------------------------------------------------------------
@torch.library.custom_op('dynamask::assign', mutates_args=(), device_types='cuda')
def assign(rois: Tensor, onehot: Optional[Tensor], num_levels: int, finest_scale: float,
           num_buckets: int) -> List[Tensor]:
    """-> [lvl int32 [K], bucket int32 [K], perm int32 [K], seg_offsets int32 [num_buckets+1]]."""
    if rois.dim() != 2 or rois.size(1) != 5:
        raise ValueError('rois must be [K,5]')
    rois = _f32c(rois, 'rois').contiguous()
    K = rois.size(0)
    dev = rois.device
    if onehot is not None:
        onehot = _f32c(onehot, 'onehot').contiguous()
        if onehot.dim() != 2 or onehot.size(0) != K or onehot.size(1) != num_buckets:
            raise ValueError('onehot must be [K,num_buckets]')
    lvl = torch.empty(K, dtype=torch.int32, device=dev)
    bucket = torch.empty(K, dtype=torch.int32, device=dev)
    perm = torch.empty(K, dtype=torch.int32, device=dev)
    seg = torch.empty(num_buckets + 1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().dm_assign(_ptr(rois), K, _ptr(onehot), num_buckets, num_levels,
                                   float(finest_scale), _ptr(lvl), _ptr(bucket), _ptr(perm),
                                   _ptr(seg), _stream(dev))
    _lib.check(rc, 'dm_assign')
    return [lvl, bucket, perm, seg]


@assign.register_fake
def _(rois, onehot, num_levels, finest_scale, num_buckets):
    K = rois.size(0)
    i32 = dict(dtype=torch.int32, device=rois.device)
    return [torch.empty(K, **i32), torch.empty(K, **i32), torch.empty(K, **i32),
            torch.empty(num_buckets + 1, **i32)]


# --------------------------------------------------------------------------------------------
# dm_roi_align_fwd / dm_roi_align_bwd
# --------------------------------------------------------------------------------------------
def _level_arrays(feats):
    L = len(feats)
    ptrs = _arr(ctypes.c_void_p, [f.data_ptr() for f in feats])
    shapes, strides = [], []
    for f in feats:
        if f.dim() != 4:
            raise ValueError('feature maps must be [N,C,H,W]')
        _f32c(f, 'feats')
        shapes += list(f.shape)
        strides += list(f.stride())
    return L, ptrs, _arr(ctypes.c_int32, shapes), _arr(ctypes.c_int64, strides)


def _tma_rows(f):
    """A level whose rows are not a multiple of 16 bytes (W = 42 on the stride-32 map of an 800x1344
    image) cannot be described by a tensor map, and its RoIs would take the slow 8-byte cp.async path
    of the forward (measured: 9 % of the RoIs, a third of a 7x7 call's time).  Such a level is copied
    once per call into a buffer with a 16-byte row pitch and handed to the library as a strided view
    (same shape, pitch rounded up to 4 floats); the pad columns are never read (the tensor map's extent
    is W, out-of-range box columns are zero-filled by the TMA unit).  A backbone that allocates the
    level with that pitch avoids the copy."""
    if f.dim() != 4 or f.size(3) % 4 == 0 or not f.is_contiguous():
        return f
    n, c, h, w = f.shape
    buf = torch.empty((n, c, h, (w + 3) & ~3), dtype=f.dtype, device=f.device)
    view = buf[..., :w]
    view.copy_(f)
    return view


def _bucket_arrays(tensors, out_hw):
    ptrs = _arr(ctypes.c_void_p, [t.data_ptr() for t in tensors])
    strides = []
    for t in tensors:
        strides += list(t.stride())
    return ptrs, _arr(ctypes.c_int32, list(out_hw)), _arr(ctypes.c_int64, strides)


@torch.library.custom_op('dynamask::roi_align_forward', mutates_args=(), device_types='cuda')
def roi_align_forward(feats: Sequence[Tensor], rois: Tensor, lvl: Optional[Tensor],
                      perm: Optional[Tensor], seg: Optional[Tensor], counts: Sequence[int],
                      out_hw: Sequence[int], spatial_scales: Sequence[float], sampling_ratio: int,
                      aligned: bool, channels_last: bool) -> List[Tensor]:
    """Multi-level, multi-bucket RoIAlign forward.

    ``counts[b]`` RoIs of bucket b (host ints), pooled size ``out_hw[2b:2b+2]``.  Returns one
    ``[counts[b], C, h, w]`` tensor per bucket (NCHW-contiguous, or channels_last on request).
    """
    nb = len(counts)
    if len(out_hw) != 2 * nb or len(spatial_scales) != len(feats):
        raise ValueError('inconsistent bucket / level arguments')
    if rois.dim() != 2 or rois.size(1) != 5:
        raise ValueError('rois must be [K,5]')
    rois = _f32c(rois, 'rois').contiguous()
    dev = rois.device
    K = rois.size(0)
    # the kernel enumerates a bucket's rows from the device-side seg_offsets; the outputs are sized from the
    # host counts, which therefore must be the seg differences of THIS call (stale counts would write past
    # the end of a bucket tensor).  The total is checkable without a device read-back:
    if sum(int(c) for c in counts) != K:
        raise ValueError('counts must sum to the number of RoIs (%d), got %s' % (K, list(counts)))
    C = feats[0].size(1)
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    outs = [torch.empty((int(counts[b]), C, int(out_hw[2 * b]), int(out_hw[2 * b + 1])),
                        dtype=torch.float32, device=dev, memory_format=fmt) for b in range(nb)]
    if K == 0:
        return outs
    if any(int(out_hw[2 * b + 1]) <= 32 for b in range(nb)):   # pooled rows of <= 8 strips: TMA patch loads
        feats = [_tma_rows(f) for f in feats]
    L, fptrs, fshapes, fstrides = _level_arrays(feats)
    optrs, ohw, ostrides = _bucket_arrays(outs, out_hw)
    scales = _arr(ctypes.c_float, [float(s) for s in spatial_scales])
    with torch.cuda.device(dev):
        sched = _sched_scratch(dev)
        rc = _lib.load().dm_roi_align_fwd(fptrs, fshapes, fstrides, scales, L, _ptr(rois), K,
                                          _ptr(lvl), _ptr(perm), _ptr(seg), nb, ohw, optrs,
                                          ostrides, int(sampling_ratio), int(bool(aligned)),
                                          _ptr(sched), _stream(dev))
    _lib.check(rc, 'dm_roi_align_fwd')
    return outs


@roi_align_forward.register_fake
def _(feats, rois, lvl, perm, seg, counts, out_hw, spatial_scales, sampling_ratio, aligned,
      channels_last):
    C = feats[0].size(1)
    return [torch.empty((counts[b], C, out_hw[2 * b], out_hw[2 * b + 1]), dtype=torch.float32,
                        device=rois.device) for b in range(len(counts))]


@torch.library.custom_op('dynamask::roi_align_backward', mutates_args=(), device_types='cuda')
def roi_align_backward(grad_outs: Sequence[Tensor], rois: Tensor, lvl: Optional[Tensor],
                       perm: Optional[Tensor], seg: Optional[Tensor],
                       feat_shapes: Sequence[int], feat_channels_last: Sequence[bool],
                       out_hw: Sequence[int], spatial_scales: Sequence[float], sampling_ratio: int,
                       aligned: bool) -> List[Tensor]:
    """Gradient w.r.t. every level's feature map (zero-initialised here, then reduced into)."""
    nb = len(grad_outs)
    L = len(spatial_scales)
    dev = rois.device
    rois = _f32c(rois, 'rois').contiguous()
    K = rois.size(0)
    grads = []
    for l in range(L):
        shp = [int(v) for v in feat_shapes[4 * l:4 * l + 4]]
        fmt = torch.channels_last if feat_channels_last[l] else torch.contiguous_format
        grads.append(torch.empty(shp, dtype=torch.float32, device=dev, memory_format=fmt))
    gos = [_f32c(g, 'grad_out') for g in grad_outs]
    _, gptrs, gshapes, gstrides = _level_arrays(grads)
    optrs, ohw, ostrides = _bucket_arrays(gos, out_hw)
    scales = _arr(ctypes.c_float, [float(s) for s in spatial_scales])
    with torch.cuda.device(dev):
        sched = _sched_scratch(dev)
        rc = _lib.load().dm_roi_align_bwd(gptrs, gshapes, gstrides, scales, L, _ptr(rois), K,
                                          _ptr(lvl), _ptr(perm), _ptr(seg), nb, ohw, optrs,
                                          ostrides, int(sampling_ratio), int(bool(aligned)), 1,
                                          _ptr(sched), _stream(dev))
    _lib.check(rc, 'dm_roi_align_bwd')
    return grads


@roi_align_backward.register_fake
def _(grad_outs, rois, lvl, perm, seg, feat_shapes, feat_channels_last, out_hw, spatial_scales,
      sampling_ratio, aligned):
    L = len(spatial_scales)
    return [torch.empty([int(v) for v in feat_shapes[4 * l:4 * l + 4]], dtype=torch.float32,
                        device=rois.device) for l in range(L)]


def _ra_setup_context(ctx, inputs, output):
    (feats, rois, lvl, perm, seg, counts, out_hw, spatial_scales, sampling_ratio, aligned,
     channels_last) = inputs
    ctx.save_for_backward(rois, *[t for t in (lvl, perm, seg) if t is not None])
    ctx.has = (lvl is not None, perm is not None, seg is not None)
    ctx.feat_shapes = [int(v) for f in feats for v in f.shape]
    ctx.feat_cl = [bool(f.dim() == 4 and not f.is_contiguous()
                        and f.is_contiguous(memory_format=torch.channels_last)) for f in feats]
    ctx.n_feats = len(feats)
    ctx.counts = list(counts)
    ctx.out_hw = list(out_hw)
    ctx.scales = list(spatial_scales)
    ctx.sr = sampling_ratio
    ctx.aligned = aligned
    ctx.needs = [f.requires_grad for f in feats]


def _ra_backward(ctx, grad_outs):
    saved = list(ctx.saved_tensors)
    rois = saved.pop(0)
    lvl = saved.pop(0) if ctx.has[0] else None
    perm = saved.pop(0) if ctx.has[1] else None
    seg = saved.pop(0) if ctx.has[2] else None
    C = ctx.feat_shapes[1]
    gos = []
    for b, g in enumerate(grad_outs):
        if g is None:
            g = torch.zeros((ctx.counts[b], C, ctx.out_hw[2 * b], ctx.out_hw[2 * b + 1]),
                            dtype=torch.float32, device=rois.device)
        gos.append(g)
    grads = roi_align_backward(gos, rois, lvl, perm, seg, ctx.feat_shapes, ctx.feat_cl,
                               ctx.out_hw, ctx.scales, ctx.sr, ctx.aligned)
    grads = [g if need else None for g, need in zip(grads, ctx.needs)]
    return (grads, None, None, None, None, None, None, None, None, None, None)


roi_align_forward.register_autograd(_ra_backward, setup_context=_ra_setup_context)


def multilevel_roi_align(feats, rois, out_sizes, spatial_scales, lvl=None, perm=None, seg=None,
                         counts=None, sampling_ratio=0, aligned=True, channels_last=False):
    """Convenience wrapper: ``out_sizes`` is a list of (h, w) per bucket."""
    out_hw = [int(v) for hw in out_sizes for v in hw]
    if counts is None:
        if len(out_sizes) != 1:
            raise ValueError('counts is required with more than one bucket')
        counts = [rois.size(0)]
    return call(roi_align_forward, list(feats), rois, lvl, perm, seg, [int(c) for c in counts], out_hw,
                [float(s) for s in spatial_scales], int(sampling_ratio), bool(aligned), bool(channels_last),
                grad_inputs=feats)


# --------------------------------------------------------------------------------------------
# dm_paste_masks
# --------------------------------------------------------------------------------------------
@torch.library.custom_op('dynamask::paste_masks', mutates_args=(), device_types='cuda')
def paste_masks(masks: Tensor, boxes: Tensor, labels: Optional[Tensor], img_h: int, img_w: int,
                region: Sequence[int], apply_sigmoid: bool, thr: float, mode: int) -> Tensor:
    """masks [N,C,S_h,S_w] fp32, boxes [N,4]; region = (x_lo, y_lo, x_hi, y_hi).

    mode 0 -> bool [N,h,w] (value >= thr); 1 -> uint8 (value*255); 2 -> float32 raw value."""
    if masks.dim() != 4:
        raise ValueError('masks must be [N,C,S_h,S_w]')
    masks = _f32c(masks, 'masks')
    if masks.stride(3) != 1 or masks.stride(2) != masks.size(3):
        masks = masks.contiguous()
    N, _, sh, sw = masks.shape
    dev = masks.device
    boxes = _f32c(boxes, 'boxes')[:, :4].contiguous()
    if boxes.size(0) != N:
        raise ValueError('boxes must be [N,4]')
    if labels is not None:
        labels = labels.to(torch.int64).contiguous()
    x_lo, y_lo, x_hi, y_hi = [int(v) for v in region]
    dt = {PASTE_BOOL: torch.bool, PASTE_U8: torch.uint8, PASTE_F32: torch.float32}[mode]
    out = torch.empty((N, y_hi - y_lo, x_hi - x_lo), dtype=dt, device=dev)
    if out.numel() == 0:
        return out
    with torch.cuda.device(dev):
        rc = _lib.load().dm_paste_masks(_ptr(masks), masks.stride(0), masks.stride(1),
                                        _ptr(labels), N, sh, sw, int(bool(apply_sigmoid)),
                                        _ptr(boxes), int(img_h), int(img_w), x_lo, y_lo, x_hi,
                                        y_hi, float(thr), int(mode), _ptr(out), _stream(dev))
    _lib.check(rc, 'dm_paste_masks')
    return out


@paste_masks.register_fake
def _(masks, boxes, labels, img_h, img_w, region, apply_sigmoid, thr, mode):
    dt = {PASTE_BOOL: torch.bool, PASTE_U8: torch.uint8, PASTE_F32: torch.float32}[mode]
    return torch.empty((masks.size(0), region[3] - region[1], region[2] - region[0]), dtype=dt,
                       device=masks.device)


# --------------------------------------------------------------------------------------------
# dm_mask_target
# --------------------------------------------------------------------------------------------
@torch.library.custom_op('dynamask::mask_target', mutates_args=(), device_types='cuda')
def mask_target(gt_blob: Tensor, img_offsets: Tensor, img_ghw: Tensor, boxes: Tensor,
                inds: Tensor, roi_img: Optional[Tensor], clip: bool,
                sizes_hw: Sequence[int]) -> List[Tensor]:
    """gt_blob uint8 (all images' [G,H,W] bitmaps back to back); -> one [K,h,w] fp32 per size."""
    if gt_blob.dtype != torch.uint8:
        raise TypeError('gt_blob must be uint8')
    dev = boxes.device
    boxes = _f32c(boxes, 'boxes')[:, :4].contiguous()
    K = boxes.size(0)
    inds = inds.to(torch.int64).contiguous()
    n_sizes = len(sizes_hw) // 2
    outs = [torch.empty((K, int(sizes_hw[2 * s]), int(sizes_hw[2 * s + 1])), dtype=torch.float32,
                        device=dev) for s in range(n_sizes)]
    if K == 0:
        return outs
    B = img_offsets.numel()
    if img_offsets.dtype != torch.int64 or img_ghw.dtype != torch.int32:
        raise TypeError('img_offsets must be int64 and img_ghw int32')
    if roi_img is not None and roi_img.dtype != torch.int32:
        raise TypeError('roi_img must be int32')
    optrs = _arr(ctypes.c_void_p, [o.data_ptr() for o in outs])
    with torch.cuda.device(dev):
        rc = _lib.load().dm_mask_target(_ptr(gt_blob), _ptr(img_offsets), _ptr(img_ghw), B,
                                        _ptr(boxes), _ptr(inds), _ptr(roi_img), K, int(bool(clip)),
                                        _arr(ctypes.c_int32, [int(v) for v in sizes_hw]), n_sizes,
                                        optrs, _stream(dev))
    _lib.check(rc, 'dm_mask_target')
    return outs


@mask_target.register_fake
def _(gt_blob, img_offsets, img_ghw, boxes, inds, roi_img, clip, sizes_hw):
    K = boxes.size(0)
    return [torch.empty((K, sizes_hw[2 * s], sizes_hw[2 * s + 1]), dtype=torch.float32,
                        device=boxes.device) for s in range(len(sizes_hw) // 2)]


# --------------------------------------------------------------------------------------------
# dm_paste_rle / dm_rle_from_canvas / dm_rle_compress_host  
# --------------------------------------------------------------------------------------------
def _rle_finish(N, rh, rw, totals, run_pass2):
    """Shared tail of the two RLE entry points.  The per-instance totals reach the host through
    pinned memory (the first of two synchronisations: the host sizes the buffers), pass 2 writes the
    transitions, ``dm_rle_strings`` turns them into pycocotools strings ON THE DEVICE, and one pinned
    copy brings ``[string offsets | strings]`` back (second synchronisation).  The host only slices
    N ``bytes`` objects out of that blob."""
    import numpy as np
    dev = totals.device
    stream = torch.cuda.current_stream(dev)
    totals_pin = torch.empty(N, dtype=torch.int32, pin_memory=True)
    totals_pin.copy_(totals, non_blocking=True)
    stream.synchronize()
    offsets_pin = torch.empty(N + 1, dtype=torch.int64, pin_memory=True)
    offsets_h = offsets_pin.numpy()
    offsets_h[0] = 0
    np.cumsum(totals_pin.numpy(), out=offsets_h[1:])
    total = int(offsets_h[-1])
    offsets = offsets_pin.to(dev, non_blocking=True)
    trans = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    run_pass2(offsets, trans)
    head = 8 * (N + 1)
    cap = 6 * total + 8 * N + 8
    blob = torch.empty(head + cap, dtype=torch.uint8, device=dev)
    str_offsets = blob[:head].view(torch.int64)
    scratch = torch.empty(max(total, 1) + 2 * N, dtype=torch.int32, device=dev)
    compact, kept, str_len = scratch[:max(total, 1)], scratch[max(total, 1):max(total, 1) + N], scratch[max(total, 1) + N:]
    with torch.cuda.device(dev):
        rc = _lib.load().dm_rle_strings(_ptr(trans), _ptr(offsets), N, int(rh) * int(rw), _ptr(compact), _ptr(kept),
                                        _ptr(str_len), _ptr(str_offsets), ctypes.c_void_p(blob.data_ptr() + head),
                                        _stream(dev))
    _lib.check(rc, 'dm_rle_strings')
    blob_pin = torch.empty(head + cap, dtype=torch.uint8, pin_memory=True)
    blob_pin.copy_(blob, non_blocking=True)
    stream.synchronize()
    so = blob_pin[:head].view(torch.int64).tolist()
    raw = blob_pin[head:head + so[-1]].numpy().tobytes()
    size = [int(rh), int(rw)]
    return [{'size': size, 'counts': raw[so[n]:so[n + 1]]} for n in range(N)]


def _paste_rle_args(masks, boxes, labels, region):
    if masks.dim() != 4:
        raise ValueError('masks must be [N,C,S_h,S_w]')
    if not masks.is_cuda:
        raise NotImplementedError('dynamask::paste_rle has no CPU implementation')
    masks = _f32c(masks, 'masks')
    if masks.stride(3) != 1 or masks.stride(2) != masks.size(3):
        masks = masks.contiguous()
    N = masks.size(0)
    boxes = _f32c(boxes, 'boxes')[:, :4].contiguous()
    if boxes.size(0) != N:
        raise ValueError('boxes must be [N,4]')
    if labels is not None:
        labels = labels.to(torch.int64).contiguous()
    x_lo, y_lo, x_hi, y_hi = [int(v) for v in region]
    return masks, boxes, labels, (x_lo, y_lo, x_hi, y_hi)


def _paste_rle_two_pass(masks, boxes, labels, img_h, img_w, reg, apply_sigmoid, thr):
    """Exact-size form: the host reads the per-instance totals between the two passes (two
    synchronisations).  Used when the capacity guess of :func:`paste_rle_async` was too small."""
    N, _, sh, sw = masks.shape
    dev = masks.device
    x_lo, y_lo, x_hi, y_hi = reg
    rh, rw = y_hi - y_lo, x_hi - x_lo
    col_counts = torch.empty((N, max(rw, 1)), dtype=torch.int32, device=dev)
    totals = torch.zeros(N, dtype=torch.int32, device=dev)

    def run(pass_no, offsets, trans):
        with torch.cuda.device(dev):
            rc = _lib.load().dm_paste_rle(_ptr(masks), masks.stride(0), masks.stride(1), _ptr(labels), N,
                                          sh, sw, int(bool(apply_sigmoid)), _ptr(boxes), int(img_h),
                                          int(img_w), x_lo, y_lo, x_hi, y_hi, float(thr), pass_no,
                                          _ptr(col_counts), _ptr(totals), _ptr(offsets), _ptr(trans),
                                          _stream(dev))
        _lib.check(rc, 'dm_paste_rle')

    run(1, None, None)
    return _rle_finish(N, rh, rw, totals, lambda off, tr: run(2, off, tr))


# transitions per instance the next paste_rle_async call provisions for (adapts to what the masks needed)
# 'slots': record transitions in the counting pass (one row segment) instead of the row-segmented two passes; off --
# the launch waits for its longest column chain, which the segments cut by the window height / 128 whatever the masks
# look like, while recorded slots only pay when NO block of the image overflows them
_RLE_HINT = {'per_inst': 4096, 'str_bytes': 1 << 16, 'slots': False, 'calls': 0}


# Pinned result buffers of paste_rle_async, by size: a buffer goes back here the moment its strings have been
# sliced out (the call's event has completed by then), so a loop over images cycles through two or three buffers
# instead of asking the caching host allocator each time (a miss there is a cudaHostAlloc: 1.2-1.4 ms).
_PIN_POOL = {}


def _pin_get(nbytes):
    free = _PIN_POOL.get(nbytes)
    if free:
        return free.pop()
    return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)


def _pin_put(buf):
    free = _PIN_POOL.setdefault(buf.numel(), [])
    if len(free) < 4:
        free.append(buf)


class PendingRle:
    """RLE strings of one :func:`paste_rle_async` call on their way to the host.  ``result()`` waits
    for the call's event (one synchronisation) and returns the list of COCO RLE dicts."""

    def __init__(self, n, size, event=None, pinned=None, head=0, prefix=0, blob=None, keep=None, redo=None, ready=None,
                 slots=False):
        self.slots = slots
        self.n, self.size, self.event, self.pinned = n, size, event, pinned
        self.head, self.prefix, self.blob, self.keep, self.redo, self._ready = head, prefix, blob, keep, redo, ready

    def result(self):
        if self._ready is not None:
            return self._ready
        self.event.synchronize()
        hdr = self.pinned[:self.head].view(torch.int64)
        status, total = int(hdr[0]) & 1, int(hdr[1])
        if self.slots and (int(hdr[0]) >> 8) > 0:
            _RLE_HINT['slots'] = False   # a block overflowed its slots: the second pass ran at full length anyway
        # provision for 1.5x what the densest image so far needed, in powers of two: the buffer sizes then stay the
        # same from call to call and come out of the caching allocators (no cudaMalloc / cudaHostAlloc per image)
        need = int(1.5 * total / max(self.n, 1)) + 256
        _RLE_HINT['per_inst'] = min(1 << 20, max(_RLE_HINT['per_inst'], 1 << (need - 1).bit_length()))
        if status != 0:
            # more transitions than provisioned: nothing was written, repeat with exact sizes
            self._ready = self.redo()
        else:
            so = hdr[2:3 + self.n].tolist()
            nbytes = so[-1]
            _RLE_HINT['str_bytes'] = max(_RLE_HINT['str_bytes'], 1 << (2 * nbytes - 1).bit_length()) if nbytes else _RLE_HINT['str_bytes']
            if nbytes <= self.prefix:
                raw = self.pinned[self.head:self.head + nbytes].numpy().tobytes()
            else:   # the strings are longer than the prefix that travelled with the header
                rest = torch.empty(nbytes - self.prefix, dtype=torch.uint8, pin_memory=True)
                rest.copy_(self.blob[self.head + self.prefix:self.head + nbytes], non_blocking=True)
                torch.cuda.current_stream(self.blob.device).synchronize()
                raw = self.pinned[self.head:self.head + self.prefix].numpy().tobytes() + rest.numpy().tobytes()
            self._ready = [{'size': self.size, 'counts': raw[so[i]:so[i + 1]]} for i in range(self.n)]
        _pin_put(self.pinned)
        self.blob = self.keep = self.redo = self.pinned = None
        return self._ready


def paste_rle_async(masks: Tensor, boxes: Tensor, labels: Optional[Tensor], img_h: int, img_w: int,
                    region: Sequence[int], apply_sigmoid: bool, thr: float) -> PendingRle:
    """Fused paste -> COCO RLE strings, enqueued without any host synchronisation
    (``dm_paste_rle_strings``: count, device-side scan, write, string building in one call; buffers
    sized from a transition capacity that adapts to the previous calls).  The header, the string
    offsets and the strings leave in ONE pinned copy behind the kernels; ``PendingRle.result()``
    waits for it.  An inference loop can therefore enqueue the next image before it collects this
    one's strings.  Same arguments and results as :func:`paste_rle`."""
    masks, boxes, labels, reg = _paste_rle_args(masks, boxes, labels, region)
    N, _, sh, sw = masks.shape
    dev = masks.device
    x_lo, y_lo, x_hi, y_hi = reg
    rh, rw = y_hi - y_lo, x_hi - x_lo
    size = [int(rh), int(rw)]
    if N == 0:
        return PendingRle(0, size, ready=[])
    lib = _lib.load()
    cap = int(N * _RLE_HINT['per_inst'])
    _RLE_HINT['calls'] += 1
    slots = bool(_RLE_HINT['slots'])
    ws_bytes = int(lib.dm_paste_rle_strings_workspace(N, max(rw, 1), max(rh, 1), cap))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    head = (8 * (N + 3) + 15) & ~15
    out_cap = 6 * cap + 8 * N + 8
    blob = torch.empty(head + out_cap, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.dm_paste_rle_strings(_ptr(masks), masks.stride(0), masks.stride(1), _ptr(labels), N, sh, sw,
                                      int(bool(apply_sigmoid)), _ptr(boxes), int(img_h), int(img_w), x_lo, y_lo,
                                      x_hi, y_hi, float(thr), int(slots), _ptr(ws), cap, _ptr(blob),
                                      ctypes.c_void_p(blob.data_ptr() + head), _stream(dev))
    _lib.check(rc, 'dm_paste_rle_strings')
    prefix = min(out_cap, int(_RLE_HINT['str_bytes']))
    pinned = _pin_get(head + prefix)
    pinned.copy_(blob[:head + prefix], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))
    redo = lambda: _paste_rle_two_pass(masks, boxes, labels, img_h, img_w, reg, apply_sigmoid, thr)  # noqa: E731
    return PendingRle(N, size, ev, pinned, head, prefix, blob, (ws, masks, boxes, labels), redo, slots=slots)


def paste_rle(masks: Tensor, boxes: Tensor, labels: Optional[Tensor], img_h: int, img_w: int,
              region: Sequence[int], apply_sigmoid: bool, thr: float):
    """Fused paste -> COCO RLE: same arguments as :func:`paste_masks` in bool mode, but returns a
    list of N ``{'size': [h, w], 'counts': bytes}`` dicts (what ``pycocotools.mask.encode`` gives for
    each pasted canvas) without ever materialising the canvases.  One host synchronisation."""
    return paste_rle_async(masks, boxes, labels, img_h, img_w, region, apply_sigmoid, thr).result()


def rle_from_canvas(canvas: Tensor):
    """COCO RLE of an existing ``[N,H,W]`` bool / uint8 device canvas (non-zero = foreground)."""
    if canvas.dim() != 3:
        raise ValueError('canvas must be [N,H,W]')
    if not canvas.is_cuda:
        raise NotImplementedError('dynamask::rle_from_canvas has no CPU implementation')
    if canvas.dtype not in (torch.bool, torch.uint8):
        raise TypeError('canvas must be bool or uint8')
    canvas = canvas.contiguous()
    N, H, W = canvas.shape
    dev = canvas.device
    if N == 0:
        return []
    col_counts = torch.empty((N, max(W, 1)), dtype=torch.int32, device=dev)
    totals = torch.zeros(N, dtype=torch.int32, device=dev)

    def run(pass_no, offsets, trans):
        with torch.cuda.device(dev):
            rc = _lib.load().dm_rle_from_canvas(_ptr(canvas), N, H, W, pass_no, _ptr(col_counts), _ptr(totals),
                                                _ptr(offsets), _ptr(trans), _stream(dev))
        _lib.check(rc, 'dm_rle_from_canvas')

    run(1, None, None)
    return _rle_finish(N, H, W, totals, lambda off, tr: run(2, off, tr))


# --------------------------------------------------------------------------------------------
# dm_simple_roi_align_fwd / _bwd  
# --------------------------------------------------------------------------------------------
@torch.library.custom_op('dynamask::simple_roi_align_forward', mutates_args=(), device_types='cuda')
def simple_roi_align_forward(feat: Tensor, rois: Tensor, out_h: int, out_w: int,
                             spatial_scale: float, aligned: bool) -> Tensor:
    """feat [N,C,H,W] fp32, rois [K,5] -> [K,C,out_h,out_w]: one zero-padded bilinear point per bin."""
    if feat.dim() != 4:
        raise ValueError('features must be [N,C,H,W]')
    if rois.dim() != 2 or rois.size(1) != 5:
        raise ValueError('rois must be [K,5]')
    _f32c(feat, 'features')
    rois = _f32c(rois, 'rois').contiguous()
    dev = feat.device
    K = rois.size(0)
    out = torch.empty((K, feat.size(1), int(out_h), int(out_w)), dtype=torch.float32, device=dev)
    if K == 0:
        return out
    with torch.cuda.device(dev):
        sched = _sched_scratch(dev)
        rc = _lib.load().dm_simple_roi_align_fwd(
            _ptr(feat), _arr(ctypes.c_int32, list(feat.shape)), _arr(ctypes.c_int64, list(feat.stride())),
            float(spatial_scale), _ptr(rois), K, int(out_h), int(out_w), _ptr(out),
            _arr(ctypes.c_int64, list(out.stride())), int(bool(aligned)), _ptr(sched), _stream(dev))
    _lib.check(rc, 'dm_simple_roi_align_fwd')
    return out


@simple_roi_align_forward.register_fake
def _(feat, rois, out_h, out_w, spatial_scale, aligned):
    return torch.empty((rois.size(0), feat.size(1), out_h, out_w), dtype=torch.float32, device=feat.device)


@torch.library.custom_op('dynamask::simple_roi_align_backward', mutates_args=(), device_types='cuda')
def simple_roi_align_backward(grad_out: Tensor, rois: Tensor, feat_shape: Sequence[int],
                              spatial_scale: float, aligned: bool) -> Tensor:
    """Gradient w.r.t. the feature map (zero-initialised here, then reduced into)."""
    grad_out = _f32c(grad_out, 'grad_out')
    rois = _f32c(rois, 'rois').contiguous()
    dev = grad_out.device
    grad = torch.empty([int(v) for v in feat_shape], dtype=torch.float32, device=dev)
    K = rois.size(0)
    with torch.cuda.device(dev):
        sched = _sched_scratch(dev)
        rc = _lib.load().dm_simple_roi_align_bwd(
            _ptr(grad), _arr(ctypes.c_int32, list(grad.shape)), _arr(ctypes.c_int64, list(grad.stride())),
            float(spatial_scale), _ptr(rois), K, int(grad_out.size(2)), int(grad_out.size(3)),
            _ptr(grad_out), _arr(ctypes.c_int64, list(grad_out.stride())), int(bool(aligned)), 1,
            _ptr(sched), _stream(dev))
    _lib.check(rc, 'dm_simple_roi_align_bwd')
    return grad


@simple_roi_align_backward.register_fake
def _(grad_out, rois, feat_shape, spatial_scale, aligned):
    return torch.empty([int(v) for v in feat_shape], dtype=torch.float32, device=grad_out.device)


def _sra_setup_context(ctx, inputs, output):
    feat, rois, out_h, out_w, spatial_scale, aligned = inputs
    ctx.save_for_backward(rois)
    ctx.feat_shape = [int(v) for v in feat.shape]
    ctx.scale = spatial_scale
    ctx.aligned = aligned


def _sra_backward(ctx, grad_out):
    (rois,) = ctx.saved_tensors
    g = simple_roi_align_backward(grad_out, rois, ctx.feat_shape, ctx.scale, ctx.aligned)
    return g, None, None, None, None, None


simple_roi_align_forward.register_autograd(_sra_backward, setup_context=_sra_setup_context)


# --------------------------------------------------------------------------------------------
# dm_refine_stages  
# --------------------------------------------------------------------------------------------
def refine_stages_(stage_preds: Sequence[Tensor]) -> Sequence[Tensor]:
    """In-place coarse-to-fine refinement of ``[N,1,S_s,S_s]`` (or ``[N,S_s,S_s]``) stage logits;
    stage 0 is left as is, every later stage is overwritten where the reference overwrites it."""
    if len(stage_preds) < 2:
        return stage_preds
    N = stage_preds[0].size(0)
    dev = stage_preds[0].device
    sizes = []
    for t in stage_preds:
        if not t.is_cuda:
            raise NotImplementedError('dynamask::refine_stages has no CPU implementation')
        _f32c(t, 'stage prediction')
        if t.dim() == 4 and t.size(1) != 1:
            raise ValueError('stage predictions must be class-selected: [N,1,S,S]')
        if t.size(0) != N or not t.is_contiguous():
            raise ValueError('stage predictions must be contiguous [N,1,S,S] tensors of one batch')
        sizes += [int(t.size(-2)), int(t.size(-1))]
    if N == 0:
        return stage_preds
    ptrs = _arr(ctypes.c_void_p, [t.data_ptr() for t in stage_preds])
    outs = _arr(ctypes.c_void_p, [None] + [t.data_ptr() for t in stage_preds[1:]])
    with torch.cuda.device(dev):
        rc = _lib.load().dm_refine_stages(ptrs, _arr(ctypes.c_int32, sizes), len(stage_preds), N, outs,
                                          _stream(dev))
    _lib.check(rc, 'dm_refine_stages')
    return stage_preds


# --------------------------------------------------------------------------------------------
# dm_polygon_target  
# --------------------------------------------------------------------------------------------
@torch.library.custom_op('dynamask::polygon_target', mutates_args=(), device_types='cuda')
def polygon_target(poly_xy: Tensor, vert_offsets: Tensor, obj_poly_offsets: Tensor, img_meta: Tensor,
                   boxes: Tensor, inds: Tensor, roi_img: Optional[Tensor], clip: bool,
                   sizes_hw: Sequence[int]) -> List[Tensor]:
    """Polygon ground truth -> one ``[K,h,w]`` float32 {0,1} target tensor per size."""
    if poly_xy.dtype != torch.float64:
        raise TypeError('poly_xy must be float64')
    if vert_offsets.dtype != torch.int64 or obj_poly_offsets.dtype != torch.int32 or img_meta.dtype != torch.int32:
        raise TypeError('vert_offsets must be int64, obj_poly_offsets and img_meta int32')
    if roi_img is not None and roi_img.dtype != torch.int32:
        raise TypeError('roi_img must be int32')
    dev = boxes.device
    boxes = _f32c(boxes, 'boxes')[:, :4].contiguous()
    K = boxes.size(0)
    inds = inds.to(torch.int64).contiguous()
    n_sizes = len(sizes_hw) // 2
    outs = [torch.empty((K, int(sizes_hw[2 * s]), int(sizes_hw[2 * s + 1])), dtype=torch.float32,
                        device=dev) for s in range(n_sizes)]
    if K == 0:
        return outs
    optrs = _arr(ctypes.c_void_p, [o.data_ptr() for o in outs])
    with torch.cuda.device(dev):
        rc = _lib.load().dm_polygon_target(_ptr(poly_xy), _ptr(vert_offsets), _ptr(obj_poly_offsets),
                                           obj_poly_offsets.numel() - 1, _ptr(img_meta),
                                           img_meta.numel() // 3, _ptr(boxes), _ptr(inds), _ptr(roi_img), K,
                                           int(bool(clip)), _arr(ctypes.c_int32, [int(v) for v in sizes_hw]),
                                           n_sizes, optrs, _stream(dev))
    _lib.check(rc, 'dm_polygon_target')
    return outs


@polygon_target.register_fake
def _(poly_xy, vert_offsets, obj_poly_offsets, img_meta, boxes, inds, roi_img, clip, sizes_hw):
    K = boxes.size(0)
    return [torch.empty((K, sizes_hw[2 * s], sizes_hw[2 * s + 1]), dtype=torch.float32,
                        device=boxes.device) for s in range(len(sizes_hw) // 2)]


# --------------------------------------------------------------------------------------------
# dm_paste_masks_select  
# --------------------------------------------------------------------------------------------
def paste_masks_switched(stage_masks: Sequence[Tensor], bucket: Tensor, boxes: Tensor,
                         labels: Optional[Tensor], img_h: int, img_w: int, apply_sigmoid: bool,
                         thr: float, mode: int) -> Tensor:
    """Paste instance n from ``stage_masks[bucket[n]]`` (each ``[N,C,S_b,S_b]``): one zero fill and
    one window launch per stage, no gather / scatter and no host synchronisation."""
    N = stage_masks[0].size(0)
    dev = stage_masks[0].device
    if not stage_masks[0].is_cuda:
        raise NotImplementedError('dynamask::paste_masks_switched has no CPU implementation')
    if bucket.dtype != torch.int32 or bucket.numel() != N:
        raise TypeError('bucket must be int32 [N]')
    boxes = _f32c(boxes, 'boxes')[:, :4].contiguous()
    if labels is not None:
        labels = labels.to(torch.int64).contiguous()
    dt = {PASTE_BOOL: torch.bool, PASTE_U8: torch.uint8, PASTE_F32: torch.float32}[mode]
    out = torch.empty((N, int(img_h), int(img_w)), dtype=dt, device=dev)
    if out.numel() == 0:
        return out
    bucket = bucket.contiguous()
    with torch.cuda.device(dev):
        for b, m in enumerate(stage_masks):
            m = _f32c(m, 'masks')
            if m.dim() != 4 or m.size(0) != N:
                raise ValueError('every stage must be [N,C,S,S]')
            if m.stride(3) != 1 or m.stride(2) != m.size(3):
                m = m.contiguous()
            rc = _lib.load().dm_paste_masks_select(
                _ptr(m), m.stride(0), m.stride(1), _ptr(labels), N, m.size(2), m.size(3),
                int(bool(apply_sigmoid)), _ptr(boxes), int(img_h), int(img_w), 0, 0, int(img_w), int(img_h),
                float(thr), int(mode), _ptr(bucket), b, 1 if b == 0 else 0, _ptr(out), _stream(dev))
            _lib.check(rc, 'dm_paste_masks_select')
    return out
