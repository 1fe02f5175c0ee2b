"""Mask-head entry points of the hot path: paste-back and target generation.

Reference: ``_do_paste_mask`` (``mmdet/models/roi_heads/mask_heads/fcn_mask_head.py:240-308``),
``DynaMaskHead.get_seg_masks`` / ``get_targets``
(``mmdet/models/roi_heads/mask_heads/dynamask_head.py:279-342`` / ``:246-271``).  The dense
convolutions of the head (``DynaMaskHead.forward``, ``SFMStage``) stay PyTorch and are out of
scope; :class:`DynaMaskHeadMixin` carries the two methods a maintainer mixes into the
reference head.
"""
import numpy as np
import torch

from . import ops
from .mask_target import multi_size_mask_targets

BYTES_PER_FLOAT = 4
# Kept for signature / behaviour parity with fcn_mask_head.py:13-16.  The fused kernel writes
# one byte per pixel and never materialises the fp32 sampling grid, so it does not chunk.
GPU_MEM_LIMIT = 1024**3


def _do_paste_mask(masks, boxes, img_h, img_w, skip_empty=True):
    """Paste instance masks according to boxes.

    Args:
        masks (Tensor): N, 1, H, W probabilities.
        boxes (Tensor): N, 4.
        img_h, img_w (int): canvas size.
        skip_empty (bool): paste only the window that tightly bounds all boxes.

    Returns:
        tuple: (Tensor[N, h', w'] float32, slices) exactly like the reference: the full canvas and
        ``()`` when ``skip_empty`` is False, else the window and its ``(slice_y, slice_x)``.
    """
    if skip_empty:
        # one 16-byte readback: the window decides the *shape* of the returned tensor
        lo = torch.clamp(boxes.min(dim=0).values.floor()[:2] - 1, min=0)
        hi = torch.stack([torch.clamp(boxes[:, 2].max().ceil() + 1, max=img_w),
                          torch.clamp(boxes[:, 3].max().ceil() + 1, max=img_h)])
        x0_int, y0_int, x1_int, y1_int = torch.cat([lo, hi]).to(torch.int32).tolist()
    else:
        x0_int, y0_int = 0, 0
        x1_int, y1_int = int(img_w), int(img_h)
    out = ops.paste_masks(masks.to(torch.float32), boxes, None, int(img_h), int(img_w),
                          [x0_int, y0_int, x1_int, y1_int], False, 0.0, ops.PASTE_F32)
    if skip_empty:
        return out, (slice(y0_int, y1_int), slice(x0_int, x1_int))
    return out, ()


def paste_masks_in_image(mask_pred, det_bboxes, det_labels, mask_thr_binary, ori_shape,
                         scale_factor, rescale):
    """Device-side body of ``get_seg_masks``: logits -> ``[N, img_h, img_w]`` bool (or uint8)."""
    bboxes = det_bboxes[:, :4]
    if rescale:
        img_h, img_w = ori_shape[:2]
    else:
        img_h = np.round(ori_shape[0] * scale_factor).astype(np.int32)
        img_w = np.round(ori_shape[1] * scale_factor).astype(np.int32)
        scale_factor = 1.0
    if not isinstance(scale_factor, (float, torch.Tensor)):
        scale_factor = bboxes.new_tensor(scale_factor)
    bboxes = bboxes / scale_factor
    img_h, img_w = int(img_h), int(img_w)
    labels = det_labels if mask_pred.shape[1] > 1 else None
    if mask_thr_binary >= 0:
        mode, thr = ops.PASTE_BOOL, float(mask_thr_binary)
    else:
        mode, thr = ops.PASTE_U8, 0.0  # for visualization and debugging
    return ops.paste_masks(mask_pred.to(torch.float32), bboxes, labels, img_h, img_w,
                           [0, 0, img_w, img_h], True, thr, mode)


def get_seg_masks(mask_pred, det_bboxes, det_labels, rcnn_test_cfg, ori_shape, scale_factor,
                  rescale):
    """Get segmentation masks from mask_pred and bboxes.

    Args:
        mask_pred (Tensor): (n, #class or 1, h, w) mask logits.
        det_bboxes (Tensor): (n, 4/5).
        det_labels (Tensor): (n, ).
        rcnn_test_cfg: config with ``mask_thr_binary``.
        ori_shape: original image size.
        scale_factor (float | ndarray | Tensor), rescale (bool): as in the reference.

    Returns:
        list[ndarray]: n masks of shape (img_h, img_w), bool (uint8 if ``mask_thr_binary < 0``).
        One device->host copy for the whole batch instead of one per instance.
    """
    im_mask = paste_masks_in_image(mask_pred, det_bboxes, det_labels,
                                   rcnn_test_cfg.mask_thr_binary, ori_shape, scale_factor, rescale)
    host = _to_host(im_mask)
    return [host[i] for i in range(host.shape[0])]


def _to_host(t):
    """One device->host copy into PINNED memory (torch's caching host allocator hands the block back
    on the next call once the previous results are dropped): ~50 GB/s instead of the ~2 GB/s of a
    pageable ``.cpu()`` for the 107 MB of canvases of one 100-detection image.  The returned array is
    a view of the pinned tensor and keeps it alive."""
    if t.numel() == 0:
        return t.cpu().numpy()
    try:
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    except RuntimeError:      # pinned memory exhausted: pageable copy, as the reference does
        return t.cpu().numpy()
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host.numpy()


def get_seg_masks_rle(mask_pred, det_bboxes, det_labels, rcnn_test_cfg, ori_shape, scale_factor,
                      rescale, wait=True):
    """``encode_mask_results`` of ``get_seg_masks`` in one device pass.

    Same arguments as :func:`get_seg_masks`; returns n COCO RLE dicts ``{'size': [h, w], 'counts':
    bytes}`` -- what ``pycocotools.mask.encode`` returns for each pasted mask
    (``mmdet/core/mask/utils.py:36-63``) -- without materialising or copying the ``[n, h, w]``
    canvases: only the run boundaries cross PCIe.  ``wait=False`` returns an ``ops.PendingRle`` instead:
    the call is enqueued without a host synchronisation and ``.result()`` collects the dicts later,
    so a test loop (``mmdet/apis/test.py:24-57``) can enqueue the next image first.
    """
    if rcnn_test_cfg.mask_thr_binary < 0:
        raise ValueError('RLE needs a binary mask: mask_thr_binary must be >= 0')
    bboxes = det_bboxes[:, :4]
    if rescale:
        img_h, img_w = ori_shape[:2]
    else:
        img_h = np.round(ori_shape[0] * scale_factor).astype(np.int32)
        img_w = np.round(ori_shape[1] * scale_factor).astype(np.int32)
        scale_factor = 1.0
    if not isinstance(scale_factor, (float, torch.Tensor)):
        scale_factor = bboxes.new_tensor(scale_factor)
    bboxes = bboxes / scale_factor
    img_h, img_w = int(img_h), int(img_w)
    labels = det_labels if mask_pred.shape[1] > 1 else None
    pending = ops.paste_rle_async(mask_pred.to(torch.float32), bboxes, labels, img_h, img_w,
                                  [0, 0, img_w, img_h], True, float(rcnn_test_cfg.mask_thr_binary))
    return pending.result() if wait else pending


def encode_mask_results(mask_results):
    """Drop-in for ``mmdet.core.mask.utils.encode_mask_results`` (``utils.py:36-63``) for results
    that are still on the device: ``mask_results`` is a per-class list of lists of ``[H, W]`` bool /
    uint8 CUDA tensors (or a ``(segms, scores)`` tuple); returns the same nesting of RLE dicts."""
    if isinstance(mask_results, tuple):
        cls_segms, cls_mask_scores = mask_results
    else:
        cls_segms = mask_results
    flat = [m for segs in cls_segms for m in segs]
    encoded = []
    if flat:
        shapes = {tuple(m.shape) for m in flat}
        if len(shapes) == 1:
            encoded = ops.rle_from_canvas(torch.stack([m.to(torch.bool) for m in flat]))
        else:
            for m in flat:
                encoded.extend(ops.rle_from_canvas(m.to(torch.bool)[None]))
    out, k = [], 0
    for segs in cls_segms:
        out.append(encoded[k:k + len(segs)])
        k += len(segs)
    if isinstance(mask_results, tuple):
        return out, cls_mask_scores
    return out


def get_seg_masks_switched(stage_instance_preds, mask_labels, det_bboxes, det_labels, rcnn_test_cfg,
                           ori_shape, scale_factor, rescale):
    """Switch-driven paste-back: detection j is pasted from the stage its
    mask-switch label selects.

    The reference only sketches this in comments (``dynamask_roi_head.py:176-203``): it pastes all
    four stage predictions of every detection (``chunk_segm_result[idx] = get_seg_masks(stage idx)``,
    four full passes) and then keeps ``chunk_segm_result[argmax(mask_labels[j])][j]``.  Here every
    detection is pasted once.  ``stage_instance_preds``: list of ``[N, C, S_b, S_b]`` logits
    (14/28/56/112); ``mask_labels``: ``[N, n_stages]`` one-hot (``get_mask_label``) or ``[N]``
    integer stage indices.  Returns what ``get_seg_masks`` returns: a list of N numpy masks."""
    from . import ops as _ops
    bboxes = det_bboxes[:, :4]
    if rescale:
        img_h, img_w = ori_shape[:2]
    else:
        img_h = np.round(ori_shape[0] * scale_factor).astype(np.int32)
        img_w = np.round(ori_shape[1] * scale_factor).astype(np.int32)
        scale_factor = 1.0
    if not isinstance(scale_factor, (float, torch.Tensor)):
        scale_factor = bboxes.new_tensor(scale_factor)
    bboxes = bboxes / scale_factor
    if mask_labels.dim() == 2:
        rois = torch.cat([bboxes.new_zeros((bboxes.size(0), 1)), bboxes.float()], 1)
        bucket = _ops.assign(rois, mask_labels.float(), 1, 56.0, mask_labels.size(1))[1]
    else:
        bucket = mask_labels.to(torch.int32)
    labels = det_labels if stage_instance_preds[0].shape[1] > 1 else None
    thr = rcnn_test_cfg.mask_thr_binary
    mode, t = (_ops.PASTE_BOOL, float(thr)) if thr >= 0 else (_ops.PASTE_U8, 0.0)
    im_mask = _ops.paste_masks_switched([p.to(torch.float32) for p in stage_instance_preds], bucket, bboxes.float(),
                                        labels, int(img_h), int(img_w), True, t, mode)
    host = _to_host(im_mask)
    return [host[i] for i in range(host.shape[0])]


def refine_stage_instance_preds(stage_instance_preds):
    """The coarse-to-fine refinement loop of ``DynaMaskRoIHead.simple_test_mask``
    (``mmdet/models/roi_heads/dynamask_roi_head.py:136-148``) as one launch.

    ``stage_instance_preds`` is the list the reference builds at ``:136``
    (``mask_results['stage_instance_preds'][1:]``: the 28 / 56 / 112 logits, each ``[N,1,S,S]``).
    Like the reference it refines the tensors **in place** -- non-boundary pixels of every finer
    stage are overwritten with the bilinearly up-sampled (already refined) coarser stage -- and
    returns the last one, the ``instance_pred`` handed to ``get_seg_masks`` at ``:148-151``.
    Non-contiguous inputs are refined through a contiguous copy that is written back."""
    preds = list(stage_instance_preds)
    if len(preds) == 0:
        raise ValueError('need at least one stage')
    work = [p if p.is_contiguous() else p.contiguous() for p in preds]
    ops.refine_stages_(work)
    for p, w in zip(preds, work):
        if w is not p:
            p.copy_(w)
    return preds[-1]


class DynaMaskHeadMixin(object):
    """``get_targets`` / ``get_seg_masks`` with the reference ``DynaMaskHead`` signatures."""

    stage_sup_size = [14, 28, 56, 112]

    def get_targets(self, pos_bboxes_list, pos_assigned_gt_inds_list, gt_masks_list):
        return multi_size_mask_targets(pos_bboxes_list, pos_assigned_gt_inds_list, gt_masks_list,
                                       self.stage_sup_size)

    def get_seg_masks(self, mask_pred, det_bboxes, det_labels, rcnn_test_cfg, ori_shape,
                      scale_factor, rescale):
        return get_seg_masks(mask_pred, det_bboxes, det_labels, rcnn_test_cfg, ori_shape,
                             scale_factor, rescale)
