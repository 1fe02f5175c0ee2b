"""Switch-driven inference end to end: MaskPre -> argmax -> run only the
stages a detection's label selects -> paste every detection once.

The reference runs all four stages of ``DynaMaskHead`` for every detection and only sketches the
dynamic variant in comments (``mmdet/models/roi_heads/dynamask_roi_head.py:160-204``: paste all
four stage predictions, then keep ``chunk_segm_result[argmax(mask_labels[j])][j]``).  Here the
mask-switch label decides how far each detection travels through the head
(``mmdet/models/roi_heads/mask_heads/dynamask_head.py:222-244``: ``instance_convs``, the three
``SFMStage`` modules at 14 / 28 / 56, ``final_instance_logits`` at 112):

* ``dm_assign`` turns the one-hot labels into bucket indices and a stable grouping ``perm`` of the
  detections by bucket (no host round trip beyond the four bucket counts);
* in that order stage ``s`` only sees the detections whose bucket is ``>= s`` -- a contiguous
  suffix -- so the (PyTorch) stage modules run on shrinking batches;
* ``get_seg_masks_switched`` pastes detection ``j`` from the stage its label selects.

The head modules themselves stay PyTorch (dense convolutions are out of scope, north-star); this
file is host-side plumbing over the reference's module interface.
"""
import torch
import torch.nn.functional as F

from . import ops
from .bbox import bbox2roi
from .mask_heads import get_seg_masks_switched
from .switch import get_mask_label


def forward_switched(mask_head, ins_feats, x, rois, roi_labels, bucket_counts):
    """``DynaMaskHead.forward`` on detections that are already grouped by bucket (ascending).

    ``ins_feats [K,C,14,14]``, ``rois [K,5]``, ``roi_labels [K]`` in grouped order;
    ``bucket_counts``: detections per bucket (host ints, ``len == len(stages) + 1``).
    Returns ``preds``: list over buckets of ``[K_b, 1, S_b, S_b]`` logits -- bucket ``b``'s detections
    at the resolution their label selected.  A detection of bucket ``b`` passes through
    ``instance_convs`` and stages ``0 .. min(b, n_stages - 1)`` only."""
    n_stages = len(mask_head.stages)
    starts = [0]
    for c in bucket_counts:
        starts.append(starts[-1] + int(c))
    feats = ins_feats
    for conv in mask_head.instance_convs:
        feats = conv(feats)
    preds = []
    lo = 0                                        # first grouped detection still travelling
    for idx, stage in enumerate(mask_head.stages):
        # detections of buckets < idx stop before this stage
        cut = starts[idx] - lo
        if cut > 0:
            feats = feats[cut:]
            lo = starts[idx]
        if feats.size(0) == 0:
            preds.append(ins_feats.new_zeros((0, 1) + (mask_head.stage_sup_size[idx], ) * 2))
            continue
        upsample_flag = mask_head.pre_upsample_last_stage or idx < n_stages - 1
        instance_preds, _, feats = stage(feats, x[-idx - 3], rois[lo:], roi_labels[lo:], upsample_flag)
        k_b = starts[idx + 1] - starts[idx]
        preds.append(instance_preds[:k_b])        # the detections that stop here
    # the last bucket goes on to the final logits (dynamask_head.py:233-242)
    cut = starts[n_stages] - lo
    feats = feats[cut:]
    labels = roi_labels[starts[n_stages]:]
    if mask_head.stage_num_classes[-1] == 1:
        labels = labels.clamp(max=0)
    if feats.size(0) > 0:
        final = mask_head.final_instance_logits(feats)[torch.arange(feats.size(0), device=feats.device), labels][:, None]
        if not mask_head.pre_upsample_last_stage:
            final = F.interpolate(final, scale_factor=2, mode='bilinear', align_corners=True)
    else:
        s = mask_head.stage_sup_size[-1]
        final = ins_feats.new_zeros((0, 1, s, s))
    preds.append(final)
    return preds


def simple_test_mask_switched(mask_head, mask_predictor, mask_roi_extractor, semantic_roi_extractor, x, img_metas,
                              det_bboxes, det_labels, test_cfg, rescale=False, interval=100, noise=None):
    """``DynaMaskRoIHead.simple_test_mask`` (``dynamask_roi_head.py:117-158``) with the switch deciding
    the stages each detection runs (the commented variant at ``:160-204``, made real).

    ``mask_predictor`` is the reference's ``MaskPre`` (``base_roi_head.py:10-27``) on the 56x56
    single-level features; ``noise`` optionally fixes the Gumbel noise ``[K,4]`` (tests).  Returns
    ``segm_result``: per class, the list of numpy masks -- the reference's return value."""
    ori_shape = img_metas[0]['ori_shape']
    scale_factor = img_metas[0]['scale_factor']
    num_classes = mask_head.stage_num_classes[0]
    segm_result = [[] for _ in range(num_classes)]
    if det_bboxes.shape[0] == 0:
        return segm_result
    if rescale and not isinstance(scale_factor, float):
        scale_factor = torch.from_numpy(scale_factor).to(det_bboxes.device)
    _bboxes = det_bboxes[:, :4] * scale_factor if rescale else det_bboxes
    mask_rois = bbox2roi([_bboxes])
    nb = len(mask_head.stages) + 1
    for i in range(0, det_labels.shape[0], interval):
        rois = mask_rois[i:i + interval]
        labels = det_labels[i:i + interval]
        sem = semantic_roi_extractor([x[0].detach()], rois)                      # [k,256,56,56]
        logits = mask_predictor(sem)
        mask_labels = get_mask_label(logits, None if noise is None else noise[i:i + interval])
        _, bucket, perm, seg = ops.assign(rois, mask_labels.detach().float().contiguous(), 1, 56.0, nb)
        seg_h = seg.cpu()
        counts = (seg_h[1:] - seg_h[:-1]).tolist()
        order = perm.long()
        g_rois, g_labels = rois[order], labels[order]
        ins = mask_roi_extractor(x[:mask_roi_extractor.num_inputs], g_rois)
        preds = forward_switched(mask_head, ins, x, g_rois, g_labels, counts)
        # back to detection order: stage b's tensor holds its detections, zeros elsewhere (never read)
        k = rois.size(0)
        full = []
        pos = 0
        for b, p in enumerate(preds):
            t = p.new_zeros((k, ) + tuple(p.shape[1:]))
            if p.size(0):
                t[order[pos:pos + p.size(0)]] = p
            pos += p.size(0)
            full.append(t)
        chunk = get_seg_masks_switched(full, bucket, _bboxes[i:i + interval], labels, test_cfg, ori_shape,
                                       scale_factor, rescale)
        for c, segm in zip(labels.tolist(), chunk):
            segm_result[c].append(segm)
    return segm_result
