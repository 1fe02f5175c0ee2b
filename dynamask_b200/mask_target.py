"""Training mask targets (reference: ``mmdet/core/mask/mask_target.py:6-62`` and the four-size
variant ``DynaMaskHead.get_targets``, ``.../mask_heads/dynamask_head.py:246-271``).

The reference walks images x sizes; every step does D2H of the proposals and indices, a clip in
numpy, an H2D upload of all the image's masks, RoIAlign, D2H of the bool result and H2D of its
float copy.  Here the whole batch is one ``dm_mask_target`` launch: proposals and indices never
leave the device, the clip is fused, the bitmaps of the batch are uploaded once per step (one pinned
staging buffer, one copy) and all sizes are produced together.
"""
import torch
from torch.nn.modules.utils import _pair

from . import ops
from .mask_structures import BitmapMasks, PolygonMasks, pack_bitmaps, pack_polygons


def _batched_targets(pos_proposals_list, pos_assigned_gt_inds_list, gt_masks_list, sizes):
    """-> list (per size) of ``[sum K_i, h, w]`` float32 targets, images in order."""
    device = pos_proposals_list[0].device
    if device.type != 'cuda':
        raise NotImplementedError('dynamask_b200 has no CPU path: proposals must be CUDA tensors')
    keep = [i for i, p in enumerate(pos_proposals_list) if p.size(0) > 0]
    sizes_hw = [int(v) for s in sizes for v in _pair(s)]
    if not keep:
        return [pos_proposals_list[0].new_zeros((0, ) + tuple(_pair(s))) for s in sizes]
    if all(isinstance(gt_masks_list[i], PolygonMasks) for i in keep):
        return _batched_polygon_targets(pos_proposals_list, pos_assigned_gt_inds_list, gt_masks_list,
                                        keep, sizes_hw, device)
    for i in keep:
        if not isinstance(gt_masks_list[i], BitmapMasks):
            raise TypeError('dynamask_b200.mask_target handles a batch of BitmapMasks or of '
                            'PolygonMasks; got %s' % type(gt_masks_list[i]).__name__)
    if len(keep) == 1:
        blob, offs, ghw = gt_masks_list[keep[0]].to_device(device)
        roi_img = None
    else:
        # a training step brings new ground truth: the bitmaps of the whole batch go through ONE pinned
        # staging buffer and ONE host -> device copy into one blob the op owns for the launch
        # (pack_bitmaps; no pointer arithmetic between separate allocations, no per-object cache to go
        # stale), with the per-image offsets / shapes beside it
        blob, offs, ghw = pack_bitmaps([gt_masks_list[i].masks for i in keep], device)
        counts = torch.tensor([pos_proposals_list[i].size(0) for i in keep])
        roi_img = torch.repeat_interleave(torch.arange(len(keep), dtype=torch.int32), counts).to(
            device, non_blocking=True)
    boxes = torch.cat([pos_proposals_list[i][:, :4] for i in keep]).float()
    inds = torch.cat([pos_assigned_gt_inds_list[i] for i in keep])
    return ops.mask_target(blob, offs, ghw, boxes, inds, roi_img, True, sizes_hw)


def _batched_polygon_targets(pos_proposals_list, pos_assigned_gt_inds_list, gt_masks_list, keep,
                             sizes_hw, device):
    """Polygon ground truth (the shipped COCO config, ``poly2mask=False``): one
    ``dm_polygon_target`` launch for all images and sizes."""
    if len(keep) == 1:
        xy, voff, ooff, meta = gt_masks_list[keep[0]].to_device(device)
        roi_img = None
    else:
        xy, voff, ooff, meta = pack_polygons([gt_masks_list[i].masks for i in keep],
                                             [(gt_masks_list[i].height, gt_masks_list[i].width) for i in keep],
                                             device)
        counts = torch.tensor([pos_proposals_list[i].size(0) for i in keep])
        roi_img = torch.repeat_interleave(torch.arange(len(keep), dtype=torch.int32), counts).to(
            device, non_blocking=True)
    boxes = torch.cat([pos_proposals_list[i][:, :4] for i in keep]).float()
    inds = torch.cat([pos_assigned_gt_inds_list[i] for i in keep])
    return ops.polygon_target(xy, voff, ooff, meta, boxes, inds, roi_img, True, sizes_hw)


def mask_target_single(pos_proposals, pos_assigned_gt_inds, gt_masks, cfg):
    """Mask targets of one image; ``cfg.mask_size`` int or (h, w). Returns ``[K,h,w]`` float32."""
    mask_size = _pair(cfg.mask_size)
    if pos_proposals.size(0) == 0:
        return pos_proposals.new_zeros((0, ) + mask_size)
    return _batched_targets([pos_proposals], [pos_assigned_gt_inds], [gt_masks], [mask_size])[0]


def mask_target(pos_proposals_list, pos_assigned_gt_inds_list, gt_masks_list, cfg):
    """Mask targets for the positive proposals of a batch, concatenated in image order."""
    if len(pos_proposals_list) == 0:
        return []
    return _batched_targets(pos_proposals_list, pos_assigned_gt_inds_list, gt_masks_list,
                            [cfg.mask_size])[0]


def multi_size_mask_targets(pos_bboxes_list, pos_assigned_gt_inds_list, gt_masks_list,
                            stage_sup_size=(14, 28, 56, 112)):
    """``DynaMaskHead.get_targets``: list over sizes of ``[sum K_i, S, S]`` float32 targets."""
    return _batched_targets(pos_bboxes_list, pos_assigned_gt_inds_list, gt_masks_list,
                            list(stage_sup_size))
