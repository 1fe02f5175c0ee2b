"""dynamask_b200 -- B200 (sm_100a) implementation of DynaMask's per-instance mask hot path.

Stages: (1) level + resolution-bucket assignment, (2) multi-level aligned RoIAlign fwd / bwd,
(3) fused sigmoid + paste + threshold, (4) mask-target generation from uint8 bitmaps.  The CUDA
kernels live in ``csrc/`` behind the C ABI of ``include/dynamask_sm100.h``; this package is the
host-side mirror of the reference's plugin surface (mmdet / mmcv names and signatures).
There is no CPU fallback: ops raise if the shared library is missing or tensors are not on CUDA.
"""
from . import ops
from .bbox import bbox2roi
from .mask_heads import (DynaMaskHeadMixin, _do_paste_mask, encode_mask_results, get_seg_masks,
                         get_seg_masks_rle, get_seg_masks_switched, paste_masks_in_image, refine_stage_instance_preds)
from .mask_structures import BitmapMasks, PolygonMasks
from .mask_target import mask_target, mask_target_single, multi_size_mask_targets
from .roi_align import RoIAlign, SimpleRoIAlign, roi_align
from .roi_extractors import (BaseRoIExtractor, BucketedRoIExtractor, BucketedRoIFeats,
                             SingleRoIExtractor)
from .sharding import checksum64, gather_checksums, image_shard, shard_rois
from .switch import get_mask_label, gumbel_softmax
from .switched import forward_switched, simple_test_mask_switched

__version__ = '0.1.0'

__all__ = [
    'ops', 'bbox2roi', 'RoIAlign', 'SimpleRoIAlign', 'roi_align', 'BaseRoIExtractor', 'SingleRoIExtractor',
    'BucketedRoIExtractor', 'BucketedRoIFeats', 'BitmapMasks', 'PolygonMasks', 'mask_target',
    'mask_target_single', 'multi_size_mask_targets', '_do_paste_mask', 'get_seg_masks',
    'paste_masks_in_image', 'get_seg_masks_rle', 'get_seg_masks_switched', 'refine_stage_instance_preds', 'encode_mask_results', 'DynaMaskHeadMixin', 'get_mask_label', 'gumbel_softmax',
    'forward_switched', 'simple_test_mask_switched', 'image_shard', 'shard_rois', 'checksum64', 'gather_checksums'
]
