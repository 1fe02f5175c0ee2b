"""Drop-in for ``mmcv.ops.RoIAlign`` / ``mmcv.ops.roi_align`` (mmcv==1.0.5 surface).

Call sites in the reference: ``BaseRoIExtractor.build_roi_layers`` constructs
``RoIAlign(spatial_scale=1/s, **cfg)`` (``.../roi_extractors/base_roi_extractor.py:49-54``) and reads
``.output_size`` as a tuple (``single_level_roi_extractor.py:56-59``);
``BitmapMasks.crop_and_resize`` calls ``roi_align(x, rois, out_shape, 1.0, 0, 'avg', True)``
positionally (``mmdet/core/mask/structures.py:281-282``).
"""
import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from . import ops


def roi_align(input, rois, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg',
              aligned=True):
    """``input [B,C,H,W]`` fp32 CUDA, ``rois [K,5]`` -> ``[K,C,ph,pw]``; differentiable w.r.t. input."""
    if pool_mode != 'avg':
        raise NotImplementedError("dynamask_b200 implements pool_mode='avg' only "
                                  "(the mask path never uses 'max')")
    if rois.size(1) != 5:
        raise AssertionError('RoI must be (idx, x1, y1, x2, y2)!')
    ph, pw = _pair(output_size)
    return ops.multilevel_roi_align([input], rois, [(ph, pw)], [float(spatial_scale)],
                                    sampling_ratio=int(sampling_ratio), aligned=bool(aligned))[0]


class RoIAlign(nn.Module):
    """RoI align pooling layer with the mmcv 1.0.5 constructor.

    Args:
        output_size (int | tuple): pooled (h, w).
        spatial_scale (float): input boxes are scaled by this number.
        sampling_ratio (int): samples per bin and axis; 0 = ceil(roi_size / output_size).
        pool_mode (str): only 'avg'.
        aligned (bool): half-pixel shift (Detectron2 ``aligned=True``).
        use_torchvision (bool): accepted for signature compatibility and ignored -- this
            module always runs the sm_100a kernel.
    """

    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg',
                 aligned=True, use_torchvision=False):
        super().__init__()
        if pool_mode != 'avg':
            raise NotImplementedError("dynamask_b200 implements pool_mode='avg' only")
        self.output_size = _pair(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.pool_mode = pool_mode
        self.aligned = aligned
        self.use_torchvision = use_torchvision

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio,
                         self.pool_mode, self.aligned)

    def __repr__(self):
        return (f'{self.__class__.__name__}(output_size={self.output_size}, '
                f'spatial_scale={self.spatial_scale}, sampling_ratio={self.sampling_ratio}, '
                f'pool_mode={self.pool_mode}, aligned={self.aligned}, '
                f'use_torchvision={self.use_torchvision})')


class SimpleRoIAlign(nn.Module):
    """Drop-in for ``mmcv.ops.SimpleRoIAlign`` as built by ``SFMStage``
    (``mmdet/models/roi_heads/mask_heads/dynamask_head.py:74``: ``SimpleRoIAlign(output_size=out_size,
    spatial_scale=1.0/semantic_out_stride)``) and called at ``:104-105`` with
    ``(semantic_feat [B,C,H,W], rois [K,5])``.

    One bilinear ``grid_sample`` point per output bin (zero padding, ``align_corners = not
    aligned``) at the bin centre of the RoI; no sampling-grid average and no border clamp.  The
    reference materialises a ``[K, P*P, 2]`` point grid per image and calls ``F.grid_sample``
    image by image; here it is one launch of the banded RoIAlign kernels in point mode.  Output
    rows follow the RoI order (identical to mmcv's per-image concatenation for RoIs sorted by image,
    which is what ``bbox2roi`` produces).  Differentiable w.r.t. ``features``.
    """

    def __init__(self, output_size, spatial_scale, aligned=True):
        super().__init__()
        self.output_size = _pair(output_size)
        self.spatial_scale = float(spatial_scale)
        # kept for signature parity with mmcv (the class carries the flag, never reads it)
        self.use_torchvision = False
        self.aligned = aligned

    def forward(self, features, rois):
        if rois.size(1) != 5:
            raise AssertionError('RoI must be (idx, x1, y1, x2, y2)!')
        return ops.simple_roi_align_forward(features, rois, self.output_size[0], self.output_size[1],
                                            self.spatial_scale, bool(self.aligned))

    def __repr__(self):
        return (f'{self.__class__.__name__}(output_size={self.output_size}, '
                f'spatial_scale={self.spatial_scale})')
