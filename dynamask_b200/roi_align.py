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
