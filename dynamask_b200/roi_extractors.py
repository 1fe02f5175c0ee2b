"""RoI feature extractors with the reference's plugin surface.

``BaseRoIExtractor`` / ``SingleRoIExtractor`` keep the constructor arguments, attributes and
``forward(feats, rois, roi_scale_factor=None)`` signature of
``mmdet/models/roi_heads/roi_extractors/{base,single_level}_roi_extractor.py``; what changes is
below the surface.  The reference maps RoIs to levels with ~6 elementwise kernels, then for each
level syncs on ``inds.any()``, gathers, runs RoIAlign and scatters back
(``single_level_roi_extractor.py:69-80``).  Here one ``dm_assign`` launch produces the level of
every RoI and one persistent ``dm_roi_align_fwd`` launch reads each RoI straight from its level
and writes it to its own output row: no host synchronisation, no gather / scatter copies.

``BucketedRoIExtractor`` adds the north-star form of stages 1-2: every RoI is pooled at the
output size chosen by the mask-switch module's one-hot label (14/28/56/112,
``mmdet/models/roi_heads/dynamask_roi_head.py:84-114``; selection sketched at ``:197-203``).
"""
from collections import namedtuple

import torch
import torch.nn as nn

from . import ops
from .roi_align import RoIAlign
from .fp16_utils import force_fp32

_LAYER_TYPES = {'RoIAlign': RoIAlign}


class BaseRoIExtractor(nn.Module):
    """Base class for RoI extractors.

    Args:
        roi_layer (dict): RoI layer type and arguments, e.g.
            ``dict(type='RoIAlign', output_size=14, sampling_ratio=0)``.
        out_channels (int): output channels of the RoI layers.
        featmap_strides (list[int]): strides of the input feature maps.
    """

    def __init__(self, roi_layer, out_channels, featmap_strides):
        super().__init__()
        self.roi_layers = self.build_roi_layers(roi_layer, featmap_strides)
        self.out_channels = out_channels
        self.featmap_strides = featmap_strides
        self.fp16_enabled = False

    @property
    def num_inputs(self):
        """int: number of input feature maps."""
        return len(self.featmap_strides)

    def init_weights(self):
        pass

    def build_roi_layers(self, layer_cfg, featmap_strides):
        """One RoI layer per stride, ``spatial_scale = 1 / stride``."""
        cfg = layer_cfg.copy()
        layer_type = cfg.pop('type')
        assert layer_type in _LAYER_TYPES, f'unknown RoI layer type {layer_type}'
        layer_cls = _LAYER_TYPES[layer_type]
        return nn.ModuleList([layer_cls(spatial_scale=1 / s, **cfg) for s in featmap_strides])

    def roi_rescale(self, rois, scale_factor):
        """Scale RoI width / height about the centre; ``[K,5] -> [K,5]``."""
        cx = (rois[:, 1] + rois[:, 3]) * 0.5
        cy = (rois[:, 2] + rois[:, 4]) * 0.5
        new_w = (rois[:, 3] - rois[:, 1]) * scale_factor
        new_h = (rois[:, 4] - rois[:, 2]) * scale_factor
        return torch.stack((rois[:, 0], cx - new_w * 0.5, cy - new_h * 0.5, cx + new_w * 0.5,
                            cy + new_h * 0.5), dim=-1)

    def forward(self, feats, rois, roi_scale_factor=None):
        raise NotImplementedError


class SingleRoIExtractor(BaseRoIExtractor):
    """Extract each RoI's features from the single FPN level its scale maps to.

    - scale < finest_scale * 2: level 0
    - finest_scale * 2 <= scale < finest_scale * 4: level 1
    - finest_scale * 4 <= scale < finest_scale * 8: level 2
    - scale >= finest_scale * 8: level 3

    Args:
        roi_layer (dict): RoI layer type and arguments.
        out_channels (int): output channels of the RoI layers.
        featmap_strides (list[int]): strides of the input feature maps.
        finest_scale (int): scale threshold of mapping to level 0. Default: 56.
    """

    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56):
        super().__init__(roi_layer, out_channels, featmap_strides)
        self.finest_scale = finest_scale

    def map_roi_levels(self, rois, num_levels):
        """``[K,5]`` rois -> ``[K]`` int64 level index (0-based), computed by ``dm_assign``."""
        lvl = ops.call(ops.assign, rois, None, int(num_levels), float(self.finest_scale), 1)[0]
        return lvl.long()

    def _layer_args(self):
        layer = self.roi_layers[0]
        return layer.output_size, layer.sampling_ratio, layer.aligned

    @force_fp32(apply_to=('feats', ), out_fp16=True)
    def forward(self, feats, rois, roi_scale_factor=None):
        out_size, sampling_ratio, aligned = self._layer_args()
        num_levels = len(feats)
        if rois.size(0) == 0:
            return feats[0].new_zeros(0, self.out_channels, *out_size)
        if num_levels == 1:
            return self.roi_layers[0](feats[0], rois)
        lvl = ops.call(ops.assign, rois, None, num_levels, float(self.finest_scale), 1)[0]
        if roi_scale_factor is not None:
            rois = self.roi_rescale(rois, roi_scale_factor)
        scales = [layer.spatial_scale for layer in self.roi_layers[:num_levels]]
        return ops.multilevel_roi_align(feats, rois, [out_size], scales, lvl=lvl,
                                        sampling_ratio=sampling_ratio, aligned=aligned)[0]


BucketedRoIFeats = namedtuple('BucketedRoIFeats',
                              ['feats', 'perm', 'seg_offsets', 'bucket', 'lvl', 'counts'])


class BucketedRoIExtractor(SingleRoIExtractor):
    """Stages 1-2 of the north-star path: pool every RoI at its selected resolution.

    Args:
        bucket_sizes (tuple[int]): pooled size of each resolution bucket, index = argmax of the
            mask-switch one-hot (``stage_sup_size`` of ``dynamask_head.py:146``).
        Remaining args as :class:`SingleRoIExtractor`; ``roi_layer['output_size']`` is the size
        used by the plain ``forward``.
    """

    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56,
                 bucket_sizes=(14, 28, 56, 112)):
        super().__init__(roi_layer, out_channels, featmap_strides, finest_scale)
        self.bucket_sizes = tuple(int(s) for s in bucket_sizes)

    @force_fp32(apply_to=('feats', ), out_fp16=True)
    def forward_bucketed(self, feats, rois, mask_labels, counts=None, channels_last=False):
        """``mask_labels [K, n_buckets]`` one-hot -> :class:`BucketedRoIFeats`.

        ``feats[b]`` is ``[K_b, C, P_b, P_b]`` holding bucket b's RoIs in their original order;
        ``perm[seg_offsets[b]:seg_offsets[b+1]]`` are their indices into ``rois``.  ``counts``
        (RoIs per bucket, host ints) may be passed when already known; otherwise it is read back
        from the device (one 20-byte copy, the only synchronisation of the call).
        """
        _, sampling_ratio, aligned = self._layer_args()
        nb = len(self.bucket_sizes)
        num_levels = len(feats)
        lvl, bucket, perm, seg = ops.call(ops.assign, rois, mask_labels, num_levels,
                                            float(self.finest_scale), nb)
        if counts is None:
            seg_host = seg.cpu()
            counts = (seg_host[1:] - seg_host[:-1]).tolist()
        scales = [layer.spatial_scale for layer in self.roi_layers[:num_levels]]
        outs = ops.multilevel_roi_align(feats, rois, [(s, s) for s in self.bucket_sizes], scales,
                                        lvl=lvl if num_levels > 1 else None, perm=perm, seg=seg,
                                        counts=counts, sampling_ratio=sampling_ratio,
                                        aligned=aligned, channels_last=channels_last)
        return BucketedRoIFeats(outs, perm, seg, bucket, lvl, list(counts))
