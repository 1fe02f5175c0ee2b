"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference in the build container.

Imports the reference's own hot-path modules from ``/root/reference`` (read-only) so that
``oracle/gen_golden.py`` can generate golden vectors with the reference's real Python code.
``/root/reference`` does not exist on the GPU box, so nothing at test / bench run time may
import this file; it is used only by ``gen_golden.py`` and by the CPU tests that are
skipped when the reference tree is absent.

The reference hard-depends on ``mmcv==1.0.5`` (``mmdet/__init__.py:19-27``) which is not
installable here (no network).  Following SURVEY.md Appendix B we register *stub* modules
for mmcv and pycocotools; the only arithmetic the stubs supply is
``mmcv.ops.roi_align`` / ``mmcv.ops.RoIAlign`` -> ``torchvision.ops.roi_align(aligned=True)``
(torchvision's CPU C++ kernel; the substitution mmdet itself makes on CPU,
``mmdet/apis/inference.py:102-108``).  Nothing in this file is reference code.
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

REF_ROOT = os.environ.get('DYNAMASK_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'mmdet'))


def _tv_roi_align(input, rois, output_size, spatial_scale=1.0, sampling_ratio=0,
                  pool_mode='avg', aligned=True):
    import torchvision.ops
    assert pool_mode == 'avg', 'stand-in kernel only has avg pooling'
    return torchvision.ops.roi_align(input, rois, _pair(output_size), spatial_scale,
                                     sampling_ratio, aligned)


class _StubRoIAlign(nn.Module):
    """Stand-in for mmcv.ops.RoIAlign (signature recalled from mmcv 1.0.5)."""

    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg',
                 aligned=True, use_torchvision=False):
        super().__init__()
        self.output_size = _pair(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.pool_mode = pool_mode
        self.aligned = aligned
        self.use_torchvision = use_torchvision

    def forward(self, input, rois):
        return _tv_roi_align(input, rois, self.output_size, self.spatial_scale,
                             self.sampling_ratio, self.pool_mode, self.aligned)


class _Dummy(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


class _Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, *a, **k):
        def deco(cls):
            self.module_dict[cls.__name__] = cls
            return cls
        return deco


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg_stub(name, relpath):
    m = types.ModuleType(name)
    m.__path__ = [os.path.join(REF_ROOT, relpath)]
    sys.modules[name] = m
    return m


_loaded = None


def load():
    """Return a namespace with the reference's own classes/functions (loaded once)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference tree not present at %s' % REF_ROOT)
    if 'mmdet' in sys.modules and not getattr(sys.modules['mmdet'], '_dm_shim', False):
        raise RuntimeError('a real mmdet is already imported; shim refuses to shadow it')

    mmcv = _stub('mmcv', __version__='1.0.5')
    ops = _stub('mmcv.ops', RoIAlign=_StubRoIAlign, roi_align=_tv_roi_align,
                DeformConv2dPack=_Dummy, SimpleRoIAlign=_Dummy, Conv2d=nn.Conv2d)
    ops.__path__ = []
    mmcv.ops = ops
    _stub('mmcv.ops.roi_align', roi_align=_tv_roi_align)
    _stub('mmcv.ops.carafe', CARAFEPack=_Dummy)
    _stub('mmcv.cnn', ConvModule=_Dummy, build_upsample_layer=lambda *a, **k: _Dummy())
    pyc = _stub('pycocotools')
    pyc.__path__ = []
    pyc.mask = _stub('pycocotools.mask')

    mmdet = _pkg_stub('mmdet', 'mmdet')
    mmdet._dm_shim = True
    core = _pkg_stub('mmdet.core', 'mmdet/core')
    _pkg_stub('mmdet.core.fp16', 'mmdet/core/fp16')
    _pkg_stub('mmdet.core.mask', 'mmdet/core/mask')
    _pkg_stub('mmdet.core.bbox', 'mmdet/core/bbox')
    _pkg_stub('mmdet.models', 'mmdet/models')
    _pkg_stub('mmdet.models.roi_heads', 'mmdet/models/roi_heads')
    _pkg_stub('mmdet.models.roi_heads.roi_extractors', 'mmdet/models/roi_heads/roi_extractors')
    _pkg_stub('mmdet.models.roi_heads.mask_heads', 'mmdet/models/roi_heads/mask_heads')
    _stub('mmdet.models.builder', ROI_EXTRACTORS=_Registry('roi_extractor'),
          HEADS=_Registry('head'), build_loss=lambda cfg: None,
          build_roi_extractor=lambda cfg: None, build_shared_head=lambda cfg: None)

    dec = importlib.import_module('mmdet.core.fp16.decorators')
    core.force_fp32 = dec.force_fp32
    core.auto_fp16 = dec.auto_fp16
    structures = importlib.import_module('mmdet.core.mask.structures')
    mt = importlib.import_module('mmdet.core.mask.mask_target')
    core.mask_target = mt.mask_target
    core.BitmapMasks = structures.BitmapMasks
    transforms = importlib.import_module('mmdet.core.bbox.transforms')
    core.bbox2roi = transforms.bbox2roi
    sre = importlib.import_module(
        'mmdet.models.roi_heads.roi_extractors.single_level_roi_extractor')
    fcn = importlib.import_module('mmdet.models.roi_heads.mask_heads.fcn_mask_head')
    dyn = importlib.import_module('mmdet.models.roi_heads.mask_heads.dynamask_head')

    ns = types.SimpleNamespace(
        RoIAlign=_StubRoIAlign, roi_align=_tv_roi_align,
        SingleRoIExtractor=sre.SingleRoIExtractor,
        BitmapMasks=structures.BitmapMasks,
        mask_target=mt.mask_target, mask_target_single=mt.mask_target_single,
        bbox2roi=transforms.bbox2roi,
        do_paste_mask=fcn._do_paste_mask, FCNMaskHead=fcn.FCNMaskHead,
        DynaMaskHead=dyn.DynaMaskHead,
    )
    _loaded = ns
    return ns


def load_refine():
    """The inference-time stage refinement of the reference, runnable on plain tensors:
    returns ``(generate_block_target, refine)`` where ``refine(stage_instance_preds)`` executes
    the *source lines* of the loop in ``DynaMaskRoIHead.simple_test_mask``
    (``mmdet/models/roi_heads/dynamask_roi_head.py:136-148``; the method itself needs a whole
    detector, the loop only needs the list of stage logits) and returns the refined last stage."""
    import textwrap
    import torch.nn.functional as F
    load()
    _pkg_stub('mmdet.models.losses', 'mmdet/models/losses')
    sys.modules['mmdet.models.builder'].LOSSES = _Registry('loss')
    cel = importlib.import_module('mmdet.models.losses.cross_entropy_loss')
    path = os.path.join(REF_ROOT, 'mmdet/models/roi_heads/dynamask_roi_head.py')
    lines = open(path).read().split('\n')
    start = next(i for i, l in enumerate(lines) if '# refine instance masks from stage 1' in l)
    end = next(i for i in range(start, len(lines)) if 'instance_pred = stage_instance_preds[-1]' in lines[i])
    src = textwrap.dedent('\n'.join(lines[start:end + 1]))
    code = compile(src, path, 'exec')

    def refine(stage_instance_preds):
        env = dict(mask_results={'stage_instance_preds': list(stage_instance_preds)}, F=F,
                   generate_block_target=cel.generate_block_target, len=len, range=range)
        exec(code, env)
        return env['instance_pred'], env['stage_instance_preds']

    return cel.generate_block_target, refine


def load_polygon():
    """The reference's own ``PolygonMasks`` (``mmdet/core/mask/structures.py:314-558``) and
    ``mask_target_single``, with the absent pycocotools supplied by the oracle's restatement of
    rleFrPoly / merge / decode (``oracle.polygon_to_bitmap``): everything around the rasteriser --
    clip, crop, scale, dtype promotions -- is then the reference's unmodified Python."""
    import numpy as np
    from oracle import oracle as O
    ns = load()
    if not hasattr(np, 'bool'):
        np.bool = bool          # structures.py:574 predates numpy 1.24
    mu = sys.modules['pycocotools.mask']
    # an "RLE" here is just the polygon list + size; merge concatenates, decode rasterises
    mu.frPyObjects = lambda polys, h, w: [dict(polys=[p], size=(h, w)) for p in polys]
    mu.merge = lambda rles: dict(polys=[p for r in rles for p in r['polys']], size=rles[0]['size'])
    mu.decode = lambda rle: O.polygon_to_bitmap(rle['polys'], rle['size'][0], rle['size'][1]).astype(np.uint8)
    structures = sys.modules['mmdet.core.mask.structures']
    return structures.PolygonMasks, ns.mask_target_single
