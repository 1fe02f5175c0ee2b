"""Pins against the REAL third-party libraries the reference takes part of this path from --
pycocotools (``mmdet/core/mask/utils.py:36-63``, ``mmdet/core/mask/structures.py:561-575``) and mmcv
(``mmcv.ops.point_sample`` / ``SimpleRoIAlign``, ``dynamask_head.py:74,104-105``; ``mmcv.ops.roi_align``).
Neither is in this image (no network), so every test here is skipped with that reason today; the
day one of them is importable these tests compare the oracle's restatements with the library itself
and the rows marked "parity unpinned" in README / DESIGN become pinned.  CPU only.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as O


def _real(name, attr, reason):
    """The real library, not the empty stand-in oracle/ref_shim.py registers under the same name so
    that the unmodified reference can be imported (the stand-ins have no __file__ and no arithmetic)."""
    mod = pytest.importorskip(name, reason=reason)
    if getattr(mod, '__file__', None) is None or not hasattr(mod, attr):
        pytest.skip(reason)
    return mod


def _masks(rng, n=6, h=37, w=53):
    out = []
    for i in range(n):
        m = np.zeros((h, w), np.uint8)
        y0, x0 = rng.integers(0, h - 5), rng.integers(0, w - 5)
        m[y0:y0 + rng.integers(2, h - y0), x0:x0 + rng.integers(2, w - x0)] = 1
        m ^= (rng.random((h, w)) < 0.05).astype(np.uint8)
        out.append(m)
    out.append(np.zeros((h, w), np.uint8))
    out.append(np.ones((h, w), np.uint8))
    return out


def test_rle_restatement_matches_pycocotools():
    mask_util = _real('pycocotools.mask', 'encode', 'pycocotools is not installed in this image')
    rng = np.random.default_rng(3)
    for m in _masks(rng):
        want = mask_util.encode(np.asfortranarray(m))
        got = O.rle_encode(m.astype(bool))
        assert got['size'] == list(want['size'])
        assert got['counts'] == want['counts']
        assert np.array_equal(O.rle_decode(want), mask_util.decode(want).astype(bool))


def test_polygon_rasteriser_matches_pycocotools():
    mask_util = _real('pycocotools.mask', 'encode', 'pycocotools is not installed in this image')
    rng = np.random.default_rng(4)
    import synth
    for polys in synth.make_polygons(12, 96, 128, rng):
        rles = mask_util.frPyObjects([p.tolist() for p in polys], 96, 128)
        want = mask_util.decode(mask_util.merge(rles)).astype(bool)
        assert np.array_equal(O.polygon_to_bitmap(polys, 96, 128).astype(bool), want)


def test_simple_roi_align_restatement_matches_mmcv():
    mmcv_ops = _real('mmcv.ops', 'point_sample', 'mmcv (mmcv-full) is not installed in this image')
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(2, 8, 50, 84, generator=g)
    x1 = torch.rand(20, generator=g) * 900
    y1 = torch.rand(20, generator=g) * 500
    rois = torch.stack([torch.randint(0, 2, (20, ), generator=g).float(), x1, y1,
                        x1 + 8 + torch.rand(20, generator=g) * 400, y1 + 8 + torch.rand(20, generator=g) * 280], 1)
    for size, scale in ((14, 1.0 / 16), (28, 1.0 / 4)):
        want = mmcv_ops.SimpleRoIAlign(size, scale)(feat, rois)
        got = O.simple_roi_align(feat, rois, size, scale)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)


def test_roi_align_restatement_matches_mmcv():
    mmcv_ops = _real('mmcv.ops', 'nms', 'mmcv (mmcv-full) is not installed in this image')
    g = torch.Generator().manual_seed(6)
    feat = torch.randn(2, 4, 40, 60, generator=g)
    x1 = torch.rand(16, generator=g) * 200
    y1 = torch.rand(16, generator=g) * 120
    rois = torch.stack([torch.randint(0, 2, (16, ), generator=g).float(), x1, y1,
                        x1 + torch.rand(16, generator=g) * 100, y1 + torch.rand(16, generator=g) * 80], 1)
    for size, sr in ((7, 0), (14, 2)):
        want = mmcv_ops.roi_align(feat, rois, (size, size), 0.25, sr, 'avg', True)
        got = O.roi_align(feat, rois, (size, size), 0.25, sr, True)
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-6)


def third_party_probe():
    """What bench.py records: which of the two libraries could be imported on this box."""
    found = {}
    for name in ('pycocotools', 'mmcv'):
        try:
            __import__(name)
            found[name] = True
        except Exception:
            found[name] = False
    return found
