"""CPU tests: the oracle is pinned against (a) the golden vectors produced by the unmodified
reference (oracle/gen_golden.py), (b) torchvision's CPU roi_align (bit-for-bit), and (c) -- when
/root/reference is present -- the reference's own Python objects run live."""
import os

import numpy as np
import pytest
import torch

import synth
from oracle import oracle as O
from oracle import ref_shim

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
STRIDES = [4, 8, 16, 32]


def gold(name):
    return np.load(os.path.join(GOLD, name))


def bits_equal(a, b):
    a = torch.as_tensor(a, dtype=torch.float32).contiguous()
    b = torch.as_tensor(b, dtype=torch.float32).contiguous()
    return a.shape == b.shape and torch.equal(a.view(torch.int32), b.view(torch.int32))


def test_golden_assign():
    d = gold('assign.npz')
    lvl, bucket, perm, seg = O.assign(d['rois'], d['onehot'], 4, 56)
    assert np.array_equal(lvl, d['lvl'])
    assert np.array_equal(bucket, d['bucket'])
    assert np.array_equal(O.map_roi_levels_c(d['rois'], 4, 56.0, 0), d['lvl'])
    # stable grouping: inside a bucket the original order is kept
    for b in range(4):
        idx = perm[seg[b]:seg[b + 1]]
        assert np.all(np.diff(idx) > 0) and np.all(d['bucket'][idx] == b)


def test_golden_extractor_forward_bit_exact():
    d = gold('extractor.npz')
    feats = [d['feat_l%d' % l] for l in range(4)]
    for p in (7, 14):
        assert bits_equal(O.single_roi_extractor(feats, d['rois'], p, STRIDES), d['out_%d' % p])
    assert bits_equal(O.single_roi_extractor(feats, d['rois'], 7, STRIDES, sampling_ratio=2,
                                             roi_scale_factor=1.25), d['out_7_sr2_rescaled'])
    assert bits_equal(O.single_roi_extractor([feats[0]], d['rois'][:6], 56, [4]), d['out_56_single_level'])
    outs, perm, seg = O.bucketed_extract(feats, d['rois'][:8], d['onehot8'], (14, 28, 56, 112), STRIDES)
    for b in range(4):
        assert bits_equal(outs[b], d['bucket_%d' % b])


def test_golden_extractor_backward():
    d = gold('extractor.npz')
    shapes = [d['feat_l%d' % l].shape for l in range(4)]
    for p in (7, 14):
        grads = O.single_roi_extractor_backward(d['gout_%d' % p], shapes, d['rois'], STRIDES)
        for l in range(4):
            # the reference accumulates RoIs in level-gathered order too, so this is bit-exact
            assert bits_equal(grads[l], d['grad_%d_l%d' % (p, l)])


def test_golden_paste():
    d = gold('paste.npz')
    n = d['logits'].shape[0]
    det = np.concatenate([d['boxes'], np.ones((n, 1), np.float32)], 1)
    labels = np.zeros(n, np.int64)
    segs = O.get_seg_masks(d['logits'], det, labels, 0.5, (120, 160, 3), 1.0, False)
    assert np.array_equal(np.stack(segs), d['segs'])
    sf = np.array([1.5] * 4, np.float32)
    det_rs = det * np.array([1.5, 1.5, 1.5, 1.5, 1.0], np.float32)
    assert np.array_equal(np.stack(O.get_seg_masks(d['logits'], det_rs, labels, 0.5, (120, 160, 3), sf, True)),
                          d['segs_rescaled'])
    assert np.array_equal(np.stack(O.get_seg_masks(d['logits'], det, labels, -1, (120, 160, 3), 1.0, False)),
                          d['segs_u8'])
    vals, sl = O.do_paste_mask(torch.from_numpy(d['logits']).sigmoid(), d['boxes'], 120, 160, skip_empty=False)
    ref = torch.from_numpy(d['values'])
    ok = ~(torch.isnan(vals) | torch.isnan(ref))
    assert sl == () and torch.equal(vals[ok], ref[ok])
    # the separable C restatement agrees with grid_sample to rounding
    c = O.paste_values_c(torch.from_numpy(d['logits']).sigmoid(), d['boxes'], 120, 160)
    ok = ~(torch.isnan(c) | torch.isnan(ref))
    assert float((c[ok] - ref[ok]).abs().max()) < 2e-6
    assert float(((c >= 0.5) == (ref >= 0.5)).float().mean()) >= 0.9999


def test_golden_mask_targets_bit_exact():
    d = gold('mask_target.npz')
    for s in (14, 28, 56, 112):
        t = O.mask_target_single(d['boxes'], d['inds'], d['masks'], s)
        assert torch.equal(t, torch.from_numpy(d['target_%d' % s]))
    assert np.array_equal(O.crop_and_resize(d['masks'], d['boxes'], (28, 28), d['inds']), d['crop_28'])


def test_roi_align_c_matches_torchvision_bitwise():
    g = torch.Generator().manual_seed(3)
    bad = 0
    for trial in range(12):
        h, w = int(torch.randint(8, 40, (1, ), generator=g)), int(torch.randint(8, 40, (1, ), generator=g))
        f = torch.randn(2, 3, h, w, generator=g)
        k = 6
        x1 = torch.rand(k, generator=g) * w * 1.2 - 3
        y1 = torch.rand(k, generator=g) * h * 1.2 - 3
        rois = torch.stack([torch.randint(0, 2, (k, ), generator=g).float(), x1, y1,
                            x1 + torch.rand(k, generator=g) * w, y1 + torch.rand(k, generator=g) * h], 1)
        for p in (2, 7, (3, 5)):
            for sr in (0, 2):
                for sc in (1.0, 0.5):
                    a, b = O.roi_align(f, rois, p, sc, sr, True), O.roi_align_tv(f, rois, p, sc, sr, True)
                    bad += int((a.view(-1).view(torch.int32) != b.view(-1).view(torch.int32)).sum())
    assert bad == 0


def test_roi_align_backward_c_matches_torchvision_autograd():
    import torchvision
    g = torch.Generator().manual_seed(4)
    f = torch.randn(1, 2, 20, 30, generator=g, requires_grad=True)
    r = torch.tensor([[0, 3.3, 2.2, 25.1, 17.9]])
    o = torchvision.ops.roi_align(f, r, (7, 7), 0.5, 0, True)
    go = torch.randn(o.shape, generator=g)
    o.backward(go)
    gi = O.roi_align_backward(go, r, f.shape, 0.5, 0, True)
    assert bits_equal(gi, f.grad)


def test_gumbel_hard_is_one_hot_argmax():
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(64, 4, generator=g)
    u = torch.rand(64, 4, generator=g)
    hard, ind = O.gumbel_softmax_hard(logits, u)
    assert torch.equal(hard.sum(1), torch.ones(64)) and torch.equal(hard.argmax(1), ind)
    assert torch.equal(O.bucket_of(hard), ind)


@pytest.mark.skipif(not ref_shim.available(), reason='reference tree not mounted')
def test_oracle_matches_live_reference():
    ns = ref_shim.load()
    g = torch.Generator().manual_seed(6)
    feats = synth.make_features(2, 4, 256, 384, g)
    rois = synth.make_rois(2, 16, 256, 384, g, s_hi=300.0)
    ext = ns.SingleRoIExtractor(dict(type='RoIAlign', output_size=14, sampling_ratio=0), 4, STRIDES)
    assert torch.equal(ext(feats, rois), O.single_roi_extractor(feats, rois, 14, STRIDES))
    assert torch.equal(ext.map_roi_levels(rois, 4), O.map_roi_levels(rois, 4, 56))
    assert torch.equal(ns.bbox2roi([rois[:3, 1:], rois[3:5, 1:]]), torch.from_numpy(O.bbox2roi([rois[:3, 1:], rois[3:5, 1:]])))
    rng = np.random.default_rng(6)
    masks = synth.make_gt_masks(4, 80, 120, rng)
    pb, pi = synth.jitter_boxes_from_masks(masks, 9, rng)

    class C:
        mask_size = 28
    ref = ns.mask_target([torch.from_numpy(pb)], [torch.from_numpy(pi)], [ns.BitmapMasks(masks, 80, 120)], C)
    assert torch.equal(ref, O.mask_target([pb], [pi], [masks], 28))


def test_simple_roi_align_oracle_matches_closed_form():
    """The mmcv restatement (affine_grid + grid_sample) against the closed form the header states:
    one bilinear point per bin at (x1 + (pw+.5)/P*(x2-x1)) * scale - .5 with zero padding."""
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(2, 3, 19, 27, generator=g)
    rois = torch.tensor([[0, 4., 6., 60., 50.], [1, -10., -5., 200., 100.], [0, 10., 10., 11., 11.],
                         [1, 300., 300., 400., 400.]])
    P, scale = 6, 0.25
    out = O.simple_roi_align(feat, rois, P, scale).numpy()
    H, W = feat.shape[2:]
    f = feat.numpy().astype(np.float64)
    for k in range(rois.size(0)):
        b, x1, y1, x2, y2 = [float(v) for v in rois[k]]
        for ph in range(P):
            for pw in range(P):
                x = (x1 + (pw + .5) / P * (x2 - x1)) * scale - .5
                y = (y1 + (ph + .5) / P * (y2 - y1)) * scale - .5
                x0, y0 = int(np.floor(x)), int(np.floor(y))
                v = np.zeros(3)
                for yy, wy in ((y0, 1 - (y - y0)), (y0 + 1, y - y0)):
                    for xx, wx in ((x0, 1 - (x - x0)), (x0 + 1, x - x0)):
                        if 0 <= yy < H and 0 <= xx < W:
                            v += wy * wx * f[int(b), :, yy, xx]
                assert np.allclose(out[k, :, ph, pw], v, rtol=1e-4, atol=1e-4)
    assert np.all(out[3] == 0)
    grad = O.simple_roi_align_backward(torch.ones(4, 3, P, P), feat.shape, rois, scale)
    # every in-map point spreads total weight <= 1 per channel
    assert float(grad.sum()) <= 4 * 3 * P * P + 1e-3


def test_golden_refine_stages_bit_exact():
    """Oracle restatement of the inference refinement loop against the reference's own source
    lines (tests/golden/refine.npz, made by oracle/gen_golden_next.py)."""
    gd = np.load(os.path.join(GOLD, 'refine.npz'))
    ins = [torch.from_numpy(gd['in_%d' % i]) for i in range(4)]
    out = O.refine_stage_preds(ins[1:])
    assert np.array_equal(out[1].numpy(), gd['out_56'])
    assert np.array_equal(out[2].numpy(), gd['out_112'])
    m28 = ins[1].squeeze(1).sigmoid() >= 0.5
    assert np.array_equal(O.generate_block_target(m28, 1).numpy(), gd['block_target_28'])
    # the 3x3 closed form the kernel uses == the two Laplacian convolutions
    assert np.array_equal(O.non_boundary_3x3(m28.numpy()), gd['block_target_28'] != 1)
    rng = np.random.default_rng(3)
    for shape in ((5, 1, 1), (4, 2, 3), (3, 9, 17), (2, 28, 28)):
        m = rng.random(shape) < 0.6
        assert np.array_equal(O.non_boundary_3x3(m), O.generate_block_target(torch.from_numpy(m), 1).numpy() != 1)


def _golden_polygons():
    gd = np.load(os.path.join(GOLD, 'polygon.npz'))
    xy, voff, ooff = gd['xy'], gd['voff'], gd['ooff']
    objs = [[xy[2 * voff[q]:2 * voff[q + 1]].copy() for q in range(ooff[g], ooff[g + 1])]
            for g in range(len(ooff) - 1)]
    return gd, objs


def test_golden_polygon_targets_bit_exact():
    """Oracle restatement of PolygonMasks.crop_and_resize + mask_target_single against the
    reference's own Python (tests/golden/polygon.npz; the rasteriser underneath is the same C
    restatement of pycocotools in both, so this pins the host arithmetic around it)."""
    gd, objs = _golden_polygons()
    H, W = [int(v) for v in gd['hw']]
    for s in (14, 28, 56, 112):
        t = O.polygon_mask_target_single(gd['boxes'], gd['inds'], objs, H, W, s)
        assert np.array_equal(t.numpy(), gd['target_%d' % s]), s
    full = np.stack([O.polygon_to_bitmap(o, H, W) for o in objs])
    assert np.array_equal(full, gd['full'])


def test_polygon_rasteriser_known_shapes():
    """Properties of pycocotools' rleFrPoly rule that can be stated without the library: an
    integer axis-aligned rectangle fills [x0,x1) x [y0,y1); a polygon covering the canvas fills
    it; vertex order and starting vertex do not matter; an empty polygon list gives zeros."""
    m = O.polygon_to_bitmap([np.array([2., 1, 6, 1, 6, 4, 2, 4])], 8, 10)
    ref = np.zeros((8, 10), bool)
    ref[1:4, 2:6] = True
    assert np.array_equal(m, ref)
    assert O.polygon_to_bitmap([np.array([-3., -2, 20, -2, 20, 30, -3, 30])], 8, 10).all()
    assert not O.polygon_to_bitmap([], 8, 10).any()
    rng = np.random.default_rng(11)
    for _ in range(20):
        p = synth.make_polygons(1, 40, 50, rng, max_parts=1)[0][0]
        a = O.polygon_to_bitmap([p], 40, 50)
        pts = p.reshape(-1, 2)
        assert np.array_equal(a, O.polygon_to_bitmap([np.roll(pts, 3, axis=0).reshape(-1)], 40, 50))
        assert np.array_equal(a, O.polygon_to_bitmap([pts[::-1].reshape(-1)], 40, 50))
        # area close to the shoelace area (clipped polygons excluded)
        if pts.min() > 1 and pts[:, 0].max() < 49 and pts[:, 1].max() < 39:
            x, y = pts[:, 0], pts[:, 1]
            area = 0.5 * abs(np.dot(x, np.roll(y, 1)) - np.dot(y, np.roll(x, 1)))
            per = np.hypot(np.diff(x, append=x[0]), np.diff(y, append=y[0])).sum()
            assert abs(a.sum() - area) <= per + 2
    # union of parts == OR of the parts
    objs = synth.make_polygons(3, 40, 50, rng, max_parts=3)
    for o in objs:
        parts = np.stack([O.polygon_to_bitmap([p], 40, 50) for p in o])
        assert np.array_equal(O.polygon_to_bitmap(o, 40, 50), parts.any(0))


# Known-answer vectors held by the reference's own tests: bitmaps that the REAL pycocotools
# produced for small polygons (reference tests/test_masks.py:339-355, :370-410, :460-470).
# These pin the oracle's restatement of rleFrPoly / merge / decode against the library itself.
_REF_TRUTH1 = np.array(
    [[0, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
     [0, 0, 1, 1, 1, 1, 0, 0, 0, 0], [0, 0, 1, 1, 1, 1, 1, 0, 0, 0],
     [0, 0, 1, 1, 1, 1, 1, 0, 0, 0], [0, 0, 1, 1, 1, 1, 1, 1, 0, 0],
     [0, 0, 0, 1, 1, 1, 1, 0, 0, 0], [0, 0, 0, 0, 1, 0, 0, 0, 0, 0],
     [0, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 0, 0]], np.uint8)
_REF_TRUTH2 = np.array(
    [[0, 1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0], [0, 0, 1, 1, 0, 0],
     [0, 0, 1, 1, 0, 0], [0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0]], np.uint8)
_REF_CROP_TRUTH = np.array([[0, 0, 0], [0, 0, 0], [0, 0, 1], [0, 1, 1]], np.uint8)
_REF_POLY1 = [np.array([1, 1, 3, 1, 4, 3, 2, 4, 1, 3], dtype=np.float64)]
_REF_POLY2 = [np.array([0., 0., 1., 0., 1., 1.]), np.array([1., 1., 2., 1., 2., 2., 1., 2.])]
_REF_POLY_CROP = [np.array([1., 3., 5., 1., 5., 6., 1, 6])]


def test_polygon_rasteriser_matches_reference_known_answers():
    # test_masks.py:370-383 -- 5x5 polygon resized x2 (polygon coordinates doubled), 1 part
    doubled = [p * 2.0 for p in _REF_POLY1]
    assert np.array_equal(O.polygon_to_bitmap(doubled, 10, 10), _REF_TRUTH1.astype(bool))
    # test_masks.py:385-399 -- two parts, union
    assert np.array_equal(O.polygon_to_bitmap([p * 2.0 for p in _REF_POLY2], 6, 6), _REF_TRUTH2.astype(bool))
    # test_masks.py:401-410 -- the 3x3 object rasterised on the 10x10 canvas
    assert np.array_equal(O.polygon_to_bitmap([p * 2.0 for p in _REF_POLY2], 10, 10),
                          np.pad(_REF_TRUTH2, ((0, 4), (0, 4)), 'constant').astype(bool))
    # test_masks.py:460-470 -- crop to [0,0,3,4]: pycocotools clips the boundary
    assert np.array_equal(O.polygon_to_bitmap(_REF_POLY_CROP, 4, 3), _REF_CROP_TRUTH.astype(bool))


def test_polygon_crop_scale_promotion_is_the_numpy2_one():
    """``PolygonMasks.crop_and_resize`` scales vertices by ``out_w / max(w, 0.1)`` with a Python int over a float32
    box extent (``mmdet/core/mask/structures.py:476-486``).  NumPy >= 2 keeps that quotient in float32 -- the form
    the kernel (``__fdiv_rn`` in ``dm_polygon.cu``), the oracle and the golden files use -- while NumPy 1.x promotes it
    to float64, where a vertex within an ulp of a .5 rounding boundary of the x5 grid may land on the other side.
    The bit-exact polygon-target claim is tied to this promotion; this test pins the assumption."""
    import numpy as np
    if int(np.__version__.split('.')[0]) < 2:
        pytest.skip('NumPy 1.x promotes the crop scale to float64: goldens / kernel follow the NumPy >= 2 float32 form')
    w = np.float32(7.3) - np.float32(0.25)
    scale = 14 / max(w, 0.1)
    assert type(scale) is np.float32
    assert scale == np.float32(14.0) / w                      # one correctly rounded float32 division
    assert float(scale) != 14.0 / float(w)                    # ... which is not the float64 quotient
