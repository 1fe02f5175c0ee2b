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
