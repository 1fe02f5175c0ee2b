"""GPU parity tests: the CUDA path, called through the plugin surface and the C ABI, against the
CPU oracle on the same seeded inputs.  Bars (BASELINE.json): level / bucket assignment bit-exact,
RoIAlign forward 1e-5 relative, backward 1e-4, pasted masks >= 99.99 % pixel agreement, mask
targets bit-exact."""
import math

import numpy as np
import pytest
import torch

import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

FWD_RTOL, FWD_ATOL = 1e-5, 1e-5   # 1e-5 relative; atol covers cancellation around zero (|feat| ~ 1)
BWD_RTOL, BWD_ATOL = 1e-4, 1e-4


def dm():
    import dynamask_b200
    return dynamask_b200


def gen(seed):
    return torch.Generator().manual_seed(seed)


def assert_close(a, b, rtol, atol, what):
    a = a.detach().cpu().float()
    b = b.detach().cpu().float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.numel() == 0:
        return
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bool(bad.any()), '%s: %d / %d outside tolerance, max err %.3e (ref max %.3e)' % (
        what, int(bad.sum()), a.numel(), float(err.max()), float(b.abs().max()))


# ------------------------------------------------------------------------------------------
# stage 1: level / bucket assignment
# ------------------------------------------------------------------------------------------
def test_assign_matches_oracle_on_seeded_rois():
    g = gen(1234)
    rois = synth.make_rois(4, 2048, 800, 1344, g)
    onehot = synth.make_onehot(rois.size(0), g)
    lvl_o, bucket_o, perm_o, seg_o = O.assign(rois, onehot, 4, 56)
    lvl, bucket, perm, seg = dm().ops.assign(rois.cuda(), onehot.cuda(), 4, 56.0, 4)
    assert torch.equal(lvl.cpu().long(), torch.from_numpy(lvl_o))
    assert torch.equal(bucket.cpu().long(), torch.from_numpy(bucket_o))
    assert torch.equal(perm.cpu().long(), torch.from_numpy(perm_o))
    assert torch.equal(seg.cpu().long(), torch.from_numpy(seg_o))


def test_assign_levels_match_reference_expression_on_cuda():
    """Authoritative check for A2: the reference expression evaluated by torch on CUDA tensors,
    including values a few ulp around every level boundary and degenerate boxes."""
    vals = []
    for k in range(0, 6):
        edge = np.float32(56.0 * 2 ** k)
        v = edge
        for _ in range(12):
            v = np.nextafter(v, np.float32(0), dtype=np.float32)
        for _ in range(25):
            vals.append(float(v))
            v = np.nextafter(v, np.float32(1e9), dtype=np.float32)
    vals = np.array(vals, np.float32)
    rois = np.zeros((len(vals) * 3 + 4, 5), np.float32)
    n = len(vals)
    rois[:n, 3] = vals; rois[:n, 4] = vals                      # squares of side v
    rois[n:2 * n, 1] = 3.25; rois[n:2 * n, 2] = 7.5
    rois[n:2 * n, 3] = 3.25 + vals * 2; rois[n:2 * n, 4] = 7.5 + vals / 2   # 2v x v/2
    rois[2 * n:3 * n, 3] = vals * vals; rois[2 * n:3 * n, 4] = 1.0         # v^2 x 1
    rois[3 * n + 0] = (0, 5, 5, 5, 5)         # zero area
    rois[3 * n + 1] = (0, 5, 5, 4, 9)         # negative width -> NaN level
    rois[3 * n + 2] = (0, 9, 9, 3, 2)         # both negative -> positive area
    rois[3 * n + 3] = (0, 0, 0, 1e4, 1e4)     # huge
    g = gen(7)
    rnd = synth.make_rois(1, 200000, 800, 1344, g).numpy()
    rois = np.concatenate([rois, rnd], 0)
    r = torch.from_numpy(rois).cuda()
    ref = O.map_roi_levels(r, 4, 56)          # same torch expression, on the CUDA tensor
    lvl = dm().ops.assign(r, None, 4, 56.0, 1)[0].long()
    nan = torch.isnan(torch.sqrt((r[:, 3] - r[:, 1]) * (r[:, 4] - r[:, 2])))
    assert int(nan.sum()) == 1
    assert torch.equal(lvl[~nan], ref[~nan])
    assert int(lvl[nan][0]) == -1
    mod = dm().SingleRoIExtractor(dict(type='RoIAlign', output_size=7, sampling_ratio=0), 8, [4, 8, 16, 32])
    assert torch.equal(mod.map_roi_levels(r, 4)[~nan], ref[~nan])


def test_assign_stable_grouping_many_rois():
    g = gen(5)
    K = 5000
    rois = synth.make_rois(1, K, 800, 1344, g)
    onehot = synth.make_onehot(K, g, probs=(0.4, 0.3, 0.2, 0.1))
    _, bucket_o, perm_o, seg_o = O.assign(rois, onehot, 4, 56)
    _, bucket, perm, seg = dm().ops.assign(rois.cuda(), onehot.cuda(), 4, 56.0, 4)
    assert torch.equal(perm.cpu().long(), torch.from_numpy(perm_o))
    assert torch.equal(seg.cpu().long(), torch.from_numpy(seg_o))
    assert sorted(perm.cpu().tolist()) == list(range(K))


# ------------------------------------------------------------------------------------------
# stage 2: RoIAlign forward
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('out_size', [7, 14, 28, (5, 3), (4, 12)])
@pytest.mark.parametrize('sampling_ratio', [0, 2])
def test_roi_align_single_level_matches_oracle(out_size, sampling_ratio):
    g = gen(11)
    feat = torch.randn(2, 6, 37, 53, generator=g)
    K = 40
    x1 = torch.rand(K, generator=g) * 240 - 20
    y1 = torch.rand(K, generator=g) * 170 - 20
    w = torch.rand(K, generator=g) * 150
    h = torch.rand(K, generator=g) * 110
    rois = torch.stack([torch.randint(0, 2, (K, ), generator=g).float(), x1, y1, x1 + w, y1 + h], 1)
    for scale in (0.25, 1.0 / 3):
        ref = O.roi_align(feat, rois, out_size, scale, sampling_ratio, True)
        out = dm().roi_align(feat.cuda(), rois.cuda(), out_size, scale, sampling_ratio, 'avg', True)
        assert_close(out, ref, FWD_RTOL, FWD_ATOL, 'roi_align %s sr=%d' % (str(out_size), sampling_ratio))


def test_roi_align_not_aligned_mode():
    g = gen(12)
    feat = torch.randn(1, 4, 30, 30, generator=g)
    rois = torch.tensor([[0, 2.0, 3.0, 2.2, 3.1], [0, 5.0, 5.0, 25.0, 18.0], [0, -4.0, -4.0, 40.0, 40.0]])
    ref = O.roi_align(feat, rois, 7, 1.0, 2, False)
    out = dm().roi_align(feat.cuda(), rois.cuda(), 7, 1.0, 2, 'avg', False)
    assert_close(out, ref, FWD_RTOL, FWD_ATOL, 'aligned=False')


def test_roi_align_edge_rois():
    g = gen(13)
    feat = torch.randn(2, 5, 25, 42, generator=g)
    rois = torch.tensor([
        [0, 10.0, 10.0, 10.0, 10.0],      # zero area
        [0, 30.0, 30.0, 20.0, 40.0],      # negative width
        [1, -500.0, -500.0, -300.0, -300.0],  # entirely outside
        [1, 5000.0, 100.0, 6000.0, 300.0],    # entirely outside (right)
        [0, -100.0, -100.0, 2000.0, 1500.0],  # covers everything
        [1, 0.0, 0.0, 1344.0, 800.0],
        [0, 1300.0, 700.0, 1400.0, 900.0],    # straddles the corner
        [3, 10.0, 10.0, 100.0, 100.0],        # batch index out of range -> zeros
    ])
    ref = O.roi_align(feat, rois, 14, 1 / 32, 0, True)
    out = dm().roi_align(feat.cuda(), rois.cuda(), 14, 1 / 32, 0, 'avg', True)
    assert_close(out, ref, FWD_RTOL, FWD_ATOL, 'edge rois')
    empty = dm().roi_align(feat.cuda(), torch.zeros(0, 5).cuda(), 14, 1 / 32, 0, 'avg', True)
    assert tuple(empty.shape) == (0, 5, 14, 14)


def test_roi_align_huge_bins_take_direct_path():
    """Whole-map RoI pooled to a tiny output at stride 1: bands do not fit shared memory."""
    g = gen(14)
    feat = torch.randn(1, 2, 300, 600, generator=g)
    rois = torch.tensor([[0, 0.0, 0.0, 600.0, 300.0], [0, 20.0, 10.0, 580.0, 290.0]])
    ref = O.roi_align(feat, rois, 2, 1.0, 0, True)
    out = dm().roi_align(feat.cuda(), rois.cuda(), 2, 1.0, 0, 'avg', True)
    assert_close(out, ref, 1e-4, 1e-5, 'huge bins')   # 45 000 samples per bin: looser sum order tolerance
    fcuda = feat.cuda().requires_grad_()
    o = dm().roi_align(fcuda, rois.cuda(), 2, 1.0, 0, 'avg', True)
    go = torch.randn(o.shape, generator=g)
    o.backward(go.cuda())
    gref = O.roi_align_backward(go, rois, feat.shape, 1.0, 0, True)
    assert_close(fcuda.grad, gref, BWD_RTOL, BWD_ATOL, 'huge bins bwd')


def test_roi_align_tall_roi_is_tiled():
    """Single-level stride-4 extractor on a big RoI (the 56x56 semantic extractor case)."""
    g = gen(15)
    feat = torch.randn(1, 3, 200, 336, generator=g)
    rois = torch.tensor([[0, 4.0, 4.0, 1330.0, 790.0], [0, 100.0, 50.0, 700.0, 780.0]])
    for p in (56, 112):
        ref = O.roi_align(feat, rois, p, 0.25, 0, True)
        out = dm().roi_align(feat.cuda(), rois.cuda(), p, 0.25, 0, 'avg', True)
        assert_close(out, ref, FWD_RTOL, FWD_ATOL, 'tall roi P=%d' % p)


def _c2_like(batch, per_img, channels, seed):
    g = gen(seed)
    feats = synth.make_features(batch, channels, 800, 1344, g)
    rois = synth.make_rois(batch, per_img, 800, 1344, g)
    return feats, rois, g


@pytest.mark.parametrize('out_size', [7, 14])
def test_single_roi_extractor_matches_oracle(out_size):
    feats, rois, _ = _c2_like(2, 64, 16, 21)
    ext = dm().SingleRoIExtractor(dict(type='RoIAlign', output_size=out_size, sampling_ratio=0), 16,
                                  [4, 8, 16, 32])
    out = ext([f.cuda() for f in feats], rois.cuda())
    ref = O.single_roi_extractor(feats, rois, out_size, [4, 8, 16, 32])
    assert_close(out, ref, FWD_RTOL, FWD_ATOL, 'SingleRoIExtractor')
    out2 = ext([f.cuda() for f in feats], rois.cuda(), roi_scale_factor=1.3)
    ref2 = O.single_roi_extractor(feats, rois, out_size, [4, 8, 16, 32], roi_scale_factor=1.3)
    assert_close(out2, ref2, FWD_RTOL, FWD_ATOL, 'SingleRoIExtractor rescaled')


def test_single_level_extractor_56():
    feats, rois, _ = _c2_like(1, 24, 8, 22)
    ext = dm().SingleRoIExtractor(dict(type='RoIAlign', output_size=56, sampling_ratio=0), 8, [4])
    out = ext([feats[0].cuda()], rois.cuda())
    ref = O.single_roi_extractor([feats[0]], rois, 56, [4])
    assert_close(out, ref, FWD_RTOL, FWD_ATOL, 'semantic extractor')
    assert ext([feats[0].cuda()], rois[:0].cuda()).shape == (0, 8, 56, 56)


def test_extractor_fp16_guard():
    feats, rois, _ = _c2_like(1, 16, 8, 23)
    ext = dm().SingleRoIExtractor(dict(type='RoIAlign', output_size=7, sampling_ratio=0), 8, [4, 8, 16, 32])
    ext.fp16_enabled = True
    out = ext([f.cuda().half() for f in feats], rois.cuda())
    assert out.dtype == torch.float16
    ref = O.single_roi_extractor([f.half().float() for f in feats], rois, 7, [4, 8, 16, 32])
    assert_close(out.float(), ref, 2e-3, 2e-3, 'fp16 guard')


@pytest.mark.parametrize('channels_last_in', [False, True])
def test_bucketed_extractor_matches_oracle(channels_last_in):
    feats, rois, g = _c2_like(2, 48, 8, 24)
    onehot = synth.make_onehot(rois.size(0), g)
    ext = dm().BucketedRoIExtractor(dict(type='RoIAlign', output_size=14, sampling_ratio=0), 8,
                                    [4, 8, 16, 32])
    fc = [f.cuda() for f in feats]
    if channels_last_in:
        fc = [f.contiguous(memory_format=torch.channels_last) for f in fc]
    res = ext.forward_bucketed(fc, rois.cuda(), onehot.cuda())
    refs, perm_o, seg_o = O.bucketed_extract(feats, rois, onehot, (14, 28, 56, 112), [4, 8, 16, 32])
    assert torch.equal(res.perm.cpu().long(), torch.from_numpy(perm_o))
    assert res.counts == [int(seg_o[i + 1] - seg_o[i]) for i in range(4)]
    for b in range(4):
        assert_close(res.feats[b], refs[b], FWD_RTOL, FWD_ATOL, 'bucket %d' % b)


def test_bucketed_extractor_channels_last_output_and_empty_bucket():
    feats, rois, g = _c2_like(1, 20, 8, 25)
    onehot = torch.zeros(20, 4)
    onehot[:, 1] = 1
    onehot[3, 1] = 0
    onehot[3, 3] = 1
    ext = dm().BucketedRoIExtractor(dict(type='RoIAlign', output_size=14, sampling_ratio=0), 8,
                                    [4, 8, 16, 32])
    res = ext.forward_bucketed([f.cuda() for f in feats], rois.cuda(), onehot.cuda(), channels_last=True)
    refs, _, _ = O.bucketed_extract(feats, rois, onehot, (14, 28, 56, 112), [4, 8, 16, 32])
    assert res.counts == [0, 19, 0, 1]
    for b in range(4):
        assert_close(res.feats[b], refs[b], FWD_RTOL, FWD_ATOL, 'cl bucket %d' % b)


# ------------------------------------------------------------------------------------------
# stage 2: RoIAlign backward
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('out_size', [7, 14, 28])
def test_extractor_backward_matches_oracle(out_size):
    feats, rois, g = _c2_like(2, 40, 8, 31)
    ext = dm().SingleRoIExtractor(dict(type='RoIAlign', output_size=out_size, sampling_ratio=0), 8,
                                  [4, 8, 16, 32])
    fc = [f.cuda().requires_grad_() for f in feats]
    out = ext(fc, rois.cuda())
    go = torch.randn(out.shape, generator=g)
    out.backward(go.cuda())
    refs = O.single_roi_extractor_backward(go, [f.shape for f in feats], rois, [4, 8, 16, 32])
    for l in range(4):
        assert_close(fc[l].grad, refs[l], BWD_RTOL, BWD_ATOL, 'grad level %d' % l)


def test_bucketed_backward_matches_oracle():
    feats, rois, g = _c2_like(1, 24, 4, 32)
    onehot = synth.make_onehot(rois.size(0), g)
    ext = dm().BucketedRoIExtractor(dict(type='RoIAlign', output_size=14, sampling_ratio=0), 4,
                                    [4, 8, 16, 32])
    fc = [f.cuda().requires_grad_() for f in feats]
    res = ext.forward_bucketed(fc, rois.cuda(), onehot.cuda())
    gos = [torch.randn(o.shape, generator=g) for o in res.feats]
    torch.autograd.backward(res.feats, [x.cuda() for x in gos])
    _, _, perm, seg = O.assign(rois, onehot, 4, 56)
    acc = [torch.zeros_like(f) for f in feats]
    for b, p in enumerate((14, 28, 56, 112)):
        idx = torch.from_numpy(perm[seg[b]:seg[b + 1]])
        gs = O.single_roi_extractor_backward(gos[b], [f.shape for f in feats], rois[idx], [4, 8, 16, 32])
        for l in range(4):
            acc[l] += gs[l]
    for l in range(4):
        assert_close(fc[l].grad, acc[l], BWD_RTOL, 2e-4, 'bucketed grad level %d' % l)


def test_roi_align_backward_sampling_ratio_and_odd_sizes():
    g = gen(33)
    feat = torch.randn(2, 3, 37, 53, generator=g)
    K = 20
    x1 = torch.rand(K, generator=g) * 200 - 10
    y1 = torch.rand(K, generator=g) * 140 - 10
    rois = torch.stack([torch.randint(0, 2, (K, ), generator=g).float(), x1, y1,
                        x1 + torch.rand(K, generator=g) * 150, y1 + torch.rand(K, generator=g) * 100], 1)
    for out_size, sr in (((5, 3), 0), (7, 2), ((4, 12), 3)):
        fcuda = feat.cuda().requires_grad_()
        o = dm().roi_align(fcuda, rois.cuda(), out_size, 0.25, sr, 'avg', True)
        go = torch.randn(o.shape, generator=g)
        o.backward(go.cuda())
        ref = O.roi_align_backward(go, rois, feat.shape, 0.25, sr, True)
        assert_close(fcuda.grad, ref, BWD_RTOL, BWD_ATOL, 'bwd %s sr=%d' % (str(out_size), sr))


# ------------------------------------------------------------------------------------------
# stage 3: paste
# ------------------------------------------------------------------------------------------
class _Cfg:
    def __init__(self, thr):
        self.mask_thr_binary = thr


def _dets(n, img_h, img_w, seed):
    g = gen(seed)
    boxes = synth.make_boxes(n, img_h, img_w, g, s_lo=8, s_hi=500)
    logits = synth.make_mask_logits(n, 112, g)
    return logits, boxes


# (431, 637): canvas bytes not a multiple of 16 -> flat tiling with tiles that straddle instances
@pytest.mark.parametrize('shape', [(800, 1333), (427, 640), (431, 637), (37, 23)])
def test_get_seg_masks_pixel_agreement(shape):
    img_h, img_w = shape
    logits, boxes = _dets(30, img_h, img_w, 41)
    det = torch.cat([boxes, torch.ones(30, 1)], 1)
    labels = torch.zeros(30, dtype=torch.long)
    ref = O.get_seg_masks(logits, det, labels, 0.5, (img_h, img_w, 3), 1.0, False)
    out = dm().get_seg_masks(logits.cuda(), det.cuda(), labels.cuda(), _Cfg(0.5), (img_h, img_w, 3), 1.0, False)
    assert len(out) == 30 and out[0].dtype == np.bool_ and out[0].shape == (img_h, img_w)
    agree = sum(int((a == b).sum()) for a, b in zip(out, ref))
    total = 30 * img_h * img_w
    assert agree / total >= 0.9999, agree / total
    assert sum(int(a.sum()) for a in out) > 0


def test_get_seg_masks_rescale_and_multiclass():
    img_h, img_w = 480, 640
    logits, boxes = _dets(12, int(img_h * 1.6), int(img_w * 1.6), 42)
    logits = torch.cat([logits, -logits, logits * 0.5], 1)          # 3 classes
    labels = torch.tensor([0, 1, 2] * 4)
    det = torch.cat([boxes, torch.ones(12, 1)], 1)
    sf = np.array([1.6, 1.6, 1.6, 1.6], np.float32)
    ref = O.get_seg_masks(logits, det, labels, 0.5, (img_h, img_w, 3), sf, True)
    out = dm().get_seg_masks(logits.cuda(), det.cuda(), labels.cuda(), _Cfg(0.5), (img_h, img_w, 3), sf, True)
    agree = sum(int((a == b).sum()) for a, b in zip(out, ref))
    assert agree / (12 * img_h * img_w) >= 0.9999
    # rescale=False with a float scale factor: canvas = round(ori * sf)
    ref2 = O.get_seg_masks(logits[:, :1], det, labels, 0.5, (img_h, img_w, 3), 1.6, False)
    out2 = dm().get_seg_masks(logits[:, :1].cuda(), det.cuda(), labels.cuda(), _Cfg(0.5), (img_h, img_w, 3), 1.6, False)
    assert out2[0].shape == ref2[0].shape == (768, 1024)
    agree = sum(int((a == b).sum()) for a, b in zip(out2, ref2))
    assert agree / (12 * 768 * 1024) >= 0.9999


def test_get_seg_masks_uint8_mode():
    logits, boxes = _dets(6, 200, 300, 43)
    det = torch.cat([boxes, torch.ones(6, 1)], 1)
    labels = torch.zeros(6, dtype=torch.long)
    ref = O.get_seg_masks(logits, det, labels, -1, (200, 300, 3), 1.0, False)
    out = dm().get_seg_masks(logits.cuda(), det.cuda(), labels.cuda(), _Cfg(-1), (200, 300, 3), 1.0, False)
    assert out[0].dtype == np.uint8
    diff = np.abs(np.stack(out).astype(np.int32) - np.stack(ref).astype(np.int32))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3


@pytest.mark.parametrize('skip_empty', [True, False])
def test_do_paste_mask_values(skip_empty):
    logits, boxes = _dets(9, 300, 420, 44)
    boxes[3] = torch.tensor([50.5, 20.0, 50.5, 90.0])       # x1 == x0: the inf -> 0 branch
    boxes[4] = torch.tensor([-30.0, -20.0, 60.0, 50.0])     # partly outside
    prob = logits.sigmoid()
    ref, sl_ref = O.do_paste_mask(prob, boxes, 300, 420, skip_empty=skip_empty)
    out, sl = dm()._do_paste_mask(prob.cuda(), boxes.cuda(), 300, 420, skip_empty=skip_empty)
    assert sl == sl_ref
    out = out.cpu()
    assert out.shape == ref.shape
    ok = ~(torch.isnan(ref) | torch.isnan(out))
    assert float((out[ok] - ref[ok]).abs().max()) < 1e-5
    assert float(ok.float().mean()) > 0.999


# ------------------------------------------------------------------------------------------
# next row: COCO RLE of the pasted masks on the device
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('shape', [(800, 1333), (431, 637), (64, 48)])
def test_get_seg_masks_rle_equals_encoding_of_pasted_masks(shape):
    img_h, img_w = shape
    n = 24
    logits, boxes = _dets(n, img_h, img_w, 51)
    boxes[1] = torch.tensor([-20.0, -30.0, img_w + 15.0, img_h + 25.0])   # covers the canvas: runs cross columns
    boxes[2] = torch.tensor([5.0, -10.0, 40.0, img_h + 10.0])             # full height
    boxes[3] = torch.tensor([3.0, 4.0, 3.0, 30.0])                        # degenerate width
    logits[1] = logits[1].abs() + 1.0                                     # all foreground
    det = torch.cat([boxes, torch.ones(n, 1)], 1).cuda()
    labels = torch.zeros(n, dtype=torch.long).cuda()
    cfg = _Cfg(0.5)
    masks = dm().get_seg_masks(logits.cuda(), det, labels, cfg, (img_h, img_w, 3), 1.0, False)
    rles = dm().get_seg_masks_rle(logits.cuda(), det, labels, cfg, (img_h, img_w, 3), 1.0, False)
    assert len(rles) == n
    for i in range(n):
        assert rles[i]['size'] == [img_h, img_w]
        # the fused kernel evaluates the same arithmetic as the paste kernel: bit-identical masks
        assert np.array_equal(O.rle_decode(rles[i]).astype(bool), masks[i]), i
        assert rles[i]['counts'] == O.rle_encode(masks[i])['counts'], i
    assert int(masks[1].sum()) == img_h * img_w


def test_get_seg_masks_rle_multiclass_and_rescale():
    img_h, img_w = 240, 320
    logits, boxes = _dets(9, int(img_h * 1.5), int(img_w * 1.5), 52)
    logits = torch.cat([logits, -logits], 1)
    labels = torch.tensor([0, 1, 0, 1, 1, 0, 0, 1, 0]).cuda()
    det = torch.cat([boxes, torch.ones(9, 1)], 1).cuda()
    sf = np.array([1.5] * 4, np.float32)
    masks = dm().get_seg_masks(logits.cuda(), det, labels, _Cfg(0.3), (img_h, img_w, 3), sf, True)
    rles = dm().get_seg_masks_rle(logits.cuda(), det, labels, _Cfg(0.3), (img_h, img_w, 3), sf, True)
    for m, r in zip(masks, rles):
        assert r == O.rle_encode(m)
    assert dm().get_seg_masks_rle(logits[:0].cuda(), det[:0], labels[:0], _Cfg(0.3), (img_h, img_w, 3), sf, True) == []


def test_rle_from_canvas_and_encode_mask_results():
    g = gen(53)
    canv = torch.rand(7, 93, 131, generator=g) < torch.rand(7, 1, 1, generator=g)
    canv[0] = False
    canv[1] = True
    canv[2, :, 10:20] = True        # full-height columns: runs continue across column boundaries
    rles = dm().ops.rle_from_canvas(canv.cuda())
    for i in range(7):
        assert rles[i] == O.rle_encode(canv[i].numpy()), i
    u8 = dm().ops.rle_from_canvas((canv.to(torch.uint8) * 255).cuda())
    assert u8 == rles
    # the reference's nesting: per class, a list of masks (mmdet/core/mask/utils.py:36-63)
    per_class = [[canv[0].cuda(), canv[3].cuda()], [], [canv[5].cuda()]]
    enc = dm().encode_mask_results(per_class)
    assert [len(e) for e in enc] == [2, 0, 1]
    assert enc[0][1] == rles[3] and enc[2][0] == rles[5]
    enc2, scores = dm().encode_mask_results((per_class, 'scores'))
    assert enc2 == enc and scores == 'scores'


# ------------------------------------------------------------------------------------------
# stage 4: mask targets
# ------------------------------------------------------------------------------------------
def _target_case(seed, img_h, img_w, g_n, k):
    rng = np.random.default_rng(seed)
    masks = synth.make_gt_masks(g_n, img_h, img_w, rng)
    boxes, inds = synth.jitter_boxes_from_masks(masks, k, rng)
    return masks, boxes, inds


def test_mask_targets_bit_exact_four_sizes():
    masks, boxes, inds = _target_case(51, 320, 480, 7, 48)
    boxes[0] = (-20, -10, 500, 340)         # clipped to the canvas
    boxes[1] = (10, 10, 10, 10)             # empty box
    boxes[2] = (0, 0, 480, 320)             # whole image: up to 35x23 samples per bin at S=14
    bm = dm().BitmapMasks(masks, 320, 480)
    out = dm().multi_size_mask_targets([torch.from_numpy(boxes).cuda()], [torch.from_numpy(inds).cuda()], [bm])
    ref = O.dyna_get_targets([boxes], [inds], [masks])
    for s in range(4):
        assert out[s].dtype == torch.float32
        assert torch.equal(out[s].cpu(), ref[s]), 'size %d: %d mismatches' % (
            s, int((out[s].cpu() != ref[s]).sum()))


def test_mask_target_batch_of_images_and_empty_image():
    cases = [_target_case(52, 200, 304, 5, 20), _target_case(53, 256, 256, 3, 0), _target_case(54, 160, 240, 9, 33)]
    props = [torch.from_numpy(c[1]).cuda() for c in cases]
    inds = [torch.from_numpy(c[2]).cuda() for c in cases]
    bms = [dm().BitmapMasks(c[0], c[0].shape[1], c[0].shape[2]) for c in cases]

    class Cfg:
        mask_size = 28
    out = dm().mask_target(props, inds, bms, Cfg)
    ref = O.mask_target([c[1] for c in cases], [c[2] for c in cases], [c[0] for c in cases], 28)
    assert torch.equal(out.cpu(), ref)
    single = dm().mask_target_single(props[1], inds[1], bms[1], Cfg)
    assert tuple(single.shape) == (0, 28, 28)


def test_bitmapmasks_crop_and_resize_matches_oracle():
    masks, boxes, inds = _target_case(55, 96, 128, 4, 17)
    bm = dm().BitmapMasks(masks, 96, 128)
    res = bm.crop_and_resize(boxes, (56, 56), inds, device='cuda')
    assert isinstance(res, dm().BitmapMasks) and res.masks.dtype == np.bool_
    assert res.height == 56 and res.width == 56 and len(res) == 17
    ref = O.crop_and_resize(masks, boxes, (56, 56), inds)
    assert np.array_equal(res.masks, ref)
    empty = dm().BitmapMasks(np.zeros((0, 96, 128), np.uint8), 96, 128)
    e = empty.crop_and_resize(boxes, (56, 56), inds, device='cuda')
    assert len(e) == 0 and e.height == 56


def test_mask_targets_exact_half_ties():
    """Axis-aligned rectangles produce averages of exactly 0.5; the >= must go the oracle's way."""
    masks = np.zeros((2, 64, 64), np.uint8)
    masks[0, 16:48, 16:48] = 1
    masks[1, :, 32:] = 1
    boxes = np.array([[0, 0, 64, 64], [16, 16, 48, 48], [8, 8, 40, 40], [0, 0, 32, 64], [15.5, 15.5, 47.5, 47.5],
                      [16, 0, 48, 64]], np.float32)
    inds = np.array([0, 0, 0, 1, 0, 1], np.int64)
    bm = dm().BitmapMasks(masks, 64, 64)
    out = dm().multi_size_mask_targets([torch.from_numpy(boxes).cuda()], [torch.from_numpy(inds).cuda()], [bm])
    ref = O.dyna_get_targets([boxes], [inds], [masks])
    for s in range(4):
        assert torch.equal(out[s].cpu(), ref[s])


# ------------------------------------------------------------------------------------------
# golden fixtures produced by the unmodified reference (oracle/gen_golden.py)
# ------------------------------------------------------------------------------------------
import os  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def gold(name):
    return np.load(os.path.join(GOLD, name))


def test_golden_assign_gpu():
    d = gold('assign.npz')
    lvl, bucket, perm, seg = dm().ops.assign(torch.from_numpy(d['rois']).cuda(), torch.from_numpy(d['onehot']).cuda(),
                                             4, 56.0, 4)
    ref = d['lvl']
    nan = np.isnan(np.sqrt((d['rois'][:, 3] - d['rois'][:, 1]) * (d['rois'][:, 4] - d['rois'][:, 2])))
    assert np.array_equal(lvl.cpu().numpy()[~nan], ref[~nan])
    assert np.array_equal(bucket.cpu().numpy(), d['bucket'])
    _, _, perm_o, seg_o = O.assign(d['rois'], d['onehot'], 4, 56)
    assert np.array_equal(perm.cpu().numpy(), perm_o) and np.array_equal(seg.cpu().numpy(), seg_o)


def test_golden_extractor_gpu():
    d = gold('extractor.npz')
    feats = [torch.from_numpy(d['feat_l%d' % l]).cuda() for l in range(4)]
    rois = torch.from_numpy(d['rois']).cuda()
    for p in (7, 14):
        ext = dm().SingleRoIExtractor(dict(type='RoIAlign', output_size=p, sampling_ratio=0), 4, [4, 8, 16, 32])
        fr = [f.clone().requires_grad_() for f in feats]
        out = ext(fr, rois)
        assert_close(out, torch.from_numpy(d['out_%d' % p]), FWD_RTOL, FWD_ATOL, 'golden fwd %d' % p)
        out.backward(torch.from_numpy(d['gout_%d' % p]).cuda())
        for l in range(4):
            assert_close(fr[l].grad, torch.from_numpy(d['grad_%d_l%d' % (p, l)]), BWD_RTOL, BWD_ATOL,
                         'golden grad %d level %d' % (p, l))
    ext = dm().SingleRoIExtractor(dict(type='RoIAlign', output_size=7, sampling_ratio=2), 4, [4, 8, 16, 32])
    assert_close(ext(feats, rois, roi_scale_factor=1.25), torch.from_numpy(d['out_7_sr2_rescaled']),
                 FWD_RTOL, FWD_ATOL, 'golden sr2 rescaled')
    sem = dm().SingleRoIExtractor(dict(type='RoIAlign', output_size=56, sampling_ratio=0), 4, [4])
    assert_close(sem([feats[0]], rois[:6]), torch.from_numpy(d['out_56_single_level']), FWD_RTOL, FWD_ATOL,
                 'golden single level 56')
    bext = dm().BucketedRoIExtractor(dict(type='RoIAlign', output_size=14, sampling_ratio=0), 4, [4, 8, 16, 32])
    res = bext.forward_bucketed(feats, rois[:8], torch.from_numpy(d['onehot8']).cuda())
    for b in range(4):
        assert_close(res.feats[b], torch.from_numpy(d['bucket_%d' % b]), FWD_RTOL, FWD_ATOL, 'golden bucket %d' % b)


def test_golden_paste_gpu():
    d = gold('paste.npz')
    n = d['logits'].shape[0]
    logits = torch.from_numpy(d['logits']).cuda()
    det = torch.cat([torch.from_numpy(d['boxes']), torch.ones(n, 1)], 1).cuda()
    labels = torch.zeros(n, dtype=torch.long).cuda()
    # instance 2 has x1 == x0: the reference's CPU branch (golden) and CUDA branch differ there by
    # construction (see oracle.get_seg_masks); the kernel follows the CUDA branch
    keep = np.array([i for i in range(n) if i != 2])
    out = np.stack(dm().get_seg_masks(logits, det, labels, _Cfg(0.5), (120, 160, 3), 1.0, False))
    assert (out[keep] == d['segs'][keep]).mean() >= 0.9999
    full = np.stack(O.get_seg_masks(d['logits'], det.cpu(), labels.cpu(), 0.5, (120, 160, 3), 1.0, False,
                                    device_mode='gpu'))
    assert (out == full).mean() >= 0.9999
    sf = np.array([1.5] * 4, np.float32)
    det_rs = det * torch.tensor([1.5, 1.5, 1.5, 1.5, 1.0]).cuda()
    out = np.stack(dm().get_seg_masks(logits, det_rs, labels, _Cfg(0.5), (120, 160, 3), sf, True))
    assert (out[keep] == d['segs_rescaled'][keep]).mean() >= 0.9999
    out = np.stack(dm().get_seg_masks(logits, det, labels, _Cfg(-1), (120, 160, 3), 1.0, False))
    # uint8 mode: the golden comes from the reference's CPU branch (skip_empty=True), which only
    # evaluates the integer box +-1 px (fcn_mask_head.py:269-276) and so drops the sub-threshold
    # fringe of half a mask pixel that the CUDA branch keeps.  Compare with the golden inside that
    # region, and with the full-canvas (CUDA-branch) oracle everywhere.
    bx = d['boxes']
    for i in keep:
        xa, ya = max(int(np.floor(bx[i, 0])) - 1, 0), max(int(np.floor(bx[i, 1])) - 1, 0)
        xb, yb = min(int(np.ceil(bx[i, 2])) + 1, 160), min(int(np.ceil(bx[i, 3])) + 1, 120)
        diff = np.abs(out[i, ya:yb, xa:xb].astype(np.int32) - d['segs_u8'][i, ya:yb, xa:xb].astype(np.int32))
        assert diff.size == 0 or diff.max() <= 1, 'uint8 paste instance %d' % i
    full8 = np.stack(O.get_seg_masks(d['logits'], det.cpu(), labels.cpu(), -1, (120, 160, 3), 1.0, False,
                                     device_mode='gpu'))
    assert np.abs(out[keep].astype(np.int32) - full8[keep].astype(np.int32)).max() <= 1
    vals, _ = dm()._do_paste_mask(logits.sigmoid(), det[:, :4], 120, 160, skip_empty=False)
    ref = torch.from_numpy(d['values'])
    ok = ~(torch.isnan(ref) | torch.isnan(vals.cpu()))
    assert float((vals.cpu()[ok] - ref[ok]).abs().max()) < 1e-5


def test_golden_mask_targets_gpu():
    d = gold('mask_target.npz')
    bm = dm().BitmapMasks(d['masks'], 96, 128)
    out = dm().multi_size_mask_targets([torch.from_numpy(d['boxes']).cuda()], [torch.from_numpy(d['inds']).cuda()], [bm])
    for s, size in enumerate((14, 28, 56, 112)):
        assert torch.equal(out[s].cpu(), torch.from_numpy(d['target_%d' % size]))
    assert np.array_equal(bm.crop_and_resize(d['boxes'], (28, 28), d['inds'], device='cuda').masks, d['crop_28'])


# ------------------------------------------------------------------------------------------
# size-independent properties at the benchmark's shapes (256 channels, 800x1344 pyramid)
# ------------------------------------------------------------------------------------------
def _full_size_case(batch, per_img, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    gc = gen(seed)
    shapes = synth.pyramid_shapes(800, 1344)
    feats = [torch.randn(batch, 256, h, w, generator=g, device='cuda') for (h, w) in shapes]
    rois = synth.make_rois(batch, per_img, 800, 1344, gc).cuda()
    onehot = synth.make_onehot(rois.size(0), gc).cuda()
    return feats, rois, onehot


def test_full_size_constant_map_is_reproduced():
    """Interpolation weights sum to one: a constant pyramid pools to the same constant."""
    feats, rois, onehot = _full_size_case(2, 256, 61)
    const = [torch.full_like(f, 3.25) for f in feats]
    ext = dm().BucketedRoIExtractor(dict(type='RoIAlign', output_size=14, sampling_ratio=0), 256, [4, 8, 16, 32])
    res = ext.forward_bucketed(const, rois, onehot)
    assert sum(res.counts) == rois.size(0)
    for o in res.feats:
        assert float((o - 3.25).abs().max()) < 1e-5


def test_full_size_forward_backward_are_adjoint_and_linear():
    """<A f, g> == <f, A^T g> and A(a f1 + f2) == a A f1 + A f2 for the bucketed operator."""
    feats, rois, onehot = _full_size_case(2, 256, 62)
    ext = dm().BucketedRoIExtractor(dict(type='RoIAlign', output_size=14, sampling_ratio=0), 256, [4, 8, 16, 32])
    fr = [f.clone().requires_grad_() for f in feats]
    res = ext.forward_bucketed(fr, rois, onehot)
    g = torch.Generator(device='cuda').manual_seed(63)
    gos = [torch.randn(o.shape, generator=g, device='cuda') for o in res.feats]
    lhs = sum(float((o.double() * go.double()).sum()) for o, go in zip(res.feats, gos))
    torch.autograd.backward(res.feats, gos)
    rhs = sum(float((f.double() * f.grad.double()).sum()) for f in fr)
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs), 1.0), (lhs, rhs)
    del fr, gos
    f2 = [torch.randn(f.shape, generator=g, device='cuda') for f in feats]
    a = ext.forward_bucketed(feats, rois, onehot, counts=res.counts).feats
    b = ext.forward_bucketed(f2, rois, onehot, counts=res.counts).feats
    c = ext.forward_bucketed([0.5 * x + y for x, y in zip(feats, f2)], rois, onehot, counts=res.counts).feats
    for x, y, z in zip(a, b, c):
        assert float((0.5 * x + y - z).abs().max()) < 2e-5


def test_full_size_paste_and_targets_properties():
    # a saturated mask pastes to (approximately) the box: pixel count ~ box area, all inside the box
    n, H, W = 64, 800, 1333
    boxes = synth.make_boxes(n, H, W, gen(64), s_lo=16, s_hi=500).cuda()
    logits = torch.full((n, 1, 112, 112), 20.0, device='cuda')
    out = dm().paste_masks_in_image(logits, boxes, torch.zeros(n, dtype=torch.long).cuda(), 0.5, (H, W, 3), 1.0, False)
    assert out.dtype == torch.bool and tuple(out.shape) == (n, H, W)
    area = ((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])).cpu()
    cnt = out.flatten(1).sum(1).cpu().float()
    perim = 2 * ((boxes[:, 2] - boxes[:, 0]) + (boxes[:, 3] - boxes[:, 1])).cpu()
    assert bool(((cnt - area).abs() <= perim + 4).all())
    # all-ones / all-zero bitmaps give all-ones / all-zero targets for boxes inside the canvas
    ones = dm().BitmapMasks(np.ones((2, 800, 1344), np.uint8), 800, 1344)
    zeros = dm().BitmapMasks(np.zeros((2, 800, 1344), np.uint8), 800, 1344)
    pb = synth.make_boxes(128, 800, 1344, gen(65), s_lo=8, s_hi=700).cuda()
    pi = torch.randint(0, 2, (128, ), generator=gen(66)).cuda()
    for t in dm().multi_size_mask_targets([pb], [pi], [ones]):
        assert float(t.min()) == 1.0
    for t in dm().multi_size_mask_targets([pb], [pi], [zeros]):
        assert float(t.max()) == 0.0


# ------------------------------------------------------------------------------------------
# next row: SimpleRoIAlign -- dynamask_head.py:74,104-105
# ------------------------------------------------------------------------------------------
def _sra_atol(feat):
    """The reference samples through grid_sample: pixel coordinates are rebuilt from NORMALISED
    fp32 coordinates, so they carry ~W * 2^-23 px of rounding noise whose sign depends on ATen
    implementation details (the CPU and CUDA linspace / unnormalise code paths already differ in
    the last ulp).  On N(0,1) features (|gradient| up to ~4 per px) that is 1e-6 * W in the output;
    the float64 closed form is checked as well."""
    return max(FWD_ATOL, 1e-6 * max(feat.shape[2:]))


def _sra_inputs(seed, B=2, C=8, H=50, W=84, K=40, img=(800, 1344)):
    g = gen(seed)
    feat = torch.randn(B, C, H, W, generator=g)
    rois = synth.make_rois(B, K // B, img[0], img[1], g)
    return feat, rois, g


@pytest.mark.parametrize('out_size,scale', [(14, 1.0 / 16), (28, 1.0 / 16), (56, 1.0 / 16), (14, 0.25), ((5, 9), 1.0 / 8)])
def test_simple_roi_align_forward_matches_oracle(out_size, scale):
    # scale 1/4 on a stride-16 map is the reference's own quirk (every SFMStage is built with
    # semantic_out_stride[-1], dynamask_head.py:192): most points then fall outside the map
    feat, rois, _ = _sra_inputs(71)
    layer = dm().SimpleRoIAlign(out_size, scale)
    out = layer(feat.cuda(), rois.cuda())
    ref = O.simple_roi_align(feat, rois, out_size, scale)
    assert_close(out, ref, FWD_RTOL, _sra_atol(feat), 'SimpleRoIAlign fwd %s' % (out_size,))
    # and the kernel is no further from the float64 closed form than the reference's own fp32 path
    exact = O.simple_roi_align_f64(feat, rois, out_size, scale)
    err_k = float((out.cpu().double() - exact).abs().max())
    err_r = float((ref.double() - exact).abs().max())
    assert err_k <= max(2.0 * err_r, 1e-5), (err_k, err_r)


def test_simple_roi_align_edges_and_align_corners():
    g = gen(72)
    feat = torch.randn(2, 4, 23, 31, generator=g)
    rois = torch.tensor([[0, -40., -30., 20., 25.],      # partly left / above the map
                         [1, 100., 60., 400., 300.],     # runs off the right / bottom edge
                         [0, 500., 500., 600., 600.],    # entirely outside -> zeros
                         [1, 10., 10., 10., 10.],        # zero extent: every point on one pixel
                         [0, 30., 20., 12., 8.],         # negative extent
                         [1, 0., 0., 124., 92.]])        # the whole map
    for aligned in (True, False):
        out = dm().SimpleRoIAlign(7, 0.25, aligned=aligned)(feat.cuda(), rois.cuda())
        ref = O.simple_roi_align(feat, rois, 7, 0.25, aligned=aligned)
        assert_close(out, ref, FWD_RTOL, _sra_atol(feat), 'SimpleRoIAlign edges aligned=%s' % aligned)
        assert float(out[2].abs().max()) == 0.0
    assert dm().SimpleRoIAlign(7, 0.25)(feat.cuda(), rois[:0].cuda()).shape == (0, 4, 7, 7)
    fc = feat.cuda().requires_grad_()
    out = dm().SimpleRoIAlign(7, 0.25)(fc, rois.cuda())
    go = torch.randn(out.shape, generator=g)
    out.backward(go.cuda())
    assert_close(fc.grad, O.simple_roi_align_backward(go, feat.shape, rois, 0.25), BWD_RTOL, BWD_ATOL,
                 'SimpleRoIAlign edges grad')


@pytest.mark.parametrize('out_size', [14, 56])
def test_simple_roi_align_backward_matches_oracle(out_size):
    feat, rois, g = _sra_inputs(73, C=4, K=24)
    fc = feat.cuda().requires_grad_()
    out = dm().SimpleRoIAlign(out_size, 1.0 / 16)(fc, rois.cuda())
    go = torch.randn(out.shape, generator=g)
    out.backward(go.cuda())
    ref = O.simple_roi_align_backward(go, feat.shape, rois, 1.0 / 16)
    assert_close(fc.grad, ref, BWD_RTOL, BWD_ATOL, 'SimpleRoIAlign grad')


def test_simple_roi_align_full_size_adjoint():
    # SFMStage sizes: 100 RoIs x 256 ch at 14 / 28 / 56 from the stride 16 / 8 / 4 maps of 800x1344
    g = gen(74)
    rois = synth.make_rois(1, 100, 800, 1344, g).cuda()
    for P, s in ((14, 16), (28, 8), (56, 4)):
        H, W = 800 // s, 1344 // s
        f = torch.randn(1, 256, H, W, generator=g).cuda().requires_grad_()
        out = dm().SimpleRoIAlign(P, 1.0 / s)(f, rois)
        go = torch.randn(out.shape, generator=g).cuda()
        out.backward(go)
        lhs = float((out.double() * go.double()).sum())
        rhs = float((f.grad.double() * f.detach().double()).sum())
        assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0) + 1e-2, (P, lhs, rhs)
        const = dm().SimpleRoIAlign(P, 1.0 / s)(torch.full_like(f, 3.0), rois)
        inside = (rois[:, 1] >= 1) & (rois[:, 2] >= 1) & (rois[:, 3] <= 1343) & (rois[:, 4] <= 799)
        assert_close(const[inside], torch.full_like(const[inside], 3.0), 1e-5, 1e-5, 'constant map')


# ------------------------------------------------------------------------------------------
# next row: fused stage-to-stage refinement -- dynamask_roi_head.py:136-148
# ------------------------------------------------------------------------------------------
def _stage_logits(n, g, sizes=(28, 56, 112), noise=1.5):
    cx = torch.rand(n, generator=g) * 0.6 - 0.3
    cy = torch.rand(n, generator=g) * 0.6 - 0.3
    rr = torch.rand(n, generator=g) * 0.5 + 0.3
    out = []
    for s in sizes:
        lin = (torch.arange(s, dtype=torch.float32) + 0.5) / s * 2 - 1
        d2 = (lin[None, None, :] - cx[:, None, None]) ** 2 + (lin[None, :, None] - cy[:, None, None]) ** 2
        blob = 6.0 * (1.0 - d2 / (rr[:, None, None] ** 2))
        out.append((blob + noise * torch.randn(n, s, s, generator=g))[:, None].contiguous())
    return out


def _assert_refined(got, ref, what):
    """Refined logits: identical up to FMA-free rounding wherever the overwrite decision agrees;
    a decision can only differ where the up-sampled mask is within rounding of 0.5."""
    got, ref = got.cpu(), ref.cpu()
    bad = (got - ref).abs() > 1e-5 + 1e-5 * ref.abs()
    assert float(bad.float().mean()) <= 1e-4, '%s: %d / %d pixels differ' % (what, int(bad.sum()), bad.numel())
    assert float(((got >= 0) != (ref >= 0)).float().mean()) <= 1e-4


def test_golden_refine_gpu():
    gd = np.load(os.path.join(GOLD, 'refine.npz'))
    preds = [torch.from_numpy(gd['in_%d' % i]).cuda() for i in (1, 2, 3)]
    keep0 = preds[0].clone()
    final = dm().refine_stage_instance_preds(preds)
    assert final.data_ptr() == preds[2].data_ptr()          # in place, like the reference
    assert torch.equal(preds[0], keep0)
    _assert_refined(preds[1], torch.from_numpy(gd['out_56']), 'golden 56')
    _assert_refined(final, torch.from_numpy(gd['out_112']), 'golden 112')


@pytest.mark.parametrize('n,sizes', [(100, (28, 56, 112)), (7, (14, 28, 56, 112)), (3, (28, 56)), (5, (10, 25, 33))])
def test_refine_stages_matches_oracle(n, sizes):
    g = gen(81 + n)
    preds = _stage_logits(n, g, sizes)
    ref = O.refine_stage_preds(preds)
    dev = [p.cuda() for p in preds]
    final = dm().refine_stage_instance_preds(dev)
    for s in range(1, len(sizes)):
        _assert_refined(dev[s], ref[s], 'stage %d' % sizes[s])
    assert final is dev[-1]
    # the refinement must actually do something on these inputs
    assert float((ref[-1] != preds[-1]).float().mean()) > 0.3


def test_refine_then_paste_agrees_with_reference_chain():
    """simple_test_mask tail: refine -> get_seg_masks, against the oracle chain."""
    g = gen(91)
    preds = _stage_logits(20, g)
    boxes = synth.make_boxes(20, 300, 400, g, s_hi=250.0)
    det = torch.cat([boxes, torch.ones(20, 1)], 1)
    labels = torch.zeros(20, dtype=torch.long)

    class Cfg:
        mask_thr_binary = 0.5
    ref_final = O.refine_stage_preds(preds)[-1]
    ref = O.get_seg_masks(ref_final, det, labels, 0.5, (300, 400, 3), 1.0, False)
    final = dm().refine_stage_instance_preds([p.cuda() for p in preds])
    out = dm().get_seg_masks(final, det.cuda(), labels.cuda(), Cfg, (300, 400, 3), 1.0, False)
    agree = sum(int((a == b).sum()) for a, b in zip(out, ref)) / (20 * 300 * 400)
    assert agree >= 0.9999, agree


# ------------------------------------------------------------------------------------------
# next row: mask targets from polygon ground truth
# ------------------------------------------------------------------------------------------
def _golden_polygons():
    gd = np.load(os.path.join(GOLD, 'polygon.npz'))
    xy, voff, ooff = gd['xy'], gd['voff'], gd['ooff']
    objs = [[xy[2 * voff[q]:2 * voff[q + 1]].copy() for q in range(ooff[g], ooff[g + 1])]
            for g in range(len(ooff) - 1)]
    return gd, objs


def test_golden_polygon_targets_gpu():
    gd, objs = _golden_polygons()
    H, W = [int(v) for v in gd['hw']]
    pm = dm().PolygonMasks(objs, H, W)
    t = dm().multi_size_mask_targets([torch.from_numpy(gd['boxes']).cuda()],
                                     [torch.from_numpy(gd['inds']).cuda()], [pm])
    for i, s in enumerate((14, 28, 56, 112)):
        assert torch.equal(t[i].cpu(), torch.from_numpy(gd['target_%d' % s])), s
    assert np.array_equal(pm.to_ndarray(), gd['full'].astype(bool))

    class C:
        mask_size = 28
    one = dm().mask_target_single(torch.from_numpy(gd['boxes']).cuda(), torch.from_numpy(gd['inds']).cuda(), pm, C)
    assert torch.equal(one.cpu(), torch.from_numpy(gd['target_28']))
    # reference call chain: crop_and_resize(...).to_ndarray()
    boxes = gd['boxes'].copy()
    boxes[:, [0, 2]] = np.clip(boxes[:, [0, 2]], 0, W)
    boxes[:, [1, 3]] = np.clip(boxes[:, [1, 3]], 0, H)
    arr = pm.crop_and_resize(boxes, (28, 28), gd['inds'], device='cuda').to_ndarray()
    assert np.array_equal(arr, gd['target_28'].astype(bool))


def test_polygon_targets_bit_exact_batch():
    rng = np.random.default_rng(95)
    H, W = 200, 304
    props, inds, pms, refs = [], [], [], [[] for _ in range(4)]
    for b in range(3):
        objs = synth.make_polygons(int(rng.integers(1, 9)), H, W, rng)
        pb, pi = synth.jitter_boxes_from_polygons(objs, 0 if b == 1 else 24, rng)
        if b == 2:
            pb[0] = (-50, -50, 400, 300)       # clipped to the canvas
            pb[1] = (100, 100, 100.2, 100.1)   # tiny box: huge scale, long off-window edges
        props.append(torch.from_numpy(pb).cuda())
        inds.append(torch.from_numpy(pi).cuda())
        pms.append(dm().PolygonMasks(objs, H, W))
        for i, s in enumerate((14, 28, 56, 112)):
            refs[i].append(O.polygon_mask_target_single(pb, pi, objs, H, W, s))
    t = dm().multi_size_mask_targets(props, inds, pms)
    for i in range(4):
        assert torch.equal(t[i].cpu(), torch.cat(refs[i])), i
    frac = float(t[3].mean())
    assert 0.1 < frac < 0.9, frac


def test_polygon_and_bitmap_targets_agree_on_rectangles():
    """Cross-check of the two ground-truth paths on shapes where both are exact."""
    objs = [[np.array([32., 24, 96, 24, 96, 72, 32, 72])]]
    bm = np.zeros((1, 96, 128), np.uint8)
    bm[0, 24:72, 32:96] = 1
    boxes = torch.tensor([[0., 0, 128, 96], [32, 24, 96, 72]]).cuda()
    inds = torch.zeros(2, dtype=torch.long).cuda()
    a = dm().multi_size_mask_targets([boxes], [inds], [dm().PolygonMasks(objs, 96, 128)], (16, 32))
    b = dm().multi_size_mask_targets([boxes], [inds], [dm().BitmapMasks(bm, 96, 128)], (16, 32))
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_polygon_masks_reference_known_answers_gpu():
    """The reference's own PolygonMasks tests (tests/test_masks.py:329-355, :358-410, :447-470):
    bitmaps the real pycocotools produced, reproduced by the device rasteriser."""
    PM = dm().PolygonMasks
    truth1 = np.array(
        [[0, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
         [0, 0, 1, 1, 1, 1, 0, 0, 0, 0], [0, 0, 1, 1, 1, 1, 1, 0, 0, 0],
         [0, 0, 1, 1, 1, 1, 1, 0, 0, 0], [0, 0, 1, 1, 1, 1, 1, 1, 0, 0],
         [0, 0, 0, 1, 1, 1, 1, 0, 0, 0], [0, 0, 0, 0, 1, 0, 0, 0, 0, 0],
         [0, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 0, 0]], np.uint8)
    truth2 = np.array(
        [[0, 1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0], [0, 0, 1, 1, 0, 0],
         [0, 0, 1, 1, 0, 0], [0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0]], np.uint8)
    # rescale / resize with 1 instance 1 part
    raw_masks1 = [[np.array([1, 1, 3, 1, 4, 3, 2, 4, 1, 3], dtype=np.float64)]]
    rescaled = PM(raw_masks1, 5, 5).rescale((12, 10))
    assert (rescaled.height, rescaled.width) == (10, 10)
    assert (rescaled.to_ndarray() == truth1).all()
    resized1 = PM(raw_masks1, 5, 5).resize((10, 10))
    assert resized1.to_ndarray().shape == (1, 10, 10)
    assert (resized1.to_ndarray() == truth1).all()
    # 1 instance 2 parts
    raw_masks2 = [[np.array([0., 0., 1., 0., 1., 1.]), np.array([1., 1., 2., 1., 2., 2., 1., 2.])]]
    resized2 = PM(raw_masks2, 3, 3).resize((6, 6))
    assert (resized2.to_ndarray() == truth2).all()
    # 2 instances
    resized3 = PM([raw_masks1[0], raw_masks2[0]], 5, 5).resize((10, 10))
    truth3 = np.stack([truth1, np.pad(truth2, ((0, 4), (0, 4)), 'constant')])
    assert (resized3.to_ndarray() == truth3).all()
    # empty
    assert PM([], 28, 28).resize((56, 72)).to_ndarray().shape == (0, 56, 72)
    # crop
    cropped = PM([[np.array([1., 3., 5., 1., 5., 6., 1, 6])]], 7, 7).crop(np.array([0, 0, 3, 4]))
    assert (cropped.height, cropped.width) == (4, 3)
    assert (cropped.to_ndarray() == np.array([[0, 0, 0], [0, 0, 0], [0, 0, 1], [0, 1, 1]])).all()
    # flip involution (test_masks.py:413-444)
    rng = np.random.default_rng(4)
    pm = PM(synth.make_polygons(3, 28, 28, rng), 28, 28)
    for d in ('horizontal', 'vertical'):
        assert (pm.to_ndarray() == pm.flip(d).flip(d).to_ndarray()).all()
    # crop_and_resize: shape / len contract (test_masks.py:506-528)
    out = pm.crop_and_resize(np.array([[2., 3, 20, 25], [0, 0, 28, 28]], np.float32), (56, 72), [0, 2])
    assert len(out) == 2 and (out.height, out.width) == (56, 72)
    assert out.to_ndarray().shape == (2, 56, 72)


# ------------------------------------------------------------------------------------------
# next row: switch-driven paste-back -- dynamask_roi_head.py:176-203 (comments)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('as_onehot', [True, False])
def test_get_seg_masks_switched_selects_stage_per_detection(as_onehot):
    g = gen(101)
    n = 24
    stages = [synth.make_mask_logits(n, s, g) for s in (14, 28, 56, 112)]
    boxes = synth.make_boxes(n, 240, 320, g, s_hi=200.0)
    det = torch.cat([boxes, torch.ones(n, 1)], 1)
    labels = torch.zeros(n, dtype=torch.long)
    onehot = synth.make_onehot(n, g)
    pick = onehot.argmax(1)

    class Cfg:
        mask_thr_binary = 0.5
    # the reference's sketch: paste every stage, keep chunk_segm_result[mask_labels[j]][j]
    per_stage = [O.get_seg_masks(s, det, labels, 0.5, (240, 320, 3), 1.0, False) for s in stages]
    ref = [per_stage[int(pick[j])][j] for j in range(n)]
    ml = onehot.cuda() if as_onehot else pick.cuda()
    out = dm().get_seg_masks_switched([s.cuda() for s in stages], ml, det.cuda(), labels.cuda(), Cfg,
                                      (240, 320, 3), 1.0, False)
    agree = sum(int((a == b).sum()) for a, b in zip(out, ref)) / (n * 240 * 320)
    assert agree >= 0.9999, agree
    # every stage is actually used and a detection pasted from the wrong stage would be caught
    assert len(set(pick.tolist())) == 4
    wrong = [per_stage[(int(pick[j]) + 1) % 4][j] for j in range(n)]
    assert sum(int((a == b).sum()) for a, b in zip(out, wrong)) / (n * 240 * 320) < 0.9999


def test_simple_roi_align_large_rois_on_fine_map():
    """Whole-image RoIs on the stride-4 map (patches far wider than a warp: the sparse path),
    mixed with small ones, forward and backward."""
    g = gen(75)
    H, W = 200, 336
    feat = torch.randn(1, 8, H, W, generator=g)
    rois = torch.tensor([[0, 0., 0., 1344., 800.], [0, 100., 50., 1200., 700.], [0, 10., 10., 60., 50.],
                         [0, 600., 5., 1340., 90.], [0, 30., 100., 80., 790.]])
    for P in (14, 56):
        fc = feat.cuda().requires_grad_()
        out = dm().SimpleRoIAlign(P, 0.25)(fc, rois.cuda())
        ref = O.simple_roi_align(feat, rois, P, 0.25)
        assert_close(out, ref, FWD_RTOL, _sra_atol(feat), 'SimpleRoIAlign large P=%d' % P)
        go = torch.randn(out.shape, generator=g)
        out.backward(go.cuda())
        assert_close(fc.grad, O.simple_roi_align_backward(go, feat.shape, rois, 0.25), BWD_RTOL, 2e-4,
                     'SimpleRoIAlign large grad P=%d' % P)


def test_single_level_56_on_large_rois_streams_wide_patch_rows():
    """The switch input (base_roi_head.py:53-58: 56x56, single level, stride 4) on RoIs up to the
    whole image: patch rows far wider than 128 floats take the shallow-ring / multi-copy walk."""
    g = gen(76)
    feat = torch.randn(2, 6, 200, 336, generator=g)
    rois = torch.tensor([[0, 0., 0., 1344., 800.], [1, 100., 50., 1200., 700.], [0, 10., 10., 60., 50.],
                         [1, 600., 5., 1340., 90.], [0, 30., 100., 80., 790.], [1, 3., 3., 900., 500.]])
    ext = dm().SingleRoIExtractor(dict(type='RoIAlign', output_size=56, sampling_ratio=0), 6, [4])
    out = ext([feat.cuda()], rois.cuda())
    ref = O.single_roi_extractor([feat], rois, 56, [4])
    assert_close(out, ref, FWD_RTOL, FWD_ATOL, 'single-level 56 large RoIs')
    out14 = dm().roi_align(feat.cuda(), rois.cuda(), 14, 0.25, 2, 'avg', True)
    assert_close(out14, O.roi_align(feat, rois, (14, 14), 0.25, 2, True), FWD_RTOL, FWD_ATOL, '14 sr2 large RoIs')
