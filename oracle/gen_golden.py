"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz with the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python oracle/gen_golden.py

Every output below is produced by the reference's own Python objects imported through
``oracle/ref_shim.py`` (``SingleRoIExtractor``, ``BitmapMasks``, ``mask_target``,
``_do_paste_mask``, ``DynaMaskHead.get_seg_masks``); the only stand-in is torchvision's CPU
``roi_align`` for the absent mmcv==1.0.5 kernel.  The fixtures travel with the repository, the
reference does not.  Inputs are stored next to the outputs so nothing depends on RNG stability.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
STRIDES = [4, 8, 16, 32]
IMG_H, IMG_W = 256, 384


def adversarial_rois():
    """Boxes whose sqrt-area sits on / next to the level boundaries 56 * 2^k, plus degenerate ones."""
    rows = []
    for k in range(0, 5):
        edge = np.float32(56.0 * 2 ** k)
        for d in (-2, -1, 0, 1, 2):
            v = edge
            for _ in range(abs(d)):
                v = np.nextafter(v, np.float32(0 if d < 0 else 1e9), dtype=np.float32)
            rows.append([0, 0, 0, v, v])
            rows.append([1, 3.5, 7.25, 3.5 + 2 * v, 7.25 + v / 2])
    rows += [[0, 5, 5, 5, 5], [0, 9, 9, 3, 2], [1, 0, 0, 1e4, 1e4], [0, 10, 10, 10.5, 200]]
    return np.asarray(rows, np.float32)


def main():
    ns = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    g = torch.Generator().manual_seed(20260101)

    # ---- stage 1: levels + buckets --------------------------------------------------------------
    rois = torch.cat([synth.make_rois(2, 192, 800, 1344, g), torch.from_numpy(adversarial_rois())])
    onehot = synth.make_onehot(rois.size(0), g, probs=(0.4, 0.3, 0.2, 0.1))
    ext = ns.SingleRoIExtractor(dict(type='RoIAlign', output_size=14, sampling_ratio=0), 4, STRIDES)
    lvl = ext.map_roi_levels(rois, 4)
    bucket = torch.argmax(onehot, dim=1)
    np.savez_compressed(os.path.join(OUT, 'assign.npz'), rois=rois.numpy(), onehot=onehot.numpy(),
                        lvl=lvl.numpy(), bucket=bucket.numpy())

    # ---- stage 2: extractor forward / backward --------------------------------------------------
    feats = synth.make_features(2, 4, IMG_H, IMG_W, g)
    r2 = synth.make_rois(2, 12, IMG_H, IMG_W, g, s_lo=6.0, s_hi=300.0)
    r2[3, 1:] = torch.tensor([-20.0, -10.0, 60.0, 40.0])     # partly outside
    r2[4, 1:] = torch.tensor([300.0, 200.0, 500.0, 400.0])   # mostly outside
    out = {}
    for p in (7, 14):
        e = ns.SingleRoIExtractor(dict(type='RoIAlign', output_size=p, sampling_ratio=0), 4, STRIDES)
        fr = [f.clone().requires_grad_() for f in feats]
        o = e(fr, r2)
        go = torch.randn(o.shape, generator=g)
        o.backward(go)
        out['out_%d' % p] = o.detach().numpy()
        out['gout_%d' % p] = go.numpy()
        for l in range(4):
            gl = fr[l].grad if fr[l].grad is not None else torch.zeros_like(feats[l])
            out['grad_%d_l%d' % (p, l)] = gl.numpy()
    e = ns.SingleRoIExtractor(dict(type='RoIAlign', output_size=7, sampling_ratio=2), 4, STRIDES)
    out['out_7_sr2_rescaled'] = e(feats, r2, roi_scale_factor=1.25).numpy()
    sem = ns.SingleRoIExtractor(dict(type='RoIAlign', output_size=56, sampling_ratio=0), 4, [4])
    out['out_56_single_level'] = sem([feats[0]], r2[:6]).numpy()
    # bucketed: each RoI at the size its label selects, reference extractor run per bucket
    r3 = r2[:8]
    oh3 = torch.zeros(8, 4)
    oh3[torch.arange(8), torch.tensor([3, 0, 1, 2, 0, 3, 2, 1])] = 1
    for b, p in enumerate((14, 28, 56, 112)):
        idx = torch.nonzero(oh3[:, b] > 0).flatten()
        e = ns.SingleRoIExtractor(dict(type='RoIAlign', output_size=p, sampling_ratio=0), 4, STRIDES)
        out['bucket_%d' % b] = e(feats, r3[idx]).numpy()
    np.savez_compressed(os.path.join(OUT, 'extractor.npz'), rois=r2.numpy(), onehot8=oh3.numpy(),
                        **{'feat_l%d' % l: feats[l].numpy() for l in range(4)}, **out)

    # ---- stage 3: paste ---------------------------------------------------------------------------
    n = 7
    logits = synth.make_mask_logits(n, 28, g)
    boxes = synth.make_boxes(n, 120, 160, g, s_lo=6, s_hi=110)
    boxes[2] = torch.tensor([40.5, 10.0, 40.5, 90.0])      # x1 == x0
    boxes[3] = torch.tensor([-15.0, -12.0, 30.0, 41.0])    # partly outside
    det = torch.cat([boxes, torch.ones(n, 1)], 1)
    labels = torch.zeros(n, dtype=torch.long)

    class Cfg:
        mask_thr_binary = 0.5
    segs = ns.DynaMaskHead.get_seg_masks(None, logits, det, labels, Cfg, (120, 160, 3), 1.0, False)
    sf = np.array([1.5, 1.5, 1.5, 1.5], np.float32)
    segs_rs = ns.DynaMaskHead.get_seg_masks(None, logits, det * torch.tensor([1.5, 1.5, 1.5, 1.5, 1.0]),
                                            labels, Cfg, (120, 160, 3), sf, True)

    class Cfg2:
        mask_thr_binary = -1
    segs_u8 = ns.DynaMaskHead.get_seg_masks(None, logits, det, labels, Cfg2, (120, 160, 3), 1.0, False)
    vals, _ = ns.do_paste_mask(logits.sigmoid(), boxes, 120, 160, skip_empty=False)
    np.savez_compressed(os.path.join(OUT, 'paste.npz'), logits=logits.numpy(), boxes=boxes.numpy(),
                        segs=np.stack(segs), segs_rescaled=np.stack(segs_rs), segs_u8=np.stack(segs_u8),
                        values=vals.numpy())

    # ---- stage 4: mask targets --------------------------------------------------------------------
    rng = np.random.default_rng(20260101)
    masks = synth.make_gt_masks(5, 96, 128, rng)
    masks[4] = 0
    masks[4, 24:72, 32:96] = 1                       # rectangle: exact 0.5 ties
    pb, pi = synth.jitter_boxes_from_masks(masks, 14, rng, jitter=6.0)
    pb[0] = (-10, -8, 140, 100)                      # clipped to the canvas
    pb[1] = (0, 0, 128, 96)
    pb[2] = (32, 24, 96, 72)
    pi[2] = 4
    pb[3] = (16, 12, 80, 60)
    pi[3] = 4
    bm = ns.BitmapMasks(masks, 96, 128)
    tg = {}
    for s in (14, 28, 56, 112):
        class C:
            mask_size = s
        tg['target_%d' % s] = ns.mask_target_single(torch.from_numpy(pb), torch.from_numpy(pi), bm, C).numpy()
    cr = bm.crop_and_resize(pb, (28, 28), pi).masks
    np.savez_compressed(os.path.join(OUT, 'mask_target.npz'), masks=masks, boxes=pb, inds=pi,
                        crop_28=cr, **tg)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
