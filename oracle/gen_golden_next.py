"""TEST INFRASTRUCTURE ONLY -- golden vectors for the SURVEY 8f "next" rows, generated with the
UNMODIFIED reference (run in the build container; needs /root/reference):

    python oracle/gen_golden_next.py

polygon.npz: PolygonMasks.crop_and_resize + to_ndarray through mask_target_single (reference
Python unmodified; pycocotools' rasteriser is the oracle's C restatement, see ref_shim.load_polygon).

refine.npz: the stage-to-stage refinement loop of DynaMaskRoIHead.simple_test_mask
(dynamask_roi_head.py:136-148, executed from its source lines through oracle/ref_shim.load_refine)
with the reference's own generate_block_target (losses/cross_entropy_loss.py:123-154).
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


def stage_logits(n, gen, sizes=(14, 28, 56, 112), noise=1.5):
    """Correlated coarse-to-fine logits: one smooth blob per instance seen at every size, plus
    independent noise per stage (so the stages disagree near the boundary, as real heads do)."""
    out = []
    cx = torch.rand(n, generator=gen) * 0.6 - 0.3
    cy = torch.rand(n, generator=gen) * 0.6 - 0.3
    rr = torch.rand(n, generator=gen) * 0.5 + 0.3
    for s in sizes:
        lin = (torch.arange(s, dtype=torch.float32) + 0.5) / s * 2 - 1
        d2 = (lin[None, None, :] - cx[:, None, None]) ** 2 + (lin[None, :, None] - cy[:, None, None]) ** 2
        blob = 6.0 * (1.0 - d2 / (rr[:, None, None] ** 2))
        out.append((blob + noise * torch.randn(n, s, s, generator=gen))[:, None].contiguous())
    return out


def main():
    gbt, refine = ref_shim.load_refine()
    g = torch.Generator().manual_seed(20260202)
    preds = stage_logits(6, g)
    # degenerate instances: all foreground, all background, a single pixel
    preds[1][1] = 3.0
    preds[2][1] = 3.0
    preds[1][2] = -3.0
    preds[1][3] = -3.0
    preds[1][3, 0, 13, 14] = 2.0
    inputs = [p.clone() for p in preds]
    final, stages = refine(preds)          # stages[0..2] are preds[1..3], refined in place
    blk = gbt((inputs[1].squeeze(1).sigmoid() >= 0.5), boundary_width=1)
    np.savez_compressed(os.path.join(OUT, 'refine.npz'),
                        **{'in_%d' % i: inputs[i].numpy() for i in range(4)},
                        out_56=stages[1].numpy(), out_112=final.numpy(),
                        block_target_28=blk.numpy())
    print('refine.npz', os.path.getsize(os.path.join(OUT, 'refine.npz')))


def main_polygon():
    """polygon.npz: the reference's PolygonMasks + mask_target_single at the four DynaMask sizes
    (only the pycocotools rasteriser underneath is the oracle's restatement)."""
    PolygonMasks, mask_target_single = ref_shim.load_polygon()
    rng = np.random.default_rng(20260303)
    H, W = 96, 128
    objs = synth.make_polygons(5, H, W, rng)
    objs.append([np.array([32., 24, 96, 24, 96, 72, 32, 72])])           # axis-aligned rectangle
    objs.append([np.array([10.5, 10.5, 60.25, 12.75, 40.0, 50.0]),        # two parts, one degenerate
                 np.array([70., 70, 70, 70, 90, 70, 90, 90, 70, 90])])     # (duplicate vertex)
    pb, pi = synth.jitter_boxes_from_polygons(objs, 16, rng, jitter=6.0)
    pb[0] = (-10, -8, 140, 100)
    pi[0] = 5
    pb[1] = (32, 24, 96, 72)
    pi[1] = 5
    pb[2] = (40, 30, 40.5, 30.2)                                          # extent < 1 -> clamped to 1
    pi[2] = 5
    pi[3] = 6
    pm = PolygonMasks(objs, H, W)
    out = {}
    for s in (14, 28, 56, 112):
        class C:
            mask_size = s
        out['target_%d' % s] = mask_target_single(torch.from_numpy(pb), torch.from_numpy(pi), pm, C).numpy()
    out['full'] = pm.to_ndarray()
    flat = np.concatenate([p for o in objs for p in o])
    voff = np.cumsum([0] + [p.size // 2 for o in objs for p in o])
    ooff = np.cumsum([0] + [len(o) for o in objs])
    np.savez_compressed(os.path.join(OUT, 'polygon.npz'), xy=flat, voff=voff, ooff=ooff, hw=np.array([H, W]),
                        boxes=pb, inds=pi, **out)
    print('polygon.npz', os.path.getsize(os.path.join(OUT, 'polygon.npz')))


if __name__ == '__main__':
    main()
    main_polygon()
