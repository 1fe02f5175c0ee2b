"""cProfile of the per-image inference tail of bench.py (host-side costs).  GPU only; not product code."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import synth  # noqa: E402
import dynamask_b200 as dm  # noqa: E402

dev = torch.device('cuda', 0)
g = torch.Generator().manual_seed(3)
rois100 = synth.make_rois(1, 100, 800, 1344, g).to(dev)
st = [torch.randn(100, 1, sz, sz, device=dev) * 3 for sz in (28, 56, 112)]
feats1 = [torch.randn(1, 256, h, w, device=dev) for (h, w) in synth.pyramid_shapes(800, 1344)]
ext = dm.SingleRoIExtractor(dict(type='RoIAlign', output_size=14, sampling_ratio=0), 256, [4, 8, 16, 32])
det100 = torch.cat([rois100[:, 1:], torch.ones(100, 1, device=dev)], 1)
lab100 = torch.zeros(100, dtype=torch.long, device=dev)


class _Cfg:
    mask_thr_binary = 0.5


def parts():
    t = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ins = ext(feats1, rois100); torch.cuda.synchronize(); t.append(time.perf_counter() - t0); t0 = time.perf_counter()
    cl = [x.clone() for x in st]; torch.cuda.synchronize(); t.append(time.perf_counter() - t0); t0 = time.perf_counter()
    final = dm.refine_stage_instance_preds(cl); torch.cuda.synchronize(); t.append(time.perf_counter() - t0); t0 = time.perf_counter()
    r = dm.get_seg_masks_rle(final, det100, lab100, _Cfg, (800, 1333, 3), 1.0, False); t.append(time.perf_counter() - t0)
    return t


def tail():
    ins = ext(feats1, rois100)
    final = dm.refine_stage_instance_preds([t.clone() for t in st])
    return ins, dm.get_seg_masks_rle(final, det100, lab100, _Cfg, (800, 1333, 3), 1.0, False)


for _ in range(3):
    tail()
for _ in range(3):
    print('extractor / clone / refine / paste->RLE ms:', ['%.3f' % (x * 1e3) for x in parts()])
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    tail()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
