"""Per-call timing of the extractor calls of one C3 training step (2 images).  GPU only."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import synth  # noqa: E402
import dynamask_b200 as dm  # noqa: E402
from dynamask_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)
g = torch.Generator().manual_seed(1234)
shapes = synth.pyramid_shapes(bench.IMG_H, bench.IMG_W)
feats = [torch.randn(2, 256, h, w, device=dev) for (h, w) in shapes]
rois = synth.make_rois(2, 512, bench.IMG_H, bench.IMG_W, g).to(dev)
r_mask = torch.cat([rois[:128], rois[512:640]])
scales = [1.0 / s for s in bench.STRIDES]
fshapes = [int(v) for f in feats for v in f.shape]
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        r = fn()
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, r


res = {}
for name, rr, P, fl, sc in (('bbox7', rois, 7, feats, scales), ('mask14', r_mask, 14, feats, scales),
                            ('sem56', r_mask, 56, feats[:1], scales[:1])):
    L = len(fl)
    t_as, asg = timed(lambda: ops.assign(rr, None, L, 56.0, 1))
    lvl = asg[0] if L > 1 else None
    K = rr.size(0)
    t_f, outs = timed(lambda: ops.roi_align_forward(fl, rr, lvl, None, None, [K], [P, P], sc, 0, True, False))
    fs = [int(v) for f in fl for v in f.shape]
    t_b, _ = timed(lambda: ops.roi_align_backward(outs, rr, lvl, None, None, fs, [False] * L, [P, P], sc, 0, True))
    res[name] = {'assign_ms': t_as, 'fwd_ms': t_f, 'bwd_ms_incl_zero_init': t_b, 'out_MB': outs[0].numel() * 4 / 1e6}
ext = dm.SingleRoIExtractor(dict(type='RoIAlign', output_size=7, sampling_ratio=0), 256, bench.STRIDES)
fr = [f.clone().requires_grad_() for f in feats]


def full():
    for f in fr:
        f.grad = None
    o = ext(fr, rois)
    o.backward(o.detach())


res['bbox7_module_fwd_bwd_ms'] = timed(full)[0]
t_z, _ = timed(lambda: [torch.zeros_like(f) for f in feats])
res['zeros_like_pyramid_ms'] = t_z
print(json.dumps(res, indent=1))
