"""Per-bucket timing of dm_roi_align_fwd / dm_roi_align_bwd on the C2 RoI population: the same
8192 RoIs' quarter (2048) pooled at ONE size per run, so the mixed-launch time of bench.py can be
attributed to resolutions.  Prints one JSON object.  GPU only; not part of the product."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import synth  # noqa: E402
from dynamask_b200 import ops  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    B, C, R = 16, 256, 512
    shapes = synth.pyramid_shapes(bench.IMG_H, bench.IMG_W)
    g = torch.Generator().manual_seed(1234)
    gd = torch.Generator(device=dev).manual_seed(1234)
    feats = [torch.randn(B, C, h, w, generator=gd, device=dev) for (h, w) in shapes]
    rois_h = synth.make_rois(B, R, bench.IMG_H, bench.IMG_W, g)
    onehot_h = synth.make_onehot(rois_h.size(0), g)
    scales = [1.0 / s for s in bench.STRIDES]
    fshapes = [int(v) for f in feats for v in f.shape]
    res = {}
    peak, _ = bench.peaks()
    for b, P in enumerate(bench.BUCKET_SIZES):
        sel = onehot_h.argmax(1) == b
        rois = rois_h[sel].to(dev)
        K = rois.size(0)
        lvl = ops.assign(rois, None, 4, 56.0, 1)[0]
        bucket = torch.full((K, ), b, dtype=torch.long)
        fb, bb = bench.algorithmic_bytes(rois_h[sel], lvl.cpu().long(), bucket, shapes, B, C)
        pyramid = 4.0 * B * C * sum(h * w for h, w in shapes)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf, tb = [], []
        for it in range(6):
            ev[0].record()
            outs = ops.roi_align_forward(feats, rois, lvl, None, None, [K], [P, P], scales, 0, True, False)
            ev[1].record()
            grads = ops.roi_align_backward(outs, rois, lvl, None, None, fshapes, [False] * 4, [P, P], scales, 0, True)
            ev[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                tf.append(ev[0].elapsed_time(ev[1]))
                tb.append(ev[1].elapsed_time(ev[2]))
            del outs, grads
        f_ms, b_ms = sum(tf) / len(tf), sum(tb) / len(tb)
        res['P%d' % P] = {'K': K, 'fwd_ms': f_ms, 'fwd_gbs': fb / f_ms / 1e6, 'fwd_frac': fb / f_ms / 1e6 / peak,
                          'bwd_ms_incl_zero_init': b_ms, 'bwd_gbs': bb / b_ms / 1e6,
                          'bwd_gbs_excl_pyramid': (bb - pyramid) / max(b_ms - pyramid / 7.4e9 * 1e3, 1e-3) / 1e6}
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
