"""Paste kernel experiments: time dm_paste_masks on the C4 shape (800 x 112^2 -> 800x1333) for
(a) COCO-shaped boxes, (b) boxes that miss the canvas (pure zero-fill path), (c) a 16-byte fill of
the same bytes.  The variant comes from DM_PASTE_VARIANT (read once by the library)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import synth  # noqa: E402
from dynamask_b200 import ops  # noqa: E402

dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(99)
n = 800
logits = synth.make_mask_logits(n, 112, g).to(dev)
boxes = synth.make_boxes(n, 800, 1333, g, s_lo=8, s_hi=500).to(dev)
dead = boxes.clone()
dead[:, 0] += 5000
dead[:, 2] += 5000


def timeit(f, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def run(bx):
    return ops.paste_masks(logits, bx, None, 800, 1333, [0, 0, 1333, 800], True, 0.5, ops.PASTE_BOOL)


t_live = timeit(lambda: run(boxes))
t_dead = timeit(lambda: run(dead))
out = run(boxes)
frac = float(out.view(torch.uint8).float().mean())
o32 = out.flatten().view(torch.int32)
t_fill = timeit(lambda: o32.zero_())
by = n * 800 * 1333
print('variant %s: coco boxes %.1f us (%.0f GB/s), off-canvas boxes %.1f us (%.0f GB/s), int32 fill %.1f us; foreground %.4f' % (
    os.environ.get('DM_PASTE_TILES', 'default'), t_live, by / t_live / 1e3, t_dead, by / t_dead / 1e3, t_fill, frac))
