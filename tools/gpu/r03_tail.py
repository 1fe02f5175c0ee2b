"""Per-image wall times of the pipelined C4 loop (bench_configs.run_tail's loop), to see where a slow pass loses its time."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench_configs as bc  # noqa: E402
import dynamask_b200 as dm  # noqa: E402

dev = torch.device('cuda:%d' % int(os.environ.get('LOCAL_RANK', 0)))
torch.cuda.set_device(dev)
mine = int(sys.argv[1]) if len(sys.argv) > 1 else 32
feats, stages, labels, ext, images = bc._tail_setup(dm, dev, 400, mine, 100, (800, 1344), 256, 0.0)
ori = (800, 1333, 3)
for i in range(2):
    bc._tail_image(dm, ext, feats, stages, labels, images[i][0], images[i][1], ori)
torch.cuda.synchronize()
NG = torch.no_grad() if os.environ.get('TAIL_NO_GRAD') else torch.enable_grad()
NG.__enter__()
for rep in range(3):
    ts = []
    pending = None
    t0 = time.perf_counter()
    for i in range(mine):
        a = time.perf_counter()
        _, nxt = bc._tail_image(dm, ext, feats, stages, labels, images[i][0], images[i][1], ori, wait=False)
        b = time.perf_counter()
        if pending is not None:
            pending.result()
        c = time.perf_counter()
        pending = nxt
        ts.append((b - a, c - b))
    pending.result()
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) * 1e3
    print('rank %s rep %d: %.3f ms per image; enqueue ms %s; collect ms %s' % (
        os.environ.get('LOCAL_RANK', '-'), rep, tot / mine, ' '.join('%.2f' % (x[0] * 1e3) for x in ts[:12]), ' '.join('%.2f' % (x[1] * 1e3) for x in ts[:12])))
NG.__exit__(None, None, None)
if os.environ.get('TAIL_PROFILE'):
    import cProfile
    import pstats
    pr = cProfile.Profile()
    with torch.no_grad():
        pr.enable()
        pending = None
        for i in range(mine):
            _, nxt = bc._tail_image(dm, ext, feats, stages, labels, images[i][0], images[i][1], ori, wait=False)
            if pending is not None:
                pending.result()
            pending = nxt
        pending.result()
        pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
print('OMP_NUM_THREADS', os.environ.get('OMP_NUM_THREADS'), 'torch threads', torch.get_num_threads())
