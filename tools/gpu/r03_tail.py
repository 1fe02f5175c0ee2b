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
C5 = len(sys.argv) > 2 and sys.argv[2] == 'c5'
if C5:
    feats, stages, labels, ext, images = bc._tail_setup(dm, dev, 500, mine, 300, (1024, 2048), 256, 0.8)
    ori = (1024, 2048, 3)
else:
    feats, stages, labels, ext, images = bc._tail_setup(dm, dev, 400, mine, 100, (800, 1344), 256, 0.0)
    ori = (800, 1333, 3)
if os.environ.get('TAIL_SLOW'):
    # report every torch.empty / library call that takes longer than 1 ms (allocator misses, blocking launches)
    _empty = torch.empty

    def empty(*a, **k):
        t = time.perf_counter()
        r = _empty(*a, **k)
        dt = (time.perf_counter() - t) * 1e3
        if dt > 1.0:
            print('  slow torch.empty %.1f ms: %s %s' % (dt, a[:1], {kk: str(v) for kk, v in k.items()}))
        return r
    torch.empty = empty
    from dynamask_b200 import _lib
    lib = _lib.load()
    for name in ('dm_paste_rle_strings', 'dm_roi_align_fwd', 'dm_refine_stages', 'dm_assign'):
        f = getattr(lib, name)

        def wrap(f=f, name=name):
            def g(*a):
                t = time.perf_counter()
                r = f(*a)
                dt = (time.perf_counter() - t) * 1e3
                if dt > 1.0:
                    print('  slow %s %.1f ms' % (name, dt))
                return r
            return g
        setattr(lib, name, wrap())
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
