"""world_size-2 gloo test of the only multi-rank logic on the path: shard by image, verify by
all-gathering checksums.  The per-rank work is stood in for by the CPU oracle
so that the test runs without GPUs; what is checked is the sharding + gather plumbing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import dynamask_b200 as dm
    import synth
    from oracle import oracle as O
    g = torch.Generator().manual_seed(77)       # same global problem on every rank
    n_img = 4
    feats = synth.make_features(n_img, 4, 128, 192, g)
    rois = synth.make_rois(n_img, 6, 128, 192, g, s_hi=150.0)
    mine = dm.image_shard(n_img, rank, world)
    local_rois, gidx = dm.shard_rois(rois, rank, world)
    local_feats = [f[mine] for f in feats]
    out = O.single_roi_extractor(local_feats, local_rois, 7, [4, 8, 16, 32])
    sums = dm.gather_checksums([dm.checksum64(out), int(local_rois.size(0))])
    if rank == 0:
        full = O.single_roi_extractor(feats, rois, 7, [4, 8, 16, 32])
        expect = []
        for r in range(world):
            _, gi = dm.shard_rois(rois, r, world)
            expect.append([dm.checksum64(full[gi]), int(gi.numel())])
        ret['ok'] = (sums == expect)
        ret['total'] = sum(s[1] for s in sums)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_shard_by_image_and_gather_checksums_gloo():
    world = 2
    mgr = mp.get_context('spawn').Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret['ok'] is True
    assert ret['total'] == 24
