"""Multi-GPU layout of the hot path: images are sharded over ranks, nothing else.

Every RoI reads only its own image's feature maps (``rois[:, 0]`` is the image index,
``mmdet/core/bbox/transforms.py:67-68``) and every pasted mask / mask target depends on one image
only, so the reference's one-process-per-GPU, sharded-by-image layout
(``mmdet/datasets/samplers/group_sampler.py:51-140``) needs no data-path collective.  The only
traffic is verification: an all-gather of 64-bit checksums so that rank 0 can compare every rank's
outputs with the oracle's.
"""
import torch
import torch.distributed as dist


def image_shard(num_images, rank, world_size):
    """Images owned by ``rank``: ``{i : i mod world_size == rank}`` (round-robin, like the sampler)."""
    return list(range(rank, num_images, world_size))


def shard_rois(rois, rank, world_size):
    """Select the RoIs of this rank's images and renumber their image index to the local batch.

    Returns ``(local_rois [k,5], global_index [k])``; ``local_rois[:, 0]`` indexes the rank's own
    feature batch (local image j = global image ``rank + j * world_size``).
    """
    img = rois[:, 0].long()
    keep = (img % world_size) == rank
    local = rois[keep].clone()
    local[:, 0] = torch.div(img[keep] - rank, world_size, rounding_mode='floor').to(rois.dtype)
    return local, torch.nonzero(keep).flatten()


def checksum64(t):
    """Order-independent 64-bit checksum of a tensor's bit pattern (int64, wraps modulo 2^64)."""
    if t.numel() == 0:
        return 0
    t = t.detach().contiguous()
    if t.dtype in (torch.float32, torch.int32):
        bits = t.view(torch.int32).to(torch.int64)
    elif t.dtype in (torch.uint8, torch.bool):
        bits = t.view(torch.uint8).to(torch.int64)
    elif t.dtype == torch.int64:
        bits = t
    else:
        bits = t.to(torch.float32).view(torch.int32).to(torch.int64)
    idx = torch.arange(1, bits.numel() + 1, device=bits.device, dtype=torch.int64)
    # position-weighted so that permuted data does not collide; int64 arithmetic wraps silently
    return int((bits.flatten() * (idx % 65521 + 1)).sum().item())


def gather_checksums(values, group=None):
    """All-gather a list of python ints (one per output tensor) -> list over ranks of lists."""
    if not (dist.is_available() and dist.is_initialized()):
        return [list(values)]
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    mine = torch.tensor(list(values), dtype=torch.int64, device=dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [o.cpu().tolist() for o in out]
