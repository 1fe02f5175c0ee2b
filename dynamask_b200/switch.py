"""Mask Switch Module glue: the hard Gumbel-softmax that turns switch logits into the one-hot
resolution label (reference: ``DynaMaskRoIHead.get_mask_label / gumbel_softmax``,
``mmdet/models/roi_heads/dynamask_roi_head.py:84-114``).  The convs / FCs that produce the
logits (``MaskPre``, ``mmdet/models/roi_heads/base_roi_head.py:10-27``) stay PyTorch modules
and are out of scope; the bucket index consumed by ``dm_assign`` is ``argmax`` of this one-hot.
"""
import torch
import torch.nn.functional as F


def sample_gumbel(shape, eps=1e-20, device=None, generator=None):
    """Gumbel(0,1) noise; the uniform draw is made on the CPU generator like the reference
    (``dynamask_roi_head.py:89-92``) and moved to ``device``."""
    u = torch.rand(shape, generator=generator)
    if device is not None:
        u = u.to(device)
    return -torch.log(-torch.log(u + eps) + eps)


def gumbel_softmax(logits, temperature=1, hard=False, noise=None):
    """Straight-through Gumbel softmax over the last dim; ``hard=True`` returns a one-hot whose
    gradient is that of the soft sample."""
    if noise is None:
        noise = sample_gumbel(logits.size(), device=logits.device)
    y = F.softmax((logits + noise) / temperature, dim=-1)
    if not hard:
        return y
    shape = y.size()
    _, ind = y.max(dim=-1)
    y_hard = torch.zeros_like(y).view(-1, shape[-1])
    y_hard.scatter_(1, ind.view(-1, 1), 1)
    y_hard = y_hard.view(*shape)
    return (y_hard - y).detach() + y


def get_mask_label(mask_logits, noise=None):
    """Switch logits ``[K,4]`` -> one-hot mask label ``[K,4]`` (temperature 0.5, hard)."""
    return gumbel_softmax(mask_logits, temperature=0.5, hard=True, noise=noise)
