"""Seeded synthetic inputs for tests and bench.  CPU tensors; callers move
them to the device.  Not part of the product package."""
import math

import numpy as np
import torch

STRIDES = (4, 8, 16, 32)


def pyramid_shapes(img_h, img_w, strides=STRIDES):
    """FPN map sizes for a padded image, e.g. 800x1344 -> 200x336, 100x168, 50x84, 25x42."""
    return [(int(math.ceil(img_h / s)), int(math.ceil(img_w / s))) for s in strides]


def make_features(batch, channels, img_h, img_w, gen, strides=STRIDES):
    return [torch.randn(batch, channels, h, w, generator=gen)
            for (h, w) in pyramid_shapes(img_h, img_w, strides)]


def make_boxes(n, img_h, img_w, gen, s_lo=8.0, s_hi=700.0, small_frac=0.0, small_range=(4.0, 128.0)):
    """COCO-shaped boxes: sqrt-area log-uniform in [s_lo, s_hi], aspect log-uniform in [1/3, 3],
    clipped to the image and placed uniformly inside it.  Returns [n,4] xyxy float32."""
    u = torch.rand(n, generator=gen)
    s = torch.exp(u * math.log(s_hi / s_lo)) * s_lo
    if small_frac > 0:
        pick = torch.rand(n, generator=gen) < small_frac
        us = torch.rand(n, generator=gen)
        small = torch.exp(us * math.log(small_range[1] / small_range[0])) * small_range[0]
        s = torch.where(pick, small, s)
    a = torch.exp((torch.rand(n, generator=gen) * 2 - 1) * math.log(3.0))
    w = torch.clamp(s * a.sqrt(), max=float(img_w))
    h = torch.clamp(s / a.sqrt(), max=float(img_h))
    cx = w / 2 + torch.rand(n, generator=gen) * (img_w - w)
    cy = h / 2 + torch.rand(n, generator=gen) * (img_h - h)
    return torch.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], dim=1).float()


def make_rois(batch, per_img, img_h, img_w, gen, **kw):
    rows = []
    for b in range(batch):
        boxes = make_boxes(per_img, img_h, img_w, gen, **kw)
        rows.append(torch.cat([torch.full((per_img, 1), float(b)), boxes], dim=1))
    return torch.cat(rows, 0)


def make_onehot(n, gen, probs=(0.25, 0.25, 0.25, 0.25)):
    idx = torch.multinomial(torch.tensor(probs), n, replacement=True, generator=gen)
    oh = torch.zeros(n, len(probs))
    oh[torch.arange(n), idx] = 1.0
    return oh


def make_mask_logits(n, size, gen):
    """[n,1,size,size] logits: radial blob 6*(1-r^2) plus N(0,1) noise (~50 % foreground)."""
    lin = (torch.arange(size, dtype=torch.float32) + 0.5) / size * 2 - 1
    r2 = lin[None, :] ** 2 + lin[:, None] ** 2
    return (6.0 * (1.0 - r2))[None, None] + torch.randn(n, 1, size, size, generator=gen)


def make_gt_masks(g, img_h, img_w, rng):
    """uint8 [g,img_h,img_w]: one random filled ellipse or rectangle per mask."""
    m = np.zeros((g, img_h, img_w), np.uint8)
    yy, xx = np.mgrid[0:img_h, 0:img_w]
    for i in range(g):
        cx, cy = rng.uniform(0, img_w), rng.uniform(0, img_h)
        rw, rh = rng.uniform(6, img_w / 3), rng.uniform(6, img_h / 3)
        if rng.random() < 0.5:
            m[i] = (((xx - cx) / rw) ** 2 + ((yy - cy) / rh) ** 2) <= 1.0
        else:
            x0, x1 = int(max(cx - rw, 0)), int(min(cx + rw, img_w))
            y0, y1 = int(max(cy - rh, 0)), int(min(cy + rh, img_h))
            m[i, y0:y1, x0:x1] = 1
    return m


def jitter_boxes_from_masks(masks, k, rng, jitter=15.0):
    """k positive proposals: bounding boxes of random gt masks jittered by +-jitter px."""
    g, h, w = masks.shape
    inds = rng.integers(0, g, size=k)
    boxes = np.zeros((k, 4), np.float32)
    for j, gi in enumerate(inds):
        ys, xs = np.nonzero(masks[gi])
        if len(xs) == 0:
            x0, y0, x1, y1 = 0, 0, w / 4, h / 4
        else:
            x0, x1, y0, y1 = xs.min(), xs.max() + 1, ys.min(), ys.max() + 1
        d = rng.uniform(-jitter, jitter, size=4)
        bx0, by0 = x0 + d[0], y0 + d[1]
        bx1, by1 = max(x1 + d[2], bx0 + 2), max(y1 + d[3], by0 + 2)
        boxes[j] = (bx0, by0, bx1, by1)
    return boxes, inds.astype(np.int64)


def make_polygons(g, img_h, img_w, rng, max_parts=3):
    """COCO-shaped polygon ground truth: per object 1..max_parts star-shaped polygons with 5..40
    float64 vertices (two decimals, like the dataset's json), some reaching past the image."""
    objs = []
    for _ in range(g):
        parts = []
        for _ in range(int(rng.integers(1, max_parts + 1))):
            cx, cy = rng.uniform(0, img_w), rng.uniform(0, img_h)
            rad = rng.uniform(4, min(img_h, img_w) / 3)
            n = int(rng.integers(5, 41))
            ang = np.sort(rng.uniform(0, 2 * np.pi, n))
            r = rad * rng.uniform(0.5, 1.0, n)
            p = np.empty(2 * n, np.float64)
            p[0::2] = np.round(cx + r * np.cos(ang), 2)
            p[1::2] = np.round(cy + r * np.sin(ang), 2)
            parts.append(p)
        objs.append(parts)
    return objs


def jitter_boxes_from_polygons(objs, k, rng, jitter=15.0):
    """k positive proposals: bounding boxes of random objects' polygons jittered by +-jitter px."""
    inds = rng.integers(0, len(objs), size=k)
    boxes = np.zeros((k, 4), np.float32)
    for j, gi in enumerate(inds):
        xs = np.concatenate([p[0::2] for p in objs[gi]])
        ys = np.concatenate([p[1::2] for p in objs[gi]])
        d = rng.uniform(-jitter, jitter, size=4)
        bx0, by0 = xs.min() + d[0], ys.min() + d[1]
        boxes[j] = (bx0, by0, max(xs.max() + d[2], bx0 + 2), max(ys.max() + d[3], by0 + 2))
    return boxes, inds.astype(np.int64)
