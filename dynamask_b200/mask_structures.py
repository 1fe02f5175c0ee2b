"""Ground-truth instance bitmaps (reference type: ``BitmapMasks``,
``mmdet/core/mask/structures.py:136-311``).

The container keeps the reference's host-side contract -- ``masks`` is a uint8 ndarray
``[N,H,W]`` and every method returns what the reference returns -- and adds a device-resident
path for ``crop_and_resize``: the bitmaps are uploaded once per (object, device) and
``dm_mask_target`` reads the uint8 planes in place, instead of the reference's per-call upload,
``index_select`` and fp32 blow-up (``structures.py:279-283``).
"""
import numpy as np
import torch

from . import ops


def pack_bitmaps(masks_list, device):
    """Concatenate per-image uint8 ``[G,H,W]`` arrays into one device blob.

    Returns ``(blob uint8 [total], img_offsets int64 [B], img_ghw int32 [B,3])`` on ``device``;
    a single pinned staging buffer and three async copies, no synchronisation.
    """
    sizes = [int(m.size) for m in masks_list]
    total = max(sum(sizes), 1)
    offs = np.zeros(len(masks_list), np.int64)
    ghw = np.zeros((len(masks_list), 3), np.int32)
    pin = torch.device(device).type == 'cuda'
    stage = torch.empty(total, dtype=torch.uint8, pin_memory=pin)
    view = stage.numpy()
    pos = 0
    for b, m in enumerate(masks_list):
        offs[b] = pos
        if m.ndim == 3:
            ghw[b] = m.shape
        view[pos:pos + sizes[b]] = np.ascontiguousarray(m, dtype=np.uint8).reshape(-1)
        pos += sizes[b]
    blob = stage.to(device, non_blocking=True)
    meta = torch.empty(len(masks_list) * 5, dtype=torch.int64, pin_memory=pin)
    meta_np = meta.numpy()
    meta_np[:len(masks_list)] = offs
    # int32 (G,H,W) triples stored behind the int64 offsets in the same pinned buffer
    meta_np[len(masks_list):].view(np.int32)[:ghw.size] = ghw.reshape(-1)
    meta_dev = meta.to(device, non_blocking=True)
    img_offsets = meta_dev[:len(masks_list)]
    img_ghw = meta_dev[len(masks_list):].view(torch.int32)[:ghw.size]
    return blob, img_offsets, img_ghw


class BitmapMasks(object):
    """Masks in the form of bitmaps.

    Args:
        masks (ndarray | list[ndarray]): masks of shape (N, H, W).
        height (int): height of masks.
        width (int): width of masks.
    """

    def __init__(self, masks, height, width):
        self.height = height
        self.width = width
        if len(masks) == 0:
            self.masks = np.empty((0, self.height, self.width), dtype=np.uint8)
        else:
            assert isinstance(masks, (list, np.ndarray))
            if isinstance(masks, list):
                assert isinstance(masks[0], np.ndarray)
                assert masks[0].ndim == 2  # (H, W)
            else:
                assert masks.ndim == 3  # (N, H, W)
            self.masks = np.stack(masks).reshape(-1, height, width)
            assert self.masks.shape[1] == self.height
            assert self.masks.shape[2] == self.width
        self._device_cache = {}

    # ---- container protocol -------------------------------------------------------------
    def __getitem__(self, index):
        masks = self.masks[index].reshape(-1, self.height, self.width)
        return BitmapMasks(masks, self.height, self.width)

    def __iter__(self):
        return iter(self.masks)

    def __repr__(self):
        return (f'{self.__class__.__name__}(num_masks={len(self.masks)}, '
                f'height={self.height}, width={self.width})')

    def __len__(self):
        return len(self.masks)

    # ---- host-side transforms (data pipeline; not on the hot path) --------------------------
    def flip(self, flip_direction='horizontal'):
        assert flip_direction in ('horizontal', 'vertical')
        if len(self.masks) == 0:
            flipped = self.masks
        else:
            axis = 2 if flip_direction == 'horizontal' else 1
            flipped = np.ascontiguousarray(np.flip(self.masks, axis=axis))
        return BitmapMasks(flipped, self.height, self.width)

    def pad(self, out_shape, pad_val=0):
        if len(self.masks) == 0:
            padded = np.empty((0, *out_shape), dtype=np.uint8)
        else:
            padded = np.full((len(self.masks), *out_shape), pad_val, dtype=self.masks.dtype)
            padded[:, :self.height, :self.width] = self.masks
        return BitmapMasks(padded, *out_shape)

    def crop(self, bbox):
        assert isinstance(bbox, np.ndarray)
        assert bbox.ndim == 1
        bbox = bbox.copy()
        bbox[0::2] = np.clip(bbox[0::2], 0, self.width)
        bbox[1::2] = np.clip(bbox[1::2], 0, self.height)
        x1, y1, x2, y2 = bbox
        w = np.maximum(x2 - x1, 1)
        h = np.maximum(y2 - y1, 1)
        if len(self.masks) == 0:
            cropped = np.empty((0, h, w), dtype=np.uint8)
        else:
            cropped = self.masks[:, y1:y1 + h, x1:x1 + w]
        return BitmapMasks(cropped, h, w)

    def expand(self, expanded_h, expanded_w, top, left):
        if len(self.masks) == 0:
            expanded = np.empty((0, expanded_h, expanded_w), dtype=np.uint8)
        else:
            expanded = np.zeros((len(self), expanded_h, expanded_w), dtype=np.uint8)
            expanded[:, top:top + self.height, left:left + self.width] = self.masks
        return BitmapMasks(expanded, expanded_h, expanded_w)

    def resize(self, out_shape, interpolation='nearest'):
        """Resize to ``(h, w)`` with OpenCV (the reference goes through ``mmcv.imresize``)."""
        import cv2
        flag = {'nearest': cv2.INTER_NEAREST, 'bilinear': cv2.INTER_LINEAR}[interpolation]
        if len(self.masks) == 0:
            resized = np.empty((0, *out_shape), dtype=np.uint8)
        else:
            resized = np.stack([
                cv2.resize(m, (out_shape[1], out_shape[0]), interpolation=flag) for m in self.masks
            ])
        return BitmapMasks(resized, *out_shape)

    def rescale(self, scale, interpolation='nearest'):
        """Rescale keeping the aspect ratio; ``scale`` is a factor or a (long, short) edge pair."""
        if isinstance(scale, (float, int)):
            factor = float(scale)
        else:
            long_e, short_e = max(scale), min(scale)
            factor = min(long_e / max(self.height, self.width), short_e / min(self.height, self.width))
        new_w = int(self.width * factor + 0.5)
        new_h = int(self.height * factor + 0.5)
        return self.resize((new_h, new_w), interpolation=interpolation)

    @property
    def areas(self):
        return self.masks.sum((1, 2))

    def to_ndarray(self):
        return self.masks

    def to_tensor(self, dtype, device):
        return torch.tensor(self.masks, dtype=dtype, device=device)

    # ---- hot path -----------------------------------------------------------------------
    def to_device(self, device):
        """Upload the bitmaps once per device; returns ``(blob, img_offsets, img_ghw)``."""
        device = torch.device(device)
        key = (device.type, device.index)
        hit = self._device_cache.get(key)
        if hit is None:
            hit = pack_bitmaps([self.masks], device)
            self._device_cache[key] = hit
        return hit

    def crop_and_resize_device(self, bboxes, out_shapes, inds, device, clip=False):
        """Targets for several output shapes in one launch, left on the device.

        ``bboxes [K,4]`` (ndarray or tensor), ``out_shapes`` list of (h, w), ``inds [K]``.
        Returns a list of float32 ``[K,h,w]`` tensors holding {0,1}.
        """
        device = torch.device(device)
        if device.type != 'cuda':
            raise NotImplementedError('dynamask_b200 has no CPU path: pass a CUDA device')
        if isinstance(bboxes, np.ndarray):
            bboxes = torch.from_numpy(np.ascontiguousarray(bboxes, dtype=np.float32))
        if isinstance(inds, np.ndarray):
            inds = torch.from_numpy(np.ascontiguousarray(inds).astype(np.int64))
        bboxes = bboxes.to(device=device, dtype=torch.float32, non_blocking=True)
        inds = inds.to(device=device, dtype=torch.int64, non_blocking=True)
        blob, offs, ghw = self.to_device(device)
        sizes = [int(v) for hw in out_shapes for v in hw]
        return ops.mask_target(blob, offs, ghw, bboxes, inds, None, bool(clip), sizes)

    def crop_and_resize(self, bboxes, out_shape, inds, device='cpu', interpolation='bilinear'):
        """Crop each box from the mask ``inds`` selects and resize it to ``out_shape``.

        Same arguments and return type as the reference (a new ``BitmapMasks`` of bool
        ``[K,h,w]`` on the host); ``device`` must be a CUDA device.
        """
        if len(self.masks) == 0:
            empty_masks = np.empty((0, *out_shape), dtype=np.uint8)
            return BitmapMasks(empty_masks, *out_shape)
        num_bbox = bboxes.shape[0]
        if num_bbox > 0:
            t = self.crop_and_resize_device(bboxes, [tuple(out_shape)], inds, device)[0]
            resized_masks = (t >= 0.5).cpu().numpy()
        else:
            resized_masks = []
        return BitmapMasks(resized_masks, *out_shape)


# ---------------------------------------------------------------------------------------------
# Polygon ground truth (reference type: ``PolygonMasks``, ``mmdet/core/mask/structures.py:314-558``)
# SURVEY.md row A10 / 8f rank 4.
# ---------------------------------------------------------------------------------------------
def pack_polygons(polys_per_image, hw_per_image, device):
    """Flatten the polygons of a batch for ``dm_polygon_target``.

    ``polys_per_image``: per image, a list (objects) of lists (polygons) of 1-D float arrays
    ``(x0, y0, x1, y1, ...)``.  Returns ``(xy float64 [2V], vert_offsets int64 [P+1],
    obj_poly_offsets int32 [G+1], img_meta int32 [B*3])`` on ``device``.
    """
    flat, voff, ooff, meta = [], [0], [0], []
    for objs, (h, w) in zip(polys_per_image, hw_per_image):
        meta += [len(ooff) - 1, int(h), int(w)]
        for polys in objs:
            for p in polys:
                a = np.asarray(p, dtype=np.float64).reshape(-1)
                flat.append(a[:2 * (a.size // 2)])
                voff.append(voff[-1] + a.size // 2)
            ooff.append(len(voff) - 1)
    xy = np.concatenate(flat) if flat else np.zeros(0, np.float64)
    pin = torch.device(device).type == 'cuda'

    def up(a, dt):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=dt))
        if pin:
            t = t.pin_memory()
        return t.to(device, non_blocking=True)

    xy_d = up(xy if xy.size else np.zeros(2, np.float64), np.float64)
    return xy_d, up(np.asarray(voff), np.int64), up(np.asarray(ooff), np.int32), up(np.asarray(meta), np.int32)


class PolygonMasks(object):
    """Masks in the form of polygons: a list (objects) of lists (polygons of the object) of 1-D
    coordinate arrays ``(x0, y0, x1, y1, ...)``.

    Same constructor, container protocol, host transforms and ``crop_and_resize`` /
    ``to_ndarray`` results as the reference class; rasterisation runs on the device
    (``dm_polygon_target``) instead of pycocotools on the host, and
    :meth:`crop_and_resize_device` produces the float targets of every size in one launch.
    """

    def __init__(self, masks, height, width):
        assert isinstance(masks, list)
        if len(masks) > 0:
            assert isinstance(masks[0], list)
            assert isinstance(masks[0][0], np.ndarray)
        self.height = height
        self.width = width
        self.masks = masks
        self._device_cache = {}

    # ---- container protocol -------------------------------------------------------------
    def __getitem__(self, index):
        if isinstance(index, np.ndarray):
            index = index.tolist()
        if isinstance(index, list):
            masks = [self.masks[i] for i in index]
        else:
            try:
                masks = self.masks[index]
            except Exception:
                raise ValueError(f'Unsupported input of type {type(index)} for indexing!')
        if len(masks) and isinstance(masks[0], np.ndarray):
            masks = [masks]  # keep three levels
        return PolygonMasks(masks, self.height, self.width)

    def __iter__(self):
        return iter(self.masks)

    def __repr__(self):
        return (f'{self.__class__.__name__}(num_masks={len(self.masks)}, '
                f'height={self.height}, width={self.width})')

    def __len__(self):
        return len(self.masks)

    # ---- host-side transforms (data pipeline; not on the hot path) --------------------------
    def _map(self, fn, height, width):
        return PolygonMasks([[fn(p.copy()) for p in obj] for obj in self.masks], height, width)

    def resize(self, out_shape, interpolation=None):
        if len(self.masks) == 0:
            return PolygonMasks([], *out_shape)
        h_scale = out_shape[0] / self.height
        w_scale = out_shape[1] / self.width

        def fn(p):
            p[0::2] *= w_scale
            p[1::2] *= h_scale
            return p
        return self._map(fn, *out_shape)

    def rescale(self, scale, interpolation=None):
        if isinstance(scale, (float, int)):
            factor = float(scale)
        else:
            long_e, short_e = max(scale), min(scale)
            factor = min(long_e / max(self.height, self.width), short_e / min(self.height, self.width))
        new_w = int(self.width * factor + 0.5)
        new_h = int(self.height * factor + 0.5)
        return self.resize((new_h, new_w))

    def flip(self, flip_direction='horizontal'):
        assert flip_direction in ('horizontal', 'vertical')
        dim, idx = (self.width, 0) if flip_direction == 'horizontal' else (self.height, 1)

        def fn(p):
            p[idx::2] = dim - p[idx::2]
            return p
        return self._map(fn, self.height, self.width)

    def crop(self, bbox):
        assert isinstance(bbox, np.ndarray)
        assert bbox.ndim == 1
        bbox = bbox.copy()
        bbox[0::2] = np.clip(bbox[0::2], 0, self.width)
        bbox[1::2] = np.clip(bbox[1::2], 0, self.height)
        x1, y1, x2, y2 = bbox
        w = np.maximum(x2 - x1, 1)
        h = np.maximum(y2 - y1, 1)

        def fn(p):
            p[0::2] -= bbox[0]
            p[1::2] -= bbox[1]
            return p
        return self._map(fn, h, w)

    def pad(self, out_shape, pad_val=0):
        """padding has no effect on polygons"""
        return PolygonMasks(self.masks, *out_shape)

    def expand(self, *args, **kwargs):
        raise NotImplementedError

    @property
    def areas(self):
        """Shoelace area of every object (sum over its polygons)."""
        out = []
        for obj in self.masks:
            a = 0.0
            for p in obj:
                x, y = p[0::2], p[1::2]
                a += 0.5 * np.abs(np.dot(x, np.roll(y, 1)) - np.dot(y, np.roll(x, 1)))
            out.append(a)
        return np.asarray(out)

    # ---- hot path -----------------------------------------------------------------------
    def to_device(self, device):
        """Upload the flattened polygons once per device; returns the ``pack_polygons`` tuple."""
        device = torch.device(device)
        key = (device.type, device.index)
        hit = self._device_cache.get(key)
        if hit is None:
            hit = pack_polygons([self.masks], [(self.height, self.width)], device)
            self._device_cache[key] = hit
        return hit

    def crop_and_resize_device(self, bboxes, out_shapes, inds, device, clip=False):
        """Float ``[K,h,w]`` {0,1} targets for several output shapes in one launch, on the device."""
        device = torch.device(device)
        if device.type != 'cuda':
            raise NotImplementedError('dynamask_b200 has no CPU path: pass a CUDA device')
        if isinstance(bboxes, np.ndarray):
            bboxes = torch.from_numpy(np.ascontiguousarray(bboxes, dtype=np.float32))
        if isinstance(inds, np.ndarray):
            inds = torch.from_numpy(np.ascontiguousarray(inds).astype(np.int64))
        bboxes = bboxes.to(device=device, dtype=torch.float32, non_blocking=True)
        inds = inds.to(device=device, dtype=torch.int64, non_blocking=True)
        xy, voff, ooff, meta = self.to_device(device)
        sizes = [int(v) for hw in out_shapes for v in hw]
        return ops.polygon_target(xy, voff, ooff, meta, bboxes, inds, None, bool(clip), sizes)

    def crop_and_resize(self, bboxes, out_shape, inds, device='cpu', interpolation='bilinear'):
        """Same return type as the reference: a new ``PolygonMasks`` whose polygons are shifted by
        the box corner and scaled to ``out_shape`` (host arithmetic identical to
        ``structures.py:478-499``); rasterise it with :meth:`to_ndarray`."""
        out_h, out_w = out_shape
        if len(self.masks) == 0:
            return PolygonMasks([], out_h, out_w)
        resized_masks = []
        for i in range(len(bboxes)):
            bbox = bboxes[i, :]
            x1, y1, x2, y2 = bbox
            w = np.maximum(x2 - x1, 1)
            h = np.maximum(y2 - y1, 1)
            h_scale = out_h / max(h, 0.1)
            w_scale = out_w / max(w, 0.1)
            obj = []
            for p in self.masks[inds[i]]:
                p = p.copy()
                p[0::2] = (p[0::2] - bbox[0]) * w_scale
                p[1::2] = (p[1::2] - bbox[1]) * h_scale
                obj.append(p)
            resized_masks.append(obj)
        return PolygonMasks(resized_masks, *out_shape)

    def to_ndarray(self, device=None):
        """Rasterise every object at ``(height, width)`` on the device -> bool ``[N,H,W]`` ndarray
        (reference: pycocotools ``frPyObjects -> merge -> decode`` per object on the host)."""
        if len(self.masks) == 0:
            return np.empty((0, self.height, self.width), dtype=np.uint8)
        if device is None:
            if not torch.cuda.is_available():
                raise NotImplementedError('dynamask_b200 has no CPU rasteriser: a CUDA device is required')
            device = torch.device('cuda', torch.cuda.current_device())
        n = len(self.masks)
        # identity crop: box (0, 0, W, H) scaled to (H, W) is exactly scale 1.0 and offset 0.0
        boxes = np.tile(np.asarray([[0, 0, self.width, self.height]], np.float32), (n, 1))
        t = self.crop_and_resize_device(boxes, [(int(self.height), int(self.width))], np.arange(n), device)[0]
        return t.to(torch.bool).cpu().numpy()

    def to_bitmap(self):
        return BitmapMasks(self.to_ndarray(), self.height, self.width)

    def to_tensor(self, dtype, device):
        if len(self.masks) == 0:
            return torch.empty((0, self.height, self.width), dtype=dtype, device=device)
        return torch.tensor(self.to_ndarray(), dtype=dtype, device=device)
