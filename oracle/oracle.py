"""TEST INFRASTRUCTURE ONLY -- CPU oracle ("port") of the DynaMask per-instance mask path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module; the product package ``dynamask_b200`` never does.

Every function restates one reference function (paths relative to ``/root/reference``) on host
memory with numpy/torch-CPU; the kernel arithmetic lives in ``dm_oracle.c`` (plain C, FMA
contraction off).  Parity pinning: the reference's own tests hold no numeric vectors for this
path, so the oracle is pinned against (a) golden vectors produced by the
unmodified reference Python run in the build container through ``oracle/ref_shim.py``
(``tests/golden/*.npz``, generator ``oracle/gen_golden.py``) and (b) torchvision's CPU
``roi_align`` -- the stand-in for the absent third-party mmcv==1.0.5 kernel.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    so = os.path.join(_HERE, 'libdm_oracle.so')
    src = os.path.join(_HERE, 'dm_oracle.c')
    if force or not os.path.exists(so) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _np32(x):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.float32)


def _p(a, t=_f32p):
    return a.ctypes.data_as(t)


# --------------------------------------------------------------------------------------------
# A0  bbox2roi -- mmdet/core/bbox/transforms.py:54-73
# --------------------------------------------------------------------------------------------
def bbox2roi(bbox_list):
    """[k_i,4+] xyxy per image -> [K,5] (image index as float, x1, y1, x2, y2)."""
    rows = []
    for img_id, b in enumerate(bbox_list):
        b = _np32(b).reshape(-1, b.shape[-1] if len(b.shape) > 1 else 4)
        col = np.full((b.shape[0], 1), img_id, dtype=np.float32)
        rows.append(np.concatenate([col, b[:, :4]], axis=1))
    if not rows:
        return np.zeros((0, 5), np.float32)
    return np.concatenate(rows, axis=0).astype(np.float32)


# --------------------------------------------------------------------------------------------
# A1  resolution bucket -- mmdet/models/roi_heads/dynamask_roi_head.py:84-114 (+ :197-203)
# --------------------------------------------------------------------------------------------
def gumbel_softmax_hard(logits, uniform, temperature=0.5, eps=1e-20):
    """Forward value of the hard (straight-through) Gumbel softmax given the uniform noise."""
    logits = torch.as_tensor(logits, dtype=torch.float32)
    u = torch.as_tensor(uniform, dtype=torch.float32)
    g = -torch.log(-torch.log(u + eps) + eps)
    y = F.softmax((logits + g) / temperature, dim=-1)
    ind = y.max(dim=-1)[1]
    hard = torch.zeros_like(y).view(-1, y.shape[-1])
    hard.scatter_(1, ind.view(-1, 1), 1)
    return hard.view_as(y), ind


def bucket_of(onehot):
    """bucket index = argmax of the one-hot mask label (first maximal index)."""
    return torch.as_tensor(onehot).argmax(dim=1).to(torch.int64)


# --------------------------------------------------------------------------------------------
# A2  map_roi_levels -- .../roi_extractors/single_level_roi_extractor.py:32-51
# --------------------------------------------------------------------------------------------
def map_roi_levels(rois, num_levels, finest_scale=56):
    """The reference expression itself, on whatever device ``rois`` lives on (torch)."""
    rois = torch.as_tensor(rois, dtype=torch.float32)
    scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
    lv = torch.floor(torch.log2(scale / finest_scale + 1e-6))
    return lv.clamp(min=0, max=num_levels - 1).long()


def map_roi_levels_c(rois, num_levels, finest_scale=56.0, recip_mode=0):
    r = _np32(rois)
    out = np.empty((r.shape[0],), np.int64)
    lib().orc_map_levels(_p(r), r.shape[0], int(num_levels), ctypes.c_float(finest_scale),
                         int(recip_mode), _p(out, _i64p))
    return out


def assign(rois, onehot, num_levels, finest_scale=56):
    """(lvl, bucket, perm, seg_offsets): stable grouping by bucket."""
    lvl = map_roi_levels(rois, num_levels, finest_scale).numpy()
    if onehot is None:
        bucket = np.zeros_like(lvl)
        nb = 1
    else:
        bucket = bucket_of(onehot).numpy()
        nb = int(torch.as_tensor(onehot).shape[1])
    perm = np.argsort(bucket, kind='stable').astype(np.int64)
    counts = np.bincount(bucket, minlength=nb)
    seg = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return lvl, bucket, perm, seg


# --------------------------------------------------------------------------------------------
# A12 roi_rescale -- .../roi_extractors/base_roi_extractor.py:57-79
# --------------------------------------------------------------------------------------------
def roi_rescale(rois, scale_factor):
    r = torch.as_tensor(rois, dtype=torch.float32)
    cx = (r[:, 1] + r[:, 3]) * 0.5
    cy = (r[:, 2] + r[:, 4]) * 0.5
    nw = (r[:, 3] - r[:, 1]) * scale_factor
    nh = (r[:, 4] - r[:, 2]) * scale_factor
    return torch.stack((r[:, 0], cx - nw * 0.5, cy - nh * 0.5, cx + nw * 0.5, cy + nh * 0.5), -1)


# --------------------------------------------------------------------------------------------
# A4/A5  RoIAlign forward / backward (mmcv.ops.roi_align; see dm_oracle.c header)
# --------------------------------------------------------------------------------------------
def _pair(v):
    return (int(v), int(v)) if np.isscalar(v) else (int(v[0]), int(v[1]))


def roi_align(feat, rois, output_size, spatial_scale=1.0, sampling_ratio=0, aligned=True):
    f = _np32(feat)
    r = _np32(rois).reshape(-1, 5)
    ph, pw = _pair(output_size)
    n, c, h, w = f.shape
    out = np.empty((r.shape[0], c, ph, pw), np.float32)
    lib().orc_roi_align_fwd(_p(f), n, c, h, w, _p(r), r.shape[0], ph, pw,
                            ctypes.c_float(spatial_scale), int(sampling_ratio), int(bool(aligned)),
                            _p(out))
    return torch.from_numpy(out)


def roi_align_backward(grad_out, rois, feat_shape, spatial_scale=1.0, sampling_ratio=0,
                       aligned=True):
    g = _np32(grad_out)
    r = _np32(rois).reshape(-1, 5)
    n, c, h, w = [int(v) for v in feat_shape]
    ph, pw = g.shape[2], g.shape[3]
    gi = np.empty((n, c, h, w), np.float32)
    lib().orc_roi_align_bwd(_p(g), n, c, h, w, _p(r), r.shape[0], ph, pw,
                            ctypes.c_float(spatial_scale), int(sampling_ratio), int(bool(aligned)),
                            _p(gi))
    return torch.from_numpy(gi)


def roi_align_tv(feat, rois, output_size, spatial_scale=1.0, sampling_ratio=0, aligned=True):
    """The stand-in library kernel (torchvision CPU); used for pinning and as the timed CPU path."""
    import torchvision.ops
    return torchvision.ops.roi_align(torch.as_tensor(feat, dtype=torch.float32),
                                     torch.as_tensor(rois, dtype=torch.float32),
                                     _pair(output_size), spatial_scale, sampling_ratio, aligned)


# --------------------------------------------------------------------------------------------
# A3  SingleRoIExtractor.forward -- .../single_level_roi_extractor.py:53-81
# --------------------------------------------------------------------------------------------
def single_roi_extractor(feats, rois, output_size, featmap_strides, finest_scale=56,
                         sampling_ratio=0, roi_scale_factor=None, kernel=roi_align):
    """Per-level select / align / scatter-back; unmatched RoIs stay zero."""
    rois = torch.as_tensor(rois, dtype=torch.float32)
    ph, pw = _pair(output_size)
    feats = [torch.as_tensor(f, dtype=torch.float32) for f in feats]
    out = torch.zeros((rois.shape[0], feats[0].shape[1], ph, pw), dtype=torch.float32)
    if len(feats) == 1:
        if rois.shape[0] == 0:
            return out
        return kernel(feats[0], rois, (ph, pw), 1.0 / featmap_strides[0], sampling_ratio, True)
    lvls = map_roi_levels(rois, len(feats), finest_scale)
    if roi_scale_factor is not None:
        rois = roi_rescale(rois, roi_scale_factor)
    for i, f in enumerate(feats):
        sel = lvls == i
        if bool(sel.any()):
            out[sel] = kernel(f, rois[sel], (ph, pw), 1.0 / featmap_strides[i], sampling_ratio,
                              True)
    return out


def single_roi_extractor_backward(grad_out, feats_shapes, rois, featmap_strides, finest_scale=56,
                                  sampling_ratio=0, roi_scale_factor=None):
    """Gradient w.r.t. every level's feature map of ``single_roi_extractor``."""
    rois = torch.as_tensor(rois, dtype=torch.float32)
    grad_out = torch.as_tensor(grad_out, dtype=torch.float32)
    if len(feats_shapes) == 1:
        return [roi_align_backward(grad_out, rois, feats_shapes[0], 1.0 / featmap_strides[0],
                                   sampling_ratio, True)]
    lvls = map_roi_levels(rois, len(feats_shapes), finest_scale)
    if roi_scale_factor is not None:
        rois = roi_rescale(rois, roi_scale_factor)
    grads = []
    for i, shp in enumerate(feats_shapes):
        sel = lvls == i
        grads.append(roi_align_backward(grad_out[sel], rois[sel], shp, 1.0 / featmap_strides[i],
                                        sampling_ratio, True))
    return grads


def bucketed_extract(feats, rois, onehot, bucket_sizes, featmap_strides, finest_scale=56,
                     sampling_ratio=0, kernel=roi_align):
    """North-star stage 1+2: each RoI aligned at the output size its one-hot label selects.

    Returns (list of [K_b,C,P_b,P_b] in original RoI order within each bucket, perm, seg)."""
    rois = torch.as_tensor(rois, dtype=torch.float32)
    _, bucket, perm, seg = assign(rois, onehot, len(feats), finest_scale)
    outs = []
    for b, p in enumerate(bucket_sizes):
        idx = torch.from_numpy(perm[seg[b]:seg[b + 1]])
        outs.append(single_roi_extractor(feats, rois[idx], p, featmap_strides, finest_scale,
                                         sampling_ratio, kernel=kernel))
    return outs, perm, seg


# --------------------------------------------------------------------------------------------
# A8/A9  BitmapMasks.crop_and_resize + mask_target -- mmdet/core/mask/structures.py:256-286,
#        mmdet/core/mask/mask_target.py:6-62, .../mask_heads/dynamask_head.py:246-271
# --------------------------------------------------------------------------------------------
def crop_and_resize(masks_u8, boxes, out_shape, inds, clip=False):
    """uint8 [G,H,W], boxes [K,4], inds [K] -> bool [K,S_h,S_w]."""
    m = np.ascontiguousarray(masks_u8, dtype=np.uint8)
    sh, sw = _pair(out_shape)
    if m.shape[0] == 0:
        return np.empty((0, sh, sw), dtype=np.uint8)
    b = _np32(boxes).reshape(-1, 4)
    i = np.ascontiguousarray(np.asarray(inds), dtype=np.int64)
    out = np.empty((b.shape[0], sh, sw), np.float32)
    lib().orc_mask_target(_p(m, _u8p), m.shape[0], m.shape[1], m.shape[2], _p(b), _p(i, _i64p),
                          b.shape[0], sh, sw, int(bool(clip)), _p(out))
    return out >= 0.5


def mask_target_single(pos_proposals, pos_assigned_gt_inds, gt_masks_u8, mask_size):
    sh, sw = _pair(mask_size)
    b = _np32(pos_proposals).reshape(-1, 4)
    if b.shape[0] == 0:
        return torch.zeros((0, sh, sw), dtype=torch.float32)
    t = crop_and_resize(gt_masks_u8, b, (sh, sw), pos_assigned_gt_inds, clip=True)
    return torch.from_numpy(t.astype(np.float32))


def mask_target(pos_proposals_list, pos_assigned_gt_inds_list, gt_masks_list, mask_size):
    ts = [mask_target_single(p, i, g, mask_size)
          for p, i, g in zip(pos_proposals_list, pos_assigned_gt_inds_list, gt_masks_list)]
    return torch.cat(ts) if ts else ts


def dyna_get_targets(pos_bboxes_list, pos_assigned_gt_inds_list, gt_masks_list,
                     stage_sup_size=(14, 28, 56, 112)):
    return [mask_target(pos_bboxes_list, pos_assigned_gt_inds_list, gt_masks_list, s)
            for s in stage_sup_size]


# --------------------------------------------------------------------------------------------
# A6/A7  _do_paste_mask + get_seg_masks -- .../mask_heads/fcn_mask_head.py:240-308,
#        .../mask_heads/dynamask_head.py:279-342
# --------------------------------------------------------------------------------------------
def paste_values_c(prob, boxes, img_h, img_w):
    """Un-thresholded full-canvas paste from probabilities, separable C restatement."""
    p = _np32(prob)
    p = p.reshape(p.shape[0], p.shape[-2], p.shape[-1])
    b = _np32(boxes).reshape(-1, 4)
    out = np.empty((p.shape[0], int(img_h), int(img_w)), np.float32)
    lib().orc_paste(_p(p), p.shape[0], p.shape[1], p.shape[2], _p(b), int(img_h), int(img_w),
                    _p(out))
    return torch.from_numpy(out)


def do_paste_mask(masks, boxes, img_h, img_w, skip_empty=True):
    """grid_sample form (the reference's own formulation) on CPU tensors."""
    masks = torch.as_tensor(masks, dtype=torch.float32)
    boxes = torch.as_tensor(boxes, dtype=torch.float32)
    if skip_empty:
        lo = torch.clamp(boxes.min(dim=0).values.floor()[:2] - 1, min=0).to(torch.int32)
        x_lo, y_lo = int(lo[0]), int(lo[1])
        x_hi = int(torch.clamp(boxes[:, 2].max().ceil() + 1, max=img_w).to(torch.int32))
        y_hi = int(torch.clamp(boxes[:, 3].max().ceil() + 1, max=img_h).to(torch.int32))
    else:
        x_lo, y_lo, x_hi, y_hi = 0, 0, int(img_w), int(img_h)
    bx0, by0, bx1, by1 = torch.split(boxes, 1, dim=1)
    n = masks.shape[0]
    ys = torch.arange(y_lo, y_hi, dtype=torch.float32) + 0.5
    xs = torch.arange(x_lo, x_hi, dtype=torch.float32) + 0.5
    ys = (ys - by0) / (by1 - by0) * 2 - 1
    xs = (xs - bx0) / (bx1 - bx0) * 2 - 1
    xs[torch.isinf(xs)] = 0
    ys[torch.isinf(ys)] = 0
    gx = xs[:, None, :].expand(n, ys.size(1), xs.size(1))
    gy = ys[:, :, None].expand(n, ys.size(1), xs.size(1))
    out = F.grid_sample(masks, torch.stack([gx, gy], dim=3), align_corners=False)
    if skip_empty:
        return out[:, 0], (slice(y_lo, y_hi), slice(x_lo, x_hi))
    return out[:, 0], ()


def get_seg_masks(mask_pred, det_bboxes, det_labels, mask_thr_binary, ori_shape, scale_factor,
                  rescale, device_mode='cpu'):
    """DynaMaskHead.get_seg_masks on CPU: list of N numpy [img_h,img_w] (bool, or uint8 if thr<0).

    ``device_mode='cpu'`` follows the reference's CPU branch (one instance per chunk,
    ``skip_empty=True``, dynamask_head.py:301-305); ``'gpu'`` follows its CUDA branch (full canvas,
    ``skip_empty=False``, :306-312).  The two agree except for degenerate boxes (x1 == x0 or
    y1 == y0), where the inf->0 patch makes the full-canvas mode paint whole rows / columns."""
    prob = torch.as_tensor(mask_pred, dtype=torch.float32).sigmoid()
    det_bboxes = torch.as_tensor(det_bboxes, dtype=torch.float32)
    boxes = det_bboxes[:, :4]
    if rescale:
        img_h, img_w = int(ori_shape[0]), int(ori_shape[1])
    else:
        img_h = int(np.round(ori_shape[0] * scale_factor).astype(np.int32))
        img_w = int(np.round(ori_shape[1] * scale_factor).astype(np.int32))
        scale_factor = 1.0
    if not isinstance(scale_factor, (float, torch.Tensor)):
        scale_factor = boxes.new_tensor(scale_factor)
    boxes = boxes / scale_factor
    n = prob.shape[0]
    if prob.shape[1] > 1:
        prob = prob[range(n), torch.as_tensor(det_labels)][:, None]
    thr = mask_thr_binary
    canvas = torch.zeros(n, img_h, img_w, dtype=torch.bool if thr >= 0 else torch.uint8)
    for i in range(n):  # one instance per chunk
        chunk, sl = do_paste_mask(prob[i:i + 1], boxes[i:i + 1], img_h, img_w,
                                  skip_empty=(device_mode == 'cpu'))
        if thr >= 0:
            chunk = (chunk >= thr).to(torch.bool)
        else:
            chunk = (chunk * 255).to(torch.uint8)
        canvas[(torch.tensor([i]),) + sl] = chunk
    return [canvas[i].numpy() for i in range(n)]


# ------------------------------------------------------------------------------------------------
# COCO run-length encoding.  The reference calls pycocotools
# (mmdet/core/mask/utils.py:36-63: mask_util.encode(np.array(m[:, :, None], order='F', dtype='uint8'))[0]);
# pycocotools (2.0.x, third party, absent from the reference tree and from this image) is restated
# here from its published algorithm (common/maskApi.c: rleEncode, rleToString, rleFrString,
# rleDecode).  PARITY UNPINNED against the real library: no pycocotools, no golden strings in the
# reference's tests; pinned only by encode/decode round trips and hand-checked small cases.
# ------------------------------------------------------------------------------------------------
def rle_counts(mask):
    """rleEncode: run lengths of the column-major flattening, starting with a run of zeros."""
    flat = np.asarray(mask).astype(np.uint8).reshape(-1, order='F') != 0
    if flat.size == 0:
        return [0]
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    bounds = np.concatenate([[0], change, [flat.size]])
    counts = np.diff(bounds).tolist()
    if flat[0]:
        counts = [0] + counts
    return counts


def rle_to_string(counts):
    """rleToString: base-32 varint, each count beyond the second stored relative to two before."""
    out = bytearray()
    for i, c in enumerate(counts):
        x = int(c)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            ch = x & 0x1f
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(ch + 48)
    return bytes(out)


def rle_from_string(s):
    """rleFrString: inverse of rle_to_string."""
    counts = []
    p = 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = s[p] - 48
            x |= (c & 0x1f) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    return counts


def rle_encode(mask):
    """pycocotools.mask.encode for one [H,W] mask: {'size': [h, w], 'counts': bytes}."""
    h, w = np.asarray(mask).shape
    return {'size': [int(h), int(w)], 'counts': rle_to_string(rle_counts(mask))}


def rle_decode(rle):
    """pycocotools.mask.decode for one RLE dict -> [H,W] uint8."""
    h, w = rle['size']
    counts = rle_from_string(rle['counts'])
    flat = np.zeros(h * w, np.uint8)
    pos, v = 0, 0
    for c in counts:
        if v:
            flat[pos:pos + c] = 1
        pos += c
        v ^= 1
    return flat.reshape((h, w), order='F')


# --------------------------------------------------------------------------------------------
# SURVEY 8f rank 2  SimpleRoIAlign -- constructed at dynamask_head.py:74, called at :104-105.
# The class lives in mmcv (mmcv/ops/point_sample.py: generate_grid, rel_roi_point_to_abs_img_point,
# abs_img_point_to_rel_img_point, point_sample, SimpleRoIAlign), a third-party dependency absent
# from /root/reference: this restates its published algorithm with the same torch calls
# (F.affine_grid, F.grid_sample); PARITY UNPINNED at the mmcv boundary, pinned only to torch's own
# grid_sample.  Output rows follow the input RoI order (mmcv concatenates per image, which is the
# same thing for RoIs sorted by image as bbox2roi produces).
# --------------------------------------------------------------------------------------------
def _generate_grid(num_grid, size):
    affine = torch.tensor([[[1., 0., 0.], [0., 1., 0.]]])
    grid = F.affine_grid(affine, torch.Size((1, 1, *size)), align_corners=False)
    grid = (grid + 1.0) / 2.0                                   # normalize
    return grid.view(1, -1, 2).expand(num_grid, -1, -1)


def _rel_roi_point_to_rel_img_point(rois, rel_roi_points, img_hw, spatial_scale):
    r = rois[:, 1:] if rois.size(1) == 5 else rois
    xs = rel_roi_points[:, :, 0] * (r[:, None, 2] - r[:, None, 0])
    ys = rel_roi_points[:, :, 1] * (r[:, None, 3] - r[:, None, 1])
    xs = xs + r[:, None, 0]
    ys = ys + r[:, None, 1]
    abs_pts = torch.stack([xs, ys], dim=2)
    h, w = img_hw
    scale = torch.tensor([w, h], dtype=torch.float).view(1, 1, 2)
    return abs_pts / scale * spatial_scale


def _point_sample(inp, points, align_corners=False):
    add_dim = points.dim() == 3
    if add_dim:
        points = points.unsqueeze(2)
    out = F.grid_sample(inp, points * 2.0 - 1.0, align_corners=align_corners)
    return out.squeeze(3) if add_dim else out


def simple_roi_align(features, rois, output_size, spatial_scale, aligned=True):
    """features [B,C,H,W], rois [K,5] -> [K,C,ph,pw] (mmcv SimpleRoIAlign.forward)."""
    features = torch.as_tensor(features, dtype=torch.float32)
    rois = torch.as_tensor(rois, dtype=torch.float32)
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    K, C = rois.size(0), features.size(1)
    out = features.new_zeros((K, C, ph, pw))
    rel = _generate_grid(K, (ph, pw))
    for b in range(features.size(0)):
        inds = rois[:, 0].long() == b
        if inds.any():
            feat = features[b:b + 1]
            pts = _rel_roi_point_to_rel_img_point(rois[inds], rel[inds], feat.shape[2:],
                                                  spatial_scale).unsqueeze(0)
            pf = _point_sample(feat, pts, align_corners=not aligned)        # [1,C,k,ph*pw]
            out[inds] = pf.squeeze(0).transpose(0, 1).reshape(-1, C, ph, pw)
    return out


def simple_roi_align_backward(grad_out, feat_shape, rois, spatial_scale, aligned=True):
    """Gradient of simple_roi_align w.r.t. the feature map, by torch autograd on the CPU."""
    f = torch.zeros(tuple(feat_shape), dtype=torch.float32, requires_grad=True)
    go = torch.as_tensor(grad_out, dtype=torch.float32)
    out = simple_roi_align(f, rois, tuple(go.shape[2:]), spatial_scale, aligned)
    out.backward(go)
    return f.grad


def simple_roi_align_f64(features, rois, output_size, spatial_scale, aligned=True):
    """Closed form of simple_roi_align in float64 (no normalised-coordinate round trip): the
    yardstick for how much of a difference is fp32 coordinate noise."""
    f = torch.as_tensor(features).double().numpy()
    r = torch.as_tensor(rois).double().numpy()
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    B, C, H, W = f.shape
    out = np.zeros((r.shape[0], C, ph, pw))
    for k in range(r.shape[0]):
        b = int(r[k, 0])
        if not 0 <= b < B:
            continue
        x = (r[k, 1] + (np.arange(pw) + .5) / pw * (r[k, 3] - r[k, 1])) * spatial_scale
        y = (r[k, 2] + (np.arange(ph) + .5) / ph * (r[k, 4] - r[k, 2])) * spatial_scale
        if aligned:
            x, y = x - .5, y - .5
        else:
            x, y = x / W * (W - 1), y / H * (H - 1)
        x0, y0 = np.floor(x).astype(np.int64), np.floor(y).astype(np.int64)
        lx, ly = x - x0, y - y0
        acc = np.zeros((C, ph, pw))
        for yy, wy in ((y0, 1 - ly), (y0 + 1, ly)):
            for xx, wx in ((x0, 1 - lx), (x0 + 1, lx)):
                oky = (yy >= 0) & (yy < H)
                okx = (xx >= 0) & (xx < W)
                v = f[b][:, np.clip(yy, 0, H - 1)][:, :, np.clip(xx, 0, W - 1)]
                acc += v * (wy * oky)[None, :, None] * (wx * okx)[None, None, :]
        out[k] = acc
    return torch.from_numpy(out)


# --------------------------------------------------------------------------------------------
# SURVEY 8f rank 3  stage-to-stage refinement at inference --
# mmdet/models/roi_heads/dynamask_roi_head.py:136-148 with generate_block_target,
# mmdet/models/losses/cross_entropy_loss.py:123-154.  Pinned by tests/golden/refine.npz (the
# reference's own source lines executed through ref_shim.load_refine).
# --------------------------------------------------------------------------------------------
def generate_block_target(mask_target, boundary_width=3):
    """0 = background, 1 = boundary band, 2 = interior foreground (two box-Laplacian convolutions)."""
    mask_target = torch.as_tensor(mask_target).float()
    k = 2 * boundary_width + 1
    lap = -torch.ones(1, 1, k, k)
    lap[0, 0, boundary_width, boundary_width] = k ** 2 - 1
    pad = F.pad(mask_target.unsqueeze(1), (boundary_width,) * 4, 'constant', 0)
    pos = F.conv2d(pad, lap, padding=0).clamp(min=0) / float(k ** 2)
    pos = (pos > 0.1).float().squeeze(1)
    neg = F.conv2d(1 - pad, lap, padding=0).clamp(min=0) / float(k ** 2)
    neg = (neg > 0.1).float().squeeze(1)
    block = torch.zeros_like(mask_target).long()
    block[(pos + neg) > 0] = 1
    block[(mask_target - pos) > 0] = 2
    return block


def non_boundary_3x3(m):
    """Closed form of ``generate_block_target(m, 1) != 1`` used by the kernel: a foreground pixel is
    boundary when its zero-padded 3x3 window holds a 0, a background pixel when it holds a 1."""
    m = np.asarray(m).astype(np.int32)
    n, h, w = m.shape
    pad = np.zeros((n, h + 2, w + 2), np.int32)
    pad[:, 1:-1, 1:-1] = m
    ones = sum(pad[:, dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3))
    boundary = np.where(m == 1, ones < 9, ones > 0)
    return ~boundary


def refine_stage_preds(stage_preds):
    """``stage_preds``: list of [N,1,S,S] logits (coarse to fine, e.g. 28/56/112).  Returns the
    refined copies (stage 0 unchanged); the last one is what get_seg_masks receives."""
    preds = [torch.as_tensor(p, dtype=torch.float32).clone() for p in stage_preds]
    for idx in range(len(preds) - 1):
        inst = preds[idx].squeeze(1).sigmoid() >= 0.5
        nb = (generate_block_target(inst, boundary_width=1) != 1).unsqueeze(1)
        nb = F.interpolate(nb.float(), preds[idx + 1].shape[-2:], mode='bilinear', align_corners=True) >= 0.5
        pre = F.interpolate(preds[idx], preds[idx + 1].shape[-2:], mode='bilinear', align_corners=True)
        preds[idx + 1][nb] = pre[nb]
    return preds


# --------------------------------------------------------------------------------------------
# SURVEY row A10 / 8f rank 4  polygon ground truth -> mask targets
#   PolygonMasks.crop_and_resize   mmdet/core/mask/structures.py:465-499
#   PolygonMasks.to_ndarray / polygon_to_bitmap   structures.py:541-575 (pycocotools, absent:
#   restated in dm_oracle.c from common/maskApi.c -- PARITY UNPINNED against the real library)
# --------------------------------------------------------------------------------------------
def polygon_to_bitmap(polygons, height, width):
    """frPyObjects -> merge -> decode of one object's polygons: bool [height,width]."""
    polys = [np.ascontiguousarray(p, dtype=np.float64).reshape(-1) for p in polygons]
    polys = [p[:2 * (p.size // 2)] for p in polys]
    voff = np.zeros(len(polys) + 1, np.int64)
    np.cumsum([p.size // 2 for p in polys], out=voff[1:])
    xy = np.concatenate(polys) if polys else np.zeros(0, np.float64)
    out = np.zeros((int(height), int(width)), np.uint8)
    f = lib().orc_polygons_to_bitmap
    f.restype = ctypes.c_int
    rc = f(xy.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), _p(voff, _i64p), ctypes.c_long(len(polys)),
           ctypes.c_long(int(height)), ctypes.c_long(int(width)), _p(out, _u8p))
    assert rc == 0
    return out.astype(bool)


def polygon_crop_and_resize(masks, bboxes, out_shape, inds):
    """The host arithmetic of PolygonMasks.crop_and_resize: list (per box) of lists of polygons."""
    out_h, out_w = out_shape
    bboxes = np.asarray(bboxes, dtype=np.float32)
    res = []
    for i in range(len(bboxes)):
        bbox = bboxes[i, :]
        x1, y1, x2, y2 = bbox
        w = np.maximum(x2 - x1, 1)
        h = np.maximum(y2 - y1, 1)
        h_scale = out_h / max(h, 0.1)
        w_scale = out_w / max(w, 0.1)
        obj = []
        for p in masks[int(inds[i])]:
            p = np.array(p, dtype=np.float64)
            p[0::2] -= bbox[0]
            p[1::2] -= bbox[1]
            p[0::2] *= w_scale
            p[1::2] *= h_scale
            obj.append(p)
        res.append(obj)
    return res


def polygon_mask_target_single(pos_proposals, pos_assigned_gt_inds, masks, height, width, mask_size):
    """mask_target_single (mask_target.py:30-62) for PolygonMasks: float32 [K,S,S] in {0,1}."""
    S = (mask_size, mask_size) if isinstance(mask_size, int) else tuple(mask_size)
    prop = _np32(pos_proposals)[:, :4].copy()
    K = prop.shape[0]
    if K == 0:
        return torch.zeros((0, ) + S)
    prop[:, [0, 2]] = np.clip(prop[:, [0, 2]], 0, width)
    prop[:, [1, 3]] = np.clip(prop[:, [1, 3]], 0, height)
    inds = np.asarray(pos_assigned_gt_inds).astype(np.int64)
    resized = polygon_crop_and_resize(masks, prop, S, inds)
    out = np.stack([polygon_to_bitmap(obj, S[0], S[1]) for obj in resized])
    return torch.from_numpy(out).float()
