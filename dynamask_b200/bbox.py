"""RoI wire format shared by every kernel (reference: ``mmdet/core/bbox/transforms.py:54-73``)."""
import torch


def bbox2roi(bbox_list):
    """list of per-image ``[k_i, 4+]`` xyxy boxes -> ``[K, 5]`` rows (image index, x1, y1, x2, y2).

    The image index is stored as a float in column 0, in image order.
    """
    rows = []
    for img_id, boxes in enumerate(bbox_list):
        if boxes.size(0) > 0:
            idx = boxes.new_full((boxes.size(0), 1), img_id)
            rows.append(torch.cat([idx, boxes[:, :4]], dim=-1))
        else:
            rows.append(boxes.new_zeros((0, 5)))
    return torch.cat(rows, 0)
