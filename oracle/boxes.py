"""Oracle: box arithmetic used by the assignment costs and the teacher decode.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Restates
  mmdet/core/bbox/transforms.py:245-270                  (cxcywh <-> xyxy)
  mmdet/core/bbox/iou_calculators/iou2d_calculator.py:192-261  (bbox_overlaps)
  mmdet/models/dense_heads/gfl_deformable_detr_head_il.py:23-60 (Integral_average)
"""
import torch


def cxcywh_to_xyxy(b):
    """transforms.py:245-256."""
    cx, cy, w, h = b.split((1, 1, 1, 1), dim=-1)
    return torch.cat([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def xyxy_to_cxcywh(b):
    """transforms.py:259-270."""
    x1, y1, x2, y2 = b.split((1, 1, 1, 1), dim=-1)
    return torch.cat([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], dim=-1)


def bbox_overlaps(b1, b2, mode='iou', is_aligned=False, eps=1e-6):
    """iou2d_calculator.py:192-261 (fp32 path; `fp16_clamp` == clamp for fp32)."""
    assert mode in ('iou', 'giou')
    rows, cols = b1.size(-2), b2.size(-2)
    if rows * cols == 0:
        return b1.new_zeros((rows,) if is_aligned else (rows, cols))
    a1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    a2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    if is_aligned:
        lt = torch.max(b1[..., :2], b2[..., :2])
        rb = torch.min(b1[..., 2:], b2[..., 2:])
        wh = (rb - lt).clamp(min=0)
        overlap = wh[..., 0] * wh[..., 1]
        union = a1 + a2 - overlap
        e_lt = torch.min(b1[..., :2], b2[..., :2])
        e_rb = torch.max(b1[..., 2:], b2[..., 2:])
    else:
        lt = torch.max(b1[..., :, None, :2], b2[..., None, :, :2])
        rb = torch.min(b1[..., :, None, 2:], b2[..., None, :, 2:])
        wh = (rb - lt).clamp(min=0)
        overlap = wh[..., 0] * wh[..., 1]
        union = a1[..., None] + a2[..., None, :] - overlap
        e_lt = torch.min(b1[..., :, None, :2], b2[..., None, :, :2])
        e_rb = torch.max(b1[..., :, None, 2:], b2[..., None, :, 2:])
    eps_t = union.new_tensor([eps])
    union = torch.max(union, eps_t)
    ious = overlap / union
    if mode == 'iou':
        return ious
    e_wh = (e_rb - e_lt).clamp(min=0)
    e_area = torch.max(e_wh[..., 0] * e_wh[..., 1], eps_t)
    return ious - (e_area - union) / e_area


def integral_average(lrtb, reg_max=16):
    """head_il.py:42-59 -- normalise each (reg_max+1)-bin group by its sum (no softmax),
    expectation over bins {0..reg_max}/reg_max/2, then (l+r, t+b)."""
    x = lrtb.reshape(-1, reg_max + 1)
    x = x / x.sum(1).unsqueeze(1).repeat(1, reg_max + 1)
    space = torch.linspace(0, reg_max, reg_max + 1).to(x.device)
    space = space / reg_max / 2
    x = x * space
    return x.sum(1).reshape(-1, 2, 2).sum(2)


def decode_cxcywh(box_pred, reg_max=16):
    """head_il.py:1427-1432: `[...,:2]` centre + Integral_average of the 4*(reg_max+1) bins."""
    lead = box_pred.shape[:-1]
    wh = integral_average(box_pred[..., 2:], reg_max).reshape(*lead, 2)
    return torch.cat((box_pred[..., :2], wh), dim=-1)
