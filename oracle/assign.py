"""Oracle: teacher keep-ids (A1), GFL Hungarian assignment (H1-H3), matched ids (A2).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Restates
  mmdet/core/utils/misc.py:143-152                                  filter_scores_and_topk
  mmdet/models/dense_heads/gfl_deformable_detr_head_il.py:1622-1668 teacher decode
  mmdet/models/detectors/deformable_detr_il.py:138-151              keep-id flattening
  mmdet/core/bbox/match_costs/match_cost.py:34-51,215-230,460-476   L1 / QFL / IoU costs
  mmdet/core/bbox/assigners/gfl_hungarian_assigner.py:102-160       assign
  mmdet/core/bbox/samplers/pseudo_sampler.py:35-41                  pos/neg split
  mmdet/models/dense_heads/gfl_deformable_detr_head_il.py:1417-1455,1765-1797 targets
"""
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

from .boxes import bbox_overlaps, cxcywh_to_xyxy, xyxy_to_cxcywh, decode_cxcywh


# ----------------------------------------------------------------------------- A1
def filter_scores_and_topk(scores, score_thr, topk):
    """misc.py:143-152 -- threshold, sort descending, keep <= topk (query, class) pairs."""
    valid = scores > score_thr
    kept = scores[valid]
    valid_idxs = torch.nonzero(valid)
    num = min(topk, valid_idxs.size(0))
    kept, order = kept.sort(descending=True)
    top = valid_idxs[order[:num]]
    keep_idxs, labels = top.unbind(dim=1)
    return kept[:num], labels, keep_idxs


def teacher_decode_single(cls_logits, box_pred, img_shape, score_thr=0.3, max_per_img=100, reg_max=16):
    """head_il.py:1622-1668 (sigmoid branch, rescale=False, need_logits=True).

    Returns (bboxes[K,4] px xyxy clamped, scores[K], labels[K], keep_id[K])."""
    prob = cls_logits.sigmoid()
    scores, labels, keep = filter_scores_and_topk(prob, score_thr, max_per_img)
    cxcywh = decode_cxcywh(box_pred[keep], reg_max)
    det = cxcywh_to_xyxy(cxcywh)
    det[:, 0::2] = det[:, 0::2] * img_shape[1]
    det[:, 1::2] = det[:, 1::2] * img_shape[0]
    det[:, 0::2].clamp_(min=0, max=img_shape[1])
    det[:, 1::2].clamp_(min=0, max=img_shape[0])
    return det, scores, labels, keep


def teacher_info_from_outputs(cls_logits, box_pred, img_shapes, score_thr=0.3, max_per_img=100, reg_max=16):
    """deformable_detr_il.py:138-151 on the last decoder layer: cls [N,Q,80], box [N,Q,70]."""
    n, q = cls_logits.shape[:2]
    bboxes, scores, labels, keep = [], [], [], []
    for i in range(n):
        b, s, l, k = teacher_decode_single(cls_logits[i], box_pred[i], img_shapes[i],
                                           score_thr, max_per_img, reg_max)
        bboxes.append(b), scores.append(s), labels.append(l), keep.append(k + i * q)
    return dict(pred_bboxes=bboxes, pred_scores=scores, pred_labels=labels,
                pred_keepid=torch.cat(keep))


# ----------------------------------------------------------------------------- H1
def cost_matrix(bbox_cxcywh, cls_logits, gt_bboxes, gt_labels, img_hw,
                w_cls=2.0, w_reg=5.0, w_iou=2.0, beta=2.0):
    """gfl_hungarian_assigner.py:120-140 for one (layer, image): [Q, G] fp32."""
    img_h, img_w = img_hw
    factor = gt_bboxes.new_tensor([img_w, img_h, img_w, img_h]).unsqueeze(0)
    gt_norm = gt_bboxes / factor
    # BBoxL1Cost(box_format='xywh'): match_cost.py:34-51
    reg = torch.cdist(bbox_cxcywh, xyxy_to_cxcywh(gt_norm), p=1) * w_reg
    # IoUCost('giou'): match_cost.py:460-476
    iou = -bbox_overlaps(cxcywh_to_xyxy(bbox_cxcywh) * factor, gt_bboxes, mode='giou') * w_iou
    # QualityFocalLossCost: match_cost.py:215-230
    sig = cls_logits.sigmoid()
    score = bbox_overlaps(cxcywh_to_xyxy(bbox_cxcywh), gt_norm)
    scale = score - sig[:, gt_labels]
    cls = F.binary_cross_entropy_with_logits(cls_logits[:, gt_labels], score,
                                             reduction='none') * scale.abs().pow(beta) * w_cls
    return cls + reg + iou


# ----------------------------------------------------------------------------- H2
def hungarian_assign(cost, gt_labels, num_query):
    """gfl_hungarian_assigner.py:102-119,142-160 given the cost; returns
    (assigned_gt_inds[Q] 1-based / 0 bg, assigned_labels[Q] / -1)."""
    gt_inds = torch.full((num_query,), -1, dtype=torch.long)
    labels = torch.full((num_query,), -1, dtype=torch.long)
    num_gts = gt_labels.numel()
    if num_gts == 0 or num_query == 0:
        if num_gts == 0:
            gt_inds[:] = 0
        return gt_inds, labels
    row, col = linear_sum_assignment(cost.detach().cpu())
    row = torch.from_numpy(row)
    col = torch.from_numpy(col)
    gt_inds[:] = 0
    gt_inds[row] = col + 1
    labels[row] = gt_labels[col]
    return gt_inds, labels


# ----------------------------------------------------------------------------- H3
def targets_single(cls_logits, bbox_cxcywh, gt_bboxes, gt_labels, img_hw, num_classes=80, **cost_kw):
    """head_il.py:1765-1797 + pseudo_sampler.py:35-41 for one image."""
    q = bbox_cxcywh.size(0)
    if gt_labels.numel() and q:
        cost = cost_matrix(bbox_cxcywh, cls_logits, gt_bboxes, gt_labels, img_hw, **cost_kw)
    else:
        cost = None
    gt_inds, _ = hungarian_assign(cost, gt_labels, q)
    pos = torch.nonzero(gt_inds > 0, as_tuple=False).squeeze(-1).unique()
    neg = torch.nonzero(gt_inds == 0, as_tuple=False).squeeze(-1).unique()
    pos_gt = gt_inds[pos] - 1
    labels = torch.full((q,), num_classes, dtype=torch.long)
    labels[pos] = gt_labels[pos_gt]
    label_weights = torch.ones(q)
    bbox_targets = torch.zeros_like(bbox_cxcywh)
    bbox_weights = torch.zeros_like(bbox_cxcywh)
    bbox_weights[pos] = 1.0
    img_h, img_w = img_hw
    factor = bbox_cxcywh.new_tensor([img_w, img_h, img_w, img_h]).unsqueeze(0)
    if pos.numel():
        bbox_targets[pos] = xyxy_to_cxcywh(gt_bboxes.view(-1, 4)[pos_gt] / factor)
    return labels, label_weights, bbox_targets, bbox_weights, pos, neg, gt_inds


def layer_targets(cls_logits, box_pred, gt_bboxes_list, gt_labels_list, img_shapes,
                  prev_labels, num_classes=80, reg_max=16, **cost_kw):
    """head_il.py:1417-1455 for ONE decoder layer: cls [N,Q,80], box [N,Q,70].

    Returns dict(labels[N*Q], label_weights, bbox_targets[N*Q,4], bbox_weights,
    teacher_only_weights[N*Q], gt_inds[N*Q])."""
    n = cls_logits.size(0)
    cxcywh = decode_cxcywh(box_pred, reg_max)
    outs = [targets_single(cls_logits[i], cxcywh[i], gt_bboxes_list[i], gt_labels_list[i],
                           img_shapes[i], num_classes, **cost_kw) for i in range(n)]
    labels = torch.cat([o[0] for o in outs])
    label_weights = torch.cat([o[1] for o in outs])
    teacher_only = label_weights.new_zeros(label_weights.shape)
    for t in prev_labels:                      # head_il.py:1453-1455
        teacher_only[labels == t, ...] = 1
    return dict(labels=labels, label_weights=label_weights,
                bbox_targets=torch.cat([o[2] for o in outs]),
                bbox_weights=torch.cat([o[3] for o in outs]),
                teacher_only_weights=teacher_only,
                gt_inds=torch.cat([o[6] for o in outs]))


def merge_pseudo_labels(teacher_bboxes, teacher_labels, gt_bboxes_list, gt_labels_list):
    """head_il.py:462-465 -- pseudo GT = cat(teacher predictions, real GT)."""
    boxes = [torch.cat([t, g], 0) for t, g in zip(teacher_bboxes, gt_bboxes_list)]
    labels = [torch.cat([t, g], 0) for t, g in zip(teacher_labels, gt_labels_list)]
    return boxes, labels
