"""Teacher keep-ids on the GPU: the part of `DeformableDETR_il.out_teacher`
(mmdet/models/detectors/deformable_detr_il.py:116-154) that turns the frozen teacher's last-layer outputs
into the `teacher_info` entries the distillation losses consume:

    get_bboxes(..., cfg=teacher_test_cfg(score_thr=0.3, max_per_img=100), need_logits=True)
      -> per image `_get_bboxes_single` (gfl_deformable_detr_head_il.py:1622-1668)
         -> `filter_scores_and_topk` (core/utils/misc.py:143-152)
    pred_keepid = cat(keep_i + i * num_query)                       (deformable_detr_il.py:151)

The reference runs this per image in Python (sigmoid, mask, nonzero, sort, gather); here it is one kernel launch
for the whole batch plus one compaction launch, and ONE 4*(N+1)-byte device->host copy to learn the ragged sizes.
"""
from typing import Dict, Optional, Sequence

import torch

from . import _lib as L


@L.guarded
def teacher_info_from_outputs(cls_scores: torch.Tensor, bbox_preds: torch.Tensor, img_shapes, score_thr: float = 0.3,
                              max_per_img: int = 100, reg_max: int = 16, need_logits: bool = False,
                              split: bool = True) -> Dict[str, object]:
    """cls_scores: [N,Q,num_classes] logits, bbox_preds: [N,Q,2+4*(reg_max+1)] sigmoid outputs of the teacher's
    last decoder layer (`head_outs[0][-1]`, `head_outs[1][-1]`); img_shapes: [N,2] (h,w).

    Returns the reference's keys: `pred_bboxes` (list of [K_i,4] px xyxy), `pred_scores`, `pred_labels` (lists),
    `pred_keepid` ([sum K] = q + Q*i), optionally `pred_logits` (list of [K_i,num_classes] sigmoid scores), plus
    the device-side ragged form the kernels take directly: `box_start` (int32 [N+1]), `cat_bboxes`, `cat_labels`.
    With split=False the per-image lists are not materialised and no host sync happens (`num_pairs` is then
    unknown on the host: the caller reads `box_start[-1]` when it needs it)."""
    lib = L.load()
    L.require_device(cls_scores)
    cls = L.f32c(cls_scores.detach())
    box = L.f32c(bbox_preds.detach())
    if cls.dim() != 3 or box.dim() != 3 or cls.shape[:2] != box.shape[:2]:
        raise L.DskdError(f'expected cls [N,Q,classes] and box [N,Q,ch], got {tuple(cls.shape)} / {tuple(box.shape)}')
    N, Q, nc = cls.shape
    want_ch = 2 + 4 * (reg_max + 1) if reg_max > 0 else 4
    if box.shape[2] != want_ch:
        raise L.DskdError(f'box channels {box.shape[2]} != {want_ch} for reg_max={reg_max} (head_il.py:153)')
    dev = cls.device
    hw = img_shapes if isinstance(img_shapes, torch.Tensor) else torch.tensor([list(s)[:2] for s in img_shapes])
    hw = hw.to(dev, torch.int32).contiguous()
    if hw.shape != (N, 2):
        raise L.DskdError(f'img_shapes must be [N,2] (h,w), got {tuple(hw.shape)}')
    M = int(max_per_img)
    st = L.stream_of(cls)
    count = torch.empty(N, dtype=torch.int32, device=dev)
    boxes = torch.empty(N, M, 4, dtype=torch.float32, device=dev)
    scores = torch.empty(N, M, dtype=torch.float32, device=dev)
    labels = torch.empty(N, M, dtype=torch.int64, device=dev)
    keepid = torch.empty(N, M, dtype=torch.int64, device=dev)
    logits = torch.zeros(N, M, nc, dtype=torch.float32, device=dev) if need_logits else None
    L.check(lib.dskd_teacher_decode(L.ptr(cls), L.ptr(box), N, Q, nc, reg_max, L.ptr(hw), float(score_thr), M,
                                    L.ptr(count), L.ptr(boxes), L.ptr(scores), L.ptr(labels), L.ptr(keepid),
                                    L.ptr(logits), st), 'dskd_teacher_decode')
    start = torch.empty(N + 1, dtype=torch.int32, device=dev)
    cat_boxes = torch.empty(N * M, 4, dtype=torch.float32, device=dev)
    cat_scores = torch.empty(N * M, dtype=torch.float32, device=dev)
    cat_labels = torch.empty(N * M, dtype=torch.int64, device=dev)
    cat_keepid = torch.empty(N * M, dtype=torch.int64, device=dev)
    L.check(lib.dskd_teacher_compact(L.ptr(count), N, M, L.ptr(boxes), L.ptr(scores), L.ptr(labels), L.ptr(keepid),
                                     L.ptr(start), L.ptr(cat_boxes), L.ptr(cat_scores), L.ptr(cat_labels),
                                     L.ptr(cat_keepid), st), 'dskd_teacher_compact')
    out: Dict[str, object] = dict(count=count, box_start=start, padded_bboxes=boxes, padded_scores=scores,
                                  padded_labels=labels, padded_keepid=keepid)
    if not split:
        out.update(cat_bboxes=cat_boxes, cat_scores=cat_scores, cat_labels=cat_labels, pred_keepid=cat_keepid)
        return out
    s = start.tolist()                                   # the one host sync
    total = s[-1]
    out.update(cat_bboxes=cat_boxes[:total], cat_scores=cat_scores[:total], cat_labels=cat_labels[:total],
               pred_keepid=cat_keepid[:total],
               pred_bboxes=[cat_boxes[s[i]:s[i + 1]] for i in range(N)],
               pred_scores=[cat_scores[s[i]:s[i + 1]] for i in range(N)],
               pred_labels=[cat_labels[s[i]:s[i + 1]] for i in range(N)])
    if need_logits:
        out['pred_logits'] = [logits[i, :s[i + 1] - s[i]] for i in range(N)]
    return out
