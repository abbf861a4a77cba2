"""GFL Hungarian assignment for DSKD's incremental head on B200.

Reference: mmdet/core/bbox/assigners/gfl_hungarian_assigner.py:59-160 (`GFLHungarianAssigner.assign`),
match costs mmdet/core/bbox/match_costs/match_cost.py:34-51,215-230,460-476, target construction
gfl_deformable_detr_head_il.py:1417-1455,1765-1797.

The reference solves 6*N problems one by one: per (layer, image) a handful of ATen ops, a forced
`.cpu()` sync and one SciPy call.  Here all cost matrices come from ONE kernel launch
(`dskd_cost_matrix`), ONE pinned device->host copy, a multi-threaded C++ solver that reproduces
SciPy's indices (`dskd_lsap_batch_f32`), ONE host->device copy and ONE target kernel
(`dskd_assign_targets`).
"""
import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib as L
from .registry import ASSIGNERS


class AssignResult:
    """The fields of mmdet's AssignResult that the IL head reads (assign_result.py)."""

    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts = num_gts
        self.gt_inds = gt_inds
        self.max_overlaps = max_overlaps
        self.labels = labels


def lsap(cost) -> tuple:
    """`scipy.optimize.linear_sum_assignment(cost)` (minimise) through `dskd_lsap_f64`.

    cost: 2-D array-like / CPU tensor; returns (row_ind, col_ind) int64 CPU tensors sorted by row."""
    c = torch.as_tensor(cost, dtype=torch.float64, device='cpu').contiguous()
    if c.dim() != 2:
        raise ValueError(f'expected a matrix, got {c.dim()}-D input')
    rows, cols = c.shape
    k = min(rows, cols)
    ri = torch.empty(k, dtype=torch.int64)
    ci = torch.empty(k, dtype=torch.int64)
    rc = L.load().dskd_lsap_f64(C.c_void_p(c.data_ptr()), rows, cols, C.c_void_p(ri.data_ptr()), C.c_void_p(ci.data_ptr()))
    if rc == L.EINFEASIBLE:
        raise ValueError('cost matrix is infeasible')        # SciPy's message for the same inputs
    L.check(rc, 'dskd_lsap_f64')
    return ri, ci


@ASSIGNERS.register_module()
class GFLHungarianAssigner:
    """Drop-in for `GFLHungarianAssigner` (gfl_hungarian_assigner.py:17-57 ctor convention: sub-dicts
    with `weight`), plus a batched entry point for all decoder layers.

    Only the cost types the IL configs use are implemented: QualityFocalLossCost + BBoxL1Cost('xywh') +
    IoUCost('giou') (chaosuan_..._40_...py:135-139)."""

    def __init__(self, cls_cost=dict(type='QualityFocalLossCost', weight=2.0),
                 reg_cost=dict(type='BBoxL1Cost', weight=5.0, box_format='xywh'),
                 iou_cost=dict(type='IoUCost', iou_mode='giou', weight=2.0),
                 num_classes=80, reg_max=16, num_threads=0, solver='device'):
        if cls_cost.get('type', 'QualityFocalLossCost') != 'QualityFocalLossCost':
            raise NotImplementedError('only QualityFocalLossCost is on the DSKD path')
        if reg_cost.get('box_format', 'xywh') != 'xywh' or iou_cost.get('iou_mode', 'giou') != 'giou':
            raise NotImplementedError("only BBoxL1Cost(box_format='xywh') and IoUCost('giou') are on the DSKD path")
        self.w_cls = float(cls_cost.get('weight', 1.0))
        self.w_reg = float(reg_cost.get('weight', 1.0))
        self.w_iou = float(iou_cost.get('weight', 1.0))
        self.num_classes = num_classes
        self.reg_max = reg_max
        self.num_threads = num_threads
        if solver not in ('device', 'host'):
            raise ValueError("solver must be 'device' (batched LSAP kernel, no host sync) or 'host' (C++ threads)")
        self.solver = solver          # `assign_batch` only; the per-image `assign` always returns host-checked results

    # ------------------------------------------------------------------ batched path
    @L.guarded
    def cost_matrices(self, cls_scores, bbox_preds, gt_bboxes_list, gt_labels_list, img_shapes, decoded=False):
        """cls_scores [P,Q,classes] (or [layers,N,Q,classes]); bbox_preds [...,Q,2+4*(reg_max+1)] sigmoid
        outputs (or decoded cxcywh [...,Q,4] with decoded=True).  Returns (cost [P,Q,max_gt] on the device,
        cols list[P], meta) -- entries beyond cols[p] are unspecified."""
        lib = L.load()
        N = len(gt_bboxes_list)
        cls = L.f32c(cls_scores.detach())
        box = L.f32c(bbox_preds.detach())
        Q = cls.shape[-2]
        cls = cls.reshape(-1, Q, cls.shape[-1])
        box = box.reshape(-1, Q, box.shape[-1])
        P = cls.shape[0]
        if P % max(N, 1) != 0:
            raise L.DskdError(f'{P} problems cannot be split over {N} images')
        dev = cls.device
        L.require_device(cls)
        glens = [int(b.shape[0]) for b in gt_bboxes_list]
        start = [0]
        for g in glens:
            start.append(start[-1] + g)
        if isinstance(img_shapes, torch.Tensor):
            img_shapes = img_shapes.tolist()
        meta = start + [int(v) for hw in img_shapes for v in hw[:2]]
        meta_t = torch.tensor(meta, dtype=torch.int32).to(dev, non_blocking=True)
        gt_start, img_hw = meta_t[:N + 1], meta_t[N + 1:]
        max_gt = max(glens) if glens else 0
        gts = L.f32c(torch.cat([b.reshape(-1, 4) for b in gt_bboxes_list], 0)) if N else torch.zeros(0, 4, device=dev)
        labs = torch.cat([l.reshape(-1) for l in gt_labels_list], 0).to(dev, torch.int64).contiguous() if N else \
            torch.zeros(0, dtype=torch.int64, device=dev)
        cost = torch.empty(P, Q, max(max_gt, 1), dtype=torch.float32, device=dev)
        reg_max = 0 if decoded else self.reg_max
        if not decoded and box.shape[-1] != 2 + 4 * (self.reg_max + 1):
            raise L.DskdError(f'bbox_preds last dim {box.shape[-1]} != 2 + 4*(reg_max+1)')
        L.check(lib.dskd_cost_matrix(L.ptr(cls), L.ptr(box), P, N, Q, cls.shape[-1], reg_max, L.ptr(gts), L.ptr(labs),
                                     L.ptr(gt_start), L.ptr(img_hw), max_gt, self.w_cls, self.w_reg, self.w_iou,
                                     L.ptr(cost), L.stream_of(cls)), 'dskd_cost_matrix')
        cols = [glens[p % N] for p in range(P)]
        return cost, cols, dict(P=P, N=N, Q=Q, max_gt=max_gt, gts=gts, labs=labs, gt_start=gt_start, img_hw=img_hw)

    def solve(self, cost: torch.Tensor, cols: Sequence[int]) -> torch.Tensor:
        """One D2H copy of every cost matrix, the C++ solver on host threads -> assigned_gt [P,Q] int64 (CPU,
        pinned): 1-based GT index or 0 (gfl_hungarian_assigner.py:153-158)."""
        P, Q, ld = cost.shape
        host = torch.empty(cost.shape, dtype=torch.float32, pin_memory=True)
        host.copy_(cost, non_blocking=True)
        cols_t = torch.tensor(list(cols), dtype=torch.int32)
        out = torch.empty(P, Q, dtype=torch.int64, pin_memory=True)
        torch.cuda.current_stream(cost.device).synchronize()       # the single host sync of the assignment
        L.check(L.load().dskd_lsap_batch_f32(C.c_void_p(host.data_ptr()), P, Q, ld, C.c_void_p(cols_t.data_ptr()),
                                             C.c_void_p(out.data_ptr()), self.num_threads), 'dskd_lsap_batch_f32')
        return out

    @L.guarded
    def solve_device(self, cost: torch.Tensor, m: dict, check: bool = False) -> torch.Tensor:
        """The same matching on the device (`dskd_lsap_batch_device`): no copy of the cost matrices, no host sync.
        `check=True` reads the per-problem status back (one sync) and raises like SciPy on an infeasible matrix."""
        P, Q, ld = cost.shape
        dev = cost.device
        assigned = torch.empty(P, Q, dtype=torch.int64, device=dev)
        status = torch.empty(P, dtype=torch.int32, device=dev)
        L.check(L.load().dskd_lsap_batch_device(L.ptr(cost), P, m['N'], Q, ld, L.ptr(m['gt_start']), m['max_gt'],
                                                L.ptr(assigned), L.ptr(status), L.stream_of(cost)), 'dskd_lsap_batch_device')
        self.last_status = status
        if check and bool((status != 0).any()):
            raise ValueError('cost matrix is infeasible')          # scipy.optimize.linear_sum_assignment's message
        return assigned

    @L.guarded
    def assign_batch(self, cls_scores, bbox_preds, gt_bboxes_list, gt_labels_list, img_shapes,
                     prev_labels: Optional[Sequence[int]] = None) -> Dict[str, torch.Tensor]:
        """All decoder layers at once (`loss_single_split` x 6 -> `get_targets`, head_il.py:504-512,1437-1455).

        Returns device tensors shaped [layers*N*Q] / [...,4] in (layer, image, query) order:
        assigned_gt_inds, labels (bg = num_classes), label_weights, bbox_targets, bbox_weights,
        teacher_only_weights."""
        lib = L.load()
        cost, cols, m = self.cost_matrices(cls_scores, bbox_preds, gt_bboxes_list, gt_labels_list, img_shapes)
        dev = cost.device
        P, N, Q = m['P'], m['N'], m['Q']
        if m['max_gt'] == 0:
            assigned = torch.zeros(P, Q, dtype=torch.int64, device=dev)
        elif self.solver == 'device' and max(Q, m['max_gt']) <= 1024:
            assigned = self.solve_device(cost, m)
        else:
            assigned = self.solve(cost, cols).to(dev, non_blocking=True)
        labels = torch.empty(P * Q, dtype=torch.int64, device=dev)
        bt = torch.empty(P * Q, 4, dtype=torch.float32, device=dev)
        bw = torch.empty(P * Q, 4, dtype=torch.float32, device=dev)
        only = torch.empty(P * Q, dtype=torch.float32, device=dev)
        from .losses import _prev_masks
        prev_mask = _prev_masks.get(prev_labels or [], self.num_classes, dev)
        L.check(lib.dskd_assign_targets(L.ptr(assigned), P, N, Q, self.num_classes, L.ptr(m['gts']), L.ptr(m['labs']),
                                        L.ptr(m['gt_start']), L.ptr(m['img_hw']), L.ptr(prev_mask), L.ptr(labels),
                                        L.ptr(bt), L.ptr(bw), L.ptr(only), L.stream_of(cost)), 'dskd_assign_targets')
        return dict(assigned_gt_inds=assigned.reshape(-1), labels=labels, label_weights=torch.ones_like(only),
                    bbox_targets=bt, bbox_weights=bw, teacher_only_weights=only)

    # ------------------------------------------------------------------ the reference's per-image signature
    @L.guarded
    def assign(self, bbox_pred, cls_pred, gt_bboxes, gt_labels, bbox_lrtb=None, img_meta=None,
               gt_bboxes_ignore=None, eps=1e-7):
        """gfl_hungarian_assigner.py:59-160: bbox_pred [Q,4] normalised cxcywh, cls_pred [Q,classes] logits,
        gt_bboxes [G,4] px xyxy, gt_labels [G]; `bbox_lrtb` is accepted and unused, like upstream."""
        assert gt_bboxes_ignore is None, 'Only case when gt_bboxes_ignore is None is supported.'
        num_gts, num_bboxes = gt_bboxes.size(0), bbox_pred.size(0)
        dev = bbox_pred.device
        gt_inds = torch.full((num_bboxes,), -1, dtype=torch.long, device=dev)
        labels = torch.full((num_bboxes,), -1, dtype=torch.long, device=dev)
        if num_gts == 0 or num_bboxes == 0:
            if num_gts == 0:
                gt_inds[:] = 0
            return AssignResult(num_gts, gt_inds, None, labels=labels)
        img_h, img_w = img_meta['img_shape'][:2]
        cost, cols, _ = self.cost_matrices(cls_pred[None], bbox_pred[None], [gt_bboxes], [gt_labels],
                                           [(img_h, img_w)], decoded=True)
        assigned = self.solve(cost, cols)[0].to(dev)
        pos = assigned > 0
        labels[pos] = gt_labels.to(dev)[assigned[pos] - 1]
        return AssignResult(num_gts, assigned, None, labels=labels)
