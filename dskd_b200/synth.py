"""Seeded synthetic COCO-shaped inputs for the distillation hot path.

Everything the reference head has in scope when it reaches its distillation block
(`gfl_deformable_detr_head_il.py:525-555,664-706`): 4-level 256-channel student / teacher
features (neck `[N,C,H,W]` per level, or encoder memory `[S,N,C]`), last-layer decoder
embeddings, teacher detections (boxes / labels / keep-ids) and the student's assigned labels.
Index structures are always drawn from a CPU generator so a CPU copy and a CUDA copy of the
same seed are identical; dense tensors are drawn on `device`.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

COCO_IMG_HW = (800, 1333)
# ResNet strides 8/16/32 + one stride-2 conv on an unpadded 800x1333 image.
COCO_LEVELS = ((100, 167), (50, 84), (25, 42), (13, 21))


def scaled_levels(tokens: int) -> Tuple[Tuple[int, int], ...]:
    """Level shapes with roughly `tokens` cells in total, same aspect / pyramid as COCO_LEVELS."""
    base = sum(h * w for h, w in COCO_LEVELS)
    f = (tokens / base) ** 0.5
    h0, w0 = max(8, round(100 * f)), max(8, round(167 * f))
    out = []
    for _ in range(4):
        out.append((h0, w0))
        h0, w0 = (h0 + 1) // 2, (w0 + 1) // 2
    return tuple(out)


@dataclass
class DistillInputs:
    student_feats: Tuple[torch.Tensor, ...]
    teacher_feats: Tuple[torch.Tensor, ...]
    hs_student: torch.Tensor            # [N,Q,C]
    hs_teacher: torch.Tensor            # [N,Q,C]
    assignments: Dict[str, object]
    spatial_shapes: torch.Tensor        # [4,2] int64 (CPU)
    levels: Tuple[Tuple[int, int], ...] = field(default=COCO_LEVELS)

    @property
    def num_images(self) -> int:
        return self.hs_student.shape[0]

    def memory(self):
        """Encoder-memory view of the same numbers: ([S,N,C] student, [S,N,C] teacher)."""
        def cat(feats):
            return torch.cat([f.flatten(2) for f in feats], dim=2).permute(2, 0, 1).contiguous()
        return cat(self.student_feats), cat(self.teacher_feats)

    def to(self, device):
        dev = torch.device(device)
        a = dict(self.assignments)
        for k in ('student_labels', 'teacher_keepid', 'teacher_labels'):      # img_shapes stays on the host,
            a[k] = a[k].to(dev)                                               # like mmdet's img_metas
        for k in ('teacher_bboxes', 'gt_bboxes'):
            a[k] = [b.to(dev) for b in a[k]]
        return DistillInputs(tuple(f.to(dev) for f in self.student_feats),
                             tuple(f.to(dev) for f in self.teacher_feats),
                             self.hs_student.to(dev), self.hs_teacher.to(dev), a,
                             self.spatial_shapes, self.levels)

    def clone_student(self, requires_grad=True):
        """Fresh leaf copies of the differentiable inputs (student feats, student embeddings)."""
        feats = tuple(f.detach().clone().requires_grad_(requires_grad) for f in self.student_feats)
        hs = self.hs_student.detach().clone().requires_grad_(requires_grad)
        return feats, hs


def _boxes(g, k, img_hw):
    h, w = img_hw
    x1 = torch.rand(k, generator=g) * 0.7 * w
    y1 = torch.rand(k, generator=g) * 0.7 * h
    bw = 8 + torch.rand(k, generator=g) * (0.3 * w - 8)
    bh = 8 + torch.rand(k, generator=g) * (0.3 * h - 8)
    b = torch.stack([x1, y1, (x1 + bw).clamp(max=w), (y1 + bh).clamp(max=h)], dim=1)
    return b.float()


def make_distill_inputs(num_images: int = 2, num_prev: int = 40, seed: int = 1234, device='cpu',
                        levels: Sequence[Tuple[int, int]] = COCO_LEVELS, num_query: int = 300,
                        channels: int = 256, num_classes: int = 80,
                        boxes_per_image: Optional[int] = None, img_hw=COCO_IMG_HW,
                        k_range=(5, 40)) -> DistillInputs:
    g = torch.Generator().manual_seed(seed)
    dev = torch.device(device)
    gd = torch.Generator(device=dev).manual_seed(seed)
    N, Q, C = num_images, num_query, channels

    def dense(*shape):
        return torch.randn(*shape, generator=gd, device=dev, dtype=torch.float32)

    student_feats = tuple(dense(N, C, h, w) for h, w in levels)
    teacher_feats = tuple(dense(N, C, h, w) for h, w in levels)
    hs_student, hs_teacher = dense(N, Q, C), dense(N, Q, C)

    labels = torch.full((N * Q,), num_classes, dtype=torch.long)
    keep, t_labels, t_boxes, gt_boxes = [], [], [], []
    for i in range(N):
        if boxes_per_image is None:
            k = int(torch.randint(k_range[0], k_range[1] + 1, (1,), generator=g))
        else:
            k = boxes_per_image
        k = min(k, Q - 10)
        t_boxes.append(_boxes(g, k, img_hw))
        tl = torch.randint(0, num_prev, (k,), generator=g)
        t_labels.append(tl)
        keep.append(torch.randperm(Q, generator=g)[:k] + i * Q)      # score order, distinct
        g_new = int(torch.randint(1, 11, (1,), generator=g))
        perm = torch.randperm(Q, generator=g)
        matched, fresh = perm[:k], perm[k:k + g_new]
        labels[i * Q + matched] = tl[torch.randperm(k, generator=g)]  # keeps n_s == n_t per class
        if num_prev < num_classes:
            labels[i * Q + fresh] = torch.randint(num_prev, num_classes, (g_new,), generator=g)
        gt_boxes.append(_boxes(g, g_new, img_hw))
    assignments = dict(
        student_labels=labels.to(dev),
        teacher_keepid=torch.cat(keep).to(dev),
        teacher_labels=torch.cat(t_labels).to(dev),
        teacher_bboxes=[b.to(dev) for b in t_boxes],
        gt_bboxes=[b.to(dev) for b in gt_boxes],
        img_shapes=torch.tensor([list(img_hw)] * N, dtype=torch.int64),      # host side (img_metas)
        prev_labels=list(range(num_prev)),
        num_classes=num_classes,
    )
    return DistillInputs(student_feats, teacher_feats, hs_student, hs_teacher, assignments,
                         torch.tensor(levels, dtype=torch.int64), tuple(tuple(l) for l in levels))


@dataclass
class AssignInputs:
    cls_logits: torch.Tensor            # [layers,N,Q,num_classes]
    box_pred: torch.Tensor              # [layers,N,Q,2+4*(reg_max+1)] sigmoid outputs
    gt_bboxes: List[torch.Tensor]       # N x [G_i,4] px xyxy (teacher pseudo boxes + new GT)
    gt_labels: List[torch.Tensor]       # N x [G_i]
    img_shapes: torch.Tensor            # [N,2] (h,w)


def make_assign_inputs(num_images: int = 2, num_prev: int = 40, seed: int = 1234, device='cpu',
                       num_layers: int = 6, num_query: int = 300, num_classes: int = 80,
                       reg_max: int = 16, img_hw=COCO_IMG_HW, k_range=(5, 40)) -> AssignInputs:
    """Hungarian bench inputs: logits ~ N(-2,1), box channels ~ U(0,1)."""
    g = torch.Generator().manual_seed(seed + 7)
    dev = torch.device(device)
    N, Q = num_images, num_query
    cls = torch.randn(num_layers, N, Q, num_classes, generator=g) - 2.0
    box = torch.rand(num_layers, N, Q, 2 + 4 * (reg_max + 1), generator=g)
    gts, labs = [], []
    for _ in range(N):
        k = int(torch.randint(k_range[0], k_range[1] + 1, (1,), generator=g))
        g_new = int(torch.randint(1, 11, (1,), generator=g))
        gts.append(_boxes(g, k + g_new, img_hw).to(dev))
        lab = torch.cat([torch.randint(0, num_prev, (k,), generator=g),
                         torch.randint(min(num_prev, num_classes - 1), num_classes, (g_new,), generator=g)])
        labs.append(lab.to(dev))
    return AssignInputs(cls.to(dev), box.to(dev), gts, labs,
                        torch.tensor([list(img_hw)] * N, dtype=torch.int64, device=dev))
