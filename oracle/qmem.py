"""Oracle: query x memory soft-ownership mask -- TEST INFRASTRUCTURE, see oracle/__init__.py.

PARITY UNPINNED BY THE REFERENCE: smilekitty7/DSKD has no query x memory contraction (no matmul / bmm /
einsum anywhere in mmdet/models/dense_heads/gfl_deformable_detr_head_il.py).  BASELINE.json's
north_star asks for one, so this file DEFINES the extension in plain PyTorch and the CUDA kernel
(dskd_b200/csrc/qmem.cu) is held to it.  What it borrows from the reference:
  * the matched teacher detections and their decoder embeddings `hs_T[pred_keepid]` (head_il.py:541-551,705),
  * the confidences `pred_scores` (deformable_detr_il.py:139),
  * the head's unused temperature argument `temp=0.5` (head_il.py:90,124) and the authors' commented
    `softmax(.../temp)` spatial-attention idiom (gfl_deformable_detr_head_il_fg_bk.py:582-591),
  * the per-(level, image) `criterion(teacher*mask, student*mask)` reduction and the sqrt of cell masks
    (head_il.py:914-923,1119-1127).
"""
import math

import torch

from .dsgfd import cell_mask_loss, memory_levels


def to_tf32(x, mode):
    """fp32 -> tf32 (10 explicit mantissa bits) by truncation ('trunc') or round-to-nearest-even ('rne')."""
    if mode is None:
        return x
    bits = x.contiguous().view(torch.int32)
    if mode == 'rne':
        bits = bits + 0x0FFF + ((bits >> 13) & 1)
    return (bits & ~0x1FFF).view(torch.float32)


def qmem_cell_weights(teacher_memory, hs_teacher, teacher_keepid, teacher_scores, box_counts, temperature=0.5,
                      tf32=None, dtype=torch.float32):
    """teacher_memory [S,N,C]; hs_teacher [rows,C]; keepid/scores [sum K] concatenated per image.

    z[s,j] = <m_s, q_j> / (sqrt(C) temp);  w[i,s] = sqrt(sum_j c_j e^z / (1 + sum_j e^z)).  Returns [N,S]."""
    S, N, C = teacher_memory.shape
    mem = to_tf32(teacher_memory, tf32).to(dtype)
    hs = to_tf32(hs_teacher.reshape(-1, C), tf32).to(dtype)
    out = torch.zeros(N, S, dtype=dtype)
    start = 0
    for i in range(N):
        k = int(box_counts[i])
        if k:
            q = hs[teacher_keepid[start:start + k]]
            c = (teacher_scores[start:start + k] if teacher_scores is not None else torch.ones(k)).to(dtype)
            z = mem[:, i, :] @ q.t() / (math.sqrt(C) * temperature)          # [S,K]
            m = z.max(dim=1).values.clamp(min=0)                               # the null logit is 0
            e = torch.exp(z - m[:, None])
            out[i] = torch.sqrt((e * c[None, :]).sum(1) / (torch.exp(-m) + e.sum(1)))
        start += k
    return out


def qmem(student_memory, teacher_memory, spatial_shapes, hs_teacher, teacher_keepid, teacher_scores, box_counts,
         criterion, temperature=0.5, weights=None):
    """DSG-FD with the soft-ownership cell mask on encoder memory [S,N,C] (mask detached)."""
    w = weights if weights is not None else qmem_cell_weights(teacher_memory, hs_teacher, teacher_keepid,
                                                              teacher_scores, box_counts, temperature)
    w = w.detach().float()
    s_lv = memory_levels(student_memory, spatial_shapes)
    t_lv = memory_levels(teacher_memory, spatial_shapes)
    masks, off = [], 0
    for h, wd in spatial_shapes:
        h, wd = int(h), int(wd)
        masks.append([w[i, off:off + h * wd].reshape(h, wd) for i in range(w.shape[0])])
        off += h * wd
    return cell_mask_loss(s_lv, t_lv, masks, criterion)
