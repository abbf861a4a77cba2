"""Oracle: Dynamically Semantic-Guided Feature Distillation (DSG-FD) and sibling masks.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Restates, loop order included
(later boxes overwrite earlier ones; the pair index runs over all images of a
level and restarts at every level):
  mmdet/models/dense_heads/gfl_deformable_detr_head_il.py:664-719   decode_v1
  .../gfl_deformable_detr_head_il.py:721-772                        decode_v2
  .../gfl_deformable_detr_head_il.py:860-925                        sg_out
  .../gfl_deformable_detr_head_il.py:1082-1129                      fg_only
  .../gfl_deformable_detr_head_il_fg_bk.py:534-578,611-625          fg_bk (area mask, MSE/C)

`criterion` is one of the oracle loss modules (oracle.losses); the reference
calls it as criterion(pred=TEACHER*mask, target=STUDENT*mask) -- the variable
names in the reference are swapped (head_il.py:709-715) and that order is kept.
"""
import torch


def box_cells(bboxes, img_hw, H, W, x_extent=None, y_extent=None):
    """head_il.py:688-696: fp32 `x / img_w * W`, floor / ceil, `.int()`.

    `x_extent/y_extent` override W/H as the scale (the fg_bk variant scales x by
    spatial_shapes[sp][0] (=H) and y by spatial_shapes[sp][1] (=W): _fg_bk.py:550-553)."""
    img_h, img_w = img_hw
    sx = W if x_extent is None else x_extent
    sy = H if y_extent is None else y_extent
    nb = torch.ones_like(bboxes)
    nb[:, 0] = bboxes[:, 0] / img_w * sx
    nb[:, 2] = bboxes[:, 2] / img_w * sx
    nb[:, 1] = bboxes[:, 1] / img_h * sy
    nb[:, 3] = bboxes[:, 3] / img_h * sy
    wmin = torch.floor(nb[:, 0]).int()
    wmax = torch.ceil(nb[:, 2]).int()
    hmin = torch.floor(nb[:, 1]).int()
    hmax = torch.ceil(nb[:, 3]).int()
    return wmin, wmax, hmin, hmax


def _decode_loss(student_feats, teacher_feats, hs_student, hs_teacher, id_soft, id_pred,
                 teacher_bboxes, img_shapes, criterion, version):
    hs_t = hs_teacher.reshape(-1, hs_teacher.shape[-1])
    hs_s = hs_student.reshape(-1, hs_student.shape[-1])
    total = 0
    for f_s, f_t in zip(student_feats, teacher_feats):
        N, C, H, W = f_s.shape
        # The reference fills one shared [N,C,H,W] tensor in place (head_il.py:683,706); a
        # separate tensor per image holds the same numbers and keeps autograd's saved views
        # valid (the reference's own MSE backward raises for N >= 2 because of that sharing).
        mask = [torch.zeros((C, H, W), device=f_s.device, dtype=f_s.dtype) for _ in range(N)]
        idx = 0
        for i in range(N):
            wmin, wmax, hmin, hmax = box_cells(teacher_bboxes[i], img_shapes[i], H, W)
            for j in range(len(teacher_bboxes[i])):
                if version == 1:
                    m = (hs_t[id_soft[idx]] - hs_s[id_pred[idx]]).abs().softmax(dim=0)
                else:
                    m = hs_t[id_soft[idx]].softmax(dim=0)
                hh, ww = hmax[j] - hmin[j], wmax[j] - wmin[j]
                mask[i][:, hmin[j]:hmax[j], wmin[j]:wmax[j]] = \
                    m.unsqueeze(1).unsqueeze(2).repeat(1, hh, ww)
                idx += 1
            pred = f_t[i] * mask[i]        # "fg_fea_s" in the reference = teacher * mask
            target = f_s[i] * mask[i]      # "fg_fea_t" = student * mask
            total = total + criterion(pred, target, weight=None, avg_factor=None)
    return total / len(img_shapes)


def decode_v1(student_feats, teacher_feats, hs_student, hs_teacher, id_soft, id_pred,
              teacher_bboxes, img_shapes, criterion):
    """head_il.py:664-719.  student_feats/teacher_feats: 4 x [N,C,H,W]; hs_*: [N,Q,C] (last
    decoder layer); id_soft = teacher keep-ids (score order), id_pred = ascending ids of
    student queries whose assigned label is a previous-task label."""
    return _decode_loss(student_feats, teacher_feats, hs_student, hs_teacher, id_soft, id_pred,
                        teacher_bboxes, img_shapes, criterion, 1)


def decode_v2(student_feats, teacher_feats, hs_teacher, id_soft, teacher_bboxes, img_shapes, criterion):
    """head_il.py:721-772 -- mask row = softmax(hs_T[id_soft]) (teacher only)."""
    return _decode_loss(student_feats, teacher_feats, hs_teacher, hs_teacher, id_soft, None,
                        teacher_bboxes, img_shapes, criterion, 2)


def memory_levels(memory, spatial_shapes):
    """head_il.py:866-880: memory [S,N,C] -> per level [N,C,H,W] views (permute(1,2,0), slice, reshape)."""
    m = memory.permute(1, 2, 0)
    N, C, _ = m.shape
    out, off = [], 0
    for h, w in spatial_shapes:
        h, w = int(h), int(w)
        out.append(m[:, :, off:off + h * w].reshape(N, C, h, w))
        off += h * w
    return out


def cell_mask_sg_out(teacher_bboxes_i, gt_bboxes_i, img_hw, H, W):
    """head_il.py:883-914: 1 inside teacher boxes (inclusive +1 ends), 0 inside GT boxes, sqrt."""
    m = torch.zeros((H, W))
    wmin, wmax, hmin, hmax = box_cells(teacher_bboxes_i, img_hw, H, W)
    for j in range(len(teacher_bboxes_i)):
        m[hmin[j]:hmax[j] + 1, wmin[j]:wmax[j] + 1] = 1
    wmin, wmax, hmin, hmax = box_cells(gt_bboxes_i, img_hw, H, W)
    for j in range(len(gt_bboxes_i)):
        m[hmin[j]:hmax[j] + 1, wmin[j]:wmax[j] + 1] = 0
    return torch.sqrt(m)


def cell_mask_area(teacher_bboxes_i, img_hw, H, W, fg_bk_scale_bug=False):
    """head_il.py:1107-1122 (fg_only) / _fg_bk.py:548-566: max over boxes of 1/((dh+1)(dw+1)), sqrt."""
    m = torch.zeros((H, W))
    if fg_bk_scale_bug:
        wmin, wmax, hmin, hmax = box_cells(teacher_bboxes_i, img_hw, H, W, x_extent=H, y_extent=W)
    else:
        wmin, wmax, hmin, hmax = box_cells(teacher_bboxes_i, img_hw, H, W)
    area = 1.0 / (hmax.view(1, -1) + 1 - hmin.view(1, -1)) / (wmax.view(1, -1) + 1 - wmin.view(1, -1))
    for j in range(len(teacher_bboxes_i)):
        m[hmin[j]:hmax[j] + 1, wmin[j]:wmax[j] + 1] = torch.maximum(
            m[hmin[j]:hmax[j] + 1, wmin[j]:wmax[j] + 1], area[0][j])
    return torch.sqrt(m)


def cell_mask_loss(student_levels, teacher_levels, cell_masks, criterion):
    """Shared tail of sg_out / fg_only (head_il.py:916-923,1121-1127): per (level, image)
    criterion(teacher*mask, student*mask), summed, / N.  cell_masks[level][i] is [H,W]."""
    total = 0
    N = student_levels[0].shape[0]
    for lvl, (f_s, f_t) in enumerate(zip(student_levels, teacher_levels)):
        for i in range(N):
            mk = cell_masks[lvl][i].unsqueeze(0)
            total = total + criterion(f_t[i] * mk, f_s[i] * mk, weight=None, avg_factor=None)
    return total / N


def sg_out(student_memory, teacher_memory, spatial_shapes, teacher_bboxes, gt_bboxes, img_shapes, criterion):
    """head_il.py:860-925 on encoder memory [S,N,C]."""
    s_lv = memory_levels(student_memory, spatial_shapes)
    t_lv = memory_levels(teacher_memory, spatial_shapes)
    masks = [[cell_mask_sg_out(teacher_bboxes[i], gt_bboxes[i], img_shapes[i], int(h), int(w))
              for i in range(len(img_shapes))] for h, w in spatial_shapes]
    return cell_mask_loss(s_lv, t_lv, masks, criterion)


def fg_only(student_memory, teacher_memory, spatial_shapes, teacher_bboxes, img_shapes, criterion):
    """head_il.py:1082-1129 on encoder memory [S,N,C]."""
    s_lv = memory_levels(student_memory, spatial_shapes)
    t_lv = memory_levels(teacher_memory, spatial_shapes)
    masks = [[cell_mask_area(teacher_bboxes[i], img_shapes[i], int(h), int(w))
              for i in range(len(img_shapes))] for h, w in spatial_shapes]
    return cell_mask_loss(s_lv, t_lv, masks, criterion)


def fg_bk(student_memory, teacher_memory, spatial_shapes, teacher_bboxes, img_shapes, criterion):
    """_fg_bk.py:534-578,611-625: area mask over all levels concatenated [N,S]; per image
    criterion(student*sqrt(mask), teacher*sqrt(mask)) / C, summed, / N.  NB here pred = student."""
    mem_s = student_memory.permute(1, 2, 0)
    mem_t = teacher_memory.permute(1, 2, 0)
    N, C, _ = mem_s.shape
    total = 0
    for i in range(N):
        rows = [cell_mask_area(teacher_bboxes[i], img_shapes[i], int(h), int(w), fg_bk_scale_bug=True).reshape(-1)
                for h, w in spatial_shapes]
        mk = torch.cat(rows).unsqueeze(0).repeat(C, 1)          # already sqrt'ed
        total = total + criterion(mem_s[i] * mk, mem_t[i] * mk, weight=None, avg_factor=None) / C
    return total / N
