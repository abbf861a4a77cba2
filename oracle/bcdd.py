"""Oracle: Between-Class Distance Distillation (BCDD).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Restates
  mmdet/models/dense_heads/gfl_deformable_detr_head_il.py:525-552    prototype sums / counts
  mmdet/models/dense_heads/gfl_deformable_detr_head_il.py:1197-1222  correlation_mat
"""
import torch


def prototypes(hs_student, student_labels, hs_teacher, teacher_keepid, teacher_labels,
               prev_labels, num_classes=80):
    """head_il.py:525-551.  hs_*: [N*Q, C] (last decoder layer, flattened);
    student_labels: [N*Q] assigned labels (bg = num_classes); teacher_keepid / teacher_labels: [sum K].
    Returns (corr_teacher, corr_student), each [num_classes, C+1] (last column = count)."""
    C = hs_student.shape[-1]
    corr_s = hs_student.new_zeros((num_classes, C + 1))
    sel = student_labels.new_zeros(student_labels.shape)
    for t in prev_labels:
        sel[student_labels == t, ...] = 1
    for idx in torch.nonzero(sel):
        lab = student_labels[idx][0]
        corr_s[lab][:-1] += hs_student[idx][0]
        corr_s[lab][-1] += 1
    corr_t = hs_student.new_zeros((num_classes, C + 1))
    for i in range(len(teacher_labels)):
        corr_t[teacher_labels[i]][:-1] += hs_teacher[teacher_keepid[i]]
        corr_t[teacher_labels[i]][-1] += 1
    return corr_t, corr_s


def distance_matrices(corr_teacher, corr_student, prev_length):
    """head_il.py:1197-1216.  NB the student rows are normalised on the TEACHER's non-zero
    index set (`idx_s = nonzero(num_t)`, :1205); pairwise distances by direct difference."""
    c_t = corr_teacher[:prev_length, :-1]
    num_t = corr_teacher[:prev_length, -1]
    idx_t = torch.nonzero(num_t).squeeze(1)
    c_t[idx_t] = c_t[idx_t] / num_t[idx_t].unsqueeze(1).repeat(1, corr_teacher.shape[1] - 1)
    c_s = corr_student[:prev_length, :-1]
    num_s = corr_student[:prev_length, -1]
    idx_s = torch.nonzero(num_t).squeeze(1)
    c_s[idx_s] = c_s[idx_s] / num_s[idx_s].unsqueeze(1).repeat(1, corr_student.shape[1] - 1)
    L = c_t.shape[0]
    mat_t = c_t.new_zeros((L, L))
    mat_s = c_t.new_zeros((L, L))
    for i in range(L):
        for j in range(L):
            mat_t[i][j] = torch.dist(c_t[i], c_t[j], p=2)
            mat_s[i][j] = torch.dist(c_s[i], c_s[j], p=2)
    return mat_t, mat_s


def correlation_loss(corr_teacher, corr_student, prev_length, criterion):
    """head_il.py:1197-1222: criterion(pred=D_teacher, target=D_student) / L."""
    mat_t, mat_s = distance_matrices(corr_teacher, corr_student, prev_length)
    return criterion(mat_t, mat_s, weight=None, avg_factor=None) / mat_t.shape[0]


def bcdd_loss(hs_student, student_labels, hs_teacher, teacher_keepid, teacher_labels,
              prev_labels, criterion, num_classes=80):
    """head_il.py:525-555 end to end."""
    corr_t, corr_s = prototypes(hs_student, student_labels, hs_teacher, teacher_keepid,
                                teacher_labels, prev_labels, num_classes)
    return correlation_loss(corr_t, corr_s, len(prev_labels), criterion)
