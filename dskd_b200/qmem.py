"""Query x memory soft-ownership cell weights on tcgen05 tensor cores.

Not part of the reference (it has no query x memory contraction); BASELINE.json's north_star asks for
one.  The definition lives in oracle/qmem.py (test infrastructure) and include/dskd_b200.h (`dskd_qmem_cell_weights`):

    z[s,j] = <memory[s,i,:], hs_T[keepid[j],:]> / (sqrt(C) * temp)        j over the K_i teacher detections of image i
    w[i,s] = sqrt( sum_j c_j e^z[s,j] / (1 + sum_j e^z[s,j]) )

The [S, K_i] score matrix stays in tensor memory; only w [N,S] is written.  `DSGFeatureDistillLoss(mask_mode='qmem',
feature_source='memory')` feeds it to the streaming masked-MSE kernel as its cell mask.
"""
from typing import Optional, Sequence

import torch

from . import _lib as L


@L.guarded
def qmem_cell_weights(teacher_memory: torch.Tensor, hs_teacher: torch.Tensor, teacher_keepid: torch.Tensor,
                      teacher_scores: Optional[torch.Tensor], box_start: torch.Tensor, max_per_image: int,
                      temperature: float = 0.5) -> torch.Tensor:
    """teacher_memory [S,N,C] fp32; hs_teacher [..., C]; teacher_keepid int64 [P] rows of hs_teacher;
    teacher_scores fp32 [P] or None; box_start int32 [N+1] (device).  Returns w [N,S] fp32."""
    lib = L.load()
    L.require_device(teacher_memory)
    mem = L.f32c(teacher_memory.detach())
    if mem.dim() != 3:
        raise L.DskdError(f'teacher_memory must be [S,N,C], got {tuple(mem.shape)}')
    S, N, C = mem.shape
    hs = L.f32c(hs_teacher.detach()).reshape(-1, C)
    dev = mem.device
    keep = teacher_keepid.to(dev, torch.int64).contiguous()
    scores = None if teacher_scores is None else L.f32c(teacher_scores.detach().to(dev))
    P = int(keep.numel())
    if scores is not None and scores.numel() != P:
        raise L.DskdError(f'{scores.numel()} teacher scores for {P} keep-ids')
    if box_start.dtype != torch.int32 or box_start.numel() != N + 1 or not box_start.is_cuda:
        raise L.DskdError('box_start must be a device int32 tensor of N+1 prefix offsets')
    out = torch.empty(N, S, dtype=torch.float32, device=dev)
    a = L.QmemArgs()
    a.N, a.C, a.S = N, C, S
    a.d_memory, a.d_hs_teacher, a.num_query_rows = mem.data_ptr(), hs.data_ptr(), hs.shape[0]
    a.d_keepid = keep.data_ptr() if P else None
    a.d_scores = scores.data_ptr() if scores is not None and P else None
    a.d_box_start = box_start.data_ptr()
    a.num_pairs, a.max_per_image, a.temperature = P, int(max_per_image), float(temperature)
    a.d_cell_weight = out.data_ptr()
    nbytes = lib.dskd_qmem_workspace_bytes(N, S, C, int(max_per_image))
    ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
    a.d_workspace, a.workspace_bytes = ws.data_ptr(), nbytes
    L.check(lib.dskd_qmem_cell_weights(a, L.stream_of(mem)), 'dskd_qmem_cell_weights')
    return out


class CellWeightMseFn(torch.autograd.Function):
    """Masked MSE with a given (constant) cell mask on encoder memory [S,N,C], fused fwd+bwd
    (`dskd_dsgfd_mse_fwd_bwd`, cell-mask mode): loss = sum_l scale_l sum (w (T - S))^2."""

    @staticmethod
    def forward(ctx, shapes, scales, cell_weight, s_mem, t_mem):
        lib = L.load()
        dev = s_mem.device
        S, N, C = s_mem.shape
        levels, cells = L.levels_struct(shapes)
        a = L.DsgfdMseArgs()
        a.layout, a.num_levels, a.N, a.C = L.LAYOUT_SNC, len(shapes), N, C
        a.levels, a.cells_per_image = levels, cells
        grad = torch.empty_like(s_mem) if ctx.needs_input_grad[3] else None
        a.d_student[0], a.d_teacher[0] = s_mem.data_ptr(), t_mem.data_ptr()
        a.d_grad_student[0] = grad.data_ptr() if grad is not None else None
        for l, sc in enumerate(scales):
            a.scale[l] = sc
        acc = torch.zeros(1, dtype=torch.float64, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        a.d_cell_weight, a.d_loss = cell_weight.data_ptr(), acc.data_ptr()
        st = L.stream_of(s_mem)
        L.check(lib.dskd_dsgfd_mse_fwd_bwd(a, st), 'dskd_dsgfd_mse_fwd_bwd')
        L.check(lib.dskd_f64_to_f32(L.ptr(acc), L.ptr(loss), 1, 1.0, st), 'dskd_f64_to_f32')
        ctx.staged = grad
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        grad = ctx.staged
        ctx.staged = None
        if grad is not None:
            g = grad_out.detach().float().reshape(1).contiguous()
            L.check(L.load().dskd_scale_inplace(L.ptr(grad), grad.numel(), L.ptr(g), L.stream_of(grad)),
                    'dskd_scale_inplace')
        return None, None, None, grad, None
