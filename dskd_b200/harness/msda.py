"""Multi-scale deformable attention on the GPU: the sampling core of every encoder / decoder layer of the harness
detector.  The reference runs mmcv's CUDA op here
(`MultiScaleDeformableAttention`, mmdet/models/utils/transformer.py:23); this is its B200 replacement behind the
same functional contract as mmcv's `MultiScaleDeformableAttnFunction.apply(value, spatial_shapes, ..., sampling_locations,
attention_weights)`, forward + backward in one CUDA kernel each (dskd_b200/csrc/msda.cu).  The published `grid_sample`
definition of the op (mmcv `multi_scale_deformable_attn_pytorch`) is restated in oracle/msda.py and is what
tests/test_gpu_msda.py holds the kernels to.  CUDA tensors only: there is no CPU path."""
import ctypes as C

import torch

from .. import _lib as L


def _levels(shapes):
    arr = (L.Level * len(shapes))()
    off = 0
    for l, (h, w) in enumerate(shapes):
        arr[l].H, arr[l].W, arr[l].cell_offset = int(h), int(w), off
        off += int(h) * int(w)
    return arr, off


class _MsdaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, value, loc, attn, shapes):
        lib = L.load()
        L.require_device(value)
        value, loc, attn = L.f32c(value), L.f32c(loc), L.f32c(attn)
        N, S, M, D = value.shape
        Lq, nl, P = loc.shape[1], loc.shape[3], loc.shape[4]
        levels, cells = _levels(shapes)
        if cells != S or nl != len(shapes):
            raise L.DskdError(f'value has {S} tokens / loc {nl} levels, the shapes give {cells} / {len(shapes)}')
        out = torch.empty(N, Lq, M, D, dtype=torch.float32, device=value.device)
        L.check(lib.dskd_msda_forward(L.ptr(value), C.cast(levels, C.c_void_p), len(shapes), L.ptr(loc), L.ptr(attn),
                                      N, S, M, D, Lq, P, L.ptr(out), L.stream_of(value)), 'dskd_msda_forward')
        ctx.save_for_backward(value, loc, attn)
        ctx.shapes = tuple((int(h), int(w)) for h, w in shapes)
        return out.view(N, Lq, M * D)

    @staticmethod
    def backward(ctx, grad_out):
        lib = L.load()
        value, loc, attn = ctx.saved_tensors
        N, S, M, D = value.shape
        Lq, P = loc.shape[1], loc.shape[4]
        levels, _ = _levels(ctx.shapes)
        go = L.f32c(grad_out)
        gv, gl, ga = torch.empty_like(value), torch.empty_like(loc), torch.empty_like(attn)
        L.check(lib.dskd_msda_backward(L.ptr(value), C.cast(levels, C.c_void_p), len(ctx.shapes), L.ptr(loc), L.ptr(attn),
                                       L.ptr(go), N, S, M, D, Lq, P, L.ptr(gv), L.ptr(gl), L.ptr(ga),
                                       L.stream_of(value)), 'dskd_msda_backward')
        return gv, gl, ga, None


@L.guarded
def ms_deform_attn(value, spatial_shapes, sampling_locations, attention_weights):
    """value [N,S,M,D]; spatial_shapes list[(H,W)]; sampling_locations [N,Lq,M,L,P,2] in [0,1] (x, y);
    attention_weights [N,Lq,M,L,P] -> [N,Lq,M*D].  Raises on CPU tensors."""
    return _MsdaFn.apply(value, sampling_locations, attention_weights, tuple(spatial_shapes))
