"""ORACLE (test infrastructure, not product code): multi-scale deformable attention as its published definition --
mmcv `multi_scale_deformable_attn_pytorch` (mmcv/ops/multi_scale_deform_attn.py, the pure-PyTorch twin of the CUDA op
the reference imports at mmdet/models/utils/transformer.py:23; mmcv-full pinned by requirements/mminstall.txt:1, not
vendored under /root/reference): `F.grid_sample(bilinear, zeros, align_corners=False)` per level, weighted by the
attention weights, summed over levels and points.  Parity unpinned by the reference's own tests (it has none for this
op); the CUDA kernels of dskd_b200/csrc/msda.cu are held to this restatement in fp32 and float64."""
import torch.nn.functional as F


def msda_torch(value, spatial_shapes, sampling_locations, attention_weights):
    """value [N,S,M,D]; spatial_shapes list[(H,W)]; sampling_locations [N,Lq,M,L,P,2] in [0,1]; attention_weights
    [N,Lq,M,L,P] -> [N,Lq,M*D]."""
    N, S, M, D = value.shape
    Lq, P = sampling_locations.shape[1], sampling_locations.shape[4]
    grids = 2 * sampling_locations - 1
    out = value.new_zeros(N * M, D, Lq)
    start = 0
    for l, (hh, ww) in enumerate(spatial_shapes):
        v = value[:, start:start + hh * ww].permute(0, 2, 3, 1).reshape(N * M, D, hh, ww)
        g = grids[:, :, :, l].permute(0, 2, 1, 3, 4).reshape(N * M, Lq, P, 2)
        sampled = F.grid_sample(v, g, mode='bilinear', padding_mode='zeros', align_corners=False)   # [N*M,D,Lq,P]
        a = attention_weights[:, :, :, l].permute(0, 2, 1, 3).reshape(N * M, 1, Lq, P)
        out = out + (sampled * a).sum(-1)
        start += hh * ww
    return out.view(N, M * D, Lq).transpose(1, 2)
