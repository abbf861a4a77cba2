"""GFL-style Deformable-DETR (R-50, 4-level ChannelMapper, 6+6 layer deformable transformer, 300 queries, 70-channel
box branch) in plain PyTorch.  Architecture anchors in the reference:
  configs/deformable_detr/chaosuan_gfl_deformable_detr_40_r50_8x4_1x_qoqo_il.py:24-152    hyper-parameters
  mmdet/models/necks/channel_mapper.py:60-110                                             neck
  mmdet/models/utils/transformer.py:893-1055                                              transformer (+ `info_all`)
  mmdet/models/dense_heads/gfl_deformable_detr_head_il.py:145-281                         branches / forward outputs
The mmcv `MultiScaleDeformableAttention` CUDA op is replaced by dskd_b200/csrc/msda.cu (`harness/msda.py`).
"""
import copy
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .msda import ms_deform_attn


_NORM_CACHE = {}


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(0, 1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


class MSDeformAttn(nn.Module):
    """Multi-scale deformable attention: projections in PyTorch, the sampling core in `harness/msda.py`."""

    def __init__(self, dim=256, heads=8, levels=4, points=4):
        super().__init__()
        self.dim, self.heads, self.levels, self.points = dim, heads, levels, points
        self.sampling_offsets = nn.Linear(dim, heads * levels * points * 2)
        self.attention_weights = nn.Linear(dim, heads * levels * points)
        self.value_proj = nn.Linear(dim, dim)
        self.output_proj = nn.Linear(dim, dim)
        nn.init.constant_(self.sampling_offsets.weight, 0.)
        thetas = torch.arange(heads, dtype=torch.float32) * (2.0 * math.pi / heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(heads, 1, 1, 2).repeat(1, levels, points, 1)
        for i in range(points):
            grid[:, :, i, :] *= i + 1
        self.sampling_offsets.bias.data = grid.view(-1)
        nn.init.constant_(self.attention_weights.weight, 0.)
        nn.init.constant_(self.attention_weights.bias, 0.)

    def forward(self, query, reference_points, value, spatial_shapes):
        """query [N,Lq,C]; reference_points [N,Lq,levels,2] in [0,1]; value [N,S,C]; spatial_shapes list[(H,W)]."""
        N, Lq, C = query.shape
        S = value.shape[1]
        h, lv, p, d = self.heads, self.levels, self.points, C // self.heads
        value = self.value_proj(value).view(N, S, h, d)
        offsets = self.sampling_offsets(query).view(N, Lq, h, lv, p, 2)
        weights = self.attention_weights(query).view(N, Lq, h, lv * p).softmax(-1).view(N, Lq, h, lv, p)
        key = (tuple(map(tuple, spatial_shapes)), query.device, query.dtype)
        norm = _NORM_CACHE.get(key)
        if norm is None:   # one small host->device copy per geometry instead of one per layer call
            norm = torch.tensor([[w, hh] for hh, w in spatial_shapes], dtype=query.dtype, device=query.device)
            _NORM_CACHE[key] = norm
        locs = reference_points[:, :, None, :, None, :] + offsets / norm[None, None, None, :, None, :]
        out = ms_deform_attn(value, spatial_shapes, locs, weights)      # CUDA op (csrc/msda.cu)
        return self.output_proj(out)


class EncoderLayer(nn.Module):
    def __init__(self, dim=256, ffn=1024, dropout=0.1):
        super().__init__()
        self.attn = MSDeformAttn(dim)
        self.n1, self.n2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn), nn.ReLU(inplace=True), nn.Dropout(dropout), nn.Linear(ffn, dim))
        self.drop = nn.Dropout(dropout)

    def forward(self, src, pos, ref, shapes):
        src = self.n1(src + self.drop(self.attn(src + pos, ref, src, shapes)))
        return self.n2(src + self.drop(self.ffn(src)))


class DecoderLayer(nn.Module):
    def __init__(self, dim=256, ffn=1024, heads=8, dropout=0.1):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(dim, heads, dropout=dropout, batch_first=True)
        self.cross = MSDeformAttn(dim)
        self.n1, self.n2, self.n3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn), nn.ReLU(inplace=True), nn.Dropout(dropout), nn.Linear(ffn, dim))
        self.drop = nn.Dropout(dropout)

    def forward(self, tgt, qpos, ref, memory, shapes):
        q = tgt + qpos
        tgt = self.n1(tgt + self.drop(self.self_attn(q, q, tgt, need_weights=False)[0]))
        tgt = self.n2(tgt + self.drop(self.cross(tgt + qpos, ref, memory, shapes)))
        return self.n3(tgt + self.drop(self.ffn(tgt)))


def sine_pos(h, w, device, feats=128, temperature=10000, offset=-0.5):
    """SinePositionalEncoding(num_feats=128, normalize=True, offset=-0.5) on an unpadded map -> [2*feats, h, w]."""
    y = (torch.arange(1, h + 1, device=device, dtype=torch.float32) + offset) / (h + 1e-6) * 2 * math.pi
    x = (torch.arange(1, w + 1, device=device, dtype=torch.float32) + offset) / (w + 1e-6) * 2 * math.pi
    dim_t = temperature ** (2 * (torch.arange(feats, device=device) // 2) / feats)
    px = x[:, None] / dim_t
    py = y[:, None] / dim_t
    px = torch.stack((px[:, 0::2].sin(), px[:, 1::2].cos()), 2).flatten(1)       # [w, feats]
    py = torch.stack((py[:, 0::2].sin(), py[:, 1::2].cos()), 2).flatten(1)       # [h, feats]
    return torch.cat((py[:, None, :].expand(h, w, feats), px[None, :, :].expand(h, w, feats)), 2).permute(2, 0, 1)


class GFLDeformableDETR(nn.Module):
    def __init__(self, num_classes=80, num_query=300, dim=256, reg_max=16, enc_layers=6, dec_layers=6, backbone='resnet50'):
        super().__init__()
        import torchvision
        net = getattr(torchvision.models, backbone)(weights=None)
        self.stem = nn.Sequential(net.conv1, net.bn1, net.relu, net.maxpool)
        self.layer1, self.layer2, self.layer3, self.layer4 = net.layer1, net.layer2, net.layer3, net.layer4
        chans = [net.layer2[-1].conv3.out_channels if hasattr(net.layer2[-1], 'conv3') else net.layer2[-1].conv2.out_channels,
                 net.layer3[-1].conv3.out_channels if hasattr(net.layer3[-1], 'conv3') else net.layer3[-1].conv2.out_channels,
                 net.layer4[-1].conv3.out_channels if hasattr(net.layer4[-1], 'conv3') else net.layer4[-1].conv2.out_channels]
        for m in list(self.stem.modules()) + list(self.layer1.modules()):      # frozen_stages=1
            for p in m.parameters(recurse=False):
                p.requires_grad_(False)
        # ChannelMapper: 1x1 conv + GN(32) per level, one extra 3x3 stride-2 conv on the last backbone map
        self.lateral = nn.ModuleList([nn.Sequential(nn.Conv2d(c, dim, 1), nn.GroupNorm(32, dim)) for c in chans])
        self.extra = nn.Sequential(nn.Conv2d(chans[-1], dim, 3, stride=2, padding=1), nn.GroupNorm(32, dim))
        self.level_embed = nn.Parameter(torch.randn(4, dim))
        self.encoder = nn.ModuleList([EncoderLayer(dim) for _ in range(enc_layers)])
        self.decoder = nn.ModuleList([DecoderLayer(dim) for _ in range(dec_layers)])
        self.query_embedding = nn.Embedding(num_query, 2 * dim)
        self.reference_points = nn.Linear(dim, 2)
        # shared branches (with_box_refine=False): head_il.py:169-173
        self.cls_branch = nn.Linear(dim, num_classes)
        self.reg_branch = nn.Sequential(nn.Linear(dim, dim), nn.ReLU(), nn.Linear(dim, dim), nn.ReLU(),
                                        nn.Linear(dim, 2 + 4 * (reg_max + 1)))
        nn.init.constant_(self.cls_branch.bias, -math.log((1 - 0.01) / 0.01))
        nn.init.constant_(self.reg_branch[-1].weight, 0.)
        nn.init.constant_(self.reg_branch[-1].bias, 0.)
        nn.init.constant_(self.reg_branch[-1].bias.data[2:], -2.0)
        self.dim, self.num_query, self.num_classes, self.reg_max = dim, num_query, num_classes, reg_max
        self.channels_last = False

    def use_channels_last(self):
        """NHWC convolutions for backbone and neck (cuDNN's tensor-core layout: no internal transposes); the neck outputs
        keep NCHW shapes with NHWC strides, the loss modules make them contiguous where they need to."""
        self.to(memory_format=torch.channels_last)
        self.channels_last = True
        return self

    def train(self, mode=True):
        super().train(mode)
        for m in self.modules():                       # norm_eval=True: BatchNorm always in eval mode
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
        return self

    def extract_feat(self, img):
        if self.channels_last:
            img = img.contiguous(memory_format=torch.channels_last)
        x = self.layer1(self.stem(img))
        c3 = self.layer2(x)
        c4 = self.layer3(c3)
        c5 = self.layer4(c4)
        feats = [lat(c) for lat, c in zip(self.lateral, (c3, c4, c5))]
        feats.append(self.extra(c5))
        return tuple(feats)

    def forward(self, img):
        """img [N,3,H,W] (no padding).  Returns the reference's `head_outs` pieces + neck features:
        dict(neck_feats (4 x [N,C,H,W]), cls [L,N,Q,80], box [L,N,Q,70] (sigmoid), memory [S,N,C], spatial_shapes, hs [L,N,Q,C])."""
        feats = self.extract_feat(img)
        N = img.shape[0]
        dev = img.device
        shapes = [tuple(f.shape[2:]) for f in feats]
        src = torch.cat([f.flatten(2).transpose(1, 2) for f in feats], 1)                         # [N,S,C]
        pos = torch.cat([(sine_pos(h, w, dev) + self.level_embed[l].view(-1, 1, 1)).flatten(1).t()
                         for l, (h, w) in enumerate(shapes)], 0)[None].expand(N, -1, -1)            # [N,S,C]
        ref_enc = torch.cat([torch.stack(torch.meshgrid((torch.arange(h, device=dev) + 0.5) / h,
                                                        (torch.arange(w, device=dev) + 0.5) / w, indexing='ij'), -1)
                             .flip(-1).reshape(-1, 2) for h, w in shapes], 0)                       # [S,2] (x,y)
        ref_enc = ref_enc[None, :, None, :].expand(N, -1, len(shapes), -1)
        memory = src
        for layer in self.encoder:
            memory = layer(memory, pos, ref_enc, shapes)
        qpos, tgt = self.query_embedding.weight.split(self.dim, 1)
        qpos, tgt = qpos[None].expand(N, -1, -1), tgt[None].expand(N, -1, -1)
        reference = self.reference_points(qpos).sigmoid()                                           # [N,Q,2]
        ref_dec = reference[:, :, None, :].expand(-1, -1, len(shapes), -1)
        hs, cls, box = [], [], []
        ref_logit = inverse_sigmoid(reference)
        for layer in self.decoder:
            tgt = layer(tgt, qpos, ref_dec, memory, shapes)
            hs.append(tgt)
            cls.append(self.cls_branch(tgt))
            tmp = self.reg_branch(tgt)
            tmp = torch.cat((tmp[..., :2] + ref_logit, tmp[..., 2:]), -1)                           # head_il.py:259-266
            box.append(tmp.sigmoid())
        return dict(neck_feats=feats, cls=torch.stack(cls), box=torch.stack(box), memory=memory.transpose(0, 1),
                    spatial_shapes=shapes, hs=torch.stack(hs))
