"""One 40+40 incremental training step hosting the distillation hot path.

Call stack mirrored (mmdet/models/detectors/deformable_detr_il.py:255-319 -> gfl_deformable_detr_head_il.py:412-1195):
  teacher forward (no_grad) -> teacher keep-ids            dskd_b200.teacher_info_from_outputs      (CUDA)
  pseudo labels gt := cat(teacher_pred, gt) (:462-465)     torch.cat
  6 x N Hungarian assignments (:504-512, :1670-1797)       GFLHungarianAssigner.assign_batch        (CUDA + C++ LSAP)
  QFL / DFL / L1 / GIoU per decoder layer (:1379-1533)     plain PyTorch below (stock detection losses, out of scope)
  BCDD  loss_corr (:525-555)                               BetweenClassDistanceLoss                 (CUDA)
  DSG-FD loss_fg_feature, decode_v1 (:664-719)             DSGFeatureDistillLoss                    (CUDA)
  backward, grad-clip 0.1, AdamW (config :214-225)         torch
"""
import copy
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import build_loss, GFLHungarianAssigner, teacher_info_from_outputs
from .model import GFLDeformableDETR


def _cxcywh_to_xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def _iou_giou(a, b, eps=1e-6):
    """aligned IoU and GIoU of xyxy boxes (iou2d_calculator.py:192-261, is_aligned=True)."""
    lt, rb = torch.max(a[:, :2], b[:, :2]), torch.min(a[:, 2:], b[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    area = lambda x: (x[:, 2] - x[:, 0]) * (x[:, 3] - x[:, 1])
    union = (area(a) + area(b) - inter).clamp(min=eps)
    iou = inter / union
    e_wh = (torch.max(a[:, 2:], b[:, 2:]) - torch.min(a[:, :2], b[:, :2])).clamp(min=0)
    e_area = (e_wh[:, 0] * e_wh[:, 1]).clamp(min=eps)
    return iou, iou - (e_area - union) / e_area


def integral_average(lrtb, reg_max=16):
    """head_il.py:42-59."""
    x = lrtb.reshape(-1, reg_max + 1)
    x = x / x.sum(1, keepdim=True)
    space = torch.linspace(0, reg_max, reg_max + 1, device=x.device) / reg_max / 2
    return (x * space).sum(1).reshape(-1, 2, 2).sum(2)


def detection_losses(cls, box70, targets, num_classes=80, reg_max=16, img_wh=None):
    """QFL (beta 2, w 2.0) + L1 (w 5.0) + GIoU (w 2.0) + DFL (w 0.5) for ONE decoder layer (head_il.py:1379-1533,
    losses/gfocal_loss.py).  cls [M,80], box70 [M,70] sigmoid outputs, targets: labels / bbox_targets / bbox_weights [M(,4)]."""
    labels, bt, bw = targets['labels'], targets['bbox_targets'], targets['bbox_weights']
    wh = integral_average(box70[:, 2:], reg_max)
    pred = torch.cat((box70[:, :2], wh), 1)
    # dense, mask-based evaluation: no nonzero() / host sync per decoder layer (the reference indexes `pos_inds`)
    pos = labels < num_classes                                                     # [M] bool
    num_pos = pos.sum().clamp(min=1).to(cls.dtype)
    iou = _iou_giou(_cxcywh_to_xyxy(pred), _cxcywh_to_xyxy(bt))[0].detach()
    score = torch.where(pos, iou, torch.zeros_like(iou))                           # IoU target of the positives
    # QualityFocalLoss
    sig = cls.sigmoid()
    neg_term = F.binary_cross_entropy_with_logits(cls, torch.zeros_like(cls), reduction='none') * sig.pow(2)
    onehot = pos[:, None] & (labels[:, None] == torch.arange(cls.shape[1], device=cls.device)[None, :])
    tgt = score[:, None].expand_as(cls)
    pos_term = F.binary_cross_entropy_with_logits(cls, tgt, reduction='none') * (tgt - sig).abs().pow(2)
    loss_cls = 2.0 * torch.where(onehot, pos_term, neg_term).sum() / num_pos
    factor = img_wh
    giou = _iou_giou(_cxcywh_to_xyxy(pred) * factor, _cxcywh_to_xyxy(bt) * factor)[1]
    loss_iou = 2.0 * ((1 - giou) * bw[:, 0]).sum() / num_pos
    loss_bbox = 5.0 * ((pred - bt).abs() * bw).sum() / num_pos
    # DistributionFocalLoss on the 4 x (reg_max+1) bins
    corners = box70[:, 2:].reshape(-1, reg_max + 1)
    target = (bt[:, 2:].unsqueeze(2).repeat(1, 1, 2).reshape(-1) / 2 * 2 * reg_max).clamp(0, reg_max - 0.01)
    dl = target.long()
    dr = dl + 1
    logp = torch.log(corners.clamp(min=1e-12) / corners.sum(1, keepdim=True).clamp(min=1e-12))
    dfl = -(logp.gather(1, dl[:, None]).squeeze(1) * (dr.float() - target) +
            logp.gather(1, dr[:, None]).squeeze(1) * (target - dl.float()))
    loss_dfl = 0.5 * (dfl * bw.reshape(-1)).sum() / (4 * num_pos)
    return loss_cls + loss_iou + loss_bbox + loss_dfl


def make_student_teacher(device, detections_per_image=30, num_prev=40, seed=0, **model_kw):
    """Random-init student; teacher = deepcopy (tools/train_increment.py:250-251), frozen.  A random detector fires no
    detection above score_thr = 0.3, so the teacher's class bias is calibrated on a probe batch to keep about
    `detections_per_image` (query, class) pairs of the PREVIOUS classes per image -- synthetic stand-in for a trained task-1 model."""
    torch.manual_seed(seed)
    student = GFLDeformableDETR(**model_kw).to(device)
    teacher = copy.deepcopy(student).eval()
    for p in teacher.parameters():
        p.requires_grad_(False)
    with torch.no_grad():
        probe = teacher(torch.randn(1, 3, 256, 320, device=device))['cls'][-1][0, :, :num_prev]      # [Q, prev]
        k = min(detections_per_image, probe.numel() - 1)
        kth = probe.flatten().topk(k + 1).values[-1]
        shift = math.log(0.3 / 0.7) - float(kth) + 1e-3
        teacher.cls_branch.bias[:num_prev] += shift
        teacher.cls_branch.bias[num_prev:] = -20.0
    return student, teacher


_TRUNK_KEYS = ('cls', 'box', 'hs')


class DetectorTrunk(nn.Module):
    """The detector as a tensors-in / tensors-out module (what `torch.cuda.make_graphed_callables` can capture): returns
    (neck level 0..3, cls, box, hs); `as_outputs` turns that back into the dict the step consumes."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, img):
        o = self.model(img)
        return (*o['neck_feats'], *(o[k] for k in _TRUNK_KEYS))

    @staticmethod
    def as_outputs(flat):
        nl = len(flat) - len(_TRUNK_KEYS)
        out = dict(zip(_TRUNK_KEYS, flat[nl:]))
        out['neck_feats'] = tuple(flat[:nl])
        return out


class GraphedTeacher:
    """The frozen teacher's forward as ONE CUDA-graph replay: static shapes, no gradient, identical every iteration.  The
    outputs are static tensors that the next replay overwrites (the step consumes them before it calls again)."""

    def __init__(self, teacher, sample_img):
        self.teacher = teacher      # the graph bakes the parameters' addresses in: the module must outlive it
        self.static_img = sample_img.detach().clone()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(2):      # lazy tables (sampling-offset normalisers, cuDNN plans) are built outside the capture
                teacher(self.static_img)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = teacher(self.static_img)

    def __call__(self, img):
        if img.shape != self.static_img.shape:
            raise ValueError(f'the teacher graph was captured for images of shape {tuple(self.static_img.shape)}')
        self.static_img.copy_(img)
        self.graph.replay()
        return self.out


def graph_detectors(student, teacher, sample_img):
    """CUDA graphs for the static-shape parts of the step: the teacher forward (one graph) and the student detector's
    forward and backward (`torch.cuda.make_graphed_callables`: one graph each, parameters stay ordinary autograd leaves, so
    DDP and the optimizer see nothing new).  The heads' losses stay eager: the number of teacher detections changes the
    shapes there every iteration.  Returns (student trunk, teacher callable); wrap the trunk in DDP AFTER this call."""
    # the parameters' AccumulateGrad nodes are first touched on the capture side stream; later backward passes run on the
    # caller's stream, which is intended (the replay is ordered on that stream)
    if hasattr(torch.autograd.graph, 'set_warn_on_accumulate_grad_stream_mismatch'):
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    g_teacher = GraphedTeacher(teacher, sample_img)
    trunk = torch.cuda.make_graphed_callables(DetectorTrunk(student), (sample_img.detach().clone(),))
    return trunk, g_teacher


class IncrementalTrainStep:
    """student / teacher + the drop-in distillation path + AdamW: `step(img, gt_bboxes, gt_labels)` runs one iteration.
    `student` is a GFLDeformableDETR or a (graphed) DetectorTrunk of one, bare or wrapped in DDP; `teacher` any callable
    returning the detector's output dict."""

    def __init__(self, student, teacher, num_prev=40, criterion='kl', lr=4e-4, sync_prototypes=False):
        self.student, self.teacher = student, teacher
        self.prev_labels = list(range(num_prev))
        self.assigner = GFLHungarianAssigner()
        self.loss_corr = build_loss(dict(type='BetweenClassDistanceLoss', reduction='mean', loss_weight=1.0,
                                         sync_prototypes=sync_prototypes))
        self.loss_fg = build_loss(dict(type='DSGFeatureDistillLoss', criterion=criterion, T=2.0, reduction='sum',
                                       loss_weight=1.0, mask_mode='decode_v1', feature_source='neck'))
        params = [p for p in self.module.parameters() if p.requires_grad]
        self.opt = torch.optim.AdamW(params, lr=lr, weight_decay=1e-4)
        self.params = params

    @property
    def module(self):
        m = self.student.module if hasattr(self.student, 'module') else self.student      # DDP
        return m.model if isinstance(m, DetectorTrunk) else m

    def _student_forward(self, img):
        out = self.student(img)
        return out if isinstance(out, dict) else DetectorTrunk.as_outputs(out)

    def step(self, img, gt_bboxes, gt_labels):
        N, _, H, W = img.shape
        dev = img.device
        img_shapes = [(H, W)] * N
        with torch.no_grad():
            t = self.teacher(img)
        s = self._student_forward(img)      # queued behind the teacher before the keep-ids' one host sync drains the stream
        with torch.no_grad():
            tinfo = teacher_info_from_outputs(t['cls'][-1], t['box'][-1], img_shapes, score_thr=0.3, max_per_img=100)
        # hard + teacher-first pseudo labels (head_il.py:462-465)
        all_b = [torch.cat([tb, gb]) for tb, gb in zip(tinfo['pred_bboxes'], gt_bboxes)]
        all_l = [torch.cat([tl, gl]) for tl, gl in zip(tinfo['pred_labels'], gt_labels)]
        tg = self.assigner.assign_batch(s['cls'].detach(), s['box'].detach(), all_b, all_l, img_shapes,
                                        prev_labels=self.prev_labels)
        L, Q = s['cls'].shape[0], s['cls'].shape[2]
        M = N * Q
        img_wh = torch.tensor([W, H, W, H], dtype=torch.float32, device=dev)
        loss = 0.
        for l in range(L):
            sl = slice(l * M, (l + 1) * M)
            loss = loss + detection_losses(s['cls'][l].reshape(M, -1), s['box'][l].reshape(M, -1),
                                           dict(labels=tg['labels'][sl], bbox_targets=tg['bbox_targets'][sl],
                                                bbox_weights=tg['bbox_weights'][sl]), img_wh=img_wh)
        assignments = dict(student_labels=tg['labels'][(L - 1) * M:], teacher_keepid=tinfo['pred_keepid'],
                           teacher_labels=tinfo['cat_labels'], teacher_bboxes=tinfo['pred_bboxes'],
                           img_shapes=img_shapes, prev_labels=self.prev_labels, num_classes=self.module.num_classes)
        queries = (s['hs'][-1], t['hs'][-1])
        loss_corr = self.loss_corr(None, None, queries, assignments)
        loss_fg = self.loss_fg(s['neck_feats'], t['neck_feats'], queries, assignments)
        total = loss + loss_corr + loss_fg
        self.opt.zero_grad(set_to_none=True)
        total.backward()
        torch.nn.utils.clip_grad_norm_(self.params, 0.1)
        self.opt.step()
        return dict(loss=total.detach(), loss_det=loss.detach(), loss_corr=loss_corr.detach(), loss_fg_feature=loss_fg.detach(),
                    num_teacher=int(tinfo['pred_keepid'].numel()))


def synthetic_batch(device, images, height=800, width=1333, seed=1234):
    """Synthetic 800x1333 images with 1-10 ground-truth boxes of the NEW classes (40..79) each."""
    g = torch.Generator(device=device).manual_seed(seed)
    img = torch.randn(images, 3, height, width, device=device, generator=g)
    gt_b, gt_l = [], []
    for _ in range(images):
        k = int(torch.randint(1, 11, (1,), device=device, generator=g))
        x1 = torch.rand(k, device=device, generator=g) * 0.7 * width
        y1 = torch.rand(k, device=device, generator=g) * 0.7 * height
        bw = 8 + torch.rand(k, device=device, generator=g) * (0.3 * width - 8)
        bh = 8 + torch.rand(k, device=device, generator=g) * (0.3 * height - 8)
        gt_b.append(torch.stack([x1, y1, (x1 + bw).clamp(max=width), (y1 + bh).clamp(max=height)], 1))
        gt_l.append(torch.randint(40, 80, (k,), device=device, generator=g))
    return img, gt_b, gt_l


def bench_train_step(device, rank, world, dist, images_per_gpu=4, criterion='kl', steps=3, warmup=2, height=800, width=1333,
                     backbone='resnet50', graphs=True, channels_last=False):
    """Time the 40+40 incremental training step on synthetic data (bench.py's `train_step` key, tools/train_step_bench.py):
    CUDA events around `steps` iterations after `warmup`, max over ranks; DDP (gradient mean) + prototype all-reduce when
    world > 1, like tools/train_increment.py:299-304.  graphs: teacher forward and student detector forward / backward as
    CUDA graphs (`graph_detectors`); if the capture fails the step runs eagerly and the error is reported."""
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    try:
        student, teacher = make_student_teacher(device, backbone=backbone)
        student.train()
        if channels_last:
            student.use_channels_last()
            teacher.use_channels_last()
        img, gt_b, gt_l = synthetic_batch(device, images_per_gpu, height, width, seed=1234 + rank)
        launch = 'eager'
        if graphs:
            try:
                student, teacher = graph_detectors(student, teacher, img)
                launch = 'cuda graphs: teacher forward, student detector forward + backward; heads / losses / optimizer eager'
            except Exception as exc:        # noqa: BLE001 -- report, fall back to eager launches
                launch = f'eager (graph capture failed: {type(exc).__name__}: {str(exc)[:200]})'
        if world > 1:
            # every parameter of the detector receives a gradient in this step (DDP itself reports so)
            student = torch.nn.parallel.DistributedDataParallel(student, device_ids=[device.index], broadcast_buffers=False)
        trainer = IncrementalTrainStep(student, teacher, num_prev=40, criterion=criterion, sync_prototypes=world > 1)
        out = None
        for _ in range(max(warmup, 1)):
            out = trainer.step(img, gt_b, gt_l)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = trainer.step(img, gt_b, gt_l)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        return {'metric': 'incremental_train_step_images_per_s', 'value': world * images_per_gpu * steps / (ms * 1e-3),
                'unit': 'images/s', 'n_gpus': world, 'steps': steps, 'warmup': max(warmup, 1), 'ms_per_step': ms / steps,
                'dtype': 'f32 (tf32 matmul/conv)', 'data': 'synthetic',
                'config': {'workload': 'coco_40+40_incremental_train_step', 'images_per_gpu': images_per_gpu,
                           'image': [height, width], 'backbone': backbone, 'criterion': criterion, 'queries': 300,
                           'decoder_layers': 6, 'parallelism': f'dp{world}', 'launch': launch,
                           'conv_layout': 'channels_last' if channels_last else 'nchw'},
                'losses': {k: float(v) for k, v in out.items()},
                'peak_mem_gb': torch.cuda.max_memory_allocated(device) / 2 ** 30}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
