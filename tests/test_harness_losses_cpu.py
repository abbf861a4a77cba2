"""The harness evaluates the detection losses densely with masks (no nonzero() / host sync per decoder layer).  The
quality-focal term must equal the indexed formulation of the reference (losses/gfocal_loss.py: zero-target BCE * sigma^2
for every (query, class), replaced at (positive query, its label) by BCE against the IoU * |IoU - sigma|^2)."""
import torch
import torch.nn.functional as F

from dskd_b200.harness import train_step as ts


def _qfl_indexed(cls, box70, labels, bt, num_classes=80, reg_max=16):
    wh = ts.integral_average(box70[:, 2:], reg_max)
    pred = torch.cat((box70[:, :2], wh), 1)
    pos = torch.nonzero(labels < num_classes).squeeze(1)
    num_pos = max(float(pos.numel()), 1.0)
    score = cls.new_zeros(labels.shape)
    score[pos] = ts._iou_giou(ts._cxcywh_to_xyxy(pred[pos]), ts._cxcywh_to_xyxy(bt[pos]))[0].detach()
    sig = cls.sigmoid()
    loss = F.binary_cross_entropy_with_logits(cls, torch.zeros_like(cls), reduction='none') * sig.pow(2)
    pl = labels[pos]
    sf = score[pos] - sig[pos, pl]
    loss[pos, pl] = F.binary_cross_entropy_with_logits(cls[pos, pl], score[pos], reduction='none') * sf.abs().pow(2)
    return 2.0 * loss.sum() / num_pos


def test_dense_quality_focal_term_equals_the_indexed_one():
    torch.manual_seed(0)
    M = 300
    cls = torch.randn(M, 80, requires_grad=True)
    box = (torch.rand(M, 70) * 0.8 + 0.1).requires_grad_(True)
    labels = torch.full((M,), 80)
    idx = torch.randperm(M)[:23]
    labels[idx] = torch.randint(0, 80, (23,))
    bt = torch.zeros(M, 4)
    bt[idx] = torch.rand(23, 4) * 0.4 + 0.2
    bw = torch.zeros(M, 4)
    bw[idx] = 1
    targets = dict(labels=labels, bbox_targets=bt, bbox_weights=bw)
    img_wh = torch.tensor([1333., 800., 1333., 800.])
    total = ts.detection_losses(cls, box, targets, img_wh=img_wh)
    g_total = torch.autograd.grad(total, cls)[0]
    g_ref = torch.autograd.grad(_qfl_indexed(cls, box, labels, bt), cls)[0]
    torch.testing.assert_close(g_total, g_ref, rtol=1e-6, atol=1e-9)     # only the QFL term depends on the logits
    # no positive at all: finite, and still the all-negative focal term
    none = dict(labels=torch.full((M,), 80), bbox_targets=torch.zeros(M, 4), bbox_weights=torch.zeros(M, 4))
    assert torch.isfinite(ts.detection_losses(cls, box, none, img_wh=img_wh))
