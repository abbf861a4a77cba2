"""Oracle: the registry loss modules and their reduction rules.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Restates
  mmdet/models/losses/utils.py:9-105       (reduce / weight / avg_factor rules)
  mmdet/models/losses/mse_loss.py:9-57     (MSELoss)
  mmdet/models/losses/kd_loss.py:12-94     (KnowledgeDistillationKLDivLoss)
  mmdet/models/losses/smooth_l1_loss.py:12-146 (SmoothL1Loss / L1Loss: the `bbox` localisation distillation term)
"""
import torch
import torch.nn.functional as F

_REDUCTIONS = (None, 'none', 'mean', 'sum')
EPS32 = float(torch.finfo(torch.float32).eps)


def reduce_elementwise(loss, weight=None, reduction='mean', avg_factor=None):
    """utils.py:30-59 -- elementwise weight, then none/mean/sum or avg_factor."""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        if reduction == 'none':
            return loss
        if reduction == 'mean':
            return loss.mean()
        if reduction == 'sum':
            return loss.sum()
        raise ValueError(f'{reduction} is not a valid value for reduction')
    if reduction == 'mean':
        return loss.sum() / (avg_factor + EPS32)
    if reduction == 'none':
        return loss
    raise ValueError('avg_factor can not be used with reduction="sum"')


def mse_elementwise(pred, target):
    """mse_loss.py:9-12."""
    return F.mse_loss(pred, target, reduction='none')


def kd_kl_elementwise(pred, soft_label, T):
    """kd_loss.py:12-43 -- softmax over dim=1, target detached, mean over dim 1, *T^2."""
    assert pred.size() == soft_label.size()
    target = F.softmax(soft_label / T, dim=1).detach()
    kl = F.kl_div(F.log_softmax(pred / T, dim=1), target, reduction='none')
    return kl.mean(1) * (T * T)


class MSELoss(torch.nn.Module):
    """mse_loss.py:15-57."""

    def __init__(self, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        assert reduction_override in _REDUCTIONS
        reduction = reduction_override if reduction_override else self.reduction
        return self.loss_weight * reduce_elementwise(
            mse_elementwise(pred, target), weight, reduction, avg_factor)


class KnowledgeDistillationKLDivLoss(torch.nn.Module):
    """kd_loss.py:46-94."""

    def __init__(self, reduction='mean', loss_weight=1.0, T=10):
        super().__init__()
        assert T >= 1
        self.reduction = reduction
        self.loss_weight = loss_weight
        self.T = T

    def forward(self, pred, soft_label, weight=None, avg_factor=None, reduction_override=None):
        assert reduction_override in _REDUCTIONS
        reduction = reduction_override if reduction_override else self.reduction
        return self.loss_weight * reduce_elementwise(
            kd_kl_elementwise(pred, soft_label, self.T), weight, reduction, avg_factor)


def smooth_l1_elementwise(pred, target, beta=1.0):
    """smooth_l1_loss.py:12-35."""
    assert beta > 0
    diff = torch.abs(pred - target)
    return torch.where(diff < beta, 0.5 * diff * diff / beta, diff - 0.5 * beta)


class SmoothL1Loss(torch.nn.Module):
    """smooth_l1_loss.py:59-101."""

    def __init__(self, beta=1.0, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.beta, self.reduction, self.loss_weight = beta, reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        assert reduction_override in _REDUCTIONS
        reduction = reduction_override if reduction_override else self.reduction
        return self.loss_weight * reduce_elementwise(smooth_l1_elementwise(pred, target, self.beta), weight, reduction,
                                                     avg_factor)


class L1Loss(torch.nn.Module):
    """smooth_l1_loss.py:104-146."""

    def __init__(self, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        assert reduction_override in _REDUCTIONS
        reduction = reduction_override if reduction_override else self.reduction
        return self.loss_weight * reduce_elementwise(torch.abs(pred - target), weight, reduction, avg_factor)


def build_loss(cfg):
    """builder.py:43-45 for the two types on the path."""
    cfg = dict(cfg)
    kind = cfg.pop('type')
    return {'MSELoss': MSELoss, 'SmoothL1Loss': SmoothL1Loss, 'L1Loss': L1Loss,
            'KnowledgeDistillationKLDivLoss': KnowledgeDistillationKLDivLoss}[kind](**cfg)
