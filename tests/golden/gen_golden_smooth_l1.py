"""Generate tests/golden/losses_smooth_l1.npz by running the reference's own SmoothL1Loss / L1Loss
(mmdet/models/losses/smooth_l1_loss.py) -- the criterion of the head's `bbox` localisation distillation term.

    python tests/golden/gen_golden_smooth_l1.py          # needs /root/reference (this container only)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import  # noqa: E402

_ref_import.install()
from mmdet.models.losses.smooth_l1_loss import L1Loss, SmoothL1Loss  # noqa: E402

g = torch.Generator().manual_seed(31)
pred = torch.randn(40, 4, generator=g) * 1.5
tgt = torch.randn(40, 4, generator=g)
w = (torch.rand(40, 1, generator=g) > 0.5).float()
d = dict(pred=pred, target=tgt, weight=w)
for beta in (1.0, 0.11, 2.5):
    for red in ('none', 'mean', 'sum'):
        d[f'sl1.b{beta}.{red}'] = SmoothL1Loss(beta=beta, reduction=red, loss_weight=10.0)(pred, tgt)
    d[f'sl1.b{beta}.mean.w.avg9'] = SmoothL1Loss(beta=beta, reduction='mean', loss_weight=10.0)(pred, tgt, weight=w, avg_factor=9.0)
    p = pred.clone().requires_grad_(True)
    SmoothL1Loss(beta=beta, reduction='sum', loss_weight=0.5)(p, tgt, weight=w).backward()
    d[f'sl1.b{beta}.grad'] = p.grad
for red in ('none', 'mean', 'sum'):
    d[f'l1.{red}'] = L1Loss(reduction=red, loss_weight=5.0)(pred, tgt)
d['l1.mean.w.avg9'] = L1Loss(reduction='mean', loss_weight=5.0)(pred, tgt, weight=w, avg_factor=9.0)
p = pred.clone().requires_grad_(True)
L1Loss(reduction='sum')(p, tgt, weight=w).backward()
d['l1.grad'] = p.grad
np.savez_compressed(os.path.join(HERE, 'losses_smooth_l1.npz'), **{k: v.detach().numpy() for k, v in d.items()})
print('losses_smooth_l1.npz', len(d), 'keys')
