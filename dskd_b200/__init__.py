"""dskd_b200 -- B200-native (sm_100a) implementation of DSKD's distillation hot path.

Drop-in registry modules for the reference's incremental Deformable-DETR head
(`mmdet/models/dense_heads/gfl_deformable_detr_head_il.py`): DSG-FD, BCDD, the GFL Hungarian
assignment, and the MSE / KD-KL loss modules, all backed by hand-written CUDA kernels behind the
C ABI in include/dskd_b200.h.  Importing this package never falls back to a CPU implementation.
"""
from .registry import LOSSES, ASSIGNERS, build_loss, build_assigner, register_into_mmdet  # noqa: F401
from .losses import (DSGFeatureDistillLoss, BetweenClassDistanceLoss, MSELoss, SmoothL1Loss, L1Loss,  # noqa: F401
                     KnowledgeDistillationKLDivLoss)
from .assigner import GFLHungarianAssigner, AssignResult, lsap  # noqa: F401
from . import synth, teacher, dist  # noqa: F401
from .teacher import teacher_info_from_outputs  # noqa: F401

__version__ = '0.1.0'
register_into_mmdet()
