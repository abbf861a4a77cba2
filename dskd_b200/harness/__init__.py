"""Training-step harness: a minimal pure-PyTorch GFL-Deformable-DETR that HOSTS the
distillation hot path in a 40+40 incremental training step.  The detector itself is plumbing (stock architecture,
random init, no parity claim -- the reference's transformer needs mmcv's CUDA op, which is absent); what is under test
is the drop-in path: teacher keep-ids -> pseudo labels -> batched Hungarian assignment -> BCDD + DSG-FD modules."""
from .model import GFLDeformableDETR  # noqa: F401
from .train_step import (DetectorTrunk, GraphedTeacher, IncrementalTrainStep, bench_train_step,  # noqa: F401
                         graph_detectors, make_student_teacher, synthetic_batch)
