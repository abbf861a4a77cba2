"""CPU oracle for the DSKD distillation hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain PyTorch (fp32, CPU) restatement of the arithmetic the
reference (smilekitty7/DSKD, an MMDetection 2.23 fork, 100 % Python) performs on
the path SURVEY.md section 8 scopes: DSG-FD (`decode_v1` & sibling masks), BCDD
(prototypes + class-distance matrix), the GFL Hungarian assignment that feeds
them, and the registry loss modules that reduce them.

Rules (enforced by tests/test_no_oracle_in_product.py):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
    `--impl reference` legs may import anything from here;
  * nothing under `dskd_b200/` imports it -- the product path is CUDA only and
    fails loudly when `libdskd_b200.so` is missing.

Parity pinning: the reference ships no test or golden vector for any of the
incremental-learning code.  The oracle is therefore
pinned against *outputs of the reference itself*: `tests/golden/gen_golden.py`
imports the unmodified reference modules from /root/reference (with a stub
`mmcv`, which only provides decorators / registries on this path), runs
`GFLDeformableDETRHead_il.loss`, `GFLHungarianAssigner.assign`,
`correlation_mat`, `MSELoss`, `KnowledgeDistillationKLDivLoss`, ... on seeded
inputs and stores inputs + outputs under `tests/golden/*.npz`.
`tests/test_oracle_golden.py` checks every oracle function against them, plus
the known answers of the reference's own doctests / unit tests
(`match_cost.py:446-453`, `losses/utils.py:72-90`,
`tests/test_metrics/test_losses.py:82-109`).

Third-party arithmetic on the path that is not vendored in the reference:
  * `scipy.optimize.linear_sum_assignment` (un-pinned in
    `requirements/optional.txt:3`; 1.18.1 installed here): modified
    Jonker-Volgenant shortest augmenting path, float64.  The oracle calls SciPy
    itself; the product re-implements it in C++ (`dskd_b200/csrc/lsap.cpp`).
  * ATen ops (softmax, kl_div, mse_loss, cdist, dist, bce_with_logits): called
    through torch 2.11 CPU.
"""

from . import boxes, losses, assign, dsgfd, bcdd, qmem  # noqa: F401
