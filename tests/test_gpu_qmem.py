"""Query x memory contraction on tcgen05 (row A5, the unpinned extension) vs oracle/qmem.py.

Tolerances: the tensor cores read the fp32 operands as tf32 by TRUNCATING the low 13 mantissa bits (measured:
the kernel agrees with the oracle evaluated on truncated operands to ~4e-7); against the plain fp32 oracle the
cell weights therefore differ by a few 1e-4 absolute.  Both bounds are asserted.  The masked-MSE built on the
kernel's own weights is held to the usual rtol 1e-4 (loss) / 1e-3 (gradients)."""
import pytest
import torch

import dskd_b200
from dskd_b200 import qmem, synth
from dskd_b200._lib import DskdError
from oracle import losses as ol
from oracle import qmem as oq

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
SMALL = ((24, 40), (12, 20), (6, 10), (3, 5))
TIGHT, LOOSE = 5e-6, 2e-3


def make(n_images, levels, num_query, channels, kpi, seed=3, scores=True):
    cpu = synth.make_distill_inputs(num_images=n_images, num_prev=40, seed=seed, levels=levels, num_query=num_query,
                                    channels=channels, boxes_per_image=kpi, img_hw=(192, 320))
    counts = [b.shape[0] for b in cpu.assignments['teacher_bboxes']]
    g = torch.Generator().manual_seed(seed + 1)
    sc = (0.3 + 0.7 * torch.rand(sum(counts), generator=g)) if scores else None
    start = torch.tensor([0] + torch.tensor(counts).cumsum(0).tolist(), dtype=torch.int32)
    return cpu, counts, sc, start


def kernel_weights(cpu, counts, sc, start, temp=0.5):
    _, t_mem = cpu.memory()
    return qmem.qmem_cell_weights(t_mem.to(DEV), cpu.hs_teacher.to(DEV), cpu.assignments['teacher_keepid'].to(DEV),
                                  None if sc is None else sc.to(DEV), start.to(DEV), max(counts) if counts else 0, temp).cpu()


@pytest.fixture
def qmem_mode(monkeypatch):
    def set_mode(mode):
        if mode is None:
            monkeypatch.delenv('DSKD_QMEM_MODE', raising=False)
        else:
            monkeypatch.setenv('DSKD_QMEM_MODE', mode)      # '1': single-CTA kernel, '2': CTA-pair (cta_group::2) kernel
    return set_mode


@pytest.mark.parametrize('mode', [None, '1', '2'], ids=['auto', 'cta1', 'cta2'])
@pytest.mark.parametrize('n,levels,q,c,kpi,scores', [
    (2, SMALL, 100, 256, 20, True),            # one query block
    (2, SMALL, 100, 64, 7, True),              # 2 channel slabs
    (3, SMALL, 100, 256, None, True),          # ragged K_i
    (2, SMALL, 320, 256, 300, False),          # 300 queries: 2 blocks of 160 (single CTA) / one 304-wide pair block
    (2, SMALL, 320, 256, 250, True),           # pair mode: one MMA of N = 256 per k-step
    (1, SMALL, 700, 256, 600, True),           # four single blocks / two pair blocks
    (2, ((13, 21), (7, 11)), 60, 128, 5, True),  # S = 350: partial last token tile
    (3, ((9, 15),), 60, 64, 9, True),          # S = 135: one token tile in pair mode, the peer CTA half empty
], ids=['1blk', 'c64', 'ragged', 'k300', 'k250', 'k600', 'tail', 'tiny'])
def test_cell_weights_vs_oracle(n, levels, q, c, kpi, scores, mode, qmem_mode):
    qmem_mode(mode)
    cpu, counts, sc, start = make(n, levels, q, c, kpi, scores=scores)
    w = kernel_weights(cpu, counts, sc, start)
    _, t_mem = cpu.memory()
    a = cpu.assignments
    ref_tf32 = oq.qmem_cell_weights(t_mem, cpu.hs_teacher, a['teacher_keepid'], sc, counts, 0.5, tf32='trunc',
                                    dtype=torch.float64)
    ref_fp32 = oq.qmem_cell_weights(t_mem, cpu.hs_teacher, a['teacher_keepid'], sc, counts, 0.5)
    assert torch.isfinite(w).all() and float(w.min()) >= 0 and float(w.max()) <= 1
    torch.testing.assert_close(w.double(), ref_tf32, rtol=0, atol=TIGHT)
    torch.testing.assert_close(w, ref_fp32, rtol=0, atol=LOOSE)


def test_temperature_and_large_scores_do_not_overflow():
    cpu, counts, sc, start = make(2, SMALL, 100, 256, 12)
    big = synth.DistillInputs(cpu.student_feats, tuple(f * 6 for f in cpu.teacher_feats), cpu.hs_student,
                              cpu.hs_teacher * 6, cpu.assignments, cpu.spatial_shapes, cpu.levels)
    w = kernel_weights(big, counts, sc, start, temp=1.0)           # |z| reaches the hundreds
    _, t_mem = big.memory()
    ref = oq.qmem_cell_weights(t_mem, big.hs_teacher, cpu.assignments['teacher_keepid'], sc, counts, 1.0, tf32='trunc',
                               dtype=torch.float64)
    assert torch.isfinite(w).all()
    # |z| in the hundreds: fp32 rounding of z itself (1e-7 * 300) shows up in e^z
    torch.testing.assert_close(w.double(), ref, rtol=0, atol=1e-4)


@pytest.mark.parametrize('k,mode', [(9, None), (9, '2'), (200, None), (400, '2')], ids=['k9', 'k9-pair', 'k200', 'k400-pair'])
def test_images_without_detections_get_zero_weight(k, mode, qmem_mode):
    qmem_mode(mode)
    cpu, counts, sc, start = make(3, SMALL, max(100, k), 256, k)
    # drop the detections of image 1
    a = cpu.assignments
    k0, k1, k2 = counts
    keep = torch.cat([a['teacher_keepid'][:k0], a['teacher_keepid'][k0 + k1:]])
    sc2 = torch.cat([sc[:k0], sc[k0 + k1:]])
    start2 = torch.tensor([0, k0, k0, k0 + k2], dtype=torch.int32)
    _, t_mem = cpu.memory()
    w = qmem.qmem_cell_weights(t_mem.to(DEV), cpu.hs_teacher.to(DEV), keep.to(DEV), sc2.to(DEV), start2.to(DEV),
                               max(k0, k2)).cpu()
    assert float(w[1].abs().max()) == 0.0 and float(w[0].max()) > 0 and float(w[2].max()) > 0
    ref = oq.qmem_cell_weights(t_mem, cpu.hs_teacher, keep, sc2, [k0, 0, k2], 0.5, tf32='trunc', dtype=torch.float64)
    torch.testing.assert_close(w.double(), ref, rtol=0, atol=TIGHT)
    none = qmem.qmem_cell_weights(t_mem.to(DEV), cpu.hs_teacher.to(DEV), keep[:0].to(DEV), None,
                                  torch.zeros(4, dtype=torch.int32, device=DEV), 0).cpu()
    assert float(none.abs().max()) == 0.0


@pytest.mark.parametrize('reduction', ['sum', 'mean'])
def test_module_loss_and_gradient(reduction):
    cpu, counts, sc, start = make(2, SMALL, 100, 256, 15)
    gpu = cpu.to(DEV)
    a = dict(gpu.assignments, teacher_scores=sc.to(DEV))
    mod = dskd_b200.build_loss(dict(type='DSGFeatureDistillLoss', criterion='mse', reduction=reduction, loss_weight=0.7,
                                    mask_mode='qmem', feature_source='memory', temp=0.5))
    s_mem, t_mem = gpu.memory()
    s_mem = s_mem.clone().requires_grad_(True)
    loss = mod((s_mem, cpu.spatial_shapes), (t_mem, cpu.spatial_shapes), (gpu.hs_student, gpu.hs_teacher), a)
    loss.backward()
    # oracle on the kernel's own weights: the streaming part must be exact to the usual tolerance
    cs, ct = cpu.memory()
    cs = cs.clone().requires_grad_(True)
    crit = ol.MSELoss(reduction, 0.7)
    ref = oq.qmem(cs, ct, cpu.levels, cpu.hs_teacher, cpu.assignments['teacher_keepid'], sc, counts, crit,
                  weights=mod.last_cell_weights.cpu())
    ref.backward()
    torch.testing.assert_close(loss.detach().cpu().double(), ref.detach().double(), rtol=1e-4, atol=1e-12)
    torch.testing.assert_close(s_mem.grad.cpu(), cs.grad, rtol=1e-3, atol=1e-6 * float(cs.grad.abs().max()))
    # end to end against the fp32 definition: tf32 operands move the loss by ~1e-3 at most
    cs2 = cpu.memory()[0].clone().requires_grad_(True)
    full = oq.qmem(cs2, ct, cpu.levels, cpu.hs_teacher, cpu.assignments['teacher_keepid'], sc, counts, crit)
    torch.testing.assert_close(loss.detach().cpu().double(), full.detach().double(), rtol=2e-3, atol=1e-12)


def test_full_size_coco_batch2_two_query_blocks():
    cpu, counts, sc, start = make(2, synth.COCO_LEVELS, 320, 256, 300, seed=11)
    w = kernel_weights(cpu, counts, sc, start, temp=2.0)
    _, t_mem = cpu.memory()
    ref = oq.qmem_cell_weights(t_mem, cpu.hs_teacher, cpu.assignments['teacher_keepid'], sc, counts, 2.0, tf32='trunc',
                               dtype=torch.float64)
    torch.testing.assert_close(w.double(), ref, rtol=0, atol=TIGHT)


def test_argument_errors():
    cpu, counts, sc, start = make(1, SMALL, 100, 48, 5)      # C = 48 is not a multiple of 32
    with pytest.raises(DskdError):
        kernel_weights(cpu, counts, sc, start)
    cpu, counts, sc, start = make(1, SMALL, 100, 64, 5)
    with pytest.raises(DskdError):
        kernel_weights(cpu, counts, sc, start, temp=0.0)
    with pytest.raises(ValueError):
        dskd_b200.build_loss(dict(type='DSGFeatureDistillLoss', mask_mode='qmem', feature_source='neck'))
