"""Next-row 4: the head's remaining distillation terms -- `soft` KL on logits
(gfl_deformable_detr_head_il.py:593-623), `bbox` / `logit` localisation KD (:625-645), whole-map `kldv` (:646-652) and
`memory` KL (:653-661) -- are calls of registry loss modules on tensors.  Each call pattern is evaluated with the CUDA
modules and with the oracle's restatement of the same modules, on the same synthetic inputs."""
import pytest
import torch

import dskd_b200
from dskd_b200 import synth
from oracle import boxes as obx
from oracle import losses as ol

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def both(cfg):
    return dskd_b200.build_loss(dict(cfg)), ol.build_loss(dict(cfg))


def check(mod_gpu, mod_cpu, pred, target, loss_rtol, **kw):
    pg = pred.to(DEV).clone().requires_grad_(True)
    pc = pred.clone().requires_grad_(True)
    kw_gpu = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
    lg = mod_gpu(pg, target.to(DEV), **kw_gpu)
    lc = mod_cpu(pc, target, **kw)
    lg.backward()
    lc.backward()
    torch.testing.assert_close(lg.detach().cpu().double(), lc.detach().double(), rtol=loss_rtol, atol=1e-12)
    torch.testing.assert_close(pg.grad.cpu(), pc.grad, rtol=1e-3, atol=1e-6 * float(pc.grad.abs().max()) + 1e-12)
    return float(lc.detach())


@pytest.fixture(scope='module')
def case():
    g = torch.Generator().manual_seed(17)
    N, Q, NC = 3, 100, 80
    inp = synth.make_distill_inputs(num_images=N, num_prev=40, seed=17, levels=((12, 20), (6, 10)), num_query=Q,
                                    img_hw=(96, 160), k_range=(4, 9))
    a = inp.assignments
    s_cls, t_cls = torch.randn(N, Q, NC, generator=g) - 2, torch.randn(N, Q, NC, generator=g) - 2
    s_box, t_box = torch.rand(N, Q, 70, generator=g), torch.rand(N, Q, 70, generator=g)
    keep = a['teacher_keepid']
    mask_student = torch.nonzero(a['student_labels'] < 40).squeeze(1)          # teacher_only_weights[-1] (:612)
    return dict(inp=inp, s_cls=s_cls, t_cls=t_cls, s_box=s_box, t_box=t_box, keep=keep, mask_student=mask_student, N=N, Q=Q)


def test_soft_kl_on_matched_logits(case):
    gpu, cpu = both(dict(type='KnowledgeDistillationKLDivLoss', loss_weight=1, T=2, reduction='mean'))   # loss_kd (:114)
    student = case['s_cls'].reshape(-1, 80)[case['mask_student']]
    teacher = case['t_cls'].reshape(-1, 80)[case['keep']]
    assert student.shape == teacher.shape
    check(gpu, cpu, student, teacher, 1e-3, weight=None, avg_factor=len(case['keep']))


def test_bbox_localisation_kd_smooth_l1(case):
    gpu, cpu = both(dict(type='SmoothL1Loss', loss_weight=10, reduction='mean'))                          # loss_ld_bbox (:118)
    wh_s = obx.integral_average(case['s_box'][:, :, 2:])
    wh_t = obx.integral_average(case['t_box'][:, :, 2:])
    pred = torch.cat((case['s_box'][:, :, :2].reshape(-1, 2), wh_s), 1)
    soft = torch.cat((case['t_box'][:, :, :2].reshape(-1, 2), wh_t), 1)
    weight = torch.zeros(pred.shape[0], 1)
    weight[case['keep']] = 1
    check(gpu, cpu, pred, soft, 1e-4, weight=weight, avg_factor=len(case['keep']))
    for beta in (0.11, 2.0):                                                   # both branches of the piecewise function
        g2, c2 = both(dict(type='SmoothL1Loss', beta=beta, loss_weight=1.0, reduction='sum'))
        check(g2, c2, pred * 3, soft, 1e-4)
    g3, c3 = both(dict(type='L1Loss', loss_weight=1.0, reduction='mean'))      # the commented alternative (:119)
    check(g3, c3, pred, soft, 1e-4, weight=weight, avg_factor=len(case['keep']))


def test_logit_localisation_kd(case):
    gpu, cpu = both(dict(type='KnowledgeDistillationKLDivLoss', loss_weight=1, T=2, reduction='mean'))   # loss_ld_logit (:121)
    pred = case['s_box'].reshape(-1, 70)
    soft = case['t_box'].reshape(-1, 70)
    weight = torch.zeros(pred.shape[0])
    weight[case['keep']] = 1
    check(gpu, cpu, pred, soft, 1e-3, weight=weight, avg_factor=len(case['keep']))


def test_whole_map_kldv_and_memory_kl(case):
    inp = case['inp']
    gpu, cpu = both(dict(type='KnowledgeDistillationKLDivLoss', loss_weight=1, T=2, reduction='sum'))    # loss_fd (:122)
    for sf, tf in zip(inp.student_feats, inp.teacher_feats):                   # softmax over dim=1 = channels of [N,C,H,W]
        check(gpu, cpu, sf, tf, 1e-3, weight=None, avg_factor=None)
    gpu, cpu = both(dict(type='KnowledgeDistillationKLDivLoss', loss_weight=2, T=2, reduction='sum'))    # loss_memory (:123)
    s_mem, t_mem = inp.memory()
    for s_img, t_img in zip(s_mem.permute(1, 2, 0), t_mem.permute(1, 2, 0)):   # [C, S] per image: softmax over the tokens
        check(gpu, cpu, s_img.contiguous(), t_img.contiguous(), 1e-3, weight=None, avg_factor=None)


def test_smooth_l1_and_l1_vs_reference_outputs():
    """The CUDA modules against the reference's own SmoothL1Loss / L1Loss outputs (tests/golden/losses_smooth_l1.npz)."""
    from conftest import Golden
    g = Golden('losses_smooth_l1.npz')
    pred, tgt, w = g.t('pred').to(DEV), g.t('target').to(DEV), g.t('weight').to(DEV)

    def close(a, b):
        torch.testing.assert_close(a.detach().cpu(), b, rtol=1e-5, atol=1e-6)
    for beta in (1.0, 0.11, 2.5):
        for red in ('none', 'mean', 'sum'):
            close(dskd_b200.SmoothL1Loss(beta, red, 10.0)(pred, tgt), g.t(f'sl1.b{beta}.{red}'))
        close(dskd_b200.SmoothL1Loss(beta, 'mean', 10.0)(pred, tgt, weight=w, avg_factor=9.0), g.t(f'sl1.b{beta}.mean.w.avg9'))
        p = pred.clone().requires_grad_(True)
        dskd_b200.SmoothL1Loss(beta, 'sum', 0.5)(p, tgt, weight=w).backward()
        close(p.grad, g.t(f'sl1.b{beta}.grad'))
    for red in ('none', 'mean', 'sum'):
        close(dskd_b200.L1Loss(red, 5.0)(pred, tgt), g.t(f'l1.{red}'))
    close(dskd_b200.L1Loss('mean', 5.0)(pred, tgt, weight=w, avg_factor=9.0), g.t('l1.mean.w.avg9'))
    p = pred.clone().requires_grad_(True)
    dskd_b200.L1Loss('sum')(p, tgt, weight=w).backward()
    close(p.grad, g.t('l1.grad'))
