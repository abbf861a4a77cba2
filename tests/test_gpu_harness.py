"""The training-step harness (next-row 1) hosts the CUDA distillation path end to end: a tiny detector, one step, every
loss finite, the teacher's detections reach the distillation losses, parameters move."""
import pytest
import torch

from dskd_b200.harness import IncrementalTrainStep, graph_detectors, make_student_teacher

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('criterion', ['kl', 'mse'])
def test_one_incremental_step_on_a_tiny_detector(criterion):
    student, teacher = make_student_teacher(DEV, detections_per_image=12, backbone='resnet18', enc_layers=1, dec_layers=2,
                                            num_query=50)
    student.train()
    trainer = IncrementalTrainStep(student, teacher, num_prev=40, criterion=criterion)
    g = torch.Generator(device=DEV).manual_seed(0)
    img = torch.randn(2, 3, 256, 320, device=DEV, generator=g)
    gt_b = [torch.tensor([[10., 20., 120., 200.], [100., 50., 300., 250.]], device=DEV), torch.zeros(0, 4, device=DEV)]
    gt_l = [torch.tensor([45, 70], device=DEV), torch.zeros(0, dtype=torch.long, device=DEV)]
    before = student.cls_branch.weight.detach().clone()
    out = trainer.step(img, gt_b, gt_l)
    assert out['num_teacher'] > 0, 'the calibrated teacher must produce pseudo labels'
    for k in ('loss', 'loss_det', 'loss_corr', 'loss_fg_feature'):
        assert torch.isfinite(out[k]), k
    assert float(out['loss_fg_feature']) >= 0 and float(out['loss_corr']) >= 0
    assert not torch.equal(before, student.cls_branch.weight.detach())
    out2 = trainer.step(img, gt_b, gt_l)
    assert torch.isfinite(out2['loss'])


def test_graphed_detectors_give_the_eager_step():
    """Teacher forward and student forward / backward as CUDA graphs (`graph_detectors`): same losses as the eager step
    on the same weights (dropout off so that the two runs are comparable), parameters move, a second replay works."""
    kw = dict(detections_per_image=12, backbone='resnet18', enc_layers=1, dec_layers=2, num_query=50)
    g = torch.Generator(device=DEV).manual_seed(0)
    img = torch.randn(2, 3, 256, 320, device=DEV, generator=g)
    gt_b = [torch.tensor([[10., 20., 120., 200.], [100., 50., 300., 250.]], device=DEV), torch.zeros(0, 4, device=DEV)]
    gt_l = [torch.tensor([45, 70], device=DEV), torch.zeros(0, dtype=torch.long, device=DEV)]
    outs = []
    for graphed in (False, True):
        student, teacher = make_student_teacher(DEV, seed=3, **kw)
        student.train()
        for m in student.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
            if isinstance(m, torch.nn.MultiheadAttention):
                m.dropout = 0.0
        model = student
        if graphed:
            student, teacher = graph_detectors(student, teacher, img)
        trainer = IncrementalTrainStep(student, teacher, num_prev=40, criterion='kl')
        before = model.cls_branch.weight.detach().clone()
        first = trainer.step(img, gt_b, gt_l)
        second = trainer.step(img, gt_b, gt_l)
        assert not torch.equal(before, model.cls_branch.weight.detach())
        outs.append((first, second))
    for k in ('loss', 'loss_det', 'loss_corr', 'loss_fg_feature'):
        torch.testing.assert_close(outs[1][0][k], outs[0][0][k], rtol=2e-3, atol=1e-5)
        assert torch.isfinite(outs[1][1][k])
    assert outs[1][0]['num_teacher'] == outs[0][0]['num_teacher'] > 0
