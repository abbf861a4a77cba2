"""Pin the oracle: every restated function against what the REFERENCE ITSELF returned
(tests/golden/*.npz, written by tests/golden/gen_golden.py) and against the known answers the
reference's own tests / doctests hold for this path.  CPU only."""
import pytest
import torch

import oracle
from oracle import assign as oa
from oracle import bcdd as ob
from oracle import boxes as obx
from oracle import dsgfd as od
from oracle import losses as ol
from conftest import Golden, load_head_case


def close(a, b, rtol=1e-6, atol=1e-7):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


# ------------------------------------------------------------------ loss modules (R1-R3)
def test_loss_modules_vs_reference():
    g = Golden('losses.npz')
    pred, tgt, w = g.t('pred'), g.t('target'), g.t('weight')
    for red in ('none', 'mean', 'sum'):
        for lw in (1.0, 0.37):
            close(ol.MSELoss(red, lw)(pred, tgt), g.t(f'mse.{red}.{lw}'))
            close(ol.MSELoss(red, lw)(pred, tgt, weight=w), g.t(f'mse.{red}.{lw}.w'))
        for T in (1, 2, 10):
            close(ol.KnowledgeDistillationKLDivLoss(red, 1.5, T)(pred, tgt), g.t(f'kd.{red}.T{T}'))
    close(ol.MSELoss('mean')(pred, tgt, weight=w, avg_factor=7.0), g.t('mse.mean.avg7'))
    close(ol.MSELoss('none')(pred, tgt, avg_factor=7.0), g.t('mse.none.avg7'))
    close(ol.MSELoss('mean')(pred, tgt, reduction_override='sum'), g.t('mse.override_sum'))
    close(ol.KnowledgeDistillationKLDivLoss('sum', 1.0, 2)(pred, tgt, weight=g.t('kd.weight')), g.t('kd.sum.T2.w'))
    close(ol.KnowledgeDistillationKLDivLoss('mean', 1.0, 2)(pred, tgt, avg_factor=3.0), g.t('kd.mean.T2.avg3'))
    close(ol.KnowledgeDistillationKLDivLoss('mean', 1.0, 10)(g.t('pred2'), g.t('target2')), g.t('kd2.mean.T10'))
    p, t = pred.clone().requires_grad_(True), tgt.clone().requires_grad_(True)
    ol.MSELoss('sum', 0.37)(p, t, weight=w).backward()
    close(p.grad, g.t('mse.grad_pred'))
    close(t.grad, g.t('mse.grad_target'))
    p, t = pred.clone().requires_grad_(True), tgt.clone().requires_grad_(True)
    ol.KnowledgeDistillationKLDivLoss('sum', 1.5, 2)(p, t).backward()
    close(p.grad, g.t('kd.grad_pred'))
    assert t.grad is None and g.v('kd.grad_target_is_none')


def test_smooth_l1_and_l1_vs_reference():
    g = Golden('losses_smooth_l1.npz')
    pred, tgt, w = g.t('pred'), g.t('target'), g.t('weight')
    for beta in (1.0, 0.11, 2.5):
        for red in ('none', 'mean', 'sum'):
            close(ol.SmoothL1Loss(beta, red, 10.0)(pred, tgt), g.t(f'sl1.b{beta}.{red}'))
        close(ol.SmoothL1Loss(beta, 'mean', 10.0)(pred, tgt, weight=w, avg_factor=9.0), g.t(f'sl1.b{beta}.mean.w.avg9'))
        p = pred.clone().requires_grad_(True)
        ol.SmoothL1Loss(beta, 'sum', 0.5)(p, tgt, weight=w).backward()
        close(p.grad, g.t(f'sl1.b{beta}.grad'))
    for red in ('none', 'mean', 'sum'):
        close(ol.L1Loss(red, 5.0)(pred, tgt), g.t(f'l1.{red}'))
    close(ol.L1Loss('mean', 5.0)(pred, tgt, weight=w, avg_factor=9.0), g.t('l1.mean.w.avg9'))
    p = pred.clone().requires_grad_(True)
    ol.L1Loss('sum')(p, tgt, weight=w).backward()
    close(p.grad, g.t('l1.grad'))


def test_loss_module_known_answers():
    # reference tests/test_metrics/test_losses.py:82-109 and tests/test_models/test_loss.py:28-88
    with pytest.raises(AssertionError):
        ol.KnowledgeDistillationKLDivLoss(T=0.5)
    kd = ol.KnowledgeDistillationKLDivLoss(loss_weight=1.0, T=1)
    with pytest.raises(AssertionError):
        kd(torch.Tensor([[5, -5, 0]]), torch.Tensor([[1, 0]]))
    assert torch.allclose(kd(torch.Tensor([[1, 2, 0]]), torch.Tensor([[1, 2, 0]])), torch.tensor(0.0))
    pred, tgt = torch.rand(1, 4), torch.rand(1, 4)
    with pytest.raises(ValueError):
        ol.MSELoss()(pred, tgt, avg_factor=10, reduction_override='sum')
    with pytest.raises(AssertionError):
        ol.MSELoss()(pred, tgt, reduction_override=True)
    ol.MSELoss()(torch.rand(0, 4), torch.rand(0, 4))
    # losses/utils.py:72-90 doctest through the same reduce rules
    l1 = (torch.Tensor([0, 2, 3]) - torch.Tensor([1, 1, 1])).abs()
    wgt = torch.Tensor([1, 0, 1])
    close(ol.reduce_elementwise(l1), torch.tensor(1.3333), rtol=1e-4, atol=1e-4)
    close(ol.reduce_elementwise(l1, wgt), torch.tensor(1.0))
    close(ol.reduce_elementwise(l1, wgt, avg_factor=2), torch.tensor(1.5))


# ------------------------------------------------------------------ boxes
def test_boxes_vs_reference():
    g = Golden('boxes.npz')
    a, b = g.t('a'), g.t('b')
    close(obx.bbox_overlaps(a, b, 'iou'), g.t('iou'), 0, 0)
    close(obx.bbox_overlaps(a, b, 'giou'), g.t('giou'), 0, 0)
    close(obx.bbox_overlaps(a[:11], b, 'iou', True), g.t('iou_aligned'), 0, 0)
    close(obx.bbox_overlaps(a[:11], b, 'giou', True), g.t('giou_aligned'), 0, 0)
    close(obx.integral_average(g.t('lrtb')), g.t('integral'), 0, 0)
    close(obx.xyxy_to_cxcywh(a), g.t('to_cxcywh'), 0, 0)
    close(obx.cxcywh_to_xyxy(obx.xyxy_to_cxcywh(a)), g.t('to_xyxy'), 0, 0)


def test_iou_cost_doctest():
    # match_cost.py:446-453
    bboxes = torch.FloatTensor([[1, 1, 2, 2], [2, 2, 3, 4]])
    gts = torch.FloatTensor([[0, 0, 2, 4], [1, 2, 3, 4]])
    close(-obx.bbox_overlaps(bboxes, gts, 'giou'),
          torch.tensor([[-0.1250, 0.1667], [0.1667, -0.5000]]), rtol=1e-3, atol=1e-4)


# ------------------------------------------------------------------ assignment (A1, H1-H3)
def test_assign_vs_reference():
    g = Golden('assign.npz')
    for c in range(g.v('num_cases')):
        p = f'case{c}.'
        cls, cx, gt, lab = g.t(p + 'cls'), g.t(p + 'cxcywh'), g.t(p + 'gt'), g.t(p + 'lab')
        close(obx.decode_cxcywh(g.t(p + 'box')), cx, 0, 0)
        if lab.numel():
            cost = oa.cost_matrix(cx, cls, gt, lab, (800, 1333), w_cls=g.v(p + 'w_cls'))
            ref = g.t(p + 'cls_cost') + g.t(p + 'reg_cost') + g.t(p + 'iou_cost')
            close(cost, ref, 0, 0)
        else:
            cost = None
        gt_inds, labels = oa.hungarian_assign(cost, lab, cx.size(0))
        assert torch.equal(gt_inds, g.t(p + 'gt_inds'))
        assert torch.equal(labels, g.t(p + 'labels'))
    s, l, k = oa.filter_scores_and_topk(g.t('topk.scores_in'), 0.3, 100)
    close(s, g.t('topk.scores'), 0, 0)
    assert torch.equal(l, g.t('topk.labels')) and torch.equal(k, g.t('topk.keep'))
    det, _, labels, keep = oa.teacher_decode_single(g.t('decode.cls'), g.t('decode.box'), (800, 1333))
    close(det, g.t('decode.det')[:, :4], 0, 0)
    assert torch.equal(labels, g.t('decode.labels')) and torch.equal(keep, g.t('decode.keep'))


# ------------------------------------------------------------------ the head's distillation block
def _head_oracle(inp, out, mode, crit):
    """Run the oracle on a stored head case; returns dict of losses / labels / grads."""
    L = inp.v('L')
    levels = [tuple(x) for x in inp.t('levels').tolist()]
    img_hw = tuple(inp.t('img_hw').tolist())
    N, Q = inp.t('s_cls').shape[1:3]
    img_shapes = [img_hw] * N
    prev = list(range(L))
    tinfo = oa.teacher_info_from_outputs(inp.t('t_cls')[-1], inp.t('t_box')[-1], img_shapes)
    gt_b, gt_l = oa.merge_pseudo_labels(tinfo['pred_bboxes'], tinfo['pred_labels'],
                                        inp.lst('gt_bboxes'), inp.lst('gt_labels'))
    layers = [oa.layer_targets(inp.t('s_cls')[k], inp.t('s_box')[k], gt_b, gt_l, img_shapes, prev,
                               w_cls=inp.v('w_cls')) for k in range(inp.t('s_cls').shape[0])]
    last = layers[-1]
    hs_s = inp.t('hs_s')[-1].clone().requires_grad_(True)
    hs_t = inp.t('hs_t')[-1]
    s_feats = [f.clone().requires_grad_(True) for f in inp.lst('s_feats')]
    t_feats = inp.lst('t_feats')
    fg = ol.MSELoss('sum', 1.0) if crit == 'mse' else ol.KnowledgeDistillationKLDivLoss('sum', 1.0, 2)
    loss_corr = ob.bcdd_loss(hs_s.reshape(N * Q, -1), last['labels'], hs_t.reshape(N * Q, -1),
                             tinfo['pred_keepid'], torch.cat(tinfo['pred_labels']), prev,
                             ol.MSELoss('mean', 1.0))
    id_pred = torch.nonzero(last['teacher_only_weights']).squeeze(1)
    s_mem = torch.cat([f.flatten(2) for f in inp.lst('s_feats')], 2).permute(2, 0, 1).contiguous().requires_grad_(True)
    t_mem = torch.cat([f.flatten(2) for f in t_feats], 2).permute(2, 0, 1).contiguous()
    if mode == 'decode_v1':
        loss_fg = od.decode_v1(s_feats, t_feats, hs_s, hs_t, tinfo['pred_keepid'], id_pred,
                               tinfo['pred_bboxes'], img_shapes, fg)
    elif mode == 'decode_v2':
        loss_fg = od.decode_v2(s_feats, t_feats, hs_t, tinfo['pred_keepid'], tinfo['pred_bboxes'], img_shapes, fg)
    elif mode == 'sg_out':
        loss_fg = od.sg_out(s_mem, t_mem, levels, tinfo['pred_bboxes'], inp.lst('gt_bboxes'), img_shapes, fg)
    elif mode == 'fg_only':
        loss_fg = od.fg_only(s_mem, t_mem, levels, tinfo['pred_bboxes'], img_shapes, fg)
    return dict(tinfo=tinfo, layers=layers, loss_corr=loss_corr, loss_fg=loss_fg, hs_s=hs_s,
                s_feats=s_feats, s_mem=s_mem)


def test_fg_bk_area_mask_mse_vs_reference():
    """`_fg_bk.py:534-578,611-625` run by the sibling head file's own `loss` (gen_golden.py fg_bk): loss and d / d memory."""
    inp, out = load_head_case('head_fg_bk_mse.npz')
    levels = [tuple(x) for x in inp.t('levels').tolist()]
    img_hw = tuple(inp.t('img_hw').tolist())
    N = inp.t('s_cls').shape[1]
    s_mem = torch.cat([f.flatten(2) for f in inp.lst('s_feats')], 2).permute(2, 0, 1).contiguous().requires_grad_(True)
    t_mem = torch.cat([f.flatten(2) for f in inp.lst('t_feats')], 2).permute(2, 0, 1).contiguous()
    loss = od.fg_bk(s_mem, t_mem, levels, out.lst('pred_bboxes'), [img_hw] * N, ol.MSELoss('sum', 1.0))
    close(loss, out.t('loss_fg_feature'), 1e-5, 1e-7)
    loss.backward()
    close(s_mem.grad, out.t('fg.grad_mem'), 1e-4, 1e-7)


HEAD_CASES = [('head_decode_v1_mse.npz', 'decode_v1', 'mse'), ('head_decode_v1_mse_n1.npz', 'decode_v1', 'mse'),
              ('head_decode_v1_kl.npz', 'decode_v1', 'kl'), ('head_decode_v1_kl_l70.npz', 'decode_v1', 'kl'),
              ('head_decode_v1_kl_l50.npz', 'decode_v1', 'kl'), ('head_decode_v1_kl_l60.npz', 'decode_v1', 'kl'),
              ('head_decode_v2_mse.npz', 'decode_v2', 'mse'), ('head_decode_v2_mse_n1.npz', 'decode_v2', 'mse'),
              ('head_sg_out_mse.npz', 'sg_out', 'mse'), ('head_sg_out_kl.npz', 'sg_out', 'kl'),
              ('head_fg_only_mse.npz', 'fg_only', 'mse'),
              ('head_fg_only_kl.npz', 'fg_only', 'kl'), ('head_decode_v2_kl.npz', 'decode_v2', 'kl')]


@pytest.mark.parametrize('name,mode,crit', HEAD_CASES)
def test_head_distillation_vs_reference(name, mode, crit):
    inp, out = load_head_case(name)
    r = _head_oracle(inp, out, mode, crit)
    # A1: teacher keep-ids / labels / boxes
    assert torch.equal(r['tinfo']['pred_keepid'], out.t('pred_keepid'))
    for a, b in zip(r['tinfo']['pred_labels'], out.lst('pred_labels')):
        assert torch.equal(a, b)
    for a, b in zip(r['tinfo']['pred_bboxes'], out.lst('pred_bboxes')):
        close(a, b, 0, 0)
    # H1-H3 / A2: assigned labels of every decoder layer are bit-exact
    assert torch.equal(torch.stack([l['labels'] for l in r['layers']]), out.t('labels_layers'))
    close(torch.stack([l['teacher_only_weights'] for l in r['layers']]), out.t('teacher_only_layers'), 0, 0)
    # B1-B3
    close(r['loss_corr'], out.t('loss_corr'), rtol=1e-6, atol=1e-8)
    g_corr = torch.autograd.grad(r['loss_corr'], r['hs_s'], retain_graph=True)[0]
    close(g_corr, out.t('corr.grad_hs'), rtol=1e-5, atol=1e-9)
    # A3
    close(r['loss_fg'], out.t('loss_fg_feature'), rtol=1e-6, atol=1e-8)
    if out.v('fg.backward_raises'):
        return                                   # the reference cannot back-propagate this case
    if not r['loss_fg'].requires_grad:
        assert out.v('fg.grad_hs_is_none') and out.v('fg.grad_feats_is_none') and out.v('fg.grad_mem_is_none')
        return
    grads = torch.autograd.grad(r['loss_fg'], [r['hs_s'], r['s_mem']] + r['s_feats'], allow_unused=True)
    assert (grads[0] is None) == out.v('fg.grad_hs_is_none')
    assert (grads[1] is None) == out.v('fg.grad_mem_is_none')
    assert (grads[2] is None) == out.v('fg.grad_feats_is_none')
    if grads[0] is not None:
        close(grads[0], out.t('fg.grad_hs'), rtol=1e-5, atol=1e-9)
    if grads[1] is not None:
        close(grads[1], out.t('fg.grad_mem'), rtol=1e-5, atol=1e-9)
    if grads[2] is not None:
        for a, b in zip(grads[2:], out.lst('fg.grad_feats')):
            close(a, b, rtol=1e-5, atol=1e-9)


def test_reference_mse_backward_limitation_is_recorded():
    """The reference's own backward raises for decode_v1 + MSELoss at N >= 2 (shared in-place mask);
    at N == 1 it works and is what pins the MSE gradients."""
    assert Golden('head_decode_v1_mse.npz').v('fg.backward_raises')
    assert not Golden('head_decode_v1_mse_n1.npz').v('fg.backward_raises')


def test_oracle_header_declares_itself_test_infrastructure():
    assert 'TEST INFRASTRUCTURE ONLY' in oracle.__doc__
