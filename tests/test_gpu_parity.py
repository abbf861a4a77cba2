"""Parity of the CUDA path (through the C ABI) against the oracle and against the reference's own
outputs (tests/golden).  Tolerances are the ones BASELINE.json's north_star states: rtol 1e-4 on
losses, 1e-3 on gradients (atol = 1e-6 * max|ref| for near-zero gradients), indices bit-exact."""
import pytest
import torch

import dskd_b200
from dskd_b200 import synth
from oracle import assign as oa
from oracle import bcdd as ob
from oracle import dsgfd as od
from oracle import losses as ol
from conftest import Golden, load_head_case

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
LOSS_RTOL, GRAD_RTOL = 1e-4, 1e-3


def assert_loss(got, ref):
    torch.testing.assert_close(got.detach().cpu().double(), ref.detach().double(), rtol=LOSS_RTOL, atol=1e-12)


def assert_grad(got, ref):
    ref = ref.detach()
    atol = 1e-6 * float(ref.abs().max()) + 1e-30
    torch.testing.assert_close(got.detach().cpu(), ref, rtol=GRAD_RTOL, atol=atol)


def crit_oracle(crit, reduction='sum', w=1.0, T=2):
    return ol.MSELoss(reduction, w) if crit == 'mse' else ol.KnowledgeDistillationKLDivLoss(reduction, w, T)


SMALL = dict(levels=((24, 40), (12, 20), (6, 10), (3, 5)), img_hw=(192, 320), num_query=100, k_range=(3, 9))
ODD = dict(levels=((25, 42), (13, 21), (7, 11), (4, 6)), img_hw=(200, 333), num_query=60, k_range=(2, 7))


def oracle_decode_f64(cpu, crit_mod):
    """decode_v1 with the dense tensors in float64 (box arithmetic stays fp32 like the reference)."""
    a = cpu.assignments
    shapes = [tuple(x) for x in a['img_shapes'].tolist()]
    id_pred = torch.nonzero(a['student_labels'] < len(a['prev_labels'])).squeeze(1)
    hs = cpu.hs_student.double().requires_grad_(True)
    loss = od.decode_v1([f.double() for f in cpu.student_feats], [f.double() for f in cpu.teacher_feats], hs,
                        cpu.hs_teacher.double(), a['teacher_keepid'], id_pred, a['teacher_bboxes'], shapes, crit_mod)
    loss.backward()
    return loss.detach(), hs.grad


# KL over H is a second-order quantity (both log-softmaxes sit near -log H and their first-order terms cancel),
# so the reference's own fp32 evaluation is only good to a few 1e-4 of the exact value (measured: 1.4e-4 at
# 800x1333, 3.7e-4 on the small case; DESIGN.md "Numerics").  The kernel is held to rtol 1e-4 against the
# float64 evaluation of the same formula and to KL_RTOL_VS_FP32 against the fp32 oracle.
KL_RTOL_VS_FP32 = 1e-3


def assert_kl_loss(got, cpu, crit_mod, ref32, rtol32=KL_RTOL_VS_FP32):
    ref64, _ = oracle_decode_f64(cpu, crit_mod)
    torch.testing.assert_close(got.detach().cpu().double(), ref64, rtol=LOSS_RTOL, atol=1e-12)
    torch.testing.assert_close(got.detach().cpu().double(), ref32.detach().double(), rtol=rtol32, atol=1e-12)


def oracle_decode(cpu, crit_mod, version=1, feats=None, hs=None):
    a = cpu.assignments
    shapes = [tuple(x) for x in a['img_shapes'].tolist()]
    id_pred = torch.nonzero(a['student_labels'] < len(a['prev_labels'])).squeeze(1)
    if version == 1:
        return od.decode_v1(feats, cpu.teacher_feats, hs, cpu.hs_teacher, a['teacher_keepid'], id_pred,
                            a['teacher_bboxes'], shapes, crit_mod)
    return od.decode_v2(feats, cpu.teacher_feats, cpu.hs_teacher, a['teacher_keepid'], a['teacher_bboxes'],
                        shapes, crit_mod)


@pytest.mark.parametrize('cfg', [SMALL, ODD], ids=['vec4', 'scalar'])
@pytest.mark.parametrize('channels', [64, 256])
@pytest.mark.parametrize('reduction', ['sum', 'mean'])
def test_dsgfd_decode_v1_mse_neck(cfg, channels, reduction):
    cpu = synth.make_distill_inputs(num_images=3, num_prev=40, seed=3, channels=channels, **cfg)
    gpu = cpu.to(DEV)
    mod = dskd_b200.DSGFeatureDistillLoss(loss_weight=0.7, reduction=reduction, criterion='mse')
    feats, hs = gpu.clone_student()
    loss = mod(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    o_feats, o_hs = cpu.clone_student()
    ref = oracle_decode(cpu, crit_oracle('mse', reduction, 0.7), 1, o_feats, o_hs)
    ref.backward()
    assert_loss(loss, ref)
    assert_grad(hs.grad, o_hs.grad)
    for g, r in zip(feats, o_feats):
        assert_grad(g.grad, r.grad)


@pytest.mark.parametrize('channels', [64, 256])
def test_dsgfd_decode_v1_mse_memory_layout(channels):
    """Same numbers through the [S,N,C] encoder-memory layout (head_il.py:866-880 views)."""
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=5, channels=channels, **ODD)
    gpu = cpu.to(DEV)
    s_mem, t_mem = gpu.memory()
    s_mem.requires_grad_(True)
    _, hs = gpu.clone_student()
    mod = dskd_b200.DSGFeatureDistillLoss(criterion='mse', feature_source='memory')
    loss = mod((s_mem, gpu.spatial_shapes), (t_mem, gpu.spatial_shapes), (hs, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    o_feats, o_hs = cpu.clone_student()
    ref = oracle_decode(cpu, crit_oracle('mse'), 1, o_feats, o_hs)
    ref.backward()
    assert_loss(loss, ref)
    assert_grad(hs.grad, o_hs.grad)
    ref_mem_grad = torch.cat([f.grad.flatten(2) for f in o_feats], 2).permute(2, 0, 1)
    assert_grad(s_mem.grad, ref_mem_grad.contiguous())


@pytest.mark.parametrize('cfg', [SMALL, ODD], ids=['vec4', 'scalar'])
@pytest.mark.parametrize('reduction', ['sum', 'mean'])
def test_dsgfd_decode_v1_kl(cfg, reduction):
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=4, channels=64, **cfg)
    gpu = cpu.to(DEV)
    mod = dskd_b200.DSGFeatureDistillLoss(loss_weight=1.3, reduction=reduction, criterion='kl', T=2.0)
    feats, hs = gpu.clone_student()
    loss = mod(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    o_feats, o_hs = cpu.clone_student()
    ref = oracle_decode(cpu, crit_oracle('kl', reduction, 1.3, 2), 1, o_feats, o_hs)
    ref.backward()
    assert_kl_loss(loss, cpu, crit_oracle('kl', reduction, 1.3, 2), ref)
    assert_grad(hs.grad, o_hs.grad)
    assert all(f.grad is None for f in feats) and all(f.grad is None for f in o_feats)


# Tall and narrow levels: 15 to 105 row blocks per column tile, column tiles narrower than a warp.
KL_TALL = {
    'parts3': dict(levels=((60, 40), (31, 20), (15, 10), (8, 5)), img_hw=(480, 320)),
    'parts6': dict(levels=((130, 36), (65, 18), (33, 9), (17, 5)), img_hw=(1040, 288)),
    'parts10': dict(levels=((230, 33), (115, 17), (58, 9), (29, 5)), img_hw=(1840, 264)),
    'strips': dict(levels=((420, 9), (210, 5), (105, 3), (53, 2)), img_hw=(3360, 72)),
    # 177 KB of owner table: the 16-warp small-batch layout does not fit beside it, the launcher falls back to 8 warps
    'very_tall': dict(levels=((1300, 3), (650, 2), (325, 1), (163, 1)), img_hw=(10400, 24)),
}


@pytest.mark.parametrize('shape', list(KL_TALL))
@pytest.mark.parametrize('T', [2.0, 3.0], ids=['T2', 'T3'])
def test_dsgfd_kl_tall_levels_and_temperatures(shape, T):
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=21, channels=40, num_query=60, k_range=(3, 8),
                                    **KL_TALL[shape])
    gpu = cpu.to(DEV)
    mod = dskd_b200.DSGFeatureDistillLoss(criterion='kl', T=T)
    feats, hs = gpu.clone_student()
    loss = mod(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    o_feats, o_hs = cpu.clone_student()
    ref = oracle_decode(cpu, crit_oracle('kl', 'sum', 1.0, T), 1, o_feats, o_hs)
    ref.backward()
    # the fp32 evaluation of the reference formula loses accuracy with the column height (softmax over H rows): at
    # H = 1300 it is 1.2e-3 off the float64 value, which the kernel matches to 1e-4
    assert_kl_loss(loss, cpu, crit_oracle('kl', 'sum', 1.0, T), ref, rtol32=5e-3 if shape == 'very_tall' else KL_RTOL_VS_FP32)
    assert_grad(hs.grad, o_hs.grad)


def test_dsgfd_kl_mask_values_underflowing_to_zero():
    """|hs_T - hs_S| spread over hundreds: most of softmax_c underflows to exactly 0 in fp32 (head_il.py:705), the
    logits of such a channel are 0 inside the box and its gradient comes from the raw teacher feature alone."""
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=9, channels=64, **ODD)
    cpu.hs_student.mul_(150.0)
    cpu.hs_teacher.mul_(150.0)
    gpu = cpu.to(DEV)
    mod = dskd_b200.DSGFeatureDistillLoss(criterion='kl')
    feats, hs = gpu.clone_student()
    loss = mod(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    o_feats, o_hs = cpu.clone_student()
    ref = oracle_decode(cpu, crit_oracle('kl'), 1, o_feats, o_hs)
    ref.backward()
    assert_kl_loss(loss, cpu, crit_oracle('kl'), ref)
    # with a nearly one-hot mask the softmax backward cancels to the last bits, in the reference's fp32 as in the
    # kernel's: hold the kernel to the usual tolerance against the float64 evaluation, or to the error the fp32
    # reference itself makes against it, whichever is larger
    _, g64 = oracle_decode_f64(cpu, crit_oracle('kl'))
    got = hs.grad.detach().cpu().double()
    err_ref = (o_hs.grad.double() - g64).abs().max()
    tol = GRAD_RTOL * g64.abs() + 1e-6 * g64.abs().max() + 4 * err_ref
    assert ((got - g64).abs() <= tol).all(), (float((got - g64).abs().max()), float(err_ref), float(g64.abs().max()))
    a = cpu.assignments
    id_pred = torch.nonzero(a['student_labels'] < len(a['prev_labels'])).squeeze(1)
    rows = torch.softmax((cpu.hs_teacher.reshape(-1, 64)[a['teacher_keepid']]
                          - cpu.hs_student.reshape(-1, 64)[id_pred]).abs(), 1)
    assert (rows == 0).any()                                   # the case really contains zero mask values


@pytest.mark.parametrize('crit', ['mse', 'kl'])
def test_dsgfd_decode_v2(crit):
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=6, channels=64, **SMALL)
    gpu = cpu.to(DEV)
    mod = dskd_b200.DSGFeatureDistillLoss(criterion=crit, mask_mode='decode_v2')
    feats, hs = gpu.clone_student()
    loss = mod(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments)
    o_feats, o_hs = cpu.clone_student()
    ref = oracle_decode(cpu, crit_oracle(crit), 2, o_feats, o_hs)
    assert_loss(loss, ref)
    if crit == 'mse':
        loss.backward()
        ref.backward()
        for g, r in zip(feats, o_feats):
            assert_grad(g.grad, r.grad)
        assert hs.grad is None or float(hs.grad.abs().max()) == 0.0
    else:
        assert not ref.requires_grad                      # constant mask, teacher pred, detached target


@pytest.mark.parametrize('mode', ['sg_out', 'fg_only', 'fg_bk'])
def test_dsgfd_cell_masks_mse_memory(mode):
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=8, channels=64, **ODD)
    gpu = cpu.to(DEV)
    a = cpu.assignments
    shapes = [tuple(x) for x in a['img_shapes'].tolist()]
    s_mem, t_mem = gpu.memory()
    s_mem.requires_grad_(True)
    mod = dskd_b200.DSGFeatureDistillLoss(criterion='mse', mask_mode=mode, feature_source='memory', loss_weight=2.0)
    loss = mod((s_mem, gpu.spatial_shapes), (t_mem, gpu.spatial_shapes), (gpu.hs_student, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    o_s, o_t = cpu.memory()
    o_s.requires_grad_(True)
    crit = ol.MSELoss('sum', 2.0)
    if mode == 'sg_out':
        ref = od.sg_out(o_s, o_t, cpu.levels, a['teacher_bboxes'], a['gt_bboxes'], shapes, crit)
    elif mode == 'fg_only':
        ref = od.fg_only(o_s, o_t, cpu.levels, a['teacher_bboxes'], shapes, crit)
    else:
        ref = od.fg_bk(o_s, o_t, cpu.levels, a['teacher_bboxes'], shapes, crit)
    ref.backward()
    assert_loss(loss, ref)
    assert_grad(s_mem.grad, o_s.grad)


@pytest.mark.parametrize('mode', ['sg_out', 'fg_only'])
def test_dsgfd_cell_masks_neck_mse_and_kl(mode):
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=9, channels=64, **SMALL)
    gpu = cpu.to(DEV)
    a = cpu.assignments
    shapes = [tuple(x) for x in a['img_shapes'].tolist()]
    o_s, o_t = cpu.memory()
    for crit in ('mse', 'kl'):
        mod = dskd_b200.DSGFeatureDistillLoss(criterion=crit, mask_mode=mode, feature_source='neck')
        feats, _ = gpu.clone_student()
        loss = mod(feats, gpu.teacher_feats, (gpu.hs_student, gpu.hs_teacher), gpu.assignments)
        c = crit_oracle(crit)
        ref = od.sg_out(o_s, o_t, cpu.levels, a['teacher_bboxes'], a['gt_bboxes'], shapes, c) if mode == 'sg_out' \
            else od.fg_only(o_s, o_t, cpu.levels, a['teacher_bboxes'], shapes, c)
        assert_loss(loss, ref)


@pytest.mark.parametrize('mode', ['sg_out', 'fg_only'])
@pytest.mark.parametrize('cfg,channels', [(SMALL, 64), (ODD, 256), (ODD, 136)], ids=['c64', 'c256', 'c136'])
def test_dsgfd_cell_masks_kl_on_memory_layout(mode, cfg, channels):
    """KL over H straight on encoder memory [S,N,C] (head_il.py:865-880,916-923): no [N,C,H,W] copy, same numbers."""
    cpu = synth.make_distill_inputs(num_images=3, num_prev=40, seed=12, channels=channels, **cfg)
    gpu = cpu.to(DEV)
    a = cpu.assignments
    shapes = [tuple(x) for x in a['img_shapes'].tolist()]
    s_mem, t_mem = gpu.memory()
    for T in (2.0, 3.0):
        mod = dskd_b200.DSGFeatureDistillLoss(criterion='kl', T=T, mask_mode=mode, feature_source='memory', loss_weight=0.8)
        loss = mod((s_mem, gpu.spatial_shapes), (t_mem, gpu.spatial_shapes), (gpu.hs_student, gpu.hs_teacher), gpu.assignments)
        o_s, o_t = cpu.memory()
        c64 = ol.KnowledgeDistillationKLDivLoss('sum', 0.8, T)
        ref64 = (od.sg_out(o_s.double(), o_t.double(), cpu.levels, a['teacher_bboxes'], a['gt_bboxes'], shapes, c64)
                 if mode == 'sg_out' else od.fg_only(o_s.double(), o_t.double(), cpu.levels, a['teacher_bboxes'], shapes, c64))
        torch.testing.assert_close(loss.detach().cpu().double(), ref64, rtol=LOSS_RTOL, atol=1e-12)
        assert not loss.requires_grad                      # constant mask, teacher pred, detached target
        neck = dskd_b200.DSGFeatureDistillLoss(criterion='kl', T=T, mask_mode=mode, feature_source='neck', loss_weight=0.8)
        same = neck(gpu.student_feats, gpu.teacher_feats, (gpu.hs_student, gpu.hs_teacher), gpu.assignments)
        torch.testing.assert_close(loss, same, rtol=2e-5, atol=0)


def test_dsgfd_kl_rejects_box_masks_on_memory():
    with pytest.raises(NotImplementedError):
        dskd_b200.DSGFeatureDistillLoss(criterion='kl', mask_mode='decode_v1', feature_source='memory')


@pytest.mark.parametrize('scale', [300.0, 3000.0])
def test_dsgfd_kl_logits_beyond_the_unshifted_range(scale):
    """Features so large that e^(feature * mask / T) leaves fp32 without the column maximum subtracted: the streaming
    kernel hands those (tile, channel) pairs to the redo launch (exact maxima, like kd_loss.py:28-34 via torch softmax)."""
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=14, channels=16, **ODD)
    for f in cpu.student_feats + cpu.teacher_feats:
        f.mul_(scale)
    gpu = cpu.to(DEV)
    mod = dskd_b200.DSGFeatureDistillLoss(criterion='kl')
    feats, hs = gpu.clone_student()
    loss = mod(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    o_feats, o_hs = cpu.clone_student()
    ref = oracle_decode(cpu, crit_oracle('kl'), 1, o_feats, o_hs)
    ref.backward()
    assert torch.isfinite(loss).all() and torch.isfinite(hs.grad).all()
    assert_kl_loss(loss, cpu, crit_oracle('kl'), ref)
    _, g64 = oracle_decode_f64(cpu, crit_oracle('kl'))
    err_ref = (o_hs.grad.double() - g64).abs().max()
    got = hs.grad.detach().cpu().double()
    tol = GRAD_RTOL * g64.abs() + 1e-6 * g64.abs().max() + 4 * err_ref
    assert ((got - g64).abs() <= tol).all(), (float((got - g64).abs().max()), float(err_ref), float(g64.abs().max()))
    # the same through the per-cell masks, both layouts (forward only)
    a = cpu.assignments
    shapes = [tuple(x) for x in a['img_shapes'].tolist()]
    o_s, o_t = cpu.memory()
    ref64 = od.fg_only(o_s.double(), o_t.double(), cpu.levels, a['teacher_bboxes'], shapes, crit_oracle('kl'))
    s_mem, t_mem = gpu.memory()
    for src, sf, tf in (('neck', gpu.student_feats, gpu.teacher_feats),
                        ('memory', (s_mem, gpu.spatial_shapes), (t_mem, gpu.spatial_shapes))):
        cell = dskd_b200.DSGFeatureDistillLoss(criterion='kl', mask_mode='fg_only', feature_source=src)
        got = cell(sf, tf, (gpu.hs_student, gpu.hs_teacher), gpu.assignments)
        torch.testing.assert_close(got.detach().cpu().double(), ref64, rtol=LOSS_RTOL, atol=1e-12)


@pytest.mark.parametrize('layout', ['0,1', '8,1', '8,2', '16,2', '16,4', '4,4', '2,2', '1,1'])
@pytest.mark.parametrize('shape', ['small', 'odd', 'parts6'])
def test_dsgfd_kl_cta_layouts(shape, layout, monkeypatch):
    """Every way the streaming kernel can lay a column tile out over a CTA (DSKD_KL_TUNE: warps, row parts per column):
    the production choices are 8 warps x 1 part (large batches) and 8 warps x 2 parts on half-size channel chunks (up to
    4 images); a run of rows that crosses a part boundary is cut into two records there."""
    cfg = dict(small=SMALL, odd=ODD, parts6=dict(num_query=60, k_range=(3, 8), **KL_TALL['parts6']))[shape]
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=21, channels=64, **cfg)
    gpu = cpu.to(DEV)
    chunk = 16 if layout != '1,1' else 2
    monkeypatch.setenv('DSKD_KL_TUNE', f'2,4,{chunk},192,0,{layout}')
    mod = dskd_b200.DSGFeatureDistillLoss(criterion='kl')
    feats, hs = gpu.clone_student()
    loss = mod(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    # forward-only kernels (5-row blocks) split their rows the same way
    monkeypatch.setenv('DSKD_KL_TUNE', f'2,5,{chunk},192,0,{layout}')
    v2 = dskd_b200.DSGFeatureDistillLoss(criterion='kl', mask_mode='decode_v2')
    loss_v2 = v2(gpu.student_feats, gpu.teacher_feats, (gpu.hs_student, gpu.hs_teacher), gpu.assignments)
    torch.cuda.synchronize()
    monkeypatch.delenv('DSKD_KL_TUNE')
    loss_v2_default = v2(gpu.student_feats, gpu.teacher_feats, (gpu.hs_student, gpu.hs_teacher), gpu.assignments)
    o_feats, o_hs = cpu.clone_student()
    ref = oracle_decode(cpu, crit_oracle('kl'), 1, o_feats, o_hs)
    ref.backward()
    assert_kl_loss(loss, cpu, crit_oracle('kl'), ref)
    assert_grad(hs.grad, o_hs.grad)
    torch.testing.assert_close(loss_v2, loss_v2_default, rtol=2e-4, atol=0)   # both within 1e-4 of the exact value


@pytest.mark.parametrize('nc', [2, 1], ids=['pairs', 'single'])
@pytest.mark.parametrize('parts', [1, 2], ids=['whole_columns', 'row_parts'])
@pytest.mark.parametrize('pool', [64, 12, 2], ids=['shared_pools', 'one_warp', 'redo'])
def test_dsgfd_kl_many_runs_per_tile(pool, parts, nc, monkeypatch):
    """Crowded images: more runs of rows per column tile than one warp's record pool holds.  With a small pool
    (DSKD_KL_TUNE) the CTA first runs fewer warps with several pools each and finally leaves the tile to the redo launch;
    all three must give the reference's numbers."""
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=15, channels=32, num_query=100, k_range=(30, 40), **{
        k: v for k, v in SMALL.items() if k not in ('k_range', 'num_query')})
    gpu = cpu.to(DEV)
    monkeypatch.setenv('DSKD_KL_TUNE', f'{nc},4,16,{pool},0,0,{parts}')   # channels per pass, ..., pool records, ..., row parts
    mod = dskd_b200.DSGFeatureDistillLoss(criterion='kl')
    feats, hs = gpu.clone_student()
    loss = mod(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    torch.cuda.synchronize()
    monkeypatch.delenv('DSKD_KL_TUNE')
    o_feats, o_hs = cpu.clone_student()
    ref = oracle_decode(cpu, crit_oracle('kl'), 1, o_feats, o_hs)
    ref.backward()
    assert_kl_loss(loss, cpu, crit_oracle('kl'), ref)
    assert_grad(hs.grad, o_hs.grad)


@pytest.mark.parametrize('crit', ['mse', 'kl'])
def test_dsgfd_fewer_matched_queries_than_teacher_boxes_is_loud(crit):
    """The reference raises IndexError at head_il.py:705 when fewer student queries carry a previous label than there are
    teacher detections.  validate=True raises the same error (one host sync); validate=False (the graph-capturable default)
    must not return a plausible number: loss and the embedding gradient are NaN."""
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=16, channels=32, **SMALL)
    a = dict(cpu.assignments)
    labels = a['student_labels'].clone()
    prev_q = torch.nonzero(labels < 40).squeeze(1)
    labels[prev_q[-3:]] = 80                       # three matched queries lose their previous label
    a['student_labels'] = labels
    cpu.assignments = a
    gpu = cpu.to(DEV)
    feats, hs = gpu.clone_student()
    with pytest.raises(IndexError):
        dskd_b200.DSGFeatureDistillLoss(criterion=crit, validate=True)(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments)
    feats, hs = gpu.clone_student()
    loss = dskd_b200.DSGFeatureDistillLoss(criterion=crit)(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    assert torch.isnan(loss)
    assert torch.isnan(hs.grad).any()


def test_dsgfd_cell_masks_need_no_queries():
    """sg_out / fg_only never touch the decoder embeddings (head_il.py:860-925,1082-1129): queries may be None."""
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=9, channels=64, **SMALL)
    gpu = cpu.to(DEV)
    for crit in ('mse', 'kl'):
        mod = dskd_b200.DSGFeatureDistillLoss(criterion=crit, mask_mode='fg_only')
        with_q = mod(gpu.student_feats, gpu.teacher_feats, (gpu.hs_student, gpu.hs_teacher), gpu.assignments)
        without = mod(gpu.student_feats, gpu.teacher_feats, None, gpu.assignments)
        half = mod(gpu.student_feats, gpu.teacher_feats, (None, gpu.hs_teacher), gpu.assignments)
        assert float(with_q) == float(without) == float(half)
    v2 = dskd_b200.DSGFeatureDistillLoss(criterion='mse', mask_mode='decode_v2')
    a = v2(gpu.student_feats, gpu.teacher_feats, (gpu.hs_student, gpu.hs_teacher), gpu.assignments)
    b = v2(gpu.student_feats, gpu.teacher_feats, (None, gpu.hs_teacher), gpu.assignments)
    assert float(a) == float(b)
    with pytest.raises(dskd_b200._lib.DskdError):
        dskd_b200.DSGFeatureDistillLoss(criterion='mse')(gpu.student_feats, gpu.teacher_feats, None, gpu.assignments)


@pytest.mark.parametrize('num_prev', [40, 50, 60, 70])
@pytest.mark.parametrize('reduction', ['mean', 'sum'])
def test_bcdd_vs_oracle(num_prev, reduction):
    cpu = synth.make_distill_inputs(num_images=4, num_prev=num_prev, seed=10, **SMALL)
    gpu = cpu.to(DEV)
    a = cpu.assignments
    mod = dskd_b200.BetweenClassDistanceLoss(loss_weight=1.5, reduction=reduction)
    _, hs = gpu.clone_student()
    loss = mod(None, None, (hs, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    _, o_hs = cpu.clone_student()
    C = o_hs.shape[-1]
    corr_t, corr_s = ob.prototypes(o_hs.reshape(-1, C), a['student_labels'], cpu.hs_teacher.reshape(-1, C),
                                   a['teacher_keepid'], a['teacher_labels'], a['prev_labels'])
    # prototype sums run in the reference's loop order: bit-exact
    assert torch.equal(mod.last_prototypes[0].cpu(), corr_t.detach())
    assert torch.equal(mod.last_prototypes[1].cpu(), corr_s.detach())
    ref = ob.correlation_loss(corr_t, corr_s, num_prev, ol.MSELoss(reduction, 1.5))
    ref.backward()
    assert_loss(loss, ref)
    assert_grad(hs.grad, o_hs.grad)
    d_t, d_s = ob.distance_matrices(*[c.detach().clone() for c in ob.prototypes(
        o_hs.detach().reshape(-1, C), a['student_labels'], cpu.hs_teacher.reshape(-1, C), a['teacher_keepid'],
        a['teacher_labels'], a['prev_labels'])], num_prev)
    torch.testing.assert_close(mod.last_distances[0].cpu(), d_t, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(mod.last_distances[1].cpu(), d_s, rtol=1e-5, atol=1e-6)


def test_bcdd_nan_edge_matches_reference_semantics():
    """A class with teacher hits but no student hits divides 0/0 in the reference (head_il.py:1205)."""
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=11, **SMALL)
    a = dict(cpu.assignments)
    lab = a['student_labels'].clone()
    victim = int(a['teacher_labels'][0])
    lab[lab == victim] = 80
    a['student_labels'] = lab
    C = cpu.hs_student.shape[-1]
    ref = ob.bcdd_loss(cpu.hs_student.reshape(-1, C), lab, cpu.hs_teacher.reshape(-1, C), a['teacher_keepid'],
                       a['teacher_labels'], a['prev_labels'], ol.MSELoss('mean', 1.0))
    gpu = cpu.to(DEV)
    ga = dict(gpu.assignments)
    ga['student_labels'] = lab.to(DEV)
    loss = dskd_b200.BetweenClassDistanceLoss()(None, None, (gpu.hs_student, gpu.hs_teacher), ga)
    assert torch.isnan(ref) and torch.isnan(loss.cpu())


# ------------------------------------------------------------------ against the reference's own outputs
def _gpu_assignments_from_golden(inp, out):
    N, Q = inp.t('s_cls').shape[1:3]
    img_hw = tuple(inp.t('img_hw').tolist())
    L = inp.v('L')
    return dict(student_labels=out.t('labels_layers')[-1].to(DEV), teacher_keepid=out.t('pred_keepid').to(DEV),
                teacher_labels=torch.cat(out.lst('pred_labels')).to(DEV),
                teacher_bboxes=[b.to(DEV) for b in out.lst('pred_bboxes')],
                gt_bboxes=[b.to(DEV) for b in inp.lst('gt_bboxes')],
                img_shapes=[img_hw] * N, prev_labels=list(range(L)), num_classes=80), N, Q


GOLDEN_CASES = [('head_decode_v1_mse.npz', 'decode_v1', 'mse', 'neck'),
                ('head_decode_v1_mse_n1.npz', 'decode_v1', 'mse', 'neck'),
                ('head_decode_v1_kl.npz', 'decode_v1', 'kl', 'neck'),
                ('head_decode_v1_kl_l70.npz', 'decode_v1', 'kl', 'neck'),
                ('head_decode_v1_kl_l50.npz', 'decode_v1', 'kl', 'neck'),      # 50+30 and 60+20 settings (BASELINE configs 4)
                ('head_decode_v1_kl_l60.npz', 'decode_v1', 'kl', 'neck'),
                ('head_decode_v2_mse_n1.npz', 'decode_v2', 'mse', 'neck'),
                ('head_sg_out_mse.npz', 'sg_out', 'mse', 'memory'),
                ('head_sg_out_kl.npz', 'sg_out', 'kl', 'neck'),
                ('head_sg_out_kl.npz', 'sg_out', 'kl', 'memory'),     # the layout the reference views (head_il.py:879-880)
                ('head_fg_only_mse.npz', 'fg_only', 'mse', 'memory'),
                # the shipped KL criterion through the sibling masks (no gradient: constant mask, detached target)
                ('head_fg_only_kl.npz', 'fg_only', 'kl', 'neck'),
                ('head_fg_only_kl.npz', 'fg_only', 'kl', 'memory'),
                ('head_decode_v2_kl.npz', 'decode_v2', 'kl', 'neck')]


@pytest.mark.parametrize('name,mode,crit,source', GOLDEN_CASES)
def test_cuda_path_vs_reference_outputs(name, mode, crit, source):
    inp, out = load_head_case(name)
    a, N, Q = _gpu_assignments_from_golden(inp, out)
    hs_s = inp.t('hs_s')[-1].to(DEV).requires_grad_(True)
    hs_t = inp.t('hs_t')[-1].to(DEV)
    s_feats = [f.to(DEV).requires_grad_(True) for f in inp.lst('s_feats')]
    t_feats = [f.to(DEV) for f in inp.lst('t_feats')]
    # BCDD
    corr = dskd_b200.BetweenClassDistanceLoss(reduction='mean')(None, None, (hs_s, hs_t), a)
    assert_loss(corr, out.t('loss_corr'))
    g, = torch.autograd.grad(corr, hs_s)
    assert_grad(g, out.t('corr.grad_hs'))
    # DSG-FD
    mod = dskd_b200.DSGFeatureDistillLoss(criterion=crit, mask_mode=mode, feature_source=source, validate=True)
    if source == 'neck':
        loss = mod(s_feats, t_feats, (hs_s, hs_t), a)
        leaves = s_feats
    else:
        shapes = inp.t('levels')
        s_mem = torch.cat([f.detach().flatten(2) for f in s_feats], 2).permute(2, 0, 1).contiguous().requires_grad_(True)
        t_mem = torch.cat([f.flatten(2) for f in t_feats], 2).permute(2, 0, 1).contiguous()
        loss = mod((s_mem, shapes), (t_mem, shapes), (hs_s, hs_t), a)
        leaves = [s_mem]
    assert_loss(loss, out.t('loss_fg_feature'))
    if out.v('fg.backward_raises') or not loss.requires_grad:
        return
    grads = torch.autograd.grad(loss, [hs_s] + leaves, allow_unused=True)
    if not out.v('fg.grad_hs_is_none'):
        assert_grad(grads[0], out.t('fg.grad_hs'))
    if source == 'memory' and not out.v('fg.grad_mem_is_none'):
        assert_grad(grads[1], out.t('fg.grad_mem'))
    if source == 'neck' and not out.v('fg.grad_feats_is_none'):
        for got, ref in zip(grads[1:], out.lst('fg.grad_feats')):
            assert_grad(got, ref)


def test_fg_bk_vs_reference_outputs():
    """The sibling head file's area-mask MSE on encoder memory (_fg_bk.py:534-578,611-625), outputs of that head's own
    `loss` (tests/golden/gen_golden.py fg_bk): loss and the gradient of the student's memory."""
    inp, out = load_head_case('head_fg_bk_mse.npz')
    N = inp.t('s_cls').shape[1]
    img_hw = tuple(inp.t('img_hw').tolist())
    a = dict(teacher_bboxes=[b.to(DEV) for b in out.lst('pred_bboxes')], img_shapes=[img_hw] * N)
    shapes = inp.t('levels')
    s_mem = torch.cat([f.flatten(2) for f in inp.lst('s_feats')], 2).permute(2, 0, 1).contiguous().to(DEV).requires_grad_(True)
    t_mem = torch.cat([f.flatten(2) for f in inp.lst('t_feats')], 2).permute(2, 0, 1).contiguous().to(DEV)
    mod = dskd_b200.DSGFeatureDistillLoss(criterion='mse', mask_mode='fg_bk', feature_source='memory')
    loss = mod((s_mem, shapes), (t_mem, shapes), None, a)
    assert_loss(loss, out.t('loss_fg_feature'))
    loss.backward()
    assert_grad(s_mem.grad, out.t('fg.grad_mem'))


def test_assignment_vs_reference_outputs():
    g = Golden('assign.npz')
    asg = dskd_b200.GFLHungarianAssigner()
    for c in range(g.v('num_cases')):
        p = f'case{c}.'
        asg.w_cls = g.v(p + 'w_cls')
        res = asg.assign(g.t(p + 'cxcywh').to(DEV), g.t(p + 'cls').to(DEV), g.t(p + 'gt').to(DEV),
                         g.t(p + 'lab').to(DEV), None, dict(img_shape=(800, 1333, 3)))
        assert torch.equal(res.gt_inds.cpu(), g.t(p + 'gt_inds')), p
        assert torch.equal(res.labels.cpu(), g.t(p + 'labels')), p
        if g.t(p + 'lab').numel():
            cost, cols, _ = asg.cost_matrices(g.t(p + 'cls').to(DEV)[None], g.t(p + 'cxcywh').to(DEV)[None],
                                              [g.t(p + 'gt').to(DEV)], [g.t(p + 'lab').to(DEV)], [(800, 1333)],
                                              decoded=True)
            ref = g.t(p + 'cls_cost') + g.t(p + 'reg_cost') + g.t(p + 'iou_cost')
            torch.testing.assert_close(cost[0, :, :cols[0]].cpu(), ref, rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize('name', ['head_decode_v1_mse.npz', 'head_decode_v1_kl_l70.npz', 'head_decode_v1_kl_l50.npz',
                                  'head_decode_v1_kl_l60.npz'])
def test_batched_assignment_all_layers_vs_reference(name):
    inp, out = load_head_case(name)
    N, Q = inp.t('s_cls').shape[1:3]
    img_hw = tuple(inp.t('img_hw').tolist())
    gt_b = [torch.cat([t, g]) for t, g in zip(out.lst('pred_bboxes'), inp.lst('gt_bboxes'))]
    gt_l = [torch.cat([t, g]) for t, g in zip(out.lst('pred_labels'), inp.lst('gt_labels'))]
    asg = dskd_b200.GFLHungarianAssigner(cls_cost=dict(type='QualityFocalLossCost', weight=inp.v('w_cls')))
    res = asg.assign_batch(inp.t('s_cls').to(DEV), inp.t('s_box').to(DEV), [b.to(DEV) for b in gt_b],
                           [l.to(DEV) for l in gt_l], [img_hw] * N, prev_labels=list(range(inp.v('L'))))
    layers = inp.t('s_cls').shape[0]
    assert torch.equal(res['labels'].cpu().view(layers, N * Q), out.t('labels_layers'))
    assert torch.equal(res['teacher_only_weights'].cpu().view(layers, N * Q), out.t('teacher_only_layers'))


def test_batched_assignment_vs_oracle_coco_shape():
    ai = synth.make_assign_inputs(num_images=4, seed=21)
    asg = dskd_b200.GFLHungarianAssigner()
    res = asg.assign_batch(ai.cls_logits.to(DEV), ai.box_pred.to(DEV), [g.to(DEV) for g in ai.gt_bboxes],
                           [l.to(DEV) for l in ai.gt_labels], ai.img_shapes, prev_labels=list(range(40)))
    shapes = [tuple(x) for x in ai.img_shapes.tolist()]
    ref = [oa.layer_targets(ai.cls_logits[k], ai.box_pred[k], ai.gt_bboxes, ai.gt_labels, shapes, list(range(40)))
           for k in range(6)]
    for key in ('labels', 'teacher_only_weights', 'bbox_weights'):
        assert torch.equal(res[key].cpu(), torch.cat([r[key] for r in ref])), key
    assert torch.equal(res['assigned_gt_inds'].cpu(), torch.cat([r['gt_inds'] for r in ref]))
    torch.testing.assert_close(res['bbox_targets'].cpu(), torch.cat([r['bbox_targets'] for r in ref]), rtol=0, atol=0)


# ------------------------------------------------------------------ registry loss modules
def test_registry_loss_modules_vs_reference_outputs():
    g = Golden('losses.npz')
    pred, tgt, w = g.t('pred').to(DEV), g.t('target').to(DEV), g.t('weight').to(DEV)

    def close(a, b, rtol=1e-5):
        torch.testing.assert_close(a.detach().cpu(), b, rtol=rtol, atol=1e-6)
    kd_rtol = 1e-4      # KL rows are second-order quantities: the fp32 reference itself carries ~2e-5 noise
    for red in ('none', 'mean', 'sum'):
        for lw in (1.0, 0.37):
            close(dskd_b200.MSELoss(red, lw)(pred, tgt), g.t(f'mse.{red}.{lw}'))
            close(dskd_b200.MSELoss(red, lw)(pred, tgt, weight=w), g.t(f'mse.{red}.{lw}.w'))
        for T in (1, 2, 10):
            close(dskd_b200.KnowledgeDistillationKLDivLoss(red, 1.5, T)(pred, tgt), g.t(f'kd.{red}.T{T}'), kd_rtol)
    close(dskd_b200.MSELoss('mean')(pred, tgt, weight=w, avg_factor=7.0), g.t('mse.mean.avg7'))
    close(dskd_b200.KnowledgeDistillationKLDivLoss('sum', 1.0, 2)(pred, tgt, weight=g.t('kd.weight').to(DEV)), g.t('kd.sum.T2.w'), kd_rtol)
    close(dskd_b200.KnowledgeDistillationKLDivLoss('mean', 1.0, 2)(pred, tgt, avg_factor=3.0), g.t('kd.mean.T2.avg3'), kd_rtol)
    close(dskd_b200.KnowledgeDistillationKLDivLoss('mean', 1.0, 10)(g.t('pred2').to(DEV), g.t('target2').to(DEV)), g.t('kd2.mean.T10'), kd_rtol)
    p, t = pred.clone().requires_grad_(True), tgt.clone().requires_grad_(True)
    dskd_b200.MSELoss('sum', 0.37)(p, t, weight=w).backward()
    close(p.grad, g.t('mse.grad_pred'))
    close(t.grad, g.t('mse.grad_target'))
    p, t = pred.clone().requires_grad_(True), tgt.clone().requires_grad_(True)
    dskd_b200.KnowledgeDistillationKLDivLoss('sum', 1.5, 2)(p, t).backward()
    close(p.grad, g.t('kd.grad_pred'))
    assert t.grad is None


def test_registry_loss_module_known_answers():
    # the reference's tests/test_metrics/test_losses.py:82-109 and tests/test_models/test_loss.py:28-88
    with pytest.raises(AssertionError):
        dskd_b200.build_loss(dict(type='KnowledgeDistillationKLDivLoss', loss_weight=1.0, T=0.5))
    kd = dskd_b200.build_loss(dict(type='KnowledgeDistillationKLDivLoss', loss_weight=1.0, T=1))
    with pytest.raises(AssertionError):
        kd(torch.Tensor([[5, -5, 0]]).to(DEV), torch.Tensor([[1, 0]]).to(DEV))
    z = kd(torch.Tensor([[1, 2, 0]]).to(DEV), torch.Tensor([[1, 2, 0]]).to(DEV))
    assert torch.allclose(z.cpu(), torch.tensor(0.0), atol=1e-7)
    pred, tgt = torch.rand(1, 4, device=DEV), torch.rand(1, 4, device=DEV)
    mse = dskd_b200.build_loss(dict(type='MSELoss'))
    with pytest.raises(ValueError):
        mse(pred, tgt, avg_factor=10, reduction_override='sum')
    with pytest.raises(AssertionError):
        mse(pred, tgt, reduction_override=True)
    mse(torch.rand(0, 4, device=DEV), torch.rand(0, 4, device=DEV))
    assert isinstance(mse(pred, tgt, avg_factor=10, reduction_override='mean'), torch.Tensor)


# ------------------------------------------------------------------ edge cases & autograd contract
def test_images_without_teacher_boxes_and_empty_batch_of_boxes():
    cpu = synth.make_distill_inputs(num_images=3, num_prev=40, seed=12, channels=64, **SMALL)
    a = cpu.assignments
    # drop every teacher detection of image 1 (ragged), keep the student labels consistent
    n0, n1 = len(a['teacher_bboxes'][0]), len(a['teacher_bboxes'][1])
    keep = torch.ones(a['teacher_keepid'].numel(), dtype=torch.bool)
    keep[n0:n0 + n1] = False
    lab = a['student_labels'].clone()
    Q = cpu.hs_student.shape[1]
    seg = lab[Q:2 * Q]
    seg[seg < 40] = 80
    a2 = dict(a, teacher_bboxes=[a['teacher_bboxes'][0], a['teacher_bboxes'][1][:0], a['teacher_bboxes'][2]],
              teacher_keepid=a['teacher_keepid'][keep], teacher_labels=a['teacher_labels'][keep], student_labels=lab)
    cpu.assignments = a2
    gpu = cpu.to(DEV)
    for crit in ('mse', 'kl'):
        mod = dskd_b200.DSGFeatureDistillLoss(criterion=crit, validate=True)
        feats, hs = gpu.clone_student()
        loss = mod(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments)
        loss.backward()
        o_feats, o_hs = cpu.clone_student()
        ref = oracle_decode(cpu, crit_oracle(crit), 1, o_feats, o_hs)
        ref.backward()
        if crit == 'kl':
            assert_kl_loss(loss, cpu, crit_oracle(crit), ref)
        else:
            assert_loss(loss, ref)
        assert_grad(hs.grad, o_hs.grad)
    # no boxes at all: loss 0, zero gradients
    a3 = dict(a2, teacher_bboxes=[b[:0] for b in a['teacher_bboxes']], teacher_keepid=a['teacher_keepid'][:0],
              teacher_labels=a['teacher_labels'][:0], student_labels=torch.full_like(lab, 80))
    cpu.assignments = a3
    gpu = cpu.to(DEV)
    feats, hs = gpu.clone_student()
    loss = dskd_b200.DSGFeatureDistillLoss(criterion='mse')(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    assert float(loss.detach()) == 0.0 and all(float(f.grad.abs().max()) == 0.0 for f in feats)


def test_validate_raises_like_reference_when_pairs_are_missing():
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=13, channels=64, **SMALL)
    gpu = cpu.to(DEV)
    a = dict(gpu.assignments)
    a['student_labels'] = torch.full_like(a['student_labels'], 80)
    with pytest.raises(IndexError):
        dskd_b200.DSGFeatureDistillLoss(validate=True)(gpu.student_feats, gpu.teacher_feats,
                                                       (gpu.hs_student, gpu.hs_teacher), a)


def test_grad_output_scaling_and_single_backward():
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=14, channels=64, **SMALL)
    gpu = cpu.to(DEV)
    mod = dskd_b200.DSGFeatureDistillLoss(criterion='mse')
    feats, hs = gpu.clone_student()
    mod(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments).backward()
    feats3, hs3 = gpu.clone_student()
    loss3 = mod(feats3, gpu.teacher_feats, (hs3, gpu.hs_teacher), gpu.assignments)
    (loss3 * 3.0).backward(retain_graph=True)
    # two forward passes differ in the last bits (the order of the fp32 atomics into the energy table is not fixed,
    # and the softmax backward subtracts nearly equal terms), hence 1e-4 with a floor relative to the largest gradient
    torch.testing.assert_close(hs3.grad, hs.grad * 3.0, rtol=1e-4, atol=1e-6 * float(hs.grad.abs().max()) * 3.0)
    torch.testing.assert_close(feats3[0].grad, feats[0].grad * 3.0, rtol=1e-6, atol=0)
    with pytest.raises(RuntimeError):
        (loss3 * 3.0).backward()
    # loss_weight is linear
    l1 = dskd_b200.DSGFeatureDistillLoss(loss_weight=1.0)(gpu.student_feats, gpu.teacher_feats, (gpu.hs_student, gpu.hs_teacher), gpu.assignments)
    l2 = dskd_b200.DSGFeatureDistillLoss(loss_weight=2.5)(gpu.student_feats, gpu.teacher_feats, (gpu.hs_student, gpu.hs_teacher), gpu.assignments)
    torch.testing.assert_close(l2, l1 * 2.5, rtol=1e-5, atol=0)


def test_cpu_tensors_are_refused():
    cpu = synth.make_distill_inputs(num_images=1, num_prev=40, seed=15, channels=64, **SMALL)
    with pytest.raises(dskd_b200._lib.DskdError):
        dskd_b200.DSGFeatureDistillLoss()(cpu.student_feats, cpu.teacher_feats, (cpu.hs_student, cpu.hs_teacher), cpu.assignments)


# ------------------------------------------------------------------ BASELINE full size
@pytest.mark.parametrize('crit', ['mse', 'kl'])
def test_full_size_coco_batch2_vs_oracle(crit):
    """BASELINE.json configs[0]: batch 2, 800x1333, 4-level 256-channel features, L = 40."""
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=1234)
    gpu = cpu.to(DEV)
    mod = dskd_b200.DSGFeatureDistillLoss(criterion=crit)
    feats, hs = gpu.clone_student()
    loss = mod(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    o_feats, o_hs = cpu.clone_student()
    ref = oracle_decode(cpu, crit_oracle(crit), 1, o_feats, o_hs)
    ref.backward()
    if crit == 'kl':
        assert_kl_loss(loss, cpu, crit_oracle(crit), ref)
    else:
        assert_loss(loss, ref)
    assert_grad(hs.grad, o_hs.grad)
    if crit == 'mse':
        for g, r in zip(feats, o_feats):
            assert_grad(g.grad, r.grad)


def test_full_size_properties_batch16():
    """Size-independent checks at the bench's size (16 images): the gradient w.r.t. the student features is
    -2*w/N*M^2*(T-S), so <grad, S - T> == 2 * loss, and it vanishes outside the teacher boxes."""
    gpu = synth.make_distill_inputs(num_images=16, num_prev=40, seed=99, device=DEV)
    feats, hs = gpu.clone_student()
    loss = dskd_b200.DSGFeatureDistillLoss(criterion='mse')(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    inner = sum(((f.grad.double()) * (f.detach().double() - t.double())).sum() for f, t in zip(feats, gpu.teacher_feats))
    torch.testing.assert_close(inner, 2.0 * loss.detach().double(), rtol=1e-5, atol=0)
    frac = sum(int((f.grad != 0).sum()) for f in feats) / sum(f.numel() for f in feats)
    assert 0.05 < frac < 0.95


# ------------------------------------------------------------------ stress: maximum detections, degenerate boxes
def _degenerate_case():
    """100 teacher detections per image (the `max_per_img` of teacher_test_cfg), heavy overlap, and every kind of
    degenerate rectangle the slice arithmetic of head_il.py:688-706 can meet: zero width / height (empty slice),
    sub-cell boxes, boxes on the image border, exact duplicates, and a teacher query kept twice (two classes above
    the score threshold: `filter_scores_and_topk` keeps (query, class) pairs)."""
    cpu = synth.make_distill_inputs(num_images=2, num_prev=40, seed=77, levels=((25, 42), (13, 21), (7, 11), (4, 6)),
                                    img_hw=(200, 333), num_query=300, channels=64, boxes_per_image=100)
    a = cpu.assignments
    h, w = 200.0, 333.0
    for i in range(2):
        b = a['teacher_bboxes'][i]
        b[0] = torch.tensor([10.0, 20.0, 10.0, 90.0])          # zero width, integer edge: empty slice
        b[1] = torch.tensor([16.0, 30.0, 80.0, 30.0])          # zero height on a cell edge
        b[2] = torch.tensor([50.3, 60.2, 50.9, 60.7])          # inside one cell
        b[3] = torch.tensor([0.0, 0.0, w, h])                  # the whole image
        b[4] = torch.tensor([w - 1.0, h - 1.0, w, h])          # last cell, touching the border
        b[5] = b[6].clone()                                    # exact duplicate of a later box
        b[7] = torch.tensor([0.0, 0.0, 0.5, 0.5])              # first cell
    keep = a['teacher_keepid'].clone()
    keep[1] = keep[0]                                          # the same teacher query kept twice
    a['teacher_keepid'] = keep
    return cpu


@pytest.mark.parametrize('crit', ['mse', 'kl'])
def test_max_detections_and_degenerate_boxes(crit):
    cpu = _degenerate_case()
    gpu = cpu.to(DEV)
    mod = dskd_b200.DSGFeatureDistillLoss(criterion=crit, validate=True)
    feats, hs = gpu.clone_student()
    loss = mod(feats, gpu.teacher_feats, (hs, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    o_feats, o_hs = cpu.clone_student()
    ref = oracle_decode(cpu, crit_oracle(crit), 1, o_feats, o_hs)
    ref.backward()
    if crit == 'kl':
        assert_kl_loss(loss, cpu, crit_oracle(crit), ref)
    else:
        assert_loss(loss, ref)
        for got, want in zip(feats, o_feats):
            assert_grad(got.grad, want.grad)
    assert_grad(hs.grad, o_hs.grad)


def test_max_detections_bcdd_with_duplicate_teacher_query():
    cpu = _degenerate_case()
    gpu = cpu.to(DEV)
    a = cpu.assignments
    _, hs = gpu.clone_student()
    mod = dskd_b200.BetweenClassDistanceLoss()
    loss = mod(None, None, (hs, gpu.hs_teacher), gpu.assignments)
    loss.backward()
    _, o_hs = cpu.clone_student()
    C = o_hs.shape[-1]
    ref = ob.bcdd_loss(o_hs.reshape(-1, C), a['student_labels'], cpu.hs_teacher.reshape(-1, C), a['teacher_keepid'],
                       a['teacher_labels'], a['prev_labels'], ol.MSELoss('mean', 1.0))
    ref.backward()
    assert_loss(loss, ref)
    assert_grad(hs.grad, o_hs.grad)
