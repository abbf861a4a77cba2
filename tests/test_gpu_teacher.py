"""Teacher keep-ids on the GPU (row A1) vs the reference's own outputs (tests/golden/assign.npz `decode.*`,
head case `pred_*`) and vs the oracle on COCO-shaped batches.  Indices bit-exact, boxes within 1 ulp-ish
(the kernel uses the same IEEE operations; scores go through CUDA's expf instead of the CPU's)."""
import pytest
import torch

from dskd_b200 import teacher
from oracle import assign as oa
from conftest import Golden, load_head_case

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _check_against(info, ref, n):
    assert torch.equal(info['pred_keepid'].cpu(), ref['pred_keepid'])
    for i in range(n):
        assert torch.equal(info['pred_labels'][i].cpu(), ref['pred_labels'][i]), i
        torch.testing.assert_close(info['pred_scores'][i].cpu(), ref['pred_scores'][i], rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(info['pred_bboxes'][i].cpu(), ref['pred_bboxes'][i], rtol=1e-6, atol=1e-4)
    start = info['box_start'].cpu().tolist()
    assert start[0] == 0 and [start[i + 1] - start[i] for i in range(n)] == [len(l) for l in ref['pred_labels']]


def test_single_image_vs_reference_get_bboxes_single():
    g = Golden('assign.npz')
    info = teacher.teacher_info_from_outputs(g.t('decode.cls')[None].to(DEV), g.t('decode.box')[None].to(DEV),
                                             [(800, 1333)], score_thr=0.3, max_per_img=100, need_logits=True)
    assert torch.equal(info['pred_keepid'].cpu(), g.t('decode.keep'))
    assert torch.equal(info['pred_labels'][0].cpu(), g.t('decode.labels'))
    det = g.t('decode.det')
    torch.testing.assert_close(info['pred_bboxes'][0].cpu(), det[:, :4], rtol=1e-6, atol=1e-4)
    torch.testing.assert_close(info['pred_scores'][0].cpu(), det[:, 4], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(info['pred_logits'][0].cpu(), g.t('decode.cls').sigmoid()[g.t('decode.keep')],
                               rtol=1e-6, atol=1e-7)


def test_filter_scores_and_topk_order_vs_reference():
    # the reference's stored `filter_scores_and_topk` case takes probabilities; feed their logits
    g = Golden('assign.npz')
    p = g.t('topk.scores_in').double()
    logits = torch.log(p / (1 - p)).float().clamp(-30, 30)
    ref = oa.filter_scores_and_topk(logits.sigmoid(), 0.3, 100)
    box = torch.rand(1, logits.shape[0], 70)
    info = teacher.teacher_info_from_outputs(logits[None].to(DEV), box.to(DEV), [(800, 1333)])
    assert torch.equal(info['pred_labels'][0].cpu(), ref[1])
    assert torch.equal(info['pred_keepid'].cpu(), ref[2])


@pytest.mark.parametrize('name', ['head_decode_v1_mse.npz', 'head_decode_v1_kl_l70.npz', 'head_decode_v1_kl_l50.npz',
                                  'head_decode_v1_kl_l60.npz'])
def test_batch_vs_reference_head_case(name):
    inp, out = load_head_case(name)
    N = inp.t('t_cls').shape[1]
    img_hw = tuple(inp.t('img_hw').tolist())
    info = teacher.teacher_info_from_outputs(inp.t('t_cls')[-1].to(DEV), inp.t('t_box')[-1].to(DEV), [img_hw] * N)
    assert torch.equal(info['pred_keepid'].cpu(), out.t('pred_keepid'))
    for i in range(N):
        assert torch.equal(info['pred_labels'][i].cpu(), out.lst('pred_labels')[i])
        torch.testing.assert_close(info['pred_bboxes'][i].cpu(), out.lst('pred_bboxes')[i], rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize('shift,max_per', [(-2.0, 100), (-6.0, 100), (-20.0, 100), (1.0, 100), (-2.0, 7), (-1.0, 1000)],
                         ids=['coco', 'few', 'none', 'all-valid', 'top7', 'top1000'])
def test_coco_shape_batch_vs_oracle(shift, max_per):
    g = torch.Generator().manual_seed(5)
    N, Q, NC = 6, 300, 80
    cls = torch.randn(N, Q, NC, generator=g) + shift
    cls[2] -= 30.0                                          # an image without any detection
    box = torch.rand(N, Q, 70, generator=g)
    shapes = [(800, 1333), (800, 1199), (608, 800), (800, 1333), (750, 1333), (800, 1067)]
    ref = oa.teacher_info_from_outputs(cls, box, shapes, 0.3, max_per)
    info = teacher.teacher_info_from_outputs(cls.to(DEV), box.to(DEV), shapes, 0.3, max_per)
    _check_against(info, ref, N)
    assert info['pred_labels'][2].numel() == 0
    nosplit = teacher.teacher_info_from_outputs(cls.to(DEV), box.to(DEV), shapes, 0.3, max_per, split=False)
    total = int(nosplit['box_start'][-1])
    assert torch.equal(nosplit['pred_keepid'][:total].cpu(), ref['pred_keepid'])


def test_ties_keep_row_major_order():
    # equal scores: the reference's sort keeps the row-major nonzero() order (stable); so does the key
    cls = torch.full((1, 50, 80), -10.0)
    cls[0, 7, 3] = cls[0, 2, 9] = cls[0, 2, 5] = cls[0, 40, 0] = 2.0
    cls[0, 11, 11] = 3.0
    box = torch.rand(1, 50, 70)
    ref = oa.teacher_info_from_outputs(cls, box, [(100, 100)])
    info = teacher.teacher_info_from_outputs(cls.to(DEV), box.to(DEV), [(100, 100)])
    assert info['pred_keepid'].cpu().tolist() == [11, 2, 2, 7, 40]
    assert info['pred_labels'][0].cpu().tolist() == [11, 5, 9, 3, 0]
    assert torch.equal(info['pred_keepid'].cpu(), ref['pred_keepid'])


def test_decoded_boxes_and_errors():
    from dskd_b200._lib import DskdError
    cls = torch.randn(2, 30, 80) - 1.0
    box4 = torch.rand(2, 30, 4) * 0.5 + 0.25
    info = teacher.teacher_info_from_outputs(cls.to(DEV), box4.to(DEV), [(64, 96)] * 2, reg_max=0, max_per_img=10)
    assert all(b.shape[0] <= 10 for b in info['pred_bboxes'])
    with pytest.raises(DskdError):
        teacher.teacher_info_from_outputs(cls.to(DEV), box4.to(DEV), [(64, 96)] * 2)       # 4 channels, reg_max 16
    with pytest.raises(DskdError):
        teacher.teacher_info_from_outputs(cls, box4, [(64, 96)] * 2, reg_max=0)            # CPU tensors
    with pytest.raises(DskdError):
        teacher.teacher_info_from_outputs(cls.to(DEV), box4.to(DEV), [(64, 96)] * 2, reg_max=0, max_per_img=5000)
