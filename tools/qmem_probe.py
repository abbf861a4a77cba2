"""Exploratory check of the tcgen05 query x memory kernel against oracle/qmem.py (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dskd_b200 import synth, qmem
from oracle import qmem as oq


def run(name, n_images, levels, num_query, channels, kpi, temp=0.5, with_scores=True):
    cpu = synth.make_distill_inputs(num_images=n_images, num_prev=40, seed=3, levels=levels, num_query=num_query,
                                    channels=channels, boxes_per_image=kpi, img_hw=(192, 320))
    a = cpu.assignments
    _, t_mem = cpu.memory()
    counts = [b.shape[0] for b in a['teacher_bboxes']]
    g = torch.Generator().manual_seed(1)
    scores = (0.3 + 0.7 * torch.rand(sum(counts), generator=g)) if with_scores else None
    start = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32)
    dev = 'cuda:0'
    w = qmem.qmem_cell_weights(t_mem.to(dev), cpu.hs_teacher.to(dev), a['teacher_keepid'].to(dev),
                               None if scores is None else scores.to(dev), start.to(dev), max(counts), temp)
    torch.cuda.synchronize()
    w = w.cpu()
    out = {}
    for mode in (None, 'trunc', 'rne'):
        ref = oq.qmem_cell_weights(t_mem, cpu.hs_teacher, a['teacher_keepid'], scores, counts, temp, tf32=mode,
                                   dtype=torch.float64)
        err = (w.double() - ref).abs().max().item()
        out[mode] = err
    print(f'{name}: S={t_mem.shape[0]} N={n_images} C={channels} K={counts[:4]} max|w|={w.abs().max():.4f} '
          f'mean w={w.mean():.4f} err fp32={out[None]:.3e} trunc={out["trunc"]:.3e} rne={out["rne"]:.3e} '
          f'nan={int(torch.isnan(w).sum())}', flush=True)


if __name__ == '__main__':
    small = ((24, 40), (12, 20), (6, 10), (3, 5))
    run('small-1blk', 2, small, 100, 256, 20)
    run('small-c64', 2, small, 100, 64, 7)
    run('small-ragged', 3, small, 100, 256, None)
    run('small-2blk', 2, small, 320, 256, 300)
    run('small-4blk', 1, small, 700, 256, 600)
    run('coco-1blk', 2, synth.COCO_LEVELS, 300, 256, 100)
    run('coco-2blk', 2, synth.COCO_LEVELS, 320, 256, 300, with_scores=False)
