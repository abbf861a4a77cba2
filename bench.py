#!/usr/bin/env python
"""Benchmark of the DSKD distillation hot path on B200 (see DESIGN.md section "Measurement").

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 20 --warmup 5
    python bench.py --impl reference --steps 3 --warmup 1        # the reference's CPU arithmetic

One "step" = distillation loss forward + backward for one batch of synthetic COCO-shaped inputs:
DSG-FD (decode_v1 mask, masked MSE) over 4-level 256-channel features + BCDD over the last-layer
decoder embeddings, `images_per_gpu` 800x1333 images per rank (weak scaling; per-class prototype
sums/counts are the only cross-rank state: one NCCL all-reduce).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'distill_loss_fwd_bwd_images_per_s'
UNIT = 'images/s'
BYTES_PER_IMAGE_MSE = 3 * 22223 * 256 * 4      # read S + read T + write dS  (SURVEY.md section 8d): 68.27 MB
WORKLOAD = 'coco_40+40_dsgfd_mse+bcdd'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='dskd_b200', choices=['dskd_b200', 'reference'])
    ap.add_argument('--images-per-gpu', type=int, default=16,
                    help='samples_per_gpu of the 40+40 config (chaosuan_..._40_r50_8x4_1x_qoqo_il.py:199)')
    ap.add_argument('--num-prev', type=int, default=40)
    ap.add_argument('--criterion', default='mse', choices=['mse', 'kl'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches only')
    ap.add_argument('--plain-graph', action='store_true', help='replay with torch.cuda.CUDAGraph (no node priorities)')
    ap.add_argument('--cpu-sample-images', type=int, default=2)
    ap.add_argument('--no-feeders', action='store_true', help='skip the teacher keep-id / assignment timings')
    ap.add_argument('--no-contraction', action='store_true', help='skip the tcgen05 query x memory kernel line')
    ap.add_argument('--contraction-queries', type=int, default=300,
                    help="queries contracted against every memory token (north_star: ~300 x 256 against ~20k x 256)")
    return ap.parse_args()


def peaks():
    """(HBM GB/s, bf16 TFLOP/s burst, source) -- measured by the driver, else the profiling guide's fallback."""
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), float(p['bf16_tflops']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1590.0, 'fallback (B200_PROFILING.md)'


def recorded_traffic(kernel, images_per_gpu):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
    very command (profiles/traffic.json); None when no capture matches the configuration."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if not os.path.exists(path):
        return None
    with open(path) as f:
        table = json.load(f)
    rec = table.get(kernel)
    if rec and int(rec.get('images_per_gpu', -1)) == int(images_per_gpu):
        return rec
    return None


def time_path_feeders(args, dev, N):
    """The rows of the hot path that FEED the losses, timed by wall clock around a device sync
    (the host solver contains one): teacher keep-ids for N images, and the Hungarian assignment of all 6*N
    (layer, image) problems with the device LSAP kernel and with the C++ host solver."""
    import dskd_b200
    from dskd_b200 import synth
    ai = synth.make_assign_inputs(num_images=N, seed=1234, device=dev)
    t_cls = ai.cls_logits[-1] + 0.5
    out = {}

    def wall(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return statistics.median(ts) * 1e3

    out['teacher_keepids_ms'] = wall(lambda: dskd_b200.teacher_info_from_outputs(t_cls, ai.box_pred[-1], ai.img_shapes))
    for solver in ('device', 'host'):
        asg = dskd_b200.GFLHungarianAssigner(solver=solver)
        out[f'assignment_{solver}_lsap_ms'] = wall(lambda: asg.assign_batch(
            ai.cls_logits, ai.box_pred, ai.gt_bboxes, ai.gt_labels, ai.img_shapes, prev_labels=list(range(args.num_prev))))
    out['problems'] = int(ai.cls_logits.shape[0] * N)
    out['note'] = ('wall clock incl. Python and a final device sync; the device solver itself needs no sync and is '
                   'index-identical to SciPy (tests/test_gpu_lsap.py)')
    return out


def time_contraction(args, dev, N, world, dist):
    """The tcgen05 query x memory contraction at the north_star shape: K queries x 256 channels
    against the 22 223 x 256 memory tokens of each of this rank's N images; CUDA-graph replay, L2 flushed between
    replays, CUDA events; FLOPs = 2*K*C*S per image."""
    from dskd_b200 import qmem
    K, S, C, Q = args.contraction_queries, 22223, 256, max(300, args.contraction_queries)
    g = torch.Generator(device=dev).manual_seed(4321)
    mem = torch.randn(S, N, C, device=dev, generator=g)
    hs = torch.randn(N, Q, C, device=dev, generator=g)
    keep = torch.cat([torch.randperm(Q, device=dev, generator=g)[:K] + i * Q for i in range(N)])
    sc = torch.rand(N * K, device=dev, generator=g)
    start = torch.arange(N + 1, device=dev, dtype=torch.int32) * K
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            qmem.qmem_cell_weights(mem, hs, keep, sc, start, K, 0.5)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        qmem.qmem_cell_weights(mem, hs, keep, sc, start, K, 0.5)
    times = []
    for _ in range(max(args.steps, 5)):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = statistics.median(times)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    flop = 2.0 * K * C * S * N
    _, bf16_peak, src = peaks()
    achieved = flop / (ms * 1e-3) / 1e12
    return {'bound': 'tensor', 'kernel': 'qmem_weight_kernel (tcgen05.mma kind::tf32, TMA-fed, TMEM accumulators)',
            'achieved': achieved, 'peak': bf16_peak, 'unit': 'TFLOP/s', 'frac': achieved / bf16_peak,
            'frac_of_tf32_rate': achieved / (bf16_peak / 2),
            'peak_source': f'{src}: cuBLAS bf16 dense; kind::tf32 issues at half that rate',
            'flop_per_launch': flop, 'call_ms': ms, 'queries': K, 'tokens': S, 'channels': C, 'images': N,
            'images_per_s': world * N / (ms * 1e-3), 'traffic': None,
            'timing': 'gather + contraction + combine, CUDA-graph replay, L2 flushed between replays, median'}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.gpu_index}', f'--query-gpu={self.QUERY}',
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    power_w_max=max(power) if power else None, samples=len(sm), reasons=sorted(reasons))


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step(cpu_inputs, criterion):
    """One fwd+bwd of the reference's arithmetic (the oracle restatement) on the host cores."""
    from oracle import bcdd as ob, dsgfd as od, losses as ol
    a = cpu_inputs.assignments
    shapes = [tuple(x) for x in a['img_shapes'].tolist()]
    L = len(a['prev_labels'])
    id_pred = torch.nonzero(a['student_labels'] < L).squeeze(1)
    feats, hs = cpu_inputs.clone_student()
    crit = ol.MSELoss('sum', 1.0) if criterion == 'mse' else ol.KnowledgeDistillationKLDivLoss('sum', 1.0, 2)
    C = hs.shape[-1]
    loss = od.decode_v1(feats, cpu_inputs.teacher_feats, hs, cpu_inputs.hs_teacher, a['teacher_keepid'], id_pred,
                        a['teacher_bboxes'], shapes, crit)
    loss = loss + ob.bcdd_loss(hs.reshape(-1, C), a['student_labels'], cpu_inputs.hs_teacher.reshape(-1, C),
                               a['teacher_keepid'], a['teacher_labels'], a['prev_labels'], ol.MSELoss('mean', 1.0))
    loss.backward()
    return float(loss)


def time_cpu_reference(args, steps, warmup):
    from dskd_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = args.cpu_sample_images
    cpu_inputs = synth.make_distill_inputs(num_images=n, num_prev=args.num_prev, seed=1234)
    for _ in range(warmup):
        cpu_reference_step(cpu_inputs, args.criterion)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_reference_step(cpu_inputs, args.criterion)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return dict(value=n * steps / total, unit=UNIT, cores=cores, kind='port',
                sample=f'{n} images/step x {steps} steps (+{warmup} warm-up) of the same workload, oracle '
                       f'restatement of head_il.py:525-555,664-719,1197-1222, torch {torch.__version__} CPU fp32'), \
        total / steps * 1e3


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    base, ms = time_cpu_reference(args, args.steps, args.warmup)
    line = {'metric': METRIC, 'value': base['value'], 'unit': UNIT, 'impl': 'reference', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'images_per_step': args.cpu_sample_images, 'num_prev': args.num_prev,
                       'criterion': args.criterion, 'levels': [[100, 167], [50, 84], [25, 42], [13, 21]],
                       'channels': 256, 'queries': 300},
            'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
        return
    import torch.distributed as dist
    import dskd_b200
    from dskd_b200 import _lib, synth

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl dskd_b200 needs a CUDA device (there is no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()

    N = args.images_per_gpu
    inputs = synth.make_distill_inputs(num_images=N, num_prev=args.num_prev, seed=1234 + rank, device=dev)
    dsg = dskd_b200.build_loss(dict(type='DSGFeatureDistillLoss', criterion=args.criterion, reduction='sum',
                                    loss_weight=1.0, T=2.0, mask_mode='decode_v1', feature_source='neck'))
    bcdd = dskd_b200.build_loss(dict(type='BetweenClassDistanceLoss', reduction='mean', loss_weight=1.0,
                                     sync_prototypes=world > 1))
    s_feats = [f.requires_grad_(True) for f in inputs.student_feats]
    hs_s = inputs.hs_student.requires_grad_(True)

    from dskd_b200 import profiling

    # high priority: the latency-bound BCDD chain (and its NCCL all-reduce) must not queue behind the 3 712 CTAs of the
    # HBM-bound DSG-FD kernel -- its few CTAs take the first SM slots that free up
    bcdd_stream = torch.cuda.Stream(priority=-1)

    def step(feats, t_feats, hs, hs_t):
        for f in feats:
            f.grad = None
        hs.grad = None
        # BCDD (latency-bound, and the only collective) runs on a side stream behind the HBM-bound DSG-FD kernel
        cur = torch.cuda.current_stream()
        bcdd_stream.wait_stream(cur)
        with torch.cuda.stream(bcdd_stream):
            loss_corr = bcdd(None, None, (hs, hs_t), inputs.assignments)
        loss_fg = dsg(feats, t_feats, (hs, hs_t), inputs.assignments)
        cur.wait_stream(bcdd_stream)
        loss = loss_fg + loss_corr
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput, eager launches (host-issue bound for this small step)
    for _ in range(max(args.warmup, 3)):
        step(s_feats, inputs.teacher_feats, hs_s, inputs.hs_teacher)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.dskd_launch_count()
    profiling.start()       # C side records an event pair around the streaming kernel of every step
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_start.record()
    for _ in range(args.steps):
        loss = step(s_feats, inputs.teacher_feats, hs_s, inputs.hs_teacher)
    e_stop.record()
    barrier()
    kernel_times = profiling.stop()
    launches = lib.dskd_launch_count() - launches0
    eager_ms = e_start.elapsed_time(e_stop)
    loss_value = float(loss.detach())

    # ---------------- the same step captured once in a CUDA graph and replayed (`value`): identical kernels,
    # identical inputs, no per-launch host cost.  Falls back to the eager number if capture is not possible.
    graph_ms = None
    graph_err = None
    graph_policy = 'torch replay (stream order)'
    if not args.no_graph:
        try:
            # fresh leaves: their AccumulateGrad nodes must first be used on the capture (side) stream
            g_feats = [f.detach().clone().requires_grad_(True) for f in s_feats]
            g_hs = hs_s.detach().clone().requires_grad_(True)
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step(g_feats, inputs.teacher_feats, g_hs, inputs.hs_teacher)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for f in g_feats:
                f.grad = None
            g_hs.grad = None
            # keep_graph: the library instantiates the captured cudaGraph_t itself, with per-node priorities (small
            # kernels -- mask build, BCDD, the NCCL all-reduce -- ahead of the streaming kernel's 3 712 CTAs)
            graph = torch.cuda.CUDAGraph(keep_graph=not args.plain_graph)
            with torch.cuda.graph(graph):
                graph_loss = step(g_feats, inputs.teacher_feats, g_hs, inputs.hs_teacher)
            if args.plain_graph:
                runner = graph
            else:
                try:
                    from dskd_b200.graphs import PrioritizedGraph
                    runner = PrioritizedGraph(graph)
                    graph_policy = f'node priorities: {runner.num_small} small kernels first, {runner.num_big} streaming'
                except Exception as exc:            # noqa: BLE001 -- fall back to PyTorch's own instantiation, say so
                    graph.instantiate()
                    runner = graph
                    graph_policy = f'torch replay (stream order); prioritized instantiation failed: {exc}'
            for _ in range(max(args.warmup, 3)):
                runner.replay()
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(args.steps):
                runner.replay()
            g1.record()
            barrier()
            graph_ms = g0.elapsed_time(g1)
            if abs(float(graph_loss.detach()) - loss_value) > 1e-3 * abs(loss_value):
                raise RuntimeError(f'graph replay loss {float(graph_loss.detach())} != eager loss {loss_value}')
        except Exception as exc:                    # noqa: BLE001 -- report, do not hide
            graph_err = f'{type(exc).__name__}: {exc}'
            graph_ms = None
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = graph_ms if graph_ms is not None else eager_ms
    t = torch.tensor([elapsed_ms, eager_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, eager_ms = float(t[0].item()), float(t[1].item())
    kernel_ms = statistics.mean(kernel_times) if kernel_times else float('nan')

    # ---------------- end-to-end through the module call with HOST buffers (`e2e`)
    e2e = None
    if not args.no_e2e:
        host_s = [torch.empty(f.shape, dtype=f.dtype, pin_memory=True).copy_(f.detach()) for f in s_feats]
        host_t = [torch.empty(f.shape, dtype=f.dtype, pin_memory=True).copy_(f) for f in inputs.teacher_feats]
        host_hs = torch.empty(hs_s.shape, pin_memory=True).copy_(hs_s.detach())
        host_ht = torch.empty(hs_s.shape, pin_memory=True).copy_(inputs.hs_teacher)
        host_loss = torch.empty((), pin_memory=True)
        h2d = sum(t_.numel() * 4 for t_ in host_s + host_t + [host_hs, host_ht])

        def e2e_step():
            fs = [h.to(dev, non_blocking=True).requires_grad_(True) for h in host_s]
            ft = [h.to(dev, non_blocking=True) for h in host_t]
            hs = host_hs.to(dev, non_blocking=True).requires_grad_(True)
            ht = host_ht.to(dev, non_blocking=True)
            l = step(fs, ft, hs, ht)
            host_loss.copy_(l.detach(), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return float(host_loss)
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {'value': world * N * e2e_steps / float(tt.item()), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': 4, 'steps': e2e_steps, 'ms_per_step': float(tt.item()) / e2e_steps * 1e3}

    contraction = None
    if not args.no_contraction:
        contraction = time_contraction(args, dev, N, world, dist)

    if rank != 0:
        leave(world, dist)
        return

    peak, _, peak_src = peaks()
    if args.criterion == 'mse':
        alg_bytes = BYTES_PER_IMAGE_MSE * N
    else:
        alg_bytes = 2 * 22223 * 256 * 4 * N
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    kernel_name = 'dsgfd_mse_nchw_kernel' if args.criterion == 'mse' else 'dsgfd_kl_col_kernel'
    traffic_rec = recorded_traffic(kernel_name, N)
    ms_per_step = elapsed_ms / args.steps
    line = {
        'metric': METRIC, 'value': world * N * args.steps / (elapsed_ms * 1e-3), 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'images_per_gpu': N, 'global_batch': world * N, 'num_prev': args.num_prev,
                   'criterion': args.criterion, 'levels': [[100, 167], [50, 84], [25, 42], [13, 21]], 'channels': 256,
                   'queries': 300, 'parallelism': f'dp{world}', 'prototype_allreduce': world > 1,
                   'launch': 'cuda_graph_replay' if graph_ms is not None else 'eager',
                   'graph_policy': graph_policy if graph_ms is not None else None,
                   'cache': f'inputs larger than L2: {2 * N * 22.76:.0f} MB of features read per step vs 126 MB L2'},
        'roofline': {'bound': 'hbm', 'kernel': 'dsgfd_mse_nchw_kernel' if args.criterion == 'mse' else 'dsgfd_kl_col_kernel',
                     'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'frac_of_nominal_8000': achieved / 8000.0, 'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': alg_bytes, 'kernel_ms': kernel_ms,
                     'kernel_share_of_step': kernel_ms / ms_per_step,
                     'traffic': (traffic_rec or {}).get('dram_bytes_per_launch'),
                     'traffic_source': (traffic_rec or {}).get('source')},
        'contraction': contraction,
        'e2e': e2e,
        'gpu_launches': int(launches),
        'eager': {'value': world * N * args.steps / (eager_ms * 1e-3), 'unit': UNIT, 'ms_per_step': eager_ms / args.steps,
                  'note': 'same step issued kernel by kernel from Python; host-issue bound', 'graph_error': graph_err},
        'clocks': clocks,
        'loss': loss_value,
    }
    if world == 1 and not args.no_feeders:
        line['feeders'] = time_path_feeders(args, dev, N)
    if world == 1 and not args.no_cpu_baseline:
        base, _ = time_cpu_reference(args, steps=3, warmup=1)
        line['cpu_baseline'] = base
    print(json.dumps(line), flush=True)
    leave(world, dist)


def leave(world, dist):
    """End of a multi-rank run.  `destroy_process_group()` blocked for minutes on the B200 box once the step (with its
    NCCL all-reduce) had been captured in a CUDA graph, so after a last barrier every rank flushes and exits directly;
    torchrun sees exit code 0."""
    if world <= 1:
        return
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == '__main__':
    main()
