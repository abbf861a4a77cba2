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

Besides the headline the line carries, measured in the same run:
  kl            the same step with the criterion every shipped config uses (KL over H, T = 2) and its kernel's roofline
  configs       BASELINE.json configs 3-5: 70+10 at 8 images/GPU, 50+30 and 60+20 at 4 images/GPU (prototype all-reduce
                at every N), and a trimmed token x query x batch sweep with the CPU port beside its smallest point
  multi_rank_parity  (N > 1) the synced BCDD loss / gradient of this rank against the single-process evaluation of the
                concatenated batch on rank 0 (SURVEY.md 8e parity definition)
  train_step    the 40+40 incremental training step hosting the losses
"""
import argparse
import gc
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
COCO_TOKENS = 22223
WORKLOAD = 'coco_40+40_dsgfd_mse+bcdd'
PARITY_NOTE = ('loss rtol 1e-4 / grad rtol 1e-3 vs the fp32 oracle and the reference golden outputs (tests/); KL-over-H loss: '
               '1e-4 vs the float64 evaluation of the reference formula, 1e-3 vs its fp32 evaluation (a second-order quantity: '
               "the reference's own fp32 value carries 1.4e-4..3.7e-4 of noise, DESIGN.md section 2)")


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
    ap.add_argument('--no-kl', action='store_true', help='skip the KL-criterion extra key')
    ap.add_argument('--no-configs', action='store_true', help='skip BASELINE configs 3-5 (L=70/50/60, sweep)')
    ap.add_argument('--no-train-step', action='store_true', help='skip the 40+40 training-step extra key')
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


def algorithmic_bytes(criterion, tokens, channels, images):
    """masked MSE fwd+bwd reads S, T and writes dS; KL fwd+bwd reads S and T (no feature gradient)."""
    return (3 if criterion == 'mse' else 2) * tokens * channels * 4 * images


STREAM_KERNEL = {'mse': 'dsgfd_mse_nchw_kernel', 'kl': 'dsgfd_kl_stream_kernel'}
CUBLAS_TF32_TFLOPS = 720.1  # torch fp32 matmul with allow_tf32, 8192^3, B200 (profiles/r2/cublas_tf32_bf16_peak.txt)


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
    K, S, C, Q = args.contraction_queries, COCO_TOKENS, 256, max(300, args.contraction_queries)
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
    del graph
    flop = 2.0 * K * C * S * N
    _, bf16_peak, src = peaks()
    achieved = flop / (ms * 1e-3) / 1e12
    kernel = 'qmem_weight_pair_kernel' if K > 160 else 'qmem_weight_kernel'  # QmemPlan: CTA pairs above 160 queries
    traffic_rec = recorded_traffic(kernel, N)
    if traffic_rec and traffic_rec.get('queries') != K:
        traffic_rec = None
    return {'bound': 'tensor', 'kernel': f'{kernel} (tcgen05.mma kind::tf32' + (' cta_group::2' if K > 160 else '') +
                                         ', TMA-fed, TMEM accumulators)',
            'achieved': achieved, 'peak': bf16_peak, 'unit': 'TFLOP/s', 'frac': achieved / bf16_peak,
            'frac_of_tf32_rate': achieved / (bf16_peak / 2),
            'frac_of_cublas_tf32': achieved / CUBLAS_TF32_TFLOPS,
            'peak_source': f'{src}: cuBLAS bf16 dense; kind::tf32 issues at half that rate; cuBLAS fp32-as-tf32 8192^3 '
                           f'GEMM measured {CUBLAS_TF32_TFLOPS} TFLOP/s on this pool (profiles/r2/cublas_tf32_bf16_peak.txt)',
            'flop_per_launch': flop, 'call_ms': ms, 'queries': K, 'tokens': S, 'channels': C, 'images': N,
            'images_per_s': world * N / (ms * 1e-3), 'traffic': (traffic_rec or {}).get('dram_bytes_per_launch'),
            'traffic_source': (traffic_rec or {}).get('source'),
            'timing': 'gather + contraction (+ combine above 320 queries), CUDA-graph replay, L2 flushed between replays, median'}


def gpu_numa_cpus(local_rank):
    """(numa node, cpu set) of the host socket this GPU hangs off, from sysfs; (None, None) when the box does not say."""
    try:
        prop = torch.cuda.get_device_properties(local_rank)
        bus = f'{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0'
        with open(f'/sys/bus/pci/devices/{bus}/numa_node') as f:
            node = int(f.read())
        if node < 0:
            return None, None
        cpus = set()
        with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
            for part in f.read().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus.update(range(int(lo), int(hi or lo) + 1))
        return node, cpus
    except (OSError, ValueError, AttributeError):
        return None, None


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
    return float(loss.detach())


def time_cpu_reference(criterion, num_prev, n, steps, warmup, levels=None, num_query=300):
    from dskd_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kw = {} if levels is None else dict(levels=levels)
    cpu_inputs = synth.make_distill_inputs(num_images=n, num_prev=num_prev, seed=1234, num_query=num_query, **kw)
    for _ in range(warmup):
        cpu_reference_step(cpu_inputs, criterion)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_reference_step(cpu_inputs, criterion)
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
    base, ms = time_cpu_reference(args.criterion, args.num_prev, args.cpu_sample_images, args.steps, args.warmup)
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
class StepBench:
    """One configuration of the distillation step on this rank: synthetic inputs, the two registry modules and the
    step closure (BCDD on a high-priority side stream behind the HBM-bound DSG-FD kernel, joined before the sum)."""

    def __init__(self, dev, rank, world, dist, images, num_prev, criterion, levels=None, num_query=300, plain_graph=False):
        import dskd_b200
        from dskd_b200 import synth
        self.dev, self.rank, self.world, self.dist = dev, rank, world, dist
        self.N, self.criterion, self.plain_graph = images, criterion, plain_graph
        kw = {} if levels is None else dict(levels=levels)
        self.inputs = synth.make_distill_inputs(num_images=images, num_prev=num_prev, seed=1234 + rank, device=dev,
                                                num_query=num_query, **kw)
        self.tokens = sum(h * w for h, w in self.inputs.levels)
        self.channels = self.inputs.hs_student.shape[-1]
        self.dsg = dskd_b200.build_loss(dict(type='DSGFeatureDistillLoss', criterion=criterion, reduction='sum',
                                             loss_weight=1.0, T=2.0, mask_mode='decode_v1', feature_source='neck'))
        self.bcdd = dskd_b200.build_loss(dict(type='BetweenClassDistanceLoss', reduction='mean', loss_weight=1.0,
                                              sync_prototypes=world > 1))
        self.s_feats = [f.requires_grad_(True) for f in self.inputs.student_feats]
        self.hs_s = self.inputs.hs_student.requires_grad_(True)
        # high priority: the latency-bound BCDD chain (and its NCCL all-reduce) must not queue behind the thousands of
        # CTAs of the HBM-bound DSG-FD kernel -- its few CTAs take the first SM slots that free up
        self.bcdd_stream = torch.cuda.Stream(priority=-1)

    def step(self, feats, t_feats, hs, hs_t, overlap=True):
        for f in feats:
            f.grad = None
        hs.grad = None
        if not overlap:     # the two module calls back to back on the current stream: what an eager caller does
            loss = self.dsg(feats, t_feats, (hs, hs_t), self.inputs.assignments) + \
                self.bcdd(None, None, (hs, hs_t), self.inputs.assignments)
            loss.backward()
            return loss
        cur = torch.cuda.current_stream()
        self.bcdd_stream.wait_stream(cur)
        with torch.cuda.stream(self.bcdd_stream):
            loss_corr = self.bcdd(None, None, (hs, hs_t), self.inputs.assignments)
        loss_fg = self.dsg(feats, t_feats, (hs, hs_t), self.inputs.assignments)
        cur.wait_stream(self.bcdd_stream)
        loss = loss_fg + loss_corr
        loss.backward()
        return loss

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def run_eager(self, steps, warmup, kernel_events=True):
        """(total ms, per-step streaming-kernel ms list, launches, loss): kernels issued one by one from Python.
        kernel_events: the C side brackets the streaming kernel of every step with a CUDA event pair (the roofline's live
        timing); off for the plain eager figure, which should not pay for event creation and four extra records per step."""
        from dskd_b200 import _lib, profiling
        lib = _lib.load()
        inp = self.inputs
        for _ in range(max(warmup, 3)):
            self.step(self.s_feats, inp.teacher_feats, self.hs_s, inp.hs_teacher, overlap=False)
        self.barrier()
        launches0 = lib.dskd_launch_count()
        if kernel_events:
            profiling.start()       # C side records an event pair around the streaming kernel of every step
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(steps):
            loss = self.step(self.s_feats, inp.teacher_feats, self.hs_s, inp.hs_teacher, overlap=False)
        e1.record()
        self.barrier()
        kernel_times = profiling.stop() if kernel_events else []
        return e0.elapsed_time(e1), kernel_times, lib.dskd_launch_count() - launches0, float(loss.detach())

    def capture(self):
        """The same step captured once in a CUDA graph: identical kernels, identical inputs, no per-launch host cost."""
        inp = self.inputs
        # fresh leaves: their AccumulateGrad nodes are first used on the capture stream
        g_feats = [f.detach().clone().requires_grad_(True) for f in self.s_feats]
        g_hs = self.hs_s.detach().clone().requires_grad_(True)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self.step(g_feats, inp.teacher_feats, g_hs, inp.hs_teacher)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for f in g_feats:
            f.grad = None
        g_hs.grad = None
        # keep_graph: the library instantiates the captured cudaGraph_t itself, with per-node priorities (small
        # kernels -- mask build, BCDD, the NCCL all-reduce -- ahead of the streaming kernel's CTAs)
        graph = torch.cuda.CUDAGraph(keep_graph=not self.plain_graph)
        with torch.cuda.graph(graph, stream=side):
            graph_loss = self.step(g_feats, inp.teacher_feats, g_hs, inp.hs_teacher)
        policy = 'torch replay (stream order)'
        runner = graph
        if not self.plain_graph:
            try:
                from dskd_b200.graphs import PrioritizedGraph
                runner = PrioritizedGraph(graph)
                policy = f'node priorities: {runner.num_small} small kernels first, {runner.num_big} streaming'
            except Exception as exc:            # noqa: BLE001 -- fall back to PyTorch's own instantiation, say so
                graph.instantiate()
                policy = f'torch replay (stream order); prioritized instantiation failed: {exc}'
        self._keep = (g_feats, g_hs, graph)
        return runner, graph_loss, policy

    def run_graph(self, steps, warmup, flush=None):
        """(ms per step, loss, policy).  flush: a buffer larger than L2 rewritten between replays (configurations whose
        inputs fit the 126 MB L2); each replay is then timed on its own and the median taken."""
        runner, graph_loss, policy = self.capture()
        for _ in range(max(warmup, 3)):
            runner.replay()
        self.barrier()
        if flush is None:
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(steps):
                runner.replay()
            g1.record()
            self.barrier()
            ms = g0.elapsed_time(g1) / steps
        else:
            ts = []
            for _ in range(steps):
                flush.zero_()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                runner.replay()
                g1.record()
                torch.cuda.synchronize()
                ts.append(g0.elapsed_time(g1))
            self.barrier()
            ms = statistics.median(ts)
        loss = float(graph_loss.detach())
        del runner
        self._keep = None
        return ms, loss, policy

    def kernel_roofline(self, kernel_ms, ms_per_step):
        peak, _, peak_src = peaks()
        alg = algorithmic_bytes(self.criterion, self.tokens, self.channels, self.N)
        achieved = alg / (kernel_ms * 1e-3) / 1e9
        rec = recorded_traffic(STREAM_KERNEL[self.criterion], self.N) if self.tokens == COCO_TOKENS else None
        return {'bound': 'hbm', 'kernel': STREAM_KERNEL[self.criterion], 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'frac_of_nominal_8000': achieved / 8000.0, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': alg, 'kernel_ms': kernel_ms,
                'kernel_share_of_step': kernel_ms / ms_per_step if ms_per_step else None,
                'traffic': (rec or {}).get('dram_bytes_per_launch'), 'traffic_source': (rec or {}).get('source')}


def measure_config(dev, rank, world, dist, images, num_prev, criterion, steps, warmup, levels=None, num_query=300,
                   flush=None, plain_graph=False):
    """Short measurement of one extra configuration: eager pass (kernel events) + graph replay; max over ranks."""
    sb = StepBench(dev, rank, world, dist, images, num_prev, criterion, levels, num_query, plain_graph)
    eager_ms, kernel_times, _, loss = sb.run_eager(max(3, steps // 2), warmup)
    kernel_ms = statistics.mean(kernel_times) if kernel_times else float('nan')
    try:
        ms, g_loss, _ = sb.run_graph(steps, warmup, flush)
        launch = 'cuda_graph_replay'
        if abs(g_loss - loss) > 1e-3 * abs(loss):
            raise RuntimeError(f'graph replay loss {g_loss} != eager loss {loss}')
    except Exception as exc:                        # noqa: BLE001 -- report, do not hide
        ms, launch = eager_ms / max(3, steps // 2), f'eager ({type(exc).__name__}: {exc})'
    ms, kernel_ms = sb.max_over_ranks(ms, kernel_ms)
    roof = sb.kernel_roofline(kernel_ms, ms)
    out = {'images_per_gpu': images, 'num_prev': num_prev, 'criterion': criterion, 'tokens': sb.tokens, 'queries': num_query,
           'value': world * images / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'launch': launch,
           'kernel_ms': kernel_ms, 'achieved_gbs': roof['achieved'], 'frac': roof['frac'], 'loss': loss,
           'l2': 'flushed between replays' if flush is not None else 'inputs larger than L2'}
    del sb
    gc.collect()
    torch.cuda.empty_cache()
    return out


def multi_rank_parity(sb, dist):
    """the W-rank result must equal the single-process evaluation of the concatenated batch.  Every rank
    evaluates BCDD with the prototype all-reduce; rank 0 gathers all ranks' embeddings / labels / keep-ids, evaluates
    BetweenClassDistanceLoss(sync_prototypes=False) on the concatenation and compares the loss (rtol 1e-4) and, for every
    rank, grad_hs / world against the matching slice (rtol 1e-3, atol 1e-6 * max)."""
    import dskd_b200
    inp, world, rank, dev = sb.inputs, sb.world, sb.rank, sb.dev
    a = inp.assignments
    hs = inp.hs_student.detach().clone().requires_grad_(True)
    loss = sb.bcdd(None, None, (hs, inp.hs_teacher), a)
    loss.backward()
    N, Q, C = hs.shape
    payload = dict(hs_s=hs.detach().cpu(), hs_t=inp.hs_teacher.cpu(), labels=a['student_labels'].cpu(),
                   keepid=a['teacher_keepid'].cpu(), tlabels=a['teacher_labels'].cpu(), grad=hs.grad.cpu(),
                   loss=float(loss.detach()))
    gathered = [None] * world
    dist.all_gather_object(gathered, payload)
    if rank != 0:
        return None
    cat = lambda k: torch.cat([g[k] for g in gathered]).to(dev)
    keep = torch.cat([g['keepid'] + r * N * Q for r, g in enumerate(gathered)]).to(dev)
    big = dict(student_labels=cat('labels'), teacher_keepid=keep, teacher_labels=cat('tlabels'),
               prev_labels=a['prev_labels'], num_classes=a.get('num_classes', 80))
    hs_all = cat('hs_s').requires_grad_(True)
    single = dskd_b200.build_loss(dict(type='BetweenClassDistanceLoss', reduction='mean', loss_weight=1.0, sync_prototypes=False))
    ref = single(None, None, (hs_all, cat('hs_t')), big)
    ref.backward()
    ref_loss = float(ref.detach())
    loss_err = max(abs(g['loss'] - ref_loss) / max(abs(ref_loss), 1e-30) for g in gathered)
    ref_grad = hs_all.grad.cpu()
    scale = float(ref_grad.abs().max())
    worst = 0.0
    for r, g in enumerate(gathered):
        got = g['grad'] / world                         # DDP's gradient mean
        want = ref_grad[r * N:(r + 1) * N]
        viol = ((got - want).abs() - (1e-3 * want.abs() + 1e-6 * scale)).max()
        worst = max(worst, float((got - want).abs().max()) / max(scale, 1e-30))
        if float(viol) > 0:
            return {'ok': False, 'loss_rel_err': loss_err, 'grad_max_err_over_max': worst, 'failed_rank': r}
    return {'ok': bool(loss_err <= 1e-4), 'loss_rel_err': loss_err, 'grad_max_err_over_max': worst, 'ranks': world,
            'what': 'synced BCDD (prototype exchange over the ranks, grad_scale = world) vs BetweenClassDistanceLoss on the '
                    'concatenated batch; loss rtol 1e-4, grad / world rtol 1e-3'}


def time_train_step(dev, rank, world, dist, images=4, criterion='kl', steps=3, warmup=2):
    """The 40+40 incremental training step (chaosuan_..._40_...py:24-152,214-238) with the shipped
    KL criterion: random-init GFL-Deformable-DETR R-50 student + frozen teacher, AdamW, the CUDA distillation path inside."""
    from dskd_b200.harness import bench_train_step
    return bench_train_step(dev, rank, world, dist, images_per_gpu=images, criterion=criterion, steps=steps, warmup=warmup)


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
        return
    import torch.distributed as dist
    from dskd_b200 import synth

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

    # the BCDD side stream (StepBench.step) is deliberate: autograd then sees the embeddings' AccumulateGrad node and the
    # BCDD backward on different streams and would warn about it on every step
    torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    N = args.images_per_gpu
    sb = StepBench(dev, rank, world, dist, N, args.num_prev, args.criterion, plain_graph=args.plain_graph)
    inputs, s_feats, hs_s, step = sb.inputs, sb.s_feats, sb.hs_s, sb.step

    # ---------------- device-resident throughput, eager launches (host-issue bound for this small step)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eager_ms, kernel_times, launches, loss_value = sb.run_eager(args.steps, args.warmup)

    # ---------------- the same step captured once in a CUDA graph and replayed (`value`)
    graph_ms = None
    graph_err = None
    graph_policy = None
    if not args.no_graph:
        try:
            ms, g_loss, graph_policy = sb.run_graph(args.steps, args.warmup)
            graph_ms = ms * args.steps
            if abs(g_loss - loss_value) > 1e-3 * abs(loss_value):
                raise RuntimeError(f'graph replay loss {g_loss} != eager loss {loss_value}')
        except Exception as exc:                    # noqa: BLE001 -- report, do not hide
            graph_err = f'{type(exc).__name__}: {exc}'
            graph_ms = None
    clocks = sampler.stop() if rank == 0 else None
    # the eager figure once more without `nvidia-smi -lms` polling the driver next to the launches (it lengthens them)
    eager_quiet_ms = sb.run_eager(args.steps, args.warmup, kernel_events=False)[0]
    elapsed_ms = graph_ms if graph_ms is not None else eager_ms
    elapsed_ms, eager_ms, eager_quiet_ms = sb.max_over_ranks(elapsed_ms, eager_ms, eager_quiet_ms)
    kernel_ms = statistics.mean(kernel_times) if kernel_times else float('nan')

    # ---------------- end-to-end through the module call with HOST buffers (`e2e`)
    e2e = None
    if not args.no_e2e:
        # the pinned staging buffers go on the host socket this GPU hangs off (first touch under that CPU affinity): at
        # 4-8 ranks the copies otherwise cross the socket interconnect; the affinity is restored right after
        affinity0 = os.sched_getaffinity(0)
        numa_node, numa_cpus = gpu_numa_cpus(local_rank)
        if numa_cpus and (affinity0 & numa_cpus):
            os.sched_setaffinity(0, affinity0 & numa_cpus)
        host_s = [torch.empty(f.shape, dtype=f.dtype, pin_memory=True).copy_(f.detach()) for f in s_feats]
        host_t = [torch.empty(f.shape, dtype=f.dtype, pin_memory=True).copy_(f) for f in inputs.teacher_feats]
        host_hs = torch.empty(hs_s.shape, pin_memory=True).copy_(hs_s.detach())
        host_ht = torch.empty(hs_s.shape, pin_memory=True).copy_(inputs.hs_teacher)
        host_loss = torch.empty((), pin_memory=True)
        os.sched_setaffinity(0, affinity0)
        h2d = sum(t_.numel() * 4 for t_ in host_s + host_t + [host_hs, host_ht])

        def upload():
            fs = [h.to(dev, non_blocking=True) for h in host_s]
            ft = [h.to(dev, non_blocking=True) for h in host_t]
            return fs, ft, host_hs.to(dev, non_blocking=True), host_ht.to(dev, non_blocking=True)

        def e2e_step():
            fs, ft, hs, ht = upload()
            l = step([f.requires_grad_(True) for f in fs], ft, hs.requires_grad_(True), ht)
            host_loss.copy_(l.detach(), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return float(host_loss)

        def h2d_only():
            upload()
            torch.cuda.current_stream().synchronize()

        def wall(fn, reps):
            for _ in range(2):
                fn()
            sb.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            sb.barrier()
            return sb.max_over_ranks(time.perf_counter() - t0)[0] / reps * 1e3
        e2e_steps = max(3, min(args.steps, 10))
        e2e_ms = wall(e2e_step, e2e_steps)
        # the same host->device copies with nothing else: the PCIe / host-memory roof of this rank count
        h2d_ms = wall(h2d_only, e2e_steps)
        e2e = {'value': world * N / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': 4, 'steps': e2e_steps, 'ms_per_step': e2e_ms, 'h2d_only_ms': h2d_ms,
               'h2d_gbs_per_rank': h2d / (h2d_ms * 1e-3) / 1e9, 'roof_frac': h2d_ms / e2e_ms,
               'roof_note': 'roof = the same pinned-host copies with no kernel behind them, all ranks at once; e2e / roof '
                            'shows what the step adds to the transfer',
               'host_buffers': (f'pinned, first touched on NUMA node {numa_node} (the socket of this GPU)'
                                if numa_cpus else 'pinned (no NUMA information on this box)')}
        del host_s, host_t, host_hs, host_ht

    contraction = None
    if not args.no_contraction:
        contraction = time_contraction(args, dev, N, world, dist)

    # ---------------- multi-rank parity of the only collective (outside every timed region)
    parity = multi_rank_parity(sb, dist) if world > 1 else None

    # ---------------- the shipped criterion: KL over H
    ms_per_step = elapsed_ms / args.steps
    extras = {}
    kernel_ms, = sb.max_over_ranks(kernel_ms)
    roofline = sb.kernel_roofline(kernel_ms, ms_per_step)
    del sb, s_feats, hs_s, inputs, step
    gc.collect()
    torch.cuda.empty_cache()
    if not args.no_kl and args.criterion != 'kl':
        kl = measure_config(dev, rank, world, dist, N, args.num_prev, 'kl', args.steps, args.warmup, plain_graph=args.plain_graph)
        peak, _, src = peaks()
        rec = recorded_traffic(STREAM_KERNEL['kl'], N)
        kl.update({'kernel': STREAM_KERNEL['kl'], 'peak': peak, 'peak_source': src,
                   'algorithmic_bytes_per_launch': algorithmic_bytes('kl', COCO_TOKENS, 256, N),
                   'traffic': (rec or {}).get('dram_bytes_per_launch'), 'traffic_source': (rec or {}).get('source'),
                   'note': 'KnowledgeDistillationKLDivLoss(T=2, sum) over H: the criterion of all four chaosuan_* configs '
                           '(:127); same step otherwise (decode_v1 + BCDD)'})
        extras['kl'] = kl

    # ---------------- BASELINE.json configs 3-5
    if not args.no_configs:
        cfg_steps = max(5, min(args.steps, 10))
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        configs = {}
        configs['coco_70+10_8img'] = measure_config(dev, rank, world, dist, 8, 70, 'kl', cfg_steps, 3)
        configs['coco_50+30_4img'] = measure_config(dev, rank, world, dist, 4, 50, 'kl', cfg_steps, 3)
        configs['coco_60+20_4img'] = measure_config(dev, rank, world, dist, 4, 60, 'kl', cfg_steps, 3)
        sweep = []
        for tokens in (5000, COCO_TOKENS, 40000):
            levels = synth.COCO_LEVELS if tokens == COCO_TOKENS else synth.scaled_levels(tokens)
            for q in (100, 300, 900):
                for n in (1, 16):
                    fl = flush if 2 * n * tokens * 256 * 4 < (160 << 20) else None
                    sweep.append(measure_config(dev, rank, world, dist, n, 40, 'mse', 5, 3, levels, q, fl))
        configs['sweep_mse'] = sweep
        if world == 1 and not args.no_cpu_baseline:
            lv = synth.scaled_levels(5000)
            base, _ = time_cpu_reference('mse', 40, 1, steps=3, warmup=1, levels=lv, num_query=100)
            configs['sweep_cpu_port_smallest'] = dict(base, tokens=sum(h * w for h, w in lv), queries=100, images=1)
        configs['note'] = ('configs 3-5 of BASELINE.json: L = 70 at samples_per_gpu = 8 (chaosuan_..._70_...py:202), L = 50 / 60 '
                           'at 4 images per GPU, KL criterion (shipped), prototype all-reduce on at n_gpus > 1; sweep: tokens x '
                           'queries x images per GPU, masked MSE + BCDD, graph replay (L2 flushed where the inputs fit it)')
        extras['configs'] = configs
        del flush

    if parity is not None:
        extras['multi_rank_parity'] = parity

    # ---------------- the 40+40 training step hosting the losses
    if not args.no_train_step:
        try:
            extras['train_step'] = time_train_step(dev, rank, world, dist)
        except Exception as exc:                    # noqa: BLE001 -- report, do not hide
            extras['train_step'] = {'error': f'{type(exc).__name__}: {exc}'}

    if rank != 0:
        leave(world, dist)
        return

    line = {
        'metric': METRIC, 'value': world * N * args.steps / (elapsed_ms * 1e-3), 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD if args.criterion == 'mse' else WORKLOAD.replace('mse', 'kl'), 'images_per_gpu': N,
                   'global_batch': world * N, 'num_prev': args.num_prev,
                   'criterion': args.criterion, 'levels': [[100, 167], [50, 84], [25, 42], [13, 21]], 'channels': 256,
                   'queries': 300, 'parallelism': f'dp{world}', 'prototype_allreduce': prototype_transport(world),
                   'launch': 'cuda_graph_replay' if graph_ms is not None else 'eager',
                   'graph_policy': graph_policy if graph_ms is not None else None,
                   'cache': f'inputs larger than L2: {2 * N * 22.76:.0f} MB of features read per step vs 126 MB L2'},
        'roofline': roofline,
        'contraction': contraction,
        'e2e': e2e,
        'gpu_launches': int(launches),
        'eager': {'value': world * N * args.steps / (eager_quiet_ms * 1e-3), 'unit': UNIT,
                  'ms_per_step': eager_quiet_ms / args.steps, 'ms_per_step_under_clock_sampler': eager_ms / args.steps,
                  'note': 'the two module calls + backward issued from Python on one stream (no graph, no side stream); '
                          'host-issue bound; measured again after the nvidia-smi clock sampler has stopped (its driver '
                          'polling lengthens every launch) and without the per-step CUDA event pair that times the '
                          'streaming kernel for `roofline`', 'graph_error': graph_err},
        'clocks': clocks,
        'loss': loss_value,
        'parity_note': PARITY_NOTE,
    }
    line.update(extras)
    if world == 1 and not args.no_feeders:
        line['feeders'] = time_path_feeders(args, dev, N)
    if world == 1 and not args.no_cpu_baseline:
        base, _ = time_cpu_reference(args.criterion, args.num_prev, args.cpu_sample_images, steps=3, warmup=1)
        line['cpu_baseline'] = base
    print(json.dumps(line), flush=True)
    leave(world, dist)


def prototype_transport(world):
    """How the BCDD prototype tables were summed over the ranks in this run."""
    if world <= 1:
        return False
    from dskd_b200 import peer
    return ('one kernel over NVLink peer memory (dskd_peer_allreduce)' if peer._exchanges
            else 'NCCL all-reduce (peer memory unavailable or DSKD_PROTO_TRANSPORT=nccl)')


def leave(world, dist):
    """End of a multi-rank run: every CUDA graph that captured the NCCL all-reduce has been dropped by now (StepBench.
    run_graph deletes its exec and torch graph), so the process group can be destroyed normally.  A watchdog still ends the
    process if the teardown should block (seen in round 1 while graph execs holding NCCL kernels were alive)."""
    if world <= 1:
        return
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    from dskd_b200 import peer
    peer.close_all()            # unmap the peers' NVLink buffers while everybody is still alive
    dist.barrier()
    sys.stdout.flush()
    sys.stderr.flush()

    def watchdog():
        time.sleep(30)
        sys.stderr.write('bench.py: destroy_process_group() did not return within 30 s; exiting\n')
        sys.stderr.flush()
        os._exit(0)
    threading.Thread(target=watchdog, daemon=True).start()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
