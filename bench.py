#!/usr/bin/env python
"""bench.py -- EGT attention-block forward+backward graphs/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C0|C0e64|C1|C2|C3|C4|C5]
    python bench.py --impl reference ...      # CPU arm: the oracle port of the reference TF path

A "step" = one EGTBlock forward + backward over one batch of synthetic dense graphs (per GPU:
B graphs of N nodes) plus, for N>1 ranks, the single all-reduce of the flat weight gradient.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# per-GPU workloads
WORKLOADS = {
    'C0':    dict(N=128, d=64, d_e=8, h=8, B=128, note='headline: N=128, h=8, bf16, d=64, d_e=8'),
    'C0e64': dict(N=128, d=64, d_e=64, h=8, B=128, note='headline with ZINC-like edge width'),
    'C1':    dict(N=37, d=64, d_e=64, h=8, B=128, note='ZINC'),
    'C2':    dict(N=75, d=64, d_e=8, h=8, B=128, note='MNIST'),
    'C3':    dict(N=190, d=96, d_e=8, h=8, B=64, note='CLUSTER d=96'),
    'C4':    dict(N=188, d=64, d_e=8, h=8, B=128, note='PATTERN'),
    'C5':    dict(N=512, d=128, d_e=32, h=16, B=64, note='synthetic roofline sweep'),
    'S64':   dict(N=64, d=64, d_e=8, h=8, B=128, note='sweep N=64'),
    'S256':  dict(N=256, d=64, d_e=8, h=8, B=128, note='sweep N=256'),
    'S512':  dict(N=512, d=64, d_e=8, h=8, B=64, note='sweep N=512'),
    # the N sweep at the widths of BASELINE config 5 (d=128, d_e=32, h=16), B/GPU = 64
    'W64':   dict(N=64, d=128, d_e=32, h=16, B=64, note='sweep N=64 at C5 widths'),
    'W128':  dict(N=128, d=128, d_e=32, h=16, B=64, note='sweep N=128 at C5 widths'),
    'W256':  dict(N=256, d=128, d_e=32, h=16, B=64, note='sweep N=256 at C5 widths'),
    'C1s':   dict(N=37, d=48, d_e=48, h=8, B=128, note='ZINC 100K widths (dk=6): staged kernels'),
}
METRIC = 'EGT-layer fwd+bwd graphs/sec'


def alg_bytes_per_graph(N, d, d_e, h, s=2):
    """fwd+bwd algorithmic HBM bytes per graph."""
    return s * (5 * N * N * d_e + 6 * N * d) + 16 * N * h + N


def alg_bytes_fwd(N, d, d_e, h, s=2):
    return s * (2 * N * N * d_e + 2 * N * d) + 8 * N * h + N


def alg_bytes_bwd(N, d, d_e, h, s=2):
    return s * (3 * N * N * d_e + 4 * N * d) + 8 * N * h


def alg_flops_per_graph(N, d, d_e, h):
    F = 2 * N * d * 3 * d + 2 * N * d * d + 4 * N * N * d + 4 * N * N * d_e * h + 2 * N * N * h * d_e
    return 3 * F


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm_gbs=j['hbm_gbs'], bf16_tflops=j['bf16_tflops'], source='measured')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source='fallback')


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            hd = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(hd, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksEventReasonHwSlowdown: 'hw_slowdown',
                     nv.nvmlClocksEventReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                     nv.nvmlClocksEventReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                     nv.nvmlClocksEventReasonSwPowerCap: 'sw_power_cap'}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(hd, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(hd)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.002)
        except Exception as ex:   # NVML missing: report that instead of inventing clocks
            self.reasons.add(f'nvml_unavailable:{type(ex).__name__}')

    def summary(self):
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(s))


# ------------------------------------------------------------------------------------------
def cpu_reference_run(w, steps, warmup, sample_graphs=None, budget_s=20.0):
    """Times the oracle (op-for-op CPU restatement of the reference TF path, fp32, eager, every
    [B,N,N,h] intermediate materialised, backward by autograd) on the host cores."""
    from oracle import egt_oracle as O
    torch.set_num_threads(os.cpu_count())
    N, d, d_e, h = w['N'], w['d'], w['d_e'], w['h']
    Bs = sample_graphs or max(1, min(w['B'], int(2.0e8 // (N * N * h * 16))))
    cfg = O.BlockConfig(model_width=d, edge_width=d_e, num_heads=h)
    params = {k: v.requires_grad_(True) for k, v in O.init_block_params(cfg).items()}
    hh, ee, mask = O.synthetic_batch(Bs, N, d, d_e)
    g = torch.Generator().manual_seed(7)
    dh, de = torch.randn(hh.shape, generator=g), torch.randn(ee.shape, generator=g)

    def step():
        hr, er = hh.clone().requires_grad_(True), ee.clone().requires_grad_(True)
        h2, e2 = O.egt_block(hr, er, mask, params, cfg)
        torch.autograd.grad([h2, e2], [hr, er] + list(params.values()), [dh, de])

    for _ in range(max(1, warmup)):
        step()
    t0 = time.perf_counter()
    n = 0
    while n < steps and (n < 1 or time.perf_counter() - t0 < budget_s):
        step()
        n += 1
    dt = time.perf_counter() - t0
    return dict(value=Bs * n / dt, unit='graphs/s', cores=torch.get_num_threads(), kind='port',
                sample=f'{n} steps x {Bs} graphs of N={N} (fp32 oracle port of the reference TF path; '
                       f'TensorFlow is not installable in this image)', ms_per_step=1e3 * dt / n, steps=n)


def run_reference(args, w):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    r = cpu_reference_run(w, args.steps, args.warmup)
    line = dict(impl='reference', metric=METRIC, value=r['value'], unit='graphs/s', n_gpus=args.gpus,
                steps=r['steps'], warmup=args.warmup, ms_per_step=r['ms_per_step'], higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload=args.workload, **{k: w[k] for k in ('N', 'd', 'd_e', 'h')},
                            graphs_per_gpu=w['B'], global_batch=w['B'] * args.gpus, parallelism=f'dp{args.gpus}',
                            note=w['note'], random_mask_prob=args.random_mask_prob, scale_degree=bool(args.scale_degree),
                            path='cpu-oracle-port'),
                cpu_baseline=dict(value=r['value'], unit='graphs/s', cores=r['cores'], kind=r['kind'],
                                  sample=r['sample']),
                e2e=dict(value=r['value'], unit='graphs/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def run_ours(args, w):
    import torch.distributed as dist
    import egt_b200
    from egt_b200 import _lib as L
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = L.load()
    N, d, d_e, h, B = w['N'], w['d'], w['d_e'], w['h'], w['B']
    training = bool(args.random_mask_prob > 0)
    blk = egt_b200.EGTBlock(model_width=d, edge_width=d_e, num_heads=h, scale_degree=bool(args.scale_degree),
                            random_mask_prob=args.random_mask_prob, seed=1000 + rank).to(dev)
    blk.train(training)
    # synthetic inputs: per-rank seed, N(0,1) activations, dense graphs (num_nodes = N)
    nsets = args.input_sets
    g = torch.Generator(device='cpu').manual_seed(20240 + rank)
    host = []
    for _ in range(nsets):
        hh = torch.randn(B, N, d, generator=g).bfloat16().pin_memory()
        ee = torch.randn(B, N, N, d_e, generator=g).bfloat16().pin_memory()
        mm = torch.ones(B, N, dtype=torch.bool).pin_memory()
        host.append((hh, ee, mm))
    devs = [tuple(t.to(dev) for t in s) for s in host]
    # upstream gradients rotate with the input sets (a single buffer would stay warm in the 126 MB L2)
    ups = [(torch.randn(B, N, d, generator=g).bfloat16().to(dev), torch.randn(B, N, N, d_e, generator=g).bfloat16().to(dev))
           for _ in range(nsets)]
    grad_host = torch.empty(blk.flat.numel(), dtype=torch.float32).pin_memory()

    def step(i, inputs=None, block=None):
        block = block or blk
        hh, ee, mm = inputs if inputs is not None else devs[i % nsets]
        dh, de = ups[i % nsets]
        hh = hh.detach().requires_grad_(True)
        ee = ee.detach().requires_grad_(True)
        block.flat.grad = None
        h2, e2 = block(hh, ee, mm)
        torch.autograd.backward([h2, e2], [dh, de])
        egt_b200.allreduce_flat_grads([block])
        return hh.grad, ee.grad

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(max(3, args.warmup)):
        step(i)
    torch.cuda.synchronize()

    # One step = one forward + backward (+ all-reduce) of the block.  The launch sequence of a step is fixed,
    # so each of the rotating input sets is captured once into a CUDA graph and the timed region replays
    # the graphs (one replay = one step); --no-graphs times the eager launches instead.
    graphs, static_out, use_graphs = [], [], not args.no_graphs
    launches_per_step = None
    if use_graphs:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(2):
                    step(i)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for i in range(nsets):
                g_ = torch.cuda.CUDAGraph()
                l0 = lib.egt_launch_count()
                with torch.cuda.graph(g_):
                    out = step(i)
                launches_per_step = lib.egt_launch_count() - l0
                graphs.append(g_)
                static_out.append((out, blk.flat.grad))
            torch.cuda.synchronize()
        except Exception as ex:   # capture not possible: measure the eager path and say so
            print(f'[bench] CUDA graph capture failed ({type(ex).__name__}: {ex}); timing eager launches', file=sys.stderr)
            graphs, static_out, use_graphs = [], [], False
            torch.cuda.synchronize()

    def run_step(i):
        if use_graphs:
            graphs[i % nsets].replay()
        else:
            step(i)

    for i in range(3):
        run_step(i)

    # ---- timed region 1: inputs resident in HBM ----
    sampler = ClockSampler(local)
    sampler.start()
    l0 = lib.egt_launch_count()
    ms = timed(run_step, args.steps)
    launches = (launches_per_step * args.steps) if use_graphs else lib.egt_launch_count() - l0

    # ---- per-kernel CUDA-event times of the same steps, launched eagerly (events cannot bracket graph nodes) ----
    lib.egt_profile_enable(1)
    timed(step, args.steps)
    prof = L.profile_read()
    lib.egt_profile_enable(0)
    path = lib.egt_last_path()

    # ---- timed region 2: end to end through the public API with HOST buffers ----
    # Every step copies its inputs from pinned host memory (on a copy stream, one step ahead of the compute
    # stream, into one of two device buffer sets) and reads the step's flat weight gradient back to the host.
    cur = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream()
    bufs = [tuple(torch.empty_like(t, device=dev) for t in host[0]) for _ in range(2)]
    e2e_graphs = []
    if use_graphs:
        for k in range(2):
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                step(0, bufs[k])
            e2e_graphs.append((g_, blk.flat.grad))
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_done[k])
            for dst, src in zip(bufs[k], host[i % nsets]):
                dst.copy_(src, non_blocking=True)
            ev_copied[k].record(copy_stream)

    def e2e_loop(steps):
        for k in range(2):
            ev_done[k].record(cur)
        prefetch(0)
        for i in range(steps):
            k = i % 2
            if i + 1 < steps:
                prefetch(i + 1)
            cur.wait_event(ev_copied[k])
            if use_graphs:
                e2e_graphs[k][0].replay()
                grad = e2e_graphs[k][1]
            else:
                step(i, bufs[k])
                grad = blk.flat.grad
            ev_done[k].record(cur)
            grad_host.copy_(grad, non_blocking=True)
            cur.synchronize()                          # the caller consumes the step's result on the host

    e2e_loop(3)

    def timed_e2e(steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_loop(steps)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_e2e = timed_e2e(args.steps)

    # ---- the shipped training setting (every reference config: random_mask_prob = 0.1, training = True) ----
    # The key mask's Philox offset = a launch argument + a device counter the captured step bumps (egt_b200.layers), so
    # the step replays from a CUDA graph and still draws a new mask every time; the eager time is reported next to it.
    # (N > 1: eager only -- NCCL work issued before a capture was seen to invalidate it, and the timed regions above use NCCL.)
    train_line = None
    if args.random_mask_prob == 0 and not args.no_train_line:
        blk_t = egt_b200.EGTBlock(model_width=d, edge_width=d_e, num_heads=h, scale_degree=bool(args.scale_degree),
                                  random_mask_prob=0.1, seed=2000 + rank).to(dev)
        blk_t.train(True)
        for i in range(3):
            step(i, block=blk_t)
        ms_eager = timed(lambda i: step(i), args.steps)
        ms_train_eager = timed(lambda i: step(i, block=blk_t), args.steps)
        ms_train, tgraphs, tkeep = ms_train_eager, [], []
        if use_graphs and world == 1:
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for i in range(2):
                        step(i, block=blk_t)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                for i in range(nsets):
                    g_ = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g_):
                        tkeep.append((step(i, block=blk_t), blk_t.flat.grad))
                    tgraphs.append(g_)
                for i in range(3):
                    tgraphs[i % nsets].replay()
                ms_train = timed(lambda i: tgraphs[i % nsets].replay(), args.steps)
            except Exception as ex:
                print(f'[bench] training-step graph capture failed ({type(ex).__name__}: {ex}); eager time reported', file=sys.stderr)
                tgraphs = []
                torch.cuda.synchronize()
        train_line = dict(random_mask_prob=0.1, training=True, cuda_graphs=bool(tgraphs), ms_per_step=ms_train / args.steps,
                          value=B * world * args.steps / (ms_train / 1e3), unit='graphs/s',
                          eager_ms_per_step=ms_train_eager / args.steps,
                          eager_ms_per_step_without_mask=ms_eager / args.steps)
        del tgraphs, tkeep
    # ---- one FULL layer: attention block + node FFN + edge FFN, forward + backward, same inputs ----
    # (graph_xformer_model_base.py:335-341: edge_update then ffn_block); CUDA graphs like the headline step
    layer_line = None
    if not args.no_layer_line and args.random_mask_prob == 0 and world == 1:   # a single-GPU diagnostic line
        lay = egt_b200.EGTStack(1, ffn=True, model_width=d, edge_width=d_e, num_heads=h, scale_degree=bool(args.scale_degree),
                                seed=3000 + rank).to(dev)
        lay.train(False)

        def layer_step(i):
            hh, ee, mm = devs[i % nsets]
            dh, de = ups[i % nsets]
            hh = hh.detach().requires_grad_(True)
            ee = ee.detach().requires_grad_(True)
            lay.flat.grad = None
            h2, e2 = lay(hh, ee, mm)
            torch.autograd.backward([h2, e2], [dh, de])
            return hh.grad, ee.grad

        for i in range(3):
            layer_step(i)
        torch.cuda.synchronize()
        lgraphs, keep = [], []
        if use_graphs:
            try:
                side = torch.cuda.Stream()             # warm up on a side stream first, as torch.cuda.graph asks
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for i in range(2):
                        layer_step(i)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                for i in range(nsets):
                    g_ = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g_):
                        keep.append((layer_step(i), lay.flat.grad))
                    lgraphs.append(g_)
                torch.cuda.synchronize()
            except Exception as ex:
                print(f'[bench] full-layer graph capture failed ({type(ex).__name__}: {ex}); timing eager launches', file=sys.stderr)
                lgraphs = []
                torch.cuda.synchronize()
        run_layer = (lambda i: lgraphs[i % nsets].replay()) if lgraphs else layer_step
        for i in range(3):
            run_layer(i)
        ms_layer = timed(run_layer, args.steps)
        lib.egt_profile_enable(1)
        timed(layer_step, args.steps)
        lprof = L.profile_read()
        lib.egt_profile_enable(0)
        layer_line = dict(what='attention block + node FFN + edge FFN, forward + backward', cuda_graphs=bool(lgraphs),
                          ms_per_step=ms_layer / args.steps, value=B * world * args.steps / (ms_layer / 1e3), unit='graphs/s',
                          kernels_ms={k: v[0] / max(1, v[1]) for k, v in lprof.items()})
        del lgraphs, keep
    allreduce_check = None
    # ---- self-check of the gradient all-reduce (N > 1): the peer-memory kernel against NCCL on the same buffer ----
    # (after every captured / timed region: NCCL work issued before a CUDA-graph capture was seen to invalidate it)
    if world > 1 and os.environ.get('EGT_BENCH_NO_ARCHECK', '0') != '1':
        step(0)
        mine = blk.flat.grad.clone()                   # already summed over ranks by the step
        blk.flat.grad = None
        hh, ee, mm = devs[0]
        hh = hh.detach().requires_grad_(True); ee = ee.detach().requires_grad_(True)
        h2, e2 = blk(hh, ee, mm)
        torch.autograd.backward([h2, e2], list(ups[0]))
        ref = blk.flat.grad.clone()                    # this rank's gradient, not reduced
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        err = float((mine - ref).abs().max())
        scale = float(ref.abs().max())
        ok = err <= 1e-3 * max(scale, 1e-6)            # weight gradients accumulate with atomics: summation order differs
        flag = torch.tensor([0 if ok else 1], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        allreduce_check = dict(max_abs_err=err, max_abs_ref=scale, ok=bool(flag.item() == 0), ranks=world)
        if not allreduce_check['ok']:
            if rank == 0:
                print(json.dumps(dict(error='gradient all-reduce mismatch', allreduce_check=allreduce_check)), flush=True)
            os._exit(3)

    sampler.stop_flag = True
    sampler.join(timeout=2)

    def finish():
        """Leave without tearing the process group down: destroying it while captured CUDA graphs still hold
        NCCL work has been seen to hang; every rank has already passed the final barrier."""
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)

    if rank != 0:
        finish()
        return
    peaks = measured_peaks()
    graphs = B * world * args.steps
    value = graphs / (ms / 1e3)
    # dominant kernel = the one with the largest total time in the profiled region
    dom = max(prof.items(), key=lambda kv: kv[1][0]) if prof else (None, (0.0, 0))
    total_prof_ms = sum(v[0] for v in prof.values()) or 1.0
    kb = {'fwd': alg_bytes_fwd(N, d, d_e, h), 'bwd': alg_bytes_bwd(N, d, d_e, h)}
    dom_name = dom[0] or ''
    per_launch_bytes = B * (kb['fwd'] if 'fwd' in dom_name else kb['bwd'])
    dom_avg_ms = dom[1][0] / max(1, dom[1][1])
    achieved = per_launch_bytes / (dom_avg_ms * 1e-3) / 1e9 if dom_avg_ms > 0 else 0.0
    step_bytes = B * alg_bytes_per_graph(N, d, d_e, h)
    step_gbs = step_bytes * args.steps / (ms * 1e-3) / 1e9          # per GPU: step_bytes are one rank's bytes
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get(f'{args.workload}:{dom_name}', tj.get(dom_name) if args.workload == 'C0' else None)
    cpu = cpu_reference_run(w, steps=1000, warmup=1, budget_s=args.cpu_budget) if world == 1 and not args.no_cpu else None
    line = dict(
        metric=METRIC, value=value, unit='graphs/s', n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
        ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16',
        data='synthetic',
        config=dict(workload=args.workload, N=N, d=d, d_e=d_e, h=h, graphs_per_gpu=B, global_batch=B * world,
                    parallelism=f'dp{world}', note=w['note'], random_mask_prob=args.random_mask_prob,
                    scale_degree=bool(args.scale_degree),
                    l2=f'rotating {nsets} input sets per rank (working set > 126 MB L2)',
                    path='fused-tcgen05' if path == 1 else 'staged', cuda_graphs=bool(use_graphs),
                    e2e='one TRAINING step per iteration: H2D of h,e,mask on a copy stream one step ahead (2 device buffer sets); the result read back is the flat weight gradient (what an optimiser step consumes) + host sync every step'),
        roofline=dict(bound='hbm', kernel=dom_name, achieved=achieved, peak=peaks['hbm_gbs'], unit='GB/s',
                      frac=achieved / peaks['hbm_gbs'], traffic=traffic, peak_source=peaks['source'],
                      kernel_share_of_step=dom[1][0] / total_prof_ms,
                      step_achieved=step_gbs, step_frac=step_gbs / peaks['hbm_gbs'],
                      kernels={k: dict(ms_total=v[0], launches=v[1]) for k, v in prof.items()}),
        e2e=dict(value=graphs / (ms_e2e / 1e3), unit='graphs/s',
                 h2d_bytes_per_step=sum(t.numel() * t.element_size() for t in host[0]),
                 d2h_bytes_per_step=grad_host.numel() * 4, ms_per_step=ms_e2e / args.steps),
        gpu_launches=int(launches), clocks=sampler.summary(),
    )
    if train_line is not None:
        line['training_step'] = train_line
    if layer_line is not None:
        line['full_layer'] = layer_line
    if allreduce_check is not None:
        line['allreduce_check'] = allreduce_check
    if cpu is not None:
        line['cpu_baseline'] = dict(value=cpu['value'], unit='graphs/s', cores=cpu['cores'], kind=cpu['kind'],
                                    sample=cpu['sample'])
    print(json.dumps(line), flush=True)
    finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='C0', choices=sorted(WORKLOADS))
    ap.add_argument('--random-mask-prob', type=float, default=0.0)
    ap.add_argument('--scale-degree', type=int, default=0)
    ap.add_argument('--input-sets', type=int, default=4)
    ap.add_argument('--cpu-budget', type=float, default=15.0)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-graphs', action='store_true')
    ap.add_argument('--no-train-line', action='store_true')
    ap.add_argument('--no-layer-line', action='store_true', help='skip the full-layer (block + FFN) measurement')
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == 'reference':
        return run_reference(args, w)
    return run_ours(args, w)


if __name__ == '__main__':
    main()
