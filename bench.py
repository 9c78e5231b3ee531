#!/usr/bin/env python
"""ROIAlign fwd+bwd throughput on B200 (BASELINE.json metric) -- one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2]

A "step" is one forward bilinear pool plus one backward gradient scatter (dX produced from
scratch, zero-fill included) over one batch of synthetic inputs already resident in HBM.
The default workload is BASELINE.json configs[3] (C4: 64 images of the Faster R-CNN box-head
shape, 32768 ROIs) for EVERY N: the 64 images are split over the N ranks with
mobulaop_b200.sharding (per-image ROI ownership, no data-path collective) = strong scaling;
N=1 runs all of it on one GPU (it fits: 3.6 GB of feature maps).  The C2 box head (configs[1],
what round 1 was quoted on) is timed per rank as the `secondary` record.  `--scaling weak
--workload C2` is the round-1 protocol (every rank owns 2 images).

value      ROIs/s over all ranks, device time (CUDA events on the launching stream), max over ranks
e2e        the same metric through mobula.op.ROIAlign[np.ndarray] with pinned HOST buffers:
           host->device and device->host copies inside the timed region
e2e_torch  the same metric through mobula.op.ROIAlign on CUDA tensors (torch autograd), wall clock
roofline   dominant kernel: algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json
cpu_baseline  the reference's own C++ (oracle/_ref, OpenMP, all host cores) on rank 0, N=1

`--impl reference` times only the reference CPU implementation (rank 0; other ranks exit).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mobulaop_b200 import workloads as WL     # noqa: E402

METRIC = 'roialign_fwd_bwd_rois_per_sec'
UNIT = 'ROIs/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='C4', choices=sorted(WL.CONFIGS))
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'])
    ap.add_argument('--no-secondary', action='store_true')
    ap.add_argument('--e2e-torch-steps', type=int, default=200)
    ap.add_argument('--mode', default='atomic', choices=['atomic', 'deterministic'])
    ap.add_argument('--algo', default='auto', choices=['auto', 'gather', 'plane', 'resident', 'sweep', 'pull'])
    ap.add_argument('--top-dtype', default='float32', choices=['float32', 'float16', 'bfloat16'],
                    help='storage type of the pooled tensor (output / dy); the arithmetic is fp32')
    ap.add_argument('--layout', default='RCHW', choices=['RCHW', 'RHWC'],
                    help='layout of the pooled tensor: [R,C,PH,PW] (reference) or channel-last [R,PH,PW,C]')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--e2e-steps', type=int, default=5)
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (of measured)'
    return 6650.0, 'B200_PROFILING.md fallback 6.65 TB/s (of fallback)'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True,
                                     timeout=5).stdout.strip().split('\n')[0]
                parts = [p.strip() for p in out.split(',')]
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for name, val in zip(names, parts[2:]):
                    if val.lower().startswith('active'):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=10)

    def summary(self):
        return dict(sm_mhz=float(np.median(self.samples)) if self.samples else None,
                    sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(self.samples))


# ---------------------------------------------------------------------------------------
# reference arm / CPU baseline (the only place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------
def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core."""
    import ctypes
    n = os.cpu_count() or 1
    os.environ['OMP_NUM_THREADS'] = str(n)
    try:
        gomp = ctypes.CDLL('libgomp.so.1')
        gomp.omp_set_num_threads(ctypes.c_int(n))
        return int(gomp.omp_get_max_threads())
    except OSError:
        return n


def reference_library():
    from oracle import roialign as O
    threads = use_all_host_cores()
    if O.RefLib.available('omp'):
        ref = O.RefLib('omp')
        return 'reference', threads, ref.forward, ref.backward
    orc = O.COracle()
    threads = orc.max_threads()
    return ('port', threads,
            lambda d, r, p, s, sr: orc.forward(d, r, p, s, sr, threads=threads),
            lambda dy, r, shp, s, sr: orc.backward(dy, r, shp, s, sr, threads=threads))


def cpu_sample(wl, budget_s):
    """A bounded sample of `wl` for the CPU arm: all images and ROIs, a channel subset
    sized so one step stays within `budget_s` (cost is linear in channels)."""
    kind, cores, fwd, bwd = reference_library()
    probe_c = min(wl.C, 8)
    probe = wl.scaled(C=probe_c)
    data, rois, dy = WL.make_inputs(probe, seed=0)
    t0 = time.perf_counter()
    fwd(data, rois, probe.pooled, probe.scale, probe.sampling_ratio)
    bwd(dy, rois, data.shape, probe.scale, probe.sampling_ratio)
    per_channel = (time.perf_counter() - t0) / probe_c
    channels = int(max(1, min(wl.C, budget_s / max(per_channel, 1e-9))))
    if channels >= wl.C:
        channels = wl.C
    return kind, cores, fwd, bwd, wl.scaled(C=channels)


def time_cpu(wl, steps, warmup, budget_s):
    kind, cores, fwd, bwd, sample = cpu_sample(wl, budget_s)
    data, rois, dy = WL.make_inputs(sample, seed=0)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        y = fwd(data, rois, sample.pooled, sample.scale, sample.sampling_ratio)
        y.sum()                                  # materialise, like the reference's asnumpy()
        bwd(dy, rois, data.shape, sample.scale, sample.sampling_ratio)   # includes dX zero-fill
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    frac = sample.C / float(wl.C)
    step_s = float(np.mean(times)) / frac        # scaled linearly to the full channel count
    desc = ('%s: all %d images and %d ROIs, %d of %d channels (cost scaled linearly by %.3g), '
            '%d timed repetitions after %d warm-up' % (wl.name, wl.N, wl.R, sample.C, wl.C, 1 / frac,
                                                         steps, warmup))
    return dict(kind=kind, cores=cores, sample=desc, step_s=step_s, value=wl.R / step_s)


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    wl = WL.CONFIGS[args.workload]
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    budget = max(0.05, min(2.0, 150.0 / (steps + warmup)))
    res = time_cpu(wl, steps, warmup, budget)
    line = dict(
        impl='reference', metric=METRIC, value=res['value'], unit=UNIT, n_gpus=args.gpus,
        steps=steps, warmup=warmup, ms_per_step=res['step_s'] * 1e3, higher_is_better=True,
        scaling=args.scaling, vs_baseline=None, dtype='f32', data='synthetic',
        config=wl.describe(),
        cpu_baseline=dict(value=res['value'], unit=UNIT, cores=res['cores'], kind=res['kind'],
                          sample=res['sample']),
        e2e=dict(value=res['value'], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        gpu_launches=0,
    )
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
class DeviceRun(object):
    """One rank's share of a workload, resident in HBM, launched through the C-ABI."""

    def __init__(self, lib, torch, dev, local_rank, wl, rois_np, mode, algo, seed, top_dtype='float32',
                 layout='RCHW'):
        import ctypes
        from mobulaop_b200 import _native
        self.lib, self.torch, self.dev, self.wl = lib, torch, dev, wl
        self._native = _native
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        if wl.literal:      # benchmark_roi_align.py:15: data = arange
            self.data = torch.arange(int(np.prod(wl.data_shape)), dtype=torch.float32,
                                     device=dev).reshape(wl.data_shape)
        else:
            self.data = torch.randn(wl.data_shape, dtype=torch.float32, device=dev, generator=gen)
        tdt = getattr(torch, top_dtype)
        oshape = wl.out_shape if layout == 'RCHW' else (wl.R,) + tuple(wl.pooled) + (wl.C,)
        self.dy = torch.randn(oshape, dtype=torch.float32, device=dev, generator=gen).to(tdt)
        self.rois = torch.from_numpy(rois_np).to(dev)
        self.out = torch.empty(oshape, dtype=tdt, device=dev)
        self.top = (dict(float32=0, float16=1, bfloat16=2)[top_dtype], dict(RCHW=0, RHWC=1)[layout])
        self.top_bytes = 4 if top_dtype == 'float32' else 2
        self.dx = torch.empty(wl.data_shape, dtype=torch.float32, device=dev)
        self.stream = torch.cuda.current_stream(dev)
        self.sp = ctypes.c_void_p(self.stream.cuda_stream)
        self.mode = _native.BWD_DETERMINISTIC if mode == 'deterministic' else _native.BWD_ATOMIC
        self.flags = _native.make_flags(algo=algo)
        self.local_rank = local_rank
        self._vp = lambda t: ctypes.c_void_p(t.data_ptr())

    def fwd(self):
        wl, vp = self.wl, self._vp
        self._native.check(self.lib.mobula_roialign_forward_ex(
            self.local_rank, self.sp, vp(self.data), vp(self.rois), vp(self.out), self.top[0], self.top[1],
            wl.R, wl.N, wl.C,
            wl.H, wl.W, wl.pooled[0], wl.pooled[1], wl.scale, wl.sampling_ratio, self.flags))

    def bwd(self):
        wl, vp = self.wl, self._vp
        self._native.check(self.lib.mobula_roialign_backward_ex(
            self.local_rank, self.sp, vp(self.dy), self.top[0], self.top[1], vp(self.rois), vp(self.dx),
            wl.R, wl.N, wl.C,
            wl.H, wl.W, wl.pooled[0], wl.pooled[1], wl.scale, wl.sampling_ratio, self.mode, self.flags))

    def free(self):
        self.data = self.dy = self.out = self.dx = self.rois = None
        self.torch.cuda.empty_cache()


def timed_steps(run, flush_l2, barrier, steps, warmup, clocks_index=None):
    """W untimed + K timed fwd+bwd steps; CUDA events on the launching stream, L2 flushed between
    steps outside the events.  Returns per-rank means in ms and the launch count."""
    torch, native = run.torch, run._native
    for _ in range(warmup):
        flush_l2()
        run.fwd()
        run.bwd()
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    launches0 = native.launch_count()
    sampler = ClockSampler(clocks_index) if clocks_index is not None else None
    if sampler:
        sampler.__enter__()
    barrier()
    wall0 = time.perf_counter()
    for k in range(steps):
        flush_l2()                          # evict L2 between steps (outside the events)
        ev[k][0].record(run.stream)
        run.fwd()
        ev[k][1].record(run.stream)
        run.bwd()
        ev[k][2].record(run.stream)
    barrier()
    wall = time.perf_counter() - wall0
    if sampler:
        sampler.__exit__()
    fwd_ms = np.array([ev[k][0].elapsed_time(ev[k][1]) for k in range(steps)])
    bwd_ms = np.array([ev[k][1].elapsed_time(ev[k][2]) for k in range(steps)])
    return dict(step_ms=float((fwd_ms + bwd_ms).mean()), fwd_ms=float(fwd_ms.mean()),
                bwd_ms=float(bwd_ms.mean()), launches=native.launch_count() - launches0, wall=wall,
                spread=dict(fwd_ms_min=float(fwd_ms.min()), fwd_ms_median=float(np.median(fwd_ms)),
                            fwd_ms_max=float(fwd_ms.max()), bwd_ms_min=float(bwd_ms.min()),
                            bwd_ms_median=float(np.median(bwd_ms)), bwd_ms_max=float(bwd_ms.max())),
                clocks=sampler.summary() if sampler else None,
                algo_fwd=native.last_algo(False), algo_bwd=native.last_algo(True))


def rank_stats(torch, dist, dev, world, local):
    """max / mean / per-rank of a few per-rank floats."""
    t = torch.tensor(local, dtype=torch.float64, device=dev)
    if world > 1:
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        allr = torch.stack(parts).cpu().numpy()
    else:
        allr = t.cpu().numpy()[None]
    return allr


def roofline_record(wl, res, peak, peak_src, traffic, top_bytes=4):
    b_fwd, b_bwd = wl.bytes_forward(top_bytes=top_bytes), wl.bytes_backward(top_bytes=top_bytes)
    fwd_ms, bwd_ms, step_ms = res['fwd_ms'], res['bwd_ms'], res['step_ms']
    if bwd_ms >= fwd_ms:
        kname, kbytes, kms = 'backward (%s)' % res['algo_bwd'], b_bwd, bwd_ms
    else:
        kname, kbytes, kms = 'forward (%s)' % res['algo_fwd'], b_fwd, fwd_ms
    achieved = kbytes / (kms * 1e-3) / 1e9
    tr = (traffic or {}).get(wl.name, {}).get(kname, {}).get('bytes')
    return dict(bound='hbm', kernel=kname, achieved=achieved, peak=peak, unit='GB/s',
                frac=achieved / peak, traffic=tr, peak_source=peak_src, algorithmic_bytes=kbytes,
                fwd=dict(ms=fwd_ms, bytes=b_fwd, gbs=b_fwd / (fwd_ms * 1e-3) / 1e9,
                         frac=b_fwd / (fwd_ms * 1e-3) / 1e9 / peak, algo=res['algo_fwd']),
                bwd=dict(ms=bwd_ms, bytes=b_bwd, gbs=b_bwd / (bwd_ms * 1e-3) / 1e9,
                         frac=b_bwd / (bwd_ms * 1e-3) / 1e9 / peak, algo=res['algo_bwd']),
                step=dict(bytes=b_fwd + b_bwd, gbs=(b_fwd + b_bwd) / (step_ms * 1e-3) / 1e9,
                          frac=(b_fwd + b_bwd) / (step_ms * 1e-3) / 1e9 / peak))


def load_traffic():
    """DRAM bytes per launch of the hot kernels, written by scripts/ncu_traffic.py from the
    `ncu --set full` capture of the same round (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mobulaop_b200 import _native, build, sharding

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    from mobulaop_b200.testing import cuda_ready
    if not cuda_ready():
        raise SystemExit('bench.py needs a CUDA device: the ROIAlign path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    build.build()
    lib = _native.load()

    base = WL.CONFIGS[args.workload]
    if args.scaling == 'strong':
        # the workload is fixed (C4: 64 images, 32768 ROIs); rank g owns images
        # [g*N/G, (g+1)*N/G) and every ROI that points into them (per-image ROI ownership)
        if base.N < world:
            raise SystemExit('%s has %d images: cannot split over %d ranks' % (base.name, base.N, world))
        glob = base
    else:
        # weak scaling: every rank owns base.N images and their ROIs
        glob = base.scaled(N=base.N * world, name=base.name)
    rois_global = WL.make_rois(glob, seed=0)
    shard = sharding.make_shard(rois_global, glob.N, rank, world)
    wl = base.scaled(N=shard.num_images, name=base.name)
    assert shard.num_rois == wl.R, (shard.num_rois, wl.R)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev)   # 256 MB, only ever read

    def flush_l2():
        # write a buffer larger than L2, then stream another one through it: the L2 ends up
        # holding clean, unrelated lines, so the timed kernels neither hit in it nor pay for
        # writing the flush buffer's dirty lines back to HBM
        flush.zero_()
        flush_rd.sum()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    K, W = max(args.steps, 1), max(args.warmup, 3)
    run = DeviceRun(lib, torch, dev, local_rank, wl, shard.rois, args.mode, args.algo, 1234 + rank,
                    args.top_dtype, args.layout)
    res = timed_steps(run, flush_l2, barrier, K, W, clocks_index=local_rank)
    allr = rank_stats(torch, dist, dev, world, [res['step_ms'], res['fwd_ms'], res['bwd_ms'],
                                                float(res['launches'])])
    step_ms, fwd_max, bwd_max = float(allr[:, 0].max()), float(allr[:, 1].max()), float(allr[:, 2].max())
    launches = int(allr[:, 3].sum())
    total_rois = glob.R
    value = total_rois / (step_ms * 1e-3)

    peak, peak_src = measured_peaks()
    traffic = load_traffic()
    run_top_bytes = run.top_bytes
    roofline = roofline_record(wl, res, peak, peak_src, traffic, run_top_bytes)      # this rank's launches

    # ---- end to end through the public API --------------------------------------------------
    e2e = e2e_torch = None
    if not args.no_e2e:
        e2e_torch = run_e2e_torch(args, wl, run, world, dev, glob.R)
    run.free()
    if not args.no_e2e:
        e2e = run_e2e(args, wl, shard, world, dev, glob.R)

    # ---- secondary record: the C2 box head per rank (the shape round 1 was quoted on) ---------
    secondary = None
    if args.workload != 'C2' and not args.no_secondary:
        c2 = WL.CONFIGS['C2']
        c2_rois = WL.make_rois(c2, seed=0)
        run2 = DeviceRun(lib, torch, dev, local_rank, c2, c2_rois, args.mode, args.algo, 99 + rank)
        res2 = timed_steps(run2, flush_l2, barrier, max(10, min(K, 50)), W)
        all2 = rank_stats(torch, dist, dev, world, [res2['step_ms'], res2['fwd_ms'], res2['bwd_ms']])
        run2.free()
        s_ms = float(all2[:, 0].max())
        secondary = dict(workload=c2.name, config=c2.describe(), scaling='weak (every rank runs C2)',
                         value=c2.R * world / (s_ms * 1e-3), unit=UNIT, ms_per_step=s_ms,
                         fwd_ms=float(all2[:, 1].max()), bwd_ms=float(all2[:, 2].max()),
                         roofline=roofline_record(c2, res2, peak, peak_src, traffic))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = time_cpu(base, steps=3, warmup=1, budget_s=4.0)
            cpu = dict(value=r['value'], unit=UNIT, cores=r['cores'], kind=r['kind'], sample=r['sample'])
        except Exception as err:          # the checker is optional for the product arm
            cpu = dict(value=None, unit=UNIT, cores=0, kind='unavailable', sample=str(err))

    if rank == 0:
        b_step = wl.bytes_forward(top_bytes=run_top_bytes) + wl.bytes_backward(top_bytes=run_top_bytes)
        details = dict(
            images_total=glob.N, rois_total=total_rois, images_per_rank=wl.N, rois_per_rank=wl.R,
            parallelism='roi-shard-by-image x%d (%s scaling)' % (world, args.scaling),
            backward_mode=args.mode, algo=args.algo, pooled_tensor=dict(dtype=args.top_dtype, layout=args.layout),
            shard_ms=dict(max=step_ms, mean=float(allr[:, 0].mean()), per_rank=[float(v) for v in allr[:, 0]]),
            fwd_ms_max=fwd_max, bwd_ms_max=bwd_max, rank0_step_spread=res['spread'],
            l2='256 MB buffer written, then another 256 MB read, between timed steps (outside the CUDA '
               'events): cold L2 without dirty flush lines; a rank\'s step moves %.0f MB > 126 MB L2' % (b_step / 1e6),
            compare_with='profiles/r1_bench_C4.json (round 1, C4 on one GPU: 3.47 M ROIs/s, 16.8 %% of '
                         'roofline); BENCH_r01.json was the C2 line -- see `secondary`'
            if args.workload == 'C4' else None)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W,
                    ms_per_step=step_ms, higher_is_better=True, scaling=args.scaling, vs_baseline=None,
                    dtype='f32', data='synthetic', config=base.describe(), details=details,
                    roofline=roofline, cpu_baseline=cpu, e2e=e2e, e2e_torch=e2e_torch,
                    secondary=secondary, gpu_launches=launches, clocks=res['clocks'],
                    wall_s_timed_region=res['wall'])
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_e2e_torch(args, wl, run, world, dev, total_rois):
    """fwd+bwd through mobula.op.ROIAlign on DEVICE tensors (torch autograd), wall clock over
    many steps: what a training loop pays, dispatch included (no L2 flush, no copies)."""
    import torch
    import torch.distributed as dist
    import mobula
    mobula.op.load('ROIAlign')
    x = run.data.requires_grad_(True)
    deterministic = args.mode == 'deterministic'

    def step():
        x.grad = None
        y = mobula.op.ROIAlign(x, run.rois, pooled_size=wl.pooled, spatial_scale=wl.scale,
                               sampling_ratio=wl.sampling_ratio, deterministic=deterministic,
                               out_dtype=None if args.top_dtype == 'float32' else args.top_dtype,
                               layout=args.layout)
        y.backward(run.dy)

    with mobula.config.TempConfig(ROIALIGN_ALGO=args.algo):
        for _ in range(3):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        n = max(args.e2e_torch_steps, 1)
        t0 = time.perf_counter()
        for _ in range(n):
            step()
        host_s = (time.perf_counter() - t0) / n          # time to ISSUE a step
        torch.cuda.synchronize(dev)
        step_s = (time.perf_counter() - t0) / n
    run.data.requires_grad_(False)
    x.grad = None
    t = torch.tensor([step_s, host_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_s, host_s = t[0].item(), t[1].item()
    return dict(value=total_rois / step_s, unit=UNIT, ms_per_step=step_s * 1e3,
                host_issue_us_per_step=host_s * 1e6, steps=n,
                api='mobula.op.ROIAlign(x, rois, ...) + y.backward(dy) on CUDA tensors (torch autograd), '
                    'wall clock, warm L2, no copies')


def run_e2e(args, wl, shard, world, dev, total_rois):
    """fwd+bwd through mobula.op.ROIAlign on NumPy views of pinned host memory."""
    import ctypes
    import torch
    import torch.distributed as dist
    import mobula
    from mobulaop_b200 import _native
    lib = _native.load()
    mobula.op.load('ROIAlign')
    if args.top_dtype == 'bfloat16':
        return None                     # NumPy has no bfloat16: the host-buffer API takes float32 / float16

    def pinned(shape, dtype=np.float32):
        n = int(np.prod(shape))
        nbytes = n * np.dtype(dtype).itemsize
        ptr = lib.mobula_host_alloc(nbytes)
        if not ptr:
            raise MemoryError('mobula_host_alloc')
        buf = (ctypes.c_char * nbytes).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape), ptr

    gen = torch.Generator(device=dev)
    gen.manual_seed(99)
    tdt = np.dtype(args.top_dtype)
    oshape = wl.out_shape if args.layout == 'RCHW' else (wl.R,) + tuple(wl.pooled) + (wl.C,)
    data, p0 = pinned(wl.data_shape)
    dy, p1 = pinned(oshape, tdt)
    # synthetic N(0,1) inputs, generated on the device and copied into the pinned host arrays
    torch.from_numpy(data).copy_(torch.randn(wl.data_shape, dtype=torch.float32, device=dev, generator=gen))
    torch.from_numpy(dy).copy_(torch.randn(oshape, dtype=torch.float32, device=dev, generator=gen))
    torch.cuda.empty_cache()
    rois = shard.rois
    op = mobula.op.ROIAlign[np.ndarray](pooled_size=wl.pooled, spatial_scale=wl.scale,
                                        sampling_ratio=wl.sampling_ratio,
                                        deterministic=args.mode == 'deterministic',
                                        out_dtype=None if args.top_dtype == 'float32' else args.top_dtype,
                                        layout=args.layout)
    with mobula.config.TempConfig(ROIALIGN_ALGO=args.algo):
        for _ in range(2):
            y = op(data, rois)
            op.backward(dy)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        n = max(args.e2e_steps, 1)
        t0 = time.perf_counter()
        for _ in range(n):
            y = op(data, rois)
            dxh, _ = op.backward(dy)
        torch.cuda.synchronize(dev)
        step_s = (time.perf_counter() - t0) / n
    t = torch.tensor([step_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_s = t.item()
    h2d = data.nbytes + dy.nbytes + 2 * rois.nbytes
    d2h = y.nbytes + dxh.nbytes
    del y, dxh
    lib.mobula_host_free(ctypes.c_void_p(p0))
    lib.mobula_host_free(ctypes.c_void_p(p1))
    return dict(value=total_rois / step_s, unit=UNIT, h2d_bytes_per_step=int(h2d),
                d2h_bytes_per_step=int(d2h), ms_per_step=step_s * 1e3, steps=n,
                api='mobula.op.ROIAlign[np.ndarray] (C-ABI, MOBULA_ROIALIGN_HOST_BUFFERS), '
                    'inputs and results in page-locked host memory (NumPy arrays); bytes are per rank')


_RESULT_FD = None


def claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries print there too (NCCL's version banner,
    for one), so file descriptor 1 is pointed at stderr for the rest of the run and the result
    line goes to a private duplicate of the original stdout."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == 'reference':
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
