#!/usr/bin/env python
"""ROIAlign fwd+bwd throughput on B200 (BASELINE.json metric) -- one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2]

A "step" is one forward bilinear pool plus one backward gradient scatter (dX produced from
scratch, zero-fill included) over one batch of synthetic inputs already resident in HBM.
N=1 runs BASELINE.json configs[1] (C2, the Faster R-CNN box head); N>1 runs the same
per-GPU work on every rank (weak scaling): the global problem has 2*N images, it is split
with mobulaop_b200.sharding (per-image ROI ownership, no data-path collective).

value      ROIs/s over all ranks, device time (CUDA events on the launching stream), max over ranks
e2e        the same metric through mobula.op.ROIAlign[np.ndarray] with pinned HOST buffers:
           host->device and device->host copies inside the timed region
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
    ap.add_argument('--workload', default='C2', choices=sorted(WL.CONFIGS))
    ap.add_argument('--mode', default='atomic', choices=['atomic', 'deterministic'])
    ap.add_argument('--algo', default='auto', choices=['auto', 'gather', 'plane', 'resident'])
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
        scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
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
def run_ours(args):
    import torch
    import torch.distributed as dist
    from mobulaop_b200 import _native, build, sharding

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the ROIAlign path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    build.build()
    lib = _native.load()

    base = WL.CONFIGS[args.workload]
    # weak scaling: every rank owns base.N images and their ROIs
    glob = base.scaled(N=base.N * world, name=base.name)
    rois_global = WL.make_rois(glob, seed=0)
    shard = sharding.make_shard(rois_global, glob.N, rank, world)
    assert shard.num_images == base.N and shard.num_rois == base.R
    wl = base
    rng = np.random.default_rng(1234 + rank)
    data = torch.from_numpy(rng.standard_normal(wl.data_shape, dtype=np.float32)).to(dev)
    dy = torch.from_numpy(rng.standard_normal(wl.out_shape, dtype=np.float32)).to(dev)
    rois = torch.from_numpy(shard.rois).to(dev)
    out = torch.empty(wl.out_shape, dtype=torch.float32, device=dev)
    dx = torch.empty(wl.data_shape, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev)   # 256 MB, only ever read

    def flush_l2():
        # write a buffer larger than L2, then stream another one through it: the L2 ends up
        # holding clean, unrelated lines, so the timed kernels neither hit in it nor pay for
        # writing the flush buffer's dirty lines back to HBM
        flush.zero_()
        flush_rd.sum()

    import ctypes
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    mode = _native.BWD_DETERMINISTIC if args.mode == 'deterministic' else _native.BWD_ATOMIC
    flags = _native.make_flags(algo=args.algo)
    PH, PW = wl.pooled

    def vp(t):
        return ctypes.c_void_p(t.data_ptr())

    def fwd():
        _native.check(lib.mobula_roialign_forward_f32(
            local_rank, sp, vp(data), vp(rois), vp(out), wl.R, wl.N, wl.C, wl.H, wl.W, PH, PW,
            wl.scale, wl.sampling_ratio, flags))

    def bwd():
        _native.check(lib.mobula_roialign_backward_f32(
            local_rank, sp, vp(dy), vp(rois), vp(dx), wl.R, wl.N, wl.C, wl.H, wl.W, PH, PW,
            wl.scale, wl.sampling_ratio, mode, flags))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        flush_l2()
        fwd()
        bwd()
    barrier()

    K = max(args.steps, 1)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    launches0 = _native.launch_count()
    with ClockSampler(local_rank) as clocks:
        barrier()
        wall0 = time.perf_counter()
        for k in range(K):
            flush_l2()                          # evict L2 between steps (outside the events)
            ev[k][0].record(stream)
            fwd()
            ev[k][1].record(stream)
            bwd()
            ev[k][2].record(stream)
        barrier()
        wall = time.perf_counter() - wall0
    launches = _native.launch_count() - launches0
    fwd_ms = np.array([ev[k][0].elapsed_time(ev[k][1]) for k in range(K)])
    bwd_ms = np.array([ev[k][1].elapsed_time(ev[k][2]) for k in range(K)])
    step_ms_local = float((fwd_ms + bwd_ms).sum() / K)

    stats = torch.tensor([step_ms_local, float(fwd_ms.mean()), float(bwd_ms.mean()), float(launches)],
                         dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        step_ms, fwd_mean, bwd_mean, launches = mx[0].item(), mx[1].item(), mx[2].item(), int(sm[3].item())
    else:
        step_ms, fwd_mean, bwd_mean = step_ms_local, float(fwd_ms.mean()), float(bwd_ms.mean())
    total_rois = wl.R * world
    value = total_rois / (step_ms * 1e-3)

    # ---- roofline of the dominant kernel (per launch, this rank) ------------------------
    peak, peak_src = measured_peaks()
    b_fwd, b_bwd = wl.bytes_forward(), wl.bytes_backward()
    if bwd_mean >= fwd_mean:
        kname, kbytes, kms = 'backward (%s)' % _native.last_algo(True), b_bwd, bwd_mean
    else:
        kname, kbytes, kms = 'forward (%s)' % _native.last_algo(False), b_fwd, fwd_mean
    achieved = kbytes / (kms * 1e-3) / 1e9
    traffic = None          # DRAM bytes of that kernel from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            traffic = json.load(f).get(wl.name, {}).get(kname, {}).get('bytes')
    except (OSError, ValueError):
        pass
    roofline = dict(bound='hbm', kernel=kname, achieved=achieved, peak=peak, unit='GB/s',
                    frac=achieved / peak, traffic=traffic, peak_source=peak_src,
                    algorithmic_bytes=kbytes,
                    fwd=dict(ms=fwd_mean, bytes=b_fwd, gbs=b_fwd / (fwd_mean * 1e-3) / 1e9,
                             frac=b_fwd / (fwd_mean * 1e-3) / 1e9 / peak,
                             algo=_native.last_algo(False)),
                    bwd=dict(ms=bwd_mean, bytes=b_bwd, gbs=b_bwd / (bwd_mean * 1e-3) / 1e9,
                             frac=b_bwd / (bwd_mean * 1e-3) / 1e9 / peak,
                             algo=_native.last_algo(True)),
                    step=dict(bytes=b_fwd + b_bwd, gbs=(b_fwd + b_bwd) / (step_ms_local * 1e-3) / 1e9,
                              frac=(b_fwd + b_bwd) / (step_ms_local * 1e-3) / 1e9 / peak))

    # ---- end to end through the public API with pinned host buffers ----------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, wl, shard, world, dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            res = time_cpu(wl, steps=3, warmup=1, budget_s=4.0)
            cpu = dict(value=res['value'], unit=UNIT, cores=res['cores'], kind=res['kind'],
                       sample=res['sample'])
        except Exception as err:          # the checker is optional for the product arm
            cpu = dict(value=None, unit=UNIT, cores=0, kind='unavailable', sample=str(err))

    if rank == 0:
        cfg = wl.describe()
        cfg.update(images_total=wl.N * world, rois_total=total_rois, parallelism='roi-shard-by-image x%d' % world,
                   backward_mode=args.mode, algo=args.algo,
                   l2='256 MB buffer written, then another 256 MB read, between timed steps (outside the CUDA '
                      'events): cold L2 without dirty flush lines; '
                      'step inputs+outputs are %.0f MB > 126 MB L2' % ((b_fwd + b_bwd) / 1e6))
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=max(args.warmup, 3),
                    ms_per_step=step_ms, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='f32', data='synthetic', config=cfg, roofline=roofline,
                    cpu_baseline=cpu, e2e=e2e, gpu_launches=int(launches), clocks=clocks.summary(),
                    wall_s_timed_region=wall)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_e2e(args, wl, shard, world, dev):
    """fwd+bwd through mobula.op.ROIAlign on NumPy views of pinned host memory."""
    import ctypes
    import torch
    import torch.distributed as dist
    import mobula
    from mobulaop_b200 import _native
    lib = _native.load()
    mobula.op.load('ROIAlign')

    def pinned(shape):
        n = int(np.prod(shape))
        ptr = lib.mobula_host_alloc(n * 4)
        if not ptr:
            raise MemoryError('mobula_host_alloc')
        buf = (ctypes.c_float * n).from_address(ptr)
        return np.frombuffer(buf, dtype=np.float32).reshape(shape), ptr

    rng = np.random.default_rng(99)
    data, p0 = pinned(wl.data_shape)
    dy, p1 = pinned(wl.out_shape)
    data[...] = rng.standard_normal(wl.data_shape, dtype=np.float32)
    dy[...] = rng.standard_normal(wl.out_shape, dtype=np.float32)
    rois = shard.rois
    op = mobula.op.ROIAlign[np.ndarray](pooled_size=wl.pooled, spatial_scale=wl.scale,
                                        sampling_ratio=wl.sampling_ratio,
                                        deterministic=args.mode == 'deterministic')
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
    lib.mobula_host_free(ctypes.c_void_p(p0))
    lib.mobula_host_free(ctypes.c_void_p(p1))
    return dict(value=wl.R * world / step_s, unit=UNIT, h2d_bytes_per_step=int(h2d),
                d2h_bytes_per_step=int(d2h), ms_per_step=step_s * 1e3, steps=n,
                api='mobula.op.ROIAlign[np.ndarray] (C-ABI, MOBULA_ROIALIGN_HOST_BUFFERS), '
                    'inputs and results in page-locked host memory (NumPy arrays)')


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
