"""Host-side cost of one call, layer by layer.

The reference's pattern for this is benchmark/benchmark_empty_op.py:15-25 (an operator
that does nothing, called in a loop); here the operator is ROIAlign on a problem so small
that the device work hides behind the host: what is left is ctypes marshalling, the
dispatcher and the launches.  Prints one JSON object; needs a GPU.

    python scripts/dispatch_overhead.py [--calls 2000]
"""
import argparse
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

import numpy as np
import torch


def per_call_us(fn, calls, dev):
    for _ in range(50):
        fn()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(calls):
        fn()
    host = time.perf_counter() - t0          # the calls have returned; the device may lag
    torch.cuda.synchronize(dev)
    total = time.perf_counter() - t0
    return {'host_us': round(host / calls * 1e6, 2), 'with_device_us': round(total / calls * 1e6, 2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--calls', type=int, default=2000)
    args = ap.parse_args()

    import mobula
    from mobulaop_b200 import _native
    mobula.op.load('ROIAlign')
    lib = _native.load()
    dev = torch.device('cuda', 0)
    N, C, H, W, R, PH, PW, scale, sr = 1, 4, 16, 16, 8, 7, 7, 1.0, 2
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((N, C, H, W), dtype=np.float32)).to(dev)
    rois_np = np.zeros((R, 5), np.float32)
    rois_np[:, 1:3] = rng.uniform(0, 6, (R, 2))
    rois_np[:, 3:5] = rois_np[:, 1:3] + rng.uniform(2, 8, (R, 2))
    rois = torch.from_numpy(rois_np).to(dev)
    y = torch.empty((R, C, PH, PW), device=dev)
    dy = torch.ones_like(y)
    dx = torch.empty_like(x)
    nthreads = int(y.numel())
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    flags = _native.make_flags()

    def raw_forward():
        lib.mobula_roialign_forward_f32(0, stream, vp(x), vp(rois), vp(y), R, N, C, H, W, PH, PW,
                                        scale, sr, flags)

    def func_forward():
        mobula.func.roi_align_forward(nthreads, x, scale, C, H, W, PH, PW, sr, rois, y)

    def func_backward():
        mobula.func.roi_align_backward(nthreads, dy, scale, C, H, W, PH, PW, sr, dx, rois, accumulate=False)

    def op_forward():
        mobula.op.ROIAlign(x, rois, pooled_size=(PH, PW), spatial_scale=scale, sampling_ratio=sr)

    xg = x.clone().requires_grad_(True)

    def op_forward_backward():
        out = mobula.op.ROIAlign(xg, rois, pooled_size=(PH, PW), spatial_scale=scale, sampling_ratio=sr)
        out.backward(dy)
        xg.grad = None

    result = {
        'problem': 'N=%d C=%d %dx%d R=%d pooled %dx%d sr=%d (device work << host work)' % (N, C, H, W, R, PH, PW, sr),
        'calls': args.calls,
        'c_abi_forward': per_call_us(raw_forward, args.calls, dev),
        'func_forward': per_call_us(func_forward, args.calls, dev),
        'func_backward': per_call_us(func_backward, args.calls, dev),
        'op_forward_torch': per_call_us(op_forward, args.calls, dev),
        'op_forward_backward_torch': per_call_us(op_forward_backward, args.calls, dev),
    }

    # the same forward + backward, captured once and replayed: no dispatcher at all
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        func_forward()
        func_backward()
    torch.cuda.synchronize(dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        func_forward()
        func_backward()
    result['graph_replay_forward_backward'] = per_call_us(graph.replay, args.calls, dev)
    result['reference_probe'] = {'func_call_us': 74, 'op_call_us': 138,
                                 'note': 'SURVEY.md 8(a) [probe]: reference dispatcher on its CPU EmptyOP'}
    print(json.dumps(result, indent=1))


if __name__ == '__main__':
    main()
