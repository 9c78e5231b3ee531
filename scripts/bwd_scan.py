"""Backward time of the C4 geometry (256 x 200 x 272 maps, 512 ROIs per image, 7 x 7, sampling ratio 2)
for a range of image counts: python scripts/bwd_scan.py 2 4 8 16 32   (needs a GPU)"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch

import mobula
from mobulaop_b200 import workloads as WL

mobula.op.load('ROIAlign')
dev = torch.device('cuda', 0)
flush = torch.empty(64 * 1024 * 1024, device=dev)
for n in [int(a) for a in sys.argv[1:]]:
    wl = WL.CONFIGS["C4"].scaled(N=n)
    data, rois, dy = WL.make_inputs(wl, seed=3)
    rois_t = torch.from_numpy(rois).to(dev)
    dy_t = torch.from_numpy(dy).to(dev)
    dx = torch.empty(data.shape, device=dev)
    C, H, W = data.shape[1:]
    PH, PW = wl.pooled
    times = []
    for it in range(13):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        mobula.func.roi_align_backward(dy_t.numel(), dy_t, wl.scale, C, H, W, PH, PW, wl.sampling_ratio, dx, rois_t,
                                       accumulate=False)
        b.record()
        torch.cuda.synchronize()
        if it >= 3:
            times.append(a.elapsed_time(b))
    print('N=%d %s=%s bwd %.1f us (median of 10)' % (n, sys.argv[0] and 'MOBULA_TABLES_Q', os.environ.get('MOBULA_TABLES_Q'),
                                                    1e3 * float(np.median(times))))
