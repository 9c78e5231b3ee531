"""Runs one case through the sweep forward several times and lists the mismatches.
    sweep_repro.py golden <name> [reps]
    sweep_repro.py wl <C2|C3|..> <C> <rois_per_image> [reps]"""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobulaop_b200 import _native, build, workloads as WL
build.build(); lib = _native.load()
if sys.argv[1] == 'golden':
    z = np.load('tests/golden/%s.npz' % sys.argv[2])
    data, rois, y_ref = z['data'], z['rois'], z['y']
    PH, PW = [int(v) for v in z['pooled']]; scale = float(z['scale']); sr = int(z['sr'])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
else:
    from oracle.roialign import COracle
    wl = WL.CONFIGS[sys.argv[2]].scaled(C=int(sys.argv[3]), rois_per_image=int(sys.argv[4]))
    data, rois, _ = WL.make_inputs(wl, seed=17, with_dy=False)
    PH, PW = wl.pooled; scale = wl.scale; sr = wl.sampling_ratio
    y_ref = COracle().forward(data, rois, wl.pooled, scale, sr, threads=0)
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
dev = torch.device('cuda', 0)
d = torch.from_numpy(data).to(dev); r = torch.from_numpy(rois).to(dev)
N, C, H, W = data.shape; R = len(rois)
p = lambda t: ctypes.c_void_p(t.data_ptr())
for it in range(reps):
    out = torch.full(y_ref.shape, float('nan'), device=dev)
    _native.check(lib.mobula_roialign_forward_f32(0, None, p(d), p(r), p(out), R, N, C, H, W, PH, PW, scale, sr,
                                                  _native.make_flags(algo='sweep')))
    torch.cuda.synchronize()
    y = out.cpu().numpy()
    bad = np.argwhere(y.view(np.int32) != y_ref.view(np.int32))
    print('run', it, 'mismatches', len(bad), 'of', y.size, [tuple(int(v) for v in b) for b in bad[:6]])
    if len(bad):
        print('   rois', len(set(bad[:, 0].tolist())), 'channels', sorted(set(bad[:, 1].tolist())), 'ph', sorted(set(bad[:, 2].tolist())),
              'nan', int(np.isnan(y).sum()))
        r0 = bad[0][0]; print('   roi', r0, rois[r0], 'bad ph of it', sorted(set(bad[bad[:, 0] == r0][:, 2].tolist())))
