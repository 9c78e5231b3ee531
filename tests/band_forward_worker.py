"""Worker of tests/test_band_forward.py: one process per kernel choice (MOBULA_FWD_BAND is read once
per process).  Runs the forward on box-head sized planes -- two channels of a 200 x 272 map do not
fit one SM's shared memory together -- and compares bit for bit with the oracle."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def boxes(rng, n_img, per_img, H, W, scale):
    img_w, img_h = W / scale, H / scale
    rows = []
    for n in range(n_img):
        cx, cy = rng.uniform(0, img_w, per_img), rng.uniform(0, img_h, per_img)
        bw = np.exp(rng.uniform(np.log(8), np.log(img_w), per_img))
        bh = np.exp(rng.uniform(np.log(8), np.log(img_h), per_img))
        r = np.stack([np.full(per_img, n), cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
        # boxes several times the plane (bin rows no band can hold), boxes outside it, degenerate boxes
        r[0] = [n, -3 * img_w, -4 * img_h, 4 * img_w, 5 * img_h]
        r[1] = [n, 10, -9 * img_h, 60, 9 * img_h]
        r[2] = [n, img_w + 50, img_h + 50, img_w + 90, img_h + 70]
        r[3] = [n, 40, 40, 30, 30]
        r[4] = [n, 5, img_h - 3, 200, img_h + 30]
        r[5] = [n, -20, -20, 90, 14]
        rows.append(r)
    rois = np.concatenate(rows).astype(np.float32)
    rois[-1, 0] = n_img + 3          # a row that points at no image
    return rois


def main():
    import torch
    from mobulaop_b200 import _native
    from oracle.roialign import COracle
    lib = _native.load()
    orc = COracle()
    rng = np.random.default_rng(11)
    dev = torch.device('cuda', 0)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    for (N, C, H, W, pooled, per_img) in [(2, 5, 200, 272, (7, 7), 150), (1, 4, 160, 336, (7, 7), 120),
                                          (2, 3, 160, 216, (5, 3), 90), (1, 2, 200, 272, (3, 11), 60)]:
        data = rng.standard_normal((N, C, H, W), dtype=np.float32)
        rois = boxes(rng, N, per_img, H, W, 0.25)
        # the last row points at no image: the library writes zeros there, the reference would read out of
        # bounds -- the oracle only sees the other rows
        ref = np.concatenate([orc.forward(data, np.ascontiguousarray(rois[:-1]), pooled, 0.25, 2),
                              np.zeros((1, C) + tuple(pooled), np.float32)])
        d, r = torch.from_numpy(data).to(dev), torch.from_numpy(rois).to(dev)
        for accumulate in (False, True):
            out = torch.full(ref.shape, 0.5 if accumulate else float('nan'), device=dev)
            _native.check(lib.mobula_roialign_forward_f32(0, None, vp(d), vp(r), vp(out), rois.shape[0], N, C, H, W,
                                                          pooled[0], pooled[1], 0.25, 2,
                                                          _native.make_flags(accumulate=accumulate, algo='resident')))
            torch.cuda.synchronize()
            got = out.cpu().numpy()
            want = (np.float32(0.5) + ref) if accumulate else ref
            same = got.view(np.uint32) == want.view(np.uint32)
            if not same.all():
                idx = np.unravel_index(int(np.argmax(~same)), got.shape)
                print('MISMATCH', (N, C, H, W, pooled), 'accumulate', accumulate, int((~same).sum()), idx, got[idx], want[idx])
                return 1
        # half-precision store variant
        out16 = torch.empty(ref.shape, dtype=torch.float16, device=dev)
        _native.check(lib.mobula_roialign_forward_ex(0, None, vp(d), vp(r), vp(out16), 1, 0, rois.shape[0], N, C, H, W,
                                                     pooled[0], pooled[1], 0.25, 2, _native.make_flags(algo='resident')))
        torch.cuda.synchronize()
        if not np.array_equal(out16.cpu().numpy().view(np.uint16), ref.astype(np.float16).view(np.uint16)):
            print('MISMATCH fp16', (N, C, H, W, pooled))
            return 1
    print('OK band=%s' % os.environ.get('MOBULA_FWD_BAND'))
    return 0


if __name__ == '__main__':
    sys.exit(main())
