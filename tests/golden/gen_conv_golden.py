#!/usr/bin/env python
"""Generate tests/golden/conv_*.npz: im2col / col2im outputs of the reference's own kernels.

Runs ONLY in the build container: the vectors come from oracle/_ref/libmobula_ref_conv.so, i.e.
/root/reference/opzoo/Convolution/Convolution.cpp compiled where it lies (oracle/Makefile,
oracle/conv_shim.cpp).  The GPU box has no /root/reference, so the vectors are committed.

    make -C oracle ref && python tests/golden/gen_conv_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.normpath(os.path.join(HERE, '..', '..')))

CASES = {
    # name: (C, H, W, kernel, pad, stride, dilation)
    'conv_ref_test': (2, 3, 4, (2, 3), (0, 1), (1, 2), (1, 1)),      # opzoo/Convolution/test_conv.py:11-15
    'conv_3x3_same': (5, 13, 17, (3, 3), (1, 1), (1, 1), (1, 1)),
    'conv_dilated': (3, 16, 20, (3, 2), (2, 1), (2, 1), (2, 3)),
    'conv_strided_nopad': (4, 11, 9, (4, 4), (0, 0), (3, 2), (1, 1)),
    'conv_1x1': (7, 6, 8, (1, 1), (0, 0), (1, 1), (1, 1)),
    'conv_wide_pad': (2, 5, 6, (3, 3), (3, 4), (1, 1), (1, 1)),
}


def main():
    from oracle.conv import RefConv, out_size
    ref = RefConv()
    rng = np.random.default_rng(20261020)
    for name, (C, H, W, k, p, s, d) in CASES.items():
        OH, OW = out_size(H, k[0], p[0], s[0], d[0]), out_size(W, k[1], p[1], s[1], d[1])
        im = rng.standard_normal((C, H, W)).astype(np.float32)
        colg = rng.standard_normal((C * k[0] * k[1], OH, OW)).astype(np.float32)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), im=im, colg=colg, kernel=k, pad=p, stride=s,
                            dilation=d, col=ref.im2col(im, k, p, s, d), im_grad=ref.col2im(colg, (C, H, W), k, p, s, d))
        print(name, (C, H, W), '->', (C * k[0] * k[1], OH, OW))


if __name__ == '__main__':
    main()
