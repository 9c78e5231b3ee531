"""Conv2D through im2col / col2im (reference: /root/reference/opzoo/Convolution/Convolution.py:6-95).

    mobula.op.load('Convolution')
    y = mobula.op.Conv2D(x, weight, bias, channels=D, kernel_size=(KH, KW), strides=..., padding=...)

x [N, C, H, W], weight [D, C, KH, KW], optional bias [D]; y [N, D, OH, OW].  Per image the
column buffer [C*KH*KW, OH, OW] comes from `mobula.func.im2col`, the product with the
[D, C*KH*KW] weight matrix from the framework (`@`), the input gradient from
`mobula.func.col2im` -- the two kernels are the prebuilt sm_100a ones.  The reference class is
written against MXNet NDArray methods (`reshape_like`, `sum(1, exclude=True)`); this one only
uses what NumPy arrays and torch tensors share.
"""
import mobulaop_b200 as mobula
from mobulaop_b200.const import req


def _size(t):
    n = 1
    for s in t.shape:
        n *= int(s)
    return n


@mobula.op.register
class Conv2D:
    def __init__(self, channels, kernel_size, strides=(1, 1), padding=(0, 0), dilation=(1, 1), groups=1):
        self.channels = int(channels)
        self.kernel_size = tuple(int(k) for k in kernel_size)
        self.strides = tuple(int(k) for k in strides)
        self.padding = tuple(int(k) for k in padding)
        self.dilation = tuple(int(k) for k in dilation)
        assert groups == 1
        self.groups = groups

    def _empty_like(self, ref, shape):
        if hasattr(ref, 'new_empty'):           # torch: same dtype and device
            return ref.new_empty(shape)
        return self.F.empty(shape, dtype=ref.dtype)

    def _geometry(self, x_shape, out_shape):
        _, _, H, W = x_shape
        KH, KW = self.kernel_size
        PH, PW = self.padding
        SH, SW = self.strides
        DH, DW = self.dilation
        OH, OW = out_shape[2], out_shape[3]
        return (H, W, KH, KW, PH, PW, SH, SW, DH, DW, OH, OW)

    def forward(self, x, weight, bias=None):
        # y = wx + b
        N, C = x.shape[0], x.shape[1]
        geom = self._geometry(x.shape, self.y.shape)
        KH, KW, OH, OW = geom[2], geom[3], geom[10], geom[11]
        D = self.y.shape[1]
        csize = C * KH * KW
        data_col = self._empty_like(x, (csize, OH, OW))
        rweight = weight.reshape((D, csize))
        rbias = bias.reshape((-1, 1, 1)) if bias is not None else None
        for i in range(N):
            mobula.func.im2col(_size(data_col), x[i], *geom, data_col)
            out = (rweight @ data_col.reshape((csize, OH * OW))).reshape((D, OH, OW))
            if rbias is not None:
                out = out + rbias
            self.assign(self.y[i], self.req[0], out)

    def backward(self, dy):
        x, weight = self.X[0], self.X[1]
        N, C = x.shape[0], x.shape[1]
        geom = self._geometry(x.shape, dy.shape)
        KH, KW, OH, OW = geom[2], geom[3], geom[10], geom[11]
        D = dy.shape[1]
        csize = C * KH * KW
        ohw = OH * OW
        rweight = weight.reshape((D, csize))
        want_dx = self.req[0] != req.null
        want_dw = self.req[1] != req.null
        out = self._empty_like(x, tuple(x.shape[1:])) if want_dx else None
        col = self._empty_like(x, (csize, OH, OW)) if want_dw else None
        dw = 0
        for i in range(N):
            rdy = dy[i].reshape((D, ohw))
            if want_dx:
                data_col = (rweight.T @ rdy).reshape((csize, OH, OW))
                mobula.func.col2im(_size(out), data_col, *geom, out)
                self.assign(self.dX[0][i], self.req[0], out)
            if want_dw:
                mobula.func.im2col(_size(col), x[i], *geom, col)
                dw = dw + rdy @ col.reshape((csize, ohw)).T
        if want_dw:
            self.assign(self.dX[1], self.req[1], dw.reshape(tuple(weight.shape)))
        if len(self.X) == 3 and self.req[2] != req.null:
            self.assign(self.dX[2], self.req[2], dy.sum((0, 2, 3)))

    def infer_shape(self, in_shape):
        assert 2 <= len(in_shape) <= 3, 'The inputs should be feature map(NCHW layout), weight and bias(optional)'
        assert len(in_shape[0]) == 4, 'input: NCHW'
        assert len(in_shape[1]) == 4, 'weight: DCKK'
        assert len(in_shape) == 2 or len(in_shape[2]) == 1, 'bias: D'
        x, weight = in_shape[:2]
        N, C, H, W = x
        KH, KW = self.kernel_size
        PH, PW = self.padding
        SH, SW = self.strides
        DH, DW = self.dilation

        def get_outsize(X, K, P, S, D):
            K += (K - 1) * (D - 1)
            return (X + 2 * P - K) // S + 1
        OH = get_outsize(H, KH, PH, SH, DH)
        OW = get_outsize(W, KW, PW, SW, DW)
        D = self.channels
        assert weight[0] == D
        assert weight[1] == C
        assert weight[2] == KH
        assert weight[3] == KW
        return in_shape, [(N, D, OH, OW)]
