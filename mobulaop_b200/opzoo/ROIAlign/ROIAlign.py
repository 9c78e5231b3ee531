"""ROIAlign operator class (reference: /root/reference/opzoo/ROIAlign/ROIAlign.py:7-51).

    mobula.op.load('ROIAlign')
    y = mobula.op.ROIAlign(data, rois, pooled_size=(7, 7), spatial_scale=0.25, sampling_ratio=2)

data [N, C, H, W] fp32, rois [R, 5] fp32 rows (batch_index, x1, y1, x2, y2) in image
coordinates, output [R, C, PH, PW].  Write requests follow the reference: 'null' skips,
'add' accumulates, anything else overwrites.  The gradient with respect to `rois` is zero.

Beyond the reference: `deterministic` (None = follow mobula.config.ROIALIGN_DETERMINISTIC)
selects the bit-reproducible backward; `out_dtype` / `layout` store the pooled tensor in
float16 / bfloat16 and / or channel-last [R, PH, PW, C] (fp32 arithmetic, one rounding at the store); `aligned=True` uses half-pixel box coordinates (the
convention of torchvision.ops.roi_align(aligned=True) and detectron2's ROIAlignV2); the dX zero-fill of the reference
(ROIAlign.py:33-34, a separate full-tensor pass) is done inside the native call.
"""
import mobulaop_b200 as mobula
from mobulaop_b200.const import req


@mobula.op.register
class ROIAlign:
    def __init__(self, pooled_size, spatial_scale, sampling_ratio, deterministic=None, aligned=False,
                 out_dtype=None, layout='RCHW'):
        if isinstance(pooled_size, int):
            pooled_size = (pooled_size, pooled_size)
        assert len(pooled_size) == 2, 'pooled_size is (pooled_height, pooled_width)'
        self.pooled_size = (int(pooled_size[0]), int(pooled_size[1]))
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.deterministic = deterministic
        self.aligned = bool(aligned)
        # storage of the pooled tensor: None = the input's type; 'float16' / 'bfloat16' (or the
        # framework's dtype object) halve its bytes, the arithmetic stays fp32; layout 'RHWC' makes
        # it channel-last [R, PH, PW, C] for consumers that want the channels innermost
        self.out_dtype = out_dtype
        assert layout in ('RCHW', 'RHWC'), "layout is 'RCHW' or 'RHWC'"
        self.layout = layout

    @staticmethod
    def _numel(t):
        n = 1
        for s in t.shape:
            n *= int(s)
        return n

    def forward(self, data, rois):
        if self.req[0] == req.null:
            return
        out = self.y
        mobula.func.roi_align_forward(
            self._numel(out), data, self.spatial_scale, data.shape[1], data.shape[2],
            data.shape[3], self.pooled_size[0], self.pooled_size[1], self.sampling_ratio, rois, out,
            accumulate=self.req[0] == req.add, aligned=self.aligned, layout=self.layout)

    def backward(self, dy):
        if self.req[0] != req.null:
            data, rois = self.X
            mode = None
            if self.deterministic is not None:
                mode = 'deterministic' if self.deterministic else 'atomic'
            mobula.func.roi_align_backward(
                self._numel(dy), dy, self.spatial_scale, data.shape[1], data.shape[2],
                data.shape[3], self.pooled_size[0], self.pooled_size[1], self.sampling_ratio,
                self.dX[0], rois, accumulate=self.req[0] == req.add, mode=mode, aligned=self.aligned,
                layout=self.layout)
        if len(self.req) > 1 and self.req[1] not in (req.null, req.add):
            self.dX[1][:] = 0

    def infer_shape(self, in_shape):
        dshape, rshape = in_shape
        assert len(dshape) == 4
        assert len(rshape) == 2
        assert rshape[1] == 5
        if self.layout == 'RHWC':
            oshape = [rshape[0], self.pooled_size[0], self.pooled_size[1], dshape[1]]
        else:
            oshape = [rshape[0], dshape[1], self.pooled_size[0], self.pooled_size[1]]
        return [dshape, rshape], [oshape]

    def infer_out_dtype(self, in_dtype, F):
        if self.out_dtype is None:
            return in_dtype
        if isinstance(self.out_dtype, str):
            return getattr(F, self.out_dtype)
        return self.out_dtype
