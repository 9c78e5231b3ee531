"""Synthetic ROIAlign workloads: the BASELINE.json configurations.

Everything is generated on the host with numpy.random.default_rng(seed) so the CUDA path,
the oracle and the reference CPU arm see identical inputs.
"""
import numpy as np


class Workload(object):
    def __init__(self, name, N, C, H, W, rois_per_image, pooled, scale, sampling_ratio,
                 min_box=16.0, max_box=512.0, stress=False, literal=False):
        self.name = name
        self.N, self.C, self.H, self.W = N, C, H, W
        self.rois_per_image = rois_per_image
        self.pooled = tuple(pooled)
        self.scale = float(scale)
        self.sampling_ratio = int(sampling_ratio)
        self.min_box, self.max_box = float(min_box), float(max_box)
        self.stress = stress
        # literal: the inputs of the reference's own benchmark script instead of random FPN boxes
        # (benchmark/benchmark_roi_align.py:15-21: data = arange, rois[i] = [i % N, 0, 0, i % H, i % W])
        self.literal = literal

    @property
    def R(self):
        return self.N * self.rois_per_image

    @property
    def data_shape(self):
        return (self.N, self.C, self.H, self.W)

    @property
    def out_shape(self):
        return (self.R, self.C) + self.pooled

    def scaled(self, N=None, C=None, rois_per_image=None, name=None):
        """Same geometry with fewer images / channels / ROIs (oracle-sized samples)."""
        return Workload(name or self.name + '-sample', N or self.N, C or self.C, self.H, self.W,
                        rois_per_image or self.rois_per_image, self.pooled, self.scale,
                        self.sampling_ratio, self.min_box, self.max_box, self.stress, self.literal)

    def describe(self):
        return dict(workload=self.name, N=self.N, C=self.C, H=self.H, W=self.W, R=self.R,
                    pooled=list(self.pooled), spatial_scale=self.scale,
                    sampling_ratio=self.sampling_ratio)

    # algorithmic (useful) bytes, SURVEY.md 8(d), with U = N*C*H*W unless given
    def bytes_forward(self, unique_elements=None, top_bytes=4):
        """`top_bytes`: bytes per element of the pooled tensor (4 = fp32, 2 = fp16 / bf16)."""
        u = self.N * self.C * self.H * self.W if unique_elements is None else unique_elements
        return 4 * u + top_bytes * self.R * self.C * self.pooled[0] * self.pooled[1] + 20 * self.R

    def bytes_backward(self, top_bytes=4):
        return (top_bytes * self.R * self.C * self.pooled[0] * self.pooled[1]
                + 4 * self.N * self.C * self.H * self.W + 20 * self.R)


def make_rois(wl, seed=0):
    """[R, 5] float32, `rois_per_image` per image (grouped by image unless `wl.literal`)."""
    if wl.literal:
        return literal_benchmark_inputs(wl.N, 1, wl.H, wl.W, wl.R)[1]
    rng = np.random.default_rng(seed + 1000003)
    img_w, img_h = wl.W / wl.scale, wl.H / wl.scale
    out = np.empty((wl.R, 5), np.float32)
    k = wl.rois_per_image
    for n in range(wl.N):
        cx = rng.uniform(0, img_w, k)
        cy = rng.uniform(0, img_h, k)
        hi_w = max(wl.min_box * 1.0001, min(wl.max_box, img_w) if not wl.stress else img_w)
        hi_h = max(wl.min_box * 1.0001, min(wl.max_box, img_h) if not wl.stress else img_h)
        bw = np.exp(rng.uniform(np.log(wl.min_box), np.log(hi_w), k))
        bh = np.exp(rng.uniform(np.log(wl.min_box), np.log(hi_h), k))
        x1 = np.clip(cx - bw / 2, -8, img_w + 8)
        x2 = np.clip(cx + bw / 2, -8, img_w + 8)
        y1 = np.clip(cy - bh / 2, -8, img_h + 8)
        y2 = np.clip(cy + bh / 2, -8, img_h + 8)
        if wl.stress:
            kind = rng.uniform(0, 1, k)
            deg = kind < 0.01                       # x2 < x1, y2 < y1 -> forced 1x1 bins
            x1, x2 = np.where(deg, x2, x1), np.where(deg, x1, x2)
            y1, y2 = np.where(deg, y2, y1), np.where(deg, y1, y2)
            outside = (kind >= 0.01) & (kind < 0.02)  # fully outside: every sample rejected
            x1 = np.where(outside, img_w + 64 + x1, x1)
            x2 = np.where(outside, img_w + 64 + x2, x2)
            y1 = np.where(outside, img_h + 64 + y1, y1)
            y2 = np.where(outside, img_h + 64 + y2, y2)
        rows = slice(n * k, (n + 1) * k)
        out[rows, 0] = n
        out[rows, 1], out[rows, 2], out[rows, 3], out[rows, 4] = x1, y1, x2, y2
    return out


def make_inputs(wl, seed=0, with_dy=True):
    rng = np.random.default_rng(seed)
    if wl.literal:
        data = np.arange(int(np.prod(wl.data_shape)), dtype=np.float32).reshape(wl.data_shape)
    else:
        data = rng.standard_normal(wl.data_shape, dtype=np.float32)
    rois = make_rois(wl, seed)
    dy = rng.standard_normal(wl.out_shape, dtype=np.float32) if with_dy else None
    return data, rois, dy


def literal_benchmark_inputs(N=2, C=1024, H=14, W=14, R=512):
    """The workload of /root/reference/benchmark/benchmark_roi_align.py:15-21."""
    data = np.arange(N * C * H * W, dtype=np.float32).reshape(N, C, H, W)
    rois = np.array([[i % N, 0, 0, i % H, i % W] for i in range(R)], np.float32)
    return data, rois


CONFIGS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable benchmark shape
    # (the script's literal inputs; BASELINE.json quotes it at 7x7 / sr 2, the script itself runs 2x2 / sr 1)
    'C1': Workload('C1-bench-7x7-sr2', 2, 1024, 14, 14, 256, (7, 7), 1.0, 2, literal=True),
    'C1-literal': Workload('C1-bench-literal-2x2-sr1', 2, 1024, 14, 14, 256, (2, 2), 1.0, 1, literal=True),
    # configs[1]: the configuration the metric is quoted on
    'C2': Workload('C2-box-head', 2, 256, 200, 272, 512, (7, 7), 0.25, 2),
    'C3': Workload('C3-mask-head', 8, 256, 128, 128, 1000, (14, 14), 0.25, 2),
    'C4': Workload('C4-batch-sharded', 64, 256, 200, 272, 512, (7, 7), 0.25, 2),
    'C5': Workload('C5-stress-adaptive', 2, 256, 200, 272, 5000, (7, 7), 0.25, 0,
                   min_box=4.0, stress=True),
}
