"""FPN level assignment around ROIAlign ("the step before the path").

The reference has no multi-level pooler; detection code that calls its ROIAlign does this
bookkeeping in Python.  `assign_levels` is equation (1) of the FPN paper as torchvision's
MultiScaleRoIAlign / detectron2's ROIPooler implement it; `roi_align_fpn` pools every ROI from its
level with `mobula.op.ROIAlign` (one launch per level, rows scattered back in the caller's order).
Bit-exact bookkeeping: a ROI's output row is exactly what a single-level call on its level gives.
"""
import numpy as np


def assign_levels(rois, k_min=2, k_max=5, canonical_scale=224.0, canonical_level=4, eps=1e-6):
    """Pyramid level of every ROI: floor(k0 + log2(sqrt(w * h) / s0) + eps), clamped to
    [k_min, k_max].  rois: [R, 5] (batch, x1, y1, x2, y2) in image coordinates, numpy or torch."""
    if hasattr(rois, 'detach'):
        import torch
        r = rois.detach()
        area = (r[:, 3] - r[:, 1]).clamp(min=0) * (r[:, 4] - r[:, 2]).clamp(min=0)
        lvl = torch.floor(canonical_level + torch.log2(torch.sqrt(area) / canonical_scale + 1e-30) + eps)
        return lvl.clamp(k_min, k_max).to(torch.int64)
    r = np.asarray(rois, np.float32)
    area = np.maximum(r[:, 3] - r[:, 1], 0) * np.maximum(r[:, 4] - r[:, 2], 0)
    with np.errstate(divide='ignore'):
        lvl = np.floor(canonical_level + np.log2(np.sqrt(area) / np.float32(canonical_scale) + np.float32(1e-30)) + eps)
    return np.clip(lvl, k_min, k_max).astype(np.int64)


def split_by_level(levels, k_min, k_max):
    """Row indices of the ROIs of every level k_min..k_max (ascending inside a level)."""
    if hasattr(levels, 'detach'):
        import torch
        return [torch.nonzero(levels == k, as_tuple=False).reshape(-1) for k in range(k_min, k_max + 1)]
    levels = np.asarray(levels)
    return [np.nonzero(levels == k)[0] for k in range(k_min, k_max + 1)]


def roi_align_fpn(features, rois, pooled_size, scales, sampling_ratio=2, k_min=2, canonical_scale=224.0,
                  canonical_level=4, **op_kwargs):
    """Pools `rois` from a feature pyramid.  features: list of [N, C, H_k, W_k] tensors for levels
    k_min, k_min + 1, ...; scales: their spatial scales (1/4, 1/8, ...).  Returns [R, C, PH, PW] in
    the order of `rois`; differentiable with torch inputs."""
    import mobulaop_b200 as mobula
    mobula.op.load('ROIAlign')
    k_max = k_min + len(features) - 1
    levels = assign_levels(rois, k_min, k_max, canonical_scale, canonical_level)
    parts = split_by_level(levels, k_min, k_max)
    if isinstance(pooled_size, int):
        pooled_size = (pooled_size, pooled_size)
    first = features[0]
    is_torch = hasattr(first, 'detach')
    R, C = int(rois.shape[0]), int(first.shape[1])
    if is_torch:
        import torch
        out = torch.zeros((R, C) + tuple(pooled_size), dtype=first.dtype, device=first.device)
    else:
        out = np.zeros((R, C) + tuple(pooled_size), first.dtype)
    for feat, scale, idx in zip(features, scales, parts):
        if len(idx) == 0:
            continue
        y = mobula.op.ROIAlign(feat, rois[idx], pooled_size=tuple(pooled_size), spatial_scale=float(scale),
                               sampling_ratio=sampling_ratio, **op_kwargs)
        if is_torch:
            out = out.index_copy(0, idx, y)
        else:
            out[idx] = y
    return out
