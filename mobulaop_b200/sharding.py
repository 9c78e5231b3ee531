"""Batch-index sharding of a ROIAlign workload over the GPUs of one node.

A ROI reads only the feature map of image `rois[r, 0]` and its gradient lands only in
that image's slice of dX (/root/reference/opzoo/ROIAlign/ROIAlign.cpp:25, 43-44,
110-111), so the path shards by image with NO data-path collective: rank g owns a
contiguous range of images, their feature maps / gradients, and every ROI that points
into them.  The reference has no multi-GPU driver -- only a per-call device guard and a
replica test (/root/reference/tests/gpu/test_gpu.py:19-43); this is the new piece.

Bookkeeping contract (bit-exact, tested on CPU with gloo):
* `image_range(g)`    = [floor(g*N/G), floor((g+1)*N/G))
* ROIs are stably partitioned by owner: inside a shard they keep their original
  relative order, `shard.roi_index` holds their original row numbers;
* the batch index of a shard's ROIs is rebased to the shard's first image;
* ROIs whose batch index is outside [0, N) belong to no shard: their output rows are
  zero and they contribute no gradient, which is what the kernels do for them.

`torch.distributed` (NCCL over NVLink on GPUs, gloo on CPU) is used only by the optional
`gather_rows`, never on the hot path.
"""
import numpy as np


class Shard(object):
    """What one rank owns."""
    __slots__ = ('rank', 'world', 'image_begin', 'image_end', 'roi_index', 'rois')

    def __init__(self, rank, world, image_begin, image_end, roi_index, rois):
        self.rank = rank
        self.world = world
        self.image_begin = image_begin
        self.image_end = image_end
        self.roi_index = roi_index      # int64 [R_g]: rows of the global roi table
        self.rois = rois                # float32 [R_g, 5], batch index rebased

    @property
    def num_images(self):
        return self.image_end - self.image_begin

    @property
    def num_rois(self):
        return int(self.roi_index.shape[0])


def image_range(rank, world, num_images):
    if not 0 <= rank < world:
        raise ValueError('rank %d outside world of %d' % (rank, world))
    return (rank * num_images) // world, ((rank + 1) * num_images) // world


def owner_of_images(batch_index, world, num_images):
    """Rank owning each batch index (int64 array); -1 for indices outside [0, N)."""
    b = np.asarray(batch_index, np.int64)
    # smallest g with b < floor((g+1)*N/G)  <=>  g = ceil((b+1)*G/N) - 1
    owner = ((b + 1) * world + num_images - 1) // max(num_images, 1) - 1
    owner = np.where((b < 0) | (b >= num_images), -1, owner)
    return owner


def batch_indices(rois):
    """Integer batch index of every ROI exactly as the kernels read it: (int)rois[r,0],
    truncation toward zero (ROIAlign.cpp:25)."""
    col = np.asarray(rois)[:, 0]
    return np.trunc(col).astype(np.int64)


def make_shard(rois, num_images, rank, world):
    rois = np.ascontiguousarray(rois, np.float32)
    if rois.ndim != 2 or rois.shape[1] != 5:
        raise ValueError('rois must be [R, 5]')
    begin, end = image_range(rank, world, num_images)
    b = batch_indices(rois)
    mine = np.nonzero((b >= begin) & (b < end))[0].astype(np.int64)   # ascending = stable
    local = rois[mine].copy()
    local[:, 0] = (b[mine] - begin).astype(np.float32)
    return Shard(rank, world, begin, end, mine, local)


def make_shards(rois, num_images, world):
    return [make_shard(rois, num_images, g, world) for g in range(world)]


def scatter_rows(shard, rows):
    """Rows of a global per-ROI array (e.g. dy) that belong to `shard`, in shard order."""
    return np.ascontiguousarray(np.asarray(rows)[shard.roi_index])


def assemble_rows(shards, parts, num_rois):
    """Inverse of the partition: global per-ROI array from per-shard pieces.  Rows of
    ROIs owned by nobody (batch index out of range) are zero."""
    first = np.asarray(parts[0])
    out = np.zeros((num_rois,) + first.shape[1:], first.dtype)
    for shard, part in zip(shards, parts):
        out[shard.roi_index] = np.asarray(part)
    return out


def assemble_images(shards, parts):
    """Global [N, ...] array (e.g. dX) from the per-rank image slices."""
    return np.concatenate([np.asarray(p) for p in parts], axis=0)


def gather_rows(local_rows, shard, num_rois, group=None):
    """Optional verification collective: all-gathers the per-ROI outputs of every rank
    (padded to the largest shard) and returns the global array on every rank.  Works on
    whatever backend `torch.distributed` was initialised with (NCCL or gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = local_rows.device
    count = torch.tensor([shard.num_rois], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    tail = tuple(local_rows.shape[1:])
    padded = torch.zeros((cap,) + tail, dtype=local_rows.dtype, device=dev)
    padded[:shard.num_rois] = local_rows
    index = torch.full((cap,), -1, dtype=torch.int64, device=dev)
    index[:shard.num_rois] = torch.from_numpy(shard.roi_index).to(dev)
    all_rows = [torch.empty_like(padded) for _ in range(world)]
    all_index = [torch.empty_like(index) for _ in range(world)]
    dist.all_gather(all_rows, padded, group=group)
    dist.all_gather(all_index, index, group=group)
    out = torch.zeros((num_rois,) + tail, dtype=local_rows.dtype, device=dev)
    for rows, idx, n in zip(all_rows, all_index, counts):
        if n:
            out[idx[:n]] = rows[:n]
    return out
