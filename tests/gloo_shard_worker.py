"""Worker of test_two_rank_gloo_driver: world_size ranks on CPU over gloo."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mobulaop_b200 import sharding as S      # noqa: E402
from oracle.roialign import COracle          # noqa: E402


def main():
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(11)          # same seed on every rank = same global problem
    N, C, H, W, R = 6, 2, 10, 14, 80
    data = rng.standard_normal((N, C, H, W)).astype(np.float32)
    rois = rng.uniform(0, 50, (R, 5)).astype(np.float32)
    rois[:, 0] = rng.integers(0, N, R)
    dy = rng.standard_normal((R, C, 2, 2)).astype(np.float32)
    oracle = COracle()
    shard = S.make_shard(rois, N, rank, world)
    local = np.ascontiguousarray(data[shard.image_begin:shard.image_end])
    y_local = oracle.forward(local, shard.rois, (2, 2), 0.25, 2)
    dx_local = oracle.backward(S.scatter_rows(shard, dy), shard.rois, local.shape, 0.25, 2)
    y_all = S.gather_rows(torch.from_numpy(y_local), shard, R).numpy()
    y_ref = oracle.forward(data, rois, (2, 2), 0.25, 2)
    dx_ref = oracle.backward(dy, rois, data.shape, 0.25, 2)
    assert np.array_equal(y_all, y_ref), 'gathered forward differs'
    assert np.array_equal(dx_local, dx_ref[shard.image_begin:shard.image_end]), 'dX slice differs'
    # weak-scaling timing protocol of bench.py: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    dist.barrier()
    if rank == 0:
        print('SHARD-OK')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
