"""Host<->device copy rates of this box with page-locked memory (the floor of the e2e number)."""
import os
import time
import torch

# under torchrun every rank probes its own GPU at the same time (aggregate host-side limit)
rank, world = int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('gloo')
dev = torch.device('cuda', rank)
torch.cuda.set_device(dev)
n = 111411200 // 4
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, dtype=torch.float32, device=dev)
d_b = torch.ones(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


gb = n * 4 / 1e9
print('rank %d of %d: ' % (rank, world) + 'H2D %.1f GB/s  D2H %.1f GB/s  duplex %.1f GB/s per direction (111 MB each way)' % (
    gb / timed(h2d), gb / timed(d2h), gb / timed(both)))
