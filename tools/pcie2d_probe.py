"""Rate of strided (2-D) page-locked copies as the channel-chunked host path issues them."""
import ctypes
import glob
import os
import time
import torch

dev = torch.device('cuda', 0)
torch.cuda.init()
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), '..', 'nvidia', 'cuda_runtime', 'lib', 'libcudart.so*')) + \
    glob.glob('/usr/local/cuda/lib64/libcudart.so*')
rt = ctypes.CDLL(cands[0])
R, C, B = 1024, 256, 49
pitch = C * B * 4
h = torch.empty(R * C * B, dtype=torch.float32).pin_memory()
d = torch.ones(R * C * B, dtype=torch.float32, device=dev)
stream = torch.cuda.Stream(dev)
rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                 ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]


def run(chunks, kind):
    cc = C // chunks
    width = cc * B * 4
    def once():
        for k in range(chunks):
            off = k * cc * B * 4
            dst, src = (h.data_ptr() + off, d.data_ptr() + off) if kind == 2 else (d.data_ptr() + off, h.data_ptr() + off)
            rc = rt.cudaMemcpy2DAsync(dst, pitch, src, pitch, width, R, kind, stream.cuda_stream)
            assert rc == 0, rc
    once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        once()
    torch.cuda.synchronize()
    return R * C * B * 4 / ((time.perf_counter() - t0) / 5) / 1e9


for chunks in (1, 4, 8, 16):
    print('chunks %2d (rows of %6d B): D2H %.1f GB/s  H2D %.1f GB/s' % (
        chunks, C // chunks * B * 4, run(chunks, 2), run(chunks, 1)))
