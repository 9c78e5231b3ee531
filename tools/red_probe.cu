// Micro-benchmark: throughput of global fp32 reductions (RED.ADD) on B200, scalar vs
// vector (v2/v4), addresses spread like a ROIAlign backward (random rows of a 111 MB map).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_probe red_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int V>
__global__ void red_kernel(float* buf, size_t elems, int iters, int spread) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  const int lane = threadIdx.x & 31;
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    // a warp hits one random row; lanes `spread` floats apart along it
    uint32_t w = __shfl_sync(0xffffffffu, s, 0);
    size_t base = ((size_t)(w >> 4) % (elems / 4096)) * 4096 + (size_t)lane * spread;
    base &= ~(size_t)(V - 1);
    float* p = buf + base;
    if (V == 1) {
      atomicAdd(p, 1.0f);
    } else if (V == 2) {
      atomicAdd(reinterpret_cast<float2*>(p), make_float2(1.f, 1.f));
    } else {
      atomicAdd(reinterpret_cast<float4*>(p), make_float4(1.f, 1.f, 1.f, 1.f));
    }
  }
}

int main() {
  const size_t elems = (size_t)2 * 256 * 200 * 272;
  float* buf;
  cudaMalloc(&buf, elems * 4 + 65536);
  cudaMemset(buf, 0, elems * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = 148 * 8, threads = 256, iters = 200;
  for (int spread : {1, 4, 18}) {
    for (int v : {1, 2, 4}) {
      float best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (v == 1) red_kernel<1><<<blocks, threads>>>(buf, elems, iters, spread);
        if (v == 2) red_kernel<2><<<blocks, threads>>>(buf, elems, iters, spread);
        if (v == 4) red_kernel<4><<<blocks, threads>>>(buf, elems, iters, spread);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      double ops = (double)blocks * threads * iters;
      printf("spread %2d  v%d: %.1f us  %.1f G lane-ops/s  %.1f G floats/s  %s\n", spread, v, best * 1e3,
             ops / (best * 1e-3) / 1e9, ops * v / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
