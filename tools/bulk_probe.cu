// Micro-benchmark: how fast can one CTA per SM pull a 200x272 fp32 plane into shared memory?
//   mode 0: cp.async.bulk, 32 KB chunks        mode 1: cp.async.bulk, one copy per row (pitch 276)
//   mode 2: LDG.128 + STS.128 by 1024 threads  mode 3: mode 2 with 512 threads
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_probe bulk_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../mobulaop_b200/csrc/ptx_sm100.cuh"
using namespace mobula_b200;

__global__ void __launch_bounds__(1024, 1)
probe(const float* __restrict__ data, float* __restrict__ sink, int planes, int height, int width, int pitch,
      int mode) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  float* plane = reinterpret_cast<float*>(smem_raw);
  const int tid = threadIdx.x;
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  __syncthreads();
  uint32_t parity = 0;
  float acc = 0.f;
  for (int p = blockIdx.x; p < planes; p += gridDim.x) {
    const float* src = data + (size_t)p * height * width;
    __syncthreads();
    if (mode <= 1) {
      if (tid < 32) {
        const uint32_t row_bytes = width * 4u;
        if (tid == 0) { ptx::fence_proxy_async(); ptx::mbar_arrive_expect_tx(&bar, row_bytes * height); }
        __syncwarp();
        if (mode == 0) {
          const uint32_t bytes = row_bytes * height, chunk = 32768u;
          for (uint32_t off = tid * chunk; off < bytes; off += 32u * chunk)
            ptx::bulk_g2s((char*)plane + off, (const char*)src + off, min(chunk, bytes - off), &bar);
        } else {
          for (int r = tid; r < height; r += 32)
            ptx::bulk_g2s(plane + (size_t)r * pitch, src + (size_t)r * width, row_bytes, &bar);
        }
      }
      ptx::mbar_wait(&bar, parity);
      parity ^= 1u;
    } else {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(plane);
      const int n4 = height * width / 4;
      for (int i = tid; i < n4; i += blockDim.x) d4[i] = __ldg(s4 + i);
      __syncthreads();
    }
    acc += plane[(tid * 53) % (height * width)];
  }
  if (acc == 12345.678f) sink[0] = acc;
}

int main() {
  const int H = 200, W = 272, planes = 512;
  float *data, *sink;
  cudaMalloc(&data, (size_t)planes * H * W * 4);
  cudaMalloc(&sink, 4);
  cudaMemset(data, 0, (size_t)planes * H * W * 4);
  char* flush; cudaMalloc(&flush, 256 << 20);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 4; ++mode) {
    const int threads = mode == 3 ? 512 : 1024;
    const int pitch = mode == 1 ? 276 : 272;
    for (int grid : {148, 296, 512}) {
      if (grid != 148 && mode != 0 && mode != 2) continue;
      float best = 1e9;
      for (int it = 0; it < 5; ++it) {
        cudaMemsetAsync(flush, it, 256 << 20);
        cudaEventRecord(e0);
        probe<<<grid, threads, (size_t)H * pitch * 4, 0>>>(data, sink, planes, H, W, pitch, mode);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      cudaError_t err = cudaGetLastError();
      printf("mode %d grid %3d threads %4d: %.1f us  (%.0f GB/s)  %s\n", mode, grid, threads, best * 1e3,
             (double)planes * H * W * 4 / (best * 1e-3) / 1e9, cudaGetErrorString(err));
    }
  }
  return 0;
}
