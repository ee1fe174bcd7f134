// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100) issue throughput on B200.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_probe tools/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float a, float b) {
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float2 x = make_float2(a, a * 1.0001f), y = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) {
        acc[i].x = fmaf(acc[i].x, x.x, y.x);
        acc[i].y = fmaf(acc[i].y, x.y, y.y);
      } else {
        acc[i] = __ffma2_rn(acc[i], x, y);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  const int blocks = 148 * 8, iters = 4096;
  cudaMalloc(&out, blocks * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) probe<0><<<blocks, 256>>>(out, iters, 0.999f, 0.001f);
      else probe<1><<<blocks, 256>>>(out, iters, 0.999f, 0.001f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)blocks * 256 * iters * 32;
      printf("%s rep %d: %.3f ms  %.1f TFLOP/s fp32\n", mode ? "FFMA2 (f32x2)" : "FFMA scalar ", rep, ms, 2 * fma / ms / 1e9);
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
