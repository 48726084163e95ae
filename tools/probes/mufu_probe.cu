// Throughput of MUFU.EX2 as f32 and as packed f16x2 on sm_100a (per SM per clock), to decide whether the softmax of the
// spatial attention should exponentiate two probabilities per MUFU instruction.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_probe mufu_probe.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(float* out, int iters, float seed) {
  float a[8];
  unsigned h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; h[i] = 0x3c003c00u + i + threadIdx.x; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("ex2.approx.f16 %0, %0;" : "+h"(*(unsigned short*)&h[i]));
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}

template <int MODE>
void run(const char* name, int elems_per_instr) {
  float* d;
  cudaMalloc(&d, 148 * 1024 * 4);
  const int iters = 4096;
  probe<MODE><<<148, 1024>>>(d, iters, 0.5f);
  cudaDeviceSynchronize();
  probe<MODE><<<148, 1024>>>(d, iters, 0.5f);
  cudaError_t e = cudaDeviceSynchronize();
  float cyc;
  cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
  const double instr = (double)iters * 8 * 1024;  // thread-instructions per SM
  printf("%-28s %s  %.2f thread-instr/clk/SM  %.2f elements/clk/SM\n", name, cudaGetErrorString(e), instr / cyc, instr * elems_per_instr / cyc);
  cudaFree(d);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("ex2.approx.f16", 1);
  return 0;
}
