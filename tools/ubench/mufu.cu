// Microbenchmark: MUFU.EX2 throughput per SM for f32 / f16x2 / bf16x2 operands (B200).  nvcc -arch=sm_100a -O3 mufu.cu -o mufu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(uint32_t* out, int iters, long long* cyc) {
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = 0x3c003800u + threadIdx.x + i;  // some half2 / float bit patterns
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r[i]));
      if (MODE == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(r[i]));
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int elems_per_inst) {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  k<MODE><<<148, 512>>>(out, iters, cyc);
  k<MODE><<<148, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double inst = (double)iters * 8 * 512;  // thread-instructions per SM
  printf("%-10s %8lld cycles/SM : %.2f thread-instr/clk/SM = %.2f elements/clk/SM  (%s)\n", name, h[0], inst / h[0], inst * elems_per_inst / h[0],
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  run<0>("f32", 1);
  run<1>("f16x2", 2);
  run<2>("bf16x2", 2);
  run<3>("tanh.f16x2", 2);
  return 0;
}
