// Microbenchmark: do the 16-bit MUFU.EX2 forms run faster than the fp32 one on sm_100a? (ex2.approx.f16 / .ftz.bf16 / x2 forms)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_h mufu_h.cu && ./mufu_h
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, int iters) {
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = 0xb800b800u + threadIdx.x + i;  // two small negative halves
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) { float f = __uint_as_float(x[i]); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f)); x[i] = __float_as_uint(f); }
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 3) { uint16_t h = (uint16_t)x[i]; asm volatile("ex2.approx.f16 %0, %0;" : "+h"(h)); x[i] = h; }
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, int results_per_instr) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  k<MODE><<<148, threads>>>(out, cyc, iters);
  k<MODE><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = (double)h[0] / iters / 16;
  int w = threads / 128;
  printf("%-24s %d warp/SMSP: %.2f clk per PTX instr per SMSP -> %.2f clk per exp result per SMSP\n", name, w, c / w, c / w / results_per_instr);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int th : {128, 256}) {
    run<0>("ex2.approx.ftz.f32", th, 1);
    run<1>("ex2.approx.f16x2", th, 2);
    run<2>("ex2.approx.ftz.bf16x2", th, 2);
    run<3>("ex2.approx.f16", th, 1);
  }
  return 0;
}
