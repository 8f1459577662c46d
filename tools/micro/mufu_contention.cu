// Does mbarrier.try_wait polling by idle warps slow the MUFU stream of the working warps? (both go through the MIO queue)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_contention mufu_contention.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)), "l"(reinterpret_cast<uint64_t&>(c))); return d; }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b))); return d; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done;
}
__device__ __forceinline__ uint32_t try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity), "r"(ns) : "memory");
  return done;
}

// MODE 0: waiters exit immediately; 1: waiters spin on try_wait; 2: try_wait with a 20 us suspend hint; 3: waiters block in bar.sync
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, int work_threads) {
  __shared__ uint64_t bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1)); }
  __syncthreads();
  if ((int)threadIdx.x >= work_threads) {
    if (MODE == 1) { while (!try_wait(b, 0)) {} }
    if (MODE == 2) { while (!try_wait_hint(b, 0, 20000)) {} }
    if (MODE == 3) { asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x)); }
    return;
  }
  float x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = -0.001f * (threadIdx.x + i);
  float2 acc0 = make_float2(0, 0), acc1 = make_float2(0, 0);
  uint32_t pk = 0;
  float2 c = make_float2(0.99f, 0.99f), m = make_float2(-0.01f, -0.01f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float2 a = ffma2(make_float2(x[i], x[i + 1]), c, m), bb = ffma2(make_float2(x[i + 2], x[i + 3]), c, m);
      a.x = ex2(a.x); a.y = ex2(a.y); bb.x = ex2(bb.x); bb.y = ex2(bb.y);
      acc0 = fadd2(acc0, a); acc1 = fadd2(acc1, bb);
      pk ^= pack(a.x, a.y) ^ pack(bb.x, bb.y);
      x[i] = a.x; x[i + 1] = a.y; x[i + 2] = bb.x; x[i + 3] = bb.y;
    }
  }
  long long t1 = clock64();
  float s = acc0.x + acc0.y + acc1.x + acc1.y + __uint_as_float(pk & 0x3f800000u);
#pragma unroll
  for (int i = 0; i < 32; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  // release the waiters
  asm volatile("bar.sync 2, %0;" ::"r"(work_threads));
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
  if (MODE == 3) asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x));
}

template <int MODE>
void run(const char* name, int work, int waiters) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 500;
  k<MODE><<<148, work + waiters>>>(out, cyc, iters, work);
  k<MODE><<<148, work + waiters>>>(out, cyc, iters, work);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = (double)h[0] / iters / 32 / (work / 128);
  printf("%-34s work warps/SMSP=%d waiter warps/SMSP=%d : %.2f cyc per MUFU warp-instr per SMSP (%s)\n", name, work / 128, waiters / 128, c, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int work : {128, 256}) {
    run<0>("no waiters", work, 0);
    for (int w : {128, 256, 512}) {
      if (work + w > 1024) continue;
      run<1>("waiters spin on try_wait", work, w);
      run<2>("waiters try_wait + 20us hint", work, w);
      run<3>("waiters blocked in bar.sync", work, w);
    }
  }
  return 0;
}
