// Microbenchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16, cta_group::1) as a function of N and of where A comes from
// (shared-memory descriptor vs TMEM). One CTA per SM, one elected thread issues `G` MMAs into distinct accumulator columns per commit.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I image_editing_framework_b200/csrc -o umma_rate tools/micro/umma_rate.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace sm100;

template <bool A_TMEM>
__global__ void __launch_bounds__(128) k(long long* out, int N, int G, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, bar = base + 49152, slot = bar + 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 49152 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot_ptr;
  const uint32_t idesc = make_idesc_f16(128, N, 1, 0, 0);
  const uint64_t dA = make_smem_desc_sw128(sA, 16, 1024), dB = make_smem_desc_sw128(sB, 16, 1024);
  if (warp == 1) {
    long long t0 = 0, t1 = 0;
    for (int it = 0; it < iters + 2; ++it) {
      if (it == 2) t0 = clock64();
      if (elect_one()) {
        for (int g = 0; g < G; ++g) {
          if (A_TMEM) umma_ts(tm + (g & 1) * 128, tm + 256 + (g & 3) * 8, dB + ((g & 3) * 2), idesc, 1);
          else umma_ss(tm + (g & 1) * 128, dA + ((g & 3) * 2), dB + ((g & 3) * 2), idesc, 1);
        }
        umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, it & 1);
    }
    t1 = clock64();
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  const int iters = 200;
  for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
    for (int G : {1, 8, 16})
      for (int N : {16, 32, 48, 64, 96, 128, 256}) {
        if (a_tmem) k<true><<<148, 128, 60000>>>(d, N, G, iters); else k<false><<<148, 128, 60000>>>(d, N, G, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        const double per = (double)h[0] / iters / G;
        printf("A from %s  N=%3d  %2d MMA per commit: %7.1f clk per MMA (%.0f MAC/clk; ideal at 3868 MAC/clk: %.1f clk)\n", a_tmem ? "TMEM" : "smem", N, G, per,
               128.0 * N * 16 / per, 128.0 * N * 16 / 3868);
      }
  return 0;
}
