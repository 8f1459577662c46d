// Microbenchmark: fraction of exponentials emulated on the FMA/ALU pipes (Cody-Waite + polynomial) next to MUFU.EX2,
// inside the softmax instruction mix (scale FFMA2 done beforehand, as in attn_tc3). Also checks the emulation's accuracy.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)), "l"(reinterpret_cast<uint64_t&>(c))); return d; }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b))); return d; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

// exp2 of a pair on the FMA/ALU pipes. x <= ~10; clamped at -126 (flushes to ~1e-38 instead of 0: invisible after bf16 rounding
// relative to a row sum >= 1). DEG 2: rel. error 1.8e-3 (below bf16's 3.9e-3 half-ulp spacing); DEG 3: 1.1e-4.
template <int DEG>
__device__ __forceinline__ float2 exp2_emul(float2 x) {
  const float2 magic = make_float2(12582912.f, 12582912.f), nmagic = make_float2(-12582912.f, -12582912.f), mone = make_float2(-1.f, -1.f);
  x.x = fmaxf(x.x, -126.f); x.y = fmaxf(x.y, -126.f);
  float2 r = fadd2(x, magic);          // integer part in the low mantissa bits
  float2 n = fadd2(r, nmagic);
  float2 f = ffma2(n, mone, x);        // fraction in [-0.5, 0.5]
  float2 p;
  if (DEG == 2) {
    p = ffma2(make_float2(0.2402265f, 0.2402265f), f, make_float2(0.6931472f, 0.6931472f));
    p = ffma2(p, f, make_float2(1.0017f, 1.0017f));   // minimax-shifted constant term
    p = ffma2(make_float2(0.2439f, 0.2439f), f, make_float2(0.69584f, 0.69584f));
    p = ffma2(p, f, make_float2(0.99992f, 0.99992f));
  } else {
    p = ffma2(make_float2(0.05550411f, 0.05550411f), f, make_float2(0.2402265f, 0.2402265f));
    p = ffma2(p, f, make_float2(0.6931472f, 0.6931472f));
    p = ffma2(p, f, make_float2(1.0f, 1.0f));
  }
  float2 o;
  o.x = __uint_as_float((__float_as_uint(r.x) << 23) + __float_as_uint(p.x));
  o.y = __uint_as_float((__float_as_uint(r.y) << 23) + __float_as_uint(p.y));
  return o;
}

// EMU = number of emulated pairs out of every 8 pairs (16 elements); DEG = polynomial degree
template <int EMU, int DEG>
__global__ void k(float* out, long long* cyc, int iters) {
  float x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = -0.37f * (threadIdx.x % 13) - 0.11f * i;
  float2 acc0 = make_float2(0, 0), acc1 = make_float2(0, 0);
  uint32_t pk = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float2 v = make_float2(x[i], x[i + 1]);
      const int pair = (i / 2) % 8;
      float2 e;
      if (pair < EMU) e = exp2_emul<DEG>(v);
      else { e.x = ex2(v.x); e.y = ex2(v.y); }
      if (pair & 1) acc1 = fadd2(acc1, e); else acc0 = fadd2(acc0, e);
      pk ^= pack(e.x, e.y);
      x[i] = e.x * 0.5f - 1.f - 0.11f * i; x[i + 1] = e.y * 0.5f - 1.3f;
    }
  }
  long long t1 = clock64();
  float s = acc0.x + acc0.y + acc1.x + acc1.y + __uint_as_float(pk & 0x3f800000u);
#pragma unroll
  for (int i = 0; i < 32; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int DEG> __global__ void acc_kernel(float* err) {
  float worst = 0.f;
  for (int i = threadIdx.x; i < 2000000; i += blockDim.x) {
    float x = -20.f + 28.f * (float)i / 2000000.f;
    float2 e = exp2_emul<DEG>(make_float2(x, x));
    float ref = exp2f(x);
    worst = fmaxf(worst, fabsf(e.x - ref) / ref);
  }
  atomicMax(reinterpret_cast<int*>(err), __float_as_int(worst));
}

template <int EMU, int DEG> void run(int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 1000;
  k<EMU, DEG><<<148, threads>>>(out, cyc, iters);
  k<EMU, DEG><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("emulated %d/8 pairs, degree %d, %d warp(s)/SMSP: %.2f cyc per element-column per SMSP\n", EMU, DEG, threads / 128, (double)h[0] / iters / 32 / (threads / 128));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  float* err; cudaMalloc(&err, 4);
  cudaMemset(err, 0, 4); acc_kernel<2><<<1, 256>>>(err); float e2; cudaMemcpy(&e2, err, 4, cudaMemcpyDeviceToHost);
  cudaMemset(err, 0, 4); acc_kernel<3><<<1, 256>>>(err); float e3; cudaMemcpy(&e3, err, 4, cudaMemcpyDeviceToHost);
  printf("max relative error of the emulation on [-20, 8]: degree 2: %.3e   degree 3: %.3e\n", e2, e3);
  for (int th : {128, 256}) {
    run<0, 3>(th); run<1, 3>(th); run<2, 3>(th); run<3, 3>(th); run<4, 3>(th);
    run<1, 2>(th); run<2, 2>(th); run<3, 2>(th);
  }
  return 0;
}
