// Microbenchmark: MUFU.EX2 throughput per SM sub-partition on sm_100a, alone and inside the softmax instruction mix.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)), "l"(reinterpret_cast<uint64_t&>(c))); return d; }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b))); return d; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = -0.001f * (threadIdx.x + i);
  float2 acc0 = make_float2(0, 0), acc1 = make_float2(0, 0);
  uint32_t pk = 0;
  float2 c = make_float2(0.99f, 0.99f), m = make_float2(-0.01f, -0.01f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // pure MUFU, 32 independent chains
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = ex2(x[i]);
    } else if (MODE == 1) {  // softmax mix: FFMA2 + 2 MUFU + FADD2 + F2FP per pair
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float2 a = ffma2(make_float2(x[i], x[i + 1]), c, m), b = ffma2(make_float2(x[i + 2], x[i + 3]), c, m);
        a.x = ex2(a.x); a.y = ex2(a.y); b.x = ex2(b.x); b.y = ex2(b.y);
        acc0 = fadd2(acc0, a); acc1 = fadd2(acc1, b);
        pk ^= pack(a.x, a.y) ^ pack(b.x, b.y);
        x[i] = a.x; x[i + 1] = a.y; x[i + 2] = b.x; x[i + 3] = b.y;
      }
    } else if (MODE == 2) {  // FMA-pipe exp2 emulation (Cody-Waite + degree-3 polynomial), packed f32x2, no MUFU
      const float2 magic = make_float2(12582912.f, 12582912.f), nmagic = make_float2(-12582912.f, -12582912.f);
      const float2 c3 = make_float2(0.0555f, 0.0555f), c2 = make_float2(0.2402f, 0.2402f), c1 = make_float2(0.6931f, 0.6931f), c0 = make_float2(1.f, 1.f);
      const float2 one = make_float2(1.f, 1.f), mone = make_float2(-1.f, -1.f);
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float2 v = ffma2(make_float2(x[i], x[i + 1]), c, m);
        float2 r = fadd2(v, magic);
        float2 fi = fadd2(r, nmagic);
        float2 f = ffma2(fi, mone, v);
        float2 p = ffma2(c3, f, c2);
        p = ffma2(p, f, c1);
        p = ffma2(p, f, c0);
        uint32_t e0 = (__float_as_uint(r.x) << 23) + __float_as_uint(p.x), e1 = (__float_as_uint(r.y) << 23) + __float_as_uint(p.y);
        float2 q = make_float2(__uint_as_float(e0), __uint_as_float(e1));
        acc0 = fadd2(acc0, q);
        pk ^= pack(q.x, q.y);
        x[i] = q.x * 1e-3f - 1.f; x[i + 1] = q.y * 1e-3f - 1.f;
      }
    } else if (MODE == 3) {  // half MUFU, half emulated
      const float2 magic = make_float2(12582912.f, 12582912.f), nmagic = make_float2(-12582912.f, -12582912.f);
      const float2 c3 = make_float2(0.0555f, 0.0555f), c2 = make_float2(0.2402f, 0.2402f), c1 = make_float2(0.6931f, 0.6931f), c0 = make_float2(1.f, 1.f);
      const float2 mone = make_float2(-1.f, -1.f);
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float2 a = ffma2(make_float2(x[i], x[i + 1]), c, m);
        a.x = ex2(a.x); a.y = ex2(a.y);
        acc0 = fadd2(acc0, a);
        pk ^= pack(a.x, a.y);
        float2 v = ffma2(make_float2(x[i + 2], x[i + 3]), c, m);
        float2 r = fadd2(v, magic);
        float2 fi = fadd2(r, nmagic);
        float2 f = ffma2(fi, mone, v);
        float2 p = ffma2(c3, f, c2);
        p = ffma2(p, f, c1);
        p = ffma2(p, f, c0);
        uint32_t e0 = (__float_as_uint(r.x) << 23) + __float_as_uint(p.x), e1 = (__float_as_uint(r.y) << 23) + __float_as_uint(p.y);
        float2 q = make_float2(__uint_as_float(e0), __uint_as_float(e1));
        acc1 = fadd2(acc1, q);
        pk ^= pack(q.x, q.y);
        x[i] = a.x; x[i + 1] = a.y; x[i + 2] = q.x * 1e-3f - 1.f; x[i + 3] = q.y * 1e-3f - 1.f;
      }
    } else if (MODE == 4) {  // max phase: FMNMX3 over 128 values, 4 chains
      float m0 = x[0], m1 = x[1], m2 = x[2], m3 = x[3];
#pragma unroll
      for (int rep = 0; rep < 4; ++rep)
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m0) : "f"(x[i]), "f"(x[i + 1]));
          asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m1) : "f"(x[i + 2]), "f"(x[i + 3]));
          asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m2) : "f"(x[i + 4]), "f"(x[i + 5]));
          asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m3) : "f"(x[i + 6]), "f"(x[i + 7]));
        }
      x[0] = m0 + m1 + m2 + m3;
    }
  }
  long long t1 = clock64();
  float s = acc0.x + acc0.y + acc1.x + acc1.y + __uint_as_float(pk & 0x3f800000u);
#pragma unroll
  for (int i = 0; i < 32; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, int elems_per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  k<MODE><<<148, threads>>>(out, cyc, iters);
  k<MODE><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = (double)h[0] / iters;
  int warps_per_smsp = threads / 128;
  printf("%-28s threads=%4d (%d warp/SMSP): %.1f cyc/iter -> %.2f cyc per element-column per warp; %.2f cyc per element-column per SMSP\n", name, threads,
         warps_per_smsp, c, c / elems_per_iter, c / elems_per_iter / warps_per_smsp);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int th : {128, 256, 512}) {
    run<0>("pure MUFU.EX2", th, 32);
    run<1>("softmax mix (MUFU)", th, 32);
    run<2>("softmax mix (FMA emulation)", th, 32);
    run<3>("softmax mix (50/50)", th, 32);
    run<4>("FMNMX3 max over 128", th, 128);
  }
  return 0;
}
