// Shared host/device helpers for libief_b200.so (B200 / sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <atomic>

#include "../../include/ief_b200.h"

// ---- error plumbing (thread-local message, see api.cu) -------------------------------------------
void ief_set_error(const char* fmt, ...);
void ief_count_launch(int n = 1);

#define IEF_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ief_set_error(__VA_ARGS__);      \
      return (code);                   \
    }                                  \
  } while (0)

#define IEF_CUDA_OK(expr)                                                               \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ief_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return IEF_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

// Launch-error check without synchronising (sticky errors surface on the next call).
#define IEF_LAUNCH_OK(name)                                                             \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      ief_set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));          \
      return IEF_ERR_CUDA;                                                              \
    }                                                                                   \
    ief_count_launch();                                                                 \
  } while (0)

// ---- element-type traits ---------------------------------------------------------------------------
template <int DTYPE> struct ElemT;
template <> struct ElemT<IEF_BF16> {
  using T = __nv_bfloat16;
  using T2 = __nv_bfloat162;
  static constexpr bool kIsBf16 = true;
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  static __device__ __forceinline__ float lo(uint32_t v) { return __uint_as_float(v << 16); }
  static __device__ __forceinline__ float hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
};
template <> struct ElemT<IEF_F16> {
  using T = __half;
  using T2 = __half2;
  static constexpr bool kIsBf16 = false;
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  static __device__ __forceinline__ float lo(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v & 0xffffu))); }
  static __device__ __forceinline__ float hi(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v >> 16))); }
};

// Per-row source table handed to the attention kernels by value (kernel parameter space).
struct IefRowTable {
  int32_t q[IEF_MAX_ROWS];
  int32_t k[IEF_MAX_ROWS];
  int32_t v[IEF_MAX_ROWS];
  int32_t k2[IEF_MAX_ROWS];
  int32_t v2[IEF_MAX_ROWS];
  int32_t pslot[IEF_MAX_ROWS];
  int32_t bias[IEF_MAX_ROWS];  // key_bias vector of the row, -1 = none
  uint8_t active[IEF_MAX_ROWS];
};

__device__ __forceinline__ float ief_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

static inline int ief_ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- per-device launch configuration ---------------------------------------------------------------
// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (kernel, device): one process may drive several GPUs, so the
// "already configured" memo is a bit per device ordinal and per call site (= per kernel template instance), never per process.
#define IEF_CONFIG_SMEM(kern, bytes)                                                                                   \
  do {                                                                                                                 \
    static std::atomic<uint64_t> done__[4];                                                                            \
    int dev__ = 0;                                                                                                     \
    IEF_CUDA_OK(cudaGetDevice(&dev__));                                                                                \
    const uint64_t bit__ = 1ull << (dev__ & 63);                                                                       \
    std::atomic<uint64_t>& w__ = done__[(dev__ >> 6) & 3];                                                             \
    if (!(w__.load(std::memory_order_acquire) & bit__)) {                                                              \
      IEF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));              \
      w__.fetch_or(bit__, std::memory_order_release);                                                                  \
    }                                                                                                                  \
  } while (0)

// SM count of the CURRENT device (cached per device ordinal; 148 on B200)
static inline int ief_sm_count() {
  static std::atomic<int> cache[256];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::atomic<int>& c = cache[dev & 255];
  int n = c.load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    c.store(n, std::memory_order_relaxed);
  }
  return n;
}

// internal launchers (one per .cu)
int ief_attn_mma_launch(const ief_attn_params* p, const IefRowTable& rows, cudaStream_t st);
int ief_attn_tc_launch(const ief_attn_params* p, const IefRowTable& rows, cudaStream_t st, float* lse_out = nullptr);
int ief_attn_probs_from_lse_launch(const ief_attn_params* p, const IefRowTable& rows, const float* lse, cudaStream_t st);
bool ief_attn_probs_via_lse(const ief_attn_params* p);  // stored maps: tcgen05 (O + lse) followed by one QK^T sweep, given a workspace
bool ief_attn_tc_supported(const ief_attn_params* p, const char** why);
