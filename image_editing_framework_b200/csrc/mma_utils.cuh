// Warp-level mma.sync / ldmatrix helpers shared by the mma flash kernel and the cross-attention edit kernel.
#pragma once
#include "ief_common.cuh"

namespace mmau {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}

template <int DTYPE>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (DTYPE == IEF_BF16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

// Cooperative copy of a [ROWS x DP] 16-bit tile (row stride LD in smem) from a strided global view.
// Rows >= n_valid and channels >= d are zero-filled. d % 8 == 0, 16-byte vectors.
template <typename T, int ROWS, int DP, int LD, int NTHREADS>
__device__ __forceinline__ void load_tile(T* s, const T* g, int64_t g_row_stride, int row0, int n_valid, int d, int tid) {
  constexpr int CPR = DP / 8;  // 16-byte chunks per row
  for (int i = tid; i < ROWS * CPR; i += NTHREADS) {
    const int r = i / CPR, c = (i % CPR) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row0 + r < n_valid && c < d) v = __ldg(reinterpret_cast<const uint4*>(g + (int64_t)(row0 + r) * g_row_stride + c));
    *reinterpret_cast<uint4*>(s + r * LD + c) = v;
  }
}

// Same tile copy with cp.async (LDGSTS): no register staging, every 16-byte piece of the tile is in flight at once, so a
// tile costs one memory round trip instead of one per loop iteration. Out-of-range pieces are zero-filled (src-size 0).
// Complete with cp_async_commit() + cp_async_wait<N>() + a barrier.
template <typename T, int ROWS, int DP, int LD, int NTHREADS>
__device__ __forceinline__ void load_tile_async(T* s, const T* g, int64_t g_row_stride, int row0, int n_valid, int d, int tid) {
  constexpr int CPR = DP / 8;
#pragma unroll
  for (int i0 = 0; i0 < ROWS * CPR; i0 += NTHREADS) {
    const int i = i0 + tid;
    if (i < ROWS * CPR) {
      const int r = i / CPR, c = (i % CPR) * 8;
      const bool ok = row0 + r < n_valid && c < d;
      const T* src = ok ? g + (int64_t)(row0 + r) * g_row_stride + c : g;
      const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(s + r * LD + c));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

}  // namespace mmau
