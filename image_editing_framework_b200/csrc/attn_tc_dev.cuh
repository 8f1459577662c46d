// Device helpers shared by the second- and third-generation tcgen05 attention kernels.
#pragma once
#include "ief_common.cuh"

template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)), "l"(reinterpret_cast<const uint64_t&>(c)));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
  return d;
}

// one 32-column chunk of scores -> probabilities (packed 16-bit pairs), accumulating the fp32 row sum. bf16: the sum is taken over the
// same 16-bit values the PV MMA multiplies (see exp_pack_chunk_mix below for why)
template <typename E>
__device__ __forceinline__ void exp_chunk(const uint32_t (&s)[32], uint32_t (&u)[16], float2 c2, float2 nmc, float2& acc0, float2& acc1) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    float2 x0 = ffma2(make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), c2, nmc);
    float2 x1 = ffma2(make_float2(__uint_as_float(s[2 * i + 2]), __uint_as_float(s[2 * i + 3])), c2, nmc);
    x0.x = ief_exp2(x0.x); x0.y = ief_exp2(x0.y);
    x1.x = ief_exp2(x1.x); x1.y = ief_exp2(x1.y);
    if constexpr (E::kIsBf16) {
      const uint32_t a0 = __float_as_uint(x0.x) & 0xffff0000u, b0 = __float_as_uint(x0.y) & 0xffff0000u;
      const uint32_t a1 = __float_as_uint(x1.x) & 0xffff0000u, b1 = __float_as_uint(x1.y) & 0xffff0000u;
      x0 = make_float2(__uint_as_float(a0), __uint_as_float(b0));
      x1 = make_float2(__uint_as_float(a1), __uint_as_float(b1));
      u[i] = __byte_perm(a0, b0, 0x7632);
      u[i + 1] = __byte_perm(a1, b1, 0x7632);
    } else {
      u[i] = E::pack(x0.x, x0.y);
      u[i + 1] = E::pack(x1.x, x1.y);
    }
    acc0 = fadd2(acc0, x0);
    acc1 = fadd2(acc1, x1);
  }
}

__device__ __forceinline__ float max_chunk(const uint32_t (&s)[32], float m) {
  float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    m0 = fmax3(m0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
    m1 = fmax3(m1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
    m2 = fmax3(m2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
    m3 = fmax3(m3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
  }
  return fmaxf(fmax3(m0, m1, m2), m3);
}

__device__ __forceinline__ void mask_chunk(uint32_t (&s)[32], int col0, int vc) {
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (col0 + i >= vc) s[i] = 0xff800000u;  // -inf
}


// The same in two steps, so the FMA-pipe half (scale and shift, in place) can run before a softmax warp takes its turn on
// the MUFU and only the exponentials themselves sit inside the turn.
__device__ __forceinline__ void scale_chunk(uint32_t (&s)[32], float2 c2, float2 nmc) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const float2 x = ffma2(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), c2, nmc);
    s[i] = __float_as_uint(x.x);
    s[i + 1] = __float_as_uint(x.y);
  }
}
template <typename E>
__device__ __forceinline__ void exp_pack_chunk(const uint32_t (&s)[32], uint32_t (&u)[16], float2& acc0, float2& acc1) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    float2 x0 = make_float2(ief_exp2(__uint_as_float(s[2 * i])), ief_exp2(__uint_as_float(s[2 * i + 1])));
    float2 x1 = make_float2(ief_exp2(__uint_as_float(s[2 * i + 2])), ief_exp2(__uint_as_float(s[2 * i + 3])));
    acc0 = fadd2(acc0, x0);
    acc1 = fadd2(acc1, x1);
    u[i] = E::pack(x0.x, x0.y);
    u[i + 1] = E::pack(x1.x, x1.y);
  }
}

// ---- part of the exponentials on the FMA + ALU pipes -------------------------------------------------------------------
// MUFU.EX2 (16/clk/SM) is the binding unit of these kernels. exp2 of a pair by Cody-Waite range reduction (the 1.5*2^23
// round-to-nearest trick) and a degree-3 minimax polynomial for 2^f on [-0.5, 0.5]: max relative error 7.5e-5, far below
// the bf16 / fp16 rounding of P (tools/micro/exp_emul.cu). It costs ~11.5 issue cycles per element against 8 MUFU cycles, so
// it only pays for a minority of the columns and when it runs OUTSIDE the MUFU turn, on the otherwise idle FMA pipe.
__device__ __forceinline__ float2 exp2_fma(float2 t) {
  const float2 magic = make_float2(12582912.f, 12582912.f), nmagic = make_float2(-12582912.f, -12582912.f), mone = make_float2(-1.f, -1.f);
  const float2 c3 = make_float2(0.0551716685f, 0.0551716685f), c2 = make_float2(0.2426111251f, 0.2426111251f),
               c1 = make_float2(0.6932609677f, 0.6932609677f), c0 = make_float2(0.9999280572f, 0.9999280572f);
  // masked (-inf) and far-away keys end up as 2^-126 ~ 0 instead of wrapping the exponent field; the upper clamp matters to the
  // unshifted loop only (a score beyond 2^127 must blow the row sum up, not wrap, so that the exact second pass is taken)
  t.x = fminf(fmaxf(t.x, -126.f), 127.f);
  t.y = fminf(fmaxf(t.y, -126.f), 127.f);
  const float2 r = fadd2(t, magic);
  const float2 fi = fadd2(r, nmagic);
  const float2 f = ffma2(fi, mone, t);
  float2 p = ffma2(c3, f, c2);
  p = ffma2(p, f, c1);
  p = ffma2(p, f, c0);
  float2 e;
  e.x = __int_as_float((__float_as_int(r.x) << 23) + __float_as_int(p.x));
  e.y = __int_as_float((__float_as_int(r.y) << 23) + __float_as_int(p.y));
  return e;
}

// EVERY = 0: plain. EVERY = n: every n-th column pair of a chunk is exponentiated here (before the turn), the others only scaled.
template <int EVERY> __device__ __forceinline__ constexpr bool emul_pair(int pair) { return EVERY > 0 && (pair % (EVERY > 0 ? EVERY : 1)) == EVERY - 1; }

#ifndef IEF_SCALAR_SCALE
#define IEF_SCALAR_SCALE 0
#endif
#ifndef IEF_TC3_EMUL_IN_TURN
#define IEF_TC3_EMUL_IN_TURN 1   // 1: the emulated pairs are interleaved with the MUFU ones inside the exp section; 0: before the turn
#endif
template <int EVERY_>
__device__ __forceinline__ void scale_chunk_mix(uint32_t (&s)[32], float2 c2, float2 nmc) {
  constexpr int EVERY = IEF_TC3_EMUL_IN_TURN ? 0 : EVERY_;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
#if IEF_SCALAR_SCALE
    float2 x = make_float2(fmaf(__uint_as_float(s[i]), c2.x, nmc.x), fmaf(__uint_as_float(s[i + 1]), c2.x, nmc.x));
#else
    float2 x = ffma2(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), c2, nmc);
#endif
    if (emul_pair<EVERY>(i / 2)) x = exp2_fma(x);
    s[i] = __float_as_uint(x.x);
    s[i + 1] = __float_as_uint(x.y);
  }
}

// PERFORMANCE DIAGNOSTIC ONLY (wrong results): -DIEF_TC3_DIAG_SKIP_EXP=n drops the exponential of every n-th column pair (n = 1: all of
// them) to measure how much of the kernel's time the MUFU pipe really accounts for.
#ifndef IEF_TC3_DIAG_SKIP_EXP
#define IEF_TC3_DIAG_SKIP_EXP 0
#endif
__device__ __forceinline__ constexpr bool diag_skip_exp(int pair) { return IEF_TC3_DIAG_SKIP_EXP > 0 && pair % (IEF_TC3_DIAG_SKIP_EXP > 0 ? IEF_TC3_DIAG_SKIP_EXP : 1) == 0; }

// Row sums in registers (head_dim > 48: no accumulator columns left for the row-sum MMA). bf16: the sum is taken over the SAME 16-bit
// values the PV MMA multiplies — the fp32 exponentials cut to their upper halves (2 LOP3 + 1 PRMT instead of 1 F2FP per pair) — so that
// O = sum(p v) / sum(p) is a ratio of consistently rounded terms. Summing the unrounded fp32 p against rounded P in the numerator costs
// up to 2^-9 |v| on peaked rows whenever the dominant p is not exactly 1 (stale lazy maximum, or the unshifted loop, where it never is):
// 0.031 observed at |v| = 5 (tools/diag/overflow_case.py). fp16 keeps round-to-nearest packing (11-bit P: a quarter of that error).
#ifndef IEF_TC3_CONSISTENT_SUM
#define IEF_TC3_CONSISTENT_SUM 1
#endif
template <typename E, int EVERY>
__device__ __forceinline__ void exp_pack_chunk_mix(const uint32_t (&s)[32], uint32_t (&u)[16], float2& acc0, float2& acc1) {
  constexpr bool cut = IEF_TC3_CONSISTENT_SUM && sizeof(typename E::T) == 2 && E::kIsBf16;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float2 x = make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]));
    if (diag_skip_exp(i)) { x.x = fmaf(x.x, 1e-3f, 0.5f); x.y = fmaf(x.y, 1e-3f, 0.5f); }
    else if (!emul_pair<EVERY>(i)) { x.x = ief_exp2(x.x); x.y = ief_exp2(x.y); }
    else if (IEF_TC3_EMUL_IN_TURN) x = exp2_fma(x);
    if constexpr (cut) {
      const uint32_t lo = __float_as_uint(x.x) & 0xffff0000u, hi = __float_as_uint(x.y) & 0xffff0000u;
      x = make_float2(__uint_as_float(lo), __uint_as_float(hi));
      u[i] = __byte_perm(lo, hi, 0x7632);
    } else {
      u[i] = E::pack(x.x, x.y);
    }
    if (i & 1) acc1 = fadd2(acc1, x); else acc0 = fadd2(acc0, x);
  }
}

// the same without the row sum (it is computed by the tensor pipe: P x ones)
template <typename E, int EVERY, int BEGIN = 0, int END = 16>
__device__ __forceinline__ void exp_pack_chunk_nosum(const uint32_t (&s)[32], uint32_t (&u)[16]) {
#pragma unroll
  for (int i = BEGIN; i < END; ++i) {
    float2 x = make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]));
    if (diag_skip_exp(i)) { x.x = fmaf(x.x, 1e-3f, 0.5f); x.y = fmaf(x.y, 1e-3f, 0.5f); }
    else if (!emul_pair<EVERY>(i)) { x.x = ief_exp2(x.x); x.y = ief_exp2(x.y); }
    else if (IEF_TC3_EMUL_IN_TURN) x = exp2_fma(x);
    u[i] = E::pack(x.x, x.y);
  }
}
