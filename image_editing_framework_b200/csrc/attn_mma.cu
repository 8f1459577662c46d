// Warp-level (mma.sync m16n8k16) controlled attention. Serves every shape ief_attn_fwd accepts and is the
// only kernel that can also emit the normalised probabilities (two sweeps: statistics, then P and PV):
//   AttentionStore self maps, N <= 32^2 (p2p/model/attention_base.py:64-68),
//   masactrl AttentionStore (masactrl/model/attention_base.py:59-66),
//   pix2pix-zero attn_probs (pix2pix-zero/model/attention_control.py:46).
// Same contract as attn_tc.cu: O[b] = softmax(scale Q[q_src[b]] K[k_src[b]]^T) V[v_src[b]], optional second
// key/value block (MasaCtrl Union, masactrl/model/attention_control.py:100-101).
#include "mma_utils.cuh"
#include <math.h>

using namespace mmau;

namespace {

constexpr int kBM = 64, kBN = 64, kThreads = 128;
constexpr float kLog2e = 1.4426950408889634f;

struct MmaArgs {
  ief_tensor4 q, k, v, o;
  int32_t B, H, Nq, Nk, d;
  int32_t nt1, nt2;
  float scale_log2;
  float* probs;         // [B,H,Nq,Nk_total] or null
  int32_t probs_accum;
  const float* key_bias;  // [n_bias, Nk] additive bias on the scaled scores, or null
  IefRowTable rows;
};

template <int DTYPE, int DP, bool WRITE_P>
__global__ void __launch_bounds__(kThreads)
attn_mma_kernel(const __grid_constant__ MmaArgs a) {
  using E = ElemT<DTYPE>;
  using T = typename E::T;
  constexpr int LD = DP + 8;
  constexpr int KS = DP / 16;  // k-steps over head_dim
  constexpr int NB = DP / 8;   // 8-wide output column blocks
  const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
  if (!a.rows.active[b]) return;
  extern __shared__ uint4 smem4[];
  T* sQ = reinterpret_cast<T*>(smem4);
  T* sKV = sQ + kBM * LD;  // two stages of [K tile | V tile]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nk_total = a.Nk * (a.nt2 > 0 ? 2 : 1);
  const int nt = a.nt1 + a.nt2;
  constexpr int NSWEEP = WRITE_P ? 2 : 1;
  const int n_iter = NSWEEP * nt;

  // key/value tile `it` of the flattened (sweep, tile) sequence -> stage it & 1, by cp.async (the statistics sweep needs no V)
  auto fetch = [&](int it) {
    const int j = it >= nt ? it - nt : it;
    const bool blk2 = j >= a.nt1;
    const int jj = blk2 ? j - a.nt1 : j;
    const int kb = blk2 ? a.rows.k2[b] : a.rows.k[b];
    const int vb = blk2 ? a.rows.v2[b] : a.rows.v[b];
    T* sK = sKV + (it & 1) * 2 * kBN * LD;
    const T* kg = reinterpret_cast<const T*>(a.k.ptr) + (int64_t)kb * a.k.stride_b + (int64_t)h * a.k.stride_h;
    load_tile_async<T, kBN, DP, LD, kThreads>(sK, kg, a.k.stride_n, jj * kBN, a.Nk, a.d, tid);
    if (!(WRITE_P && it < nt)) {
      const T* vg = reinterpret_cast<const T*>(a.v.ptr) + (int64_t)vb * a.v.stride_b + (int64_t)h * a.v.stride_h;
      load_tile_async<T, kBN, DP, LD, kThreads>(sK + kBN * LD, vg, a.v.stride_n, jj * kBN, a.Nk, a.d, tid);
    }
    cp_async_commit();
  };

  const T* qg = reinterpret_cast<const T*>(a.q.ptr) + (int64_t)a.rows.q[b] * a.q.stride_b + (int64_t)h * a.q.stride_h;
  load_tile_async<T, kBM, DP, LD, kThreads>(sQ, qg, a.q.stride_n, qt * kBM, a.Nq, a.d, tid);
  fetch(0);
  uint32_t qf[KS][4];

  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f}, inv_l[2] = {1.f, 1.f};
  float o[NB][4];
#pragma unroll
  for (int i = 0; i < NB; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  // rows with a key bias work on t = s * scale_log2 + bias * log2(e) from the start (so the running max sees the bias)
  const float* bias = (a.key_bias != nullptr && a.rows.bias[b] >= 0) ? a.key_bias + (int64_t)a.rows.bias[b] * a.Nk : nullptr;
  const float c2 = bias ? 1.f : a.scale_log2;
  const int grow0 = qt * kBM + warp * 16 + g;  // rows grow0 and grow0 + 8

#pragma unroll 1
  for (int it = 0; it < n_iter; ++it) {
    const int sweep = it >= nt ? 1 : 0;
    const int j = it - sweep * nt;
    {
      const bool stats_only = WRITE_P && sweep == 0;
      const bool emit = WRITE_P && sweep == 1;
      const bool blk2 = j >= a.nt1;
      const int jj = blk2 ? j - a.nt1 : j;
      const T* sK = sKV + (it & 1) * 2 * kBN * LD;
      const T* sV = sK + kBN * LD;
      cp_async_wait<0>();
      __syncthreads();                    // tile `it` has landed; every warp is done with tile it-1
      if (it + 1 < n_iter) fetch(it + 1);  // overlaps with the math below
      if (it == 0) {
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) ldsm_x4(qf[kk], &sQ[(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + kk * 16 + (lane >> 4) * 8]);
      }
      // S = Q K^T  (16 rows x 64 keys per warp)
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
        for (int nb2 = 0; nb2 < 4; ++nb2) {
          uint32_t bf[4];
          ldsm_x4(bf, &sK[(nb2 * 16 + (lane & 7) + (lane >> 4) * 8) * LD + kk * 16 + ((lane >> 3) & 1) * 8]);
          mma16816<DTYPE>(s[2 * nb2], qf[kk], bf[0], bf[1]);
          mma16816<DTYPE>(s[2 * nb2 + 1], qf[kk], bf[2], bf[3]);
        }
      }
      const int vc = min(kBN, a.Nk - jj * kBN);
      if (bias) {
        // "masked" keys carry finfo.min in the reference; clamped so the product with log2(e) stays finite and a row whose
        // keys are all masked still comes out as the uniform average, as it does there
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          const int c = nb * 8 + 2 * t;
          const float b0 = c < vc ? fmaxf(__ldg(bias + jj * kBN + c) * kLog2e, -3.0e38f) : 0.f;
          const float b1 = c + 1 < vc ? fmaxf(__ldg(bias + jj * kBN + c + 1) * kLog2e, -3.0e38f) : 0.f;
          s[nb][0] = fmaf(s[nb][0], a.scale_log2, b0);
          s[nb][1] = fmaf(s[nb][1], a.scale_log2, b1);
          s[nb][2] = fmaf(s[nb][2], a.scale_log2, b0);
          s[nb][3] = fmaf(s[nb][3], a.scale_log2, b1);
        }
      }
      if (vc < kBN) {
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          const int c = nb * 8 + 2 * t;
          if (c >= vc) s[nb][0] = s[nb][2] = -INFINITY;
          if (c + 1 >= vc) s[nb][1] = s[nb][3] = -INFINITY;
        }
      }
      if (!emit) {
        float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          tm0 = fmaxf(tm0, fmaxf(s[nb][0], s[nb][1]));
          tm1 = fmaxf(tm1, fmaxf(s[nb][2], s[nb][3]));
        }
        tm0 = quad_max(tm0);
        tm1 = quad_max(tm1);
        const float mn0 = fmaxf(m[0], tm0), mn1 = fmaxf(m[1], tm1);
        const float al0 = ief_exp2((m[0] - mn0) * c2), al1 = ief_exp2((m[1] - mn1) * c2);
        m[0] = mn0;
        m[1] = mn1;
        l[0] *= al0;
        l[1] *= al1;
        if (!stats_only) {
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            o[i][0] *= al0; o[i][1] *= al0; o[i][2] *= al1; o[i][3] *= al1;
          }
        }
      }
      const float mc0 = m[0] * c2, mc1 = m[1] * c2;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        s[nb][0] = ief_exp2(fmaf(s[nb][0], c2, -mc0)) * inv_l[0];
        s[nb][1] = ief_exp2(fmaf(s[nb][1], c2, -mc0)) * inv_l[0];
        s[nb][2] = ief_exp2(fmaf(s[nb][2], c2, -mc1)) * inv_l[1];
        s[nb][3] = ief_exp2(fmaf(s[nb][3], c2, -mc1)) * inv_l[1];
        ps0 += s[nb][0] + s[nb][1];
        ps1 += s[nb][2] + s[nb][3];
      }
      if (!emit) { l[0] += ps0; l[1] += ps1; }
      if (emit && a.rows.pslot[b] >= 0) {
        // normalised probabilities -> global fp32 (each quad writes one full 32-byte sector per row)
        const int64_t prow = ((int64_t)a.rows.pslot[b] * a.H + h) * a.Nq;
        const int col0 = (blk2 ? a.Nk : 0) + jj * kBN;
        if (vc == kBN && ((nk_total | col0) & 1) == 0) {
          // full tile, 8-byte aligned pairs; accumulating: the old pairs of four column blocks are requested together before the first
          // of their stores (load / add / store per element strings one DRAM round trip per element together, see
          // attn_probs_from_lse_kernel)
          float2* d0 = reinterpret_cast<float2*>(a.probs + (prow + grow0) * nk_total + col0 + 2 * t);
          float2* d1 = reinterpret_cast<float2*>(a.probs + (prow + grow0 + 8) * nk_total + col0 + 2 * t);
          const bool r0 = grow0 < a.Nq, r1 = grow0 + 8 < a.Nq;
#pragma unroll
          for (int g4 = 0; g4 < 8; g4 += 4) {
            float2 o0[4], o1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              o0[i] = (a.probs_accum && r0) ? d0[(g4 + i) * 4] : make_float2(0.f, 0.f);
              o1[i] = (a.probs_accum && r1) ? d1[(g4 + i) * 4] : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (r0) d0[(g4 + i) * 4] = make_float2(o0[i].x + s[g4 + i][0], o0[i].y + s[g4 + i][1]);
              if (r1) d1[(g4 + i) * 4] = make_float2(o1[i].x + s[g4 + i][2], o1[i].y + s[g4 + i][3]);
            }
          }
        } else {
#pragma unroll
          for (int nb = 0; nb < 8; ++nb) {
            const int c = nb * 8 + 2 * t;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int r = grow0 + hh * 8;
              if (r < a.Nq) {
                float* dst = a.probs + (prow + r) * nk_total + col0 + c;
                const float p0 = s[nb][2 * hh], p1 = s[nb][2 * hh + 1];
                if (c < vc) dst[0] = a.probs_accum ? dst[0] + p0 : p0;
                if (c + 1 < vc) dst[1] = a.probs_accum ? dst[1] + p1 : p1;
              }
            }
          }
        }
      }
      if (!stats_only) {
        // O += P V
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t pa[4];
          pa[0] = E::pack(s[2 * kk][0], s[2 * kk][1]);
          pa[1] = E::pack(s[2 * kk][2], s[2 * kk][3]);
          pa[2] = E::pack(s[2 * kk + 1][0], s[2 * kk + 1][1]);
          pa[3] = E::pack(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
          for (int nb2 = 0; nb2 < KS; ++nb2) {
            uint32_t vf[4];
            ldsm_x4_t(vf, &sV[(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + nb2 * 16 + (lane >> 4) * 8]);
            mma16816<DTYPE>(o[2 * nb2], pa, vf[0], vf[1]);
            mma16816<DTYPE>(o[2 * nb2 + 1], pa, vf[2], vf[3]);
          }
        }
      }
    }
    if (it == nt - 1) {  // end of the first (or only) sweep
      l[0] = quad_sum(l[0]);
      l[1] = quad_sum(l[1]);
      if (WRITE_P) { inv_l[0] = 1.f / l[0]; inv_l[1] = 1.f / l[1]; }
    }
  }
  const float f0 = WRITE_P ? 1.f : 1.f / l[0], f1 = WRITE_P ? 1.f : 1.f / l[1];
  T* og = reinterpret_cast<T*>(a.o.ptr) + (int64_t)b * a.o.stride_b + (int64_t)h * a.o.stride_h;
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    const int c = nb * 8 + 2 * t;
    if (c < a.d) {
      if (grow0 < a.Nq) *reinterpret_cast<uint32_t*>(og + (int64_t)grow0 * a.o.stride_n + c) = E::pack(o[nb][0] * f0, o[nb][1] * f0);
      if (grow0 + 8 < a.Nq) *reinterpret_cast<uint32_t*>(og + (int64_t)(grow0 + 8) * a.o.stride_n + c) = E::pack(o[nb][2] * f1, o[nb][3] * f1);
    }
  }
}

template <int DTYPE, int DP, bool WRITE_P>
int launch_one(const MmaArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int smem = (kBM + 4 * kBN) * (DP + 8) * 2;  // Q + two stages of K, V
  auto kern = attn_mma_kernel<DTYPE, DP, WRITE_P>;
  IEF_CONFIG_SMEM(kern, smem);
  kern<<<grid, kThreads, smem, st>>>(a);
  IEF_LAUNCH_OK("attn_mma_kernel");
  return IEF_OK;
}

template <int DTYPE, bool WRITE_P>
int launch_dp(const MmaArgs& a, dim3 grid, cudaStream_t st) {
  const int d = a.d;
  if (d <= 32) return launch_one<DTYPE, 32, WRITE_P>(a, grid, st);
  if (d <= 48) return launch_one<DTYPE, 48, WRITE_P>(a, grid, st);
  if (d <= 64) return launch_one<DTYPE, 64, WRITE_P>(a, grid, st);
  if (d <= 80) return launch_one<DTYPE, 80, WRITE_P>(a, grid, st);
  if (d <= 96) return launch_one<DTYPE, 96, WRITE_P>(a, grid, st);
  if (d <= 128) return launch_one<DTYPE, 128, WRITE_P>(a, grid, st);
  return launch_one<DTYPE, 160, WRITE_P>(a, grid, st);
}

// Probabilities only, from a known row log-sum-exp (log2 units): P = exp2(scale_log2 * Q K^T - lse). One sweep over the keys, no V,
// no running statistics: used behind the tcgen05 kernels, which produce O and the lse, so that stored maps cost one QK^T on
// mma.sync plus the HBM traffic of the maps instead of the two-sweep kernel above.
template <int DTYPE, int DP, bool ACCUM>
__global__ void __launch_bounds__(kThreads)
attn_probs_from_lse_kernel(const __grid_constant__ MmaArgs a, const float* __restrict__ lse, int nqt, int ksplit) {
  using E = ElemT<DTYPE>;
  using T = typename E::T;
  constexpr int LD = DP + 8, KS = DP / 16;
  // grid.x = query tiles x key splits: there is no reduction over the keys, so the key range is spread over several CTAs to
  // keep enough of them in flight (the sweep is bound by the latency of the map read-modify-write, not by arithmetic)
  const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x % nqt, split = blockIdx.x / nqt;
  if (!a.rows.active[b] || a.rows.pslot[b] < 0) return;
  extern __shared__ uint4 smem4[];
  T* sQ = reinterpret_cast<T*>(smem4);
  T* sKs = sQ + kBM * LD;  // two stages of K
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nk_total = a.Nk * (a.nt2 > 0 ? 2 : 1);
  const int nt_all = a.nt1 + a.nt2;
  const int j_begin = (int)((long)nt_all * split / ksplit), j_end = (int)((long)nt_all * (split + 1) / ksplit);
  auto fetch = [&](int j) {
    const bool blk2 = j >= a.nt1;
    const int jj = blk2 ? j - a.nt1 : j;
    const int kb = blk2 ? a.rows.k2[b] : a.rows.k[b];
    const T* kg = reinterpret_cast<const T*>(a.k.ptr) + (int64_t)kb * a.k.stride_b + (int64_t)h * a.k.stride_h;
    load_tile_async<T, kBN, DP, LD, kThreads>(sKs + (j & 1) * kBN * LD, kg, a.k.stride_n, jj * kBN, a.Nk, a.d, tid);
    cp_async_commit();
  };
  const T* qg = reinterpret_cast<const T*>(a.q.ptr) + (int64_t)a.rows.q[b] * a.q.stride_b + (int64_t)h * a.q.stride_h;
  load_tile_async<T, kBM, DP, LD, kThreads>(sQ, qg, a.q.stride_n, qt * kBM, a.Nq, a.d, tid);
  if (j_begin >= j_end) return;
  fetch(j_begin);
  const int grow0 = qt * kBM + warp * 16 + g;
  const int64_t lrow = ((int64_t)b * a.H + h) * a.Nq;
  const float lse0 = grow0 < a.Nq ? __ldg(lse + lrow + grow0) : 0.f, lse1 = grow0 + 8 < a.Nq ? __ldg(lse + lrow + grow0 + 8) : 0.f;
  const float c2 = a.scale_log2;
  const int64_t prow = ((int64_t)a.rows.pslot[b] * a.H + h) * a.Nq;
  uint32_t qf[KS][4];
#pragma unroll 1
  for (int j = j_begin; j < j_end; ++j) {
    const bool blk2 = j >= a.nt1;
    const int jj = blk2 ? j - a.nt1 : j;
    const T* sK = sKs + (j & 1) * kBN * LD;
    cp_async_wait<0>();
    __syncthreads();
    if (j + 1 < j_end) fetch(j + 1);
    if (j == j_begin) {
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) ldsm_x4(qf[kk], &sQ[(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + kk * 16 + (lane >> 4) * 8]);
    }
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
      for (int nb2 = 0; nb2 < 4; ++nb2) {
        uint32_t bf[4];
        ldsm_x4(bf, &sK[(nb2 * 16 + (lane & 7) + (lane >> 4) * 8) * LD + kk * 16 + ((lane >> 3) & 1) * 8]);
        mma16816<DTYPE>(s[2 * nb2], qf[kk], bf[0], bf[1]);
        mma16816<DTYPE>(s[2 * nb2 + 1], qf[kk], bf[2], bf[3]);
      }
    }
    const int vc = min(kBN, a.Nk - jj * kBN);
    const int col0 = (blk2 ? a.Nk : 0) + jj * kBN;
    if (vc == kBN && ((nk_total | col0) & 1) == 0) {
      // Full tile, 8-byte aligned pairs. Accumulating: ALL sixteen old pairs of this thread are requested before the first store —
      // written as load / add / store per pair, every load had to wait behind the previous store to the same array (they may alias
      // as far as the compiler knows), i.e. sixteen serial DRAM round trips per tile: that chain, not bandwidth, bounded the sweep.
      float2* d0 = reinterpret_cast<float2*>(a.probs + (prow + grow0) * nk_total + col0 + 2 * t);
      float2* d1 = reinterpret_cast<float2*>(a.probs + (prow + grow0 + 8) * nk_total + col0 + 2 * t);
      const bool r0 = grow0 < a.Nq, r1 = grow0 + 8 < a.Nq;
      float2 o0[8], o1[8];
      if constexpr (ACCUM) {
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          o0[nb] = r0 ? d0[nb * 4] : make_float2(0.f, 0.f);
          o1[nb] = r1 ? d1[nb * 4] : make_float2(0.f, 0.f);
        }
      }
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        float2 p0 = make_float2(ief_exp2(fmaf(s[nb][0], c2, -lse0)), ief_exp2(fmaf(s[nb][1], c2, -lse0)));
        float2 p1 = make_float2(ief_exp2(fmaf(s[nb][2], c2, -lse1)), ief_exp2(fmaf(s[nb][3], c2, -lse1)));
        if constexpr (ACCUM) {
          p0.x += o0[nb].x; p0.y += o0[nb].y;
          p1.x += o1[nb].x; p1.y += o1[nb].y;
        }
        if (r0) d0[nb * 4] = p0;
        if (r1) d1[nb * 4] = p1;
      }
      continue;
    }
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {   // ragged last tile / odd key counts: element-wise
      const int c = nb * 8 + 2 * t;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int r = grow0 + hh * 8;
        if (r < a.Nq) {
          const float ls = hh ? lse1 : lse0;
          float* dst = a.probs + (prow + r) * nk_total + col0 + c;
          const float p0 = ief_exp2(fmaf(s[nb][2 * hh], c2, -ls)), p1 = ief_exp2(fmaf(s[nb][2 * hh + 1], c2, -ls));
          if (c < vc) dst[0] = ACCUM ? dst[0] + p0 : p0;
          if (c + 1 < vc) dst[1] = ACCUM ? dst[1] + p1 : p1;
        }
      }
    }
  }
}

template <int DTYPE, int DP, bool ACCUM>
int launch_probs_one(const MmaArgs& a, const float* lse, dim3 grid, cudaStream_t st) {
  constexpr int smem = (kBM + 2 * kBN) * (DP + 8) * 2;
  auto kern = attn_probs_from_lse_kernel<DTYPE, DP, ACCUM>;
  IEF_CONFIG_SMEM(kern, smem);
  // enough CTAs for ~6 per SM: split the key range when the (row, head, query tile) grid alone is too small
  int stored_rows = 0;
  for (int i = 0; i < a.B; ++i) stored_rows += (a.rows.active[i] && a.rows.pslot[i] >= 0) ? 1 : 0;
  const int nqt = grid.x, nt_all = a.nt1 + a.nt2;
  const long base_ctas = (long)nqt * a.H * (stored_rows > 0 ? stored_rows : 1);
  int ksplit = (int)((148L * 6 + base_ctas - 1) / base_ctas);
  ksplit = ksplit < 1 ? 1 : (ksplit > nt_all ? nt_all : ksplit);
  grid.x = nqt * ksplit;
  kern<<<grid, kThreads, smem, st>>>(a, lse, nqt, ksplit);
  IEF_LAUNCH_OK("attn_probs_from_lse_kernel");
  return IEF_OK;
}

template <int DTYPE, bool ACCUM>
int launch_probs_dp2(const MmaArgs& a, const float* lse, dim3 grid, cudaStream_t st) {
  const int d = a.d;
  if (d <= 32) return launch_probs_one<DTYPE, 32, ACCUM>(a, lse, grid, st);
  if (d <= 48) return launch_probs_one<DTYPE, 48, ACCUM>(a, lse, grid, st);
  if (d <= 64) return launch_probs_one<DTYPE, 64, ACCUM>(a, lse, grid, st);
  if (d <= 80) return launch_probs_one<DTYPE, 80, ACCUM>(a, lse, grid, st);
  if (d <= 96) return launch_probs_one<DTYPE, 96, ACCUM>(a, lse, grid, st);
  return launch_probs_one<DTYPE, 128, ACCUM>(a, lse, grid, st);
}

template <int DTYPE>
int launch_probs_dp(const MmaArgs& a, const float* lse, dim3 grid, cudaStream_t st) {
  return a.probs_accum ? launch_probs_dp2<DTYPE, true>(a, lse, grid, st) : launch_probs_dp2<DTYPE, false>(a, lse, grid, st);
}

}  // namespace

int ief_attn_mma_launch(const ief_attn_params* p, const IefRowTable& rows, cudaStream_t st) {
  IEF_REQUIRE(p->dtype == IEF_BF16 || p->dtype == IEF_F16, IEF_ERR_UNSUPPORTED, "ief_attn_fwd: dtype must be bf16 or f16");
  IEF_REQUIRE(p->d % 8 == 0 && p->d >= 8 && p->d <= 160, IEF_ERR_UNSUPPORTED, "ief_attn_fwd(mma): head_dim %d not a multiple of 8 in [8,160]", p->d);
  const ief_tensor4* ts[4] = {&p->q, &p->k, &p->v, &p->o};
  for (auto tt : ts) {
    IEF_REQUIRE((reinterpret_cast<uintptr_t>(tt->ptr) & 15) == 0, IEF_ERR_INVALID, "ief_attn_fwd: pointer not 16-byte aligned");
    IEF_REQUIRE(tt->stride_n % 8 == 0 && tt->stride_h % 8 == 0 && tt->stride_b % 8 == 0, IEF_ERR_UNSUPPORTED,
                "ief_attn_fwd: strides must be multiples of 8 elements");
  }
  MmaArgs a;
  a.q = p->q; a.k = p->k; a.v = p->v; a.o = p->o;
  a.B = p->B; a.H = p->H; a.Nq = p->Nq; a.Nk = p->Nk; a.d = p->d;
  a.nt1 = ief_ceil_div(p->Nk, kBN);
  bool any2 = false;
  for (int i = 0; i < p->B; ++i) any2 |= rows.k2[i] >= 0;
  if (any2)
    for (int i = 0; i < p->B; ++i)
      IEF_REQUIRE(rows.k2[i] >= 0 && rows.v2[i] >= 0, IEF_ERR_UNSUPPORTED, "k_src2/v_src2 must be set for every row or none");
  a.nt2 = any2 ? a.nt1 : 0;
  a.scale_log2 = p->scale * kLog2e;
  a.probs = p->probs_out;
  a.probs_accum = p->probs_accum;
  a.key_bias = p->key_bias;
  a.rows = rows;
  dim3 grid(ief_ceil_div(p->Nq, kBM), p->H, p->B);
  if (p->probs_out) {
    return p->dtype == IEF_BF16 ? launch_dp<IEF_BF16, true>(a, grid, st) : launch_dp<IEF_F16, true>(a, grid, st);
  }
  return p->dtype == IEF_BF16 ? launch_dp<IEF_BF16, false>(a, grid, st) : launch_dp<IEF_F16, false>(a, grid, st);
}

// Stored maps behind a tcgen05 launch that produced O and the row log-sum-exp (log2 units, [B, H, Nq]): one QK^T sweep.
int ief_attn_probs_from_lse_launch(const ief_attn_params* p, const IefRowTable& rows, const float* lse, cudaStream_t st) {
  IEF_REQUIRE(p->d % 8 == 0 && p->d >= 8 && p->d <= 128, IEF_ERR_UNSUPPORTED, "ief_attn_fwd(probs from lse): head_dim %d", p->d);
  MmaArgs a;
  a.q = p->q; a.k = p->k; a.v = p->v; a.o = p->o;
  a.B = p->B; a.H = p->H; a.Nq = p->Nq; a.Nk = p->Nk; a.d = p->d;
  a.nt1 = ief_ceil_div(p->Nk, kBN);
  bool any2 = false;
  for (int i = 0; i < p->B; ++i) any2 |= rows.k2[i] >= 0;
  a.nt2 = any2 ? a.nt1 : 0;
  a.scale_log2 = p->scale * kLog2e;
  a.probs = p->probs_out;
  a.probs_accum = p->probs_accum;
  a.key_bias = nullptr;
  a.rows = rows;
  dim3 grid(ief_ceil_div(p->Nq, kBM), p->H, p->B);
  return p->dtype == IEF_BF16 ? launch_probs_dp<IEF_BF16>(a, lse, grid, st) : launch_probs_dp<IEF_F16>(a, lse, grid, st);
}
