// Backward of the 77-key cross-attention for Pix2Pix-zero's guidance pass (pix2pix-zero/model/sd_utils.py:163-174: the loss is a sum
// over the cross-attention layers of ||P - P_ref||^2, differentiated with respect to the latents).
//   forward     P = softmax(scale * Q K^T),  O = P V                               (attention_control.py:43-49)
//   given       dO [B, N, H*d]  and  dP_ext [B*H, N, Nk]  (the loss' direct gradient on the probabilities, optional)
//   computes    dP = dO V^T + dP_ext;   dS = P * (dP - rowsum(P * dP)) * scale;   dQ = dS K
//   and, on request, stores dS (fp32, [B*H, N, Nk]) so that the caller can form the two small reductions over the query axis,
//   dK = dS^T Q and dV = P^T dO (77 x d each), with two batched GEMMs.
// Same tiling as cross_attn.cu: a CTA is 64 query rows of one (row, head), 4 warps x 16 rows, P recomputed from Q and K and kept in
// mma.sync accumulator fragments; dO V^T is the same MMA shape as Q K^T, dS K the same as P V. All tiles arrive in one cp.async round.
#include "mma_utils.cuh"
#include <math.h>

using namespace mmau;

namespace {

constexpr int kBM = 64, kNKP = 80, kThreads = 128;
constexpr float kLog2e = 1.4426950408889634f;

struct BwdArgs {
  ief_tensor4 q, k, v, dout, dq;
  int32_t B, H, Nq, Nk, d;
  float scale, scale_log2;
  const float* dprobs;  // [B*H, Nq, Nk] or null
  float* ds_out;        // [B*H, Nq, Nk] or null
};

// acc[nb][0..1] = row g, columns nb*8 + 2t, +1;  acc[nb][2..3] = row g+8:  A[16 x DP] (this warp's rows) times B[80 x DP]^T
template <int DTYPE, int DP>
__device__ __forceinline__ void rows_times_keys(const typename ElemT<DTYPE>::T* sA, const typename ElemT<DTYPE>::T* sB, float (&acc)[kNKP / 8][4],
                                                int warp, int lane) {
  constexpr int LD = DP + 8, KS = DP / 16;
#pragma unroll
  for (int i = 0; i < kNKP / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    uint32_t af[4];
    ldsm_x4(af, &sA[(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + kk * 16 + (lane >> 4) * 8]);
#pragma unroll
    for (int nb2 = 0; nb2 < kNKP / 16; ++nb2) {
      uint32_t bf[4];
      ldsm_x4(bf, &sB[(nb2 * 16 + (lane & 7) + (lane >> 4) * 8) * LD + kk * 16 + ((lane >> 3) & 1) * 8]);
      mma16816<DTYPE>(acc[2 * nb2], af, bf[0], bf[1]);
      mma16816<DTYPE>(acc[2 * nb2 + 1], af, bf[2], bf[3]);
    }
  }
}

template <int DTYPE, int DP>
__global__ void __launch_bounds__(kThreads)
cross_attn_bwd_kernel(const __grid_constant__ BwdArgs a) {
  using E = ElemT<DTYPE>;
  using T = typename E::T;
  constexpr int LD = DP + 8, KS = DP / 16, NB = DP / 8;
  const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
  extern __shared__ uint4 smem4[];
  T* sQ = reinterpret_cast<T*>(smem4);
  T* sG = sQ + kBM * LD;   // dO tile
  T* sK = sG + kBM * LD;
  T* sV = sK + kNKP * LD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int Nk = a.Nk;
  auto at = [&](const ief_tensor4& x) { return reinterpret_cast<const T*>(x.ptr) + (int64_t)b * x.stride_b + (int64_t)h * x.stride_h; };
  load_tile_async<T, kBM, DP, LD, kThreads>(sQ, at(a.q), a.q.stride_n, qt * kBM, a.Nq, a.d, tid);
  load_tile_async<T, kBM, DP, LD, kThreads>(sG, at(a.dout), a.dout.stride_n, qt * kBM, a.Nq, a.d, tid);
  load_tile_async<T, kNKP, DP, LD, kThreads>(sK, at(a.k), a.k.stride_n, 0, Nk, a.d, tid);
  load_tile_async<T, kNKP, DP, LD, kThreads>(sV, at(a.v), a.v.stride_n, 0, Nk, a.d, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  // P = softmax(scale Q K^T), normalised, in fragments
  float p[kNKP / 8][4];
  rows_times_keys<DTYPE, DP>(sQ, sK, p, warp, lane);
  {
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nb = 0; nb < kNKP / 8; ++nb) {
      const int c = nb * 8 + 2 * t;
      if (c >= Nk) p[nb][0] = p[nb][2] = -INFINITY;
      if (c + 1 >= Nk) p[nb][1] = p[nb][3] = -INFINITY;
      m0 = fmaxf(m0, fmaxf(p[nb][0], p[nb][1]));
      m1 = fmaxf(m1, fmaxf(p[nb][2], p[nb][3]));
    }
    m0 = quad_max(m0) * a.scale_log2;
    m1 = quad_max(m1) * a.scale_log2;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nb = 0; nb < kNKP / 8; ++nb) {
      p[nb][0] = ief_exp2(fmaf(p[nb][0], a.scale_log2, -m0));
      p[nb][1] = ief_exp2(fmaf(p[nb][1], a.scale_log2, -m0));
      p[nb][2] = ief_exp2(fmaf(p[nb][2], a.scale_log2, -m1));
      p[nb][3] = ief_exp2(fmaf(p[nb][3], a.scale_log2, -m1));
      l0 += p[nb][0] + p[nb][1];
      l1 += p[nb][2] + p[nb][3];
    }
    const float i0 = 1.f / quad_sum(l0), i1 = 1.f / quad_sum(l1);
#pragma unroll
    for (int nb = 0; nb < kNKP / 8; ++nb) {
      p[nb][0] *= i0; p[nb][1] *= i0; p[nb][2] *= i1; p[nb][3] *= i1;
    }
  }
  // dP = dO V^T (+ the loss' direct gradient on the probabilities)
  float dp[kNKP / 8][4];
  rows_times_keys<DTYPE, DP>(sG, sV, dp, warp, lane);
  const int grow0 = qt * kBM + warp * 16 + g;
  const int64_t prow = ((int64_t)b * a.H + h) * a.Nq + grow0;
  if (a.dprobs != nullptr) {
    const float* e0 = a.dprobs + prow * Nk;
    const float* e1 = e0 + (int64_t)8 * Nk;
#pragma unroll
    for (int nb = 0; nb < kNKP / 8; ++nb) {
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        const int n = nb * 8 + 2 * t + e2;
        if (n < Nk) {
          if (grow0 < a.Nq) dp[nb][e2] += __ldg(e0 + n);
          if (grow0 + 8 < a.Nq) dp[nb][2 + e2] += __ldg(e1 + n);
        }
      }
    }
  }
  // dS = P * (dP - sum_k P dP) * scale      (columns >= Nk have P = 0)
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int nb = 0; nb < kNKP / 8; ++nb) {
    d0 += p[nb][0] * dp[nb][0] + p[nb][1] * dp[nb][1];
    d1 += p[nb][2] * dp[nb][2] + p[nb][3] * dp[nb][3];
  }
  d0 = quad_sum(d0);
  d1 = quad_sum(d1);
#pragma unroll
  for (int nb = 0; nb < kNKP / 8; ++nb) {
    p[nb][0] = p[nb][0] * (dp[nb][0] - d0) * a.scale;
    p[nb][1] = p[nb][1] * (dp[nb][1] - d0) * a.scale;
    p[nb][2] = p[nb][2] * (dp[nb][2] - d1) * a.scale;
    p[nb][3] = p[nb][3] * (dp[nb][3] - d1) * a.scale;
  }
  if (a.ds_out != nullptr) {
    float* o0 = a.ds_out + prow * Nk;
    float* o1 = o0 + (int64_t)8 * Nk;
#pragma unroll
    for (int nb = 0; nb < kNKP / 8; ++nb) {
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        const int n = nb * 8 + 2 * t + e2;
        if (n < Nk) {
          if (grow0 < a.Nq) o0[n] = p[nb][e2];
          if (grow0 + 8 < a.Nq) o1[n] = p[nb][2 + e2];
        }
      }
    }
  }
  // dQ = dS K   (dS straight from the fragments, K read "V-style": rows = keys)
  float o[NB][4];
#pragma unroll
  for (int i = 0; i < NB; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < kNKP / 16; ++kk) {
    uint32_t pa[4];
    pa[0] = E::pack(p[2 * kk][0], p[2 * kk][1]);
    pa[1] = E::pack(p[2 * kk][2], p[2 * kk][3]);
    pa[2] = E::pack(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    pa[3] = E::pack(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int nb2 = 0; nb2 < KS; ++nb2) {
      uint32_t kf[4];
      ldsm_x4_t(kf, &sK[(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + nb2 * 16 + (lane >> 4) * 8]);
      mma16816<DTYPE>(o[2 * nb2], pa, kf[0], kf[1]);
      mma16816<DTYPE>(o[2 * nb2 + 1], pa, kf[2], kf[3]);
    }
  }
  T* og = reinterpret_cast<T*>(a.dq.ptr) + (int64_t)b * a.dq.stride_b + (int64_t)h * a.dq.stride_h;
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    const int c = nb * 8 + 2 * t;
    if (c < a.d) {
      if (grow0 < a.Nq) *reinterpret_cast<uint32_t*>(og + (int64_t)grow0 * a.dq.stride_n + c) = E::pack(o[nb][0], o[nb][1]);
      if (grow0 + 8 < a.Nq) *reinterpret_cast<uint32_t*>(og + (int64_t)(grow0 + 8) * a.dq.stride_n + c) = E::pack(o[nb][2], o[nb][3]);
    }
  }
}

template <int DTYPE, int DP>
int launch_one(const BwdArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int smem = 2 * (kBM + kNKP) * (DP + 8) * 2;
  auto kern = cross_attn_bwd_kernel<DTYPE, DP>;
  IEF_CONFIG_SMEM(kern, smem);
  kern<<<grid, kThreads, smem, st>>>(a);
  IEF_LAUNCH_OK("cross_attn_bwd_kernel");
  return IEF_OK;
}

template <int DTYPE>
int launch_dp(const BwdArgs& a, dim3 grid, cudaStream_t st) {
  const int d = a.d;
  if (d <= 32) return launch_one<DTYPE, 32>(a, grid, st);
  if (d <= 48) return launch_one<DTYPE, 48>(a, grid, st);
  if (d <= 64) return launch_one<DTYPE, 64>(a, grid, st);
  if (d <= 80) return launch_one<DTYPE, 80>(a, grid, st);
  if (d <= 96) return launch_one<DTYPE, 96>(a, grid, st);
  if (d <= 128) return launch_one<DTYPE, 128>(a, grid, st);
  return launch_one<DTYPE, 160>(a, grid, st);
}

}  // namespace

extern "C" int ief_cross_attn_bwd(const ief_cross_bwd_params* p, void* stream) {
  IEF_REQUIRE(p != nullptr, IEF_ERR_INVALID, "ief_cross_attn_bwd: null params");
  IEF_REQUIRE(p->q.ptr && p->k.ptr && p->v.ptr && p->dout.ptr && p->dq.ptr, IEF_ERR_INVALID, "ief_cross_attn_bwd: null tensor pointer");
  IEF_REQUIRE(p->dtype == IEF_BF16 || p->dtype == IEF_F16, IEF_ERR_UNSUPPORTED, "ief_cross_attn_bwd: dtype must be bf16 or f16");
  IEF_REQUIRE(p->B >= 1 && p->B <= 65535 && p->H >= 1 && p->H <= 65535 && p->Nq >= 1, IEF_ERR_INVALID, "ief_cross_attn_bwd: bad B/H/Nq");
  IEF_REQUIRE(p->Nk >= 1 && p->Nk <= kNKP, IEF_ERR_UNSUPPORTED, "ief_cross_attn_bwd: Nk=%d, at most %d keys supported", p->Nk, kNKP);
  IEF_REQUIRE(p->d % 8 == 0 && p->d >= 8 && p->d <= 160, IEF_ERR_UNSUPPORTED, "ief_cross_attn_bwd: head_dim %d unsupported", p->d);
  const ief_tensor4* ts[5] = {&p->q, &p->k, &p->v, &p->dout, &p->dq};
  for (auto tt : ts) {
    IEF_REQUIRE((reinterpret_cast<uintptr_t>(tt->ptr) & 15) == 0, IEF_ERR_INVALID, "ief_cross_attn_bwd: pointer not 16-byte aligned");
    IEF_REQUIRE(tt->stride_n % 8 == 0 && tt->stride_h % 8 == 0 && tt->stride_b % 8 == 0, IEF_ERR_UNSUPPORTED,
                "ief_cross_attn_bwd: strides must be multiples of 8 elements");
  }
  BwdArgs a;
  a.q = p->q; a.k = p->k; a.v = p->v; a.dout = p->dout; a.dq = p->dq;
  a.B = p->B; a.H = p->H; a.Nq = p->Nq; a.Nk = p->Nk; a.d = p->d;
  a.scale = p->scale;
  a.scale_log2 = p->scale * kLog2e;
  a.dprobs = p->dprobs;
  a.ds_out = p->ds_out;
  dim3 grid(ief_ceil_div(p->Nq, kBM), p->H, p->B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return p->dtype == IEF_BF16 ? launch_dp<IEF_BF16>(a, grid, st) : launch_dp<IEF_F16>(a, grid, st);
}
