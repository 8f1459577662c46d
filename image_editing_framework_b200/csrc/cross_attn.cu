// 77-key cross-attention with the Prompt-to-Prompt probability edit fused (ief_cross_attn_edit_fwd).
// Replaces, per cross-attention layer, the chain
//   get_attention_scores (p2p/model/register.py:47) -> AttentionControl.__call__ (attention_base.py:16-28)
//   -> AttentionControlEdit.forward (attention_base.py:113-125) -> replace_cross_attention
//      (attention_control.py:15-16 replace, :28-31 refine, :42-46 reweight) -> bmm (register.py:50)
//   and AttentionStore.forward's capture of the post-edit maps (attention_base.py:64-68)
// by one kernel: the <=80 key probabilities of a 64-query tile live in shared memory (fp32), the edit is
// applied there, the maps are (optionally) streamed to / accumulated into the store with coalesced
// writes, and P'V runs on the tensor cores. HBM-bound (AI ~ 76 flop/B): Q and O are touched once.
#include "mma_utils.cuh"
#include <math.h>

using namespace mmau;

namespace {

constexpr int kBM = 64, kNKP = 80, kPLD = 81, kThreads = 128;
constexpr float kLog2e = 1.4426950408889634f;

struct CrossArgs {
  ief_tensor4 q, k, v, o;
  int32_t B, H, Nq, Nk, d, mode;
  float scale_log2;
  const float* mapper;
  const int32_t* mapper_idx;
  const float* refine_alpha;
  const float* equalizer;
  const float* step_alpha;
  float* probs;
  int32_t probs_accum;
  int32_t base_row[IEF_MAX_ROWS], edit_slot[IEF_MAX_ROWS], store_slot[IEF_MAX_ROWS];
};

// softmax(scale * Q[row] K[row]^T) for this CTA's 64-query tile -> dst (fp32 [64][kPLD], cols >= Nk zero)
template <int DTYPE, int DP>
__device__ __forceinline__ void tile_probs(const CrossArgs& a, int src_row, int h, int qt, typename ElemT<DTYPE>::T* sQ,
                                           typename ElemT<DTYPE>::T* sK, float* dst, int tid) {
  using T = typename ElemT<DTYPE>::T;
  constexpr int LD = DP + 8, KS = DP / 16;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  __syncthreads();  // previous users of sQ/sK are done
  const T* qg = reinterpret_cast<const T*>(a.q.ptr) + (int64_t)src_row * a.q.stride_b + (int64_t)h * a.q.stride_h;
  const T* kg = reinterpret_cast<const T*>(a.k.ptr) + (int64_t)src_row * a.k.stride_b + (int64_t)h * a.k.stride_h;
  load_tile<T, kBM, DP, LD, kThreads>(sQ, qg, a.q.stride_n, qt * kBM, a.Nq, a.d, tid);
  load_tile<T, kNKP, DP, LD, kThreads>(sK, kg, a.k.stride_n, 0, a.Nk, a.d, tid);
  __syncthreads();
  float s[kNKP / 8][4];
#pragma unroll
  for (int i = 0; i < kNKP / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    uint32_t qf[4];
    ldsm_x4(qf, &sQ[(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + kk * 16 + (lane >> 4) * 8]);
#pragma unroll
    for (int nb2 = 0; nb2 < kNKP / 16; ++nb2) {
      uint32_t bf[4];
      ldsm_x4(bf, &sK[(nb2 * 16 + (lane & 7) + (lane >> 4) * 8) * LD + kk * 16 + ((lane >> 3) & 1) * 8]);
      mma16816<DTYPE>(s[2 * nb2], qf, bf[0], bf[1]);
      mma16816<DTYPE>(s[2 * nb2 + 1], qf, bf[2], bf[3]);
    }
  }
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nb = 0; nb < kNKP / 8; ++nb) {
    const int c = nb * 8 + 2 * t;
    if (c >= a.Nk) s[nb][0] = s[nb][2] = -INFINITY;
    if (c + 1 >= a.Nk) s[nb][1] = s[nb][3] = -INFINITY;
    m0 = fmaxf(m0, fmaxf(s[nb][0], s[nb][1]));
    m1 = fmaxf(m1, fmaxf(s[nb][2], s[nb][3]));
  }
  m0 = quad_max(m0) * a.scale_log2;
  m1 = quad_max(m1) * a.scale_log2;
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int nb = 0; nb < kNKP / 8; ++nb) {
    s[nb][0] = ief_exp2(fmaf(s[nb][0], a.scale_log2, -m0));
    s[nb][1] = ief_exp2(fmaf(s[nb][1], a.scale_log2, -m0));
    s[nb][2] = ief_exp2(fmaf(s[nb][2], a.scale_log2, -m1));
    s[nb][3] = ief_exp2(fmaf(s[nb][3], a.scale_log2, -m1));
    l0 += s[nb][0] + s[nb][1];
    l1 += s[nb][2] + s[nb][3];
  }
  const float i0 = 1.f / quad_sum(l0), i1 = 1.f / quad_sum(l1);
  float* d0 = dst + (warp * 16 + g) * kPLD;
  float* d1 = d0 + 8 * kPLD;
#pragma unroll
  for (int nb = 0; nb < kNKP / 8; ++nb) {
    const int c = nb * 8 + 2 * t;
    d0[c] = s[nb][0] * i0;
    d0[c + 1] = s[nb][1] * i0;
    d1[c] = s[nb][2] * i1;
    d1[c + 1] = s[nb][3] * i1;
  }
}

template <int DTYPE, int DP>
__global__ void __launch_bounds__(kThreads)
cross_attn_edit_kernel(const __grid_constant__ CrossArgs a) {
  using E = ElemT<DTYPE>;
  using T = typename E::T;
  constexpr int LD = DP + 8, KS = DP / 16, NB = DP / 8;
  const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
  extern __shared__ uint4 smem4[];
  T* sQ = reinterpret_cast<T*>(smem4);
  T* sK = sQ + kBM * LD;
  T* sV = sK + kNKP * LD;
  float* sP = reinterpret_cast<float*>(sV + kNKP * LD);
  float* sPb = sP + kBM * kPLD;
  float* sM = sPb + kBM * kPLD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int base = a.base_row[b], slot = a.edit_slot[b], Nk = a.Nk;

  if (base >= 0) {
    if (a.mode == IEF_EDIT_REPLACE)
      for (int i = tid; i < Nk * Nk; i += kThreads) sM[i] = __ldg(a.mapper + (int64_t)slot * Nk * Nk + i);
    tile_probs<DTYPE, DP>(a, base, h, qt, sQ, sK, sPb, tid);
  }
  tile_probs<DTYPE, DP>(a, b, h, qt, sQ, sK, sP, tid);
  {
    const T* vg = reinterpret_cast<const T*>(a.v.ptr) + (int64_t)b * a.v.stride_b + (int64_t)h * a.v.stride_h;
    load_tile<T, kNKP, DP, LD, kThreads>(sV, vg, a.v.stride_n, 0, Nk, a.d, tid);
  }
  __syncthreads();
  if (base >= 0) {
    // P' = edit(base, P) * alpha + (1 - alpha) * P      (attention_base.py:119-120)
    const float* al = a.step_alpha + (int64_t)slot * Nk;
    const float* eq = a.equalizer ? a.equalizer + (int64_t)slot * Nk : nullptr;
    for (int i = tid; i < kBM * Nk; i += kThreads) {
      const int r = i / Nk, n = i - r * Nk;
      const float pb = sP[r * kPLD + n];
      float e;
      if (a.mode == IEF_EDIT_REPLACE) {
        e = 0.f;
        for (int w = 0; w < Nk; ++w) e = fmaf(sPb[r * kPLD + w], sM[w * Nk + n], e);
      } else if (a.mode == IEF_EDIT_REFINE) {
        int idx = __ldg(a.mapper_idx + (int64_t)slot * Nk + n);
        if (idx < 0) idx += Nk;  // torch advanced indexing wraps -1 to the last column (attention_control.py:29)
        const float ra = __ldg(a.refine_alpha + (int64_t)slot * Nk + n);
        e = sPb[r * kPLD + idx] * ra + pb * (1.f - ra);
      } else {
        e = sPb[r * kPLD + n];
      }
      if (eq) e *= __ldg(eq + n);
      const float av = __ldg(al + n);
      sP[r * kPLD + n] = e * av + (1.f - av) * pb;
    }
    __syncthreads();
  }
  const int sslot = a.store_slot[b];
  if (a.probs != nullptr && sslot >= 0) {
    // rows of this tile are contiguous in the [slot, H, Nq, Nk] store: fully coalesced
    const int rows = min(kBM, a.Nq - qt * kBM);
    float* dst = a.probs + (((int64_t)sslot * a.H + h) * a.Nq + (int64_t)qt * kBM) * Nk;
    for (int i = tid; i < rows * Nk; i += kThreads) {
      const int r = i / Nk, n = i - r * Nk;
      const float v = sP[r * kPLD + n];
      dst[i] = a.probs_accum ? dst[i] + v : v;
    }
  }
  // O = P' V
  float o[NB][4];
#pragma unroll
  for (int i = 0; i < NB; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  const float* p0 = sP + (warp * 16 + g) * kPLD;
  const float* p1 = p0 + 8 * kPLD;
#pragma unroll
  for (int kk = 0; kk < kNKP / 16; ++kk) {
    const int c = kk * 16 + 2 * t;
    uint32_t pa[4];
    pa[0] = E::pack(p0[c], p0[c + 1]);
    pa[1] = E::pack(p1[c], p1[c + 1]);
    pa[2] = E::pack(p0[c + 8], p0[c + 9]);
    pa[3] = E::pack(p1[c + 8], p1[c + 9]);
#pragma unroll
    for (int nb2 = 0; nb2 < KS; ++nb2) {
      uint32_t vf[4];
      ldsm_x4_t(vf, &sV[(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + nb2 * 16 + (lane >> 4) * 8]);
      mma16816<DTYPE>(o[2 * nb2], pa, vf[0], vf[1]);
      mma16816<DTYPE>(o[2 * nb2 + 1], pa, vf[2], vf[3]);
    }
  }
  const int grow0 = qt * kBM + warp * 16 + g;
  T* og = reinterpret_cast<T*>(a.o.ptr) + (int64_t)b * a.o.stride_b + (int64_t)h * a.o.stride_h;
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    const int c = nb * 8 + 2 * t;
    if (c < a.d) {
      if (grow0 < a.Nq) *reinterpret_cast<uint32_t*>(og + (int64_t)grow0 * a.o.stride_n + c) = E::pack(o[nb][0], o[nb][1]);
      if (grow0 + 8 < a.Nq) *reinterpret_cast<uint32_t*>(og + (int64_t)(grow0 + 8) * a.o.stride_n + c) = E::pack(o[nb][2], o[nb][3]);
    }
  }
}

template <int DTYPE, int DP>
int launch_one(const CrossArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int smem = (kBM + 2 * kNKP) * (DP + 8) * 2 + 2 * kBM * kPLD * 4 + kNKP * kNKP * 4;
  auto kern = cross_attn_edit_kernel<DTYPE, DP>;
  static bool configured = false;
  if (!configured) {
    IEF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  kern<<<grid, kThreads, smem, st>>>(a);
  IEF_LAUNCH_OK("cross_attn_edit_kernel");
  return IEF_OK;
}

template <int DTYPE>
int launch_dp(const CrossArgs& a, dim3 grid, cudaStream_t st) {
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

extern "C" int ief_cross_attn_edit_fwd(const ief_cross_params* p, void* stream) {
  IEF_REQUIRE(p != nullptr, IEF_ERR_INVALID, "ief_cross_attn_edit_fwd: null params");
  IEF_REQUIRE(p->q.ptr && p->k.ptr && p->v.ptr && p->o.ptr, IEF_ERR_INVALID, "ief_cross_attn_edit_fwd: null tensor pointer");
  IEF_REQUIRE(p->dtype == IEF_BF16 || p->dtype == IEF_F16, IEF_ERR_UNSUPPORTED, "ief_cross_attn_edit_fwd: dtype must be bf16 or f16");
  IEF_REQUIRE(p->B >= 1 && p->B <= IEF_MAX_ROWS && p->H >= 1 && p->Nq >= 1, IEF_ERR_INVALID, "ief_cross_attn_edit_fwd: bad B/H/Nq");
  IEF_REQUIRE(p->Nk >= 1 && p->Nk <= kNKP, IEF_ERR_UNSUPPORTED, "ief_cross_attn_edit_fwd: Nk=%d, at most %d keys supported", p->Nk, kNKP);
  IEF_REQUIRE(p->d % 8 == 0 && p->d >= 8 && p->d <= 160, IEF_ERR_UNSUPPORTED, "ief_cross_attn_edit_fwd: head_dim %d unsupported", p->d);
  const ief_tensor4* ts[4] = {&p->q, &p->k, &p->v, &p->o};
  for (auto tt : ts) {
    IEF_REQUIRE((reinterpret_cast<uintptr_t>(tt->ptr) & 15) == 0, IEF_ERR_INVALID, "ief_cross_attn_edit_fwd: pointer not 16-byte aligned");
    IEF_REQUIRE(tt->stride_n % 8 == 0 && tt->stride_h % 8 == 0 && tt->stride_b % 8 == 0, IEF_ERR_UNSUPPORTED,
                "ief_cross_attn_edit_fwd: strides must be multiples of 8 elements");
  }
  CrossArgs a;
  a.q = p->q; a.k = p->k; a.v = p->v; a.o = p->o;
  a.B = p->B; a.H = p->H; a.Nq = p->Nq; a.Nk = p->Nk; a.d = p->d; a.mode = p->mode;
  a.scale_log2 = p->scale * kLog2e;
  a.mapper = p->mapper; a.mapper_idx = p->mapper_idx; a.refine_alpha = p->refine_alpha;
  a.equalizer = p->equalizer; a.step_alpha = p->step_alpha;
  a.probs = p->probs_out; a.probs_accum = p->probs_accum;
  bool any_edit = false;
  for (int i = 0; i < p->B; ++i) {
    a.base_row[i] = p->base_row ? p->base_row[i] : -1;
    a.edit_slot[i] = p->edit_slot ? p->edit_slot[i] : 0;
    a.store_slot[i] = p->store_slot ? p->store_slot[i] : i;
    IEF_REQUIRE(a.base_row[i] < p->B, IEF_ERR_INVALID, "ief_cross_attn_edit_fwd: base_row[%d]=%d out of range", i, a.base_row[i]);
    if (a.base_row[i] >= 0) {
      any_edit = true;
      IEF_REQUIRE(a.edit_slot[i] >= 0 && a.edit_slot[i] < p->n_slots, IEF_ERR_INVALID, "ief_cross_attn_edit_fwd: edit_slot[%d] out of range", i);
    }
  }
  if (any_edit) {
    IEF_REQUIRE(p->step_alpha != nullptr, IEF_ERR_INVALID, "ief_cross_attn_edit_fwd: step_alpha required when a row is edited");
    IEF_REQUIRE(p->mode != IEF_EDIT_REPLACE || p->mapper, IEF_ERR_INVALID, "ief_cross_attn_edit_fwd: mapper required for REPLACE");
    IEF_REQUIRE(p->mode != IEF_EDIT_REFINE || (p->mapper_idx && p->refine_alpha), IEF_ERR_INVALID,
                "ief_cross_attn_edit_fwd: mapper_idx and refine_alpha required for REFINE");
    IEF_REQUIRE(p->mode >= IEF_EDIT_NONE && p->mode <= IEF_EDIT_REFINE, IEF_ERR_INVALID, "ief_cross_attn_edit_fwd: bad mode %d", p->mode);
  }
  dim3 grid(ief_ceil_div(p->Nq, kBM), p->H, p->B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return p->dtype == IEF_BF16 ? launch_dp<IEF_BF16>(a, grid, st) : launch_dp<IEF_F16>(a, grid, st);
}
