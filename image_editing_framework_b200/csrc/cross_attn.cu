// 77-key cross-attention with the Prompt-to-Prompt probability edit fused (ief_cross_attn_edit_fwd).
// Replaces, per cross-attention layer, the chain
//   get_attention_scores (p2p/model/register.py:47) -> AttentionControl.__call__ (attention_base.py:16-28)
//   -> AttentionControlEdit.forward (attention_base.py:113-125) -> replace_cross_attention
//      (attention_control.py:15-16 replace, :28-31 refine, :42-46 reweight) -> bmm (register.py:50)
//   and AttentionStore.forward's capture of the post-edit maps (attention_base.py:64-68)
// by one kernel. HBM / latency bound (AI ~ 76 flop/B): Q and O are touched once.
//
// A CTA is 64 query rows of one (batch row, head); each of its 4 warps owns 16 rows and keeps their <= 80 probabilities
// in mma.sync accumulator registers from QK^T to P'V (no shared-memory round trip). Edited rows first compute the BASE
// row's probabilities for the same queries and park them in a per-warp fp32 staging area (only __syncwarp needed), then
// apply the edit in registers:
//     E = edit(base, P);  E *= equalizer;  P' = E * alpha + (1 - alpha) * P
// The replace mapper is applied in its sparse form (a word swap touches 1-4 source tokens per target token); a dense
// 77x77 fp32 multiply on the CUDA cores is kept as a fallback for mappers with more than 8 non-zeros in a column.
// Maps are streamed to / accumulated into the store straight from the registers (each quad writes 32-byte pieces).
// Shared memory is sized per mode (25 KB for plain rows at head_dim 40), so 4-8 CTAs are resident per SM.
#include "mma_utils.cuh"
#include <math.h>
#include <stdlib.h>

bool ief_cross_tc_supported(const ief_cross_params* p);
int ief_cross_tc_launch(const ief_cross_params* p, cudaStream_t st, const int32_t* rows = nullptr, int n_rows = 0, int pdl = 0);
int ief_cross_tc_edit_launch(const ief_cross_params* p, const int32_t* rows, int n_rows, cudaStream_t st);

using namespace mmau;

namespace {

constexpr int kBM = 64, kNKP = 80, kPLD = 81, kThreads = 128, kNZ = 8;
constexpr float kLog2e = 1.4426950408889634f;

enum : int { kPlain = 0, kEditGather = 1, kEditDense = 2 };  // shared-memory / code flavour of a launch

struct CrossArgs {
  ief_tensor4 q, k, v, o;
  int32_t B, H, Nq, Nk, d, mode;
  float scale_log2;
  const float* mapper;          // dense [slots, Nk, Nk]
  const int32_t* mapper_nz_idx; // sparse [slots, Nk, kNZ] (source token per non-zero, -1 padded)
  const float* mapper_nz_w;     // sparse [slots, Nk, kNZ]
  const int32_t* mapper_idx;    // refine gather [slots, Nk]
  const float* refine_alpha;
  const float* equalizer;
  const float* step_alpha;
  float* probs;
  int32_t probs_accum;
  int32_t base_row[IEF_MAX_ROWS], edit_slot[IEF_MAX_ROWS], store_slot[IEF_MAX_ROWS];
};

// softmax(scale * Q[row] K[row]^T) of this warp's 16 queries as normalised fp32 accumulator fragments:
// s[nb][0..1] = row g, columns nb*8 + 2t, +1;  s[nb][2..3] = row g+8. Columns >= Nk come out as 0.
template <int DTYPE, int DP>
__device__ __forceinline__ void warp_probs(const CrossArgs& a, const typename ElemT<DTYPE>::T* sQ, const typename ElemT<DTYPE>::T* sK,
                                           float (&s)[kNKP / 8][4], int warp, int lane) {
  constexpr int LD = DP + 8, KS = DP / 16;
  const int t = lane & 3;
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
#pragma unroll
  for (int nb = 0; nb < kNKP / 8; ++nb) {
    s[nb][0] *= i0; s[nb][1] *= i0; s[nb][2] *= i1; s[nb][3] *= i1;
  }
}

template <int FLAVOUR, int DP> constexpr int cross_smem_bytes() {
  int b = (kBM + 2 * kNKP) * (DP + 8) * 2;                       // Q, K, V tiles
  if (FLAVOUR != kPlain) b += (kBM + kNKP) * (DP + 8) * 2;        // base row's Q, K tiles (fetched in the same round trip)
  if (FLAVOUR != kPlain) b += 4 * 16 * kPLD * 4;                  // per-warp base-probability staging
  if (FLAVOUR == kEditGather) b += kNKP * kNZ * 8;                // sparse mapper (idx + weight)
  if (FLAVOUR == kEditDense) b += kNKP * kNKP * 4;                // dense mapper
  return b;
}

template <int DTYPE, int DP, int FLAVOUR>
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
  T* sQb = sV + kNKP * LD;                                        // base row's tiles (edit flavours)
  T* sKb = sQb + kBM * LD;
  float* sPb = reinterpret_cast<float*>(FLAVOUR == kPlain ? sQb : sKb + kNKP * LD);  // [4 warps][16][kPLD]   (edit flavours)
  float* sM = sPb + 4 * 16 * kPLD;                                // dense mapper / sparse weights
  int32_t* sMi = reinterpret_cast<int32_t*>(sM + kNKP * kNZ);     // sparse indices (gather flavour)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int base = FLAVOUR == kPlain ? -1 : a.base_row[b];
  const int slot = a.edit_slot[b], Nk = a.Nk;
  const bool edited = base >= 0;

  auto qptr = [&](int row) { return reinterpret_cast<const T*>(a.q.ptr) + (int64_t)row * a.q.stride_b + (int64_t)h * a.q.stride_h; };
  auto kptr = [&](int row) { return reinterpret_cast<const T*>(a.k.ptr) + (int64_t)row * a.k.stride_b + (int64_t)h * a.k.stride_h; };
  // Every tile this CTA needs is requested up front with cp.async: group 0 = the base row's Q/K (edited rows), group 1 = own
  // Q/K/V. One memory round trip per CTA; the base probabilities are computed while group 1 is still landing.
  if (FLAVOUR != kPlain && edited) {
    load_tile_async<T, kBM, DP, LD, kThreads>(sQb, qptr(base), a.q.stride_n, qt * kBM, a.Nq, a.d, tid);
    load_tile_async<T, kNKP, DP, LD, kThreads>(sKb, kptr(base), a.k.stride_n, 0, Nk, a.d, tid);
  }
  cp_async_commit();
  load_tile_async<T, kBM, DP, LD, kThreads>(sQ, qptr(b), a.q.stride_n, qt * kBM, a.Nq, a.d, tid);
  load_tile_async<T, kNKP, DP, LD, kThreads>(sK, kptr(b), a.k.stride_n, 0, Nk, a.d, tid);
  load_tile_async<T, kNKP, DP, LD, kThreads>(sV, reinterpret_cast<const T*>(a.v.ptr) + (int64_t)b * a.v.stride_b + (int64_t)h * a.v.stride_h, a.v.stride_n, 0,
                                             Nk, a.d, tid);
  cp_async_commit();
  float s[kNKP / 8][4];
  float* wPb = sPb + warp * 16 * kPLD;
  if (FLAVOUR != kPlain && edited) {
    if (FLAVOUR == kEditDense && a.mode == IEF_EDIT_REPLACE)
      for (int i = tid; i < Nk * Nk; i += kThreads) sM[i] = __ldg(a.mapper + (int64_t)slot * Nk * Nk + i);
    if (FLAVOUR == kEditGather && a.mode == IEF_EDIT_REPLACE)
      for (int i = tid; i < Nk * kNZ; i += kThreads) {
        sM[i] = __ldg(a.mapper_nz_w + (int64_t)slot * Nk * kNZ + i);
        sMi[i] = __ldg(a.mapper_nz_idx + (int64_t)slot * Nk * kNZ + i);
      }
    cp_async_wait<1>();
    __syncthreads();
    warp_probs<DTYPE, DP>(a, sQb, sKb, s, warp, lane);
#pragma unroll
    for (int nb = 0; nb < kNKP / 8; ++nb) {  // park the base probabilities of this warp's 16 rows (read back by this warp only)
      const int c = nb * 8 + 2 * t;
      wPb[g * kPLD + c] = s[nb][0];
      wPb[g * kPLD + c + 1] = s[nb][1];
      wPb[(g + 8) * kPLD + c] = s[nb][2];
      wPb[(g + 8) * kPLD + c + 1] = s[nb][3];
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  warp_probs<DTYPE, DP>(a, sQ, sK, s, warp, lane);

  if (FLAVOUR != kPlain && edited) {
    // P' = edit(base, P) * alpha + (1 - alpha) * P      (attention_base.py:119-120), in registers
    const float* al = a.step_alpha + (int64_t)slot * Nk;
    const float* eq = a.equalizer ? a.equalizer + (int64_t)slot * Nk : nullptr;
#pragma unroll
    for (int nb = 0; nb < kNKP / 8; ++nb) {
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        const int n = nb * 8 + 2 * t + e2;
        if (n < Nk) {
          const float av = __ldg(al + n), eqv = eq ? __ldg(eq + n) : 1.f;
          float ev[2];
          if (a.mode == IEF_EDIT_REPLACE) {
            ev[0] = ev[1] = 0.f;
            if (FLAVOUR == kEditDense) {
              for (int w = 0; w < Nk; ++w) {
                const float m = sM[w * Nk + n];
                ev[0] = fmaf(wPb[g * kPLD + w], m, ev[0]);
                ev[1] = fmaf(wPb[(g + 8) * kPLD + w], m, ev[1]);
              }
            } else {
#pragma unroll
              for (int z = 0; z < kNZ; ++z) {
                const int w = sMi[n * kNZ + z];
                if (w >= 0) {
                  const float m = sM[n * kNZ + z];
                  ev[0] = fmaf(wPb[g * kPLD + w], m, ev[0]);
                  ev[1] = fmaf(wPb[(g + 8) * kPLD + w], m, ev[1]);
                }
              }
            }
          } else if (a.mode == IEF_EDIT_REFINE) {
            int idx = __ldg(a.mapper_idx + (int64_t)slot * Nk + n);
            if (idx < 0) idx += Nk;  // torch advanced indexing wraps -1 to the last column (attention_control.py:29)
            const float ra = __ldg(a.refine_alpha + (int64_t)slot * Nk + n);
            ev[0] = wPb[g * kPLD + idx] * ra + s[nb][e2] * (1.f - ra);
            ev[1] = wPb[(g + 8) * kPLD + idx] * ra + s[nb][2 + e2] * (1.f - ra);
          } else {
            ev[0] = wPb[g * kPLD + n];
            ev[1] = wPb[(g + 8) * kPLD + n];
          }
          s[nb][e2] = ev[0] * eqv * av + (1.f - av) * s[nb][e2];
          s[nb][2 + e2] = ev[1] * eqv * av + (1.f - av) * s[nb][2 + e2];
        }
      }
    }
  }
  const int grow0 = qt * kBM + warp * 16 + g;
  const int sslot = a.store_slot[b];
  if (a.probs != nullptr && sslot >= 0) {
    // post-edit maps -> store (overwrite or accumulate); each quad covers 8 consecutive floats of a row
    float* p0 = a.probs + (((int64_t)sslot * a.H + h) * a.Nq + grow0) * Nk;
    float* p1 = p0 + (int64_t)8 * Nk;
#pragma unroll
    for (int nb = 0; nb < kNKP / 8; ++nb) {
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        const int n = nb * 8 + 2 * t + e2;
        if (n < Nk) {
          if (grow0 < a.Nq) p0[n] = a.probs_accum ? p0[n] + s[nb][e2] : s[nb][e2];
          if (grow0 + 8 < a.Nq) p1[n] = a.probs_accum ? p1[n] + s[nb][2 + e2] : s[nb][2 + e2];
        }
      }
    }
  }
  // O = P' V  (probabilities straight from the accumulator fragments)
  float o[NB][4];
#pragma unroll
  for (int i = 0; i < NB; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < kNKP / 16; ++kk) {
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

template <int DTYPE, int DP, int FLAVOUR>
int launch_one(const CrossArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int smem = cross_smem_bytes<FLAVOUR, DP>();
  auto kern = cross_attn_edit_kernel<DTYPE, DP, FLAVOUR>;
  IEF_CONFIG_SMEM(kern, smem);
  kern<<<grid, kThreads, smem, st>>>(a);
  IEF_LAUNCH_OK("cross_attn_edit_kernel");
  return IEF_OK;
}

template <int DTYPE, int FLAVOUR>
int launch_dp(const CrossArgs& a, dim3 grid, cudaStream_t st) {
  const int d = a.d;
  if (d <= 32) return launch_one<DTYPE, 32, FLAVOUR>(a, grid, st);
  if (d <= 48) return launch_one<DTYPE, 48, FLAVOUR>(a, grid, st);
  if (d <= 64) return launch_one<DTYPE, 64, FLAVOUR>(a, grid, st);
  if (d <= 80) return launch_one<DTYPE, 80, FLAVOUR>(a, grid, st);
  if (d <= 96) return launch_one<DTYPE, 96, FLAVOUR>(a, grid, st);
  if (d <= 128) return launch_one<DTYPE, 128, FLAVOUR>(a, grid, st);
  return launch_one<DTYPE, 160, FLAVOUR>(a, grid, st);
}

template <int DTYPE>
int launch_flavour(const CrossArgs& a, int flavour, dim3 grid, cudaStream_t st) {
  if (flavour == kPlain) return launch_dp<DTYPE, kPlain>(a, grid, st);
  if (flavour == kEditGather) return launch_dp<DTYPE, kEditGather>(a, grid, st);
  return launch_dp<DTYPE, kEditDense>(a, grid, st);
}

}  // namespace

namespace { thread_local const char* g_last_cross_impl = "none"; }
extern "C" const char* ief_last_cross_impl(void) { return g_last_cross_impl; }

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
  a.mapper = p->mapper; a.mapper_nz_idx = p->mapper_nz_idx; a.mapper_nz_w = p->mapper_nz_w;
  a.mapper_idx = p->mapper_idx; a.refine_alpha = p->refine_alpha;
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
  int flavour = kPlain;
  if (any_edit) {
    IEF_REQUIRE(p->step_alpha != nullptr, IEF_ERR_INVALID, "ief_cross_attn_edit_fwd: step_alpha required when a row is edited");
    IEF_REQUIRE(p->mode >= IEF_EDIT_NONE && p->mode <= IEF_EDIT_REFINE, IEF_ERR_INVALID, "ief_cross_attn_edit_fwd: bad mode %d", p->mode);
    IEF_REQUIRE(p->mode != IEF_EDIT_REPLACE || p->mapper || (p->mapper_nz_idx && p->mapper_nz_w), IEF_ERR_INVALID,
                "ief_cross_attn_edit_fwd: REPLACE needs mapper or mapper_nz_idx + mapper_nz_w");
    IEF_REQUIRE(p->mode != IEF_EDIT_REFINE || (p->mapper_idx && p->refine_alpha), IEF_ERR_INVALID,
                "ief_cross_attn_edit_fwd: mapper_idx and refine_alpha required for REFINE");
    flavour = (p->mode == IEF_EDIT_REPLACE && !(p->mapper_nz_idx && p->mapper_nz_w)) ? kEditDense : kEditGather;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // Tensor-pipe kernels (>= 128 queries): un-edited rows without a map output on cross_tc.cu, edited and / or stored rows on
  // cross_tc_edit.cu (sparse replace mapper, refine gather, reweight, alpha blend, store epilogue). What stays on this file's
  // mma.sync kernel: layers with fewer than 128 queries (8x8 latents) and the dense 77x77 mapper fallback.
  if (flavour != kEditDense && ief_cross_tc_supported(p)) {
    static int edit_on = -1;
    if (edit_on < 0) { const char* e = getenv("IEF_CROSS_TC_EDIT"); edit_on = (e && e[0] == '0') ? 0 : 1; }
    int32_t special[IEF_MAX_ROWS], plain[IEF_MAX_ROWS];
    int ns = 0, np = 0;
    for (int i = 0; i < p->B; ++i)   // edited rows first: their CTAs run longest
      if (a.base_row[i] >= 0) special[ns++] = i;
    for (int i = 0; i < p->B; ++i) {
      if (a.base_row[i] >= 0) continue;
      if (p->probs_out != nullptr && a.store_slot[i] >= 0) special[ns++] = i; else plain[np++] = i;
    }
    if (ns == 0) {
      g_last_cross_impl = "tcgen05";
      return ief_cross_tc_launch(p, st);
    }
    if (edit_on) {
      g_last_cross_impl = "tcgen05-edit";
      // Where do the call's plain rows run? (a) in the edit kernel's launch, after the edited ones (longest CTAs first), or (b) on the
      // leaner plain kernel (128 TMEM columns: four CTAs per SM instead of two) in a second launch that is a PROGRAMMATIC DEPENDENT of
      // the edit launch: disjoint rows, same inputs, so the two grids run side by side. Measured, P2P replace at B=4 (profiles/
      // r02_cross_launch_modes.txt): N=4096,d=40  (a) 29.7 us  (b) 25.6;  N=1024,d=80  19.4 / 17.6;  N=256,d=160  15.3 / 15.4;
      // with stored maps N=1024  23.6 / 25.6 — and a plain, serialised second launch is the slowest everywhere (33.7 / 23.6 / 23.5).
      // Hence (b) for >= 1024 queries without a map output, (a) otherwise. IEF_CROSS_TC_ONE_LAUNCH=1|2|0 forces (a) / (b) / serialised.
      static int one = -1;
      if (one < 0) { const char* e = getenv("IEF_CROSS_TC_ONE_LAUNCH"); one = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : -1; }
      const int how = one >= 0 ? one : ((p->probs_out == nullptr && p->Nq >= 1024 && np > 0) ? 2 : 1);
      if (how == 1) {
        for (int i = 0; i < np; ++i) special[ns + i] = plain[i];
        return ief_cross_tc_edit_launch(p, special, ns + np, st);
      }
      int rc = ief_cross_tc_edit_launch(p, special, ns, st);
      if (rc != IEF_OK || np == 0) return rc;
      return ief_cross_tc_launch(p, st, plain, np, how == 2 ? 1 : 0);
    }
  }
  g_last_cross_impl = "mma";
  dim3 grid(ief_ceil_div(p->Nq, kBM), p->H, p->B);
  return p->dtype == IEF_BF16 ? launch_flavour<IEF_BF16>(a, flavour, grid, st) : launch_flavour<IEF_F16>(a, flavour, grid, st);
}
