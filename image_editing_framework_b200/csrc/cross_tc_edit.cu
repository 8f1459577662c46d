// 77-key cross-attention rows that are EDITED (Prompt-to-Prompt replace / refine / reweight, alpha blend) and / or whose post-edit
// probability maps are STORED (AttentionStore), on the tensor pipe. The plain, unstored rows of the same call run on cross_tc.cu.
//
//   P      = softmax(scale Q[b]    K[b]^T)                                  (p2p/model/register.py:47)
//   P_base = softmax(scale Q[base] K[base]^T)                               (attention_base.py:115-117: attn_base = attn[0])
//   E      = replace: sum_j w[n][j] P_base[idx[n][j]]   (attention_control.py:15-16, sparse form of the mapper einsum)
//            refine : P_base[mapper_idx[n]] ra[n] + P[n] (1 - ra[n])        (attention_control.py:28-31; -1 wraps to the last key)
//            none   : P_base[n]
//   P'     = E eq[n] alpha[n] + (1 - alpha[n]) P                            (attention_control.py:42-46, attention_base.py:119-120)
//   store[slot] (+)= P'                                                     (attention_base.py:64-68 after the edit)
//   O      = P' V                                                           (register.py:50)
//
// One CTA = one 128-row query tile of one (row, head). Warp 4 drives TMA (own Q/K/V and the base row's Q/K in one round trip) and
// issues S = Q K^T, S_base = Q_base K_base^T and O = P' V as tcgen05.mma groups; warps 0-3 own a score row per thread (TMEM lane):
//   * exponentials are written back over the scores in TMEM (fp32), so that the three passes (maximum, exponentials, edit) are short
//     rolled loops over 16-column pieces: the code runs once per CTA and its size is its cost,
//   * the base row's exponentials in TMEM serve as the per-row gather table: the
//     mapper index of a target token is the same for every row, so `P_base[idx[n]]` is one tcgen05.ld.x1 at a warp-uniform,
//     run-time column address — no shared-memory staging of probabilities and no dynamic register indexing,
//   * P' (normalised, 16-bit pairs) overwrites the consumed score columns and feeds the PV MMA from TMEM; O leaves unscaled,
//   * stored maps go through a 2 KB per-warp transposition buffer so that global accesses are 64-byte row pieces, not 32 scattered words.
// TMEM: S 0-79 | S_base 80-159 | O from 160 (EDIT) — 256 columns up to head_dim 96, 512 above; without EDIT: S 0-79 | O from 80.
#include "ief_common.cuh"
#include "ptx_sm100.cuh"
#include "attn_tc_host.cuh"
#include <math.h>
#include <stdlib.h>

using namespace sm100;

namespace {

constexpr int kBM = 128, kNK = 80, kThreads = 160, kNZ = 8;
constexpr int kQChunk = 128 * 128, kKVChunk = kNK * 128;   // bytes per 64-channel chunk
constexpr int kStageLd = 17;                               // floats per staged row piece (16 + 1 against bank conflicts)
// EDIT tables: int4-sized record per target token {first gather column, A, B, sources beyond the first} with
//   P'[n] = P_base[idx0] * A + P[n] * B,  A = w0 eq alpha,  B = keep eq alpha + (1 - alpha)      (+ further sources, below)
// then the flat list of further sources {token within its 16-token piece, column, weight * eq * alpha}, ordered by token, and the
// list position at which each 16-token piece starts
constexpr int kMaxExtra = kNK * (kNZ - 1);
constexpr int kTableBytes = kNK * 16 + kMaxExtra * 12 + 8 * 4 + kNK * 4;
constexpr float kLog2e = 1.4426950408889634f;

struct CrossTcEditArgs {
  void* o;
  int64_t o_sb, o_sn, o_sh;
  int32_t B, H, Nq, Nk, d, ksteps_qk, dv_mma, mode;
  uint32_t idesc_qk, idesc_pv;
  float scale_log2;
  int32_t perm_q[3], perm_k[3], perm_v[3];
  const int32_t* mapper_nz_idx;
  const float* mapper_nz_w;
  const int32_t* mapper_idx;
  const float* refine_alpha;
  const float* equalizer;
  const float* step_alpha;
  float* probs;
  int32_t probs_accum;
  // the rows this launch covers (blockIdx.z indexes these arrays): edited rows first so that their longer CTAs start first
  int32_t row[IEF_MAX_ROWS], base_row[IEF_MAX_ROWS], edit_slot[IEF_MAX_ROWS], store_slot[IEF_MAX_ROWS];
};

template <int DCH, bool EDIT, bool STORE> constexpr int cross_tc_edit_smem() {
  int b = DCH * (kQChunk + 2 * kKVChunk);
  if (EDIT) b += DCH * (kQChunk + kKVChunk);          // base row's Q and K tiles
  if (EDIT) b += kTableBytes;                          // per-token tables + the list of further sources
  if (STORE) b += 4 * 32 * kStageLd * 4;               // per-warp transposition buffer
  return b + 1024 + 128;
}

__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}

template <int DTYPE, int DCH, int TCOLS, bool EDIT, bool STORE>
__global__ void __launch_bounds__(kThreads)
cross_tc_edit_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CrossTcEditArgs a) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // nothing of this grid is consumed by a dependent launch (cross_attn.cu)
  using E = ElemT<DTYPE>;
  constexpr int OFF_B = kNK, OFF_O = EDIT ? 2 * kNK : kNK;
  const int zi = blockIdx.z, b = a.row[zi], h = blockIdx.y, qt = blockIdx.x;
  const int base_b = EDIT ? a.base_row[zi] : -1;
  const bool edited = EDIT && base_b >= 0;
  const int slot = a.edit_slot[zi];
  const int sslot = STORE ? a.store_slot[zi] : -1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base, sK = sQ + DCH * kQChunk, sV = sK + DCH * kKVChunk;
  const uint32_t sQb = sV + DCH * kKVChunk, sKb = sQb + DCH * kQChunk;                 // EDIT only
  constexpr int kTilesEnd = DCH * (kQChunk + 2 * kKVChunk) + (EDIT ? DCH * (kQChunk + kKVChunk) : 0);
  int4* sTok = reinterpret_cast<int4*>(base_ptr + kTilesEnd);                           // EDIT tables, see kTableBytes
  int32_t* sExTok = reinterpret_cast<int32_t*>(sTok + kNK);
  int32_t* sExIdx = sExTok + kMaxExtra;
  float* sExW = reinterpret_cast<float*>(sExIdx + kMaxExtra);
  int32_t* sExBeg = reinterpret_cast<int32_t*>(sExW + kMaxExtra);                       // [kNK / 16 + 1] (+ padding)
  int32_t* sExCnt = sExBeg + 8;                                                         // [kNK] scratch: further sources per token
  constexpr int kTablesEnd = kTilesEnd + (EDIT ? kTableBytes : 0);
  float* stage = reinterpret_cast<float*>(base_ptr + kTablesEnd);                      // STORE only: [4 warps][32][kStageLd]
  const uint32_t bar0 = base + kTablesEnd + (STORE ? 4 * 32 * kStageLd * 4 : 0);
  const uint32_t bar_in = bar0, bar_s = bar0 + 8, bar_p = bar0 + 16, bar_o = bar0 + 24;
  const uint32_t tmem_slot = bar0 + 32;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = a.Nk;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_in, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, TCOLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 4) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_in, DCH * (kQChunk + 2 * kKVChunk) + (edited ? DCH * (kQChunk + kKVChunk) : 0));
#pragma unroll
      for (int c = 0; c < DCH; ++c) {
        if (edited) {  // the base row's tiles first: its softmax is the first thing the row threads need after their own scores
          tc_tma_tile(sQb + c * kQChunk, &tmQ, bar_in, c * 64, qt * kBM, h, base_b, a.perm_q);
          tc_tma_tile(sKb + c * kKVChunk, &tmK, bar_in, c * 64, 0, h, base_b, a.perm_k);
        }
        tc_tma_tile(sQ + c * kQChunk, &tmQ, bar_in, c * 64, qt * kBM, h, b, a.perm_q);
        tc_tma_tile(sK + c * kKVChunk, &tmK, bar_in, c * 64, 0, h, b, a.perm_k);
        tc_tma_tile(sV + c * kKVChunk, &tmV, bar_in, c * 64, 0, h, b, a.perm_v);
      }
    }
    __syncwarp();
    mbar_wait(bar_in, 0);
    tc_fence_after();
    if (elect_one()) {
      for (int k = 0; k < a.ksteps_qk; ++k)
        umma_ss(tmem_base, make_smem_desc_sw128(sQ + (k >> 2) * kQChunk + (k & 3) * 32, 16, 1024),
                make_smem_desc_sw128(sK + (k >> 2) * kKVChunk + (k & 3) * 32, 16, 1024), a.idesc_qk, k > 0);
      if (edited)
        for (int k = 0; k < a.ksteps_qk; ++k)
          umma_ss(tmem_base + OFF_B, make_smem_desc_sw128(sQb + (k >> 2) * kQChunk + (k & 3) * 32, 16, 1024),
                  make_smem_desc_sw128(sKb + (k >> 2) * kKVChunk + (k & 3) * 32, 16, 1024), a.idesc_qk, k > 0);
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_p, 0);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < kNK / 16; ++k)
        umma_ts(tmem_base + OFF_O, tmem_base + k * 8, make_smem_desc_sw128(sV + k * 2048, kKVChunk, 1024), a.idesc_pv, k > 0);
      umma_commit(bar_o);
    }
    __syncwarp();
  } else {
    const int tid = threadIdx.x;
    if constexpr (EDIT) {
      // per-token tables, built while the tiles are in flight
      int nex = 0, ex_i[kNZ];
      float ex_w[kNZ], eqa = 0.f;
      if (tid < kNK) {
        const int n = tid;
        int idx0 = 0;
        float w0 = 0.f, keep = 0.f, oma = 0.f;
        if (edited && n < nk) {
          const int64_t tn = (int64_t)slot * nk + n;
          const float al = __ldg(a.step_alpha + tn), eq = a.equalizer ? __ldg(a.equalizer + tn) : 1.f;
          eqa = eq * al;
          oma = 1.f - al;
          if (a.mode == IEF_EDIT_REPLACE) {
            int nnz = 0;
#pragma unroll
            for (int z = 0; z < kNZ; ++z) {
              const int w = __ldg(a.mapper_nz_idx + tn * kNZ + z);
              const float wt = __ldg(a.mapper_nz_w + tn * kNZ + z);
              if (w >= 0 && w < nk) {
                if (nnz == 0) { idx0 = w; w0 = wt; } else { ex_i[nex] = w; ex_w[nex] = wt; ++nex; }
                ++nnz;
              }
            }
          } else if (a.mode == IEF_EDIT_REFINE) {
            int idx = __ldg(a.mapper_idx + tn);
            if (idx < 0) idx += nk;  // torch advanced indexing wraps -1 to the last column (attention_control.py:29)
            const float ra = __ldg(a.refine_alpha + tn);
            idx0 = idx;
            w0 = ra;
            keep = 1.f - ra;
          } else {
            idx0 = n;
            w0 = 1.f;
          }
        }
        sTok[n] = make_int4(idx0, __float_as_int(w0 * eqa), __float_as_int(keep * eqa + oma), 0);
        sExCnt[n] = nex;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid < kNK) {
        int off = 0;
        for (int m2 = 0; m2 < tid; ++m2) off += sExCnt[m2];
        if ((tid & 15) == 0) sExBeg[tid >> 4] = off;
        if (tid == kNK - 1) sExBeg[kNK / 16] = off + nex;
        for (int j = 0; j < nex; ++j) {
          sExTok[off + j] = tid & 15;
          sExIdx[off + j] = ex_i[j];
          sExW[off + j] = ex_w[j] * eqa;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    const int row = warp * 32 + lane;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const uint32_t tS = tmem_base + lane_off, tB = tS + OFF_B, tO = tS + OFF_O;
    const float c2 = a.scale_log2;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    // scores -> exponentials written back over the scores (fp32), returns 1 / row sum. Rolled loops: this kernel runs its code once
    // per CTA, so its size is what it costs (an unrolled first version was 18 K instructions and instruction-fetch bound).
    auto softmax_in_place = [&](uint32_t t0) -> float {
      float m = -INFINITY;
      uint32_t r[16];
#pragma unroll 1
      for (int c = 0; c < kNK / 16; ++c) {
        tmem_ld16(t0 + 16 * c, r);
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (16 * c + e < nk) m = fmaxf(m, __uint_as_float(r[e]));
      }
      const float mc = m * c2;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll 1
      for (int c = 0; c < kNK / 16; ++c) {
        tmem_ld16(t0 + 16 * c, r);
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          float p0 = ief_exp2(fmaf(__uint_as_float(r[e]), c2, -mc)), p1 = ief_exp2(fmaf(__uint_as_float(r[e + 1]), c2, -mc));
          if (16 * c + e >= nk) p0 = 0.f;
          if (16 * c + e + 1 >= nk) p1 = 0.f;
          l0 += p0; l1 += p1;
          r[e] = __float_as_uint(p0);
          r[e + 1] = __float_as_uint(p1);
        }
        tmem_st16(t0 + 16 * c, r);
      }
      tc_wait_st();
      return 1.f / (l0 + l1);
    };
    const float inv_own = softmax_in_place(tS);
    const float inv_base = edited ? softmax_in_place(tB) : 0.f;
    // ---- P' in 16-token pieces: gather, edit, blend, store, pack. P' (8 packed columns per piece) goes over own-row columns whose
    // exponentials have been consumed (piece c overwrites columns 8c .. 8c+7, read as part of pieces <= c).
    float* wst = STORE ? stage + warp * 32 * kStageLd : nullptr;
    const int grow_w = qt * kBM + warp * 32;                 // first query row of this warp
#pragma unroll 1
    for (int c = 0; c < kNK / 16; ++c) {
      float pp[16];
      {
        uint32_t r[16];
        tmem_ld16(tS + 16 * c, r);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) pp[i] = __uint_as_float(r[i]) * inv_own;
      }
      if (edited) {
        uint32_t g[16];
        int4 tk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          tk[i] = sTok[16 * c + i];
          g[i] = tmem_ld1(tB + tk[i].x);
        }
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) pp[i] = __uint_as_float(g[i]) * inv_base * __int_as_float(tk[i].y) + pp[i] * __int_as_float(tk[i].z);
        // further sources of a replaced token (multi-token words): rare; the list is the same for every row
        const int xe = sExBeg[c + 1];
        for (int x = sExBeg[c]; x < xe; ++x) {
          const uint32_t gx = tmem_ld1(tB + sExIdx[x]);
          tc_wait_ld();
          const float val = __uint_as_float(gx) * inv_base * sExW[x];
          const int ii = sExTok[x];
#pragma unroll
          for (int i = 0; i < 16; ++i) pp[i] += i == ii ? val : 0.f;
        }
      }
      if constexpr (STORE) {
        if (sslot >= 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) wst[lane * kStageLd + i] = pp[i];
          __syncwarp();
          const int rsub = lane >> 4, n = 16 * c + (lane & 15);
          float* dst = a.probs + (((int64_t)sslot * a.H + h) * a.Nq + grow_w) * nk + n;
          if (n < nk) {
            float old[16];
#pragma unroll
            for (int it = 0; it < 16; ++it) {
              const int r = 2 * it + rsub;
              old[it] = (a.probs_accum && grow_w + r < a.Nq) ? dst[(int64_t)r * nk] : 0.f;
            }
#pragma unroll
            for (int it = 0; it < 16; ++it) {
              const int r = 2 * it + rsub;
              if (grow_w + r < a.Nq) dst[(int64_t)r * nk] = old[it] + wst[r * kStageLd + (lane & 15)];
            }
          }
          __syncwarp();
        }
      }
      uint32_t u8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) u8[e] = E::pack(pp[2 * e], pp[2 * e + 1]);
      tmem_st8(tS + 8 * c, u8);
    }
    tc_wait_st();
    tc_fence_before();
    mbar_arrive(bar_p);
    // ---- epilogue: O = P' V is already normalised
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const int grow = qt * kBM + row;
    typename E::T* op = reinterpret_cast<typename E::T*>(a.o) + (int64_t)b * a.o_sb + (int64_t)grow * a.o_sn + (int64_t)h * a.o_sh;
    const int nchunk_d = (a.d + 15) >> 4;
    for (int cc = 0; cc < nchunk_d; ++cc) {
      uint32_t r[16];
      tmem_ld16(tO + 16 * cc, r);
      tc_wait_ld();
      if (grow < a.Nq) {
        uint4 v0, v1;
        v0.x = E::pack(__uint_as_float(r[0]), __uint_as_float(r[1]));
        v0.y = E::pack(__uint_as_float(r[2]), __uint_as_float(r[3]));
        v0.z = E::pack(__uint_as_float(r[4]), __uint_as_float(r[5]));
        v0.w = E::pack(__uint_as_float(r[6]), __uint_as_float(r[7]));
        v1.x = E::pack(__uint_as_float(r[8]), __uint_as_float(r[9]));
        v1.y = E::pack(__uint_as_float(r[10]), __uint_as_float(r[11]));
        v1.z = E::pack(__uint_as_float(r[12]), __uint_as_float(r[13]));
        v1.w = E::pack(__uint_as_float(r[14]), __uint_as_float(r[15]));
        if (16 * cc + 8 <= a.d) *reinterpret_cast<uint4*>(op + 16 * cc) = v0;
        if (16 * cc + 16 <= a.d) *reinterpret_cast<uint4*>(op + 16 * cc + 8) = v1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
}

template <int DTYPE, int DCH, bool EDIT, bool STORE>
int launch_flavour(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CrossTcEditArgs& a, dim3 grid, cudaStream_t st) {
  // TMEM columns: EDIT 160 + dv, else 80 + dv, rounded up to a power of two
  const int need = (EDIT ? 2 * kNK : kNK) + a.dv_mma;
  constexpr int smem = cross_tc_edit_smem<DCH, EDIT, STORE>();
#define IEF_LAUNCH_TCOLS(TC)                                                      \
  do {                                                                            \
    auto kern = cross_tc_edit_kernel<DTYPE, DCH, TC, EDIT, STORE>;                \
    IEF_CONFIG_SMEM(kern, smem);                                                  \
    kern<<<grid, kThreads, smem, st>>>(mq, mk, mv, a);                            \
    IEF_LAUNCH_OK("cross_tc_edit_kernel");                                        \
    return IEF_OK;                                                                \
  } while (0)
  // only the (chunks, columns) pairs that can occur are instantiated
  if constexpr (!EDIT && DCH == 1) {
    if (need <= 128) IEF_LAUNCH_TCOLS(128);
  }
  if constexpr (DCH < 3 && !(EDIT && DCH == 3)) {
    if (need <= 256) IEF_LAUNCH_TCOLS(256);
  }
  if constexpr (!EDIT && DCH == 3) IEF_LAUNCH_TCOLS(256);
  if constexpr (EDIT && DCH >= 2) IEF_LAUNCH_TCOLS(512);
  ief_set_error("cross_tc_edit: no kernel for head_dim %d", a.d);
  return IEF_ERR_UNSUPPORTED;
#undef IEF_LAUNCH_TCOLS
}

template <int DTYPE, int DCH>
int launch_dch(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CrossTcEditArgs& a, bool edit, bool store, dim3 grid,
               cudaStream_t st) {
  if (edit) return store ? launch_flavour<DTYPE, DCH, true, true>(mq, mk, mv, a, grid, st) : launch_flavour<DTYPE, DCH, true, false>(mq, mk, mv, a, grid, st);
  return launch_flavour<DTYPE, DCH, false, true>(mq, mk, mv, a, grid, st);
}

template <int DTYPE>
int launch_dtype(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CrossTcEditArgs& a, bool edit, bool store, dim3 grid,
                 cudaStream_t st) {
  if (a.d <= 64) return launch_dch<DTYPE, 1>(mq, mk, mv, a, edit, store, grid, st);
  if (a.d <= 128) return launch_dch<DTYPE, 2>(mq, mk, mv, a, edit, store, grid, st);
  return launch_dch<DTYPE, 3>(mq, mk, mv, a, edit, store, grid, st);
}

}  // namespace

// Rows `rows[0..n_rows)` of the call (edited and / or stored ones), edited rows listed first by the caller.
int ief_cross_tc_edit_launch(const ief_cross_params* p, const int32_t* rows, int n_rows, cudaStream_t st) {
  CrossTcEditArgs a;
  a.o = p->o.ptr; a.o_sb = p->o.stride_b; a.o_sn = p->o.stride_n; a.o_sh = p->o.stride_h;
  a.B = p->B; a.H = p->H; a.Nq = p->Nq; a.Nk = p->Nk; a.d = p->d; a.mode = p->mode;
  a.ksteps_qk = ief_ceil_div(p->d, 16);
  a.dv_mma = ((p->d + 15) / 16) * 16;
  const int fmt = p->dtype == IEF_BF16 ? 1 : 0;
  a.idesc_qk = make_idesc_f16(kBM, kNK, fmt, 0, 0);
  a.idesc_pv = make_idesc_f16(kBM, a.dv_mma, fmt, 0, 1);
  a.scale_log2 = p->scale * kLog2e;
  a.mapper_nz_idx = p->mapper_nz_idx; a.mapper_nz_w = p->mapper_nz_w; a.mapper_idx = p->mapper_idx; a.refine_alpha = p->refine_alpha;
  a.equalizer = p->equalizer; a.step_alpha = p->step_alpha;
  a.probs = p->probs_out; a.probs_accum = p->probs_accum;
  bool edit = false, store = false;
  for (int i = 0; i < IEF_MAX_ROWS; ++i) {
    const int b = i < n_rows ? rows[i] : 0;
    a.row[i] = b;
    a.base_row[i] = (i < n_rows && p->base_row) ? p->base_row[b] : -1;
    a.edit_slot[i] = (i < n_rows && p->edit_slot) ? p->edit_slot[b] : 0;
    a.store_slot[i] = (i < n_rows && p->probs_out) ? (p->store_slot ? p->store_slot[b] : b) : -1;
    if (i < n_rows && a.base_row[i] >= 0) edit = true;
    if (i < n_rows && a.store_slot[i] >= 0) store = true;
  }
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = ief_tc_make_map(&mq, p->dtype, p->q, p->d, p->Nq, p->H, p->B, a.perm_q, kBM)) != IEF_OK) return rc;
  if ((rc = ief_tc_make_map(&mk, p->dtype, p->k, p->d, p->Nk, p->H, p->B, a.perm_k, kNK)) != IEF_OK) return rc;
  if ((rc = ief_tc_make_map(&mv, p->dtype, p->v, p->d, p->Nk, p->H, p->B, a.perm_v, kNK)) != IEF_OK) return rc;
  dim3 grid(ief_ceil_div(p->Nq, kBM), p->H, n_rows);
  return p->dtype == IEF_BF16 ? launch_dtype<IEF_BF16>(mq, mk, mv, a, edit, store, grid, st) : launch_dtype<IEF_F16>(mq, mk, mv, a, edit, store, grid, st);
}
