// Controlled self-attention for sm_100a, second generation: one CTA per SM, TWO 128-row query tiles in flight.
//
//   O[b] = softmax(scale * Q[q_src[b]] K[k_src[b]]^T) V[v_src[b]]        (ief_attn_fwd, include/ief_b200.h)
//
// Why a second kernel: exp throughput (MUFU, 16/clk/SM) bounds a 128x128 score tile at 1024 cycles while its two GEMMs
// need only 384-512 tensor-pipe cycles at head_dim 40-64. The first kernel (attn_tc.cu) serialised QK^T -> softmax -> PV
// per CTA and relied on two co-resident CTAs to overlap; ncu showed them marching in lock-step (XU pipe 45 %, tensor
// pipe 17 %). Here the MMA warp ping-pongs between tiles A and B so that one tile's exponentials always overlap the
// other tile's GEMMs (the schedule FlashAttention-4 uses on this hardware):
//
//   warp 0      TMA producer: Q_A, Q_B once; K and V rings (separate barriers, K is freed after both QK^T, V after both PV)
//   warp 1      tcgen05.mma issuer (warps 2,3 idle: they only exist so the first warpgroup can donate registers)
//   warps 4-7   softmax of tile A (one query row per thread), lazy O rescale, epilogue
//   warps 8-11  softmax of tile B
// setmaxnreg moves registers from the first warpgroup (48/thread) to the softmax warpgroups (224/thread) so that the
// 128 scores of a row live in registers without spilling.
//
// TMEM (512 columns), two layouts:
//   head_dim <= 64  (split P):  S_A 0 | S_B 128 | P_A 256 | P_B 320 | O_A 384 | O_B 448
//       P has its own columns, so the softmax releases S as soon as it is in registers ("consumed") and the MMA warp
//       issues QK(j+1) a whole softmax ahead; a softmax group never waits for its next score tile.
//       order per key tile j:  [cons_A(j)] QK_A(j+1) [cons_B(j)] QK_B(j+1) [P_A(j)] PV_A(j) [P_B(j)] PV_B(j)
//   head_dim 65..128 (aliased): S_A 0 | S_B 128 | O_A 256 | O_B 384, P overwrites the first 64 columns of its S
//       order per key tile j:  [P_A(j)] PV_A(j) QK_A(j+1) [P_B(j)] PV_B(j) QK_B(j+1)
// The softmax reads its 128 scores from TMEM once and keeps them in registers; max uses 3-input FMNMX3, the scale/shift and
// the row sum use packed FFMA2 / FADD2, so the per-element issue cost stays well under the MUFU time.
#include "ief_common.cuh"
#include "ptx_sm100.cuh"
#include "attn_tc_host.cuh"
#include "attn_tc_dev.cuh"
#include <math.h>

using namespace sm100;

// clock64 phase stamps for tools/tc2_trace.py are compiled out by default (their per-tile predicate tests cost ~3 %):
// rebuild with IEF_EXTRA_NVCC_FLAGS="-DIEF_TC2_TRACE=1" to get them back
#ifndef IEF_TC2_TRACE
#define IEF_TC2_TRACE 0
#endif

namespace {

constexpr int kBM = 128;                 // rows per query tile (two tiles per CTA)
constexpr int kBN = 128;                 // keys per KV tile
constexpr int kThreads = 384;
constexpr int kRegsLow = 64, kRegsHigh = 216;  // 128*64 + 256*216 = 63488 <= 65536
constexpr float kRescaleThreshold = 8.0f;

template <int DCH> struct Cfg2 {
  static constexpr int kStages = (DCH == 1) ? 3 : 2;
  static constexpr int kTileBytes = DCH * kTcChunkBytes;
  static constexpr int kSmemData = kTileBytes * (2 + 2 * kStages);
  static constexpr int kSmemBytes = kSmemData + 1024 + 256;
  static constexpr int kTmemCols = 512;
  static constexpr bool kSplitP = (DCH == 1);
  static constexpr int kColP = kSplitP ? 256 : 0;     // + 64 * tile (split) / + 128 * tile (aliased onto S)
  static constexpr int kStrideP = kSplitP ? 64 : 128;
  static constexpr int kColO = kSplitP ? 384 : 256;
  static constexpr int kStrideO = kSplitP ? 64 : 128;
};


template <int DTYPE, int DCH>
__global__ void __launch_bounds__(kThreads, 1)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                const __grid_constant__ TcArgs a) {
  using Cfg = Cfg2<DCH>;
  using E = ElemT<DTYPE>;
  constexpr int ST = Cfg::kStages;

  const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;  // qt indexes 256-row query blocks
  if (!a.rows.active[b]) return;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto sQ = [&](int t) { return base + Cfg::kTileBytes * t; };
  auto sK = [&](int s) { return base + Cfg::kTileBytes * (2 + s); };
  auto sV = [&](int s) { return base + Cfg::kTileBytes * (2 + ST + s); };
  const uint32_t bar0 = base + Cfg::kSmemData;
  const uint32_t bar_q = bar0;
  auto bar_s = [&](int t) { return bar0 + 8 + 8 * t; };    // S_t (or final O_t) complete in TMEM
  auto bar_p = [&](int t) { return bar0 + 24 + 8 * t; };   // P_t written by the 128 softmax threads of tile t
  auto bar_c = [&](int t) { return bar0 + 40 + 8 * t; };   // S_t consumed (in registers) by its 4 softmax warps   [split P]
  auto bar_o = [&](int t) { return bar0 + 56 + 8 * t; };   // PV_t(j) complete: P_t reusable, O_t stable            [split P]
  auto bar_kf = [&](int s) { return bar0 + 72 + 8 * s; };
  auto bar_ke = [&](int s) { return bar0 + 72 + 8 * (ST + s); };
  auto bar_vf = [&](int s) { return bar0 + 72 + 8 * (2 * ST + s); };
  auto bar_ve = [&](int s) { return bar0 + 72 + 8 * (3 * ST + s); };
  const uint32_t tmem_slot = bar0 + 72 + 32 * ST;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = a.nt1 + a.nt2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_q, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_s(t), 1);
      mbar_init(bar_p(t), 128);
      mbar_init(bar_c(t), 4);
      mbar_init(bar_o(t), 1);
    }
    for (int s = 0; s < ST; ++s) {
      mbar_init(bar_kf(s), 1);
      mbar_init(bar_ke(s), 1);
      mbar_init(bar_vf(s), 1);
      mbar_init(bar_ve(s), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp waits, one elected lane issues)
    reg_dec<kRegsLow>();
    const int qb = a.rows.q[b];
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_q, 2 * Cfg::kTileBytes);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int c = 0; c < DCH; ++c) tc_tma_tile(sQ(t) + c * kTcChunkBytes, &tmQ, bar_q, c * 64, (2 * qt + t) * kBM, h, qb, a.perm_q);
    }
    __syncwarp();
    for (int j = 0; j < nt; ++j) {
      const int s = j % ST, ph = (j / ST) & 1;
      const bool blk2 = j >= a.nt1;
      const int jj = blk2 ? j - a.nt1 : j;
      const int kb = blk2 ? a.rows.k2[b] : a.rows.k[b];
      const int vb = blk2 ? a.rows.v2[b] : a.rows.v[b];
      mbar_wait(bar_ke(s), ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_kf(s), Cfg::kTileBytes);
#pragma unroll
        for (int c = 0; c < DCH; ++c) tc_tma_tile(sK(s) + c * kTcChunkBytes, &tmK, bar_kf(s), c * 64, jj * kBN, h, kb, a.perm_k);
      }
      __syncwarp();
      mbar_wait(bar_ve(s), ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_vf(s), Cfg::kTileBytes);
#pragma unroll
        for (int c = 0; c < DCH; ++c) tc_tma_tile(sV(s) + c * kTcChunkBytes, &tmV, bar_vf(s), c * 64, jj * kBN, h, vb, a.perm_v);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp waits, one elected lane issues)
    reg_dec<kRegsLow>();
    // descriptors: the high word (SBO, version, swizzle mode) is constant; per MMA only the 14-bit start address changes
    const uint64_t desc_k = make_smem_desc_sw128(0, 16, 1024);              // K-major operands: Q and K tiles
    const uint64_t desc_v = make_smem_desc_sw128(0, kTcChunkBytes, 1024);   // MN-major operand: V tile
    auto issue_qk = [&](int t, int s) {
      for (int k = 0; k < a.ksteps_qk; ++k) {
        const uint32_t off = (k >> 2) * kTcChunkBytes + (k & 3) * 32;
        umma_ss(tmem_base + 128 * t, desc_k | ((sQ(t) + off) >> 4), desc_k | ((sK(s) + off) >> 4), a.idesc_qk, k > 0);
      }
      umma_commit(bar_s(t));
    };
    auto issue_pv = [&](int t, int s, bool acc) {
#pragma unroll
      for (int k = 0; k < kBN / 16; ++k)
        umma_ts(tmem_base + Cfg::kColO + Cfg::kStrideO * t, tmem_base + Cfg::kColP + Cfg::kStrideP * t + k * 8, desc_v | ((sV(s) + k * 2048) >> 4),
                a.idesc_pv, acc || (k > 0));
    };
    mbar_wait(bar_q, 0);
    mbar_wait(bar_kf(0), 0);
    tc_fence_after();
    if (elect_one()) {
      issue_qk(0, 0);
      issue_qk(1, 0);
      umma_commit(bar_ke(0));
    }
    __syncwarp();
    for (int j = 0; j < nt; ++j) {
      const int s = j % ST, ph = (j / ST) & 1;
      const int s1 = (j + 1) % ST, ph1 = ((j + 1) / ST) & 1;
      const bool more = j + 1 < nt;
#if IEF_TC2_TRACE
      const bool trace = a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j < 64;
#else
      constexpr bool trace = false;
#endif
      if constexpr (Cfg::kSplitP) {
        // next score tiles first: they only need S_t(j) to be in the softmax registers
        if (more) {
          mbar_wait(bar_kf(s1), ph1);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            mbar_wait(bar_c(t), j & 1);
            tc_fence_after();
            if (elect_one()) {
              issue_qk(t, s1);
              if (t == 1) umma_commit(bar_ke(s1));
            }
            __syncwarp();
          }
        }
        mbar_wait(bar_vf(s), ph);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (trace && lane == 0) a.dbg[1024 + (t * 64 + j) * 4 + 0] = clock64();
          mbar_wait(bar_p(t), j & 1);
          tc_fence_after();
          if (trace && lane == 0) a.dbg[1024 + (t * 64 + j) * 4 + 1] = clock64();
          if (elect_one()) {
            issue_pv(t, s, j > 0);
            umma_commit(bar_o(t));
            if (t == 1) umma_commit(bar_ve(s));
          }
          __syncwarp();
          if (trace && lane == 0) a.dbg[1024 + (t * 64 + j) * 4 + 3] = clock64();
        }
      } else {
        mbar_wait(bar_vf(s), ph);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(bar_p(t), j & 1);
          if (more && t == 0) mbar_wait(bar_kf(s1), ph1);
          tc_fence_after();
          if (elect_one()) {
            issue_pv(t, s, j > 0);
            if (t == 1) umma_commit(bar_ve(s));
            if (more) {
              issue_qk(t, s1);   // overwrites S_t / P_t: ordered behind PV_t(j) on the in-order tensor pipe
              if (t == 1) umma_commit(bar_ke(s1));
            } else {
              umma_commit(bar_s(t));  // final O_t complete
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 4) {
    reg_dec<kRegsLow>();  // idle warps of the first warpgroup: setmaxnreg is warpgroup-collective
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue of tile t
    reg_inc<kRegsHigh>();
    const int t = (warp - 4) >> 2;
    const int sub = warp & 3;  // TMEM sub-partition this warp may access (warp id % 4)
    const int row = sub * 32 + lane;
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    const uint32_t tS = tmem_base + 128 * t + lane_off, tP = tmem_base + Cfg::kColP + Cfg::kStrideP * t + lane_off,
                   tO = tmem_base + Cfg::kColO + Cfg::kStrideO * t + lane_off;
    const float c2 = a.scale_log2;
    float m_used = -INFINITY, l = 0.f;
    const int nchunk_o = a.dv_mma >> 4;
    for (int j = 0; j < nt; ++j) {
      const bool blk2 = j >= a.nt1;
      const int jj = blk2 ? j - a.nt1 : j;
      const int vc = min(kBN, a.Nk - jj * kBN);
#if IEF_TC2_TRACE
      const bool trace = a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && row == 0 && j < 64;
      long long* tr = trace ? a.dbg + (t * 64 + j) * 8 : nullptr;
#else
      constexpr bool trace = false;
      long long* tr = nullptr;
#endif
      if (trace) tr[0] = clock64();
      mbar_wait(bar_s(t), j & 1);
      tc_fence_after();
      if (trace) tr[1] = clock64();
      uint32_t s0[32], s1[32], s2[32], s3[32];
      tmem_ld32(tS, s0);
      tmem_ld32(tS + 32, s1);
      tmem_ld32(tS + 64, s2);
      tmem_ld32(tS + 96, s3);
      tc_wait_ld();
      if constexpr (Cfg::kSplitP) {  // S_t(j) is in registers: let the tensor pipe overwrite it with S_t(j+1) right away
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_c(t));
      }
      if (trace) tr[2] = clock64();
      if (vc < kBN) {
        mask_chunk(s0, 0, vc);
        mask_chunk(s1, 32, vc);
        mask_chunk(s2, 64, vc);
        mask_chunk(s3, 96, vc);
      }
      const float tmax = fmaxf(fmaxf(max_chunk(s0, -INFINITY), max_chunk(s1, -INFINITY)), fmaxf(max_chunk(s2, -INFINITY), max_chunk(s3, -INFINITY)));
      // PV_t(j-1) must be complete before P_t is rewritten or O_t rescaled (split-P layout). The wait is deferred
      // until the first chunk of exponentials has been computed, unless a rescale needs O earlier.
      bool o_ready = !Cfg::kSplitP || j == 0;
      if (j == 0) {
        m_used = tmax;
      } else {
        const float m_new = fmaxf(m_used, tmax);
        const bool need = (m_new - m_used) * c2 > kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          if (!o_ready) {
            mbar_wait(bar_o(t), (j - 1) & 1);
            tc_fence_after();
            o_ready = true;
          }
          const float alpha = ief_exp2((m_used - m_new) * c2);
          for (int cc = 0; cc < nchunk_o; ++cc) {
            uint32_t r[16];
            tmem_ld16(tO + 16 * cc, r);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st16(tO + 16 * cc, r);
          }
          l *= alpha;
          m_used = m_new;
        }
      }
      if (trace) tr[3] = clock64();
      const float mc = m_used * c2;
      const float2 c2v = make_float2(c2, c2), nmc = make_float2(-mc, -mc);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
      uint32_t u[16];
      exp_chunk<E>(s0, u, c2v, nmc, acc0, acc1);
      if (!o_ready) {
        mbar_wait(bar_o(t), (j - 1) & 1);
        tc_fence_after();
      }
      tmem_st16(tP, u);
      exp_chunk<E>(s1, u, c2v, nmc, acc0, acc1);
      tmem_st16(tP + 16, u);
      exp_chunk<E>(s2, u, c2v, nmc, acc0, acc1);
      tmem_st16(tP + 32, u);
      exp_chunk<E>(s3, u, c2v, nmc, acc0, acc1);
      tmem_st16(tP + 48, u);
      if (trace) tr[4] = clock64();
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p(t));
      if (trace) tr[5] = clock64();
      acc0 = fadd2(acc0, acc1);
      l += acc0.x + acc0.y;
    }
    // epilogue: O / l -> 16-bit -> global
    if constexpr (Cfg::kSplitP) mbar_wait(bar_o(t), (nt - 1) & 1);
    else mbar_wait(bar_s(t), nt & 1);
    tc_fence_after();
    const float inv = 1.f / l;
    const int grow = (2 * qt + t) * kBM + row;
    if (a.lse_out != nullptr && grow < a.Nq) a.lse_out[((int64_t)b * a.H + h) * a.Nq + grow] = fmaf(m_used, c2, log2f(l));
    typename E::T* op = reinterpret_cast<typename E::T*>(a.o) + (int64_t)b * a.o_sb + (int64_t)grow * a.o_sn + (int64_t)h * a.o_sh;
    const int nchunk_d = (a.d + 15) >> 4;
    for (int cc = 0; cc < nchunk_d; ++cc) {
      uint32_t r[16];
      tmem_ld16(tO + 16 * cc, r);
      tc_wait_ld();
      if (grow < a.Nq) {
        uint4 v0, v1;
        v0.x = E::pack(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
        v0.y = E::pack(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
        v0.z = E::pack(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
        v0.w = E::pack(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
        v1.x = E::pack(__uint_as_float(r[8]) * inv, __uint_as_float(r[9]) * inv);
        v1.y = E::pack(__uint_as_float(r[10]) * inv, __uint_as_float(r[11]) * inv);
        v1.z = E::pack(__uint_as_float(r[12]) * inv, __uint_as_float(r[13]) * inv);
        v1.w = E::pack(__uint_as_float(r[14]) * inv, __uint_as_float(r[15]) * inv);
        if (16 * cc + 8 <= a.d) *reinterpret_cast<uint4*>(op + 16 * cc) = v0;
        if (16 * cc + 16 <= a.d) *reinterpret_cast<uint4*>(op + 16 * cc + 8) = v1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int DTYPE, int DCH>
int launch_tc2(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const TcArgs& a, dim3 grid, cudaStream_t st) {
  auto kern = attn_tc2_kernel<DTYPE, DCH>;
  IEF_CONFIG_SMEM(kern, Cfg2<DCH>::kSmemBytes);
  kern<<<grid, kThreads, Cfg2<DCH>::kSmemBytes, st>>>(mq, mk, mv, a);
  IEF_LAUNCH_OK("attn_tc2_kernel");
  return IEF_OK;
}

}  // namespace

int ief_attn_tc2_launch(const ief_attn_params* p, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const TcArgs& a, cudaStream_t st) {
  const int dch = ief_ceil_div(p->d, 64);
  dim3 grid(ief_ceil_div(p->Nq, 2 * kBM), p->H, p->B);
  if (p->dtype == IEF_BF16) return dch == 1 ? launch_tc2<IEF_BF16, 1>(mq, mk, mv, a, grid, st) : launch_tc2<IEF_BF16, 2>(mq, mk, mv, a, grid, st);
  return dch == 1 ? launch_tc2<IEF_F16, 1>(mq, mk, mv, a, grid, st) : launch_tc2<IEF_F16, 2>(mq, mk, mv, a, grid, st);
}
