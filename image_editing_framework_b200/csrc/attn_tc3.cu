// Controlled self-attention for sm_100a, third generation (head_dim <= 64): the second-generation schedule (attn_tc2.cu,
// two 128-row query tiles ping-ponging through a split S | P | O TMEM layout) with each tile's softmax spread over TWO
// warpgroups that own the left / right 64 score columns of every row.
//
//   O[b] = softmax(scale * Q[q_src[b]] K[k_src[b]]^T) V[v_src[b]]        (ief_attn_fwd, include/ief_b200.h)
//
// Why: the exp pipe (MUFU, 8 cycles per warp instruction per SM sub-partition) is the binding unit at head_dim 40-64.
// Measured (tools/micro/mufu.cu): one softmax warp per sub-partition keeps it 84 % busy, two keep it 98 % busy. With one
// warpgroup per tile only one warp per sub-partition is in its exp phase most of the time; with two, whichever tile is
// in its exp phase saturates the pipe on its own while the other tile loads / reduces / synchronises.
//
//   warp 0        TMA producer            warp 1   tcgen05.mma issuer       warps 2,3  idle (register donors)
//   warps 4-11    tile A: columns 0-63 (warps 4-7) and 64-127 (warps 8-11) of every score row
//   warps 12-19   tile B likewise
// The two halves of a row exchange their partial row maximum through shared memory (one 256-thread named barrier per
// key tile) so both scale with the same reference; partial row sums are merged once in the epilogue. Each half rescales
// and writes back its share of the O columns.
#include "ief_common.cuh"
#include "ptx_sm100.cuh"
#include "attn_tc_host.cuh"
#include "attn_tc_dev.cuh"
#include <math.h>

using namespace sm100;

namespace {

constexpr int kBM = 128, kBN = 128;
constexpr int kThreads = 640;
constexpr int kRegsLow = 32, kRegsHigh = 112;  // pool = 640 threads x 96 regs at launch (61440): 128*32 + 512*112 = 61440
constexpr int kStages = 3;
constexpr int kTileBytes = kTcChunkBytes;                    // head_dim <= 64: one 64-channel chunk per tile
constexpr int kSmemData = kTileBytes * (2 + 2 * kStages);    // Q_A Q_B | K ring | V ring
constexpr int kSmemXchg = 6 * 1024;                          // row-max exchange [2 parities][2 tiles][2 halves][128] + row-sum [2][2][128]
constexpr int kSmemBytes = kSmemData + kSmemXchg + 1024 + 256;
constexpr float kRescaleThreshold = 8.0f;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <int DTYPE>
__global__ void __launch_bounds__(kThreads, 1)
attn_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                const __grid_constant__ TcArgs a) {
  using E = ElemT<DTYPE>;
  constexpr int ST = kStages;
  const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
  if (!a.rows.active[b]) return;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  auto sQ = [&](int t) { return base + kTileBytes * t; };
  auto sK = [&](int s) { return base + kTileBytes * (2 + s); };
  auto sV = [&](int s) { return base + kTileBytes * (2 + ST + s); };
  float* xmax = reinterpret_cast<float*>(base_ptr + kSmemData);         // [parity][tile][half][128]
  float* xsum = xmax + 2 * 2 * 2 * 128;                                 // [tile][half][128]
  const uint32_t bar0 = base + kSmemData + kSmemXchg;
  const uint32_t bar_q = bar0;
  auto bar_s = [&](int t) { return bar0 + 8 + 8 * t; };    // S_t complete in TMEM
  auto bar_p = [&](int t) { return bar0 + 24 + 8 * t; };   // P_t written by the 256 softmax threads of tile t
  auto bar_c = [&](int t) { return bar0 + 40 + 8 * t; };   // S_t in registers of its 8 softmax warps
  auto bar_o = [&](int t) { return bar0 + 56 + 8 * t; };   // PV_t(j) complete
  auto bar_x = [&](int t) { return bar0 + 72 + 8 * t; };   // exp turn of tile t (granted by the 8 warps of the other tile)
  auto bar_kf = [&](int s) { return bar0 + 88 + 8 * s; };
  auto bar_ke = [&](int s) { return bar0 + 88 + 8 * (ST + s); };
  auto bar_vf = [&](int s) { return bar0 + 88 + 8 * (2 * ST + s); };
  auto bar_ve = [&](int s) { return bar0 + 88 + 8 * (3 * ST + s); };
  const uint32_t tmem_slot = bar0 + 88 + 32 * ST;
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
      mbar_init(bar_p(t), 256);
      mbar_init(bar_c(t), 8);
      mbar_init(bar_o(t), 1);
      mbar_init(bar_x(t), 8);
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
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // TMEM columns: S_A 0 | S_B 128 | P_A 256 | P_B 320 | O_A 384 | O_B 448

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    reg_dec<kRegsLow>();
    const int qb = a.rows.q[b];
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_q, 2 * kTileBytes);
      tc_tma_tile(sQ(0), &tmQ, bar_q, 0, (2 * qt) * kBM, h, qb, a.perm_q);
      tc_tma_tile(sQ(1), &tmQ, bar_q, 0, (2 * qt + 1) * kBM, h, qb, a.perm_q);
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
        mbar_arrive_expect_tx(bar_kf(s), kTileBytes);
        tc_tma_tile(sK(s), &tmK, bar_kf(s), 0, jj * kBN, h, kb, a.perm_k);
      }
      __syncwarp();
      mbar_wait(bar_ve(s), ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_vf(s), kTileBytes);
        tc_tma_tile(sV(s), &tmV, bar_vf(s), 0, jj * kBN, h, vb, a.perm_v);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp waits, one elected lane issues)
    reg_dec<kRegsLow>();
    const uint64_t desc_k = make_smem_desc_sw128(0, 16, 1024);
    const uint64_t desc_v = make_smem_desc_sw128(0, kTcChunkBytes, 1024);
    auto issue_qk = [&](int t, int s) {
      for (int k = 0; k < a.ksteps_qk; ++k)
        umma_ss(tmem_base + 128 * t, desc_k | ((sQ(t) + k * 32) >> 4), desc_k | ((sK(s) + k * 32) >> 4), a.idesc_qk, k > 0);
      umma_commit(bar_s(t));
    };
    auto issue_pv = [&](int t, int s, bool acc) {
#pragma unroll
      for (int k = 0; k < kBN / 16; ++k)
        umma_ts(tmem_base + 384 + 64 * t, tmem_base + 256 + 64 * t + k * 8, desc_v | ((sV(s) + k * 2048) >> 4), a.idesc_pv, acc || (k > 0));
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
      if (j + 1 < nt) {  // next score tiles first: they only need S_t(j) to be in the softmax registers
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
        mbar_wait(bar_p(t), j & 1);
        tc_fence_after();
        if (elect_one()) {
          issue_pv(t, s, j > 0);
          umma_commit(bar_o(t));
          if (t == 1) umma_commit(bar_ve(s));
        }
        __syncwarp();
      }
    }
  } else if (warp < 4) {
    reg_dec<kRegsLow>();
  } else {
    // ------------------------------------------------------------------ softmax of tile t, column half `half`
    reg_inc<kRegsHigh>();
    const int idx = warp - 4;
    const int t = idx >> 3, half = (idx >> 2) & 1;
    const int sub = warp & 3;  // TMEM sub-partition of this warp
    const int row = sub * 32 + lane;
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    const uint32_t tS = tmem_base + 128 * t + 64 * half + lane_off;
    const uint32_t tP = tmem_base + 256 + 64 * t + 32 * half + lane_off;
    const uint32_t tO = tmem_base + 384 + 64 * t + lane_off;
    const float c2 = a.scale_log2;
    float m_used = -INFINITY, l = 0.f;
    const int nchunk_o = a.dv_mma >> 4;
    for (int j = 0; j < nt; ++j) {
      const bool blk2 = j >= a.nt1;
      const int jj = blk2 ? j - a.nt1 : j;
      const int vc = min(kBN, a.Nk - jj * kBN);
      const bool trace = a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && row == 0 && half == 0 && j < 64;
      long long* tr = trace ? a.dbg + (t * 64 + j) * 8 : nullptr;
      if (trace) tr[0] = clock64();
      mbar_wait(bar_s(t), j & 1);
      tc_fence_after();
      if (trace) tr[1] = clock64();
      uint32_t s0[32], s1[32];
      tmem_ld32(tS, s0);
      tmem_ld32(tS + 32, s1);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_c(t));  // S_t(j) is in registers: the tensor pipe may overwrite it with S_t(j+1)
      if (trace) tr[2] = clock64();
      if (vc < kBN) {
        mask_chunk(s0, 64 * half, vc);
        mask_chunk(s1, 64 * half + 32, vc);
      }
      // row maximum: own 64 columns, then exchange with the other half of the row
      const float lmax = fmaxf(max_chunk(s0, -INFINITY), max_chunk(s1, -INFINITY));
      float* xm = xmax + ((j & 1) * 4 + t * 2) * 128;
      xm[half * 128 + row] = lmax;
      named_bar_sync(1 + t, 256);
      const float tmax = fmaxf(lmax, xm[(half ^ 1) * 128 + row]);
      bool o_ready = j == 0;
      if (j == 0) {
        m_used = tmax;
      } else {
        const float m_new = fmaxf(m_used, tmax);
        const bool need = (m_new - m_used) * c2 > kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {  // both halves of a row see identical values and take the same decision
          mbar_wait(bar_o(t), (j - 1) & 1);
          tc_fence_after();
          o_ready = true;
          const float alpha = ief_exp2((m_used - m_new) * c2);
          for (int cc = half; cc < nchunk_o; cc += 2) {  // this half's share of the O columns
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
      // Ordered exp sections (the ping-pong of FlashAttention-3/4): tile A's and tile B's exponentials strictly alternate,
      // so each runs on an uncontended MUFU while the other tile loads / reduces / synchronises. Left to themselves the two
      // tiles fall into lock-step (both idle the MUFU during their non-exp phases, then share it).
      const float mc = m_used * c2;
      const float2 c2v = make_float2(c2, c2), nmc = make_float2(-mc, -mc);
      scale_chunk(s0, c2v, nmc);   // FMA pipe work, outside the MUFU turn
      scale_chunk(s1, c2v, nmc);
      if (!o_ready) {  // PV_t(j-1) must have finished reading P_t; wait for it BEFORE taking the exp turn, not inside it
        mbar_wait(bar_o(t), (j - 1) & 1);
        tc_fence_after();
        o_ready = true;
      }
      // Ordered exp sections (the ping-pong of FlashAttention-3/4): tile A's and tile B's exponentials strictly alternate,
      // so each runs on an uncontended MUFU while the other tile loads / reduces / synchronises.
      if (a.skew_cycles != 0 && (t == 1 || j > 0)) mbar_wait(bar_x(t), (t == 1 ? j : j - 1) & 1);
      if (trace) tr[3] = clock64();
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
      uint32_t u[16];
      exp_pack_chunk<E>(s0, u, acc0, acc1);
      tmem_st16(tP, u);
      if (a.skew_cycles != 0) {  // hand the MUFU over one chunk early: the other tile's wake-up overlaps our last 32 columns
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_x(t ^ 1));
      }
      exp_pack_chunk<E>(s1, u, acc0, acc1);
      tmem_st16(tP + 16, u);
      if (trace) tr[4] = clock64();
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p(t));
      if (trace) tr[5] = clock64();
      acc0 = fadd2(acc0, acc1);
      l += acc0.x + acc0.y;
    }
    // epilogue: merge the two halves' row sums, then each half writes its share of O / l
    mbar_wait(bar_o(t), (nt - 1) & 1);
    tc_fence_after();
    float* xs = xsum + t * 2 * 128;
    xs[half * 128 + row] = l;
    named_bar_sync(1 + t, 256);
    const float inv = 1.f / (l + xs[(half ^ 1) * 128 + row]);
    const int grow = (2 * qt + t) * kBM + row;
    typename E::T* op = reinterpret_cast<typename E::T*>(a.o) + (int64_t)b * a.o_sb + (int64_t)grow * a.o_sn + (int64_t)h * a.o_sh;
    const int nchunk_d = (a.d + 15) >> 4;
    for (int cc = half; cc < nchunk_d; cc += 2) {
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
    tmem_dealloc(tmem_base, 512);
  }
}

template <int DTYPE>
int launch_tc3(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const TcArgs& a, dim3 grid, cudaStream_t st) {
  auto kern = attn_tc3_kernel<DTYPE>;
  static bool configured = false;
  if (!configured) {
    IEF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  kern<<<grid, kThreads, kSmemBytes, st>>>(mq, mk, mv, a);
  IEF_LAUNCH_OK("attn_tc3_kernel");
  return IEF_OK;
}

}  // namespace

int ief_attn_tc3_launch(const ief_attn_params* p, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const TcArgs& a, cudaStream_t st) {
  dim3 grid(ief_ceil_div(p->Nq, 2 * kBM), p->H, p->B);
  return p->dtype == IEF_BF16 ? launch_tc3<IEF_BF16>(mq, mk, mv, a, grid, st) : launch_tc3<IEF_F16>(mq, mk, mv, a, grid, st);
}
