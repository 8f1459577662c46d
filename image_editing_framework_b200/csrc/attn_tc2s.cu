// Controlled self-attention for sm_100a, second generation, SPLIT-KV flavour (head_dim <= 64).
//
//   O[b] = softmax(scale * Q[q_src[b]] K[k_src[b]]^T) V[v_src[b]]        (ief_attn_fwd, include/ief_b200.h)
//
// Same machinery as attn_tc2.cu (split S | P | O TMEM layout, early QK issue, elected-lane MMA issue, setmaxnreg), but the
// two softmax warpgroups of a CTA work on the SAME 128 query rows and split the key range in two halves; their partial
// (O, max, sum) are merged through shared memory at the end. A CTA is therefore one 128-row tile instead of 256 rows:
// twice as many, half as long CTAs. That is the cure for wave quantisation on the headline shape — SD-1.5's 64x64 layer
// at B=4 is 512 CTAs of 256 rows = 3.46 waves on 148 SMs (86 % occupancy of the last wave) but 1024 CTAs of 128 rows =
// 6.92 waves (98.8 %). The price is K/V traffic from L2 (both halves stream their own K/V; Q is shared instead).
// The launcher picks this flavour when it improves the wave efficiency (attn_tc.cu).
#include "ief_common.cuh"
#include "ptx_sm100.cuh"
#include "attn_tc_host.cuh"
#include "attn_tc_dev.cuh"
#include <math.h>

using namespace sm100;

namespace {

constexpr int kBM = 128, kBN = 128;
constexpr int kThreads = 384;
constexpr int kRegsLow = 64, kRegsHigh = 216;  // 128*64 + 256*216 = 63488 <= 65536 (pool at launch: 384 x 168)
constexpr int ST = 3;
constexpr int kTile = kTcChunkBytes;                       // 16 KiB: one [128 x 64ch] box (head_dim <= 64)
constexpr int kSmemData = kTile * (1 + 4 * ST);            // Q | K ring [ST][2] | V ring [ST][2]
constexpr int kSmemBytes = kSmemData + 1024 + 256;
constexpr float kRescaleThreshold = 8.0f;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <int DTYPE>
__global__ void __launch_bounds__(kThreads, 1)
attn_tc2s_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ TcArgs a) {
  using E = ElemT<DTYPE>;
  const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;  // qt indexes 128-row query tiles
  if (!a.rows.active[b]) return;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base;
  auto sK = [&](int s, int t) { return base + kTile * (1 + 2 * s + t); };
  auto sV = [&](int s, int t) { return base + kTile * (1 + 2 * ST + 2 * s + t); };
  float* stage_o = reinterpret_cast<float*>(base_ptr + kTile);            // merge staging, reuses K ring stage 0 (32 KiB): [col][row]
  float* stage_ml = reinterpret_cast<float*>(base_ptr + kTile * 3);       // K ring stage 1: [2][128] (max, sum)
  const uint32_t bar0 = base + kSmemData;
  const uint32_t bar_q = bar0;
  auto bar_s = [&](int t) { return bar0 + 8 + 8 * t; };
  auto bar_p = [&](int t) { return bar0 + 24 + 8 * t; };
  auto bar_c = [&](int t) { return bar0 + 40 + 8 * t; };
  auto bar_o = [&](int t) { return bar0 + 56 + 8 * t; };
  auto bar_kf = [&](int s) { return bar0 + 72 + 8 * s; };
  auto bar_ke = [&](int s) { return bar0 + 72 + 8 * (ST + s); };
  auto bar_vf = [&](int s) { return bar0 + 72 + 8 * (2 * ST + s); };
  auto bar_ve = [&](int s) { return bar0 + 72 + 8 * (3 * ST + s); };
  const uint32_t tmem_slot = bar0 + 72 + 32 * ST;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = a.nt1 + a.nt2;
  const int ntA = (nt + 1) >> 1, ntB = nt - ntA;  // key tiles [0, ntA) -> warpgroup A, [ntA, nt) -> warpgroup B

  // global key-tile index -> (row of the source tables, tile inside that key block)
  auto kv_coord = [&](int g, int& jj, int& kb, int& vb) {
    const bool blk2 = g >= a.nt1;
    jj = blk2 ? g - a.nt1 : g;
    kb = blk2 ? a.rows.k2[b] : a.rows.k[b];
    vb = blk2 ? a.rows.v2[b] : a.rows.v[b];
  };

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
      mbar_arrive_expect_tx(bar_q, kTile);
      tc_tma_tile(sQ, &tmQ, bar_q, 0, qt * kBM, h, qb, a.perm_q);
    }
    __syncwarp();
    for (int j = 0; j < ntA; ++j) {
      const int s = j % ST, ph = (j / ST) & 1;
      const int nload = j < ntB ? 2 : 1;
      mbar_wait(bar_ke(s), ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_kf(s), nload * kTile);
        for (int t = 0; t < nload; ++t) {
          int jj, kb, vb;
          kv_coord(t == 0 ? j : ntA + j, jj, kb, vb);
          tc_tma_tile(sK(s, t), &tmK, bar_kf(s), 0, jj * kBN, h, kb, a.perm_k);
        }
      }
      __syncwarp();
      mbar_wait(bar_ve(s), ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_vf(s), nload * kTile);
        for (int t = 0; t < nload; ++t) {
          int jj, kb, vb;
          kv_coord(t == 0 ? j : ntA + j, jj, kb, vb);
          tc_tma_tile(sV(s, t), &tmV, bar_vf(s), 0, jj * kBN, h, vb, a.perm_v);
        }
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
        umma_ss(tmem_base + 128 * t, desc_k | ((sQ + k * 32) >> 4), desc_k | ((sK(s, t) + k * 32) >> 4), a.idesc_qk, k > 0);
      umma_commit(bar_s(t));
    };
    auto issue_pv = [&](int t, int s, bool acc) {
#pragma unroll
      for (int k = 0; k < kBN / 16; ++k)
        umma_ts(tmem_base + 384 + 64 * t, tmem_base + 256 + 64 * t + k * 8, desc_v | ((sV(s, t) + k * 2048) >> 4), a.idesc_pv, acc || (k > 0));
    };
    mbar_wait(bar_q, 0);
    mbar_wait(bar_kf(0), 0);
    tc_fence_after();
    if (elect_one()) {
      issue_qk(0, 0);
      if (ntB > 0) issue_qk(1, 0);
      umma_commit(bar_ke(0));
    }
    __syncwarp();
    for (int j = 0; j < ntA; ++j) {
      const int s = j % ST, ph = (j / ST) & 1;
      const int s1 = (j + 1) % ST, ph1 = ((j + 1) / ST) & 1;
      if (j + 1 < ntA) {
        mbar_wait(bar_kf(s1), ph1);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (j + 1 < (t == 0 ? ntA : ntB)) {
            mbar_wait(bar_c(t), j & 1);
            tc_fence_after();
            if (elect_one()) issue_qk(t, s1);
            __syncwarp();
          }
        }
        if (elect_one()) umma_commit(bar_ke(s1));
        __syncwarp();
      }
      mbar_wait(bar_vf(s), ph);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (j < (t == 0 ? ntA : ntB)) {
          mbar_wait(bar_p(t), j & 1);
          tc_fence_after();
          if (elect_one()) {
            issue_pv(t, s, j > 0);
            umma_commit(bar_o(t));
          }
          __syncwarp();
        }
      }
      if (elect_one()) umma_commit(bar_ve(s));
      __syncwarp();
    }
  } else if (warp < 4) {
    reg_dec<kRegsLow>();
  } else {
    // ------------------------------------------------------------------ softmax of key half t, then merge
    reg_inc<kRegsHigh>();
    const int t = (warp - 4) >> 2;
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    const uint32_t tS = tmem_base + 128 * t + lane_off, tP = tmem_base + 256 + 64 * t + lane_off, tO = tmem_base + 384 + 64 * t + lane_off;
    const float c2 = a.scale_log2;
    float m_used = -INFINITY, l = 0.f;
    const int nchunk_o = a.dv_mma >> 4;
    const int my_nt = t == 0 ? ntA : ntB;

    for (int j = 0; j < my_nt; ++j) {
      int jj, kb_unused, vb_unused;
      kv_coord(t == 0 ? j : ntA + j, jj, kb_unused, vb_unused);
      const int vc = min(kBN, a.Nk - jj * kBN);
      mbar_wait(bar_s(t), j & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32], s2[32], s3[32];
      tmem_ld32(tS, s0);
      tmem_ld32(tS + 32, s1);
      tmem_ld32(tS + 64, s2);
      tmem_ld32(tS + 96, s3);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_c(t));
      if (vc < kBN) {
        mask_chunk(s0, 0, vc);
        mask_chunk(s1, 32, vc);
        mask_chunk(s2, 64, vc);
        mask_chunk(s3, 96, vc);
      }
      const float tmax = fmaxf(fmaxf(max_chunk(s0, -INFINITY), max_chunk(s1, -INFINITY)), fmaxf(max_chunk(s2, -INFINITY), max_chunk(s3, -INFINITY)));
      bool o_ready = j == 0;
      if (j == 0) {
        m_used = tmax;
      } else {
        const float m_new = fmaxf(m_used, tmax);
        const bool need = (m_new - m_used) * c2 > kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(bar_o(t), (j - 1) & 1);
          tc_fence_after();
          o_ready = true;
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
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p(t));
      acc0 = fadd2(acc0, acc1);
      l += acc0.x + acc0.y;
    }
    if (my_nt > 0) {
      mbar_wait(bar_o(t), (my_nt - 1) & 1);
      tc_fence_after();
    }
    // ---- merge the two key halves. When warpgroup B's last PV has completed, every QK MMA of the CTA has completed too
    // (program order of the issuing thread), so the K ring is free to serve as staging; the V ring may still be in use.
    const int nchunk_d = (a.d + 15) >> 4;
    if (t == 1 && ntB > 0) {
      stage_ml[row] = m_used;
      stage_ml[128 + row] = l;
      for (int cc = 0; cc < nchunk_d; ++cc) {
        uint32_t r[16];
        tmem_ld16(tO + 16 * cc, r);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) stage_o[(16 * cc + i) * 128 + row] = __uint_as_float(r[i]);
      }
    }
    named_bar_sync(1, 256);
    if (t == 0) {
      float wA = 1.f, wB = 0.f, lsum = l;
      if (ntB > 0) {
        const float mB = stage_ml[row], lB = stage_ml[128 + row];
        const float m = fmaxf(m_used, mB);
        wA = ief_exp2((m_used - m) * c2);
        wB = ief_exp2((mB - m) * c2);
        lsum = l * wA + lB * wB;
      }
      const float inv = 1.f / lsum;
      wA *= inv;
      wB *= inv;
      const int grow = qt * kBM + row;
      typename E::T* op = reinterpret_cast<typename E::T*>(a.o) + (int64_t)b * a.o_sb + (int64_t)grow * a.o_sn + (int64_t)h * a.o_sh;
      for (int cc = 0; cc < nchunk_d; ++cc) {
        uint32_t r[16];
        tmem_ld16(tO + 16 * cc, r);
        tc_wait_ld();
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          f[i] = __uint_as_float(r[i]) * wA;
          if (ntB > 0) f[i] = fmaf(stage_o[(16 * cc + i) * 128 + row], wB, f[i]);
        }
        if (grow < a.Nq) {
          uint4 v0, v1;
          v0.x = E::pack(f[0], f[1]); v0.y = E::pack(f[2], f[3]); v0.z = E::pack(f[4], f[5]); v0.w = E::pack(f[6], f[7]);
          v1.x = E::pack(f[8], f[9]); v1.y = E::pack(f[10], f[11]); v1.z = E::pack(f[12], f[13]); v1.w = E::pack(f[14], f[15]);
          if (16 * cc + 8 <= a.d) *reinterpret_cast<uint4*>(op + 16 * cc) = v0;
          if (16 * cc + 16 <= a.d) *reinterpret_cast<uint4*>(op + 16 * cc + 8) = v1;
        }
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
int launch_tc2s(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const TcArgs& a, dim3 grid, cudaStream_t st) {
  auto kern = attn_tc2s_kernel<DTYPE>;
  static bool configured = false;
  if (!configured) {
    IEF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  kern<<<grid, kThreads, kSmemBytes, st>>>(mq, mk, mv, a);
  IEF_LAUNCH_OK("attn_tc2s_kernel");
  return IEF_OK;
}

}  // namespace

int ief_attn_tc2s_launch(const ief_attn_params* p, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const TcArgs& a, cudaStream_t st) {
  dim3 grid(ief_ceil_div(p->Nq, kBM), p->H, p->B);
  return p->dtype == IEF_BF16 ? launch_tc2s<IEF_BF16>(mq, mk, mv, a, grid, st) : launch_tc2s<IEF_F16>(mq, mk, mv, a, grid, st);
}
