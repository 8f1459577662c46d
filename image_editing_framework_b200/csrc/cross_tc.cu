// 77-key cross-attention of PLAIN rows (no probability edit, no map output) on the tensor pipe.
//   O[b] = softmax(scale * Q[b] K[b]^T) V[b],  Nk <= 80        (p2p/model/register.py:43-51 on an un-edited layer call;
//   masactrl/model/register.py:35-47, pnp/model/register.py:65-76, pix2pix-zero attention_control.py:43-49 for cross layers)
//
// The mma.sync kernel (cross_attn.cu) spreads a score row over a quad and pays ~56 warp instructions per query row; here a
// thread owns a whole row (TMEM lane), S = QK^T and O = PV are two tcgen05.mma groups, and a row costs ~8 warp instructions:
// the kernel moves from instruction-issue-bound towards the HBM roofline it nominally has (AI ~ 76 flop/B).
//   warp 4      TMA (Q 128 x d, K/V 80 x d, SWIZZLE_128B, zero-filled out of bounds) and the MMA issue (one elected lane)
//   warps 0-3   softmax of 128 rows: S (80 fp32 columns) -> registers, exp2, P (16-bit pairs) written over S's first 40 columns
//               (a row is read and rewritten by the same thread, so the alias needs no synchronisation), epilogue O / l
// TMEM: S/P columns [0, 80), O columns [80, 80 + dv): 128 columns for head_dim <= 48 (4 CTAs per SM), else 256.
#include "ief_common.cuh"
#include "ptx_sm100.cuh"
#include "attn_tc_host.cuh"
#include <math.h>
#include <stdlib.h>

using namespace sm100;

namespace {

constexpr int kBM = 128, kNK = 80, kThreads = 160;
constexpr int kQChunk = 128 * 128, kKVChunk = kNK * 128;   // bytes per 64-channel chunk
constexpr float kLog2e = 1.4426950408889634f;

struct CrossTcArgs {
  void* o;
  int64_t o_sb, o_sn, o_sh;
  int32_t B, H, Nq, Nk, d, ksteps_qk, dv_mma;
  int32_t tiles_per_cta;   // consecutive 128-row query tiles of one (row, head) handled by a CTA (K/V loaded once, Q double-buffered)
  uint32_t idesc_qk, idesc_pv;
  float scale_log2;
  int32_t perm_q[3], perm_k[3], perm_v[3];
  int32_t row[IEF_MAX_ROWS];   // blockIdx.z -> batch row (a call may hand its edited / stored rows to cross_tc_edit.cu)
  int32_t pdl;                 // launched as a programmatic dependent of the edit kernel's launch of the same call
};

template <int DCH> constexpr int cross_tc_smem() { return DCH * (2 * kQChunk + 2 * kKVChunk) + 1024 + 128; }

template <int DTYPE, int DCH, int TCOLS>
__global__ void __launch_bounds__(kThreads)
cross_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                const __grid_constant__ CrossTcArgs a) {
  using E = ElemT<DTYPE>;
  constexpr int kTmemCols = TCOLS;   // 128 when 80 + dv <= 128 (head_dim <= 48), else 256
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // nothing of this grid is consumed by a dependent launch
  const int b = a.row[blockIdx.z], h = blockIdx.y;
  const int nqt = (a.Nq + kBM - 1) / kBM;
  const int qt0 = blockIdx.x * a.tiles_per_cta, ntile = min(a.tiles_per_cta, nqt - qt0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto sQ = [&](int buf) { return base + buf * DCH * kQChunk; };
  const uint32_t sK = base + 2 * DCH * kQChunk, sV = sK + DCH * kKVChunk;
  const uint32_t bar0 = sV + DCH * kKVChunk;
  const uint32_t bar_kv = bar0, bar_s = bar0 + 8, bar_p = bar0 + 16, bar_o = bar0 + 24, bar_oe = bar0 + 32;
  auto bar_q = [&](int buf) { return bar0 + 40 + 8 * buf; };
  const uint32_t tmem_slot = bar0 + 56;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_kv, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    mbar_init(bar_oe, 128);
    mbar_init(bar_q(0), 1);
    mbar_init(bar_q(1), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + kNK;

  if (warp == 4) {
    auto load_q = [&](int i) {  // query tile i of this CTA -> buffer i & 1
      mbar_arrive_expect_tx(bar_q(i & 1), DCH * kQChunk);
#pragma unroll
      for (int c = 0; c < DCH; ++c) tc_tma_tile(sQ(i & 1) + c * kQChunk, &tmQ, bar_q(i & 1), c * 64, (qt0 + i) * kBM, h, b, a.perm_q);
    };
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_kv, 2 * DCH * kKVChunk);
#pragma unroll
      for (int c = 0; c < DCH; ++c) {
        tc_tma_tile(sK + c * kKVChunk, &tmK, bar_kv, c * 64, 0, h, b, a.perm_k);
        tc_tma_tile(sV + c * kKVChunk, &tmV, bar_kv, c * 64, 0, h, b, a.perm_v);
      }
      load_q(0);
    }
    __syncwarp();
    mbar_wait(bar_kv, 0);
    for (int i = 0; i < ntile; ++i) {
      // Q(i+1) goes into the buffer QK(i-1) read; that MMA group has completed (its scores were consumed before P(i-1) arrived)
      if (i + 1 < ntile && elect_one()) load_q(i + 1);
      __syncwarp();
      mbar_wait(bar_q(i & 1), (i >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        // S[128 x 80] = Q K^T : both operands K-major, contraction over head_dim in steps of 16. Issued after PV(i-1), which
        // reads P(i-1) from the columns S(i) overwrites: MMAs of one thread execute in order
        for (int k = 0; k < a.ksteps_qk; ++k)
          umma_ss(tmem_S, make_smem_desc_sw128(sQ(i & 1) + (k >> 2) * kQChunk + (k & 3) * 32, 16, 1024),
                  make_smem_desc_sw128(sK + (k >> 2) * kKVChunk + (k & 3) * 32, 16, 1024), a.idesc_qk, k > 0);
        umma_commit(bar_s);
      }
      __syncwarp();
      mbar_wait(bar_p, i & 1);
      if (i > 0) mbar_wait(bar_oe, (i - 1) & 1);  // the epilogue of tile i-1 has read its O out of TMEM
      tc_fence_after();
      if (elect_one()) {
        // O[128 x dv] = P V : A = P from TMEM (8 columns per 16 keys), B = V (MN-major: rows = keys, 64-channel chunks kKVChunk apart)
#pragma unroll
        for (int k = 0; k < kNK / 16; ++k)
          umma_ts(tmem_O, tmem_S + k * 8, make_smem_desc_sw128(sV + k * 2048, kKVChunk, 1024), a.idesc_pv, k > 0);
        umma_commit(bar_o);
      }
      __syncwarp();
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const uint32_t tS = tmem_S + lane_off, tO = tmem_O + lane_off;
    const float c2 = a.scale_log2;
    const int nk = a.Nk;
    const int nchunk_d = (a.d + 15) >> 4;
    for (int i = 0; i < ntile; ++i) {
      mbar_wait(bar_s, i & 1);
      tc_fence_after();
      // two passes over the score row in TMEM (maximum, then exp2 + pack) keep the live registers low: the kernel is latency
      // bound, so what matters is how many CTAs fit on an SM (TMEM: 4 at 128 columns)
      float m = -INFINITY;
      {
        uint32_t r[16];
#pragma unroll 1
        for (int c = 0; c < kNK / 16; ++c) {
          tmem_ld16(tS + 16 * c, r);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (16 * c + e < nk) m = fmaxf(m, __uint_as_float(r[e]));
        }
      }
      const float mc = m * c2;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll 1
      for (int c = 0; c < kNK / 16; ++c) {
        uint32_t r[16], u8[8];
        tmem_ld16(tS + 16 * c, r);
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float p0 = ief_exp2(fmaf(__uint_as_float(r[2 * e]), c2, -mc)), p1 = ief_exp2(fmaf(__uint_as_float(r[2 * e + 1]), c2, -mc));
          if (16 * c + 2 * e >= nk) p0 = 0.f;
          if (16 * c + 2 * e + 1 >= nk) p1 = 0.f;
          l0 += p0; l1 += p1;
          u8[e] = E::pack(p0, p1);
        }
        // P columns [8c, 8c+8) overwrite score columns this thread has already consumed (score chunk c/2 <= c)
        tmem_st8(tS + 8 * c, u8);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p);
      const float inv = 1.f / (l0 + l1);
      mbar_wait(bar_o, i & 1);
      tc_fence_after();
      const int grow = (qt0 + i) * kBM + row;
      typename E::T* op = reinterpret_cast<typename E::T*>(a.o) + (int64_t)b * a.o_sb + (int64_t)grow * a.o_sn + (int64_t)h * a.o_sh;
      for (int cc = 0; cc < nchunk_d; ++cc) {
        uint32_t r[16];
        tmem_ld16(tO + 16 * cc, r);
        tc_wait_ld();
        if (cc == nchunk_d - 1) {  // O(i) is out of TMEM: PV(i+1) may overwrite it
          tc_fence_before();
          mbar_arrive(bar_oe);
        }
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
  }
  if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");  // a dependent launch must not complete before the launch it depends on
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int DTYPE, int DCH, int TCOLS>
int launch_cross_tc(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CrossTcArgs& a, dim3 grid, cudaStream_t st) {
  auto kern = cross_tc_kernel<DTYPE, DCH, TCOLS>;
  IEF_CONFIG_SMEM(kern, cross_tc_smem<DCH>());
  if (a.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = cross_tc_smem<DCH>();
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    IEF_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, mq, mk, mv, a));
  } else {
    kern<<<grid, kThreads, cross_tc_smem<DCH>(), st>>>(mq, mk, mv, a);
  }
  IEF_LAUNCH_OK("cross_tc_kernel");
  return IEF_OK;
}

}  // namespace

// Plain cross-attention rows on the tensor pipe. Returns IEF_ERR_UNSUPPORTED (without setting the error text) when the shape is
// outside this kernel's range so that the caller can fall through to the mma.sync kernel.
bool ief_cross_tc_supported(const ief_cross_params* p) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("IEF_CROSS_TC"); on = (e && e[0] == '0') ? 0 : 1; }
  return on && p->Nk <= kNK && p->d % 8 == 0 && p->d >= 8 && p->d <= 160 && p->Nq >= kBM && (p->dtype == IEF_BF16 || p->dtype == IEF_F16);
}

int ief_cross_tc_launch(const ief_cross_params* p, cudaStream_t st, const int32_t* rows, int n_rows, int pdl) {
  CrossTcArgs a;
  a.pdl = pdl;
  const int nb = rows ? n_rows : p->B;
  for (int i = 0; i < IEF_MAX_ROWS; ++i) a.row[i] = i < nb ? (rows ? rows[i] : i) : 0;
  a.o = p->o.ptr; a.o_sb = p->o.stride_b; a.o_sn = p->o.stride_n; a.o_sh = p->o.stride_h;
  a.B = p->B; a.H = p->H; a.Nq = p->Nq; a.Nk = p->Nk; a.d = p->d;
  a.ksteps_qk = ief_ceil_div(p->d, 16);
  a.dv_mma = ((p->d + 15) / 16) * 16;
  const int fmt = p->dtype == IEF_BF16 ? 1 : 0;
  a.idesc_qk = make_idesc_f16(kBM, kNK, fmt, 0, 0);
  a.idesc_pv = make_idesc_f16(kBM, a.dv_mma, fmt, 0, 1);
  a.scale_log2 = p->scale * kLog2e;
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = ief_tc_make_map(&mq, p->dtype, p->q, p->d, p->Nq, p->H, p->B, a.perm_q, kBM)) != IEF_OK) return rc;
  if ((rc = ief_tc_make_map(&mk, p->dtype, p->k, p->d, p->Nk, p->H, p->B, a.perm_k, kNK)) != IEF_OK) return rc;
  if ((rc = ief_tc_make_map(&mv, p->dtype, p->v, p->d, p->Nk, p->H, p->B, a.perm_v, kNK)) != IEF_OK) return rc;
  const int cfg = p->d <= 48 ? 0 : (p->d <= 64 ? 1 : (p->d <= 128 ? 2 : 3));  // (chunks, TMEM columns): (1,128) (1,256) (2,256) (3,256)
  // One wave: as many consecutive query tiles per CTA as it takes for all CTAs to be resident at once (4 per SM at 128 TMEM
  // columns, 2 at 256); the per-CTA prologue (barriers, TMEM allocation, K/V load) is then paid once per 1-8 tiles.
  const int sms = ief_sm_count();
  const int nqt = ief_ceil_div(p->Nq, kBM);
  const long slots = (long)sms * (cfg == 0 ? 4 : 2);
  int tpc = (int)(((long)nqt * p->H * nb + slots - 1) / slots);
  tpc = tpc < 1 ? 1 : (tpc > 8 ? 8 : tpc);
  if (tpc > nqt) tpc = nqt;
  a.tiles_per_cta = tpc;
  dim3 grid(ief_ceil_div(nqt, tpc), p->H, nb);
  if (p->dtype == IEF_BF16) {
    switch (cfg) {
      case 0: return launch_cross_tc<IEF_BF16, 1, 128>(mq, mk, mv, a, grid, st);
      case 1: return launch_cross_tc<IEF_BF16, 1, 256>(mq, mk, mv, a, grid, st);
      case 2: return launch_cross_tc<IEF_BF16, 2, 256>(mq, mk, mv, a, grid, st);
      default: return launch_cross_tc<IEF_BF16, 3, 256>(mq, mk, mv, a, grid, st);
    }
  }
  switch (cfg) {
    case 0: return launch_cross_tc<IEF_F16, 1, 128>(mq, mk, mv, a, grid, st);
    case 1: return launch_cross_tc<IEF_F16, 1, 256>(mq, mk, mv, a, grid, st);
    case 2: return launch_cross_tc<IEF_F16, 2, 256>(mq, mk, mv, a, grid, st);
    default: return launch_cross_tc<IEF_F16, 3, 256>(mq, mk, mv, a, grid, st);
  }
}
