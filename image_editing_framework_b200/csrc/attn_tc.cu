// Controlled self-attention, flash style, for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM.
//
//   O[b] = softmax(scale * Q[q_src[b]] K[k_src[b]]^T) V[v_src[b]]        (ief_attn_fwd, include/ief_b200.h)
//
// replaces the materialised-probability arithmetic of the reference closures
//   p2p/model/register.py:47-50, masactrl/model/register.py:35-44 + attention_control.py:37-68,
//   pnp/model/register.py:44-75, pix2pix-zero/model/attention_control.py:43-48
// with one online-softmax pass; the edit is nothing but the per-row (q,k,v) source indices.
//
// CTA = one 128-row query tile of one (batch row, head). 6 warps:
//   warp 0      TMA producer   (Q once, then a ring of K/V tiles of 128 keys)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..5  softmax: one query row per thread (TMEM lane == row), lazy O rescale, epilogue
// TMEM columns: S (128 x fp32) at 0, P (bf16 pairs) aliases S[0,64), O at 128.
// Two CTAs are resident per SM for head_dim <= 64, so one CTA's exp work overlaps the other's MMAs.
#include "ief_common.cuh"
#include "ptx_sm100.cuh"
#include "attn_tc_host.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <math.h>

using namespace sm100;

namespace {

constexpr int kBM = 128;          // query rows per CTA
constexpr int kBN = 128;          // keys per KV tile
constexpr int kChunkBytes = kTcChunkBytes;
constexpr int kThreads = 192;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units: P may grow to 2^8 before O is rescaled

template <int DCH> struct TcCfg {
  static constexpr int kStages = (DCH == 3) ? 1 : 2;
  static constexpr int kTileBytes = DCH * kChunkBytes;
  static constexpr int kSmemData = kTileBytes * (1 + 2 * kStages);
  static constexpr int kSmemBytes = kSmemData + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = (DCH == 3) ? 512 : 256;
  static constexpr int kMinBlocks = (DCH == 1) ? 2 : 1;
};

#define tma_tile tc_tma_tile

template <int DTYPE, int DCH>
__global__ void __launch_bounds__(kThreads, TcCfg<DCH>::kMinBlocks)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
               const __grid_constant__ TcArgs a) {
  using Cfg = TcCfg<DCH>;
  using E = ElemT<DTYPE>;
  constexpr int ST = Cfg::kStages;

  const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
  if (!a.rows.active[b]) return;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;
  auto sK = [&](int s) { return base + Cfg::kTileBytes * (1 + 2 * s); };
  auto sV = [&](int s) { return base + Cfg::kTileBytes * (2 + 2 * s); };
  const uint32_t bar0 = base + Cfg::kSmemData;
  const uint32_t bar_q = bar0;                 // Q landed
  const uint32_t bar_s = bar0 + 8;             // S tile (or final O) complete in TMEM
  const uint32_t bar_p = bar0 + 16;            // P written by all 128 softmax threads
  auto bar_full = [&](int s) { return bar0 + 24 + 8 * s; };
  auto bar_empty = [&](int s) { return bar0 + 24 + 8 * ST + 8 * s; };
  const uint32_t tmem_slot = bar0 + 24 + 16 * ST;  // 4 bytes
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = a.nt1 + a.nt2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_q, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    for (int s = 0; s < ST; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
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
  const uint32_t tmem_S = tmem_base;        // columns [0,128)
  const uint32_t tmem_P = tmem_base;        // aliases S columns [0,64)
  const uint32_t tmem_O = tmem_base + 128;  // columns [128, 128 + dv_mma)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int qb = a.rows.q[b];
      mbar_arrive_expect_tx(bar_q, Cfg::kTileBytes);
#pragma unroll
      for (int c = 0; c < DCH; ++c) tma_tile(sQ + c * kChunkBytes, &tmQ, bar_q, c * 64, qt * kBM, h, qb, a.perm_q);
      for (int j = 0; j < nt; ++j) {
        const int s = j % ST, ph = (j / ST) & 1;
        mbar_wait(bar_empty(s), ph ^ 1);
        const bool blk2 = j >= a.nt1;
        const int jj = blk2 ? j - a.nt1 : j;
        const int kb = blk2 ? a.rows.k2[b] : a.rows.k[b];
        const int vb = blk2 ? a.rows.v2[b] : a.rows.v[b];
        mbar_arrive_expect_tx(bar_full(s), 2 * Cfg::kTileBytes);
#pragma unroll
        for (int c = 0; c < DCH; ++c) tma_tile(sK(s) + c * kChunkBytes, &tmK, bar_full(s), c * 64, jj * kBN, h, kb, a.perm_k);
#pragma unroll
        for (int c = 0; c < DCH; ++c) tma_tile(sV(s) + c * kChunkBytes, &tmV, bar_full(s), c * 64, jj * kBN, h, vb, a.perm_v);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      mbar_wait(bar_q, 0);
      for (int j = 0; j < nt; ++j) {
        const int s = j % ST, ph = (j / ST) & 1;
        mbar_wait(bar_full(s), ph);
        tc_fence_after();
        // S = Q K^T : A = Q tile (K-major), B = K tile (K-major), contraction over head_dim in steps of 16
        for (int k = 0; k < a.ksteps_qk; ++k) {
          const uint32_t off = (k >> 2) * kChunkBytes + (k & 3) * 32;
          umma_ss(tmem_S, make_smem_desc_sw128(sQ + off, 16, 1024), make_smem_desc_sw128(sK(s) + off, 16, 1024), a.idesc_qk, k > 0);
        }
        umma_commit(bar_s);
        // O += P V : A = P from TMEM (bf16 pairs, 8 columns per 16 keys), B = V tile (MN-major: rows = keys)
        mbar_wait(bar_p, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kBN / 16; ++k) {
          umma_ts(tmem_O, tmem_P + k * 8, make_smem_desc_sw128(sV(s) + k * 2048, kChunkBytes, 1024), a.idesc_pv, (j > 0) || (k > 0));
        }
        umma_commit(bar_empty(s));  // K/V stage free once the PV MMAs have read it
      }
      umma_commit(bar_s);  // final O complete
    }
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    const int sub = warp & 3;  // TMEM sub-partition this warp may access
    const int row = sub * 32 + lane;
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    const uint32_t tS = tmem_S + lane_off, tP = tmem_P + lane_off, tO = tmem_O + lane_off;
    const float c2 = a.scale_log2;
    float m_used = -INFINITY, l = 0.f;
    const int nchunk_o = a.dv_mma >> 4;

    for (int j = 0; j < nt; ++j) {
      const bool blk2 = j >= a.nt1;
      const int jj = blk2 ? j - a.nt1 : j;
      const int vc = min(kBN, a.Nk - jj * kBN);  // valid keys in this tile
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      // pass 1: row max of the raw scores
      float tmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tS + 32 * c, r);
        tc_wait_ld();
        if (vc >= kBN) {
#pragma unroll
          for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, (32 * c + i < vc) ? __uint_as_float(r[i]) : -INFINITY);
        }
      }
      if (j == 0) {
        m_used = tmax;
      } else {
        const float m_new = fmaxf(m_used, tmax);
        const bool need = (m_new - m_used) * c2 > kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
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
      // pass 2: P = exp2(s*c2 - m*c2), packed to 16-bit pairs over the S columns already consumed
      const float mc = m_used * c2;
      float lsum0 = 0.f, lsum1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tS + 32 * c, r);
        tc_wait_ld();
        uint32_t u[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = ief_exp2(fmaf(__uint_as_float(r[2 * i]), c2, -mc));
          float p1 = ief_exp2(fmaf(__uint_as_float(r[2 * i + 1]), c2, -mc));
          if (vc < kBN) {
            if (32 * c + 2 * i >= vc) p0 = 0.f;
            if (32 * c + 2 * i + 1 >= vc) p1 = 0.f;
          }
          if constexpr (E::kIsBf16) {  // the row sum over the same 16-bit values the PV MMA multiplies (attn_tc_dev.cuh, exp_pack_chunk_mix)
            const uint32_t b0 = __float_as_uint(p0) & 0xffff0000u, b1 = __float_as_uint(p1) & 0xffff0000u;
            p0 = __uint_as_float(b0);
            p1 = __uint_as_float(b1);
            u[i] = __byte_perm(b0, b1, 0x7632);
          } else {
            u[i] = E::pack(p0, p1);
          }
          lsum0 += p0;
          lsum1 += p1;
        }
        tmem_st16(tP + 16 * c, u);
      }
      l += lsum0 + lsum1;
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    // epilogue: O / l -> 16-bit -> global (one row per thread, 16-byte stores)
    mbar_wait(bar_s, nt & 1);
    tc_fence_after();
    const float inv = 1.f / l;
    const int grow = qt * kBM + row;
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
  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Build a 4-D tensor map over a [rows, tokens, heads, d] view. The three outer dimensions are ordered by
// ascending stride (TMA is happiest with monotone strides); perm[i] says which of (token, head, row) is
// TMA dimension i+1. Box = 64 channels x 128 tokens, SWIZZLE_128B, out-of-bounds reads return zero —
// that is how head_dim 40/80/160 is padded to the 64-element swizzle atom and how ragged token counts
// are padded to the 128-row tile.
int make_map(CUtensorMap* m, int dtype, const ief_tensor4& t, int d, int N, int H, int B, int32_t perm[3], int box_rows = kBN) {
  EncodeTiledFn enc = get_encode();
  IEF_REQUIRE(enc != nullptr, IEF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  struct Dim { int kind; uint64_t size; int64_t stride; uint32_t box; } dims[3] = {
      {0, (uint64_t)N, t.stride_n, (uint32_t)box_rows}, {1, (uint64_t)H, t.stride_h, 1u}, {2, (uint64_t)B, t.stride_b, 1u}};
  // size-1 dimensions may carry arbitrary strides: sort them outermost and synthesise a legal stride below
  for (auto& x : dims) if (x.size == 1) x.stride = INT64_MAX / 4;
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (dims[j].stride < dims[i].stride) { Dim tmp = dims[i]; dims[i] = dims[j]; dims[j] = tmp; }
  cuuint64_t gdim[4] = {(cuuint64_t)d, dims[0].size, dims[1].size, dims[2].size};
  cuuint64_t gstr[3];
  int64_t prev = (int64_t)d;  // elements spanned so far, used to synthesise strides of size-1 dims
  for (int i = 0; i < 3; ++i) {
    int64_t s = dims[i].size == 1 ? ((prev + 7) / 8) * 8 : dims[i].stride;
    IEF_REQUIRE(s > 0 && (s * 2) % 16 == 0, IEF_ERR_UNSUPPORTED, "tensor stride %lld elements is not a multiple of 16 bytes", (long long)s);
    gstr[i] = (cuuint64_t)s * 2;
    prev = s * (int64_t)dims[i].size;
    perm[i] = dims[i].kind;
  }
  cuuint32_t box[4] = {64u, dims[0].box, dims[1].box, dims[2].box};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  IEF_REQUIRE((reinterpret_cast<uintptr_t>(t.ptr) & 15) == 0, IEF_ERR_INVALID, "tensor base pointer is not 16-byte aligned");
  CUresult r = enc(m, dtype == IEF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, t.ptr, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IEF_REQUIRE(r == CUDA_SUCCESS, IEF_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return IEF_OK;
}

template <int DTYPE, int DCH>
int launch_tc(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const TcArgs& a, dim3 grid, cudaStream_t st) {
  auto kern = attn_tc_kernel<DTYPE, DCH>;
  IEF_CONFIG_SMEM(kern, TcCfg<DCH>::kSmemBytes);
  kern<<<grid, kThreads, TcCfg<DCH>::kSmemBytes, st>>>(mq, mk, mv, a);
  IEF_LAUNCH_OK("attn_tc_kernel");
  return IEF_OK;
}

}  // namespace

int ief_tc_make_map(CUtensorMap* m, int dtype, const ief_tensor4& t, int d, int N, int H, int B, int32_t perm[3], int box_rows) {
  return make_map(m, dtype, t, d, N, H, B, perm, box_rows);
}

// max_j |k_j| per 128-key tile of every (row, head), inflated by 0.1 %: with |q_i| it bounds the tile's scores (Cauchy-Schwarz),
// which lets attn_tc3 (bf16) skip the running-maximum pass on most tiles. One thread per key, 16-byte loads.
__global__ void __launch_bounds__(128) key_norm_kernel(const __nv_bfloat16* k, int64_t sb, int64_t sn, int64_t sh, int Nk, int d, int H, int ntb,
                                                       float* out) {
  const int tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z, key = tile * 128 + threadIdx.x;
  float acc = 0.f;
  if (key < Nk) {
    const __nv_bfloat16* row = k + (int64_t)b * sb + (int64_t)key * sn + (int64_t)h * sh;
    for (int c = 0; c < d; c += 8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + c));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
        acc = fmaf(lo, lo, acc);
        acc = fmaf(hi, hi, acc);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc = fmaxf(acc, __shfl_xor_sync(0xffffffffu, acc, o));
  __shared__ float part[4];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) out[((int64_t)b * H + h) * ntb + tile] = sqrtf(fmaxf(fmaxf(part[0], part[1]), fmaxf(part[2], part[3]))) * 1.001f;
}

// The key-norm pre-pass (guarded skipping of the maximum pass, attn_tc3 MAXMODE 1) gained 3-7 % at head dims 49..64 and lost 3-13 %
// elsewhere (profiles/r01_kernel_microbench_gen3b.jsonl); the optimistic unshifted loop (MAXMODE 2) supersedes it without a
// pre-pass. It stays reachable for A/B runs with IEF_TC3_SKIPMAX=2.
static bool key_norm_prepass_pays(const ief_attn_params* p) {
  static int force = -1;
  if (force < 0) { const char* e = getenv("IEF_TC3_SKIPMAX"); force = (e && atoi(e) == 2) ? 1 : 0; }
  return force && p->dtype == IEF_BF16 && p->d <= 64 && p->Nq >= 512 && p->key_bias == nullptr && p->probs_out == nullptr;
}

// Stored maps (probs_out) of layers the tcgen05 kernels serve: O and the row log-sum-exp from them, then one QK^T sweep that writes
// the maps (attn_probs_from_lse_kernel) — instead of the two-sweep mma.sync kernel. Needs B*H*Nq floats of workspace.
bool ief_attn_probs_via_lse(const ief_attn_params* p) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("IEF_PROBS_VIA_LSE"); on = (e && e[0] == '0') ? 0 : 1; }
  if (!on || p->probs_out == nullptr || p->key_bias != nullptr || p->d > 128 || p->Nq < 512 || p->Nk < 256) return false;
  ief_attn_params q = *p;
  q.probs_out = nullptr;
  return ief_attn_tc_supported(&q, nullptr);
}

extern "C" int64_t ief_attn_workspace_bytes(const ief_attn_params* p) {
  if (p == nullptr) return 0;
  if (ief_attn_probs_via_lse(p)) return (int64_t)p->B * p->H * p->Nq * (int64_t)sizeof(float);
  if (!key_norm_prepass_pays(p)) return 0;
  return (int64_t)p->B * p->H * ief_ceil_div(p->Nk, 128) * (int64_t)sizeof(float);
}

bool ief_attn_tc_supported(const ief_attn_params* p, const char** why) {
  static const char* w = "";
  if (why) *why = w;
#define NO(msg) do { if (why) *why = msg; return false; } while (0)
  if (p->probs_out) NO("probs_out requires the mma two-sweep kernel");
  if (p->key_bias && p->bias_sel && (p->d > 64 || p->k_src2)) NO("key_bias: only the head_dim <= 64 generation of the tcgen05 kernel takes it, and no second key block");
  if (p->dtype != IEF_BF16 && p->dtype != IEF_F16) NO("dtype must be bf16 or f16");
  if (p->d % 8 != 0 || p->d < 8 || p->d > 192) NO("head_dim must be a multiple of 8 in [8,192]");
  if (p->Nq < 1 || p->Nk < 1) NO("empty sequence");
  const ief_tensor4* ts[4] = {&p->q, &p->k, &p->v, &p->o};
  for (auto t : ts) {
    if ((reinterpret_cast<uintptr_t>(t->ptr) & 15) != 0) NO("pointer not 16-byte aligned");
    if (t->stride_n % 8 || t->stride_h % 8 || t->stride_b % 8) NO("strides must be multiples of 8 elements");
  }
#undef NO
  return true;
}

int ief_attn_tc_launch(const ief_attn_params* p, const IefRowTable& rows, cudaStream_t st, float* lse_out) {
  const char* why = "";
  IEF_REQUIRE(ief_attn_tc_supported(p, &why), IEF_ERR_UNSUPPORTED, "tcgen05 attention: %s", why);
  TcArgs a;
  a.o = p->o.ptr; a.o_sb = p->o.stride_b; a.o_sn = p->o.stride_n; a.o_sh = p->o.stride_h;
  a.B = p->B; a.H = p->H; a.Nq = p->Nq; a.Nk = p->Nk; a.d = p->d;
  a.nt1 = ief_ceil_div(p->Nk, kBN);
  bool any2 = false;
  for (int i = 0; i < p->B; ++i) any2 |= rows.k2[i] >= 0;
  a.nt2 = any2 ? a.nt1 : 0;
  if (any2)
    for (int i = 0; i < p->B; ++i)
      IEF_REQUIRE(rows.k2[i] >= 0 && rows.v2[i] >= 0, IEF_ERR_UNSUPPORTED, "k_src2/v_src2 must be set for every row or none");
  a.ksteps_qk = ief_ceil_div(p->d, 16);
  const int dch = ief_ceil_div(p->d, 64);
  static int force_n64 = -1;
  if (force_n64 < 0) { const char* e = getenv("IEF_TC_PV_N64"); force_n64 = (e && e[0] == '1') ? 1 : 0; }
  a.dv_mma = force_n64 ? dch * 64 : ((p->d + 15) / 16) * 16;
  const int fmt = p->dtype == IEF_BF16 ? 1 : 0;
  a.idesc_qk = make_idesc_f16(kBM, kBN, fmt, 0, 0);
  a.idesc_pv = make_idesc_f16(kBM, a.dv_mma, fmt, 0, 1);
  a.idesc_sum = make_idesc_f16(kBM, 16, fmt, 0, 0);
  a.knorm = nullptr;
  a.knorm_tiles = 0;
  a.pdl = 0;
  a.lse_out = lse_out;
  a.key_bias = nullptr;
  for (int i = 0; i < p->B; ++i)
    if (rows.bias[i] >= 0) a.key_bias = p->key_bias;
  a.sum_mma = (a.dv_mma <= 48 && getenv("IEF_TC3_NO_SUM_MMA") == nullptr) ? 1 : 0;
  static int env_nomax = -1;  // IEF_TC3_NOMAX=0: always the exact running-maximum loop (A/B)
  if (env_nomax < 0) { const char* e = getenv("IEF_TC3_NOMAX"); env_nomax = (e && e[0] == '0') ? 0 : 1; }
  a.nomax = (env_nomax && p->dtype == IEF_BF16 && a.key_bias == nullptr) ? 1 : 0;
  a.scale_log2 = p->scale * kLog2e;
  a.rows = rows;
  a.dbg = ief_debug_trace_buffer();
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_map(&mq, p->dtype, p->q, p->d, p->Nq, p->H, p->B, a.perm_q)) != IEF_OK) return rc;
  if ((rc = make_map(&mk, p->dtype, p->k, p->d, p->Nk, p->H, p->B, a.perm_k)) != IEF_OK) return rc;
  if ((rc = make_map(&mv, p->dtype, p->v, p->d, p->Nk, p->H, p->B, a.perm_v)) != IEF_OK) return rc;
  // Kernel generations (IEF_TC_VERSION=1|2|3 caps the generation for A/B measurements):
  //   head_dim <= 64 : third generation (attn_tc3: column-split softmax, ordered exp sections), as 128-row split-KV CTAs when
  //                    that shortens the estimated wave time, else as 256-row CTAs; generation 2: attn_tc2 (256-row CTAs only)
  //   head_dim <= 128: second generation (attn_tc2, P aliased onto S)        above: first generation (this file)
  static int env_version = -1;
  if (env_version < 0) { const char* e = getenv("IEF_TC_VERSION"); env_version = (e && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : 3; }
  // a requested row log-sum-exp needs a generation whose epilogue writes it: 3 for head_dim <= 64, 2 (pair form) up to 128
  // ... and a key bias is only implemented by generation 3
  const int version = (lse_out != nullptr || a.key_bias != nullptr) ? 3 : env_version;
  IEF_REQUIRE(lse_out == nullptr || dch <= 2, IEF_ERR_UNSUPPORTED, "ief_attn_fwd: row log-sum-exp output needs head_dim <= 128");
  if (version >= 2 && dch == 1) {
    // 256-row CTAs (two query tiles share K/V) or 128-row CTAs (two key halves share Q)? Estimated time = waves x (key steps
    // per CTA + fixed prologue/epilogue, about three steps' worth): take the smaller. IEF_TC_SPLITKV=0|1|2 forces pair / split / hybrid.
    static int force = -2;  // process-wide A/B switch read from the environment once; nothing device-dependent is cached here
    if (force == -2) {
      const char* e = getenv("IEF_TC_SPLITKV");
      force = e ? atoi(e) : -1;
    }
    const int sms = ief_sm_count();
    // estimated time in half key steps (one step ~2.25 K cycles). Measured costs (tools/tc3_trace.py, profiles/r02_attn_tc3_persistent.txt):
    // a one-item 256-row CTA pays ~4 steps of prologue + epilogue, an item of a persistent CTA ~1.5 (+2.5 once per launch), a split-KV
    // CTA one more than a pair CTA for its merge, and a second launch ~3.5 steps (gap + cold start)
    const int nt = a.nt1 + a.nt2;
    const long pairs = (long)ief_ceil_div(p->Nq, 2 * kBM) * p->H * p->B;
    const long full = pairs / sms, rest = pairs % sms;
    auto pair_cost = [&](long waves, long items) {  // `items` 256-row items spread over the SMs in `waves` rounds
      return ief_attn_tc3_persistent(items, nt) ? waves * (2 * nt + 3) + 5 : waves * (2 * nt + 8);
    };
    const long split_wave = 2 * ((nt + 1) / 2) + 10;
    const long t_pair = pair_cost(full + (rest ? 1 : 0), pairs);
    const long t_split = ((2 * pairs + sms - 1) / sms) * split_wave;
    const long t_hybrid = pair_cost(full, full * sms) + ((2 * rest + sms - 1) / sms) * split_wave + (rest ? 7 : 0);
    int mode = 0;  // 0 pair, 1 split, 2 hybrid
    if (force >= 0) mode = force;
    else if (t_hybrid < t_pair && t_hybrid <= t_split && full > 0 && rest > 0) mode = 2;
    else if (t_split < t_pair) mode = 1;
    if (version >= 3) {
      const int ntb = ief_ceil_div(p->Nk, kBN);
      if (key_norm_prepass_pays(p) && p->workspace != nullptr && p->workspace_bytes >= (int64_t)p->B * p->H * ntb * (int64_t)sizeof(float)) {
        IEF_REQUIRE((reinterpret_cast<uintptr_t>(p->workspace) & 15) == 0, IEF_ERR_INVALID, "ief_attn_fwd: workspace must be 16-byte aligned");
        float* kn = static_cast<float*>(p->workspace);
        key_norm_kernel<<<dim3(ntb, p->H, p->B), 128, 0, st>>>(static_cast<const __nv_bfloat16*>(p->k.ptr), p->k.stride_b, p->k.stride_n, p->k.stride_h,
                                                               p->Nk, p->d, p->H, ntb, kn);
        IEF_LAUNCH_OK("key_norm_kernel");
        a.knorm = kn;
        a.knorm_tiles = ntb;
      }
      return ief_attn_tc3_launch(p, mq, mk, mv, a, mode, st);
    }
  }
  if (version >= 2 && dch <= 2) return ief_attn_tc2_launch(p, mq, mk, mv, a, st);
  dim3 grid(ief_ceil_div(p->Nq, kBM), p->H, p->B);
  if (p->dtype == IEF_BF16) {
    if (dch == 1) return launch_tc<IEF_BF16, 1>(mq, mk, mv, a, grid, st);
    if (dch == 2) return launch_tc<IEF_BF16, 2>(mq, mk, mv, a, grid, st);
    return launch_tc<IEF_BF16, 3>(mq, mk, mv, a, grid, st);
  } else {
    if (dch == 1) return launch_tc<IEF_F16, 1>(mq, mk, mv, a, grid, st);
    if (dch == 2) return launch_tc<IEF_F16, 2>(mq, mk, mv, a, grid, st);
    return launch_tc<IEF_F16, 3>(mq, mk, mv, a, grid, st);
  }
}

// ------------------------------------------------------------------------------------------------------------
// ief_umma_probe: one CTA, D[128,N] = A[128,K] * B through exactly the descriptor helpers used above.
namespace {
struct ProbeArgs {
  const void* a; float* d;
  int32_t N, K, b_mn, a_tmem;
  uint32_t idesc;
  int32_t perm_a[3], perm_b[3];
};

template <int DTYPE>
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ProbeArgs a) {
  using E = ElemT<DTYPE>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;                     // 2 chunks
  const uint32_t sB = base + 2 * kChunkBytes;   // up to 3 chunks
  const uint32_t bar_ld = base + 5 * kChunkBytes, bar_mma = bar_ld + 8, slot = bar_ld + 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_ld, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  const int kch = (a.K + 63) / 64;
  const int nch = a.b_mn ? (a.N + 63) / 64 : kch;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_ld, (kch + nch) * kChunkBytes);
    for (int c = 0; c < kch; ++c) tma_tile(sA + c * kChunkBytes, &tmA, bar_ld, c * 64, 0, 0, 0, a.perm_a);
    for (int c = 0; c < nch; ++c) tma_tile(sB + c * kChunkBytes, &tmB, bar_ld, c * 64, 0, 0, 0, a.perm_b);
  }
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  if (a.a_tmem) {  // stage A into TMEM columns [256, 256 + K/2) as 16-bit pairs, one row per thread
    const typename E::T* arow = reinterpret_cast<const typename E::T*>(a.a) + (int64_t)(warp * 32 + lane) * a.K;
    for (int c = 0; c < (a.K + 31) / 32; ++c) {
      uint32_t u[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int col = 32 * c + 2 * i;
        u[i] = col < a.K ? *reinterpret_cast<const uint32_t*>(arow + col) : 0u;
      }
      tmem_st16(tmem + lane_off + 256 + 16 * c, u);
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    mbar_wait(bar_ld, 0);
    tc_fence_after();
    for (int k = 0; k < a.K / 16; ++k) {
      const uint32_t koff = (k >> 2) * kChunkBytes + (k & 3) * 32;
      const uint64_t bdesc = a.b_mn ? make_smem_desc_sw128(sB + k * 2048, kChunkBytes, 1024) : make_smem_desc_sw128(sB + koff, 16, 1024);
      if (a.a_tmem) umma_ts(tmem, tmem + 256 + k * 8, bdesc, a.idesc, k > 0);
      else umma_ss(tmem, make_smem_desc_sw128(sA + koff, 16, 1024), bdesc, a.idesc, k > 0);
    }
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  float* drow = a.d + (int64_t)(warp * 32 + lane) * a.N;
  for (int cc = 0; cc < a.N / 16; ++cc) {
    uint32_t r[16];
    tmem_ld16(tmem + lane_off + 16 * cc, r);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) drow[16 * cc + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
}  // namespace

extern "C" int ief_umma_probe(const ief_umma_probe_params* p, void* stream) {
  IEF_REQUIRE(p && p->a && p->b && p->d, IEF_ERR_INVALID, "ief_umma_probe: null pointer");
  IEF_REQUIRE(p->N % 16 == 0 && p->N >= 16 && p->N <= (p->b_mn_major ? 192 : 128), IEF_ERR_UNSUPPORTED, "ief_umma_probe: N=%d unsupported", p->N);
  IEF_REQUIRE(p->K % 16 == 0 && p->K >= 16 && p->K <= 128, IEF_ERR_UNSUPPORTED, "ief_umma_probe: K=%d unsupported", p->K);
  IEF_REQUIRE(p->dtype == IEF_BF16 || p->dtype == IEF_F16, IEF_ERR_UNSUPPORTED, "ief_umma_probe: dtype");
  ProbeArgs a;
  a.a = p->a; a.d = p->d; a.N = p->N; a.K = p->K; a.b_mn = p->b_mn_major; a.a_tmem = p->a_from_tmem;
  a.idesc = make_idesc_f16(128, p->N, p->dtype == IEF_BF16 ? 1 : 0, 0, p->b_mn_major ? 1 : 0);
  CUtensorMap ma, mb;
  ief_tensor4 ta{const_cast<void*>(p->a), (int64_t)128 * p->K, (int64_t)p->K, (int64_t)p->K};
  int rc;
  if ((rc = make_map(&ma, p->dtype, ta, p->K, 128, 1, 1, a.perm_a)) != IEF_OK) return rc;
  if (p->b_mn_major) {
    ief_tensor4 tb{const_cast<void*>(p->b), (int64_t)p->K * p->N, (int64_t)p->N, (int64_t)p->N};
    if ((rc = make_map(&mb, p->dtype, tb, p->N, p->K, 1, 1, a.perm_b)) != IEF_OK) return rc;
  } else {
    ief_tensor4 tb{const_cast<void*>(p->b), (int64_t)p->N * p->K, (int64_t)p->K, (int64_t)p->K};
    if ((rc = make_map(&mb, p->dtype, tb, p->K, p->N, 1, 1, a.perm_b)) != IEF_OK) return rc;
  }
  const int smem = 5 * kChunkBytes + 1024 + 64;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p->dtype == IEF_BF16) {
    IEF_CUDA_OK(cudaFuncSetAttribute(umma_probe_kernel<IEF_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_probe_kernel<IEF_BF16><<<1, 128, smem, st>>>(ma, mb, a);
  } else {
    IEF_CUDA_OK(cudaFuncSetAttribute(umma_probe_kernel<IEF_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_probe_kernel<IEF_F16><<<1, 128, smem, st>>>(ma, mb, a);
  }
  IEF_LAUNCH_OK("umma_probe_kernel");
  return IEF_OK;
}
