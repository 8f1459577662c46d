// Definitions shared by the two tcgen05 attention kernels (attn_tc.cu, attn_tc2.cu).
#pragma once
#include "ief_common.cuh"
#include "ptx_sm100.cuh"
#include <cuda.h>

constexpr int kTcChunkBytes = 128 * 128;  // one [128 rows x 64 elem] SWIZZLE_128B box = 16 KiB

struct TcArgs {
  void* o;
  int64_t o_sb, o_sn, o_sh;
  int32_t B, H, Nq, Nk, d;
  int32_t nt1, nt2;        // KV tiles in block 1 / block 2 (Union)
  int32_t ksteps_qk;       // ceil(d/16)
  int32_t dv_mma;          // N of the PV MMA (>= d, multiple of 16)
  uint32_t idesc_qk, idesc_pv;
  uint32_t idesc_sum;      // attn_tc3: [128 x 16] = P x ones (row sums on the tensor pipe), K-major B operand
  int32_t sum_mma;         // attn_tc3: 1 when dv_mma <= 48: the 16 columns after O in each 64-column accumulator hold the row sums
  int32_t nomax;           // attn_tc3, bf16: optimistic unshifted softmax with an in-CTA exact second pass (MAXMODE 2)
  float scale_log2;
  int32_t perm_q[3], perm_k[3], perm_v[3];  // which of (token, head, row) feeds TMA coordinate 1..3
  int32_t work_offset;  // attn_tc3: first linear work item (q block + nq_blocks * (head + H * row)) of this launch
  int32_t nq_blocks;    // attn_tc3: query blocks per (row, head) in this launch's flavour
  uint64_t rcp_nq, rcp_nq_h;  // attn_tc3: ceil(2^32 / nq_blocks), ceil(2^32 / (nq_blocks * H)): work item -> (q block, head, row) without integer division
  uint64_t active_mask; // attn_tc3: bit b = rows.active[b] (a scalar instead of an indexed constant-bank load between two items)
  int32_t pdl;          // attn_tc3: launched as a programmatic dependent of the previous attn_tc3 launch (the remainder of a hybrid call)
  int32_t n_items;      // attn_tc3: work items of this launch (a persistent CTA takes blockIdx.x, blockIdx.x + gridDim.x, ...)
  float* lse_out;         // attn_tc2 / attn_tc3: row log-sum-exp in log2 units, [B][H][Nq] (for the stored-maps sweep), or null
  const float* key_bias;  // attn_tc3: [n_bias][Nk] additive bias on the scaled scores (masked MasaCtrl), rows.bias[b] selects, or null
  const float* knorm;   // attn_tc3, bf16: [B][H][knorm_tiles] max key norm per 128-key tile (pre-pass), or null
  int32_t knorm_tiles;
  IefRowTable rows;
  long long* dbg;  // optional clock64 trace of CTA (0,0,0), see ief_debug_set_trace_buffer
};

__device__ __forceinline__ void tc_tma_tile(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int tok, int head, int row,
                                            const int32_t (&perm)[3]) {
  int cc[3] = {tok, head, row};
  sm100::tma_load_4d(dst, m, bar, c0, cc[perm[0]], cc[perm[1]], cc[perm[2]]);
}

long long* ief_debug_trace_buffer();
// 4-D SWIZZLE_128B tensor map over a [rows, tokens, heads, d] view, box = 64 channels x box_rows tokens (attn_tc.cu)
int ief_tc_make_map(CUtensorMap* m, int dtype, const ief_tensor4& t, int d, int N, int H, int B, int32_t perm[3], int box_rows);
int ief_attn_tc3_launch(const ief_attn_params* p, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const TcArgs& a, int mode,
                        cudaStream_t st);  // mode: 0 pair, 1 split, 2 hybrid (full waves as pairs, remainder split)
bool ief_attn_tc3_persistent(long items, int nt);  // would a 256-row launch of `items` items with nt key tiles each run as persistent CTAs?
int ief_attn_tc2_launch(const ief_attn_params* p, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const TcArgs& a, cudaStream_t st);
