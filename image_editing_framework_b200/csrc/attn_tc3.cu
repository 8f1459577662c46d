// Controlled self-attention for sm_100a, third generation (head_dim <= 64).
//
//   O[b] = softmax(scale * Q[q_src[b]] K[k_src[b]]^T) V[v_src[b]]        (ief_attn_fwd, include/ief_b200.h)
//
// One CTA per SM keeps TWO independent online-softmax streams (t = 0,1) in flight over a split TMEM layout
//   S_0 0 | S_1 128 | P_0 256 | P_1 320 | O_0 384 | O_1 448      (512 columns)
// so that QK(j+1) of a stream is issued as soon as its S(j) sits in the softmax registers. Two flavours, one template:
//   SPLIT = false  the streams are two 128-row QUERY tiles sharing the K/V ring           (256-row CTAs)
//   SPLIT = true   the streams are the two halves of the KEY range of ONE query tile, sharing Q; their partial (O, max, sum)
//                  are merged through shared memory at the end                            (128-row CTAs: twice as many,
//                  half as long — the cure for wave quantisation, e.g. SD-1.5's 64x64 layer: 3.46 -> 6.92 waves)
//
// Why 20 warps: the exp pipe (MUFU, 8 cycles per warp instruction per SM sub-partition) is the binding unit at head_dim
// 40-64 and one softmax warp per sub-partition only keeps it 84 % busy (two: 98 %, tools/micro/mufu.cu). Each stream's
// softmax is therefore spread over TWO warpgroups owning the left / right 64 score columns of every row:
//   warp 0   TMA producer      warps 1, 2   tcgen05.mma issuers, one per stream (elected lane)      warp 3   idle register donor (setmaxnreg)
//   warps 4-11   stream 0: columns 0-63 (warps 4-7) and 64-127 (warps 8-11)      warps 12-19   stream 1 likewise
// The halves of a row exchange their partial row maximum through shared memory (one 256-thread named barrier per key
// tile); partial row sums are merged in the epilogue; each half rescales / writes back its share of the O columns.
// The two streams' exp sections are strictly ORDERED (the ping-pong of FlashAttention-3/4): left alone they fall into
// lock-step and idle the MUFU together. Scale/shift (FMA pipe) happens before a stream takes its turn, and the turn is
// handed over one 32-column chunk early so the other stream's wake-up overlaps the tail.
//
// What bounded generations 3-3c (round 2, profiles/r02_tc3_no_exp_diagnostic.txt): with every exponential removed the kernel was NOT
// faster. A tcgen05.mma group needs ~300 clk from issue to its barrier, and ONE issuer warp walking QK_0 QK_1 PV_0 PV_1 with a blocking
// wait before each strung four such round trips (each answered by the softmax side) together per key tile: ~2350 clk, the same as two
// ordered exp sections. Hence one issuer warp per stream (kDualIssue) — after which the exp pipe binds and a quarter of the
// exponentials go to the FMA pipe (IEF_TC3_EMUL). The compile-time variants below delete softmax-warp instructions:
//   SUMMMA  head_dim <= 48: row sums come from an extra [128 x 16] = P x ones MMA into the accumulator columns left free after O
//           (same fp32 accumulation and lazy rescale as O) instead of 32 packed adds per row and tile
//   MAXMODE 1 (SKIP)   bf16 with a key-norm pre-pass (TcArgs::knorm): tiles whose Cauchy-Schwarz score bound is provably harmless
//           skip the running-maximum pass, the exchange between the column halves and the rescale decision (A/B only now)
//   MAXMODE 2 (NOMAX)  bf16: OPTIMISTIC UNSHIFTED SOFTMAX. bf16 P and the fp32 accumulators carry 8 exponent bits, so
//           P = exp2(scale*log2e*s) needs no running maximum at all while every row sum stays inside [2^-100, 2^100] - true for any
//           attention logits within +-69 (natural units). The CTA runs its whole key loop without the maximum pass, the exchange
//           between the column halves, the rescale decision and the shift, then looks at its row sums: in range -> normalise and
//           store (the result is the softmax up to rounding: O/l is invariant to the reference point); out of range (inf, NaN,
//           underflow) -> the SAME CTA repeats its key loop with the exact running-maximum loop (second pass; the mbarrier phase
//           parities of that pass are offset by the number of phases the first pass completed, PbPass2). No pre-pass, no second
//           launch; what the trace showed to bound the kernel - the serial chain load S -> max -> exchange -> decide -> scale ->
//           waits between two exp sections of a stream (~1300-1500 clk against a 1024 clk exp section of the other stream) -
//           shrinks to load S -> scale -> waits.
// (the same switches as run-time branches inside one loop cost 8-13 %: the fast and the exact loop are separate instantiations.)
//
// Launch forms (generation 3e): 256-row CTAs run PERSISTENT when a launch has more items than SMs (kPersist below: the next item's Q
// load, first QK^T and exp turn overlap the epilogue of the current one); the remainder of a hybrid call (fewer than #SM/2 256-row items,
// run as split-KV CTAs) is a programmatic dependent launch that backfills SMs as the persistent CTAs leave.
#include "ief_common.cuh"
#include "ptx_sm100.cuh"
#include "attn_tc_host.cuh"
#include "attn_tc_dev.cuh"
#include <math.h>
#include <stdlib.h>

using namespace sm100;

namespace {

constexpr int kBM = 128, kBN = 128;
constexpr int kThreads = 640;
constexpr int kRegsLow = 32, kRegsHigh = 112;  // pool = 640 threads x 96 regs at launch (61440): 128*32 + 512*112 = 61440
#ifndef IEF_TC3_STAGES
#define IEF_TC3_STAGES 3
#endif
#ifndef IEF_TC3_STAGES_PAIR
#define IEF_TC3_STAGES_PAIR IEF_TC3_STAGES  // ring depth of the 256-row flavour (one K + one V tile per stage: 4 stages fit, the split flavour's 3 x 4 tiles do not)
#endif
constexpr int kTile = kTcChunkBytes;           // 16 KiB: one [128 x 64ch] box (head_dim <= 64)
constexpr int kSmemOnes = 2 * 1024;            // [16 x 64] tile of 1.0 (K-major, SWIZZLE_128B footprint): B operand of the row-sum MMA
constexpr int kSmemXchg = 8 * 1024;            // floats: row-max exchange [2 parities][2 streams][2 halves][128] | row sums [2][2][128] | split merge [2][128]
// clock64 phase stamps of CTA 0 for tools/tc3_trace.py: 0 = compiled out (the per-tile predicate tests alone cost ~5 % of the
// softmax warps' instructions), 1 = per-tile phases, 2 = plus the sub-phases of the pre-turn work.
// Build a traced library with IEF_EXTRA_NVCC_FLAGS="-DIEF_TC3_TRACE=2" csrc/build.sh
#ifndef IEF_TC3_TRACE
#define IEF_TC3_TRACE 0
#endif
#define IEF_TC3_FINE_TRACE (IEF_TC3_TRACE >= 2)
#ifndef IEF_TC3_TRACE_ITEM
#define IEF_TC3_TRACE_ITEM 0  // which of a persistent CTA's items gets the per-tile stamps
#endif
#ifndef IEF_TC3_LD64
#define IEF_TC3_LD64 0  // experiment: the 64 score columns of a thread in one tcgen05.ld.x64 instead of two x32
#endif
#ifndef IEF_TC3_HANDOVER_EARLY
#define IEF_TC3_HANDOVER_EARLY 0  // experiment: hand the turn over on entering the exp section
#endif
#ifndef IEF_TC3_HANDOVER_LATE
#define IEF_TC3_HANDOVER_LATE 0  // experiment: hand the turn over after the whole exp section instead of one chunk early
#endif
// Every IEF_TC3_EMUL-th column pair is exponentiated on the FMA / ALU pipes (Cody-Waite + degree-3 minimax, 7.5e-5, attn_tc_dev.cuh)
// inside the exp section, interleaved with the MUFU ones. Round 1 measured no gain from this (and a loss when placed before the turn):
// the kernel was then equally bound by the single MMA issuer's chain of round trips. With two issuers (kDualIssue) the exp pipe is
// the binding unit and a quarter of the pairs on the FMA pipe buys +9-10 % at head_dim 40, +3 % at 64 (1/6: +8 / +1, 1/3: +6 / -1,
// 1/2: +3 / -8 %; profiles/r02_attn_tc3_dual_issue_emulation.txt).
#ifndef IEF_TC3_EMUL
#define IEF_TC3_EMUL 4
#endif
#ifndef IEF_TC3_EMUL_D64
#define IEF_TC3_EMUL_D64 IEF_TC3_EMUL  // the same share for head_dim 49-64 (row sums in registers: the FMA pipe is busier there)
#endif
template <bool SUMMMA> constexpr int kDefaultEmul = SUMMMA ? IEF_TC3_EMUL : IEF_TC3_EMUL_D64;
constexpr float kRescaleThreshold = 8.0f;
// bf16 only: a tile whose Cauchy-Schwarz score bound stays within 2^kSkipMargin of the first tile's smallest row maximum needs no
// row-maximum pass at all (P and the fp32 accumulators have 8 exponent bits; see the softmax section)
constexpr float kSkipMargin = 80.0f;
// MAXMODE 2: a row sum outside (2^-100, 2^100) sends the CTA into its exact second pass
constexpr float kNoMaxLo = 7.8886090522101181e-31f, kNoMaxHi = 1.2676506002282294e+30f;
#ifndef IEF_TC3_FAST_ORDERED
#define IEF_TC3_FAST_ORDERED 1     // 0: the unshifted loop runs its exp sections unordered (A/B)
#endif
// Two switches of the unshifted loop, per accumulator layout (measured on the persistent form, profiles/r02_attn_tc3_persistent.txt):
//   scale inside the exp section instead of in front of the turn   head_dim <= 48: -1 %      49-64: +1-2 %
//   wait for PV_t(j-1) right before the first store of P (half an exp section into the turn) instead of in front of the turn
//                                                                   head_dim <= 48: +2.5 %    49-64: -6 %
// (head_dim 49-64 sums its rows in registers: ~340 instructions per thread and tile inside the exp section against ~240, its
// sections are bound by instruction issue as much as by the MUFU — with NO exponential on the FMA pipe it runs at the same speed)
#ifndef IEF_TC3_FAST_SCALE_IN_TURN
#define IEF_TC3_FAST_SCALE_IN_TURN 0
#endif
#ifndef IEF_TC3_FAST_SCALE_IN_TURN_D64
#define IEF_TC3_FAST_SCALE_IN_TURN_D64 1
#endif
#ifndef IEF_TC3_PV_WAIT_LATE
#define IEF_TC3_PV_WAIT_LATE 1
#endif
#ifndef IEF_TC3_PV_WAIT_LATE_D64
#define IEF_TC3_PV_WAIT_LATE_D64 0
#endif
constexpr bool kFastOrdered = IEF_TC3_FAST_ORDERED != 0;
// Two tcgen05.mma issuer warps, one per stream (warp 1: stream 0, warp 2: stream 1), instead of one warp walking QK_0 QK_1 PV_0 PV_1 in
// a fixed order with a blocking wait before each: a commit group needs ~300 clk from issue to its barrier (tools/micro/umma_rate.cu) and
// the softmax side reacts to each of them, so the single in-order issuer strung four such round trips together per key tile — ~2350 clk
// per tile pair even with every exponential removed (profiles/r02_tc3_no_exp_diagnostic.txt), as much as the exp sections themselves.
#ifndef IEF_TC3_DUAL_ISSUE
#define IEF_TC3_DUAL_ISSUE 1
#endif
constexpr bool kDualIssue = IEF_TC3_DUAL_ISSUE != 0;
// Persistent 256-row CTAs: one CTA per SM walks its share of the work items (item = first + blockIdx.x + i * gridDim.x) with TMEM,
// barriers and tensor maps set up once. Nothing is drained between items: the producer fetches the next Q as soon as the last QK^T of
// the current item has completed and runs ahead in the K / V ring, each issuer puts S(0) of the next item into TMEM as soon as the
// last score tile of the current one sits in the softmax registers, so that a stream's epilogue (~2.5 K cycles) overlaps the other
// stream's exp sections and the next item's first exp section starts right behind it. All barriers simply keep counting: the phase
// parities of item number k are those of item 0 xor (k odd ? phases per item, per barrier : 0) (PbSeq), the same mechanism as the
// exact second pass. An item whose unshifted row sums leave the safe range (MAXMODE 2) is only flagged; flagged items are repeated
// with the exact loop after the CTA's last item (one rendezvous per CTA instead of one per item).
#ifndef IEF_TC3_PERSIST
#define IEF_TC3_PERSIST 1
#endif
constexpr bool kPersist = IEF_TC3_PERSIST != 0;
#ifndef IEF_TC3_PIN_BAR0
#define IEF_TC3_PIN_BAR0 1
#endif
constexpr int kMaxItemsPerCta = 32;  // bits of the redo mask

// mbarrier phase-parity bases of a pass: zero for the first (or only) pass; for the exact second pass of a MAXMODE 2 CTA the number
// of phases each barrier completed during the first pass, mod 2
struct PbZero {
  __device__ __forceinline__ int stage(int) const { return 0; }
  __device__ __forceinline__ int strm(int) const { return 0; }
  __device__ __forceinline__ int turn(int) const { return 0; }
  __device__ __forceinline__ int item() const { return 0; }
};
struct PbPass2 {
  uint32_t m;  // bits 0..3: ring stages, 4..5: per-stream barriers (S, P, C, O), 6..7: exp-turn barriers
  __device__ __forceinline__ int stage(int s) const { return (m >> s) & 1; }
  __device__ __forceinline__ int strm(int t) const { return (m >> (4 + t)) & 1; }
  __device__ __forceinline__ int turn(int t) const { return (m >> (6 + t)) & 1; }
  __device__ __forceinline__ int item() const { return 0; }  // the second pass re-uses the Q tile: no new phase of the Q barrier
};
struct PbSeq {  // persistent CTAs: bases of the k-th item of a CTA (k items done, each advanced every barrier by the same number of phases)
  uint32_t stages;  // bit s: ring stage s
  int kk;           // Q loaded / Q free / O drained: one phase per item
  __device__ __forceinline__ int stage(int s) const { return (stages >> s) & 1; }
  __device__ __forceinline__ int strm(int) const { return 0; }  // per-stream (S, P, C, O) and exp-turn barriers: nt phases per item, nt even
  __device__ __forceinline__ int turn(int) const { return 0; }
  __device__ __forceinline__ int item() const { return kk; }
};
template <bool V> struct BoolTag { static constexpr bool value = V; };

template <bool SPLIT> struct Cfg3 {
  static constexpr int kQTiles = SPLIT ? 1 : 2;
  static constexpr int kRingTiles = SPLIT ? 2 : 1;                       // K (and V) tiles per ring stage
  static constexpr int kStages = SPLIT ? IEF_TC3_STAGES : IEF_TC3_STAGES_PAIR;  // K / V ring depth (4: no gain, profiles/r02 notes)
  static constexpr int kSmemData = kTile * (kQTiles + 2 * kStages * kRingTiles);
  static constexpr int kSmemBytes = kSmemData + kSmemOnes + kSmemXchg + 1024 + 256;
};

// sqrt.approx (MUFU): sqrtf() would call the IEEE slow-path subroutine, which does not fit the 32-register helper warps
__device__ __forceinline__ float approx_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <int DTYPE, bool SPLIT, int EMUL, bool SUMMMA, int MAXMODE, bool BIAS, bool PERSIST>
__global__ void __launch_bounds__(kThreads, 1)
attn_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                const __grid_constant__ TcArgs a) {
  using E = ElemT<DTYPE>;
  using Cfg = Cfg3<SPLIT>;
  constexpr int RT = Cfg::kRingTiles;
  constexpr int ST = Cfg::kStages;
  constexpr bool SKIP = MAXMODE == 1, NOMAX = MAXMODE == 2;
  static_assert(!(NOMAX && BIAS), "the unshifted loop has no key-bias form");
  // persistent form: 256-row flavour only (the split flavour stages its merge in the K ring), not for the key-norm variant (its
  // softmax warps read the Q tile), the unshifted loop must take exp turns like the exact one (same phases per item), and the launcher
  // only picks it for an EVEN number of key tiles per item: the per-tile barriers then advance an even number of phases per item and
  // the softmax warps, whose instruction issue bounds the kernel, need no parity bases at all
  static_assert(!PERSIST || (!SPLIT && !SKIP && kFastOrdered), "persistent form: 256-row flavour, not the key-norm variant, ordered exp sections");
  // 1-D grid over linearised work items so that a launch can cover any contiguous range of them (hybrid pair + split launches);
  // a persistent CTA takes the items blockIdx.x, blockIdx.x + gridDim.x, ... of the launch's n_items
  int qt, h, b;  // 256-row block (pair) or 128-row tile (split), head, batch row of the current item: every thread walks the same sequence
  auto set_item = [&](int it) -> bool {
    // division by the two launch constants through host-made reciprocals (exact for lin * divisor < 2^32, checked by the launcher):
    // this runs on the softmax warps between two items, next to the other stream's exp section
    const uint32_t lin = (uint32_t)(a.work_offset + (int)blockIdx.x + it * (int)gridDim.x);
    const uint32_t bb = (uint32_t)(((uint64_t)lin * a.rcp_nq_h) >> 32), rem = lin - bb * (uint32_t)(a.nq_blocks * a.H);
    const uint32_t hh = (uint32_t)(((uint64_t)rem * a.rcp_nq) >> 32);
    b = (int)bb;
    h = (int)hh;
    qt = (int)(rem - hh * (uint32_t)a.nq_blocks);
    return ((a.active_mask >> bb) & 1ull) != 0;
  };
  const int n_local = PERSIST ? (a.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 1;
  if constexpr (!PERSIST) {
    if (!set_item(0)) return;
  }
  // The two launches of a hybrid call (full waves | remainder) read the same inputs and write disjoint rows: the remainder is launched
  // as a programmatic dependent, so its CTAs move onto an SM as soon as that SM's CTA of this launch has left instead of after the
  // whole grid has drained and a launch gap (nothing here is consumed by the dependent: the trigger can be given at once).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const bool cta_trace = a.dbg != nullptr && blockIdx.x + a.work_offset == 0;
  if (cta_trace && threadIdx.x == 0) a.dbg[1536] = clock64();

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  auto sQ = [&](int t) { return base + kTile * (SPLIT ? 0 : t); };
  auto sK = [&](int s, int t) { return base + kTile * (Cfg::kQTiles + RT * s + (SPLIT ? t : 0)); };
  auto sV = [&](int s, int t) { return base + kTile * (Cfg::kQTiles + RT * ST + RT * s + (SPLIT ? t : 0)); };
  float* stage_o = reinterpret_cast<float*>(base_ptr + kTile * Cfg::kQTiles);  // split merge staging: K ring stage 0 (32 KiB), [col][row]
  const uint32_t sOnes = base + Cfg::kSmemData;
  float* xmax = reinterpret_cast<float*>(base_ptr + Cfg::kSmemData + kSmemOnes);  // [parity][stream][half][128]
  float* xsum = xmax + 2 * 2 * 2 * 128;                                       // [stream][half][128]
  float* xml = xsum + 2 * 2 * 128;                                            // split merge: [0..127] max, [128..255] sum of stream 1
  float* xq = xml + 256;                                   // [2 streams][8 warps][2]: one-time (max |q|, min row max) exchange
  uint32_t bar0_ = base + Cfg::kSmemData + kSmemOnes + kSmemXchg;
#if IEF_TC3_PIN_BAR0
  if constexpr (PERSIST || IEF_TC3_PIN_BAR0 >= 2) asm volatile("" : "+r"(bar0_));  // one register instead of re-deriving the shared-memory window address inside the key loop
#endif
  const uint32_t bar0 = bar0_;
  const uint32_t bar_q = bar0;
  auto bar_s = [&](int t) { return bar0 + 8 + 8 * t; };    // S_t complete in TMEM
  auto bar_p = [&](int t) { return bar0 + 24 + 8 * t; };   // P_t written by the 256 softmax threads of stream t
  auto bar_c = [&](int t) { return bar0 + 40 + 8 * t; };   // S_t in the registers of its 8 softmax warps
  auto bar_o = [&](int t) { return bar0 + 56 + 8 * t; };   // PV_t(j) complete
  auto bar_x = [&](int t) { return bar0 + 72 + 8 * t; };   // exp turn of stream t (granted by the 8 warps of the other stream)
  auto bar_kf = [&](int s) { return bar0 + 88 + 8 * s; };
  auto bar_ke = [&](int s) { return bar0 + 88 + 8 * (ST + s); };
  auto bar_vf = [&](int s) { return bar0 + 88 + 8 * (2 * ST + s); };
  auto bar_ve = [&](int s) { return bar0 + 88 + 8 * (3 * ST + s); };
  const uint32_t bar_qe = bar0 + 88 + 32 * ST;                            // persistent: every QK^T of the current item has completed (Q tile free)
  auto bar_d = [&](int t) { return bar0 + 88 + 32 * ST + 8 + 8 * t; };   // persistent: the epilogue of stream t has read O_t out of TMEM
  const uint32_t tmem_slot = bar0 + 88 + 32 * ST + 24;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  volatile uint32_t* redo_flag = tmem_slot_ptr + 2;  // MAXMODE 2: some row sum of the unshifted pass left the safe range (persistent: bit i = i-th item)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = a.nt1 + a.nt2;
  // bf16 + per-tile key-norm bounds from the pre-pass (ief_attn_tc_key_norms): tiles after the first skip the row-maximum pass
  // when the bound proves that harmless (see the softmax section)
  constexpr bool skip_guard = SKIP;
  // key tiles per stream: pair -> both streams see all nt tiles; split -> [0, nt0) and [nt0, nt)
  const int nt0 = SPLIT ? (nt + 1) >> 1 : nt, nt1 = SPLIT ? nt - nt0 : nt;
  auto stream_nt = [&](int t) { return t == 0 ? nt0 : nt1; };
  auto global_tile = [&](int t, int j) { return SPLIT && t == 1 ? nt0 + j : j; };
  auto kv_coord = [&](int g, int& jj, int& kb, int& vb) {  // global key tile -> (tile inside its key block, source rows)
    const bool blk2 = g >= a.nt1;
    jj = blk2 ? g - a.nt1 : g;
    kb = blk2 ? a.rows.k2[b] : a.rows.k[b];
    vb = blk2 ? a.rows.v2[b] : a.rows.v[b];
  };
  // parity bases of the exact second pass (MAXMODE 2): phases completed by the first pass, per barrier
  auto pass2_bases = [&]() {
    PbPass2 pb;
    pb.m = 0;
#pragma unroll
    for (int s = 0; s < ST; ++s) pb.m |= (uint32_t)((nt0 > s ? (nt0 - s + ST - 1) / ST : 0) & 1) << s;
    pb.m |= (uint32_t)(nt0 & 1) << 4 | (uint32_t)(nt1 & 1) << 5;
    if (kFastOrdered) pb.m |= (uint32_t)(nt1 & 1) << 6 | (uint32_t)(nt0 & 1) << 7;  // bar_x(t) completes once per section of stream t^1
    return pb;
  };
  auto seq_bases = [&](int k) {  // persistent (256-row flavour: nt0 == nt1 == nt, even)
    PbSeq pb;
    pb.stages = (k & 1) ? (pass2_bases().m & 0xfu) : 0u;
    pb.kk = k & 1;
    return pb;
  };
  // The items of a persistent CTA: entry e < n_local is item e in its first (or only) form; with the unshifted loop, entries n_local
  // ... 2 n_local - 1 are the exact repeats of the flagged items, after one CTA-wide rendezvous. Returns the item index or -1 (skip),
  // -2 (done). Every thread of the CTA walks the same sequence; the state carried across an item is just (e, k).
  auto next_entry = [&](int e, bool& exact) -> int {
    exact = !NOMAX;
    if (e < n_local) return e;
    if (e == n_local) named_bar_sync(0, kThreads);  // the softmax warps have looked at the row sums of every item
    const uint32_t redo = *redo_flag;
    if (redo == 0) return -2;
    exact = true;
    return ((redo >> (e - n_local)) & 1) ? e - n_local : -1;
  };

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
      mbar_init(bar_d(t), 8);
    }
    mbar_init(bar_qe, kDualIssue ? 2 : 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(bar_kf(s), 1);
      mbar_init(bar_ke(s), kDualIssue ? 2 : 1);
      mbar_init(bar_vf(s), 1);
      mbar_init(bar_ve(s), kDualIssue ? 2 : 1);
    }
    *redo_flag = 0;
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if constexpr (SUMMMA) {
    // Row sums on the tensor pipe: l = P x ones lands in the 16 accumulator columns after O (free when head_dim <= 48), in
    // the same fp32 accumulation and under the same lazy rescale as O. That removes the 32 packed adds per row and tile from
    // the softmax warps, whose instruction issue — not the MUFU alone — bounds this kernel.
    uint32_t* ones = reinterpret_cast<uint32_t*>(base_ptr + Cfg::kSmemData);
    const uint32_t one2 = DTYPE == IEF_BF16 ? 0x3f803f80u : 0x3c003c00u;
    for (int i = threadIdx.x; i < kSmemOnes / 4; i += kThreads) ones[i] = one2;
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (cta_trace && threadIdx.x == 0) a.dbg[1537] = clock64();

  // ------------------------------------------------------------------ TMA producer (warp 0)
  auto producer = [&](auto pb, bool load_q, int k) {
    const int qb = a.rows.q[b];
    if (load_q) {
      if constexpr (PERSIST) {
        if (k > 0) mbar_wait(bar_qe, (k - 1) & 1);  // the previous item's last QK^T has read the Q tile
      }
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_q, Cfg::kQTiles * kTile);
        for (int t = 0; t < Cfg::kQTiles; ++t) tc_tma_tile(sQ(t), &tmQ, bar_q, 0, (Cfg::kQTiles * qt + t) * kBM, h, qb, a.perm_q);
      }
      __syncwarp();
    }
    for (int j = 0; j < nt0; ++j) {
      const int s = j % ST, ph = ((j / ST) + pb.stage(s)) & 1;
      const int nload = SPLIT ? (j < nt1 ? 2 : 1) : 1;
      mbar_wait(bar_ke(s), ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_kf(s), nload * kTile);
        for (int t = 0; t < nload; ++t) {
          int jj, kb, vb;
          kv_coord(global_tile(t, j), jj, kb, vb);
          tc_tma_tile(sK(s, t), &tmK, bar_kf(s), 0, jj * kBN, h, kb, a.perm_k);
        }
      }
      __syncwarp();
      mbar_wait(bar_ve(s), ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_vf(s), nload * kTile);
        for (int t = 0; t < nload; ++t) {
          int jj, kb, vb;
          kv_coord(global_tile(t, j), jj, kb, vb);
          tc_tma_tile(sV(s, t), &tmV, bar_vf(s), 0, jj * kBN, h, vb, a.perm_v);
        }
      }
      __syncwarp();
    }
  };

  // ------------------------------------------------------------------ MMA issuer (warp 1: whole warp waits, one elected lane issues)
  auto mma_issuer = [&](auto pb, int t_lo, int t_hi, int k) {
    const uint64_t desc_k = make_smem_desc_sw128(0, 16, 1024);
    const uint64_t desc_v = make_smem_desc_sw128(0, kTcChunkBytes, 1024);
    auto issue_qk = [&](int t, int s) {
      for (int k = 0; k < a.ksteps_qk; ++k)
        umma_ss(tmem_base + 128 * t, desc_k | ((sQ(t) + k * 32) >> 4), desc_k | ((sK(s, t) + k * 32) >> 4), a.idesc_qk, k > 0);
      umma_commit(bar_s(t));
    };
    auto issue_pv = [&](int t, int s, bool acc) {
#pragma unroll
      for (int k = 0; k < kBN / 16; ++k)
        umma_ts(tmem_base + 384 + 64 * t, tmem_base + 256 + 64 * t + k * 8, desc_v | ((sV(s, t) + k * 2048) >> 4), a.idesc_pv, acc || (k > 0));
      if constexpr (SUMMMA) {
#pragma unroll
        for (int k = 0; k < kBN / 16; ++k)  // every 16-key step multiplies the same all-ones [16 x 16] block
          umma_ts(tmem_base + 384 + 64 * t + a.dv_mma, tmem_base + 256 + 64 * t + k * 8, desc_k | ((sOnes + (k & 3) * 32) >> 4), a.idesc_sum, acc || (k > 0));
      }
    };
    // This warp issues for the streams [t_lo, t_hi): both (one issuer) or one each (two issuers, kDualIssue). Every issuer arrives
    // once per ring stage on the "K consumed" / "V consumed" barriers (their count = number of issuers), through a commit when it
    // has MMAs in flight on that stage and through a plain arrive when its stream has no tile there.
    auto release = [&](uint32_t bar, bool issued) {
      if (issued) umma_commit(bar); else mbar_arrive(bar);
    };
    mbar_wait(bar_q, pb.item());
    mbar_wait(bar_kf(0), pb.stage(0));
    if constexpr (PERSIST) {
      if (k > 0)  // S_t still holds the previous item's last score tile until its softmax warps have it in registers
        for (int t = t_lo; t < t_hi; ++t) mbar_wait(bar_c(t), pb.strm(t) ^ 1);
    }
    tc_fence_after();
    if (elect_one()) {
      bool any = false;
      for (int t = t_lo; t < t_hi; ++t)
        if (stream_nt(t) > 0) { issue_qk(t, 0); any = true; }
      release(bar_ke(0), any);
      if constexpr (PERSIST) {
        if (nt0 == 1) umma_commit(bar_qe);
      }
    }
    __syncwarp();
    for (int j = 0; j < nt0; ++j) {
      const int s = j % ST, ph = ((j / ST) + pb.stage(s)) & 1;
      const int s1 = (j + 1) % ST, ph1 = (((j + 1) / ST) + pb.stage(s1)) & 1;
      if (j + 1 < nt0) {  // next score tiles first: they only need S_t(j) to be in the softmax registers
        mbar_wait(bar_kf(s1), ph1);
        bool any = false;
        for (int t = t_lo; t < t_hi; ++t) {
          if (j + 1 < stream_nt(t)) {
            mbar_wait(bar_c(t), (j + pb.strm(t)) & 1);
            tc_fence_after();
            if (elect_one()) issue_qk(t, s1);
            __syncwarp();
            any = true;
          }
        }
        if (elect_one()) {
          release(bar_ke(s1), any);
          if constexpr (PERSIST) {
            if (j + 2 == nt0) umma_commit(bar_qe);  // that was the item's last QK^T: the producer may fetch the next item's Q
          }
        }
        __syncwarp();
      }
      mbar_wait(bar_vf(s), ph);
      bool any = false;
      for (int t = t_lo; t < t_hi; ++t) {
        if (j < stream_nt(t)) {
          mbar_wait(bar_p(t), (j + pb.strm(t)) & 1);
          if constexpr (PERSIST) {
            if (j == 0 && k > 0) mbar_wait(bar_d(t), (k - 1) & 1);  // PV(0) overwrites O_t: the previous item's epilogue must have read it
          }
          tc_fence_after();
          if (elect_one()) {
            issue_pv(t, s, j > 0);
            umma_commit(bar_o(t));
          }
          __syncwarp();
          any = true;
        }
      }
      if (elect_one()) release(bar_ve(s), any);
      __syncwarp();
    }
  };

  if (warp < 4) {
    reg_dec<kRegsLow>();
    auto role = [&](auto pb, bool load_q, int k) {
      if (warp == 0) producer(pb, load_q, k);
      else if (warp == 1) mma_issuer(pb, 0, kDualIssue ? 1 : 2, k);
      else if (warp == 2 && kDualIssue) mma_issuer(pb, 1, 2, k);
    };
    if constexpr (PERSIST) {
      // the pipeline roles do not care which loop the softmax warps run: every item is the same sequence of barrier phases
      int k = 0;
      for (int e = 0; e < (NOMAX ? 2 : 1) * n_local; ++e) {
        bool exact;
        const int it = next_entry(e, exact);
        if (it == -2) break;
        if (it < 0 || !set_item(it)) continue;
        role(seq_bases(k), true, k);
        ++k;
      }
    } else {
      role(PbZero{}, true, 0);
      if constexpr (NOMAX) {
        named_bar_sync(0, kThreads);  // the softmax warps have looked at the row sums of the unshifted pass
        if (*redo_flag) role(pass2_bases(), false, 0);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax of stream t, column half `half`
    reg_inc<kRegsHigh>();
    const int idx = warp - 4;
    const int t = idx >> 3, half = (idx >> 2) & 1;
    const int sub = warp & 3;  // TMEM sub-partition of this warp
    const int row = sub * 32 + lane;
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    const uint32_t tS = tmem_base + 128 * t + 64 * half + lane_off;
    const uint32_t tP = tmem_base + 256 + 64 * t + 32 * half + lane_off;
    const uint32_t tO = tmem_base + 384 + 64 * t + lane_off;
    // BIAS variant, rows with a key bias: the scores are turned into t = s * scale_log2 + bias * log2(e) right after the load, so the
    // maximum sees the bias and everything downstream works with a unit scale
    const float* bias_row = nullptr;
    float c2 = a.scale_log2;
    auto item_setup = [&]() {  // per item: the batch row decides whether a key bias applies
      if constexpr (BIAS) {
        bias_row = a.rows.bias[b] >= 0 ? a.key_bias + (int64_t)a.rows.bias[b] * a.Nk : nullptr;
        c2 = bias_row != nullptr ? 1.f : a.scale_log2;
      }
    };
    if constexpr (!PERSIST) item_setup();
    float m_used = -INFINITY, l = 0.f;
    float qn_max = INFINITY, m_floor = -INFINITY;  // stream-wide max |q_row| and min first-tile row maximum (scaled): the skip test
    float qn_row = 0.f;
    if constexpr (skip_guard) {
      // |q_row| from the Q tile in shared memory (chunk order rotated by lane against bank conflicts), before the score
      // registers go live
      mbar_wait(bar_q, 0);
      const uint8_t* qrow = base_ptr + (sQ(SPLIT ? 0 : t) - base) + row * 128;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(qrow + (((c + lane) & 7) << 4));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
          acc = fmaf(lo, lo, acc);
          acc = fmaf(hi, hi, acc);
        }
      }
      qn_row = approx_sqrt(acc) * 1.001f;
    }
    constexpr bool sum_mma = SUMMMA;
    const int nchunk_o = (a.dv_mma >> 4) + (sum_mma ? 1 : 0);  // accumulator chunks touched by the lazy rescale (the row sums ride along in the chunk after O)
    const int my_nt = stream_nt(t);

    // One pass over this stream's key tiles. FAST (MAXMODE 2, first pass): unshifted exponentials, no maximum, no exchange, no
    // rescale. Otherwise the exact online softmax.
    // Turn protocol: exp sections alternate 0(0) 1(0) 0(1) 1(1) ...; stream 0 waits for stream 1's previous section, stream
    // 1 for stream 0's current one. In the split flavour stream 1 may be one step short (odd tile count): stream 0's last
    // grant then simply goes unused.
    auto run_pass = [&](auto fast_tag, auto pb, int k) {
      constexpr bool FAST = decltype(fast_tag)::value;
      constexpr bool ordered = !FAST || kFastOrdered;
      for (int j = 0; j < my_nt; ++j) {
        int jj, kb_tile, vb_unused;
        kv_coord(global_tile(t, j), jj, kb_tile, vb_unused);
        const int vc = min(kBN, a.Nk - jj * kBN);
        float kn_cur = 0.f;  // requested here, needed after the score load: its latency hides behind the barrier wait and tcgen05.ld
        if constexpr (skip_guard && !FAST) kn_cur = __ldg(a.knorm + ((int64_t)kb_tile * a.H + h) * a.knorm_tiles + jj);
#if IEF_TC3_TRACE
        const bool trace = cta_trace && k == IEF_TC3_TRACE_ITEM && row == 0 && half == 0 && j < 64;
        long long* tr = trace ? a.dbg + (t * 64 + j) * 8 : nullptr;
#else
        constexpr bool trace = false;
        long long* tr = nullptr;
#endif
#if IEF_TC3_FINE_TRACE
        long long* tr2 = trace ? a.dbg + 1024 + (t * 64 + j) * 4 : nullptr;  // sub-phases of the pre-turn work
#endif
        if (trace) tr[0] = clock64();
        mbar_wait(bar_s(t), (j + pb.strm(t)) & 1);
        tc_fence_after();
        if (trace) tr[1] = clock64();
        uint32_t s0[32], s1[32];
#if IEF_TC3_LD64
        tmem_ld64(tS, s0, s1);
#else
        tmem_ld32(tS, s0);
        tmem_ld32(tS + 32, s1);
#endif
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_c(t));  // S_t(j) is in registers: the tensor pipe may overwrite it with S_t(j+1)
        if (trace) tr[2] = clock64();
        if constexpr (BIAS) {
          if (bias_row != nullptr) {
            // masked keys carry finfo.min in the reference; clamped so that the product with log2(e) stays finite and a row whose keys
            // are all masked still comes out as the uniform average, as it does there. Out-of-range columns are masked below.
            const float* bp = bias_row + jj * kBN + 64 * half;
            const float sl = a.scale_log2;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float b0 = 64 * half + i < vc ? fmaxf(__ldg(bp + i) * 1.4426950408889634f, -3.0e38f) : 0.f;
              const float b1 = 64 * half + 32 + i < vc ? fmaxf(__ldg(bp + 32 + i) * 1.4426950408889634f, -3.0e38f) : 0.f;
              s0[i] = __float_as_uint(fmaf(__uint_as_float(s0[i]), sl, b0));
              s1[i] = __float_as_uint(fmaf(__uint_as_float(s1[i]), sl, b1));
            }
          }
        }
        if (vc < kBN) {
          mask_chunk(s0, 64 * half, vc);
          mask_chunk(s1, 64 * half + 32, vc);
        }
        bool o_ready = j == 0;
        if constexpr (!FAST) {
          // bf16: P and the fp32 accumulators carry 8 exponent bits, so the running maximum only has to keep exp2() in range, not
          // near 1. qn_max * kn(tile) * c2 bounds every scaled score of the tile (Cauchy-Schwarz); while it stays within
          // 2^kSkipMargin of the smallest first-tile row maximum, m_used is left alone and the tile needs no maximum pass, no
          // exchange between the column halves and no rescale decision. The test is identical in all 256 threads of the stream.
          bool exact = true;
          if constexpr (skip_guard) exact = j == 0 || qn_max * kn_cur * c2 - m_floor > kSkipMargin;
          if (exact) {
            // row maximum: own 64 columns, then exchange with the other half of the row
            const float lmax = fmaxf(max_chunk(s0, -INFINITY), max_chunk(s1, -INFINITY));
#if IEF_TC3_FINE_TRACE
            if (trace) tr2[0] = clock64();
#endif
            float* xm = xmax + ((j & 1) * 4 + t * 2) * 128;
            xm[half * 128 + row] = lmax;
            named_bar_sync(1 + t, 256);
#if IEF_TC3_FINE_TRACE
            if (trace) tr2[1] = clock64();
#endif
            const float tmax = fmaxf(lmax, xm[(half ^ 1) * 128 + row]);
            if (j == 0) {
              m_used = tmax;
              if constexpr (skip_guard) {
                // one-time: stream-wide max |q_row| and min first-tile row maximum through shared memory
                float qn = qn_row, mf = tmax * c2;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                  qn = fmaxf(qn, __shfl_xor_sync(0xffffffffu, qn, o));
                  mf = fminf(mf, __shfl_xor_sync(0xffffffffu, mf, o));
                }
                float* xw = xq + t * 16;
                if (lane == 0) {
                  xw[2 * (idx & 7)] = qn;
                  xw[2 * (idx & 7) + 1] = mf;
                }
                named_bar_sync(1 + t, 256);
                qn_max = xw[0];
                m_floor = xw[1];
#pragma unroll
                for (int w8 = 1; w8 < 8; ++w8) {
                  qn_max = fmaxf(qn_max, xw[2 * w8]);
                  m_floor = fminf(m_floor, xw[2 * w8 + 1]);
                }
              }
            } else {
              const float m_new = fmaxf(m_used, tmax);
              const bool need = (m_new - m_used) * c2 > kRescaleThreshold;
              if (__any_sync(0xffffffffu, need)) {  // both halves of a row see identical values and take the same decision
                mbar_wait(bar_o(t), (j - 1 + pb.strm(t)) & 1);
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
          }
        }
#if IEF_TC3_FINE_TRACE
        if (trace) tr2[2] = clock64();
#endif
        const float mc = FAST ? 0.f : m_used * c2;
        const float2 c2v = make_float2(c2, c2), nmc = make_float2(-mc, -mc);
        constexpr bool scale_in_turn = FAST && (SUMMMA ? IEF_TC3_FAST_SCALE_IN_TURN : IEF_TC3_FAST_SCALE_IN_TURN_D64);
        if constexpr (!scale_in_turn) {
          scale_chunk_mix<EMUL>(s0, c2v, nmc);   // FMA-pipe work (incl. the emulated share of the exponentials), outside the MUFU turn
          scale_chunk_mix<EMUL>(s1, c2v, nmc);
        }
#if IEF_TC3_FINE_TRACE
        if (trace) tr[6] = clock64();
#endif
        constexpr bool pv_wait_late = (SUMMMA ? IEF_TC3_PV_WAIT_LATE : IEF_TC3_PV_WAIT_LATE_D64) != 0;
        if (!o_ready && !(FAST && pv_wait_late)) {  // PV_t(j-1) must have finished reading P_t; wait for it BEFORE taking the exp turn, not inside it
          mbar_wait(bar_o(t), (j - 1 + pb.strm(t)) & 1);
          tc_fence_after();
        }
#if IEF_TC3_FINE_TRACE
        if (trace) tr[7] = clock64();
#endif
        if constexpr (ordered) {
          // ordered exp sections. Persistent: stream 0's first section of an item follows stream 1's last section of the previous one
          // (every grant is consumed, so no phase of a turn barrier can complete unobserved); stream 1 grants stream 0's very first
          // section before its first item, so stream 0 waits for phase j like stream 1
          if constexpr (PERSIST) mbar_wait(bar_x(t), j & 1);
          else if (t == 1 || j > 0) mbar_wait(bar_x(t), ((t == 1 ? j : j - 1) + pb.turn(t)) & 1);
        }
        if (trace) tr[3] = clock64();
#if IEF_TC3_HANDOVER_EARLY
        if constexpr (ordered) {  // experiment: the other stream may start as soon as this one HAS started (sections only staggered)
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_x(t ^ 1));
        }
#endif
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
        uint32_t u[16];
        if constexpr (scale_in_turn) scale_chunk_mix<EMUL>(s0, c2v, nmc);
        if constexpr (sum_mma) exp_pack_chunk_nosum<E, EMUL>(s0, u); else exp_pack_chunk_mix<E, EMUL>(s0, u, acc0, acc1);
        if constexpr (FAST && pv_wait_late) {
          // unshifted loop: only the first store of P needs PV_t(j-1) to have read the previous P_t; by now (half an exp section into
          // the turn) it practically always has, so the wait costs one barrier poll instead of ~200 cycles in front of the turn
          if (!o_ready) {
            mbar_wait(bar_o(t), (j - 1 + pb.strm(t)) & 1);
            tc_fence_after();
          }
        }
        tmem_st16(tP, u);
#if !IEF_TC3_HANDOVER_LATE && !IEF_TC3_HANDOVER_EARLY
        if constexpr (ordered) {
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_x(t ^ 1));  // hand the MUFU over one chunk early
        }
#endif
        if constexpr (scale_in_turn) scale_chunk_mix<EMUL>(s1, c2v, nmc);
        if constexpr (sum_mma) exp_pack_chunk_nosum<E, EMUL>(s1, u); else exp_pack_chunk_mix<E, EMUL>(s1, u, acc0, acc1);
        tmem_st16(tP + 16, u);
#if IEF_TC3_HANDOVER_LATE
        if constexpr (ordered) {
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_x(t ^ 1));
        }
#endif
        if (trace) tr[4] = clock64();
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(bar_p(t));
        if (trace) tr[5] = clock64();
        if constexpr (!sum_mma) {
          acc0 = fadd2(acc0, acc1);
          l += acc0.x + acc0.y;
        }
      }
      if (cta_trace && warp == 4 && lane == 0) a.dbg[1538] = clock64();
      if (my_nt > 0) {
        mbar_wait(bar_o(t), (my_nt - 1 + pb.strm(t)) & 1);
        tc_fence_after();
      }
      if (cta_trace && warp == 4 && lane == 0) a.dbg[1539] = clock64();
    };
    // row sum = first accumulator column after O (row-sum MMA) or the merge of the two column halves' partial sums
    auto row_sum = [&]() -> float {
      if constexpr (sum_mma) {
        uint32_t r[16];
        tmem_ld16(tO + a.dv_mma, r);
        tc_wait_ld();
        return my_nt > 0 ? __uint_as_float(r[0]) : 0.f;
      } else {
        float* xs = xsum + t * 2 * 128;
        xs[half * 128 + row] = l;
        named_bar_sync(1 + t, 256);
        return l + xs[(half ^ 1) * 128 + row];
      }
    };

    // ---- epilogue of one item: merge (split flavour), normalise, store; persistent CTAs then hand O_t back to the issuer
    auto epilogue = [&](float lrow) {
      const int nchunk_d = (a.d + 15) >> 4;
      float wmine = 1.f, wother = 0.f;
      if constexpr (SPLIT) {
        // ... then the two key halves. When stream 1's last PV has completed, every QK MMA of the CTA has completed too
        // (program order of the issuing thread), so the K ring is free to serve as staging; the V ring may still be in use.
        if (t == 1 && nt1 > 0) {
          if (half == 0) {
            xml[row] = m_used;
            xml[128 + row] = lrow;
          }
          for (int cc = half; cc < nchunk_d; cc += 2) {
            uint32_t r[16];
            tmem_ld16(tO + 16 * cc, r);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) stage_o[(16 * cc + i) * 128 + row] = __uint_as_float(r[i]);
          }
        }
        named_bar_sync(3, 512);
        if (t == 0 && nt1 > 0) {
          const float mB = xml[row], lB = xml[128 + row];
          const float m = fmaxf(m_used, mB);
          wmine = ief_exp2((m_used - m) * c2);
          wother = ief_exp2((mB - m) * c2);
          lrow = lrow * wmine + lB * wother;
        }
      }
      if (!SPLIT || t == 0) {
        if (a.lse_out != nullptr && half == 0) {  // row log-sum-exp (log2 units) for the stored-maps sweep
          const int r = (Cfg::kQTiles * qt + (SPLIT ? 0 : t)) * kBM + row;
          // split flavour: lrow already is the merged sum relative to the joint maximum m_used + log2(1 / wmine) / c2
          const float m_ref = (SPLIT && nt1 > 0) ? m_used * c2 - log2f(wmine) : m_used * c2;
          if (r < a.Nq) a.lse_out[((int64_t)b * a.H + h) * a.Nq + r] = m_ref + log2f(lrow);
        }
        const float inv = 1.f / lrow;
        wmine *= inv;
        wother *= inv;
        const int grow = (Cfg::kQTiles * qt + (SPLIT ? 0 : t)) * kBM + row;
        typename E::T* op = reinterpret_cast<typename E::T*>(a.o) + (int64_t)b * a.o_sb + (int64_t)grow * a.o_sn + (int64_t)h * a.o_sh;
        for (int cc = half; cc < nchunk_d; cc += 2) {
          uint32_t r[16];
          tmem_ld16(tO + 16 * cc, r);
          tc_wait_ld();
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            f[i] = __uint_as_float(r[i]) * wmine;
            if (SPLIT && nt1 > 0) f[i] = fmaf(stage_o[(16 * cc + i) * 128 + row], wother, f[i]);
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
      if constexpr (PERSIST) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_d(t));
      }
    };

    if constexpr (PERSIST) {
      if (t == 1) {  // stream 0's first exp section of the CTA has no predecessor: granted here
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_x(0));
      }
      int k = 0;
#pragma nounroll
      for (int e = 0; e < (NOMAX ? 2 : 1) * n_local; ++e) {
        bool exact;
#if IEF_TC3_TRACE
        const long long t_top = clock64();
#endif
        int it = next_entry(e, exact);
        if (it == -2) break;
        if (it < 0 || !set_item(it)) continue;
        item_setup();
        l = 0.f;
#if IEF_TC3_TRACE
        long long* trk = cta_trace && row == 0 && half == 0 && k < 16 ? a.dbg + 1600 + (k * 2 + t) * 4 : nullptr;  // item start | loop end | epilogue end | loop top
        if (trk) { trk[0] = clock64(); trk[3] = t_top; }
#endif
        if constexpr (NOMAX) {
          if (!exact) {
            m_used = 0.f;
            run_pass(BoolTag<true>{}, seq_bases(k), k);
          }
        }
        if (exact) {
          m_used = -INFINITY;
          run_pass(BoolTag<false>{}, seq_bases(k), k);
        }
#if IEF_TC3_TRACE
        if (trk) trk[1] = clock64();
#endif
        const float lrow = row_sum();
        if constexpr (NOMAX) {  // inf, NaN, zero or close to the edge of the fp32 / bf16 range: flag the item for the exact sweep
          if (!exact && !(lrow > kNoMaxLo && lrow < kNoMaxHi)) atomicOr(const_cast<uint32_t*>(redo_flag), 1u << it);
        }
        // the item's coordinates are recomputed here instead of being carried through the key loop in registers
        asm volatile("" : "+r"(it));
        set_item(it);
        epilogue(lrow);
#if IEF_TC3_TRACE
        if (trk) trk[2] = clock64();
#endif
        ++k;
      }
    } else {
      float lrow;
      if constexpr (NOMAX) {
        m_used = 0.f;
        run_pass(BoolTag<true>{}, PbZero{}, 0);
        lrow = row_sum();
        if (my_nt > 0 && !(lrow > kNoMaxLo && lrow < kNoMaxHi)) *redo_flag = 1;  // inf, NaN, zero or close to the edge of the fp32 / bf16 range
        named_bar_sync(0, kThreads);
        if (*redo_flag) {
          m_used = -INFINITY;
          l = 0.f;
          run_pass(BoolTag<false>{}, pass2_bases(), 0);
          lrow = row_sum();
        }
      } else {
        run_pass(BoolTag<false>{}, PbZero{}, 0);
        lrow = row_sum();
      }
      epilogue(lrow);
    }
  }
  if (cta_trace && warp == 4 && lane == 0) a.dbg[1540] = clock64();
  // a dependent launch must not COMPLETE before the launch it depends on: what follows in the stream waits for this grid only
  if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (cta_trace && threadIdx.x == 32) a.dbg[1541] = clock64();
}

template <int DTYPE, bool SPLIT, bool SUMMMA, int MAXMODE, bool BIAS, bool PERSIST>
int launch_tc3k(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const TcArgs& a, int grid, cudaStream_t st) {
  auto kern = attn_tc3_kernel<DTYPE, SPLIT, kDefaultEmul<SUMMMA>, SUMMMA, MAXMODE, BIAS, PERSIST>;
  IEF_CONFIG_SMEM(kern, Cfg3<SPLIT>::kSmemBytes);
  if (a.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg3<SPLIT>::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    IEF_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, mq, mk, mv, a));
  } else {
    kern<<<grid, kThreads, Cfg3<SPLIT>::kSmemBytes, st>>>(mq, mk, mv, a);
  }
  IEF_LAUNCH_OK("attn_tc3_kernel");
  return IEF_OK;
}

template <int DTYPE, bool SPLIT, bool SUMMMA, int MAXMODE, bool BIAS = false>
int launch_tc3s(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, TcArgs a, int first, int count, int nq_blocks, cudaStream_t st) {
  if (count <= 0) return IEF_OK;
  a.work_offset = first;
  a.nq_blocks = nq_blocks;
  a.n_items = count;
  // ceil(2^32 / divisor): floor(lin * rcp / 2^32) == lin / divisor for every lin < 2^32 / divisor
  const uint64_t div_h = (uint64_t)nq_blocks * (uint64_t)a.H, last = (uint64_t)first + (uint64_t)count;
  IEF_REQUIRE(last * div_h < (1ull << 32), IEF_ERR_UNSUPPORTED, "tcgen05 attention: %llu work items exceed the kernel's index arithmetic", (unsigned long long)last);
  static_assert(IEF_MAX_ROWS <= 64, "TcArgs::active_mask holds one bit per batch row");
  a.active_mask = 0;
  for (int i = 0; i < a.B && i < 64; ++i) a.active_mask |= (uint64_t)(a.rows.active[i] != 0) << i;
  a.rcp_nq = ((1ull << 32) + (uint64_t)nq_blocks - 1) / (uint64_t)nq_blocks;
  a.rcp_nq_h = ((1ull << 32) + div_h - 1) / div_h;
  // Persistent form (one CTA per SM walks count / #SM items): 256-row flavour with more items than SMs and an even number of key
  // tiles per item (see the kernel); everything else runs one item per CTA
  if constexpr (!SPLIT && MAXMODE != 1) {
    if (ief_attn_tc3_persistent(count, a.nt1 + a.nt2)) {
      const int sms = ief_sm_count();
      const int grid = sms * kMaxItemsPerCta >= count ? sms : ief_ceil_div(count, kMaxItemsPerCta);
      return launch_tc3k<DTYPE, SPLIT, SUMMMA, MAXMODE, BIAS, true>(mq, mk, mv, a, grid, st);
    }
  }
  return launch_tc3k<DTYPE, SPLIT, SUMMMA, MAXMODE, BIAS, false>(mq, mk, mv, a, count, st);
}

template <int DTYPE, bool SPLIT>
int launch_tc3(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, TcArgs a, int first, int count, int nq_blocks, cudaStream_t st) {
  if (a.key_bias != nullptr)  // masked MasaCtrl passes: the variant that adds a per-key bias before the maximum
    return a.sum_mma ? launch_tc3s<DTYPE, SPLIT, true, 0, true>(mq, mk, mv, a, first, count, nq_blocks, st)
                     : launch_tc3s<DTYPE, SPLIT, false, 0, true>(mq, mk, mv, a, first, count, nq_blocks, st);
  if constexpr (DTYPE == IEF_BF16) {
    if (a.knorm != nullptr)  // key-norm pre-pass available: the variant that skips provably harmless row-maximum passes (A/B only)
      return a.sum_mma ? launch_tc3s<DTYPE, SPLIT, true, 1>(mq, mk, mv, a, first, count, nq_blocks, st)
                       : launch_tc3s<DTYPE, SPLIT, false, 1>(mq, mk, mv, a, first, count, nq_blocks, st);
    if (a.nomax)             // optimistic unshifted softmax with the in-CTA exact second pass
      return a.sum_mma ? launch_tc3s<DTYPE, SPLIT, true, 2>(mq, mk, mv, a, first, count, nq_blocks, st)
                       : launch_tc3s<DTYPE, SPLIT, false, 2>(mq, mk, mv, a, first, count, nq_blocks, st);
  }
  return a.sum_mma ? launch_tc3s<DTYPE, SPLIT, true, 0>(mq, mk, mv, a, first, count, nq_blocks, st)
                   : launch_tc3s<DTYPE, SPLIT, false, 0>(mq, mk, mv, a, first, count, nq_blocks, st);
}

template <int DTYPE>
int launch_mode(const ief_attn_params* p, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const TcArgs& a, int mode, cudaStream_t st) {
  const int nqp = ief_ceil_div(p->Nq, 2 * kBM);         // 256-row blocks per (row, head)
  const int pairs = nqp * p->H * p->B;
  if (mode == 0) return launch_tc3<DTYPE, false>(mq, mk, mv, a, 0, pairs, nqp, st);
  if (mode == 1) return launch_tc3<DTYPE, true>(mq, mk, mv, a, 0, 2 * pairs, 2 * nqp, st);
  // hybrid: the full waves as 256-row CTAs, the remaining r < #SM/2 pairs as 2r half-length split-KV CTAs (pair L = split 2L, 2L+1)
  const int sms = ief_sm_count();
  const int full = (pairs / sms) * sms, rest = pairs - full;
  static int diag_part = -1;  // IEF_TC3_DIAG_PART=1|2: launch only the full waves / only the remainder (WRONG results; timing of the two parts)
  if (diag_part < 0) { const char* e = getenv("IEF_TC3_DIAG_PART"); diag_part = e ? atoi(e) : 0; }
  int rc = diag_part == 2 ? IEF_OK : launch_tc3<DTYPE, false>(mq, mk, mv, a, 0, full, nqp, st);
  if (rc != IEF_OK || diag_part == 1) return rc;
  static int env_pdl = -1;  // IEF_TC3_PDL=0: the remainder as an ordinary second launch (A/B)
  if (env_pdl < 0) { const char* e = getenv("IEF_TC3_PDL"); env_pdl = (e && e[0] == '0') ? 0 : 1; }
  TcArgs a2 = a;
  a2.pdl = (env_pdl && full > 0 && diag_part == 0) ? 1 : 0;
  return launch_tc3<DTYPE, true>(mq, mk, mv, a2, 2 * full, 2 * rest, 2 * nqp, st);
}

}  // namespace

bool ief_attn_tc3_persistent(long items, int nt) {
  static int env = -1;  // IEF_TC3_PERSIST=0: one item per CTA everywhere (A/B switch read once; nothing device-dependent is cached)
  if (env < 0) { const char* e = getenv("IEF_TC3_PERSIST"); env = (e && e[0] == '0') ? 0 : 1; }
  return kPersist && kFastOrdered && env && items > ief_sm_count() && (nt & 1) == 0;
}

int ief_attn_tc3_launch(const ief_attn_params* p, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const TcArgs& a, int mode,
                        cudaStream_t st) {
  return p->dtype == IEF_BF16 ? launch_mode<IEF_BF16>(p, mq, mk, mv, a, mode, st) : launch_mode<IEF_F16>(p, mq, mk, mv, a, mode, st);
}
