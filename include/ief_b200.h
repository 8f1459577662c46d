/*
 * ief_b200.h — C ABI of libief_b200.so: the B200 (sm_100a) controlled-attention hot path.
 *
 * The reference (AY-Liu/Image-Editing-Framework) is pure Python and has NO FFI; each entry
 * point below replaces a block of torch arithmetic inside the reference's attention
 * closures / controllers (cited per function, paths relative to the reference root).
 * The reference-side binding is a ctypes stub, shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C, no torch types; every pointer is a DEVICE pointer unless it says "host".
 *   - the caller owns every buffer; the library allocates nothing persistent, keeps no
 *     pointer after return, launches only on `stream` (a cudaStream_t) and never syncs.
 *   - return 0 on success, negative ief_status on failure; ief_last_error() returns a
 *     thread-local message. Unsupported shapes/dtypes are errors — there is no CPU
 *     fallback and no silent slow path.
 *   - tensors are addressed as [batch row][token][head][channel] with explicit strides
 *     (in ELEMENTS) and a contiguous channel axis, so the [B,N,H*d] output of nn.Linear
 *     is consumed in place: no head_to_batch_dim / batch_to_head_dim copies
 *     (p2p/model/register.py:43-45,51; masactrl/model/register.py:33; pnp/model/register.py:54-63,76).
 */
#ifndef IEF_B200_H_
#define IEF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IEF_ABI_VERSION 5
#define IEF_MAX_ROWS 64  /* max UNet batch rows per call (reference uses 1, 2 or 4) */
#define IEF_MAX_WORDS 77 /* CLIP context length, p2p/model/ptp_utils.py:8 MAX_NUM_WORDS */

typedef enum ief_status {
  IEF_OK = 0,
  IEF_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, misalignment) */
  IEF_ERR_UNSUPPORTED = -2, /* shape / dtype / feature outside what the kernels implement */
  IEF_ERR_CUDA = -3,        /* a CUDA runtime / driver call failed */
  IEF_ERR_NO_DEVICE = -4    /* not running on an sm_100 device */
} ief_status;

typedef enum ief_dtype { IEF_BF16 = 0, IEF_F16 = 1, IEF_F32 = 2 } ief_dtype;

/* which kernel family serves ief_attn_fwd */
typedef enum ief_attn_impl {
  IEF_IMPL_AUTO = 0,   /* tcgen05 when the shape allows it, else mma.sync */
  IEF_IMPL_MMA = 1,    /* warp-level mma.sync flash kernel (any Nq/Nk, d%8==0, d<=160) */
  IEF_IMPL_TCGEN05 = 2 /* TMA + tcgen05.mma + TMEM flash kernel */
} ief_attn_impl;

/* A strided [rows, tokens, heads, d] view; strides in elements, channel stride 1. */
typedef struct ief_tensor4 {
  void* ptr;
  int64_t stride_b; /* batch-row stride  */
  int64_t stride_n; /* token stride      */
  int64_t stride_h; /* head stride       */
} ief_tensor4;

/*
 * ief_attn_fwd — O[b] = softmax(scale * Q[q_src[b]] K[k_src[b]]^T) V[v_src[b]]
 * one flash-style pass, no materialised probabilities. The per-row source indices
 * express every self-attention edit of the reference:
 *   P2P self replace  (p2p/model/attention_base.py:132-136):   row 3 -> (q,k,v) = (2,2,3)
 *   MasaCtrl mutual   (masactrl/model/attention_control.py:37-68): row 1 -> (1,0,0), row 3 -> (3,2,2)
 *   PnP q/k injection (pnp/model/register.py:44-52):            rows 1,3 -> (2,2,self)
 *   plain attention   (p2p/model/register.py:47-50, pix2pix-zero/model/attention_control.py:43-48)
 * k_src2/v_src2 >= 0 appends a second key/value block (MasaCtrl Union,
 * masactrl/model/attention_control.py:100-101): keys = [K[k_src]; K[k_src2]].
 * probs_out (optional, fp32 [n_stored,H,Nq,Nk_total] contiguous) receives the normalised
 * probabilities (AttentionStore self maps N<=1024, p2p/model/attention_base.py:64-68;
 * pix2pix-zero attn_probs, pix2pix-zero/model/attention_control.py:46); when set the
 * mma.sync two-sweep kernel is used. probs_accum != 0 adds into probs_out instead of
 * overwriting (AttentionStore.between_steps, p2p/model/attention_base.py:76-82).
 */
typedef struct ief_attn_params {
  ief_tensor4 q, k, v, o; /* o is written; dtype for all four = dtype */
  int32_t dtype;          /* IEF_BF16 or IEF_F16 */
  int32_t B, H, Nq, Nk, d;
  float scale;            /* dim_head ** -0.5 in the reference */
  int32_t impl;           /* ief_attn_impl */
  const int32_t* q_src;   /* HOST arrays of length B, NULL = identity */
  const int32_t* k_src;
  const int32_t* v_src;
  const int32_t* k_src2;  /* HOST, NULL or entries <0 = no second block */
  const int32_t* v_src2;
  float* probs_out;       /* optional fp32 [n_stored,H,Nq,Nk_total] */
  int32_t probs_accum;
  const int32_t* probs_slot; /* HOST [B] or NULL (= b): row b's maps go to probs_out[probs_slot[b]]; <0 = not stored */
  const uint8_t* row_mask; /* HOST [B] or NULL: rows with 0 are skipped (output untouched) */
  const float* key_bias;   /* device fp32 [n_bias, Nk] or NULL: additive per-key bias on the scaled scores,
                              softmax(scale*QK^T + key_bias[bias_sel[b]]) — the fore-/background key masks of
                              MutualSelfAttentionControlMask / MaskAuto (masactrl/model/attention_control.py:139-147,
                              238-246; "masked" keys carry finfo.min, as there). Served by the mma kernel and, for head_dim <= 64, by a
                              variant of the tcgen05 kernel; no second K/V block. */
  const int32_t* bias_sel; /* HOST [B] or NULL (= no bias): row b uses key_bias[bias_sel[b]]; <0 = no bias for that row */
  int32_t n_bias;
  void* workspace;         /* optional device scratch of >= ief_attn_workspace_bytes(p) bytes (16-byte aligned), or NULL. With it (a) the
                              bf16 tcgen05 kernel runs a small pre-pass (max key norm per 128-key tile) whose Cauchy-Schwarz score
                              bound lets most tiles skip the running-maximum pass, (b) a call with probs_out on a layer the tcgen05
                              kernels serve gets O and the row log-sum-exp from them and writes the maps in one further sweep
                              instead of the two-sweep mma kernel. Same results either way; contents only live during the call. */
  int64_t workspace_bytes;
} ief_attn_params;

/* Scratch size ief_attn_fwd can make use of for these parameters (0 = none). */
int64_t ief_attn_workspace_bytes(const ief_attn_params* p);

int ief_attn_fwd(const ief_attn_params* p, void* stream);

/*
 * ief_cross_attn_edit_fwd — 77-key cross-attention with the P2P probability edit fused.
 * For every batch row b:  P_b = softmax(scale * Q[b] K[b]^T)           (Nk <= 80 keys)
 * rows with base_row[b] >= 0 are "replace" rows of AttentionControlEdit.forward
 * (p2p/model/attention_base.py:113-125) with base = P[base_row[b]]:
 *     E  = edit(base, P_b)                 edit selected by `mode`:
 *            IEF_EDIT_REPLACE  base @ mapper[e]                        (attention_control.py:15-16)
 *            IEF_EDIT_REFINE   base[:, idx[e]] * ra[e] + P_b * (1-ra[e]) (attention_control.py:28-31)
 *            IEF_EDIT_NONE     base
 *          then, if equalizer != NULL:  E *= equalizer[e]              (attention_control.py:42-46)
 *     P_b' = E * alpha[e] + (1 - alpha[e]) * P_b     alpha = cross_replace_alpha[cur_step] (:119-120)
 * where e = edit_slot[b] indexes the per-target-prompt tables. Rows with base_row < 0
 * keep P_b. Then O[b] = P_b' V[b].
 * probs_out (optional fp32 [B,H,Nq,Nk]) receives P_b' (post-edit, what a reader of
 * AttentionStore sees, attention_base.py:67 stores by alias) of every row with store_slot[b] >= 0
 * (probs_out is [n_stored, H, Nq, Nk]), overwriting or accumulating (probs_accum).
 * Kernels: with >= 128 queries the plain, unstored rows run on cross_tc.cu and the edited / stored rows on
 * cross_tc_edit.cu (both tcgen05 / TMEM / TMA; the edit needs the sparse form of a REPLACE mapper); fewer queries
 * or a dense-only mapper take the mma.sync kernel (cross_attn.cu). Same results within rounding either way.
 */
typedef enum ief_edit_mode { IEF_EDIT_NONE = 0, IEF_EDIT_REPLACE = 1, IEF_EDIT_REFINE = 2 } ief_edit_mode;

typedef struct ief_cross_params {
  ief_tensor4 q, k, v, o;
  int32_t dtype;
  int32_t B, H, Nq, Nk, d;
  float scale;
  int32_t mode;             /* ief_edit_mode */
  const int32_t* base_row;  /* HOST [B]; NULL = no edit on any row */
  const int32_t* edit_slot; /* HOST [B]; NULL = slot 0 */
  int32_t n_slots;
  const float* mapper;       /* device fp32 [n_slots, Nk, Nk]  (REPLACE, dense form) */
  const int32_t* mapper_nz_idx; /* optional sparse form of mapper: device int32 [n_slots, Nk, 8], for target token n the source
                                   tokens w with mapper[w][n] != 0 in ascending order, padded with -1 (a word swap touches a
                                   handful of tokens). When given (with mapper_nz_w) it is used instead of the dense form and
                                   yields bit-identical results; a column with more than 8 non-zeros needs the dense form. */
  const float* mapper_nz_w;     /* device fp32 [n_slots, Nk, 8]: the matching mapper[w][n] values */
  const int32_t* mapper_idx; /* device int32 [n_slots, Nk], may hold -1 (REFINE; -1 wraps to Nk-1 as torch indexing does) */
  const float* refine_alpha; /* device fp32 [n_slots, Nk]      (REFINE) */
  const float* equalizer;    /* device fp32 [n_slots, Nk] or NULL */
  const float* step_alpha;   /* device fp32 [n_slots, Nk]: cross_replace_alpha[cur_step] */
  float* probs_out;          /* optional fp32 [n_stored,H,Nq,Nk] */
  int32_t probs_accum;
  const int32_t* store_slot; /* HOST [B] or NULL (= b): row b's maps go to probs_out[store_slot[b]]; <0 = not stored */
} ief_cross_params;

int ief_cross_attn_edit_fwd(const ief_cross_params* p, void* stream);

/*
 * ief_cross_attn_bwd — backward of the (un-edited) <= 80-key cross-attention, for Pix2Pix-zero's guidance pass
 * (pix2pix-zero/model/sd_utils.py:163-174: loss = sum over attn2 layers of ||attn_probs - ref||^2, differentiated
 * with respect to the latents; forward: pix2pix-zero/model/attention_control.py:43-49).
 *     P = softmax(scale Q K^T) is recomputed;  dP = dO V^T + dprobs;  dS = P * (dP - rowsum(P * dP)) * scale;  dQ = dS K
 * q, k, v, dout, dq: strided [B, N, H, d] views (bf16 / fp16); dprobs: fp32 [B*H, Nq, Nk], the gradient arriving directly on
 * the probabilities (NULL = none); ds_out: fp32 [B*H, Nq, Nk] or NULL — when given, dS is stored so that the caller can
 * form dK = dS^T Q and dV = P^T dO (two 77 x d reductions over the query axis) with batched GEMMs.
 */
typedef struct ief_cross_bwd_params {
  ief_tensor4 q, k, v, dout, dq;
  int32_t dtype;
  int32_t B, H, Nq, Nk, d;
  float scale;
  const float* dprobs;
  float* ds_out;
} ief_cross_bwd_params;

int ief_cross_attn_bwd(const ief_cross_bwd_params* p, void* stream);

/*
 * ief_store_accumulate — dst[i][:] += src[i][:] for n tensors in ONE launch
 * (AttentionStore.between_steps python loop, p2p/model/attention_base.py:76-82;
 *  masactrl/model/attention_base.py:47-55). dst/src are HOST arrays of device fp32
 * pointers, numel a HOST array of element counts. n <= 64.
 */
int ief_store_accumulate(float* const* dst, const float* const* src, const int64_t* numel, int32_t n, void* stream);

/*
 * ief_local_blend — LocalBlend.__call__ (p2p/model/ptp_utils.py:20-32).
 * maps: n_maps device fp32 tensors, each [n_prompts*heads_i, res*res, 77] (HOST array of
 * pointers, heads_i in map_heads) — the summed store the reference passes at
 * attention_base.py:129 (the step average would give the same mask: it is max-normalised).
 * word_alpha: device fp32 [n_prompts,77]
 * (alpha_layers). x_t: fp32 [n_prompts,C,Hx,Wx], updated in place:
 *     x_t = x_t[0] + mask * (x_t - x_t[0]),  mask = OR over prompts of (norm maxpool3x3 map > threshold)
 */
typedef struct ief_local_blend_params {
  const float* const* maps; /* HOST array [n_maps] of device pointers */
  const int32_t* map_heads; /* HOST [n_maps] */
  int32_t n_maps, n_prompts, res, n_words;
  const float* word_alpha;
  float threshold;
  float* x_t;
  int32_t C, Hx, Wx;
  float* workspace; /* device fp32 [n_prompts * res * res] scratch (contents ignored) */
  float* mask_out;  /* optional device fp32 [n_prompts, Hx*Wx]: per-prompt mask before the OR (tests) */
} ief_local_blend_params;

int ief_local_blend(const ief_local_blend_params* p, void* stream);

/*
 * ief_cfg_ddim_step — classifier-free guidance + DDIM update in one launch
 * (p2p/model/sd_utils.py:75-76 and the same two lines in every driver; diffusers
 * DDIMScheduler.step, eta=0, epsilon prediction, no clipping):
 *     eps = eps_u + g (eps_c - eps_u)
 *     x0  = (x - sqrt(1-a_t) eps) / sqrt(a_t);   x' = sqrt(a_p) x0 + sqrt(1-a_p) eps
 * eps_uncond/eps_cond/x/x_out: same dtype (`dtype`: IEF_F32, IEF_BF16 or IEF_F16), n elements.
 * eps_cond == NULL means no guidance (eps = eps_uncond): the DDIM reverse step
 * (ddim_reverse, inversion/ddim.py:9-18) is the same formula with a_t/a_prev swapped by the caller.
 */
int ief_cfg_ddim_step(const void* eps_uncond, const void* eps_cond, const void* x, void* x_out, int64_t n, int32_t dtype,
                      float guidance, float alpha_t, float alpha_prev, void* stream);

/*
 * ief_mask_blend — spatial fore-/background blend of two attention outputs
 * (masactrl/model/attention_control.py:176-177, 318-319):
 *     fg[b,n,:] = fg[b,n,:] * w[n] + bg[b,n,:] * (1 - w[n])      for rows b with row_mask[b] != 0
 * fg (in/out) and bg: contiguous [B, N, C] of `dtype` (IEF_BF16 / IEF_F16 / IEF_F32); w: device fp32 [N];
 * row_mask: HOST [B] or NULL (= all rows). Arithmetic in fp32, one rounding on the store.
 */
int ief_mask_blend(void* fg, const void* bg, const float* w, int32_t dtype, int32_t B, int64_t N, int64_t C,
                   const uint8_t* row_mask, void* stream);

/*
 * ief_umma_probe — diagnostic: one CTA runs TMA -> smem -> tcgen05.mma -> TMEM -> global with
 * caller-supplied descriptor fields, so tests can pin the sm_100a encodings the attention
 * kernel relies on. D[128,N] (fp32) = A[128,K] * B, A from smem (K-major) or TMEM.
 */
typedef struct ief_umma_probe_params {
  const void* a;     /* bf16 [128, K] row-major */
  const void* b;     /* bf16: b_mn_major==0 -> [N, K] row-major; ==1 -> [K, N] row-major */
  float* d;          /* fp32 [128, N] */
  int32_t N, K;      /* N%16==0, N<=256; K%16==0, K<=128 */
  int32_t b_mn_major;
  int32_t a_from_tmem;
  int32_t dtype;     /* IEF_BF16 / IEF_F16 */
} ief_umma_probe_params;

int ief_umma_probe(const ief_umma_probe_params* p, void* stream);

/* library / device introspection */
int ief_abi_version(void);
const char* ief_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t ief_launch_count(void);
/* name of the kernel family the last ief_attn_fwd call dispatched to ("tcgen05" / "mma") */
const char* ief_last_attn_impl(void);
/* the same for the last ief_cross_attn_edit_fwd call of this thread: "tcgen05" (cross_tc.cu only), "tcgen05-edit"
 * (cross_tc_edit.cu, plus cross_tc.cu for the call's plain rows), "mma" (cross_attn.cu) */
const char* ief_last_cross_impl(void);
/* 0 when a CUDA device with compute capability 10.x is current */
int ief_check_device(void);

#ifdef __cplusplus
}
#endif
#endif /* IEF_B200_H_ */
