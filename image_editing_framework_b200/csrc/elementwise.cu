// HBM / latency-bound pieces of the hot path:
//   ief_cfg_ddim_step     CFG combine + DDIM (reverse) update   p2p/model/sd_utils.py:75-76, inversion/ddim.py:9-18
//   ief_store_accumulate  AttentionStore.between_steps "+="      p2p/model/attention_base.py:76-82
//   ief_local_blend       LocalBlend.__call__                    p2p/model/ptp_utils.py:20-32
//   ief_mask_blend        fg/bg output blend of masked MasaCtrl  masactrl/model/attention_control.py:176-177, 318-319
#include "ief_common.cuh"
#include <math.h>

namespace {

// ------------------------------------------------------------------------------------------------ DDIM step
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// Every operation is a separately rounded fp32 op in the reference's order (no FMA contraction), so an
// fp32 call is bit-identical to torch eager:  eps = eu + g*(ec-eu);  x0 = (x - sqrt(1-at)*eps)/sqrt(at);
// out = sqrt(ap)*x0 + sqrt(1-ap)*eps.
struct StepCoef { float g, sb_t, sa_t, sa_p, sb_p; int has_cond; };

__device__ __forceinline__ float ddim_one(float eu, float ec, float x, const StepCoef& c) {
  float eps = eu;
  if (c.has_cond) eps = __fadd_rn(eu, __fmul_rn(c.g, __fsub_rn(ec, eu)));
  const float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(c.sb_t, eps)), c.sa_t);
  return __fadd_rn(__fmul_rn(c.sa_p, x0), __fmul_rn(c.sb_p, eps));
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
cfg_ddim_kernel(const T* __restrict__ eu, const T* __restrict__ ec, const T* __restrict__ x, T* __restrict__ out, int64_t n, StepCoef c) {
  struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };
  const int64_t nv = n / VEC;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    Pack a = reinterpret_cast<const Pack*>(eu)[i], b, xx = reinterpret_cast<const Pack*>(x)[i], o;
    if (c.has_cond) b = reinterpret_cast<const Pack*>(ec)[i]; else b = a;
#pragma unroll
    for (int k = 0; k < VEC; ++k) o.v[k] = from_f<T>(ddim_one(to_f<T>(a.v[k]), to_f<T>(b.v[k]), to_f<T>(xx.v[k]), c));
    reinterpret_cast<Pack*>(out)[i] = o;
  }
  // scalar tail
  for (int64_t i = nv * VEC + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = from_f<T>(ddim_one(to_f<T>(eu[i]), c.has_cond ? to_f<T>(ec[i]) : 0.f, to_f<T>(x[i]), c));
}

template <typename T>
int launch_ddim(const void* eu, const void* ec, const void* x, void* out, int64_t n, const StepCoef& c, cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  const bool aligned = ((reinterpret_cast<uintptr_t>(eu) | reinterpret_cast<uintptr_t>(ec) | reinterpret_cast<uintptr_t>(x) |
                         reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const int64_t work = aligned ? (n + VEC - 1) / VEC : n;
  int blocks = (int)((work + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (aligned)
    cfg_ddim_kernel<T, VEC><<<blocks, 256, 0, st>>>((const T*)eu, (const T*)ec, (const T*)x, (T*)out, n, c);
  else
    cfg_ddim_kernel<T, 1><<<blocks, 256, 0, st>>>((const T*)eu, (const T*)ec, (const T*)x, (T*)out, n, c);
  IEF_LAUNCH_OK("cfg_ddim_kernel");
  return IEF_OK;
}

// ------------------------------------------------------------------------------------------------ store accumulate
struct AccumTable {
  float* dst[64];
  const float* src[64];
  int64_t numel[64];
};

__global__ void __launch_bounds__(256)
accumulate_kernel(const __grid_constant__ AccumTable t) {
  const int k = blockIdx.y;
  float* __restrict__ d = t.dst[k];
  const float* __restrict__ s = t.src[k];
  const int64_t n = t.numel[k];
  const bool aligned = ((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(s)) & 15) == 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (aligned) {
    const int64_t nv = n >> 2;
    float4* d4 = reinterpret_cast<float4*>(d);
    const float4* s4 = reinterpret_cast<const float4*>(s);
    int64_t i = i0;
    // two independent 16-byte streams per thread per iteration (memory-level parallelism)
    for (; i + stride < nv; i += 2 * stride) {
      float4 a0 = d4[i], b0 = __ldg(s4 + i), a1 = d4[i + stride], b1 = __ldg(s4 + i + stride);
      a0.x += b0.x; a0.y += b0.y; a0.z += b0.z; a0.w += b0.w;
      a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
      d4[i] = a0;
      d4[i + stride] = a1;
    }
    for (; i < nv; i += stride) {
      float4 a0 = d4[i], b0 = __ldg(s4 + i);
      a0.x += b0.x; a0.y += b0.y; a0.z += b0.z; a0.w += b0.w;
      d4[i] = a0;
    }
    for (int64_t j = (nv << 2) + i0; j < n; j += stride) d[j] += s[j];
  } else {
    for (int64_t j = i0; j < n; j += stride) d[j] += s[j];
  }
}

// ------------------------------------------------------------------------------------------------ LocalBlend
struct LbTable {
  const float* maps[16];
  int32_t heads[16];
  int32_t head_off[16];
  int32_t n_maps, n_prompts, res, n_words, total_heads;
  const float* word_alpha;
  float* work;  // [n_prompts, res*res], fully written by lb_reduce_kernel
};

// work[p][pix] = sum over all maps' heads and over the selected words of map[p*heads+h][pix][w] * alpha[p][w], DETERMINISTIC: a CTA owns
// 32 pixels of one prompt; its 8 warps split the (map, head) list round-robin in a fixed order and their partial sums are added in
// warp order — no atomics, so the thresholded mask (and with it an edit) is bit-reproducible run to run, like the reference's
// `sum(-1).mean(1)`.
__global__ void __launch_bounds__(256)
lb_reduce_kernel(const __grid_constant__ LbTable t) {
  __shared__ int s_words[IEF_MAX_WORDS];
  __shared__ float s_alpha[IEF_MAX_WORDS];
  __shared__ int s_nw;
  __shared__ float s_part[8][32];
  const int p = blockIdx.y;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int w = 0; w < t.n_words; ++w) {
      const float a = t.word_alpha[p * t.n_words + w];
      if (a != 0.f) { s_words[n] = w; s_alpha[n] = a; ++n; }
    }
    s_nw = n;
  }
  __syncthreads();
  const int npix = t.res * t.res, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pix = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (pix < npix) {
    for (int j = warp; j < t.total_heads; j += 8) {
      int mi = 0;
      while (mi + 1 < t.n_maps && j >= t.head_off[mi + 1]) ++mi;
      const int heads = t.heads[mi], hh = j - t.head_off[mi];
      const float* row = t.maps[mi] + (((int64_t)p * heads + hh) * npix + pix) * t.n_words;
      float part = 0.f;
      for (int k = 0; k < s_nw; ++k) part = fmaf(__ldg(row + s_words[k]), s_alpha[k], part);
      acc += part;
    }
  }
  s_part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && pix < npix) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += s_part[w][lane];
    t.work[p * npix + pix] = sum;
  }
}

struct LbApply {
  const float* work;
  float* x_t;
  float* mask_out;
  int32_t n_prompts, res, C, Hx, Wx, total_heads;
  float threshold;
};

// one CTA: 3x3 max-pool (stride 1, pad 1) of the head-mean map, per-prompt max normalisation, nearest
// upsampling to the latent size, threshold, OR over prompts, then x_t = x_t[0] + mask * (x_t - x_t[0]).
__global__ void __launch_bounds__(1024)
lb_apply_kernel(const __grid_constant__ LbApply a) {
  extern __shared__ float sm[];
  const int npix = a.res * a.res;
  float* pooled = sm;                       // [n_prompts][npix]
  float* pmax = sm + a.n_prompts * npix;    // [n_prompts]
  __shared__ float red[32];
  const float inv_h = 1.f / (float)a.total_heads;
  for (int p = 0; p < a.n_prompts; ++p) {
    float lmax = -INFINITY;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
      const int y = i / a.res, x = i - y * a.res;
      float v = -INFINITY;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int yy = y + dy, xx = x + dx;
          if (yy >= 0 && yy < a.res && xx >= 0 && xx < a.res) v = fmaxf(v, a.work[p * npix + yy * a.res + xx] * inv_h);
        }
      pooled[p * npix + i] = v;
      lmax = fmaxf(lmax, v);
    }
    for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lmax;
    __syncthreads();
    if (threadIdx.x == 0) {
      float mm = -INFINITY;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mm = fmaxf(mm, red[w]);
      pmax[p] = mm;
    }
    __syncthreads();
  }
  const int HW = a.Hx * a.Wx;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const int y = i / a.Wx, x = i - y * a.Wx;
    // torch nearest: src = floor(dst * in / out)
    const int sy = min((int)floorf((float)y * ((float)a.res / (float)a.Hx)), a.res - 1);
    const int sx = min((int)floorf((float)x * ((float)a.res / (float)a.Wx)), a.res - 1);
    bool any = false;
    for (int p = 0; p < a.n_prompts; ++p) {
      const bool mk = __fdiv_rn(pooled[p * npix + sy * a.res + sx], pmax[p]) > a.threshold;
      if (a.mask_out) a.mask_out[p * HW + i] = mk ? 1.f : 0.f;
      any |= mk;
    }
    const float mf = any ? 1.f : 0.f;
    for (int c = 0; c < a.C; ++c) {
      const float x0 = a.x_t[(int64_t)c * HW + i];
      for (int p = 1; p < a.n_prompts; ++p) {
        float* px = a.x_t + ((int64_t)p * a.C + c) * HW + i;
        *px = __fadd_rn(x0, __fmul_rn(mf, __fsub_rn(*px, x0)));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ fg/bg blend
struct BlendRows { uint8_t on[IEF_MAX_ROWS]; };

// fg = fg * w[n] + bg * (1 - w[n]) over [B, N, C]; a thread owns 8 (16-bit) or 4 (fp32) consecutive channels: 16-byte accesses.
// The three products/sums are rounded separately in fp32, as torch evaluates `fg * mask + bg * (1 - mask)`.
template <typename T>
__global__ void __launch_bounds__(256) mask_blend_kernel(T* fg, const T* bg, const float* w, int64_t N, int64_t C, int64_t vec_per_row,
                                                         int64_t total, BlendRows rows) {
  constexpr int V = 16 / sizeof(T);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / vec_per_row;  // b * N + n
    const int64_t b = row / N, n = row - b * N;
    if (!rows.on[b]) continue;
    const float wn = __ldg(w + n), wb = __fsub_rn(1.f, wn);
    const int64_t off = row * C + (i - row * vec_per_row) * V;
    uint4 f4 = *reinterpret_cast<const uint4*>(fg + off);
    const uint4 b4 = __ldg(reinterpret_cast<const uint4*>(bg + off));
    T* fe = reinterpret_cast<T*>(&f4);
    const T* be = reinterpret_cast<const T*>(&b4);
#pragma unroll
    for (int e = 0; e < V; ++e) fe[e] = from_f<T>(__fadd_rn(__fmul_rn(to_f<T>(fe[e]), wn), __fmul_rn(to_f<T>(be[e]), wb)));
    *reinterpret_cast<uint4*>(fg + off) = f4;
  }
}

template <typename T>
int launch_blend(void* fg, const void* bg, const float* w, int32_t B, int64_t N, int64_t C, const BlendRows& rows, cudaStream_t st) {
  constexpr int V = 16 / sizeof(T);
  const int64_t vec_per_row = C / V, total = (int64_t)B * N * vec_per_row;
  const int64_t blocks = (total + 255) / 256;
  const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
  mask_blend_kernel<T><<<grid, 256, 0, st>>>(static_cast<T*>(fg), static_cast<const T*>(bg), w, N, C, vec_per_row, total, rows);
  IEF_LAUNCH_OK("mask_blend_kernel");
  return IEF_OK;
}

}  // namespace

extern "C" int ief_mask_blend(void* fg, const void* bg, const float* w, int32_t dtype, int32_t B, int64_t N, int64_t C,
                              const uint8_t* row_mask, void* stream) {
  IEF_REQUIRE(fg && bg && w, IEF_ERR_INVALID, "ief_mask_blend: null pointer");
  IEF_REQUIRE(B >= 1 && B <= IEF_MAX_ROWS && N >= 1 && C >= 1, IEF_ERR_INVALID, "ief_mask_blend: bad shape B=%d N=%lld C=%lld", B, (long long)N,
              (long long)C);
  const int V = dtype == IEF_F32 ? 4 : 8;
  IEF_REQUIRE(C % V == 0, IEF_ERR_UNSUPPORTED, "ief_mask_blend: C=%lld must be a multiple of %d", (long long)C, V);
  IEF_REQUIRE(((reinterpret_cast<uintptr_t>(fg) | reinterpret_cast<uintptr_t>(bg)) & 15) == 0, IEF_ERR_INVALID, "ief_mask_blend: pointers must be 16-byte aligned");
  BlendRows rows;
  for (int i = 0; i < IEF_MAX_ROWS; ++i) rows.on[i] = i < B ? (row_mask ? row_mask[i] : 1) : 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dtype) {
    case IEF_F32: return launch_blend<float>(fg, bg, w, B, N, C, rows, st);
    case IEF_BF16: return launch_blend<__nv_bfloat16>(fg, bg, w, B, N, C, rows, st);
    case IEF_F16: return launch_blend<__half>(fg, bg, w, B, N, C, rows, st);
    default: ief_set_error("ief_mask_blend: unknown dtype %d", dtype); return IEF_ERR_UNSUPPORTED;
  }
}

extern "C" int ief_cfg_ddim_step(const void* eps_uncond, const void* eps_cond, const void* x, void* x_out, int64_t n, int32_t dtype,
                                 float guidance, float alpha_t, float alpha_prev, void* stream) {
  IEF_REQUIRE(eps_uncond && x && x_out, IEF_ERR_INVALID, "ief_cfg_ddim_step: null pointer");
  IEF_REQUIRE(n >= 0, IEF_ERR_INVALID, "ief_cfg_ddim_step: negative element count");
  IEF_REQUIRE(alpha_t > 0.f && alpha_t <= 1.f && alpha_prev > 0.f && alpha_prev <= 1.f, IEF_ERR_INVALID,
              "ief_cfg_ddim_step: alphas_cumprod values must lie in (0,1], got %g and %g", alpha_t, alpha_prev);
  if (n == 0) return IEF_OK;
  StepCoef c;
  c.g = guidance;
  c.has_cond = eps_cond != nullptr;
  // fp32 scalar math exactly as torch does on 0-dim fp32 tensors: (1-a)**0.5, a**0.5
  c.sb_t = sqrtf(1.f - alpha_t);
  c.sa_t = sqrtf(alpha_t);
  c.sa_p = sqrtf(alpha_prev);
  c.sb_p = sqrtf(1.f - alpha_prev);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!c.has_cond) eps_cond = eps_uncond;
  switch (dtype) {
    case IEF_F32: return launch_ddim<float>(eps_uncond, eps_cond, x, x_out, n, c, st);
    case IEF_BF16: return launch_ddim<__nv_bfloat16>(eps_uncond, eps_cond, x, x_out, n, c, st);
    case IEF_F16: return launch_ddim<__half>(eps_uncond, eps_cond, x, x_out, n, c, st);
    default: ief_set_error("ief_cfg_ddim_step: unknown dtype %d", dtype); return IEF_ERR_UNSUPPORTED;
  }
}

extern "C" int ief_store_accumulate(float* const* dst, const float* const* src, const int64_t* numel, int32_t n, void* stream) {
  IEF_REQUIRE(n >= 0 && n <= 64, IEF_ERR_INVALID, "ief_store_accumulate: n=%d outside [0,64]", n);
  if (n == 0) return IEF_OK;
  IEF_REQUIRE(dst && src && numel, IEF_ERR_INVALID, "ief_store_accumulate: null table");
  AccumTable t;
  int64_t maxn = 0;
  for (int i = 0; i < n; ++i) {
    IEF_REQUIRE(dst[i] && src[i] && numel[i] >= 0, IEF_ERR_INVALID, "ief_store_accumulate: bad entry %d", i);
    t.dst[i] = dst[i]; t.src[i] = src[i]; t.numel[i] = numel[i];
    if (numel[i] > maxn) maxn = numel[i];
  }
  if (maxn == 0) return IEF_OK;
  int64_t bx = (maxn / 4 + 2 * 256 - 1) / (2 * 256);
  if (bx < 1) bx = 1;
  if (bx > 148 * 4) bx = 148 * 4;
  dim3 grid((unsigned)bx, (unsigned)n);
  accumulate_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t);
  IEF_LAUNCH_OK("accumulate_kernel");
  return IEF_OK;
}

extern "C" int ief_local_blend(const ief_local_blend_params* p, void* stream) {
  IEF_REQUIRE(p && p->maps && p->map_heads && p->word_alpha && p->x_t && p->workspace, IEF_ERR_INVALID, "ief_local_blend: null pointer");
  IEF_REQUIRE(p->n_maps >= 1 && p->n_maps <= 16, IEF_ERR_INVALID, "ief_local_blend: n_maps=%d outside [1,16]", p->n_maps);
  // the reference's final broadcast (mask[:1] + mask[1:]) against x_t only type-checks for two prompts
  IEF_REQUIRE(p->n_prompts == 2, IEF_ERR_UNSUPPORTED, "ief_local_blend: n_prompts=%d, the reference broadcast only admits 2", p->n_prompts);
  IEF_REQUIRE(p->res >= 1 && p->res <= 64 && p->n_words >= 1 && p->n_words <= IEF_MAX_WORDS, IEF_ERR_UNSUPPORTED, "ief_local_blend: bad res/n_words");
  IEF_REQUIRE(p->C >= 1 && p->Hx >= 1 && p->Wx >= 1, IEF_ERR_INVALID, "ief_local_blend: bad latent shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LbTable t;
  int total = 0;
  for (int i = 0; i < p->n_maps; ++i) {
    IEF_REQUIRE(p->maps[i] && p->map_heads[i] >= 1, IEF_ERR_INVALID, "ief_local_blend: bad map %d", i);
    t.maps[i] = p->maps[i]; t.heads[i] = p->map_heads[i]; t.head_off[i] = total;
    total += p->map_heads[i];
  }
  t.n_maps = p->n_maps; t.n_prompts = p->n_prompts; t.res = p->res; t.n_words = p->n_words; t.total_heads = total;
  t.word_alpha = p->word_alpha; t.work = p->workspace;
  const int npix = p->res * p->res;
  dim3 grid((npix + 31) / 32, p->n_prompts);
  lb_reduce_kernel<<<grid, 256, 0, st>>>(t);
  IEF_LAUNCH_OK("lb_reduce_kernel");
  LbApply a;
  a.work = p->workspace; a.x_t = p->x_t; a.mask_out = p->mask_out; a.n_prompts = p->n_prompts; a.res = p->res;
  a.C = p->C; a.Hx = p->Hx; a.Wx = p->Wx; a.total_heads = total; a.threshold = p->threshold;
  const int smem = sizeof(float) * (p->n_prompts * npix + p->n_prompts);
  lb_apply_kernel<<<1, 1024, smem, st>>>(a);
  IEF_LAUNCH_OK("lb_apply_kernel");
  return IEF_OK;
}
