// C-ABI glue of libief_b200.so: error state, dispatch of ief_attn_fwd, introspection.
#include "ief_common.cuh"
#include <atomic>
#include <string.h>

namespace {
thread_local char g_err[512] = "";
thread_local const char* g_last_impl = "none";
std::atomic<int64_t> g_launches{0};
#if defined(IEF_TC3_TRACE) && IEF_TC3_TRACE
long long* g_trace = nullptr;  // debug builds only (IEF_EXTRA_NVCC_FLAGS=-DIEF_TC3_TRACE=1): the shipped library keeps no mutable global state
#endif
}  // namespace

void ief_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
#if defined(IEF_TC3_TRACE) && IEF_TC3_TRACE
long long* ief_debug_trace_buffer() { return g_trace; }
// Debug builds only (not part of the ABI header): device buffer of >= 1600 int64 receiving clock64 phase stamps of CTA 0.
extern "C" void ief_debug_set_trace_buffer(long long* p) { g_trace = p; }
#else
long long* ief_debug_trace_buffer() { return nullptr; }
#endif
void ief_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int ief_abi_version(void) { return IEF_ABI_VERSION; }
extern "C" const char* ief_last_error(void) { return g_err; }
extern "C" int64_t ief_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" const char* ief_last_attn_impl(void) { return g_last_impl; }

extern "C" int ief_check_device(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    ief_set_error("no CUDA device is current");
    return IEF_ERR_NO_DEVICE;
  }
  int major = 0;
  IEF_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  IEF_REQUIRE(major == 10, IEF_ERR_NO_DEVICE, "device %d has compute capability %d.x; libief_b200 is built for sm_100a only", dev, major);
  return IEF_OK;
}

extern "C" int ief_attn_fwd(const ief_attn_params* p, void* stream) {
  IEF_REQUIRE(p != nullptr, IEF_ERR_INVALID, "ief_attn_fwd: null params");
  IEF_REQUIRE(p->q.ptr && p->k.ptr && p->v.ptr && p->o.ptr, IEF_ERR_INVALID, "ief_attn_fwd: null tensor pointer");
  IEF_REQUIRE(p->B >= 1 && p->B <= IEF_MAX_ROWS, IEF_ERR_INVALID, "ief_attn_fwd: B=%d outside [1,%d]", p->B, IEF_MAX_ROWS);
  IEF_REQUIRE(p->H >= 1 && p->H <= 65535, IEF_ERR_INVALID, "ief_attn_fwd: H=%d", p->H);
  IEF_REQUIRE(p->Nq >= 1 && p->Nk >= 1, IEF_ERR_INVALID, "ief_attn_fwd: empty sequence (Nq=%d, Nk=%d)", p->Nq, p->Nk);
  IEF_REQUIRE(p->d >= 8, IEF_ERR_INVALID, "ief_attn_fwd: head_dim %d", p->d);
  IEF_REQUIRE(p->scale > 0.f, IEF_ERR_INVALID, "ief_attn_fwd: scale must be positive");
  IefRowTable rows;
  bool any_bias = false;
  for (int i = 0; i < p->B; ++i) {
    rows.q[i] = p->q_src ? p->q_src[i] : i;
    rows.k[i] = p->k_src ? p->k_src[i] : i;
    rows.v[i] = p->v_src ? p->v_src[i] : i;
    rows.k2[i] = p->k_src2 ? p->k_src2[i] : -1;
    rows.v2[i] = p->v_src2 ? p->v_src2[i] : -1;
    rows.pslot[i] = p->probs_slot ? p->probs_slot[i] : i;
    rows.active[i] = p->row_mask ? p->row_mask[i] : 1;
    rows.bias[i] = (p->key_bias && p->bias_sel) ? p->bias_sel[i] : -1;
    IEF_REQUIRE(rows.bias[i] < p->n_bias, IEF_ERR_INVALID, "ief_attn_fwd: bias_sel[%d]=%d but n_bias=%d", i, rows.bias[i], p->n_bias);
    if (rows.bias[i] >= 0) any_bias = true;
    IEF_REQUIRE(rows.q[i] >= 0 && rows.q[i] < p->B && rows.k[i] >= 0 && rows.k[i] < p->B && rows.v[i] >= 0 && rows.v[i] < p->B,
                IEF_ERR_INVALID, "ief_attn_fwd: source row index out of range for row %d", i);
    IEF_REQUIRE(rows.k2[i] < p->B && rows.v2[i] < p->B, IEF_ERR_INVALID, "ief_attn_fwd: second-block row index out of range for row %d", i);
    IEF_REQUIRE((rows.k2[i] >= 0) == (rows.v2[i] >= 0), IEF_ERR_INVALID, "ief_attn_fwd: k_src2 and v_src2 must be set together (row %d)", i);
  }
  for (int i = p->B; i < IEF_MAX_ROWS; ++i) {
    rows.q[i] = rows.k[i] = rows.v[i] = 0;
    rows.k2[i] = rows.v2[i] = rows.pslot[i] = rows.bias[i] = -1;
    rows.active[i] = 0;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int impl = p->impl;
  if (p->probs_out) {
    IEF_REQUIRE(impl != IEF_IMPL_TCGEN05, IEF_ERR_UNSUPPORTED, "ief_attn_fwd: probs_out is only produced by the mma kernel");
    if (impl == IEF_IMPL_AUTO && ief_attn_probs_via_lse(p) && p->workspace != nullptr &&
        p->workspace_bytes >= (int64_t)p->B * p->H * p->Nq * (int64_t)sizeof(float)) {
      // O and the row log-sum-exp from the tcgen05 kernel, then one QK^T sweep writes / accumulates the maps
      IEF_REQUIRE((reinterpret_cast<uintptr_t>(p->workspace) & 15) == 0, IEF_ERR_INVALID, "ief_attn_fwd: workspace must be 16-byte aligned");
      ief_attn_params o_only = *p;
      o_only.probs_out = nullptr;
      o_only.workspace = nullptr;
      o_only.workspace_bytes = 0;
      float* lse = static_cast<float*>(p->workspace);
      g_last_impl = "tcgen05+probs";
      const int rc = ief_attn_tc_launch(&o_only, rows, st, lse);
      if (rc != IEF_OK) return rc;
      return ief_attn_probs_from_lse_launch(p, rows, lse, st);
    }
    impl = IEF_IMPL_MMA;
  }
  if (any_bias) {
    IEF_REQUIRE(p->k_src2 == nullptr, IEF_ERR_UNSUPPORTED, "ief_attn_fwd: key_bias cannot be combined with a second key/value block");
    IEF_REQUIRE(impl != IEF_IMPL_TCGEN05 || p->d <= 64, IEF_ERR_UNSUPPORTED,
                "ief_attn_fwd: key_bias on the tcgen05 kernel needs head_dim <= 64 (got %d)", p->d);
    if (p->d > 64) impl = IEF_IMPL_MMA;
  }
  if (impl == IEF_IMPL_AUTO) {
    // the tcgen05 kernel pays off once a head has at least a few 128-key tiles; tiny layers
    // (8x8, 16x16 latents) stay on the warp-level kernel whose CTAs are 4x finer.
    const char* why = nullptr;
    impl = (ief_attn_tc_supported(p, &why) && p->Nq >= 512 && p->Nk >= 256) ? IEF_IMPL_TCGEN05 : IEF_IMPL_MMA;
  }
  if (impl == IEF_IMPL_TCGEN05) {
    g_last_impl = "tcgen05";
    return ief_attn_tc_launch(p, rows, st);
  }
  IEF_REQUIRE(impl == IEF_IMPL_MMA, IEF_ERR_INVALID, "ief_attn_fwd: unknown impl %d", p->impl);
  g_last_impl = "mma";
  return ief_attn_mma_launch(p, rows, st);
}
