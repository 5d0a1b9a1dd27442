// C ABI of libmontage_render.so -- argument validation and kernel dispatch.
// Declarations and the reference interfaces they replace: include/montage_render.h.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>

#include "montage_render.h"
#include "mgr_common.cuh"
#include "mgr_errors.h"
#include "warp_ops.cuh"

#include "launchers_decl.h"
#include "augment_geom.cuh"

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
std::atomic<int> g_debug_path{0};
}

namespace mgr {
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(MGR_ERR_CUDA_BASE + (int)e, "%s: %s", what, cudaGetErrorString(e));
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int side_stream(SideStream* out) {
  constexpr int kMaxDev = 64;
  thread_local SideStream cache[kMaxDev];
  thread_local bool have[kMaxDev] = {};
  int dev = 0;
  MGR_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) return fail(MGR_ERR_UNSUPPORTED, "device ordinal %d", dev);
  if (!have[dev]) {
    SideStream ss;
    MGR_CUDA(cudaStreamCreateWithFlags(&ss.side, cudaStreamNonBlocking));
    MGR_CUDA(cudaEventCreateWithFlags(&ss.fork_ev, cudaEventDisableTiming));
    MGR_CUDA(cudaEventCreateWithFlags(&ss.join_ev, cudaEventDisableTiming));
    cache[dev] = ss;
    have[dev] = true;
  }
  *out = cache[dev];
  return MGR_OK;
}
int side_fork(const SideStream& ss, cudaStream_t s) {
  MGR_CUDA(cudaEventRecord(ss.fork_ev, s));
  MGR_CUDA(cudaStreamWaitEvent(ss.side, ss.fork_ev, 0));
  return MGR_OK;
}
int side_join(const SideStream& ss, cudaStream_t s) {
  MGR_CUDA(cudaEventRecord(ss.join_ev, ss.side));
  MGR_CUDA(cudaStreamWaitEvent(s, ss.join_ev, 0));
  return MGR_OK;
}
int debug_path() { return g_debug_path.load(std::memory_order_relaxed); }
}  // namespace mgr

namespace {
using mgr::fail;

int check_common(const void* x, const int64_t* xs, int B, int L, int H, int W, int dtype, int range_mode,
                 mgr::Geometry* g) {
  if (!x) return fail(MGR_ERR_INVALID_ARGUMENT, "x is NULL");
  if (B < 0 || L < 1 || H < 1 || W < 1)
    return fail(MGR_ERR_INVALID_ARGUMENT, "bad shape B=%d L=%d H=%d W=%d (need B>=0, L,H,W>=1)", B, L, H, W);
  if (dtype != MGR_F32 && dtype != MGR_BF16 && dtype != MGR_F16)
    return fail(MGR_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
  if (range_mode != MGR_RANGE_M11 && range_mode != MGR_RANGE_01)
    return fail(MGR_ERR_INVALID_ARGUMENT, "bad range_mode %d", range_mode);
  if (B > 65535) return fail(MGR_ERR_UNSUPPORTED, "B=%d exceeds 65535 per call; split the batch", B);
  if (L > 32) return fail(MGR_ERR_UNSUPPORTED, "L=%d exceeds 32 layers", L);
  g->B = B; g->L = L; g->H = H; g->W = W;
  g->m11 = (range_mode == MGR_RANGE_M11);
  g->vec8 = 0;
  if (xs) {
    if (xs[4] != 1) return fail(MGR_ERR_UNSUPPORTED, "x stride along W must be 1 (got %lld)", (long long)xs[4]);
    g->sb = xs[0]; g->sl = xs[1]; g->sc = xs[2]; g->sh = xs[3];
  } else {
    g->sh = W; g->sc = (long long)H * W; g->sl = 4 * g->sc; g->sb = (long long)L * g->sl;
  }
  if (g->sh < W || g->sh * (long long)H > 0x7fffffffLL)
    return fail(MGR_ERR_UNSUPPORTED, "row stride %lld out of range for H=%d W=%d", g->sh, H, W);
  return MGR_OK;
}

}  // namespace

extern "C" {

int mgr_abi_version(void) { return MGR_ABI_VERSION; }

#define MGR_STR2(v) #v
#define MGR_STR(v) MGR_STR2(v)
const char* mgr_build_info(void) {
  return "libmontage_render abi " MGR_STR(MGR_ABI_VERSION) " sm_100a cuda " MGR_STR(__CUDACC_VER_MAJOR__) "." MGR_STR(
      __CUDACC_VER_MINOR__) " built " __DATE__ " " __TIME__;
}

const char* mgr_last_error(void) { return g_err; }

int mgr_set_debug_path(int path) {
  if (path < 0 || path > 4) return fail(MGR_ERR_INVALID_ARGUMENT, "debug path %d unknown", path);
  g_debug_path.store(path, std::memory_order_relaxed);
  return MGR_OK;
}

long long mgr_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t mgr_saved_alpha_bytes(int B, int L, int H, int W, int dtype) {
  if (B <= 0 || L <= 0 || H <= 0 || W <= 0) return 0;
  // the alpha samples [B,L,H,W], then one int per sample: "all placements are pure translations" (written by the
  // forward before its kernels run, so that a CTA learns with ONE load whether the sample is its kernel's)
  return mgr::saved_alpha_flags_offset(B, L, H, W, dtype == MGR_F32 ? 4 : 2) + sizeof(int) * (size_t)B;
}

int mgr_render_forward(const void* x, const int64_t* x_strides, const float* theta, void* out, void* saved_alpha,
                       int B, int L, int H, int W, int dtype, int range_mode, void* stream) {
  mgr::Geometry g;
  if (int rc = check_common(x, x_strides, B, L, H, W, dtype, range_mode, &g)) return rc;
  if (!out) return fail(MGR_ERR_INVALID_ARGUMENT, "out is NULL");
  if (B == 0) return MGR_OK;
  // the backward of a warped stack addresses layers through a 16-bit grid dimension: refuse here, before any work is
  // queued, what mgr_render_backward would refuse later
  if (theta && (long long)B * L > 65535) return fail(MGR_ERR_UNSUPPORTED, "B*L=%lld exceeds 65535 per call; split the batch", (long long)B * L);
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case MGR_F32: return mgr_fwd_f32(x, theta, out, saved_alpha, g, s);
    case MGR_BF16: return mgr_fwd_bf16(x, theta, out, saved_alpha, g, s);
    default: return mgr_fwd_f16(x, theta, out, saved_alpha, g, s);
  }
}

namespace {
// tiled two-pass path: fp32 records (T_l a_l, d a_l) per layer-pixel + G_P per pixel
// + per layer: inverse plan (128 B), launch-order entry, work-list entry; + 4 counters; + a flag per sample
size_t tiled_workspace(int B, int L, int H, int W) {
  return ((size_t)B * L * H * W) * 8 + ((size_t)B * H * W) * 16 + (size_t)B * L * (128 + 4 + 4) + 32 + (size_t)B * 4;
}
}  // namespace

size_t mgr_render_backward_workspace_bytes(int B, int L, int H, int W, int dtype, int has_theta, int flags) {
  if (B <= 0 || L <= 0 || H <= 0 || W <= 0 || !has_theta) return 0;
  size_t need = tiled_workspace(B, L, H, W);
  // general direct-gather path with 16-bit storage: fp32 scatter accumulator
  if ((flags & MGR_NEED_GRAD_X) && dtype != MGR_F32) {
    const size_t scatter = sizeof(float) * (size_t)B * L * 4 * H * W;
    if (scatter > need) need = scatter;
  }
  return need;
}

size_t mgr_render_backward_workspace_bytes_for(const void* x, const int64_t* x_strides, int has_theta, int has_saved_alpha,
                                               int B, int L, int H, int W, int dtype, int flags) {
  if (B <= 0 || L <= 0 || H <= 0 || W <= 0 || !has_theta) return 0;
  mgr::Geometry g;
  if (check_common(x, x_strides, B, L, H, W, dtype, MGR_RANGE_M11, &g) != MGR_OK) return 0;
  const bool tiled = has_saved_alpha && (dtype == MGR_F32 ? mgr_tiled_ok_f32(x, g) : dtype == MGR_BF16 ? mgr_tiled_ok_bf16(x, g) : mgr_tiled_ok_f16(x, g));
  if (tiled) return tiled_workspace(B, L, H, W);
  if ((flags & MGR_NEED_GRAD_X) && dtype != MGR_F32) return sizeof(float) * (size_t)B * L * 4 * H * W;   // fp32 scatter accumulator
  return 0;
}

int mgr_render_backward(const void* x, const int64_t* x_strides, const float* theta, const void* out,
                        const void* grad_out, const void* saved_alpha, void* grad_x, float* grad_theta, void* workspace,
                        size_t workspace_bytes, int B, int L, int H, int W, int dtype, int range_mode, int flags,
                        void* stream) {
  mgr::Geometry g;
  if (int rc = check_common(x, x_strides, B, L, H, W, dtype, range_mode, &g)) return rc;
  if (!out || !grad_out) return fail(MGR_ERR_INVALID_ARGUMENT, "out / grad_out is NULL");
  if (!theta) flags &= ~MGR_NEED_GRAD_THETA;
  if ((flags & MGR_NEED_GRAD_X) && !grad_x) return fail(MGR_ERR_INVALID_ARGUMENT, "grad_x is NULL but requested");
  if ((flags & MGR_NEED_GRAD_THETA) && !grad_theta)
    return fail(MGR_ERR_INVALID_ARGUMENT, "grad_theta is NULL but requested");
  if (!(flags & (MGR_NEED_GRAD_X | MGR_NEED_GRAD_THETA)) || B == 0) return MGR_OK;
  // what the path this call takes needs (the tiled kernels for aligned tensors with saved alphas, else the scatter path)
  const size_t need = mgr_render_backward_workspace_bytes_for(x, x_strides, theta != nullptr, saved_alpha != nullptr, B, L, H, W, dtype, flags);
  if (need > 0 && (!workspace || workspace_bytes < need))
    return fail(MGR_ERR_WORKSPACE_TOO_SMALL, "workspace %zu bytes < required %zu", workspace_bytes, need);
  if (theta && (long long)B * L > 65535) return fail(MGR_ERR_UNSUPPORTED, "B*L=%lld exceeds 65535 per call; split the batch", (long long)B * L);
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case MGR_F32: return mgr_bwd_f32(x, theta, out, grad_out, saved_alpha, grad_x, grad_theta, workspace, g, flags, s);
    case MGR_BF16: return mgr_bwd_bf16(x, theta, out, grad_out, saved_alpha, grad_x, grad_theta, workspace, g, flags, s);
    default: return mgr_bwd_f16(x, theta, out, grad_out, saved_alpha, grad_x, grad_theta, workspace, g, flags, s);
  }
}


int mgr_warp_forward(const void* x, const int64_t* x_strides, const float* theta, void* out, int B, int L, int H,
                     int W, int dtype, int range_mode, void* stream) {
  mgr::Geometry g;
  if (int rc = check_common(x, x_strides, B, L, H, W, dtype, range_mode, &g)) return rc;
  if (!theta || !out) return fail(MGR_ERR_INVALID_ARGUMENT, "theta / out is NULL");
  if (B == 0) return MGR_OK;
  if ((long long)B * L > 65535) return fail(MGR_ERR_UNSUPPORTED, "B*L=%lld exceeds 65535 per call", (long long)B * L);
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case MGR_F32: return mgr_warp_fwd_f32(x, theta, out, g, s);
    case MGR_BF16: return mgr_warp_fwd_bf16(x, theta, out, g, s);
    default: return mgr_warp_fwd_f16(x, theta, out, g, s);
  }
}

size_t mgr_warp_backward_workspace_bytes(int B, int L, int H, int W, int dtype, int flags) {
  if (B <= 0 || L <= 0 || H <= 0 || W <= 0) return 0;
  // gather path: per layer an inverse placement, a launch-order entry and a work-list entry; counters; a flag per sample
  size_t need = (size_t)B * L * (128 + 4 + 4) + 32 + (size_t)B * 4;
  // scatter path (odd widths / strides): an fp32 accumulator for 16-bit tensors
  if ((flags & MGR_NEED_GRAD_X) && dtype != MGR_F32) {
    const size_t scatter = sizeof(float) * (size_t)B * L * 4 * H * W;
    if (scatter > need) need = scatter;
  }
  return need;
}

int mgr_warp_backward(const void* x, const int64_t* x_strides, const float* theta, const void* grad_out,
                      void* grad_x, float* grad_theta, void* workspace, size_t workspace_bytes, int B, int L, int H,
                      int W, int dtype, int range_mode, int flags, void* stream) {
  mgr::Geometry g;
  if (int rc = check_common(x, x_strides, B, L, H, W, dtype, range_mode, &g)) return rc;
  if (!theta || !grad_out) return fail(MGR_ERR_INVALID_ARGUMENT, "theta / grad_out is NULL");
  if ((flags & MGR_NEED_GRAD_X) && !grad_x) return fail(MGR_ERR_INVALID_ARGUMENT, "grad_x is NULL but requested");
  if ((flags & MGR_NEED_GRAD_THETA) && !grad_theta)
    return fail(MGR_ERR_INVALID_ARGUMENT, "grad_theta is NULL but requested");
  if (!(flags & (MGR_NEED_GRAD_X | MGR_NEED_GRAD_THETA)) || B == 0) return MGR_OK;
  if ((long long)B * L > 65535) return fail(MGR_ERR_UNSUPPORTED, "B*L=%lld exceeds 65535 per call", (long long)B * L);
  const size_t need = mgr_warp_backward_workspace_bytes(B, L, H, W, dtype, flags);
  if (need > 0 && (!workspace || workspace_bytes < need))
    return fail(MGR_ERR_WORKSPACE_TOO_SMALL, "workspace %zu bytes < required %zu", workspace_bytes, need);
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case MGR_F32: return mgr_warp_bwd_f32(x, theta, grad_out, grad_x, grad_theta, workspace, g, flags, s);
    case MGR_BF16: return mgr_warp_bwd_bf16(x, theta, grad_out, grad_x, grad_theta, workspace, g, flags, s);
    default: return mgr_warp_bwd_f16(x, theta, grad_out, grad_x, grad_theta, workspace, g, flags, s);
  }
}

int mgr_translation_to_theta(const float* translation, float* theta, long long n, void* stream) {
  if (n < 0) return fail(MGR_ERR_INVALID_ARGUMENT, "n=%lld", n);
  if (n == 0) return MGR_OK;
  if (!translation || !theta) return fail(MGR_ERR_INVALID_ARGUMENT, "translation / theta is NULL");
  mgr::translation_to_theta_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(translation, theta, n);
  MGR_CUDA(cudaGetLastError());
  mgr::count_launch();
  return MGR_OK;
}

int mgr_pad_stack_layer(const void* src, const int64_t* src_strides, void* dst, int B, int L, int l, int h, int w,
                        int H, int W, float pad_value, int dtype, void* stream) {
  if (!src || !dst) return fail(MGR_ERR_INVALID_ARGUMENT, "src / dst is NULL");
  if (B < 0 || L < 1 || l < 0 || l >= L || h < 1 || w < 1 || H < h || W < w)
    return fail(MGR_ERR_INVALID_ARGUMENT, "bad shape B=%d L=%d l=%d h=%d w=%d H=%d W=%d", B, L, l, h, w, H, W);
  if (dtype != MGR_F32 && dtype != MGR_BF16 && dtype != MGR_F16) return fail(MGR_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
  if (B == 0) return MGR_OK;
  long long ss[4];
  if (src_strides) { for (int k = 0; k < 4; ++k) ss[k] = src_strides[k]; }
  else { ss[3] = 1; ss[2] = w; ss[1] = (long long)h * w; ss[0] = 4 * ss[1]; }
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case MGR_F32: return mgr_pad_stack_f32(src, ss, dst, B, L, l, h, w, H, W, pad_value, s);
    case MGR_BF16: return mgr_pad_stack_bf16(src, ss, dst, B, L, l, h, w, H, W, pad_value, s);
    default: return mgr_pad_stack_f16(src, ss, dst, B, L, l, h, w, H, W, pad_value, s);
  }
}


int mgr_composite_jvp(const void* x, const int64_t* x_strides, const void* tangent, void* out_tangent, int B, int L,
                      int H, int W, int dtype, int range_mode, void* stream) {
  mgr::Geometry g;
  if (int rc = check_common(x, x_strides, B, L, H, W, dtype, range_mode, &g)) return rc;
  if (!tangent || !out_tangent) return fail(MGR_ERR_INVALID_ARGUMENT, "tangent / out_tangent is NULL");
  if (B == 0) return MGR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case MGR_F32: return mgr_jvp_f32(x, tangent, out_tangent, g, s);
    case MGR_BF16: return mgr_jvp_bf16(x, tangent, out_tangent, g, s);
    default: return mgr_jvp_f16(x, tangent, out_tangent, g, s);
  }
}

namespace {
int ragged_geometry(const MgrLayer* layers, const float* theta, int B, int L, int H, int W, int dtype, int range_mode,
                    mgr::Geometry* g, mgr::SrcLayers* src) {
  if (!layers) return fail(MGR_ERR_INVALID_ARGUMENT, "layers is NULL");
  if (!theta) return fail(MGR_ERR_UNSUPPORTED, "a ragged stack needs theta (composite-only stacks use the canvas layout)");
  if (B < 0 || L < 1 || H < 1 || W < 1)
    return fail(MGR_ERR_INVALID_ARGUMENT, "bad shape B=%d L=%d H=%d W=%d (need B>=0, L,H,W>=1)", B, L, H, W);
  if (dtype != MGR_F32 && dtype != MGR_BF16 && dtype != MGR_F16) return fail(MGR_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
  if (range_mode != MGR_RANGE_M11 && range_mode != MGR_RANGE_01) return fail(MGR_ERR_INVALID_ARGUMENT, "bad range_mode %d", range_mode);
  if (B > 65535) return fail(MGR_ERR_UNSUPPORTED, "B=%d exceeds 65535 per call; split the batch", B);
  if (L < 2 || L > mgr::kMaxTiledLayers) return fail(MGR_ERR_UNSUPPORTED, "a ragged stack needs 2..32 layers, got %d", L);
  if ((long long)B * L > 65535) return fail(MGR_ERR_UNSUPPORTED, "B*L=%lld exceeds 65535 per call; split the batch", (long long)B * L);
  g->B = B; g->L = L; g->H = H; g->W = W;
  g->m11 = (range_mode == MGR_RANGE_M11);
  g->vec8 = 0;
  g->sh = W; g->sc = (long long)H * W; g->sl = 4 * g->sc; g->sb = (long long)L * g->sl;     // unused by the ragged kernels
  *src = mgr::SrcLayers{};
  for (int l = 0; l < L; ++l)
    src->s[l] = mgr::SrcLayer{layers[l].ptr, layers[l].sb, layers[l].sc, layers[l].sh, layers[l].h, layers[l].w, layers[l].top, layers[l].left};
  return MGR_OK;
}
}  // namespace

int mgr_render_forward_ragged(const MgrLayer* layers, const float* theta, void* out, void* saved_alpha, int B, int L, int H,
                              int W, int dtype, int range_mode, void* stream) {
  mgr::Geometry g;
  mgr::SrcLayers src;
  if (int rc = ragged_geometry(layers, theta, B, L, H, W, dtype, range_mode, &g, &src)) return rc;
  if (!out) return fail(MGR_ERR_INVALID_ARGUMENT, "out is NULL");
  if (B == 0) return MGR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case MGR_F32: return mgr_fwd_ragged_f32(src, theta, out, saved_alpha, g, s);
    case MGR_BF16: return mgr_fwd_ragged_bf16(src, theta, out, saved_alpha, g, s);
    default: return mgr_fwd_ragged_f16(src, theta, out, saved_alpha, g, s);
  }
}

int mgr_render_backward_ragged(const MgrLayer* layers, const float* theta, const void* out, const void* grad_out,
                               const void* saved_alpha, const MgrLayer* grads, float* grad_theta, void* workspace,
                               size_t workspace_bytes, int B, int L, int H, int W, int dtype, int range_mode, int flags,
                               void* stream) {
  mgr::Geometry g;
  mgr::SrcLayers src;
  if (int rc = ragged_geometry(layers, theta, B, L, H, W, dtype, range_mode, &g, &src)) return rc;
  if (!out || !grad_out || !saved_alpha) return fail(MGR_ERR_INVALID_ARGUMENT, "out / grad_out / saved_alpha is NULL");
  if ((flags & MGR_NEED_GRAD_X) && !grads) return fail(MGR_ERR_INVALID_ARGUMENT, "grads is NULL but requested");
  if ((flags & MGR_NEED_GRAD_THETA) && !grad_theta) return fail(MGR_ERR_INVALID_ARGUMENT, "grad_theta is NULL but requested");
  if (!(flags & (MGR_NEED_GRAD_X | MGR_NEED_GRAD_THETA))) return MGR_OK;
  if (B == 0) return MGR_OK;
  const size_t need = mgr_render_backward_workspace_bytes(B, L, H, W, MGR_F32, 1, flags);   // the tiled path's records only
  if (!workspace || workspace_bytes < need)
    return fail(MGR_ERR_WORKSPACE_TOO_SMALL, "workspace %zu bytes < required %zu", workspace_bytes, need);
  mgr::DstLayers dst{};
  if (flags & MGR_NEED_GRAD_X)
    for (int l = 0; l < L; ++l)
      dst.s[l] = mgr::DstLayer{grads[l].ptr, grads[l].sb, grads[l].sc, grads[l].sh, grads[l].h, grads[l].w, grads[l].top, grads[l].left};
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case MGR_F32: return mgr_bwd_ragged_f32(src, theta, out, grad_out, saved_alpha, dst, grad_theta, workspace, g, flags, s);
    case MGR_BF16: return mgr_bwd_ragged_bf16(src, theta, out, grad_out, saved_alpha, dst, grad_theta, workspace, g, flags, s);
    default: return mgr_bwd_ragged_f16(src, theta, out, grad_out, saved_alpha, dst, grad_theta, workspace, g, flags, s);
  }
}

int mgr_composite_u8(const void* x, const int64_t* x_strides, float* out_f32, unsigned char* out_u8, int B, int L, int H,
                     int W, int dtype, int range_mode, void* stream) {
  mgr::Geometry g;
  if (int rc = check_common(x, x_strides, B, L, H, W, dtype, range_mode, &g)) return rc;
  if (!out_f32 && !out_u8) return fail(MGR_ERR_INVALID_ARGUMENT, "out_f32 and out_u8 are both NULL");
  if (B == 0) return MGR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case MGR_F32: return mgr_pil_f32(x, out_f32, out_u8, g, s);
    case MGR_BF16: return mgr_pil_bf16(x, out_f32, out_u8, g, s);
    default: return mgr_pil_f16(x, out_f32, out_u8, g, s);
  }
}

// ---- AugmentPipe's geometric execution block (augment_geom.cuh) ------------------------------------------------
namespace {
int aug_geometry(int B, int C, int H, int W, int mx0, int my0, int mx1, int my1, mgr::AugGeom* a) {
  if (B < 0 || C < 1 || H < 2 || W < 2) return fail(MGR_ERR_INVALID_ARGUMENT, "bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  if (mx0 < 0 || mx1 < 0 || my0 < 0 || my1 < 0 || mx0 > W - 1 || mx1 > W - 1 || my0 > H - 1 || my1 > H - 1)
    return fail(MGR_ERR_INVALID_ARGUMENT, "reflect margins (%d,%d,%d,%d) must lie in [0, size - 1]", mx0, my0, mx1, my1);
  a->B = B; a->C = C; a->H = H; a->W = W;
  a->mx0 = mx0; a->my0 = my0; a->mx1 = mx1; a->my1 = my1;
  a->Hp = H + my0 + my1; a->Wp = W + mx0 + mx1;
  a->Hu = 2 * a->Hp; a->Wu = 2 * a->Wp;
  a->Hs = 2 * (H + 6); a->Ws = 2 * (W + 6);
  if ((long long)B * C * a->Hu * a->Wu >= (1LL << 40)) return fail(MGR_ERR_UNSUPPORTED, "problem too large");
  return MGR_OK;
}
// intermediate of the separable FIR stages, per (b, c) plane: rows x cols of the half-filtered signal
size_t aug_temp_elems(const mgr::AugGeom& a) {
  size_t m = (size_t)a.H * a.Wu;                               // upsample: along x first
  if ((size_t)a.Hs * a.W > m) m = (size_t)a.Hs * a.W;          // decimation (and its adjoint): x of the tall signal
  if ((size_t)a.Hu * a.Wp > m) m = (size_t)a.Hu * a.Wp;        // adjoint of the upsample: along x first
  return m;
}
unsigned aug_grid(long long total) {
  const long long blocks = (total + 255) / 256;
  return (unsigned)(blocks < 148 * 32 ? (blocks > 0 ? blocks : 1) : 148 * 32);
}
}  // namespace

size_t mgr_augment_geom_workspace_bytes(int B, int C, int H, int W, int mx0, int my0, int mx1, int my1) {
  mgr::AugGeom a;
  if (aug_geometry(B, C, H, W, mx0, my0, mx1, my1, &a) != MGR_OK) return 0;
  const size_t bc = (size_t)B * C;
  return sizeof(float) * bc * ((size_t)a.Hu * a.Wu + (size_t)a.Hs * a.Ws + (size_t)a.Hp * a.Wp + aug_temp_elems(a));
}

int mgr_augment_geom_forward(const float* images, const float* theta, float* out, void* workspace, size_t workspace_bytes,
                             int B, int C, int H, int W, int mx0, int my0, int mx1, int my1, void* stream) {
  mgr::AugGeom a;
  if (int rc = aug_geometry(B, C, H, W, mx0, my0, mx1, my1, &a)) return rc;
  if (!images || !theta || !out) return fail(MGR_ERR_INVALID_ARGUMENT, "images / theta / out is NULL");
  if (B == 0) return MGR_OK;
  const size_t need = mgr_augment_geom_workspace_bytes(B, C, H, W, mx0, my0, mx1, my1);
  if (!workspace || workspace_bytes < need) return fail(MGR_ERR_WORKSPACE_TOO_SMALL, "workspace %zu bytes < required %zu", workspace_bytes, need);
  cudaStream_t s = (cudaStream_t)stream;
  const long long bc = (long long)B * C;
  float* U = (float*)workspace;
  float* S = U + bc * a.Hu * a.Wu;
  float* T = S + bc * a.Hs * a.Ws + bc * a.Hp * a.Wp;
  using namespace mgr;
  // reflect pad + x2 upsample: along x ([H, W] -> [H, Wu]), then along y (-> [Hu, Wu])
  aug_fir_1d<kUp, true><<<aug_grid(bc * H * a.Wu), 256, 0, s>>>(images, T, bc, W, a.Wu, H, mx0);
  aug_up_y_blocked<<<aug_grid(bc * ((a.Hu + kUpRows - 1) / kUpRows) * a.Wu), 256, 0, s>>>(T, U, bc, H, a.Hu, a.Wu, my0);
  aug_sample<false><<<aug_grid((long long)B * a.Hs * a.Ws), 256, 0, s>>>(theta, U, S, a);
  // low-pass + decimate: along x ([Hs, Ws] -> [Hs, W]), then along y (-> [H, W])
  aug_fir_1d<kDown, true><<<aug_grid(bc * a.Hs * W), 256, 0, s>>>(S, T, bc, a.Ws, W, a.Hs, 0);
  aug_fir_1d<kDown, false><<<aug_grid(bc * H * W), 256, 0, s>>>(T, out, bc, a.Hs, H, W, 0);
  MGR_CUDA(cudaGetLastError());
  mgr::count_launch(5);
  return MGR_OK;
}

int mgr_augment_geom_backward(const float* grad_out, const float* theta, float* grad_images, void* workspace,
                              size_t workspace_bytes, int B, int C, int H, int W, int mx0, int my0, int mx1, int my1,
                              void* stream) {
  mgr::AugGeom a;
  if (int rc = aug_geometry(B, C, H, W, mx0, my0, mx1, my1, &a)) return rc;
  if (!grad_out || !theta || !grad_images) return fail(MGR_ERR_INVALID_ARGUMENT, "grad_out / theta / grad_images is NULL");
  if (B == 0) return MGR_OK;
  const size_t need = mgr_augment_geom_workspace_bytes(B, C, H, W, mx0, my0, mx1, my1);
  if (!workspace || workspace_bytes < need) return fail(MGR_ERR_WORKSPACE_TOO_SMALL, "workspace %zu bytes < required %zu", workspace_bytes, need);
  cudaStream_t s = (cudaStream_t)stream;
  const long long bc = (long long)B * C;
  float* gU = (float*)workspace;
  float* gS = gU + bc * a.Hu * a.Wu;
  float* gxp = gS + bc * a.Hs * a.Ws;
  float* T = gxp + bc * a.Hp * a.Wp;
  using namespace mgr;
  // adjoint of the decimation: along y ([H, W] -> [Hs, W]), then along x (-> [Hs, Ws])
  aug_adj_y_blocked<kDownAdj><<<aug_grid(bc * ((a.Hs + 7) / 8) * W), 256, 0, s>>>(grad_out, T, bc, H, a.Hs, W);
  aug_fir_1d<kDownAdj, true><<<aug_grid(bc * a.Hs * a.Ws), 256, 0, s>>>(T, gS, bc, W, a.Ws, a.Hs, 0);
  MGR_CUDA(cudaMemsetAsync(gU, 0, sizeof(float) * bc * a.Hu * a.Wu, s));
  aug_sample<true><<<aug_grid((long long)B * a.Hs * a.Ws), 256, 0, s>>>(theta, gS, gU, a);
  // adjoint of the upsample w.r.t. the padded image: along x ([Hu, Wu] -> [Hu, Wp]), then along y (-> [Hp, Wp]); fold
  aug_fir_1d<kUpAdj, true><<<aug_grid(bc * a.Hu * a.Wp), 256, 0, s>>>(gU, T, bc, a.Wu, a.Wp, a.Hu, 0);
  aug_adj_y_blocked<kUpAdj><<<aug_grid(bc * ((a.Hp + 3) / 4) * a.Wp), 256, 0, s>>>(T, gxp, bc, a.Hu, a.Hp, a.Wp);
  aug_fold_reflect<<<aug_grid(bc * H * W), 256, 0, s>>>(gxp, grad_images, a);
  MGR_CUDA(cudaGetLastError());
  mgr::count_launch(6);
  return MGR_OK;
}

// ---- end-to-end entry point with HOST buffers: chunked, double-buffered, three streams ----------------
namespace {
size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
struct HostSlot { char *x, *theta, *go, *out, *sav, *gx, *gt, *ws; };
// Two copy streams and the events that order them, created once per device and kept for the life of the process
// (creating and destroying them per call cost ~60 us of host time and leaked on error paths).
struct HostPipe {
  std::mutex mu;
  bool ready = false;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t start = nullptr, done_in = nullptr, done_out = nullptr, h2d[2] = {}, comp[2] = {}, freed[2] = {};
  int init() {                                      // caller holds mu
    if (ready) return MGR_OK;
    cudaEvent_t* ev[] = {&start, &done_in, &done_out, &h2d[0], &h2d[1], &comp[0], &comp[1], &freed[0], &freed[1]};
    cudaError_t e = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking);
    for (cudaEvent_t* p : ev)
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(p, cudaEventDisableTiming);
    if (e != cudaSuccess) {                         // all or nothing
      for (cudaEvent_t* p : ev) { if (*p) cudaEventDestroy(*p); *p = nullptr; }
      if (s_in) cudaStreamDestroy(s_in);
      if (s_out) cudaStreamDestroy(s_out);
      s_in = s_out = nullptr;
      return mgr::cuda_fail(e, "host pipeline: stream / event creation");
    }
    ready = true;
    return MGR_OK;
  }
};
constexpr int kMaxDevices = 64;
HostPipe* host_pipe(int device) {
  static HostPipe pipes[kMaxDevices];
  return (device >= 0 && device < kMaxDevices) ? &pipes[device] : nullptr;
}
size_t host_slot_bytes(int cb, int L, int H, int W, int dtype, HostSlot* s, char* base) {
  const size_t es = dtype == MGR_F32 ? 4 : 2;
  const size_t nx = (size_t)cb * L * 4 * H * W * es, no = (size_t)cb * 4 * H * W * es, nt = (size_t)cb * L * 6 * 4;
  const size_t nsav = mgr_saved_alpha_bytes(cb, L, H, W, dtype);
  const size_t nws = mgr_render_backward_workspace_bytes(cb, L, H, W, dtype, 1, 3);
  size_t off = 0;
  auto take = [&](char** p, size_t n) { if (s) *p = base + off; off += align256(n); };
  take(s ? &s->x : nullptr, nx); take(s ? &s->theta : nullptr, nt); take(s ? &s->go : nullptr, no);
  take(s ? &s->out : nullptr, no); take(s ? &s->sav : nullptr, nsav); take(s ? &s->gx : nullptr, nx);
  take(s ? &s->gt : nullptr, nt); take(s ? &s->ws : nullptr, nws);
  return off;
}
}  // namespace

size_t mgr_render_host_workspace_bytes(int chunk_B, int L, int H, int W, int dtype) {
  if (chunk_B <= 0 || L <= 0 || H <= 0 || W <= 0) return 0;
  return 2 * host_slot_bytes(chunk_B, L, H, W, dtype, nullptr, nullptr);
}

int mgr_render_fwd_bwd_host(const void* h_x, const float* h_theta, const void* h_grad_out, void* h_out,
                            void* h_grad_x, float* h_grad_theta, void* d_workspace, size_t d_workspace_bytes,
                            int chunk_B, int B, int L, int H, int W, int dtype, int range_mode, void* stream) {
  if (!h_x || !h_theta || !h_grad_out || !h_out || !h_grad_x || !h_grad_theta)
    return fail(MGR_ERR_INVALID_ARGUMENT, "a host buffer is NULL");
  if (chunk_B < 1 || B < 0 || L < 1 || H < 1 || W < 1) return fail(MGR_ERR_INVALID_ARGUMENT, "bad shape");
  if (dtype != MGR_F32 && dtype != MGR_BF16 && dtype != MGR_F16) return fail(MGR_ERR_INVALID_ARGUMENT, "bad dtype %d", dtype);
  if (B == 0) return MGR_OK;
  const size_t need = mgr_render_host_workspace_bytes(chunk_B, L, H, W, dtype);
  if (!d_workspace || d_workspace_bytes < need)
    return fail(MGR_ERR_WORKSPACE_TOO_SMALL, "device workspace %zu bytes < required %zu", d_workspace_bytes, need);
  HostSlot slot[2];
  const size_t per = host_slot_bytes(chunk_B, L, H, W, dtype, &slot[0], (char*)d_workspace);
  host_slot_bytes(chunk_B, L, H, W, dtype, &slot[1], (char*)d_workspace + per);
  const size_t es = dtype == MGR_F32 ? 4 : 2;
  const size_t sx = (size_t)L * 4 * H * W * es, so = (size_t)4 * H * W * es, st_ = (size_t)L * 6 * 4;   // bytes per sample
  cudaStream_t user = (cudaStream_t)stream;
  int device = 0;
  MGR_CUDA(cudaGetDevice(&device));
  HostPipe* pipe = host_pipe(device);
  if (!pipe) return fail(MGR_ERR_UNSUPPORTED, "device ordinal %d out of range for the host pipeline", device);
  // the pipe's events are re-recorded by every call: one call at a time per device enqueues (the enqueue is host work
  // of ~100 us; the device side of consecutive calls still overlaps as far as the caller's streams allow)
  std::lock_guard<std::mutex> lock(pipe->mu);
  if (int rc = pipe->init()) return rc;
  cudaStream_t s_in = pipe->s_in, s_out = pipe->s_out;
  cudaEvent_t* h2d = pipe->h2d; cudaEvent_t* comp = pipe->comp; cudaEvent_t* freed = pipe->freed;
  // the copy streams start after whatever the caller has queued on `user`
  MGR_CUDA(cudaEventRecord(pipe->start, user));
  MGR_CUDA(cudaStreamWaitEvent(s_in, pipe->start, 0));
  MGR_CUDA(cudaStreamWaitEvent(s_out, pipe->start, 0));
  // From here on every exit path -- success, a failed launch, a failed copy -- makes `user` wait for both copy streams,
  // so the caller's next work on `user` (or the next call's reuse of the device slots) never races copies in flight.
  struct Join {
    cudaStream_t user, s_in, s_out;
    cudaEvent_t e_in, e_out;
    ~Join() {
      if (cudaEventRecord(e_in, s_in) == cudaSuccess) cudaStreamWaitEvent(user, e_in, 0);
      if (cudaEventRecord(e_out, s_out) == cudaSuccess) cudaStreamWaitEvent(user, e_out, 0);
    }
  } join{user, s_in, s_out, pipe->done_in, pipe->done_out};
  // The last D2H runs with the H2D engine idle (and the first H2D with the D2H engine idle), so the tail of the
  // batch is cut into progressively smaller chunks: chunk_B, ..., chunk_B/2, chunk_B/4, chunk_B/4.
  const int min_cb = chunk_B >= 4 ? chunk_B / 4 : 1;
  int b0 = 0;
  for (int c = 0; b0 < B; ++c) {
    const int k = c & 1, left = B - b0;
    int cb = left < chunk_B ? left : chunk_B;
    if (left <= chunk_B && cb > min_cb) cb = (cb / 2 > min_cb) ? cb / 2 : min_cb;
    const HostSlot& S = slot[k];
    // The slot's INPUT buffers are free once the kernels of chunk c-2 have run; its OUTPUT buffers once the D2H of
    // chunk c-2 has drained them.  Waiting for each separately keeps both copy engines busy back to back.
    if (c >= 2) MGR_CUDA(cudaStreamWaitEvent(s_in, comp[k], 0));
    MGR_CUDA(cudaMemcpyAsync(S.x, (const char*)h_x + b0 * sx, cb * sx, cudaMemcpyHostToDevice, s_in));
    MGR_CUDA(cudaMemcpyAsync(S.theta, (const char*)h_theta + b0 * st_, cb * st_, cudaMemcpyHostToDevice, s_in));
    MGR_CUDA(cudaMemcpyAsync(S.go, (const char*)h_grad_out + b0 * so, cb * so, cudaMemcpyHostToDevice, s_in));
    MGR_CUDA(cudaEventRecord(h2d[k], s_in));
    MGR_CUDA(cudaStreamWaitEvent(user, h2d[k], 0));
    if (c >= 2) MGR_CUDA(cudaStreamWaitEvent(user, freed[k], 0));
    if (int rc = mgr_render_forward(S.x, nullptr, (const float*)S.theta, S.out, S.sav, cb, L, H, W, dtype, range_mode, user)) return rc;
    if (int rc = mgr_render_backward(S.x, nullptr, (const float*)S.theta, S.out, S.go, S.sav, S.gx, (float*)S.gt, S.ws,
                                     mgr_render_backward_workspace_bytes(cb, L, H, W, dtype, 1, 3), cb, L, H, W, dtype,
                                     range_mode, 3, user))
      return rc;
    MGR_CUDA(cudaEventRecord(comp[k], user));
    MGR_CUDA(cudaStreamWaitEvent(s_out, comp[k], 0));
    MGR_CUDA(cudaMemcpyAsync((char*)h_out + b0 * so, S.out, cb * so, cudaMemcpyDeviceToHost, s_out));
    MGR_CUDA(cudaMemcpyAsync((char*)h_grad_x + b0 * sx, S.gx, cb * sx, cudaMemcpyDeviceToHost, s_out));
    MGR_CUDA(cudaMemcpyAsync((char*)h_grad_theta + b0 * st_, S.gt, cb * st_, cudaMemcpyDeviceToHost, s_out));
    MGR_CUDA(cudaEventRecord(freed[k], s_out));
    b0 += cb;
  }
  return MGR_OK;                  // ~Join: `user` completes only when the last results have landed on the host
}

}  // extern "C"
