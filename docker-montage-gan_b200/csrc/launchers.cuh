// Typed launchers: compiled once per storage dtype (inst_*.cu) so the dtypes build in parallel.
#pragma once
#include "mgr_errors.h"
#include "render_direct.cuh"

namespace mgr {

template <typename T>
int launch_forward(const void* x, const float* theta, void* out, const mgr::Geometry& g, cudaStream_t s) {
  dim3 grid((g.W + mgr::kTileW - 1) / mgr::kTileW, (g.H + mgr::kTileH - 1) / mgr::kTileH, g.B);
  const size_t smem = sizeof(mgr::TileAffine) * g.L;
  if (theta)
    mgr::render_fwd_direct<T, true><<<grid, mgr::kDirectThreads, smem, s>>>((const T*)x, theta, (T*)out, g);
  else
    mgr::render_fwd_direct<T, false><<<grid, mgr::kDirectThreads, 0, s>>>((const T*)x, nullptr, (T*)out, g);
  MGR_CUDA(cudaGetLastError());
  count_launch();
  return MGR_OK;
}

template <typename T, int LMAX, bool kWarp>
int launch_backward_l(const void* x, const float* theta, const void* out, const void* gout, float* gx32,
                      void* gx, float* gtheta, const mgr::Geometry& g, int flags, cudaStream_t s) {
  dim3 grid((g.W + mgr::kTileW - 1) / mgr::kTileW, (g.H + mgr::kTileH - 1) / mgr::kTileH, g.B);
  const size_t smem = kWarp ? (sizeof(mgr::TileAffine) + 6 * sizeof(float)) * g.L : 0;
  const bool nx = flags & MGR_NEED_GRAD_X, nt = kWarp && (flags & MGR_NEED_GRAD_THETA);
#define MGR_LAUNCH(NX, NT)                                                                              \
  mgr::render_bwd_direct<T, LMAX, kWarp, NX, NT><<<grid, mgr::kDirectThreads, smem, s>>>(               \
      (const T*)x, theta, (const T*)out, (const T*)gout, gx32, (T*)gx, gtheta, g)
  if (nx && nt) MGR_LAUNCH(true, true);
  else if (nx) MGR_LAUNCH(true, false);
  else if (nt) MGR_LAUNCH(false, true);
#undef MGR_LAUNCH
  MGR_CUDA(cudaGetLastError());
  count_launch();
  return MGR_OK;
}

template <typename T, bool kWarp>
int launch_backward(const void* x, const float* theta, const void* out, const void* gout, float* gx32, void* gx,
                    float* gtheta, const mgr::Geometry& g, int flags, cudaStream_t s) {
  if (g.L <= 8) return launch_backward_l<T, 8, kWarp>(x, theta, out, gout, gx32, gx, gtheta, g, flags, s);
  return launch_backward_l<T, 32, kWarp>(x, theta, out, gout, gx32, gx, gtheta, g, flags, s);
}

template <typename T>
int backward_typed(const void* x, const float* theta, const void* out, const void* gout, void* gx, float* gtheta,
                   void* ws, const mgr::Geometry& g, int flags, cudaStream_t s) {
  const long long n = (long long)g.B * g.L * 4 * g.H * g.W;
  if (theta) {
    if (flags & MGR_NEED_GRAD_THETA) MGR_CUDA(cudaMemsetAsync(gtheta, 0, sizeof(float) * 6 * g.B * g.L, s));
    float* gx32 = nullptr;
    if (flags & MGR_NEED_GRAD_X) {
      gx32 = (sizeof(T) == 4) ? (float*)gx : (float*)ws;
      MGR_CUDA(cudaMemsetAsync(gx32, 0, sizeof(float) * n, s));
    }
    int rc = launch_backward<T, true>(x, theta, out, gout, gx32, gx, gtheta, g, flags, s);
    if (rc) return rc;
    if ((flags & MGR_NEED_GRAD_X) && sizeof(T) != 4) {
      const int threads = 256;
      const long long blocks = (n + threads - 1) / threads;
      mgr::cast_from_f32<T><<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), threads, 0, s>>>(gx32, (T*)gx, n);
      MGR_CUDA(cudaGetLastError());
      count_launch();
    }
    return MGR_OK;
  }
  return launch_backward<T, false>(x, nullptr, out, gout, nullptr, gx, nullptr, g, flags & MGR_NEED_GRAD_X, s);
}


}  // namespace mgr

#include "launchers_decl.h"
#define MGR_INSTANTIATE(SUFFIX, T)                                                                           \
  int mgr_fwd_##SUFFIX(const void* x, const float* theta, void* out, const mgr::Geometry& g, cudaStream_t s) { \
    return mgr::launch_forward<T>(x, theta, out, g, s);                                                      \
  }                                                                                                          \
  int mgr_bwd_##SUFFIX(const void* x, const float* theta, const void* out, const void* gout, void* gx,       \
                       float* gtheta, void* ws, const mgr::Geometry& g, int flags, cudaStream_t s) {         \
    return mgr::backward_typed<T>(x, theta, out, gout, gx, gtheta, ws, g, flags, s);                         \
  }
