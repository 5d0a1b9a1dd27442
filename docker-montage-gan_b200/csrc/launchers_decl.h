// Per-dtype launcher entry points (defined in inst_*.cu).
#pragma once
#include <cuda_runtime.h>
#include "mgr_common.cuh"
#define MGR_DECLARE(SUFFIX)                                                                                  \
  bool mgr_tiled_ok_##SUFFIX(const void* x, const mgr::Geometry& g);                                         \
  int mgr_fwd_##SUFFIX(const void* x, const float* theta, void* out, void* sav, const mgr::Geometry& g,     \
                       cudaStream_t s);                                                                      \
  int mgr_bwd_##SUFFIX(const void* x, const float* theta, const void* out, const void* gout, const void* sav, \
                       void* gx, float* gtheta, void* ws, const mgr::Geometry& g, int flags, cudaStream_t s); \
  int mgr_warp_fwd_##SUFFIX(const void* x, const float* theta, void* out, const mgr::Geometry& g, cudaStream_t s); \
  int mgr_warp_bwd_##SUFFIX(const void* x, const float* theta, const void* gout, void* gx, float* gtheta,    \
                            void* ws, const mgr::Geometry& g, int flags, cudaStream_t s);                    \
  int mgr_pad_stack_##SUFFIX(const void* src, const long long* ss, void* dst, int B, int L, int l, int h,    \
                             int w, int H, int W, float pad, cudaStream_t s);                               \
  int mgr_jvp_##SUFFIX(const void* x, const void* tx, void* tout, const mgr::Geometry& g, cudaStream_t s);   \
  int mgr_pil_##SUFFIX(const void* x, float* of, unsigned char* ou, const mgr::Geometry& g, cudaStream_t s); \
  int mgr_fwd_ragged_##SUFFIX(const mgr::SrcLayers& src, const float* theta, void* out, void* sav,           \
                              const mgr::Geometry& g, cudaStream_t s);                                       \
  int mgr_bwd_ragged_##SUFFIX(const mgr::SrcLayers& src, const float* theta, const void* out, const void* gout, \
                              const void* sav, const mgr::DstLayers& dst, float* gtheta, void* ws,           \
                              const mgr::Geometry& g, int flags, cudaStream_t s);
MGR_DECLARE(f32)
MGR_DECLARE(bf16)
MGR_DECLARE(f16)
