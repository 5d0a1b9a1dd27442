// Per-dtype launcher entry points (defined in inst_*.cu).
#pragma once
#include <cuda_runtime.h>
#include "mgr_common.cuh"
#define MGR_DECLARE(SUFFIX)                                                                                  \
  int mgr_fwd_##SUFFIX(const void* x, const float* theta, void* out, void* sav, const mgr::Geometry& g,     \
                       cudaStream_t s);                                                                      \
  int mgr_bwd_##SUFFIX(const void* x, const float* theta, const void* out, const void* gout, const void* sav, \
                       void* gx, float* gtheta, void* ws, const mgr::Geometry& g, int flags, cudaStream_t s);
MGR_DECLARE(f32)
MGR_DECLARE(bf16)
MGR_DECLARE(f16)
