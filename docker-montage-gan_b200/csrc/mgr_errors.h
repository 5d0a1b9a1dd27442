// Error plumbing shared by every translation unit of libmontage_render.so.
#pragma once
#include <cuda_runtime.h>
#include "montage_render.h"

namespace mgr {
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);
int debug_path();                // 0 auto, 1 force the direct-gather kernels (tests / A-B timing)   // process-wide count of kernels launched by this library
}  // namespace mgr

#define MGR_CUDA(expr)                                        \
  do {                                                        \
    cudaError_t e_ = (expr);                                  \
    if (e_ != cudaSuccess) return mgr::cuda_fail(e_, #expr);  \
  } while (0)
