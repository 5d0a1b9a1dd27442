// Error plumbing shared by every translation unit of libmontage_render.so.
#pragma once
#include <cuda_runtime.h>
#include "montage_render.h"

namespace mgr {
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);
// Fork / join onto a cached per-device side stream: the two kernel families that share a batch (general placements and
// pure translations) partition its samples, so their launches are independent; run back to back on one stream the family
// with nothing to do still costs its launch (4096 idle CTAs of 512 threads: 9-19 us per kernel, ncu).  fork() makes the
// side stream wait for everything queued on `s` so far; join() makes `s` wait for the side stream.  Both are plain event
// record / wait pairs, so they are legal under stream capture (the side stream joins the capture and leaves it at join()).
struct SideStream { cudaStream_t side; cudaEvent_t fork_ev, join_ev; };
int side_stream(SideStream* out);                    // cached per device and host thread; MGR_OK or an error code
int side_fork(const SideStream& ss, cudaStream_t s);
int side_join(const SideStream& ss, cudaStream_t s);
int debug_path();                // 0 auto, 1 force the direct-gather kernels (tests / A-B timing)   // process-wide count of kernels launched by this library
}  // namespace mgr

#define MGR_CUDA(expr)                                        \
  do {                                                        \
    cudaError_t e_ = (expr);                                  \
    if (e_ != cudaSuccess) return mgr::cuda_fail(e_, #expr);  \
  } while (0)
