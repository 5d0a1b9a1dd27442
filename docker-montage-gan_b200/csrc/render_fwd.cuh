// The forward kernel: one launch, two code paths chosen per CTA from the sample's placements -- the stencil path for
// stacks of pure translations (render_shift.cuh), the general tiled path otherwise (render_tiled.cuh).  One launch
// instead of one per path: a path that has nothing to do in a batch used to cost ~10 us of empty CTAs.
#pragma once
#include "render_shift.cuh"
#include "render_tiled.cuh"

namespace mgr {

template <typename T, bool kSave, bool kRagged>
__global__ void __launch_bounds__(kTiledThreads, MGR_FWD_BLOCKS)
render_fwd(const T* __restrict__ x, const __grid_constant__ SrcLayers src, const float* __restrict__ theta, T* __restrict__ out,
           typename SavedAlpha<T>::type* __restrict__ sav, Geometry g, int use_stencil) {
  if (use_stencil && cta_all_shift(theta + (long long)blockIdx.z * g.L * 6, g.L, threadIdx.x, kTiledThreads))
    fwd_shift_body<T, kSave>(src, theta, out, sav, g);
  else
    fwd_tiled_body<T, kSave, kRagged>(x, src, theta, out, sav, g);
}

// The two paths as launches of their own: fp32 stacks keep them apart, because the merged launch must reserve the
// general path's 45 KB staging buffer for every CTA and the fp32 stencil forward (the kernel closest to the HBM roof,
// 56-58 % of peak at 512 x 512) then loses 13 % -- less L1 next to the larger shared-memory carve-out.
template <typename T, bool kSave, bool kRagged>
__global__ void __launch_bounds__(kTiledThreads, MGR_FWD_BLOCKS)
render_fwd_general_only(const T* __restrict__ x, const __grid_constant__ SrcLayers src, const float* __restrict__ theta,
                        T* __restrict__ out, typename SavedAlpha<T>::type* __restrict__ sav, Geometry g, int use_stencil) {
  if (use_stencil && cta_all_shift(theta + (long long)blockIdx.z * g.L * 6, g.L, threadIdx.x, kTiledThreads)) return;
  fwd_tiled_body<T, kSave, kRagged>(x, src, theta, out, sav, g);
}
template <typename T, bool kSave>
__global__ void __launch_bounds__(kTiledThreads, MGR_SHF_BLOCKS)
render_fwd_stencil_only(const __grid_constant__ SrcLayers src, const float* __restrict__ theta, T* __restrict__ out,
                        typename SavedAlpha<T>::type* __restrict__ sav, Geometry g, const int* __restrict__ shift_flags) {
  // is the sample this kernel's?  One load of the forward's per-sample flag if there is one, else every warp finds out
  // by itself from the placements (no CTA barrier before an idle CTA leaves)
  if (shift_flags ? shift_flags[blockIdx.z] == 0
                  : !__all_sync(0xffffffffu, (int)(threadIdx.x & 31) >= g.L || is_pure_shift(theta + ((long long)blockIdx.z * g.L + (threadIdx.x & 31)) * 6)))
    return;
  fwd_shift_body<T, kSave>(src, theta, out, sav, g);
}

}  // namespace mgr
