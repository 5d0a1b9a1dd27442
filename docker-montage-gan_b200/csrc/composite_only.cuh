// Composite without a warp (theta == NULL): the real-image branch of the global discriminator
// (custom/loss_aio.py:313-320) and what AnalyticRenderer does to already-warped layers
// (diff_rendering/networks.py:36-44 signature; custom_utils/image_utils.py:112-163 arithmetic).  No gather: a pure
// streaming pass, so four adjacent pixels per thread and one 8- / 16-byte access per channel plane.  The math is that of
// render_direct.cuh (division-free adjoint); this is the forward's vector form for aligned tensors.
#pragma once
#include "mgr_common.cuh"

namespace mgr {

template <typename T>
__global__ void __launch_bounds__(256)
composite_fwd_vec(const T* __restrict__ x, T* __restrict__ out, Geometry g) {
  const int wv = g.W / 4;
  const long long hw = (long long)g.H * g.W;
  const long long total = (long long)g.B * g.H * wv;
  const float shift = g.m11 ? 1.f : 0.f, scale = g.m11 ? 0.5f : 1.f;
  const float os = g.m11 ? 2.f : 1.f, ob = g.m11 ? -1.f : 0.f;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int jv = (int)(k % wv);
    const int i = (int)((k / wv) % g.H);
    const long long b = k / ((long long)wv * g.H);
    const T* xb = x + b * g.sb + (long long)i * g.sh + 4 * jv;
    float S[3][4] = {}, R[4] = {};
    for (int l = 0; l < g.L; ++l) {
      const T* xl = xb + (long long)l * g.sl;
      float z[4][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ld_vec<T, 4>(xl + c * g.sc, z[c]);
#pragma unroll
        for (int q = 0; q < 4; ++q) z[c][q] = scale * (z[c][q] + shift);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float a = z[3][q], om = 1.f - a;
#pragma unroll
        for (int c = 0; c < 3; ++c) S[c][q] = fmaf(om, S[c][q], a * z[c][q]);
        R[q] = fmaf(om, R[q], a);
      }
    }
    T* o = out + b * 4 * hw + (long long)i * g.W + 4 * jv;
    float v[4], inv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) inv[q] = (R[q] != 0.f) ? 1.f / R[q] : 0.f;       // nan_to_num(0/0) = 0 (image_utils.py:132)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = fmaf(S[c][q] * inv[q], os, ob);
      st_vec4<T>(o + c * hw, v);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = fmaf(R[q], os, ob);
    st_vec4<T>(o + 3 * hw, v);
  }
}

// (A four-pixel backward was measured no faster than render_bwd_direct's one pixel per thread -- 108 registers for the
//  per-layer transmittances leave too few loads in flight -- so the backward stays there: 3.1 / 5.6 TB/s bf16 / fp32.)

}  // namespace mgr
