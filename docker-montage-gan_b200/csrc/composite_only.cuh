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

// Backward for 16-bit tensors, two pixels per thread (one 32-bit access per plane; render_bwd_direct's one pixel per thread
// issues 16-bit loads and stores, half a sector per warp request).  Four pixels per thread were measured no faster than
// one: the per-layer transmittances then need ~108 registers and too few loads stay in flight.
//   sweep 1 (front -> back): T_l in front of every layer and A = sum T_l a_l from the alpha plane alone;
//   sweep 2 (back -> front): d c_l = G_P T_l a_l,  d a_l = T_l [G_P.(c_l - S_l) + G_A (1 - R_l)]
template <typename T, int LMAX>
__global__ void __launch_bounds__(256)
composite_bwd_vec2(const T* __restrict__ x, const T* __restrict__ out, const T* __restrict__ gout, T* __restrict__ gx, Geometry g) {
  constexpr int kV = 2;
  const int wv = g.W / kV;
  const long long hw = (long long)g.H * g.W;
  const long long total = (long long)g.B * g.H * wv;
  const float shift = g.m11 ? 1.f : 0.f, scale = g.m11 ? 0.5f : 1.f;
  const float gs = g.m11 ? 2.f : 1.f, is = g.m11 ? 0.5f : 1.f, ib = g.m11 ? 0.5f : 0.f;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int jv = (int)(k % wv);
    const int i = (int)((k / wv) % g.H);
    const long long b = k / ((long long)wv * g.H);
    const T* xb = x + b * g.sb + (long long)i * g.sh + kV * jv;
    float Tl[LMAX][kV];
    float Tc[kV] = {1.f, 1.f}, A[kV] = {};
#pragma unroll
    for (int l = LMAX - 1; l >= 0; --l) {
      if (l < g.L) {
        float a[kV];
        ld_vec<T, kV>(xb + (long long)l * g.sl + 3 * g.sc, a);
#pragma unroll
        for (int q = 0; q < kV; ++q) {
          const float aq = scale * (a[q] + shift);
          Tl[l][q] = Tc[q];
          A[q] = fmaf(Tc[q], aq, A[q]);
          Tc[q] *= (1.f - aq);
        }
      }
    }
    const long long po = b * 4 * hw + (long long)i * g.W + kV * jv;
    float G[4][kV], GA[kV];                     // G[c]: G_P (c < 3), G[3]: upstream alpha gradient
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      ld_vec<T, kV>(gout + po + c * hw, G[c]);
#pragma unroll
      for (int q = 0; q < kV; ++q) G[c][q] *= gs;
    }
    {
      float o[3][kV];
#pragma unroll
      for (int c = 0; c < 3; ++c) ld_vec<T, kV>(out + po + c * hw, o[c]);
#pragma unroll
      for (int q = 0; q < kV; ++q) {
        if (A[q] != 0.f) {                      // A == 0: every gradient is defined as 0
          const float inv = 1.f / A[q];
          const float dot = G[0][q] * fmaf(o[0][q], is, ib) + G[1][q] * fmaf(o[1][q], is, ib) + G[2][q] * fmaf(o[2][q], is, ib);
          GA[q] = G[3][q] - dot * inv;
          G[0][q] *= inv; G[1][q] *= inv; G[2][q] *= inv;
        } else {
          GA[q] = 0.f; G[0][q] = G[1][q] = G[2][q] = 0.f;
        }
      }
    }
    float S[3][kV] = {}, R[kV] = {};
    T* gb = gx + (b * g.L * 4) * hw + (long long)i * g.W + kV * jv;      // grad_x is contiguous [B,L,4,H,W]
#pragma unroll
    for (int l = 0; l < LMAX; ++l) {
      if (l < g.L) {
        const T* xl = xb + (long long)l * g.sl;
        float z[4][kV], gz[4][kV];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          ld_vec<T, kV>(xl + c * g.sc, z[c]);
#pragma unroll
          for (int q = 0; q < kV; ++q) z[c][q] = scale * (z[c][q] + shift);
        }
#pragma unroll
        for (int q = 0; q < kV; ++q) {
          const float a = z[3][q], om = 1.f - a, T_l = Tl[l][q], ta = T_l * a;
          gz[0][q] = G[0][q] * ta; gz[1][q] = G[1][q] * ta; gz[2][q] = G[2][q] * ta;
          gz[3][q] = T_l * (G[0][q] * (z[0][q] - S[0][q]) + G[1][q] * (z[1][q] - S[1][q]) + G[2][q] * (z[2][q] - S[2][q]) +
                            GA[q] * (1.f - R[q]));
#pragma unroll
          for (int c = 0; c < 3; ++c) S[c][q] = fmaf(om, S[c][q], a * z[c][q]);
          R[q] = fmaf(om, R[q], a);
        }
        T* gl = gb + (long long)l * 4 * hw;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float v[kV];
#pragma unroll
          for (int q = 0; q < kV; ++q) v[q] = scale * gz[c][q];             // d z / d x = scale
          st_vec2(gl + c * hw, v);
        }
      }
    }
  }
}

}  // namespace mgr
