// The callers either side of the fused renderer:
//   * materialised warp of every layer  -- what STNv2c / STNv2b return to their callers
//     (fukuwarai/networks.py:250-257, 219-225) and what random_position computes
//     (custom_utils/image_utils.py:281-294): snapshot / EMA / metrics paths still want the warped
//     layers themselves;
//   * translation -> 2x3 theta (image_utils.py:316-335, a B*L Python loop of tiny H2D copies there);
//   * centre-pad one generator output into the [B,L,4,H,W] canvas (image_utils.py:216-243).
// These are streaming / gather kernels off the critical path; they share the placement arithmetic
// of the renderer (mgr_common.cuh) so that warp() followed by a composite equals render().
#pragma once
#include "render_direct.cuh"

namespace mgr {

// ---- warp forward: one thread per output pixel of one layer ------------------------------------
template <typename T>
__global__ void __launch_bounds__(kDirectThreads)
warp_fwd_kernel(const T* __restrict__ x, const float* __restrict__ theta, T* __restrict__ out, Geometry g) {
  __shared__ TileAffine aff;
  const int n = blockIdx.z;                      // b * L + l
  const int b = n / g.L, l = n - b * g.L;
  const int j0 = blockIdx.x * kTileW, i0 = blockIdx.y * kTileH;
  const int dj = threadIdx.x % kTileW, di = threadIdx.x / kTileW;
  if (threadIdx.x == 0) aff = make_tile_affine(theta + (long long)n * 6, g.H, g.W, j0, i0);
  __syncthreads();
  const int j = j0 + dj, i = i0 + di;
  if (j >= g.W || i >= g.H) return;
  const float shift = g.m11 ? 1.f : 0.f;
  const T* img = x + (long long)b * g.sb + (long long)l * g.sl;
  const Taps p = make_taps(aff, dj, di, g.H, g.W, g.sh);
  float z[4];
  sample_rgba(img, g.sc, p, shift, 1.f, z);      // sum_k w_k (x_k + shift), zeros padding
  const long long hw = (long long)g.H * g.W;
  T* q = out + (long long)n * 4 * hw + (long long)i * g.W + j;
#pragma unroll
  for (int c = 0; c < 4; ++c) st(q + c * hw, z[c] - shift);   // STNv2c: grid_sample(x + 1) - 1
}

// ---- warp backward: scatter to grad_x (fp32 atomics) and reduce grad_theta ------------------------
template <typename T, bool kNeedX, bool kNeedTheta>
__global__ void __launch_bounds__(kDirectThreads)
warp_bwd_kernel(const T* __restrict__ x, const float* __restrict__ theta, const T* __restrict__ gout,
                float* __restrict__ gx32, float* __restrict__ gtheta, Geometry g) {
  __shared__ TileAffine aff;
  __shared__ float acc[6];
  const int n = blockIdx.z;
  const int b = n / g.L, l = n - b * g.L;
  const int j0 = blockIdx.x * kTileW, i0 = blockIdx.y * kTileH;
  const int dj = threadIdx.x % kTileW, di = threadIdx.x / kTileW;
  if (threadIdx.x == 0) aff = make_tile_affine(theta + (long long)n * 6, g.H, g.W, j0, i0);
  if (threadIdx.x < 6) acc[threadIdx.x] = 0.f;
  __syncthreads();
  const int j = j0 + dj, i = i0 + di;
  const bool live = j < g.W && i < g.H;
  const long long hw = (long long)g.H * g.W;
  float part[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (live) {
    const float shift = g.m11 ? 1.f : 0.f;
    const T* img = x + (long long)b * g.sb + (long long)l * g.sl;
    const Taps p = make_taps(aff, dj, di, g.H, g.W, g.sh);
    const float ex = 1.f - p.fx, ey = 1.f - p.fy;
    float dix = 0.f, diy = 0.f;
    const T* go = gout + (long long)n * 4 * hw + (long long)i * g.W + j;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float gc = ld(go + c * hw);
      if (kNeedX) {
        float* q = gx32 + ((long long)n * 4 + c) * hw + (long long)p.y0 * g.W + p.x0;
        if (p.mask & 1u) atomicAdd(q, gc * p.w00);
        if (p.mask & 2u) atomicAdd(q + 1, gc * p.w01);
        if (p.mask & 4u) atomicAdd(q + g.W, gc * p.w10);
        if (p.mask & 8u) atomicAdd(q + g.W + 1, gc * p.w11);
      }
      if (kNeedTheta) {
        const T* pl = img + c * g.sc;
        const float v00 = (p.mask & 1u) ? ld(pl + p.o00) + shift : 0.f;
        const float v01 = (p.mask & 2u) ? ld(pl + p.o01) + shift : 0.f;
        const float v10 = (p.mask & 4u) ? ld(pl + p.o10) + shift : 0.f;
        const float v11 = (p.mask & 8u) ? ld(pl + p.o11) + shift : 0.f;
        dix = fmaf(gc, (v01 - v00) * ey + (v11 - v10) * p.fy, dix);
        diy = fmaf(gc, (v10 - v00) * ex + (v11 - v01) * p.fx, diy);
      }
    }
    if (kNeedTheta) {
      const float ggx = dix * (0.5f * g.W), ggy = diy * (0.5f * g.H);
      const float xj = norm_coord(j, g.W), yi = norm_coord(i, g.H);
      part[0] = ggx * xj; part[1] = ggx * yi; part[2] = ggx;
      part[3] = ggy * xj; part[4] = ggy * yi; part[5] = ggy;
    }
  }
  if (kNeedTheta) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const float s = warp_sum(part[k]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&acc[k], s);
    }
    __syncthreads();
    if (threadIdx.x < 6) atomicAdd(gtheta + (long long)n * 6 + threadIdx.x, acc[threadIdx.x]);
  }
}

// ---- translation [n,2] -> theta [n,2,3] = [[1,0,dx],[0,1,dy]] --------------------------------------
static __global__ void translation_to_theta_kernel(const float* __restrict__ tr, float* __restrict__ theta, long long n) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const float dx = tr[2 * k], dy = tr[2 * k + 1];
  float* t = theta + 6 * k;
  t[0] = 1.f; t[1] = 0.f; t[2] = dx;
  t[3] = 0.f; t[4] = 1.f; t[5] = dy;
}

// ---- centre-pad one layer's generator output [B,4,h,w] into the canvas dst[:, l] of [B,L,4,H,W] ----
// Grid (ceil(W / 256), ceil(H / 4), B * 4): a thread owns four adjacent canvas pixels of one row (no divisions;
// one 8- / 16-byte store when kVec4, i.e. W % 4 == 0 and an aligned dst).
template <typename T, bool kVec4>
__global__ void __launch_bounds__(256)
pad_stack_kernel(const T* __restrict__ src, long long ssb, long long ssc, long long ssh, long long ssw,
                 T* __restrict__ dst, int L, int l, int h, int w, int H, int W, float pad) {
  const int top = (H - h) / 2, left = (W - w) / 2;            // pad_256: pad_x1 = pad_x // 2 (image_utils.py:222-225)
  const int X0 = 4 * (blockIdx.x * 64 + (threadIdx.x & 63)), Y = blockIdx.y * 4 + (threadIdx.x >> 6);
  const int b = blockIdx.z >> 2, c = blockIdx.z & 3;
  if (X0 >= W || Y >= H) return;
  const int y = Y - top;
  const bool yin = (unsigned)y < (unsigned)h;
  const T* srow = src + b * ssb + c * ssc + (long long)y * ssh;
  __align__(16) T v[4];
  T padv;
  st(&padv, pad);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int xx = X0 + q - left;
    v[q] = (yin && (unsigned)xx < (unsigned)w) ? srow[xx * ssw] : padv;
  }
  T* q0 = dst + ((((long long)b * L + l) * 4 + c) * H + Y) * W + X0;
  if (kVec4) {
    if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(q0) = *reinterpret_cast<const float4*>(v);
    else *reinterpret_cast<uint2*>(q0) = *reinterpret_cast<const uint2*>(v);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (X0 + q < W) q0[q] = v[q];
  }
}

// ---- forward-mode derivative (JVP) of the composite: d out along a tangent of x ------------------------
// Needed for the double backward of the composite-only branch (R1 penalty on the real layers,
// custom/loss_aio.py:327-338): the backward is linear in grad_out, so its adjoint w.r.t. grad_out is this JVP.
//   S' = a' c + a c' - a' S + (1-a) S',  R' = a' - a' R + (1-a) R'  (back -> front),  o = S/R:
//   o'_rgb = S'/R - S R'/R^2 (0 where R == 0),  o'_a = R'
template <typename T>
__global__ void composite_jvp_kernel(const T* __restrict__ x, const T* __restrict__ tx, T* __restrict__ tout, Geometry g) {
  const long long hw = (long long)g.H * g.W;
  const long long total = (long long)g.B * hw;
  const float zs = g.m11 ? 0.5f : 1.f, zb = g.m11 ? 0.5f : 0.f;   // z = zs x + zb, z' = zs x', out' = o' / zs
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const long long b = k / hw, pix = k - b * hw;
    const int i = (int)(pix / g.W), j = (int)(pix - (long long)i * g.W);
    const T* xb = x + b * g.sb + (long long)i * g.sh + j;
    const T* tb = tx + (b * g.L * 4) * hw + pix;               // tangent is contiguous
    float S[3] = {0.f, 0.f, 0.f}, dS[3] = {0.f, 0.f, 0.f}, R = 0.f, dR = 0.f;
    float c[3] = {0.f, 0.f, 0.f}, dc[3] = {0.f, 0.f, 0.f}, a = 0.f, da = 0.f;
    for (int l = 0; l < g.L; ++l) {
      const T* xl = xb + (long long)l * g.sl;
      const T* tl = tb + (long long)l * 4 * hw;
      a = fmaf(ld(xl + 3 * g.sc), zs, zb); da = zs * ld(tl + 3 * hw);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        c[q] = fmaf(ld(xl + q * g.sc), zs, zb); dc[q] = zs * ld(tl + q * hw);
        dS[q] = da * (c[q] - S[q]) + a * dc[q] + (1.f - a) * dS[q];
        S[q] = fmaf(1.f - a, S[q], a * c[q]);
      }
      dR = da * (1.f - R) + (1.f - a) * dR;
      R = fmaf(1.f - a, R, a);
    }
    const float os = g.m11 ? 2.f : 1.f;                         // out = os o + const
    T* o = tout + b * 4 * hw + pix;
    if (g.L == 1) {                                             // single layer is returned untouched
#pragma unroll
      for (int q = 0; q < 3; ++q) st(o + q * hw, os * dc[q]);
      st(o + 3 * hw, os * da);
    } else {
      const float inv = (R != 0.f) ? 1.f / R : 0.f;
#pragma unroll
      for (int q = 0; q < 3; ++q) st(o + q * hw, os * (dS[q] - S[q] * inv * dR) * inv);
      st(o + 3 * hw, os * dR);
    }
  }
}

}  // namespace mgr
