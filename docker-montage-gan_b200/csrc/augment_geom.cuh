// AugmentPipe's geometric execution block (reference training/augment.py:306-342 with the FIR
// arithmetic of torch_utils/ops/upfirdn2d.py:168-222): reflect pad -> x2 upsample with the 12-tap sym6 low-pass ->
// affine bilinear resampling onto a 2(H+6) x 2(W+6) grid -> low-pass, decimate x2, crop to H x W.  The chain is linear
// in the images, so the backward is the chain of adjoints and needs no saved activations.  fp32 (the reference runs
// its augmentations in fp32).  First cut: one thread per output element, gather form everywhere except the adjoint
// of the resampling (fp32 atomics); these kernels sit between the renderer and the discriminator, off this round's
// tuned path.  Oracle: oracle/augment_geom.py, golden vectors from the reference pipe itself.
#pragma once
#include "mgr_common.cuh"

namespace mgr {

// sym6 decomposition low-pass, normalised to unit DC gain in fp32 exactly as upfirdn2d.setup_filter does (f / f.sum())
__constant__ float c_sym6[12] = {0x1.64eba6p-7f, 0x1.43869ap-9f, -0x1.55bc6p-4f, -0x1.17d9fcp-5f, 0x1.638ffcp-2f, 0x1.1d2814p-1f,
                                 0x1.e95fc2p-3f, -0x1.a4c2eep-5f, -0x1.e7fa18p-7f, 0x1.031304p-5f, 0x1.47ab76p-10f, -0x1.697e02p-8f};
constexpr int kFir = 12;

struct AugGeom {
  int B, C, H, W;            // images
  int mx0, my0, mx1, my1;    // reflect padding
  int Hp, Wp;                // padded image        H + my0 + my1, W + mx0 + mx1
  int Hu, Wu;                // upsampled           2 Hp, 2 Wp
  int Hs, Ws;                // resampling grid     2 (H + 6), 2 (W + 6)
};

__device__ __forceinline__ int reflect_index(int p, int n) { return p < 0 ? -p : (p >= n ? 2 * (n - 1) - p : p); }

// U[Y, X] = sum_k sum_l g[k] g[l] u0[Y + k - 6, X + l - 6],  g = 2 * flip(f),  u0[2i, 2j] = reflect-padded image, else 0
static __global__ void __launch_bounds__(256)
aug_up_fwd(const float* __restrict__ img, float* __restrict__ U, AugGeom a) {
  const long long total = (long long)a.B * a.C * a.Hu * a.Wu;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(k % a.Wu), Y = (int)((k / a.Wu) % a.Hu);
    const long long bc = k / ((long long)a.Wu * a.Hu);
    const float* p = img + bc * a.H * a.W;
    float acc = 0.f;
    for (int ky = Y & 1; ky < kFir; ky += 2) {                   // taps that land on an even (= non-zero) row
      const int ty = Y + ky - 6;
      if (ty < 0 || ty >= a.Hu) continue;
      const int iy = reflect_index((ty >> 1) - a.my0, a.H);
      const float gy = 2.f * c_sym6[kFir - 1 - ky];
      float row = 0.f;
      for (int kx = X & 1; kx < kFir; kx += 2) {
        const int tx = X + kx - 6;
        if (tx < 0 || tx >= a.Wu) continue;
        const int ix = reflect_index((tx >> 1) - a.mx0, a.W);
        row = fmaf(2.f * c_sym6[kFir - 1 - kx], p[iy * a.W + ix], row);
      }
      acc = fmaf(gy, row, acc);
    }
    U[k] = acc;
  }
}

// pixel (Xs, Ys) of the resampling grid -> source coordinates in U (affine_grid + grid_sample unnormalisation)
__device__ __forceinline__ void aug_source_coords(const float* __restrict__ th, const AugGeom& a, int Xs, int Ys, float& ix, float& iy) {
  const float xn = (2.f * Xs + 1.f) / a.Ws - 1.f, yn = (2.f * Ys + 1.f) / a.Hs - 1.f;
  const float gx = fmaf(th[0], xn, fmaf(th[1], yn, th[2])), gy = fmaf(th[3], xn, fmaf(th[4], yn, th[5]));
  ix = ((gx + 1.f) * a.Wu - 1.f) * 0.5f;
  iy = ((gy + 1.f) * a.Hu - 1.f) * 0.5f;
}

template <bool kAdjoint>     // false: S = sample(U);  true: gU += sample^T(gS)  (gU zeroed by the caller)
static __global__ void __launch_bounds__(256)
aug_sample(const float* __restrict__ theta, const float* __restrict__ src, float* __restrict__ dst, AugGeom a) {
  const long long total = (long long)a.B * a.Hs * a.Ws;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int Xs = (int)(k % a.Ws), Ys = (int)((k / a.Ws) % a.Hs);
    const int b = (int)(k / ((long long)a.Ws * a.Hs));
    float ix, iy;
    aug_source_coords(theta + b * 6, a, Xs, Ys, ix, iy);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float fx = ix - fx0, fy = iy - fy0;
    const int x0 = (int)fminf(fmaxf(fx0, -2.f), (float)a.Wu), y0 = (int)fminf(fmaxf(fy0, -2.f), (float)a.Hu);
    const bool xin0 = (unsigned)x0 < (unsigned)a.Wu, xin1 = (unsigned)(x0 + 1) < (unsigned)a.Wu;
    const bool yin0 = (unsigned)y0 < (unsigned)a.Hu, yin1 = (unsigned)(y0 + 1) < (unsigned)a.Hu;
    const float w00 = (1.f - fx) * (1.f - fy), w01 = fx * (1.f - fy), w10 = (1.f - fx) * fy, w11 = fx * fy;
    const long long uo = (long long)y0 * a.Wu + x0;
    for (int c = 0; c < a.C; ++c) {
      const long long ub = ((long long)b * a.C + c) * a.Hu * a.Wu + uo;
      const long long so = ((long long)b * a.C + c) * a.Hs * a.Ws + (long long)Ys * a.Ws + Xs;
      if (!kAdjoint) {
        float v = 0.f;
        if (xin0 && yin0) v = fmaf(w00, src[ub], v);
        if (xin1 && yin0) v = fmaf(w01, src[ub + 1], v);
        if (xin0 && yin1) v = fmaf(w10, src[ub + a.Wu], v);
        if (xin1 && yin1) v = fmaf(w11, src[ub + a.Wu + 1], v);
        dst[so] = v;
      } else {
        const float g = src[so];
        if (xin0 && yin0) atomicAdd(dst + ub, w00 * g);
        if (xin1 && yin0) atomicAdd(dst + ub + 1, w01 * g);
        if (xin0 && yin1) atomicAdd(dst + ub + a.Wu, w10 * g);
        if (xin1 && yin1) atomicAdd(dst + ub + a.Wu + 1, w11 * g);
      }
    }
  }
}

// out[y, x] = sum_k sum_l f[k] f[l] S[2 y + k + 1, 2 x + l + 1]      (crop 1, correlate, keep every second sample)
static __global__ void __launch_bounds__(256)
aug_down_fwd(const float* __restrict__ S, float* __restrict__ out, AugGeom a) {
  const long long total = (long long)a.B * a.C * a.H * a.W;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(k % a.W), y = (int)((k / a.W) % a.H);
    const long long bc = k / ((long long)a.W * a.H);
    const float* p = S + bc * a.Hs * a.Ws + (long long)(2 * y + 1) * a.Ws + (2 * x + 1);
    float acc = 0.f;
    for (int ky = 0; ky < kFir; ++ky) {
      float row = 0.f;
#pragma unroll
      for (int kx = 0; kx < kFir; ++kx) row = fmaf(c_sym6[kx], p[ky * a.Ws + kx], row);
      acc = fmaf(c_sym6[ky], row, acc);
    }
    out[k] = acc;
  }
}

// adjoint of aug_down_fwd: gS[Y, X] = sum over (y, k): 2 y + k + 1 = Y, (x, l): 2 x + l + 1 = X of f[k] f[l] gout[y, x]
static __global__ void __launch_bounds__(256)
aug_down_bwd(const float* __restrict__ gout, float* __restrict__ gS, AugGeom a) {
  const long long total = (long long)a.B * a.C * a.Hs * a.Ws;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(k % a.Ws), Y = (int)((k / a.Ws) % a.Hs);
    const long long bc = k / ((long long)a.Ws * a.Hs);
    const float* p = gout + bc * a.H * a.W;
    float acc = 0.f;
    for (int ky = (Y - 1) & 1; ky < kFir; ky += 2) {             // 2 y = Y - 1 - ky must be even and inside
      const int y2 = Y - 1 - ky;
      if (y2 < 0 || (y2 >> 1) >= a.H) continue;
      float row = 0.f;
      for (int kx = (X - 1) & 1; kx < kFir; kx += 2) {
        const int x2 = X - 1 - kx;
        if (x2 < 0 || (x2 >> 1) >= a.W) continue;
        row = fmaf(c_sym6[kx], p[(y2 >> 1) * a.W + (x2 >> 1)], row);
      }
      acc = fmaf(c_sym6[ky], row, acc);
    }
    gS[k] = acc;
  }
}

// adjoint of aug_up_fwd, first half: gradient w.r.t. the reflect-PADDED image,
//   gxp[i, j] = sum_k sum_l g[k] g[l] gU[2 i + 6 - k, 2 j + 6 - l]
static __global__ void __launch_bounds__(256)
aug_up_bwd_padded(const float* __restrict__ gU, float* __restrict__ gxp, AugGeom a) {
  const long long total = (long long)a.B * a.C * a.Hp * a.Wp;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(k % a.Wp), i = (int)((k / a.Wp) % a.Hp);
    const long long bc = k / ((long long)a.Wp * a.Hp);
    const float* p = gU + bc * a.Hu * a.Wu;
    float acc = 0.f;
    for (int ky = 0; ky < kFir; ++ky) {
      const int Y = 2 * i + 6 - ky;
      if (Y < 0 || Y >= a.Hu) continue;
      float row = 0.f;
      for (int kx = 0; kx < kFir; ++kx) {
        const int X = 2 * j + 6 - kx;
        if (X < 0 || X >= a.Wu) continue;
        row = fmaf(2.f * c_sym6[kFir - 1 - kx], p[(long long)Y * a.Wu + X], row);
      }
      acc = fmaf(2.f * c_sym6[kFir - 1 - ky], row, acc);
    }
    gxp[k] = acc;
  }
}

// second half: fold the reflect padding back, gimg[y, x] = sum of gxp over the padded positions that mirror onto (y, x)
__device__ __forceinline__ int reflect_sources(int x, int n, int m0, int m1, int (&src)[3]) {
  int cnt = 0;
  src[cnt++] = x + m0;                                          // the pixel itself
  if (x >= 1 && x <= m0) src[cnt++] = m0 - x;                   // mirrored across the first pixel
  if (x <= n - 2 && x >= n - 1 - m1) src[cnt++] = m0 + 2 * (n - 1) - x;   // mirrored across the last pixel
  return cnt;
}
static __global__ void __launch_bounds__(256)
aug_fold_reflect(const float* __restrict__ gxp, float* __restrict__ gimg, AugGeom a) {
  const long long total = (long long)a.B * a.C * a.H * a.W;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(k % a.W), y = (int)((k / a.W) % a.H);
    const long long bc = k / ((long long)a.W * a.H);
    int sx[3], sy[3];
    const int nx = reflect_sources(x, a.W, a.mx0, a.mx1, sx), ny = reflect_sources(y, a.H, a.my0, a.my1, sy);
    const float* p = gxp + bc * a.Hp * a.Wp;
    float acc = 0.f;
    for (int q = 0; q < ny; ++q)
      for (int r = 0; r < nx; ++r) acc += p[(long long)sy[q] * a.Wp + sx[r]];
    gimg[k] = acc;
  }
}

}  // namespace mgr
