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

// k -> (plane, row, col) of a [planes][rows][cols] array; 32-bit divisions whenever the index fits (64-bit ones cost
// more than the 12-tap filters these kernels run)
__device__ __forceinline__ void split3(long long k, int cols, int rows, int& c, int& r, long long& pl) {
  if (k < (1LL << 31)) {
    const unsigned kk = (unsigned)k, q = kk / (unsigned)cols, q2 = q / (unsigned)rows;
    c = (int)(kk - q * (unsigned)cols); r = (int)(q - q2 * (unsigned)rows); pl = q2;
  } else {
    const long long q = k / cols;
    c = (int)(k - q * cols); pl = q / rows; r = (int)(q - pl * rows);
  }
}

// Every FIR stage is separable (upfirdn2d.py:211-216 runs two 1-D convolutions as well): one kernel per axis, a thread
// per output element, the intermediate in the workspace.  n_out / n_in are the lengths along the filtered axis; the
// other axis has `other` elements; kAlongX selects which of the two innermost axes is filtered.
//   kUp:       dst[X] = sum_{k = X mod 2, step 2} g[k] src[reflect((X + k - 6) / 2 - m0)],  g = 2 flip(f)   (x2 upsample of
//              the reflect-padded signal; zeros between samples never get multiplied)
//   kDown:     dst[x] = sum_k f[k] src[2 x + k + 1]                                          (crop 1, correlate, decimate)
//   kDownAdj:  dst[X] = sum_{k: X - 1 - k even} f[k] src[(X - 1 - k) / 2]                    (adjoint of kDown)
//   kUpAdj:    dst[j] = sum_k g[k] src[2 j + 6 - k]            (adjoint of kUp w.r.t. the PADDED signal; fold follows)
enum { kUp = 0, kDown = 1, kDownAdj = 2, kUpAdj = 3 };
template <int kOp, bool kAlongX>
static __global__ void __launch_bounds__(256)
aug_fir_1d(const float* __restrict__ src, float* __restrict__ dst, long long planes, int n_in, int n_out, int other, int m0) {
  // layout: [planes][rows][cols]; kAlongX: rows = other, cols filtered; else rows filtered, cols = other
  const int cols_out = kAlongX ? n_out : other, rows_out = kAlongX ? other : n_out;
  const int cols_in = kAlongX ? n_in : other, rows_in = kAlongX ? other : n_in;
  const long long total = planes * rows_out * cols_out;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    int c, r;
    long long pl;
    split3(k, cols_out, rows_out, c, r, pl);
    const int o = kAlongX ? c : r;                             // position along the filtered axis
    const float* p = src + pl * rows_in * cols_in + (kAlongX ? (long long)r * cols_in : c);
    const int step = kAlongX ? 1 : cols_in;
    float acc = 0.f;
    if (kOp == kUp) {
      // every tap is evaluated (out-of-range ones with weight 0 on a clamped address): no data-dependent branches, so
      // the loads of one output issue back to back instead of one dependent round trip per tap
#pragma unroll
      for (int q = 0; q < kFir / 2; ++q) {
        const int t = (o & 1) + 2 * q;
        const int u = o + t - 6;                               // position in the zero-inserted padded signal (even here)
        const bool ok = u >= 0 && u < n_out;
        const int idx = min(max(reflect_index((max(u, 0) >> 1) - m0, n_in), 0), n_in - 1);
        acc = fmaf(ok ? 2.f * c_sym6[kFir - 1 - t] : 0.f, p[(long long)idx * step], acc);
      }
    } else if (kOp == kDown) {
#pragma unroll
      for (int t = 0; t < kFir; ++t) acc = fmaf(c_sym6[t], p[(long long)(2 * o + t + 1) * step], acc);
    } else if (kOp == kDownAdj) {                              // (the branch-free form was measured slower for the adjoints)
      for (int t = (o - 1) & 1; t < kFir; t += 2) {
        const int x2 = o - 1 - t;
        if (x2 < 0 || (x2 >> 1) >= n_in) continue;
        acc = fmaf(c_sym6[t], p[(long long)(x2 >> 1) * step], acc);
      }
    } else {
      for (int t = 0; t < kFir; ++t) {
        const int X = 2 * o + 6 - t;
        if (X < 0 || X >= n_in) continue;
        acc = fmaf(2.f * c_sym6[kFir - 1 - t], p[(long long)X * step], acc);
      }
    }
    dst[k] = acc;
  }
}

// The x2 upsample along y, register-blocked: a thread owns one column and kUpRows consecutive output rows, loads the
// kUpRows / 2 + 6 input rows they need once (reflect-indexed) and writes the outputs row by row (coalesced across the
// warp).  The generic pass above re-reads six input rows per output row and was 54 % of the forward.
constexpr int kUpRows = 8;
static __global__ void __launch_bounds__(256)
aug_up_y_blocked(const float* __restrict__ src, float* __restrict__ dst, long long planes, int n_in, int n_out, int cols, int m0) {
  const int row_blocks = (n_out + kUpRows - 1) / kUpRows;
  const long long total = planes * row_blocks * cols;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    int c, rb;
    long long pl;
    split3(k, cols, row_blocks, c, rb, pl);
    const int Y0 = rb * kUpRows;                               // even
    const float* p = src + pl * n_in * cols + c;
    // padded-signal rows (Y0 - 6) / 2 .. (Y0 + kUpRows + 4) / 2 feed these outputs
    constexpr int kIn = kUpRows / 2 + 6;
    const int i0 = (Y0 - 6) / 2;                               // exact: Y0 - 6 is even
    const int np = n_out / 2;                                  // padded length
    float v[kIn];
#pragma unroll
    for (int q = 0; q < kIn; ++q) {
      const int i = i0 + q;
      v[q] = (i >= 0 && i < np) ? p[(long long)reflect_index(i - m0, n_in) * cols] : 0.f;
    }
    float* o = dst + pl * n_out * cols + (long long)Y0 * cols + c;
#pragma unroll
    for (int dy = 0; dy < kUpRows; ++dy) {
      if (Y0 + dy >= n_out) break;
      float acc = 0.f;
#pragma unroll
      for (int t = dy & 1; t < kFir; t += 2)                   // u = Y0 + dy + t - 6 even -> padded row (u / 2) = i0 + (dy + t) / 2
        acc = fmaf(2.f * c_sym6[kFir - 1 - t], v[(dy + t) >> 1], acc);
      o[(long long)dy * cols] = acc;
    }
  }
}

// The two adjoint passes along y, blocked the same way.
//   kDownAdj: outputs Y0 .. Y0 + 7 (Y0 even) read input rows (Y0 - 12) / 2 .. (Y0 + 6) / 2     -> 10 rows for 8 outputs
//   kUpAdj:   outputs j0 .. j0 + 3 read input rows 2 j0 - 5 .. 2 j0 + 12                        -> 18 rows for 4 outputs
template <int kOp>
static __global__ void __launch_bounds__(256)
aug_adj_y_blocked(const float* __restrict__ src, float* __restrict__ dst, long long planes, int n_in, int n_out, int cols) {
  constexpr int kRows = kOp == kDownAdj ? 8 : 4;
  constexpr int kIn = kOp == kDownAdj ? 10 : 18;
  const int row_blocks = (n_out + kRows - 1) / kRows;
  const long long total = planes * row_blocks * cols;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    int c, rb;
    long long pl;
    split3(k, cols, row_blocks, c, rb, pl);
    const int o0 = rb * kRows;
    const float* p = src + pl * n_in * cols + c;
    const int i0 = kOp == kDownAdj ? (o0 - 12) / 2 : 2 * o0 - 5;
    float v[kIn];
#pragma unroll
    for (int q = 0; q < kIn; ++q) {
      const int i = i0 + q;
      v[q] = (i >= 0 && i < n_in) ? p[(long long)i * cols] : 0.f;
    }
    float* o = dst + pl * n_out * cols + (long long)o0 * cols + c;
#pragma unroll
    for (int d = 0; d < kRows; ++d) {
      if (o0 + d >= n_out) break;
      float acc = 0.f;
      if (kOp == kDownAdj) {
#pragma unroll
        for (int t = (d + 1) & 1; t < kFir; t += 2)            // x2 = o0 + d - 1 - t even; input row x2 / 2 = i0 + (d + 11 - t) / 2
          acc = fmaf(c_sym6[t], v[(d + 11 - t) >> 1], acc);
      } else {
#pragma unroll
        for (int t = 0; t < kFir; ++t)                         // input row 2 (o0 + d) + 6 - t = i0 + 2 d + 11 - t
          acc = fmaf(2.f * c_sym6[kFir - 1 - t], v[2 * d + 11 - t], acc);
      }
      o[(long long)d * cols] = acc;
    }
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
    int Xs, Ys;
    long long bl;
    split3(k, a.Ws, a.Hs, Xs, Ys, bl);
    const int b = (int)bl;
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

// fold the reflect padding back after the kUpAdj passes, gimg[y, x] = sum of gxp over the padded positions that mirror onto (y, x)
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
    int x, y;
    long long bc;
    split3(k, a.W, a.H, x, y, bc);
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
