// Direct-gather render kernels: one thread per output pixel, taps read straight from global
// memory (L1/L2 catch the 2x2 reuse).  This is the general path -- any theta, any strides -- and
// the fallback of the tiled kernels when a layer's footprint does not fit shared memory.
//
// Math: SURVEY.md Appendix A.  Reference semantics: fukuwarai/networks.py:250-257 (warp),
// custom_utils/image_utils.py:128-146 (over), custom/loss_aio.py:251 (range shifts).
#pragma once
#include "mgr_common.cuh"

namespace mgr {

constexpr int kTileW = 32;   // output tile of one CTA: 32 x 8 pixels, one pixel per thread
constexpr int kTileH = 8;
constexpr int kDirectThreads = kTileW * kTileH;

// Sample one layer at this thread's pixel, in the [0,1] compositing domain.
//   m11:  z = 0.5 * sum_k w_k * (x_k + 1)   (out-of-bounds taps contribute 0: transparent black)
//   01 :  z =       sum_k w_k *  x_k
template <typename T>
__device__ __forceinline__ void sample_rgba(const T* __restrict__ img, long long sc, const Taps& p,
                                            float shift, float scale, float (&z)[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const T* pl = img + c * sc;
    const float v00 = (p.mask & 1u) ? ld(pl + p.o00) + shift : 0.f;
    const float v01 = (p.mask & 2u) ? ld(pl + p.o01) + shift : 0.f;
    const float v10 = (p.mask & 4u) ? ld(pl + p.o10) + shift : 0.f;
    const float v11 = (p.mask & 8u) ? ld(pl + p.o11) + shift : 0.f;
    z[c] = scale * fmaf(v11, p.w11, fmaf(v10, p.w10, fmaf(v01, p.w01, v00 * p.w00)));
  }
}

template <typename T>
__device__ __forceinline__ float sample_alpha(const T* __restrict__ img, long long sc, const Taps& p,
                                              float shift, float scale) {
  const T* pl = img + 3 * sc;
  const float v00 = (p.mask & 1u) ? ld(pl + p.o00) + shift : 0.f;
  const float v01 = (p.mask & 2u) ? ld(pl + p.o01) + shift : 0.f;
  const float v10 = (p.mask & 4u) ? ld(pl + p.o10) + shift : 0.f;
  const float v11 = (p.mask & 8u) ? ld(pl + p.o11) + shift : 0.f;
  return scale * fmaf(v11, p.w11, fmaf(v10, p.w10, fmaf(v01, p.w01, v00 * p.w00)));
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <typename T, bool kWarp>
__global__ void __launch_bounds__(kDirectThreads)
render_fwd_direct(const T* __restrict__ x, const float* __restrict__ theta, T* __restrict__ out, Geometry g) {
  extern __shared__ unsigned char smem_raw[];
  TileAffine* aff = reinterpret_cast<TileAffine*>(smem_raw);   // [L]
  const int b = blockIdx.z;
  const int j0 = blockIdx.x * kTileW, i0 = blockIdx.y * kTileH;
  const int dj = threadIdx.x % kTileW, di = threadIdx.x / kTileW;
  const int j = j0 + dj, i = i0 + di;
  if (kWarp) {
    for (int l = threadIdx.x; l < g.L; l += kDirectThreads)
      aff[l] = make_tile_affine(theta + ((long long)b * g.L + l) * 6, g.H, g.W, j0, i0);
    __syncthreads();
  }
  if (j >= g.W || i >= g.H) return;
  const float shift = g.m11 ? 1.f : 0.f, scale = g.m11 ? 0.5f : 1.f;
  const T* xb = x + (long long)b * g.sb;

  float S0 = 0.f, S1 = 0.f, S2 = 0.f, R = 0.f;    // premultiplied colour and alpha of the canvas
  float z[4] = {0.f, 0.f, 0.f, 0.f};
  for (int l = 0; l < g.L; ++l) {
    const T* img = xb + (long long)l * g.sl;
    if (kWarp) {
      const Taps p = make_taps(aff[l], dj, di, g.H, g.W, g.sh);
      sample_rgba(img, g.sc, p, shift, scale, z);
    } else {
      const long long o = (long long)i * g.sh + j;
#pragma unroll
      for (int c = 0; c < 4; ++c) z[c] = scale * (ld(img + c * g.sc + o) + shift);
    }
    const float a = z[3], om = 1.f - a;
    S0 = fmaf(om, S0, a * z[0]);
    S1 = fmaf(om, S1, a * z[1]);
    S2 = fmaf(om, S2, a * z[2]);
    R = fmaf(om, R, a);
  }
  float o0, o1, o2;
  if (g.L == 1) {            // the reference returns the single layer untouched (image_utils.py:142)
    o0 = z[0]; o1 = z[1]; o2 = z[2];
  } else {
    const float inv = (R != 0.f) ? 1.f / R : 0.f;     // nan_to_num(0/0) = 0 (image_utils.py:132)
    o0 = S0 * inv; o1 = S1 * inv; o2 = S2 * inv;
  }
  const float os = g.m11 ? 2.f : 1.f, ob = g.m11 ? -1.f : 0.f;
  const long long hw = (long long)g.H * g.W;
  T* ob_ptr = out + (long long)b * 4 * hw + (long long)i * g.W + j;
  st(ob_ptr, fmaf(o0, os, ob));
  st(ob_ptr + hw, fmaf(o1, os, ob));
  st(ob_ptr + 2 * hw, fmaf(o2, os, ob));
  st(ob_ptr + 3 * hw, fmaf(R, os, ob));
}

// ---------------------------------------------------------------------------------------------
// backward
//   sweep 1 (front -> back): alpha taps only, stash the transmittance T_l in front of each layer
//   sweep 2 (back -> front): all taps, running S_l / R_l behind the layer;
//       d c_l = G_P T_l a_l ;  d a_l = T_l [ G_P.(c_l - S_l) + G_A (1 - R_l) ]      (A.3)
//   then the bilinear adjoint: scatter to grad_x (fp32 atomics) and reduce grad_theta (A.1).
// ---------------------------------------------------------------------------------------------
template <typename T, int LMAX, bool kWarp, bool kNeedX, bool kNeedTheta>
__global__ void __launch_bounds__(kDirectThreads)
render_bwd_direct(const T* __restrict__ x, const float* __restrict__ theta, const T* __restrict__ out,
                  const T* __restrict__ gout, float* __restrict__ gx32, T* __restrict__ gx_direct,
                  float* __restrict__ gtheta, Geometry g) {
  extern __shared__ unsigned char smem_raw[];
  TileAffine* aff = reinterpret_cast<TileAffine*>(smem_raw);                      // [L]
  float* gth_acc = reinterpret_cast<float*>(smem_raw + sizeof(TileAffine) * g.L);  // [L][6]
  const int b = blockIdx.z;
  const int j0 = blockIdx.x * kTileW, i0 = blockIdx.y * kTileH;
  const int dj = threadIdx.x % kTileW, di = threadIdx.x / kTileW;
  const int j = j0 + dj, i = i0 + di;
  const bool live = (j < g.W) && (i < g.H);
  if (kWarp) {
    for (int l = threadIdx.x; l < g.L; l += kDirectThreads)
      aff[l] = make_tile_affine(theta + ((long long)b * g.L + l) * 6, g.H, g.W, j0, i0);
    if (kNeedTheta)
      for (int k = threadIdx.x; k < g.L * 6; k += kDirectThreads) gth_acc[k] = 0.f;
    __syncthreads();
  }
  const float shift = g.m11 ? 1.f : 0.f, scale = g.m11 ? 0.5f : 1.f;
  const T* xb = x + (long long)b * g.sb;
  const long long hw = (long long)g.H * g.W;

  float Tl[LMAX];
  float GP0 = 0.f, GP1 = 0.f, GP2 = 0.f, GA = 0.f;
  if (live) {
    // ---- sweep 1: transmittance in front of every layer, and the composited alpha
    //      A = sum_l T_l a_l  (front-to-back form keeps tiny coverage that 1 - prod(1-a) would lose)
    float Tcur = 1.f, Aacc = 0.f;
#pragma unroll
    for (int l = LMAX - 1; l >= 0; --l) {
      if (l < g.L) {
        const T* img = xb + (long long)l * g.sl;
        float a;
        if (kWarp) {
          const Taps p = make_taps(aff[l], dj, di, g.H, g.W, g.sh);
          a = sample_alpha(img, g.sc, p, shift, scale);
        } else {
          a = scale * (ld(img + 3 * g.sc + (long long)i * g.sh + j) + shift);
        }
        Tl[l] = Tcur;
        Aacc = fmaf(Tcur, a, Aacc);
        Tcur *= (1.f - a);
      }
    }
    // ---- upstream gradient in the compositing domain
    const long long po = (long long)b * 4 * hw + (long long)i * g.W + j;
    const float gs = g.m11 ? 2.f : 1.f;                 // d out / d o
    const float g0 = gs * ld(gout + po), g1 = gs * ld(gout + po + hw), g2 = gs * ld(gout + po + 2 * hw),
                g3 = gs * ld(gout + po + 3 * hw);
    if (g.L == 1) {
      GP0 = g0; GP1 = g1; GP2 = g2; GA = g3;            // identity (handled below via Tl = 1, a-terms)
    } else {
      const float A = Aacc;
      if (A != 0.f) {
        const float inv = 1.f / A;
        // o_rgb from the saved forward output
        const float is = g.m11 ? 0.5f : 1.f, ib = g.m11 ? 0.5f : 0.f;
        const float o0 = fmaf(ld(out + po), is, ib), o1 = fmaf(ld(out + po + hw), is, ib),
                    o2 = fmaf(ld(out + po + 2 * hw), is, ib);
        GP0 = g0 * inv; GP1 = g1 * inv; GP2 = g2 * inv;
        GA = g3 - (g0 * o0 + g1 * o1 + g2 * o2) * inv;
      }
    }
  }

  // ---- sweep 2
  float S0 = 0.f, S1 = 0.f, S2 = 0.f, R = 0.f;
  const float xj = norm_coord(j, g.W), yi = norm_coord(i, g.H);
#pragma unroll
  for (int l = 0; l < LMAX; ++l) {
    if (l < g.L) {                                    // uniform across the block
      float part[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (live) {
        const T* img = xb + (long long)l * g.sl;
        const float T_l = Tl[l];
        float z[4];
        Taps p;
        float v[4][4];
        if (kWarp) {
          p = make_taps(aff[l], dj, di, g.H, g.W, g.sh);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const T* pl = img + c * g.sc;
            v[c][0] = (p.mask & 1u) ? ld(pl + p.o00) + shift : 0.f;
            v[c][1] = (p.mask & 2u) ? ld(pl + p.o01) + shift : 0.f;
            v[c][2] = (p.mask & 4u) ? ld(pl + p.o10) + shift : 0.f;
            v[c][3] = (p.mask & 8u) ? ld(pl + p.o11) + shift : 0.f;
            z[c] = scale * fmaf(v[c][3], p.w11, fmaf(v[c][2], p.w10, fmaf(v[c][1], p.w01, v[c][0] * p.w00)));
          }
        } else {
          const long long o = (long long)i * g.sh + j;
#pragma unroll
          for (int c = 0; c < 4; ++c) z[c] = scale * (ld(img + c * g.sc + o) + shift);
        }
        const float a = z[3], om = 1.f - a;
        float gz[4];
        if (g.L == 1) {
          gz[0] = GP0; gz[1] = GP1; gz[2] = GP2; gz[3] = GA;
        } else {
          const float ta = T_l * a;
          gz[0] = GP0 * ta; gz[1] = GP1 * ta; gz[2] = GP2 * ta;
          gz[3] = T_l * (GP0 * (z[0] - S0) + GP1 * (z[1] - S1) + GP2 * (z[2] - S2) + GA * (1.f - R));
        }
        S0 = fmaf(om, S0, a * z[0]);
        S1 = fmaf(om, S1, a * z[1]);
        S2 = fmaf(om, S2, a * z[2]);
        R = fmaf(om, R, a);
        // d z / d (raw sample sum) = scale
        if (kWarp) {
          float dix = 0.f, diy = 0.f;
          const float ex = 1.f - p.fx, ey = 1.f - p.fy;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float gsum = gz[c] * scale;
            if (kNeedX) {
              float* pl = gx32 + (((long long)b * g.L + l) * 4 + c) * hw;
              float* q = pl + (long long)p.y0 * g.W + p.x0;   // grad_x is contiguous
              if (p.mask & 1u) atomicAdd(q, gsum * p.w00);
              if (p.mask & 2u) atomicAdd(q + 1, gsum * p.w01);
              if (p.mask & 4u) atomicAdd(q + g.W, gsum * p.w10);
              if (p.mask & 8u) atomicAdd(q + g.W + 1, gsum * p.w11);
            }
            if (kNeedTheta) {
              dix = fmaf(gsum, (v[c][1] - v[c][0]) * ey + (v[c][3] - v[c][2]) * p.fy, dix);
              diy = fmaf(gsum, (v[c][2] - v[c][0]) * ex + (v[c][3] - v[c][1]) * p.fx, diy);
            }
          }
          if (kNeedTheta) {
            const float ggx = dix * (0.5f * g.W), ggy = diy * (0.5f * g.H);
            part[0] = ggx * xj; part[1] = ggx * yi; part[2] = ggx;
            part[3] = ggy * xj; part[4] = ggy * yi; part[5] = ggy;
          }
        } else if (kNeedX) {
          const long long o = (((long long)b * g.L + l) * 4) * hw + (long long)i * g.W + j;
#pragma unroll
          for (int c = 0; c < 4; ++c) st(gx_direct + o + c * hw, gz[c] * scale);
        }
      }
      if (kWarp && kNeedTheta) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const float s = warp_sum(part[k]);
          if ((threadIdx.x & 31) == 0) atomicAdd(&gth_acc[l * 6 + k], s);
        }
      }
    }
  }
  if (kWarp && kNeedTheta) {
    __syncthreads();
    for (int k = threadIdx.x; k < g.L * 6; k += kDirectThreads)
      atomicAdd(gtheta + (long long)b * g.L * 6 + k, gth_acc[k]);
  }
}

// fp32 scratch -> storage dtype (only needed when grad_x is 16-bit and was accumulated in fp32)
template <typename T>
__global__ void cast_from_f32(const float* __restrict__ src, T* __restrict__ dst, long long n) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; k < n; k += stride) st(dst + k, src[k]);
}

}  // namespace mgr
