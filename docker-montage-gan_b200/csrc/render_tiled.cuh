// Tiled forward: one CTA per 32x32 output tile of one sample, looping over the layers back to
// front.  Per layer the CTA stages the source footprint of the tile into shared memory
// (tile_common.cuh), every thread samples its four pixels from shared memory with packed
// fp32x2 arithmetic, and the running premultiplied colour / alpha stay in registers.  One LDS.64
// (16-bit storage) or LDS.128 (fp32 storage) per bilinear tap, no bounds checks in the sampling
// loop, no intermediate tensor (grid, warped layers, range-shifted copies) ever reaches memory.
//
// A layer whose footprint misses the image is skipped; one whose footprint does not fit the
// staging buffer (strong minification / rotation) is sampled straight from global memory with the
// bounds-checked taps of render_direct.cuh -- decided per (tile, layer), uniform across the CTA.
//
// Optionally writes the sampled alpha of every (layer, pixel) for the backward pass (`sav`), which
// spares the backward a whole front-to-back sampling sweep.
//
// Math: SURVEY.md Appendix A.  Reference semantics: fukuwarai/networks.py:250-257 (warp),
// custom_utils/image_utils.py:128-146 (over), custom/loss_aio.py:251 (range shifts).
#pragma once
#include "render_direct.cuh"
#include "tile_common.cuh"

namespace mgr {

// storage type of the saved alpha samples: fp32 for fp32 tensors, fp16 otherwise (alpha lives in
// [0,1]: fp16 keeps 11 significant bits, enough for 16-bit gradients)
template <typename T> struct SavedAlpha { using type = __half; };
template <> struct SavedAlpha<float> { using type = float; };
__device__ __forceinline__ void st_alpha(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_alpha(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ float ld_alpha(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_alpha(const __half* p) { return __half2float(__ldg(p)); }

// Bounds-checked sample of one pixel straight from global memory (fallback for huge footprints).
template <typename T>
__device__ __noinline__ float4 sample_pixel_direct(const T* __restrict__ img, const TileAffine& t, int dj, int di,
                                                   int H, int W, long long sh, long long sc, float shift, float scale) {
  const Taps tp = make_taps(t, dj, di, H, W, sh);
  float zz[4];
  sample_rgba(img, sc, tp, shift, scale, zz);
  return make_float4(zz[0], zz[1], zz[2], zz[3]);
}

// L >= 2 only (a single layer is returned untouched by the reference; the host routes L == 1 to
// the direct kernel).  kSave: also write the sampled alpha of every (layer, pixel) to `sav`.
#ifndef MGR_FWD_BLOCKS
#define MGR_FWD_BLOCKS 3
#endif
template <typename T, bool kSave, bool kRagged>
__device__ __forceinline__ void fwd_tiled_body(const T* __restrict__ x, const SrcLayers& src, const float* __restrict__ theta,
                                               T* __restrict__ out, typename SavedAlpha<T>::type* __restrict__ sav,
                                               const Geometry& g) {
  using Vec = typename Texel<T>::Vec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Vec* buf = reinterpret_cast<Vec*>(smem_raw);                                            // [kCapTexels]
  LayerPlan* plan = reinterpret_cast<LayerPlan*>(smem_raw + sizeof(Vec) * kCapTexels);    // [L]
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int j0 = blockIdx.x * kTW, i0 = blockIdx.y * kTH;
  const int tx = tid & 31, ty = tid >> 5;
  for (int l = tid; l < g.L; l += kTiledThreads)
    plan[l] = plan_layer(theta + ((long long)b * g.L + l) * 6, g.H, g.W, j0, i0, kStageVec, layer_rect<kRagged>(g, src, l));
  __syncthreads();

  const f32x2 zs2 = bc(g.m11 ? 0.5f : 1.f), zb2 = bc(g.m11 ? 0.5f : 0.f);     // z = zs * raw + zb
  const int hw = g.H * g.W;                                   // one plane fits 32 bits (host-checked)
  const int j = j0 + tx;
  const int pix0 = (i0 + ty) * g.W + j;                       // pixel k lives 8*k rows further down
  const int row8 = kRowStep * g.W;                          // a thread's pixels are kRowStep rows apart
  bool live[kPx];
#pragma unroll
  for (int k = 0; k < kPx; ++k) live[k] = j < g.W && i0 + ty + kRowStep * k < g.H;
  const float djf = (float)(tx - kTW / 2);
  float S0[kPx], S1[kPx], S2[kPx], R[kPx];
#pragma unroll
  for (int k = 0; k < kPx; ++k) S0[k] = S1[k] = S2[k] = R[k] = 0.f;

  for (int l = 0; l < g.L; ++l) {
    const LayerPlan& p = plan[l];
    const int mode = p.mode;
    typename SavedAlpha<T>::type* sv = nullptr;
    if (kSave) sv = sav + ((long long)b * g.L + l) * hw + pix0;
    if (mode == kSkip) {                     // fully transparent layer: the canvas is unchanged
      if (kSave) {
#pragma unroll
        for (int k = 0; k < kPx; ++k)
          if (live[k]) st_alpha(sv + k * row8, 0.f);
      }
      continue;
    }
    const SrcView sv_ = layer_view<T, kRagged>(x, g, src, b, l);
    if (mode == kStaged) {
      // fp32 footprints take the "readers are done" barrier with their loads already in flight (measured: -5 % forward,
      // -2 % pass 1 at 512 x 512 fp32; neutral or slightly negative for 16-bit texels, which keep the plain order)
      if constexpr (sizeof(T) == 4) {
        stage_footprint<T, true>(g.m11 != 0, sv_, p, buf, tid);
      } else {
        __syncthreads();                       // the previous layer's readers are done with buf
        stage_footprint<T>(g.m11 != 0, sv_, p, buf, tid);
      }
      __syncthreads();
    }
    const float a01 = p.aff.a01, a11 = p.aff.a11;
    const float bx = fmaf(p.aff.a00, djf, p.lrx), by = fmaf(p.aff.a10, djf, p.lry);
    const int pitch = p.bw;
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
      float r_, g_, b_, a;
      if (mode == kStaged) {
        const float dif = (float)(ty + kRowStep * k - kTH / 2);
        const float ix = fmaf(a01, dif, bx), iy = fmaf(a11, dif, by);
        const float fxf = floorf(ix), fyf = floorf(iy);
        const Sample s = sample_staged<T>(buf + (int)fyf * pitch + (int)fxf, pitch, ix - fxf, iy - fyf);
        upk(fma2(s.rg, zs2, zb2), r_, g_);
        upk(fma2(s.ba, zs2, zb2), b_, a);
      } else {
        const float4 z = sample_pixel_direct<T>(reinterpret_cast<const T*>(sv_.base), p.aff, tx - kTW / 2,
                                                ty + kRowStep * k - kTH / 2, sv_.h, sv_.w, sv_.rowbytes / sizeof(T), sv_.plane / sizeof(T),
                                                g.m11 ? 1.f : 0.f, g.m11 ? 0.5f : 1.f);
        r_ = z.x; g_ = z.y; b_ = z.z; a = z.w;
      }
      if (kSave) { if (live[k]) st_alpha(sv + k * row8, a); }
      const float om = 1.f - a;
      S0[k] = fmaf(om, S0[k], a * r_);
      S1[k] = fmaf(om, S1[k], a * g_);
      S2[k] = fmaf(om, S2[k], a * b_);
      R[k] = fmaf(om, R[k], a);
    }
  }

  const float os = g.m11 ? 2.f : 1.f, obias = g.m11 ? -1.f : 0.f;   // out = os * o + obias
  T* outp = out + (long long)b * 4 * hw + pix0;
#pragma unroll
  for (int k = 0; k < kPx; ++k) {
    if (live[k]) {
      const float inv = (R[k] != 0.f) ? 1.f / R[k] : 0.f;     // nan_to_num(0/0) = 0 (image_utils.py:132)
      T* q = outp + k * row8;
      st(q, fmaf(S0[k] * inv, os, obias));
      st(q + hw, fmaf(S1[k] * inv, os, obias));
      st(q + 2 * hw, fmaf(S2[k] * inv, os, obias));
      st(q + 3 * hw, fmaf(R[k], os, obias));
    }
  }
}

inline size_t tiled_smem_bytes(int L, size_t vec_bytes) { return vec_bytes * kCapTexels + sizeof(LayerPlan) * L; }

}  // namespace mgr
