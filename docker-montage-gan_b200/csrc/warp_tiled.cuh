// Materialised warp, tiled: the staging / sampling machinery of the fused renderer (tile_common.cuh) with one layer
// per CTA and the sampled RGBA written out instead of composited.  What STNv2c / STNv2b return to their callers
// (fukuwarai/networks.py:250-257) and what random_position computes (custom_utils/image_utils.py:281-294): snapshot,
// EMA and metrics paths want the warped layers themselves.  grid_sample(x + 1) - 1 with zeros padding is the plain
// bilinear lerp of the raw values with out-of-image texels read as -1 (m11) / 0 (01), which is exactly what the staged
// footprint holds, so the output is the raw lerp.
#pragma once
#include "render_bwd_tiled.cuh"

namespace mgr {

template <typename T>
__global__ void __launch_bounds__(kTiledThreads, 3)
warp_fwd_tiled(const T* __restrict__ x, const float* __restrict__ theta, T* __restrict__ out, Geometry g, int skip_shift) {
  using Vec = typename Texel<T>::Vec;
  if (skip_shift && is_pure_shift(theta + (long long)blockIdx.z * 6)) return;      // warp_fwd_shift_tma's layer
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Vec* buf = reinterpret_cast<Vec*>(smem_raw);                // [kCapTexels]
  __shared__ LayerPlan plan;
  const int tid = threadIdx.x;
  const int n = blockIdx.z;                                   // b * L + l
  const int b = n / g.L, l = n - b * g.L;
  const int j0 = blockIdx.x * kTW, i0 = blockIdx.y * kTH;
  const int tx = tid & 31, ty = tid >> 5;
  if (tid == 0) plan = plan_layer(theta + (long long)n * 6, g.H, g.W, j0, i0, kStageVec, SrcRect{0, 0, g.W, g.H});
  __syncthreads();
  const LayerPlan& p = plan;
  const int hw = g.H * g.W;
  const int j = j0 + tx;
  const int row8 = kRowStep * g.W;
  T* op = out + (long long)n * 4 * hw + (i0 + ty) * g.W + j;
  const float padv = g.m11 ? -1.f : 0.f;
  const SrcLayers none{};
  const SrcView sv_ = layer_view<T, false>(x, g, none, b, l);
  if (p.mode == kStaged) stage_footprint<T>(g.m11 != 0, sv_, p, buf, tid);
  __syncthreads();
  const float djf = (float)(tx - kTW / 2);
  const float bx = fmaf(p.aff.a00, djf, p.lrx), by = fmaf(p.aff.a10, djf, p.lry);
#pragma unroll
  for (int k = 0; k < kPx; ++k) {
    if (j >= g.W || i0 + ty + kRowStep * k >= g.H) continue;
    float r_ = padv, g_ = padv, b_ = padv, a_ = padv;         // kSkip: the taps miss the image
    if (p.mode == kStaged) {
      const float dif = (float)(ty + kRowStep * k - kTH / 2);
      const float ix = fmaf(p.aff.a01, dif, bx), iy = fmaf(p.aff.a11, dif, by);
      const float fxf = floorf(ix), fyf = floorf(iy);
      const Sample s = sample_staged<T>(buf + (int)fyf * p.bw + (int)fxf, p.bw, ix - fxf, iy - fyf);
      upk(s.rg, r_, g_);
      upk(s.ba, b_, a_);
    } else if (p.mode == kDirect) {
      const float shift = g.m11 ? 1.f : 0.f;
      const float4 z = sample_pixel_direct<T>(reinterpret_cast<const T*>(sv_.base), p.aff, tx - kTW / 2, ty + kRowStep * k - kTH / 2,
                                              sv_.h, sv_.w, sv_.rowbytes / sizeof(T), sv_.plane / sizeof(T), shift, 1.f);
      r_ = z.x - shift; g_ = z.y - shift; b_ = z.z - shift; a_ = z.w - shift;
    }
    T* q = op + k * row8;
    st(q, r_); st(q + hw, g_); st(q + 2 * hw, b_); st(q + 3 * hw, a_);
  }
}

// ---- backward of the materialised warp, atomics-free -----------------------------------------------------------
// grad_x: the gather-form bilinear adjoint of the fused renderer's pass 2 (render_bwd_tiled.cuh: pass2_block), reading
// the upstream gradient of the warped layers (four planes of T per layer) instead of composite records.  Needs the
// inverse placements and the work list (inverse_plans_kernel, sample_flags_kernel with every layer kept).
template <typename T, int kTX>
__global__ void __launch_bounds__(256, MGR_P2_BLOCKS)
warp_bwd_gather(const InverseLayer* __restrict__ plans, const int* __restrict__ work, const int* __restrict__ wcnt,
                const T* __restrict__ gw, T* __restrict__ gx, Geometry g, int skip_shift) {
  if ((int)blockIdx.z >= wcnt[0]) return;
  if (skip_shift && plans[work[blockIdx.z]].shift_only) return;      // a translation layer: warp_fwd_shift_tma<T, true> writes its grad_x
  pass2_block<T, false, kTX>(plans, work[blockIdx.z], blockIdx.x * P2Shape<kTX>::kW, blockIdx.y * P2Shape<kTX>::kH,
                             PlanarGrads<T>{gw, 0}, 1.f, gx, nullptr, g);
}

// grad_theta: one (layer, 32 x 32 tile) per CTA -- stage the footprint, take d sample / d (ix, iy) from the lerp
// differences, contract with the upstream gradient, reduce over the tile, six atomics per CTA.
template <typename T>
__global__ void __launch_bounds__(kTiledThreads, 3)
warp_bwd_theta_tiled(const T* __restrict__ x, const float* __restrict__ theta, const T* __restrict__ gw,
                     float* __restrict__ gtheta, Geometry g, int skip_shift) {
  using Vec = typename Texel<T>::Vec;
  if (skip_shift && is_pure_shift(theta + (long long)blockIdx.z * 6)) return;      // warp_bwd_theta_shift_tma's layer
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Vec* buf = reinterpret_cast<Vec*>(smem_raw);                                   // [kCapTexels]
  float* stash = reinterpret_cast<float*>(smem_raw + sizeof(Vec) * kCapTexels);  // [kPx][256] partial sums
  __shared__ LayerPlan plan;
  const int tid = threadIdx.x;
  const int n = blockIdx.z;
  const int b = n / g.L, l = n - b * g.L;
  const int j0 = blockIdx.x * kTW, i0 = blockIdx.y * kTH;
  const int tx = tid & 31, ty = tid >> 5;
  if (tid == 0) plan = plan_layer(theta + (long long)n * 6, g.H, g.W, j0, i0, kStageVec, SrcRect{0, 0, g.W, g.H});
  __syncthreads();
  const LayerPlan& p = plan;
  if (p.mode == kSkip) return;                                  // the taps miss the image: no dependence on theta
  const int hw = g.H * g.W;
  const int j = j0 + tx;
  const SrcLayers none{};
  const SrcView sv_ = layer_view<T, false>(x, g, none, b, l);
  if (p.mode == kStaged) stage_footprint<T>(g.m11 != 0, sv_, p, buf, tid);
  __syncthreads();
  const T* gp_ = gw + (long long)n * 4 * hw + (i0 + ty) * g.W + j;
  const float djf = (float)(tx - kTW / 2);
  const float bx = fmaf(p.aff.a00, djf, p.lrx), by = fmaf(p.aff.a10, djf, p.lry);
  float accx = 0.f, accxy = 0.f, accy = 0.f, accyy = 0.f;
#pragma unroll
  for (int k = 0; k < kPx; ++k) {
    const int i = i0 + ty + kRowStep * k;
    if (j >= g.W || i >= g.H) continue;
    float dxr, dxg, dxb, dxa, dyr, dyg, dyb, dya;
    if (p.mode == kStaged) {
      const float dif = (float)(ty + kRowStep * k - kTH / 2);
      const float ix = fmaf(p.aff.a01, dif, bx), iy = fmaf(p.aff.a11, dif, by);
      const float fxf = floorf(ix), fyf = floorf(iy);
      const SampleGrad s = sample_staged_grad<T>(buf + (int)fyf * p.bw + (int)fxf, p.bw, ix - fxf, iy - fyf);
      upk(s.dx_rg, dxr, dxg); upk(s.dx_ba, dxb, dxa);
      upk(s.dy_rg, dyr, dyg); upk(s.dy_ba, dyb, dya);
    } else {                                                    // huge footprint: bounds-checked taps from global memory
      const Taps tp = make_taps(p.aff, tx - kTW / 2, ty + kRowStep * k - kTH / 2, sv_.h, sv_.w, sv_.rowbytes / sizeof(T));
      const float shift = g.m11 ? 1.f : 0.f;
      const float ex = 1.f - tp.fx, ey = 1.f - tp.fy;
      float dd[4][2];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const T* pl = reinterpret_cast<const T*>(sv_.base + (size_t)c * sv_.plane);
        const float v00 = (tp.mask & 1u) ? ld(pl + tp.o00) + shift : 0.f, v01 = (tp.mask & 2u) ? ld(pl + tp.o01) + shift : 0.f;
        const float v10 = (tp.mask & 4u) ? ld(pl + tp.o10) + shift : 0.f, v11 = (tp.mask & 8u) ? ld(pl + tp.o11) + shift : 0.f;
        dd[c][0] = (v01 - v00) * ey + (v11 - v10) * tp.fy;
        dd[c][1] = (v10 - v00) * ex + (v11 - v01) * tp.fx;
      }
      dxr = dd[0][0]; dxg = dd[1][0]; dxb = dd[2][0]; dxa = dd[3][0];
      dyr = dd[0][1]; dyg = dd[1][1]; dyb = dd[2][1]; dya = dd[3][1];
    }
    const T* q = gp_ + k * kRowStep * g.W;
    const float g0 = ld(q), g1 = ld(q + hw), g2 = ld(q + 2 * hw), g3 = ld(q + 3 * hw);
    const float dix = fmaf(g0, dxr, fmaf(g1, dxg, fmaf(g2, dxb, g3 * dxa)));
    const float diy = fmaf(g0, dyr, fmaf(g1, dyg, fmaf(g2, dyb, g3 * dya)));
    const float yi = norm_coord(i, g.H);
    accx += dix; accxy = fmaf(dix, yi, accxy);
    accy += diy; accyy = fmaf(diy, yi, accyy);
  }
  park_theta_partials(stash + tid, accx, accxy, accy, accyy);
  __syncthreads();
  reduce_theta_partials(stash, 1, tid, norm_coord(j, g.W), 0.5f * (float)g.W, 0.5f * (float)g.H, gtheta + (long long)n * 6);
}

}  // namespace mgr
