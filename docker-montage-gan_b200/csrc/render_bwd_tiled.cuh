// Tiled backward in two passes, no atomics on the pixel-gradient path.
//
// Pass 1  (render_bwd_pass1, one CTA per 32x32 OUTPUT tile, same staging as the forward):
//   * pre-pass over the alpha samples saved by the forward: transmittance T_l in front of every
//     layer (front -> back) and the composited alpha A = sum_l T_l a_l;
//   * back -> front sweep with the layers re-sampled from staged footprints, carrying the canvas
//     S_l, R_l behind the layer:   d c_l = G_P T_l a_l,   d a_l = T_l [G_P.(c_l - S_l) + G_A (1 - R_l)]
//     (division-free in (1 - a_l), so exact at opaque texels);
//   * writes one gradient record per (layer, pixel): (T_l a_l, d a_l), plus G_P per pixel -- the
//     gradient w.r.t. the warped sample is (G_P * T_l a_l, d a_l);
//   * grad_theta: d sample / d(ix, iy) comes for free from the lerp differences; reduced
//     thread -> warp shuffle -> shared -> one atomicAdd per (CTA, layer, coefficient)  (A.1).
//
// Pass 2  (render_bwd_pass2, one thread per SOURCE texel): the bilinear adjoint in gather form.
//   grad_x[t] = sum over output pixels p of  hat(ix(p) - x_t) * hat(iy(p) - y_t) * g(p),
//   hat(u) = max(0, 1 - |u|).  The pixels that can touch texel t are the integer points of the
//   parallelogram M^-1((x_t, y_t) + (-1,1)^2); they are enumerated over its bounding box, so every
//   grad_x element is written exactly once, coalesced, in the storage dtype -- no zero-fill, no
//   fp32 scatter buffer, no atomics, deterministic.
//
// Reference semantics: autograd of fukuwarai/networks.py:250-257 + custom/loss_aio.py:251
// (ATen grid_sampler_2d_backward + affine_grid backward + the a_over_b chain).
#pragma once
#include "render_tiled.cuh"

namespace mgr {

// workspace layout of the tiled backward (all fp32):
//   rec [B*L][H*W] float2 = (T_l a_l, d a_l)        gp [B][H*W] float4 = (G_P.rgb, G_A)
//   inverse plans [B*L] InverseLayer (128 B each)        order [B*L] int + 2 counters
//   work [B*L] int + 2 counters (pass 2's compact layer list)   sample_all_shift [B] int

#ifndef MGR_P1_BLOCKS
#define MGR_P1_BLOCKS 2
#endif
template <typename T, bool kNeedTheta, bool kGPSmem, bool kRagged>
__global__ void __launch_bounds__(kTiledThreads, kGPSmem ? MGR_P1_BLOCKS : 2)
render_bwd_pass1(const T* __restrict__ x, const __grid_constant__ SrcLayers src, const float* __restrict__ theta, const T* __restrict__ out,
                 const T* __restrict__ gout, const typename SavedAlpha<T>::type* __restrict__ sav,
                 float2* __restrict__ rec, float4* __restrict__ gp, float* __restrict__ gtheta, Geometry g,
                 const int* __restrict__ sample_all_shift, int skip_shift) {
  using Vec = typename Texel<T>::Vec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Vec* buf = reinterpret_cast<Vec*>(smem_raw);                                            // [kCapTexels]
  LayerPlan* plan = reinterpret_cast<LayerPlan*>(smem_raw + sizeof(Vec) * kCapTexels);    // [L]
  const int tid = threadIdx.x;
  // [L][kPx][256]: transmittance in front of layer l, later the layer's theta-gradient partials (16-byte aligned)
  float* stash = reinterpret_cast<float*>(smem_raw + align16(sizeof(Vec) * kCapTexels + sizeof(LayerPlan) * g.L));
  float* Tst = stash + tid;
  // (G_P, G_A) per pixel: a shared-memory copy when it still leaves room for 2 CTAs/SM (host decides), else the
  // thread re-reads its own entries of the global gp buffer (L1/L2 hits)
  float4* GPs = reinterpret_cast<float4*>(stash + (size_t)g.L * kPx * kTiledThreads) + tid;   // [kPx][256]
  const int b = blockIdx.z;
  if (skip_shift && sample_all_shift[b]) return;            // render_bwd_shift's sample (flag from sample_flags_kernel)
  const int j0 = blockIdx.x * kTW, i0 = blockIdx.y * kTH;
  const int tx = tid & 31, ty = tid >> 5;
  for (int l = tid; l < g.L; l += kTiledThreads)
    plan[l] = plan_layer(theta + ((long long)b * g.L + l) * 6, g.H, g.W, j0, i0, kStageVec, layer_rect<kRagged>(g, src, l));
  __syncthreads();

  const float zs = g.m11 ? 0.5f : 1.f;
  const f32x2 zs2 = bc(zs), zb2 = bc(g.m11 ? 0.5f : 0.f);     // z = zs * raw + zb
  const int hw = g.H * g.W;
  const int j = j0 + tx;
  const int pix0 = (i0 + ty) * g.W + j;                       // pixel k lives 8*k rows further down
  const int row8 = kRowStep * g.W;                          // a thread's pixels are kRowStep rows apart
  bool live[kPx];
#pragma unroll
  for (int k = 0; k < kPx; ++k) live[k] = j < g.W && i0 + ty + kRowStep * k < g.H;
  float2* recb = rec + (long long)b * g.L * hw + pix0;
  const typename SavedAlpha<T>::type* savb = sav + (long long)b * g.L * hw + pix0;

  float4* gpp = gp + (long long)b * hw + pix0;                 // (G_P, G_A) of this thread's pixels: gpp[k * row8]
  // ---- pre-pass: T_l (stashed in shared memory) and A ------------------------------------------------
  {
    float GP0[kPx], GP1[kPx], GP2[kPx], GA[kPx];
    float Tc[kPx], A[kPx];
    // upstream gradient and forward output: issued first so that their latency hides behind the alpha sweep
    float gv[kPx][4], ov[kPx][3];
    {
      const T* gob = gout + (long long)b * 4 * hw + pix0;
      const T* ob_ = out + (long long)b * 4 * hw + pix0;
#pragma unroll
      for (int k = 0; k < kPx; ++k) {
#pragma unroll
        for (int c = 0; c < 4; ++c) gv[k][c] = live[k] ? ld(gob + k * row8 + c * hw) : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) ov[k][c] = live[k] ? ld(ob_ + k * row8 + c * hw) : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < kPx; ++k) { Tc[k] = 1.f; A[k] = 0.f; }
    // running pointers (front layer first): keeps the loop free of 64-bit index arithmetic
    const typename SavedAlpha<T>::type* sa[kPx];
#pragma unroll
    for (int k = 0; k < kPx; ++k) sa[k] = savb + (long long)(g.L - 1) * hw + k * row8;
    float* tp = Tst + (g.L - 1) * kPx * kTiledThreads;
    for (int l = g.L - 1; l >= 0; --l) {
#pragma unroll
      for (int k = 0; k < kPx; ++k) {
        tp[k * kTiledThreads] = live[k] ? Tc[k] : 0.f;
        if (live[k]) {
          const float a = ld_alpha(sa[k]);
          A[k] = fmaf(Tc[k], a, A[k]);
          Tc[k] *= (1.f - a);
        }
        sa[k] -= hw;
      }
      tp -= kPx * kTiledThreads;
    }
    // upstream gradient in the compositing domain: o = P / A
    const float gs = g.m11 ? 2.f : 1.f;                       // d out / d o
    const float is = g.m11 ? 0.5f : 1.f, ib = g.m11 ? 0.5f : 0.f;
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
      GP0[k] = GP1[k] = GP2[k] = GA[k] = 0.f;
      if (live[k]) {
        const float g0 = gs * gv[k][0], g1 = gs * gv[k][1],
                    g2 = gs * gv[k][2], g3 = gs * gv[k][3];
        if (A[k] != 0.f) {                                    // A == 0: every gradient is defined as 0
          const float inv = 1.f / A[k];
          const float o0 = fmaf(ov[k][0], is, ib), o1 = fmaf(ov[k][1], is, ib),
                      o2 = fmaf(ov[k][2], is, ib);
          GP0[k] = g0 * inv; GP1[k] = g1 * inv; GP2[k] = g2 * inv;
          GA[k] = g3 - (g0 * o0 + g1 * o1 + g2 * o2) * inv;
        }
        gpp[k * row8] = make_float4(GP0[k], GP1[k], GP2[k], GA[k]);      // pass 2 reads it (and this thread, see above)
      }
      if (kGPSmem) GPs[k * kTiledThreads] = make_float4(GP0[k], GP1[k], GP2[k], GA[k]);
    }
  }

  // ---- back -> front sweep ------------------------------------------------------------------------
  const float djf = (float)(tx - kTW / 2);
  const float xj = norm_coord(j, g.W);
  const float hW = 0.5f * (float)g.W * zs, hH = 0.5f * (float)g.H * zs;   // d ix / d gx (and the range scale)
  float S0[kPx], S1[kPx], S2[kPx], R[kPx];
#pragma unroll
  for (int k = 0; k < kPx; ++k) S0[k] = S1[k] = S2[k] = R[k] = 0.f;

  for (int l = 0; l < g.L; ++l) {
    const LayerPlan& p = plan[l];
    const int mode = p.mode;
    float2* rl = recb + (long long)l * hw;
    if (mode == kSkip) {                     // a = 0, c undefined: no colour gradient; d a_l is still defined
#pragma unroll
      for (int k = 0; k < kPx; ++k) {
        if (live[k]) {
          const float T_l = Tst[(l * kPx + k) * kTiledThreads];
          // c_l = 0 in the compositing domain (transparent black)
          const float4 G4 = kGPSmem ? GPs[k * kTiledThreads] : gpp[k * row8];
          const float ga = T_l * (-(G4.x * S0[k] + G4.y * S1[k] + G4.z * S2[k]) + G4.w * (1.f - R[k]));
          rl[k * row8] = make_float2(0.f, ga);
        }
      }
      if (kNeedTheta) park_theta_partials(Tst + l * kPx * kTiledThreads, 0.f, 0.f, 0.f, 0.f);
      continue;                              // the footprint misses the image: no texel, no theta gradient
    }
    const SrcView sv_ = layer_view<T, kRagged>(x, g, src, b, l);
    if (mode == kStaged) {
      // fp32 footprints take the "readers are done" barrier with their loads already in flight (measured: -5 % forward,
      // -2 % pass 1 at 512 x 512 fp32; neutral or slightly negative for 16-bit texels, which keep the plain order)
      if constexpr (sizeof(T) == 4) {
        stage_footprint<T, true>(g.m11 != 0, sv_, p, buf, tid);
      } else {
        __syncthreads();
        stage_footprint<T>(g.m11 != 0, sv_, p, buf, tid);
      }
      __syncthreads();
    }
    const float a01 = p.aff.a01, a11 = p.aff.a11;
    const float bx = fmaf(p.aff.a00, djf, p.lrx), by = fmaf(p.aff.a10, djf, p.lry);
    const int pitch = p.bw;
    float accx = 0.f, accxy = 0.f, accy = 0.f, accyy = 0.f;
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
      float r_, g_, b_, a;
      float dxr, dxg, dxb, dxa, dyr, dyg, dyb, dya;
      if (mode == kStaged) {
        const float dif = (float)(ty + kRowStep * k - kTH / 2);
        const float ix = fmaf(a01, dif, bx), iy = fmaf(a11, dif, by);
        const float fxf = floorf(ix), fyf = floorf(iy);
        const SampleGrad s = sample_staged_grad<T>(buf + (int)fyf * pitch + (int)fxf, pitch, ix - fxf, iy - fyf);
        upk(fma2(s.rg, zs2, zb2), r_, g_);
        upk(fma2(s.ba, zs2, zb2), b_, a);
        upk(s.dx_rg, dxr, dxg); upk(s.dx_ba, dxb, dxa);
        upk(s.dy_rg, dyr, dyg); upk(s.dy_ba, dyb, dya);
      } else {
        // huge footprint: bounds-checked taps straight from global memory
        const Taps tp = make_taps(p.aff, tx - kTW / 2, ty + kRowStep * k - kTH / 2, sv_.h, sv_.w, sv_.rowbytes / sizeof(T));
        const float shift = g.m11 ? 1.f : 0.f;
        float v[4][4], zz[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const T* pl = reinterpret_cast<const T*>(sv_.base + (size_t)c * sv_.plane);
          v[c][0] = (tp.mask & 1u) ? ld(pl + tp.o00) + shift : 0.f;
          v[c][1] = (tp.mask & 2u) ? ld(pl + tp.o01) + shift : 0.f;
          v[c][2] = (tp.mask & 4u) ? ld(pl + tp.o10) + shift : 0.f;
          v[c][3] = (tp.mask & 8u) ? ld(pl + tp.o11) + shift : 0.f;
          zz[c] = zs * fmaf(v[c][3], tp.w11, fmaf(v[c][2], tp.w10, fmaf(v[c][1], tp.w01, v[c][0] * tp.w00)));
        }
        const float ex = 1.f - tp.fx, ey = 1.f - tp.fy;
        r_ = zz[0]; g_ = zz[1]; b_ = zz[2]; a = zz[3];
        dxr = (v[0][1] - v[0][0]) * ey + (v[0][3] - v[0][2]) * tp.fy; dyr = (v[0][2] - v[0][0]) * ex + (v[0][3] - v[0][1]) * tp.fx;
        dxg = (v[1][1] - v[1][0]) * ey + (v[1][3] - v[1][2]) * tp.fy; dyg = (v[1][2] - v[1][0]) * ex + (v[1][3] - v[1][1]) * tp.fx;
        dxb = (v[2][1] - v[2][0]) * ey + (v[2][3] - v[2][2]) * tp.fy; dyb = (v[2][2] - v[2][0]) * ex + (v[2][3] - v[2][1]) * tp.fx;
        dxa = (v[3][1] - v[3][0]) * ey + (v[3][3] - v[3][2]) * tp.fy; dya = (v[3][2] - v[3][0]) * ex + (v[3][3] - v[3][1]) * tp.fx;
      }
      const float T_l = Tst[(l * kPx + k) * kTiledThreads];
      const float ta = T_l * a;
      const float4 G4 = kGPSmem ? GPs[k * kTiledThreads] : (live[k] ? gpp[k * row8] : make_float4(0.f, 0.f, 0.f, 0.f));
      const float ga = T_l * (G4.x * (r_ - S0[k]) + G4.y * (g_ - S1[k]) + G4.z * (b_ - S2[k]) + G4.w * (1.f - R[k]));
      if (live[k]) rl[k * row8] = make_float2(ta, ga);
      if (kNeedTheta) {
        const float gr = G4.x * ta, gg = G4.y * ta, gb = G4.z * ta;
        const float dix = fmaf(gr, dxr, fmaf(gg, dxg, fmaf(gb, dxb, ga * dxa)));
        const float diy = fmaf(gr, dyr, fmaf(gg, dyg, fmaf(gb, dyb, ga * dya)));
        const float yi = norm_coord(i0 + ty + kRowStep * k, g.H);
        accx += dix; accxy = fmaf(dix, yi, accxy);
        accy += diy; accyy = fmaf(diy, yi, accyy);
      }
      const float om = 1.f - a;
      S0[k] = fmaf(om, S0[k], a * r_);
      S1[k] = fmaf(om, S1[k], a * g_);
      S2[k] = fmaf(om, S2[k], a * b_);
      R[k] = fmaf(om, R[k], a);
    }
    // the T_l slots of this layer are dead: park the thread's theta-gradient partials there (tile_common.cuh)
    if (kNeedTheta) park_theta_partials(Tst + l * kPx * kTiledThreads, accx, accxy, accy, accyy);
  }
  if (kNeedTheta) {
    __syncthreads();
    // a thread's four pixels share the column: (ggx x_j, ggx y_i, ggx, ggy x_j, ggy y_i, ggy)
    reduce_theta_partials(stash, g.L, tid, xj, hW, hH, gtheta + (long long)b * g.L * 6);
  }
}

// ---------------------------------------------------------------------------------------------
// pass 2: gather-form bilinear adjoint, one thread per source texel, block = 32 x 8 texels
// ---------------------------------------------------------------------------------------------
// Per-layer inverse placement, computed once per (b, l) by inverse_plans_kernel (double precision)
// so that the thousands of pass-2 blocks of a layer do not each redo the divisions.
struct InverseLayer {
  double i00, i01, i10, i11;    // A^-1, pixel space
  double c0, c1;                // (ix, iy) = A (j, i) + c
  float a00, a01, a10, a11;     // A
  float f00, f01, f10, f11;     // A^-1 rounded to fp32 (per-thread use)
  float rj, ri;                 // half extents of the pre-image of a texel's (-1,1)^2 support (+ slack)
  float r00, r10;               // 1/a00, 1/a10 (0 if ~0): per-row interval refinement for wide windows
  int valid;                    // 0: non-finite or singular placement -> grad_x of this layer is 0
  int wide;                     // 1: window wider than 4 somewhere -> refine every row
  int shift_only;               // 1: pure translation (a00 = a11 = 1, a01 = a10 = 0 exactly): 2x2 stencil adjoint
  int X, Y;                     // shift_only: ix = j + X + fx, iy = i + Y + fy
  float fx, fy;
  int all_shift;                // every layer of this sample is a pure translation (set by sample_flags_kernel)
};

static_assert(sizeof(InverseLayer) <= 128, "workspace reserves 128 B per layer plan");

// Also builds the launch order of pass 2: layers whose per-texel window is huge (strongly magnifying or
// near-singular placements, a few hundred candidates per texel) are scheduled FIRST so that their long-running
// blocks overlap the rest of the grid instead of forming a tail (longest-processing-time-first).
// order[0..n): heavy layers from the front, the others from the back; cnt[2] zeroed by the caller.
// the inverse placement of one layer; heavy: its per-texel window is huge (scheduled first by pass 2)
__device__ __forceinline__ InverseLayer make_inverse_plan(const float* __restrict__ th, int H, int W, bool& heavy) {
  const double w = W, h = H;
  const double a00 = th[0], a01 = th[1] * (w / h), a10 = th[3] * (h / w), a11 = th[4];
  InverseLayer q;
  q.c0 = a00 * (0.5 - 0.5 * w) + a01 * (0.5 - 0.5 * h) + th[2] * 0.5 * w + 0.5 * (w - 1.0);
  q.c1 = a10 * (0.5 - 0.5 * w) + a11 * (0.5 - 0.5 * h) + th[5] * 0.5 * h + 0.5 * (h - 1.0);
  const double inv = 1.0 / (a00 * a11 - a01 * a10);
  q.i00 = a11 * inv; q.i01 = -a01 * inv; q.i10 = -a10 * inv; q.i11 = a00 * inv;
  q.a00 = (float)a00; q.a01 = (float)a01; q.a10 = (float)a10; q.a11 = (float)a11;
  q.f00 = (float)q.i00; q.f01 = (float)q.i01; q.f10 = (float)q.i10; q.f11 = (float)q.i11;
  const double rj = fabs(q.i00) + fabs(q.i01), ri = fabs(q.i10) + fabs(q.i11);
  q.valid = isfinite(q.c0) && isfinite(q.c1) && isfinite(rj) && isfinite(ri) && rj < 1.0e6 && ri < 1.0e6;
  q.rj = (float)rj * 1.0001f + 1e-3f; q.ri = (float)ri * 1.0001f + 1e-3f;
  q.r00 = fabs(a00) > 1e-6 ? (float)(1.0 / a00) : 0.f;
  q.r10 = fabs(a10) > 1e-6 ? (float)(1.0 / a10) : 0.f;
  q.wide = (2.0 * rj > 3.5) || (2.0 * ri > 3.5);
  heavy = q.valid && fmin(rj, (double)W) * fmin(ri, (double)H) > 64.0;
  // what STNv2c emits (convert_translate_to_2x3, image_utils.py:316-335): the adjoint is a fixed 2x2 stencil
  q.shift_only = is_pure_shift(th);
  const double fX = floor(q.c0), fY = floor(q.c1);
  q.X = q.shift_only ? (int)fX : 0; q.Y = q.shift_only ? (int)fY : 0;
  q.fx = q.shift_only ? (float)(q.c0 - fX) : 0.f; q.fy = q.shift_only ? (float)(q.c1 - fY) : 0.f;
  q.all_shift = 0;
  return q;
}

static __global__ void inverse_plans_kernel(const float* __restrict__ theta, InverseLayer* __restrict__ plans, int n, int H, int W,
                                            int* __restrict__ order, int* __restrict__ cnt) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  bool heavy;
  const InverseLayer q = make_inverse_plan(theta + (long long)k * 6, H, W, heavy);
  if (heavy) order[atomicAdd(&cnt[0], 1)] = k;
  else order[n - 1 - atomicAdd(&cnt[1], 1)] = k;
  plans[k] = q;
}

// The forward's flags: one warp per sample, a lane per layer (L <= 32 on the tiled path): flags[b] = every layer of the
// sample is a pure translation.  One round of loads instead of a thread walking its sample's 6 L floats (6.2 -> ~3 us).
static __global__ void __launch_bounds__(256)
sample_shift_flags_kernel(const float* __restrict__ theta, int B, int L, int* __restrict__ flags) {
  const int b = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  bool ok = true;
  for (int l = lane; l < L; l += 32) ok = ok && is_pure_shift(theta + ((long long)b * L + l) * 6);
  const bool all = __all_sync(0xffffffffu, ok);
  if (lane == 0) flags[b] = all ? 1 : 0;
}

// One CTA.  Phase 1, a thread per sample: are all of its layers pure translations?  (those samples belong to
// render_bwd_shift).  Phase 2: the layers pass 2 has to process, in launch order (order[] minus the layers of
// all-translation samples when the stencil kernels are on), compacted into work[0 .. wcnt[0]).  A batch of
// translations leaves pass 2 with nothing to do.
__device__ __forceinline__ void sample_flags_cta(InverseLayer* __restrict__ plans, int B, int L, const int* __restrict__ order,
                                                 int* __restrict__ work, int* __restrict__ wcnt, int* __restrict__ sample_all_shift,
                                                 int skip_shift) {
  __shared__ int s_warp[8], s_base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int b = tid; b < B; b += 256) {
    int all = 1;
    for (int l = 0; l < L; ++l) all &= plans[b * L + l].shift_only;
    for (int l = 0; l < L; ++l) plans[b * L + l].all_shift = all;
    sample_all_shift[b] = all;                  // what pass 1 and the stencil backward read to claim / decline a sample
  }
  if (tid == 0) s_base = 0;
  __syncthreads();                              // flags visible to the whole CTA
  const int n = B * L;
  for (int start = 0; start < n; start += 256) {
    const int i = start + tid;
    const int layer = i < n ? order[i] : 0;
    const bool keep = i < n && !(skip_shift && plans[layer].all_shift);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[wid] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < wid; ++w) off += s_warp[w];
    if (keep) work[off + __popc(m & ((1u << lane) - 1u))] = layer;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += s_warp[w];
      s_base += tot;
    }
    __syncthreads();
  }
  if (tid == 0) { wcnt[0] = s_base; wcnt[1] = 0; }
}

static __global__ void __launch_bounds__(256)
sample_flags_kernel(InverseLayer* __restrict__ plans, int B, int L, const int* __restrict__ order, int* __restrict__ work,
                    int* __restrict__ wcnt, int* __restrict__ sample_all_shift, int skip_shift) {
  sample_flags_cta(plans, B, L, order, work, wcnt, sample_all_shift, skip_shift);
}

// Both steps in ONE single-CTA launch for batches of up to kSmallPlacements layers (every configuration of the
// reference): placements, launch order (counters in shared memory, so no memset), flags, work list, and -- if asked --
// the zeroing of grad_theta.  Four tiny launches become one; they sit on the critical path in front of pass 1.
constexpr int kSmallPlacements = 1024;
static __global__ void __launch_bounds__(256)
placements_small_kernel(const float* __restrict__ theta, InverseLayer* __restrict__ plans, int B, int L, int H, int W,
                        int* __restrict__ order, int* __restrict__ work, int* __restrict__ wcnt, int* __restrict__ sample_all_shift,
                        float* __restrict__ gtheta_to_zero, int skip_shift) {
  __shared__ int s_cnt[2];
  const int n = B * L;
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  for (int k = threadIdx.x; k < n; k += 256) {
    bool heavy;
    const InverseLayer q = make_inverse_plan(theta + (long long)k * 6, H, W, heavy);
    if (heavy) order[atomicAdd(&s_cnt[0], 1)] = k;
    else order[n - 1 - atomicAdd(&s_cnt[1], 1)] = k;
    plans[k] = q;
    if (gtheta_to_zero) {
#pragma unroll
      for (int c = 0; c < 6; ++c) gtheta_to_zero[(long long)k * 6 + c] = 0.f;
    }
  }
  __syncthreads();                              // plans and order are visible to the whole CTA
  sample_flags_cta(plans, B, L, order, work, wcnt, sample_all_shift, skip_shift);
}

// Block of 256 threads = 32 x 8 threads, each owning a 2 x 2 block of texels -> 64 x 16 texels per CTA.
// A 2 x 2 block shares its candidate pixels (the union of four windows is barely larger than one), the
// record loads and the per-candidate coordinate arithmetic; the accumulators are packed fp32x2 pairs.
// Block shape: kTX threads along x (each thread 2 x 2 texels).  kTX = 16 -> 32 x 32 texels: the candidate pixels of
// neighbouring threads overlap more (-2.4 % backward at B*L >= 448 layers); kTX = 32 -> 64 x 16 texels: shorter blocks,
// which matters when a near-singular layer's blocks are the tail of a small launch (+15 us at B = 8..16 otherwise).
// The host picks by the number of blocks in the launch (launchers.cuh).
template <int kTX> struct P2Shape { static constexpr int kTY = 256 / kTX, kW = 2 * kTX, kH = 2 * kTY; };

template <typename T> struct Pack2;      // two horizontally adjacent texels of one channel -> one (aligned) store
template <> struct Pack2<float> {
  __device__ static __forceinline__ void store(float* p, float a, float b) {
    *reinterpret_cast<float2*>(p) = make_float2(a, b);
  }
};
template <> struct Pack2<__nv_bfloat16> {
  __device__ static __forceinline__ void store(__nv_bfloat16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
  }
};
template <> struct Pack2<__half> {
  __device__ static __forceinline__ void store(__half* p, float a, float b) {
    *reinterpret_cast<__half2*>(p) = __floats2half2_rn(a, b);
  }
};

// hat(u) = max(0, 1 - |u|): one FADD.SAT (1 - |u| never exceeds 1, so the upper clamp of the saturation is idle; a NaN
// coordinate yields 0 instead of NaN -- such placements are flagged invalid before the loops anyway)
__device__ __forceinline__ float hat(float u) { return __saturatef(1.f - fabsf(u)); }

#ifndef MGR_P2_BLOCKS
#define MGR_P2_BLOCKS 4
#endif
// Where the gather reads the gradient w.r.t. one warped sample (g_rgb, g_a) of an output pixel of the layer:
//   CompositeRecords: the fused renderer's pass-1 output -- (T_l a_l, d a_l) per layer-pixel times (G_P, .) per pixel;
//   PlanarGrads<T>:   the upstream gradient of MATERIALISED warped layers, four planes of T (mgr_warp_backward).
// A cursor points at a pixel; at(off) / advance(off) move it by pixels, load(mm) reads the pixel mm further on.
struct CompositeRecords {
  const float2* r;
  const float4* g;
  __device__ __forceinline__ CompositeRecords layer(int n, int b, int hw) const {    // from the tensors' first element
    return CompositeRecords{r + (long long)n * hw, g + (long long)b * hw};
  }
  __device__ __forceinline__ CompositeRecords at(int off) const { return CompositeRecords{r + off, g + off}; }
  __device__ __forceinline__ void advance(int off) { r += off; g += off; }
  __device__ __forceinline__ void load(int mm, float (&gv)[4]) const {
    const float2 rr = __ldg(r + mm);
    const float4 G = __ldg(g + mm);
    gv[0] = G.x * rr.x; gv[1] = G.y * rr.x; gv[2] = G.z * rr.x; gv[3] = rr.y;
  }
};
template <typename T>
struct PlanarGrads {
  const T* p;
  int hw;
  __device__ __forceinline__ PlanarGrads layer(int n, int, int hw_) const { return PlanarGrads{p + (long long)n * 4 * hw_, hw_}; }
  __device__ __forceinline__ PlanarGrads at(int off) const { return PlanarGrads{p + off, hw}; }
  __device__ __forceinline__ void advance(int off) { p += off; }
  __device__ __forceinline__ void load(int mm, float (&gv)[4]) const {
    gv[0] = ld(p + mm); gv[1] = ld(p + hw + mm); gv[2] = ld(p + 2 * hw + mm); gv[3] = ld(p + 3 * hw + mm);
  }
};

// one block of 2 kTX x 2 (256 / kTX) texels of layer n = b * L + l; `src` points at the first element of the
// gradient-record tensors; zs = d(sampled value) / d(texel) apart from the bilinear weight
template <typename T, bool kRagged, int kP2TX, typename Src>
__device__ __forceinline__ void pass2_block(const InverseLayer* __restrict__ plans, int n, int x0b, int y0b, const Src src,
                                            float zs, T* __restrict__ gx, const DstLayers* dst, const Geometry& g) {
  constexpr int kP2W = P2Shape<kP2TX>::kW, kP2H = P2Shape<kP2TX>::kH;
  constexpr bool kComposite = std::is_same<Src, CompositeRecords>::value;
  __shared__ float s_jcf, s_icf;
  __shared__ int s_JC, s_IC, s_ok;
  const int b = n / g.L;
  const int tx = threadIdx.x % kP2TX, ty = threadIdx.x / kP2TX;
  const InverseLayer& L_ = plans[n];
  const int hw = g.H * g.W;
  const int x = x0b + 2 * tx, y = y0b + 2 * ty;                // top-left texel of this thread's 2 x 2 block (canvas coordinates)
  // the layer's own pixels: a rectangle of the canvas (the whole canvas in the [B,L,4,H,W] layout); left, w even
  DstLayer dl;
  if (kRagged) dl = dst->s[n - b * g.L];
  else dl = DstLayer{gx + (long long)(n - b * g.L) * 4 * hw, (long long)g.L * 4 * hw, hw, g.W, g.H, g.W, 0, 0};
  const int xl = x - dl.left, yl = y - dl.top;
  const bool mine = (unsigned)xl < (unsigned)dl.w && yl >= -1 && yl < dl.h;      // at least one of the block's rows is inside
  // CTA-uniform: the whole 64 x 16 texel block lies outside this layer's rectangle (small layers of a ragged stack)
  if (x0b + kP2W <= dl.left || x0b >= dl.left + dl.w || y0b + kP2H <= dl.top || y0b >= dl.top + dl.h) return;
  T* gxp = reinterpret_cast<T*>(dl.ptr) + (long long)b * dl.sb + (long long)yl * dl.sh + xl;
  // acc[row][channel] = (texel column 0, texel column 1): the two columns share the candidate's gradient value, so a
  // candidate costs eight packed FMAs whose scalar operand is the gradient and whose pair operand is (w_y * w_x0, w_y * w_x1)
  f32x2 acc[2][4];
#pragma unroll
  for (int q = 0; q < 8; ++q) (&acc[0][0])[q] = 0ull;
  auto accumulate = [](f32x2 (&a)[2][4], float wx0, float wx1, float wy0, float wy1, const float (&gv)[4]) {
    const f32x2 wx = pk(wx0, wx1);
    const f32x2 w0 = mul2(wx, bc(wy0)), w1 = mul2(wx, bc(wy1));
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      a[0][c] = fma2(bc(gv[c]), w0, a[0][c]);
      a[1][c] = fma2(bc(gv[c]), w1, a[1][c]);
    }
  };

  if (L_.shift_only) {
    // pure translation: texel (x, y) is tap (dx, dy) of pixel (x - X - dx, y - Y - dy) with the layer-wide
    // weights (dx ? fx : 1 - fx)(dy ? fy : 1 - fy): a fixed 2 x 2 stencil; the 2 x 2 block reads 3 x 3 records
    if (!mine) return;
    const float wx[2] = {1.f - L_.fx, L_.fx}, wy[2] = {1.f - L_.fy, L_.fy};
#pragma unroll
    for (int di = -1; di <= 1; ++di) {
      const int i = y - L_.Y + di;
      if ((unsigned)i >= (unsigned)g.H) continue;
#pragma unroll
      for (int dj = -1; dj <= 1; ++dj) {
        const int j = x - L_.X + dj;
        if ((unsigned)j >= (unsigned)g.W) continue;
        float gv[4];
        src.layer(n, b, hw).load(i * g.W + j, gv);
        // tap index of texel (row ky, column kx) for this pixel: (ky - di, kx - dj), weight 0 outside {0, 1}
        const float ux0 = (dj == 0) ? wx[0] : (dj == -1 ? wx[1] : 0.f), ux1 = (dj == 1) ? wx[0] : (dj == 0 ? wx[1] : 0.f);
        const float uy0 = (di == 0) ? wy[0] : (di == -1 ? wy[1] : 0.f), uy1 = (di == 1) ? wy[0] : (di == 0 ? wy[1] : 0.f);
        accumulate(acc, ux0, ux1, uy0, uy1, gv);
      }
    }
  } else {
    if (threadIdx.x == 0) {
      // pre-image of the CTA's centre, split into integer + fraction in double precision
      const double xc = x0b + 0.5 * kP2W - L_.c0, yc = y0b + 0.5 * kP2H - L_.c1;
      const double jc = L_.i00 * xc + L_.i01 * yc, ic = L_.i10 * xc + L_.i11 * yc;
      const int ok = L_.valid && isfinite(jc) && isfinite(ic) && fabs(jc) < 1.0e8 && fabs(ic) < 1.0e8;
      const double fj = floor(jc), fi = floor(ic);
      s_JC = ok ? (int)fj : 0; s_IC = ok ? (int)fi : 0;
      s_jcf = ok ? (float)(jc - fj) : 0.f; s_icf = ok ? (float)(ic - fi) : 0.f;
      s_ok = ok;
    }
    const float a00 = L_.a00, a01 = L_.a01, a10 = L_.a10, a11 = L_.a11;
    const float i00 = L_.f00, i01 = L_.f01, i10 = L_.f10, i11 = L_.f11;
    // half extents of the pre-image of the 2 x 2 block's support: a texel's (-1,1)^2 grown by +-0.5
    const float rj = L_.rj + 0.5f * (fabsf(i00) + fabsf(i01)), ri = L_.ri + 0.5f * (fabsf(i10) + fabsf(i11));
    const int wide = L_.wide;
    const float r00 = L_.r00, r10 = L_.r10;
    __syncthreads();
    const int JC = s_JC, IC = s_IC;
    const float jcf = s_jcf, icf = s_icf;
    // centre of this thread's 2 x 2 block relative to the CTA centre, and its pre-image relative to (JC, IC)
    const float dxl = (float)(2 * tx) + 0.5f - 0.5f * kP2W, dyl = (float)(2 * ty) + 0.5f - 0.5f * kP2H;
    const float pj = jcf + i00 * dxl + i01 * dyl, pi = icf + i10 * dxl + i11 * dyl;
    float mlo = fmaxf(ceilf(pj - rj), (float)(-JC)), mhi = fminf(floorf(pj + rj), (float)(g.W - 1 - JC));
    float nlo = fmaxf(ceilf(pi - ri), (float)(-IC)), nhi = fminf(floorf(pi + ri), (float)(g.H - 1 - IC));
    const bool has = s_ok && mine && mlo <= mhi && nlo <= nhi;
    if (!has) { mlo = 0.f; mhi = -1.f; nlo = 0.f; nhi = -1.f; }
    float x0l = dxl - 0.5f, y0l = dyl - 0.5f;                  // texel (0,0) of the block relative to the CTA centre
    const Src src0 = src.layer(n, b, hw).at(IC * g.W + JC);

    // one candidate row nn of the window (mlo_..mhi_) of the block at (x0l_, y0l_), accumulated into acc_
    auto row = [&](int nn, float mlo_, float mhi_, float x0l_, float y0l_, f32x2 (&acc_)[2][4]) {
      const float di = (float)nn - icf;
      const float ub = fmaf(a01, di, -x0l_), vb = fmaf(a11, di, -y0l_);
      float lo = mlo_, hi = mhi_;
      {
        // magnifying or near-singular placement: solve |a00 dj + ub - kx| < 1, |a10 dj + vb - ky| < 1 for the row
        // (dj = mm - jcf; kx, ky in {0, 1}) and keep one candidate of slack on either side (its weight is 0)
        float l2 = -3.0e9f, h2 = 3.0e9f;
        if (r00 != 0.f) {
          const float t0 = (-1.f - ub) * r00, t1 = (2.f - ub) * r00;
          l2 = fmaxf(l2, fminf(t0, t1)); h2 = fminf(h2, fmaxf(t0, t1));
        } else if (ub <= -1.f || ub >= 2.f) { h2 = -3.0e9f; }
        if (r10 != 0.f) {
          const float t0 = (-1.f - vb) * r10, t1 = (2.f - vb) * r10;
          l2 = fmaxf(l2, fminf(t0, t1)); h2 = fminf(h2, fmaxf(t0, t1));
        } else if (vb <= -1.f || vb >= 2.f) { h2 = -3.0e9f; }
        lo = fmaxf(floorf(l2 + jcf) - 1.f, mlo_);
        hi = fminf(ceilf(h2 + jcf) + 1.f, mhi_);
        if (!(lo <= hi)) return;
      }
      const int ma = (int)lo, mb = (int)hi;
      const float dja = lo - jcf;
      float u = fmaf(a00, dja, ub);                            // ix(candidate) - x of texel column 0
      float v = fmaf(a10, dja, vb);                            // iy(candidate) - y of texel row 0
      const Src cur = src0.at(nn * g.W);
#pragma unroll 1
      for (int mm = ma; mm <= mb; ++mm, u += a00, v += a10) {
        float gv[4];
        cur.load(mm, gv);
        accumulate(acc_, hat(u), hat(u - 1.f), hat(v), hat(v - 1.f), gv);
      }
    };

    if (!wide) {
      // the common case (|det| ~ 1): a 3..5 x 3..5 candidate window, plain serial loops, no refinement
      if (has) {
        const int m0 = (int)mlo, m1 = (int)mhi, n0 = (int)nlo, n1 = (int)nhi;
        const float dj0 = mlo - jcf;
        float di = nlo - icf;
        // (the two row pointers are spelled out: a cursor object costs the fused renderer's hot loop six instructions
        //  after register allocation)
        Src cur = src0.at(n0 * g.W);
        const float2* recn = nullptr;
        const float4* gpb = nullptr;
        if constexpr (kComposite) { recn = src0.r + n0 * g.W; gpb = src0.g + n0 * g.W; }
        for (int nn = n0; nn <= n1; ++nn, di += 1.f) {
          float u = fmaf(a00, dj0, fmaf(a01, di, -x0l));
          float v = fmaf(a10, dj0, fmaf(a11, di, -y0l));
          if constexpr (kComposite) {
#pragma unroll 1
            for (int mm = m0; mm <= m1; ++mm, u += a00, v += a10) {
              const float2 r = __ldg(recn + mm);
              const float4 G = __ldg(gpb + mm);
              const float gv[4] = {G.x * r.x, G.y * r.x, G.z * r.x, r.y};
              accumulate(acc, hat(u), hat(u - 1.f), hat(v), hat(v - 1.f), gv);
            }
            recn += g.W; gpb += g.W;
          } else {
#pragma unroll 1
            for (int mm = m0; mm <= m1; ++mm, u += a00, v += a10) {
              float gv[4];
              cur.load(mm, gv);
              accumulate(acc, hat(u), hat(u - 1.f), hat(v), hat(v - 1.f), gv);
            }
            cur.advance(g.W);
          }
        }
      }
    } else {
      // Windows with many rows (near-singular / strongly magnifying placements) occur on a handful of lanes
      // per warp; left alone they serialise hundreds of dependent row iterations on those lanes.  The warp
      // takes them one at a time instead: 32 lanes share the rows of ONE block, then a butterfly adds the partials.
      const int lane = threadIdx.x & 31;
      const bool heavy = has && (nhi - nlo >= 31.f);
      unsigned todo = __ballot_sync(0xffffffffu, heavy);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const float mlo_s = __shfl_sync(0xffffffffu, mlo, src), mhi_s = __shfl_sync(0xffffffffu, mhi, src);
        const int n0_s = (int)__shfl_sync(0xffffffffu, nlo, src), n1_s = (int)__shfl_sync(0xffffffffu, nhi, src);
        const float x0l_s = __shfl_sync(0xffffffffu, x0l, src), y0l_s = __shfl_sync(0xffffffffu, y0l, src);
        f32x2 part[2][4];
#pragma unroll
        for (int q = 0; q < 8; ++q) (&part[0][0])[q] = 0ull;
        for (int nn = n0_s + lane; nn <= n1_s; nn += 32) row(nn, mlo_s, mhi_s, x0l_s, y0l_s, part);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float p0, p1;
          upk((&part[0][0])[q], p0, p1);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            p0 += __shfl_xor_sync(0xffffffffu, p0, o);
            p1 += __shfl_xor_sync(0xffffffffu, p1, o);
          }
          if (lane == src) (&acc[0][0])[q] = pk(p0, p1);
        }
      }
      if (has && !heavy) {
        const int n0 = (int)nlo, n1 = (int)nhi;
        for (int nn = n0; nn <= n1; ++nn) row(nn, mlo, mhi, x0l, y0l, acc);
      }
    }
    if (!mine) return;
  }
#pragma unroll
  for (int ky = 0; ky < 2; ++ky) {
    if ((unsigned)(yl + ky) >= (unsigned)dl.h) continue;
    T* o = gxp + ky * dl.sh;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float t0, t1;                                    // texel columns 0 and 1 of channel c
      upk(acc[ky][c], t0, t1);
      Pack2<T>::store(o + c * dl.sc, zs * t0, zs * t1);    // w % 4 == 0 and left % 4 == 0 on this path: xl + 1 < w
    }
  }
}

// One CTA per (layer of the work list, 64 x 16 texel block).  The grid's z extent is B * L (the host cannot know how
// many layers the work list holds); CTAs beyond wcnt[0] leave after that one load -- for a batch of translations that
// is all of them.  (Walking the items inside persistent or grid-stride CTAs was measured 3-8 % slower on general batches.)
template <typename T, bool kRagged, int kTX>
__global__ void __launch_bounds__(256, MGR_P2_BLOCKS)
render_bwd_pass2(const InverseLayer* __restrict__ plans, const int* __restrict__ work, const int* __restrict__ wcnt,
                 const float2* __restrict__ rec, const float4* __restrict__ gp, T* __restrict__ gx,
                 const __grid_constant__ DstLayers dst, Geometry g) {
  if ((int)blockIdx.z >= wcnt[0]) return;
  pass2_block<T, kRagged, kTX>(plans, work[blockIdx.z], blockIdx.x * P2Shape<kTX>::kW, blockIdx.y * P2Shape<kTX>::kH,
                               CompositeRecords{rec, gp}, g.m11 ? 0.5f : 1.f, gx, &dst, g);
}

}  // namespace mgr
