// Kernels for stacks whose placements are all PURE TRANSLATIONS -- exactly what the reference's placement
// network emits (STNv2c builds theta with convert_translate_to_2x3: [[1,0,dx],[0,1,dy]],
// fukuwarai/networks.py:246-247, custom_utils/image_utils.py:316-335).  Then
//     ix = j + sx,  iy = i + sy,   sx = dx * W/2, sy = dy * H/2,
// the tap origin is (j + X, i + Y) with X = floor(sx), Y = floor(sy) and the bilinear weights
// (fx, fy) are the SAME for every pixel of the layer: the warp is a 2x2 stencil.
//
//  * forward: a thread owns a vertical strip of four pixels, so five staged rows serve all four
//    (10 tap loads and 10 unpacks per 4 pixels instead of 16 + 16), horizontal lerps are shared
//    between vertically adjacent pixels, and no per-pixel floor / address arithmetic is left.
//  * backward: ONE fused kernel, no records, no second pass, no atomics on grad_x.  The adjoint of
//    a 2x2 stencil is a 2x2 stencil: texel (a + X, b + Y) collects from pixels (a-1..a, b-1..b).
//    CTAs own 31x31 "anchors" (a, b) and compute the composite adjoint on a 32x32 pixel tile that
//    overlaps its left/top neighbours by one pixel, exchange the per-pixel gradient records through
//    shared memory, and write every grad_x element exactly once (texels no pixel touches are
//    zero-filled by the CTA that owns the same coordinates unshifted).
//
// A CTA decides by itself (all layers of its sample are pure translations, checked on the theta
// values) whether it runs; the general kernels make the opposite decision, so the two launches
// partition the batch.  Math and reference semantics as in render_tiled.cuh / render_bwd_tiled.cuh.
#pragma once
#include "render_tiled.cuh"

// resident CTAs per SM the stencil kernels are compiled for (developer knobs, tools/ab.py)
#ifndef MGR_SHB_BLOCKS
#define MGR_SHB_BLOCKS 2
#endif
#ifndef MGR_SHF_BLOCKS
#define MGR_SHF_BLOCKS 3
#endif

namespace mgr {

struct ShiftPlan {     // 16 bytes: arrays behind a ShiftPlan[L] stay 16-byte aligned
  int X, Y;          // integer part of the shift (pixels)
  float fx, fy;      // fractional part, in [0, 1)
};

static_assert(sizeof(ShiftPlan) == 16, "shared-memory layout of the stencil kernels");

__device__ __forceinline__ ShiftPlan make_shift_plan(const float* __restrict__ th, int H, int W) {
  const double sx = (double)th[2] * 0.5 * W, sy = (double)th[5] * 0.5 * H;
  const double fX = floor(sx), fY = floor(sy);
  ShiftPlan p;
  p.X = (int)fX; p.Y = (int)fY; p.fx = (float)(sx - fX); p.fy = (float)(sy - fY);
  return p;
}

// footprint of the pixel tile whose top-left pixel is (j0, i0) under a shift plan, as a LayerPlan
__device__ __forceinline__ LayerPlan shift_footprint(const ShiftPlan& sp, int j0, int i0, const SrcLayer& src) {   // src: rectangle only
  LayerPlan p;
  p.x_lo = (j0 + sp.X) & ~(kStageVec - 1);
  p.y_lo = i0 + sp.Y;
  p.bw = ((j0 + sp.X + kTW + 1) - p.x_lo + kStageVec - 1) & ~(kStageVec - 1);     // columns x_lo .. j0+X+32
  p.bh = kTH + 1;
  const bool miss = (p.x_lo + p.bw <= src.left) || (p.x_lo >= src.left + src.w) || (p.y_lo + p.bh <= src.top) || (p.y_lo >= src.top + src.h);
  p.mode = miss ? kSkip : kStaged;
  p.lrx = p.lry = 0.f; p.pitch = p.bw; p.dX = p.dY = 0; p.pad2_[0] = p.pad2_[1] = 0;
  return p;
}

constexpr int kShiftCap = (kTW + 2 * kStageVec) * (kTH + 1);        // staged texels per layer (40 x 33)

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <typename T, bool kSave>
__device__ __forceinline__ void fwd_shift_body(const SrcLayers& src, const float* __restrict__ theta, T* __restrict__ out,
                                               typename SavedAlpha<T>::type* __restrict__ sav, const Geometry& g) {
  using Vec = typename Texel<T>::Vec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Vec* buf = reinterpret_cast<Vec*>(smem_raw);                                           // [kShiftCap]
  ShiftPlan* splan = reinterpret_cast<ShiftPlan*>(smem_raw + sizeof(Vec) * kShiftCap);   // [L]
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const float* thb = theta + (long long)b * g.L * 6;
  const int j0 = blockIdx.x * kTW, i0 = blockIdx.y * kTH;
  const int tx = tid & 31, ty = tid >> 5;                    // pixels (tx, 4 ty + k), k = 0..3
  for (int l = tid; l < g.L; l += kTiledThreads) splan[l] = make_shift_plan(thb + 6 * l, g.H, g.W);
  __syncthreads();

  const f32x2 zs2 = bc(g.m11 ? 0.5f : 1.f), zb2 = bc(g.m11 ? 0.5f : 0.f);     // z = zs * raw + zb
  const int hw = g.H * g.W;
  const int j = j0 + tx;
  const int pix0 = (i0 + kPx * ty) * g.W + j;
  bool live[kPx];
#pragma unroll
  for (int k = 0; k < kPx; ++k) live[k] = j < g.W && i0 + kPx * ty + k < g.H;
  float S0[kPx], S1[kPx], S2[kPx], R[kPx];
#pragma unroll
  for (int k = 0; k < kPx; ++k) S0[k] = S1[k] = S2[k] = R[k] = 0.f;

  for (int l = 0; l < g.L; ++l) {
    const ShiftPlan sp = splan[l];
    const LayerPlan p = shift_footprint(sp, j0, i0, src.s[l]);
    typename SavedAlpha<T>::type* sv = nullptr;
    if (kSave) sv = sav + ((long long)b * g.L + l) * hw + pix0;
    if (p.mode == kSkip) {
      if (kSave) {
#pragma unroll
        for (int k = 0; k < kPx; ++k)
          if (live[k]) st_alpha(sv + k * g.W, 0.f);
      }
      continue;
    }
    stage_footprint<T, true>(g.m11 != 0, src_view<T>(src.s[l], b), p, buf, tid);   // takes the "readers are done" barrier with its loads in flight
    __syncthreads();
    const f32x2 fx2 = bc(sp.fx), fy2 = bc(sp.fy);
    const Vec* q = buf + (kPx * ty) * p.bw + (tx + j0 + sp.X - p.x_lo);
    f32x2 h_rg, h_ba;                                         // horizontally interpolated row
    {
      f32x2 a_rg, a_ba, b_rg, b_ba;
      Texel<T>::unpack(q[0], a_rg, a_ba);
      Texel<T>::unpack(q[1], b_rg, b_ba);
      h_rg = fma2(fx2, sub2(b_rg, a_rg), a_rg);
      h_ba = fma2(fx2, sub2(b_ba, a_ba), a_ba);
    }
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
      q += p.bw;
      f32x2 a_rg, a_ba, b_rg, b_ba;
      Texel<T>::unpack(q[0], a_rg, a_ba);
      Texel<T>::unpack(q[1], b_rg, b_ba);
      const f32x2 n_rg = fma2(fx2, sub2(b_rg, a_rg), a_rg), n_ba = fma2(fx2, sub2(b_ba, a_ba), a_ba);
      float r_, g_, b_, a;
      upk(fma2(fma2(fy2, sub2(n_rg, h_rg), h_rg), zs2, zb2), r_, g_);
      upk(fma2(fma2(fy2, sub2(n_ba, h_ba), h_ba), zs2, zb2), b_, a);
      h_rg = n_rg; h_ba = n_ba;
      if (kSave) { if (live[k]) st_alpha(sv + k * g.W, a); }
      const float om = 1.f - a;
      S0[k] = fmaf(om, S0[k], a * r_);
      S1[k] = fmaf(om, S1[k], a * g_);
      S2[k] = fmaf(om, S2[k], a * b_);
      R[k] = fmaf(om, R[k], a);
    }
  }

  const float os = g.m11 ? 2.f : 1.f, obias = g.m11 ? -1.f : 0.f;
  T* outp = out + (long long)b * 4 * hw + pix0;
#pragma unroll
  for (int k = 0; k < kPx; ++k) {
    if (live[k]) {
      const float inv = (R[k] != 0.f) ? 1.f / R[k] : 0.f;     // nan_to_num(0/0) = 0 (image_utils.py:132)
      T* o = outp + k * g.W;
      st(o, fmaf(S0[k] * inv, os, obias));
      st(o + hw, fmaf(S1[k] * inv, os, obias));
      st(o + 2 * hw, fmaf(S2[k] * inv, os, obias));
      st(o + 3 * hw, fmaf(R[k], os, obias));
    }
  }
}

inline size_t shift_fwd_smem_bytes(int L, size_t vec_bytes) { return vec_bytes * kShiftCap + sizeof(ShiftPlan) * L; }

// ---------------------------------------------------------------------------------------------
// fused backward: composite adjoint + theta gradient + stencil adjoint, anchors 31 x 31 per CTA
// ---------------------------------------------------------------------------------------------
constexpr int kAnchor = kTW - 1;

template <typename T, bool kNeedX, bool kNeedTheta, bool kGPSmem>
__global__ void __launch_bounds__(kTiledThreads, kGPSmem ? MGR_SHB_BLOCKS : 2)
render_bwd_shift(const __grid_constant__ SrcLayers src, const float* __restrict__ theta, const T* __restrict__ out,
                 const T* __restrict__ gout, const typename SavedAlpha<T>::type* __restrict__ sav,
                 const __grid_constant__ DstLayers dst, float* __restrict__ gtheta, float4* __restrict__ gp, Geometry g,
                 const int* __restrict__ sample_all_shift) {
  using Vec = typename Texel<T>::Vec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Vec* buf = reinterpret_cast<Vec*>(smem_raw);                                           // [kShiftCap]
  float4* G = reinterpret_cast<float4*>(smem_raw + sizeof(Vec) * kShiftCap);             // [32][32] gradient records
  ShiftPlan* splan = reinterpret_cast<ShiftPlan*>(G + kTW * kTH);                        // [L], 16 B each
  const int tid = threadIdx.x;
  float* stash = reinterpret_cast<float*>(splan + g.L);       // [L][kPx][256]: T_l, later the theta-gradient partials
  float* Tst = stash + tid;
  const int b = blockIdx.z;
  const float* thb = theta + (long long)b * g.L * 6;
  if (!sample_all_shift[b]) return;                           // per-sample flag from sample_flags_kernel
  // pixel tile: origin one pixel left/up of the anchors it owns; neighbouring tiles overlap by one pixel
  const int j0 = blockIdx.x * kAnchor - 1, i0 = blockIdx.y * kAnchor - 1;
  const int tx = tid & 31, ty = tid >> 5;
  for (int l = tid; l < g.L; l += kTiledThreads) splan[l] = make_shift_plan(thb + 6 * l, g.H, g.W);
  __syncthreads();

  const float zs = g.m11 ? 0.5f : 1.f;
  const f32x2 zs2 = bc(zs), zb2 = bc(g.m11 ? 0.5f : 0.f);
  const int hw = g.H * g.W;
  const int j = j0 + tx;
  const int ibase = i0 + kPx * ty;
  const int pix0 = ibase * g.W + j;                           // may be "negative" for the halo; only used when live
  bool live[kPx];       // a real pixel of the image
  bool own[kPx];        // a pixel whose theta-gradient this CTA accounts for (halo pixels belong to the neighbour)
#pragma unroll
  for (int k = 0; k < kPx; ++k) {
    const int i = ibase + k;
    live[k] = (unsigned)j < (unsigned)g.W && (unsigned)i < (unsigned)g.H;
    own[k] = live[k] && tx >= 1 && (kPx * ty + k) >= 1;
  }
  const typename SavedAlpha<T>::type* savb = sav + (long long)b * g.L * hw + pix0;
  // (G_P, G_A) per pixel lives in the workspace (re-read per layer: L1/L2 hits).  Halo pixels are written by
  // two or four overlapping tiles with identical values.
  float4* gpp = gp + (long long)b * hw + pix0;
  float4* GPs = reinterpret_cast<float4*>(stash + (size_t)g.L * kPx * kTiledThreads) + tid;   // [kPx][256], if it fits

  // ---- pre-pass: T_l and A from the saved alpha samples ---------------------------------------------
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
        for (int c = 0; c < 4; ++c) gv[k][c] = live[k] ? ld(gob + k * g.W + c * hw) : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) ov[k][c] = live[k] ? ld(ob_ + k * g.W + c * hw) : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < kPx; ++k) { Tc[k] = 1.f; A[k] = 0.f; }
    // running pointers (front layer first): keeps the loop free of 64-bit index arithmetic
    const typename SavedAlpha<T>::type* sa[kPx];
#pragma unroll
    for (int k = 0; k < kPx; ++k) sa[k] = savb + (long long)(g.L - 1) * hw + k * g.W;
    float* tp = Tst + (g.L - 1) * kPx * kTiledThreads;
    for (int l = g.L - 1; l >= 0; --l) {
#pragma unroll
      for (int k = 0; k < kPx; ++k) {
        tp[k * kTiledThreads] = live[k] ? Tc[k] : 0.f;
        if (live[k]) {
          const float a = ld_alpha(sa[k]);
          A[k] = __fmaf_rn(Tc[k], a, A[k]);
          Tc[k] = __fmul_rn(Tc[k], __fsub_rn(1.f, a));
        }
        sa[k] -= hw;
      }
      tp -= kPx * kTiledThreads;
    }
    const float gs = g.m11 ? 2.f : 1.f;                       // d out / d o
    const float is = g.m11 ? 0.5f : 1.f, ib = g.m11 ? 0.5f : 0.f;
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
      GP0[k] = GP1[k] = GP2[k] = GA[k] = 0.f;
      if (live[k]) {
        const float g0 = __fmul_rn(gs, gv[k][0]), g1 = __fmul_rn(gs, gv[k][1]),
                    g2 = __fmul_rn(gs, gv[k][2]), g3 = __fmul_rn(gs, gv[k][3]);
        if (A[k] != 0.f) {
          const float inv = 1.f / A[k];
          const float o0 = __fmaf_rn(ov[k][0], is, ib), o1 = __fmaf_rn(ov[k][1], is, ib),
                      o2 = __fmaf_rn(ov[k][2], is, ib);
          // Explicitly rounded steps: a halo pixel is computed by up to four overlapping CTAs (as different unrolled
          // instances k), and when (G_P, G_A) lives in global memory they all store it to the same address -- the
          // stores must carry identical bits, so the compiler may not contract or re-associate per instance.
          GP0[k] = __fmul_rn(g0, inv); GP1[k] = __fmul_rn(g1, inv); GP2[k] = __fmul_rn(g2, inv);
          GA[k] = __fsub_rn(g3, __fmul_rn(__fmaf_rn(g2, o2, __fmaf_rn(g1, o1, __fmul_rn(g0, o0))), inv));
        }
      }
      if (kGPSmem) GPs[k * kTiledThreads] = make_float4(GP0[k], GP1[k], GP2[k], GA[k]);
      else if (live[k]) gpp[k * g.W] = make_float4(GP0[k], GP1[k], GP2[k], GA[k]);
    }
  }

  const float xj = norm_coord(j, g.W);
  const float hW = 0.5f * (float)g.W * zs, hH = 0.5f * (float)g.H * zs;
  float S0[kPx], S1[kPx], S2[kPx], R[kPx];
#pragma unroll
  for (int k = 0; k < kPx; ++k) S0[k] = S1[k] = S2[k] = R[k] = 0.f;
  float4* Gt = G + (kPx * ty) * kTW + tx;                     // this thread's records: Gt[k * kTW]
  const bool colok = tx >= 1 && j <= g.W;                     // anchor column a in [0, W]: pixel W is virtual (record 0)

  for (int l = 0; l < g.L; ++l) {
    const ShiftPlan sp = splan[l];
    const LayerPlan p = shift_footprint(sp, j0, i0, src.s[l]);
    if (p.mode == kSkip) {
      // the footprint misses the image: a transparent-black layer for this tile (a = 0): the canvas is unchanged,
      // no theta gradient, and none of this tile's anchor texels lies inside the image.  What remains is to zero
      // the texels at this tile's own (unshifted) coordinates that no pixel touches.
      if (kNeedX) {
        const DstLayer& dl = dst.s[l];
        T* gxl = reinterpret_cast<T*>(dl.ptr) + (long long)b * dl.sb;
        const int xl = j - dl.left;                            // this thread's own column inside the layer's rectangle
#pragma unroll
        for (int k = 0; k < kPx; ++k) {
          const int brow = ibase + k, yl = brow - dl.top;
          if (tx >= 1 && (kPx * ty + k) >= 1 && (unsigned)xl < (unsigned)dl.w && (unsigned)yl < (unsigned)dl.h) {
            const int pa = j - sp.X, pb = brow - sp.Y;
            if (pa < 0 || pa > g.W || pb < 0 || pb > g.H) {
              T* o = gxl + yl * dl.sh + xl;
              st(o, 0.f); st(o + dl.sc, 0.f); st(o + 2 * dl.sc, 0.f); st(o + 3 * dl.sc, 0.f);
            }
          }
        }
      }
      if (kNeedTheta) park_theta_partials(Tst + l * kPx * kTiledThreads, 0.f, 0.f, 0.f, 0.f);
      continue;
    }
    __syncthreads();                                          // previous layer: buf readers and G readers are done
    stage_footprint<T>(g.m11 != 0, src_view<T>(src.s[l], b), p, buf, tid);
    __syncthreads();
    float accx = 0.f, accxy = 0.f, accy = 0.f, accyy = 0.f;
    {
      const f32x2 fx2 = bc(sp.fx), fy2 = bc(sp.fy);
      const Vec* q = buf + (kPx * ty) * p.bw + (tx + j0 + sp.X - p.x_lo);
      f32x2 h_rg, h_ba, d_rg, d_ba;                           // row: lerped value and horizontal difference
      {
        f32x2 a_rg, a_ba, b_rg, b_ba;
        Texel<T>::unpack(q[0], a_rg, a_ba);
        Texel<T>::unpack(q[1], b_rg, b_ba);
        d_rg = sub2(b_rg, a_rg); d_ba = sub2(b_ba, a_ba);
        h_rg = fma2(fx2, d_rg, a_rg); h_ba = fma2(fx2, d_ba, a_ba);
      }
#pragma unroll
      for (int k = 0; k < kPx; ++k) {
        q += p.bw;
        f32x2 a_rg, a_ba, b_rg, b_ba;
        Texel<T>::unpack(q[0], a_rg, a_ba);
        Texel<T>::unpack(q[1], b_rg, b_ba);
        const f32x2 e_rg = sub2(b_rg, a_rg), e_ba = sub2(b_ba, a_ba);
        const f32x2 n_rg = fma2(fx2, e_rg, a_rg), n_ba = fma2(fx2, e_ba, a_ba);
        const f32x2 dy_rg = sub2(n_rg, h_rg), dy_ba = sub2(n_ba, h_ba);                // d raw / d iy
        const f32x2 dx_rg = fma2(fy2, sub2(e_rg, d_rg), d_rg), dx_ba = fma2(fy2, sub2(e_ba, d_ba), d_ba);
        float r_, g_, b_, a;
        upk(fma2(fma2(fy2, dy_rg, h_rg), zs2, zb2), r_, g_);
        upk(fma2(fma2(fy2, dy_ba, h_ba), zs2, zb2), b_, a);
        h_rg = n_rg; h_ba = n_ba; d_rg = e_rg; d_ba = e_ba;
        const float T_l = Tst[(l * kPx + k) * kTiledThreads];
        const float ta = T_l * a;
        const float4 G4 = kGPSmem ? GPs[k * kTiledThreads] : (live[k] ? gpp[k * g.W] : make_float4(0.f, 0.f, 0.f, 0.f));
        const float ga = T_l * (G4.x * (r_ - S0[k]) + G4.y * (g_ - S1[k]) + G4.z * (b_ - S2[k]) + G4.w * (1.f - R[k]));
        const float gr = G4.x * ta, gg = G4.y * ta, gb = G4.z * ta;
        if (kNeedX) Gt[k * kTW] = make_float4(gr, gg, gb, ga);
        if (kNeedTheta) {
          float dxr, dxg, dxb, dxa, dyr, dyg, dyb, dya;
          upk(dx_rg, dxr, dxg); upk(dx_ba, dxb, dxa);
          upk(dy_rg, dyr, dyg); upk(dy_ba, dyb, dya);
          if (own[k]) {
            const float dix = fmaf(gr, dxr, fmaf(gg, dxg, fmaf(gb, dxb, ga * dxa)));
            const float diy = fmaf(gr, dyr, fmaf(gg, dyg, fmaf(gb, dyb, ga * dya)));
            const float yi = norm_coord(ibase + k, g.H);
            accx += dix; accxy = fmaf(dix, yi, accxy);
            accy += diy; accyy = fmaf(diy, yi, accyy);
          }
        }
        const float om = 1.f - a;
        S0[k] = fmaf(om, S0[k], a * r_);
        S1[k] = fmaf(om, S1[k], a * g_);
        S2[k] = fmaf(om, S2[k], a * b_);
        R[k] = fmaf(om, R[k], a);
      }
    }
    if (kNeedTheta) park_theta_partials(Tst + l * kPx * kTiledThreads, accx, accxy, accy, accyy);   // T_l slots are dead
    if (kNeedX) {
      __syncthreads();                                        // records of the whole tile are in G
      // anchors (a, b) = (j, ibase + k), tx >= 1, row >= 1: texel (a + X, b + Y) gets
      //   (1-fy)[(1-fx) g(a, b) + fx g(a-1, b)] + fy[(1-fx) g(a, b-1) + fx g(a-1, b-1)]
      const DstLayer& dl = dst.s[l];
      const int X = j + sp.X - dl.left;                       // texel column inside the layer's rectangle
      const bool xin = colok && (unsigned)X < (unsigned)dl.w;
      const f32x2 wx0 = bc((1.f - sp.fx) * zs), wx1 = bc(sp.fx * zs), wy0 = bc(1.f - sp.fy), wy1 = bc(sp.fy);
      T* gxl = reinterpret_cast<T*>(dl.ptr) + (long long)b * dl.sb;
      const int dsc = (int)dl.sc, dsh = (int)dl.sh;            // one layer of one sample spans < 2^31 elements (host-checked)
      const int xl = j - dl.left;                             // own column, for the zero-fill of untouched texels
      // CTA-uniform: can a texel at this tile's own (unshifted) coordinates be out of every pixel's reach?
      const bool zfill = j0 + 1 - sp.X < 0 || j0 + kAnchor - sp.X > g.W || i0 + 1 - sp.Y < 0 || i0 + kAnchor - sp.Y > g.H;
      f32x2 hp_rg = bc(0.f), hp_ba = bc(0.f);
      if (tx >= 1 && ty >= 1) {                               // row kPx*ty - 1 is inside the tile
        const float4 c = Gt[-kTW], d = Gt[-kTW - 1];
        hp_rg = fma2(wx1, pk(d.x, d.y), mul2(wx0, pk(c.x, c.y)));
        hp_ba = fma2(wx1, pk(d.z, d.w), mul2(wx0, pk(c.z, c.w)));
      }
#pragma unroll
      for (int k = 0; k < kPx; ++k) {
        f32x2 hc_rg = bc(0.f), hc_ba = bc(0.f);
        if (tx >= 1) {
          const float4 c = Gt[k * kTW], d = Gt[k * kTW - 1];
          hc_rg = fma2(wx1, pk(d.x, d.y), mul2(wx0, pk(c.x, c.y)));
          hc_ba = fma2(wx1, pk(d.z, d.w), mul2(wx0, pk(c.z, c.w)));
        }
        const int brow = ibase + k;                           // anchor row b
        const bool rowok = (kPx * ty + k) >= 1 && brow <= g.H;
        const int Y = brow + sp.Y - dl.top;
        if (xin && rowok && (unsigned)Y < (unsigned)dl.h) {
          float v0, v1, v2, v3;
          upk(fma2(wy1, hp_rg, mul2(wy0, hc_rg)), v0, v1);
          upk(fma2(wy1, hp_ba, mul2(wy0, hc_ba)), v2, v3);
          T* o = gxl + (Y * dsh + X);
          st(o, v0); st(o + dsc, v1); st(o + 2 * dsc, v2); st(o + 3 * dsc, v3);
        }
        // the same coordinates taken as a TEXEL of this layer: if no pixel touches it, it is ours to zero
        const int yl = brow - dl.top;
        if (zfill && colok && rowok && (unsigned)xl < (unsigned)dl.w && (unsigned)yl < (unsigned)dl.h) {
          const int pa = j - sp.X, pb = brow - sp.Y;          // the anchor that would own texel (j, brow)
          if (pa < 0 || pa > g.W || pb < 0 || pb > g.H) {
            T* o = gxl + (yl * dsh + xl);
            st(o, 0.f); st(o + dsc, 0.f); st(o + 2 * dsc, 0.f); st(o + 3 * dsc, 0.f);
          }
        }
        hp_rg = hc_rg; hp_ba = hc_ba;
      }
    }
  }
  if (kNeedTheta) {
    __syncthreads();
    reduce_theta_partials(stash, g.L, tid, xj, hW, hH, gtheta + (long long)b * g.L * 6);
  }
}

inline size_t shift_bwd_smem_bytes(int L, size_t vec_bytes) {
  return vec_bytes * kShiftCap + sizeof(float4) * kTW * kTH + sizeof(ShiftPlan) * L +
         sizeof(float) * (size_t)L * kPx * kTiledThreads;
}

}  // namespace mgr
