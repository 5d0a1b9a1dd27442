// Fused backward for stacks of pure translations on TMA box copies (the staged kernel of render_shift.cuh keeps ragged
// stacks and strides that break TMA's 16-byte rule).  Same mathematics as render_bwd_shift -- composite adjoint in
// the scalar form of SURVEY.md Appendix A.3, the adjoint of the 2x2 stencil as a 2x2 stencil over per-pixel gradient records
// exchanged through shared memory, every grad_x element written exactly once, no atomics on grad_x -- rebuilt around what
// the forward of render_shift_tma.cuh showed: the box copy replaces the staging loop, a thread owns a 4-wide, 2-tall strip
// of pixels and works on planar rows with packed fp32x2 arithmetic on pixel pairs.
//
//   * pixel tile 64 x 16 at origin (63 bx - 1, 15 by - 1): tiles overlap their left / top neighbours by one pixel, a CTA
//     owns the 63 x 15 anchors that are not in its first column / row (texel (a + X, b + Y) collects from pixels
//     (a-1..a, b-1..b));
//   * one producer lane: at CTA start the saved alphas [L][16][64+], grad_out and out [4][16][64+] of the tile (three box
//     copies on one mbarrier), then the x footprints {64+ x 17 x 4} of the layers through a two-stage ring;
//   * pre-pass: transmittances T_l front to back from the saved alphas, written back IN PLACE of the alpha samples (the
//     main loop re-samples alpha from x and only needs T_l), and (G_P, G_A) per pixel into a thread-private shared slot
//     (kGlobalT -- fp32 tensors, 16-bit stacks past 19 layers: no pixel-space boxes, T_l parked in the workspace instead);
//   * per layer: sample alpha, then the colours channel by channel folding each into u = G_P.c + G_A and into the
//     G_P-weighted coordinate derivatives, so no channel's samples stay live; records (T a G_P, dA) to shared memory, one
//     barrier, stencil adjoint with the layer-uniform weights, stores; the six theta sums of the layer are reduced in the warp
//     at once (transposing butterfly) and parked per (layer, warp): one atomic per (CTA, layer, coefficient) at the end.
//
// Reference semantics: fukuwarai/networks.py:250-257 (warp), custom_utils/image_utils.py:128-146 (over) and their autograd.
#pragma once
#include "render_shift_tma.cuh"

#ifndef MGR_STB_STAGES
#define MGR_STB_STAGES 2
#endif
#ifndef MGR_STB_BLOCKS
#define MGR_STB_BLOCKS 3
#endif

namespace mgr {

#ifndef MGR_STB_TILE_H
#define MGR_STB_TILE_H 16
#endif
constexpr int kBW = 64, kBH = MGR_STB_TILE_H;     // pixel tile (height a multiple of 4: whole warps of 16 x 2 strips)
constexpr int kBAncW = kBW - 1, kBAncH = kBH - 1; // anchors a CTA owns
constexpr int kBConsumers = 8 * kBH;              // 16 x (kBH / 2) strips of 4 x 2 pixels
constexpr int kBThreads = kBConsumers + 32;
constexpr int kBStages = MGR_STB_STAGES;

template <typename T> struct BwdBox {
  using SA = typename SavedAlpha<T>::type;
  static_assert(sizeof(SA) == sizeof(T), "the alpha box shares the x box's alignment rule");
  static constexpr int kAlign = 16 / (int)sizeof(T);
  static constexpr int W = kBW + kAlign;
  static constexpr int kXRows = kBH + 1;
  static constexpr int kXPlane = W * kXRows;
  static constexpr int kXBytes = kXPlane * 4 * (int)sizeof(T);
  static constexpr int kXStage = (kXBytes + 127) & ~127;
  static constexpr int kPPlane = W * kBH;                       // one plane of a pixel-space box (alpha, grad_out, out)
  static constexpr int kGBytes = kPPlane * 4 * (int)sizeof(T);
  static constexpr int kGStage = (kGBytes + 127) & ~127;
  static constexpr int kRecBytes = 4 * kBH * kBW * 4;           // gradient records, planar fp32; aliases the grad_out + out boxes
  static_assert(kRecBytes <= 2 * kGStage, "records alias the grad_out / out boxes");
  static constexpr int kGPBytes = kBConsumers * 8 * 16;         // (G_P, G_A): 8 float4 per thread
  static __host__ __device__ constexpr int alpha_bytes(int L) { return (kPPlane * L * (int)sizeof(SA) + 127) & ~127; }
};

struct BwdSmem {          // byte offsets inside the dynamic shared memory window
  int x, gout, out, gp, alpha, red, plan, bars, total;
};
template <typename T>
__host__ __device__ inline BwdSmem bwd_tma_layout(int L, bool global_t) {
  using Box = BwdBox<T>;
  BwdSmem m;
  m.x = 0;
  m.gout = m.x + kBStages * Box::kXStage;
  if (global_t) {                          // no pixel-space boxes (see kGlobalT); `gout` is the record area, no alpha tile
    m.out = m.gout;
    m.gp = m.gout + Box::kRecBytes;
    m.alpha = m.gp + Box::kGPBytes;
    m.red = m.alpha;
  } else {
    m.out = m.gout + Box::kGStage;
    m.gp = m.out + Box::kGStage;
    m.alpha = m.gp + Box::kGPBytes;
    m.red = m.alpha + Box::alpha_bytes(L);
  }
  m.plan = m.red + L * (kBConsumers / 32) * 8 * (int)sizeof(float);
  m.bars = m.plan + L * (int)sizeof(ShiftPlan);
  m.total = m.bars + 64;
  return m;
}

__device__ __forceinline__ bool bwd_box_misses(int x0, int y0, int W, int H) {
  return x0 + kBW < 0 || x0 >= W || y0 + kBH < 0 || y0 >= H;
}

// [-1,1] range mode: out-of-image texels of an x box become -1 (see shift_patch_oob); kBConsumers threads
template <typename T>
__device__ __forceinline__ void bwd_patch_oob(T* stage, int x0, int y0, int W, int H, int tid) {
  constexpr int BW = BwdBox<T>::W, BH = BwdBox<T>::kXRows;
  const int nT = min(max(-y0, 0), BH), nB = min(max(y0 + BH - H, 0), BH);
  const int nL = min(max(-x0, 0), BW), nR = min(max(x0 + BW - W, 0), BW);
  const T m1 = raw_minus_one<T>();
  const int cl = tid & 7;                                     // 8 column lanes x 16 (row, channel) lanes
#pragma unroll 1
  for (int rc = tid >> 3; rc < 4 * BH; rc += kBConsumers / 8) {
    const int r = rc >> 2, c = rc & 3;
    T* row = stage + c * BwdBox<T>::kXPlane + r * BW;
    if (r < nT || r >= BH - nB) {
#pragma unroll 1
      for (int k = cl; k < BW; k += 8) row[k] = m1;
    } else {
#pragma unroll 1
      for (int k = cl; k < nL; k += 8) row[k] = m1;
#pragma unroll 1
      for (int k = BW - nR + cl; k < BW; k += 8) row[k] = m1;
    }
  }
}

// One channel of a 4 x 2 strip: raw samples and their derivatives along x and y (two rows of two pixel pairs)
template <typename T, int R>
__device__ __forceinline__ void bwd_sample_strip(const T* p, int sh, f32x2 fx2, f32x2 fy2, f32x2 (&v)[2][2], f32x2 (&dx)[2][2], f32x2 (&dy)[2][2]) {
  constexpr int BW = BwdBox<T>::W;
  f32x2 h0[3], h1[3], d0[3], d1[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    f32x2 L0, L1, R0, R1;
    Row5<T>::template taps<R>(p + r * BW, sh, L0, L1, R0, R1);
    d0[r] = sub2(R0, L0); d1[r] = sub2(R1, L1);
    h0[r] = fma2(fx2, d0[r], L0); h1[r] = fma2(fx2, d1[r], L1);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    dy[r][0] = sub2(h0[r + 1], h0[r]); dy[r][1] = sub2(h1[r + 1], h1[r]);
    v[r][0] = fma2(fy2, dy[r][0], h0[r]); v[r][1] = fma2(fy2, dy[r][1], h1[r]);
    dx[r][0] = fma2(fy2, sub2(d0[r + 1], d0[r]), d0[r]); dx[r][1] = fma2(fy2, sub2(d1[r + 1], d1[r]), d1[r]);
  }
}

__device__ __forceinline__ float sa_to_float(float v) { return v; }
__device__ __forceinline__ float sa_to_float(__half v) { return __half2float(v); }
template <typename SA> __device__ __forceinline__ SA float_to_sa(float v);
template <> __device__ __forceinline__ float float_to_sa<float>(float v) { return v; }
template <> __device__ __forceinline__ __half float_to_sa<__half>(float v) { return __float2half_rn(v); }
__device__ __forceinline__ float t_to_float(float v) { return v; }
__device__ __forceinline__ float t_to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float t_to_float(__half v) { return __half2float(v); }
__device__ __forceinline__ float hsum(f32x2 v) { float a, b; upk(v, a, b); return a + b; }

// the left-tap pairs of a strip's two pixel pairs: gl = the record of the pixel left of the strip
template <typename T>
__device__ __forceinline__ void left_pairs(float gl, f32x2 C0, f32x2 C1, f32x2& L0, f32x2& L1) {
  if (Row5<T>::kStrided) { L0 = pk(gl, lo(C1)); L1 = C0; }             // C0 = (g0, g2), C1 = (g1, g3): lefts (gl, g1), (g0, g2)
  else { L0 = pk(gl, lo(C0)); L1 = pk(hi(C0), lo(C1)); }               // C0 = (g0, g1), C1 = (g2, g3): lefts (gl, g0), (g1, g2)
}

__device__ __forceinline__ void st_vec2(float* p, const float (&v)[2]) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }

// four adjacent texels of one row, element k written iff ok[k]: one double-width store where two neighbours share an aligned
// word of twice the element size (`odd`: the first texel's column is odd), single-element stores for the rest.  Rows start on even element offsets (W % 4 == 0).
// Predicated per element on purpose: a masked fast / slow split sent most warps of a border tile down both paths (+16 %).
template <typename T>
__device__ __forceinline__ void store4(T* o, const float (&v)[4], const bool (&ok)[4], bool odd) {
  {
    auto pair = [&](int k) {                                  // elements k, k + 1 share a word
      if (ok[k] && ok[k + 1]) { const float t[2] = {v[k], v[k + 1]}; st_vec2(o + k, t); }
      else { if (ok[k]) st(o + k, v[k]); if (ok[k + 1]) st(o + k + 1, v[k + 1]); }
    };
    if (odd) { if (ok[0]) st(o, v[0]); pair(1); if (ok[3]) st(o + 3, v[3]); }
    else { pair(0); pair(2); }
  }
}
template <typename T, bool kNeedX, bool kNeedTheta, bool kGlobalT>
__global__ void __launch_bounds__(kBThreads, MGR_STB_BLOCKS)
render_bwd_shift_tma(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap amap,
                     const __grid_constant__ CUtensorMap gmap, const __grid_constant__ CUtensorMap omap,
                     const float* __restrict__ theta, T* __restrict__ gx, long long gx_sb, long long gx_sl,
                     float* __restrict__ gtheta, Geometry g, const int* __restrict__ sample_all_shift,
                     const typename SavedAlpha<T>::type* __restrict__ sav, const T* __restrict__ gout, const T* __restrict__ outp,
                     float* __restrict__ tws) {
  using Box = BwdBox<T>;
  using SA = typename SavedAlpha<T>::type;
  constexpr int BW = Box::W;
  // kGlobalT (fp32 tensors, and 16-bit stacks with many layers): the alpha / grad_out / out boxes and the transmittances of a tile
  // would leave room for too few CTAs per SM, so
  // the pre-pass reads the tile's pixels straight from global memory (once per CTA) and parks T_l in the workspace (`tws`, the
  // record area of the general passes, 8 bytes per layer-pixel: a translation sample's [L][H*W] floats take the first half of ITS OWN
  // slice, so nothing collides with the general samples' records being written on the other stream); the layer loop reads T_l back
  // (L2 hits, issued before the layer's sampling).  Pixels shared by overlapping tiles get the same bits from each of them.
  static_assert(kGlobalT || sizeof(T) == 2, "fp32 tensors use the workspace for T_l");
  const int b = blockIdx.z;
  if (!sample_all_shift[b]) return;
  extern __shared__ __align__(128) unsigned char smem[];
  const BwdSmem lay = bwd_tma_layout<T>(g.L, kGlobalT);
  SA* atile = reinterpret_cast<SA*>(smem + lay.alpha);
  const T* gtile = reinterpret_cast<const T*>(smem + lay.gout);
  const T* otile = reinterpret_cast<const T*>(smem + lay.out);
  float* rec = reinterpret_cast<float*>(smem + lay.gout);                // [2 buffers][2: T a, dA][kBH][kBW] after the pre-pass
  float4* GPs = reinterpret_cast<float4*>(smem + lay.gp);               // [(comp * 2 + row)][thread]
  float* red = reinterpret_cast<float*>(smem + lay.red);                // [L][warp][8]
  ShiftPlan* splan = reinterpret_cast<ShiftPlan*>(smem + lay.plan);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + lay.bars);        // [kBStages]
  uint64_t* empty = full + kBStages;                                    // [kBStages]
  uint64_t* bar0 = empty + kBStages;
  const int tid = threadIdx.x;
  const int j0 = blockIdx.x * kBAncW - 1, i0 = blockIdx.y * kBAncH - 1;
  const int xp = j0 & ~(Box::kAlign - 1);                   // origin of the pixel-space boxes (16-byte rule)
  if (tid == kBConsumers) {                       // the producer lane: barriers, then the pixel-space boxes at once (they do not depend on theta)
#pragma unroll
    for (int s = 0; s < kBStages; ++s) { tma_mbar_init(&full[s], 1); tma_mbar_init(&empty[s], kBConsumers / 32); }
    tma_mbar_init(bar0, 1);
    tma_fence_barrier_init();
    if (!kGlobalT) {
      tma_mbar_expect_tx(bar0, (uint32_t)(Box::kPPlane * g.L * (int)sizeof(SA) + 2 * Box::kGBytes));
      tma_load_4d(atile, &amap, bar0, xp, i0, 0, b);
      tma_load_4d(smem + lay.gout, &gmap, bar0, xp, i0, 0, b);
      tma_load_4d(smem + lay.out, &omap, bar0, xp, i0, 0, b);
    }
  }
  for (int l = tid; l < g.L; l += kBThreads) splan[l] = make_shift_plan(theta + ((long long)b * g.L + l) * 6, g.H, g.W);
  if (kNeedTheta)
    for (int k = tid; k < g.L * (kBConsumers / 32) * 8; k += kBThreads) red[k] = 0.f;
  __syncthreads();

  if (tid >= kBConsumers) {                       // ---- producer warp ----
    if (tid == kBConsumers) {
      int it = 0;
      for (int l = 0; l < g.L; ++l) {
        const int x0 = j0 + splan[l].X, y0 = i0 + splan[l].Y;
        if (bwd_box_misses(x0, y0, g.W, g.H)) continue;
        const int s = it % kBStages;
        if (it >= kBStages) tma_mbar_wait(&empty[s], (uint32_t)((it / kBStages) - 1) & 1u);
        tma_mbar_expect_tx(&full[s], (uint32_t)Box::kXBytes);
        tma_load_5d(smem + lay.x + (size_t)Box::kXStage * s, &xmap, &full[s], x0 & ~(Box::kAlign - 1), y0, 0, l, b);
        ++it;
      }
    }
    return;
  }

  // ---- consumers ----
  const int tx = tid & 15, ty = tid >> 4, lane = tid & 31, wid = tid >> 5;
  const int jt = j0 + 4 * tx, it0 = i0 + 2 * ty;           // first pixel of the strip (may be -1 or past the image: records are 0 there)
  const int poff = (2 * ty) * BW + (j0 - xp) + 4 * tx;      // the strip inside a plane of the pixel-space boxes
  const float zs = g.m11 ? 0.5f : 1.f;
  const f32x2 zs2 = bc(zs), zb2 = bc(g.m11 ? 0.5f : 0.f);

  // ---- pre-pass: T_l in place of the alpha samples (fp32: in the workspace), (G_P, G_A) per pixel ----
  const int hw = g.H * g.W;
  bool livep[2][4];                                           // the strip's pixels that exist
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) livep[r][k] = (unsigned)(jt + k) < (unsigned)g.W && (unsigned)(it0 + r) < (unsigned)g.H;
  const int pix0 = it0 * g.W + jt;                            // the strip's first pixel inside a plane (only used where live)
  if (!kGlobalT) tma_mbar_wait(bar0, 0);
  {
    float A[2][4], Tc[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) { A[r][k] = 0.f; Tc[r][k] = 1.f; }
    if constexpr (kGlobalT) {
      const SA* sb_ = sav + (long long)b * g.L * hw + pix0;
      float* tb_ = tws + 2LL * b * g.L * hw + pix0;          // the first half of THIS sample's record area: [L][H*W] floats
      float an[2][4];                                         // next layer's alphas: loaded one layer ahead of their use
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) an[r][k] = livep[r][k] ? ld_alpha(sb_ + (long long)(g.L - 1) * hw + r * g.W + k) : 0.f;
      for (int l = g.L - 1; l >= 0; --l) {
        float ac[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            ac[r][k] = an[r][k];
            if (l > 0) an[r][k] = livep[r][k] ? ld_alpha(sb_ + (long long)(l - 1) * hw + r * g.W + k) : 0.f;
          }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (livep[r][k]) tb_[(long long)l * hw + r * g.W + k] = Tc[r][k];
            A[r][k] = fmaf(Tc[r][k], ac[r][k], A[r][k]);
            Tc[r][k] *= 1.f - ac[r][k];
          }
      }
    } else {
      for (int l = g.L - 1; l >= 0; --l) {
        SA* al = atile + l * Box::kPPlane + poff;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float a = sa_to_float(al[r * BW + k]);
            al[r * BW + k] = float_to_sa<SA>(Tc[r][k]);
            A[r][k] = fmaf(Tc[r][k], a, A[r][k]);
            Tc[r][k] *= 1.f - a;
          }
      }
    }
    const float gs = g.m11 ? 2.f : 1.f;                       // d out / d o
    const float is = g.m11 ? 0.5f : 1.f, ib = g.m11 ? 0.5f : 0.f;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float gpv[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = poff + r * BW + k;
        gpv[0][k] = gpv[1][k] = gpv[2][k] = gpv[3][k] = 0.f;
        if (A[r][k] != 0.f) {                                 // (a pixel outside the image has A == 0)
          const float inv = 1.f / A[r][k];
          float gr_[4], or_[3];
          if constexpr (kGlobalT) {
            const T* gq = gout + (long long)b * 4 * hw + pix0 + r * g.W + k;
            const T* oq = outp + (long long)b * 4 * hw + pix0 + r * g.W + k;
#pragma unroll
            for (int c = 0; c < 4; ++c) gr_[c] = ld(gq + c * hw);
#pragma unroll
            for (int c = 0; c < 3; ++c) or_[c] = ld(oq + c * hw);
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) gr_[c] = t_to_float(gtile[e + c * Box::kPPlane]);
#pragma unroll
            for (int c = 0; c < 3; ++c) or_[c] = t_to_float(otile[e + c * Box::kPPlane]);
          }
          const float g0 = gs * gr_[0], g1 = gs * gr_[1], g2 = gs * gr_[2], g3 = gs * gr_[3];
          const float o0 = fmaf(or_[0], is, ib), o1 = fmaf(or_[1], is, ib), o2 = fmaf(or_[2], is, ib);
          gpv[0][k] = g0 * inv; gpv[1][k] = g1 * inv; gpv[2][k] = g2 * inv;
          gpv[3][k] = g3 - fmaf(g2, o2, fmaf(g1, o1, g0 * o0)) * inv;
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        f32x2 q0, q1;
        strip_pack<T>(gpv[c], q0, q1);
        float4 v;
        upk(q0, v.x, v.y); upk(q1, v.z, v.w);
        GPs[(c * 2 + r) * kBConsumers + tid] = v;
      }
    }
  }
  named_barrier(1, kBConsumers);                              // grad_out / out boxes are dead: their bytes become the records

  // ownership masks for the theta sums (the tile's first column / row belong to the neighbours) and the strip's coordinates
  f32x2 msk[2][2], xq[2];
  float yr[2];
  if (kNeedTheta) {
    const float inv_w = 1.f / (float)g.W, inv_h = 1.f / (float)g.H;   // normalised pixel centres (2k + 1) / n - 1, to an ulp
    float mc[4], xs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { mc[k] = (4 * tx + k >= 1) ? 1.f : 0.f; xs[k] = fmaf((float)(2 * (jt + k) + 1), inv_w, -1.f); }
    f32x2 m0, m1;
    strip_pack<T>(mc, m0, m1);
    strip_pack<T>(xs, xq[0], xq[1]);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float mr = (2 * ty + r >= 1) ? 1.f : 0.f;
      msk[r][0] = mul2(m0, bc(mr)); msk[r][1] = mul2(m1, bc(mr));
      yr[r] = fmaf((float)(2 * (it0 + r) + 1), inv_h, -1.f);
    }
  }

  f32x2 q[2][2];
#pragma unroll
  for (int r = 0; r < 2; ++r) q[r][0] = q[r][1] = bc(0.f);
  const int toff = (2 * ty) * BW + 4 * tx;
  T* gxb = gx + (long long)b * gx_sb;
  unsigned cmask = 0, rmask = 0;                              // anchors of the strip this CTA owns and that exist (a <= W, b <= H)
#pragma unroll
  for (int k = 0; k < 4; ++k) cmask |= (4 * tx + k >= 1 && jt + k <= g.W) ? (1u << k) : 0u;
#pragma unroll
  for (int r = 0; r < 2; ++r) rmask |= (2 * ty + r >= 1 && it0 + r <= g.H) ? (1u << r) : 0u;
  const int zx0 = blockIdx.x * kBW, zy0 = blockIdx.y * kBH;   // the aligned block this CTA zero-fills
  const bool zf_block = zx0 < g.W && zy0 < g.H;
  const int roff = (2 * ty) * kBW + 4 * tx;                    // this strip inside a record plane

  int it = 0;
  for (int l = 0; l < g.L; ++l) {
    const ShiftPlan sp = splan[l];
    const int x0 = j0 + sp.X, y0 = i0 + sp.Y;
    T* gxl = gxb + (long long)l * gx_sl;
    // Texels no pixel reaches (columns outside [X, X + W], rows outside [Y, Y + H]) are zeroed by the CTA that owns the
    // ALIGNED 64 x 16 block with its grid coordinates (the anchor tiling has at least as many tiles per axis): 8-byte
    // stores, two chunk rows per thread, nothing to do for the blocks a layer covers entirely (CTA-uniform test).
    if (kNeedX && zf_block && (zx0 < sp.X || zx0 + kBW - 1 > sp.X + g.W || zy0 < sp.Y || zy0 + kBH - 1 > sp.Y + g.H)) {
      const int cx = zx0 + 4 * tx;
      if (cx < g.W) {
        unsigned cm = 0;                                      // bit k: texel column cx + k is out of every pixel's reach
#pragma unroll
        for (int k = 0; k < 4; ++k) cm |= (cx + k < sp.X || cx + k > sp.X + g.W) ? (1u << k) : 0u;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int yy = zy0 + ty + (kBH / 2) * rr;
          if (yy >= g.H) continue;
          const unsigned m = (yy < sp.Y || yy > sp.Y + g.H) ? 0xfu : cm;
          if (m == 0u) continue;
          T* o = gxl + yy * g.W + cx;
          if (m == 0xfu) {
            const float z4[4] = {0.f, 0.f, 0.f, 0.f};
            st_vec4<T>(o, z4); st_vec4<T>(o + hw, z4); st_vec4<T>(o + 2 * hw, z4); st_vec4<T>(o + 3 * hw, z4);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (m & (1u << k)) { st(o + k, 0.f); st(o + hw + k, 0.f); st(o + 2 * hw + k, 0.f); st(o + 3 * hw + k, 0.f); }
          }
        }
      }
    }
    if (bwd_box_misses(x0, y0, g.W, g.H)) continue;           // transparent layer for this tile
    const int s = it % kBStages;
    T* stage = reinterpret_cast<T*>(smem + lay.x + (size_t)Box::kXStage * s);
    const int xa = x0 & ~(Box::kAlign - 1), dx0 = x0 - xa;
    tma_mbar_wait(&full[s], (uint32_t)(it / kBStages) & 1u);
    if (g.m11 && (xa < 0 || xa + BW > g.W || y0 < 0 || y0 + Box::kXRows > g.H)) {       // CTA-uniform
      bwd_patch_oob<T>(stage, xa, y0, g.W, g.H, tid);
      tma_fence_proxy_async();
      named_barrier(1, kBConsumers);
    }
    const f32x2 fx2 = bc(sp.fx), fy2 = bc(sp.fy);
    const T* p = stage + toff + (sizeof(T) == 2 ? (dx0 & ~1) : 0);   // 16-bit: the even element at or left of e0 (Row5)
    const int sh = sizeof(T) == 2 ? (dx0 & 1) * 16 : dx0;           // 16-bit: funnel shift; fp32: offset inside the aligned chunk
    const SA* tl = atile + l * Box::kPPlane + poff;
    float tg[2][4];                                           // fp32: T_l from the workspace, in flight while alpha is sampled
    if constexpr (kGlobalT) {
      const float* tq = tws + (2LL * b * g.L + l) * hw + pix0;
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) tg[r][k] = livep[r][k] ? __ldcg(tq + r * g.W + k) : 1.f;
    }
    float* recw = rec + (it & 1) * 2 * kBH * kBW;               // this layer's record buffer

    float th6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto body = [&](auto rtag) {
      constexpr int R = decltype(rtag)::value;
      f32x2 a[2][2], dxa[2][2], dya[2][2], ta[2][2], tt[2][2];
      {
        f32x2 v[2][2];
        bwd_sample_strip<T, R>(p + 3 * Box::kXPlane, sh, fx2, fy2, v, dxa, dya);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          float tv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) tv[k] = kGlobalT ? tg[r][k] : sa_to_float(tl[r * BW + k]);
          strip_pack<T>(tv, tt[r][0], tt[r][1]);
          a[r][0] = fma2(v[r][0], zs2, zb2); a[r][1] = fma2(v[r][1], zs2, zb2);
          ta[r][0] = mul2(tt[r][0], a[r][0]); ta[r][1] = mul2(tt[r][1], a[r][1]);
        }
      }
      // colours: u = G_P . c + G_A, and the G_P-weighted derivatives; records T a G_P per channel as they come
      f32x2 u[2][2], Dx[2][2], Dy[2][2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float4 ga4 = GPs[(3 * 2 + r) * kBConsumers + tid];
        u[r][0] = pk(ga4.x, ga4.y); u[r][1] = pk(ga4.z, ga4.w);
        Dx[r][0] = Dx[r][1] = Dy[r][0] = Dy[r][1] = bc(0.f);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        f32x2 v[2][2], dxc[2][2], dyc[2][2];
        bwd_sample_strip<T, R>(p + c * Box::kXPlane, sh, fx2, fy2, v, dxc, dyc);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float4 gp4 = GPs[(c * 2 + r) * kBConsumers + tid];
          const f32x2 gp0 = pk(gp4.x, gp4.y), gp1 = pk(gp4.z, gp4.w);
          u[r][0] = fma2(gp0, fma2(v[r][0], zs2, zb2), u[r][0]);
          u[r][1] = fma2(gp1, fma2(v[r][1], zs2, zb2), u[r][1]);
          if (kNeedTheta) {
            Dx[r][0] = fma2(gp0, dxc[r][0], Dx[r][0]); Dx[r][1] = fma2(gp1, dxc[r][1], Dx[r][1]);
            Dy[r][0] = fma2(gp0, dyc[r][0], Dy[r][0]); Dy[r][1] = fma2(gp1, dyc[r][1], Dy[r][1]);
          }
        }
      }
      // dA_l = T_l (u_l - q_l), q_{l+1} = q_l + a_l (u_l - q_l); the layer's theta sums
      f32x2 sx[2], sy[2];                                     // per row: sum over the strip's pixels of dL/dix, dL/diy (masked)
      f32x2 cx[2], cy[2];                                     // per pixel pair: the same summed over the two rows
      cx[0] = cx[1] = cy[0] = cy[1] = bc(0.f);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const f32x2 d0 = sub2(u[r][0], q[r][0]), d1 = sub2(u[r][1], q[r][1]);
        const f32x2 ga0 = mul2(tt[r][0], d0), ga1 = mul2(tt[r][1], d1);
        q[r][0] = fma2(a[r][0], d0, q[r][0]); q[r][1] = fma2(a[r][1], d1, q[r][1]);
        if (kNeedX) {
          float4 o;
          upk(ta[r][0], o.x, o.y); upk(ta[r][1], o.z, o.w);
          *reinterpret_cast<float4*>(recw + roff + r * kBW) = o;
          upk(ga0, o.x, o.y); upk(ga1, o.z, o.w);
          *reinterpret_cast<float4*>(recw + kBH * kBW + roff + r * kBW) = o;
        }
        if (kNeedTheta) {
          const f32x2 ix0 = mul2(msk[r][0], fma2(ta[r][0], Dx[r][0], mul2(ga0, dxa[r][0])));
          const f32x2 ix1 = mul2(msk[r][1], fma2(ta[r][1], Dx[r][1], mul2(ga1, dxa[r][1])));
          const f32x2 iy0 = mul2(msk[r][0], fma2(ta[r][0], Dy[r][0], mul2(ga0, dya[r][0])));
          const f32x2 iy1 = mul2(msk[r][1], fma2(ta[r][1], Dy[r][1], mul2(ga1, dya[r][1])));
          sx[r] = add2(ix0, ix1); sy[r] = add2(iy0, iy1);
          cx[0] = add2(cx[0], ix0); cx[1] = add2(cx[1], ix1);
          cy[0] = add2(cy[0], iy0); cy[1] = add2(cy[1], iy1);
        }
      }
      if (kNeedTheta) {
        const float sx0 = hsum(sx[0]), sx1 = hsum(sx[1]), sy0 = hsum(sy[0]), sy1 = hsum(sy[1]);
        th6[0] = hsum(fma2(cx[0], xq[0], mul2(cx[1], xq[1])));   // sum dix * x_j
        th6[1] = fmaf(sx0, yr[0], sx1 * yr[1]);                  // sum dix * y_i
        th6[2] = sx0 + sx1;
        th6[3] = hsum(fma2(cy[0], xq[0], mul2(cy[1], xq[1])));
        th6[4] = fmaf(sy0, yr[0], sy1 * yr[1]);
        th6[5] = sy0 + sy1;
      }
    };
    body(std::integral_constant<int, -1>{});               // one body for every alignment of the box (Row5)
    __syncwarp();
    if (lane == 0) tma_mbar_arrive(&empty[s]);                // the x box is free for the copy after next
    ++it;
    if (kNeedTheta) {
      const float sum = warp_sum6(th6, lane);
      const int qi = warp_sum6_index(lane);
      if ((lane & 3) == 0 && qi < 6) red[(l * (kBConsumers / 32) + wid) * 8 + qi] = sum;
    }
    if (kNeedX) {
      named_barrier(1, kBConsumers);                          // the tile's records are in shared memory
      // anchors (a, b) = the strip's pixels (not the tile's first column / row): texel (a + X, b + Y) gets
      //   (1-fy)[(1-fx) g(a, b) + fx g(a-1, b)] + fy[(1-fx) g(a, b-1) + fx g(a-1, b-1)]     (x zs: d z / d raw)
      const f32x2 wx0 = bc((1.f - sp.fx) * zs), wx1 = bc(sp.fx * zs), wy0 = bc(1.f - sp.fy), wy1 = bc(sp.fy);
      const int rtop = ty > 0 ? 2 * ty - 1 : 0;               // row above the strip (unused garbage for ty == 0)
      const int tl_ = tid > 0 ? tid - 1 : 0, tu = tid >= 16 ? tid - 16 : 0, tul = tid >= 17 ? tid - 17 : 0;   // left / upper / upper-left strips
      bool colp[4], rowp[2];
#pragma unroll
      for (int k = 0; k < 4; ++k) colp[k] = ((cmask >> k) & 1u) && (unsigned)(jt + k + sp.X) < (unsigned)g.W;
#pragma unroll
      for (int r = 0; r < 2; ++r) rowp[r] = ((rmask >> r) & 1u) && (unsigned)(it0 + r + sp.Y) < (unsigned)g.H;
      const int tbase = (it0 + sp.Y) * g.W + jt + sp.X;       // texel of the strip's first anchor (used only where valid)
      const bool odd = (jt + sp.X) & 1;                       // CTA-uniform: which neighbours share a 32-bit word of grad_x
      // the records of the strip's rows -1, 0, 1 and of the pixel left of each: (T a) and dA
      f32x2 taC[3][2], gaC[3][2];
      float taL[3], gaL[3];
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
        const int row = rr == 0 ? rtop : 2 * ty + rr - 1;
        const float4 t4 = *reinterpret_cast<const float4*>(recw + row * kBW + 4 * tx);
        const float4 g4 = *reinterpret_cast<const float4*>(recw + kBH * kBW + row * kBW + 4 * tx);
        taC[rr][0] = pk(t4.x, t4.y); taC[rr][1] = pk(t4.z, t4.w);
        gaC[rr][0] = pk(g4.x, g4.y); gaC[rr][1] = pk(g4.z, g4.w);
        const int cl = tx > 0 ? 4 * tx - 1 : 0;               // (unused garbage for tx == 0)
        taL[rr] = recw[row * kBW + cl];
        gaL[rr] = recw[kBH * kBW + row * kBW + cl];
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        f32x2 h[3][2];                                        // rows -1, 0, 1 of the strip, x-stencil applied
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
          f32x2 C0, C1;
          float gl;
          if (c < 3) {                                        // colour: record = G_P (per pixel, layer-invariant) x (T a)
            const float4 gp4 = GPs[(c * 2 + (rr == 0 ? 1 : rr - 1)) * kBConsumers + (rr == 0 ? tu : tid)];
            const float gpl = reinterpret_cast<const float*>(&GPs[(c * 2 + (rr == 0 ? 1 : rr - 1)) * kBConsumers + (rr == 0 ? tul : tl_)])[3];
            C0 = mul2(pk(gp4.x, gp4.y), taC[rr][0]); C1 = mul2(pk(gp4.z, gp4.w), taC[rr][1]);
            gl = gpl * taL[rr];
          } else { C0 = gaC[rr][0]; C1 = gaC[rr][1]; gl = gaL[rr]; }
          f32x2 L0, L1;
          left_pairs<T>(gl, C0, C1, L0, L1);
          h[rr][0] = fma2(wx1, L0, mul2(wx0, C0));
          h[rr][1] = fma2(wx1, L1, mul2(wx0, C1));
        }
        T* oc = gxl + c * hw + tbase;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          float v4[4];
          strip_unpack<T>(fma2(wy1, h[r][0], mul2(wy0, h[r + 1][0])), fma2(wy1, h[r][1], mul2(wy0, h[r + 1][1])), v4);
          if (rowp[r]) store4<T>(oc + r * g.W, v4, colp, odd);
        }
      }
      // no second barrier: the next layer writes the OTHER record buffer, and nobody writes this one again before the
      // barrier after that, which every thread reaches only when it is done reading here
    }
  }

  if (kNeedTheta) {
    named_barrier(1, kBConsumers);
    const float hW = 0.5f * (float)g.W * zs, hH = 0.5f * (float)g.H * zs;
    for (int k = tid; k < g.L * 6; k += kBConsumers) {
      const int l = k / 6, qi = k - 6 * l;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kBConsumers / 32; ++w) v += red[(l * (kBConsumers / 32) + w) * 8 + qi];
      atomicAdd(gtheta + ((long long)b * g.L + l) * 6 + qi, v * (qi < 3 ? hW : hH));
    }
  }
}

// tensor maps of the backward: x {W,H,4,L,B}, saved alpha {W,H,L,B}, grad_out / out {W,H,4,B} (contiguous tensors)
template <typename T>
inline bool bwd_tma_maps(CUtensorMap* xmap, CUtensorMap* amap, CUtensorMap* gmap, CUtensorMap* omap, const void* x, const void* sav,
                         const void* gout, const void* out, const Geometry& g, bool global_t) {
  using Box = BwdBox<T>;
  if (!shift_tma_x_map<T>(xmap, x, g, Box::W, Box::kXRows)) return false;
  if (global_t) { *amap = *gmap = *omap = *xmap; return true; }           // the pixel-space tensors are read directly (kGlobalT)
  const long long hw = (long long)g.H * g.W;
  const long long ad[4] = {g.W, g.H, g.L, g.B}, as[4] = {1, g.W, hw, hw * g.L};
  const int ab[4] = {Box::W, kBH, g.L, 1};
  if (!tma_make_map(amap, sav, (int)sizeof(typename Box::SA), 4, ad, as, ab)) return false;
  const long long gd[4] = {g.W, g.H, 4, g.B}, gst[4] = {1, g.W, hw, 4 * hw};
  const int gb[4] = {Box::W, kBH, 4, 1};
  return tma_make_map(gmap, gout, (int)sizeof(T), 4, gd, gst, gb) && tma_make_map(omap, out, (int)sizeof(T), 4, gd, gst, gb);
}

}  // namespace mgr
