// Building blocks of the tiled render kernels (forward and backward pass 1):
//   * packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2 through PTX .f32x2),
//   * the per-(tile, layer) plan: placement split + source footprint of a 32x32 output tile,
//   * staging of that footprint from the planar layer into shared memory as channel-interleaved
//     RGBA texels (vector loads, one half-warp per footprint row),
//   * the bilinear sampler on staged texels (value and, for the backward, d/dix and d/diy).
#pragma once
#include <type_traits>

#include "mgr_common.cuh"

namespace mgr {

// ---- packed fp32x2 (sm_100+: one instruction, two fp32 lanes) ---------------------------------
using f32x2 = unsigned long long;
__device__ __forceinline__ f32x2 pk(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(f32x2 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ float lo(f32x2 v) { float a, b; upk(v, a, b); return a; }
__device__ __forceinline__ float hi(f32x2 v) { float a, b; upk(v, a, b); return b; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 bc(float a) { return pk(a, a); }

// ---- tile geometry ------------------------------------------------------------------------------
constexpr int kTW = 32, kTH = 32;           // output tile
constexpr int kTiledThreads = 256;          // 8 warps; thread (tx, ty) owns pixels (tx, ty + kRowStep * k), k < kPx
constexpr int kPx = kTW * kTH / kTiledThreads;
constexpr int kRowStep = kTiledThreads / kTW;
__host__ __device__ constexpr size_t align16(size_t n) { return (n + 15) & ~(size_t)15; }
constexpr int kCapTexels = 2816;            // staging capacity (e.g. a 52 x 54 footprint)

enum { kSkip = 0, kStaged = 1, kDirect = 2 };

struct LayerPlan {
  TileAffine aff;   // placement relative to the tile CENTRE with a global integer origin (direct path)
  float lrx, lry;   // staged path: the tile centre in footprint-local coordinates
  int x_lo, y_lo;   // footprint origin in the source image (x_lo multiple of the staging vector width)
  int bw, bh;       // footprint size in texels (bw multiple of the vector width) == shared pitch / rows
  int mode;
  int pitch;        // warp-specialised kernels: row pitch of the staged footprint in texels (>= bw)
  int dX, dY;       // staged path: tap column / row inside the footprint = floor(tile-centre-relative coordinate) + (dX, dY)
  int pad2_[2];
};

// One layer's pixels for the CTA's sample: SrcLayer with the batch offset applied and byte strides (they fit 32 bits,
// host-checked).  Built from the kernel-parameter descriptors, so every field is CTA-uniform (uniform registers).
struct SrcView {
  const char* base;
  unsigned plane, rowbytes;
  int left, top, w, h;
};
// kRagged = false: the canvas layout, straight from the Geometry (compile-time zero offsets: the code the canvas path
// always had); kRagged = true: from the per-layer descriptors.
template <typename T, bool kRagged>
__device__ __forceinline__ SrcView layer_view(const T* x, const Geometry& g, const SrcLayers& src, int b, int l) {
  SrcView v;
  if (kRagged) {
    const SrcLayer& s = src.s[l];
    v.base = reinterpret_cast<const char*>(reinterpret_cast<const T*>(s.ptr) + (long long)b * s.sb);
    v.plane = (unsigned)s.sc * (unsigned)sizeof(T);
    v.rowbytes = (unsigned)s.sh * (unsigned)sizeof(T);
    v.left = s.left; v.top = s.top; v.w = s.w; v.h = s.h;
  } else {
    v.base = reinterpret_cast<const char*>(x + (long long)b * g.sb + (long long)l * g.sl);
    v.plane = (unsigned)g.sc * (unsigned)sizeof(T);
    v.rowbytes = (unsigned)g.sh * (unsigned)sizeof(T);
    v.left = 0; v.top = 0; v.w = g.W; v.h = g.H;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ SrcView src_view(const SrcLayer& s, int b) {
  SrcView v;
  v.base = reinterpret_cast<const char*>(reinterpret_cast<const T*>(s.ptr) + (long long)b * s.sb);
  v.plane = (unsigned)s.sc * (unsigned)sizeof(T);
  v.rowbytes = (unsigned)s.sh * (unsigned)sizeof(T);
  v.left = s.left; v.top = s.top; v.w = s.w; v.h = s.h;
  return v;
}

// Plan one layer for the tile whose top-left pixel is (j0, i0).  Coordinates are carried relative
// to the tile centre (|dj|, |di| <= 16) so that the fp32 per-pixel arithmetic stays accurate to
// ~2e-6 px (the reference's own fp32 grid carries ~3e-5 px at 256x256).
struct SrcRect { int left, top, w, h; };
template <bool kRagged>
__device__ __forceinline__ SrcRect layer_rect(const Geometry& g, const SrcLayers& src, int l) {
  if (kRagged) return SrcRect{src.s[l].left, src.s[l].top, src.s[l].w, src.s[l].h};
  return SrcRect{0, 0, g.W, g.H};
}

__device__ __forceinline__ LayerPlan plan_layer(const float* __restrict__ th, int H, int W, int j0, int i0, int vec,
                                                const SrcRect src) {
  LayerPlan p;
  p.aff = make_tile_affine(th, H, W, j0 + kTW / 2, i0 + kTH / 2);
  const TileAffine& t = p.aff;
  const float lo_ = -(float)(kTW / 2), hi_ = (float)(kTW / 2 - 1);   // dj in [-kTW/2, kTW/2 - 1]
  const float lov = -(float)(kTH / 2), hiv = (float)(kTH / 2 - 1);   // di in [-kTH/2, kTH/2 - 1]
  const float eps = 2e-3f;   // fp32 rounding of per-pixel coordinates is ~1e-5 px; stay well clear
  const float xmin = t.rx + fminf(t.a00 * lo_, t.a00 * hi_) + fminf(t.a01 * lov, t.a01 * hiv) - eps;
  const float xmax = t.rx + fmaxf(t.a00 * lo_, t.a00 * hi_) + fmaxf(t.a01 * lov, t.a01 * hiv) + eps;
  const float ymin = t.ry + fminf(t.a10 * lo_, t.a10 * hi_) + fminf(t.a11 * lov, t.a11 * hiv) - eps;
  const float ymax = t.ry + fmaxf(t.a10 * lo_, t.a10 * hi_) + fmaxf(t.a11 * lov, t.a11 * hiv) + eps;
  p.x_lo = p.y_lo = p.bw = p.bh = 0;
  p.lrx = p.lry = 0.f;
  p.pitch = 0; p.dX = p.dY = 0; p.pad2_[0] = p.pad2_[1] = 0;
  p.mode = kDirect;
  // NaN / huge placements take the bounds-checked direct path
  // (the direct path addresses the layer's own pixels: origin moved to the rectangle's corner)
  const int X0 = t.X0, Y0 = t.Y0;
  p.aff.X0 = X0 - src.left; p.aff.Y0 = Y0 - src.top;
  if (!(fabsf(xmin) < 1.0e6f && fabsf(xmax) < 1.0e6f && fabsf(ymin) < 1.0e6f && fabsf(ymax) < 1.0e6f)) return p;
  if (abs(X0) > (1 << 28) || abs(Y0) > (1 << 28)) return p;
  int x_lo = X0 + (int)floorf(xmin), x_hi = X0 + (int)floorf(xmax) + 1;   // inclusive tap columns
  int y_lo = Y0 + (int)floorf(ymin), y_hi = Y0 + (int)floorf(ymax) + 1;
  // the taps miss the layer's rectangle: a fully transparent layer for this tile
  if (x_hi < src.left || x_lo >= src.left + src.w || y_hi < src.top || y_lo >= src.top + src.h) { p.mode = kSkip; return p; }
  // staged or direct: decided on the footprint as the WIDEST staging vector (8 texels) would lay it out, so that the
  // decision -- and with it the arithmetic that produces every bit of the result -- does not depend on the alignment
  // the tensor happens to allow (a ragged stack and its padded canvas must agree bit for bit)
  const int bw8 = (x_hi - (x_lo & ~7) + 8) & ~7;
  x_lo &= ~(vec - 1);                                   // vec = staging vector width in texels (4 or 8)
  const int bw = (x_hi - x_lo + vec) & ~(vec - 1), bh = y_hi - y_lo + 1;
  p.x_lo = x_lo; p.y_lo = y_lo; p.bw = bw; p.bh = bh; p.pitch = bw;
  p.lrx = t.rx + (float)(X0 - x_lo);
  p.lry = t.ry + (float)(Y0 - y_lo);
  p.dX = X0 - x_lo; p.dY = Y0 - y_lo;
  p.mode = ((long long)bw8 * bh <= kCapTexels) ? kStaged : kDirect;
  return p;
}

// ---- storage-dtype traits: one interleaved RGBA texel in shared memory ---------------------------
template <typename T> struct Texel;

template <> struct Texel<float> {
  using Vec = float4;
  __device__ static __forceinline__ void unpack(const float4& v, f32x2& rg, f32x2& ba) { rg = pk(v.x, v.y); ba = pk(v.z, v.w); }
};

template <> struct Texel<__nv_bfloat16> {
  using Vec = uint2;                                    // r | g << 16,  b | a << 16
  __device__ static __forceinline__ void unpack(const uint2& v, f32x2& rg, f32x2& ba) {
    rg = pk(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u));
    ba = pk(__uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
  }
  __device__ static __forceinline__ uint32_t oob2(bool m11) { return m11 ? 0xBF80BF80u : 0u; }   // bf16 -1.0 twice
};

template <> struct Texel<__half> {
  using Vec = uint2;
  __device__ static __forceinline__ void unpack(const uint2& v, f32x2& rg, f32x2& ba) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    rg = pk(a.x, a.y); ba = pk(b.x, b.y);
  }
  __device__ static __forceinline__ uint32_t oob2(bool m11) { return m11 ? 0xBC00BC00u : 0u; }   // fp16 -1.0 twice
};

// ---- staging: planar global rows -> interleaved shared texels -------------------------------------
// A half-warp cooperates on one footprint row, each lane moving four texels per channel (64-bit
// loads for 16-bit storage, 128-bit for fp32); kRowsInFlight rows are loaded before the first
// shared store so a thread keeps 4 * kRowsInFlight vector loads in flight (staging is latency-bound).
// Texels outside the image are written as "transparent black" in the raw range (-1 in m11 mode,
// 0 in 01 mode): that IS padding_mode='zeros', and it keeps bounds checks out of the sampling loop.
// Host-checked requirements (tiled_ok): W and every stride are multiples of 4 and the base pointer
// is aligned to the vector, so a lane's four texels are all inside or all outside the image; one
// layer spans < 2 GiB so byte offsets from the (uniform) layer base fit 32 bits.
constexpr int kStageVec = 4;            // texels per lane and channel
constexpr int kStageLanes = 16;         // lanes per footprint row
constexpr int kStageRows = kTiledThreads / kStageLanes;
template <typename T> struct StageCfg { static constexpr int kRowsInFlight = 3; };
template <> struct StageCfg<float> { static constexpr int kRowsInFlight = 2; };

__device__ __forceinline__ void interleave_store(float4* dst, const float4& r, const float4& g, const float4& b, const float4& a) {
  dst[0] = make_float4(r.x, g.x, b.x, a.x);
  dst[1] = make_float4(r.y, g.y, b.y, a.y);
  dst[2] = make_float4(r.z, g.z, b.z, a.z);
  dst[3] = make_float4(r.w, g.w, b.w, a.w);
}
// 16-bit: each source word holds two horizontally adjacent texels of one channel
__device__ __forceinline__ uint4 interleave2(uint32_t r, uint32_t g, uint32_t b, uint32_t a) {
  uint4 t;
  t.x = __byte_perm(r, g, 0x5410); t.y = __byte_perm(b, a, 0x5410);   // texel 0: r|g<<16, b|a<<16
  t.z = __byte_perm(r, g, 0x7632); t.w = __byte_perm(b, a, 0x7632);   // texel 1
  return t;
}
__device__ __forceinline__ void interleave_store(uint2* dst, const uint2& r, const uint2& g, const uint2& b, const uint2& a) {
  uint4* d = reinterpret_cast<uint4*>(dst);
  d[0] = interleave2(r.x, g.x, b.x, a.x);
  d[1] = interleave2(r.y, g.y, b.y, a.y);
}
__device__ __forceinline__ void fill_store(float4* dst, float o) {
  const float4 v = make_float4(o, o, o, o);
  dst[0] = v; dst[1] = v; dst[2] = v; dst[3] = v;
}
__device__ __forceinline__ void fill_store(uint2* dst, uint32_t o) {
  uint4* d = reinterpret_cast<uint4*>(dst);
  const uint4 v = make_uint4(o, o, o, o);
  d[0] = v; d[1] = v;
}

// kBarrier: the caller's "previous readers are done with buf" barrier is taken INSIDE, after the loads of the first
// block of rows have been issued and before the first shared store -- warps that reach the barrier early wait with
// their global loads already in flight (one barrier per layer then hides part of the staging latency for free).
template <typename T, bool kBarrier = false>
__device__ __forceinline__ void stage_footprint(bool m11, const SrcView& sv, const LayerPlan& p,
                                                typename Texel<T>::Vec* __restrict__ buf, int tid) {
  using Ld = typename std::conditional<sizeof(T) == 4, float4, uint2>::type;     // four texels of one channel
  using Vec = typename Texel<T>::Vec;
  constexpr int kIn = StageCfg<T>::kRowsInFlight;
  const int rsub = tid / kStageLanes, q = tid % kStageLanes;
  const int nv = p.bw / kStageVec;                            // vectors per footprint row
  typename std::conditional<sizeof(T) == 4, float, uint32_t>::type ob;
  if constexpr (sizeof(T) == 4) ob = m11 ? -1.f : 0.f; else ob = Texel<T>::oob2(m11);
  const char* base = sv.base;
  const unsigned plane = sv.plane, rowbytes = sv.rowbytes;
  const int dstep = kStageRows * p.bw;
  Ld R[kIn], G[kIn], Bl[kIn], A[kIn];
  bool inside[kIn];
  // rows r0, r0 + kStageRows, ... of vector column cv: issue the loads
  auto load_block = [&](int cv, int r0) {
    const int x = p.x_lo + kStageVec * cv - sv.left;           // column inside the layer's rectangle
    const bool xin = cv < nv && (unsigned)x < (unsigned)sv.w;
#pragma unroll
    for (int it = 0; it < kIn; ++it) {
      const int r = r0 + it * kStageRows;
      const int y = p.y_lo + r - sv.top;
      inside[it] = xin && r < p.bh && (unsigned)y < (unsigned)sv.h;
#ifdef MGR_EXPERIMENT_NO_STAGE_LOADS
      inside[it] = false;
#endif
      if (inside[it]) {
        const unsigned off = (unsigned)y * rowbytes + (unsigned)x * (unsigned)sizeof(T);
        R[it] = __ldg(reinterpret_cast<const Ld*>(base + off));
        G[it] = __ldg(reinterpret_cast<const Ld*>(base + (off + plane)));
        Bl[it] = __ldg(reinterpret_cast<const Ld*>(base + (off + 2 * plane)));
        A[it] = __ldg(reinterpret_cast<const Ld*>(base + (off + 3 * plane)));
      }
    }
  };
  // ... and interleave them into shared memory (texels outside the layer: the padding value)
  auto store_block = [&](int cv, int r0) {
    Vec* dst = buf + r0 * p.bw + kStageVec * cv;
#pragma unroll
    for (int it = 0; it < kIn; ++it) {
      if (inside[it]) interleave_store(dst, R[it], G[it], Bl[it], A[it]);
      else if (cv < nv && r0 + it * kStageRows < p.bh) fill_store(dst, ob);
      dst += dstep;
    }
  };
  int first_r0 = rsub;
  if (kBarrier) {                                             // peeled first block: every thread takes the barrier
    load_block(q, rsub);
    __syncthreads();
    store_block(q, rsub);
    first_r0 = rsub + kStageRows * kIn;
  }
  for (int c0 = 0; c0 < nv; c0 += kStageLanes) {              // one pass unless the footprint is > 64 texels wide
    for (int r0 = (c0 == 0 ? first_r0 : rsub); r0 < p.bh; r0 += kStageRows * kIn) {
      load_block(c0 + q, r0);
      store_block(c0 + q, r0);
    }
  }
}

// ---- bilinear sampling of a staged footprint ---------------------------------------------------
// Lerp form on raw storage values: top = v00 + fx (v01 - v00), bot likewise, s = top + fy (bot - top).
// A stack of identical taps returns that value exactly, so fully transparent texels (raw -1 in m11
// mode) give z == 0 exactly after the range shift -- as in the reference, where (x + 1) is exactly
// 0 before the weights are applied (fukuwarai/networks.py:253).
struct Sample {
  f32x2 rg, ba;         // interpolated raw value (r, g), (b, a)
};
struct SampleGrad {
  f32x2 rg, ba;
  f32x2 dx_rg, dx_ba;   // d value / d ix
  f32x2 dy_rg, dy_ba;   // d value / d iy
};

template <typename T>
__device__ __forceinline__ Sample sample_staged(const typename Texel<T>::Vec* __restrict__ q, int pitch, float fx, float fy) {
  f32x2 a_rg, a_ba, b_rg, b_ba, c_rg, c_ba, d_rg, d_ba;
  Texel<T>::unpack(q[0], a_rg, a_ba);
  Texel<T>::unpack(q[1], b_rg, b_ba);
  Texel<T>::unpack(q[pitch], c_rg, c_ba);
  Texel<T>::unpack(q[pitch + 1], d_rg, d_ba);
  const f32x2 fx2 = bc(fx), fy2 = bc(fy);
  const f32x2 top_rg = fma2(fx2, sub2(b_rg, a_rg), a_rg), top_ba = fma2(fx2, sub2(b_ba, a_ba), a_ba);
  const f32x2 bot_rg = fma2(fx2, sub2(d_rg, c_rg), c_rg), bot_ba = fma2(fx2, sub2(d_ba, c_ba), c_ba);
  Sample s;
  s.rg = fma2(fy2, sub2(bot_rg, top_rg), top_rg);
  s.ba = fma2(fy2, sub2(bot_ba, top_ba), top_ba);
  return s;
}

template <typename T>
__device__ __forceinline__ SampleGrad sample_staged_grad(const typename Texel<T>::Vec* __restrict__ q, int pitch,
                                                         float fx, float fy) {
  f32x2 a_rg, a_ba, b_rg, b_ba, c_rg, c_ba, d_rg, d_ba;
  Texel<T>::unpack(q[0], a_rg, a_ba);
  Texel<T>::unpack(q[1], b_rg, b_ba);
  Texel<T>::unpack(q[pitch], c_rg, c_ba);
  Texel<T>::unpack(q[pitch + 1], d_rg, d_ba);
  const f32x2 fx2 = bc(fx), fy2 = bc(fy);
  const f32x2 dxt_rg = sub2(b_rg, a_rg), dxt_ba = sub2(b_ba, a_ba);
  const f32x2 dxb_rg = sub2(d_rg, c_rg), dxb_ba = sub2(d_ba, c_ba);
  const f32x2 top_rg = fma2(fx2, dxt_rg, a_rg), top_ba = fma2(fx2, dxt_ba, a_ba);
  const f32x2 bot_rg = fma2(fx2, dxb_rg, c_rg), bot_ba = fma2(fx2, dxb_ba, c_ba);
  SampleGrad s;
  s.dy_rg = sub2(bot_rg, top_rg);
  s.dy_ba = sub2(bot_ba, top_ba);
  s.rg = fma2(fy2, s.dy_rg, top_rg);
  s.ba = fma2(fy2, s.dy_ba, top_ba);
  s.dx_rg = fma2(fy2, sub2(dxb_rg, dxt_rg), dxt_rg);
  s.dx_ba = fma2(fy2, sub2(dxb_ba, dxt_ba), dxt_ba);
  return s;
}

// ---- theta-gradient reduction of the tiled backward kernels --------------------------------------------
// During the layer sweep a thread parks its four partial sums (sum dix, sum dix*y_i, sum diy, sum diy*y_i over its
// pixels) of layer l in the four shared-memory slots that held its transmittances T_l: they are dead once the layer
// has been swept, and the owner is the only thread that ever touched them, so no barrier is needed.  After the sweep
// (and one barrier) warp w folds layers w, w + 8, ...: lane t adds the eight threads of column t (they share x_j),
// forms the six affine coefficients' terms and the warp finishes with one transposing butterfly and six global
// atomics per (CTA, layer) -- instead of a butterfly plus shared-memory atomics per (warp, layer).
__device__ __forceinline__ void park_theta_partials(float* Tst_l, float accx, float accxy, float accy, float accyy) {
  Tst_l[0 * kTiledThreads] = accx;
  Tst_l[1 * kTiledThreads] = accxy;
  Tst_l[2 * kTiledThreads] = accy;
  Tst_l[3 * kTiledThreads] = accyy;
}
// `stash` = base of the [L][kPx][kTiledThreads] array (NOT offset by tid); call after __syncthreads()
__device__ __forceinline__ void reduce_theta_partials(const float* stash, int L, int tid, float xj, float hW, float hH,
                                                      float* __restrict__ gth) {
  static_assert(kPx == 4, "four partial sums are parked in the kPx slots of a layer");
  const int lane = tid & 31;
  for (int l = tid >> 5; l < L; l += kTiledThreads / 32) {
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float* s = stash + (l * kPx + q) * kTiledThreads + lane;
      float a = 0.f;
#pragma unroll
      for (int r = 0; r < kTiledThreads / 32; ++r) a += s[32 * r];
      v[q] = a;
    }
    const float part[6] = {hW * v[0] * xj, hW * v[1], hW * v[0], hH * v[2] * xj, hH * v[3], hH * v[2]};
    const float sum = warp_sum6(part, lane);
    const int q = warp_sum6_index(lane);
    if ((lane & 3) == 0 && q < 6) atomicAdd(gth + l * 6 + q, sum);
  }
}

}  // namespace mgr
