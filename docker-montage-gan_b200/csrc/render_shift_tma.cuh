// Stencil kernels on TMA-staged planar footprints: the production case (the reference's placement network emits pure
// translations: fukuwarai/networks.py:246-247 through custom_utils/image_utils.py:316-335).
//
// Under a translation (X + fx, Y + fy) the footprint of the output tile with origin (j0, i0) is the rectangle of each
// channel plane starting at texel (j0 + X, i0 + Y): one 5-D tensor-map box copy {W: tile + 1 + alignment slack,
// H: tile + 1, C: 4, layer, sample} per (tile, layer), issued by one lane of a producer warp into a ring of
// shared-memory stages.  The copy unit wants the box to start on a 16-byte boundary of the row (measured: any other
// innermost coordinate is an illegal-instruction fault, negative and past-the-end coordinates are fine --
// tools/micro/tma_probe.cu), so the box starts at the 16-byte boundary at or left of the footprint and a thread's
// five texels sit at one of four offsets inside two aligned shared-memory chunks: the per-layer body is compiled for
// each offset (CTA-uniform switch), the extraction costs nothing at run time; the consumer warps wait on the stage's mbarrier and never touch global memory for x.  What the
// copy removes, measured on the round-2 stencil forward (ncu): the staging loop (index arithmetic, 128-bit loads held in
// registers, byte permutes to interleave the channels, bank-conflicted shared stores) was more than half of its 89
// warp-instructions per layer-pixel.  Out-of-range texels arrive as zeros = transparent black in the [0,1] range mode; in
// the [-1,1] mode the padding value is -1, so tiles that touch the layer's border patch the out-of-range part of the box
// before sampling (CTA-uniform, border tiles only) -- the lerp arithmetic then sees exactly what the staged kernels saw
// (a run of equal taps returns that value exactly: transparent stays exactly transparent).
//
// A thread owns a 4-wide, 2-tall strip of pixels: three rows of five texels per channel, fetched as one 64-bit (16-bit
// storage) or 128-bit (fp32) shared load plus one 32-bit load per row, lerped along x once per row and along y once per
// pixel with packed fp32x2 arithmetic on pairs of horizontally adjacent pixels.
#pragma once
#include "render_shift.cuh"
#include "tma.cuh"

#ifndef MGR_STF_STAGES16
#define MGR_STF_STAGES16 4
#endif
#ifndef MGR_STF_BLOCKS
#define MGR_STF_BLOCKS 4
#endif

namespace mgr {

#ifndef MGR_STF_TILE_H
#define MGR_STF_TILE_H 16
#endif
constexpr int kSW = 64, kSH = MGR_STF_TILE_H;     // output tile of the TMA stencil forward (height a multiple of 4)
constexpr int kSConsumers = 8 * kSH;              // 16 x (kSH / 2) strips of 4 x 2 pixels
constexpr int kSThreads = kSConsumers + 32;       // + the producer warp
constexpr int kSMaxStages = 4;

template <typename T, int TH = kSH> struct ShiftBox {
  static constexpr int kAlign = 16 / (int)sizeof(T);        // the box starts on a 16-byte boundary of the row
  static constexpr int W = kSW + kAlign;                    // tile + 1 tap columns + up to kAlign - 1 columns of alignment slack
  static constexpr int H = TH + 1;
  static constexpr int kPlane = W * H;                      // elements per channel plane
  static constexpr int kBytes = kPlane * 4 * (int)sizeof(T);
  static constexpr int kStageBytes = (kBytes + 127) & ~127; // stages start on 128-byte boundaries
  static constexpr int kStages = sizeof(T) == 4 ? 3 : MGR_STF_STAGES16;
};

// the taps of the tile (columns x0 .. x0 + kSW, rows y0 .. y0 + kSH) all miss the image: the layer is transparent here
template <int TH = kSH>
__device__ __forceinline__ bool shift_box_misses(int x0, int y0, int W, int H) {
  return x0 + kSW < 0 || x0 >= W || y0 + TH < 0 || y0 >= H;
}

// ---- five adjacent texels e0..e4 of one channel row -> the x-lerped values of the strip's four pixels ------------------
// `p` points at the aligned chunk (four elements: 8 bytes of 16-bit storage, 16 bytes of fp32) that holds e0 at element
// offset R in 0..3; e0..e4 lie inside that chunk and the next one.
// Packed arithmetic wants its operands in aligned register pairs.  16-bit storage: every element is produced by its own
// unpack instruction, which can write any register, so the pixels are paired (0, 2) and (1, 3): the left / right taps are
// then the three pairs (e0, e2), (e1, e3), (e2, e4) -- seven unpacks per row, no register moves (pairing (0, 1), (2, 3)
// needs (e0,e1), (e1,e2), (e2,e3), (e3,e4): the compiler materialised them with 9-10 moves per row).  fp32 storage: the
// pixels are paired (0, 1) and (2, 3) (the 128-bit loads leave neighbours in neighbouring registers).
// Pair order of the results: Row5<T>::kStrided ? {(px0, px2), (px1, px3)} : {(px0, px1), (px2, px3)}.
template <typename T> struct Row5;
template <> struct Row5<float> {
  static constexpr bool kStrided = false;
  // R = 0..3: e0 at compile-time offset R of the aligned 16-byte chunk at p; R = -1: the offset is `off` at run time (CTA-uniform):
  // eleven selects instead of a copy of the layer body per offset (the backward's loop has to fit the instruction cache)
  template <int R>
  static __device__ __forceinline__ void taps(const float* p, int off, f32x2& L0, f32x2& L1, f32x2& R0, f32x2& R1) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if constexpr (R < 0) {
      const bool two = off & 2, one = off & 1;
      float w[6], e[5];
#pragma unroll
      for (int k = 0; k < 6; ++k) w[k] = two ? v[k + 2] : v[k];
#pragma unroll
      for (int k = 0; k < 5; ++k) e[k] = one ? w[k + 1] : w[k];
      L0 = pk(e[0], e[1]); L1 = pk(e[2], e[3]); R0 = pk(e[1], e[2]); R1 = pk(e[3], e[4]);
    } else {
      constexpr int Q = R < 0 ? 0 : R;
      L0 = pk(v[Q], v[Q + 1]); L1 = pk(v[Q + 2], v[Q + 3]);
      R0 = pk(v[Q + 1], v[Q + 2]); R1 = pk(v[Q + 3], v[Q + 4]);
    }
  }
};
// 16-bit storage: `p` points at the EVEN element at or left of e0 (a 32-bit word boundary), `sh` = 16 if e0 is the odd half
// of that word, else 0.  Three 32-bit loads and three funnel shifts give the words (e0,e1), (e2,e3), (e4,.) whatever the
// alignment, so ONE copy of the layer body serves every offset (four copies, one per offset inside an aligned 8-byte chunk,
// put the backward's layer loop at 44 KB of code: ncu showed 32 % of the stall samples waiting for instructions).
struct Words3 { uint32_t a0, a1, a2; };
template <typename T>
__device__ __forceinline__ Words3 row_words(const T* p, int sh) {
  const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
  const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
  Words3 r;
  r.a0 = __funnelshift_r(w0, w1, sh); r.a1 = __funnelshift_r(w1, w2, sh); r.a2 = w2 >> sh;
  return r;
}
// R = -1: funnel-shifted words (p = even element, runtime sh); R = 0..3: e0 at compile-time offset R of the aligned 8-byte
// chunk at p (two 64-bit loads, no shifts: what the forward uses -- its four bodies fit the instruction cache)
template <> struct Row5<__nv_bfloat16> {
  static constexpr bool kStrided = true;
  template <int M> static __device__ __forceinline__ float elem(const uint32_t (&w)[4]) {
    return __uint_as_float((M & 1) ? (w[M >> 1] & 0xffff0000u) : (w[M >> 1] << 16));
  }
  template <int R>
  static __device__ __forceinline__ void taps(const __nv_bfloat16* p, int sh, f32x2& L0, f32x2& L1, f32x2& R0, f32x2& R1) {
    if constexpr (R < 0) {
      const Words3 w = row_words(p, sh);
      L0 = pk(__uint_as_float(w.a0 << 16), __uint_as_float(w.a1 << 16));
      L1 = R0 = pk(__uint_as_float(w.a0 & 0xffff0000u), __uint_as_float(w.a1 & 0xffff0000u));
      R1 = pk(__uint_as_float(w.a1 << 16), __uint_as_float(w.a2 << 16));
    } else {
      const uint2 c0 = *reinterpret_cast<const uint2*>(p), c1 = *reinterpret_cast<const uint2*>(p + 4);
      const uint32_t w[4] = {c0.x, c0.y, c1.x, c1.y};
      constexpr int Q = R < 0 ? 0 : R;
      L0 = pk(elem<Q>(w), elem<Q + 2>(w));
      L1 = R0 = pk(elem<Q + 1>(w), elem<Q + 3>(w));
      R1 = pk(elem<Q + 2>(w), elem<Q + 4>(w));
    }
  }
};
template <> struct Row5<__half> {
  static constexpr bool kStrided = true;
  template <int M> static __device__ __forceinline__ float elem(const uint32_t (&w)[4]) {
    const __half2 h = *reinterpret_cast<const __half2*>(&w[M >> 1]);
    return (M & 1) ? __high2float(h) : __low2float(h);
  }
  template <int R>
  static __device__ __forceinline__ void taps(const __half* p, int sh, f32x2& L0, f32x2& L1, f32x2& R0, f32x2& R1) {
    if constexpr (R < 0) {
      const Words3 w = row_words(p, sh);
      const __half2 h0 = *reinterpret_cast<const __half2*>(&w.a0), h1 = *reinterpret_cast<const __half2*>(&w.a1),
                    h2 = *reinterpret_cast<const __half2*>(&w.a2);
      L0 = pk(__low2float(h0), __low2float(h1));
      L1 = R0 = pk(__high2float(h0), __high2float(h1));
      R1 = pk(__low2float(h1), __low2float(h2));
    } else {
      const uint2 c0 = *reinterpret_cast<const uint2*>(p), c1 = *reinterpret_cast<const uint2*>(p + 4);
      const uint32_t w[4] = {c0.x, c0.y, c1.x, c1.y};
      constexpr int Q = R < 0 ? 0 : R;
      L0 = pk(elem<Q>(w), elem<Q + 2>(w));
      L1 = R0 = pk(elem<Q + 1>(w), elem<Q + 3>(w));
      R1 = pk(elem<Q + 2>(w), elem<Q + 4>(w));
    }
  }
};
// two packed pairs <-> the four pixels in column order
template <typename T>
__device__ __forceinline__ void strip_unpack(f32x2 q0, f32x2 q1, float (&v)[4]) {
  if (Row5<T>::kStrided) { upk(q0, v[0], v[2]); upk(q1, v[1], v[3]); }
  else { upk(q0, v[0], v[1]); upk(q1, v[2], v[3]); }
}
template <typename T>
__device__ __forceinline__ void strip_pack(const float (&v)[4], f32x2& q0, f32x2& q1) {
  if (Row5<T>::kStrided) { q0 = pk(v[0], v[2]); q1 = pk(v[1], v[3]); }
  else { q0 = pk(v[0], v[1]); q1 = pk(v[2], v[3]); }
}

template <typename T> __device__ __forceinline__ T raw_minus_one();
template <> __device__ __forceinline__ float raw_minus_one<float>() { return -1.f; }
template <> __device__ __forceinline__ __nv_bfloat16 raw_minus_one<__nv_bfloat16>() { return __ushort_as_bfloat16((unsigned short)0xBF80u); }
template <> __device__ __forceinline__ __half raw_minus_one<__half>() { return __ushort_as_half((unsigned short)0xBC00u); }

// [-1,1] range mode: texels of the box outside the image become the padding value -1 (the copy zero-fills them).
// Threads 0..255 of the consumer group.
template <typename T, int TH = kSH>
__device__ __forceinline__ void shift_patch_oob(T* stage, int x0, int y0, int W, int H, int tid) {   // (x0, y0): texel of box element (0, 0)
  constexpr int BW = ShiftBox<T, TH>::W, BH = ShiftBox<T, TH>::H, NC = ShiftBox<T, TH>::W;
  const int nT = min(max(-y0, 0), BH), nB = min(max(y0 + BH - H, 0), BH);
  const int nL = min(max(-x0, 0), NC), nR = min(max(x0 + NC - W, 0), NC);
  const T m1 = raw_minus_one<T>();
  const int cl = tid & 7;                                     // 8 column lanes x 32 (row, channel) lanes
#pragma unroll 1
  for (int rc = tid >> 3; rc < 4 * BH; rc += TH) {                // 8 TH threads take part
    const int r = rc >> 2, c = rc & 3;
    T* row = stage + c * ShiftBox<T, TH>::kPlane + r * BW;
    if (r < nT || r >= BH - nB) {
#pragma unroll 1
      for (int k = cl; k < NC; k += 8) row[k] = m1;
    } else {
#pragma unroll 1
      for (int k = cl; k < nL; k += 8) row[k] = m1;
#pragma unroll 1
      for (int k = NC - nR + cl; k < NC; k += 8) row[k] = m1;
    }
  }
}

template <typename SA> __device__ __forceinline__ void st_alpha4(SA* p, float a0, float a1, float a2, float a3);
template <> __device__ __forceinline__ void st_alpha4<float>(float* p, float a0, float a1, float a2, float a3) {
  *reinterpret_cast<float4*>(p) = make_float4(a0, a1, a2, a3);
}
template <> __device__ __forceinline__ void st_alpha4<__half>(__half* p, float a0, float a1, float a2, float a3) {
  const __half2 lo_ = __floats2half2_rn(a0, a1), hi_ = __floats2half2_rn(a2, a3);
  uint2 t;
  t.x = *reinterpret_cast<const uint32_t*>(&lo_); t.y = *reinterpret_cast<const uint32_t*>(&hi_);
  *reinterpret_cast<uint2*>(p) = t;
}

// One channel of a 4 x 2 strip: raw bilinear samples of two rows of two pixel pairs (lerp form, x then y).
template <typename T, int R>
__device__ __forceinline__ void shift_sample_strip(const T* p, int sh, f32x2 fx2, f32x2 fy2, f32x2 (&v)[2][2]) {
  constexpr int BW = ShiftBox<T>::W;
  f32x2 h0[3], h1[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    f32x2 L0, L1, R0, R1;
    Row5<T>::template taps<R>(p + r * BW, sh, L0, L1, R0, R1);
    h0[r] = fma2(fx2, sub2(R0, L0), L0);
    h1[r] = fma2(fx2, sub2(R1, L1), L1);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    v[r][0] = fma2(fy2, sub2(h0[r + 1], h0[r]), h0[r]);
    v[r][1] = fma2(fy2, sub2(h1[r + 1], h1[r]), h1[r]);
  }
}

struct ShiftRing {
  uint64_t full[kSMaxStages];       // the copy's bytes have landed (tx count), consumers wait
  uint64_t empty[kSMaxStages];      // one arrival per consumer warp, the producer waits
};

template <typename T>
inline size_t shift_tma_fwd_smem_bytes(int L) {
  return (size_t)ShiftBox<T>::kStageBytes * ShiftBox<T>::kStages + sizeof(ShiftPlan) * (size_t)L + sizeof(ShiftRing);
}

template <typename T, bool kSave>
__global__ void __launch_bounds__(kSThreads, MGR_STF_BLOCKS)
render_fwd_shift_tma(const __grid_constant__ CUtensorMap xmap, const float* __restrict__ theta, T* __restrict__ out,
                     typename SavedAlpha<T>::type* __restrict__ sav, Geometry g, const int* __restrict__ shift_flags) {
  using SA = typename SavedAlpha<T>::type;
  using Box = ShiftBox<T>;
  constexpr int S = Box::kStages;
  const int b = blockIdx.z;
  // is the sample this kernel's?  (flag written by sample_shift_flags_kernel; without one every warp looks at the placements)
  if (shift_flags ? shift_flags[b] == 0
                  : !__all_sync(0xffffffffu, (int)(threadIdx.x & 31) >= g.L || is_pure_shift(theta + ((long long)b * g.L + (threadIdx.x & 31)) * 6)))
    return;
  extern __shared__ __align__(128) unsigned char smem[];     // no static shared memory in this kernel: the window starts aligned
  ShiftPlan* splan = reinterpret_cast<ShiftPlan*>(smem + (size_t)Box::kStageBytes * S);
  ShiftRing* ring = reinterpret_cast<ShiftRing*>(splan + g.L);
  const int tid = threadIdx.x;
  const int j0 = blockIdx.x * kSW, i0 = blockIdx.y * kSH;
  for (int l = tid; l < g.L; l += kSThreads) splan[l] = make_shift_plan(theta + ((long long)b * g.L + l) * 6, g.H, g.W);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) { tma_mbar_init(&ring->full[s], 1); tma_mbar_init(&ring->empty[s], kSConsumers / 32); }
    tma_fence_barrier_init();
  }
  __syncthreads();

  if (tid >= kSConsumers) {                       // ---- producer warp: one lane issues the box copies ----
    if (tid == kSConsumers) {
      tma_prefetch_map(&xmap);
      int it = 0;
      for (int l = 0; l < g.L; ++l) {
        const int x0 = j0 + splan[l].X, y0 = i0 + splan[l].Y;
        if (shift_box_misses(x0, y0, g.W, g.H)) continue;
        const int s = it % S;
        if (it >= S) tma_mbar_wait(&ring->empty[s], (uint32_t)((it / S) - 1) & 1u);
        tma_mbar_expect_tx(&ring->full[s], (uint32_t)Box::kBytes);
        tma_load_5d(smem + (size_t)Box::kStageBytes * s, &xmap, &ring->full[s], x0 & ~(Box::kAlign - 1), y0, 0, l, b);
        ++it;
      }
    }
    return;
  }

  // ---- consumers ----
  const int tx = tid & 15, ty = tid >> 4;
  const int j = j0 + 4 * tx, i = i0 + 2 * ty;
  const bool col_live = j < g.W;                            // W is a multiple of 4: a strip is inside or outside as a whole
  const bool live0 = col_live && i < g.H, live1 = col_live && i + 1 < g.H;
  const int hw = g.H * g.W;
  const f32x2 zs2 = bc(g.m11 ? 0.5f : 1.f), zb2 = bc(g.m11 ? 0.5f : 0.f), one2 = bc(1.f);
  f32x2 Sc[3][2][2], R_[2][2];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int p = 0; p < 2; ++p) { Sc[0][r][p] = Sc[1][r][p] = Sc[2][r][p] = R_[r][p] = bc(0.f); }
  const int toff = (2 * ty) * Box::W + 4 * tx;              // this strip's aligned chunk inside a plane of the box, before the layer's offset
  const int poff = i * g.W + j;                               // this strip's first pixel inside a plane (H * W < 2^29)

  int it = 0;
  for (int l = 0; l < g.L; ++l) {
    const ShiftPlan sp = splan[l];
    const int x0 = j0 + sp.X, y0 = i0 + sp.Y;
    SA* sv = kSave ? sav + ((long long)b * g.L + l) * hw + poff : nullptr;      // CTA-uniform base + one 32-bit offset
    if (shift_box_misses(x0, y0, g.W, g.H)) {
      if (kSave) {
        if (live0) st_alpha4<SA>(sv, 0.f, 0.f, 0.f, 0.f);
        if (live1) st_alpha4<SA>(sv + g.W, 0.f, 0.f, 0.f, 0.f);
      }
      continue;
    }
    const int s = it % S;
    T* stage = reinterpret_cast<T*>(smem + (size_t)Box::kStageBytes * s);
    const int xa = x0 & ~(Box::kAlign - 1), dx = x0 - xa;     // box origin and the footprint's offset inside it
    tma_mbar_wait(&ring->full[s], (uint32_t)(it / S) & 1u);
    if (g.m11 && (xa < 0 || xa + Box::W > g.W || y0 < 0 || y0 + Box::H > g.H)) {       // CTA-uniform
      shift_patch_oob<T>(stage, xa, y0, g.W, g.H, tid);
      tma_fence_proxy_async();
      named_barrier(1, kSConsumers);
    }
    const f32x2 fx2 = bc(sp.fx), fy2 = bc(sp.fy);
    // e0 sits at offset R = dx & 3 of the strip's aligned chunk (8 bytes of 16-bit storage, 16 of fp32): the body is compiled per R
    const T* p = stage + toff + (dx & ~3);
    const int sh = 0;
    auto body = [&](auto rtag) {
      constexpr int R = decltype(rtag)::value;
      f32x2 a[2][2], om[2][2];
      {
        f32x2 v[2][2];
        shift_sample_strip<T, R>(p + 3 * Box::kPlane, sh, fx2, fy2, v);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            a[r][q] = fma2(v[r][q], zs2, zb2);
            om[r][q] = sub2(one2, a[r][q]);
            R_[r][q] = fma2(om[r][q], R_[r][q], a[r][q]);
          }
      }
      if (kSave) {
        float av[4];
        strip_unpack<T>(a[0][0], a[0][1], av);
        if (live0) st_alpha4<SA>(sv, av[0], av[1], av[2], av[3]);
        strip_unpack<T>(a[1][0], a[1][1], av);
        if (live1) st_alpha4<SA>(sv + g.W, av[0], av[1], av[2], av[3]);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        f32x2 v[2][2];
        shift_sample_strip<T, R>(p + c * Box::kPlane, sh, fx2, fy2, v);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int q = 0; q < 2; ++q) Sc[c][r][q] = fma2(om[r][q], Sc[c][r][q], mul2(a[r][q], fma2(v[r][q], zs2, zb2)));
      }
    };
    switch (dx & 3) {
      case 0: body(std::integral_constant<int, 0>{}); break;
      case 1: body(std::integral_constant<int, 1>{}); break;
      case 2: body(std::integral_constant<int, 2>{}); break;
      default: body(std::integral_constant<int, 3>{}); break;
    }
    __syncwarp();
    if ((tid & 31) == 0) tma_mbar_arrive(&ring->empty[s]);
    ++it;
  }

  const float os = g.m11 ? 2.f : 1.f, obias = g.m11 ? -1.f : 0.f;
  T* outp = out + (long long)b * 4 * hw + poff;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (!(r ? live1 : live0)) continue;
    float Rv[4], inv[4];
    strip_unpack<T>(R_[r][0], R_[r][1], Rv);
#pragma unroll
    for (int k = 0; k < 4; ++k) inv[k] = (Rv[k] != 0.f) ? 1.f / Rv[k] : 0.f;       // nan_to_num(0/0) = 0 (image_utils.py:132)
    T* o = outp + r * g.W;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float sv4[4], ov[4];
      strip_unpack<T>(Sc[c][r][0], Sc[c][r][1], sv4);
#pragma unroll
      for (int k = 0; k < 4; ++k) ov[k] = fmaf(sv4[k] * inv[k], os, obias);
      st_vec4<T>(o + c * hw, ov);
    }
    float av[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) av[k] = fmaf(Rv[k], os, obias);
    st_vec4<T>(o + 3 * hw, av);
  }
}

// the 5-D map {W, H, 4, L, B} over the canvas layout x[B,L,4,H,W] with the forward's box; false: keep the staged kernel
template <typename T>
inline bool shift_tma_x_map(CUtensorMap* map, const void* x, const Geometry& g, int box_w, int box_h) {
  const long long dims[5] = {g.W, g.H, 4, g.L, g.B};
  const long long strides[5] = {1, g.sh, g.sc, g.sl, g.sb};
  const int box[5] = {box_w, box_h, 4, 1, 1};
  return tma_make_map(map, x, (int)sizeof(T), 5, dims, strides, box);
}

}  // namespace mgr

constexpr int kWTH = 32, kWConsumers = 8 * kWTH;          // the materialised-warp kernels keep 64 x 32 tiles (64 x 16: +7 % on the warp forward)

// ---- materialised warp of pure-translation LAYERS (what STNv2c / STNv2b return, fukuwarai/networks.py:250-257, and what
// random_position computes, custom_utils/image_utils.py:281-294) on the same box copy: one (layer, 64 x 32 tile) per CTA, the
// raw bilinear lerp of the four planes written out.  The decision is per LAYER here (each layer is warped on its own): CTAs
// of layers with a general placement leave at once, warp_fwd_tiled makes the opposite choice.  (A CTA walking a band of
// tiles with the next box in flight was measured no faster: 109 / 171 us against 109 / 163 us, bf16 / fp32 at C2.)
namespace mgr {

// kAdjoint: the backward of the same op w.r.t. the layer.  The adjoint of a 2 x 2 stencil with weights (1-fx, fx) x (1-fy, fy)
// at integer shift (X, Y) is the 2 x 2 stencil with weights (fx, 1-fx) x (fy, 1-fy) at shift (-X-1, -Y-1) applied to the
// upstream gradient, zeros outside the image in either range mode: the same kernel on another plan (`xmap` is then the map
// over the gradient of the warped layers, `out` is grad_x).  A whole-pixel shift (fx == 0) stays a whole-pixel shift (-X).
template <typename T, bool kAdjoint>
__global__ void __launch_bounds__(kWConsumers, 6)
warp_fwd_shift_tma(const __grid_constant__ CUtensorMap xmap, const float* __restrict__ theta, T* __restrict__ out, Geometry g) {
  using Box = ShiftBox<T, kWTH>;
  const int n = blockIdx.z;                                   // b * L + l
  const float* th = theta + (long long)n * 6;
  if (!is_pure_shift(th)) return;
  if (kAdjoint) g.m11 = 0;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  T* stage = reinterpret_cast<T*>(smem);
  const int tid = threadIdx.x;
  const int b = n / g.L, l = n - b * g.L;
  const int j0 = blockIdx.x * kSW, i0 = blockIdx.y * kWTH;
  ShiftPlan sp = make_shift_plan(th, g.H, g.W);               // every thread: no broadcast needed, the values are uniform
  if (kAdjoint) {
    sp.X = sp.fx == 0.f ? -sp.X : -sp.X - 1; sp.fx = sp.fx == 0.f ? 0.f : 1.f - sp.fx;
    sp.Y = sp.fy == 0.f ? -sp.Y : -sp.Y - 1; sp.fy = sp.fy == 0.f ? 0.f : 1.f - sp.fy;
  }
  const int x0 = j0 + sp.X, y0 = i0 + sp.Y;
  const bool miss = shift_box_misses<kWTH>(x0, y0, g.W, g.H);
  const int xa = x0 & ~(Box::kAlign - 1), dx = x0 - xa;
  if (tid == 0 && !miss) {
    tma_mbar_init(&bar, 1);
    tma_fence_barrier_init();
    tma_mbar_expect_tx(&bar, (uint32_t)Box::kBytes);
    tma_load_5d(stage, &xmap, &bar, xa, y0, 0, l, b);
  }
  __syncthreads();                                            // the barrier is initialised before anyone waits on it
  const int tx = tid & 15, ty = tid >> 4;
  const int j = j0 + 4 * tx, i = i0 + 2 * ty;
  const bool col_live = j < g.W;
  const int hw = g.H * g.W;
  T* op = out + (long long)n * 4 * hw + (long long)i * g.W + j;
  const float padv = g.m11 ? -1.f : 0.f;
  if (miss) {                                                 // every tap is outside the image: the padding value
    const float p4[4] = {padv, padv, padv, padv};
#pragma unroll
    for (int r = 0; r < 2; ++r)
      if (col_live && i + r < g.H) {
#pragma unroll
        for (int c = 0; c < 4; ++c) st_vec4<T>(op + c * hw + r * g.W, p4);
      }
    return;
  }
  tma_mbar_wait(&bar, 0);
  if (g.m11 && (xa < 0 || xa + Box::W > g.W || y0 < 0 || y0 + Box::H > g.H)) {         // CTA-uniform
    shift_patch_oob<T, kWTH>(stage, xa, y0, g.W, g.H, tid);
    __syncthreads();
  }
  const f32x2 fx2 = bc(sp.fx), fy2 = bc(sp.fy);
  const T* p = stage + (2 * ty) * Box::W + 4 * tx + (dx & ~3);
  auto body = [&](auto rtag) {
    constexpr int R = decltype(rtag)::value;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      f32x2 v[2][2];
      shift_sample_strip<T, R>(p + c * Box::kPlane, 0, fx2, fy2, v);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float o4[4];
        strip_unpack<T>(v[r][0], v[r][1], o4);
        if (col_live && i + r < g.H) st_vec4<T>(op + c * hw + r * g.W, o4);
      }
    }
  };
  switch (dx & 3) {
    case 0: body(std::integral_constant<int, 0>{}); break;
    case 1: body(std::integral_constant<int, 1>{}); break;
    case 2: body(std::integral_constant<int, 2>{}); break;
    default: body(std::integral_constant<int, 3>{}); break;
  }
}

// grad_theta of the materialised warp for a translation layer: the box copy of the footprint, d sample / d (ix, iy) from the
// lerp differences of the four planes (layer-wide weights), contracted with the upstream gradient of the strip's eight pixels
// (two 8- / 16-byte loads per plane), six sums per thread -> transposing butterfly per warp -> one atomic per (CTA, coefficient).
template <typename T>
__global__ void __launch_bounds__(kWConsumers, 3)
warp_bwd_theta_shift_tma(const __grid_constant__ CUtensorMap xmap, const float* __restrict__ theta, const T* __restrict__ gw,
                         float* __restrict__ gtheta, Geometry g) {
  using Box = ShiftBox<T, kWTH>;
  const int n = blockIdx.z;
  const float* th = theta + (long long)n * 6;
  if (!is_pure_shift(th)) return;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ float s_red[kWConsumers / 32][8];
  T* stage = reinterpret_cast<T*>(smem);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = n / g.L, l = n - b * g.L;
  const int j0 = blockIdx.x * kSW, i0 = blockIdx.y * kWTH;
  const ShiftPlan sp = make_shift_plan(th, g.H, g.W);
  const int x0 = j0 + sp.X, y0 = i0 + sp.Y;
  if (shift_box_misses<kWTH>(x0, y0, g.W, g.H)) return;               // the taps miss the image: no dependence on theta (CTA-uniform)
  const int xa = x0 & ~(Box::kAlign - 1), dx = x0 - xa;
  if (tid == 0) {
    tma_mbar_init(&bar, 1);
    tma_fence_barrier_init();
    tma_mbar_expect_tx(&bar, (uint32_t)Box::kBytes);
    tma_load_5d(stage, &xmap, &bar, xa, y0, 0, l, b);
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  const int j = j0 + 4 * tx, i = i0 + 2 * ty;
  const int hw = g.H * g.W;
  const bool live[2] = {j < g.W && i < g.H, j < g.W && i + 1 < g.H};
  // the upstream gradient of the strip, in flight while the box lands
  f32x2 gq[4][2][2];
  const T* gp_ = gw + (long long)n * 4 * hw + (long long)i * g.W + j;
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float v4[4] = {0.f, 0.f, 0.f, 0.f};
      if (live[r]) ld_vec<T, 4>(gp_ + c * hw + r * g.W, v4);
      strip_pack<T>(v4, gq[c][r][0], gq[c][r][1]);
    }
  tma_mbar_wait(&bar, 0);
  if (g.m11 && (xa < 0 || xa + Box::W > g.W || y0 < 0 || y0 + Box::H > g.H)) {         // CTA-uniform
    shift_patch_oob<T, kWTH>(stage, xa, y0, g.W, g.H, tid);
    __syncthreads();
  }
  const f32x2 fx2 = bc(sp.fx), fy2 = bc(sp.fy);
  const T* p = stage + (2 * ty) * Box::W + 4 * tx + (dx & ~3);
  f32x2 ix[2][2], iy[2][2];
#pragma unroll
  for (int r = 0; r < 2; ++r) ix[r][0] = ix[r][1] = iy[r][0] = iy[r][1] = bc(0.f);
  auto body = [&](auto rtag) {
    constexpr int R = decltype(rtag)::value;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      constexpr int BW = Box::W;
      f32x2 h0[3], h1[3], d0[3], d1[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        f32x2 L0, L1, R0, R1;
        Row5<T>::template taps<R>(p + c * Box::kPlane + r * BW, 0, L0, L1, R0, R1);
        d0[r] = sub2(R0, L0); d1[r] = sub2(R1, L1);
        h0[r] = fma2(fx2, d0[r], L0); h1[r] = fma2(fx2, d1[r], L1);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        // d/diy = h[r+1] - h[r];  d/dix = lerp_y of the row differences
        iy[r][0] = fma2(gq[c][r][0], sub2(h0[r + 1], h0[r]), iy[r][0]);
        iy[r][1] = fma2(gq[c][r][1], sub2(h1[r + 1], h1[r]), iy[r][1]);
        ix[r][0] = fma2(gq[c][r][0], fma2(fy2, sub2(d0[r + 1], d0[r]), d0[r]), ix[r][0]);
        ix[r][1] = fma2(gq[c][r][1], fma2(fy2, sub2(d1[r + 1], d1[r]), d1[r]), ix[r][1]);
      }
    }
  };
  switch (dx & 3) {
    case 0: body(std::integral_constant<int, 0>{}); break;
    case 1: body(std::integral_constant<int, 1>{}); break;
    case 2: body(std::integral_constant<int, 2>{}); break;
    default: body(std::integral_constant<int, 3>{}); break;
  }
  // six sums over the strip (pixels outside the image carry a zero upstream gradient)
  const float inv_w = 1.f / (float)g.W, inv_h = 1.f / (float)g.H;
  float xs[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) xs[k] = fmaf((float)(2 * (j + k) + 1), inv_w, -1.f);
  f32x2 xq0, xq1;
  strip_pack<T>(xs, xq0, xq1);
  const float y0n = fmaf((float)(2 * i + 1), inv_h, -1.f), y1n = fmaf((float)(2 * i + 3), inv_h, -1.f);
  auto hs = [](f32x2 v) { float a, b; upk(v, a, b); return a + b; };
  const float sx0 = hs(add2(ix[0][0], ix[0][1])), sx1 = hs(add2(ix[1][0], ix[1][1]));
  const float sy0 = hs(add2(iy[0][0], iy[0][1])), sy1 = hs(add2(iy[1][0], iy[1][1]));
  const float hW = 0.5f * (float)g.W, hH = 0.5f * (float)g.H;
  const float th6[6] = {hW * hs(fma2(add2(ix[0][0], ix[1][0]), xq0, mul2(add2(ix[0][1], ix[1][1]), xq1))), hW * fmaf(sx0, y0n, sx1 * y1n), hW * (sx0 + sx1),
                        hH * hs(fma2(add2(iy[0][0], iy[1][0]), xq0, mul2(add2(iy[0][1], iy[1][1]), xq1))), hH * fmaf(sy0, y0n, sy1 * y1n), hH * (sy0 + sy1)};
  const float sum = warp_sum6(th6, lane);
  const int qi = warp_sum6_index(lane);
  if ((lane & 3) == 0 && qi < 6) s_red[wid][qi] = sum;
  __syncthreads();
  if (tid < 6) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kWConsumers / 32; ++w) v += s_red[w][tid];
    atomicAdd(gtheta + (long long)n * 6 + tid, v);
  }
}

}  // namespace mgr
