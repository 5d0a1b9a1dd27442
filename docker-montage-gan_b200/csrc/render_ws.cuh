// Warp-specialised general-placement kernels (forward and backward pass 1).
//
// A CTA still owns one 32x32 output tile of one sample and walks the layers back to front, but staging and sampling
// are decoupled: warps 8..11 (PRODUCERS) move each layer's source footprint from the planar global layout into a ring
// of two shared-memory buffers of channel-interleaved texels, warps 0..7 (CONSUMERS) sample and composite.  The two
// sides meet only at mbarriers (full[s] / empty[s] per ring slot), so the global-load latency of layer l + 1 hides
// behind the arithmetic of layer l instead of sitting between two __syncthreads() as in render_tiled.cuh
// (ncu, round 1: staging was 34 % of pass 1's warp-time and a third of the forward's instructions).
//
// The staging loop itself is flattened: the footprint is a dense list of (row, 4-texel vector) items dealt to the
// producer threads round-robin, so every lane has work whatever the footprint's width (the half-warp-per-row scheme
// left 30-40 % of the lanes idle and re-evaluated bounds per lane and row).
//
// Math: SURVEY.md Appendix A (backward: the scalar form q_l = G_P.S_l + G_A R_l of the canvas behind the layer, so a
// pixel carries one running value instead of four).  Reference semantics: fukuwarai/networks.py:250-257 (warp),
// custom_utils/image_utils.py:128-146 (over), custom/loss_aio.py:251 (range shifts) and their autograd.
#pragma once
#include "render_bwd_tiled.cuh"
#include "render_shift.cuh"
#include "render_tiled.cuh"

#ifndef MGR_WS_WAIT_HINT
#define MGR_WS_WAIT_HINT 0x989680u
#endif
#ifndef MGR_WSF_BLOCKS
#define MGR_WSF_BLOCKS 2
#endif
#ifndef MGR_WSB_BLOCKS
#define MGR_WSB_BLOCKS 2
#endif
#ifndef MGR_WS_PRODW
#define MGR_WS_PRODW 8
#endif
// register budgets of the two roles (setmaxnreg; multiples of 8).  The CTA's pool is what it was launched with:
// 4 producer warps: 384 threads x 80 = 30720 >= 128 x 72 + 256 x 80 = 128 x 64 + 256 x 88;  8 producer warps: 512 x 64 = 32768 = 256 x 48 + 256 x 80.
#ifndef MGR_WSF_PREG
#define MGR_WSF_PREG (MGR_WS_PRODW == 4 ? 72 : 48)
#endif
#ifndef MGR_WSF_CREG
#define MGR_WSF_CREG 80
#endif
#ifndef MGR_WSB_PREG
#define MGR_WSB_PREG (MGR_WS_PRODW == 4 ? 64 : 40)
#endif
#ifndef MGR_WSB_CREG
#define MGR_WSB_CREG 88
#endif

namespace mgr {

constexpr int kConsThreads = kTiledThreads;       // 8 consumer warps: the pixel mapping of the tiled kernels
constexpr int kProdThreads = 32 * MGR_WS_PRODW;   // producer warps in whole warpgroups (setmaxnreg granularity): 4 or 8
constexpr int kWsThreads = kConsThreads + kProdThreads;
constexpr int kMaxStages = 3;
// ring depth (padding included, a slot is 27.5 KB for 16-bit texels and 55 KB for fp32)
#ifndef MGR_WS_STAGES16
#define MGR_WS_STAGES16 2
#endif
template <typename T> struct WsStages { static constexpr int value = sizeof(T) == 4 ? 2 : MGR_WS_STAGES16; };

// ---- mbarrier (shared::cta) -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {       // release.cta: prior shared-memory accesses are ordered before it
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {   // acquire.cta
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"      // suspend-time hint: a waiting warp sleeps until the phase
      "@P1 bra DONE;\n"                                                      // completes instead of polling in the issue slots of
      "bra LAB_WAIT;\n"                                                      // the warps it is waiting for
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(MGR_WS_WAIT_HINT) : "memory");
}
// Register re-balancing between the roles (whole warpgroups: warps 0-7 consume, 8-11 produce).  The kernel is compiled
// for kWsThreads x (65536 / (kWsThreads * CTAs per SM)) registers; producers hand theirs back, consumers take them.
template <int kRegs> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }

struct WsSync {
  uint64_t full[kMaxStages];      // producers arrive (kProdThreads), consumers wait
  uint64_t empty[kMaxStages];     // consumers arrive (kConsThreads), producers wait
};
constexpr size_t kWsSyncBytes = 64;
static_assert(sizeof(WsSync) <= kWsSyncBytes, "ring barriers");

// ---- slot layout and flattened staging ------------------------------------------------------------------------------
// A footprint row is a sequence of ITEMS: kVec texels (all four channels interleaved) = the unit one producer lane loads
// (one vector load per channel plane) and stores.  Items are kVec + kPad texels apart in the slot:
//
//   16-bit storage, rows 16-byte aligned (Geometry::vec8):  kVec = 8, kPad = 2  -> 64 bytes of texels every 80 bytes
//   fp32 storage:                                            kVec = 4, kPad = 1  -> 64 bytes of texels every 80 bytes
//   16-bit storage, any 4-texel alignment:                   kVec = 2, kPad = 0  -> 16-byte items, dense
//
// With a 64-byte lane stride the four 16-byte stores of an item would hit the same four banks in every lane of a
// quarter-warp (ncu: 7.2 M conflict wavefronts out of 10 M store wavefronts in one forward, the L1 / shared pipe the
// busiest unit of the kernel); at 80 bytes the eight lanes of a quarter-warp cover all 32 banks.  Narrow dense items are
// conflict-free too, but cost four times the index / address arithmetic per texel (a third of the kernel's instructions),
// so they only serve tensors whose rows are not 16-byte aligned.  Texel column c of a row lives at unit c + kPad * (c / kVec).
//
// The footprint is dealt to the producer threads as a dense list: item v = vector cv = v % nv of row r = v / nv (nv =
// bw / kVec; the division is a multiply by ceil(2^20 / nv), exact for v * (nv - 1) < 2^20), so every lane has work
// whatever the footprint's width.  A thread issues the loads of kU items before it interleaves and stores the first.
// Items outside the layer's rectangle get the padding value (-1 in m11 mode, 0 in 01 mode: transparent black, which IS
// padding_mode='zeros' after the range shift).
// (Tried and dropped: cp.async of the raw planes into the item's own bytes + an in-place interleave -- every staged byte
//  then crosses shared memory three times: forward +3 % bf16, +48 % fp32; prefetch.global.L2 of the tile's footprints at
//  CTA start: +10 %; a symmetric software pipeline, every warp prefetching the next layer into registers while it samples:
//  122-128 registers, 16 warps per SM, no faster than the two-barrier kernel it replaced.)
constexpr int kSlotUnits = kCapTexels * 5 / 4;      // texel units per slot, padding included

template <typename T, int kVec> struct FlatChunk;      // one channel's kVec texels
template <> struct FlatChunk<float, 4> { using type = float4; };
template <> struct FlatChunk<__nv_bfloat16, 8> { using type = uint4; };
template <> struct FlatChunk<__half, 8> { using type = uint4; };
template <> struct FlatChunk<__nv_bfloat16, 2> { using type = uint32_t; };
template <> struct FlatChunk<__half, 2> { using type = uint32_t; };

__device__ __forceinline__ void interleave_store(uint2* dst, const uint4& r, const uint4& g, const uint4& b, const uint4& a) {
  uint4* d = reinterpret_cast<uint4*>(dst);                   // eight texels: four 16-byte stores of two texels each
  d[0] = interleave2(r.x, g.x, b.x, a.x);
  d[1] = interleave2(r.y, g.y, b.y, a.y);
  d[2] = interleave2(r.z, g.z, b.z, a.z);
  d[3] = interleave2(r.w, g.w, b.w, a.w);
}
__device__ __forceinline__ void interleave_store(uint2* dst, uint32_t r, uint32_t g, uint32_t b, uint32_t a) {
  *reinterpret_cast<uint4*>(dst) = interleave2(r, g, b, a);   // two texels
}
template <int kVec>
__device__ __forceinline__ void fill_texels(uint2* dst, uint32_t o) {
  uint4* d = reinterpret_cast<uint4*>(dst);
  const uint4 v = make_uint4(o, o, o, o);
#pragma unroll
  for (int k = 0; k < kVec / 2; ++k) d[k] = v;
}
template <int kVec>
__device__ __forceinline__ void fill_texels(float4* dst, float o) { fill_store(dst, o); }

template <typename T, int kVec, int kPad, int kU, int kThreads>
__device__ __forceinline__ void stage_flat(bool m11, const SrcView& sv, const LayerPlan& p,
                                           typename Texel<T>::Vec* __restrict__ buf, int ptid) {
  using Chunk = typename FlatChunk<T, kVec>::type;
  const int pitch = p.pitch, nv = p.bw / kVec, V = nv * p.bh;
  const unsigned M = ((1u << 20) + (unsigned)nv - 1u) / (unsigned)nv;
  const int x_org = p.x_lo - sv.left, y_org = p.y_lo - sv.top;
  const unsigned w = (unsigned)sv.w, h = (unsigned)sv.h, rowbytes = sv.rowbytes;
  const size_t plane = sv.plane;
  const char* base = sv.base;
  for (int v0 = ptid; v0 < V; v0 += kThreads * kU) {
    Chunk R[kU], G[kU], Bl[kU], A[kU];
    int so[kU];                    // slot offset in texel units, -1: no item
    bool in[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int v = v0 + u * kThreads;
      const int r = (int)(((unsigned)v * M) >> 20);
      const int cv = v - r * nv;
      const int x = x_org + kVec * cv, y = y_org + r;
      so[u] = v < V ? r * pitch + (kVec + kPad) * cv : -1;
      in[u] = v < V && (unsigned)x < w && (unsigned)y < h;
      if (in[u]) {
        const char* q = base + ((unsigned)y * rowbytes + (unsigned)x * (unsigned)sizeof(T));
        R[u] = __ldg(reinterpret_cast<const Chunk*>(q));
        G[u] = __ldg(reinterpret_cast<const Chunk*>(q + plane));
        Bl[u] = __ldg(reinterpret_cast<const Chunk*>(q + 2 * plane));
        A[u] = __ldg(reinterpret_cast<const Chunk*>(q + 3 * plane));
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (in[u]) {
        interleave_store(buf + so[u], R[u], G[u], Bl[u], A[u]);
      } else if (so[u] >= 0) {
        if constexpr (sizeof(T) == 4) fill_texels<kVec>(buf + so[u], m11 ? -1.f : 0.f);
        else fill_texels<kVec>(buf + so[u], Texel<T>::oob2(m11));
      }
    }
  }
}

// items in flight per producer thread (developer knobs, tools/ab.py): wide items (64 bytes each), narrow items (16 bytes)
#ifndef MGR_WS_KU
#define MGR_WS_KU 1
#endif
#ifndef MGR_WS_KU_NARROW
#define MGR_WS_KU_NARROW 4
#endif

// texel column c of a footprint row -> texel unit inside the row
template <typename T>
__device__ __forceinline__ int slot_unit(int c, bool vec8) {
  if constexpr (sizeof(T) == 4) return c + (c >> 2);
  return vec8 ? c + 2 * (c >> 3) : c;
}
// bilinear sample from explicit tap pointers: t0 / t1 = the two taps of the upper row, the lower row is `pitch` further
template <typename T>
__device__ __forceinline__ Sample sample_taps(const typename Texel<T>::Vec* __restrict__ t0, const typename Texel<T>::Vec* __restrict__ t1,
                                              int pitch, float fx, float fy) {
  f32x2 a_rg, a_ba, b_rg, b_ba, c_rg, c_ba, d_rg, d_ba;
  Texel<T>::unpack(t0[0], a_rg, a_ba);
  Texel<T>::unpack(t1[0], b_rg, b_ba);
  Texel<T>::unpack(t0[pitch], c_rg, c_ba);
  Texel<T>::unpack(t1[pitch], d_rg, d_ba);
  const f32x2 fx2 = bc(fx), fy2 = bc(fy);
  const f32x2 top_rg = fma2(fx2, sub2(b_rg, a_rg), a_rg), top_ba = fma2(fx2, sub2(b_ba, a_ba), a_ba);
  const f32x2 bot_rg = fma2(fx2, sub2(d_rg, c_rg), c_rg), bot_ba = fma2(fx2, sub2(d_ba, c_ba), c_ba);
  Sample s;
  s.rg = fma2(fy2, sub2(bot_rg, top_rg), top_rg);
  s.ba = fma2(fy2, sub2(bot_ba, top_ba), top_ba);
  return s;
}
template <typename T>
__device__ __forceinline__ SampleGrad sample_taps_grad(const typename Texel<T>::Vec* __restrict__ t0,
                                                       const typename Texel<T>::Vec* __restrict__ t1, int pitch, float fx, float fy) {
  f32x2 a_rg, a_ba, b_rg, b_ba, c_rg, c_ba, d_rg, d_ba;
  Texel<T>::unpack(t0[0], a_rg, a_ba);
  Texel<T>::unpack(t1[0], b_rg, b_ba);
  Texel<T>::unpack(t0[pitch], c_rg, c_ba);
  Texel<T>::unpack(t1[pitch], d_rg, d_ba);
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

// the slot's geometry for a footprint of bw x bh texels: row pitch in texel units
template <typename T>
__device__ __forceinline__ int slot_pitch(int bw, bool vec8) {
  if constexpr (sizeof(T) == 4) return bw / 4 * 5;
  return vec8 ? bw / 8 * 10 : bw;
}

// producers: every kStaged layer of the tile goes through the ring, at most WsStages slots ahead of the consumers
template <typename T, bool kRagged, int kThreads>
__device__ __forceinline__ void ws_producer(const T* __restrict__ x, const Geometry& g, const SrcLayers& src, int b,
                                            const LayerPlan* __restrict__ plan, typename Texel<T>::Vec* __restrict__ buf,
                                            WsSync* sy, int ptid, int& n) {
  constexpr int kS = WsStages<T>::value;
  for (int l = 0; l < g.L; ++l) {
    const LayerPlan& p = plan[l];
    if (p.mode != kStaged) continue;
    const int s = n % kS;
    if (n >= kS) mbar_wait(&sy->empty[s], (unsigned)((n / kS - 1) & 1));
    const SrcView sv_ = layer_view<T, kRagged>(x, g, src, b, l);
    typename Texel<T>::Vec* slot = buf + s * kSlotUnits;
    if constexpr (sizeof(T) == 4) {
      stage_flat<T, 4, 1, MGR_WS_KU, kThreads>(g.m11 != 0, sv_, p, slot, ptid);
    } else {
      if (g.vec8) stage_flat<T, 8, 2, MGR_WS_KU, kThreads>(g.m11 != 0, sv_, p, slot, ptid);
      else stage_flat<T, 2, 0, MGR_WS_KU_NARROW, kThreads>(g.m11 != 0, sv_, p, slot, ptid);
    }
    mbar_arrive(&sy->full[s]);
    ++n;
  }
}

// the plans of one tile, computed by the first L producer threads
template <typename T, bool kRagged>
__device__ __forceinline__ void ws_plans(LayerPlan* plan, const float* __restrict__ theta_b, const Geometry& g,
                                         const SrcLayers& src, int j0, int i0, int ptid) {
  for (int l = ptid; l < g.L; l += kProdThreads) {
    LayerPlan p = plan_layer(theta_b + l * 6, g.H, g.W, j0, i0, (sizeof(T) == 2 && g.vec8) ? 8 : kStageVec, layer_rect<kRagged>(g, src, l));
    p.pitch = slot_pitch<T>(p.bw, g.vec8 != 0);
    plan[l] = p;
  }
}

__host__ __device__ constexpr size_t ws_ring_bytes(size_t vec_bytes) {
  return kWsSyncBytes + vec_bytes * kSlotUnits * (vec_bytes == 16 ? 2 : MGR_WS_STAGES16);      // == WsStages<T>::value slots
}
inline size_t ws_fwd_smem_bytes(int L, size_t vec_bytes) { return ws_ring_bytes(vec_bytes) + sizeof(LayerPlan) * L; }

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
// every lane of the warp learns whether all L <= 32 layers of the sample are pure translations (no CTA barrier)
__device__ __forceinline__ bool warp_all_shift(const float* __restrict__ theta_b, int L, int lane) {
  return __all_sync(0xffffffffu, lane >= L || is_pure_shift(theta_b + 6 * lane));
}

// one tile's consumer work: sample the staged layers back to front, keep the canvas in registers, write the pixels
template <typename T, bool kSave, bool kRagged>
__device__ __forceinline__ void fwd_consume_tile(const T* __restrict__ x, const SrcLayers& src, T* __restrict__ out,
                                                 typename SavedAlpha<T>::type* __restrict__ sav, const Geometry& g, int b, int j0, int i0,
                                                 const LayerPlan* __restrict__ plan, const typename Texel<T>::Vec* __restrict__ buf,
                                                 WsSync* sy, int tid, int& n) {
  using Vec = typename Texel<T>::Vec;
  using SA = typename SavedAlpha<T>::type;
  const int tx = tid & 31, ty = tid >> 5;
  const bool vec8 = g.vec8 != 0;
  const f32x2 zs2 = bc(g.m11 ? 0.5f : 1.f), zb2 = bc(g.m11 ? 0.5f : 0.f);     // z = zs * raw + zb
  const int hw = g.H * g.W;                                   // one plane fits 32 bits (host-checked)
  const int j = j0 + tx;
  const int pix0 = (i0 + ty) * g.W + j;                       // pixel k lives 8*k rows further down
  const int row8 = kRowStep * g.W;
  unsigned live = 0;
#pragma unroll
  for (int k = 0; k < kPx; ++k) live |= (j < g.W && i0 + ty + kRowStep * k < g.H) ? (1u << k) : 0u;
  const float djf = (float)(tx - kTW / 2), dif0 = (float)(ty - kTH / 2);
  float S0[kPx], S1[kPx], S2[kPx], R[kPx];
#pragma unroll
  for (int k = 0; k < kPx; ++k) S0[k] = S1[k] = S2[k] = R[k] = 0.f;
  SA* svl = kSave ? sav + (long long)b * g.L * hw + pix0 : nullptr;     // this thread's pixel 0 of layer l (bumped per layer)

  for (int l = 0; l < g.L; ++l, svl += hw) {
    const LayerPlan& p = plan[l];
    const int mode = p.mode;
    if (mode == kSkip) {                     // fully transparent layer: the canvas is unchanged
      if (kSave) {
#pragma unroll
        for (int k = 0; k < kPx; ++k)
          if (live & (1u << k)) st_alpha(svl + k * row8, 0.f);
      }
      continue;
    }
    constexpr int kS = WsStages<T>::value;
    const int s = n % kS;
    const Vec* bufs = buf + s * kSlotUnits;
    if (mode == kStaged) mbar_wait(&sy->full[s], (unsigned)((n / kS) & 1));
    const float a01 = p.aff.a01, a11 = p.aff.a11;
    // coordinates relative to the tile centre (the same numbers whatever the footprint's alignment: a ragged stack and
    // its padded canvas stage different rectangles but sample identical bits); (dX, dY) moves the tap into the footprint
    const float bx = fmaf(a01, dif0, fmaf(p.aff.a00, djf, p.aff.rx)), by = fmaf(a11, dif0, fmaf(p.aff.a10, djf, p.aff.ry));
    const int pitch = p.pitch;
    const Vec* bufo = bufs + p.dY * pitch;
    const int dX = p.dX;
    float al[kPx];
    auto over = [&](int k, float r_, float g_, float b_, float a) {       // the canvas takes one more layer (straight alpha)
      al[k] = a;
      const float om = 1.f - a;
      S0[k] = fmaf(om, S0[k], a * r_);
      S1[k] = fmaf(om, S1[k], a * g_);
      S2[k] = fmaf(om, S2[k], a * b_);
      R[k] = fmaf(om, R[k], a);
    };
    // the mode is tested OUTSIDE the pixel loop: the staged path is one straight-line block in which the four pixels'
    // address / load / lerp chains interleave (a branch per pixel kept them apart: one dependent chain per warp)
    if (mode == kStaged) {
#pragma unroll
      for (int k = 0; k < kPx; ++k) {
        const float ix = fmaf(a01, (float)(kRowStep * k), bx), iy = fmaf(a11, (float)(kRowStep * k), by);
        const float fxf = floorf(ix), fyf = floorf(iy);
        const int cx = (int)fxf + dX;                             // tap column inside the footprint
        const Vec* row = bufo + (int)fyf * pitch;
        const Sample sm = sample_taps<T>(row + slot_unit<T>(cx, vec8), row + slot_unit<T>(cx + 1, vec8), pitch, ix - fxf, iy - fyf);
        float r_, g_, b_, a;
        upk(fma2(sm.rg, zs2, zb2), r_, g_);
        upk(fma2(sm.ba, zs2, zb2), b_, a);
        over(k, r_, g_, b_, a);
      }
      mbar_arrive(&sy->empty[s]);                               // the slot's texels are in registers
      ++n;
    } else {
      const SrcView sv_ = layer_view<T, kRagged>(x, g, src, b, l);
#pragma unroll
      for (int k = 0; k < kPx; ++k) {
        const float4 z = sample_pixel_direct<T>(reinterpret_cast<const T*>(sv_.base), p.aff, tx - kTW / 2,
                                                ty + kRowStep * k - kTH / 2, sv_.h, sv_.w, sv_.rowbytes / sizeof(T), sv_.plane / sizeof(T),
                                                g.m11 ? 1.f : 0.f, g.m11 ? 0.5f : 1.f);
        over(k, z.x, z.y, z.z, z.w);
      }
    }
    if (kSave) {
#pragma unroll
      for (int k = 0; k < kPx; ++k)
        if (live & (1u << k)) st_alpha(svl + k * row8, al[k]);
    }
  }

  const float os = g.m11 ? 2.f : 1.f, obias = g.m11 ? -1.f : 0.f;   // out = os * o + obias
  T* outp = out + (long long)b * 4 * hw + pix0;
#pragma unroll
  for (int k = 0; k < kPx; ++k) {
    if (live & (1u << k)) {
      const float inv = (R[k] != 0.f) ? 1.f / R[k] : 0.f;     // nan_to_num(0/0) = 0 (image_utils.py:132)
      T* q = outp + k * row8;
      st(q, fmaf(S0[k] * inv, os, obias));
      st(q + hw, fmaf(S1[k] * inv, os, obias));
      st(q + 2 * hw, fmaf(S2[k] * inv, os, obias));
      st(q + 3 * hw, fmaf(R[k], os, obias));
    }
  }
}

template <typename T, bool kSave, bool kRagged>
__global__ void __launch_bounds__(kWsThreads, MGR_WSF_BLOCKS)
render_fwd_ws(const T* __restrict__ x, const __grid_constant__ SrcLayers src, const float* __restrict__ theta, T* __restrict__ out,
              typename SavedAlpha<T>::type* __restrict__ sav, Geometry g, int skip_all_shift, const int* __restrict__ shift_flags) {
  using Vec = typename Texel<T>::Vec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const float* theta_b = theta + (long long)b * g.L * 6;
  // a stack of pure translations belongs to the stencil kernel (render_fwd_stencil_only), launched next to this one:
  // one load of the per-sample flag when the forward wrote it (it does when there is a saved-alpha buffer to hold it),
  // else every warp finds out by itself from the placements -- either way without a CTA barrier
  if (skip_all_shift && (shift_flags ? shift_flags[b] != 0 : warp_all_shift(theta_b, g.L, tid & 31))) return;
  WsSync* sy = reinterpret_cast<WsSync*>(smem_raw);
  Vec* buf = reinterpret_cast<Vec*>(smem_raw + kWsSyncBytes);                               // [WsStages][kSlotUnits]
  LayerPlan* plan = reinterpret_cast<LayerPlan*>(smem_raw + ws_ring_bytes(sizeof(Vec)));    // [L]
  const int j0 = blockIdx.x * kTW, i0 = blockIdx.y * kTH;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&sy->full[s], kProdThreads); mbar_init(&sy->empty[s], kConsThreads); }
  }
  if (tid >= kConsThreads) ws_plans<T, kRagged>(plan, theta_b, g, src, j0, i0, tid - kConsThreads);
  __syncthreads();
  int n = 0;
  if (tid >= kConsThreads) {
    setmaxnreg_dec<MGR_WSF_PREG>();
    ws_producer<T, kRagged, kProdThreads>(x, g, src, b, plan, buf, sy, tid - kConsThreads, n);
    return;
  }
  setmaxnreg_inc<MGR_WSF_CREG>();
  fwd_consume_tile<T, kSave, kRagged>(x, src, out, sav, g, b, j0, i0, plan, buf, sy, tid, n);
}

// ---------------------------------------------------------------------------------------------------------------
// backward, pass 1: composite adjoint + theta gradient per output tile, gradient records for pass 2
// ---------------------------------------------------------------------------------------------------------------
// Per pixel, with (G_P, G_A) the upstream gradient in the compositing domain:
//   u_l = G_P . c_l + G_A              q_0 = 0,  q_{l+1} = q_l + a_l (u_l - q_l)      (q_l = G_P . S_l + G_A R_l)
//   d a_l = T_l (u_l - q_l)            d c_l = G_P T_l a_l
// -- the same numbers as T_l [G_P . (c_l - S_l) + G_A (1 - R_l)] with one running scalar instead of the four of (S_l, R_l).
// Shared memory: [ring barriers][2 ring slots][plans][T_l stash, later theta partials: L x 4 x 256 floats][(G_P, G_A): 4 x 256 float4]
inline size_t ws_bwd_smem_bytes(int L, size_t vec_bytes, bool gp_smem) {
  return align16(ws_ring_bytes(vec_bytes) + sizeof(LayerPlan) * L) + sizeof(float) * (size_t)L * kPx * kConsThreads +
         (gp_smem ? sizeof(float4) * kPx * kConsThreads : 0);
}

template <typename T, bool kNeedTheta, bool kGPSmem, bool kRagged>
__global__ void __launch_bounds__(kWsThreads, MGR_WSB_BLOCKS)
render_bwd_pass1_ws(const T* __restrict__ x, const __grid_constant__ SrcLayers src, const float* __restrict__ theta,
                    const T* __restrict__ out, const T* __restrict__ gout, const typename SavedAlpha<T>::type* __restrict__ sav,
                    float2* __restrict__ rec, float4* __restrict__ gp, float* __restrict__ gtheta, Geometry g,
                    const int* __restrict__ sample_all_shift, int skip_shift) {
  using Vec = typename Texel<T>::Vec;
  using SA = typename SavedAlpha<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  if (skip_shift && sample_all_shift[b]) return;            // render_bwd_shift's sample (flag from the placement kernels)
  WsSync* sy = reinterpret_cast<WsSync*>(smem_raw);
  Vec* buf = reinterpret_cast<Vec*>(smem_raw + kWsSyncBytes);
  LayerPlan* plan = reinterpret_cast<LayerPlan*>(smem_raw + ws_ring_bytes(sizeof(Vec)));
  float* stash = reinterpret_cast<float*>(smem_raw + align16(ws_ring_bytes(sizeof(Vec)) + sizeof(LayerPlan) * g.L));
  const int j0 = blockIdx.x * kTW, i0 = blockIdx.y * kTH;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&sy->full[s], kProdThreads); mbar_init(&sy->empty[s], kConsThreads); }
  }
  if (tid >= kConsThreads) ws_plans<T, kRagged>(plan, theta + (long long)b * g.L * 6, g, src, j0, i0, tid - kConsThreads);
  __syncthreads();
  if (tid >= kConsThreads) {
    setmaxnreg_dec<MGR_WSB_PREG>();
    int np = 0;
    ws_producer<T, kRagged, kProdThreads>(x, g, src, b, plan, buf, sy, tid - kConsThreads, np);
    return;
  }
  setmaxnreg_inc<MGR_WSB_CREG>();

  // ---- consumers ----
  float* Tst = stash + tid;                                   // [L][kPx][256]: T_l, later this thread's theta partials
  float4* GPs = reinterpret_cast<float4*>(stash + (size_t)g.L * kPx * kConsThreads) + tid;   // [kPx][256]
  const int tx = tid & 31, ty = tid >> 5;
  const bool vec8 = g.vec8 != 0;
  const float zs = g.m11 ? 0.5f : 1.f;
  const f32x2 zs2 = bc(zs), zb2 = bc(g.m11 ? 0.5f : 0.f);     // z = zs * raw + zb
  const int hw = g.H * g.W;
  const int j = j0 + tx;
  const int pix0 = (i0 + ty) * g.W + j;                       // pixel k lives 8*k rows further down
  const int row8 = kRowStep * g.W;
  unsigned live = 0;
#pragma unroll
  for (int k = 0; k < kPx; ++k) live |= (j < g.W && i0 + ty + kRowStep * k < g.H) ? (1u << k) : 0u;
  float4* gpp = gp + (long long)b * hw + pix0;                 // (G_P, G_A) of this thread's pixels: gpp[k * row8]

  // ---- pre-pass: T_l from the saved alphas (front -> back), A, then (G_P, G_A) -----------------------------------
  float4 G4r[kGPSmem ? 1 : kPx];                               // registers when the shared copy does not fit
  {
    float gv[kPx][4], ov[kPx][3];                              // issued first: their latency hides behind the alpha sweep
    const T* gob = gout + (long long)b * 4 * hw + pix0;
    const T* ob_ = out + (long long)b * 4 * hw + pix0;
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
      const bool lv = live & (1u << k);
#pragma unroll
      for (int c = 0; c < 4; ++c) gv[k][c] = lv ? ld(gob + k * row8 + c * hw) : 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) ov[k][c] = lv ? ld(ob_ + k * row8 + c * hw) : 0.f;
    }
    float Tc[kPx], A[kPx];
#pragma unroll
    for (int k = 0; k < kPx; ++k) { Tc[k] = 1.f; A[k] = 0.f; }
    const SA* sa = sav + ((long long)b * g.L + (g.L - 1)) * hw + pix0;     // front layer first
    float* tp = Tst + (g.L - 1) * kPx * kConsThreads;
    for (int l = g.L - 1; l >= 0; --l, sa -= hw, tp -= kPx * kConsThreads) {
#pragma unroll
      for (int k = 0; k < kPx; ++k) {
        const bool lv = live & (1u << k);
        tp[k * kConsThreads] = lv ? Tc[k] : 0.f;
        const float a = lv ? ld_alpha(sa + k * row8) : 0.f;
        A[k] = fmaf(Tc[k], a, A[k]);
        Tc[k] *= (1.f - a);
      }
    }
    const float gs = g.m11 ? 2.f : 1.f;                       // d out / d o
    const float is = g.m11 ? 0.5f : 1.f, ib = g.m11 ? 0.5f : 0.f;
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
      float4 G4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((live & (1u << k)) && A[k] != 0.f) {                // A == 0: every gradient is defined as 0
        const float g0 = gs * gv[k][0], g1 = gs * gv[k][1], g2 = gs * gv[k][2], g3 = gs * gv[k][3];
        const float inv = 1.f / A[k];
        const float o0 = fmaf(ov[k][0], is, ib), o1 = fmaf(ov[k][1], is, ib), o2 = fmaf(ov[k][2], is, ib);
        G4 = make_float4(g0 * inv, g1 * inv, g2 * inv, g3 - (g0 * o0 + g1 * o1 + g2 * o2) * inv);
      }
      if (live & (1u << k)) gpp[k * row8] = G4;               // pass 2 reads it
      if (kGPSmem) GPs[k * kConsThreads] = G4; else G4r[k] = G4;
    }
  }

  // ---- back -> front sweep ------------------------------------------------------------------------------------------
  const float djf = (float)(tx - kTW / 2), dif0 = (float)(ty - kTH / 2);
  float q[kPx];
#pragma unroll
  for (int k = 0; k < kPx; ++k) q[k] = 0.f;
  float2* rl = rec + (long long)b * g.L * hw + pix0;          // this thread's pixel 0 of layer l (bumped per layer)
  float yi[kPx];
#pragma unroll
  for (int k = 0; k < kPx; ++k) yi[k] = norm_coord(i0 + ty + kRowStep * k, g.H);

  int n = 0;
  for (int l = 0; l < g.L; ++l, rl += hw) {
    const LayerPlan& p = plan[l];
    const int mode = p.mode;
    float* Tl = Tst + l * kPx * kConsThreads;
    if (mode == kSkip) {                     // a = 0, c = 0 (transparent black): u = G_A, q unchanged, no colour gradient
#pragma unroll
      for (int k = 0; k < kPx; ++k) {
        if (live & (1u << k)) {
          const float4 G4 = kGPSmem ? GPs[k * kConsThreads] : G4r[kGPSmem ? 0 : k];
          rl[k * row8] = make_float2(0.f, Tl[k * kConsThreads] * (G4.w - q[k]));
        }
      }
      if (kNeedTheta) park_theta_partials(Tl, 0.f, 0.f, 0.f, 0.f);
      continue;                              // the footprint misses the layer: no texel, no theta gradient
    }
    constexpr int kS = WsStages<T>::value;
    const int s = n % kS;
    const Vec* bufs = buf + s * kSlotUnits;
    if (mode == kStaged) mbar_wait(&sy->full[s], (unsigned)((n / kS) & 1));
    const float a01 = p.aff.a01, a11 = p.aff.a11;
    // coordinates relative to the tile centre (the same numbers whatever the footprint's alignment: a ragged stack and
    // its padded canvas stage different rectangles but sample identical bits); (dX, dY) moves the tap into the footprint
    const float bx = fmaf(a01, dif0, fmaf(p.aff.a00, djf, p.aff.rx)), by = fmaf(a11, dif0, fmaf(p.aff.a10, djf, p.aff.ry));
    const int pitch = p.pitch;
    const Vec* bufo = bufs + p.dY * pitch;
    const int dX = p.dX;
    float accx = 0.f, accxy = 0.f, accy = 0.f, accyy = 0.f;
    // composite adjoint of pixel k given the layer's sample (compositing domain) and its raw derivatives
    auto adjoint = [&](int k, float r_, float g_, float b_, float a, float dxr, float dxg, float dxb, float dxa,
                       float dyr, float dyg, float dyb, float dya) {
      const float T_l = Tl[k * kConsThreads];
      const float4 G4 = kGPSmem ? GPs[k * kConsThreads] : G4r[kGPSmem ? 0 : k];
      const float u = fmaf(G4.x, r_, fmaf(G4.y, g_, fmaf(G4.z, b_, G4.w)));
      const float d = u - q[k];
      const float ta = T_l * a, ga = T_l * d;
      q[k] = fmaf(a, d, q[k]);
      if (live & (1u << k)) rl[k * row8] = make_float2(ta, ga);
      if (kNeedTheta) {
        const float gr = G4.x * ta, gg = G4.y * ta, gb = G4.z * ta;
        const float dix = fmaf(gr, dxr, fmaf(gg, dxg, fmaf(gb, dxb, ga * dxa)));
        const float diy = fmaf(gr, dyr, fmaf(gg, dyg, fmaf(gb, dyb, ga * dya)));
        accx += dix; accxy = fmaf(dix, yi[k], accxy);
        accy += diy; accyy = fmaf(diy, yi[k], accyy);
      }
    };
    // the mode is tested OUTSIDE the pixel loop (see render_fwd_ws): four independent chains per thread
    if (mode == kStaged) {
#pragma unroll
      for (int k = 0; k < kPx; ++k) {
        const float ix = fmaf(a01, (float)(kRowStep * k), bx), iy = fmaf(a11, (float)(kRowStep * k), by);
        const float fxf = floorf(ix), fyf = floorf(iy);
        const int cx = (int)fxf + dX;                             // tap column inside the footprint
        const Vec* row = bufo + (int)fyf * pitch;
        const SampleGrad sg = sample_taps_grad<T>(row + slot_unit<T>(cx, vec8), row + slot_unit<T>(cx + 1, vec8), pitch, ix - fxf, iy - fyf);
        float r_, g_, b_, a, dxr, dxg, dxb, dxa, dyr, dyg, dyb, dya;
        upk(fma2(sg.rg, zs2, zb2), r_, g_);
        upk(fma2(sg.ba, zs2, zb2), b_, a);
        upk(sg.dx_rg, dxr, dxg); upk(sg.dx_ba, dxb, dxa);
        upk(sg.dy_rg, dyr, dyg); upk(sg.dy_ba, dyb, dya);
        adjoint(k, r_, g_, b_, a, dxr, dxg, dxb, dxa, dyr, dyg, dyb, dya);
      }
      mbar_arrive(&sy->empty[s]);                               // the slot's texels are in registers
      ++n;
    } else {
      // huge footprint: bounds-checked taps straight from global memory
      const SrcView sv_ = layer_view<T, kRagged>(x, g, src, b, l);
#pragma unroll
      for (int k = 0; k < kPx; ++k) {
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
        adjoint(k, zz[0], zz[1], zz[2], zz[3],
                (v[0][1] - v[0][0]) * ey + (v[0][3] - v[0][2]) * tp.fy, (v[1][1] - v[1][0]) * ey + (v[1][3] - v[1][2]) * tp.fy,
                (v[2][1] - v[2][0]) * ey + (v[2][3] - v[2][2]) * tp.fy, (v[3][1] - v[3][0]) * ey + (v[3][3] - v[3][2]) * tp.fy,
                (v[0][2] - v[0][0]) * ex + (v[0][3] - v[0][1]) * tp.fx, (v[1][2] - v[1][0]) * ex + (v[1][3] - v[1][1]) * tp.fx,
                (v[2][2] - v[2][0]) * ex + (v[2][3] - v[2][1]) * tp.fx, (v[3][2] - v[3][0]) * ex + (v[3][3] - v[3][1]) * tp.fx);
      }
    }
    // the T_l slots of this layer are dead: park the thread's theta-gradient partials there (tile_common.cuh)
    if (kNeedTheta) park_theta_partials(Tl, accx, accxy, accy, accyy);
  }
  if (kNeedTheta) {
    // consumers only: the producers may have left already (named barrier 1, 256 threads)
    asm volatile("bar.sync 1, %0;" ::"n"(kConsThreads) : "memory");
    const float hW = 0.5f * (float)g.W * zs, hH = 0.5f * (float)g.H * zs;   // d ix / d gx (and the range scale)
    reduce_theta_partials(stash, g.L, tid, norm_coord(j, g.W), hW, hH, gtheta + (long long)b * g.L * 6);
  }
}

}  // namespace mgr
