// Typed launchers: compiled once per storage dtype (inst_*.cu) so the dtypes build in parallel.
#pragma once
#include "mgr_errors.h"
#include "render_direct.cuh"
#include "render_bwd_tiled.cuh"
#include "render_shift.cuh"
#include "render_tiled.cuh"
#include "render_fwd.cuh"
#include "warp_ops.cuh"
#include "warp_tiled.cuh"
#include "pil_composite.cuh"
#include "composite_only.cuh"
#include "render_ws.cuh"
#include "render_shift_tma.cuh"
#include "render_shift_tma_bwd.cuh"

#include <cstdlib>
#include <mutex>
#include <map>
#include <utility>

#ifndef MGR_STB_SMEM_CAP
#define MGR_STB_SMEM_CAP (100 * 1024)      // the stencil backward's tile state must leave room for two CTAs per SM
#endif

namespace mgr {

// Opt a kernel in to more than 48 KB of dynamic shared memory when a launch first needs it (and again only if a later
// launch needs more) instead of on every launch: cudaFuncSetAttribute is a driver call of a few microseconds.
template <typename Kern>
int ensure_dynamic_smem(Kern kern, size_t bytes) {
  if (bytes <= 40 * 1024) return MGR_OK;          // static shared memory counts against the default 48 KB too
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> granted;
  int dev = 0;
  MGR_CUDA(cudaGetDevice(&dev));
  const std::pair<const void*, int> key(reinterpret_cast<const void*>(kern), dev);
  std::lock_guard<std::mutex> lock(mu);
  auto it = granted.find(key);
  if (it != granted.end() && it->second >= bytes) return MGR_OK;
  MGR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  granted[key] = bytes;
  return MGR_OK;
}

// The tiled kernels stage footprints with 128-bit vector loads (4 fp32 / 8 16-bit texels): every
// plane row must start on a vector boundary and W must be a multiple of the vector width (a lane's
// texels are then all inside or all outside the image).
template <typename T>
bool tiled_ok(const void* x, const Geometry& g) {
  if (debug_path() == 1) return false;
  const int v = kStageVec;
  const long long layer_bytes = (3 * g.sc + (long long)g.H * g.sh) * (long long)sizeof(T);   // 32-bit offsets inside a layer
  return (g.L <= kMaxTiledLayers) && (g.W % v == 0) && (g.sh % v == 0) && (g.sc % v == 0) && (g.sl % v == 0) && (g.sb % v == 0) &&
         (reinterpret_cast<uintptr_t>(x) % (kStageVec * sizeof(T)) == 0) && g.sc >= 0 && (long long)g.H * g.W < (1LL << 29) && layer_bytes < (1LL << 31);
}

// four adjacent pixels per thread (composite_only.cuh): rows, planes and base 4-element aligned
template <typename T>
bool vec4_ok(const void* x, const Geometry& g) {
  return g.W % 4 == 0 && g.sh % 4 == 0 && g.sc % 4 == 0 && g.sl % 4 == 0 && g.sb % 4 == 0 &&
         reinterpret_cast<uintptr_t>(x) % (4 * sizeof(T)) == 0;
}

// the canvas layout x[B,L,4,H,W] / grad_x[B,L,4,H,W] as per-layer descriptors (every layer covers the whole canvas)
template <typename T>
SrcLayers canvas_src(const void* x, const Geometry& g) {
  SrcLayers r{};
  for (int l = 0; l < g.L && l < kMaxTiledLayers; ++l)
    r.s[l] = SrcLayer{reinterpret_cast<const T*>(x) + (long long)l * g.sl, g.sb, g.sc, g.sh, g.H, g.W, 0, 0};
  return r;
}
template <typename T>
DstLayers canvas_dst(void* gx, const Geometry& g) {
  DstLayers r{};
  const long long hw = (long long)g.H * g.W;
  for (int l = 0; l < g.L && l < kMaxTiledLayers; ++l)
    r.s[l] = DstLayer{gx ? reinterpret_cast<T*>(gx) + (long long)l * 4 * hw : nullptr, (long long)g.L * 4 * hw, hw, g.W, g.H, g.W, 0, 0};
  return r;
}

// Ragged stacks (one tensor per layer) only exist on the tiled path: every rectangle must satisfy
// the vector-staging rules the canvas layout satisfies by construction.
template <typename T>
const char* ragged_problem(const SrcLayers& src, const DstLayers* dst, const Geometry& g) {
  const int v = kStageVec;
  if (g.L < 2 || g.L > kMaxTiledLayers) return "a ragged stack needs 2..32 layers";
  if (g.W % v || (long long)g.H * g.W >= (1LL << 29)) return "canvas width must be a multiple of 4 (and H*W < 2^29)";
  for (int l = 0; l < g.L; ++l) {
    const SrcLayer& a = src.s[l];
    if (!a.ptr) return "a layer pointer is NULL";
    if (a.h < 1 || a.w < 1 || a.top < 0 || a.left < 0 || a.top + a.h > g.H || a.left + a.w > g.W) return "a layer rectangle leaves the canvas";
    if (a.w % v || a.left % v) return "layer width and left offset must be multiples of 4";
    if (a.sh % v || a.sc % v || a.sb % v || a.sc < 0 || a.sh < a.w) return "layer strides must be non-negative multiples of 4 elements";
    if (reinterpret_cast<uintptr_t>(a.ptr) % (v * sizeof(T))) return "layer base pointer must be aligned to 4 elements";
    if ((3 * a.sc + (long long)a.h * a.sh) * (long long)sizeof(T) >= (1LL << 31)) return "one layer of one sample must span < 2 GiB";
    if (dst) {
      const DstLayer& d = dst->s[l];
      if (!d.ptr) return "a grad pointer is NULL";
      if (d.h != a.h || d.w != a.w || d.top != a.top || d.left != a.left) return "grad rectangle differs from the layer's";
      if (d.sh % 2 || d.sc % 2 || d.sb % 2 || reinterpret_cast<uintptr_t>(d.ptr) % (2 * sizeof(T))) return "grad strides / pointer must be even";
      if (3 * d.sc + (long long)d.h * d.sh >= (1LL << 31)) return "one grad layer of one sample must span < 2^31 elements";
    }
  }
  return nullptr;
}

// fork the general kernels onto the side stream?  Only where an idle launch costs more than the fork / join (four runtime
// calls): from a million layer-pixels (measured: at 3.7 M -- C1 -- translations gain 10 %, general placements lose nothing; at 0.9 M
// general placements lose 7 %); MGR_NO_SIDE_STREAM=1 in the environment keeps everything on one stream
inline bool use_side_stream(const Geometry& g) {
  static const bool off = [] { const char* e = getenv("MGR_NO_SIDE_STREAM"); return e && e[0] == '1'; }();
  static const long long min_px = [] { const char* e = getenv("MGR_SIDE_STREAM_MIN"); return e ? atoll(e) : (1LL << 20); }();   // developer knob
  return !off && (long long)g.B * g.L * g.H * g.W >= min_px;
}

// joins on every exit path (an early error return must not leave the side stream dangling off `s`, least of all in a capture)
struct ForkGuard {
  SideStream ss{};
  cudaStream_t s = nullptr;
  bool active = false;
  int fork(cudaStream_t main) {
    if (int rc = side_stream(&ss)) return rc;
    if (int rc = side_fork(ss, main)) return rc;
    s = main; active = true;
    return MGR_OK;
  }
  cudaStream_t side_or(cudaStream_t main) const { return active ? ss.side : main; }
  int join() {
    if (!active) return MGR_OK;
    active = false;
    return side_join(ss, s);
  }
  ~ForkGuard() { if (active) (void)side_join(ss, s); }
};

// 16-bit footprints are staged in 8-texel items (128-bit loads) when every row of every layer starts on a 16-byte
// boundary and the rectangles are multiples of 8 texels (render_ws.cuh: stage_flat); narrow items otherwise
template <typename T>
Geometry with_vec8(const SrcLayers& src, const Geometry& g) {
  Geometry r = g;
  r.vec8 = sizeof(T) == 2 && g.W % 8 == 0;
  for (int l = 0; r.vec8 && l < g.L && l < kMaxTiledLayers; ++l) {
    const SrcLayer& a = src.s[l];
    if (a.w % 8 || a.left % 8 || a.sh % 8 || a.sc % 8 || a.sb % 8 || reinterpret_cast<uintptr_t>(a.ptr) % 16) r.vec8 = 0;
  }
  return r;
}

template <typename T, bool kRagged>
int launch_forward_tiled(const void* x, const SrcLayers& src, const float* theta, void* out, void* sav, const mgr::Geometry& g,
                         cudaStream_t s) {
  using Vec = typename Texel<T>::Vec;
  const size_t smem_a = tiled_smem_bytes(g.L, sizeof(Vec)), smem_b = shift_fwd_smem_bytes(g.L, sizeof(Vec));
  const size_t smem = smem_a > smem_b ? smem_a : smem_b;      // one launch serves both code paths (render_fwd.cuh)
  // the forward kernels stay inside the default 48 KB of dynamic shared memory (fp32, L = 32: 45 056 + 2 048 bytes), so no
  // per-function, per-device opt-in is needed
  static_assert(sizeof(typename Texel<float>::Vec) * kCapTexels + sizeof(LayerPlan) * kMaxTiledLayers <= 48 * 1024, "forward smem");
  dim3 grid((g.W + kTW - 1) / kTW, (g.H + kTH - 1) / kTH, g.B);
  using SA = typename SavedAlpha<T>::type;
  const int stencil = debug_path() != 2;        // all-translation samples take the stencil path
  if (debug_path() != 3) {
    // warp-specialised general path (render_ws.cuh) + the stencil kernel for all-translation samples; every CTA of
    // either launch reads its sample's placements and leaves at once if the sample is the other kernel's
    const size_t smem_w = ws_fwd_smem_bytes(g.L, sizeof(Vec));
    // per-sample "all translations" flags in the tail of the saved-alpha buffer (mgr_saved_alpha_bytes reserves it)
    int* flags = nullptr;
    if (stencil && sav) {
      flags = reinterpret_cast<int*>(reinterpret_cast<char*>(sav) + saved_alpha_flags_offset(g.B, g.L, g.H, g.W, sizeof(SA)));
      sample_shift_flags_kernel<<<(g.B + 7) / 8, 256, 0, s>>>(theta, g.B, g.L, flags);
      MGR_CUDA(cudaGetLastError());
      count_launch();
    }
    // the two families partition the batch: the general kernel goes to the side stream, the stencil kernel stays on `s`
    // (big batches only: the fork / join is four runtime calls)
    ForkGuard fg;
    if (stencil && use_side_stream(g))
      if (int rc = fg.fork(s)) return rc;
    cudaStream_t sg = fg.side_or(s);
    auto launch = [&](auto kern) -> int {
      if (int rc = ensure_dynamic_smem(kern, smem_w)) return rc;
      kern<<<grid, kWsThreads, smem_w, sg>>>((const T*)x, src, theta, (T*)out, (SA*)sav, with_vec8<T>(src, g), stencil, flags);
      return MGR_OK;
    };
    if (int rc = sav ? launch(render_fwd_ws<T, true, kRagged>) : launch(render_fwd_ws<T, false, kRagged>)) return rc;
    if (stencil) {
      MGR_CUDA(cudaGetLastError());
      count_launch();
      // canvas layout that meets TMA's alignment rules: box copies into a ring, planar stencil (render_shift_tma.cuh);
      // ragged stacks and odd strides keep the staged stencil kernel
      CUtensorMap xmap;
      if (!kRagged && debug_path() != 4 && shift_tma_x_map<T>(&xmap, x, g, ShiftBox<T>::W, ShiftBox<T>::H)) {
        const size_t smem_t = shift_tma_fwd_smem_bytes<T>(g.L);
        dim3 gridt((g.W + kSW - 1) / kSW, (g.H + kSH - 1) / kSH, g.B);
        auto launch_t = [&](auto kern) -> int {
          if (int rc = ensure_dynamic_smem(kern, smem_t)) return rc;
          kern<<<gridt, kSThreads, smem_t, s>>>(xmap, theta, (T*)out, (SA*)sav, g, flags);
          return MGR_OK;
        };
        if (int rc = sav ? launch_t(render_fwd_shift_tma<T, true>) : launch_t(render_fwd_shift_tma<T, false>)) return rc;
      } else if (sav) render_fwd_stencil_only<T, true><<<grid, kTiledThreads, smem_b, s>>>(src, theta, (T*)out, (SA*)sav, g, flags);
      else render_fwd_stencil_only<T, false><<<grid, kTiledThreads, smem_b, s>>>(src, theta, (T*)out, nullptr, g, flags);
    }
    MGR_CUDA(cudaGetLastError());
    count_launch();
    return fg.join();                                       // both kernels are queued: `s` continues when the side stream is done too
  }
  if constexpr (sizeof(T) == 4) {               // fp32: one launch per path (see render_fwd.cuh)
    if (sav) render_fwd_general_only<T, true, kRagged><<<grid, kTiledThreads, smem_a, s>>>((const T*)x, src, theta, (T*)out, (SA*)sav, g, stencil);
    else render_fwd_general_only<T, false, kRagged><<<grid, kTiledThreads, smem_a, s>>>((const T*)x, src, theta, (T*)out, nullptr, g, stencil);
    if (stencil) {
      MGR_CUDA(cudaGetLastError());
      count_launch();
      if (sav) render_fwd_stencil_only<T, true><<<grid, kTiledThreads, smem_b, s>>>(src, theta, (T*)out, (SA*)sav, g, nullptr);
      else render_fwd_stencil_only<T, false><<<grid, kTiledThreads, smem_b, s>>>(src, theta, (T*)out, nullptr, g, nullptr);
    }
  } else {
    if (sav) render_fwd<T, true, kRagged><<<grid, kTiledThreads, smem, s>>>((const T*)x, src, theta, (T*)out, (SA*)sav, g, stencil);
    else render_fwd<T, false, kRagged><<<grid, kTiledThreads, smem, s>>>((const T*)x, src, theta, (T*)out, nullptr, g, stencil);
  }
  MGR_CUDA(cudaGetLastError());
  count_launch();
  return MGR_OK;
}

template <typename T>
int launch_forward(const void* x, const float* theta, void* out, void* sav, const mgr::Geometry& g, cudaStream_t s) {
  if (theta && g.L >= 2 && tiled_ok<T>(x, g)) return launch_forward_tiled<T, false>(x, canvas_src<T>(x, g), theta, out, sav, g, s);
  if (!theta && g.L >= 2 && vec4_ok<T>(x, g) && reinterpret_cast<uintptr_t>(out) % (4 * sizeof(T)) == 0 && debug_path() != 1) {
    const long long blocks = ((long long)g.B * g.H * (g.W / 4) + 255) / 256;     // streaming composite, 4 px per thread
    composite_fwd_vec<T><<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, s>>>((const T*)x, (T*)out, g);
    MGR_CUDA(cudaGetLastError());
    count_launch();
    return MGR_OK;
  }
  dim3 grid((g.W + mgr::kTileW - 1) / mgr::kTileW, (g.H + mgr::kTileH - 1) / mgr::kTileH, g.B);
  const size_t smem = sizeof(mgr::TileAffine) * g.L;
  if (theta)
    mgr::render_fwd_direct<T, true><<<grid, mgr::kDirectThreads, smem, s>>>((const T*)x, theta, (T*)out, g);
  else
    mgr::render_fwd_direct<T, false><<<grid, mgr::kDirectThreads, 0, s>>>((const T*)x, nullptr, (T*)out, g);
  MGR_CUDA(cudaGetLastError());
  count_launch();
  return MGR_OK;
}

template <typename T, int LMAX, bool kWarp>
int launch_backward_l(const void* x, const float* theta, const void* out, const void* gout, float* gx32,
                      void* gx, float* gtheta, const mgr::Geometry& g, int flags, cudaStream_t s) {
  dim3 grid((g.W + mgr::kTileW - 1) / mgr::kTileW, (g.H + mgr::kTileH - 1) / mgr::kTileH, g.B);
  const size_t smem = kWarp ? (sizeof(mgr::TileAffine) + 6 * sizeof(float)) * g.L : 0;
  const bool nx = flags & MGR_NEED_GRAD_X, nt = kWarp && (flags & MGR_NEED_GRAD_THETA);
#define MGR_LAUNCH(NX, NT)                                                                              \
  mgr::render_bwd_direct<T, LMAX, kWarp, NX, NT><<<grid, mgr::kDirectThreads, smem, s>>>(               \
      (const T*)x, theta, (const T*)out, (const T*)gout, gx32, (T*)gx, gtheta, g)
  if (nx && nt) MGR_LAUNCH(true, true);
  else if (nx) MGR_LAUNCH(true, false);
  else if (nt) MGR_LAUNCH(false, true);
#undef MGR_LAUNCH
  MGR_CUDA(cudaGetLastError());
  count_launch();
  return MGR_OK;
}

template <typename T, bool kWarp>
int launch_backward(const void* x, const float* theta, const void* out, const void* gout, float* gx32, void* gx,
                    float* gtheta, const mgr::Geometry& g, int flags, cudaStream_t s) {
  if (g.L <= 8) return launch_backward_l<T, 8, kWarp>(x, theta, out, gout, gx32, gx, gtheta, g, flags, s);
  return launch_backward_l<T, 32, kWarp>(x, theta, out, gout, gx32, gx, gtheta, g, flags, s);
}

// two-pass tiled backward (+ the fused stencil backward for all-translation samples): records in the workspace, no
// atomics on grad_x; sources and gradients addressed through per-layer descriptors (canvas or ragged)
template <typename T, bool kRagged>
int backward_tiled(const void* x, const SrcLayers& src, const float* theta, const void* out, const void* gout, const void* sav,
                   void* gx, const DstLayers& dst, float* gtheta, void* ws, const mgr::Geometry& g, int flags, cudaStream_t s) {
  // two-pass tiled backward: records in the workspace, no atomics on grad_x
  using Vec = typename Texel<T>::Vec;
  using SA = typename SavedAlpha<T>::type;
  const bool nx = flags & MGR_NEED_GRAD_X, nt = flags & MGR_NEED_GRAD_THETA;
  float2* rec = reinterpret_cast<float2*>(ws);
  float4* gp = reinterpret_cast<float4*>(reinterpret_cast<char*>(ws) + sizeof(float2) * (size_t)g.B * g.L * g.H * g.W);
  InverseLayer* inv = reinterpret_cast<InverseLayer*>(reinterpret_cast<char*>(gp) + sizeof(float4) * (size_t)g.B * g.H * g.W);
  int* order = reinterpret_cast<int*>(inv + (size_t)g.B * g.L);          // [B*L] + 2 counters, then work [B*L] + 2 counters
  int* work = order + (size_t)g.B * g.L + 2;
  int* wcnt = work + (size_t)g.B * g.L;                                  // [0] layers in the work list
  int* sflag = wcnt + 2;                                                 // [B] sample is all translations
  if ((long long)g.B * g.L > 65535) return fail(MGR_ERR_UNSUPPORTED, "B*L=%lld exceeds 65535 per backward call; split the batch", (long long)g.B * g.L);
  // placements first: inverse plans + launch order for pass 2, and the per-sample "all translations" flags every
  // kernel below uses to claim or decline a sample with one load
  if (g.B * g.L <= kSmallPlacements) {
    placements_small_kernel<<<1, 256, 0, s>>>(theta, inv, g.B, g.L, g.H, g.W, order, work, wcnt, sflag, nt ? gtheta : nullptr,
                                              debug_path() != 2);
    MGR_CUDA(cudaGetLastError());
    count_launch();
  } else {
    if (nt) MGR_CUDA(cudaMemsetAsync(gtheta, 0, sizeof(float) * 6 * g.B * g.L, s));
    MGR_CUDA(cudaMemsetAsync(order + g.B * g.L, 0, 2 * sizeof(int), s));
    inverse_plans_kernel<<<(g.B * g.L + 127) / 128, 128, 0, s>>>(theta, inv, g.B * g.L, g.H, g.W, order, order + g.B * g.L);
    sample_flags_kernel<<<1, 256, 0, s>>>(inv, g.B, g.L, order, work, wcnt, sflag, debug_path() != 2);
    MGR_CUDA(cudaGetLastError());
    count_launch(2);
  }
  dim3 grid((g.W + kTW - 1) / kTW, (g.H + kTH - 1) / kTH, g.B);
  const int shift = debug_path() != 2;
  const size_t gp_bytes = sizeof(float4) * kPx * kTiledThreads;
  // the general passes (side stream) and the stencil backward (`s`) touch disjoint samples; see SideStream
  ForkGuard fg;
  if (shift && use_side_stream(g))
    if (int rc = fg.fork(s)) return rc;
  cudaStream_t sg = fg.side_or(s);
  // warp-specialised pass 1 for 16-bit tensors; fp32 keeps the two-barrier kernel (its 45 KB slots leave room for one
  // CTA per SM only next to the transmittance stash: measured 4-12 % slower)
  if (debug_path() != 3 && sizeof(T) == 2) {
    // warp-specialised pass 1 (render_ws.cuh); (G_P, G_A) copy in shared memory if two CTAs per SM still fit
    const bool gp_smem = ws_bwd_smem_bytes(g.L, sizeof(Vec), true) * MGR_WSB_BLOCKS + 2 * 1024 <= 227 * 1024;
    const size_t smem = ws_bwd_smem_bytes(g.L, sizeof(Vec), gp_smem);
    auto launch = [&](auto kern) -> int {
      if (int rc = ensure_dynamic_smem(kern, smem)) return rc;
      kern<<<grid, kWsThreads, smem, sg>>>((const T*)x, src, theta, (const T*)out, (const T*)gout, (const SA*)sav, rec, gp,
                                           nt ? gtheta : nullptr, with_vec8<T>(src, g), sflag, shift);
      return MGR_OK;
    };
    int rc;
    if (nt) rc = gp_smem ? launch(render_bwd_pass1_ws<T, true, true, kRagged>) : launch(render_bwd_pass1_ws<T, true, false, kRagged>);
    else rc = gp_smem ? launch(render_bwd_pass1_ws<T, false, true, kRagged>) : launch(render_bwd_pass1_ws<T, false, false, kRagged>);
    if (rc) return rc;
  } else {
    size_t smem = align16(tiled_smem_bytes(g.L, sizeof(Vec))) +
                  sizeof(float) * (size_t)g.L * kPx * kTiledThreads;                // + transmittance stash
    // (G_P, G_A) copy in shared memory if two CTAs per SM still fit (the kernels are compiled for two: at three they
    // spill, and the spills cost more than the third CTA hides)
    const bool gp_smem = (smem + gp_bytes) * 2 + 2 * 1024 <= 227 * 1024;
    if (gp_smem) smem += gp_bytes;
    auto launch = [&](auto kern) -> int {
      if (int rc = ensure_dynamic_smem(kern, smem)) return rc;
      kern<<<grid, kTiledThreads, smem, sg>>>((const T*)x, src, theta, (const T*)out, (const T*)gout, (const SA*)sav, rec, gp,
                                              nt ? gtheta : nullptr, g, sflag, shift);
      return MGR_OK;
    };
    int rc;
    if (nt) rc = gp_smem ? launch(render_bwd_pass1<T, true, true, kRagged>) : launch(render_bwd_pass1<T, true, false, kRagged>);
    else rc = gp_smem ? launch(render_bwd_pass1<T, false, true, kRagged>) : launch(render_bwd_pass1<T, false, false, kRagged>);
    if (rc) return rc;
  }
  MGR_CUDA(cudaGetLastError());
  count_launch();
  if (nx) {
    // square 32 x 32 texel blocks for big launches, 64 x 16 when a heavy layer's blocks would be the tail (see P2Shape)
    const long long blocks = (long long)((g.W + 31) / 32) * ((g.H + 31) / 32) * g.B * g.L;
    if (blocks >= 16384) {
      dim3 grid2((g.W + 31) / 32, (g.H + 31) / 32, g.B * g.L);
      render_bwd_pass2<T, kRagged, 16><<<grid2, 256, 0, sg>>>(inv, work, wcnt, rec, gp, (T*)gx, dst, g);
    } else {
      dim3 grid2((g.W + 63) / 64, (g.H + 15) / 16, g.B * g.L);
      render_bwd_pass2<T, kRagged, 32><<<grid2, 256, 0, sg>>>(inv, work, wcnt, rec, gp, (T*)gx, dst, g);
    }
    MGR_CUDA(cudaGetLastError());
    count_launch();
  }
  bool shift_tma = false;
  if constexpr (!kRagged) {
    // canvas layout, contiguous grad_x: the stencil backward on TMA box copies (render_shift_tma_bwd.cuh)
    CUtensorMap xmap, amap, gmap, omap;
    // transmittances in shared memory (16-bit tensors while two CTAs per SM fit: up to 19 layers; measured at L = 9, where
    // the tile needs 76 KB: 172 us against 215 us with the workspace variant at three CTAs per SM) or in the workspace
#ifdef MGR_STB_FORCE_GLOBAL_T
    const bool global_t = true;
#else
    const bool global_t = sizeof(T) == 4 || bwd_tma_layout<T>(g.L, false).total > MGR_STB_SMEM_CAP;
#endif
    const BwdSmem lay = bwd_tma_layout<T>(g.L, global_t);
    if (shift && debug_path() != 4 && (size_t)lay.total <= MGR_STB_SMEM_CAP && (!nx || reinterpret_cast<uintptr_t>(dst.s[0].ptr) % 16 == 0) &&
        bwd_tma_maps<T>(&xmap, &amap, &gmap, &omap, x, sav, gout, out, g, global_t)) {
      shift_tma = true;
      dim3 grid4((g.W + 1 + kBAncW - 1) / kBAncW, (g.H + 1 + kBAncH - 1) / kBAncH, g.B);
      const DstLayer& d0 = dst.s[0];
      const long long gx_sl = g.L > 1 ? (reinterpret_cast<T*>(dst.s[1].ptr) - reinterpret_cast<T*>(d0.ptr)) : 0;
      auto launch4 = [&](auto kern) -> int {
        if (int rc = ensure_dynamic_smem(kern, (size_t)lay.total)) return rc;
        kern<<<grid4, kBThreads, lay.total, s>>>(xmap, amap, gmap, omap, theta, (T*)d0.ptr, d0.sb, gx_sl, nt ? gtheta : nullptr, g, sflag,
                                                 (const SA*)sav, (const T*)gout, (const T*)out, reinterpret_cast<float*>(rec));
        return MGR_OK;
      };
      int rc;
      if constexpr (sizeof(T) == 4) {
        if (nx && nt) rc = launch4(render_bwd_shift_tma<T, true, true, true>);
        else if (nx) rc = launch4(render_bwd_shift_tma<T, true, false, true>);
        else rc = launch4(render_bwd_shift_tma<T, false, true, true>);
      } else if (global_t) {
        if (nx && nt) rc = launch4(render_bwd_shift_tma<T, true, true, true>);
        else if (nx) rc = launch4(render_bwd_shift_tma<T, true, false, true>);
        else rc = launch4(render_bwd_shift_tma<T, false, true, true>);
      } else {
        if (nx && nt) rc = launch4(render_bwd_shift_tma<T, true, true, false>);
        else if (nx) rc = launch4(render_bwd_shift_tma<T, true, false, false>);
        else rc = launch4(render_bwd_shift_tma<T, false, true, false>);
      }
      if (rc) return rc;
      MGR_CUDA(cudaGetLastError());
      count_launch();
    }
  }
  if (shift && !shift_tma) {
    size_t smem3 = shift_bwd_smem_bytes(g.L, sizeof(Vec));
    const bool gp_smem3 = (smem3 + gp_bytes) * 2 + 2 * 1024 <= 227 * 1024;
    if (gp_smem3) smem3 += gp_bytes;
    dim3 grid3((g.W + 1 + kAnchor - 1) / kAnchor, (g.H + 1 + kAnchor - 1) / kAnchor, g.B);
    auto launch3 = [&](auto kern) -> int {
      if (int rc = ensure_dynamic_smem(kern, smem3)) return rc;
      kern<<<grid3, kTiledThreads, smem3, s>>>(src, theta, (const T*)out, (const T*)gout, (const SA*)sav, dst,
                                               nt ? gtheta : nullptr, gp, g, sflag);
      return MGR_OK;
    };
    int rc;
    if (nx && nt) rc = gp_smem3 ? launch3(render_bwd_shift<T, true, true, true>) : launch3(render_bwd_shift<T, true, true, false>);
    else if (nx) rc = gp_smem3 ? launch3(render_bwd_shift<T, true, false, true>) : launch3(render_bwd_shift<T, true, false, false>);
    else rc = gp_smem3 ? launch3(render_bwd_shift<T, false, true, true>) : launch3(render_bwd_shift<T, false, true, false>);
    if (rc) return rc;
    MGR_CUDA(cudaGetLastError());
    count_launch();
  }
  return fg.join();
}

template <typename T>
int backward_typed(const void* x, const float* theta, const void* out, const void* gout, const void* sav, void* gx,
                   float* gtheta, void* ws, const mgr::Geometry& g, int flags, cudaStream_t s) {
  const long long n = (long long)g.B * g.L * 4 * g.H * g.W;
  if (theta && sav && g.L >= 2 && tiled_ok<T>(x, g))
    return backward_tiled<T, false>(x, canvas_src<T>(x, g), theta, out, gout, sav, gx, canvas_dst<T>(gx, g), gtheta, ws, g, flags, s);
  if (theta) {
    if (flags & MGR_NEED_GRAD_THETA) MGR_CUDA(cudaMemsetAsync(gtheta, 0, sizeof(float) * 6 * g.B * g.L, s));
    float* gx32 = nullptr;
    if (flags & MGR_NEED_GRAD_X) {
      gx32 = (sizeof(T) == 4) ? (float*)gx : (float*)ws;
      MGR_CUDA(cudaMemsetAsync(gx32, 0, sizeof(float) * n, s));
    }
    int rc = launch_backward<T, true>(x, theta, out, gout, gx32, gx, gtheta, g, flags, s);
    if (rc) return rc;
    if ((flags & MGR_NEED_GRAD_X) && sizeof(T) != 4) {
      const int threads = 256;
      const long long blocks = (n + threads - 1) / threads;
      mgr::cast_from_f32<T><<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), threads, 0, s>>>(gx32, (T*)gx, n);
      MGR_CUDA(cudaGetLastError());
      count_launch();
    }
    return MGR_OK;
  }
  if constexpr (sizeof(T) == 2) {               // 16-bit: two pixels per thread (composite_only.cuh)
    const bool ok2 = g.W % 2 == 0 && g.sh % 2 == 0 && g.sc % 2 == 0 && g.sl % 2 == 0 && g.sb % 2 == 0 &&
                     (reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gout) |
                      reinterpret_cast<uintptr_t>(gx)) % 4 == 0;
    if ((flags & MGR_NEED_GRAD_X) && g.L >= 2 && g.L <= 32 && ok2 && debug_path() != 1) {
      const long long blocks = ((long long)g.B * g.H * (g.W / 2) + 255) / 256;
      const unsigned grid = (unsigned)(blocks < 148 * 32 ? blocks : 148 * 32);
      if (g.L <= 8) composite_bwd_vec2<T, 8><<<grid, 256, 0, s>>>((const T*)x, (const T*)out, (const T*)gout, (T*)gx, g);
      else if (g.L <= 16) composite_bwd_vec2<T, 16><<<grid, 256, 0, s>>>((const T*)x, (const T*)out, (const T*)gout, (T*)gx, g);
      else composite_bwd_vec2<T, 32><<<grid, 256, 0, s>>>((const T*)x, (const T*)out, (const T*)gout, (T*)gx, g);
      MGR_CUDA(cudaGetLastError());
      count_launch();
      return MGR_OK;
    }
  }
  return launch_backward<T, false>(x, nullptr, out, gout, nullptr, gx, nullptr, g, flags & MGR_NEED_GRAD_X, s);
}

// ---- materialised warp and staging helpers (warp_ops.cuh) -----------------------------------------
template <typename T>
int launch_warp_forward(const void* x, const float* theta, void* out, const Geometry& g, cudaStream_t s) {
  if (tiled_ok<T>(x, g)) {                                     // staged footprints, one layer per CTA (warp_tiled.cuh)
    // layers that are pure translations: box copies + planar stencil (render_shift_tma.cuh); the two kernels partition the
    // layers.  One stream: forking was measured to cost a general batch more than it saves a translation batch here
    // (fp32 class-swap step at C2: 1422 vs 1348 us general, 908 vs 912 us translations -- tools/dropin_bench.py)
    CUtensorMap xmap;
    const bool tma = debug_path() != 4 && debug_path() != 2 && reinterpret_cast<uintptr_t>(out) % 16 == 0 &&
                     shift_tma_x_map<T>(&xmap, x, g, ShiftBox<T, kWTH>::W, ShiftBox<T, kWTH>::H);
    dim3 gridt((g.W + kTW - 1) / kTW, (g.H + kTH - 1) / kTH, g.B * g.L);
    warp_fwd_tiled<T><<<gridt, kTiledThreads, sizeof(typename Texel<T>::Vec) * kCapTexels, s>>>((const T*)x, theta, (T*)out, g, tma ? 1 : 0);
    MGR_CUDA(cudaGetLastError());
    count_launch();
    if (tma) {
      dim3 grids((g.W + kSW - 1) / kSW, (g.H + kWTH - 1) / kWTH, g.B * g.L);
      warp_fwd_shift_tma<T, false><<<grids, kWConsumers, ShiftBox<T, kWTH>::kStageBytes, s>>>(xmap, theta, (T*)out, g);
      MGR_CUDA(cudaGetLastError());
      count_launch();
    }
    return MGR_OK;
  }
  dim3 grid((g.W + kTileW - 1) / kTileW, (g.H + kTileH - 1) / kTileH, g.B * g.L);
  warp_fwd_kernel<T><<<grid, kDirectThreads, 0, s>>>((const T*)x, theta, (T*)out, g);
  MGR_CUDA(cudaGetLastError());
  count_launch();
  return MGR_OK;
}

template <typename T>
int launch_warp_backward(const void* x, const float* theta, const void* gout, void* gx, float* gtheta, void* ws,
                         const Geometry& g, int flags, cudaStream_t s) {
  const long long n = (long long)g.B * g.L * 4 * g.H * g.W;
  const bool nx = flags & MGR_NEED_GRAD_X, nt = flags & MGR_NEED_GRAD_THETA;
  if (tiled_ok<T>(x, g)) {
    // atomics-free: gather-form adjoint for grad_x, staged footprints + per-CTA reduction for grad_theta (warp_tiled.cuh)
    using Vec = typename Texel<T>::Vec;
    InverseLayer* inv = reinterpret_cast<InverseLayer*>(ws);
    int* order = reinterpret_cast<int*>(inv + (size_t)g.B * g.L);
    int* work = order + (size_t)g.B * g.L + 2;
    int* wcnt = work + (size_t)g.B * g.L;
    int* sflag = wcnt + 2;
    // grad_x of the layers that are pure translations: the forward's box-copy kernel run as its own adjoint on the upstream
    // gradient (contiguous [B,L,4,H,W]); the gather below leaves those layers alone.
    CUtensorMap gmap;
    Geometry gc = g;
    gc.sh = g.W; gc.sc = (long long)g.H * g.W; gc.sl = 4 * gc.sc; gc.sb = g.L * gc.sl;
    const bool tma = nx && debug_path() != 4 && debug_path() != 2 && reinterpret_cast<uintptr_t>(gx) % 16 == 0 &&
                     shift_tma_x_map<T>(&gmap, gout, gc, ShiftBox<T, kWTH>::W, ShiftBox<T, kWTH>::H);
    if (tma) {
      dim3 grids((g.W + kSW - 1) / kSW, (g.H + kWTH - 1) / kWTH, g.B * g.L);
      warp_fwd_shift_tma<T, true><<<grids, kWConsumers, ShiftBox<T, kWTH>::kStageBytes, s>>>(gmap, theta, (T*)gx, g);
      MGR_CUDA(cudaGetLastError());
      count_launch();
    }
    cudaStream_t sg = s;
    if (nx) {
      if (g.B * g.L <= kSmallPlacements) {                       // every layer stays in the work list (skip_shift = 0)
        placements_small_kernel<<<1, 256, 0, sg>>>(theta, inv, g.B, g.L, g.H, g.W, order, work, wcnt, sflag, nullptr, 0);
        MGR_CUDA(cudaGetLastError());
        count_launch();
      } else {
        MGR_CUDA(cudaMemsetAsync(order + g.B * g.L, 0, 2 * sizeof(int), sg));
        inverse_plans_kernel<<<(g.B * g.L + 127) / 128, 128, 0, sg>>>(theta, inv, g.B * g.L, g.H, g.W, order, order + g.B * g.L);
        sample_flags_kernel<<<1, 256, 0, sg>>>(inv, g.B, g.L, order, work, wcnt, sflag, 0);
        MGR_CUDA(cudaGetLastError());
        count_launch(2);
      }
      const long long blocks = (long long)((g.W + 31) / 32) * ((g.H + 31) / 32) * g.B * g.L;
      if (blocks >= 16384) {
        dim3 grid2((g.W + 31) / 32, (g.H + 31) / 32, g.B * g.L);
        warp_bwd_gather<T, 16><<<grid2, 256, 0, sg>>>(inv, work, wcnt, (const T*)gout, (T*)gx, g, tma ? 1 : 0);
      } else {
        dim3 grid2((g.W + 63) / 64, (g.H + 15) / 16, g.B * g.L);
        warp_bwd_gather<T, 32><<<grid2, 256, 0, sg>>>(inv, work, wcnt, (const T*)gout, (T*)gx, g, tma ? 1 : 0);
      }
      MGR_CUDA(cudaGetLastError());
      count_launch();
    }
    if (nt) {
      MGR_CUDA(cudaMemsetAsync(gtheta, 0, sizeof(float) * 6 * g.B * g.L, s));
      const size_t smem = sizeof(Vec) * kCapTexels + sizeof(float) * kPx * kTiledThreads;
      if (int rc = ensure_dynamic_smem(warp_bwd_theta_tiled<T>, smem)) return rc;
      dim3 gridt((g.W + kTW - 1) / kTW, (g.H + kTH - 1) / kTH, g.B * g.L);
      CUtensorMap xmap;
      const bool tma_t = debug_path() != 4 && debug_path() != 2 && reinterpret_cast<uintptr_t>(gout) % 16 == 0 &&
                         shift_tma_x_map<T>(&xmap, x, g, ShiftBox<T, kWTH>::W, ShiftBox<T, kWTH>::H);
      warp_bwd_theta_tiled<T><<<gridt, kTiledThreads, smem, s>>>((const T*)x, theta, (const T*)gout, gtheta, g, tma_t ? 1 : 0);
      MGR_CUDA(cudaGetLastError());
      count_launch();
      if (tma_t) {                                             // translation layers: box copies (render_shift_tma.cuh)
        dim3 grids((g.W + kSW - 1) / kSW, (g.H + kWTH - 1) / kWTH, g.B * g.L);
        warp_bwd_theta_shift_tma<T><<<grids, kWConsumers, ShiftBox<T, kWTH>::kStageBytes, s>>>(xmap, theta, (const T*)gout, gtheta, g);
        MGR_CUDA(cudaGetLastError());
        count_launch();
      }
    }
    return MGR_OK;
  }
  float* gx32 = nullptr;
  if (nx) {
    gx32 = (sizeof(T) == 4) ? (float*)gx : (float*)ws;
    MGR_CUDA(cudaMemsetAsync(gx32, 0, sizeof(float) * n, s));
  }
  if (nt) MGR_CUDA(cudaMemsetAsync(gtheta, 0, sizeof(float) * 6 * g.B * g.L, s));
  dim3 grid((g.W + kTileW - 1) / kTileW, (g.H + kTileH - 1) / kTileH, g.B * g.L);
  if (nx && nt) warp_bwd_kernel<T, true, true><<<grid, kDirectThreads, 0, s>>>((const T*)x, theta, (const T*)gout, gx32, gtheta, g);
  else if (nx) warp_bwd_kernel<T, true, false><<<grid, kDirectThreads, 0, s>>>((const T*)x, theta, (const T*)gout, gx32, gtheta, g);
  else if (nt) warp_bwd_kernel<T, false, true><<<grid, kDirectThreads, 0, s>>>((const T*)x, theta, (const T*)gout, gx32, gtheta, g);
  MGR_CUDA(cudaGetLastError());
  count_launch();
  if (nx && sizeof(T) != 4) {
    const long long blocks = (n + 255) / 256;
    cast_from_f32<T><<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, s>>>(gx32, (T*)gx, n);
    MGR_CUDA(cudaGetLastError());
    count_launch();
  }
  return MGR_OK;
}

template <typename T>
int launch_pad_stack(const void* src, const long long* ss, void* dst, int B, int L, int l, int h, int w, int H, int W,
                     float pad, cudaStream_t s) {
  if ((long long)B * 4 > 65535 || (H + 3) / 4 > 65535) return fail(MGR_ERR_UNSUPPORTED, "pad_stack: B=%d or H=%d too large for one call", B, H);
  dim3 grid((W + 255) / 256, (H + 3) / 4, B * 4);
  const bool vec = W % 4 == 0 && reinterpret_cast<uintptr_t>(dst) % (4 * sizeof(T)) == 0;
  if (vec) pad_stack_kernel<T, true><<<grid, 256, 0, s>>>((const T*)src, ss[0], ss[1], ss[2], ss[3], (T*)dst, L, l, h, w, H, W, pad);
  else pad_stack_kernel<T, false><<<grid, 256, 0, s>>>((const T*)src, ss[0], ss[1], ss[2], ss[3], (T*)dst, L, l, h, w, H, W, pad);
  MGR_CUDA(cudaGetLastError());
  count_launch();
  return MGR_OK;
}

template <typename T>
int launch_composite_jvp(const void* x, const void* tx, void* tout, const Geometry& g, cudaStream_t s) {
  const long long total = (long long)g.B * g.H * g.W;
  const long long blocks = (total + 255) / 256;
  composite_jvp_kernel<T><<<(unsigned)(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, s>>>((const T*)x, (const T*)tx, (T*)tout, g);
  MGR_CUDA(cudaGetLastError());
  count_launch();
  return MGR_OK;
}

template <typename T>
int launch_pil_composite(const void* x, float* out_f32, uint8_t* out_u8, const Geometry& g, cudaStream_t s) {
  // four pixels per thread when rows, planes and bases are 4-element aligned (the reference's contiguous layout is)
  const bool vec = g.W % 4 == 0 && g.sh % 4 == 0 && g.sc % 4 == 0 && g.sl % 4 == 0 && g.sb % 4 == 0 &&
                   reinterpret_cast<uintptr_t>(x) % (4 * sizeof(T)) == 0 &&
                   (!out_f32 || reinterpret_cast<uintptr_t>(out_f32) % 16 == 0) &&
                   (!out_u8 || reinterpret_cast<uintptr_t>(out_u8) % 4 == 0);
  const long long total = (long long)g.B * g.H * (vec ? g.W / 4 : g.W);
  const long long blocks = (total + 255) / 256;
  const unsigned grid = (unsigned)(blocks < 148 * 16 ? blocks : 148 * 16);     // grid-stride, 16 CTAs per SM at most
  if (vec) pil_composite_kernel<T, 4><<<grid, 256, 0, s>>>((const T*)x, out_f32, out_u8, g);
  else pil_composite_kernel<T, 1><<<grid, 256, 0, s>>>((const T*)x, out_f32, out_u8, g);
  MGR_CUDA(cudaGetLastError());
  count_launch();
  return MGR_OK;
}

// ---- ragged stacks: one [B,4,h,w] tensor per layer (tiled kernels only) -----------------------------------------
template <typename T>
int launch_forward_ragged(const SrcLayers& src, const float* theta, void* out, void* sav, const Geometry& g, cudaStream_t s) {
  if (const char* why = ragged_problem<T>(src, nullptr, g)) return fail(MGR_ERR_UNSUPPORTED, "ragged stack: %s", why);
  return launch_forward_tiled<T, true>(nullptr, src, theta, out, sav, g, s);
}
template <typename T>
int launch_backward_ragged(const SrcLayers& src, const float* theta, const void* out, const void* gout, const void* sav,
                           const DstLayers& dst, float* gtheta, void* ws, const Geometry& g, int flags, cudaStream_t s) {
  if (const char* why = ragged_problem<T>(src, (flags & MGR_NEED_GRAD_X) ? &dst : nullptr, g))
    return fail(MGR_ERR_UNSUPPORTED, "ragged stack: %s", why);
  return backward_tiled<T, true>(nullptr, src, theta, out, gout, sav, nullptr, dst, gtheta, ws, g, flags, s);
}

}  // namespace mgr

#include "launchers_decl.h"
#define MGR_INSTANTIATE(SUFFIX, T)                                                                           \
  bool mgr_tiled_ok_##SUFFIX(const void* x, const mgr::Geometry& g) { return g.L >= 2 && mgr::tiled_ok<T>(x, g); } \
  int mgr_fwd_##SUFFIX(const void* x, const float* theta, void* out, void* sav, const mgr::Geometry& g,     \
                       cudaStream_t s) {                                                                     \
    return mgr::launch_forward<T>(x, theta, out, sav, g, s);                                                 \
  }                                                                                                          \
  int mgr_bwd_##SUFFIX(const void* x, const float* theta, const void* out, const void* gout, const void* sav, \
                       void* gx, float* gtheta, void* ws, const mgr::Geometry& g, int flags, cudaStream_t s) { \
    return mgr::backward_typed<T>(x, theta, out, gout, sav, gx, gtheta, ws, g, flags, s);                    \
  }                                                                                                          \
  int mgr_warp_fwd_##SUFFIX(const void* x, const float* theta, void* out, const mgr::Geometry& g, cudaStream_t s) { \
    return mgr::launch_warp_forward<T>(x, theta, out, g, s);                                                 \
  }                                                                                                          \
  int mgr_warp_bwd_##SUFFIX(const void* x, const float* theta, const void* gout, void* gx, float* gtheta,    \
                            void* ws, const mgr::Geometry& g, int flags, cudaStream_t s) {                   \
    return mgr::launch_warp_backward<T>(x, theta, gout, gx, gtheta, ws, g, flags, s);                        \
  }                                                                                                          \
  int mgr_pad_stack_##SUFFIX(const void* src, const long long* ss, void* dst, int B, int L, int l, int h,    \
                             int w, int H, int W, float pad, cudaStream_t s) {                               \
    return mgr::launch_pad_stack<T>(src, ss, dst, B, L, l, h, w, H, W, pad, s);                              \
  }                                                                                                          \
  int mgr_jvp_##SUFFIX(const void* x, const void* tx, void* tout, const mgr::Geometry& g, cudaStream_t s) {  \
    return mgr::launch_composite_jvp<T>(x, tx, tout, g, s);                                                  \
  }                                                                                                          \
  int mgr_pil_##SUFFIX(const void* x, float* of, unsigned char* ou, const mgr::Geometry& g, cudaStream_t s) { \
    return mgr::launch_pil_composite<T>(x, of, ou, g, s);                                                    \
  }                                                                                                          \
  int mgr_fwd_ragged_##SUFFIX(const mgr::SrcLayers& src, const float* theta, void* out, void* sav,           \
                              const mgr::Geometry& g, cudaStream_t s) {                                      \
    return mgr::launch_forward_ragged<T>(src, theta, out, sav, g, s);                                        \
  }                                                                                                          \
  int mgr_bwd_ragged_##SUFFIX(const mgr::SrcLayers& src, const float* theta, const void* out, const void* gout, \
                              const void* sav, const mgr::DstLayers& dst, float* gtheta, void* ws,           \
                              const mgr::Geometry& g, int flags, cudaStream_t s) {                           \
    return mgr::launch_backward_ragged<T>(src, theta, out, gout, sav, dst, gtheta, ws, g, flags, s);         \
  }
