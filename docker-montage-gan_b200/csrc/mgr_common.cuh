// Shared device helpers for the montage renderer kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mgr {

// ---- element access: storage dtype <-> fp32 registers -------------------------------------
template <typename T> __device__ __forceinline__ float ld(const T* p);
template <> __device__ __forceinline__ float ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __uint_as_float(static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
}
template <> __device__ __forceinline__ float ld<__half>(const __half* p) {
  return __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short*>(p))));
}

template <typename T> __device__ __forceinline__ void st(T* p, float v);
template <> __device__ __forceinline__ void st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void st<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// ---- problem description handed to every kernel -------------------------------------------
struct Geometry {
  int B, L, H, W;
  long long sb, sl, sc, sh;  // element strides of x for [B,L,4,H,W]; W stride is 1
  int m11;                   // 1: [-1,1] range mode, 0: [0,1]
  int vec8;                  // warp-specialised kernels: 16-bit footprints staged in 8-texel items (set by the launchers)
};

// layout of the forward's saved-alpha buffer: [B,L,H,W] alpha samples, padded to 256 bytes, then int flags[B]
__host__ __device__ inline size_t saved_alpha_flags_offset(int B, int L, int H, int W, size_t elem_bytes) {
  return (((size_t)B * L * H * W * elem_bytes) + 255) & ~(size_t)255;
}

// kVec adjacent elements -> fp32 (kVec = 4: one 16- or 8-byte load; the caller guarantees the alignment)
template <typename T, int kVec> struct VecIO;
template <typename T> struct VecIO<T, 1> {
  static __device__ __forceinline__ void ld(const T* p, float (&v)[1]) { v[0] = mgr::ld(p); }
};
template <> struct VecIO<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <> struct VecIO<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  }
};
template <> struct VecIO<__half, 4> {
  static __device__ __forceinline__ void ld(const __half* p, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
};
template <typename T, int kVec> __device__ __forceinline__ void ld_vec(const T* p, float (&v)[kVec]) { VecIO<T, kVec>::ld(p, v); }
template <int kVec> __device__ __forceinline__ void st_vec_f32(float* p, const float (&v)[kVec]) {
  if constexpr (kVec == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else p[0] = v[0];
}

// two adjacent 16-bit elements: one 32-bit access
template <> struct VecIO<__nv_bfloat16, 2> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[2]) {
    const uint32_t t = __ldg(reinterpret_cast<const uint32_t*>(p));
    v[0] = __uint_as_float(t << 16); v[1] = __uint_as_float(t & 0xffff0000u);
  }
};
template <> struct VecIO<__half, 2> {
  static __device__ __forceinline__ void ld(const __half* p, float (&v)[2]) {
    const uint32_t t = __ldg(reinterpret_cast<const uint32_t*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t));
    v[0] = a.x; v[1] = a.y;
  }
};
__device__ __forceinline__ void st_vec2(__nv_bfloat16* p, const float (&v)[2]) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v[0], v[1]);
}
__device__ __forceinline__ void st_vec2(__half* p, const float (&v)[2]) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(v[0], v[1]); }

template <typename T> __device__ __forceinline__ void st_vec4(T* p, const float (&v)[4]);      // four adjacent elements, aligned
template <> __device__ __forceinline__ void st_vec4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void st_vec4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<const uint32_t*>(&a); t.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}
template <> __device__ __forceinline__ void st_vec4<__half>(__half* p, const float (&v)[4]) {
  const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<const uint32_t*>(&a); t.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

// Where one layer's pixels live.  The canvas layout x[B,L,4,H,W] is the special case {x + l*sl, sb, sc, sh, H, W, 0, 0};
// the ragged layout (the local generators' outputs at native size, custom_utils/image_utils.py:216-243
// without the padded copy) gives every layer its own [B,4,h,w] tensor centred at (left, top) of the H x W canvas.
// Texels outside the rectangle are the padding value (-1 in m11 mode, 0 in 01 mode: transparent black), which is also
// what padding_mode='zeros' yields outside the canvas -- so "ragged" only means "per-layer bounds and base pointer".
constexpr int kMaxTiledLayers = 32;
struct SrcLayer {
  const void* ptr;           // element (b = 0, c = 0, y = 0, x = 0) of this layer
  long long sb, sc, sh;      // element strides: batch, channel, row (column stride 1)
  int h, w, top, left;
};
struct SrcLayers { SrcLayer s[kMaxTiledLayers]; };
struct DstLayer {            // grad_x of one layer: [B,4,h,w], rows contiguous
  void* ptr;
  long long sb, sc, sh;
  int h, w, top, left;
};
struct DstLayers { DstLayer s[kMaxTiledLayers]; };

// Pixel-space placement of one layer relative to a tile origin (j0,i0), SURVEY.md A.1:
//   ix(j,i) = a00*(j-j0) + a01*(i-i0) + (X0 + rx),  iy likewise.
// The integer part X0/Y0 is split off in double precision so the fp32 per-pixel arithmetic only
// carries tile-local magnitudes (fraction accurate to ~1e-6 px instead of ulp(256) = 3e-5 px).
struct TileAffine {
  float a00, a01, rx;
  float a10, a11, ry;
  int X0, Y0;
};

__device__ __forceinline__ TileAffine make_tile_affine(const float* __restrict__ th, int H, int W,
                                                       int j0, int i0) {
  // theta maps normalised output coords to normalised input coords (align_corners=False):
  //   gx = t00*x_j + t01*y_i + t02,  x_j = (2j+1)/W - 1,  ix = ((gx+1)*W - 1)/2
  // => ix = t00*(j + .5 - W/2) + t01*(W/H)*(i + .5 - H/2) + t02*W/2 + (W-1)/2
  const double t00 = th[0], t01 = th[1], t02 = th[2], t10 = th[3], t11 = th[4], t12 = th[5];
  const double w = W, h = H;
  const double a00 = t00, a01 = t01 * (w / h);
  const double a10 = t10 * (h / w), a11 = t11;
  double cx = a00 * (j0 + 0.5 - 0.5 * w) + a01 * (i0 + 0.5 - 0.5 * h) + t02 * 0.5 * w + 0.5 * (w - 1.0);
  double cy = a10 * (j0 + 0.5 - 0.5 * w) + a11 * (i0 + 0.5 - 0.5 * h) + t12 * 0.5 * h + 0.5 * (h - 1.0);
  // keep the integer split representable for absurd thetas (NaN falls through the min/max as-is)
  cx = fmin(fmax(cx, -1.0e9), 1.0e9);
  cy = fmin(fmax(cy, -1.0e9), 1.0e9);
  const double fx = floor(cx), fy = floor(cy);
  TileAffine t;
  t.a00 = (float)a00; t.a01 = (float)a01; t.a10 = (float)a10; t.a11 = (float)a11;
  t.X0 = (int)fx; t.Y0 = (int)fy;
  t.rx = (float)(cx - fx); t.ry = (float)(cy - fy);
  return t;
}

// Four in-bounds-masked bilinear taps of one channel plane.
struct Taps {
  int o00, o01, o10, o11;        // element offsets inside an x plane (valid only if the mask bit is set)
  int x0, y0;                    // integer tap origin (top-left), may be outside the image
  float w00, w01, w10, w11;      // bilinear weights (not masked)
  float fx, fy;
  unsigned mask;                 // bit0..3: tap 00,01,10,11 inside the image
};

__device__ __forceinline__ Taps make_taps(const TileAffine& t, int dj, int di, int H, int W, long long sh) {
  const float ix = fmaf(t.a00, (float)dj, fmaf(t.a01, (float)di, t.rx));
  const float iy = fmaf(t.a10, (float)dj, fmaf(t.a11, (float)di, t.ry));
  const float fxf = floorf(ix), fyf = floorf(iy);
  Taps p;
  p.fx = ix - fxf; p.fy = iy - fyf;
  // saturating float->int conversion keeps huge coordinates out of range instead of wrapping
  const int x0 = t.X0 + __float2int_rd(fminf(fmaxf(fxf, -1.0e9f), 1.0e9f));
  const int y0 = t.Y0 + __float2int_rd(fminf(fmaxf(fyf, -1.0e9f), 1.0e9f));
  const bool xin0 = (unsigned)x0 < (unsigned)W, xin1 = (unsigned)(x0 + 1) < (unsigned)W;
  const bool yin0 = (unsigned)y0 < (unsigned)H, yin1 = (unsigned)(y0 + 1) < (unsigned)H;
  p.mask = (xin0 && yin0 ? 1u : 0u) | (xin1 && yin0 ? 2u : 0u) | (xin0 && yin1 ? 4u : 0u) | (xin1 && yin1 ? 8u : 0u);
  const int base = y0 * (int)sh + x0;   // plane offsets fit in int (H*W <= 2^31 enforced by the API)
  p.x0 = x0; p.y0 = y0;
  p.o00 = base; p.o01 = base + 1; p.o10 = base + (int)sh; p.o11 = base + (int)sh + 1;
  const float ex = 1.f - p.fx, ey = 1.f - p.fy;
  p.w00 = ex * ey; p.w01 = p.fx * ey; p.w10 = ex * p.fy; p.w11 = p.fx * p.fy;
  return p;
}

// A placement that is exactly [[1,0,dx],[0,1,dy]] with a sane shift: what convert_translate_to_2x3 builds
// (custom_utils/image_utils.py:316-335).  The shift kernels and the general kernels partition the batch on it.
__device__ __forceinline__ bool is_pure_shift(const float* __restrict__ th) {
  return th[0] == 1.f && th[1] == 0.f && th[3] == 0.f && th[4] == 1.f && fabsf(th[2]) < 1.0e6f && fabsf(th[5]) < 1.0e6f;
}

// true iff every layer of a sample is a pure translation (uniform over the CTA; contains a barrier)
__device__ __forceinline__ bool cta_all_shift(const float* __restrict__ theta_b, int L, int tid, int nthreads) {
  bool ok = true;
  for (int l = tid; l < L; l += nthreads) ok = ok && is_pure_shift(theta_b + 6 * l);
  return __syncthreads_and(ok);
}

// normalised output coordinate of pixel index k along an axis of size n: (2k+1)/n - 1
__device__ __forceinline__ float norm_coord(int k, int n) { return (float)(2 * k + 1) / (float)n - 1.f; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum six per-lane values over the warp with a transposing butterfly: after the first three exchange steps each
// lane carries one of eight partial sums (two are padding), so the whole reduction takes 9 shuffles instead of
// 30.  On return, lane q * 4 (q = 0..5) holds the warp sum of v[q] in `out`; other lanes hold garbage.
__device__ __forceinline__ float warp_sum6(const float (&v)[6], int lane) {
  // step 1 (xor 16): lanes < 16 keep v0..v3 (index 0..3), lanes >= 16 keep v4, v5, 0, 0
  const bool hi16 = lane & 16;
  float a0 = hi16 ? v[4] : v[0], a1 = hi16 ? v[5] : v[1], a2 = hi16 ? 0.f : v[2], a3 = hi16 ? 0.f : v[3];
  const float s0 = hi16 ? v[0] : v[4], s1 = hi16 ? v[1] : v[5], s2 = hi16 ? v[2] : 0.f, s3 = hi16 ? v[3] : 0.f;
  a0 += __shfl_xor_sync(0xffffffffu, s0, 16);
  a1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  a2 += __shfl_xor_sync(0xffffffffu, s2, 16);
  a3 += __shfl_xor_sync(0xffffffffu, s3, 16);
  // step 2 (xor 8): lanes with bit 3 clear keep a0, a1; set keep a2, a3
  const bool hi8 = lane & 8;
  float b0 = hi8 ? a2 : a0, b1 = hi8 ? a3 : a1;
  const float t0 = hi8 ? a0 : a2, t1 = hi8 ? a1 : a3;
  b0 += __shfl_xor_sync(0xffffffffu, t0, 8);
  b1 += __shfl_xor_sync(0xffffffffu, t1, 8);
  // step 3 (xor 4): bit 2 clear keeps b0, set keeps b1
  const bool hi4 = lane & 4;
  float c = hi4 ? b1 : b0;
  const float w = hi4 ? b0 : b1;
  c += __shfl_xor_sync(0xffffffffu, w, 4);
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  return c;     // lane bits (16, 8, 4) = (q >= 4, (q & 2) != 0, q & 1): value index q = 4*bit16 + 2*bit3 + bit2
}
// the value index held by a lane after warp_sum6 (0..7; 6 and 7 are padding)
__device__ __forceinline__ int warp_sum6_index(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }

}  // namespace mgr
