// Non-differentiable 8-bit "over" composite, bit-exact with the reference's Pillow path:
//   custom_utils/image_utils.py:74-96  alpha_composite(blchw_lchw):
//     ToPILImage()(chw)            float [0,1] -> byte = trunc(v * 255)   (fp32 product, truncation)
//     canvas.alpha_composite(img)  Pillow's integer "over" (libImaging/AlphaComposite.c, third-party, Pillow 12.2.0
//                                  in this image; the algorithm is restated in oracle/restatement.py and pinned
//                                  against Pillow itself by tests/golden/pil_composite_golden.npz)
//     ToTensor()(canvas)           byte -> fp32 byte / 255
// Consumers: snapshots / metrics / renderer-training targets (custom/loss_aio.py:351,362,
// custom/training_loop_aio.py:531,765,775, metrics/metric_utils.py:233,304) -- there a per-sample, per-layer
// CPU loop with a D2H and an H2D copy around it; here one streaming kernel, layers read once, byte arithmetic in
// registers.  HBM-bound: B*L*4*H*W*s_x bytes in, B*4*H*W*(4 [+1]) bytes out.
#pragma once
#include "mgr_common.cuh"

namespace mgr {

// Pillow: #define SHIFTFORDIV255(a) ((((a) >> 8) + a) >> 8), PRECISION_BITS 7
__device__ __forceinline__ uint32_t div255(uint32_t a) { return ((a >> 8) + a) >> 8; }

// float in the caller's range -> Pillow byte.  m11: normalize_zero1 first, (t + 1) / 2 in fp32 (image_utils.py:184-187).
// Values outside [0,1] saturate (the reference's uint8 cast wraps there; it never feeds such values).
__device__ __forceinline__ uint32_t to_byte(float v, bool m11) {
  if (m11) v = __fmul_rn(__fadd_rn(v, 1.f), 0.5f);
  const float s = __fmul_rn(v, 255.f);                        // no contraction: the product is rounded to fp32 first
  return min(__float2uint_rz(s), 255u);                       // truncation; the conversion saturates, negatives and NaN -> 0
}

// floor(n / d) for n < 2^31, 0 < d <= 65025 with a quotient below 2^16: fp32 estimate (relative error ~2e-7, so at
// most one off) corrected with one integer multiply -- a third of the instructions of the generic 32-bit division.
__device__ __forceinline__ uint32_t div_small_quotient(uint32_t n, uint32_t d) {
  uint32_t q = __float2uint_rz(__fmul_rn((float)n, __frcp_rn((float)d)));
  const int r = (int)(n - q * d);
  if (r < 0) --q;
  else if (r >= (int)d) ++q;
  return q;
}

// dst <- src over dst, 8-bit straight alpha, per AlphaComposite.c
__device__ __forceinline__ void pil_over(uint32_t (&d)[4], const uint32_t (&s)[4]) {
  if (s[3] == 0) return;                                      // transparent source: destination copied through
  const uint32_t blend = d[3] * (255u - s[3]);
  const uint32_t outa255 = s[3] * 255u + blend;               // > 0 here
  const uint32_t coef1 = div_small_quotient(s[3] * (255u * 255u * 128u), outa255);   // numerator <= 255^3 * 128 < 2^31
  const uint32_t coef2 = 255u * 128u - coef1;
#pragma unroll
  for (int c = 0; c < 3; ++c) d[c] = div255(s[c] * coef1 + d[c] * coef2 + (0x80u << 7)) >> 7;
  d[3] = div255(outa255 + 0x80u);
}

// kVec adjacent pixels per thread (kVec = 4 needs 4-element-aligned rows and strides; kVec = 1 always works)
template <typename T, int kVec>
__global__ void __launch_bounds__(256)
pil_composite_kernel(const T* __restrict__ x, float* __restrict__ out_f32, uint8_t* __restrict__ out_u8, Geometry g) {
  const int wv = g.W / kVec;
  const long long hw = (long long)g.H * g.W;
  const long long total = (long long)g.B * g.H * wv;
  const bool m11 = g.m11 != 0;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int jv = (int)(k % wv);
    const int i = (int)((k / wv) % g.H);
    const long long b = k / ((long long)wv * g.H);
    const T* xb = x + b * g.sb + (long long)i * g.sh + (long long)jv * kVec;
    uint32_t d[kVec][4];
    for (int l = 0; l < g.L; ++l) {
      const T* xl = xb + (long long)l * g.sl;
      uint32_t s[kVec][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[kVec];
        ld_vec<T, kVec>(xl + c * g.sc, v);
#pragma unroll
        for (int q = 0; q < kVec; ++q) s[q][c] = to_byte(v[q], m11);
      }
#pragma unroll
      for (int q = 0; q < kVec; ++q) {
        if (l == 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) d[q][c] = s[q][c];      // canvas = layer 0 (image_utils.py:85)
        } else {
          pil_over(d[q], s[q]);
        }
      }
    }
    const long long o = b * 4 * hw + (long long)i * g.W + (long long)jv * kVec;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (out_f32) {
        float r[kVec];
#pragma unroll
        for (int q = 0; q < kVec; ++q) r[q] = __fdiv_rn((float)d[q][c], 255.f);     // ToTensor: byte / 255 in fp32
        st_vec_f32<kVec>(out_f32 + o + c * hw, r);
      }
      if (out_u8) {
        if (kVec == 4) {
          *reinterpret_cast<uint32_t*>(out_u8 + o + c * hw) = d[0][c] | (d[1][c] << 8) | (d[2][c] << 16) | (d[3][c] << 24);
        } else {
#pragma unroll
          for (int q = 0; q < kVec; ++q) out_u8[o + c * hw + q] = (uint8_t)d[q][c];
        }
      }
    }
  }
}

}  // namespace mgr
