// Microbenchmarks behind the round-2 backward redesign (developer tool, not product code).
//   1. shared-memory scatter of the bilinear adjoint: int32 fixed-point ATOMS.ADD (the only native shared atomic on
//      sm_100a; fp32 is a CAS loop) vs a non-atomic read-modify-write (wrong, rate reference) vs the arithmetic alone;
//      planar [channel][texel] vs interleaved [texel][channel] accumulators.
//   2. flushing per-(tile, layer) footprint sums into grad_x with vector REDs: REDG.BF16x4 (8 B) and REDG.F32x4 (16 B)
//      vs plain stores of the same addresses, on the C2 gradient (448 planes-quads of 256 x 256), whole tensor and in
//      L2-sized chunks of 16 samples.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scatter_bench scatter_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int kCap = 2816;
constexpr int kPitch = 52;

// MODE 0: ATOMS planar, 1: ATOMS interleaved, 2: non-atomic RMW planar, 3: arithmetic only
template <int MODE>
__global__ void __launch_bounds__(256, 2)
scatter_k(float* __restrict__ sink, int layers, float a00, float a01, float a10, float a11, long long* cyc) {
  __shared__ int acc[4 * kCap];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  for (int i = tid; i < 4 * kCap; i += 256) acc[i] = 0;
  __syncthreads();
  const long long t0 = clock64();
  float keep = 0.f;
  const float djf = (float)(tx - 16);
  for (int l = 0; l < layers; ++l) {
    const float rx = 24.3f + 0.37f * l + 0.01f * blockIdx.x, ry = 20.7f + 0.21f * l;
    const float scale = 1048576.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float dif = (float)(ty + 8 * k - 16);
      const float ix = fmaf(a00, djf, fmaf(a01, dif, rx)), iy = fmaf(a10, djf, fmaf(a11, dif, ry));
      const float fxf = floorf(ix), fyf = floorf(iy);
      const float fx = ix - fxf, fy = iy - fyf;
      int cell = (int)fyf * kPitch + (int)fxf;
      cell = min(max(cell, 0), kCap - kPitch - 2);
      const float ex = 1.f - fx, ey = 1.f - fy;
      const float w[4] = {ex * ey * scale, fx * ey * scale, ex * fy * scale, fx * fy * scale};
      const float g[4] = {0.1f + 0.001f * l + 0.0001f * tid, -0.2f + 0.001f * k, 0.3f * fx, 0.05f + fy};
      const int off[4] = {0, 1, kPitch, kPitch + 1};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float m = fmaf(w[t], g[c], 12582912.f);            // 1.5 * 2^23: integer lands in the low mantissa bits
          const int v = __float_as_int(m) - 0x4B400000;
          if (MODE == 0) atomicAdd(&acc[c * kCap + cell + off[t]], v);
          else if (MODE == 1) atomicAdd(&acc[(cell + off[t]) * 4 + c], v);
          else if (MODE == 2) { int* p = &acc[c * kCap + cell + off[t]]; *(volatile int*)p = *(volatile int*)p + v; }
          else keep += (float)v;
        }
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  float s = keep;
  for (int i = tid; i < 4 * kCap; i += 256) s += (float)acc[i];
  if (s == 123.456f) sink[0] = s;
  if (tid == 0) atomicAdd((unsigned long long*)cyc, (unsigned long long)(t1 - t0));
}

// ---- flush: one CTA per (plane-quad n, 32 x 32 tile); footprint = rows x vecs vectors of 4 texels per channel ----
__device__ __forceinline__ void red_bf16x4(__nv_bfloat16* p, uint32_t a, uint32_t b) {
  asm volatile("red.global.add.noftz.v2.bf16x2 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void red_f32x4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// MODE 0: RED, 1: plain store.   T = __nv_bfloat16 or float
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
flush_k(T* __restrict__ g, int H, int W, int rows, int vecs, int n0) {
  const int n = n0 + blockIdx.z;
  const int x_lo = (int)blockIdx.x * 32 - 4, y_lo = (int)blockIdx.y * 32 - 4;
  T* base = g + (size_t)n * 4 * H * W;
  const int per_ch = rows * vecs;
  for (int i = threadIdx.x; i < 4 * per_ch; i += 256) {
    const int c = i / per_ch, r = (i - c * per_ch) / vecs, v = i - c * per_ch - r * vecs;
    const int x = x_lo + 4 * v, y = y_lo + r;
    if (x < 0 || x + 4 > W || y < 0 || y >= H) continue;
    T* p = base + ((size_t)c * H + y) * W + x;
    if constexpr (sizeof(T) == 2) {
      if (MODE == 0) red_bf16x4(p, 0x3c003c00u + i, 0x3c003c00u);
      else *reinterpret_cast<uint2*>(p) = make_uint2(0x3c003c00u + i, 0x3c003c00u);
    } else {
      if (MODE == 0) red_f32x4(p, 1.f, 2.f, 3.f, (float)i);
      else *reinterpret_cast<float4*>(p) = make_float4(1.f, 2.f, 3.f, (float)i);
    }
  }
}

template <typename F>
float time_us(F&& f, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / reps;
}

int main() {
  float* sink; long long* cyc;
  CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&cyc, 8));
  const float cfg[4][4] = {{1.f, 0.02f, -0.02f, 1.f}, {0.9f, 0.25f, -0.2f, 0.85f}, {1.2f, 0.3f, -0.3f, 1.1f}, {0.5f, 0.1f, -0.1f, 0.5f}};
  const int ctas = 4096, layers = 7;             // == C2: 64 samples x 64 tiles, 7 layers
  const double lpx = (double)ctas * layers * 1024;
  const char* names[4] = {"ATOMS planar", "ATOMS interleaved", "non-atomic RMW planar", "arithmetic only"};
  for (int c = 0; c < 4; ++c) {
    for (int mode = 0; mode < 4; ++mode) {
      CK(cudaMemset(cyc, 0, 8));
      auto run = [&]() {
        if (mode == 0) scatter_k<0><<<ctas, 256>>>(sink, layers, cfg[c][0], cfg[c][1], cfg[c][2], cfg[c][3], cyc);
        else if (mode == 1) scatter_k<1><<<ctas, 256>>>(sink, layers, cfg[c][0], cfg[c][1], cfg[c][2], cfg[c][3], cyc);
        else if (mode == 2) scatter_k<2><<<ctas, 256>>>(sink, layers, cfg[c][0], cfg[c][1], cfg[c][2], cfg[c][3], cyc);
        else scatter_k<3><<<ctas, 256>>>(sink, layers, cfg[c][0], cfg[c][1], cfg[c][2], cfg[c][3], cyc);
      };
      const float us = time_us(run, 5);
      long long h = 0;
      CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
      printf("scatter cfg %d %-22s: %8.1f us for the C2 layer-pixel count  (%.0f layer-Mpix/s; %.0f cycles per CTA)\n", c, names[mode], us,
             lpx / us, (double)h / (6.0 * ctas));
    }
  }
  CK(cudaGetLastError());

  // ---- flush ----
  const int N = 448, H = 256, W = 256;
  void* g;
  CK(cudaMalloc(&g, (size_t)N * 4 * H * W * 4));
  const int rows = 42, vecs = 11;                // a 44 x 42 footprint per 32 x 32 tile: 1.8 texels per pixel
  for (int dt = 0; dt < 2; ++dt) {
    const size_t bytes = (size_t)N * 4 * H * W * (dt ? 4 : 2);
    const float ms_set = time_us([&]() { cudaMemsetAsync(g, 0, bytes, 0); }, 5);
    printf("flush %s: memset of %zu MB: %.1f us\n", dt ? "f32 " : "bf16", bytes >> 20, ms_set);
    for (int mode = 0; mode < 2; ++mode) {
      for (int chunk = 0; chunk < 2; ++chunk) {
        const int per = chunk ? 16 * 7 : N;         // plane-quads per launch
        auto run = [&]() {
          for (int n0 = 0; n0 < N; n0 += per) {
            dim3 grid(W / 32, H / 32, per);
            if (chunk) cudaMemsetAsync((char*)g + (size_t)n0 * 4 * H * W * (dt ? 4 : 2), 0, (size_t)per * 4 * H * W * (dt ? 4 : 2), 0);
            if (dt == 0) { if (mode == 0) flush_k<__nv_bfloat16, 0><<<grid, 256>>>((__nv_bfloat16*)g, H, W, rows, vecs, n0);
                           else flush_k<__nv_bfloat16, 1><<<grid, 256>>>((__nv_bfloat16*)g, H, W, rows, vecs, n0); }
            else { if (mode == 0) flush_k<float, 0><<<grid, 256>>>((float*)g, H, W, rows, vecs, n0);
                   else flush_k<float, 1><<<grid, 256>>>((float*)g, H, W, rows, vecs, n0); }
          }
        };
        const float us = time_us(run, 5);
        printf("flush %s %-6s %-34s: %8.1f us\n", dt ? "f32 " : "bf16", mode ? "store" : "RED",
               chunk ? "chunks of 16 samples (+ their memset)" : "whole tensor (no memset)", us);
      }
    }
  }
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
