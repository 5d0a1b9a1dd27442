// Microbenchmark: L2 atomic throughput for a bilinear scatter, scalar REDs into planar fp32 vs
// vector red.v4.f32 into interleaved fp32.   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void red_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__global__ void scalar_k(float* g, int N, int H, int W, float a00, float a01, float a10, float a11) {
  int j = blockIdx.x * 32 + (threadIdx.x & 31), i = blockIdx.y * 8 + (threadIdx.x >> 5), n = blockIdx.z;
  float ix = a00 * j + a01 * i + 3.3f + n * 0.37f, iy = a10 * j + a11 * i + 2.7f;
  int x0 = (int)floorf(ix), y0 = (int)floorf(iy);
  float fx = ix - x0, fy = iy - y0;
  float* base = g + (size_t)n * 4 * H * W;
  for (int c = 0; c < 4; ++c) {
    float v = 1.f + c;
    float* p = base + (size_t)c * H * W;
    if (x0 >= 0 && x0 + 1 < W && y0 >= 0 && y0 + 1 < H) {
      atomicAdd(p + y0 * W + x0, v * (1 - fx) * (1 - fy));
      atomicAdd(p + y0 * W + x0 + 1, v * fx * (1 - fy));
      atomicAdd(p + (y0 + 1) * W + x0, v * (1 - fx) * fy);
      atomicAdd(p + (y0 + 1) * W + x0 + 1, v * fx * fy);
    }
  }
}
__global__ void vector_k(float* g, int N, int H, int W, float a00, float a01, float a10, float a11) {
  int j = blockIdx.x * 32 + (threadIdx.x & 31), i = blockIdx.y * 8 + (threadIdx.x >> 5), n = blockIdx.z;
  float ix = a00 * j + a01 * i + 3.3f + n * 0.37f, iy = a10 * j + a11 * i + 2.7f;
  int x0 = (int)floorf(ix), y0 = (int)floorf(iy);
  float fx = ix - x0, fy = iy - y0;
  float* base = g + (size_t)n * 4 * H * W;
  if (x0 >= 0 && x0 + 1 < W && y0 >= 0 && y0 + 1 < H) {
    float w00 = (1 - fx) * (1 - fy), w01 = fx * (1 - fy), w10 = (1 - fx) * fy, w11 = fx * fy;
    float* p = base + ((size_t)y0 * W + x0) * 4;
    red_v4(p, w00, 2 * w00, 3 * w00, 4 * w00);
    red_v4(p + 4, w01, 2 * w01, 3 * w01, 4 * w01);
    red_v4(p + 4 * W, w10, 2 * w10, 3 * w10, 4 * w10);
    red_v4(p + 4 * W + 4, w11, 2 * w11, 3 * w11, 4 * w11);
  }
}
int main() {
  const int N = 448, H = 256, W = 256;
  float* g;
  cudaMalloc(&g, (size_t)N * 4 * H * W * 4);
  cudaMemset(g, 0, (size_t)N * 4 * H * W * 4);
  dim3 grid(W / 32, H / 8, N);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float cfg[3][4] = {{1, 0, 0, 1}, {0.9f, 0.2f, -0.2f, 0.9f}, {0.7f, 0.7f, -0.7f, 0.7f}};
  for (int c = 0; c < 3; ++c) {
    for (int mode = 0; mode < 2; ++mode) {
      float ms = 0;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) scalar_k<<<grid, 256>>>(g, N, H, W, cfg[c][0], cfg[c][1], cfg[c][2], cfg[c][3]);
        else vector_k<<<grid, 256>>>(g, N, H, W, cfg[c][0], cfg[c][1], cfg[c][2], cfg[c][3]);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("cfg %d %s: %.1f us  (%.1f Gpx-layer/s)\n", c, mode ? "red.v4 interleaved" : "scalar planar    ", ms * 1e3,
             (double)N * H * W / ms / 1e6);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
