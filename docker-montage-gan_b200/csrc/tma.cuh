// TMA plumbing for the translation (stencil) kernels: tensor maps over the planar tensors (host) and the
// cp.async.bulk.tensor / mbarrier wrappers (device).
//
// Why TMA here: under a pure translation the footprint of an output tile is an axis-aligned RECTANGLE of each channel
// plane whose origin is an arbitrary integer texel (j0 + X, i0 + Y).  A tiled tensor-map copy takes exactly that -- any
// element coordinate, negative or past the end, with the out-of-range part zero-filled -- and lands the box in shared
// memory with no staging instructions, no registers in flight and no alignment case analysis (BASELINE north_star:
// "TMA where the footprint is rectangular").  The planes stay planar in shared memory: with layer-uniform bilinear weights
// a thread works on runs of horizontally adjacent texels of one channel, so there is nothing to interleave.
//
// The driver entry point is fetched through the runtime (cudaGetDriverEntryPoint): the library does not link libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mgr_common.cuh"

namespace mgr {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tma_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// A tiled map over `rank` dimensions (innermost first).  Elements are moved as opaque 2- or 4-byte integers.  Returns
// false when the tensor does not meet TMA's rules (16-byte aligned base, strides multiples of 16 bytes, box <= 256 per
// dimension, inner box extent a multiple of 16 bytes): the caller then keeps the non-TMA kernel.
inline bool tma_make_map(CUtensorMap* map, const void* base, int elem_bytes, int rank, const long long* dims,
                         const long long* strides_elems /* [rank], strides_elems[0] == 1 */, const int* box) {
  EncodeTiledFn fn = tma_encode_fn();
  if (!fn || rank < 1 || rank > 5) return false;
  if (reinterpret_cast<uintptr_t>(base) % 16) return false;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int d = 0; d < rank; ++d) {
    if (dims[d] < 1 || dims[d] > 0xffffffffLL || box[d] < 1 || box[d] > 256) return false;
    gdim[d] = (cuuint64_t)dims[d]; bx[d] = (cuuint32_t)box[d]; es[d] = 1;
    if (d > 0) {
      long long sb = strides_elems[d] * elem_bytes;
      if (dims[d] == 1)      // never stepped: any legal value (views of one sample / one layer carry arbitrary strides here)
        sb = d == 1 ? ((dims[0] * elem_bytes + 15) & ~15LL) : (long long)gstr[d - 2] * dims[d - 1];
      if (sb <= 0 || sb % 16 || sb >= (1LL << 40)) return false;
      gstr[d - 1] = (cuuint64_t)sb;
    }
  }
  if (strides_elems[0] != 1 || (box[0] * elem_bytes) % 16) return false;
  const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16;
  return fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- device side ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tma_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barriers visible to the async proxy (the TMA unit arrives on them)
__device__ __forceinline__ void tma_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// order this thread's generic-proxy shared-memory accesses before later async-proxy (TMA) accesses
__device__ __forceinline__ void tma_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tma_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "TMA_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"      // suspends up to the hint, wakes when the phase completes
      "@P1 bra TMA_DONE;\n"
      "bra TMA_WAIT;\n"
      "TMA_DONE:\n"
      "}" ::"r"(tma_smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// box copies global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(tma_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(tma_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(tma_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(tma_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// named barrier among `count` threads (count a multiple of 32); id 0 is __syncthreads()
__device__ __forceinline__ void named_barrier(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

}  // namespace mgr
