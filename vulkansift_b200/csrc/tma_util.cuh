/*
 * tma_util.cuh -- shared TMA / mbarrier helpers (sm_100a) and the host-side tensor-map encoder.
 * cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint so the library never links libcuda.
 */
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vks
{

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

/* NULL when the driver does not export the symbol */
static inline PFN_encodeTiled tma_encoder()
{
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried)
  {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

/* fp32 (or, with fp16 = true, binary16) tensor of up to 3 dims (x fastest), no swizzle, out-of-bounds elements read as zero */
static inline bool tma_make_map_f32(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                                    const uint32_t *box, bool fp16 = false)
{
  PFN_encodeTiled enc = tma_encoder();
  if (!enc)
    return false;
  cuuint64_t gdim[3];
  cuuint64_t gstr[2];
  cuuint32_t bx[3], es[3];
  for (int i = 0; i < rank; i++)
  {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0)
      gstr[i - 1] = strides_bytes[i - 1];
  }
  return enc(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void *)base, gdim, gstr, bx, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t tma_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tma_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint32_t bar, uint32_t parity)
{
  uint32_t done;
  do
  {
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t"
                 "}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(map),
               "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_f32(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map),
               "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
/* order prior generic-proxy accesses to shared memory before later async-proxy (TMA) writes */
__device__ __forceinline__ void tma_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

} // namespace vks
