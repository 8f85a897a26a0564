/*
 * layer_io.cuh -- element access to the scale-space layers, which hold fp32 or binary16 values
 * (VKSIFT_PYRAMID_PRECISION_FLOAT16: the reference stores R16_SFLOAT images, sift_memory.c:139).  Arithmetic is fp32 in both
 * modes; a binary16 layer value is the fp32 result rounded to nearest even once, when it is stored (SURVEY B-D11).
 */
#pragma once

#include <cuda_fp16.h>
#include <stdint.h>

namespace vks
{

__device__ __forceinline__ float layer_ld(const void *base, size_t idx, int fp16)
{
  return fp16 ? __half2float(__ldg(reinterpret_cast<const __half *>(base) + idx)) : __ldg(reinterpret_cast<const float *>(base) + idx);
}
/* v is already rounded through binary16 in fp16 mode (the callers round once and reuse the rounded value) */
__device__ __forceinline__ void layer_st(void *base, size_t idx, float v, int fp16)
{
  if (fp16)
    reinterpret_cast<__half *>(base)[idx] = __float2half_rn(v);
  else
    reinterpret_cast<float *>(base)[idx] = v;
}
__host__ __device__ __forceinline__ const void *layer_ptr(const void *base, size_t idx, int fp16)
{
  return reinterpret_cast<const char *>(base) + idx * (fp16 ? 2u : 4u);
}
__host__ __device__ __forceinline__ void *layer_ptr(void *base, size_t idx, int fp16) { return reinterpret_cast<char *>(base) + idx * (fp16 ? 2u : 4u); }

} // namespace vks
