/*
 * blur_arith.cuh -- packed fp32 helpers shared by the blur kernels (pyramid.cu, pyramid_strip.cu).
 * add/mul/fma.rn.f32x2 are two IEEE operations each, so the per-pixel operation sequence of include/vksift_arith.h
 * (mul by the centre tap, then fma((a+b), tap_i, acc)) is unchanged by the packing.
 */
#pragma once

#include <cuda_fp16.h>
#include <stdint.h>

namespace vks
{

typedef unsigned long long pk2; /* two packed fp32 */
__device__ __forceinline__ pk2 pk_fma(pk2 a, pk2 b, pk2 c)
{
  pk2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ pk2 pk_add(pk2 a, pk2 b)
{
  pk2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ pk2 pk_sub(pk2 a, pk2 b)
{
  pk2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ pk2 pk_mul(pk2 a, pk2 b)
{
  pk2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float pk_lo(pk2 a) { return __uint_as_float((uint32_t)a); }
__device__ __forceinline__ float pk_hi(pk2 a) { return __uint_as_float((uint32_t)(a >> 32)); }
__device__ __forceinline__ pk2 pk_make(float lo, float hi)
{
  pk2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}

/* predicated 8-byte global stores: the vertical pass stores rows under `row < rows_valid`; as predicated instructions they
 * cost nothing, as branches (what the compiler makes of an if around a store and its operands) three instructions per row */
__device__ __forceinline__ void pk_stg_if(void *ptr, pk2 v, bool ok)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.b64 [%0], %1;\n\t}" ::"l"(ptr), "l"(v), "r"((uint32_t)ok) : "memory");
}
__device__ __forceinline__ void pk_stcs_if(void *ptr, pk2 v, bool ok)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.cs.b64 [%0], %1;\n\t}" ::"l"(ptr), "l"(v), "r"((uint32_t)ok) : "memory");
}
__device__ __forceinline__ void f32_stg_if(void *ptr, float v, bool ok)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.f32 [%0], %1;\n\t}" ::"l"(ptr), "f"(v), "r"((uint32_t)ok) : "memory");
}

/* fp16 precision mode: a value goes through binary16 (round to nearest even) on its way to memory */
__device__ __forceinline__ float round_half1(float v) { return __half2float(__float2half_rn(v)); }
__device__ __forceinline__ pk2 round_half2(pk2 v)
{
  const __half2 h = __floats2half2_rn(pk_lo(v), pk_hi(v));
  const float2 f = __half22float2(h);
  return pk_make(f.x, f.y);
}

/* Programmatic dependent launch: a kernel lets the next launch of its stream be scheduled while it is still
 * running (its CTAs take over SMs as ours drain and run their prologue), and waits for the completion and
 * memory flush of the previous launch before it touches anything that launch wrote. */
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }


} // namespace vks
