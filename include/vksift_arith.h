/*
 * vksift_arith.h -- normative fp32 arithmetic of the SIFT detect path.
 *
 * The reference evaluates exp/atan/sin/cos/pow/log2 with GLSL builtins whose
 * results are only ULP-bounded by the Vulkan spec (exp: 3+2|x| ULP, atan: 4096
 * ULP) and therefore differ between GPU drivers.  Keypoint accept/reject,
 * fixed-point histogram bins and u8 descriptor bytes depend on those values, so
 * a bit-exact CPU<->GPU comparison needs ONE definition of each function that
 * compiles to the same IEEE-754 operation sequence under gcc and under nvcc.
 * This header is that definition: only +,-,*,/ sqrt (all correctly rounded),
 * explicit fused multiply-add and integer bit manipulation are used.
 *
 * Consumers: the CUDA kernels (vulkansift_b200/csrc) and the CPU oracle
 * (oracle/sift_oracle.c).  Build rules that make it hold:
 *   gcc : -ffp-contract=off (no implicit fusion), no -ffast-math
 *   nvcc: -fmad=false, default -prec-div=true -prec-sqrt=true -ftz=false
 * Reference call sites that these functions replace are cited per function
 * (paths relative to the reference repository root).
 */
#ifndef VKSIFT_ARITH_H
#define VKSIFT_ARITH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define VKS_HD __host__ __device__ __forceinline__
#else
#define VKS_HD static inline
#endif

#define VKS_PI_F 3.14159274101257324f     /* float(3.14159265358979323846), shaders' PI */
#define VKS_TWO_PI_F 6.28318548202514648f /* 2.f*PI evaluated in fp32 */
#define VKS_SQRT2_F 1.41421353816986084f  /* sqrt(2) rounded to fp32 */

/* ---- correctly rounded primitives (never contracted) -------------------- */
VKS_HD float vks_fma(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
VKS_HD float vks_mul(float a, float b)
{
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
VKS_HD float vks_add(float a, float b)
{
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
VKS_HD float vks_sub(float a, float b)
{
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
VKS_HD float vks_div(float a, float b)
{
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
VKS_HD float vks_sqrt(float a)
{
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}
/* round-half-to-even, the decision taken for GLSL round() (SURVEY B-D8) */
VKS_HD float vks_rint(float a) { return rintf(a); }

VKS_HD uint32_t vks_f2u(float f)
{
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
VKS_HD float vks_u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

/* exact 2^e for e in [-126,127]; replaces pow(2.f, octave_idx)
 * (ExtractKeypoints.comp:213, ComputeOrientation.comp:69, ComputeDescriptors.comp:106) */
VKS_HD float vks_pow2i(int e) { return vks_u2f((uint32_t)(e + 127) << 23); }

/* ceil(log2(m)) for finite m > 0, exact on the bit pattern; replaces
 * ceil(log2(max_elem_val)) (ComputeOrientation.comp:81, ComputeDescriptors.comp:124) */
VKS_HD int vks_ceil_log2(float m)
{
  uint32_t u = vks_f2u(m);
  int e = (int)((u >> 23) & 0xffu) - 127;
  return (u & 0x7fffffu) ? e + 1 : e;
}

/* e^x.  x <= -87 returns 0 (true value < 2^-125).  ~1 ULP.
 * Replaces GLSL exp() (ComputeOrientation.comp:79,105, ComputeDescriptors.comp:119-121,161) */
VKS_HD float vks_expf(float x)
{
  if (x < -87.0f)
    return 0.0f;
  if (x > 88.0f)
    x = 88.0f;
  float kf = vks_rint(vks_mul(x, 1.44269502162933350f));
  float r = vks_fma(kf, -0.693145751953125f, x);  /* ln2 high part: k*hi is exact */
  r = vks_fma(kf, -1.42860682030941723e-6f, r);   /* ln2 low part */
  float p = 0x1.a1907ep-13f;
  p = vks_fma(p, r, 0x1.6da0ccp-10f);
  p = vks_fma(p, r, 0x1.11109cp-7f);
  p = vks_fma(p, r, 0x1.555474p-5f);
  p = vks_fma(p, r, 0x1.555556p-3f);
  p = vks_fma(p, r, 0.5f);
  float e = vks_fma(vks_mul(r, r), p, r); /* r + r^2 p(r) */
  e = vks_add(e, 1.0f);
  return vks_mul(e, vks_pow2i((int)kf));
}

/* 2^x for x in [-100,100]; replaces pow(2.f, subpix_s/nb_scales) (ExtractKeypoints.comp:218) */
VKS_HD float vks_exp2f(float x)
{
  float kf = vks_rint(x);
  float r = vks_sub(x, kf); /* exact */
  float p = 0x1.00e142p-16f;
  p = vks_fma(p, r, 0x1.4466b0p-13f);
  p = vks_fma(p, r, 0x1.5d873ep-10f);
  p = vks_fma(p, r, 0x1.3b29e6p-7f);
  p = vks_fma(p, r, 0x1.c6b08ep-5f);
  p = vks_fma(p, r, 0x1.ebfbe0p-3f);
  p = vks_fma(p, r, 0x1.62e430p-1f);
  float e = vks_fma(p, r, 1.0f);
  return vks_mul(e, vks_pow2i((int)kf));
}

/* atan2(y,x) in (-pi,pi], atan2(0,0) := 0.  ~2 ULP.
 * Replaces GLSL atan(y,x) (ComputeOrientation.comp:106, ComputeDescriptors.comp:145) */
VKS_HD float vks_atan2f(float y, float x)
{
  float ax = fabsf(x), ay = fabsf(y);
  float mx = ax > ay ? ax : ay;
  float mn = ax > ay ? ay : ax;
  if (mx == 0.0f)
    return 0.0f;
  float a = vks_div(mn, mx);
  float s = vks_mul(a, a);
  float p = 0x1.84c188p-9f;
  p = vks_fma(p, s, -0x1.0f33dcp-6f);
  p = vks_fma(p, s, 0x1.648358p-5f);
  p = vks_fma(p, s, -0x1.366f3ap-4f);
  p = vks_fma(p, s, 0x1.b5689ap-4f);
  p = vks_fma(p, s, -0x1.231c9ep-3f);
  p = vks_fma(p, s, 0x1.997b36p-3f);
  p = vks_fma(p, s, -0x1.5554eap-2f);
  float r = vks_fma(vks_mul(a, s), p, a);
  if (ay > ax)
    r = vks_sub(1.57079637050628662f, r);
  if (x < 0.0f)
    r = vks_sub(VKS_PI_F, r);
  if (y < 0.0f)
    r = -r;
  return r;
}

/* sin and cos of t, |t| <= 64.  ~1 ULP on the path's range [0, 2pi].
 * Replaces GLSL cos()/sin() (ComputeDescriptors.comp:110-111) */
VKS_HD void vks_sincosf(float t, float *sn, float *cs)
{
  float kf = vks_rint(vks_mul(t, 0.636619746685028076f));
  float r = vks_fma(kf, -1.57079637050628662f, t);
  r = vks_fma(kf, 4.37113882867379223e-8f, r); /* pi/2 = hi - 4.37e-8 */
  float s = vks_mul(r, r);
  float ps = 0x1.6cca8ep-19f;
  ps = vks_fma(ps, s, -0x1.a00f6ep-13f);
  ps = vks_fma(ps, s, 0x1.111108p-7f);
  ps = vks_fma(ps, s, -0x1.555556p-3f);
  float sv = vks_fma(vks_mul(r, s), ps, r);
  float pc = 0x1.99e24cp-16f;
  pc = vks_fma(pc, s, -0x1.6c0c1ep-10f);
  pc = vks_fma(pc, s, 0x1.55554ap-5f);
  float cv = vks_fma(vks_mul(s, s), pc, vks_fma(s, -0.5f, 1.0f));
  int q = ((int)kf) & 3;
  float so = (q & 1) ? cv : sv;
  float co = (q & 1) ? sv : cv;
  if (q == 1 || q == 2)
    co = -co;
  if (q >= 2)
    so = -so;
  *sn = so;
  *cs = co;
}

/* ---- scale-space arithmetic --------------------------------------------- */

/* MIRRORED_REPEAT addressing of the blur sampler (sift_detector.c:214-216):
 * -1->0, -2->1, n->n-1, n+1->n-2, period 2n. */
VKS_HD int vks_mirror(int i, int n)
{
  int p = 2 * n;
  int m = i % p;
  if (m < 0)
    m += p;
  return m < n ? m : p - 1 - m;
}

/* R8_UNORM -> float of vkCmdCopyBufferToImage + sampling (sift_detector.c:881) */
VKS_HD float vks_unorm8(uint8_t v) { return vks_div((float)v, 255.0f); }

/* a*(1-f) + b*f with one rounding less: fma(b, f, a*(1-f)); the LINEAR blit
 * (sift_detector.c:909-916) only ever uses f in {0, .25, .75} */
VKS_HD float vks_lerp(float a, float b, float f) { return vks_fma(b, f, vks_mul(a, vks_sub(1.0f, f))); }

/* One separable blur tap pair: acc += (a+b)*k
 * (GaussianBlur.comp:35,41; GaussianBlurInterpolated.comp:36,42 through the
 * effective-tap table) */
VKS_HD float vks_blur_tap(float acc, float a, float b, float k) { return vks_fma(vks_add(a, b), k, acc); }

#endif /* VKSIFT_ARITH_H */
