/*
 * glsl_emu.h -- just enough GLSL 450 compute semantics to run the reference's shaders on the CPU.
 *
 * TEST INFRASTRUCTURE ONLY.  oracle/build_ref.py reads the reference's *.comp files in place from
 * /root/reference, applies mechanical syntax rewrites (layout qualifiers, interface blocks, float
 * literals, array declarators, signed % -> OpSMod) and compiles them with this prelude into
 * oracle/_ref/libvksift_ref.so.  The algorithm that runs is the reference's own source text.
 *
 * Semantics chosen where GLSL/Vulkan leave room (all documented in SURVEY.md appendix B):
 *   - exp/atan/sin/cos/pow/log2 use include/vksift_arith.h (the builtins are only ULP-bounded)
 *   - round() is round-half-even (B-D8); signed % is OpSMod (B-D7)
 *   - out-of-bounds imageLoad returns 0, out-of-bounds imageStore is dropped (B-D3)
 *   - linear sampling uses exact fp32 weights (B-D1), MIRRORED_REPEAT addressing
 *   - a work group runs as cooperative fibers: every invocation runs until its next barrier(),
 *     in invocation order, so atomics resolve in invocation order
 */
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <ucontext.h>
#include <vector>

#include "../include/vksift_arith.h"

namespace glsl
{
typedef uint32_t uint;

template <typename T>
struct tvec2
{
  union
  {
    T x, r;
  };
  union
  {
    T y, g;
  };
  tvec2() : x(0), y(0) {}
  template <typename A, typename B>
  tvec2(A a, B b) : x((T)a), y((T)b) {}
  template <typename U>
  explicit tvec2(const U &v) : x((T)v.x), y((T)v.y) {}
};
template <typename T>
struct tvec3
{
  union
  {
    T x, r;
  };
  union
  {
    T y, g;
  };
  union
  {
    T z, b;
  };
  tvec3() : x(0), y(0), z(0) {}
  template <typename A, typename B, typename C>
  tvec3(A a, B b_, C c) : x((T)a), y((T)b_), z((T)c) {}
  template <typename U, typename C>
  tvec3(const tvec2<U> &v, C c) : x((T)v.x), y((T)v.y), z((T)c) {}
  template <typename U>
  explicit tvec3(const tvec3<U> &v) : x((T)v.x), y((T)v.y), z((T)v.z) {}
};
template <typename T>
struct tvec4
{
  union
  {
    T x, r;
  };
  union
  {
    T y, g;
  };
  union
  {
    T z, b;
  };
  union
  {
    T w, a;
  };
  tvec4() : x(0), y(0), z(0), w(0) {}
  template <typename A, typename B, typename C, typename D>
  tvec4(A a_, B b_, C c, D d) : x((T)a_), y((T)b_), z((T)c), w((T)d) {}
};
typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec2<int> ivec2;
typedef tvec3<int> ivec3;
typedef tvec3<uint> uvec3;

template <typename T>
inline tvec3<T> operator+(const tvec3<T> &a, const tvec3<T> &b)
{
  return tvec3<T>(a.x + b.x, a.y + b.y, a.z + b.z);
}
template <typename T>
inline tvec3<T> operator-(const tvec3<T> &a, const tvec3<T> &b)
{
  return tvec3<T>(a.x - b.x, a.y - b.y, a.z - b.z);
}

/* ---- images --------------------------------------------------------------- */
struct image2DArray
{
  float *data = nullptr;
  int w = 0, h = 0, layers = 0;
};
struct sampler2DArray
{
  const float *data = nullptr;
  int w = 0, h = 0, layers = 0;
};
inline ivec3 imageSize(const image2DArray &im) { return ivec3(im.w, im.h, im.layers); }
inline ivec3 textureSize(const sampler2DArray &s, int) { return ivec3(s.w, s.h, s.layers); }
inline vec4 imageLoad(const image2DArray &im, const ivec3 &c)
{
  if (c.x < 0 || c.x >= im.w || c.y < 0 || c.y >= im.h || c.z < 0 || c.z >= im.layers)
    return vec4(0.f, 0.f, 0.f, 0.f);
  return vec4(im.data[((size_t)c.z * im.h + c.y) * im.w + c.x], 0.f, 0.f, 1.f);
}
inline void imageStore(image2DArray &im, const ivec3 &c, const vec4 &v)
{
  if (c.x < 0 || c.x >= im.w || c.y < 0 || c.y >= im.h || c.z < 0 || c.z >= im.layers)
    return;
  im.data[((size_t)c.z * im.h + c.y) * im.w + c.x] = v.x;
}
/* VkSampler of sift_detector.c:208-225: LINEAR filter, MIRRORED_REPEAT, normalized coordinates.
 * Vulkan: u = s*W, sample around u-0.5 with weights (1-f, f); the array layer is round(z). */
inline vec4 textureLod(const sampler2DArray &s, const vec3 &c, int)
{
  const float u = c.x * (float)s.w - 0.5f, v = c.y * (float)s.h - 0.5f;
  const float uf = floorf(u), vf = floorf(v);
  const float fx = u - uf, fy = v - vf;
  const int i0 = vks_mirror((int)uf, s.w), i1 = vks_mirror((int)uf + 1, s.w);
  const int j0 = vks_mirror((int)vf, s.h), j1 = vks_mirror((int)vf + 1, s.h);
  int l = (int)rintf(c.z);
  l = l < 0 ? 0 : (l >= s.layers ? s.layers - 1 : l);
  const float *L = s.data + (size_t)l * s.w * s.h;
  const float t00 = L[(size_t)j0 * s.w + i0], t10 = L[(size_t)j0 * s.w + i1];
  const float t01 = L[(size_t)j1 * s.w + i0], t11 = L[(size_t)j1 * s.w + i1];
  const float top = vks_lerp(t00, t10, fx), bot = vks_lerp(t01, t11, fx);
  return vec4(vks_lerp(top, bot, fy), 0.f, 0.f, 1.f);
}

/* ---- builtins --------------------------------------------------------------- */
inline float g_abs(float a) { return fabsf(a); }
inline int g_abs(int a) { return a < 0 ? -a : a; }
inline float g_exp(float a) { return vks_expf(a); }
inline float g_sqrt(float a) { return sqrtf(a); }
inline float g_sqrt(int a) { return sqrtf((float)a); }
inline float g_atan(float y, float x) { return vks_atan2f(y, x); }
inline float g_cos(float a)
{
  float s, c;
  vks_sincosf(a, &s, &c);
  return c;
}
inline float g_sin(float a)
{
  float s, c;
  vks_sincosf(a, &s, &c);
  return s;
}
inline float g_floor(float a) { return floorf(a); }
inline float g_ceil(float a) { return ceilf(a); }
inline float g_round(float a) { return rintf(a); }
/* exact log2 for the only use, ceil(log2(m)): integer for powers of two, strictly between integers otherwise */
inline float g_log2(float m)
{
  const int c = vks_ceil_log2(m);
  const bool pow2 = (vks_f2u(m) & 0x7fffffu) == 0;
  return pow2 ? (float)c : (float)c - 0.5f;
}
inline float g_pow(float a, float b)
{
  if (a == 2.f)
    return vks_exp2f(b);
  if (b == 2.f)
    return a * a;
  return powf(a, b);
}
inline float g_pow(float a, int b) { return g_pow(a, (float)b); }
inline uint g_min(uint a, uint b) { return a < b ? a : b; }
inline int g_min(int a, int b) { return a < b ? a : b; }
inline float g_min(float a, float b) { return a < b ? a : b; }
/* OpSMod: result takes the sign of the divisor (what glslang emits for signed %) */
inline int glsl_mod(int a, int b)
{
  int m = a % b;
  if (m != 0 && ((m < 0) != (b < 0)))
    m += b;
  return m;
}
inline uint glsl_mod(uint a, uint b) { return a % b; }
inline int glsl_mod(int a, uint b) { return glsl_mod(a, (int)b); }
inline uint glsl_mod(uint a, int b) { return a % (uint)b; }

template <typename V>
inline uint atomicAdd(uint &m, V v)
{
  const uint o = m;
  m += (uint)v;
  return o;
}
inline uint atomicMax(uint &m, uint v)
{
  const uint o = m;
  if (v > m)
    m = v;
  return o;
}
template <typename V>
inline uint atomicOr(uint &m, V v)
{
  const uint o = m;
  m |= (uint)v;
  return o;
}
inline void memoryBarrierShared() {}

/* ---- invocation state + fibers ---------------------------------------------- */
extern uvec3 gl_GlobalInvocationID, gl_LocalInvocationID, gl_WorkGroupID, gl_NumWorkGroups;
void barrier();

typedef void (*shader_main_fn)();
/* run groups_x*groups_y*groups_z work groups of local size (lx,ly,lz); with_barriers selects the fiber scheduler */
void dispatch(shader_main_fn fn, uint gx, uint gy, uint gz, uint lx, uint ly, uint lz, bool with_barriers);

} // namespace glsl

/* GLSL builtin names resolve to the emulation (plain names would be ambiguous with <cmath>);
 * defined last so that no standard header sees them */
#define abs g_abs
#define exp g_exp
#define sqrt g_sqrt
#define atan g_atan
#define cos g_cos
#define sin g_sin
#define floor g_floor
#define ceil g_ceil
#define round g_round
#define log2 g_log2
#define pow g_pow
#define min g_min
