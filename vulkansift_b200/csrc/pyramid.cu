/*
 * pyramid.cu -- Gaussian scale space + DoG for sm_100a.
 *
 * One launch ("step") runs one or more blur passes; a pass produces one
 * Gaussian layer from the previous one and everything that hangs off it:
 *   - H and V separable passes fused through shared memory
 *     (reference: two dispatches of GaussianBlur[Interpolated].comp per layer,
 *      sift_detector.c:955-1001)
 *   - DoG[s-1] = G[s] - G[s-1] written from the same tile
 *     (reference: DifferenceOfGaussian.comp:12-15, sift_detector.c:1039-1079)
 *   - layer ns additionally seeds the next octave: NEAREST blit, dst(i,j) =
 *     src(2i+1,2j+1)  (reference: vkCmdBlitImage, sift_detector.c:1003-1034)
 *   - the first pass of octave 0 reads the u8 input and applies the UNORM
 *     conversion and the LINEAR 2x blit on the fly
 *     (reference: sift_detector.c:860-916)
 * Per-pixel arithmetic is the normative sequence of include/vksift_arith.h
 * (mul by the centre tap, then fma((a+b), tap_i, acc) for i = 1..radius), so
 * results do not depend on the tiling.
 */
#include "vksift_internal.h"

namespace vks
{

/* ---- source fetch with MIRRORED_REPEAT addressing ----------------------- */
__device__ __forceinline__ float fetch_u8_up2(const uint8_t *__restrict__ img, int sw, int sh, int x, int y)
{
  /* LINEAR blit, scale 0.5: u = (x+0.5)*0.5 - 0.5 -> even x: (k-1,k) f=.75 ; odd x: (k,k+1) f=.25 */
  const int kx = x >> 1, ky = y >> 1;
  int x0, x1, y0, y1;
  float fx, fy;
  if (x & 1)
  {
    x0 = kx;
    x1 = kx + 1;
    fx = 0.25f;
  }
  else
  {
    x0 = kx - 1;
    x1 = kx;
    fx = 0.75f;
  }
  if (y & 1)
  {
    y0 = ky;
    y1 = ky + 1;
    fy = 0.25f;
  }
  else
  {
    y0 = ky - 1;
    y1 = ky;
    fy = 0.75f;
  }
  x0 = max(x0, 0);
  y0 = max(y0, 0);
  x1 = min(x1, sw - 1);
  y1 = min(y1, sh - 1);
  const float t00 = vks_unorm8(img[(size_t)y0 * sw + x0]);
  const float t10 = vks_unorm8(img[(size_t)y0 * sw + x1]);
  const float t01 = vks_unorm8(img[(size_t)y1 * sw + x0]);
  const float t11 = vks_unorm8(img[(size_t)y1 * sw + x1]);
  const float top = vks_lerp(t00, t10, fx);
  const float bot = vks_lerp(t01, t11, fx);
  return vks_lerp(top, bot, fy);
}

__device__ __forceinline__ float fetch_src(const BlurPass &p, int x, int y)
{
  x = vks_mirror(x, p.w);
  y = vks_mirror(y, p.h);
  if (p.src_kind == BLUR_SRC_LAYER)
    return ((const float *)p.src)[(size_t)y * p.src_pitch + x];
  if (p.src_kind == BLUR_SRC_U8_UP2)
    return fetch_u8_up2((const uint8_t *)p.src, p.src_w, p.src_h, x, y);
  return vks_unorm8(((const uint8_t *)p.src)[(size_t)y * p.src_w + x]);
}

/* ---- compact tile kernel ----------------------------------------------------
 * Any radius (<= 20), 32x16 output tile, 256 threads, rolled tap loops: a few hundred bytes of
 * code.  It runs the small octaves, where a launch is latency bound and the unrolled fast kernel
 * below loses more time fetching its straight-line code than computing, and every pass whose
 * radius the fast kernel does not cover.  Same per-pixel operation sequence as everywhere. */
#define SB_W 32
#define SB_H 16
#define SB_RMAX 20
#define SB_IN_W (SB_W + 2 * SB_RMAX)
#define SB_IN_H (SB_H + 2 * SB_RMAX)

/* reflect once on either side: valid for -n <= i < 2n */
__device__ __forceinline__ int mirror_once(int i, int n)
{
  i = i < 0 ? -1 - i : i;
  return i >= n ? 2 * n - 1 - i : i;
}

__global__ void __launch_bounds__(256) blur_step_small_kernel(const __grid_constant__ BlurStep S)
{
  __shared__ float s_in[SB_IN_H * SB_IN_W];
  __shared__ float s_mid[SB_IN_H * SB_W];

  int pi = 0;
#pragma unroll
  for (int i = 1; i < VKS_MAX_PASSES_PER_STEP; i++)
    if (i < S.n_pass && (int)blockIdx.x >= S.pass[i].tile_begin)
      pi = i;
  const BlurPass &p = S.pass[pi];
  const int t = (int)blockIdx.x - p.tile_begin;
  const int x0 = (t % p.tiles_x) * SB_W;
  const int y0 = (t / p.tiles_x) * SB_H;
  const int R = p.radius;
  const int in_w = SB_W + 2 * R;
  const int rows_valid = min(SB_H, p.h - y0);
  const int in_h = rows_valid + 2 * R;
  const int tid = threadIdx.x;
  const int n_el = in_w * in_h;

  if (p.src_kind == BLUR_SRC_LAYER)
  {
    const bool once = (x0 - R >= -p.w) && (x0 + SB_W + R <= 2 * p.w) && (y0 - R >= -p.h) && (y0 + SB_H + R <= 2 * p.h);
    const float *__restrict__ src = (const float *)p.src;
    for (int i0 = tid; i0 < n_el; i0 += 8 * 256)
    {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; k++)
      {
        const int i = i0 + k * 256;
        v[k] = 0.f;
        if (i < n_el)
        {
          const int m = i / in_w, c = i - m * in_w;
          int gx = x0 - R + c, gy = y0 - R + m;
          gx = once ? mirror_once(gx, p.w) : vks_mirror(gx, p.w);
          gy = once ? mirror_once(gy, p.h) : vks_mirror(gy, p.h);
          v[k] = __ldg(src + (size_t)gy * p.src_pitch + gx);
        }
      }
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (i0 + k * 256 < n_el)
          s_in[i0 + k * 256] = v[k];
    }
  }
  else
  {
    for (int i = tid; i < n_el; i += 256)
    {
      const int m = i / in_w, c = i - m * in_w;
      s_in[i] = fetch_src(p, x0 - R + c, y0 - R + m);
    }
  }
  __syncthreads();

  for (int i = tid; i < in_h * SB_W; i += 256)
  {
    const int yy = i / SB_W, xx = i - yy * SB_W;
    const float *row = s_in + yy * in_w + xx + R;
    float acc = vks_mul(row[0], p.taps[0]);
    for (int k = 1; k <= R; k++)
      acc = vks_blur_tap(acc, row[k], row[-k], p.taps[k]);
    s_mid[i] = acc;
  }
  __syncthreads();

  for (int i = tid; i < rows_valid * SB_W; i += 256)
  {
    const int yy = i / SB_W, xx = i - yy * SB_W;
    const int x = x0 + xx, y = y0 + yy;
    if (x >= p.w)
      continue;
    const float *col = s_mid + (yy + R) * SB_W + xx;
    float acc = vks_mul(col[0], p.taps[0]);
    for (int k = 1; k <= R; k++)
      acc = vks_blur_tap(acc, col[k * SB_W], col[-k * SB_W], p.taps[k]);
    p.dst_g[(size_t)y * p.dst_pitch + x] = acc;
    if (p.dst_d)
      p.dst_d[(size_t)y * p.dst_pitch + x] = vks_sub(acc, s_in[(yy + R) * in_w + xx + R]);
    if (p.dst_next && (x & 1) && (y & 1))
    {
      const int nx = x >> 1, ny = y >> 1;
      if (nx < p.next_w && ny < p.next_h)
        p.dst_next[(size_t)ny * p.next_pitch + nx] = acc;
    }
  }
}

/* ==========================================================================
 * Fast tile kernel: even radius <= 12 (every scale of the default configuration).
 *
 * FP32 issue rate, not HBM, is what limits this stage on B200 (about 2R+1 fp32
 * operations per pixel and pass), so the arithmetic runs on the packed
 * FADD2/FFMA2/FMUL2 pipe (add/fma/mul.rn.f32x2, two IEEE results per
 * instruction, bit-identical to the scalar sequence of vksift_arith.h):
 *   - 64x128 output tile, 256 threads, two CTAs per SM
 *   - input tile + halo in smem with ROW PAIRS interleaved ([y/2][x][y&1]) so the
 *     horizontal pass packs (row y, row y+1) and reads aligned pairs for every tap
 *   - horizontal pass: one thread = 2 rows x 8 columns, sliding window of 8+2R
 *     packed values in registers; result row-major in smem
 *   - vertical pass: one thread = 2 columns x 8 rows, packs (x, x+1); writes G with
 *     64-bit stores, DoG = G - centre from the input tile, and the decimated seed
 *   - taps live in uniform registers as (k,k) pairs straight from the kernel parameters
 * ========================================================================== */
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
__device__ __forceinline__ pk2 pk_make(float lo, float hi) { return (pk2)__float_as_uint(lo) | ((pk2)__float_as_uint(hi) << 32); }

#define FT_W 64
#define FT_H 128
#define FT_THREADS 256
#define FT_MS (FT_W + 2) /* row stride of the horizontal-pass result, floats */

__host__ __device__ constexpr int ft_rx(int R) { return (R + 3) & ~3; }                              /* x halo rounded to float4 */
__host__ __device__ constexpr int ft_ss(int R) { return (((FT_W + 2 * ft_rx(R)) * 2 / 4) | 1) * 4; } /* floats per row pair, stride/4 odd */
__host__ __device__ constexpr int ft_smem_floats(int R, int TH) { return ((TH + 2 * R) / 2) * ft_ss(R) + (TH + 2 * R) * FT_MS; }

template <int R, int TH>
__device__ __forceinline__ void blur_tile_fast(const BlurPass &p, const pk2 *__restrict__ taps2, float *smem, int x0, int y0)
{
  constexpr int RX = ft_rx(R);
  constexpr int SS = ft_ss(R);
  constexpr int IN_W = FT_W + 2 * RX; /* columns loaded */
  constexpr int IN_H = TH + 2 * R;    /* rows loaded */
  constexpr int NRP = IN_H / 2;
  float *s_in = smem;
  float *s_mid = smem + NRP * SS;
  const int tid = threadIdx.x;

  /* rows / columns of this tile that can influence a pixel inside the image */
  const int rows_valid = min(TH, p.h - y0);                 /* output rows */
  const int nrp = min(NRP, (rows_valid + 2 * R + 1) / 2);   /* row pairs the vertical pass will read */
  const int ncg = min(FT_W / 8, (p.w - x0 + 7) / 8);        /* 8-column groups with at least one pixel */

  /* ---- stage 1: source -> smem, row pairs interleaved ---- */
  const bool interior = (p.src_kind == BLUR_SRC_LAYER) && (x0 - RX >= 0) && (x0 + FT_W + RX <= p.w) && (y0 - R >= 0) && (y0 + TH + R <= p.h);
  if (interior)
  {
    const float *__restrict__ src = (const float *)p.src + (size_t)(y0 - R) * p.src_pitch + (x0 - RX);
    constexpr int C4 = IN_W / 4;
    constexpr int ITEMS = NRP * C4;
#pragma unroll 4
    for (int it = tid; it < ITEMS; it += FT_THREADS)
    {
      const int rp = it / C4, c4 = it - rp * C4;
      const float4 a = __ldg((const float4 *)(src + (size_t)(2 * rp) * p.src_pitch) + c4);
      const float4 b = __ldg((const float4 *)(src + (size_t)(2 * rp + 1) * p.src_pitch) + c4);
      float4 *d = (float4 *)(s_in + rp * SS + c4 * 8);
      d[0] = make_float4(a.x, b.x, a.y, b.y);
      d[1] = make_float4(a.z, b.z, a.w, b.w);
    }
  }
  else
  {
    const bool once = (x0 - RX >= -p.w) && (x0 + FT_W + RX <= 2 * p.w) && (y0 - R >= -p.h) && (y0 + TH + R <= 2 * p.h);
    const int n_el = IN_W * 2 * nrp;
    if (p.src_kind == BLUR_SRC_LAYER)
    {
      const float *__restrict__ src = (const float *)p.src;
#pragma unroll 8
      for (int i = tid; i < n_el; i += FT_THREADS)
      {
        const int m = i / IN_W, c = i - m * IN_W;
        int gx = x0 - RX + c, gy = y0 - R + m;
        gx = once ? mirror_once(gx, p.w) : vks_mirror(gx, p.w);
        gy = once ? mirror_once(gy, p.h) : vks_mirror(gy, p.h);
        s_in[(m >> 1) * SS + c * 2 + (m & 1)] = __ldg(src + (size_t)gy * p.src_pitch + gx);
      }
    }
    else
    {
      /* octave 0 seed: stage the u8 source window as float (one UNORM division per source pixel),
       * then apply the LINEAR 2x blit (or the 1:1 copy) from shared memory */
      const bool up = (p.src_kind == BLUR_SRC_U8_UP2);
      const uint8_t *__restrict__ img = (const uint8_t *)p.src;
      /* mirrored destination coordinates stay inside [lo, hi] of the image */
      const int dx_lo = max(0, min(x0 - RX, p.w - 1)), dx_hi = min(p.w - 1, max(0, x0 + FT_W + RX - 1));
      const int dy_lo = max(0, min(y0 - R, p.h - 1)), dy_hi = min(p.h - 1, max(0, y0 + 2 * nrp - R - 1));
      const bool full = !once; /* tiny images reflect more than once: stage the whole source */
      const int sx_lo = full ? 0 : (up ? max(0, (dx_lo >> 1) - 1) : dx_lo);
      const int sx_hi = full ? p.src_w - 1 : (up ? min(p.src_w - 1, (dx_hi >> 1) + 1) : dx_hi);
      const int sy_lo = full ? 0 : (up ? max(0, (dy_lo >> 1) - 1) : dy_lo);
      const int sy_hi = full ? p.src_h - 1 : (up ? min(p.src_h - 1, (dy_hi >> 1) + 1) : dy_hi);
      const int sw = sx_hi - sx_lo + 1, sh = sy_hi - sy_lo + 1;
      const bool staged = (sw * sh <= IN_H * FT_MS);
      if (staged)
      {
#pragma unroll 4
        for (int i = tid; i < sw * sh; i += FT_THREADS)
        {
          const int yy = i / sw, xx = i - yy * sw;
          s_mid[i] = vks_unorm8(__ldg(img + (size_t)(sy_lo + yy) * p.src_w + sx_lo + xx));
        }
        __syncthreads();
        for (int i = tid; i < n_el; i += FT_THREADS)
        {
          const int m = i / IN_W, c = i - m * IN_W;
          int gx = x0 - RX + c, gy = y0 - R + m;
          gx = once ? mirror_once(gx, p.w) : vks_mirror(gx, p.w);
          gy = once ? mirror_once(gy, p.h) : vks_mirror(gy, p.h);
          float v;
          if (up)
          {
            const int kx = gx >> 1, ky = gy >> 1;
            /* the clamps to the staged window only ever bind for cells that no in-image output reads */
            const int ax = min(max(max(((gx & 1) ? kx : kx - 1), 0) - sx_lo, 0), sw - 1);
            const int bx = min(max(min(((gx & 1) ? kx + 1 : kx), p.src_w - 1) - sx_lo, 0), sw - 1);
            const int ay = min(max(max(((gy & 1) ? ky : ky - 1), 0) - sy_lo, 0), sh - 1);
            const int by = min(max(min(((gy & 1) ? ky + 1 : ky), p.src_h - 1) - sy_lo, 0), sh - 1);
            const float fx = (gx & 1) ? 0.25f : 0.75f, fy = (gy & 1) ? 0.25f : 0.75f;
            const float top = vks_lerp(s_mid[ay * sw + ax], s_mid[ay * sw + bx], fx);
            const float bot = vks_lerp(s_mid[by * sw + ax], s_mid[by * sw + bx], fx);
            v = vks_lerp(top, bot, fy);
          }
          else
            v = s_mid[min(max(gy - sy_lo, 0), sh - 1) * sw + min(max(gx - sx_lo, 0), sw - 1)];
          s_in[(m >> 1) * SS + c * 2 + (m & 1)] = v;
        }
      }
      else
      {
        for (int i = tid; i < n_el; i += FT_THREADS)
        {
          const int m = i / IN_W, c = i - m * IN_W;
          s_in[(m >> 1) * SS + c * 2 + (m & 1)] = fetch_src(p, x0 - RX + c, y0 - R + m);
        }
      }
    }
  }
  __syncthreads();

  /* ---- stage 2: horizontal pass, unit = (row pair, 8 columns), lanes along row pairs ---- */
  for (int u = tid; u < nrp * ncg; u += FT_THREADS)
  {
    const int cg = u / nrp, rp = u - cg * nrp;
    const ulonglong2 *wsrc = (const ulonglong2 *)(s_in + rp * SS + (cg * 8 + RX - R) * 2);
    pk2 wv[8 + 2 * R];
#pragma unroll
    for (int j = 0; j < (8 + 2 * R) / 2; j++)
    {
      const ulonglong2 v = wsrc[j];
      wv[2 * j] = v.x;
      wv[2 * j + 1] = v.y;
    }
    pk2 o[8];
#pragma unroll
    for (int q = 0; q < 8; q++)
    {
      pk2 acc = pk_mul(wv[R + q], taps2[0]);
#pragma unroll
      for (int i = 1; i <= R; i++)
        acc = pk_fma(pk_add(wv[R + q + i], wv[R + q - i]), taps2[i], acc);
      o[q] = acc;
    }
    float *r0 = s_mid + (2 * rp) * FT_MS + cg * 8;
    float *r1 = r0 + FT_MS;
#pragma unroll
    for (int q = 0; q < 8; q += 2)
    {
      *(float2 *)(r0 + q) = make_float2(pk_lo(o[q]), pk_lo(o[q + 1]));
      *(float2 *)(r1 + q) = make_float2(pk_hi(o[q]), pk_hi(o[q + 1]));
    }
  }
  __syncthreads();

  /* ---- stage 3: vertical pass, unit = (column pair, 8 rows), lanes along column pairs ---- */
  for (int u = tid; u < (FT_W / 2) * (TH / 8); u += FT_THREADS)
  {
    const int rg = u / (FT_W / 2), cp = u - rg * (FT_W / 2);
    const int x = x0 + 2 * cp;
    const int yb = y0 + rg * 8;
    if (x >= p.w || yb >= p.h)
      continue;
    pk2 wv[8 + 2 * R];
#pragma unroll
    for (int j = 0; j < 8 + 2 * R; j++)
      wv[j] = *(const pk2 *)(s_mid + (rg * 8 + j) * FT_MS + 2 * cp);
    pk2 o[8];
#pragma unroll
    for (int q = 0; q < 8; q++)
    {
      pk2 acc = pk_mul(wv[R + q], taps2[0]);
#pragma unroll
      for (int i = 1; i <= R; i++)
        acc = pk_fma(pk_add(wv[R + q + i], wv[R + q - i]), taps2[i], acc);
      o[q] = acc;
    }
    const bool pair_ok = (x + 1 < p.w);
#pragma unroll
    for (int q = 0; q < 8; q += 2)
    {
      /* centre values of rows q, q+1: {in[y][x], in[y+1][x], in[y][x+1], in[y+1][x+1]} */
      const float4 c = *(const float4 *)(s_in + ((R + rg * 8 + q) >> 1) * SS + (RX + 2 * cp) * 2);
#pragma unroll
      for (int e = 0; e < 2; e++)
      {
        const int y = yb + q + e;
        if (y >= p.h)
          continue;
        const pk2 g = o[q + e];
        const pk2 cen = e ? pk_make(c.y, c.w) : pk_make(c.x, c.z);
        float *gp = p.dst_g + (size_t)y * p.dst_pitch + x;
        if (pair_ok)
          *(pk2 *)gp = g;
        else
          *gp = pk_lo(g);
        if (p.dst_d)
        {
          const pk2 d = pk_sub(g, cen);
          float *dp = p.dst_d + (size_t)y * p.dst_pitch + x;
          if (pair_ok)
            *(pk2 *)dp = d;
          else
            *dp = pk_lo(d);
        }
        if (p.dst_next && (y & 1) && pair_ok)
        {
          /* x is even: the odd column of the pair feeds next(x>>1, y>>1) */
          const int nx = x >> 1, ny = y >> 1;
          if (nx < p.next_w && ny < p.next_h)
            p.dst_next[(size_t)ny * p.next_pitch + nx] = pk_hi(g);
        }
      }
    }
  }
}

struct BlurStepFast
{
  BlurStep s;
  float2 taps2[VKS_MAX_PASSES_PER_STEP][14]; /* (k,k) pairs, zero padded to the even radius */
};

template <int TH>
__device__ __forceinline__ void blur_tile_dispatch(const BlurPass &p, const pk2 *taps2, float *smem, int x0, int y0)
{
  switch ((p.radius + 1) & ~1)
  {
  case 2:
    blur_tile_fast<2, TH>(p, taps2, smem, x0, y0);
    break;
  case 4:
    blur_tile_fast<4, TH>(p, taps2, smem, x0, y0);
    break;
  case 6:
    blur_tile_fast<6, TH>(p, taps2, smem, x0, y0);
    break;
  case 8:
    blur_tile_fast<8, TH>(p, taps2, smem, x0, y0);
    break;
  case 10:
    blur_tile_fast<10, TH>(p, taps2, smem, x0, y0);
    break;
  default:
    blur_tile_fast<12, TH>(p, taps2, smem, x0, y0);
    break;
  }
}

__global__ void __launch_bounds__(FT_THREADS, 2) blur_step_fast_kernel(const __grid_constant__ BlurStepFast S)
{
  extern __shared__ __align__(16) float ft_smem[];
  int pi = 0;
#pragma unroll
  for (int i = 1; i < VKS_MAX_PASSES_PER_STEP; i++)
    if (i < S.s.n_pass && (int)blockIdx.x >= S.s.pass[i].tile_begin)
      pi = i;
  const BlurPass &p = S.s.pass[pi];
  const pk2 *taps2 = reinterpret_cast<const pk2 *>(S.taps2[pi]);
  const int t = (int)blockIdx.x - p.tile_begin;
  const int x0 = (t % p.tiles_x) * FT_W;
  const int y0 = (t / p.tiles_x) * FT_H;
  blur_tile_dispatch<FT_H>(p, taps2, ft_smem, x0, y0);
}

/* A pass goes to the fast kernel when its radius is covered and the layer is large enough for
 * throughput to matter (more 64x128 tiles than SMs); everything else runs on the compact kernel. */
bool blur_pass_is_fast(const BlurPass &bp)
{
  if (bp.radius < 1 || bp.radius > 12)
    return false;
  return ((bp.w + FT_W - 1) / FT_W) * ((bp.h + FT_H - 1) / FT_H) >= 148;
}

bool blur_step_is_fast(const BlurStep &step) { return step.n_pass > 0 && blur_pass_is_fast(step.pass[0]); }

void blur_step_tiles(BlurStep *step)
{
  /* tile geometry depends on the kernel that will run the step; a step never mixes the two kinds */
  const bool fast = blur_step_is_fast(*step);
  const int tw = fast ? FT_W : SB_W, th = fast ? FT_H : SB_H;
  int begin = 0;
  for (int i = 0; i < step->n_pass; i++)
  {
    BlurPass &bp = step->pass[i];
    bp.tile_h = th;
    bp.tiles_x = (bp.w + tw - 1) / tw;
    bp.tiles_y = (bp.h + th - 1) / th;
    bp.tile_begin = begin;
    begin += bp.tiles_x * bp.tiles_y;
  }
  step->n_tiles = begin;
}

static cudaError_t launch_blur_step_fast(const BlurStep &step, cudaStream_t st)
{
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev])
  {
    cudaError_t e =
        cudaFuncSetAttribute(blur_step_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(float) * ft_smem_floats(12, FT_H));
    if (e != cudaSuccess)
      return e;
    attr_done[dev] = true;
  }
  BlurStepFast F;
  F.s = step;
  size_t smem = 0;
  for (int i = 0; i < step.n_pass; i++)
  {
    const int re = (step.pass[i].radius + 1) & ~1;
    const size_t need = sizeof(float) * (size_t)ft_smem_floats(re < 2 ? 2 : re, FT_H);
    smem = need > smem ? need : smem;
    for (int k = 0; k < 14; k++)
    {
      const float v = (k <= step.pass[i].radius) ? step.pass[i].taps[k] : 0.f;
      F.taps2[i][k] = make_float2(v, v);
    }
  }
  blur_step_fast_kernel<<<step.n_tiles, FT_THREADS, smem, st>>>(F);
  return cudaGetLastError();
}

cudaError_t launch_blur_step(const BlurStep &step, cudaStream_t st)
{
  if (step.n_tiles <= 0)
    return cudaSuccess;
  if (blur_step_is_fast(step))
    return launch_blur_step_fast(step, st);
  blur_step_small_kernel<<<step.n_tiles, 256, 0, st>>>(step);
  return cudaGetLastError();
}

} // namespace vks
