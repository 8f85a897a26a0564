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
#include "tma_util.cuh"
#include "blur_arith.cuh"
#include "layer_io.cuh"

#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>
#include <vector>

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
    return layer_ld(p.src, (size_t)y * p.src_pitch + x, p.fp16);
  if (p.src_kind == BLUR_SRC_U8_UP2)
    return fetch_u8_up2(*(const uint8_t *const *)p.src, p.src_w, p.src_h, x, y);
  return vks_unorm8((*(const uint8_t *const *)p.src)[(size_t)y * p.src_w + x]);
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
    const void *__restrict__ src = p.src;
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
          v[k] = layer_ld(src, (size_t)gy * p.src_pitch + gx, p.fp16);
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
    if (p.fp16)
      acc = __half2float(__float2half_rn(acc));
    layer_st(p.dst_g, (size_t)y * p.dst_pitch + x, acc, p.fp16);
    if (p.dst_d)
      layer_st(p.dst_d, (size_t)y * p.dst_pitch + x, vks_sub(acc, s_in[(yy + R) * in_w + xx + R]), p.fp16);
    if (p.dst_next && (x & 1) && (y & 1))
    {
      const int nx = x >> 1, ny = y >> 1;
      if (nx < p.next_w && ny < p.next_h)
        layer_st(p.dst_next, (size_t)ny * p.next_pitch + nx, acc, p.fp16);
    }
  }
}

/* ==========================================================================
 * Fast tile kernel: radius <= 12 (every scale of the default configuration).
 *
 * Measured on B200 (tools/ubench/fp32_rate.cu): FFMA2/FADD2 retire the same number of fp32 results per
 * cycle as FFMA/FADD (128 per SM), they only halve the issue slots.  A blur pass needs 2(2R+1) fp32
 * operations per pixel; the kernel is bound by issue slots, the FMA pipe and, above all, by how many of its
 * CTAs an SM holds, not by HBM:
 *   - 64x64 output tile, 256 threads, five (R = 4) or four CTAs per SM: a tile runs in phases (TMA wait,
 *     horizontal pass, barrier, vertical pass) and only the other CTAs of the SM fill the gaps
 *   - the source tile + halo arrives by TMA (cp.async.bulk.tensor.2d, zero thread instructions) in four
 *     row bands with one mbarrier each, so the horizontal pass starts when the first band has landed;
 *     tiles on the image border get their MIRRORED_REPEAT halo patched in shared memory
 *   - horizontal pass: one thread = 1 row x 16 columns, window in registers straight from LDS.128.
 *     Output pairs (x, x+1): even taps use the aligned register pairs of the window (FADD2), odd taps add
 *     the two scalars into a fresh pair (2 FADD), both feed one FFMA2 -- no register shuffles
 *   - vertical pass: one thread = 2 columns x 8 rows, pairs (x, x+1), sliding window of LDS.64;
 *     writes G, DoG = G - centre (from the source tile) and the decimated seed with predicated 64-bit stores
 *   - one launch = one layer, one kernel per (radius, kind): every parameter and tap is a constant operand
 *   - everything that runs rarely (border patch, partial tile columns) is kept in rolled loops: the kernels of
 *     five to ten launches share an SM's instruction cache, and straight-line cold code cost 1.4 % of the image time
 * Per-pixel operation sequence is the one of vksift_arith.h; add/fma.rn.f32x2 are two IEEE operations.
 * ========================================================================== */
#define FT_W 64
#define FT_H 128 /* only the unit of the "large enough" test in blur_pass_is_fast; tiles are ft_tile_h(R) = 64 rows high */
#define FT_THREADS 256
#define FT_MS (FT_W + 4) /* row stride of the horizontal-pass result, floats (stride/4 odd) */
#define FT_NB 4          /* TMA row bands per tile */

enum
{
  FT_KIND_SEED = 0, /* u8 source (octave 0, layer 0): no DoG */
  FT_KIND_LAYER = 1, /* float source, G + DoG */
  FT_KIND_NEXT = 2,  /* float source, G + DoG + decimated seed of the next octave (layer ns) */
  FT_KIND_FIRST = 3  /* float source (the expanded input image, expand_input_kernel), G only: layer 0 of octave 0 */
};
__host__ __device__ constexpr bool ft_has_dog(int KIND) { return KIND == FT_KIND_LAYER || KIND == FT_KIND_NEXT; }

__host__ __device__ constexpr int ft_rx(int R) { return (R + 3) & ~3; } /* x halo rounded to float4 */
/* source tile row stride (= TMA box width), floats: covers 64 + 2 halos, stride/4 odd -> LDS.128 down a column is conflict free */
__host__ __device__ constexpr int ft_s(int R) { return (((FT_W + 2 * ft_rx(R)) / 4) | 1) * 4; }
__host__ __device__ constexpr int ft_in_h(int R, int TH = FT_H) { return TH + 2 * R; }
#define FT_BOX_H 8 /* rows per TMA request: 8 rows of a stride that is a multiple of 4 floats keep every destination 128-byte aligned */
__host__ __device__ constexpr int ft_n_box(int R, int TH = FT_H) { return (ft_in_h(R, TH) + FT_BOX_H - 1) / FT_BOX_H; }
__host__ __device__ constexpr int ft_in_ha(int R, int TH = FT_H) { return ft_n_box(R, TH) * FT_BOX_H; } /* rows allocated for the source tile */
/* Tile height and residency of the per-layer launches.  A tile runs in phases (TMA wait, horizontal pass, barrier, vertical
 * pass) and only other CTAs of the SM fill the gaps, so residency decides (measured with 8 detections in flight, ms per
 * 1920x1080 image): 64-row tiles with 4 / 4 / 3 CTAs per SM for R = 4 / 6-8 / 10-12: 0.3151;  96-row tiles for R >= 10 (fewer
 * halo rows, balanced horizontal pass, still 3 CTAs): 0.3125;  five CTAs for R = 4 (48 registers, 41 KB): 0.3080;  64-row tiles
 * and FOUR CTAs for R >= 10 (64 registers without spills, 55-57 KB): 0.3046 -- occupancy beats the 12 % of instructions the
 * taller tiles save. */
__host__ __device__ constexpr int ft_tile_h(int R) { return R >= 0 ? 64 : 64; }
__host__ __device__ constexpr int ft_ctas_per_sm(int R) { return R <= 4 ? 5 : 4; }
__host__ __device__ constexpr int ft_bar_off(int R, int TH) { return ft_in_ha(R, TH) * ft_s(R) + ft_in_h(R, TH) * FT_MS; }
/* the 1 KB UNORM table only exists in the seed pass: without it four CTAs of the R = 12 layer kernel fit an SM (57 344 bytes each) */
__host__ __device__ constexpr int ft_smem_bytes(int R, int TH, int KIND) { return 4 * ft_bar_off(R, TH) + 8 * FT_NB + (KIND == 0 ? 1024 : 0); }
/* source tile, horizontal-pass result, TMA barriers, UNORM table of the seed pass */
/* one shared-memory layout for every radius (a persistent CTA runs tiles of several layers): source tile and
 * horizontal-pass result sized by R inside the first FT_BAR_OFF floats, then the TMA barriers and the UNORM table */
#define FT_BAR_OFF (ft_in_ha(12) * ft_s(12) + ft_in_h(12) * FT_MS)
#define FT_SMEM_BYTES (4 * FT_BAR_OFF + 8 * FT_NB + 1024)
__host__ __device__ constexpr int ft_even(int r) { return r < 2 ? 2 : ((r + 1) & ~1); }

struct BlurPassFast
{
  BlurPass p;
  float2 taps2[14]; /* (k,k) pairs, zero padded to the even radius */
};

#define FT_TAP(i) (*reinterpret_cast<const pk2 *>(&taps2[i]))

/* One 64 x TH tile of one layer.  The TMA barriers at the end of the shared memory block are initialised by the
 * caller; `parity` is the phase they complete next (a persistent caller flips it for every float-source tile).
 * `taps2` must point into the kernel parameters at a compile-time offset (constant operands). */
template <int R, int KIND, int TH, int BAR_OFF, bool H16>
__device__ __forceinline__ void blur_tile(const BlurPass &p, const CUtensorMap *tmap, const float2 *__restrict__ taps2, float *ft_smem, int x0, int y0,
                                          uint32_t parity, bool pdl)
{
  constexpr int VR = TH / (FT_THREADS / 32); /* output rows per warp in the vertical pass */
  static_assert(VR % 4 == 0, "vertical pass works in blocks of 4 rows");
  constexpr int RX = ft_rx(R);
  constexpr int S = ft_s(R);
  constexpr int IN_H = ft_in_h(R, TH);
  constexpr int IN_HA = ft_in_ha(R, TH);
  constexpr int NBOX = ft_n_box(R, TH); /* TMA requests per tile; request j signals barrier j*FT_NB/NBOX */
  float *s_in = ft_smem;
  float *s_mid = ft_smem + IN_HA * S;
  const uint32_t bar0 = tma_smem_u32(ft_smem + BAR_OFF);
  const int tid = threadIdx.x;
  const int wi = tid >> 5, lane = tid & 31;

  /* rows of this tile that can influence a pixel inside the image */
  const int rows_valid = min(TH, p.h - y0); /* output rows */
  const int rows_in = rows_valid + 2 * R;     /* source rows the vertical pass will read */
  int bands_seen = FT_NB;                      /* TMA bands this thread has already waited for */

  /* ---- stage 1: source tile -> smem ---- */
  if (KIND != FT_KIND_SEED && H16)
  {
    /* binary16 layers: the tile is fetched with vectorised half2 loads (the pair of a column pair is 4-byte aligned: tile
     * origin and pitch are even) and widened to the fp32 tile the passes work on; MIRRORED_REPEAT resolved on the way */
    if (pdl)
      pdl_wait();
    const __half *__restrict__ src = reinterpret_cast<const __half *>(p.src);
    const bool once = (x0 - RX >= -p.w) && (x0 - RX + S <= 2 * p.w) && (y0 - R >= -p.h) && (y0 + TH + R <= 2 * p.h);
    /* one warp per tile row, a lane per group of four columns: one 8-byte load (four halves; tile origin, pitch and group are
     * multiples of four elements), groups that touch the image border go element by element through the mirror */
    for (int r = wi; r < rows_in; r += FT_THREADS / 32)
    {
      const int gy = once ? mirror_once(y0 - R + r, p.h) : vks_mirror(y0 - R + r, p.h);
      const __half *row = src + (size_t)gy * p.src_pitch;
      for (int c = 4 * lane; c < S; c += 128)
      {
        const int gx = x0 - RX + c;
        float4 v;
        if (gx >= 0 && gx + 3 < p.w)
        {
          const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(row + gx));
          const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x)), b = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
          v = make_float4(a.x, a.y, b.x, b.y);
        }
        else
        {
          float e[4];
#pragma unroll
          for (int k = 0; k < 4; k++)
            e[k] = __half2float(__ldg(row + (once ? mirror_once(gx + k, p.w) : vks_mirror(gx + k, p.w))));
          v = make_float4(e[0], e[1], e[2], e[3]);
        }
        *reinterpret_cast<float4 *>(s_in + r * S + c) = v;
      }
    }
    __syncthreads();
  }
  else if (KIND != FT_KIND_SEED)
  {
    if (tid == 0)
    {
      if (pdl)
        pdl_wait(); /* the source layer is written by the previous launch of the stream */
      tma_fence_proxy_async(); /* earlier generic-proxy accesses to the tile buffer are ordered before the TMA writes */
#pragma unroll
      for (int b = 0; b < FT_NB; b++)
      {
        constexpr int per = FT_BOX_H * S * 4;
        const int n_req = ((b + 1) * NBOX + FT_NB - 1) / FT_NB - (b * NBOX + FT_NB - 1) / FT_NB; /* requests j with j*FT_NB/NBOX == b */
        tma_mbar_expect_tx(bar0 + 8 * b, (uint32_t)(n_req * per));
      }
#pragma unroll
      for (int j = 0; j < NBOX; j++)
        tma_load_2d_f32(tma_smem_u32(s_in + j * FT_BOX_H * S), tmap, x0 - RX, y0 - R + j * FT_BOX_H, bar0 + 8 * (j * FT_NB / NBOX));
    }
    bands_seen = 0;
    const bool border = (x0 - RX < 0) || (x0 - RX + S > p.w) || (y0 - R < 0) || (y0 - R + IN_H > p.h);
    if (border)
    {
      /* MIRRORED_REPEAT: cells outside the image (zero filled by TMA) take the value of their mirror
       * cell, which lies inside the image and inside this tile for every cell an in-image output reads.
       * Columns first (rows that exist in the image), then whole rows. */
#pragma unroll
      for (int b = 0; b < FT_NB; b++)
        tma_mbar_wait(bar0 + 8 * b, parity);
      bands_seen = FT_NB;
      const int cx = x0 - RX, cy = y0 - R;          /* image coordinates of cell (0,0) */
      const int nl = max(0, -cx);                   /* columns left of the image */
      const int cr = min(S, max(0, p.w - cx));      /* first column right of the image */
      const int ncol = nl + (S - cr);
      const int r_lo = max(0, -cy), r_hi = min(rows_in, p.h - cy); /* rows inside the image */
      if (ncol > 0)
      {
#pragma unroll 1
        for (int i = tid; i < (r_hi - r_lo) * ncol; i += FT_THREADS)
        {
          const int rr = i / ncol, k = i - rr * ncol;
          const int r = r_lo + rr;
          const int c = k < nl ? k : cr + (k - nl);
          const int mx = min(max(mirror_once(cx + c, p.w) - cx, 0), S - 1);
          s_in[r * S + c] = s_in[r * S + mx];
        }
        __syncthreads();
      }
      const int nrow = r_lo + (rows_in - r_hi);
      if (nrow > 0)
      {
#pragma unroll 1
        for (int i = tid; i < nrow * (S / 4); i += FT_THREADS)
        {
          const int k = i / (S / 4), c4 = i - k * (S / 4);
          const int r = k < r_lo ? k : r_hi + (k - r_lo);
          const int my = min(max(mirror_once(cy + r, p.h) - cy, 0), IN_H - 1);
          ((float4 *)(s_in + r * S))[c4] = ((const float4 *)(s_in + my * S))[c4];
        }
      }
      __syncthreads();
    }
  }
  else
  {
    if (pdl)
      pdl_wait();
    /* octave 0 seed.  s_mid serves as scratch: the u8 source window as float (one UNORM division per
     * source pixel, MIRRORED_REPEAT / clamp-to-edge resolved here), then the LINEAR 2x blit (or the 1:1
     * copy) out of shared memory.  Destination cell (m, c) <-> image pixel (x0-RX+c, y0-R+m). */
    const bool up = (p.src_kind == BLUR_SRC_U8_UP2);
    const bool once = (x0 - RX >= -p.w) && (x0 - RX + S <= 2 * p.w) && (y0 - R >= -p.h) && (y0 + TH + R <= 2 * p.h);
    const uint8_t *__restrict__ img = *(const uint8_t *const *)p.src; /* u8 sources are reached through a pointer slot (see BlurPass) */
    const int n_el = S * rows_in;
    /* mirrored destination coordinates stay inside [lo, hi] of the image */
    const int dx_lo = max(0, min(x0 - RX, p.w - 1)), dx_hi = min(p.w - 1, max(0, x0 - RX + S - 1));
    const int dy_lo = max(0, min(y0 - R, p.h - 1)), dy_hi = min(p.h - 1, max(0, y0 + rows_in - R - 1));
    const bool full = !once; /* tiny images reflect more than once: stage the whole source */
    const int sx_lo = full ? 0 : (up ? max(0, (dx_lo >> 1) - 1) : dx_lo);
    const int sx_hi = full ? p.src_w - 1 : (up ? min(p.src_w - 1, (dx_hi >> 1) + 1) : dx_hi);
    const int sy_lo = full ? 0 : (up ? max(0, (dy_lo >> 1) - 1) : dy_lo);
    const int sy_hi = full ? p.src_h - 1 : (up ? min(p.src_h - 1, (dy_hi >> 1) + 1) : dy_hi);
    const int sw = sx_hi - sx_lo + 1, sh = sy_hi - sy_lo + 1;
    if (sw * sh <= IN_H * FT_MS && once)
    {
      /* UNORM conversion through a 256-entry table (one IEEE division per CTA thread instead of one per pixel) */
      float *lut = ft_smem + BAR_OFF + 2 * FT_NB;
      lut[tid] = vks_unorm8((uint8_t)tid);
      __syncthreads();
      /* one warp per staged row, lanes along the row: no integer division in the loop */
      for (int yy = wi; yy < sh; yy += FT_THREADS / 32)
      {
        const uint8_t *__restrict__ srow = img + (size_t)(sy_lo + yy) * p.src_w + sx_lo;
        for (int xx = lane; xx < sw; xx += 32)
          s_mid[yy * sw + xx] = lut[__ldg(srow + xx)];
      }
      __syncthreads();
      const int cx = x0 - RX, cy = y0 - R; /* both even */
      const bool inner = up && cx >= 2 && cx + S <= p.w - 2 && cy >= 2 && cy + rows_in <= p.h - 2 && (rows_in & 1) == 0;
      if (inner)
      {
        /* Tile away from the image border: no mirroring, no clamping, sx_lo = cx/2 - 1, sy_lo = cy/2 - 1.
         * With H[t][c] = the horizontal lerp of staged row t at destination column c, destination row
         * 2j = lerp(H[j], H[j+1], .75) and row 2j+1 = lerp(H[j+1], H[j+2], .25).  One thread owns a column
         * pair (even c: staged columns (c/2, c/2+1), f=.75; odd c+1: (c/2+1, c/2+2), f=.25) and walks down. */
        constexpr int NCP = S / 2;
        constexpr int NG = FT_THREADS / NCP;
        const int cp = tid % NCP, grp = tid / NCP;
        const int n_rp = rows_in >> 1;
        const int per = (n_rp + NG - 1) / NG;
        const int j0 = grp * per, j1 = min(n_rp, j0 + per);
        if (grp < NG && j0 < j1)
        {
          const float *sp = s_mid + j0 * sw + cp;
          float a0 = sp[0], a1 = sp[1], a2 = sp[2];
          float he0 = vks_lerp(a0, a1, 0.75f), ho0 = vks_lerp(a1, a2, 0.25f);
          sp += sw;
          a0 = sp[0], a1 = sp[1], a2 = sp[2];
          float he1 = vks_lerp(a0, a1, 0.75f), ho1 = vks_lerp(a1, a2, 0.25f);
          float *dp = s_in + (2 * j0) * S + 2 * cp;
          for (int j = j0; j < j1; j++)
          {
            sp += sw;
            a0 = sp[0], a1 = sp[1], a2 = sp[2];
            const float he2 = vks_lerp(a0, a1, 0.75f), ho2 = vks_lerp(a1, a2, 0.25f);
            *(float2 *)dp = make_float2(vks_lerp(he0, he1, 0.75f), vks_lerp(ho0, ho1, 0.75f));
            *(float2 *)(dp + S) = make_float2(vks_lerp(he1, he2, 0.25f), vks_lerp(ho1, ho2, 0.25f));
            dp += 2 * S;
            he0 = he1, ho0 = ho1, he1 = he2, ho1 = ho2;
          }
        }
      }
      else
      if (up)
      {
        /* one thread = one destination column pair (even x, odd x): their source columns are
         * (k-1, k) with f=.75 and (k, k+1) with f=.25; walk down the rows */
        for (int i = tid; i < (S / 2) * rows_in; i += FT_THREADS)
        {
          const int m = i / (S / 2), cp = i - m * (S / 2);
          const int gy = mirror_once(y0 - R + m, p.h);
          const int ky = gy >> 1;
          const int ay = min(max(max(((gy & 1) ? ky : ky - 1), 0) - sy_lo, 0), sh - 1);
          const int by = min(max(min(((gy & 1) ? ky + 1 : ky), p.src_h - 1) - sy_lo, 0), sh - 1);
          const float fy = (gy & 1) ? 0.25f : 0.75f;
          float v[2];
#pragma unroll
          for (int e = 0; e < 2; e++)
          {
            const int gx = mirror_once(x0 - RX + 2 * cp + e, p.w);
            const int kx = gx >> 1;
            /* the clamps to the staged window only ever bind for cells that no in-image output reads */
            const int ax = min(max(max(((gx & 1) ? kx : kx - 1), 0) - sx_lo, 0), sw - 1);
            const int bx = min(max(min(((gx & 1) ? kx + 1 : kx), p.src_w - 1) - sx_lo, 0), sw - 1);
            const float fx = (gx & 1) ? 0.25f : 0.75f;
            const float top = vks_lerp(s_mid[ay * sw + ax], s_mid[ay * sw + bx], fx);
            const float bot = vks_lerp(s_mid[by * sw + ax], s_mid[by * sw + bx], fx);
            v[e] = vks_lerp(top, bot, fy);
          }
          *(float2 *)(s_in + m * S + 2 * cp) = make_float2(v[0], v[1]);
        }
      }
      else
      {
        for (int i = tid; i < n_el; i += FT_THREADS)
        {
          const int m = i / S, c = i - m * S;
          const int gx = mirror_once(x0 - RX + c, p.w), gy = mirror_once(y0 - R + m, p.h);
          s_in[i] = s_mid[min(max(gy - sy_lo, 0), sh - 1) * sw + min(max(gx - sx_lo, 0), sw - 1)];
        }
      }
    }
    else
    {
      for (int i = tid; i < n_el; i += FT_THREADS)
      {
        const int m = i / S, c = i - m * S;
        s_in[i] = fetch_src(p, x0 - RX + c, y0 - R + m);
      }
    }
    __syncthreads();
  }

  /* ---- stage 2: horizontal pass, warp unit = (8 rows, 64 columns) = the rows of one TMA request;
   *      lane = (row, 16-column group): 8 consecutive lanes read 8 rows, conflict free because stride/4 is odd ---- */
  {
    constexpr int WN = 16 + 2 * RX; /* window floats, float4 aligned */
    const int n_units = (rows_in + 7) >> 3;
    for (int wu = wi; wu < n_units; wu += FT_THREADS / 32)
    {
      const int g = lane >> 3;
      const int r = min(wu * 8 + (lane & 7), IN_H - 1);
      if (KIND != FT_KIND_SEED)
      {
        const int need = min(wu, NBOX - 1) * FT_NB / NBOX;
        while (bands_seen <= need)
        {
          tma_mbar_wait(bar0 + 8 * bands_seen, parity);
          bands_seen++;
        }
      }
      const ulonglong2 *wsrc = (const ulonglong2 *)(s_in + r * S + g * 16);
      pk2 wp[WN / 2]; /* wp[j] = (in[2j], in[2j+1]) relative to column g*16 - RX */
#pragma unroll
      for (int j = 0; j < WN / 4; j++)
      {
        const ulonglong2 v = wsrc[j];
        wp[2 * j] = v.x;
        wp[2 * j + 1] = v.y;
      }
      ulonglong2 *dst = (ulonglong2 *)(s_mid + r * FT_MS + g * 16);
#pragma unroll
      for (int hb = 0; hb < 2; hb++)
      {
        /* four output pairs at a time: four independent accumulator chains */
        pk2 acc[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
          acc[j] = pk_mul(wp[RX / 2 + 4 * hb + j], FT_TAP(0));
#pragma unroll
        for (int i = 1; i <= R; i++)
        {
#pragma unroll
          for (int j = 0; j < 4; j++)
          {
            const int c = RX / 2 + 4 * hb + j; /* pair index of the centre */
            pk2 sum;
            if ((i & 1) == 0)
              sum = pk_add(wp[c + i / 2], wp[c - i / 2]);
            else
            {
              /* (in[x+i], in[x+1+i]) straddles two register pairs: add the scalars into a fresh pair */
              const float s0 = __fadd_rn(pk_hi(wp[c + (i - 1) / 2]), pk_hi(wp[c - (i + 1) / 2]));
              const float s1 = __fadd_rn(pk_lo(wp[c + (i + 1) / 2]), pk_lo(wp[c - (i - 1) / 2]));
              sum = pk_make(s0, s1);
            }
            acc[j] = pk_fma(sum, FT_TAP(i), acc[j]);
          }
        }
        dst[2 * hb] = make_ulonglong2(acc[0], acc[1]);
        dst[2 * hb + 1] = make_ulonglong2(acc[2], acc[3]);
      }
    }
  }
  __syncthreads();

  /* ---- stage 3: vertical pass, unit = (column pair, 16 rows), lanes along column pairs ---- */
  const int ry = wi * VR; /* first output row of this warp inside the tile */
  const int nrows = min(VR, rows_valid - ry);
  const int yb = y0 + ry; /* even */
  if (nrows <= 0)
  {
  }
  else if (x0 + FT_W <= p.w)
  {
    /* every lane owns a full pixel pair */
    const int x = x0 + 2 * lane;
    const float *mcol = s_mid + ry * FT_MS + 2 * lane;
    const float *ccol = s_in + (R + ry) * S + RX + 2 * lane;
    /* G and DoG share pitch and offset: one 32-bit byte offset walks down both layers */
    char *const gbase = (char *)p.dst_g;
    char *const dbase = (char *)p.dst_d;
    char *const nbase = (char *)p.dst_next;
    uint32_t off = ((uint32_t)yb * (uint32_t)p.dst_pitch + (uint32_t)x) * 4u;
    uint32_t noff = ((uint32_t)(yb >> 1) * (uint32_t)p.next_pitch + (uint32_t)(x >> 1)) * 4u;
    const uint32_t pitch4 = (uint32_t)p.dst_pitch * 4u, npitch4 = (uint32_t)p.next_pitch * 4u;
    pk2 wv[VR + 2 * R];
#pragma unroll
    for (int j = 0; j < 2 * R; j++)
      wv[j] = *(const pk2 *)(mcol + j * FT_MS);
#pragma unroll
    for (int qb = 0; qb < VR; qb += 4)
    {
#pragma unroll
      for (int j = 0; j < 4; j++)
        wv[2 * R + qb + j] = *(const pk2 *)(mcol + (2 * R + qb + j) * FT_MS);
      pk2 acc[4];
#pragma unroll
      for (int j = 0; j < 4; j++)
        acc[j] = pk_mul(wv[R + qb + j], FT_TAP(0));
#pragma unroll
      for (int i = 1; i <= R; i++)
      {
#pragma unroll
        for (int j = 0; j < 4; j++)
          acc[j] = pk_fma(pk_add(wv[R + qb + j + i], wv[R + qb + j - i]), FT_TAP(i), acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; j++)
      {
        const int q = qb + j;
        if (H16)
        {
          if (q < nrows)
          {
            /* binary16 layers: one rounding to nearest even, the rounded value is what the DoG is taken from (it is what
             * the layer holds), half2 stores */
            const __half2 hg = __floats2half2_rn(pk_lo(acc[j]), pk_hi(acc[j]));
            *reinterpret_cast<__half2 *>(gbase + (off >> 1)) = hg;
            const float2 fg = __half22float2(hg);
            if (ft_has_dog(KIND))
            {
              const pk2 d = pk_sub(pk_make(fg.x, fg.y), *(const pk2 *)(ccol + q * S));
              __stcs(reinterpret_cast<__half2 *>(dbase + (off >> 1)), __floats2half2_rn(pk_lo(d), pk_hi(d)));
            }
            if (KIND == FT_KIND_NEXT && (q & 1))
            {
              *reinterpret_cast<__half *>(nbase + (noff >> 1)) = __high2half(hg);
              noff += npitch4;
            }
          }
        }
        else
        {
          /* stores as predicated instructions (rows past the image end in the last tile row); the operands are computed
           * for every row: the DoG source row is in shared memory whether the output row exists or not */
          const bool ok = q < nrows;
          pk_stg_if(gbase + off, acc[j], ok);
          if (ft_has_dog(KIND))
          {
            /* streaming: the DoG layer is not read before the extrema scan, the L2 lines are better spent on G, which the
             * next layer's launch reads back */
            pk_stcs_if(dbase + off, pk_sub(acc[j], *(const pk2 *)(ccol + q * S)), ok);
          }
          if (KIND == FT_KIND_NEXT && (q & 1))
          {
            /* x even, y odd: the odd column of the pair feeds next(x>>1, y>>1) */
            f32_stg_if(nbase + noff, pk_hi(acc[j]), ok);
            noff += npitch4;
          }
        }
        off += pitch4;
      }
    }
  }
  else
  {
    /* last tile column of a layer whose width is not a multiple of 64: plain scalar loop, kept rolled (it runs in one tile
     * column of some layers and must not take instruction-cache space from the path above) */
#pragma unroll 1
    for (int idx = lane; idx < FT_W * nrows; idx += 32)
    {
      const int col = idx & (FT_W - 1), q = idx >> 6;
      const int x = x0 + col, y = yb + q;
      if (x >= p.w)
        continue;
      const float *mc = s_mid + (ry + q + R) * FT_MS + col;
      float acc = vks_mul(mc[0], taps2[0].x);
#pragma unroll 1
      for (int i = 1; i <= R; i++)
        acc = vks_blur_tap(acc, mc[i * FT_MS], mc[-i * FT_MS], taps2[i].x);
      if (H16)
        acc = round_half1(acc);
      layer_st(p.dst_g, (size_t)y * p.dst_pitch + x, acc, H16);
      if (ft_has_dog(KIND))
        layer_st(p.dst_d, (size_t)y * p.dst_pitch + x, vks_sub(acc, s_in[(R + ry + q) * S + RX + col]), H16);
      if (KIND == FT_KIND_NEXT && (x & 1) && (y & 1))
      {
        const int nx = x >> 1, ny = y >> 1;
        if (nx < p.next_w && ny < p.next_h)
          layer_st(p.dst_next, (size_t)ny * p.next_pitch + nx, acc, H16);
      }
    }
  }
}

/* one launch = one pass (one Gaussian layer): every parameter sits at a fixed constant-bank address, the
 * taps are constant operands of the FFMA2s.  Passes that are independent (different octaves) run
 * concurrently from different streams; the block scheduler fills the tail of one with the head of another. */
template <int R, int KIND, bool H16>
__global__ void __launch_bounds__(FT_THREADS, ft_ctas_per_sm(R)) blur_pass_fast_kernel(const __grid_constant__ BlurPassFast P)
{
  extern __shared__ __align__(128) float ft_smem[];
  constexpr int TH = ft_tile_h(R);
  constexpr int BAR_OFF = ft_bar_off(R, TH);
  pdl_launch_dependents();
  if (threadIdx.x == 0)
  {
    const uint32_t bar0 = tma_smem_u32(ft_smem + BAR_OFF);
#pragma unroll
    for (int b = 0; b < FT_NB; b++)
      tma_mbar_init(bar0 + 8 * b, 1);
    tma_mbar_fence_init();
  }
  __syncthreads(); /* barriers initialised before anybody polls them */
  /* grid = (tiles_x, tiles_y): no division on the way to the tile origin */
  const int x0 = (int)blockIdx.x * FT_W;
  const int y0 = (int)blockIdx.y * TH;
  blur_tile<R, KIND, TH, BAR_OFF, H16>(P.p, &P.p.tmap, P.taps2, ft_smem, x0, y0, 0u, true);
}

/* ==========================================================================
 * Fused kernel for the small octaves.
 *
 * Below ~1000x600 a layer is a handful of tiles and the pyramid becomes a chain of dependent launches
 * (layer s needs layer s-1, the next octave needs layer ns): latency, not throughput.  Two measured facts
 * shape this kernel (B200, profiles/): a launch on a cold SM pays roughly 0.5-1 us per KB of straight-line
 * code it runs once, so the unrolled kernels above need 10+ us for a one-tile grid; and the per-launch
 * dependency gap is ~2.5 us.  So: ONE launch produces up to FZ_MAXL consecutive layers of an octave, and
 * its code is a few hundred bytes of rolled loops that stay in the instruction cache.
 * A CTA owns a 32x32 output tile, loads it with a halo of the summed radii and recomputes the halo of the
 * intermediate layers itself (the small octaves hold ~6 % of the pyramid's pixels; small tiles keep the
 * per-CTA critical path short and spread an octave over many SMs).  Recomputing a layer outside the
 * image on the MIRRORED_REPEAT extension of its source yields bit for bit the value of the mirror pixel
 * (the tap pairs a+b only swap their operands), so borders need no special handling beyond the mirrored
 * load.  One thread = one pixel, lanes along x in both passes (conflict-free LDS.32); the loops are bound
 * by the shared-memory pipe, which is fine at this size.
 * ========================================================================== */
#define FZ_T 32       /* output tile edge */
#define FZ_MAXHALO 24 /* summed radii of a launch */
#define FZ_HB (FZ_T + 2 * FZ_MAXHALO)
#define FZ_WB (FZ_HB + 1)
#define FZ_THREADS 1024
#define FZ_BUF (FZ_HB * FZ_WB)
#define FZ_SMEM (3 * FZ_BUF * 4)

__global__ void __launch_bounds__(FZ_THREADS, 1) octave_fused_kernel(const __grid_constant__ FusedLaunch P)
{
  extern __shared__ __align__(16) float fz_smem[];
  __shared__ float s_taps[FZ_MAXL][16];
  float *cur = fz_smem, *mid = fz_smem + FZ_BUF, *nxt = fz_smem + 2 * FZ_BUF;
  const int tid = threadIdx.x;
  const int wi = tid >> 5, lane = tid & 31;
  const int t = (int)blockIdx.x;
  const int x0 = (t % P.tiles_x) * FZ_T, y0 = (t / P.tiles_x) * FZ_T;
#ifdef VKS_FUSED_TIMING
  unsigned long long tm[12];
  int ntm = 0;
#define FZ_MARK()                                                                                                                                    \
  {                                                                                                                                                  \
    unsigned long long t__;                                                                                                                          \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                                                                                        \
    tm[ntm++] = t__;                                                                                                                                 \
  }
#else
#define FZ_MARK()
#endif
  FZ_MARK();
  int halo = 0;
  for (int k = 0; k < P.n_layers; k++)
    halo += P.radius[k];
  if (tid < FZ_MAXL * 16)
    s_taps[tid >> 4][tid & 15] = (tid & 15) < 14 ? P.taps2[tid >> 4][tid & 15].x : 0.f;
  pdl_launch_dependents();
  pdl_wait(); /* the source layer is written by the previous launch of the stream */
  /* buffer cell (r, c) <-> image pixel (x0 - halo + c, y0 - halo + r) */
  {
    const int n = FZ_T + 2 * halo;
    /* a halo of at most 24 pixels reflects once unless the octave itself is tiny */
    const bool once = (FZ_MAXHALO <= P.w) && (FZ_MAXHALO <= P.h) && (x0 + FZ_T + FZ_MAXHALO <= 2 * P.w) && (y0 + FZ_T + FZ_MAXHALO <= 2 * P.h);
    int gxs[3]; /* n <= 80: at most three columns per lane */
#pragma unroll
    for (int j = 0; j < 3; j++)
      gxs[j] = once ? mirror_once(x0 - halo + lane + 32 * j, P.w) : vks_mirror(x0 - halo + lane + 32 * j, P.w);
    for (int rr = wi; rr < n; rr += FZ_THREADS / 32)
    {
      const int gy = once ? mirror_once(y0 - halo + rr, P.h) : vks_mirror(y0 - halo + rr, P.h);
      const size_t row = (size_t)gy * P.pitch;
      float *dst = cur + rr * FZ_WB + lane;
#pragma unroll
      for (int j = 0; j < 3; j++)
        if (lane + 32 * j < n)
          dst[32 * j] = layer_ld(P.src, row + gxs[j], P.fp16);
    }
  }
  __syncthreads();
  FZ_MARK();
  int lo = 0; /* the region of the current source is [lo, lo+n)^2 in buffer cells */
#pragma unroll 1
  for (int k = 0; k < P.n_layers; k++)
  {
    const int R = P.radius[k];
    const float *tp = s_taps[k];
    const int n_src = FZ_T + 2 * (halo - lo);
    const int n = n_src - 2 * R; /* region of this layer: [lo+R, lo+R+n)^2 */
    /* horizontal: rows [lo, lo+n_src), columns [lo+R, lo+R+n) */
#pragma unroll 1
    for (int r = lo + wi; r < lo + n_src; r += FZ_THREADS / 32)
#pragma unroll 1
      for (int c = lo + R + lane; c < lo + R + n; c += 32)
      {
        const float *q = cur + r * FZ_WB + c;
        float acc = vks_mul(q[0], tp[0]);
#pragma unroll 2
        for (int i = 1; i <= R; i++)
          acc = vks_blur_tap(acc, q[i], q[-i], tp[i]);
        mid[r * FZ_WB + c] = acc;
      }
    __syncthreads();
    FZ_MARK();
    /* vertical over [lo+R, lo+R+n)^2, plus the global stores of the pixels that lie in the tile core and in the image */
    {
      const size_t lofs = (size_t)k * P.layer_stride; /* elements */
      const bool is_next = (k == P.next_k);
#pragma unroll 1
      for (int r = lo + R + wi; r < lo + R + n; r += FZ_THREADS / 32)
      {
        const int gy = y0 - halo + r;
        const bool row_ok = (r >= halo) && (r < halo + FZ_T) && (gy < P.h);
#pragma unroll 1
        for (int c = lo + R + lane; c < lo + R + n; c += 32)
        {
          const float *q = mid + r * FZ_WB + c;
          float acc = vks_mul(q[0], tp[0]);
#pragma unroll 2
          for (int i = 1; i <= R; i++)
            acc = vks_blur_tap(acc, q[i * FZ_WB], q[-i * FZ_WB], tp[i]);
          if (P.fp16)
            acc = round_half1(acc);
          nxt[r * FZ_WB + c] = acc;
          const int gx = x0 - halo + c;
          if (row_ok && c >= halo && c < halo + FZ_T && gx < P.w)
          {
            const size_t o = lofs + (size_t)gy * P.pitch + gx;
            layer_st(P.g0, o, acc, P.fp16);
            layer_st(P.d0, o, vks_sub(acc, cur[r * FZ_WB + c]), P.fp16);
            if (is_next && (gx & 1) && (gy & 1))
            {
              const int nx = gx >> 1, ny = gy >> 1;
              if (nx < P.next_w && ny < P.next_h)
                layer_st(P.dst_next, (size_t)ny * P.next_pitch + nx, acc, P.fp16);
            }
          }
        }
      }
    }
    __syncthreads();
    FZ_MARK();
    float *tmp = cur;
    cur = nxt;
    nxt = tmp;
    lo += R;
  }
#ifdef VKS_FUSED_TIMING
  if (blockIdx.x == 0 && tid == 0 && P.w < 100)
  {
    printf("fused timing (ns) w=%d layers=%d:", P.w, P.n_layers);
    for (int i = 1; i < ntm; i++)
      printf(" %llu", tm[i] - tm[0]);
    printf("\n");
  }
#endif
}

/* launch with programmatic stream serialization: the kernel may be scheduled before the previous kernel of the
 * stream has finished; it orders itself with pdl_wait() */
template <typename P>
static cudaError_t launch_pdl(void (*kernel)(P), dim3 grid, int block, size_t smem, cudaStream_t st, const P &params)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool no_pdl = [] { const char *e = getenv("VKSIFT_NO_PDL"); return e && e[0] == '1'; }();
  cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, params);
}

/* Groups the passes of one octave (consecutive layers, same octave, float sources) into fused launches: a group
 * ends after the layer that seeds the next octave, after FZ_MAXL layers or before the summed radii exceed FZ_MAXHALO.
 * Returns false when a pass cannot be fused (radius > 12): the caller falls back to the compact kernel. */
bool fused_plan_octave(const BlurPass *passes, int n_pass, std::vector<FusedLaunch> *out)
{
  int i = 0;
  while (i < n_pass)
  {
    FusedLaunch F;
    memset(&F, 0, sizeof(F));
    const BlurPass &first = passes[i];
    if (first.src_kind != BLUR_SRC_LAYER)
      return false;
    F.src = first.src;
    F.g0 = first.dst_g;
    F.d0 = first.dst_d;
    F.w = first.w;
    F.h = first.h;
    F.pitch = first.dst_pitch;
    F.next_k = -1;
    F.fp16 = first.fp16;
    F.tiles_x = (F.w + FZ_T - 1) / FZ_T;
    int halo = 0;
    while (i < n_pass && F.n_layers < FZ_MAXL)
    {
      const BlurPass &bp = passes[i];
      if (bp.radius < 1 || bp.radius > 12)
        return false;
      const int re = bp.radius;
      if (halo + re > FZ_MAXHALO)
        break;
      const int k = F.n_layers;
      if (k == 1)
        F.layer_stride = (int)(((const char *)bp.dst_g - (const char *)F.g0) / (first.fp16 ? 2 : 4));
      F.radius[k] = re;
      for (int j = 0; j < 14; j++)
      {
        const float v = (j <= bp.radius) ? bp.taps[j] : 0.f;
        F.taps2[k][j] = make_float2(v, v);
      }
      halo += re;
      F.n_layers++;
      i++;
      if (bp.dst_next)
      {
        F.next_k = k;
        F.dst_next = bp.dst_next;
        F.next_w = bp.next_w;
        F.next_h = bp.next_h;
        F.next_pitch = bp.next_pitch;
        break;
      }
    }
    if (F.n_layers == 0)
      return false;
    out->push_back(F);
  }
  return true;
}

cudaError_t launch_fused(const FusedLaunch &F, cudaStream_t st)
{
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev])
  {
    cudaError_t e = cudaFuncSetAttribute(octave_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FZ_SMEM);
    if (e != cudaSuccess)
      return e;
    attr_done[dev] = true;
  }
  const int tiles = F.tiles_x * ((F.h + FZ_T - 1) / FZ_T);
  return launch_pdl(octave_fused_kernel, dim3((unsigned)tiles), FZ_THREADS, FZ_SMEM, st, F);
}

/* A pass goes to the fast per-layer kernel when its radius is covered and the layer is large enough for
 * throughput to matter (at least 24 blocks of 64x128 pixels); smaller octaves are latency bound and take the
 * fused kernel (or, when their radii do not fit it, the compact kernel). */
bool blur_pass_is_fast(const BlurPass &bp)
{
  if (bp.radius < 1 || bp.radius > 12)
    return false;
  return ((bp.w + FT_W - 1) / FT_W) * ((bp.h + FT_H - 1) / FT_H) >= 24;
}

bool blur_step_tiles(BlurStep *step)
{
  int begin = 0;
  for (int i = 0; i < step->n_pass; i++)
  {
    BlurPass &bp = step->pass[i];
    bp.tile_h = SB_H;
    bp.tiles_x = (bp.w + SB_W - 1) / SB_W;
    bp.tiles_y = (bp.h + SB_H - 1) / SB_H;
    bp.tile_begin = begin;
    begin += bp.tiles_x * bp.tiles_y;
  }
  step->n_tiles = begin;
  return true;
}

bool blur_pass_prepare_fast(BlurPass *bpp)
{
  BlurPass &bp = *bpp;
  bp.tile_h = ft_tile_h(ft_even(bp.radius));
  bp.tiles_x = (bp.w + FT_W - 1) / FT_W;
  bp.tiles_y = (bp.h + bp.tile_h - 1) / bp.tile_h;
  bp.tile_begin = 0;
  if (bp.src_kind == BLUR_SRC_LAYER && !bp.fp16)
  {
    /* one TMA box = source tile + halo; cells outside the layer read as zero and are patched by the kernel
     * (binary16 layers are fetched with half2 loads instead) */
    const int re = ft_even(bp.radius);
    const uint64_t dims[2] = {(uint64_t)bp.w, (uint64_t)bp.h};
    const uint64_t strides[1] = {(uint64_t)bp.src_pitch * 4};
    const uint32_t box[2] = {(uint32_t)ft_s(re), (uint32_t)FT_BOX_H};
    if (!tma_make_map_f32(&bp.tmap, (const float *)bp.src, 2, dims, strides, box))
      return false;
  }
  return true;
}

template <int R, int KIND, bool H16>
static cudaError_t launch_fast_rkh(const BlurPassFast &F, dim3 n_tiles, cudaStream_t st)
{
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev])
  {
    cudaError_t e =
        cudaFuncSetAttribute(blur_pass_fast_kernel<R, KIND, H16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ft_smem_bytes(R, ft_tile_h(R), KIND));
    if (e != cudaSuccess)
      return e;
    attr_done[dev] = true;
  }
  return launch_pdl(blur_pass_fast_kernel<R, KIND, H16>, n_tiles, FT_THREADS, ft_smem_bytes(R, ft_tile_h(R), KIND), st, F);
}

template <int R, int KIND>
static cudaError_t launch_fast_rk(const BlurPassFast &F, dim3 n_tiles, cudaStream_t st)
{
  return F.p.fp16 ? launch_fast_rkh<R, KIND, true>(F, n_tiles, st) : launch_fast_rkh<R, KIND, false>(F, n_tiles, st);
}

template <int R>
static cudaError_t launch_fast_r(const BlurPassFast &F, dim3 n_tiles, cudaStream_t st)
{
  if (F.p.src_kind != BLUR_SRC_LAYER)
    return launch_fast_rk<R, FT_KIND_SEED>(F, n_tiles, st);
  if (!F.p.dst_d)
    return launch_fast_rkh<R, FT_KIND_FIRST, false>(F, n_tiles, st); /* the expanded input is fp32 (blur_plan only uses it without binary16 storage) */
  if (F.p.dst_next)
    return launch_fast_rk<R, FT_KIND_NEXT>(F, n_tiles, st);
  return launch_fast_rk<R, FT_KIND_LAYER>(F, n_tiles, st);
}

/* ---- input expansion --------------------------------------------------------
 * Layer 0 of octave 0 is the blur of the input image after the UNORM conversion and (with upsampling) the LINEAR 2x blit
 * (sift_detector.c:860-953).  The fused seed pass above does conversion + blit + blur per tile and spends two thirds of its
 * instructions on staging the u8 window and on the blit; with several detections in flight instructions are what counts, so
 * for large images the expanded fp32 image is written once by this kernel (5.5 instructions per pixel) and the first blur is
 * an ordinary TMA-fed layer launch without a DoG output (FT_KIND_FIRST).  Same operation sequence per pixel as fetch_u8_up2:
 * horizontal lerp of the two rows, then the vertical one.  One thread = one output column pair, eight output rows. */
#define XP_ROWS 32
__global__ void __launch_bounds__(256) expand_input_kernel(const uint8_t *const *__restrict__ src_slot, const int sw, const int sh, const int up,
                                                           float *__restrict__ dst, const int pitch, const int w, const int h)
{
  __shared__ float lut[256];
  lut[threadIdx.x] = vks_unorm8((uint8_t)threadIdx.x);
  pdl_launch_dependents();
  __syncthreads();
  const uint8_t *__restrict__ img = *src_slot;
  const int cp = blockIdx.x * blockDim.x + threadIdx.x; /* column pair */
  const int x = 2 * cp, y0 = (int)blockIdx.y * XP_ROWS;
  if (x >= w)
    return;
  if (!up)
  {
    for (int y = y0; y < min(h, y0 + XP_ROWS); y++)
    {
      const uint8_t *row = img + (size_t)y * sw;
      float *o = dst + (size_t)y * pitch + x;
      o[0] = lut[row[x]];
      if (x + 1 < w)
        o[1] = lut[row[x + 1]];
    }
    return;
  }
  /* destination columns (x, x+1) = (2k, 2k+1): sources (k-1, k) with f = .75 and (k, k+1) with f = .25, clamped to the image
   * (w = 2 sw is even: a column pair is always complete) */
  const int k = x >> 1;
  const int c0 = max(k - 1, 0), c1 = k, c2 = min(k + 1, sw - 1);
  auto hrow = [&](int sy, float &he, float &ho) {
    const uint8_t *row = img + (uint32_t)(min(max(sy, 0), sh - 1) * sw);
    const float a0 = lut[row[c0]], a1 = lut[row[c1]], a2 = lut[row[c2]];
    he = vks_lerp(a0, a1, 0.75f);
    ho = vks_lerp(a1, a2, 0.25f);
  };
  /* destination rows (2j, 2j+1): source rows (j-1, j) with f = .75 and (j, j+1) with f = .25 */
  const int j0 = y0 >> 1;
  const int n_pairs = min(XP_ROWS, h - y0) >> 1; /* h = 2 sh is even */
  float he0, ho0, he1, ho1, he2, ho2;
  hrow(j0 - 1, he0, ho0);
  hrow(j0, he1, ho1);
  float *o = dst + (size_t)y0 * pitch + x;
#pragma unroll 4
  for (int jj = 0; jj < n_pairs; jj++)
  {
    hrow(j0 + jj + 1, he2, ho2);
    *reinterpret_cast<float2 *>(o) = make_float2(vks_lerp(he0, he1, 0.75f), vks_lerp(ho0, ho1, 0.75f));
    *reinterpret_cast<float2 *>(o + pitch) = make_float2(vks_lerp(he1, he2, 0.25f), vks_lerp(ho1, ho2, 0.25f));
    o += 2 * pitch;
    he0 = he1, ho0 = ho1, he1 = he2, ho1 = ho2;
  }
}

cudaError_t launch_expand_input(const void *src_slot, int sw, int sh, int up, float *dst, int pitch, int w, int h, cudaStream_t st)
{
  const dim3 grid((unsigned)(((w + 1) / 2 + 255) / 256), (unsigned)((h + XP_ROWS - 1) / XP_ROWS));
  expand_input_kernel<<<grid, 256, 0, st>>>((const uint8_t *const *)src_slot, sw, sh, up, dst, pitch, w, h);
  return cudaGetLastError();
}

cudaError_t launch_blur_pass_fast(const BlurPass &bp, cudaStream_t st)
{
  BlurPassFast F;
  F.p = bp;
  for (int k = 0; k < 14; k++)
  {
    const float v = (k <= bp.radius) ? bp.taps[k] : 0.f;
    F.taps2[k] = make_float2(v, v);
  }
  const dim3 n_tiles((unsigned)bp.tiles_x, (unsigned)bp.tiles_y);
  switch (ft_even(bp.radius))
  {
  case 2:
    return launch_fast_r<2>(F, n_tiles, st);
  case 4:
    return launch_fast_r<4>(F, n_tiles, st);
  case 6:
    return launch_fast_r<6>(F, n_tiles, st);
  case 8:
    return launch_fast_r<8>(F, n_tiles, st);
  case 10:
    return launch_fast_r<10>(F, n_tiles, st);
  default:
    return launch_fast_r<12>(F, n_tiles, st);
  }
}

/* binary16 layer -> dense fp32 image (vksift_downloadScaleSpaceImage / vksift_downloadDoGImage with VKSIFT_PYRAMID_PRECISION_FLOAT16) */
__global__ void widen_layer_kernel(const __half *__restrict__ src, int w, int h, int pitch, float *__restrict__ dst)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < w && y < h)
    dst[(size_t)y * w + x] = __half2float(src[(size_t)y * pitch + x]);
}

cudaError_t launch_widen_layer(const void *src, int w, int h, int pitch, float *dst, cudaStream_t st)
{
  widen_layer_kernel<<<dim3((w + 255) / 256, h, 1), 256, 0, st>>>((const __half *)src, w, h, pitch, dst);
  return cudaGetLastError();
}

cudaError_t launch_blur_step(const BlurStep &step, cudaStream_t st)
{
  if (step.n_tiles <= 0)
    return cudaSuccess;
  blur_step_small_kernel<<<step.n_tiles, 256, 0, st>>>(step);
  return cudaGetLastError();
}

} // namespace vks
