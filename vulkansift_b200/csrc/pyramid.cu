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

/* ---- baseline tile kernel ------------------------------------------------
 * 64x32 output tile per CTA, 256 threads.  Stage 1 loads the tile plus a halo
 * of `radius` on every side, stage 2 blurs rows (for tile rows plus the
 * vertical halo), stage 3 blurs columns and writes G, DoG and the decimated
 * seed of the next octave. */
#define BT_W 64
#define BT_H 32
#define BT_RMAX 19
#define BT_IN_W (BT_W + 2 * BT_RMAX + 2) /* 104, even row length */
#define BT_IN_H (BT_H + 2 * BT_RMAX)     /* 70 */

__global__ void __launch_bounds__(256) blur_step_kernel(const __grid_constant__ BlurStep S)
{
  __shared__ float s_in[BT_IN_H][BT_IN_W];
  __shared__ float s_mid[BT_IN_H][BT_W];

  int pi = 0;
#pragma unroll
  for (int i = 1; i < VKS_MAX_PASSES_PER_STEP; i++)
    if (i < S.n_pass && (int)blockIdx.x >= S.pass[i].tile_begin)
      pi = i;
  const BlurPass &p = S.pass[pi];
  const int t = (int)blockIdx.x - p.tile_begin;
  const int x0 = (t % p.tiles_x) * BT_W;
  const int y0 = (t / p.tiles_x) * BT_H;
  const int R = p.radius;
  const int in_w = BT_W + 2 * R, in_h = BT_H + 2 * R;
  const int tid = threadIdx.x;

  for (int i = tid; i < in_w * in_h; i += 256)
  {
    const int yy = i / in_w, xx = i - yy * in_w;
    s_in[yy][xx] = fetch_src(p, x0 - R + xx, y0 - R + yy);
  }
  __syncthreads();

  for (int i = tid; i < in_h * BT_W; i += 256)
  {
    const int yy = i / BT_W, xx = i - yy * BT_W;
    const float *row = &s_in[yy][xx + R];
    float acc = vks_mul(row[0], p.taps[0]);
    for (int k = 1; k <= R; k++)
      acc = vks_blur_tap(acc, row[k], row[-k], p.taps[k]);
    s_mid[yy][xx] = acc;
  }
  __syncthreads();

  for (int i = tid; i < BT_H * BT_W; i += 256)
  {
    const int yy = i / BT_W, xx = i - yy * BT_W;
    const int x = x0 + xx, y = y0 + yy;
    if (x >= p.w || y >= p.h)
      continue;
    float acc = vks_mul(s_mid[yy + R][xx], p.taps[0]);
    for (int k = 1; k <= R; k++)
      acc = vks_blur_tap(acc, s_mid[yy + R + k][xx], s_mid[yy + R - k][xx], p.taps[k]);
    p.dst_g[(size_t)y * p.dst_pitch + x] = acc;
    if (p.dst_d)
      p.dst_d[(size_t)y * p.dst_pitch + x] = vks_sub(acc, s_in[yy + R][xx + R]);
    if (p.dst_next && (x & 1) && (y & 1))
    {
      const int nx = x >> 1, ny = y >> 1;
      if (nx < p.next_w && ny < p.next_h)
        p.dst_next[(size_t)ny * p.next_pitch + nx] = acc;
    }
  }
}

cudaError_t launch_blur_step(const BlurStep &step, cudaStream_t st)
{
  if (step.n_tiles <= 0)
    return cudaSuccess;
  blur_step_kernel<<<step.n_tiles, 256, 0, st>>>(step);
  return cudaGetLastError();
}

} // namespace vks
