/*
 * sift_oracle.c -- plain-C restatement of the reference detect + match path.
 * TEST INFRASTRUCTURE ONLY (see sift_oracle.h).  Single source of arithmetic
 * truth for the parity tests; follows the reference file:line cited at every
 * function (paths relative to the reference repository root, "SURVEY B-Dn"
 * = the documented decision for a behaviour the reference leaves undefined).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp (oracle/Makefile).
 */
#include "sift_oracle.h"

#include "../include/vksift_arith.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

struct vkso_Context
{
  vkso_Config cfg;
  int ns;          /* nb_scales_per_octave */
  int max_octaves; /* sift_memory.c:655-660 */
  uint32_t max_image_size;
  /* tables, sift_detector.c:52-145 */
  uint32_t ksize[VKSO_MAX_KERNEL];
  float ktab[VKSO_MAX_KERNEL][VKSO_MAX_KERNEL]; /* [scale][20], scale < ns+3 <= 20 */
  uint32_t radius[VKSO_MAX_KERNEL];
  float etap[VKSO_MAX_KERNEL][VKSO_MAX_KERNEL + 1];
  /* current pyramid */
  int n_oct;
  uint32_t ow[VKSO_MAX_OCTAVES], oh[VKSO_MAX_OCTAVES];
  float *G[VKSO_MAX_OCTAVES]; /* (ns+3) layers */
  float *D[VKSO_MAX_OCTAVES]; /* (ns+2) layers */
  size_t alloc_px[VKSO_MAX_OCTAVES];
  /* sections */
  uint32_t cap[VKSO_MAX_OCTAVES], found[VKSO_MAX_OCTAVES], kept[VKSO_MAX_OCTAVES], prim[VKSO_MAX_OCTAVES];
  vkso_Feature *feat[VKSO_MAX_OCTAVES];
  double t_stage[4];
};

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void vkso_default_config(vkso_Config *c)
{
  /* vulkansift.c:47-64 */
  c->input_image_max_size = 1920u * 1080u;
  c->max_nb_sift_per_buffer = 100000u;
  c->use_input_upsampling = 1;
  c->nb_octaves = 0;
  c->nb_scales_per_octave = 3;
  c->input_image_blur_level = 0.5f;
  c->seed_scale_sigma = 1.6f;
  c->intensity_threshold = 0.04f;
  c->edge_threshold = 10.f;
  c->max_nb_orientation_per_keypoint = 4;
  c->use_vlfeat_format = 0;
  c->use_interpolated_blur = 1;
  c->use_fp16_pyramid = 0;
  c->nb_threads = 0;
}

/* ---- tables ------------------------------------------------------------- */

/* sift_detector.c:52-145 (setupGaussianKernels) + the effective-tap expansion
 * of the paired-tap table consumed by GaussianBlurInterpolated.comp:34-44
 * (SURVEY A.2, B-D1, B-D2). */
static void setup_kernels(vkso_Context *ctx)
{
  const uint32_t nb_scales = (uint32_t)ctx->ns;
  const float seed_sigma = ctx->cfg.seed_scale_sigma;
  memset(ctx->ktab, 0, sizeof(ctx->ktab));
  memset(ctx->etap, 0, sizeof(ctx->etap));
  for (uint32_t s = 0; s < nb_scales + 3; s++)
  {
    float sigma;
    if (s == 0)
    {
      float init_blur = ctx->cfg.use_input_upsampling ? ctx->cfg.input_image_blur_level * 2.f : ctx->cfg.input_image_blur_level;
      sigma = sqrtf((seed_sigma * seed_sigma) - (init_blur * init_blur));
    }
    else
    {
      float sig_prev = powf(powf(2.f, 1.f / nb_scales), (float)(s - 1)) * seed_sigma;
      float sig_total = sig_prev * powf(2.f, 1.f / nb_scales);
      sigma = sqrtf(sig_total * sig_total - sig_prev * sig_prev);
    }
    uint32_t ksz = (uint32_t)(int)(ceilf(sigma * 4.f) + 1.f);
    if (ksz > VKSO_MAX_KERNEL)
      ksz = VKSO_MAX_KERNEL;
    ctx->ksize[s] = ksz;

    float c[VKSO_MAX_KERNEL];
    c[0] = 1.f;
    float sum = c[0];
    for (uint32_t i = 1; i < ksz; i++)
    {
      c[i] = (float)exp(-0.5 * powf((float)i, 2.f) / powf(sigma, 2.f));
      sum += 2 * c[i];
    }
    for (uint32_t i = 0; i < ksz; i++)
      c[i] /= sum;

    float *k = ctx->ktab[s];
    float *e = ctx->etap[s];
    if (ctx->cfg.use_interpolated_blur)
    {
      k[0] = c[0];
      k[1] = 0.f;
      e[0] = c[0];
      uint32_t r = 0;
      for (uint32_t d = 1, ki = 1; (d + 1) < ksz; d += 2, ki++)
      {
        float w = c[d] + c[d + 1];
        float off = (((float)d * c[d]) + ((float)(d + 1) * c[d + 1])) / (c[d] + c[d + 1]);
        k[ki * 2] = w;
        k[ki * 2 + 1] = off;
        /* bilinear fetch at texel offset `off`: weight (1-f) on tap d, f on tap d+1 */
        float f = off - (float)d;
        e[d] = w * (1.0f - f);
        e[d + 1] = w * f;
        r = d + 1;
      }
      ctx->radius[s] = r; /* an unpaired last tap is dropped, not renormalised */
    }
    else
    {
      for (uint32_t i = 0; i < ksz; i++)
      {
        k[i] = c[i];
        e[i] = c[i];
      }
      ctx->radius[s] = ksz - 1;
    }
  }
}

/* sift_memory.c:15-38 (updateScaleSpaceInfo) */
static void update_scale_space(vkso_Context *ctx, uint32_t w, uint32_t h)
{
  uint32_t lowest = (w > h) ? h : w;
  int up = ctx->cfg.use_input_upsampling ? 1 : 0;
  /* the reference stores the float expression into a uint32_t */
  float nf = log2f((float)lowest) - 4 + (float)up;
  uint32_t n = (nf > 0.f) ? (uint32_t)nf : 0u;
  if ((uint32_t)ctx->max_octaves < n)
    n = (uint32_t)ctx->max_octaves;
  ctx->n_oct = (int)n;
  float sf = up ? 0.5f : 1.f;
  for (int o = 0; o < ctx->n_oct; o++)
  {
    ctx->ow[o] = (uint32_t)((1.f / (powf(2.f, (float)o) * sf)) * (float)w);
    ctx->oh[o] = (uint32_t)((1.f / (powf(2.f, (float)o) * sf)) * (float)h);
  }
}

/* sift_memory.c:40-87 (updateBufferInfo): per-octave section capacities */
static void update_sections(vkso_Context *ctx)
{
  memset(ctx->cap, 0, sizeof(ctx->cap));
  float maxf = (float)ctx->cfg.max_nb_sift_per_buffer;
  float halves = maxf - powf(0.5f, (float)ctx->n_oct) * maxf;
  float corr = maxf / halves;
  for (int i = 0; i < ctx->n_oct; i++)
    ctx->cap[i] = (uint32_t)floorf((powf(0.5f, (float)(i + 1)) * maxf) * corr);
}

vkso_Context *vkso_create(const vkso_Config *cfg)
{
  vkso_Context *ctx = (vkso_Context *)calloc(1, sizeof(vkso_Context));
  ctx->cfg = *cfg;
  ctx->ns = cfg->nb_scales_per_octave;
  /* sift_memory.c:644-660 */
  uint32_t side = (uint32_t)ceilf(sqrtf((float)cfg->input_image_max_size));
  ctx->max_image_size = side * side;
  float mf = log2f((float)side) - 4 + (cfg->use_input_upsampling ? 1 : 0);
  ctx->max_octaves = (mf > 0.f) ? (int)(uint32_t)mf : 0;
  if (cfg->nb_octaves > 0 && cfg->nb_octaves < ctx->max_octaves)
    ctx->max_octaves = cfg->nb_octaves;
  if (ctx->max_octaves > VKSO_MAX_OCTAVES)
    ctx->max_octaves = VKSO_MAX_OCTAVES;
  setup_kernels(ctx);
  update_scale_space(ctx, side, side);
  update_sections(ctx);
  return ctx;
}

void vkso_destroy(vkso_Context *ctx)
{
  if (!ctx)
    return;
  for (int o = 0; o < VKSO_MAX_OCTAVES; o++)
  {
    free(ctx->G[o]);
    free(ctx->D[o]);
    free(ctx->feat[o]);
  }
  free(ctx);
}

int vkso_max_octaves(const vkso_Context *ctx) { return ctx->max_octaves; }
void vkso_kernel_table(const vkso_Context *ctx, uint32_t *ksize, float *k)
{
  for (int s = 0; s < ctx->ns + 3; s++)
  {
    ksize[s] = ctx->ksize[s];
    memcpy(k + s * VKSO_MAX_KERNEL, ctx->ktab[s], sizeof(float) * VKSO_MAX_KERNEL);
  }
}
void vkso_effective_taps(const vkso_Context *ctx, uint32_t *radius, float *e)
{
  for (int s = 0; s < ctx->ns + 3; s++)
  {
    radius[s] = ctx->radius[s];
    memcpy(e + s * (VKSO_MAX_KERNEL + 1), ctx->etap[s], sizeof(float) * (VKSO_MAX_KERNEL + 1));
  }
}

/* ---- fp16 storage mode (SURVEY B-D11): round-to-nearest-even through binary16 */
static float round_through_half(float f)
{
  uint32_t x = vks_f2u(f);
  uint32_t sign = x & 0x80000000u;
  uint32_t ax = x & 0x7fffffffu;
  if (ax >= 0x7f800000u)
    return f; /* inf/nan */
  if (ax >= 0x477ff000u)
    return vks_u2f(sign | 0x7f800000u); /* >= 65520 rounds to inf */
  if (ax < 0x33000001u)
    return vks_u2f(sign); /* < 2^-25 (or == 2^-25, tie to even) -> 0 */
  int e = (int)(ax >> 23) - 127;
  int drop = (e >= -14) ? 13 : (13 + (-14 - e)); /* mantissa bits lost */
  uint32_t m = (ax & 0x7fffffu) | 0x800000u;
  uint32_t half_ulp = 1u << (drop - 1);
  uint32_t rem = m & ((1u << drop) - 1);
  m >>= drop;
  if (rem > half_ulp || (rem == half_ulp && (m & 1)))
    m++;
  /* rebuild: value = m * 2^(e-23+drop) */
  float v = (float)m * vks_pow2i(e - 23 + drop > -126 ? e - 23 + drop : -126);
  if (e - 23 + drop < -126)
    v = 0.f; /* unreachable for half range */
  return vks_u2f(vks_f2u(v) | sign);
}
static void store_precision(vkso_Context *ctx, float *p, size_t n)
{
  if (!ctx->cfg.use_fp16_pyramid)
    return;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; i++)
    p[i] = round_through_half(p[i]);
}

/* ---- scale space -------------------------------------------------------- */

/* GaussianBlur.comp:32-44 / GaussianBlurInterpolated.comp:32-44 with the
 * MIRRORED_REPEAT sampler of sift_detector.c:208-225.  One separable pass. */
static void blur_pass(const float *in, float *out, int w, int h, const float *e, int r, int vertical)
{
#pragma omp parallel for schedule(static)
  for (int y = 0; y < h; y++)
  {
    for (int x = 0; x < w; x++)
    {
      float acc = vks_mul(in[(size_t)y * w + x], e[0]);
      for (int i = 1; i <= r; i++)
      {
        float a, b;
        if (!vertical)
        {
          a = in[(size_t)y * w + vks_mirror(x + i, w)];
          b = in[(size_t)y * w + vks_mirror(x - i, w)];
        }
        else
        {
          a = in[(size_t)vks_mirror(y + i, h) * w + x];
          b = in[(size_t)vks_mirror(y - i, h) * w + x];
        }
        acc = vks_blur_tap(acc, a, b, e[i]);
      }
      out[(size_t)y * w + x] = acc;
    }
  }
}

/* sift_detector.c:860-916: u8 staging -> R8_UNORM image -> vkCmdBlitImage LINEAR
 * (clamp-to-edge) into octave 0 layer 0.  Vulkan blit: u = (i+0.5)*srcW/dstW,
 * bilinear around u-0.5. */
static void seed_image(const uint8_t *img, int sw, int sh, float *dst, int dw, int dh)
{
  if (dw == sw && dh == sh)
  {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)sw * sh; i++)
      dst[i] = vks_unorm8(img[i]);
    return;
  }
  float sx = (float)sw / (float)dw, sy = (float)sh / (float)dh;
#pragma omp parallel for schedule(static)
  for (int j = 0; j < dh; j++)
  {
    float v = ((float)j + 0.5f) * sy - 0.5f;
    float vf = floorf(v);
    float fy = v - vf;
    int y0 = (int)vf, y1 = y0 + 1;
    y0 = y0 < 0 ? 0 : (y0 > sh - 1 ? sh - 1 : y0);
    y1 = y1 < 0 ? 0 : (y1 > sh - 1 ? sh - 1 : y1);
    for (int i = 0; i < dw; i++)
    {
      float u = ((float)i + 0.5f) * sx - 0.5f;
      float uf = floorf(u);
      float fx = u - uf;
      int x0 = (int)uf, x1 = x0 + 1;
      x0 = x0 < 0 ? 0 : (x0 > sw - 1 ? sw - 1 : x0);
      x1 = x1 < 0 ? 0 : (x1 > sw - 1 ? sw - 1 : x1);
      float top = vks_lerp(vks_unorm8(img[(size_t)y0 * sw + x0]), vks_unorm8(img[(size_t)y0 * sw + x1]), fx);
      float bot = vks_lerp(vks_unorm8(img[(size_t)y1 * sw + x0]), vks_unorm8(img[(size_t)y1 * sw + x1]), fx);
      dst[(size_t)j * dw + i] = vks_lerp(top, bot, fy);
    }
  }
}

/* sift_detector.c:1003-1034: vkCmdBlitImage NEAREST of layer ns into the next
 * octave's layer 0: src = floor((i+0.5)*srcW/dstW) */
static void downsample_nearest(const float *src, int sw, int sh, float *dst, int dw, int dh)
{
  float sx = (float)sw / (float)dw, sy = (float)sh / (float)dh;
#pragma omp parallel for schedule(static)
  for (int j = 0; j < dh; j++)
  {
    int ys = (int)floorf(((float)j + 0.5f) * sy);
    if (ys > sh - 1)
      ys = sh - 1;
    for (int i = 0; i < dw; i++)
    {
      int xs = (int)floorf(((float)i + 0.5f) * sx);
      if (xs > sw - 1)
        xs = sw - 1;
      dst[(size_t)j * dw + i] = src[(size_t)ys * sw + xs];
    }
  }
}

/* sift_detector.c:893-1079: blur chain per octave, then DoG (DifferenceOfGaussian.comp:12-15) */
static void build_pyramid(vkso_Context *ctx, const uint8_t *img, int w, int h)
{
  const int ns = ctx->ns;
  size_t max_px = 0;
  for (int o = 0; o < ctx->n_oct; o++)
  {
    size_t px = (size_t)ctx->ow[o] * ctx->oh[o];
    if (px > max_px)
      max_px = px;
    if (px > ctx->alloc_px[o])
    {
      free(ctx->G[o]);
      free(ctx->D[o]);
      ctx->G[o] = (float *)malloc(sizeof(float) * px * (size_t)(ns + 3));
      ctx->D[o] = (float *)malloc(sizeof(float) * px * (size_t)(ns + 2));
      ctx->alloc_px[o] = px;
    }
  }
  float *tmp = (float *)malloc(sizeof(float) * (max_px ? max_px : 1));
  for (int o = 0; o < ctx->n_oct; o++)
  {
    const int ow = (int)ctx->ow[o], oh = (int)ctx->oh[o];
    const size_t px = (size_t)ow * oh;
    float *G = ctx->G[o];
    if (o == 0)
    {
      seed_image(img, w, h, G, ow, oh);
      blur_pass(G, tmp, ow, oh, ctx->etap[0], (int)ctx->radius[0], 0);
      blur_pass(tmp, G, ow, oh, ctx->etap[0], (int)ctx->radius[0], 1);
    }
    else
    {
      downsample_nearest(ctx->G[o - 1] + (size_t)ns * ctx->ow[o - 1] * ctx->oh[o - 1], (int)ctx->ow[o - 1], (int)ctx->oh[o - 1], G, ow, oh);
    }
    store_precision(ctx, G, px);
    for (int s = 1; s < ns + 3; s++)
    {
      blur_pass(G + (size_t)(s - 1) * px, tmp, ow, oh, ctx->etap[s], (int)ctx->radius[s], 0);
      blur_pass(tmp, G + (size_t)s * px, ow, oh, ctx->etap[s], (int)ctx->radius[s], 1);
      store_precision(ctx, G + (size_t)s * px, px);
    }
    float *D = ctx->D[o];
    for (int s = 0; s < ns + 2; s++)
    {
      const float *a = G + (size_t)(s + 1) * px, *b = G + (size_t)s * px;
      float *d = D + (size_t)s * px;
#pragma omp parallel for schedule(static)
      for (long i = 0; i < (long)px; i++)
        d[i] = vks_sub(a[i], b[i]);
      store_precision(ctx, d, px);
    }
  }
  free(tmp);
}

/* ---- extrema + refinement (ExtractKeypoints.comp) ----------------------- */

typedef struct
{
  const float *D;
  int w, h, ns;
} dog_view;

/* imageLoad on the DoG array; layer ns+2 does not exist -> 0 (SURVEY B-D3) */
static inline float dogv(const dog_view *v, int s, int x, int y)
{
  if (s < 0 || s >= v->ns + 2 || x < 0 || x >= v->w || y < 0 || y >= v->h)
    return 0.f;
  return v->D[((size_t)s * v->h + y) * v->w + x];
}

/* ExtractKeypoints.comp:57-116 */
static int is_extremum(const dog_view *v, int s, int x, int y, float prefilter)
{
  float c = dogv(v, s, x, y);
  if (!(fabsf(c) > prefilter))
    return 0;
  int gt = 1, lt = 1;
  for (int ds = -1; ds <= 1; ds++)
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++)
      {
        if (!ds && !dy && !dx)
          continue;
        float n = dogv(v, s + ds, x + dx, y + dy);
        gt &= (c > n);
        lt &= (c < n);
      }
  return gt | lt;
}

/* ExtractKeypoints.comp:118-224.  Returns 1 and fills f when accepted. */
static int refine_keypoint(const dog_view *v, int x, int y, int s, int octave_idx, float sigma0, float thr, float edge_limit, vkso_Feature *f)
{
  const int w = v->w, h = v->h, ns = v->ns;
  float oX = 0.f, oY = 0.f, oS = 0.f;
  float gX = 0.f, gY = 0.f, gS = 0.f;
  int rx = x, ry = y, rs = s;
  for (int step = 0; step < 5; step++)
  {
    float c = dogv(v, rs, rx, ry);
    float sp = dogv(v, rs + 1, rx, ry), sm = dogv(v, rs - 1, rx, ry);
    float xp = dogv(v, rs, rx + 1, ry), xm = dogv(v, rs, rx - 1, ry);
    float yp = dogv(v, rs, rx, ry + 1), ym = dogv(v, rs, rx, ry - 1);
    gS = 0.5f * (sp - sm);
    gX = 0.5f * (xp - xm);
    gY = 0.5f * (yp - ym);
    float h11 = sp + sm - 2.f * c;
    float h22 = xp + xm - 2.f * c;
    float h33 = yp + ym - 2.f * c;
    float h12 = 0.25f * (dogv(v, rs + 1, rx + 1, ry) - dogv(v, rs + 1, rx - 1, ry) - dogv(v, rs - 1, rx + 1, ry) + dogv(v, rs - 1, rx - 1, ry));
    float h13 = 0.25f * (dogv(v, rs + 1, rx, ry + 1) - dogv(v, rs + 1, rx, ry - 1) - dogv(v, rs - 1, rx, ry + 1) + dogv(v, rs - 1, rx, ry - 1));
    float h23 = 0.25f * (dogv(v, rs, rx + 1, ry + 1) - dogv(v, rs, rx + 1, ry - 1) - dogv(v, rs, rx - 1, ry + 1) + dogv(v, rs, rx - 1, ry - 1));

    float det = h11 * ((h22 * h33) - (h23 * h23)) - h12 * ((h12 * h33) - (h13 * h23)) + h13 * ((h12 * h23) - (h13 * h22));
    if (det != 0.0f)
    {
      float i11 = ((h22 * h33) - (h23 * h23)) / det;
      float i12 = -1.f * ((h12 * h33) - (h13 * h23)) / det;
      float i13 = ((h12 * h23) - (h13 * h22)) / det;
      float i22 = ((h11 * h33) - (h13 * h13)) / det;
      float i23 = -1.f * ((h11 * h23) - (h13 * h12)) / det;
      float i33 = ((h11 * h22) - (h12 * h12)) / det;
      oS = -i11 * gS - i12 * gX - i13 * gY;
      oX = -i12 * gS - i22 * gX - i23 * gY;
      oY = -i13 * gS - i23 * gX - i33 * gY;
    }
    else
    {
      return 0;
    }
    if (fabsf(oX) < 0.6f && fabsf(oY) < 0.6f && fabsf(oS) < 0.6f)
      break;
    else if (step < 4)
    {
      rx += ((oX >= 0.6f && rx < (w - 2)) ? 1 : 0) + ((oX <= -0.6f && rx > 1) ? -1 : 0);
      ry += ((oY >= 0.6f && ry < (h - 2)) ? 1 : 0) + ((oY <= -0.6f && ry > 1) ? -1 : 0);
      rs += ((oS >= 0.6f && rs < (ns + 1)) ? 1 : 0) + ((oS <= -0.6f && rs > 1) ? -1 : 0);
    }
  }
  float px = (float)rx + oX, py = (float)ry + oY, ps = (float)rs + oS;
  float val = dogv(v, rs, rx, ry) + 0.5f * (gX * oX + gY * oY + gS * oS);
  if (!(fabsf(val) > thr && fabsf(oX) < 1.5f && fabsf(oY) < 1.5f && fabsf(oS) < 1.5f && px >= 0.f && px < (float)w && py >= 0.f && py < (float)h &&
        ps >= 0.f && ps <= (float)(ns + 1)))
    return 0;
  float c = dogv(v, rs, rx, ry);
  float e11 = dogv(v, rs, rx + 1, ry) + dogv(v, rs, rx - 1, ry) - 2.f * c;
  float e22 = dogv(v, rs, rx, ry + 1) + dogv(v, rs, rx, ry - 1) - 2.f * c;
  float e12 = 0.25f * (dogv(v, rs, rx + 1, ry + 1) - dogv(v, rs, rx + 1, ry - 1) - dogv(v, rs, rx - 1, ry + 1) + dogv(v, rs, rx - 1, ry - 1));
  float edgeness = ((e11 + e22) * (e11 + e22)) / ((e11 * e22) - (e12 * e12));
  if (!((edgeness < edge_limit) && (edgeness >= 0.f)))
    return 0;
  float sf = vks_pow2i(octave_idx);
  memset(f, 0, sizeof(*f));
  f->scale_x = px;
  f->scale_y = py;
  f->scale_idx = (uint32_t)vks_rint(ps);
  f->octave_idx = octave_idx;
  f->sigma = sigma0 * vks_exp2f(ps / (float)ns) * sf;
  f->orientation = 0.f;
  f->intensity = val;
  f->x = px * sf;
  f->y = py * sf;
  return 1;
}

/* ---- orientation (ComputeOrientation.comp) ------------------------------ */

typedef struct
{
  const float *G;
  int w, h, nl;
} gauss_view;

/* imageLoad on the Gaussian array, out of bounds -> 0 (SURVEY B-D3) */
static inline float gv(const gauss_view *v, int s, int x, int y)
{
  if (s < 0 || s >= v->nl || x < 0 || x >= v->w || y < 0 || y >= v->h)
    return 0.f;
  return v->G[((size_t)s * v->h + y) * v->w + x];
}

/* Returns the number of orientations (0..36) in bin order, written to ori[]. */
static int compute_orientations(const gauss_view *v, const vkso_Feature *kp, float *ori)
{
  uint32_t hist[36], tmp[36];
  memset(hist, 0, sizeof(hist));
  float sf = vks_pow2i(kp->octave_idx);
  float lambda = 1.5f * (kp->sigma / sf);
  int r = (int)floorf(3 * lambda);
  float es = -1.f / (2.f * lambda * lambda);
  /* :75-81 fixed-point scale */
  float m = 0.f;
  for (int i = -r; i <= r; i++)
    for (int j = -r; j <= r; j++)
      m += vks_expf(es * (float)((i * i) + (j * j))) * VKS_SQRT2_F;
  float fp = (float)(1u << (uint32_t)(30 - vks_ceil_log2(m)));
  float rsx = vks_rint(kp->scale_x), rsy = vks_rint(kp->scale_y);
  int s = (int)kp->scale_idx;
  for (int dy = -r; dy <= r; dy++)
    for (int dx = -r; dx <= r; dx++)
    {
      int gx = (int)rsx + dx, gy = (int)rsy + dy;
      float sdx = (rsx + (float)dx) - kp->scale_x;
      float sdy = (rsy + (float)dy) - kp->scale_y;
      float d2 = (sdx * sdx) + (sdy * sdy);
      if ((gx < 1 || gx >= (v->w - 1) || gy < 1 || gy >= (v->h - 1)) && (d2 > (float)(r * r)))
        continue;
      float gX = 0.5f * (gv(v, s, gx + 1, gy) - gv(v, s, gx - 1, gy));
      float gY = 0.5f * (gv(v, s, gx, gy + 1) - gv(v, s, gx, gy - 1));
      float mag = vks_expf(d2 * es) * vks_sqrt((gX * gX) + (gY * gY));
      float th = vks_atan2f(gY, gX);
      if (th < 0.f)
        th += VKS_TWO_PI_F;
      else if (th > VKS_TWO_PI_F)
        th -= VKS_TWO_PI_F;
      int bin = (int)((th * 36.f) / VKS_TWO_PI_F);
      if (bin < 0)
        bin += 36;
      else if (bin >= 36)
        bin -= 36;
      hist[bin] += (uint32_t)(mag * fp);
    }
  /* :130-147 */
  for (int it = 0; it < 3; it++)
  {
    for (int i = 0; i < 36; i++)
      tmp[i] = (uint32_t)((float)(hist[(i + 35) % 36] + hist[i] + hist[(i + 1) % 36]) / 3.f);
    for (int i = 0; i < 36; i++)
      hist[i] = (uint32_t)((float)(tmp[(i + 35) % 36] + tmp[i] + tmp[(i + 1) % 36]) / 3.f);
  }
  uint32_t mx = 0;
  for (int i = 0; i < 36; i++)
    if (hist[i] > mx)
      mx = hist[i];
  int n = 0;
  for (int i = 0; i < 36; i++)
  {
    uint32_t hp = hist[(i + 35) % 36], hn = hist[(i + 1) % 36], hc = hist[i];
    if (((float)hc >= (0.8f * (float)mx)) && (hc > hp) && (hc > hn))
    {
      /* uint32 differences wrap before the float conversion (SURVEY B-D12) */
      float num = (float)(uint32_t)(hp - hn);
      float den = (float)(uint32_t)(hp - (2u * hc) + hn);
      float idx = (float)i + 0.5f * (num / den);
      ori[n++] = ((idx + 0.5f) * VKS_TWO_PI_F) / 36.f;
    }
  }
  return n;
}

/* ---- descriptor (ComputeDescriptors.comp) ------------------------------- */
static void compute_descriptor(const gauss_view *v, vkso_Feature *kp, int vlfeat)
{
  uint32_t desc[128];
  memset(desc, 0, sizeof(desc));
  float sf = vks_pow2i(kp->octave_idx);
  float lambda = 3.0f * (kp->sigma / sf);
  float radius = VKS_SQRT2_F * lambda * 5.f * 0.5f;
  int R = (int)floorf(radius + 0.5f);
  float sn, cs;
  vks_sincosf(kp->orientation, &sn, &cs);
  float kc = cs / lambda, ks = sn / lambda;
  const float es = -0.125f;
  float m = 0.f;
  for (int i = 0; i < R / 2; i++)
  {
    m += vks_expf(es * (float)((i * i) + (i * i))) * VKS_SQRT2_F;
    for (int j = i + 1; j < R / 2; j++)
      m += vks_expf(es * (float)((i * i) + (j * j))) * VKS_SQRT2_F * 2.f;
  }
  float fp = (float)(1u << (uint32_t)(16 - vks_ceil_log2(m)));
  float rsx = vks_rint(kp->scale_x), rsy = vks_rint(kp->scale_y);
  int s = (int)kp->scale_idx;
  for (int dy = -R; dy <= R; dy++)
    for (int dx = -R; dx <= R; dx++)
    {
      int ix = (int)rsx + dx, iy = (int)rsy + dy;
      float sdx = (rsx + (float)dx) - kp->scale_x;
      float sdy = (rsy + (float)dy) - kp->scale_y;
      if (ix < 1 || ix >= (v->w - 1) || iy < 1 || iy >= (v->h - 1))
        continue;
      float ox = kc * sdx + ks * sdy;
      float oy = kc * sdy - ks * sdx;
      float gX = 0.5f * (gv(v, s, ix + 1, iy) - gv(v, s, ix - 1, iy));
      float gY = 0.5f * (gv(v, s, ix, iy + 1) - gv(v, s, ix, iy - 1));
      float th = vks_atan2f(gY, gX);
      if (th < 0.f)
        th += VKS_TWO_PI_F;
      else if (th > VKS_TWO_PI_F)
        th -= VKS_TWO_PI_F;
      th = th - kp->orientation;
      if (th < 0.f)
        th += VKS_TWO_PI_F;
      else if (th > VKS_TWO_PI_F)
        th -= VKS_TWO_PI_F;
      float mag = vks_expf(es * ((ox * ox) + (oy * oy))) * vks_sqrt((gX * gX) + (gY * gY));
      float fx = ox + 2.f, fy = oy + 2.f;
      float fb = vlfeat ? ((th * 8.f) / VKS_TWO_PI_F) : ((-th * 8.f) / VKS_TWO_PI_F);
      int hx = (int)floorf(fx - 0.5f), hy = (int)floorf(fy - 0.5f), hb = (int)floorf(fb);
      float rx = fx - ((float)hx + 0.5f), ry = fy - ((float)hy + 0.5f), rb = fb - (float)hb;
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
          for (int k = 0; k < 2; k++)
            if ((i + hx) >= 0 && (i + hx) < 4 && (j + hy) >= 0 && (j + hy) < 4)
            {
              /* signed GLSL % lowers to OpSMod: non-negative result (SURVEY B-D7) */
              int b = (((k + hb) % 8) + 8) % 8;
              int idx = (j + hy) * 32 + (i + hx) * 8 + b;
              float val = fabsf(1.f - (float)i - rx) * fabsf(1.f - (float)j - ry) * fabsf(1.f - (float)k - rb) * mag;
              desc[idx] += (uint32_t)(val * fp);
            }
    }
  /* :209-265 */
  uint32_t acc = 0;
  for (int i = 0; i < 128; i++)
    acc += desc[i] * desc[i];
  float norm = vks_sqrt((float)acc);
  uint32_t clampv = (uint32_t)(norm * 0.2f);
  acc = 0;
  for (int i = 0; i < 128; i++)
  {
    if (desc[i] > clampv)
      desc[i] = clampv;
    acc += desc[i] * desc[i];
  }
  norm = vks_sqrt((float)acc);
  float inv = 512.f / norm;
  for (int i = 0; i < 128; i++)
  {
    float d = (float)desc[i] * inv;
    uint8_t o;
    if (!(d == d))
      o = 0; /* 0 * inf of an empty descriptor */
    else if (d > 255.f)
      o = 255;
    else
      o = (uint8_t)(uint32_t)d;
    kp->descriptor[i] = o;
  }
}

/* ---- full detection ----------------------------------------------------- */
uint32_t vkso_detect(vkso_Context *ctx, const uint8_t *image, uint32_t width, uint32_t height)
{
#ifdef _OPENMP
  if (ctx->cfg.nb_threads > 0)
    omp_set_num_threads(ctx->cfg.nb_threads);
#endif
  const int ns = ctx->ns;
  update_scale_space(ctx, width, height);
  update_sections(ctx);
  double t0 = now_s();
  build_pyramid(ctx, image, (int)width, (int)height);
  double t1 = now_s();
  ctx->t_stage[0] = t1 - t0;
  ctx->t_stage[1] = ctx->t_stage[2] = ctx->t_stage[3] = 0.0;

  const float thr = ctx->cfg.intensity_threshold / (float)ns; /* sift_detector.c:1136 */
  const float prefilter = thr * 0.8f;
  const float er = ctx->cfg.edge_threshold;
  const float edge_limit = ((er + 1.f) * (er + 1.f)) / er; /* pow(edge+1,2)/edge */
  uint32_t total = 0;
  for (int o = 0; o < ctx->n_oct; o++)
  {
    const int w = (int)ctx->ow[o], h = (int)ctx->oh[o];
    const int octave_idx = o - (ctx->cfg.use_input_upsampling ? 1 : 0); /* sift_detector.c:1134 */
    const uint32_t cap = ctx->cap[o];
    free(ctx->feat[o]);
    ctx->feat[o] = (vkso_Feature *)malloc(sizeof(vkso_Feature) * (cap ? cap : 1));
    vkso_Feature *sec = ctx->feat[o];
    uint32_t count = 0; /* keeps counting past capacity like nb_elem */
    double ta = now_s();
    dog_view dv = {ctx->D[o], w, h, ns};
    /* detection order (s, y, x) = canonical order of SURVEY B-D4 */
    for (int s = 1; s <= ns; s++)
    {
      /* rows in parallel, appended in row order afterwards */
      vkso_Feature **rowf = (vkso_Feature **)calloc((size_t)h, sizeof(vkso_Feature *));
      int *rown = (int *)calloc((size_t)h, sizeof(int));
#pragma omp parallel for schedule(dynamic, 8)
      for (int y = 1; y < h - 1; y++)
      {
        int n = 0, capn = 0;
        vkso_Feature *lst = NULL;
        for (int x = 1; x < w - 1; x++)
        {
          if (!is_extremum(&dv, s, x, y, prefilter))
            continue;
          vkso_Feature f;
          if (!refine_keypoint(&dv, x, y, s, octave_idx, ctx->cfg.seed_scale_sigma, thr, edge_limit, &f))
            continue;
          if (n == capn)
          {
            capn = capn ? capn * 2 : 8;
            lst = (vkso_Feature *)realloc(lst, sizeof(vkso_Feature) * (size_t)capn);
          }
          lst[n++] = f;
        }
        rowf[y] = lst;
        rown[y] = n;
      }
      for (int y = 1; y < h - 1; y++)
      {
        for (int i = 0; i < rown[y]; i++)
        {
          if (count < cap)
            sec[count] = rowf[y][i];
          count++;
        }
        free(rowf[y]);
      }
      free(rowf);
      free(rown);
    }
    double tb = now_s();
    ctx->t_stage[1] += tb - ta;
    uint32_t nprim = count < cap ? count : cap;
    ctx->prim[o] = nprim;

    /* orientation: first peak in place, the others appended after all
     * primaries in (parent, bin) order (SURVEY B-D4) */
    gauss_view gvw = {ctx->G[o], w, h, ns + 3};
    float *oris = (float *)malloc(sizeof(float) * 36 * (nprim ? nprim : 1));
    int *nori = (int *)malloc(sizeof(int) * (nprim ? nprim : 1));
#pragma omp parallel for schedule(dynamic, 4)
    for (long i = 0; i < (long)nprim; i++)
      nori[i] = compute_orientations(&gvw, &sec[i], oris + 36 * i);
    const uint32_t max_ori = ctx->cfg.max_nb_orientation_per_keypoint;
    for (uint32_t i = 0; i < nprim; i++)
    {
      for (int k = 0; k < nori[i]; k++)
      {
        if (k == 0)
          sec[i].orientation = oris[36 * i];
        else if (max_ori == 0 || (uint32_t)k < max_ori)
        {
          if (count < cap)
          {
            sec[count] = sec[i];
            sec[count].orientation = oris[36 * i + k];
          }
          count++;
        }
      }
    }
    free(oris);
    free(nori);
    double tc = now_s();
    ctx->t_stage[2] += tc - tb;
    uint32_t nkept = count < cap ? count : cap;
    ctx->found[o] = count;
    ctx->kept[o] = nkept;
#pragma omp parallel for schedule(dynamic, 4)
    for (long i = 0; i < (long)nkept; i++)
      compute_descriptor(&gvw, &sec[i], ctx->cfg.use_vlfeat_format);
    ctx->t_stage[3] += now_s() - tc;
    total += nkept;
  }
  for (int o = ctx->n_oct; o < VKSO_MAX_OCTAVES; o++)
    ctx->found[o] = ctx->kept[o] = ctx->prim[o] = 0;
  return total;
}

int vkso_nb_octaves(const vkso_Context *ctx) { return ctx->n_oct; }
void vkso_octave_resolution(const vkso_Context *ctx, int o, uint32_t *w, uint32_t *h)
{
  *w = ctx->ow[o];
  *h = ctx->oh[o];
}
void vkso_section_capacity(const vkso_Context *ctx, uint32_t *caps) { memcpy(caps, ctx->cap, sizeof(uint32_t) * (size_t)ctx->n_oct); }
void vkso_section_counts(const vkso_Context *ctx, uint32_t *found, uint32_t *kept)
{
  memcpy(found, ctx->found, sizeof(uint32_t) * (size_t)ctx->n_oct);
  memcpy(kept, ctx->kept, sizeof(uint32_t) * (size_t)ctx->n_oct);
}
void vkso_primary_counts(const vkso_Context *ctx, uint32_t *p) { memcpy(p, ctx->prim, sizeof(uint32_t) * (size_t)ctx->n_oct); }
/* sift_memory.c:1108-1195: sections concatenated in octave order */
void vkso_get_features(const vkso_Context *ctx, vkso_Feature *out)
{
  size_t n = 0;
  for (int o = 0; o < ctx->n_oct; o++)
  {
    memcpy(out + n, ctx->feat[o], sizeof(vkso_Feature) * ctx->kept[o]);
    n += ctx->kept[o];
  }
}
const float *vkso_gaussian_layer(const vkso_Context *ctx, int o, int s) { return ctx->G[o] + (size_t)s * ctx->ow[o] * ctx->oh[o]; }
const float *vkso_dog_layer(const vkso_Context *ctx, int o, int s) { return ctx->D[o] + (size_t)s * ctx->ow[o] * ctx->oh[o]; }
void vkso_stage_seconds(const vkso_Context *ctx, double *t4) { memcpy(t4, ctx->t_stage, sizeof(double) * 4); }

/* ---- 2-NN match (Get2NearestNeighbors.comp:43-104) ---------------------- */
static inline float desc_dist(const uint8_t *a, const uint8_t *b)
{
  float dist = 0.f;
  for (int i = 0; i < 128; i++)
  {
    uint32_t d = (uint32_t)a[i] - (uint32_t)b[i];
    dist += (float)(uint32_t)(d * d); /* exact: sum <= 8 323 200 < 2^24 */
  }
  return sqrtf(dist);
}
static void match_strided(const uint8_t *a, size_t sa, uint32_t na, const uint8_t *b, size_t sb, uint32_t nb, vkso_Match *out, int nb_threads)
{
#ifdef _OPENMP
  if (nb_threads > 0)
    omp_set_num_threads(nb_threads);
#else
  (void)nb_threads;
#endif
#pragma omp parallel for schedule(static)
  for (long ia = 0; ia < (long)na; ia++)
  {
    const uint8_t *da = a + (size_t)ia * sa;
    float d0 = desc_dist(da, b), d1 = desc_dist(da, b + sb);
    float bd, sd;
    uint32_t bi, si;
    if (d0 < d1)
    {
      bd = d0, bi = 0, sd = d1, si = 1;
    }
    else
    {
      bd = d1, bi = 1, sd = d0, si = 0;
    }
    for (uint32_t ib = 2; ib < nb; ib++)
    {
      float d = desc_dist(da, b + (size_t)ib * sb);
      if (d < bd)
      {
        sd = bd, si = bi, bd = d, bi = ib;
      }
      else if (d < sd)
      {
        sd = d, si = ib;
      }
    }
    out[ia].idx_a = (uint32_t)ia;
    out[ia].idx_b1 = bi;
    out[ia].idx_b2 = si;
    out[ia].dist_a_b1 = bd;
    out[ia].dist_a_b2 = sd;
  }
}
void vkso_match(const uint8_t *a, uint32_t na, const uint8_t *b, uint32_t nb, vkso_Match *out, int nb_threads)
{
  match_strided(a, 128, na, b, 128, nb, out, nb_threads);
}
void vkso_match_features(const vkso_Feature *a, uint32_t na, const vkso_Feature *b, uint32_t nb, vkso_Match *out, int nb_threads)
{
  match_strided(a->descriptor, sizeof(vkso_Feature), na, b->descriptor, sizeof(vkso_Feature), nb, out, nb_threads);
}

/* ---- arithmetic probes -------------------------------------------------- */
/* the two Vulkan fixed-function steps of the path (restated from the Vulkan specification, not in the reference sources),
 * exposed so that a test can chain the reference's own shaders between them (tests/test_ref_chain.py) */
void vkso_seed_image(const uint8_t *img, int sw, int sh, float *dst, int dw, int dh) { seed_image(img, sw, sh, dst, dw, dh); }
void vkso_downsample_nearest(const float *src, int sw, int sh, float *dst, int dw, int dh) { downsample_nearest(src, sw, sh, dst, dw, dh); }

float vkso_expf(float x) { return vks_expf(x); }
float vkso_exp2f(float x) { return vks_exp2f(x); }
float vkso_atan2f(float y, float x) { return vks_atan2f(y, x); }
void vkso_sincosf(float t, float *s, float *c) { vks_sincosf(t, s, c); }
int vkso_ceil_log2(float m) { return vks_ceil_log2(m); }
int vkso_mirror(int i, int n) { return vks_mirror(i, n); }
