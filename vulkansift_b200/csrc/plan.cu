/*
 * plan.cu -- host-side tables of the detection pipeline.
 *
 * Product restatement of the reference's host arithmetic (the oracle has its
 * own, oracle/sift_oracle.c; tests compare the two through
 * vksiftx_getEffectiveTaps / vksiftx_getSectionCapacities):
 *   plan_gaussian_taps  <- sift_detector.c:52-145  setupGaussianKernels
 *   plan_max_octaves    <- sift_memory.c:644-660
 *   plan_octaves        <- sift_memory.c:15-38     updateScaleSpaceInfo
 *   plan_sections       <- sift_memory.c:40-87     updateBufferInfo
 * Host code is built with -ffp-contract=off so that gcc and g++ agree.
 */
#include "vksift_internal.h"

#include <cmath>
#include <cstring>

namespace vks
{

static const char TAG[] = "SiftPlan";

void plan_gaussian_taps(ScalePlan *sp, const vksift_Config *cfg)
{
  const uint32_t ns = cfg->nb_scales_per_octave;
  const float s0 = cfg->seed_scale_sigma;
  std::memset(sp, 0, sizeof(*sp));
  sp->ns = (int)ns;
  for (uint32_t layer = 0; layer < ns + 3; layer++)
  {
    /* incremental blur that takes layer-1 to layer (layer 0: input blur -> seed sigma) */
    float sigma;
    if (layer == 0)
    {
      const float in_blur = cfg->use_input_upsampling ? cfg->input_image_blur_level * 2.f : cfg->input_image_blur_level;
      sigma = sqrtf((s0 * s0) - (in_blur * in_blur));
    }
    else
    {
      const float prev = powf(powf(2.f, 1.f / ns), (float)(layer - 1)) * s0;
      const float next = prev * powf(2.f, 1.f / ns);
      sigma = sqrtf(next * next - prev * prev);
    }
    uint32_t ksize = (uint32_t)(int)(ceilf(sigma * 4.f) + 1.f);
    if (ksize > VKS_MAX_TAPS - 1)
    {
      LOGW(TAG, "Gaussian kernel for scale %u needs %u coefficients, more than the supported %d: the tail is ignored (use a smaller seed_scale_sigma).",
           layer, ksize, VKS_MAX_TAPS - 1);
      ksize = VKS_MAX_TAPS - 1;
    }
    sp->ksize[layer] = ksize;

    float half[VKS_MAX_TAPS];
    half[0] = 1.f;
    float norm = half[0];
    for (uint32_t i = 1; i < ksize; i++)
    {
      half[i] = (float)exp(-0.5 * powf((float)i, 2.f) / powf(sigma, 2.f));
      norm += 2 * half[i];
    }
    for (uint32_t i = 0; i < ksize; i++)
      half[i] /= norm;

    float *e = sp->taps[layer];
    if (cfg->use_hardware_interpolated_blur)
    {
      /* The reference folds taps (1,2),(3,4).. into one bilinear fetch each
       * (weight w at fractional offset off).  A bilinear fetch at offset
       * d+f is (1-f)*I[d] + f*I[d+1], so the convolution it performs has the
       * effective taps below; an unpaired last tap is dropped. */
      e[0] = half[0];
      uint32_t r = 0;
      for (uint32_t d = 1; (d + 1) < ksize; d += 2)
      {
        const float w = half[d] + half[d + 1];
        const float off = (((float)d * half[d]) + ((float)(d + 1) * half[d + 1])) / (half[d] + half[d + 1]);
        const float f = off - (float)d;
        e[d] = w * (1.0f - f);
        e[d + 1] = w * f;
        r = d + 1;
      }
      sp->radius[layer] = r;
    }
    else
    {
      for (uint32_t i = 0; i < ksize; i++)
        e[i] = half[i];
      sp->radius[layer] = ksize - 1;
    }
    LOGD(TAG, "scale %u: sigma=%f ksize=%u radius=%u", layer, sigma, ksize, sp->radius[layer]);
  }
}

uint32_t plan_max_octaves(const vksift_Config *cfg, uint32_t *side_out)
{
  const uint32_t side = (uint32_t)ceilf(sqrtf((float)cfg->input_image_max_size));
  if (side_out)
    *side_out = side;
  const float f = log2f((float)side) - 4 + (cfg->use_input_upsampling ? 1 : 0);
  uint32_t n = f > 0.f ? (uint32_t)f : 0u;
  if (cfg->nb_octaves > 0 && cfg->nb_octaves < n)
    n = cfg->nb_octaves;
  if (n > VKS_MAX_OCT)
    n = VKS_MAX_OCT;
  return n;
}

uint32_t plan_octaves(uint32_t w, uint32_t h, bool upsample, uint32_t max_octaves, uint32_t *ow, uint32_t *oh)
{
  const uint32_t lowest = w > h ? h : w;
  const float f = log2f((float)lowest) - 4 + (upsample ? 1 : 0);
  uint32_t n = f > 0.f ? (uint32_t)f : 0u;
  if (max_octaves < n)
    n = max_octaves;
  const float sf = upsample ? 0.5f : 1.f;
  for (uint32_t o = 0; o < n; o++)
  {
    ow[o] = (uint32_t)((1.f / (powf(2.f, (float)o) * sf)) * (float)w);
    oh[o] = (uint32_t)((1.f / (powf(2.f, (float)o) * sf)) * (float)h);
  }
  return n;
}

void plan_sections(uint32_t max_feats, uint32_t n_oct, uint32_t *cap)
{
  for (uint32_t i = 0; i < VKS_MAX_OCT; i++)
    cap[i] = 0;
  if (n_oct == 0)
    return;
  const float maxf = (float)max_feats;
  const float halves = maxf - powf(0.5f, (float)n_oct) * maxf;
  const float corr = maxf / halves;
  for (uint32_t i = 0; i < n_oct; i++)
    cap[i] = (uint32_t)floorf((powf(0.5f, (float)(i + 1)) * maxf) * corr);
}

} // namespace vks
