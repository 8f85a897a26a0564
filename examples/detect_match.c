/* Minimal C caller of the vksift_* API (the shape of the reference's README example, README.md:92-135):
 *   gcc examples/detect_match.c -Iinclude -Lvulkansift_b200/lib -lvulkansift -Wl,-rpath,$PWD/vulkansift_b200/lib -lm -o detect_match
 * Detects features on two synthetic images and matches them. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vulkansift/vulkansift.h>

static void blobs(uint8_t *img, int w, int h, int shift)
{
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
    {
      float v = 0.5f;
      for (int k = 0; k < 40; k++)
      {
        const float cx = (float)((k * 97 + shift) % w), cy = (float)((k * 57) % h), s = 3.f + (float)(k % 5);
        const float d2 = (x - cx) * (x - cx) + (y - cy) * (y - cy);
        v += ((k & 1) ? 0.4f : -0.4f) * expf(-d2 / (2.f * s * s));
      }
      v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
      img[y * w + x] = (uint8_t)(255.f * v + 0.5f);
    }
}

int main(void)
{
  const int w = 640, h = 480;
  uint8_t *a = malloc((size_t)w * h), *b = malloc((size_t)w * h);
  blobs(a, w, h, 0);
  blobs(b, w, h, 6);
  if (vksift_loadVulkan() != VKSIFT_SUCCESS)
    return 1;
  vksift_Config config = vksift_getDefaultConfig();
  config.input_image_max_size = (uint32_t)(w * h);
  vksift_Instance inst = NULL;
  if (vksift_createInstance(&inst, &config) != VKSIFT_SUCCESS)
    return 1;
  vksift_detectFeatures(inst, a, w, h, 0u);
  vksift_detectFeatures(inst, b, w, h, 1u);
  const uint32_t na = vksift_getFeaturesNumber(inst, 0u), nb = vksift_getFeaturesNumber(inst, 1u);
  vksift_matchFeatures(inst, 0u, 1u);
  const uint32_t nm = vksift_getMatchesNumber(inst);
  vksift_Match_2NN *m = malloc(sizeof(*m) * (nm ? nm : 1));
  vksift_downloadMatches(inst, m);
  uint32_t good = 0;
  for (uint32_t i = 0; i < nm; i++)
    good += (m[i].dist_a_b1 / m[i].dist_a_b2) < 0.75f;
  printf("features: %u / %u, matches passing the ratio test: %u of %u\n", na, nb, good, nm);
  free(m);
  vksift_destroyInstance(&inst);
  vksift_unloadVulkan();
  free(a);
  free(b);
  return 0;
}
