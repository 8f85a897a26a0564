/* The reference's runtime measurement (src/perf/perf_runtime.cpp:5-6,62-79 driving src/perf/wrappers/vulkansift_wrapper.cpp:5-33)
 * as a plain-C program against this library: 50 warm-up and 500 measured iterations of
 *     vksift_detectFeatures(buffer 0) -> vksift_getFeaturesNumber -> vksift_downloadFeatures
 * on one image, wall clock per iteration, mean written as "mean_ms;nb_features" to runtime_results_vulkansift.txt like the
 * reference does.  A second loop submits image i+1 into the other buffer before fetching the features of image i: the two
 * detections run on two lanes of the instance (include/vksift_b200_ext.h) and overlap on the GPU.
 *
 *   gcc -O2 examples/perf_runtime.c -Iinclude -Lvulkansift_b200/lib -lvulkansift -Wl,-rpath,$PWD/vulkansift_b200/lib -lm -o perf_runtime
 *   ./perf_runtime [image.pgm]        (binary P5 PGM, 8 bit; without an argument: a synthetic 1920x1080 blob field)
 */
#define _POSIX_C_SOURCE 199309L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <vulkansift/vulkansift.h>

#define NB_ITER_WARMUP 50
#define NB_ITER_MEAS 500

static double now_ms(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

static uint8_t *read_pgm(const char *path, int *w, int *h)
{
  FILE *f = fopen(path, "rb");
  if (!f)
    return NULL;
  int maxv = 0;
  char magic[3] = {0};
  if (fscanf(f, "%2s", magic) != 1 || strcmp(magic, "P5") != 0)
  {
    fclose(f);
    return NULL;
  }
  int vals[3], n = 0;
  while (n < 3)
  {
    int c = fgetc(f);
    if (c == '#')
      while (c != '\n' && c != EOF)
        c = fgetc(f);
    else if (c >= '0' && c <= '9')
    {
      ungetc(c, f);
      if (fscanf(f, "%d", &vals[n++]) != 1)
        break;
    }
    else if (c == EOF)
      break;
  }
  fgetc(f); /* the single whitespace after maxval */
  *w = vals[0];
  *h = vals[1];
  maxv = vals[2];
  if (n < 3 || maxv > 255 || *w <= 0 || *h <= 0)
  {
    fclose(f);
    return NULL;
  }
  uint8_t *img = malloc((size_t)*w * *h);
  if (fread(img, 1, (size_t)*w * *h, f) != (size_t)*w * *h)
  {
    free(img);
    img = NULL;
  }
  fclose(f);
  return img;
}

/* seeded field of Gaussian blobs (the shape of the benchmark's synthetic input, not the same generator) */
static uint8_t *blob_field(int w, int h, int n_blobs, uint32_t seed)
{
  float *acc = malloc(sizeof(float) * (size_t)w * h);
  for (size_t i = 0; i < (size_t)w * h; i++)
    acc[i] = 0.5f;
  uint32_t s = seed;
  for (int k = 0; k < n_blobs; k++)
  {
    s = s * 1664525u + 1013904223u;
    const float cx = (float)(s >> 8) / 16777216.f * (float)w;
    s = s * 1664525u + 1013904223u;
    const float cy = (float)(s >> 8) / 16777216.f * (float)h;
    s = s * 1664525u + 1013904223u;
    const float sig = 2.f * powf(6.f, (float)(s >> 8) / 16777216.f);
    s = s * 1664525u + 1013904223u;
    const float amp = (0.15f + 0.45f * (float)(s >> 8) / 16777216.f) * ((s & 1u) ? 1.f : -1.f);
    const int r = (int)(4.f * sig);
    for (int y = (int)cy - r; y <= (int)cy + r; y++)
      for (int x = (int)cx - r; x <= (int)cx + r; x++)
        if (x >= 0 && x < w && y >= 0 && y < h)
          acc[(size_t)y * w + x] += amp * expf(-(((float)x - cx) * ((float)x - cx) + ((float)y - cy) * ((float)y - cy)) / (2.f * sig * sig));
  }
  uint8_t *img = malloc((size_t)w * h);
  for (size_t i = 0; i < (size_t)w * h; i++)
  {
    const float v = acc[i] < 0.f ? 0.f : (acc[i] > 1.f ? 1.f : acc[i]);
    img[i] = (uint8_t)(255.f * v + 0.5f);
  }
  free(acc);
  return img;
}

int main(int argc, char **argv)
{
  int w = 1920, h = 1080;
  uint8_t *image = (argc > 1) ? read_pgm(argv[1], &w, &h) : blob_field(w, h, 2400, 42u);
  if (!image)
  {
    fprintf(stderr, "Failed to read image %s\n", argv[1]);
    return -1;
  }
  /* vulkansift_wrapper.cpp:5-19 */
  vksift_setLogLevel(VKSIFT_LOG_WARNING);
  if (vksift_loadVulkan() != VKSIFT_SUCCESS)
    return -1;
  vksift_Config config = vksift_getDefaultConfig();
  config.use_hardware_interpolated_blur = true;
  config.input_image_max_size = 1920u * 2u * 1080u * 2u;
  vksift_Instance inst = NULL;
  if (vksift_createInstance(&inst, &config) != VKSIFT_SUCCESS)
    return -1;
  vksift_Feature *feats = malloc(sizeof(vksift_Feature) * config.max_nb_sift_per_buffer);
  uint32_t n = 0;

  /* perf_runtime.cpp:62-79 with the wrapper's detectSIFT (vulkansift_wrapper.cpp:25-33) */
  for (int i = 0; i < NB_ITER_WARMUP; i++)
  {
    vksift_detectFeatures(inst, image, (uint32_t)w, (uint32_t)h, 0u);
    n = vksift_getFeaturesNumber(inst, 0u);
    vksift_downloadFeatures(inst, feats, 0u);
  }
  double sum = 0.;
  for (int i = 0; i < NB_ITER_MEAS; i++)
  {
    const double t0 = now_ms();
    vksift_detectFeatures(inst, image, (uint32_t)w, (uint32_t)h, 0u);
    n = vksift_getFeaturesNumber(inst, 0u);
    vksift_downloadFeatures(inst, feats, 0u);
    sum += now_ms() - t0;
  }
  const double mean_ms = sum / NB_ITER_MEAS;
  printf("serial (reference protocol): %dx%d, %u features, %.4f ms per image (upload -> detect -> download)\n", w, h, n, mean_ms);
  FILE *res = fopen("runtime_results_vulkansift.txt", "w");
  if (res)
  {
    fprintf(res, "%f;%u\n", mean_ms, n);
    fclose(res);
  }

  /* the same work with the two buffers of the default configuration used alternately: image i+1 is submitted before
   * the features of image i are fetched */
  for (int i = 0; i < 4; i++)
    vksift_detectFeatures(inst, image, (uint32_t)w, (uint32_t)h, (uint32_t)(i & 1));
  const double t0 = now_ms();
  vksift_detectFeatures(inst, image, (uint32_t)w, (uint32_t)h, 0u);
  for (int i = 1; i <= NB_ITER_MEAS; i++)
  {
    if (i < NB_ITER_MEAS)
      vksift_detectFeatures(inst, image, (uint32_t)w, (uint32_t)h, (uint32_t)(i & 1));
    const uint32_t b = (uint32_t)((i - 1) & 1);
    n = vksift_getFeaturesNumber(inst, b);
    vksift_downloadFeatures(inst, feats, b);
  }
  printf("two buffers alternately    : %dx%d, %u features, %.4f ms per image\n", w, h, n, (now_ms() - t0) / NB_ITER_MEAS);

  free(feats);
  vksift_destroyInstance(&inst);
  vksift_unloadVulkan();
  free(image);
  return 0;
}
