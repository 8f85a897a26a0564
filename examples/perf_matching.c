/* The reference's matching-quality measurement (src/perf/perf_matching.cpp:30-79,143-207 with matchFeatures of
 * src/perf/perf_common.cpp:109-170, do_crosscheck = false) as a plain-C program against this library.
 *
 * For an image pair related by a known homography H (image N = image 1 warped by H) it detects features on both, runs
 * vksift_matchFeatures(1 -> N), keeps the matches whose distance ratio is below LOWES_RATIO = 0.75 and reports the
 * reference's metrics: putative match ratio = kept / features(1), precision = inliers / kept (a kept match is an inlier when
 * the matched point lies within PIXEL_DIST_THRESHOLD = 2.5 px of H * point), matching score = inliers / features(1), and a
 * repeatability figure (share of image-1 keypoints with an image-N keypoint within 2.5 px of their projection; the reference
 * takes cv::evaluateFeatureDetector's overlap-based number instead, which needs OpenCV).  One line per pair goes to
 * matching_results_vulkansift.txt in the reference's "dataset;1;n;repeatability;putative;precision;score" format.
 *
 *   gcc -O2 examples/perf_matching.c -Iinclude -Lvulkansift_b200/lib -lvulkansift -Wl,-rpath,$PWD/vulkansift_b200/lib -lm -o perf_matching
 *   ./perf_matching                 synthetic pairs: a 1280x960 blob field and five translated + slightly scaled copies
 *   ./perf_matching DATASET_PATH    Oxford affine-covariant sets (bark, bikes, boat, ...; binary PGM only, so "boat")
 */
#define _POSIX_C_SOURCE 199309L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vulkansift/vulkansift.h>

#define LOWES_RATIO 0.75f
#define PIXEL_DIST_THRESHOLD 2.5f

static float lcg01(uint32_t *s)
{
  *s = *s * 1664525u + 1013904223u;
  return (float)(*s >> 8) / 16777216.f;
}

/* blob field sampled through the inverse of H: out(x, y) = field(Hinv * (x, y)) */
typedef struct
{
  float cx, cy, sig, amp;
} Blob;

static void render(uint8_t *img, int w, int h, const Blob *b, int nb, const float *Hinv)
{
  float *acc = malloc(sizeof(float) * (size_t)w * h);
  for (size_t i = 0; i < (size_t)w * h; i++)
    acc[i] = 0.5f;
  /* forward-map every blob centre with H = inverse(Hinv) of an affine map, then splat it (the maps used here are
   * similarity transforms, so a Gaussian blob stays a Gaussian blob with sigma scaled) */
  const float a = Hinv[0], c = Hinv[3], tx = Hinv[2], ty = Hinv[5], bb = Hinv[1], d = Hinv[4];
  const float det = a * d - bb * c;
  const float scale = 1.f / sqrtf(fabsf(det));
  for (int k = 0; k < nb; k++)
  {
    const float px = b[k].cx - tx, py = b[k].cy - ty;
    const float cx = (d * px - bb * py) / det, cy = (-c * px + a * py) / det;
    const float sig = b[k].sig * scale;
    const int r = (int)(4.f * sig);
    for (int y = (int)cy - r; y <= (int)cy + r; y++)
      for (int x = (int)cx - r; x <= (int)cx + r; x++)
        if (x >= 0 && x < w && y >= 0 && y < h)
          acc[(size_t)y * w + x] += b[k].amp * expf(-(((float)x - cx) * ((float)x - cx) + ((float)y - cy) * ((float)y - cy)) / (2.f * sig * sig));
  }
  for (size_t i = 0; i < (size_t)w * h; i++)
  {
    const float v = acc[i] < 0.f ? 0.f : (acc[i] > 1.f ? 1.f : acc[i]);
    img[i] = (uint8_t)(255.f * v + 0.5f);
  }
  free(acc);
}

static uint8_t *read_pgm(const char *path, int *w, int *h)
{
  FILE *f = fopen(path, "rb");
  if (!f)
    return NULL;
  char magic[3] = {0};
  int vals[3], n = 0;
  if (fscanf(f, "%2s", magic) != 1 || strcmp(magic, "P5") != 0)
  {
    fclose(f);
    return NULL;
  }
  while (n < 3)
  {
    int ch = fgetc(f);
    if (ch == '#')
      while (ch != '\n' && ch != EOF)
        ch = fgetc(f);
    else if (ch >= '0' && ch <= '9')
    {
      ungetc(ch, f);
      if (fscanf(f, "%d", &vals[n++]) != 1)
        break;
    }
    else if (ch == EOF)
      break;
  }
  fgetc(f);
  if (n < 3 || vals[2] > 255)
  {
    fclose(f);
    return NULL;
  }
  *w = vals[0];
  *h = vals[1];
  uint8_t *img = malloc((size_t)*w * *h);
  if (fread(img, 1, (size_t)*w * *h, f) != (size_t)*w * *h)
  {
    free(img);
    img = NULL;
  }
  fclose(f);
  return img;
}

static int read_homography(const char *path, float *H)
{
  FILE *f = fopen(path, "r");
  if (!f)
    return 0;
  int ok = 1;
  for (int i = 0; i < 9; i++)
    ok &= fscanf(f, "%f", &H[i]) == 1;
  fclose(f);
  return ok;
}

/* perf_matching.cpp:30-79 */
static void compute_metrics(const vksift_Feature *f1, uint32_t n1, const vksift_Feature *f2, uint32_t n2, const vksift_Match_2NN *m, uint32_t nm,
                            const float *H, float *repeat, float *putative, float *precision, float *score, uint32_t *kept_out, uint32_t *inl_out)
{
  uint32_t kept = 0, inliers = 0;
  for (uint32_t i = 0; i < nm; i++)
  {
    if (!((m[i].dist_a_b1 / m[i].dist_a_b2) < LOWES_RATIO)) /* perf_common.cpp:158 */
      continue;
    kept++;
    const vksift_Feature *a = &f1[m[i].idx_a], *b = &f2[m[i].idx_b1];
    const float winv = 1.f / ((H[6] * a->x) + (H[7] * a->y) + H[8]);
    const float gx = ((H[0] * a->x) + (H[1] * a->y) + H[2]) * winv, gy = ((H[3] * a->x) + (H[4] * a->y) + H[5]) * winv;
    if (sqrtf((b->x - gx) * (b->x - gx) + (b->y - gy) * (b->y - gy)) < PIXEL_DIST_THRESHOLD)
      inliers++;
  }
  uint32_t rep = 0;
  for (uint32_t i = 0; i < n1; i++)
  {
    const float winv = 1.f / ((H[6] * f1[i].x) + (H[7] * f1[i].y) + H[8]);
    const float gx = ((H[0] * f1[i].x) + (H[1] * f1[i].y) + H[2]) * winv, gy = ((H[3] * f1[i].x) + (H[4] * f1[i].y) + H[5]) * winv;
    for (uint32_t j = 0; j < n2; j++)
      if ((f2[j].x - gx) * (f2[j].x - gx) + (f2[j].y - gy) * (f2[j].y - gy) < PIXEL_DIST_THRESHOLD * PIXEL_DIST_THRESHOLD)
      {
        rep++;
        break;
      }
  }
  *repeat = n1 ? (float)rep / (float)n1 : 0.f;
  *putative = n1 ? (float)kept / (float)n1 : 0.f;
  *precision = kept ? (float)inliers / (float)kept : 0.f;
  *score = n1 ? (float)inliers / (float)n1 : 0.f;
  *kept_out = kept;
  *inl_out = inliers;
}

int main(int argc, char **argv)
{
  vksift_setLogLevel(VKSIFT_LOG_WARNING);
  if (vksift_loadVulkan() != VKSIFT_SUCCESS)
    return -1;
  vksift_Config config = vksift_getDefaultConfig();
  config.input_image_max_size = 1920u * 1080u;
  vksift_Instance inst = NULL;
  if (vksift_createInstance(&inst, &config) != VKSIFT_SUCCESS)
    return -1;
  FILE *res = fopen("matching_results_vulkansift.txt", "w");
  vksift_Feature *f1 = malloc(sizeof(vksift_Feature) * config.max_nb_sift_per_buffer);
  vksift_Feature *f2 = malloc(sizeof(vksift_Feature) * config.max_nb_sift_per_buffer);
  vksift_Match_2NN *m = malloc(sizeof(vksift_Match_2NN) * config.max_nb_sift_per_buffer);
  uint32_t total_kept = 0, total_matches = 0, total_inl = 0;

  for (int n = 2; n <= 6; n++)
  {
    int w = 1280, h = 960, w2 = 1280, h2 = 960;
    uint8_t *img1 = NULL, *img2 = NULL;
    float H[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const char *set = "synthetic";
    if (argc > 1)
    {
      char p[1024];
      set = "boat";
      snprintf(p, sizeof(p), "%s/boat/img1.pgm", argv[1]);
      img1 = read_pgm(p, &w, &h);
      snprintf(p, sizeof(p), "%s/boat/img%d.pgm", argv[1], n);
      img2 = read_pgm(p, &w2, &h2);
      snprintf(p, sizeof(p), "%s/boat/H1to%dp", argv[1], n);
      if (!img1 || !img2 || !read_homography(p, H))
      {
        fprintf(stderr, "Failed to read the boat set under %s\n", argv[1]);
        return 0;
      }
    }
    else
    {
      /* image N = image 1 under a similarity: scale 1 + 0.02 (n-1), translation (7 (n-1), 5 (n-1)) */
      Blob *b = malloc(sizeof(Blob) * 900);
      uint32_t s = 12345u;
      for (int k = 0; k < 900; k++)
      {
        b[k].cx = lcg01(&s) * (float)w;
        b[k].cy = lcg01(&s) * (float)h;
        b[k].sig = 2.f * powf(6.f, lcg01(&s));
        b[k].amp = (0.15f + 0.45f * lcg01(&s)) * ((s & 1u) ? 1.f : -1.f);
      }
      const float sc = 1.f + 0.02f * (float)(n - 1), tx = 7.f * (float)(n - 1), ty = 5.f * (float)(n - 1);
      H[0] = sc, H[4] = sc, H[2] = tx, H[5] = ty; /* p2 = sc * p1 + t */
      const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      const float Hinv[9] = {1.f / sc, 0, -tx / sc, 0, 1.f / sc, -ty / sc, 0, 0, 1}; /* p1 = (p2 - t) / sc */
      img1 = malloc((size_t)w * h);
      img2 = malloc((size_t)w * h);
      render(img1, w, h, b, 900, I);
      render(img2, w, h, b, 900, Hinv);
      free(b);
    }
    vksift_detectFeatures(inst, img1, (uint32_t)w, (uint32_t)h, 0u);
    vksift_detectFeatures(inst, img2, (uint32_t)w2, (uint32_t)h2, 1u);
    const uint32_t n1 = vksift_getFeaturesNumber(inst, 0u), n2 = vksift_getFeaturesNumber(inst, 1u);
    vksift_downloadFeatures(inst, f1, 0u);
    vksift_downloadFeatures(inst, f2, 1u);
    float rep = 0, put = 0, prec = 0, score = 0;
    uint32_t kept = 0, inl = 0, nm = 0;
    if (n1 > 0 && n2 >= 2)
    {
      vksift_matchFeatures(inst, 0u, 1u);
      nm = vksift_getMatchesNumber(inst);
      vksift_downloadMatches(inst, m);
      compute_metrics(f1, n1, f2, n2, m, nm, H, &rep, &put, &prec, &score, &kept, &inl);
    }
    printf("%s 1->%d: %u / %u features, repeatability %.3f, putative match ratio %.3f, precision %.3f, matching score %.3f\n", set, n, n1, n2, rep,
           put, prec, score);
    if (res)
      fprintf(res, "%s;%d;%d;%f;%f;%f;%f\n", set, 1, n + 1, rep, put, prec, score);
    total_kept += kept;
    total_matches += nm;
    total_inl += inl;
    free(img1);
    free(img2);
  }
  printf("matches: %u filtered of %u; inliers within 2.5 px of the known homography: %u\n", total_kept, total_matches, total_inl);
  if (res)
    fclose(res);
  free(f1);
  free(f2);
  free(m);
  vksift_destroyInstance(&inst);
  vksift_unloadVulkan();
  return 0;
}
