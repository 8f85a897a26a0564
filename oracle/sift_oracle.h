/*
 * sift_oracle.h -- CPU oracle of the vksift detect + 2-NN match path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.
 *
 * PARITY PINNING: **pinned against the reference's own code run here.**  The reference ships no tests,
 * golden vectors or fixtures for this path and its build needs the Vulkan SDK, glslc and a Vulkan device
 * (all absent), so oracle/build_ref.py compiles the reference's seven GLSL shaders (read in place from
 * /root/reference, syntactic rewrites only, on top of oracle/glsl_emu.h) and its three host functions
 * (setupGaussianKernels, updateScaleSpaceInfo, updateBufferInfo) into oracle/_ref/libvksift_ref.so and runs
 * them on the CPU.  tests/golden/ref_vectors.npz holds vectors generated from that build
 * (tests/golden/make_golden.py); tests/test_ref_parity.py requires this oracle to reproduce them: host tables,
 * DoG, keypoints, orientations, descriptors and matches bit-exactly, blur layers to 1e-6 absolute (the
 * reference samples through normalized texture coordinates).  Fixed-function Vulkan steps that are not in the
 * reference sources (UNORM conversion, LINEAR / NEAREST blits) follow the Vulkan specification and remain
 * unpinned; OpenCV SIFT is the statistical cross-check for the whole pipeline (tests/test_oracle.py).
 */
#ifndef SIFT_ORACLE_H
#define SIFT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#define VKSO_MAX_KERNEL 20 /* VKSIFT_DETECTOR_MAX_GAUSSIAN_KERNEL_SIZE, sift_detector.h */
#define VKSO_MAX_OCTAVES 32

  /* same layout as vksift_Feature (include/vulkansift/vulkansift_types.h) */
  typedef struct
  {
    float x, y, scale_x, scale_y;
    uint32_t scale_idx;
    int32_t octave_idx;
    float sigma, orientation, intensity;
    uint8_t descriptor[128];
  } vkso_Feature;

  typedef struct
  {
    uint32_t idx_a, idx_b1, idx_b2;
    float dist_a_b1, dist_a_b2;
  } vkso_Match;

  /* plain-int mirror of the vksift_Config fields the path reads */
  typedef struct
  {
    uint32_t input_image_max_size;
    uint32_t max_nb_sift_per_buffer;
    int32_t use_input_upsampling;
    int32_t nb_octaves;
    int32_t nb_scales_per_octave;
    float input_image_blur_level;
    float seed_scale_sigma;
    float intensity_threshold;
    float edge_threshold;
    uint32_t max_nb_orientation_per_keypoint;
    int32_t use_vlfeat_format;
    int32_t use_interpolated_blur;
    int32_t use_fp16_pyramid;
    int32_t nb_threads; /* 0 = all (OpenMP); arithmetic per element is unaffected */
  } vkso_Config;

  typedef struct vkso_Context vkso_Context;

  void vkso_default_config(vkso_Config *cfg);
  vkso_Context *vkso_create(const vkso_Config *cfg);
  void vkso_destroy(vkso_Context *ctx);

  /* host-side tables */
  int vkso_max_octaves(const vkso_Context *ctx);
  /* raw table exactly as the reference pushes it to the shaders: ksize[s], k[s][20] */
  void vkso_kernel_table(const vkso_Context *ctx, uint32_t *ksize, float *k);
  /* effective symmetric taps actually convolved: ntaps[s] (= radius), e[s][21] */
  void vkso_effective_taps(const vkso_Context *ctx, uint32_t *radius, float *e);

  /* full detection; returns the number of features kept (sections clamped) */
  uint32_t vkso_detect(vkso_Context *ctx, const uint8_t *image, uint32_t width, uint32_t height);
  int vkso_nb_octaves(const vkso_Context *ctx);
  void vkso_octave_resolution(const vkso_Context *ctx, int octave, uint32_t *w, uint32_t *h);
  void vkso_section_capacity(const vkso_Context *ctx, uint32_t *caps);
  /* per octave: features found (may exceed capacity) and kept */
  void vkso_section_counts(const vkso_Context *ctx, uint32_t *found, uint32_t *kept);
  /* per octave: number of primary keypoints kept (before extra orientations) */
  void vkso_primary_counts(const vkso_Context *ctx, uint32_t *primaries);
  void vkso_get_features(const vkso_Context *ctx, vkso_Feature *out);
  const float *vkso_gaussian_layer(const vkso_Context *ctx, int octave, int scale);
  const float *vkso_dog_layer(const vkso_Context *ctx, int octave, int scale);
  /* stage timings of the last vkso_detect, seconds: pyramid+DoG, extrema, orientation, descriptor */
  void vkso_stage_seconds(const vkso_Context *ctx, double *t4);

  /* 2-NN brute force on 128-byte descriptors (Get2NearestNeighbors.comp).  nb >= 2. */
  void vkso_match(const uint8_t *desc_a, uint32_t na, const uint8_t *desc_b, uint32_t nb, vkso_Match *out, int nb_threads);
  void vkso_match_features(const vkso_Feature *a, uint32_t na, const vkso_Feature *b, uint32_t nb, vkso_Match *out, int nb_threads);

  /* the Vulkan fixed-function steps (u8 UNORM upload + LINEAR blit, NEAREST blit), sift_detector.c:860-916, 1003-1034 */
  void vkso_seed_image(const uint8_t *img, int sw, int sh, float *dst, int dw, int dh);
  void vkso_downsample_nearest(const float *src, int sw, int sh, float *dst, int dw, int dh);

  /* arithmetic probes for tests (include/vksift_arith.h) */
  float vkso_expf(float x);
  float vkso_exp2f(float x);
  float vkso_atan2f(float y, float x);
  void vkso_sincosf(float t, float *s, float *c);
  int vkso_ceil_log2(float m);
  int vkso_mirror(int i, int n);

#ifdef __cplusplus
}
#endif
#endif
