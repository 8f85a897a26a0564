/*
 * vulkansift.h -- the vksift_* C ABI, implemented on CUDA for NVIDIA B200 (sm_100a).
 *
 * Drop-in boundary: these are the 20 entry points a program built against the
 * reference library binds (reference: include/vulkansift/vulkansift.h:23-111).
 * Names, argument meaning, blocking behaviour and error reporting follow the
 * reference implementation (src/vulkansift/vulkansift.c); everything below the
 * boundary is new: CUDA streams/events instead of Vulkan queues/fences, and
 * hand-written sm_100a kernels instead of the GLSL compute shaders.
 *
 * Threading: like the reference, an instance is not thread-safe, and at most
 * one detection or matching pipeline is in flight per instance.
 */
#ifndef VULKAN_SIFT_H
#define VULKAN_SIFT_H

#include <stdbool.h>
#include <stdint.h>

#include "vulkansift/vulkansift_types.h"

#if defined(_WIN32) || defined(_WIN64)
#define VKSIFT_EXPORT __declspec(dllexport)
#else
#define VKSIFT_EXPORT __attribute__((__visibility__("default")))
#endif

#ifdef __cplusplus
extern "C"
{
#endif

  typedef struct vksift_Instance_T *vksift_Instance;

  /* ---- process-wide setup (reference vulkansift.h:23-31) ------------------
   * vksift_loadVulkan() initialises the CUDA driver and checks that an sm_100
   * device is visible; VKSIFT_VULKAN_ERROR if not.  Must precede
   * vksift_createInstance().  vksift_unloadVulkan() reverses it. */
  VKSIFT_EXPORT vksift_Result vksift_loadVulkan();
  VKSIFT_EXPORT void vksift_unloadVulkan();

  /* Two-call enumeration: gpu_names == NULL writes the device count to
   * *gpu_count; otherwise *gpu_count names are copied out. */
  VKSIFT_EXPORT void vksift_getAvailableGPUs(uint32_t *gpu_count, VKSIFT_GPU_NAME *gpu_names);
  VKSIFT_EXPORT void vksift_setLogLevel(const vksift_LogLevel level);

  /* ---- instance (reference vulkansift.h:33-39) ----------------------------
   * *instance_ptr must be NULL on entry; it is NULL again after any failure
   * and after vksift_destroyInstance().  One instance drives one GPU. */
  VKSIFT_EXPORT vksift_Result vksift_createInstance(vksift_Instance *instance_ptr, const vksift_Config *config);
  VKSIFT_EXPORT void vksift_destroyInstance(vksift_Instance *instance_ptr);
  VKSIFT_EXPORT vksift_Config vksift_getDefaultConfig();

  /* ---- pipelines (reference vulkansift.h:41-58) ---------------------------
   * Both calls return once the work is enqueued.  A pipeline already running
   * on the instance is waited for first.  image_data (row-major 8-bit
   * grayscale) is only read during the call. */
  VKSIFT_EXPORT void vksift_detectFeatures(vksift_Instance instance, const uint8_t *image_data, const uint32_t image_width, const uint32_t image_height,
                                           const uint32_t gpu_buffer_id);
  /* For every feature of buffer A: the two nearest descriptors of buffer B. */
  VKSIFT_EXPORT void vksift_matchFeatures(vksift_Instance instance, const uint32_t gpu_buffer_id_A, const uint32_t gpu_buffer_id_B);

  /* ---- transfers (reference vulkansift.h:60-92) ---------------------------
   * Blocking: each waits until the buffer it touches is idle. */
  VKSIFT_EXPORT uint32_t vksift_getFeaturesNumber(vksift_Instance instance, const uint32_t gpu_buffer_id);
  /* feats_ptr must hold vksift_getFeaturesNumber() records. */
  VKSIFT_EXPORT void vksift_downloadFeatures(vksift_Instance instance, vksift_Feature *feats_ptr, const uint32_t gpu_buffer_id);
  VKSIFT_EXPORT void vksift_uploadFeatures(vksift_Instance instance, const vksift_Feature *feats_ptr, const uint32_t nb_feats,
                                           const uint32_t gpu_buffer_id);
  /* Number of rows of the last match = feature count of its buffer A; never blocks. */
  VKSIFT_EXPORT uint32_t vksift_getMatchesNumber(vksift_Instance instance);
  /* matches must hold vksift_getMatchesNumber() records. */
  VKSIFT_EXPORT void vksift_downloadMatches(vksift_Instance instance, vksift_Match_2NN *matches);
  /* Non-blocking poll: false while a running pipeline reads or writes the buffer. */
  VKSIFT_EXPORT bool vksift_isBufferAvailable(vksift_Instance instance, const uint32_t gpu_buffer_id);

  /* ---- scale-space inspection (reference vulkansift.h:94-100) ------------- */
  VKSIFT_EXPORT uint8_t vksift_getScaleSpaceNbOctaves(vksift_Instance instance);
  VKSIFT_EXPORT void vksift_getScaleSpaceOctaveResolution(vksift_Instance instance, const uint8_t octave, uint32_t *octave_images_width,
                                                          uint32_t *octave_images_height);
  /* scale in [0, nb_scales_per_octave+3) */
  VKSIFT_EXPORT void vksift_downloadScaleSpaceImage(vksift_Instance instance, const uint8_t octave, const uint8_t scale, float *blurred_image);
  /* scale in [0, nb_scales_per_octave+2) */
  VKSIFT_EXPORT void vksift_downloadDoGImage(vksift_Instance instance, const uint8_t octave, const uint8_t scale, float *dog_image);

  /* ---- debug (reference vulkansift.h:102-111) -----------------------------
   * The reference presents an empty frame so graphics debuggers can delimit a
   * capture.  ncu needs none: logs a warning and returns. */
  VKSIFT_EXPORT void vksift_presentDebugFrame(vksift_Instance instance);

#ifdef __cplusplus
}
#endif

#endif /* VULKAN_SIFT_H */
