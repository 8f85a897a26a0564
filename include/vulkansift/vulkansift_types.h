/*
 * vulkansift_types.h -- POD types of the vksift_* C ABI, B200/CUDA build.
 *
 * Binary-compatible with the reference's public types so existing callers
 * relink unchanged (reference: include/vulkansift/vulkansift_types.h:15-162).
 * Sizes on x86-64/gcc, checked by tests/test_abi.py:
 *   vksift_Feature 164 B (align 4), vksift_Match_2NN 20 B, vksift_Config 88 B.
 */
#ifndef VKSIFT_TYPES_H
#define VKSIFT_TYPES_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

/* descriptor geometry: 4x4 spatial cells x 8 orientation bins = 128 bytes */
#define VKSIFT_FEATURE_NB_HIST 4
#define VKSIFT_FEATURE_NB_ORI 8

  /* device name slot filled by vksift_getAvailableGPUs() (reference :15) */
  typedef char VKSIFT_GPU_NAME[256];

  /* One SIFT feature as exchanged with the host (reference :17-31).
   * Device-side storage is struct-of-arrays; this AoS record only exists in
   * host transfers (vksift_downloadFeatures / vksift_uploadFeatures). */
  typedef struct
  {
    float x, y;             /* position in input-image pixels */
    float scale_x, scale_y; /* position in the octave image the keypoint was found in */
    uint32_t scale_idx;     /* Gaussian layer index inside the octave */
    int32_t octave_idx;     /* -1 is the 2x-upsampled octave */
    float sigma;            /* blur level, in input-image units */
    float orientation;      /* radians */
    float intensity;        /* interpolated DoG response */
    uint8_t descriptor[VKSIFT_FEATURE_NB_HIST * VKSIFT_FEATURE_NB_HIST * VKSIFT_FEATURE_NB_ORI];
  } vksift_Feature;

  /* Result row of the brute-force 2-nearest-neighbour search (reference :33-40).
   * Distances are true L2 norms (not squared) of the u8 descriptors. */
  typedef struct
  {
    uint32_t idx_a;
    uint32_t idx_b1;
    uint32_t idx_b2;
    float dist_a_b1;
    float dist_a_b2;
  } vksift_Match_2NN;

  typedef enum
  {
    VKSIFT_NO_LOG,
    VKSIFT_LOG_ERROR,
    VKSIFT_LOG_WARNING,
    VKSIFT_LOG_INFO,
    VKSIFT_LOG_DEBUG
  } vksift_LogLevel; /* reference :42-49 */

  typedef enum
  {
    VKSIFT_DESCRIPTOR_FORMAT_UBC,   /* Lowe / OpenCV / SiftGPU bin direction */
    VKSIFT_DESCRIPTOR_FORMAT_VLFEAT /* VLFeat / PopSift bin direction */
  } vksift_DescriptorFormat; /* reference :51-55 */

  typedef enum
  {
    VKSIFT_PYRAMID_PRECISION_FLOAT32,
    VKSIFT_PYRAMID_PRECISION_FLOAT16 /* layers stored as fp16, arithmetic stays fp32 */
  } vksift_PyramidPrecisionMode; /* reference :57-61 */

  typedef enum
  {
    VKSIFT_SUCCESS,
    /* Rejected before any side effect; the instance stays usable. */
    VKSIFT_INVALID_INPUT_ERROR,
    /* Device-side failure (here: a CUDA error).  The name is kept for ABI
     * compatibility; the instance must be destroyed afterwards. */
    VKSIFT_VULKAN_ERROR
  } vksift_Result; /* reference :63-74 */

  /* Window handles for the reference's frame-capture debug presenter
   * (reference :91-95).  Accepted and ignored: CUDA profilers need no frames. */
  typedef struct
  {
    void *context;
    void *window;
  } vksift_ExternalWindowInfo;

  /* Instance configuration (reference :97-162); field order and widths are ABI. */
  typedef struct
  {
    uint32_t input_image_max_size;   /* max width*height of an input image     (1920*1080) */
    uint32_t sift_buffer_count;      /* number of device feature buffers       (2) */
    uint32_t max_nb_sift_per_buffer; /* capacity of one feature buffer         (100000) */

    bool use_input_upsampling;    /* build octave 0 from a 2x upsampled input  (true) */
    uint8_t nb_octaves;           /* 0 = derive from the image size            (0) */
    uint8_t nb_scales_per_octave; /* DoG scales searched per octave            (3) */
    float input_image_blur_level; /* blur assumed in the input                 (0.5) */
    float seed_scale_sigma;       /* blur of the first pyramid layer           (1.6) */
    float intensity_threshold;    /* DoG contrast threshold, divided by nb_scales_per_octave (0.04) */
    float edge_threshold;         /* principal-curvature ratio limit           (10) */
    uint32_t max_nb_orientation_per_keypoint; /* 0 = unlimited                 (4) */
    vksift_DescriptorFormat descriptor_format; /*                              (UBC) */

    int32_t gpu_device_index;            /* CUDA ordinal, <0 = pick automatically (-1) */
    bool use_hardware_interpolated_blur; /* paired-tap blur table              (true) */
    vksift_PyramidPrecisionMode pyramid_precision_mode; /*                     (FLOAT32) */

    /* Called by every function that cannot return a vksift_Result when it
     * detects an error.  May throw (C++ callers).  Default aborts. */
    void (*on_error_callback_function)(vksift_Result);

    bool use_gpu_debug_functions; /* accepted, no effect                       (false) */
    vksift_ExternalWindowInfo gpu_debug_external_window_info;
  } vksift_Config;

#ifdef __cplusplus
}
#endif

#endif /* VKSIFT_TYPES_H */
