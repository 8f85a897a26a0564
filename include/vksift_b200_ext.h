/*
 * vksift_b200_ext.h -- vksiftx_* extension entry points of the B200 build.
 *
 * The reference API (include/vulkansift/vulkansift.h) only takes and returns
 * HOST memory and keeps one pipeline in flight (this build keeps one detection
 * in flight per detection lane, see vksiftx_getLaneCount).  These additions let a caller
 * keep inputs and results resident in HBM (device-timed benchmarks, NCCL
 * descriptor exchange between GPUs) and read per-stage device timings.  They
 * are plain C ABI like the rest: pointers and sizes, no CUDA or torch types.
 * Device pointers are passed as void* and are CUDA device addresses valid on
 * the instance's GPU.
 */
#ifndef VKSIFT_B200_EXT_H
#define VKSIFT_B200_EXT_H

#include "vulkansift/vulkansift.h"

#ifdef __cplusplus
extern "C"
{
#endif

  /* Library build tag, e.g. "vulkansift-b200 0.1 sm_100a". */
  VKSIFT_EXPORT const char *vksiftx_getVersionString();

  /* CUDA ordinal the instance runs on, and its main stream (a cudaStream_t). */
  VKSIFT_EXPORT int32_t vksiftx_getDeviceIndex(vksift_Instance instance);
  VKSIFT_EXPORT void *vksiftx_getStream(vksift_Instance instance);

  /* Same as vksift_detectFeatures (vulkansift.c:315-344) but the image is
   * already in device memory (row-major u8, width*height bytes, must stay valid
   * until the buffer is available again).  No host<->device copy is made. */
  VKSIFT_EXPORT void vksiftx_detectFeaturesDevice(vksift_Instance instance, const void *d_image, const uint32_t image_width,
                                                  const uint32_t image_height, const uint32_t gpu_buffer_id);

  /* Block until every pipeline of the instance has finished. */
  VKSIFT_EXPORT void vksiftx_waitIdle(vksift_Instance instance);

  /* Detection lanes.  The reference makes vksift_detectFeatures wait for the previous detection because the instance owns
   * one scale space and one command buffer (vulkansift.c:326-327).  This build gives the instance min(sift_buffer_count, 8)
   * lanes (environment override VKSIFT_LANES=n), each with its own scale space, scratch memory and streams; a detection
   * into buffer b runs on lane b % lanes and waits for that lane only, so detections into different buffers overlap on the
   * GPU.  Results are identical to the one-lane schedule.  A detection enqueued while other lanes hold detections the caller
   * has not waited for runs its latency-bound kernels (extrema scan, orientation) on small grids, which leaves the SMs to the
   * other detections' kernels (throughput); one that finds every lane idle uses the full grids (latency).
   * vksift_getScaleSpace* / vksift_download*Image show the scale space of the most recent detection.
   * vksiftx_joinLanes makes the instance stream (vksiftx_getStream) wait, on the device, for every detection enqueued so
   * far: an event recorded on that stream afterwards times them all. */
  VKSIFT_EXPORT uint32_t vksiftx_getLaneCount(vksift_Instance instance);
  VKSIFT_EXPORT void vksiftx_joinLanes(vksift_Instance instance);

  /* Packed device view of a feature buffer: n descriptors [n][128] u8 (row
   * pitch 128 B) and n 36-byte feature heads (the vksift_Feature fields before
   * `descriptor`).  Read-only (the matcher caches per-buffer norms).  Waits for the buffer; pointers stay valid until the next
   * detect/upload on it.  Any out pointer may be NULL. */
  VKSIFT_EXPORT void vksiftx_getBufferDeviceView(vksift_Instance instance, const uint32_t gpu_buffer_id, uint32_t *nb_feats, void **d_descriptors,
                                                 void **d_heads);

  /* Replace the content of a feature buffer by nb_feats descriptors read from
   * device memory ([nb_feats][128] u8); heads are zeroed.  This is the device
   * twin of vksift_uploadFeatures (vulkansift.c:398-415) and the landing point
   * of the NCCL descriptor all-gather. */
  VKSIFT_EXPORT void vksiftx_uploadDescriptorsDevice(vksift_Instance instance, const void *d_descriptors, const uint32_t nb_feats,
                                                     const uint32_t gpu_buffer_id);

  /* Copy the packed descriptors of a feature buffer into caller-owned device memory ([capacity][128] u8,
   * e.g. the send slot of an NCCL all-gather); rows past the feature count are zero-filled up to
   * `capacity`.  Returns the feature count.  Blocking like the other transfer functions. */
  VKSIFT_EXPORT uint32_t vksiftx_copyDescriptorsToDevice(vksift_Instance instance, const uint32_t gpu_buffer_id, void *d_dst, const uint32_t capacity);

  /* vksift_matchFeatures with the B side read in place from caller-owned device memory ([nb_feats_B][128] u8, 128-byte
   * aligned, e.g. a peer's block inside the receive buffer of the NCCL descriptor all-gather): no copy into a feature
   * buffer.  Same result records, tie rule, asynchrony and error behaviour as vksift_matchFeatures; the memory must stay
   * valid until the match has been downloaded. */
  VKSIFT_EXPORT void vksiftx_matchFeaturesAgainstDevice(vksift_Instance instance, const uint32_t gpu_buffer_id_A, const void *d_descriptors_B,
                                                        const uint32_t nb_feats_B);

  /* All-pairs step of cross-image matching (one image per GPU, descriptor blocks exchanged with an all-gather): 2-NN of
   * buffer A's features against EACH of n_blocks descriptor blocks read in place from caller-owned device memory (block j =
   * counts[j] rows of 128 bytes at d_blocks + j * block_stride_bytes, 128-byte aligned), all searches enqueued back to back
   * with no host round trip in between.  Blocks with fewer than 2 rows and block `skip_block` (the caller's own block; pass
   * 0xffffffff to skip none) are left out.  Asynchronous like vksift_matchFeatures.  vksiftx_downloadMatchesBlocks blocks
   * until the searches are done and writes n_blocks consecutive lists of vksift_getFeaturesNumber(A) records (list j =
   * matches against block j, same record semantics as vksift_downloadMatches; the lists of left-out blocks are zero-filled)
   * with ONE device-to-host copy. */
  VKSIFT_EXPORT void vksiftx_matchFeaturesAgainstBlocks(vksift_Instance instance, const uint32_t gpu_buffer_id_A, const void *d_blocks,
                                                        const uint32_t n_blocks, const uint64_t block_stride_bytes, const uint32_t *counts,
                                                        const uint32_t skip_block);
  VKSIFT_EXPORT void vksiftx_downloadMatchesBlocks(vksift_Instance instance, vksift_Match_2NN *matches, const uint32_t n_blocks);

  /* Descriptor exchange between the GPUs of one node over NVLink peer memory (one process and one instance per GPU).  The
   * reference has no multi-GPU path (one instance = one GPU, vulkansift.h:32-34); this is the exchange step of cross-image
   * matching: every rank pushes the descriptors of one feature buffer straight into a receive slot on every peer (one kernel,
   * posted NVLink stores, completion flags in peer memory) and then matches against the slots in place.
   *   vksiftx_exchangeCreate      allocates this rank's receive region (2 x world_size slots of slot_rows descriptors:
   *                               the most a rank may hold when it exchanges) and writes its 64-byte CUDA IPC handle to handle_out;
   *   vksiftx_exchangeConnect     maps the peers' regions: `handles` = world_size x 64 bytes, rank order, gathered by the
   *                               caller with whatever it has (MPI, torch.distributed, a file);
   *   vksiftx_exchangeAllGather   COLLECTIVE: pushes the descriptors of gpu_buffer_id to every peer, waits until every peer's
   *                               block of the same round has arrived, returns the row counts (counts[world_size]) and the
   *                               device address / stride of the received blocks (block r = rank r's descriptors; the own
   *                               block is not filled).  The blocks stay valid until the next-but-one exchange;
   *   vksiftx_exchangeMatchAllPeers  the same followed by vksiftx_matchFeaturesAgainstBlocks(own buffer, received blocks,
   *                               skip = own rank); fetch the result with vksiftx_downloadMatchesBlocks(world_size).
   * Every rank must make the same sequence of exchange calls; a peer that does not show up within two seconds is reported
   * through the error callback (VKSIFT_VULKAN_ERROR) instead of hanging the GPU.  All peers must have called Connect
   * before the first AllGather, and Destroy only after the last one has completed everywhere (barriers are the caller's). */
  VKSIFT_EXPORT bool vksiftx_exchangeCreate(vksift_Instance instance, const uint32_t rank, const uint32_t world_size, const uint32_t slot_rows,
                                            void *handle_out);
  VKSIFT_EXPORT bool vksiftx_exchangeConnect(vksift_Instance instance, const void *handles);
  VKSIFT_EXPORT bool vksiftx_exchangeAllGather(vksift_Instance instance, const uint32_t gpu_buffer_id, uint32_t *counts, void **d_blocks,
                                               uint64_t *block_stride_bytes);
  VKSIFT_EXPORT bool vksiftx_exchangeMatchAllPeers(vksift_Instance instance, const uint32_t gpu_buffer_id, uint32_t *counts);
  VKSIFT_EXPORT void vksiftx_exchangeDestroy(vksift_Instance instance);

  /* Device pointer of the last match result: vksift_getMatchesNumber() rows of vksift_Match_2NN. */
  VKSIFT_EXPORT void *vksiftx_getMatchesDevice(vksift_Instance instance);

  /* Stage timing with CUDA events on the instance stream.  When enabled, every
   * detect/match records events around its stages; read them back (ms) after
   * the pipeline finished.  Stage order for detection:
   *   0 pyramid+DoG, 1 extrema+refine+order, 2 orientation, 3 descriptor (+assemble), 4 whole pipeline
   * and for matching: 5 prepare (norms), 6 2-NN kernel, 7 whole pipeline. */
#define VKSIFTX_NB_STAGES 8
  VKSIFT_EXPORT void vksiftx_setProfiling(vksift_Instance instance, const bool enabled);
  VKSIFT_EXPORT void vksiftx_getStageTimesMs(vksift_Instance instance, float *times_ms);

  /* Launch trace of the scale-space stage of the last detection: with tracing enabled every launch of that stage is
   * bracketed by a CUDA event pair on its stream (this costs a few microseconds per launch, so traced detections are for
   * analysis, not for timing the pipeline).  vksiftx_getLaunchTrace waits for the detection, fills up to `capacity`
   * entries (name: 32 chars, start/end in microseconds from the start of the detection) and returns the launch count. */
  /* Cross-checked, ratio-tested matching on the device: what the reference's callers do on the CPU after two
   * vksift_matchFeatures / vksift_downloadMatches round trips (src/examples/test_sift_match.cpp:73-107,
   * src/perf/perf_common.cpp:122-170).  Runs A->B and B->A 2-NN searches, keeps the pairs (i, j) with j = nn1(i), i = nn1(j) and
   * dist1/dist2 < lowe_ratio in both directions, in increasing i.  Blocking; writes at most `capacity` pairs
   * (idx in A, idx in B) to host memory and returns the number of pairs found.  The instance's retained match result
   * (vksift_downloadMatches) is the A->B list afterwards. */
  VKSIFT_EXPORT uint32_t vksiftx_matchFeaturesCrossChecked(vksift_Instance instance, const uint32_t gpu_buffer_id_A, const uint32_t gpu_buffer_id_B,
                                                           const float lowe_ratio, uint32_t *pairs, const uint32_t capacity);

  VKSIFT_EXPORT void vksiftx_setLaunchTrace(vksift_Instance instance, const bool enabled);
  /* Analysis aid: run every launch of the detection on one stream, in dependency order.  With the default schedule the
   * octaves overlap on several streams and the time between a launch's two events includes the kernels it shares the GPU
   * with; in the serial schedule it is that kernel's own duration. */
  VKSIFT_EXPORT void vksiftx_setSerialSchedule(vksift_Instance instance, const bool enabled);
  VKSIFT_EXPORT uint32_t vksiftx_getLaunchTrace(vksift_Instance instance, char (*names)[32], float *start_us, float *end_us, const uint32_t capacity);

  /* Number of kernels this library launched on the instance since creation
   * (graph nodes count as launches). */
  VKSIFT_EXPORT uint64_t vksiftx_getKernelLaunchCount(vksift_Instance instance);

  /* Host tables of the instance, for parity tests against the oracle:
   * effective blur taps per scale (radius[s], taps[s][21]) -- sift_detector.c:52-145 --
   * and section capacities per octave -- sift_memory.c:40-87. */
  VKSIFT_EXPORT void vksiftx_getEffectiveTaps(vksift_Instance instance, uint32_t *radius, float *taps);
  VKSIFT_EXPORT void vksiftx_getSectionCapacities(vksift_Instance instance, const uint32_t gpu_buffer_id, uint32_t *caps);

  /* Matcher implementation switch for verification: 0 = tcgen05 tensor-core
   * kernel (default, the product path), 1 = SIMT dp4a cross-check kernel. */
  VKSIFT_EXPORT void vksiftx_setMatcherImpl(vksift_Instance instance, const int32_t impl);

#ifdef VKS_ANALYSIS
  /* ANALYSIS BUILD ONLY (vulkansift_b200/lib/libvulkansift_analysis.so, compiled with -DVKS_ANALYSIS; the product library
   * does not contain this entry point nor the code behind it).  Ablation timing: leave stages of the detection out to
   * measure what each one costs in the pipelined schedule; the results of detections made while a bit is set are INVALID.
   * Bits: 1 descriptors, 2 orientation, 4 extrema scan + refinement + ordering, 8 scale space. */
  VKSIFT_EXPORT void vksiftx_setDebugSkip(vksift_Instance instance, const int32_t mask);
#endif

#ifdef __cplusplus
}
#endif

#endif /* VKSIFT_B200_EXT_H */
