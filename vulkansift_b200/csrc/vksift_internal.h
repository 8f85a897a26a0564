/*
 * vksift_internal.h -- shared declarations of the B200 SIFT library.
 *
 * Layering (replaces reference L1-L4):
 *   api.cu        C ABI, validation, stream/event orchestration   (vulkansift.c)
 *   plan.cu       host tables: octaves, taps, sections            (sift_memory.c:15-87, sift_detector.c:52-145)
 *   pyramid.cu    seed + separable blur + DoG + next-octave seed  (GaussianBlur*.comp, DifferenceOfGaussian.comp, blits)
 *   extrema.cu    3x3x3 scan, refinement, deterministic ordering  (ExtractKeypoints.comp)
 *   describe.cu   orientation, assembly, descriptor               (ComputeOrientation.comp, ComputeDescriptors.comp)
 *   match.cu      2-NN brute force on tcgen05 tensor cores        (Get2NearestNeighbors.comp)
 */
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "vksift_arith.h"
#include "vksift_b200_ext.h"
#include "vulkansift/vulkansift.h"

#define VKS_MAX_OCT 16    /* log2(65536) - 4 + 1 */
#define VKS_MAX_LAYERS 24 /* nb_scales_per_octave + 3 <= 24 */
#define VKS_MAX_TAPS 21   /* VKSIFT_DETECTOR_MAX_GAUSSIAN_KERNEL_SIZE (20) + centre */
#define VKS_MAX_ORI 36    /* orientation histogram bins */

namespace vks
{

/* ---- logging (vkenv/logger.c) ------------------------------------------- */
void log_msg(int level, const char *tag, const char *fmt, ...);
#define LOGE(tag, ...) ::vks::log_msg(VKSIFT_LOG_ERROR, tag, __VA_ARGS__)
#define LOGW(tag, ...) ::vks::log_msg(VKSIFT_LOG_WARNING, tag, __VA_ARGS__)
#define LOGI(tag, ...) ::vks::log_msg(VKSIFT_LOG_INFO, tag, __VA_ARGS__)
#define LOGD(tag, ...) ::vks::log_msg(VKSIFT_LOG_DEBUG, tag, __VA_ARGS__)

/* ---- device-visible plain structs --------------------------------------- */

/* the 36 bytes of vksift_Feature in front of `descriptor` */
struct FeatHead
{
  float x, y, scale_x, scale_y;
  uint32_t scale_idx;
  int32_t octave_idx;
  float sigma, orientation, intensity;
};
static_assert(sizeof(FeatHead) == 36, "FeatHead layout");

/* one octave of the scale space in HBM: (ns+3) Gaussian and (ns+2) DoG layers of fp32 values, or of binary16 values with
 * VKSIFT_PYRAMID_PRECISION_FLOAT16 (the reference allocates VK_FORMAT_R16_SFLOAT images then, sift_memory.c:139); rows padded to
 * `pitch` elements (a multiple of 128 bytes).  Kernels read and write layers through layer_ld / layer_st (layer_io.cuh). */
struct OctaveView
{
  void *G;
  void *D;
  int w, h, pitch;
  int layer_stride; /* elements between layers = pitch*h */
  int fp16;         /* element type: 0 float, 1 __half */
};

struct DetectParams
{
  OctaveView oct[VKS_MAX_OCT];
  uint32_t cap[VKS_MAX_OCT];     /* section capacity (updateBufferInfo) */
  uint32_t sec_off[VKS_MAX_OCT]; /* prefix sum of cap */
  int n_oct, ns, upsample;
  int ob, oe; /* octave range [ob, oe) a launch of the extrema / ordering / orientation kernels works on */
  float sigma0, thr, prefilter, edge_limit;
  uint32_t max_ori;    /* 0 = unlimited */
  uint32_t ori_stride; /* orientation slots per primary */
  int vlfeat;
  uint32_t max_feats;
  /* ordered compaction of the keypoints (extrema.cu): one bit per (s, y, x) with s in [1, ns], rows of bm_rw[o] 32-bit words;
   * bit (s, y, x) of octave o = bit (x & 31) of word bm_off[o] + ((s-1)*h + y) * bm_rw[o] + (x >> 5) */
  uint32_t *acc_bm;   /* keypoints accepted by the refinement (cleared before every detection) */
  uint32_t *row_cnt;  /* accepted keypoints per (s, y) row, turned into the row's first rank by row_scan_kernel */
  uint32_t bm_off[VKS_MAX_OCT], bm_rw[VKS_MAX_OCT], row_off[VKS_MAX_OCT];
  /* queue of strict extrema per octave: key (s << 40 | y << 20 | x), bit 63 = accepted; q_heads[i] = refined record of entry i */
  unsigned long long *raw_q;
  FeatHead *q_heads;
  uint32_t q_off[VKS_MAX_OCT], q_cap[VKS_MAX_OCT];
};

/* per-buffer device counters, zeroed at the start of every detection */
struct DetectCounters
{
  uint32_t n_raw[VKS_MAX_OCT];    /* strict extrema found = queue fill (keeps counting past the queue capacity) */
  uint32_t n_cand[VKS_MAX_OCT];   /* accepted keypoints found (may exceed capacity) */
  uint32_t n_prim[VKS_MAX_OCT];   /* primaries kept = min(n_cand, cap) */
  uint32_t n_found[VKS_MAX_OCT];  /* primaries + extra orientations found */
  uint32_t n_kept[VKS_MAX_OCT];   /* min(n_found, cap) */
  uint32_t out_off[VKS_MAX_OCT];  /* packed output offset per octave */
  uint32_t prim_off[VKS_MAX_OCT]; /* offset of the octave's primaries in the work list */
  uint32_t n_prim_total;
  uint32_t n_total; /* packed feature count */
  uint32_t desc_next; /* work counter of the descriptor kernel (features are handed out one at a time) */
};

/* ---- separable blur work description (pyramid.cu) ----------------------- */
enum BlurSrcKind
{
  BLUR_SRC_LAYER = 0,  /* float layer of the same octave */
  BLUR_SRC_U8_UP2 = 1, /* u8 input, 2x LINEAR blit on the fly (octave 0, upsampling) */
  BLUR_SRC_U8 = 2      /* u8 input converted 1:1 */
};

struct BlurPass
{
  const void *src; /* layer (float or __half elements, see fp16), or (u8 kinds) the address of a device slot holding the image
                      pointer: the launch sequence is captured in a CUDA graph once and replayed for every image */
  void *dst_g;     /* Gaussian layer written */
  void *dst_d;     /* DoG layer (dst_g - src) or NULL */
  void *dst_next;  /* next octave layer 0 (NEAREST blit of this layer) or NULL */
  int w, h;        /* layer size */
  int src_pitch, dst_pitch, next_pitch;
  int src_w, src_h; /* u8 input size for the seed pass */
  int next_w, next_h;
  int src_kind;
  int fp16; /* VKSIFT_PYRAMID_PRECISION_FLOAT16: layers hold binary16 elements; arithmetic stays fp32, a value is rounded to nearest
               even when it is stored (SURVEY B-D11) */
  int radius;
  int tiles_x, tiles_y, tile_begin; /* CTA range of this pass inside the launch */
  int tile_h;                       /* output rows per tile chosen for this pass */
  float taps[VKS_MAX_TAPS];
  alignas(64) CUtensorMap tmap; /* source layer, box = tile + halo (fast kernel, float sources) */
};
#define VKS_MAX_PASSES_PER_STEP 4
struct BlurStep
{
  BlurPass pass[VKS_MAX_PASSES_PER_STEP];
  int n_pass;
  int n_tiles;
};

/* up to FZ_MAXL consecutive layers of one (small) octave produced by one launch of the fused kernel */
#define FZ_MAXL 3
struct FusedLaunch
{
  const void *src;  /* layer s_begin-1 of the octave */
  void *g0;         /* Gaussian layer s_begin */
  void *d0;         /* DoG layer s_begin-1 */
  int layer_stride; /* elements between consecutive layers */
  int w, h, pitch;
  int n_layers;
  int fp16;
  int radius[FZ_MAXL]; /* rounded up to even, taps zero padded */
  int next_k;          /* layer (0-based inside the launch) whose NEAREST decimation seeds the next octave, or -1 */
  void *dst_next;
  int next_w, next_h, next_pitch;
  int tiles_x;
  float2 taps2[FZ_MAXL][14];
};

/* a chain of consecutive layers of one large octave produced by one launch of the streaming strip kernel (pyramid_strip.cu) */
struct StripLaunch
{
  int kind;        /* which instantiation: 0 = three layers with radii (4, 6, 8), 1 = two layers with radii (10, 12) */
  int grid;        /* strips x segments */
  int first_layer; /* layer index of the chain's first layer inside the octave (trace label) */
  int n_layers;
  alignas(16) unsigned char params[512]; /* StripParams */
};

/* ---- host-side plan ------------------------------------------------------ */
struct ScalePlan
{
  int ns;
  uint32_t ksize[VKS_MAX_LAYERS];
  uint32_t radius[VKS_MAX_LAYERS];
  float taps[VKS_MAX_LAYERS][VKS_MAX_TAPS];
};
/* sift_detector.c:52-145 */
void plan_gaussian_taps(ScalePlan *sp, const vksift_Config *cfg);
/* sift_memory.c:655-660 */
uint32_t plan_max_octaves(const vksift_Config *cfg, uint32_t *side_out);
/* sift_memory.c:15-38 */
uint32_t plan_octaves(uint32_t w, uint32_t h, bool upsample, uint32_t max_octaves, uint32_t *ow, uint32_t *oh);
/* sift_memory.c:40-87 */
void plan_sections(uint32_t max_feats, uint32_t n_oct, uint32_t *cap);

/* ---- kernels' host launchers -------------------------------------------- */
struct FeatureBuffer;
struct Instance;

/* fills tiles_x/tiles_y/tile_begin/n_tiles of a step for the kernel that will run it */
bool blur_step_tiles(BlurStep *step);
/* true when the pass runs on the unrolled packed-fp32 kernel, false for the compact kernel */
bool blur_pass_is_fast(const BlurPass &bp);
/* tile geometry + TMA tensor map of a pass for the fast kernel */
bool blur_pass_prepare_fast(BlurPass *bp);
/* the input image as fp32 at the resolution of octave 0 (UNORM conversion, LINEAR 2x blit when `up`): source of the first blur */
cudaError_t launch_expand_input(const void *src_slot, int sw, int sh, int up, float *dst, int pitch, int w, int h, cudaStream_t st);
cudaError_t launch_blur_pass_fast(const BlurPass &bp, cudaStream_t st);
cudaError_t launch_blur_step(const BlurStep &step, cudaStream_t st);
cudaError_t launch_widen_layer(const void *src, int w, int h, int pitch, float *dst, cudaStream_t st);
/* groups consecutive layer passes of one octave into fused launches; false when a pass cannot be fused */
bool fused_plan_octave(const BlurPass *passes, int n_pass, std::vector<FusedLaunch> *out);
cudaError_t launch_fused(const FusedLaunch &F, cudaStream_t st);
/* plan a chain of passes (consecutive layers of one octave, float source) for the strip kernel; false when the chain or the
 * octave size is not one the kernel is built for (the caller keeps the per-layer launches) */
bool strip_plan(const BlurPass *passes, int n, StripLaunch *out);
cudaError_t launch_strip(const StripLaunch &L, cudaStream_t st);
struct ExtremaPlan; /* TMA tensor maps over the DoG layers of the current pyramid */
cudaError_t extrema_plan_build(const DetectParams &P, ExtremaPlan **plan_io);
void extrema_plan_destroy(ExtremaPlan *pl);
/* scan + refinement + ordered compaction of the octaves [P.ob, P.oe): prim[sec_off[o] + rank] for rank < cap[o] */
/* scan_ctas: persistent CTAs of the extrema scan (0 = two per SM, the choice for one detection at a time) */
cudaError_t launch_extrema(const DetectParams &P, const ExtremaPlan *pl, DetectCounters *cnt, FeatHead *prim, cudaStream_t st, uint64_t *launch_count,
                           int scan_ctas);
/* words of each bitmap and rows of row_cnt the current pyramid needs; fills bm_off / bm_rw / row_off of P */
/* also: the octaves' shares (q_off / q_cap) of an extrema queue of queue_total entries */
void extrema_layout(DetectParams *P, size_t *bm_words, size_t *rows, size_t *queue_entries, uint32_t queue_total);
bool extrema_scales_supported(int ns);
cudaError_t launch_orientation(const DetectParams &P, DetectCounters *cnt, const FeatHead *prim, float *ori, uint32_t *n_ori, cudaStream_t st,
                               int ctas_per_sm);
cudaError_t launch_assemble(const DetectParams &P, DetectCounters *cnt, const uint32_t *n_ori, uint32_t *feat_src, uint32_t *host_counts,
                            cudaStream_t st);
/* table of the descriptor's fixed-point scale sums M(R/2) (ComputeDescriptors.comp:116-124 only depends on the window radius) */
#define VKS_DESC_M_TABLE 128
cudaError_t launch_descriptors(const DetectParams &P, DetectCounters *cnt, const float *m_table, const FeatHead *prim, const float *ori, const uint32_t *feat_src,
                               FeatHead *out_heads, uint8_t *out_desc, cudaStream_t st);
/* AoS <-> SoA for host transfers */
cudaError_t launch_pack_aos(const FeatHead *heads, const uint8_t *desc, uint32_t n, uint8_t *aos, cudaStream_t st);
cudaError_t launch_unpack_aos(const uint8_t *aos, uint32_t n, FeatHead *heads, uint8_t *desc, cudaStream_t st);

/* matcher */
struct MatchWorkspace;
cudaError_t match_workspace_create(MatchWorkspace **ws, uint32_t max_feats);
void match_workspace_destroy(MatchWorkspace *ws);
/* |x|^2 of n descriptors in both forms the matcher uses: plain (A side) and packed nbk (B side, padded to a multiple of 128 rows) */
cudaError_t launch_norms(const uint8_t *desc, uint32_t n, uint32_t *out_plain, uint32_t *out_packed, cudaStream_t st);
/* norm_a / norm_b: cached norms of the two sides (launch_norms) or nullptr, in which case they are computed into the workspace.
 * inputs_settled: descriptors and norms were NOT written by the work enqueued on `st` right before this call (cached norms of
 * unchanged buffers); the search may then overlap the tail of a previous search on the same stream. */
cudaError_t launch_match(MatchWorkspace *ws, int impl, const uint8_t *desc_a, uint32_t na, const uint32_t *norm_a, const uint8_t *desc_b, uint32_t nb,
                         const uint32_t *norm_b, vksift_Match_2NN *out, cudaStream_t st, cudaEvent_t ev_after_prepare, bool inputs_settled,
                         uint64_t *launch_count);

/* B-side norms of up to VKS_MAX_MATCH_BLOCKS descriptor blocks in one launch (all-pairs step) */
#define VKS_MAX_MATCH_BLOCKS 64
struct MatchBlockCounts
{
  uint32_t n[VKS_MAX_MATCH_BLOCKS];
};
cudaError_t launch_norms_blocks(const uint8_t *desc, const MatchBlockCounts &counts, uint32_t n_blocks, uint32_t stride_rows, uint32_t *out_packed,
                                cudaStream_t st);

/* ONE search of A against n_groups of the blocks (group g = block blk[g] with cnt[g] rows, blocks at a common stride of
 * stride_rows rows, packed norms laid out the same way; result list of block j at out + j * out_stride).  Tensor-core path only.
 * Contract: in every searched block the rows between its count and the largest count rounded up to 128 are ZERO. */
cudaError_t launch_match_blocks(MatchWorkspace *ws, const uint8_t *desc_a, uint32_t na, const uint32_t *norm_a, const uint8_t *blocks,
                                const uint32_t *norm_blocks, uint32_t stride_rows, const uint32_t *blk, const uint32_t *cnt, uint32_t n_groups,
                                vksift_Match_2NN *out, uint32_t out_stride, cudaStream_t st, bool inputs_settled, uint64_t *launch_count);

cudaError_t launch_match_filter(const vksift_Match_2NN *m12, uint32_t na, const vksift_Match_2NN *m21, uint32_t nb, float ratio, uint32_t *pairs,
                                uint32_t capacity, uint32_t *count, cudaStream_t st);

/* descriptor all-gather over NVLink peer memory (exchange.cu) */
#define VKS_MAX_PEERS 16
struct PeerExchange;
cudaError_t exchange_create(PeerExchange **out, int rank, int world, uint32_t slot_rows, void *handle64);
cudaError_t exchange_connect(PeerExchange *x, const void *handles);
/* push + wait; the rows of every received block between its count and the largest count rounded up to 128 are zeroed
 * (launch_match_blocks relies on it) */
cudaError_t exchange_allgather(PeerExchange *x, const uint8_t *desc, uint32_t n, cudaStream_t st, uint64_t *launch_count);
const uint32_t *exchange_host_counts(const PeerExchange *x);
uint32_t exchange_timeout_mask(const PeerExchange *x);
const uint8_t *exchange_blocks(const PeerExchange *x, uint64_t *stride_bytes);
int exchange_world(const PeerExchange *x);
uint32_t exchange_slot_rows(const PeerExchange *x);
int exchange_rank(const PeerExchange *x);
void exchange_destroy(PeerExchange *x);

} // namespace vks
