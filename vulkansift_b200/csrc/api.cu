/*
 * api.cu -- the vksift_* C ABI on CUDA.
 *
 * Mirrors the behaviour of the reference's src/vulkansift/vulkansift.c (entry
 * points, validation, blocking rules, error callback) with CUDA plumbing:
 *   Vulkan instance/loader  -> CUDA runtime initialisation check
 *   general queue + fences  -> one stream + two events (detect / match)
 *   staging buffers         -> pinned host buffers, cudaMemcpyAsync
 *   pre-recorded cmd buffer -> a launch plan rebuilt when the resolution changes
 * Feature buffers are packed on the device (sections are applied as counts,
 * see describe.cu), so the reference's pack step (sift_memory.c:957-1047)
 * has nothing to move.
 */
#include "vksift_internal.h"

#include <nvtx3/nvToolsExt.h> /* header-only: ranges are no-ops unless a profiler injects its library */

#include <cassert>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace vks
{

/* ---- logger (vkenv/logger.c:55-84) --------------------------------------- */
static int g_log_level = VKSIFT_LOG_INFO;
static bool g_api_loaded = false;

void log_msg(int level, const char *tag, const char *fmt, ...)
{
  if (level > g_log_level || level == VKSIFT_NO_LOG)
    return;
  static const char *names[] = {"", "ERROR", "WARNING", "INFO", "DEBUG"};
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  fprintf(level <= VKSIFT_LOG_WARNING ? stderr : stdout, "[%s][%s] %s\n", names[level], tag, buf);
}

static const char TAG[] = "VulkanSift";

/* ---- instance state -------------------------------------------------------- */
struct FeatureBuffer
{
  FeatHead *heads = nullptr; /* packed [max] */
  uint8_t *desc = nullptr;   /* packed [max][128] */
  DetectCounters *cnt = nullptr;
  uint32_t *host_counts = nullptr; /* pinned+mapped: [0]=n, [1..16]=found, [17..32]=kept */
  uint32_t *host_counts_dev = nullptr;
  uint32_t cap[VKS_MAX_OCT] = {0};
  uint32_t n_oct = 0;
  uint32_t cur_w = 0, cur_h = 0;
  bool uploaded = false; /* content came from vksift_uploadFeatures (is_packed in the reference) */
  uint32_t n_uploaded = 0;
  /* |d|^2 of the descriptors in the two forms the matcher reads (A side plain, B side packed), computed by the first match
   * that uses the buffer and kept until its content changes: matching one image against many (all-pairs) or the same pair
   * repeatedly does not recompute them */
  uint32_t *norm_plain = nullptr, *norm_packed = nullptr;
  bool norms_valid = false;
  uint32_t norms_n = 0;
};

struct Pyramid
{
  uint32_t n_oct = 0;
  uint32_t w[VKS_MAX_OCT] = {0}, h[VKS_MAX_OCT] = {0}, pitch[VKS_MAX_OCT] = {0};
  void *G[VKS_MAX_OCT] = {nullptr}; /* fp32 elements, or binary16 with VKSIFT_PYRAMID_PRECISION_FLOAT16 (sift_memory.c:139) */
  void *D[VKS_MAX_OCT] = {nullptr};
  size_t alloc_bytes_g[VKS_MAX_OCT] = {0};
  size_t alloc_bytes_d[VKS_MAX_OCT] = {0};
  int es = 4; /* bytes per element */
};

enum
{
  EV_D0 = 0, /* detect start */
  EV_D1,     /* after pyramid (of the large octaves when the small ones finish on the side stream) */
  EV_D1B,    /* after the pyramid of the small octaves (side stream) */
  EV_D2,     /* after extrema+order */
  EV_D3,     /* after orientation */
  EV_D4,     /* after descriptors */
  EV_M0,
  EV_M1,
  EV_M2,
  EV_COUNT
};

} // namespace vks

using namespace vks;

struct vksift_Instance_T
{
  vksift_Config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_detect_done = nullptr, ev_match_done = nullptr;
  bool detect_pending = false, match_pending = false;
  uint32_t detect_buffer = 0, match_a = 0, match_b = 0;

  uint32_t max_image_size = 0;
  uint32_t max_octaves = 0;
  uint32_t cur_w = 0, cur_h = 0;
  ScalePlan scales;
  Pyramid pyr;
  ExtremaPlan *extrema_plan = nullptr;
  std::vector<std::vector<BlurPass>> fast_oct; /* octaves [0,k) on the fast kernel: one stream per octave, one launch per layer */
  std::vector<std::vector<StripLaunch>> strip_oct; /* per fast octave: the multi-layer strip launches that replace its per-layer launches
                                                      of layers >= 1 (empty: the octave keeps the per-layer launches) */
  bool use_strip = false; /* VKSIFT_STRIP=1: multi-layer strip kernel for the large octaves (bit-exact; measured slower than the per-layer launches) */
  std::vector<BlurStep> steps_side;            /* small octaves [k,n) that cannot be fused: compact kernel, wavefront steps on the side stream */
  struct FusedOct
  {
    BlurStep seed_step;              /* octave 0 only: the u8 -> layer 0 pass (compact kernel) */
    bool has_seed_step = false;
    std::vector<FusedLaunch> chain;  /* launches up to the one that seeds the next octave: the critical path */
    std::vector<FusedLaunch> rest;   /* remaining layers, second side stream */
  };
  std::vector<FusedOct> fused_oct;             /* small octaves [k,n): fused kernel, one or two launches per octave */
  bool serial = false;   /* vksiftx_setSerialSchedule */
#ifdef VKS_ANALYSIS
  int debug_skip = 0;    /* analysis build only (libvulkansift_analysis.so): stages left out for ablation timing, results invalid: 1 descriptors, 2 orientation, 4 extrema+order, 8 scale space */
#endif
  bool no_split = false; /* VKSIFT_NO_SPLIT=1: extrema/orientation of all octaves after the whole pyramid (debug) */
  cudaStream_t side_stream = nullptr, side2_stream = nullptr;
  cudaEvent_t ev_chain[VKS_MAX_OCT] = {nullptr}; /* chain launches of octave o enqueued on the side stream */
  cudaEvent_t ev_join2 = nullptr;
  cudaStream_t oct_stream[VKS_MAX_OCT] = {nullptr}; /* [0] unused: octave 0 runs on the main stream */
  cudaEvent_t ev_seed[VKS_MAX_OCT] = {nullptr};     /* layer 0 of octave o written (by the pass producing layer ns of octave o-1) */
  cudaEvent_t ev_oct_done[VKS_MAX_OCT] = {nullptr};
  cudaEvent_t ev_pyr[VKS_MAX_OCT] = {nullptr}; /* timing: scale space of octave o complete (profiling, split schedule) */
  int n_pyr_events = 0;
  cudaEvent_t ev_join = nullptr;

  uint8_t *h_image = nullptr; /* pinned */
  uint8_t *d_image = nullptr;
  const uint8_t **h_src_slot = nullptr; /* pinned: pointer of the image the next detection reads */
  const uint8_t **d_src_slot = nullptr; /* device copy, read by the seed pass */
  /* the detection pipeline of a buffer as a CUDA graph (the reference's pre-recorded command buffer, sift_detector.c:1369-1393):
   * ~25 launches and ~15 cross-stream event operations cost more CPU time than the GPU needs to run them */
  struct DetectGraph
  {
    cudaGraphExec_t exec = nullptr;
    bool prof = false;
    uint32_t uses = 0; /* the first use runs eagerly (lazy kernel attributes), the second one captures */
    uint64_t launches = 0;
  };
  std::vector<DetectGraph> graphs;
  bool use_graph = false;

  /* Detection lanes.  The reference owns one scale space and one command buffer, so a detection first waits for the
   * previous one (vulkansift.c:326-327).  With 180 GB of HBM a second and third scale space cost nothing: the instance the
   * caller holds (the primary, lane 0) owns `lanes.size() - 1` secondary instances, each with its own pyramid, scratch,
   * streams and staging, all aliasing the primary's feature buffers.  A detection into buffer b runs on lane b % lanes and
   * only waits for that lane, so detections into different buffers overlap on the GPU (the latency chain of the small
   * octaves of one image hides behind the large octaves of the next) and the launch work of one overlaps the execution of
   * the other.  Results are unchanged; only the blocking rule is relaxed. */
  std::vector<vksift_Instance> lanes; /* primary: [0] = this; secondaries: empty */
  vksift_Instance primary = nullptr;  /* secondaries: the owner */
  uint32_t last_lane = 0;             /* lane of the most recent detection: the scale space the download functions show */
  cudaEvent_t ev_h2d = nullptr;       /* image upload from caller-pinned memory complete */

  std::vector<FeatureBuffer> own_buffers; /* primary only */
  FeatureBuffer *buffers = nullptr;       /* own_buffers.data() of the primary */
  uint32_t n_buffers = 0;
  uint32_t *compact_mem = nullptr; /* [accepted bitmap | row counts] of the ordered compaction (extrema.cu), cleared per detection */
  size_t compact_alloc_words = 0, bm_words = 0, compact_rows = 0;
  unsigned long long *raw_q = nullptr; /* extrema queue, all octaves */
  FeatHead *q_heads = nullptr;
  size_t q_alloc = 0;
  uint32_t q_total = 0; /* queue entries shared by the octaves (4 x max_nb_sift_per_buffer, at least 16 k; VKSIFT_RAW_QUEUE overrides) */
  FeatHead *prim = nullptr;
  float *ori = nullptr;
  uint32_t *n_ori = nullptr;
  uint32_t *feat_src = nullptr;
  float *desc_m_table = nullptr; /* fixed-point scale sums of the descriptor by window radius */
  uint32_t ori_stride = 4;

  uint8_t *d_aos = nullptr; /* device AoS staging for feature up/download */
  MatchWorkspace *match_ws = nullptr;
  vksift_Match_2NN *d_matches = nullptr;
  vksift_Match_2NN *d_matches_rev = nullptr; /* B->A list of the cross-checked matcher */
  vksift_Match_2NN *d_matches_blocks = nullptr; /* [blocks_cap][max_nb_sift_per_buffer]: results of vksiftx_matchFeaturesAgainstBlocks */
  uint32_t blocks_cap = 0, blocks_n = 0, blocks_na = 0;
  uint32_t *d_block_norms = nullptr; /* packed B-side norms of all blocks of such a call, one launch */
  float *d_expanded = nullptr;       /* the input image as fp32 at octave-0 resolution (launch_expand_input), source of the first blur */
  size_t expanded_bytes = 0;
  bool use_expand = false;           /* octave 0 runs on the fast kernels with fp32 storage: expand once, then an ordinary layer launch */
  int ori_ctas = 4;                  /* resident CTAs per SM of the orientation kernel */
  int scan_ctas = 0;                 /* persistent CTAs of the extrema scan, 0 = two per SM */
  /* Set per detection: other lanes hold detections the caller has not waited for, i.e. the GPU is shared.  The latency-bound
   * kernels then run on small grids (scan_ctas, ori_ctas) and leave the SMs to the issue-bound kernels of the other detections;
   * a detection that has the GPU to itself uses the full grids.  Either mode produces the same results. */
  bool crowded = false;
  PeerExchange *exchange = nullptr;  /* NVLink peer-memory all-gather of descriptor blocks (vksiftx_exchange*) */
  size_t block_norms_cap = 0;
  uint32_t *d_pairs = nullptr;               /* [2*max + 1]: filtered pairs, count at the end */
  uint32_t nb_matches = 0;
  int matcher_impl = 0;

  bool profiling = false;
  /* VKSIFT_TRACE=1: an event pair around every launch of the scale-space stage, printed as a timeline (debug aid) */
  struct TraceMark
  {
    char name[32];
    cudaEvent_t e0, e1;
  };
  bool trace = false;
  bool trace_dump_stderr = false;
  std::vector<TraceMark> trace_marks;
  size_t trace_used = 0;
  cudaEvent_t ev[EV_COUNT] = {nullptr};
  bool ev_detect_valid = false, ev_match_valid = false;
  bool ev_d1b_valid = false;
  uint64_t launches = 0;
};

namespace
{

struct DeviceGuard
{
  int prev = -1;
  explicit DeviceGuard(int dev)
  {
    cudaGetDevice(&prev);
    if (prev != dev)
      cudaSetDevice(dev);
    else
      prev = -1;
  }
  ~DeviceGuard()
  {
    if (prev >= 0)
      cudaSetDevice(prev);
  }
};

#define CU_TRY(expr)                                                                                                                                 \
  do                                                                                                                                                 \
  {                                                                                                                                                  \
    cudaError_t e__ = (expr);                                                                                                                        \
    if (e__ != cudaSuccess)                                                                                                                          \
    {                                                                                                                                                \
      LOGE(TAG, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e__), __FILE__, __LINE__, #expr);                                                    \
      return false;                                                                                                                                  \
    }                                                                                                                                                \
  } while (0)

void default_error_callback(vksift_Result r)
{
  LOGD(TAG, r == VKSIFT_INVALID_INPUT_ERROR ? "Aborting after invalid input error..." : "Aborting after device error...");
  abort();
}

/* ---- validation (vulkansift.c:541-661) ------------------------------------ */
bool config_valid(const vksift_Config *c)
{
  bool ok = true;
  auto need = [&](bool cond, const char *msg) {
    if (!cond)
    {
      LOGE(TAG, "%s", msg);
      ok = false;
    }
  };
  const bool seed_ok = ((c->use_input_upsampling ? 2.f : 1.f) * c->input_image_blur_level) <= c->seed_scale_sigma;
  need(c->input_image_max_size >= 1024, "Invalid configuration: input image size must be greater than or equal to 1024");
  need(c->sift_buffer_count > 0, "Invalid configuration: number of SIFT buffers must be greater than zero");
  need(c->max_nb_sift_per_buffer > 0, "Invalid configuration: number of SIFT features per buffers must be greater than zero");
  need(c->nb_scales_per_octave > 0, "Invalid configuration: number of scales per octave must be greater than zero");
  need(c->nb_scales_per_octave + 3 <= VKS_MAX_LAYERS && extrema_scales_supported(c->nb_scales_per_octave),
       "Invalid configuration: too many scales per octave for this build (the extrema scan stages all DoG layers of a tile in shared memory: at most 9)");
  need(c->input_image_blur_level >= 0.f, "Invalid configuration: input image blur level cannot be negative");
  need(c->seed_scale_sigma >= 0, "Invalid configuration: seed scale blur level cannot be negative");
  need(seed_ok, "Invalid configuration: the input image blur level (2x if upscaling activated) must be less than the seed scale blur level");
  need(c->intensity_threshold >= 0.f, "Invalid configuration: the DoG intensity threshold cannot be negative");
  need(c->edge_threshold >= 0.f, "Invalid configuration: the DoG edge threshold cannot be negative");
  need(c->on_error_callback_function != NULL, "Invalid configuration: the error callback function must not be NULL");
  switch (c->pyramid_precision_mode)
  {
  case VKSIFT_PYRAMID_PRECISION_FLOAT32:
    break;
  case VKSIFT_PYRAMID_PRECISION_FLOAT16:
    break;
  default:
    need(false, "Invalid configuration: invalid scale-space pyramid format precision specified");
    break;
  }
  return ok;
}

bool buffer_idx_valid(vksift_Instance inst, uint32_t idx)
{
  /* the reference tests '>' (vulkansift.c:588); its own example documents '>=' as the intent (SURVEY B-D5) */
  if (idx >= inst->n_buffers)
  {
    LOGE(TAG, "Provided target buffer index is (%u) but the number of reserved buffers is (%u).", idx, inst->n_buffers);
    return false;
  }
  return true;
}

bool resolution_valid(vksift_Instance inst, uint32_t w, uint32_t h)
{
  const uint32_t size = w * h;
  if (size > inst->max_image_size)
  {
    LOGE(TAG, "Provided input image size (%u*%u=%u) is greater than the configured maximum image size (%u).", w, h, size, inst->max_image_size);
    return false;
  }
  if (size < 1024u)
  {
    LOGE(TAG, "Invalid input image size (%u*%u=%u). Input image size must be greater than or equal to 1024", w, h, size);
    return false;
  }
  return true;
}

/* ---- memory ---------------------------------------------------------------- */
bool alloc_pyramid(vksift_Instance inst)
{
  Pyramid &p = inst->pyr;
  const size_t ns = inst->cfg.nb_scales_per_octave;
  p.es = (inst->cfg.pyramid_precision_mode == VKSIFT_PYRAMID_PRECISION_FLOAT16) ? 2 : 4;
  const uint32_t row_align = 128u / (uint32_t)p.es; /* rows start on 128-byte boundaries */
  for (uint32_t o = 0; o < p.n_oct; o++)
  {
    p.pitch[o] = (p.w[o] + row_align - 1u) & ~(row_align - 1u);
    const size_t layer = (size_t)p.pitch[o] * p.h[o] * (size_t)p.es;
    const size_t need_g = layer * (ns + 3), need_d = layer * (ns + 2);
    if (need_g > p.alloc_bytes_g[o])
    {
      if (p.G[o])
        CU_TRY(cudaFree(p.G[o]));
      p.G[o] = nullptr;
      CU_TRY(cudaMalloc(&p.G[o], need_g));
      p.alloc_bytes_g[o] = need_g;
    }
    if (need_d > p.alloc_bytes_d[o])
    {
      if (p.D[o])
        CU_TRY(cudaFree(p.D[o]));
      p.D[o] = nullptr;
      CU_TRY(cudaMalloc(&p.D[o], need_d));
      p.alloc_bytes_d[o] = need_d;
    }
  }
  return true;
}

/* Launch plan of the scale space for the current resolution: the CUDA twin of
 * recScaleSpaceConstructionCmds + recDifferenceOfGaussianCmds (sift_detector.c:893-1079). */
bool build_blur_plan(vksift_Instance inst)
{
  const Pyramid &p = inst->pyr;
  const int ns = inst->cfg.nb_scales_per_octave;
  auto make_pass = [&](uint32_t o, int s) {
    BlurPass bp;
    memset(&bp, 0, sizeof(bp));
    const size_t layer = (size_t)p.pitch[o] * p.h[o] * (size_t)p.es; /* bytes */
    bp.w = (int)p.w[o];
    bp.h = (int)p.h[o];
    bp.dst_pitch = (int)p.pitch[o];
    bp.dst_g = (char *)p.G[o] + layer * s;
    if (s == 0)
    {
      bp.src = inst->d_src_slot; /* slot holding the address of the image (staging copy or caller's device buffer) */
      bp.src_kind = inst->cfg.use_input_upsampling ? BLUR_SRC_U8_UP2 : BLUR_SRC_U8;
      bp.src_w = (int)inst->cur_w;
      bp.src_h = (int)inst->cur_h;
      bp.src_pitch = (int)inst->cur_w;
      bp.dst_d = nullptr;
    }
    else
    {
      bp.src = (const char *)p.G[o] + layer * (s - 1);
      bp.src_kind = BLUR_SRC_LAYER;
      bp.src_pitch = (int)p.pitch[o];
      bp.dst_d = (char *)p.D[o] + layer * (s - 1);
    }
    if (s == ns && o + 1 < p.n_oct)
    {
      bp.dst_next = p.G[o + 1];
      bp.next_pitch = (int)p.pitch[o + 1];
      bp.next_w = (int)p.w[o + 1];
      bp.next_h = (int)p.h[o + 1];
    }
    bp.fp16 = (inst->cfg.pyramid_precision_mode == VKSIFT_PYRAMID_PRECISION_FLOAT16) ? 1 : 0;
    bp.radius = (int)inst->scales.radius[s];
    memcpy(bp.taps, inst->scales.taps[s], sizeof(bp.taps));
    return bp;
  };
  /* Schedule.  Layer s of octave o needs layer s-1 of the same octave, and layer 0 of octave o >= 1 is
   * written by the pass that produces layer ns of octave o-1 (reference order: sift_detector.c:1369-1378
   * records octave after octave, 72 serial dispatches).
   * Octaves [0,k) that are at least one tile wide run on the fast kernel, one launch per layer, each octave
   * in its own stream that starts once its seed exists: the late scales of an octave run next to the early
   * scales of the following ones and the block scheduler fills the tail of one launch with the head of another.
   * The remaining small octaves [k,n) are latency bound; they run on the compact kernel as one wavefront
   * (all passes that are ready share a launch) in a side stream and join before the extrema scan. */
  inst->fast_oct.clear();
  inst->strip_oct.clear();
  inst->steps_side.clear();
  if (p.n_oct == 0)
    return true;
  bool ok = true;
  int k = 0;
  for (; k < (int)p.n_oct; k++)
  {
    bool fast = true;
    for (int s = (k == 0 ? 0 : 1); s < ns + 3 && fast; s++)
    {
      const BlurPass bp = make_pass((uint32_t)k, s);
      fast = blur_pass_is_fast(bp);
    }
    if (!fast)
      break;
  }
  inst->use_expand = false;
  for (int o = 0; o < k; o++)
  {
    std::vector<BlurPass> passes;
    for (int s = (o == 0 ? 0 : 1); s < ns + 3; s++)
    {
      BlurPass bp = make_pass((uint32_t)o, s);
      if (o == 0 && s == 0 && !bp.fp16 && !getenv("VKSIFT_NO_EXPAND"))
      {
        /* large image: conversion and 2x blit by a kernel of their own, the first blur reads the expanded image through TMA */
        const size_t need = (size_t)p.pitch[0] * p.h[0] * sizeof(float);
        if (need > inst->expanded_bytes)
        {
          if (inst->d_expanded)
            CU_TRY(cudaFree(inst->d_expanded));
          inst->d_expanded = nullptr;
          inst->expanded_bytes = 0;
          CU_TRY(cudaMalloc(&inst->d_expanded, need));
          inst->expanded_bytes = need;
        }
        inst->use_expand = true;
        bp.src = inst->d_expanded;
        bp.src_kind = BLUR_SRC_LAYER;
        bp.src_pitch = (int)p.pitch[0];
      }
      if (!blur_pass_prepare_fast(&bp))
      {
        LOGE(TAG, "cuTensorMapEncodeTiled failed for a scale-space layer (%dx%d)", bp.w, bp.h);
        ok = false;
      }
      passes.push_back(bp);
    }
    inst->fast_oct.push_back(passes);
    /* layers 1..ns and ns+1, ns+2 as two launches of the streaming strip kernel where the octave is large enough and the
     * radii are the ones it is built for (default configuration); the seed pass of octave 0 stays a launch of its own */
    std::vector<StripLaunch> strips;
    const int first = (o == 0) ? 1 : 0; /* index of layer 1 in `passes` */
    if (inst->use_strip && ns == 3 && (int)passes.size() == first + 5)
    {
      StripLaunch a, c;
      if (strip_plan(passes.data() + first, 3, &a) && strip_plan(passes.data() + first + 3, 2, &c))
      {
        a.first_layer = 1;
        c.first_layer = 4;
        strips.push_back(a);
        strips.push_back(c);
      }
    }
    inst->strip_oct.push_back(strips);
  }
  auto wavefront = [&](int o_begin, int o_end, std::vector<BlurStep> &out) {
    if (o_end <= o_begin)
      return;
    const int n_steps = ns * (o_end - o_begin - 1) + ns + 3;
    for (int t = 0; t < n_steps; t++)
    {
      BlurStep step;
      memset(&step, 0, sizeof(step));
      for (int o = o_end - 1; o >= o_begin; o--)
      {
        const int s = t - ns * (o - o_begin);
        if (s < (o == 0 ? 0 : 1) || s > ns + 2 || step.n_pass >= VKS_MAX_PASSES_PER_STEP)
          continue;
        step.pass[step.n_pass++] = make_pass((uint32_t)o, s);
      }
      if (step.n_pass == 0)
        continue;
      blur_step_tiles(&step);
      out.push_back(step);
    }
  };
  /* small octaves: fused launches (chain = up to the layer seeding the next octave, rest = the others) */
  inst->fused_oct.clear();
  bool fused_ok = true;
  for (int o = k; o < (int)p.n_oct && fused_ok; o++)
  {
    vksift_Instance_T::FusedOct fo;
    std::vector<BlurPass> passes;
    for (int s = 1; s < ns + 3; s++)
      passes.push_back(make_pass((uint32_t)o, s));
    if (o == 0)
    {
      memset(&fo.seed_step, 0, sizeof(fo.seed_step));
      fo.seed_step.pass[0] = make_pass(0, 0);
      fo.seed_step.n_pass = 1;
      blur_step_tiles(&fo.seed_step);
      fo.has_seed_step = true;
    }
    std::vector<FusedLaunch> all;
    fused_ok = fused_plan_octave(passes.data(), (int)passes.size(), &all);
    bool seen_next = (o + 1 >= (int)p.n_oct); /* the last octave seeds nothing: everything is off the critical path */
    for (const FusedLaunch &F : all)
    {
      if (seen_next && o + 1 < (int)p.n_oct)
        fo.rest.push_back(F);
      else
        fo.chain.push_back(F);
      if (F.next_k >= 0)
        seen_next = true;
    }
    inst->fused_oct.push_back(fo);
  }
  if (!fused_ok)
  {
    inst->fused_oct.clear();
    wavefront(k, (int)p.n_oct, inst->steps_side);
  }
  return ok;
}

void fill_detect_params(vksift_Instance inst, const FeatureBuffer &fb, DetectParams *P);

void invalidate_graphs(vksift_Instance inst);

bool set_resolution(vksift_Instance inst, uint32_t w, uint32_t h)
{
  invalidate_graphs(inst); /* captured launches carry the old geometry */
  inst->cur_w = w;
  inst->cur_h = h;
  Pyramid &p = inst->pyr;
  p.n_oct = plan_octaves(w, h, inst->cfg.use_input_upsampling, inst->max_octaves, p.w, p.h);
  if (!alloc_pyramid(inst))
    return false;
  if (!build_blur_plan(inst))
    return false;
  {
    /* tensor maps of the extrema scan and the bitmaps of the ordered compaction only depend on the pyramid geometry */
    DetectParams P;
    FeatureBuffer dummy;
    fill_detect_params(inst, dummy, &P);
    CU_TRY(extrema_plan_build(P, &inst->extrema_plan));
    size_t words = 0, rows = 0, qn = 0;
    extrema_layout(&P, &words, &rows, &qn, inst->q_total);
    if (qn > inst->q_alloc)
    {
      if (inst->raw_q)
        CU_TRY(cudaFree(inst->raw_q));
      if (inst->q_heads)
        CU_TRY(cudaFree(inst->q_heads));
      inst->raw_q = nullptr;
      inst->q_heads = nullptr;
      inst->q_alloc = 0;
      CU_TRY(cudaMalloc(&inst->raw_q, sizeof(unsigned long long) * qn));
      CU_TRY(cudaMalloc(&inst->q_heads, sizeof(FeatHead) * qn));
      inst->q_alloc = qn;
    }
    if (words >= (1ull << 31))
    {
      LOGE(TAG, "scale space too large for the keypoint bitmaps (%zu words)", words);
      return false;
    }
    inst->bm_words = words;
    inst->compact_rows = rows;
    if (words + rows > inst->compact_alloc_words)
    {
      if (inst->compact_mem)
        CU_TRY(cudaFree(inst->compact_mem));
      inst->compact_mem = nullptr;
      inst->compact_alloc_words = 0;
      CU_TRY(cudaMalloc(&inst->compact_mem, sizeof(uint32_t) * (words + rows)));
      inst->compact_alloc_words = words + rows;
    }
  }
  return true;
}

void update_buffer_sections(vksift_Instance inst, FeatureBuffer &fb)
{
  invalidate_graphs(inst);
  fb.cur_w = inst->cur_w;
  fb.cur_h = inst->cur_h;
  fb.n_oct = inst->pyr.n_oct;
  plan_sections(inst->cfg.max_nb_sift_per_buffer, fb.n_oct, fb.cap);
}

void fill_detect_params(vksift_Instance inst, const FeatureBuffer &fb, DetectParams *P)
{
  memset(P, 0, sizeof(*P));
  const Pyramid &p = inst->pyr;
  const vksift_Config &c = inst->cfg;
  uint32_t off = 0;
  for (uint32_t o = 0; o < p.n_oct; o++)
  {
    P->oct[o].G = p.G[o];
    P->oct[o].D = p.D[o];
    P->oct[o].w = (int)p.w[o];
    P->oct[o].h = (int)p.h[o];
    P->oct[o].pitch = (int)p.pitch[o];
    P->oct[o].layer_stride = (int)(p.pitch[o] * p.h[o]);
    P->oct[o].fp16 = (p.es == 2) ? 1 : 0;
    P->cap[o] = fb.cap[o];
    P->sec_off[o] = off;
    off += fb.cap[o];
  }
  P->n_oct = (int)p.n_oct;
  P->ob = 0;
  P->oe = (int)p.n_oct;
  P->ns = c.nb_scales_per_octave;
  P->upsample = c.use_input_upsampling ? 1 : 0;
  P->sigma0 = c.seed_scale_sigma;
  P->thr = c.intensity_threshold / c.nb_scales_per_octave; /* sift_detector.c:1136 */
  P->prefilter = P->thr * 0.8f;                             /* ExtractKeypoints.comp:58 */
  P->edge_limit = ((c.edge_threshold + 1.f) * (c.edge_threshold + 1.f)) / c.edge_threshold; /* :203 */
  P->max_ori = c.max_nb_orientation_per_keypoint;
  P->ori_stride = inst->ori_stride;
  P->vlfeat = (c.descriptor_format == VKSIFT_DESCRIPTOR_FORMAT_VLFEAT) ? 1 : 0;
  P->max_feats = c.max_nb_sift_per_buffer;
  size_t words = 0, rows = 0, qn = 0;
  extrema_layout(P, &words, &rows, &qn, inst->q_total);
  P->acc_bm = inst->compact_mem;
  P->row_cnt = inst->compact_mem + inst->bm_words;
  P->raw_q = inst->raw_q;
  P->q_heads = inst->q_heads;
}

void wait_lane(vksift_Instance lane)
{
  if (lane->detect_pending)
  {
    cudaEventSynchronize(lane->ev_detect_done);
    lane->detect_pending = false;
  }
}

vksift_Instance lane_of_buffer(vksift_Instance inst, uint32_t buf) { return inst->lanes.empty() ? inst : inst->lanes[buf % inst->lanes.size()]; }
vksift_Instance last_lane(vksift_Instance inst) { return inst->lanes.empty() ? inst : inst->lanes[inst->last_lane]; }

/* detect: every lane of the instance */
void wait_pipelines(vksift_Instance inst, bool detect, bool match)
{
  if (detect)
  {
    wait_lane(inst);
    for (size_t k = 1; k < inst->lanes.size(); k++)
      wait_lane(inst->lanes[k]);
  }
  if (match && inst->match_pending)
  {
    cudaEventSynchronize(inst->ev_match_done);
    inst->match_pending = false;
  }
}

uint32_t buffer_count(vksift_Instance inst, uint32_t idx, bool log_lost)
{
  FeatureBuffer &fb = inst->buffers[idx];
  if (fb.uploaded)
    return fb.n_uploaded;
  if (log_lost)
  {
    uint32_t lost = 0;
    for (uint32_t o = 0; o < fb.n_oct && o < VKS_MAX_OCT; o++) /* slots of higher octaves may be stale from an earlier, larger resolution */
      lost += fb.host_counts[1 + o] - fb.host_counts[1 + VKS_MAX_OCT + o];
    if (lost > 0)
      LOGE(TAG,
           "%u feature(s) lost because the SIFT buffer was full, consider increasing the maximum number of SIFT features per buffer in the "
           "configuration.",
           lost);
  }
  return fb.host_counts[0];
}

void destroy_instance(vksift_Instance inst)
{
  DeviceGuard g(inst->device);
  cudaDeviceSynchronize();
  for (size_t k = 1; k < inst->lanes.size(); k++)
    destroy_instance(inst->lanes[k]);
  inst->lanes.clear();
  if (inst->ev_h2d)
    cudaEventDestroy(inst->ev_h2d);
  for (auto &fb : inst->own_buffers)
  {
    cudaFree(fb.heads);
    cudaFree(fb.desc);
    cudaFree(fb.cnt);
    cudaFree(fb.norm_plain);
    cudaFree(fb.norm_packed);
    if (fb.host_counts)
      cudaFreeHost(fb.host_counts);
  }
  for (uint32_t o = 0; o < VKS_MAX_OCT; o++)
  {
    cudaFree(inst->pyr.G[o]);
    cudaFree(inst->pyr.D[o]);
  }
  if (inst->h_image)
    cudaFreeHost(inst->h_image);
  cudaFree(inst->d_image);
  invalidate_graphs(inst);
  if (inst->h_src_slot)
    cudaFreeHost(inst->h_src_slot);
  cudaFree(inst->d_src_slot);
  cudaFree(inst->compact_mem);
  cudaFree(inst->raw_q);
  cudaFree(inst->q_heads);
  cudaFree(inst->prim);
  cudaFree(inst->ori);
  cudaFree(inst->n_ori);
  cudaFree(inst->feat_src);
  if (!inst->primary)
    cudaFree(inst->desc_m_table);
  cudaFree(inst->d_aos);
  cudaFree(inst->d_matches);
  cudaFree(inst->d_matches_rev);
  cudaFree(inst->d_matches_blocks);
  cudaFree(inst->d_block_norms);
  cudaFree(inst->d_expanded);
  exchange_destroy(inst->exchange);
  inst->exchange = nullptr;
  cudaFree(inst->d_pairs);
  match_workspace_destroy(inst->match_ws);
  extrema_plan_destroy(inst->extrema_plan);
  for (int i = 0; i < EV_COUNT; i++)
    if (inst->ev[i])
      cudaEventDestroy(inst->ev[i]);
  if (inst->ev_detect_done)
    cudaEventDestroy(inst->ev_detect_done);
  if (inst->ev_match_done)
    cudaEventDestroy(inst->ev_match_done);
  if (inst->ev_join)
    cudaEventDestroy(inst->ev_join);
  if (inst->ev_join2)
    cudaEventDestroy(inst->ev_join2);
  for (int o = 0; o < VKS_MAX_OCT; o++)
    if (inst->ev_chain[o])
      cudaEventDestroy(inst->ev_chain[o]);
  if (inst->side2_stream)
    cudaStreamDestroy(inst->side2_stream);
  for (int o = 0; o < VKS_MAX_OCT; o++)
  {
    if (inst->ev_seed[o])
      cudaEventDestroy(inst->ev_seed[o]);
    if (inst->ev_oct_done[o])
      cudaEventDestroy(inst->ev_oct_done[o]);
    if (inst->ev_pyr[o])
      cudaEventDestroy(inst->ev_pyr[o]);
    if (inst->oct_stream[o])
      cudaStreamDestroy(inst->oct_stream[o]);
  }
  if (inst->side_stream)
    cudaStreamDestroy(inst->side_stream);
  if (inst->stream)
    cudaStreamDestroy(inst->stream);
  delete inst;
}

bool create_resources(vksift_Instance inst)
{
  const vksift_Config &c = inst->cfg;
  CU_TRY(cudaStreamCreateWithFlags(&inst->stream, cudaStreamNonBlocking));
  /* The dependency chain of the pyramid runs octave 0 -> 1 -> ... (each needs layer ns of the previous one), so the
   * smaller an octave, the later it starts and the more the pipeline waits for it: smaller octaves get the higher
   * stream priority and their CTAs are placed before the queued CTAs of the big, throughput-bound launches. */
  int prio_lo = 0, prio_hi = 0;
  CU_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi)); /* numerically lower = more urgent */
  auto prio = [&](int level) { return prio_lo - level < prio_hi ? prio_hi : prio_lo - level; };
  CU_TRY(cudaStreamCreateWithPriority(&inst->side_stream, cudaStreamNonBlocking, prio(3)));
  CU_TRY(cudaStreamCreateWithPriority(&inst->side2_stream, cudaStreamNonBlocking, prio(1)));
  CU_TRY(cudaEventCreateWithFlags(&inst->ev_join2, cudaEventDisableTiming));
  for (int o = 0; o < VKS_MAX_OCT; o++)
    CU_TRY(cudaEventCreateWithFlags(&inst->ev_chain[o], cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&inst->ev_join, cudaEventDisableTiming));
  for (int o = 1; o < VKS_MAX_OCT; o++)
  {
    CU_TRY(cudaStreamCreateWithPriority(&inst->oct_stream[o], cudaStreamNonBlocking, prio(o >= 2 ? 2 : 1)));
    CU_TRY(cudaEventCreateWithFlags(&inst->ev_seed[o], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&inst->ev_oct_done[o], cudaEventDisableTiming));
    CU_TRY(cudaEventCreate(&inst->ev_pyr[o]));
  }
  CU_TRY(cudaEventCreateWithFlags(&inst->ev_detect_done, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&inst->ev_match_done, cudaEventDisableTiming));
  /* waited for inside vksift_detectFeatures while a 2 MB copy runs (tens of microseconds): spin, a blocking wait costs a sleep and a
   * late wake-up per image and makes the end-to-end loop sensitive to the host's scheduler */
  CU_TRY(cudaEventCreateWithFlags(&inst->ev_h2d, cudaEventDisableTiming));
  for (int i = 0; i < EV_COUNT; i++)
    CU_TRY(cudaEventCreate(&inst->ev[i]));

  uint32_t side = 0;
  inst->max_octaves = plan_max_octaves(&c, &side);
  inst->max_image_size = side * side; /* sift_memory.c:641 */
  plan_gaussian_taps(&inst->scales, &c);

  CU_TRY(cudaHostAlloc(&inst->h_image, inst->max_image_size, cudaHostAllocDefault));
  CU_TRY(cudaMalloc(&inst->d_image, inst->max_image_size));
  CU_TRY(cudaHostAlloc(&inst->h_src_slot, sizeof(void *), cudaHostAllocDefault));
  CU_TRY(cudaMalloc(&inst->d_src_slot, sizeof(void *)));
  /* detection lanes: VKSIFT_LANES overrides the default of one lane per feature buffer, at most 8
   * (1920x1080: 0.45 / 0.35 / 0.315 / 0.309 ms per image with 1 / 2 / 4 / 8 lanes; 640x480: 0.224 / 0.119 / 0.080 / 0.070) */
  uint32_t n_lanes = c.sift_buffer_count < 8u ? c.sift_buffer_count : 8u;
  if (const char *e = getenv("VKSIFT_LANES"))
  {
    const long v = strtol(e, nullptr, 10);
    if (v >= 1)
      n_lanes = (uint32_t)v < c.sift_buffer_count ? (uint32_t)v : c.sift_buffer_count;
  }
  {
    /* Alternative schedules of the same kernels (measurements in DESIGN.md):
     * VKSIFT_GRAPH=0/1: replay of the detection as a CUDA graph, the analogue of the reference's pre-recorded command
     * buffer.  A replay costs 8 us of CPU time instead of 210 us but loses the programmatic-dependent-launch edges and
     * the stream priorities, so one detection alone is ~20 us slower; with several lanes the throughput is what counts
     * and the replay wins (0.315 against 0.321 ms per 1920x1080 image), so it is the default exactly then. */
    const char *g = getenv("VKSIFT_GRAPH");
    if (const char *sp = getenv("VKSIFT_STRIP"))
      inst->use_strip = (sp[0] == '1');
    const char *nsp = getenv("VKSIFT_NO_SPLIT");
    inst->no_split = (nsp && nsp[0] == '1');
    inst->use_graph = (g ? g[0] == '1' : n_lanes > 1);
    inst->ori_ctas = n_lanes > 1 ? 2 : 4; /* throughput with several detections in flight, latency with one (launch_orientation) */
    /* extrema scan CTAs (launch_extrema): 296 alone, 74 with two lanes, 37 with four, 24 from six lanes on */
    inst->scan_ctas = n_lanes > 1 ? (int)(148u / n_lanes > 24u ? 148u / n_lanes : 24u) : 0;
    if (inst->primary)
      inst->use_graph = inst->primary->use_graph;
  }

  const size_t maxf = c.max_nb_sift_per_buffer;
  {
    /* extrema queue: raw extrema outnumber accepted keypoints 2-3x; an overflow is handled (slow path), never dropped */
    unsigned long long q = 4ull * c.max_nb_sift_per_buffer;
    q = q < 16384ull ? 16384ull : (q > (1ull << 22) ? (1ull << 22) : q);
    if (const char *e = getenv("VKSIFT_RAW_QUEUE"))
    {
      const long v = strtol(e, nullptr, 10);
      if (v >= 1)
        q = (unsigned long long)v;
    }
    inst->q_total = (uint32_t)q;
  }
  inst->ori_stride = (c.max_nb_orientation_per_keypoint == 0 || c.max_nb_orientation_per_keypoint > VKS_MAX_ORI) ? VKS_MAX_ORI
                                                                                                                   : c.max_nb_orientation_per_keypoint;
  CU_TRY(cudaMalloc(&inst->prim, sizeof(FeatHead) * (maxf + 1)));
  CU_TRY(cudaMalloc(&inst->ori, sizeof(float) * (maxf + 1) * inst->ori_stride));
  CU_TRY(cudaMalloc(&inst->n_ori, sizeof(uint32_t) * (maxf + 1)));
  CU_TRY(cudaMalloc(&inst->feat_src, sizeof(uint32_t) * (maxf + 1)));
  if (inst->primary)
    inst->desc_m_table = inst->primary->desc_m_table; /* read-only table, built once per instance */
  else
  {
    /* M(hr) of ComputeDescriptors.comp:116-124 only depends on the window radius: tabulated on the host with the shared
     * arithmetic of include/vksift_arith.h (same IEEE sequence under g++ -ffp-contract=off as in the kernels) */
    float h_table[VKS_DESC_M_TABLE];
    for (int hr = 0; hr < VKS_DESC_M_TABLE; hr++)
    {
      float m = 0.f;
      for (int i = 0; i < hr; i++)
        for (int j = i; j < hr; j++)
        {
          float t = vks_mul(vks_expf(vks_mul(-0.125f, (float)((i * i) + (j * j)))), VKS_SQRT2_F);
          if (j > i)
            t = vks_mul(t, 2.f);
          m = vks_add(m, t);
        }
      h_table[hr] = m;
    }
    CU_TRY(cudaMalloc(&inst->desc_m_table, sizeof(float) * VKS_DESC_M_TABLE));
    CU_TRY(cudaMemcpy(inst->desc_m_table, h_table, sizeof(h_table), cudaMemcpyHostToDevice));
  }
  CU_TRY(cudaMalloc(&inst->d_aos, sizeof(vksift_Feature) * maxf));
  inst->graphs.resize(2 * (size_t)c.sift_buffer_count);
  if (inst->primary)
  {
    /* a secondary lane: detection only, on the primary's feature buffers */
    inst->buffers = inst->primary->buffers;
    inst->n_buffers = inst->primary->n_buffers;
    return set_resolution(inst, side, side);
  }
  CU_TRY(cudaMalloc(&inst->d_matches, sizeof(vksift_Match_2NN) * maxf));
  CU_TRY(cudaMalloc(&inst->d_matches_rev, sizeof(vksift_Match_2NN) * maxf));
  CU_TRY(cudaMalloc(&inst->d_pairs, sizeof(uint32_t) * (2 * maxf + 1)));
  CU_TRY(match_workspace_create(&inst->match_ws, c.max_nb_sift_per_buffer));

  inst->own_buffers.resize(c.sift_buffer_count);
  inst->buffers = inst->own_buffers.data();
  inst->n_buffers = c.sift_buffer_count;
  for (auto &fb : inst->own_buffers)
  {
    /* +256 rows: the matcher's TMA boxes and |b|^2 loads may run past the last feature */
    CU_TRY(cudaMalloc(&fb.heads, sizeof(FeatHead) * (maxf + 256)));
    CU_TRY(cudaMalloc(&fb.desc, 128 * (maxf + 256)));
    CU_TRY(cudaMemset(fb.desc, 0, 128 * (maxf + 256)));
    CU_TRY(cudaMalloc(&fb.cnt, sizeof(DetectCounters)));
    CU_TRY(cudaMemset(fb.cnt, 0, sizeof(DetectCounters)));
    CU_TRY(cudaMalloc(&fb.norm_plain, sizeof(uint32_t) * (maxf + 256)));
    CU_TRY(cudaMalloc(&fb.norm_packed, sizeof(uint32_t) * (maxf + 256)));
    CU_TRY(cudaHostAlloc(&fb.host_counts, sizeof(uint32_t) * (1 + 2 * VKS_MAX_OCT), cudaHostAllocMapped));
    memset(fb.host_counts, 0, sizeof(uint32_t) * (1 + 2 * VKS_MAX_OCT));
    CU_TRY(cudaHostGetDevicePointer(&fb.host_counts_dev, fb.host_counts, 0));
  }
  /* default resolution = square of the maximum size (sift_memory.c:637-640) */
  if (!set_resolution(inst, side, side))
    return false;
  for (auto &fb : inst->own_buffers)
    update_buffer_sections(inst, fb);
  inst->lanes.push_back(inst);
  for (uint32_t k = 1; k < n_lanes; k++)
  {
    vksift_Instance lane = new vksift_Instance_T();
    lane->cfg = inst->cfg;
    lane->device = inst->device;
    lane->primary = inst;
    inst->lanes.push_back(lane); /* owned from here on: destroy_instance(inst) frees it */
    if (!create_resources(lane))
      return false;
  }
  return true;
}

#ifdef VKS_ANALYSIS
#define VKS_SKIP(inst) ((inst)->debug_skip)
#else
#define VKS_SKIP(inst) 0
#endif

/* Named ranges around the enqueue of each stage, with the region names the reference gives its debug markers
 * (beginMarkerRegion, sift_detector.c:29-50, 865-1297; sift_matcher.c:251) and, like there, only when the instance was
 * created with use_gpu_debug_functions: an Nsight timeline of this library reads like a RenderDoc / Nsight Graphics capture
 * of the reference. */
struct MarkerRegion
{
  bool on;
  MarkerRegion(vksift_Instance inst, const char *name) : on(inst->cfg.use_gpu_debug_functions)
  {
    if (on)
      nvtxRangePushA(name);
  }
  ~MarkerRegion()
  {
    if (on)
      nvtxRangePop();
  }
};

struct TraceScope
{
  vksift_Instance inst;
  cudaStream_t st;
  size_t idx = (size_t)-1;
  TraceScope(vksift_Instance i, cudaStream_t s, const char *fmt, int a, int b, int c = 0) : inst(i), st(s)
  {
    if (!inst->trace)
      return;
    if (inst->trace_used == inst->trace_marks.size())
    {
      vksift_Instance_T::TraceMark m;
      cudaEventCreate(&m.e0);
      cudaEventCreate(&m.e1);
      inst->trace_marks.push_back(m);
    }
    idx = inst->trace_used++;
    snprintf(inst->trace_marks[idx].name, sizeof(inst->trace_marks[idx].name), fmt, a, b, c);
    cudaEventRecord(inst->trace_marks[idx].e0, st);
  }
  ~TraceScope()
  {
    if (idx != (size_t)-1)
      cudaEventRecord(inst->trace_marks[idx].e1, st);
  }
};

void trace_dump(vksift_Instance inst)
{
  if (!inst->trace || !inst->trace_dump_stderr || inst->trace_used == 0 || !inst->profiling)
    return;
  cudaEventSynchronize(inst->ev[EV_D4]);
  fprintf(stderr, "[trace] %-24s %9s %9s %9s\n", "launch", "start_us", "end_us", "dur_us");
  for (size_t i = 0; i < inst->trace_used; i++)
  {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, inst->ev[EV_D0], inst->trace_marks[i].e0);
    cudaEventElapsedTime(&b, inst->ev[EV_D0], inst->trace_marks[i].e1);
    fprintf(stderr, "[trace] %-24s %9.1f %9.1f %9.1f\n", inst->trace_marks[i].name, a * 1e3f, b * 1e3f, (b - a) * 1e3f);
  }
  float d1 = 0.f;
  cudaEventElapsedTime(&d1, inst->ev[EV_D0], inst->ev[EV_D1]);
  fprintf(stderr, "[trace] scale space done at %.1f us\n", d1 * 1e3f);
}

/* every operation of one detection on the instance's streams, in an order that stream capture accepts
 * (the side streams fork from and join the main stream through events) */
bool record_detection(vksift_Instance inst, uint32_t buf)
{
  FeatureBuffer &fb = inst->buffers[buf];
  cudaStream_t st = inst->stream;
  DetectParams P;
  fill_detect_params(inst, fb, &P);
  const bool prof = inst->profiling;
  /* stage-time events must become event-record nodes when the sequence is captured into a graph */
  cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
  CU_TRY(cudaStreamIsCapturing(st, &cap_status));
  const bool capturing = (cap_status == cudaStreamCaptureStatusActive);

  inst->trace_used = 0;
  {
    MarkerRegion mr(inst, "Clear buffer data");
    CU_TRY(cudaMemcpyAsync(inst->d_src_slot, inst->h_src_slot, sizeof(void *), cudaMemcpyHostToDevice, st));
  }
  MarkerRegion mr_pyr(inst, "Scale space construction + DoG computation");
  if (prof)
    CU_TRY(cudaEventRecordWithFlags(inst->ev[EV_D0], st, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
  CU_TRY(cudaMemsetAsync(fb.cnt, 0, sizeof(DetectCounters), st));
  CU_TRY(cudaMemsetAsync(inst->compact_mem, 0, sizeof(uint32_t) * (inst->bm_words + inst->compact_rows), st)); /* accepted bitmap + row counters */
  const int ns = inst->cfg.nb_scales_per_octave;
  const int n_fast = (int)inst->fast_oct.size();
  /* Split schedule: the extrema scan, ordering and orientation pass of an octave (their per-octave sections are
   * independent) follow that octave's scale space on the same stream, so they overlap the scale space of the later
   * octaves, a latency chain that leaves the GPU nearly idle at its end; everything joins before the feature assembly. */
  /* serial schedule (analysis only): every launch on the main stream in dependency order, so that the event pair around
   * a launch times that kernel alone */
  const bool serial = inst->serial;
  cudaStream_t side2 = serial ? st : inst->side2_stream;
  const bool split = (n_fast > 0) && !inst->fused_oct.empty() && n_fast < P.n_oct && !inst->no_split && !serial;
  inst->ev_d1b_valid = split && prof;
  inst->n_pyr_events = (split && prof) ? n_fast : 0;
  auto post_chain = [&](int ob, int oe, cudaStream_t s) -> bool {
    MarkerRegion mr(inst, "ExtractKeypoints + ComputeOrientation");
    DetectParams Q = P;
    Q.ob = ob;
    Q.oe = oe;
    if (!(VKS_SKIP(inst) & 4))
    {
      CU_TRY(launch_extrema(Q, inst->extrema_plan, fb.cnt, inst->prim, s, &inst->launches, inst->crowded ? inst->scan_ctas : 0));
    }
    if (!(VKS_SKIP(inst) & 2))
      CU_TRY(launch_orientation(Q, fb.cnt, inst->prim, inst->ori, inst->n_ori, s, inst->crowded ? inst->ori_ctas : 4));
    inst->launches += 1;
    return true;
  };
  for (int o = 0; o < n_fast; o++)
  {
    cudaStream_t so = (o == 0 || serial) ? st : inst->oct_stream[o];
    if (o > 0)
      CU_TRY(cudaStreamWaitEvent(so, inst->ev_seed[o], 0));
    const bool strips = !inst->strip_oct[o].empty();
    for (BlurPass &bp : inst->fast_oct[o])
    {
      const bool is_seed = (bp.src_kind != BLUR_SRC_LAYER) || bp.dst_d == nullptr; /* layer 0 of octave 0 */
      if (strips && !is_seed)
        break; /* layers >= 1 come from the strip launches below */
      if (is_seed && inst->use_expand && !(VKS_SKIP(inst) & 8))
      {
        TraceScope ts(inst, so, "expand o%d r%d", o, 0);
        CU_TRY(launch_expand_input(inst->d_src_slot, (int)inst->cur_w, (int)inst->cur_h, inst->cfg.use_input_upsampling ? 1 : 0, inst->d_expanded,
                                   (int)inst->pyr.pitch[0], (int)inst->pyr.w[0], (int)inst->pyr.h[0], so));
        inst->launches++;
      }
      if (!(VKS_SKIP(inst) & 8))
      {
        TraceScope ts(inst, so, is_seed ? "seed o%d r%d" : "fast o%d r%d", o, bp.radius);
        CU_TRY(launch_blur_pass_fast(bp, so));
      }
      inst->launches++;
      if (bp.dst_next && o + 1 < VKS_MAX_OCT)
        CU_TRY(cudaEventRecord(inst->ev_seed[o + 1], so)); /* written by the pass producing layer ns */
    }
    if (strips)
      for (const StripLaunch &L : inst->strip_oct[o])
      {
        if (!(VKS_SKIP(inst) & 8))
        {
          TraceScope ts(inst, so, "strip o%d g%d+%d", o, L.first_layer, L.n_layers);
          CU_TRY(launch_strip(L, so));
        }
        inst->launches++;
        if (L.first_layer + L.n_layers - 1 == ns && o + 1 < (int)inst->pyr.n_oct)
          CU_TRY(cudaEventRecord(inst->ev_seed[o + 1], so)); /* the chain ending in layer ns wrote the next octave's seed */
      }
    if (o > 0)
    {
      if (split)
      {
        if (prof)
          CU_TRY(cudaEventRecordWithFlags(inst->ev_pyr[o], so, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
        if (!post_chain(o, o + 1, so))
          return false;
      }
      CU_TRY(cudaEventRecord(inst->ev_oct_done[o], so));
    }
  }
  if (!inst->fused_oct.empty())
  {
    cudaStream_t ss = (n_fast == 0 || serial) ? st : inst->side_stream;
    if (n_fast > 0)
      CU_TRY(cudaStreamWaitEvent(ss, inst->ev_seed[n_fast], 0));
    bool used2 = false;
    for (size_t j = 0; j < inst->fused_oct.size(); j++)
    {
      auto &fo = inst->fused_oct[j];
      if (fo.has_seed_step)
      {
        CU_TRY(launch_blur_step(fo.seed_step, ss));
        inst->launches++;
      }
      for (const FusedLaunch &F : fo.chain)
      {
        TraceScope ts(inst, ss, "fused chain o%d n%d", n_fast + (int)j, F.n_layers);
        if (!(VKS_SKIP(inst) & 8))
          CU_TRY(launch_fused(F, ss));
        inst->launches++;
      }
      if (!fo.rest.empty())
      {
        CU_TRY(cudaEventRecord(inst->ev_chain[j], ss));
        CU_TRY(cudaStreamWaitEvent(side2, inst->ev_chain[j], 0));
        for (const FusedLaunch &F : fo.rest)
        {
          TraceScope ts(inst, side2, "fused rest o%d n%d", n_fast + (int)j, F.n_layers);
          if (!(VKS_SKIP(inst) & 8))
            CU_TRY(launch_fused(F, side2));
          inst->launches++;
        }
        used2 = true;
      }
    }
    if (split)
    {
      if (used2)
      {
        CU_TRY(cudaEventRecord(inst->ev_join2, side2));
        CU_TRY(cudaStreamWaitEvent(ss, inst->ev_join2, 0));
      }
      if (prof)
        CU_TRY(cudaEventRecordWithFlags(inst->ev[EV_D1B], ss, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
      if (!post_chain(n_fast, P.n_oct, ss))
        return false;
      CU_TRY(cudaEventRecord(inst->ev_join, ss));
    }
    else
    {
      if (n_fast > 0)
      {
        CU_TRY(cudaEventRecord(inst->ev_join, ss));
        CU_TRY(cudaStreamWaitEvent(st, inst->ev_join, 0));
      }
      if (used2)
      {
        CU_TRY(cudaEventRecord(inst->ev_join2, side2));
        CU_TRY(cudaStreamWaitEvent(st, inst->ev_join2, 0));
      }
    }
  }
  if (!inst->steps_side.empty())
  {
    cudaStream_t ss = (n_fast == 0 || serial) ? st : inst->side_stream;
    if (n_fast > 0)
      CU_TRY(cudaStreamWaitEvent(ss, inst->ev_seed[n_fast], 0));
    for (BlurStep &step : inst->steps_side)
    {
      CU_TRY(launch_blur_step(step, ss));
      inst->launches++;
    }
    if (n_fast > 0)
    {
      CU_TRY(cudaEventRecord(inst->ev_join, ss));
      CU_TRY(cudaStreamWaitEvent(st, inst->ev_join, 0));
    }
  }
  if (!split)
    for (int o = 1; o < n_fast; o++)
      CU_TRY(cudaStreamWaitEvent(st, inst->ev_oct_done[o], 0));
  if (prof)
    CU_TRY(cudaEventRecordWithFlags(inst->ev[EV_D1], st, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
  {
    const cudaError_t pe = cudaPeekAtLastError();
    if (pe != cudaSuccess)
      LOGE(TAG, "pending CUDA error before the extrema scan: %s", cudaGetErrorName(pe));
  }
  DetectParams PA = P;
  if (split)
    PA.oe = 1; /* octave 0 (its layers are complete on this stream); the other octaves follow on the side stream */
  MarkerRegion mr_feat(inst, "ExtractKeypoints + ComputeOrientation + ComputeDescriptors + CopySiftCount");
  if (!(VKS_SKIP(inst) & 4))
  {
    CU_TRY(launch_extrema(PA, inst->extrema_plan, fb.cnt, inst->prim, st, &inst->launches, inst->crowded ? inst->scan_ctas : 0));
  }
  if (prof)
    CU_TRY(cudaEventRecordWithFlags(inst->ev[EV_D2], st, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
  if (!(VKS_SKIP(inst) & 2))
    CU_TRY(launch_orientation(PA, fb.cnt, inst->prim, inst->ori, inst->n_ori, st, inst->crowded ? inst->ori_ctas : 4));
  inst->launches++;
  if (split)
  {
    /* the other octaves' keypoints are oriented too */
    for (int o = 1; o < n_fast; o++)
      CU_TRY(cudaStreamWaitEvent(st, inst->ev_oct_done[o], 0));
    CU_TRY(cudaStreamWaitEvent(st, inst->ev_join, 0));
  }
  if (prof)
    CU_TRY(cudaEventRecordWithFlags(inst->ev[EV_D3], st, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
  CU_TRY(launch_assemble(P, fb.cnt, inst->n_ori, inst->feat_src, fb.host_counts_dev, st));
  if (!(VKS_SKIP(inst) & 1))
    CU_TRY(launch_descriptors(P, fb.cnt, inst->desc_m_table, inst->prim, inst->ori, inst->feat_src, fb.heads, fb.desc, st));
  inst->launches += 2;
  if (prof)
    CU_TRY(cudaEventRecordWithFlags(inst->ev[EV_D4], st, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
  return true;
}

void invalidate_graphs(vksift_Instance inst)
{
  for (auto &g : inst->graphs)
  {
    if (g.exec)
      cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
    g.uses = 0;
  }
}

/* enqueue the whole detection pipeline (sift_detector.c:1369-1393): replay of the buffer's graph, captured on its second use */
bool enqueue_detection(vksift_Instance inst, const uint8_t *d_image, uint32_t buf)
{
  FeatureBuffer &fb = inst->buffers[buf];
  cudaStream_t st = inst->stream;
  *inst->h_src_slot = d_image;
  auto &g = inst->graphs[2 * buf + (inst->crowded ? 1 : 0)]; /* one graph per grid policy */
  /* stage profiling and launch tracing are analysis modes of the eager schedule (the one a single-lane instance runs) */
  const bool want_graph = inst->use_graph && !inst->trace && !inst->profiling;
  if (g.exec && (g.prof != inst->profiling || !want_graph))
  {
    cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
  }
  if (want_graph && !g.exec && g.uses >= 1)
  {
    const uint64_t l0 = inst->launches;
    CU_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const bool ok = record_detection(inst, buf);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &graph);
    g.launches = inst->launches - l0;
    inst->launches = l0;
    if (ok && e == cudaSuccess && graph)
    {
      if (cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess)
      {
        g.exec = nullptr;
        cudaGetLastError();
      }
      g.prof = inst->profiling;
    }
    else
    {
      LOGW(TAG, "CUDA graph capture of the detection pipeline failed (%s), launching eagerly", cudaGetErrorName(e));
      cudaGetLastError();
      inst->use_graph = false;
    }
    if (graph)
      cudaGraphDestroy(graph);
  }
  g.uses++;
  if (g.exec)
  {
    CU_TRY(cudaGraphLaunch(g.exec, st));
    inst->launches += g.launches;
  }
  else if (!record_detection(inst, buf))
    return false;
  if (inst->profiling)
    inst->ev_detect_valid = true;
  CU_TRY(cudaEventRecord(inst->ev_detect_done, st));
  inst->detect_pending = true;
  inst->detect_buffer = buf;
  fb.uploaded = false;
  fb.norms_valid = false;
  return true;
}

bool detect_common(vksift_Instance inst, const uint8_t *host_image, const uint8_t *dev_image, uint32_t w, uint32_t h, uint32_t buf)
{
  /* A running matching pipeline is waited for first (vulkansift.c:325-327), and so is the previous detection of the lane
   * this buffer maps to (its scale space and scratch are about to be overwritten); detections on other lanes keep running. */
  wait_pipelines(inst, false, true);
  vksift_Instance lane = lane_of_buffer(inst, buf);
  wait_lane(lane);
  inst->last_lane = inst->lanes.empty() ? 0u : buf % (uint32_t)inst->lanes.size();
  /* vksift_prepareSiftMemoryForDetection (sift_memory.c:891-955) */
  if (lane->cur_w != w || lane->cur_h != h)
  {
    if (!set_resolution(lane, w, h))
      return false;
  }
  FeatureBuffer &fb = lane->buffers[buf];
  if (fb.cur_w != lane->cur_w || fb.cur_h != lane->cur_h)
    update_buffer_sections(lane, fb);
  const uint8_t *src = dev_image;
  bool direct = false;
  if (host_image)
  {
    /* The caller's buffer is only borrowed for the duration of the call (sift_memory.c:943 copies it to a staging buffer).
     * Page-locked caller memory is read by the copy engine directly and the call returns once that copy has completed,
     * after the pipeline has been enqueued behind it; pageable memory goes through the lane's pinned staging buffer. */
    cudaPointerAttributes attr;
    direct = (cudaPointerGetAttributes(&attr, host_image) == cudaSuccess && attr.type == cudaMemoryTypeHost);
    if (!direct)
    {
      cudaGetLastError();
      memcpy(lane->h_image, host_image, (size_t)w * h);
    }
    CU_TRY(cudaMemcpyAsync(lane->d_image, direct ? host_image : lane->h_image, (size_t)w * h, cudaMemcpyHostToDevice, lane->stream));
    if (direct)
      CU_TRY(cudaEventRecord(lane->ev_h2d, lane->stream));
    src = lane->d_image;
  }
  lane->crowded = false;
  for (vksift_Instance other : inst->lanes)
    if (other != lane && other->detect_pending)
      lane->crowded = true;
  const bool ok = enqueue_detection(lane, src, buf);
  if (direct)
    CU_TRY(cudaEventSynchronize(lane->ev_h2d));
  return ok;
}

/* the lane whose pending detection writes this buffer is waited for; a pending match that reads the buffer too */
void wait_buffer(vksift_Instance inst, uint32_t buf)
{
  vksift_Instance lane = lane_of_buffer(inst, buf);
  if (lane->detect_pending && lane->detect_buffer == buf)
    wait_lane(lane);
  if (inst->match_pending && (buf == inst->match_a || buf == inst->match_b))
    wait_pipelines(inst, false, true);
}

} // namespace

/* =========================================================================== */
extern "C"
{

  vksift_Config vksift_getDefaultConfig()
  {
    /* vulkansift.c:47-64 */
    vksift_Config c;
    memset(&c, 0, sizeof(c));
    c.input_image_max_size = 1920u * 1080u;
    c.sift_buffer_count = 2u;
    c.max_nb_sift_per_buffer = 100000u;
    c.use_input_upsampling = true;
    c.nb_octaves = 0;
    c.nb_scales_per_octave = 3u;
    c.input_image_blur_level = 0.5f;
    c.seed_scale_sigma = 1.6f;
    c.intensity_threshold = 0.04f;
    c.edge_threshold = 10.f;
    c.max_nb_orientation_per_keypoint = 4;
    c.descriptor_format = VKSIFT_DESCRIPTOR_FORMAT_UBC;
    c.gpu_device_index = -1;
    c.use_hardware_interpolated_blur = true;
    c.pyramid_precision_mode = VKSIFT_PYRAMID_PRECISION_FLOAT32;
    c.on_error_callback_function = default_error_callback;
    c.use_gpu_debug_functions = false;
    c.gpu_debug_external_window_info.context = NULL;
    c.gpu_debug_external_window_info.window = NULL;
    return c;
  }

  vksift_Result vksift_loadVulkan()
  {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
    {
      LOGE(TAG, "vksift_loadVulkan() failure: no CUDA device available (%s).", cudaGetErrorName(e));
      cudaGetLastError();
      return VKSIFT_VULKAN_ERROR;
    }
    bool any_sm100 = false;
    for (int i = 0; i < n; i++)
    {
      int major = 0;
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i);
      any_sm100 |= (major == 10);
    }
    if (!any_sm100)
    {
      LOGE(TAG, "vksift_loadVulkan() failure: this build only carries sm_100a (NVIDIA B200) kernels and no such device is visible.");
      return VKSIFT_VULKAN_ERROR;
    }
    g_api_loaded = true;
    LOGI(TAG, "vksift_loadVulkan() success (CUDA runtime, %d device(s))", n);
    return VKSIFT_SUCCESS;
  }

  void vksift_unloadVulkan() { g_api_loaded = false; }

  void vksift_getAvailableGPUs(uint32_t *gpu_count, VKSIFT_GPU_NAME *gpu_names)
  {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
      cudaGetLastError();
      n = 0;
    }
    if (gpu_names == NULL)
    {
      *gpu_count = (uint32_t)n;
      return;
    }
    for (uint32_t i = 0; i < *gpu_count && i < (uint32_t)n; i++)
    {
      cudaDeviceProp prop;
      memset(gpu_names[i], 0, sizeof(VKSIFT_GPU_NAME));
      if (cudaGetDeviceProperties(&prop, (int)i) == cudaSuccess)
        strncpy(gpu_names[i], prop.name, sizeof(VKSIFT_GPU_NAME) - 1);
    }
  }

  void vksift_setLogLevel(const vksift_LogLevel level)
  {
    switch (level)
    {
    case VKSIFT_NO_LOG:
    case VKSIFT_LOG_ERROR:
    case VKSIFT_LOG_WARNING:
    case VKSIFT_LOG_INFO:
    case VKSIFT_LOG_DEBUG:
      g_log_level = (int)level;
      break;
    default:
      LOGE(TAG, "vksift_LogLevel in vksift_setLogLevel() is not handled");
      break;
    }
  }

  vksift_Result vksift_createInstance(vksift_Instance *instance_ptr, const vksift_Config *config)
  {
    assert(instance_ptr != NULL);
    assert(*instance_ptr == NULL);
    assert(config != NULL);
    if (!g_api_loaded)
    {
      LOGE(TAG, "vksift_createInstance() failure: GPU API not available. vksift_loadVulkan() must be called before using this function.");
      return VKSIFT_VULKAN_ERROR;
    }
    if (!config_valid(config))
    {
      LOGE(TAG, "vksift_createInstance() failure: Invalid configuration detected.");
      return VKSIFT_INVALID_INPUT_ERROR;
    }
    int n = 0;
    cudaGetDeviceCount(&n);
    int dev = config->gpu_device_index;
    if (dev < 0)
    {
      /* automatic selection (vkenv/vulkan_device.c:394-494 scores devices): most SMs wins */
      int best = -1, best_sms = -1;
      for (int i = 0; i < n; i++)
      {
        int major = 0, sms = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, i);
        if (major == 10 && sms > best_sms)
        {
          best = i;
          best_sms = sms;
        }
      }
      dev = best;
    }
    if (dev < 0 || dev >= n)
    {
      LOGE(TAG, "vksift_createInstance() failure: GPU device index %d is not available (%d device(s)).", config->gpu_device_index, n);
      return VKSIFT_VULKAN_ERROR;
    }
    vksift_Instance inst = new vksift_Instance_T();
    inst->cfg = *config;
    inst->device = dev;
    *instance_ptr = inst;
    DeviceGuard g(dev);
    if (!create_resources(inst))
    {
      LOGE(TAG, "vksift_createInstance() failure: Failed to setup the required device objects");
      cudaGetLastError();
      destroy_instance(inst);
      *instance_ptr = NULL;
      return VKSIFT_VULKAN_ERROR;
    }
    if (config->use_gpu_debug_functions)
      LOGI(TAG, "use_gpu_debug_functions: NVTX ranges with the reference's marker region names are emitted around every stage (the debug "
                "presenter and frame delimiters have no CUDA counterpart).");
    LOGI(TAG, "vksift_createInstance() success");
    return VKSIFT_SUCCESS;
  }

  void vksift_destroyInstance(vksift_Instance *instance_ptr)
  {
    assert(instance_ptr != NULL);
    assert(*instance_ptr != NULL);
    destroy_instance(*instance_ptr);
    *instance_ptr = NULL;
  }

  bool vksift_isBufferAvailable(vksift_Instance inst, const uint32_t gpu_buffer_id)
  {
    /* vulkansift.c:295-313 */
    DeviceGuard g(inst->device);
    vksift_Instance lane = lane_of_buffer(inst, gpu_buffer_id);
    if (lane->detect_pending && gpu_buffer_id == lane->detect_buffer)
    {
      if (cudaEventQuery(lane->ev_detect_done) == cudaErrorNotReady)
        return false;
      lane->detect_pending = false;
    }
    if (inst->match_pending && (gpu_buffer_id == inst->match_a || gpu_buffer_id == inst->match_b))
    {
      if (cudaEventQuery(inst->ev_match_done) == cudaErrorNotReady)
        return false;
      inst->match_pending = false;
    }
    return true;
  }

  void vksift_detectFeatures(vksift_Instance inst, const uint8_t *image_data, const uint32_t image_width, const uint32_t image_height,
                             const uint32_t gpu_buffer_id)
  {
    if (!buffer_idx_valid(inst, gpu_buffer_id) || !resolution_valid(inst, image_width, image_height))
    {
      LOGE(TAG, "vksift_detectFeatures() error: invalid input.");
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return;
    }
    bool ok;
    {
      DeviceGuard g(inst->device);
      ok = detect_common(inst, image_data, nullptr, image_width, image_height, gpu_buffer_id);
    }
    if (!ok)
    {
      LOGE(TAG, "vksift_detectFeatures() error: Failed to start the detection pipeline.");
      inst->cfg.on_error_callback_function(VKSIFT_VULKAN_ERROR);
    }
  }

  void vksiftx_detectFeaturesDevice(vksift_Instance inst, const void *d_image, const uint32_t image_width, const uint32_t image_height,
                                    const uint32_t gpu_buffer_id)
  {
    if (!buffer_idx_valid(inst, gpu_buffer_id) || !resolution_valid(inst, image_width, image_height) || d_image == NULL)
    {
      LOGE(TAG, "vksiftx_detectFeaturesDevice() error: invalid input.");
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return;
    }
    bool ok;
    {
      DeviceGuard g(inst->device);
      ok = detect_common(inst, nullptr, (const uint8_t *)d_image, image_width, image_height, gpu_buffer_id);
    }
    if (!ok)
    {
      LOGE(TAG, "vksiftx_detectFeaturesDevice() error: Failed to start the detection pipeline.");
      inst->cfg.on_error_callback_function(VKSIFT_VULKAN_ERROR);
    }
  }

  uint32_t vksift_getFeaturesNumber(vksift_Instance inst, const uint32_t gpu_buffer_id)
  {
    if (!buffer_idx_valid(inst, gpu_buffer_id))
    {
      LOGE(TAG, "vksift_getFeaturesNumber() error: invalid input.");
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return 0;
    }
    DeviceGuard g(inst->device);
    wait_buffer(inst, gpu_buffer_id);
    return buffer_count(inst, gpu_buffer_id, true);
  }

  void vksift_downloadFeatures(vksift_Instance inst, vksift_Feature *feats_ptr, const uint32_t gpu_buffer_id)
  {
    if (!buffer_idx_valid(inst, gpu_buffer_id))
    {
      LOGE(TAG, "vksift_downloadFeatures() error: invalid input.");
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return;
    }
    bool ok = true;
    {
      DeviceGuard g(inst->device);
      wait_buffer(inst, gpu_buffer_id);
      FeatureBuffer &fb = inst->buffers[gpu_buffer_id];
      const uint32_t n = buffer_count(inst, gpu_buffer_id, false);
      if (n > 0)
      {
        /* on the stream and staging area of the buffer's lane, so that the transfer does not queue behind another lane */
        vksift_Instance lane = lane_of_buffer(inst, gpu_buffer_id);
        auto run = [&]() -> bool {
          CU_TRY(launch_pack_aos(fb.heads, fb.desc, n, lane->d_aos, lane->stream));
          inst->launches++;
          CU_TRY(cudaMemcpyAsync(feats_ptr, lane->d_aos, sizeof(vksift_Feature) * (size_t)n, cudaMemcpyDeviceToHost, lane->stream));
          CU_TRY(cudaStreamSynchronize(lane->stream));
          return true;
        };
        ok = run();
      }
    }
    if (!ok)
    {
      LOGE(TAG, "vksift_downloadFeatures() error when downloading detection results.");
      inst->cfg.on_error_callback_function(VKSIFT_VULKAN_ERROR);
    }
  }

  void vksift_uploadFeatures(vksift_Instance inst, const vksift_Feature *feats_ptr, const uint32_t nb_feats, const uint32_t gpu_buffer_id)
  {
    if (!buffer_idx_valid(inst, gpu_buffer_id) || nb_feats > inst->cfg.max_nb_sift_per_buffer)
    {
      if (nb_feats > inst->cfg.max_nb_sift_per_buffer)
        LOGE(TAG, "Provided features count (%u) is greater than the configured maximum number of features per GPU buffer size (%u).", nb_feats,
             inst->cfg.max_nb_sift_per_buffer);
      LOGE(TAG, "vksift_uploadFeatures() error: invalid input.");
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return;
    }
    bool ok = true;
    {
      DeviceGuard g(inst->device);
      if (!vksift_isBufferAvailable(inst, gpu_buffer_id))
        wait_pipelines(inst, true, true);
      FeatureBuffer &fb = inst->buffers[gpu_buffer_id];
      auto run = [&]() -> bool {
        if (nb_feats > 0)
        {
          CU_TRY(cudaMemcpyAsync(inst->d_aos, feats_ptr, sizeof(vksift_Feature) * (size_t)nb_feats, cudaMemcpyHostToDevice, inst->stream));
          CU_TRY(launch_unpack_aos(inst->d_aos, nb_feats, fb.heads, fb.desc, inst->stream));
          inst->launches++;
        }
        CU_TRY(cudaStreamSynchronize(inst->stream)); /* blocking transfer, feats_ptr is free after return */
        return true;
      };
      ok = run();
      if (ok)
      {
        fb.uploaded = true;
        fb.n_uploaded = nb_feats;
      }
      fb.norms_valid = false;
    }
    if (!ok)
    {
      LOGE(TAG, "vksift_uploadFeatures() error when uploading SIFT features to GPU memory.");
      inst->cfg.on_error_callback_function(VKSIFT_VULKAN_ERROR);
    }
  }

  void vksiftx_uploadDescriptorsDevice(vksift_Instance inst, const void *d_descriptors, const uint32_t nb_feats, const uint32_t gpu_buffer_id)
  {
    if (!buffer_idx_valid(inst, gpu_buffer_id) || nb_feats > inst->cfg.max_nb_sift_per_buffer || (nb_feats > 0 && d_descriptors == NULL))
    {
      LOGE(TAG, "vksiftx_uploadDescriptorsDevice() error: invalid input.");
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return;
    }
    bool ok = true;
    {
      DeviceGuard g(inst->device);
      if (!vksift_isBufferAvailable(inst, gpu_buffer_id))
        wait_pipelines(inst, true, true);
      FeatureBuffer &fb = inst->buffers[gpu_buffer_id];
      auto run = [&]() -> bool {
        if (nb_feats > 0)
        {
          CU_TRY(cudaMemcpyAsync(fb.desc, d_descriptors, 128 * (size_t)nb_feats, cudaMemcpyDeviceToDevice, inst->stream));
          CU_TRY(cudaMemsetAsync(fb.heads, 0, sizeof(FeatHead) * (size_t)nb_feats, inst->stream));
        }
        CU_TRY(cudaStreamSynchronize(inst->stream));
        return true;
      };
      ok = run();
      if (ok)
      {
        fb.uploaded = true;
        fb.n_uploaded = nb_feats;
      }
      fb.norms_valid = false;
    }
    if (!ok)
    {
      LOGE(TAG, "vksiftx_uploadDescriptorsDevice() error when copying descriptors.");
      inst->cfg.on_error_callback_function(VKSIFT_VULKAN_ERROR);
    }
  }

  /* *fresh is set when the norms had to be computed now (the search that follows must not overlap earlier work then) */
  static bool ensure_norms(vksift_Instance inst, FeatureBuffer &fb, uint32_t n, bool *fresh = nullptr)
  {
    if (fb.norms_valid && fb.norms_n == n)
      return true;
    if (fresh)
      *fresh = true;
    CU_TRY(launch_norms(fb.desc, n, fb.norm_plain, fb.norm_packed, inst->stream));
    inst->launches++;
    fb.norms_valid = true;
    fb.norms_n = n;
    return true;
  }

  /* shared by vksift_matchFeatures and vksiftx_matchFeaturesAgainstDevice: B is a feature buffer or caller-owned descriptors */
  static void match_common(vksift_Instance inst, const uint32_t gpu_buffer_id_A, const uint32_t gpu_buffer_id_B, const uint8_t *d_desc_b,
                           const uint32_t nb_ext, const char *fn)
  {
    bool ok = true, invalid = false;
    {
      DeviceGuard g(inst->device);
      /* vulkansift.c:427-428 waits for the detection AND the previous matching pipeline (one command buffer, one fence).  The
       * detections are waited for here too (the feature counts come from them); a previous search is ordered before this one
       * by the stream instead of by the host: searches enqueued back to back run back to back on the GPU, the next one's MMA
       * phase overlapping the previous one's merge.  Results and every later blocking rule (downloads, detections into a
       * buffer that a search still reads) are unchanged: they wait for the LAST search, which completes after all others. */
      wait_pipelines(inst, true, false);
      const uint32_t na = buffer_count(inst, gpu_buffer_id_A, false);
      const uint32_t nb = d_desc_b ? nb_ext : buffer_count(inst, gpu_buffer_id_B, false);
      if (nb < 2 && na > 0)
      {
        /* the shader reads B[0] and B[1] unconditionally (Get2NearestNeighbors.comp:66-67), SURVEY B-D13 */
        LOGE(TAG, "%s() error: buffer B holds %u feature(s), the 2-nearest-neighbour search needs at least 2.", fn, nb);
        invalid = true;
      }
      else
      {
        inst->nb_matches = na; /* sift_memory.c:1056 */
        FeatureBuffer &A = inst->buffers[gpu_buffer_id_A];
        const uint8_t *b_desc = d_desc_b ? d_desc_b : inst->buffers[gpu_buffer_id_B].desc;
        auto run = [&]() -> bool {
          MarkerRegion mr(inst, "Matching");
          const bool prof = inst->profiling;
          if (prof)
            CU_TRY(cudaEventRecord(inst->ev[EV_M0], inst->stream));
          bool fresh = false;
          if (na > 0 && !ensure_norms(inst, A, na, &fresh))
            return false;
          const uint32_t *nb_cached = nullptr;
          if (na > 0 && !d_desc_b)
          {
            FeatureBuffer &B = inst->buffers[gpu_buffer_id_B];
            if (!ensure_norms(inst, B, nb, &fresh))
              return false;
            nb_cached = B.norm_packed;
          }
          /* profiling brackets the kernels with events: no overlap then, the stage times stay per search */
          CU_TRY(launch_match(inst->match_ws, inst->matcher_impl, A.desc, na, A.norm_plain, b_desc, nb, nb_cached, inst->d_matches, inst->stream,
                              prof ? inst->ev[EV_M1] : nullptr, !fresh && !prof, &inst->launches));
          if (prof)
          {
            if (na == 0)
              CU_TRY(cudaEventRecord(inst->ev[EV_M1], inst->stream));
            CU_TRY(cudaEventRecord(inst->ev[EV_M2], inst->stream));
            inst->ev_match_valid = true;
          }
          CU_TRY(cudaEventRecord(inst->ev_match_done, inst->stream));
          return true;
        };
        ok = run();
        inst->match_pending = ok;
        inst->match_a = gpu_buffer_id_A;
        inst->match_b = d_desc_b ? gpu_buffer_id_A : gpu_buffer_id_B;
      }
    }
    if (invalid)
    {
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return;
    }
    if (!ok)
    {
      LOGE(TAG, "%s() error: Failed to start the matching pipeline.", fn);
      inst->cfg.on_error_callback_function(VKSIFT_VULKAN_ERROR);
    }
  }

  void vksift_matchFeatures(vksift_Instance inst, const uint32_t gpu_buffer_id_A, const uint32_t gpu_buffer_id_B)
  {
    if (!buffer_idx_valid(inst, gpu_buffer_id_A) || !buffer_idx_valid(inst, gpu_buffer_id_B))
    {
      LOGE(TAG, "vksift_matchFeatures() error: invalid input.");
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return;
    }
    match_common(inst, gpu_buffer_id_A, gpu_buffer_id_B, nullptr, 0, "vksift_matchFeatures");
  }

  void vksiftx_matchFeaturesAgainstDevice(vksift_Instance inst, const uint32_t gpu_buffer_id_A, const void *d_descriptors_B, const uint32_t nb_feats_B)
  {
    if (!buffer_idx_valid(inst, gpu_buffer_id_A) || d_descriptors_B == NULL || ((uintptr_t)d_descriptors_B & 127u) != 0 ||
        nb_feats_B > inst->cfg.max_nb_sift_per_buffer)
    {
      LOGE(TAG, "vksiftx_matchFeaturesAgainstDevice() error: invalid input (B must be a 128-byte aligned device pointer holding at most "
                "max_nb_sift_per_buffer descriptors).");
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return;
    }
    match_common(inst, gpu_buffer_id_A, 0, (const uint8_t *)d_descriptors_B, nb_feats_B, "vksiftx_matchFeaturesAgainstDevice");
  }

  /* zero_padded: the caller guarantees zeros between every block's count and the largest count rounded up to 128 rows (the
   * exchange does): all blocks are then searched by ONE launch of the tensor-core kernel instead of one per block */
  static void match_blocks_common(vksift_Instance inst, const uint32_t gpu_buffer_id_A, const void *d_blocks, const uint32_t n_blocks,
                                  const uint64_t block_stride_bytes, const uint32_t *counts, const uint32_t skip_block, const bool zero_padded)
  {
    if (!buffer_idx_valid(inst, gpu_buffer_id_A) || d_blocks == NULL || counts == NULL || n_blocks == 0 || n_blocks > 1024u ||
        ((uintptr_t)d_blocks & 127u) != 0 || (block_stride_bytes & 127u) != 0)
    {
      LOGE(TAG, "vksiftx_matchFeaturesAgainstBlocks() error: invalid input (blocks must be 128-byte aligned device memory, at most 1024 blocks).");
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return;
    }
    for (uint32_t j = 0; j < n_blocks; j++)
      if (counts[j] > inst->cfg.max_nb_sift_per_buffer || (uint64_t)counts[j] * 128u > block_stride_bytes)
      {
        LOGE(TAG, "vksiftx_matchFeaturesAgainstBlocks() error: block %u holds %u descriptors (more than a block or max_nb_sift_per_buffer).", j, counts[j]);
        inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
        return;
      }
    bool ok = true;
    {
      DeviceGuard g(inst->device);
      wait_pipelines(inst, true, true);
      const uint32_t na = buffer_count(inst, gpu_buffer_id_A, false);
      FeatureBuffer &A = inst->buffers[gpu_buffer_id_A];
      const size_t maxf = inst->cfg.max_nb_sift_per_buffer;
      auto run = [&]() -> bool {
        if (n_blocks > inst->blocks_cap)
        {
          if (inst->d_matches_blocks)
            CU_TRY(cudaFree(inst->d_matches_blocks));
          inst->d_matches_blocks = nullptr;
          inst->blocks_cap = 0;
          CU_TRY(cudaMalloc(&inst->d_matches_blocks, sizeof(vksift_Match_2NN) * maxf * n_blocks));
          inst->blocks_cap = n_blocks;
        }
        inst->blocks_n = n_blocks;
        inst->blocks_na = na;
        if (na == 0)
          return true;
        if (!ensure_norms(inst, A, na))
          return false;
        /* |b|^2 of every block in one launch (blocks whose stride is a whole number of 128-row tiles, at most 64 of them) */
        const uint32_t stride_rows = (uint32_t)(block_stride_bytes / 128u);
        const bool batched_norms = n_blocks <= VKS_MAX_MATCH_BLOCKS && (stride_rows % 128u) == 0;
        if (batched_norms)
        {
          const size_t need = (size_t)n_blocks * stride_rows;
          if (need > inst->block_norms_cap)
          {
            if (inst->d_block_norms)
              CU_TRY(cudaFree(inst->d_block_norms));
            inst->d_block_norms = nullptr;
            inst->block_norms_cap = 0;
            CU_TRY(cudaMalloc(&inst->d_block_norms, sizeof(uint32_t) * need));
            inst->block_norms_cap = need;
          }
          MatchBlockCounts bc;
          for (uint32_t j = 0; j < VKS_MAX_MATCH_BLOCKS; j++)
            bc.n[j] = (j < n_blocks && j != skip_block) ? counts[j] : 0u;
          CU_TRY(launch_norms_blocks((const uint8_t *)d_blocks, bc, n_blocks, stride_rows, inst->d_block_norms, inst->stream));
          inst->launches++;
        }
        if (batched_norms && zero_padded && inst->matcher_impl == 0)
        {
          uint32_t blk[VKS_MAX_MATCH_BLOCKS], cnt[VKS_MAX_MATCH_BLOCKS], n_groups = 0;
          for (uint32_t j = 0; j < n_blocks; j++)
          {
            if (j == skip_block || counts[j] < 2)
            {
              CU_TRY(cudaMemsetAsync(inst->d_matches_blocks + (size_t)j * maxf, 0, sizeof(vksift_Match_2NN) * (size_t)na, inst->stream));
              continue;
            }
            blk[n_groups] = j;
            cnt[n_groups++] = counts[j];
          }
          CU_TRY(launch_match_blocks(inst->match_ws, A.desc, na, A.norm_plain, (const uint8_t *)d_blocks, inst->d_block_norms, stride_rows, blk, cnt,
                                     n_groups, inst->d_matches_blocks, (uint32_t)maxf, inst->stream, false, &inst->launches));
          CU_TRY(cudaEventRecord(inst->ev_match_done, inst->stream));
          return true;
        }
        bool first_search = true;
        for (uint32_t j = 0; j < n_blocks; j++)
        {
          vksift_Match_2NN *out = inst->d_matches_blocks + (size_t)j * maxf;
          if (j == skip_block || counts[j] < 2)
          {
            CU_TRY(cudaMemsetAsync(out, 0, sizeof(vksift_Match_2NN) * (size_t)na, inst->stream));
            continue;
          }
          const uint8_t *b = (const uint8_t *)d_blocks + (size_t)j * block_stride_bytes;
          /* the searches share the workspace (B-side norms, partial keys): they are ordered by the stream */
          CU_TRY(launch_match(inst->match_ws, inst->matcher_impl, A.desc, na, A.norm_plain, b, counts[j],
                              batched_norms ? inst->d_block_norms + (size_t)j * stride_rows : nullptr, out, inst->stream, nullptr, !first_search,
                              &inst->launches));
          first_search = false; /* the first search follows the norm launches, the others only earlier searches */
        }
        CU_TRY(cudaEventRecord(inst->ev_match_done, inst->stream));
        return true;
      };
      ok = run();
      inst->match_pending = ok && na > 0;
      inst->match_a = inst->match_b = gpu_buffer_id_A;
    }
    if (!ok)
    {
      LOGE(TAG, "vksiftx_matchFeaturesAgainstBlocks() error: Failed to start the matching pipeline.");
      inst->cfg.on_error_callback_function(VKSIFT_VULKAN_ERROR);
    }
  }

  void vksiftx_matchFeaturesAgainstBlocks(vksift_Instance inst, const uint32_t gpu_buffer_id_A, const void *d_blocks, const uint32_t n_blocks,
                                          const uint64_t block_stride_bytes, const uint32_t *counts, const uint32_t skip_block)
  {
    match_blocks_common(inst, gpu_buffer_id_A, d_blocks, n_blocks, block_stride_bytes, counts, skip_block, false);
  }

  void vksiftx_downloadMatchesBlocks(vksift_Instance inst, vksift_Match_2NN *matches, const uint32_t n_blocks)
  {
    if (matches == NULL || n_blocks != inst->blocks_n)
    {
      LOGE(TAG, "vksiftx_downloadMatchesBlocks() error: invalid input (%u blocks asked, the last search had %u).", n_blocks, inst->blocks_n);
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return;
    }
    bool ok = true;
    {
      DeviceGuard g(inst->device);
      wait_pipelines(inst, false, true);
      if (inst->blocks_na > 0)
      {
        auto run = [&]() -> bool {
          /* [n_blocks][na] on the host from [n_blocks][max] on the device: one strided copy */
          CU_TRY(cudaMemcpy2DAsync(matches, sizeof(vksift_Match_2NN) * (size_t)inst->blocks_na, inst->d_matches_blocks,
                                   sizeof(vksift_Match_2NN) * (size_t)inst->cfg.max_nb_sift_per_buffer, sizeof(vksift_Match_2NN) * (size_t)inst->blocks_na,
                                   n_blocks, cudaMemcpyDeviceToHost, inst->stream));
          CU_TRY(cudaStreamSynchronize(inst->stream));
          return true;
        };
        ok = run();
      }
    }
    if (!ok)
    {
      LOGE(TAG, "vksiftx_downloadMatchesBlocks() error when downloading SIFT matches from GPU memory.");
      inst->cfg.on_error_callback_function(VKSIFT_VULKAN_ERROR);
    }
  }

  /* ---- descriptor exchange between the GPUs of a node over NVLink peer memory (exchange.cu) ---- */
  bool vksiftx_exchangeCreate(vksift_Instance inst, const uint32_t rank, const uint32_t world_size, const uint32_t slot_rows, void *handle_out)
  {
    if (handle_out == NULL || inst->exchange != nullptr || slot_rows == 0)
    {
      LOGE(TAG, "vksiftx_exchangeCreate() error: invalid input (one exchange per instance, slots of at least one row).");
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return false;
    }
    DeviceGuard g(inst->device);
    const uint32_t rows = (slot_rows + 127u) & ~127u;
    cudaError_t e = exchange_create(&inst->exchange, (int)rank, (int)world_size, rows, handle_out);
    if (e != cudaSuccess)
    {
      LOGE(TAG, "vksiftx_exchangeCreate() error: %s (rank %u of %u, at most %d ranks).", cudaGetErrorName(e), rank, world_size, VKS_MAX_PEERS);
      inst->cfg.on_error_callback_function(e == cudaErrorInvalidValue ? VKSIFT_INVALID_INPUT_ERROR : VKSIFT_VULKAN_ERROR);
      return false;
    }
    return true;
  }

  bool vksiftx_exchangeConnect(vksift_Instance inst, const void *handles)
  {
    if (handles == NULL || inst->exchange == nullptr)
    {
      LOGE(TAG, "vksiftx_exchangeConnect() error: invalid input (no exchange created).");
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return false;
    }
    DeviceGuard g(inst->device);
    cudaError_t e = exchange_connect(inst->exchange, handles);
    if (e != cudaSuccess)
    {
      LOGE(TAG, "vksiftx_exchangeConnect() error: cannot map a peer's exchange memory (%s).", cudaGetErrorName(e));
      inst->cfg.on_error_callback_function(VKSIFT_VULKAN_ERROR);
      return false;
    }
    return true;
  }

  bool vksiftx_exchangeAllGather(vksift_Instance inst, const uint32_t gpu_buffer_id, uint32_t *counts, void **d_blocks, uint64_t *block_stride_bytes)
  {
    if (!buffer_idx_valid(inst, gpu_buffer_id) || inst->exchange == nullptr || counts == NULL)
    {
      LOGE(TAG, "vksiftx_exchangeAllGather() error: invalid input (buffer index, or no exchange created).");
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return false;
    }
    bool ok = true;
    uint32_t late = 0;
    {
      DeviceGuard g(inst->device);
      wait_pipelines(inst, true, false); /* the feature count comes from the detection; earlier searches are ordered by the stream */
      const uint32_t n = buffer_count(inst, gpu_buffer_id, false);
      FeatureBuffer &fb = inst->buffers[gpu_buffer_id];
      auto run = [&]() -> bool {
        MarkerRegion mr(inst, "DescriptorExchange");
        if (n > exchange_slot_rows(inst->exchange))
        {
          LOGE(TAG, "vksiftx_exchangeAllGather() error: the buffer holds %u features, an exchange slot %u rows.", n, exchange_slot_rows(inst->exchange));
          return false;
        }
        CU_TRY(exchange_allgather(inst->exchange, fb.desc, n, inst->stream, &inst->launches));
        CU_TRY(cudaStreamSynchronize(inst->stream));
        return true;
      };
      ok = run();
      if (ok)
      {
        late = exchange_timeout_mask(inst->exchange);
        const uint32_t *hc = exchange_host_counts(inst->exchange);
        for (int p = 0; p < exchange_world(inst->exchange); p++)
          counts[p] = hc[p];
        uint64_t stride = 0;
        const uint8_t *blocks = exchange_blocks(inst->exchange, &stride);
        if (d_blocks)
          *d_blocks = (void *)blocks;
        if (block_stride_bytes)
          *block_stride_bytes = stride;
      }
    }
    if (!ok || late != 0)
    {
      if (late != 0)
        LOGE(TAG, "vksiftx_exchangeAllGather() error: peers (mask 0x%x) did not publish their descriptors within two seconds.", late);
      else
        LOGE(TAG, "vksiftx_exchangeAllGather() error: Failed to run the descriptor exchange.");
      inst->cfg.on_error_callback_function(VKSIFT_VULKAN_ERROR);
      return false;
    }
    return true;
  }

  bool vksiftx_exchangeMatchAllPeers(vksift_Instance inst, const uint32_t gpu_buffer_id, uint32_t *counts)
  {
    void *blocks = nullptr;
    uint64_t stride = 0;
    if (!vksiftx_exchangeAllGather(inst, gpu_buffer_id, counts, &blocks, &stride))
      return false;
    match_blocks_common(inst, gpu_buffer_id, blocks, (uint32_t)exchange_world(inst->exchange), stride, counts, (uint32_t)exchange_rank(inst->exchange),
                        true);
    return true;
  }

  void vksiftx_exchangeDestroy(vksift_Instance inst)
  {
    DeviceGuard g(inst->device);
    cudaDeviceSynchronize();
    exchange_destroy(inst->exchange);
    inst->exchange = nullptr;
  }

  uint32_t vksift_getMatchesNumber(vksift_Instance inst) { return inst->nb_matches; }

  void vksift_downloadMatches(vksift_Instance inst, vksift_Match_2NN *matches)
  {
    bool ok = true;
    {
      DeviceGuard g(inst->device);
      wait_pipelines(inst, false, true);
      if (inst->nb_matches > 0)
      {
        auto run = [&]() -> bool {
          CU_TRY(cudaMemcpyAsync(matches, inst->d_matches, sizeof(vksift_Match_2NN) * (size_t)inst->nb_matches, cudaMemcpyDeviceToHost, inst->stream));
          CU_TRY(cudaStreamSynchronize(inst->stream));
          return true;
        };
        ok = run();
      }
    }
    if (!ok)
    {
      LOGE(TAG, "vksift_downloadMatches() error when downloading SIFT matches from GPU memory.");
      inst->cfg.on_error_callback_function(VKSIFT_VULKAN_ERROR);
    }
  }

  /* ---- scale-space access (vulkansift.c:463-519) --------------------------- */
  /* the scale space shown is the one of the most recent detection (every lane starts at the default resolution) */
  uint8_t vksift_getScaleSpaceNbOctaves(vksift_Instance inst) { return (uint8_t)last_lane(inst)->pyr.n_oct; }

  void vksift_getScaleSpaceOctaveResolution(vksift_Instance inst, const uint8_t octave, uint32_t *octave_images_width, uint32_t *octave_images_height)
  {
    const Pyramid &pyr = last_lane(inst)->pyr;
    if (octave >= pyr.n_oct)
    {
      LOGE(TAG, "vksift_getScaleSpaceOctaveResolution() error: invalid input. Requested octave idx is %d but the current number of octave is %d", octave,
           pyr.n_oct);
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return;
    }
    *octave_images_width = pyr.w[octave];
    *octave_images_height = pyr.h[octave];
  }

  static void download_layer(vksift_Instance inst, const uint8_t octave, const uint8_t scale, bool dog, float *out, const char *fn)
  {
    const uint32_t nscales = inst->cfg.nb_scales_per_octave + (dog ? 2u : 3u);
    const Pyramid &p = last_lane(inst)->pyr;
    if (octave >= p.n_oct || scale >= nscales)
    {
      if (octave >= p.n_oct)
        LOGE(TAG, "Requested octave idx is %d but the current number of octaves is %d", octave, p.n_oct);
      else
        LOGE(TAG, "Requested scale idx is %d but the number of %s scales is %d", scale, dog ? "DoG" : "blurred", nscales);
      LOGE(TAG, "%s() error: invalid input.", fn);
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return;
    }
    bool ok = true;
    {
      DeviceGuard g(inst->device);
      wait_pipelines(inst, true, false);
      const size_t layer = (size_t)p.pitch[octave] * p.h[octave] * (size_t)p.es;
      const char *src = (const char *)(dog ? p.D[octave] : p.G[octave]) + layer * scale;
      auto run = [&]() -> bool {
        if (p.es == 2)
        {
          /* binary16 layers are widened on the device (exact), the caller always receives floats (vulkansift.h:98-100) */
          float *tmp = nullptr;
          CU_TRY(cudaMalloc(&tmp, sizeof(float) * (size_t)p.w[octave] * p.h[octave]));
          cudaError_t e = launch_widen_layer(src, (int)p.w[octave], (int)p.h[octave], (int)p.pitch[octave], tmp, inst->stream);
          if (e == cudaSuccess)
            e = cudaMemcpyAsync(out, tmp, sizeof(float) * (size_t)p.w[octave] * p.h[octave], cudaMemcpyDeviceToHost, inst->stream);
          if (e == cudaSuccess)
            e = cudaStreamSynchronize(inst->stream);
          cudaFree(tmp);
          CU_TRY(e);
          return true;
        }
        CU_TRY(cudaMemcpy2DAsync(out, sizeof(float) * p.w[octave], src, sizeof(float) * p.pitch[octave], sizeof(float) * p.w[octave], p.h[octave],
                                 cudaMemcpyDeviceToHost, inst->stream));
        CU_TRY(cudaStreamSynchronize(inst->stream));
        return true;
      };
      ok = run();
    }
    if (!ok)
    {
      LOGE(TAG, "%s() error when downloading a pyramid image from GPU memory.", fn);
      inst->cfg.on_error_callback_function(VKSIFT_VULKAN_ERROR);
    }
  }

  void vksift_downloadScaleSpaceImage(vksift_Instance inst, const uint8_t octave, const uint8_t scale, float *blurred_image)
  {
    download_layer(inst, octave, scale, false, blurred_image, "vksift_downloadScaleSpaceImage");
  }
  void vksift_downloadDoGImage(vksift_Instance inst, const uint8_t octave, const uint8_t scale, float *dog_image)
  {
    download_layer(inst, octave, scale, true, dog_image, "vksift_downloadDoGImage");
  }

  void vksift_presentDebugFrame(vksift_Instance inst)
  {
    (void)inst;
    LOGW(TAG, "vksift_presentDebugFrame() was called but this build has no debug presenter (use ncu / compute-sanitizer instead).");
  }

  /* ---- extensions (include/vksift_b200_ext.h) ------------------------------ */
  const char *vksiftx_getVersionString() { return "vulkansift-b200 0.1 sm_100a"; }
  int32_t vksiftx_getDeviceIndex(vksift_Instance inst) { return inst->device; }
  void *vksiftx_getStream(vksift_Instance inst) { return (void *)inst->stream; }

  void vksiftx_waitIdle(vksift_Instance inst)
  {
    DeviceGuard g(inst->device);
    wait_pipelines(inst, true, true);
    cudaStreamSynchronize(inst->stream);
  }

  uint32_t vksiftx_getLaneCount(vksift_Instance inst) { return (uint32_t)inst->lanes.size(); }

  void vksiftx_joinLanes(vksift_Instance inst)
  {
    /* device-side join: work enqueued on the instance stream after this call starts after every enqueued detection */
    DeviceGuard g(inst->device);
    for (size_t k = 1; k < inst->lanes.size(); k++)
      if (inst->lanes[k]->detect_pending)
        cudaStreamWaitEvent(inst->stream, inst->lanes[k]->ev_detect_done, 0);
  }

  void vksiftx_getBufferDeviceView(vksift_Instance inst, const uint32_t gpu_buffer_id, uint32_t *nb_feats, void **d_descriptors, void **d_heads)
  {
    if (!buffer_idx_valid(inst, gpu_buffer_id))
    {
      LOGE(TAG, "vksiftx_getBufferDeviceView() error: invalid input.");
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return;
    }
    DeviceGuard g(inst->device);
    if (!vksift_isBufferAvailable(inst, gpu_buffer_id))
      wait_pipelines(inst, true, true);
    if (nb_feats)
      *nb_feats = buffer_count(inst, gpu_buffer_id, false);
    if (d_descriptors)
      *d_descriptors = inst->buffers[gpu_buffer_id].desc;
    if (d_heads)
      *d_heads = inst->buffers[gpu_buffer_id].heads;
  }

  uint32_t vksiftx_copyDescriptorsToDevice(vksift_Instance inst, const uint32_t gpu_buffer_id, void *d_dst, const uint32_t capacity)
  {
    if (!buffer_idx_valid(inst, gpu_buffer_id) || d_dst == NULL)
    {
      LOGE(TAG, "vksiftx_copyDescriptorsToDevice() error: invalid input.");
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return 0;
    }
    bool ok = true, too_small = false;
    uint32_t n = 0;
    {
      DeviceGuard g(inst->device);
      if (!vksift_isBufferAvailable(inst, gpu_buffer_id))
        wait_pipelines(inst, true, true);
      n = buffer_count(inst, gpu_buffer_id, false);
      if (n > capacity)
        too_small = true;
      else
      {
        auto run = [&]() -> bool {
          if (n > 0)
            CU_TRY(cudaMemcpyAsync(d_dst, inst->buffers[gpu_buffer_id].desc, 128 * (size_t)n, cudaMemcpyDeviceToDevice, inst->stream));
          if (capacity > n)
            CU_TRY(cudaMemsetAsync((uint8_t *)d_dst + 128 * (size_t)n, 0, 128 * (size_t)(capacity - n), inst->stream));
          CU_TRY(cudaStreamSynchronize(inst->stream));
          return true;
        };
        ok = run();
      }
    }
    if (too_small)
    {
      LOGE(TAG, "vksiftx_copyDescriptorsToDevice() error: destination holds %u descriptors, the buffer has %u.", capacity, n);
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return 0;
    }
    if (!ok)
    {
      LOGE(TAG, "vksiftx_copyDescriptorsToDevice() error when copying descriptors.");
      inst->cfg.on_error_callback_function(VKSIFT_VULKAN_ERROR);
      return 0;
    }
    return n;
  }

  void *vksiftx_getMatchesDevice(vksift_Instance inst) { return inst->d_matches; }

  void vksiftx_setProfiling(vksift_Instance inst, const bool enabled)
  {
    const char *t = getenv("VKSIFT_TRACE");
    for (vksift_Instance lane : inst->lanes)
    {
      lane->profiling = enabled;
      if (t && t[0] == '1')
        lane->trace = enabled;
      lane->trace_dump_stderr = (t && t[0] == '1');
    }
  }

  uint32_t vksiftx_matchFeaturesCrossChecked(vksift_Instance inst, const uint32_t gpu_buffer_id_A, const uint32_t gpu_buffer_id_B,
                                             const float lowe_ratio, uint32_t *pairs, const uint32_t capacity)
  {
    if (!buffer_idx_valid(inst, gpu_buffer_id_A) || !buffer_idx_valid(inst, gpu_buffer_id_B) || (pairs == NULL && capacity > 0))
    {
      LOGE(TAG, "vksiftx_matchFeaturesCrossChecked() error: invalid input.");
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return 0;
    }
    bool ok = true, invalid = false;
    uint32_t n_pairs = 0;
    {
      DeviceGuard g(inst->device);
      wait_pipelines(inst, true, true);
      const uint32_t na = buffer_count(inst, gpu_buffer_id_A, false);
      const uint32_t nb = buffer_count(inst, gpu_buffer_id_B, false);
      if (na < 2 || nb < 2)
      {
        LOGE(TAG, "vksiftx_matchFeaturesCrossChecked() error: both buffers need at least 2 features (%u, %u).", na, nb);
        invalid = true;
      }
      else
      {
        FeatureBuffer &A = inst->buffers[gpu_buffer_id_A], &B = inst->buffers[gpu_buffer_id_B];
        const size_t maxf = inst->cfg.max_nb_sift_per_buffer;
        auto run = [&]() -> bool {
          bool fresh = false;
          if (!ensure_norms(inst, A, na, &fresh) || !ensure_norms(inst, B, nb, &fresh))
            return false;
          CU_TRY(launch_match(inst->match_ws, inst->matcher_impl, B.desc, nb, B.norm_plain, A.desc, na, A.norm_packed, inst->d_matches_rev, inst->stream,
                              nullptr, !fresh, &inst->launches));
          CU_TRY(launch_match(inst->match_ws, inst->matcher_impl, A.desc, na, A.norm_plain, B.desc, nb, B.norm_packed, inst->d_matches, inst->stream,
                              nullptr, true, &inst->launches));
          CU_TRY(launch_match_filter(inst->d_matches, na, inst->d_matches_rev, nb, lowe_ratio, inst->d_pairs, (uint32_t)maxf, inst->d_pairs + 2 * maxf,
                                     inst->stream));
          inst->launches++;
          CU_TRY(cudaMemcpyAsync(&n_pairs, inst->d_pairs + 2 * maxf, sizeof(uint32_t), cudaMemcpyDeviceToHost, inst->stream));
          CU_TRY(cudaStreamSynchronize(inst->stream));
          const uint32_t n_copy = n_pairs < capacity ? n_pairs : capacity;
          if (n_copy > 0)
            CU_TRY(cudaMemcpy(pairs, inst->d_pairs, sizeof(uint32_t) * 2 * n_copy, cudaMemcpyDeviceToHost));
          return true;
        };
        ok = run();
        inst->nb_matches = na;
        inst->match_a = gpu_buffer_id_A;
        inst->match_b = gpu_buffer_id_B;
      }
    }
    if (invalid)
    {
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return 0;
    }
    if (!ok)
    {
      LOGE(TAG, "vksiftx_matchFeaturesCrossChecked() error: Failed to run the matching pipeline.");
      inst->cfg.on_error_callback_function(VKSIFT_VULKAN_ERROR);
      return 0;
    }
    return n_pairs;
  }

  void vksiftx_setSerialSchedule(vksift_Instance inst, const bool enabled)
  {
    wait_pipelines(inst, true, true);
    for (vksift_Instance lane : inst->lanes)
    {
      lane->serial = enabled;
      invalidate_graphs(lane);
    }
  }

  void vksiftx_setLaunchTrace(vksift_Instance inst, const bool enabled)
  {
    for (vksift_Instance lane : inst->lanes)
    {
      lane->trace = enabled;
      if (enabled)
        lane->profiling = true; /* the trace is relative to the detection's start event */
    }
  }

  uint32_t vksiftx_getLaunchTrace(vksift_Instance inst, char (*names)[32], float *start_us, float *end_us, const uint32_t capacity)
  {
    DeviceGuard g(inst->device);
    wait_pipelines(inst, true, true);
    inst = last_lane(inst); /* the most recent detection */
    if (!inst->ev_detect_valid)
      return 0;
    cudaEventSynchronize(inst->ev[EV_D4]);
    const uint32_t n = (uint32_t)inst->trace_used;
    for (uint32_t i = 0; i < n && i < capacity; i++)
    {
      float a = 0.f, b = 0.f;
      cudaEventElapsedTime(&a, inst->ev[EV_D0], inst->trace_marks[i].e0);
      cudaEventElapsedTime(&b, inst->ev[EV_D0], inst->trace_marks[i].e1);
      memcpy(names[i], inst->trace_marks[i].name, 32);
      start_us[i] = a * 1e3f;
      end_us[i] = b * 1e3f;
    }
    return n;
  }

  void vksiftx_getStageTimesMs(vksift_Instance inst, float *t)
  {
    DeviceGuard g(inst->device);
    wait_pipelines(inst, true, true);
    for (int i = 0; i < VKSIFTX_NB_STAGES; i++)
      t[i] = 0.f;
    vksift_Instance primary = inst;
    inst = last_lane(inst); /* detection stages: the most recent detection */
    if (inst->ev_detect_valid)
    {
      cudaEventSynchronize(inst->ev[EV_D4]);
      trace_dump(inst);
      cudaEventElapsedTime(&t[0], inst->ev[EV_D0], inst->ev[EV_D1]);
      if (inst->ev_d1b_valid)
      {
        /* scale space = large octaves (main stream) and small octaves (side stream), whichever finishes last */
        float tb = 0.f;
        if (cudaEventElapsedTime(&tb, inst->ev[EV_D0], inst->ev[EV_D1B]) == cudaSuccess && tb > t[0])
          t[0] = tb;
        for (int o = 1; o < inst->n_pyr_events; o++)
          if (cudaEventElapsedTime(&tb, inst->ev[EV_D0], inst->ev_pyr[o]) == cudaSuccess && tb > t[0])
            t[0] = tb;
      }
      cudaEventElapsedTime(&t[1], inst->ev[EV_D1], inst->ev[EV_D2]);
      cudaEventElapsedTime(&t[2], inst->ev[EV_D2], inst->ev[EV_D3]);
      cudaEventElapsedTime(&t[3], inst->ev[EV_D3], inst->ev[EV_D4]);
      cudaEventElapsedTime(&t[4], inst->ev[EV_D0], inst->ev[EV_D4]);
    }
    inst = primary;
    if (inst->ev_match_valid)
    {
      cudaEventSynchronize(inst->ev[EV_M2]);
      cudaEventElapsedTime(&t[5], inst->ev[EV_M0], inst->ev[EV_M1]);
      cudaEventElapsedTime(&t[6], inst->ev[EV_M1], inst->ev[EV_M2]);
      cudaEventElapsedTime(&t[7], inst->ev[EV_M0], inst->ev[EV_M2]);
    }
  }

  uint64_t vksiftx_getKernelLaunchCount(vksift_Instance inst)
  {
    uint64_t n = inst->launches;
    for (size_t k = 1; k < inst->lanes.size(); k++)
      n += inst->lanes[k]->launches;
    return n;
  }

  void vksiftx_getEffectiveTaps(vksift_Instance inst, uint32_t *radius, float *taps)
  {
    const int n = inst->cfg.nb_scales_per_octave + 3;
    for (int s = 0; s < n; s++)
    {
      radius[s] = inst->scales.radius[s];
      memcpy(taps + s * VKS_MAX_TAPS, inst->scales.taps[s], sizeof(float) * VKS_MAX_TAPS);
    }
  }

  void vksiftx_getSectionCapacities(vksift_Instance inst, const uint32_t gpu_buffer_id, uint32_t *caps)
  {
    if (!buffer_idx_valid(inst, gpu_buffer_id))
    {
      inst->cfg.on_error_callback_function(VKSIFT_INVALID_INPUT_ERROR);
      return;
    }
    const FeatureBuffer &fb = inst->buffers[gpu_buffer_id];
    for (uint32_t o = 0; o < fb.n_oct; o++)
      caps[o] = fb.cap[o];
  }

  void vksiftx_setMatcherImpl(vksift_Instance inst, const int32_t impl) { inst->matcher_impl = impl; }

#ifdef VKS_ANALYSIS
  void vksiftx_setDebugSkip(vksift_Instance inst, const int32_t mask)
  {
    wait_pipelines(inst, true, true);
    for (vksift_Instance lane : inst->lanes)
    {
      lane->debug_skip = mask;
      invalidate_graphs(lane);
    }
  }
#endif

} /* extern "C" */
