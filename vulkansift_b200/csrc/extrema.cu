/*
 * extrema.cu -- DoG extrema scan, sub-pixel refinement and deterministic ordered compaction.
 *
 * Replaces ExtractKeypoints.comp (reference: shaders/ExtractKeypoints.comp:44-231,
 * dispatched by sift_detector.c:1106-1189).  The reference appends accepted
 * keypoints with atomicAdd, i.e. in a race-dependent order; here the order is the
 * one of the detecting thread ids, key = (s, y, x) -- the canonical order of
 * SURVEY.md B-D4, which the oracle produces by construction -- and nothing on the
 * way has a capacity that depends on the image:
 *   extrema_kernel   queues every strict extremum (key (s, y, x)) in a per-octave queue
 *   refine_kernel    one thread per queued extremum (small CTAs: a refinement is a long latency chain for one lane, so the
 *                    kernel must not hold many thread slots): ExtractKeypoints.comp:118-224; an accepted keypoint keeps its
 *                    refined record beside the queue entry, sets its bit in a (s, y, x) bitmap (ns*w*h bits per octave)
 *                    and counts into its (s, y) row
 *   row_scan_kernel  exclusive scan of the row counts -> first rank of every row, n_cand, n_prim
 *   rank_kernel      rank of an accepted keypoint = row start + accepted bits in front of it in its row; the keypoints whose
 *                    rank is below the section capacity are written to slot `rank`
 * The queue is sized generously (4 x max_nb_sift_per_buffer shared by the octaves, at least 16 k) but an image may produce
 * more strict extrema than any fixed queue holds.  When an octave's queue overflows, refine_kernel and rank_kernel take a
 * slow path for that octave that needs no queue: every (s, y, x) is tested again straight from global memory, accepted
 * keypoints only set their bit, and the ranking pass walks the bitmap and refines the kept ones once more (deterministic,
 * same inputs).  Either way section overflow keeps exactly the lowest keys among the ACCEPTED keypoints, like the oracle,
 * however many raw extrema the image produces (the first version dropped queue overflow in atomic-arrival order).
 *
 * All float expressions keep the association order of the shader; the library
 * is compiled with -fmad=false so nothing is contracted.
 */
#include "vksift_internal.h"

#include <cstdio>
#include <cstdlib>

#include "layer_io.cuh"
#include "tma_util.cuh"

namespace vks
{

/* imageLoad on the DoG array; layers beyond ns+1 do not exist -> 0 (SURVEY B-D3) */
__device__ __forceinline__ float dog_at(const OctaveView &ov, int ns, int s, int x, int y)
{
  if (s < 0 || s >= ns + 2 || x < 0 || x >= ov.w || y < 0 || y >= ov.h)
    return 0.f;
  return layer_ld(ov.D, (size_t)s * ov.layer_stride + (size_t)y * ov.pitch + x, ov.fp16);
}

/* ExtractKeypoints.comp:118-224 */
__device__ bool refine_keypoint(const DetectParams &P, const OctaveView &ov, int o, int x, int y, int s, FeatHead *out)
{
  const int ns = P.ns, w = ov.w, h = ov.h;
  float oX = 0.f, oY = 0.f, oS = 0.f, gX = 0.f, gY = 0.f, gS = 0.f;
  int rx = x, ry = y, rs = s;
  for (int step = 0; step < 5; step++)
  {
    const float c = dog_at(ov, ns, rs, rx, ry);
    const float sp = dog_at(ov, ns, rs + 1, rx, ry), sm = dog_at(ov, ns, rs - 1, rx, ry);
    const float xp = dog_at(ov, ns, rs, rx + 1, ry), xm = dog_at(ov, ns, rs, rx - 1, ry);
    const float yp = dog_at(ov, ns, rs, rx, ry + 1), ym = dog_at(ov, ns, rs, rx, ry - 1);
    gS = 0.5f * (sp - sm);
    gX = 0.5f * (xp - xm);
    gY = 0.5f * (yp - ym);
    const float h11 = sp + sm - 2.f * c;
    const float h22 = xp + xm - 2.f * c;
    const float h33 = yp + ym - 2.f * c;
    const float h12 =
        0.25f * (dog_at(ov, ns, rs + 1, rx + 1, ry) - dog_at(ov, ns, rs + 1, rx - 1, ry) - dog_at(ov, ns, rs - 1, rx + 1, ry) + dog_at(ov, ns, rs - 1, rx - 1, ry));
    const float h13 =
        0.25f * (dog_at(ov, ns, rs + 1, rx, ry + 1) - dog_at(ov, ns, rs + 1, rx, ry - 1) - dog_at(ov, ns, rs - 1, rx, ry + 1) + dog_at(ov, ns, rs - 1, rx, ry - 1));
    const float h23 =
        0.25f * (dog_at(ov, ns, rs, rx + 1, ry + 1) - dog_at(ov, ns, rs, rx + 1, ry - 1) - dog_at(ov, ns, rs, rx - 1, ry + 1) + dog_at(ov, ns, rs, rx - 1, ry - 1));

    const float det = h11 * ((h22 * h33) - (h23 * h23)) - h12 * ((h12 * h33) - (h13 * h23)) + h13 * ((h12 * h23) - (h13 * h22));
    if (det != 0.0f)
    {
      const float i11 = ((h22 * h33) - (h23 * h23)) / det;
      const float i12 = -1.f * ((h12 * h33) - (h13 * h23)) / det;
      const float i13 = ((h12 * h23) - (h13 * h22)) / det;
      const float i22 = ((h11 * h33) - (h13 * h13)) / det;
      const float i23 = -1.f * ((h11 * h23) - (h13 * h12)) / det;
      const float i33 = ((h11 * h22) - (h12 * h12)) / det;
      oS = -i11 * gS - i12 * gX - i13 * gY;
      oX = -i12 * gS - i22 * gX - i23 * gY;
      oY = -i13 * gS - i23 * gX - i33 * gY;
    }
    else
    {
      return false;
    }
    if (fabsf(oX) < 0.6f && fabsf(oY) < 0.6f && fabsf(oS) < 0.6f)
      break;
    else if (step < 4)
    {
      rx += ((oX >= 0.6f && rx < (w - 2)) ? 1 : 0) + ((oX <= -0.6f && rx > 1) ? -1 : 0);
      ry += ((oY >= 0.6f && ry < (h - 2)) ? 1 : 0) + ((oY <= -0.6f && ry > 1) ? -1 : 0);
      rs += ((oS >= 0.6f && rs < (ns + 1)) ? 1 : 0) + ((oS <= -0.6f && rs > 1) ? -1 : 0);
    }
  }
  const float px = (float)rx + oX, py = (float)ry + oY, ps = (float)rs + oS;
  const float cval = dog_at(ov, ns, rs, rx, ry);
  const float val = cval + 0.5f * (gX * oX + gY * oY + gS * oS);
  if (!(fabsf(val) > P.thr && fabsf(oX) < 1.5f && fabsf(oY) < 1.5f && fabsf(oS) < 1.5f && px >= 0.f && px < (float)w && py >= 0.f && py < (float)h &&
        ps >= 0.f && ps <= (float)(ns + 1)))
    return false;
  const float e11 = dog_at(ov, ns, rs, rx + 1, ry) + dog_at(ov, ns, rs, rx - 1, ry) - 2.f * cval;
  const float e22 = dog_at(ov, ns, rs, rx, ry + 1) + dog_at(ov, ns, rs, rx, ry - 1) - 2.f * cval;
  const float e12 =
      0.25f * (dog_at(ov, ns, rs, rx + 1, ry + 1) - dog_at(ov, ns, rs, rx + 1, ry - 1) - dog_at(ov, ns, rs, rx - 1, ry + 1) + dog_at(ov, ns, rs, rx - 1, ry - 1));
  const float edgeness = ((e11 + e22) * (e11 + e22)) / ((e11 * e22) - (e12 * e12));
  if (!((edgeness < P.edge_limit) && (edgeness >= 0.f)))
    return false;
  const int octave_idx = o - (P.upsample ? 1 : 0);
  const float sf = vks_pow2i(octave_idx);
  out->scale_x = px;
  out->scale_y = py;
  out->scale_idx = (uint32_t)vks_rint(ps);
  out->octave_idx = octave_idx;
  out->sigma = P.sigma0 * vks_exp2f(ps / (float)ns) * sf;
  out->orientation = 0.f;
  out->intensity = val;
  out->x = px * sf;
  out->y = py * sf;
  return true;
}

/* ExtractKeypoints.comp:44-116.  The reference runs one thread per (x,y,s) and reads 27 texels each.
 * Here a persistent CTA walks 240x8 tiles; for each tile ONE TMA request (cp.async.bulk.tensor.3d, box
 * 256 x 10 x (ns+2), zero fill outside the image) brings the tile plus a 1-pixel halo of ALL DoG layers of the
 * octave into shared memory, double buffered so that the next tile lands while the current one is tested:
 * every DoG value is read from HBM once and no thread spends instructions on the staging.  The ns inner
 * scales are tested from smem: prefilter |v| > 0.8*thr, strict 26-neighbour extremum, and for the few
 * survivors the refinement (global reads, it walks outside the tile).  All octaves run in one launch. */
#define EX_TW 240
#define EX_TH 8
#define EX_SW 256 /* smem row: columns x0-4 .. x0+251.  TMA moves a box row by row at a fixed cost per row, so rows
                     are made as long as a box allows (256 elements = 1 KB): 72-float rows measured 6 B/clk/SM */
#define EX_SH (EX_TH + 2)
#define EX_THREADS_DEFAULT 512 /* 256 columns x 2 row groups (VKSIFT_EX_THREADS=256: one thread per column of 8 rows) */

struct ExtremaMaps
{
  CUtensorMap m[VKS_MAX_OCT]; /* 3-D fp32 maps over D[o]: (x, y, layer) */
};

__device__ __forceinline__ bool extrema_tile_coords(const DetectParams &P, int t, int *o_out, int *x0, int *y0)
{
  for (int o = 0; o < P.n_oct; o++)
  {
    const int tx = (P.oct[o].w + EX_TW - 1) / EX_TW;
    const int n = tx * ((P.oct[o].h + EX_TH - 1) / EX_TH);
    if (t < n)
    {
      *o_out = o;
      *x0 = (t % tx) * EX_TW;
      *y0 = (t / tx) * EX_TH;
      return true;
    }
    t -= n;
  }
  return false;
}

__device__ __forceinline__ float ex_to_float(float v) { return v; }
__device__ __forceinline__ float ex_to_float(__half v) { return __half2float(v); }

/* T = element type of the DoG layers (float, or __half with VKSIFT_PYRAMID_PRECISION_FLOAT16): the tile is staged and compared in
 * that type (binary16 comparisons order exactly like the fp32 values they convert to) */
template <int EX_THREADS, class T>
__global__ void __launch_bounds__(EX_THREADS) extrema_kernel(const __grid_constant__ DetectParams P, const __grid_constant__ ExtremaMaps maps, int t_begin,
                                                      int n_tiles, DetectCounters *__restrict__ cnt)
{
  constexpr int EX_RPT = EX_TH / (EX_THREADS / 256); /* rows per thread */
  constexpr int EX_HALO = 16 / (int)sizeof(T);       /* columns left of the tile in a staged row: a TMA box starts on a 16-byte boundary of the layer */
  extern __shared__ __align__(128) unsigned char ex_smem_raw[];
  T *const ex_smem = reinterpret_cast<T *>(ex_smem_raw);
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ int s_tile[2][4]; /* (octave, x0, y0) of the tile in each buffer: computed once by the thread that requests it */
  const int ns = P.ns, nl = P.ns + 2;
  const float prefilter = P.prefilter;
  const uint32_t tile_bytes = (uint32_t)(nl * EX_SH * EX_SW) * (uint32_t)sizeof(T);
  const int buf_floats = (nl * EX_SH * EX_SW + 63) & ~63; /* elements per buffer: TMA destinations must be 128-byte aligned */
  const int tid = threadIdx.x;
  const uint32_t bar0 = tma_smem_u32(&s_bar[0]);

  if (tid == 0)
  {
    tma_mbar_init(bar0, 1);
    tma_mbar_init(bar0 + 8, 1);
    tma_mbar_fence_init();
    int o, x0, y0;
    if (t_begin + (int)blockIdx.x < n_tiles && extrema_tile_coords(P, t_begin + (int)blockIdx.x, &o, &x0, &y0))
    {
      s_tile[0][0] = o;
      s_tile[0][1] = x0;
      s_tile[0][2] = y0;
      tma_mbar_expect_tx(bar0, tile_bytes);
      for (int l = 0; l < nl; l++) /* one request per layer: the TMA unit pipelines independent requests */
        tma_load_3d(tma_smem_u32(ex_smem + l * EX_SH * EX_SW), &maps.m[o], (x0 - EX_HALO) / (int)(4 / sizeof(T)), y0 - 1, l, bar0);
    }
  }
  __syncthreads();

  int it = 0;
  for (int t = t_begin + (int)blockIdx.x; t < n_tiles; t += gridDim.x, it++) /* tiles [t_begin, n_tiles) = the octaves [P.ob, P.oe) */
  {
    const int cur = it & 1;
    /* written by thread 0 before the barrier that ended the previous iteration (or the one after the prologue) */
    const int o = s_tile[cur][0], x0 = s_tile[cur][1], y0 = s_tile[cur][2];
    if (tid == 0)
    {
      /* prefetch the next tile into the other buffer (its previous readers passed the barrier below) */
      int on, xn, yn;
      if (t + (int)gridDim.x < n_tiles && extrema_tile_coords(P, t + (int)gridDim.x, &on, &xn, &yn))
      {
        s_tile[cur ^ 1][0] = on;
        s_tile[cur ^ 1][1] = xn;
        s_tile[cur ^ 1][2] = yn;
        tma_fence_proxy_async();
        tma_mbar_expect_tx(bar0 + 8 * (cur ^ 1), tile_bytes);
        for (int l = 0; l < nl; l++)
          tma_load_3d(tma_smem_u32(ex_smem + (cur ^ 1) * buf_floats + l * EX_SH * EX_SW), &maps.m[on], (xn - EX_HALO) / (int)(4 / sizeof(T)), yn - 1, l,
                      bar0 + 8 * (cur ^ 1));
      }
    }
    tma_mbar_wait(bar0 + 8 * cur, (uint32_t)(it >> 1) & 1u);

    const OctaveView &ov = P.oct[o];
    const T *tile = ex_smem + cur * buf_floats;
    const int lx = tid & 255, ly = (tid >> 8) * EX_RPT; /* one column of EX_RPT rows per thread */
    const int x = x0 + lx;
    const int ow = ov.w, oh = ov.h;
    if (lx < EX_TW && x >= 1 && x < ow - 1)
    {
      /* rows of this thread that are inside [1, h-2] */
      const int r_lo = max(0, 1 - (y0 + ly)), r_hi = min(EX_RPT, (oh - 1) - (y0 + ly));
      const uint32_t row_ok = (r_hi > r_lo) ? (((1u << r_hi) - 1u) & ~((1u << r_lo) - 1u)) : 0u;
      for (int s = 1; s <= ns; s++)
      {
        const T *col = tile + (s * EX_SH + ly + 1) * EX_SW + (lx + EX_HALO);
        /* prefilter all centre values first (independent loads), then visit only the survivors */
        T cv[EX_RPT];
#pragma unroll
        for (int r = 0; r < EX_RPT; r++)
          cv[r] = col[r * EX_SW];
        uint32_t mask = 0;
#pragma unroll
        for (int r = 0; r < EX_RPT; r++)
          mask |= (fabsf(ex_to_float(cv[r])) > prefilter) ? (1u << r) : 0u;
        mask &= row_ok;
        while (mask)
        {
          const int r = __ffs(mask) - 1;
          mask &= mask - 1;
          const T *pc = col + r * EX_SW;
          const T c = pc[0];
          /* strict 26-neighbour test (:59-116), staged so that the many non-extrema leave after 4 compares */
          bool gt = true, lt = true;
#define EX_CMP(off)                                                                                                                                  \
  {                                                                                                                                                  \
    const T n = pc[off];                                                                                                                             \
    gt = gt && (c > n);                                                                                                                              \
    lt = lt && (c < n);                                                                                                                              \
  }
          EX_CMP(-1) EX_CMP(1) EX_CMP(-EX_SW) EX_CMP(EX_SW)
          if (!(gt || lt))
            continue;
          EX_CMP(-EX_SW - 1) EX_CMP(-EX_SW + 1) EX_CMP(EX_SW - 1) EX_CMP(EX_SW + 1)
          if (!(gt || lt))
            continue;
#pragma unroll
          for (int ds = -1; ds <= 1; ds += 2)
          {
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
#pragma unroll
              for (int dx = -1; dx <= 1; dx++)
                EX_CMP(ds * (EX_SH * EX_SW) + dy * EX_SW + dx)
            if (!(gt || lt))
              break;
          }
#undef EX_CMP
          if (!(gt || lt))
            continue;
          /* a strict extremum: queue it.  The sub-pixel refinement (a long serial chain of dependent global loads and
           * divisions for one lane) runs in its own kernel with one thread per queued extremum instead of stalling this
           * tile's whole CTA at the next barrier.  The counter keeps counting past the queue capacity: that is how the
           * later kernels see an overflow (and switch to the path that needs no queue). */
          const int y = y0 + ly + r;
          const uint32_t slot = atomicAdd(&cnt->n_raw[o], 1u);
          if (slot < P.q_cap[o])
            P.raw_q[P.q_off[o] + slot] = ((unsigned long long)s << 40) | ((unsigned long long)y << 20) | (unsigned long long)x;
        }
      }
    }
    __syncthreads(); /* everyone is done with buffer `cur` before it is refilled two iterations later */
  }
}

/* ---- ordered compaction ------------------------------------------------------------------------------------------ */
#define RF_THREADS 64 /* small CTAs: most of a refinement is waiting for one lane's dependent loads */
#define RF_CTAS 256   /* per octave: 16 k threads, one queued extremum each for all but pathological images */
#define RQ_ACCEPTED (1ull << 63)

/* strict extremum test straight from global memory: the slow path's twin of the shared-memory test in extrema_kernel */
__device__ bool is_extremum_global(const DetectParams &P, const OctaveView &ov, int x, int y, int s)
{
  const float c = dog_at(ov, P.ns, s, x, y);
  if (!(fabsf(c) > P.prefilter))
    return false;
  bool gt = true, lt = true;
  for (int ds = -1; ds <= 1; ds++)
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++)
      {
        if (ds == 0 && dy == 0 && dx == 0)
          continue;
        const float n = dog_at(ov, P.ns, s + ds, x + dx, y + dy);
        gt = gt && (c > n);
        lt = lt && (c < n);
      }
  return gt || lt;
}

__device__ __forceinline__ void mark_accepted(const DetectParams &P, int o, int x, int y, int s)
{
  const uint32_t row = (uint32_t)((s - 1) * P.oct[o].h + y);
  atomicOr(P.acc_bm + P.bm_off[o] + (size_t)row * P.bm_rw[o] + (uint32_t)(x >> 5), 1u << (x & 31));
  atomicAdd(P.row_cnt + P.row_off[o] + row, 1u);
}

/* ExtractKeypoints.comp:118-224: grid (RF_CTAS, octaves of the launch) */
__global__ void __launch_bounds__(RF_THREADS) refine_kernel(const __grid_constant__ DetectParams P, const DetectCounters *__restrict__ cnt)
{
  const int o = P.ob + (int)blockIdx.y;
  const OctaveView &ov = P.oct[o];
  const uint32_t n = cnt->n_raw[o];
  const uint32_t tid = blockIdx.x * RF_THREADS + threadIdx.x, stride = gridDim.x * RF_THREADS;
  if (n <= P.q_cap[o])
  {
    unsigned long long *__restrict__ q = P.raw_q + P.q_off[o];
    FeatHead *__restrict__ qh = P.q_heads + P.q_off[o];
    for (uint32_t i = tid; i < n; i += stride)
    {
      const unsigned long long key = q[i];
      const int s = (int)(key >> 40), y = (int)((key >> 20) & 0xfffffu), x = (int)(key & 0xfffffu);
      FeatHead hd;
      if (!refine_keypoint(P, ov, o, x, y, s, &hd))
        continue;
      qh[i] = hd;
      q[i] = key | RQ_ACCEPTED;
      mark_accepted(P, o, x, y, s);
    }
    return;
  }
  /* queue overflow: every inner (s, y, x) again, from global memory */
  const unsigned long long total = (unsigned long long)P.ns * (unsigned long long)(ov.h - 2) * (unsigned long long)(ov.w - 2);
  for (unsigned long long i = tid; i < total; i += stride)
  {
    const int x = 1 + (int)(i % (unsigned long long)(ov.w - 2));
    const unsigned long long t = i / (unsigned long long)(ov.w - 2);
    const int y = 1 + (int)(t % (unsigned long long)(ov.h - 2)), s = 1 + (int)(t / (unsigned long long)(ov.h - 2));
    FeatHead hd;
    if (is_extremum_global(P, ov, x, y, s) && refine_keypoint(P, ov, o, x, y, s, &hd))
      mark_accepted(P, o, x, y, s);
  }
}

/* one CTA per octave: row_cnt -> exclusive prefix (first rank of each row), n_cand = accepted keypoints (the
 * reference's counter keeps counting past the capacity, ExtractKeypoints.comp:208-211), n_prim = min(n_cand, cap) */
#define SCAN_THREADS 1024
__global__ void __launch_bounds__(SCAN_THREADS) row_scan_kernel(const __grid_constant__ DetectParams P, DetectCounters *__restrict__ cnt)
{
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const int o = P.ob + (int)blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
  const uint32_t n_rows = (uint32_t)(P.ns * P.oct[o].h);
  uint32_t *__restrict__ rc = P.row_cnt + P.row_off[o];
  if (tid == 0)
    s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_rows; base += SCAN_THREADS)
  {
    const uint32_t i = base + tid;
    const uint32_t e = (i < n_rows) ? rc[i] : 0u;
    uint32_t v = e;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d)
        v += t;
    }
    if (lane == 31)
      s_warp[wi] = v;
    __syncthreads();
    if (wi == 0)
    {
      uint32_t wv = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1)
      {
        const uint32_t t = __shfl_up_sync(0xffffffffu, wv, d);
        if (lane >= d)
          wv += t;
      }
      s_warp[lane] = wv;
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    if (i < n_rows)
      rc[i] = carry + (wi ? s_warp[wi - 1] : 0u) + (v - e);
    __syncthreads();
    if (tid == SCAN_THREADS - 1)
      s_carry = carry + s_warp[31];
    __syncthreads();
  }
  if (tid == 0)
  {
    cnt->n_cand[o] = s_carry;
    cnt->n_prim[o] = min(s_carry, P.cap[o]);
  }
}

/* rank of an accepted keypoint = first rank of its (s, y) row + accepted bits in front of it in that row */
__device__ __forceinline__ uint32_t rank_of(const DetectParams &P, int o, uint32_t row, int x)
{
  const uint32_t *__restrict__ bm = P.acc_bm + P.bm_off[o] + (size_t)row * P.bm_rw[o];
  uint32_t rank = P.row_cnt[P.row_off[o] + row];
  const uint32_t xw = (uint32_t)x >> 5;
  for (uint32_t k = 0; k < xw; k++)
    rank += (uint32_t)__popc(bm[k]);
  return rank + (uint32_t)__popc(bm[xw] & ((1u << (x & 31)) - 1u));
}

__global__ void __launch_bounds__(RF_THREADS) rank_kernel(const __grid_constant__ DetectParams P, const DetectCounters *__restrict__ cnt,
                                                          FeatHead *__restrict__ prim)
{
  const int o = P.ob + (int)blockIdx.y;
  const OctaveView &ov = P.oct[o];
  const uint32_t n = cnt->n_raw[o];
  const uint32_t tid = blockIdx.x * RF_THREADS + threadIdx.x, stride = gridDim.x * RF_THREADS;
  if (n <= P.q_cap[o])
  {
    const unsigned long long *__restrict__ q = P.raw_q + P.q_off[o];
    const FeatHead *__restrict__ qh = P.q_heads + P.q_off[o];
    for (uint32_t i = tid; i < n; i += stride)
    {
      const unsigned long long key = q[i];
      if (!(key & RQ_ACCEPTED))
        continue;
      const int s = (int)((key >> 40) & 0xfffffu), y = (int)((key >> 20) & 0xfffffu), x = (int)(key & 0xfffffu);
      const uint32_t rank = rank_of(P, o, (uint32_t)((s - 1) * ov.h + y), x);
      if (rank < P.cap[o])
        prim[P.sec_off[o] + rank] = qh[i];
    }
    return;
  }
  /* queue overflow: walk the accepted bitmap, refine the kept keypoints once more (same inputs, same result) */
  const uint32_t rw = P.bm_rw[o];
  const uint32_t n_words = (uint32_t)(P.ns * ov.h) * rw;
  const uint32_t *__restrict__ bm = P.acc_bm + P.bm_off[o];
  for (uint32_t wi = tid; wi < n_words; wi += stride)
  {
    uint32_t bits = bm[wi];
    if (bits == 0)
      continue;
    const uint32_t row = wi / rw, xw = wi - row * rw;
    uint32_t rank = rank_of(P, o, row, (int)(xw * 32u));
    const int s = (int)(row / (uint32_t)ov.h) + 1, y = (int)(row % (uint32_t)ov.h);
    while (bits && rank < P.cap[o])
    {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      FeatHead hd;
      refine_keypoint(P, ov, o, (int)(xw * 32u) + b, y, s, &hd);
      prim[P.sec_off[o] + rank] = hd;
      rank++;
    }
  }
}

struct ExtremaPlan
{
  ExtremaMaps maps;
  int n_tiles;
  bool valid;
};

/* (re)build the tensor maps for the current pyramid; called when the resolution changes */
cudaError_t extrema_plan_build(const DetectParams &P, ExtremaPlan **plan_io)
{
  if (*plan_io == nullptr)
    *plan_io = new ExtremaPlan();
  ExtremaPlan *pl = *plan_io;
  pl->valid = false;
  pl->n_tiles = 0;
  for (int o = 0; o < P.n_oct; o++)
  {
    const OctaveView &ov = P.oct[o];
    const uint64_t dims[3] = {(uint64_t)ov.w, (uint64_t)ov.h, (uint64_t)(P.ns + 2)};
    /* binary16 layers are described to the TMA unit as fp32 tensors of half the width (a pair of halves = one 32-bit element;
     * tile origins are even): same box bytes, same zero fill outside the image */
    const uint64_t es = ov.fp16 ? 2 : 4;
    const uint64_t strides[2] = {(uint64_t)ov.pitch * es, (uint64_t)ov.layer_stride * es};
    const uint32_t box[3] = {(uint32_t)(ov.fp16 ? EX_SW / 2 : EX_SW), EX_SH, 1}; /* one layer per request */
    const uint64_t dims32[3] = {(uint64_t)(ov.fp16 ? (ov.w + 1) / 2 : ov.w), dims[1], dims[2]};
    if (!tma_make_map_f32(&pl->maps.m[o], ov.D, 3, dims32, strides, box))
      return cudaErrorInvalidValue;
    pl->n_tiles += ((ov.w + EX_TW - 1) / EX_TW) * ((ov.h + EX_TH - 1) / EX_TH);
  }
  pl->valid = true;
  return cudaSuccess;
}

void extrema_plan_destroy(ExtremaPlan *pl) { delete pl; }

void extrema_layout(DetectParams *P, size_t *bm_words, size_t *rows, size_t *queue_entries, uint32_t queue_total)
{
  size_t words = 0, nrows = 0, qn = 0;
  double px_total = 0.;
  for (int o = 0; o < P->n_oct; o++)
    px_total += (double)P->oct[o].w * (double)P->oct[o].h;
  for (int o = 0; o < P->n_oct; o++)
  {
    P->bm_off[o] = (uint32_t)words;
    P->bm_rw[o] = (uint32_t)((P->oct[o].w + 31) / 32);
    P->row_off[o] = (uint32_t)nrows;
    const size_t r = (size_t)P->ns * (size_t)P->oct[o].h;
    words += (r * P->bm_rw[o] + 3) & ~(size_t)3;
    nrows += r;
    /* the octave's share of the extrema queue, by pixel count, at least 256 entries */
    size_t share = (size_t)((double)queue_total * ((double)P->oct[o].w * (double)P->oct[o].h) / (px_total > 0. ? px_total : 1.));
    if (share < 256)
      share = 256;
    P->q_off[o] = (uint32_t)qn;
    P->q_cap[o] = (uint32_t)share;
    qn += share;
  }
  *bm_words = words;
  *rows = nrows;
  *queue_entries = qn;
}

cudaError_t launch_extrema(const DetectParams &P, const ExtremaPlan *pl, DetectCounters *cnt, FeatHead *prim, cudaStream_t st, uint64_t *launch_count,
                           int scan_ctas)
{
  if (!pl || !pl->valid || pl->n_tiles == 0 || P.oe <= P.ob)
    return cudaSuccess;
  int t_begin = 0, t_end = 0;
  for (int o = 0; o < P.oe; o++)
  {
    const int n = ((P.oct[o].w + EX_TW - 1) / EX_TW) * ((P.oct[o].h + EX_TH - 1) / EX_TH);
    if (o < P.ob)
      t_begin += n;
    t_end += n;
  }
  const bool fp16 = P.oct[P.ob].fp16 != 0;
  const size_t smem = 2 * (fp16 ? sizeof(__half) : sizeof(float)) * (size_t)((((P.ns + 2) * EX_SH * EX_SW) + 63) & ~63);
  if (smem > 220 * 1024)
    return cudaErrorInvalidConfiguration; /* rejected at instance creation (extrema_scales_supported) */
  static bool attr_done[64] = {false};
  static int threads = 0;
  if (threads == 0)
  {
    const char *e = getenv("VKSIFT_EX_THREADS");
    threads = (e && atoi(e) == 256) ? 256 : EX_THREADS_DEFAULT;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev])
  {
    cudaError_t e = cudaFuncSetAttribute(extrema_kernel<256, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(extrema_kernel<512, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(extrema_kernel<256, __half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(extrema_kernel<512, __half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess)
      return e;
    attr_done[dev] = true;
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int per_sm = (int)((220 * 1024) / smem) < 1 ? 1 : (int)((220 * 1024) / smem);
  /* Persistent scan CTAs (each holds a double-buffered 100 KB tile): two per SM scan a layer set at 4 TB/s, which is what one
   * detection alone wants.  With several detections in flight the scan is not on the critical path of the GPU (it moves bytes,
   * the blur and descriptor kernels of the other lanes need the issue slots and the shared memory), and FEWER scan CTAs give the
   * higher throughput: 8 lanes, ms per 1920x1080 image, 296 / 148 / 74 / 48 / 24 / 16 CTAs: 0.2919 / 0.2874 / 0.2833 / 0.2806 /
   * 0.2784 / 0.2799; 2 lanes: 0.3401 / 0.3296 / 0.3281 for 296 / 148 / 74.  The caller passes the count for its lane count. */
  static int env_grid = -1;
  if (env_grid < 0)
  {
    const char *e = getenv("VKSIFT_EX_GRID");
    env_grid = e ? atoi(e) : 0;
  }
  int grid = sms * (per_sm > 2 ? 2 : per_sm);
  if (scan_ctas > 0 && scan_ctas < grid)
    grid = scan_ctas;
  if (env_grid > 0)
    grid = env_grid;
  if (grid > t_end - t_begin)
    grid = t_end - t_begin;
  if (threads == 256 && fp16)
    extrema_kernel<256, __half><<<grid, 256, smem, st>>>(P, pl->maps, t_begin, t_end, cnt);
  else if (threads == 256)
    extrema_kernel<256, float><<<grid, 256, smem, st>>>(P, pl->maps, t_begin, t_end, cnt);
  else if (fp16)
    extrema_kernel<512, __half><<<grid, 512, smem, st>>>(P, pl->maps, t_begin, t_end, cnt);
  else
    extrema_kernel<512, float><<<grid, 512, smem, st>>>(P, pl->maps, t_begin, t_end, cnt);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return e;
  const dim3 rgrid(RF_CTAS, P.oe - P.ob, 1);
  refine_kernel<<<rgrid, RF_THREADS, 0, st>>>(P, cnt);
  row_scan_kernel<<<P.oe - P.ob, SCAN_THREADS, 0, st>>>(P, cnt);
  rank_kernel<<<rgrid, RF_THREADS, 0, st>>>(P, cnt, prim);
  *launch_count += 4;
  return cudaGetLastError();
}

/* largest nb_scales_per_octave whose double-buffered scan tile fits in shared memory */
bool extrema_scales_supported(int ns) { return 2 * sizeof(float) * (size_t)((((ns + 2) * EX_SH * EX_SW) + 31) & ~31) <= 220 * 1024; }

} // namespace vks
