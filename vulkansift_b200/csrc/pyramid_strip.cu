/*
 * pyramid_strip.cu -- multi-layer streaming blur kernel for the large octaves (sm_100a).
 *
 * One launch produces a chain of NL consecutive Gaussian layers of one octave and everything that hangs off them
 * (DoG layers, the NEAREST-decimated seed of the next octave) from ONE read of the layer in front of the chain:
 *   reference: 2 dispatches of GaussianBlur[Interpolated].comp per layer + DifferenceOfGaussian.comp + vkCmdBlitImage,
 *   sift_detector.c:955-1079, each of which goes through device memory.
 * The per-layer kernel of pyramid.cu re-reads its 33 MB source for every layer (1.5x the bytes that have to move);
 * here the intermediate layers stay in shared memory.
 *
 * Geometry.  A CTA (512 threads, one per SM) owns a strip of W output columns and a segment of `hseg` rows and walks down
 * it in steps of 8 rows.  Layer k of the chain is computed on the strip plus the summed radii of the layers behind it
 * (x halo, recomputed by the neighbouring strip as well), rows flow through a software pipeline:
 *
 *   source rows --1-D bulk copies (TMA unit), 2 steps ahead--> source ring (6 blocks of 8 rows)
 *   H_k : horizontal pass of layer k on the 8 newest rows of layer k-1          -> H_k ring (2 blocks, double buffer)
 *   V_k : vertical pass of layer k; a thread owns a column pair and keeps the 2 R_k previous H_k rows IN REGISTERS
 *         (rolling window), so every H value is read from shared memory once     -> G_k ring (input of H_k+1 and of the
 *         DoG of layer k+1), global G_k, global DoG = G_k - G_k-1, decimated seed
 *
 * In step s every pass works on the rows the previous steps completed (H_k(s) on rows [8s - offH_k, +8), V_k(s) on rows
 * [8s - offV_k, +8)), so ONE __syncthreads per step is the only synchronisation and all 2 NL passes of a step run
 * concurrently on different warps: warps have fixed roles (one V role: the register window is persistent state; one H
 * role), paired so that every warp carries the same number of fp32 operations per step.
 *
 * Borders.  MIRRORED_REPEAT in y: the source rows are fetched at mirrored coordinates.  In x: the cells of a source row
 * outside the image are patched with their mirror cell after the row has landed.  Intermediate layers computed outside
 * the image on the mirrored extension equal their mirror pixel bit for bit (the tap pairs a+b only swap operands), so
 * nothing else needs care.  Per-pixel arithmetic is the sequence of include/vksift_arith.h, as everywhere.
 */
#include "vksift_internal.h"

#include "blur_arith.cuh"
#include "tma_util.cuh"

#include <cstdlib>
#include <cstring>

namespace vks
{

#define ST_B 8          /* rows per step */
#define ST_THREADS 512
#define ST_SRC_BLOCKS 6 /* source ring, blocks of ST_B rows */
#define ST_ISSUER (ST_THREADS - 32) /* lane 0 of the last warp issues the bulk copies */

template <int NL_, int R0, int R1, int R2, int W_>
struct StripGeo
{
  static constexpr int NL = NL_;
  static constexpr int W = W_;
  static constexpr int B = ST_B;
  __host__ __device__ static constexpr int R(int k) { return k == 0 ? R0 : (k == 1 ? R1 : R2); }
  /* summed radii of the layers behind layer k (k = -1: of the whole chain) */
  __host__ __device__ static constexpr int halo(int k) { return (k < 0 ? R0 : 0) + ((k < 1 && NL > 1) ? R1 : 0) + ((k < 2 && NL > 2) ? R2 : 0); }
  static constexpr int HX = halo(-1); /* halo of the source layer, x and y */
  static constexpr int HY = HX;
  static constexpr int HXA = (HX + 3) & ~3; /* region column of the strip's first output column (16-byte aligned rows) */
  /* region columns [clo, chi) of layer k's output */
  __host__ __device__ static constexpr int clo(int k) { return HXA - halo(k); }
  __host__ __device__ static constexpr int chi(int k) { return HXA + W + halo(k); }
  __host__ __device__ static constexpr int rx(int k) { return (R(k) + 3) & ~3; }
  /* H units are 16 columns wide and start on a multiple of 4 columns (LDS.128 windows) */
  __host__ __device__ static constexpr int hu_lo(int k) { return clo(k) & ~3; }
  __host__ __device__ static constexpr int hu_n(int k) { return (chi(k) - hu_lo(k) + 63) / 64; } /* warps of 8 rows x 64 columns */
  __host__ __device__ static constexpr int vu_n(int k) { return (chi(k) - clo(k) + 63) / 64; }   /* warps of 64 columns */
  static constexpr int src_lo = (HXA - HX) & ~3;
  static constexpr int src_hi = (HXA + W + HX + 3) & ~3;
  __host__ __device__ static constexpr int imax(int a, int b) { return a > b ? a : b; }
  __host__ __device__ static constexpr int need_cols(int k) { return imax(hu_lo(k) + 64 * hu_n(k) + rx(k), clo(k) + 64 * vu_n(k)); }
  static constexpr int NEED = imax(imax(need_cols(0), NL > 1 ? need_cols(1) : 0), imax(NL > 2 ? need_cols(2) : 0, src_hi));
  static constexpr int S = ((NEED + 7) / 8) * 8 + 4; /* row stride of every ring, floats: S/4 odd -> LDS.128 down 8 rows is conflict free */
  /* pipeline offsets: H_k(s) works on region rows [8s - offH(k), +8), V_k(s) emits rows [8s - offV(k), +8) */
  __host__ __device__ static constexpr int offH(int k) { return (k > 0 ? 2 * B + R0 : 0) + (k > 1 ? 2 * B + R1 : 0); }
  __host__ __device__ static constexpr int offV(int k) { return offH(k) + B + R(k); }
  /* rows of the G_k ring (k < NL-1): written by V_k, read by H_k+1 one step later and by the DoG of V_k+1 */
  __host__ __device__ static constexpr int DG(int k) { return (3 * B + R(k + 1) + 7) & ~7; }
  static constexpr int SRC_ROWS = ST_SRC_BLOCKS * B;
  /* shared memory layout, floats */
  static constexpr int OFF_SRC = 0;
  __host__ __device__ static constexpr int off_h(int k) { return SRC_ROWS * S + k * 2 * B * S; }
  __host__ __device__ static constexpr int off_g(int k) { return SRC_ROWS * S + NL * 2 * B * S + (k > 0 ? DG(0) * S : 0); }
  static constexpr int FLOATS = SRC_ROWS * S + NL * 2 * B * S + (NL > 1 ? DG(0) * S : 0) + (NL > 2 ? DG(1) * S : 0);
  static constexpr int SMEM_BYTES = FLOATS * 4 + ST_SRC_BLOCKS * 8 + 16;
  static_assert(hu_n(0) <= 4 && vu_n(0) <= 4, "a pass is spread over four warps");
  static_assert(hu_lo(0) - rx(0) >= 0 && (NL < 2 || hu_lo(1) - rx(1) >= 0) && (NL < 3 || hu_lo(2) - rx(2) >= 0), "H window left of the region");
  static_assert(R0 <= 16 && R1 <= 16 && R2 <= 16 && (R0 % 2) == 0 && (R1 % 2) == 0 && (R2 % 2) == 0, "even radii up to 16");
  static_assert(W % 4 == 0, "strip origin on a 16-byte boundary");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

struct StripParams
{
  const float *src;  /* layer in front of the chain */
  float *g[3];       /* Gaussian layers written */
  float *d[3];       /* DoG layers written: d[k] = g[k] - (k ? g[k-1] : src) */
  float *next;       /* layer 0 of the next octave (NEAREST decimation of the chain's last layer) or NULL */
  int w, h, pitch;
  int next_w, next_h, next_pitch;
  int hseg, n_strips;
  int fp16;
  float2 taps2[3][14]; /* (k,k) pairs, zero padded to the even radius */
};

__device__ __forceinline__ int st_mirror_once(int i, int n)
{
  i = i < 0 ? -1 - i : i;
  return i >= n ? 2 * n - 1 - i : i;
}

__device__ __forceinline__ void st_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

#define ST_TAP(i) (*reinterpret_cast<const pk2 *>(&taps2[i]))

/* horizontal pass: 16 outputs of one row from a window of 16 + 2 RX inputs (LDS.128), written with STS.128.
 * Output pairs (x, x+1): even taps use the aligned register pairs of the window, odd taps add the two scalars into a
 * fresh pair; both feed one FFMA2 -- the arrangement of the per-layer kernel (pyramid.cu). */
template <int R>
__device__ __forceinline__ void strip_h_unit(const float *__restrict__ in, float *__restrict__ out, const float2 *__restrict__ taps2)
{
  constexpr int RX = (R + 3) & ~3;
  constexpr int WN = 16 + 2 * RX;
  pk2 wp[WN / 2]; /* wp[j] = (in[2j], in[2j+1]) */
  const ulonglong2 *wsrc = reinterpret_cast<const ulonglong2 *>(in);
#pragma unroll
  for (int j = 0; j < WN / 4; j++)
  {
    const ulonglong2 v = wsrc[j];
    wp[2 * j] = v.x;
    wp[2 * j + 1] = v.y;
  }
  ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(out);
#pragma unroll
  for (int hb = 0; hb < 2; hb++)
  {
    pk2 acc[4];
#pragma unroll
    for (int j = 0; j < 4; j++)
      acc[j] = pk_mul(wp[RX / 2 + 4 * hb + j], ST_TAP(0));
#pragma unroll
    for (int i = 1; i <= R; i++)
    {
#pragma unroll
      for (int j = 0; j < 4; j++)
      {
        const int c = RX / 2 + 4 * hb + j;
        pk2 sum;
        if ((i & 1) == 0)
          sum = pk_add(wp[c + i / 2], wp[c - i / 2]);
        else
        {
          const float s0 = __fadd_rn(pk_hi(wp[c + (i - 1) / 2]), pk_hi(wp[c - (i + 1) / 2]));
          const float s1 = __fadd_rn(pk_lo(wp[c + (i + 1) / 2]), pk_lo(wp[c - (i - 1) / 2]));
          sum = pk_make(s0, s1);
        }
        acc[j] = pk_fma(sum, ST_TAP(i), acc[j]);
      }
    }
    dst[2 * hb] = make_ulonglong2(acc[0], acc[1]);
    dst[2 * hb + 1] = make_ulonglong2(acc[2], acc[3]);
  }
}

template <class G, int K>
__device__ __forceinline__ void strip_h_role(const StripParams &P, float *smem, int s, int cb, int lane, int ry_end)
{
  constexpr int R = G::R(K);
  constexpr int RX = G::rx(K);
  if (cb >= G::hu_n(K))
    return;
  const int a = ST_B * s - G::offH(K);
  /* rows of H_k that a stored pixel depends on */
  if (a + ST_B <= G::HY - G::halo(K) - R || a >= ry_end + G::halo(K) + R)
    return;
  const int r = lane & 7, g = lane >> 3;
  const int ry = max(a + r, 0);
  const int cs = G::hu_lo(K) + 64 * cb + 16 * g;
  const float *in_row = (K == 0) ? smem + G::OFF_SRC + (ry % G::SRC_ROWS) * G::S : smem + G::off_g(K > 0 ? K - 1 : 0) + (ry % G::DG(K > 0 ? K - 1 : 0)) * G::S;
  float *out = smem + G::off_h(K) + ((s & 1) * ST_B + r) * G::S + cs;
  strip_h_unit<R>(in_row + cs - RX, out, P.taps2[K]);
}

/* vertical pass of layer K for one column pair: `win` holds the H rows [a - R, a + R) in front of the 8 rows loaded here */
template <class G, int K, int WN>
__device__ __forceinline__ void strip_v_role(const StripParams &P, float *smem, int s, int cb, int lane, int ry_end, int XO, int yb, pk2 (&win)[WN])
{
  static_assert(WN >= 2 * G::R(K) + ST_B, "window too small for this layer");
  constexpr int R = G::R(K);
  constexpr bool LAST = (K == G::NL - 1);
  if (cb >= G::vu_n(K) || s < 1)
    return;
  const float2 *__restrict__ taps2 = P.taps2[K];
  const int a = ST_B * s - G::offV(K); /* first output row (region coordinates) */
  const int col_raw = G::clo(K) + 64 * cb + 2 * lane;
  const bool lane_ok = col_raw < G::chi(K);
  const int col = lane_ok ? col_raw : G::clo(K);
  /* the 8 newest H rows: written by H_K in the previous step */
  const float *hblk = smem + G::off_h(K) + (((s - 1) & 1) * ST_B) * G::S + col;
#pragma unroll
  for (int i = 0; i < ST_B; i++)
    win[2 * R + i] = *reinterpret_cast<const pk2 *>(hblk + i * G::S);
  if (a + ST_B > G::HY - G::halo(K) && a < ry_end + G::halo(K))
  {
    const int x = XO + col;
    const bool col_store = lane_ok && col >= G::HXA && col < G::HXA + G::W && x < P.w;
    const bool fp16 = P.fp16 != 0;
#pragma unroll
    for (int qb = 0; qb < ST_B; qb += 4)
    {
      pk2 acc[4];
#pragma unroll
      for (int j = 0; j < 4; j++)
        acc[j] = pk_mul(win[R + qb + j], ST_TAP(0));
#pragma unroll
      for (int i = 1; i <= R; i++)
      {
#pragma unroll
        for (int j = 0; j < 4; j++)
          acc[j] = pk_fma(pk_add(win[R + qb + j + i], win[R + qb + j - i]), ST_TAP(i), acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; j++)
      {
        const int ry = a + qb + j;
        if (ry < 0)
          continue;
        if (fp16)
          acc[j] = round_half2(acc[j]);
        if (!LAST)
        {
          if (lane_ok)
            *reinterpret_cast<pk2 *>(smem + G::off_g(LAST ? 0 : K) + (ry % G::DG(LAST ? 0 : K)) * G::S + col) = acc[j];
        }
        if (col_store && ry >= G::HY && ry < ry_end)
        {
          const int y = yb + ry;
          const size_t o = (size_t)y * (size_t)P.pitch + (size_t)x;
          const float *crow = (K == 0) ? smem + G::OFF_SRC + (ry % G::SRC_ROWS) * G::S : smem + G::off_g(K > 0 ? K - 1 : 0) + (ry % G::DG(K > 0 ? K - 1 : 0)) * G::S;
          pk2 dd = pk_sub(acc[j], *reinterpret_cast<const pk2 *>(crow + col));
          if (fp16)
            dd = round_half2(dd);
          if (LAST)
            *reinterpret_cast<pk2 *>(P.g[K] + o) = acc[j]; /* read back by the next chain's launch: keep it in L2 */
          else
            __stcs(reinterpret_cast<pk2 *>(P.g[K] + o), acc[j]);
          __stcs(reinterpret_cast<pk2 *>(P.d[K] + o), dd);
          if (LAST && P.next != nullptr && (y & 1))
          {
            /* x even, y odd: the odd column of the pair feeds next(x >> 1, y >> 1) */
            const int nx = x >> 1, ny = y >> 1;
            if (nx < P.next_w && ny < P.next_h)
              P.next[(size_t)ny * (size_t)P.next_pitch + (size_t)nx] = pk_hi(acc[j]);
          }
        }
      }
    }
  }
  /* roll the window: the newest 2R rows stay */
#pragma unroll
  for (int i = 0; i < 2 * R; i++)
    win[i] = win[i + ST_B];
}

template <int NL, int R0, int R1, int R2, int W>
__global__ void __launch_bounds__(ST_THREADS, 1) pyramid_strip_kernel(const __grid_constant__ StripParams P)
{
  using G = StripGeo<NL, R0, R1, R2, W>;
  extern __shared__ __align__(128) float st_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = tma_smem_u32(st_smem + G::FLOATS);

  const int strip = (int)blockIdx.x % P.n_strips, seg = (int)blockIdx.x / P.n_strips;
  const int x0 = strip * W, XO = x0 - G::HXA;
  const int y0s = seg * P.hseg, y1s = min(P.h, y0s + P.hseg);
  const int yb = y0s - G::HY;               /* image row of region row 0 */
  const int ry_end = G::HY + (y1s - y0s);   /* region rows [HY, ry_end) are stored */
  const int n_steps = (ry_end + G::offV(NL - 1) + ST_B - 1) / ST_B;

  /* source fetch: columns [fx0, fx1) of the layer, 16-byte aligned on both sides, clipped to the padded row */
  const int fx0 = max(XO + G::src_lo, 0), fx1 = min(XO + G::src_hi, P.pitch);
  const uint32_t row_bytes = (uint32_t)(fx1 - fx0) * 4u;
  const int dcol = fx0 - XO;
  const bool border_l = (XO + G::HXA - G::HX) < 0, border_r = (XO + G::HXA + W + G::HX) > P.w;

  pdl_launch_dependents();
  if (tid == 0)
  {
#pragma unroll
    for (int b = 0; b < ST_SRC_BLOCKS; b++)
      tma_mbar_init(bar0 + 8 * b, 1);
    tma_mbar_fence_init();
  }
  pdl_wait(); /* the source layer is written by the previous launch of the stream */
  __syncthreads();

  auto issue = [&](int b) {
    const uint32_t bar = bar0 + 8u * (uint32_t)(b % ST_SRC_BLOCKS);
    tma_mbar_expect_tx(bar, row_bytes * ST_B);
#pragma unroll
    for (int i = 0; i < ST_B; i++)
    {
      const int ry = ST_B * b + i;
      const int gy = min(max(st_mirror_once(yb + ry, P.h), 0), P.h - 1); /* MIRRORED_REPEAT; rows far outside feed nothing that is stored */
      st_bulk_load(tma_smem_u32(st_smem + G::OFF_SRC + (ry % G::SRC_ROWS) * G::S + dcol), P.src + (size_t)gy * (size_t)P.pitch + (size_t)fx0, row_bytes, bar);
    }
  };
  /* cells of the source rows of block b that lie outside the image take the value of their mirror cell (one warp) */
  auto patch = [&](int b) {
    tma_mbar_wait(bar0 + 8u * (uint32_t)(b % ST_SRC_BLOCKS), (uint32_t)(b / ST_SRC_BLOCKS) & 1u);
    const int c_first = G::HXA - G::HX, c_last = G::HXA + W + G::HX; /* columns a stored pixel can depend on */
    const int nl = border_l ? min(-(XO + c_first), c_last - c_first) : 0; /* cells left of the image */
    const int cr = border_r ? max(P.w - XO, c_first) : c_last;            /* first cell right of the image */
    const int n = nl + (c_last - cr);
    for (int idx = lane; idx < ST_B * n; idx += 32)
    {
      const int r = idx / n, k = idx - r * n;
      const int c = k < nl ? c_first + k : cr + (k - nl);
      const int mc = st_mirror_once(XO + c, P.w) - XO;
      float *row = st_smem + G::OFF_SRC + ((ST_B * b + r) % G::SRC_ROWS) * G::S;
      if (mc >= 0 && mc < G::S)
        row[c] = row[mc];
    }
  };

  if (tid == ST_ISSUER)
  {
    tma_fence_proxy_async();
    issue(0);
    if (1 < n_steps)
      issue(1);
  }
  if (warp == ST_THREADS / 32 - 1 && (border_l || border_r))
    patch(0);
  __syncthreads();

  /* role of this warp: cb = column block (64 columns) inside the pass */
  const int cb = warp & 3, grp = warp >> 2;
  /* persistent register window of this warp's V role (a warp has one): the H rows in front of the rows it loads next */
  constexpr int RMAX = G::imax(G::R(0), G::imax(NL > 1 ? G::R(1) : 0, NL > 2 ? G::R(2) : 0));
  pk2 win[2 * RMAX + ST_B];
#pragma unroll
  for (int i = 0; i < 2 * RMAX + ST_B; i++)
    win[i] = 0ull;

#pragma unroll 1
  for (int s = 0; s < n_steps; s++)
  {
    if (tid == ST_ISSUER && s + 2 < n_steps)
    {
      tma_fence_proxy_async(); /* the slot's last readers passed the barrier that ended the previous step */
      issue(s + 2);
    }
    tma_mbar_wait(bar0 + 8u * (uint32_t)(s % ST_SRC_BLOCKS), (uint32_t)(s / ST_SRC_BLOCKS) & 1u); /* source rows of this step have landed */
    if (NL == 3)
    {
      /* fp32 operations per step and warp: V0 + H1, V1 + H0, V2, H2 -- e.g. (9 + 13, 13 + 9, 17, 17) x 512 pixels for radii (4, 6, 8) */
      if (grp == 0)
      {
        strip_v_role<G, 0>(P, st_smem, s, cb, lane, ry_end, XO, yb, win);
        strip_h_role<G, (NL > 1 ? 1 : 0)>(P, st_smem, s, cb, lane, ry_end);
      }
      else if (grp == 1)
      {
        strip_v_role<G, (NL > 1 ? 1 : 0)>(P, st_smem, s, cb, lane, ry_end, XO, yb, win);
        strip_h_role<G, 0>(P, st_smem, s, cb, lane, ry_end);
      }
      else if (grp == 2)
        strip_v_role<G, (NL > 2 ? 2 : 0)>(P, st_smem, s, cb, lane, ry_end, XO, yb, win);
      else
        strip_h_role<G, (NL > 2 ? 2 : 0)>(P, st_smem, s, cb, lane, ry_end);
    }
    else
    {
      if (grp == 0)
        strip_v_role<G, 0>(P, st_smem, s, cb, lane, ry_end, XO, yb, win);
      else if (grp == 1)
        strip_v_role<G, (NL > 1 ? 1 : 0)>(P, st_smem, s, cb, lane, ry_end, XO, yb, win);
      else if (grp == 2)
        strip_h_role<G, 0>(P, st_smem, s, cb, lane, ry_end);
      else
        strip_h_role<G, (NL > 1 ? 1 : 0)>(P, st_smem, s, cb, lane, ry_end);
    }
    if (warp == ST_THREADS / 32 - 1 && (border_l || border_r) && s + 1 < n_steps)
      patch(s + 1);
    __syncthreads();
  }
}

/* ---- host side -------------------------------------------------------------------------------------------------- */
typedef StripGeo<3, 4, 6, 8, 224> GeoA;  /* layers 1..3 of the default configuration (radii 4, 6, 8) */
typedef StripGeo<2, 10, 12, 0, 232> GeoC; /* layers 4, 5 (radii 10, 12) */

static int strip_even(int r) { return r < 2 ? 2 : ((r + 1) & ~1); }

/* kind of chain the kernel is built for, or -1: (0) three layers with radii 4, 6, 8; (1) two layers with radii 10, 12 */
int strip_chain_kind(const BlurPass *passes, int n)
{
  for (int i = 0; i < n; i++)
    if (passes[i].src_kind != BLUR_SRC_LAYER || passes[i].radius < 1)
      return -1;
  if (n == 3 && strip_even(passes[0].radius) == 4 && strip_even(passes[1].radius) == 6 && strip_even(passes[2].radius) == 8)
    return 0;
  if (n == 2 && strip_even(passes[0].radius) == 10 && strip_even(passes[1].radius) == 12)
    return 1;
  return -1;
}

bool strip_plan(const BlurPass *passes, int n, StripLaunch *out)
{
  const int kind = strip_chain_kind(passes, n);
  if (kind < 0)
    return false;
  const BlurPass &first = passes[0];
  /* large octaves only: a CTA walks its segment row block by row block (about a microsecond per step), which only pays when
   * the octave fills the GPU with strips x segments; smaller octaves keep the per-layer launches */
  if (first.w < 1024 || first.h < 256)
    return false;
  memset(out, 0, sizeof(*out));
  StripParams &P = *reinterpret_cast<StripParams *>(out->params);
  static_assert(sizeof(StripParams) <= sizeof(out->params), "StripLaunch::params too small");
  P.src = (const float *)first.src;
  P.w = first.w;
  P.h = first.h;
  P.pitch = first.dst_pitch;
  P.fp16 = first.fp16;
  for (int k = 0; k < n; k++)
  {
    if (passes[k].dst_pitch != first.dst_pitch || passes[k].src_pitch != first.dst_pitch || passes[k].dst_d == nullptr)
      return false;
    if (k > 0 && passes[k].src != passes[k - 1].dst_g)
      return false;
    if (passes[k].dst_next && k != n - 1)
      return false;
    P.g[k] = passes[k].dst_g;
    P.d[k] = passes[k].dst_d;
    for (int j = 0; j < 14; j++)
    {
      const float v = (j <= passes[k].radius) ? passes[k].taps[j] : 0.f;
      P.taps2[k][j] = make_float2(v, v);
    }
  }
  const BlurPass &last = passes[n - 1];
  P.next = last.dst_next;
  P.next_w = last.next_w;
  P.next_h = last.next_h;
  P.next_pitch = last.next_pitch;
  const int W = kind == 0 ? GeoA::W : GeoC::W;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  P.n_strips = (first.w + W - 1) / W;
  /* one CTA per SM, one wave: as many segments as fit, but segments of at least 128 rows (every segment recomputes 2 HY rows) */
  int n_segs = sms / P.n_strips;
  const int max_segs = (first.h + 127) / 128;
  if (n_segs > max_segs)
    n_segs = max_segs;
  if (n_segs < 1)
    n_segs = 1;
  int hseg = (first.h + n_segs - 1) / n_segs;
  hseg = (hseg + 1) & ~1;
  n_segs = (first.h + hseg - 1) / hseg;
  P.hseg = hseg;
  out->kind = kind;
  out->grid = P.n_strips * n_segs;
  out->first_layer = 0; /* filled by the caller */
  out->n_layers = n;
  return true;
}

template <class G, class K>
static cudaError_t strip_launch_kind(K kernel, const StripParams &P, int grid, cudaStream_t st)
{
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev])
  {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
    if (e != cudaSuccess)
      return e;
    attr_done[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(ST_THREADS);
  cfg.dynamicSmemBytes = G::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool no_pdl = [] { const char *e = getenv("VKSIFT_NO_PDL"); return e && e[0] == '1'; }();
  cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, P);
}

cudaError_t launch_strip(const StripLaunch &L, cudaStream_t st)
{
  const StripParams &P = *reinterpret_cast<const StripParams *>(L.params);
  if (L.kind == 0)
    return strip_launch_kind<GeoA>(pyramid_strip_kernel<3, 4, 6, 8, GeoA::W>, P, L.grid, st);
  return strip_launch_kind<GeoC>(pyramid_strip_kernel<2, 10, 12, 0, GeoC::W>, P, L.grid, st);
}

} // namespace vks
