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
#define ST_H_BLOCKS 3   /* H ring: written by H_k, consumed by V_k one step later, which leaves the DoG rows in its place for the store */

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
  /* rows of the G_k ring (k < NL-1), a power of two >= 3B + R(k+1): written by V_k (slot of region row ry = (ry + offV(k)) & (DG-1),
   * so that the eight rows of a step are one aligned block), read by H_k+1 one step later and by the DoG of V_k+1 */
  __host__ __device__ static constexpr int DG(int k) { return (3 * B + R(k + 1)) <= 32 ? 32 : 64; }
  static constexpr int SRC_ROWS = ST_SRC_BLOCKS * B;
  /* shared memory layout, floats */
  static constexpr int OFF_SRC = 0;
  __host__ __device__ static constexpr int off_h(int k) { return SRC_ROWS * S + k * ST_H_BLOCKS * B * S; }
  __host__ __device__ static constexpr int off_g(int k) { return SRC_ROWS * S + NL * ST_H_BLOCKS * B * S + (k > 0 ? DG(0) * S : 0); }
  static constexpr int FLOATS = SRC_ROWS * S + NL * ST_H_BLOCKS * B * S + (NL > 1 ? DG(0) * S : 0) + (NL > 2 ? DG(1) * S : 0);
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
/* shared -> global bulk copy (TMA unit): one instruction stores a whole row segment */
__device__ __forceinline__ void st_bulk_store(void *dst, uint32_t src, uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void st_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void st_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

#define ST_TAP(i) (*reinterpret_cast<const pk2 *>(&taps2[i]))

/* horizontal pass: 16 outputs of one row from a window of 16 + 2 RX inputs (LDS.128), written with STS.128.
 * Output pairs (x, x+1).  wp[j] = (in[2j], in[2j+1]) are the loaded register pairs, wq[j] = (in[2j+1], in[2j+2]) the same
 * window shifted by one element (two register moves per pair, once per window): even taps add aligned pairs of wp, odd
 * taps aligned pairs of wq, so every tap of every output pair is one FADD2 + one FFMA2. */
template <int R>
__device__ __forceinline__ void strip_h_unit(const float *__restrict__ in, float *__restrict__ out, const float2 *__restrict__ taps2)
{
  constexpr int RX = (R + 3) & ~3;
  constexpr int WN = 16 + 2 * RX;
  pk2 wp[WN / 2], wq[WN / 2 - 1];
  const ulonglong2 *wsrc = reinterpret_cast<const ulonglong2 *>(in);
#pragma unroll
  for (int j = 0; j < WN / 4; j++)
  {
    const ulonglong2 v = wsrc[j];
    wp[2 * j] = v.x;
    wp[2 * j + 1] = v.y;
  }
#pragma unroll
  for (int j = 0; j < WN / 2 - 1; j++)
    wq[j] = pk_make(pk_hi(wp[j]), pk_lo(wp[j + 1]));
  ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(out);
  pk2 acc[8]; /* output pair j covers columns (2j, 2j+1) of the unit = window elements RX + 2j, RX + 2j + 1 */
#pragma unroll
  for (int j = 0; j < 8; j++)
    acc[j] = pk_mul(wp[RX / 2 + j], ST_TAP(0));
#pragma unroll
  for (int i = 1; i <= R; i++)
  {
#pragma unroll
    for (int j = 0; j < 8; j++)
    {
      const int c = RX / 2 + j; /* pair index of the centre */
      pk2 sum;
      if ((i & 1) == 0)
        sum = pk_add(wp[c + i / 2], wp[c - i / 2]);
      else /* (in[2c+i], in[2c+1+i]) = wq[c + (i-1)/2], (in[2c-i], in[2c+1-i]) = wq[c - (i+1)/2] */
        sum = pk_add(wq[c + (i - 1) / 2], wq[c - (i + 1) / 2]);
      acc[j] = pk_fma(sum, ST_TAP(i), acc[j]);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; q++)
    dst[q] = make_ulonglong2(acc[2 * q], acc[2 * q + 1]);
}

/* per-CTA constants of a launch */
struct StripCtx
{
  int XO, yb, ry_end, x0;
  int store_cols; /* columns of a stored row segment (multiple of 4) */
};

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
  const int cs = G::hu_lo(K) + 64 * cb + 16 * g;
  /* the eight input rows are one aligned block of the ring they live in */
  const float *in_row;
  if (K == 0)
    in_row = smem + G::OFF_SRC + ((s % ST_SRC_BLOCKS) * ST_B + r) * G::S;
  else
    in_row = smem + G::off_g(K > 0 ? K - 1 : 0) + (((ST_B * (s - 1)) & (G::DG(K > 0 ? K - 1 : 0) - 1)) + r) * G::S;
  float *out = smem + G::off_h(K) + ((s % ST_H_BLOCKS) * ST_B + r) * G::S + cs;
  strip_h_unit<R>(in_row + cs - RX, out, P.taps2[K]);
}

/* vertical pass of layer K for one column pair: `win` holds the H rows [a - R, a + R) in front of the 8 rows loaded here.
 * Results: G_K into its ring (not for the chain's last layer, which goes to global memory from registers), DoG in place
 * of the consumed H rows; both are stored to global memory by bulk copies one step later (strip_store_rows). */
template <class G, int K, bool FP16>
__device__ __forceinline__ void strip_v_role(const StripParams &P, float *smem, int s, int cb, int lane, const StripCtx &C, pk2 (&win)[2 * G::R(K) + ST_B])
{
  constexpr int R = G::R(K);
  constexpr bool LAST = (K == G::NL - 1);
  constexpr int KP = K > 0 ? K - 1 : 0;
  if (cb >= G::vu_n(K) || s < 1)
    return;
  const float2 *__restrict__ taps2 = P.taps2[K];
  const int a = ST_B * s - G::offV(K); /* first output row (region coordinates) */
  const int col_raw = G::clo(K) + 64 * cb + 2 * lane;
  const bool lane_ok = col_raw < G::chi(K);
  const int col = lane_ok ? col_raw : G::clo(K);
  /* the 8 newest H rows: written by H_K in the previous step */
  float *hblk = smem + G::off_h(K) + (((s - 1) % ST_H_BLOCKS) * ST_B) * G::S + col;
#pragma unroll
  for (int i = 0; i < ST_B; i++)
    win[2 * R + i] = *reinterpret_cast<const pk2 *>(hblk + i * G::S);
  if (a + ST_B > G::HY - G::halo(K) && a < C.ry_end + G::halo(K))
  {
    pk2 acc[ST_B];
#pragma unroll
    for (int j = 0; j < ST_B; j++)
      acc[j] = pk_mul(win[R + j], ST_TAP(0));
#pragma unroll
    for (int i = 1; i <= R; i++)
    {
#pragma unroll
      for (int j = 0; j < ST_B; j++)
        acc[j] = pk_fma(pk_add(win[R + j + i], win[R + j - i]), ST_TAP(i), acc[j]);
    }
    if (FP16)
    {
#pragma unroll
      for (int j = 0; j < ST_B; j++)
        acc[j] = round_half2(acc[j]);
    }
    if (lane_ok)
    {
      /* DoG centre = layer K-1 at the same pixel: rows a + j of the source ring (blocks of 8 rows) or of the G_K-1 ring */
      const int cb0 = (K == 0) ? (s - 1 + 2 * ST_SRC_BLOCKS) : 0; /* block of row 8(s-1); rows a + j = 8(s-1) + (j - R) for K = 0 */
      const int g_base = (a + G::offV(KP)) & (G::DG(KP) - 1);      /* slot of row a in the G_K-1 ring (K > 0) */
      float *gring = smem + G::off_g(LAST ? 0 : K) + ((ST_B * s) & (G::DG(LAST ? 0 : K) - 1)) * G::S + col;
#pragma unroll
      for (int j = 0; j < ST_B; j++)
      {
        if (!LAST)
          *reinterpret_cast<pk2 *>(gring + j * G::S) = acc[j];
        const float *crow;
        if (K == 0)
        {
          const int q = j - R;                                    /* row 8(s-1) + q */
          const int fb = (q >= 0) ? q / 8 : -((7 - q) / 8);       /* floor(q / 8), compile time */
          const int rr = q - 8 * fb;
          crow = smem + G::OFF_SRC + (((cb0 + fb) % ST_SRC_BLOCKS) * ST_B + rr) * G::S;
        }
        else
          crow = smem + G::off_g(KP) + ((g_base + j) & (G::DG(KP) - 1)) * G::S;
        pk2 dd = pk_sub(acc[j], *reinterpret_cast<const pk2 *>(crow + col));
        if (FP16)
          dd = round_half2(dd);
        *reinterpret_cast<pk2 *>(hblk + j * G::S) = dd;
      }
      tma_fence_proxy_async(); /* these rows leave through the TMA unit (async proxy) after the step's barrier */
    }
    if (LAST)
    {
      /* the chain's last Gaussian layer and the decimated seed of the next octave: straight from the registers */
      const int x = C.XO + col;
      const bool col_store = lane_ok && col >= G::HXA && col < G::HXA + G::W && x < P.w;
      const int j_lo = max(G::HY - a, 0), j_n = min(C.ry_end - a, ST_B) - j_lo; /* rows [j_lo, j_lo + j_n) are stored */
      const int y0 = C.yb + a;
      float *gp = P.g[K] + (size_t)y0 * (size_t)P.pitch + (size_t)x;
      const bool nx_ok = P.next != nullptr && (x >> 1) < P.next_w;
      if (col_store)
      {
#pragma unroll
        for (int j = 0; j < ST_B; j++)
        {
          if ((unsigned)(j - j_lo) < (unsigned)j_n)
          {
            *reinterpret_cast<pk2 *>(gp + (size_t)j * (size_t)P.pitch) = acc[j]; /* read back by the next chain's launch: keep it in L2 */
            const int y = y0 + j;
            if (nx_ok && (y & 1) && (y >> 1) < P.next_h) /* x even, y odd: the odd column of the pair feeds next(x >> 1, y >> 1) */
              P.next[(size_t)(y >> 1) * (size_t)P.next_pitch + (size_t)(x >> 1)] = pk_hi(acc[j]);
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

/* One lane = one row: bulk copies (TMA unit, shared -> global) of the rows V_K emitted in the previous step.
 * what = 0: Gaussian layer K from its ring; what = 1: DoG layer K from the H block V_K consumed. */
template <class G, int K>
__device__ __forceinline__ void strip_store_rows(const StripParams &P, float *smem, int s, int lane, const StripCtx &C, int what)
{
  if (lane >= ST_B || s < 2)
    return;
  const int sp = s - 1;                       /* the step that produced the rows */
  const int ry = ST_B * sp - G::offV(K) + lane;
  if (ry < G::HY || ry >= C.ry_end)
    return;
  const float *src;
  float *dst;
  if (what == 0)
  {
    src = smem + G::off_g(K < G::NL - 1 ? K : 0) + (((ST_B * sp) & (G::DG(K < G::NL - 1 ? K : 0) - 1)) + lane) * G::S + G::HXA;
    dst = P.g[K];
  }
  else
  {
    src = smem + G::off_h(K) + (((sp - 1) % ST_H_BLOCKS) * ST_B + lane) * G::S + G::HXA;
    dst = P.d[K];
  }
  dst += (size_t)(C.yb + ry) * (size_t)P.pitch + (size_t)C.x0;
  tma_fence_proxy_async(); /* the rows were written with ordinary stores before the barrier */
  st_bulk_store(dst, tma_smem_u32(src), (uint32_t)C.store_cols * 4u);
  st_bulk_commit();
}

template <int NL, int R0, int R1, int R2, int W, bool FP16>
__global__ void __launch_bounds__(ST_THREADS, 1) pyramid_strip_kernel(const __grid_constant__ StripParams P)
{
  using G = StripGeo<NL, R0, R1, R2, W>;
  extern __shared__ __align__(128) float st_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = tma_smem_u32(st_smem + G::FLOATS);

  const int strip = (int)blockIdx.x % P.n_strips, seg = (int)blockIdx.x / P.n_strips;
  StripCtx C;
  C.x0 = strip * W;
  C.XO = C.x0 - G::HXA;
  const int XO = C.XO;
  const int y0s = seg * P.hseg, y1s = min(P.h, y0s + P.hseg);
  C.yb = y0s - G::HY;             /* image row of region row 0 */
  C.ry_end = G::HY + (y1s - y0s); /* region rows [HY, ry_end) are stored */
  C.store_cols = min(W, (P.w - C.x0 + 3) & ~3);
  const int yb = C.yb;
  const int n_steps = (C.ry_end + G::offV(NL - 1) + ST_B - 1) / ST_B;

  /* source fetch: columns [fx0, fx1) of the layer, 16-byte aligned on both sides, clipped to the padded row */
  const int fx0 = max(XO + G::src_lo, 0), fx1 = min(XO + G::src_hi, P.pitch);
  const uint32_t row_bytes = (uint32_t)(fx1 - fx0) * 4u;
  const int dcol = fx0 - XO;
  const bool border_l = (XO + G::HXA - G::HX) < 0, border_r = (XO + G::HXA + W + G::HX) > P.w;

  /* Roles.  A warp scheduler (SMSP = warp % 4) hosts two role groups only, so its four warps share code, and the groups are
   * paired so that the four schedulers carry about the same number of instructions per step:
   *   NL = 3:  group 0 = V0 + H1, group 1 = V1 + H0, group 2 = V2, group 3 = H2   (SMSP 0, 1: groups 0 and 3; SMSP 2, 3: groups 1 and 2)
   *   NL = 2:  group 0 = V0,      group 1 = V1,      group 2 = H0, group 3 = H1
   * cb = block of 64 columns inside the pass.  The lightest group also drives the TMA unit (source loads, border patch, row stores). */
  const int smsp = warp & 3, widx = warp >> 2;
  const int grp = (smsp < 2) ? (widx < 2 ? 0 : 3) : (widx < 2 ? 1 : 2);
  const int cb = (smsp & 1) * 2 + (widx & 1);
  constexpr int XGRP = (NL == 3) ? 3 : 2;   /* group with the fewest instructions per step */
  const bool is_issuer = (grp == XGRP && cb == 0 && lane == 0);
  const bool is_patcher = (grp == XGRP && cb == 1);
  /* store duty: (what, layer) combinations spread over warps; NL = 3: G0, G1, D0, D1, D2; NL = 2: G0, D0, D1 */
  int store_what = -1, store_k = 0;
  if (grp == XGRP)
  {
    if (cb == 2)
      store_what = 1, store_k = 0;
    else if (cb == 3)
      store_what = 1, store_k = 1;
    else if (cb == 0)
      store_what = 0, store_k = 0;
    else if (NL == 3)
      store_what = 0, store_k = 1;
  }
  else if (NL == 3 && grp == 2 && cb == 0)
    store_what = 1, store_k = 2;

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
      st_bulk_load(tma_smem_u32(st_smem + G::OFF_SRC + ((b % ST_SRC_BLOCKS) * ST_B + i) * G::S + dcol), P.src + (size_t)gy * (size_t)P.pitch + (size_t)fx0,
                   row_bytes, bar);
    }
  };
  /* cells of the source rows of block b that lie outside the image take the value of their mirror cell (one warp: four lanes
   * per row).  Only the HX cells next to the image matter: nothing that is stored depends on cells further out. */
  auto patch = [&](int b) {
    tma_mbar_wait(bar0 + 8u * (uint32_t)(b % ST_SRC_BLOCKS), (uint32_t)(b / ST_SRC_BLOCKS) & 1u);
    const int nl = border_l ? min(-XO, G::HXA) - (G::HXA - G::HX) : 0; /* cells [HXA - HX, -XO) are left of the image */
    const int cr = P.w - XO;                                          /* first cell right of the image */
    const int nr = border_r ? min(G::HX, G::HXA + W + G::HX - cr) : 0;
    float *row = st_smem + G::OFF_SRC + ((b % ST_SRC_BLOCKS) * ST_B + (lane >> 2)) * G::S;
    for (int k = lane & 3; k < nl + nr; k += 4)
    {
      const int c = k < nl ? (G::HXA - G::HX) + k : cr + (k - nl);
      const int mc = st_mirror_once(XO + c, P.w) - XO;
      if (mc >= 0 && mc < G::S)
        row[c] = row[mc];
    }
  };

  if (is_issuer)
  {
    tma_fence_proxy_async();
    issue(0);
    if (1 < n_steps)
      issue(1);
  }
  if (is_patcher && (border_l || border_r))
    patch(0);
  __syncthreads();

  /* what every warp does at the top and at the bottom of a step */
  auto step_begin = [&](int s) {
    if (is_issuer && s + 2 < n_steps)
    {
      tma_fence_proxy_async(); /* the slot's last readers passed the barrier that ended the previous step */
      issue(s + 2);
    }
    if (store_what >= 0)
    {
      if (store_k == 0)
        strip_store_rows<G, 0>(P, st_smem, s, lane, C, store_what);
      else if (store_k == 1)
        strip_store_rows<G, (NL > 1 ? 1 : 0)>(P, st_smem, s, lane, C, store_what);
      else
        strip_store_rows<G, (NL > 2 ? 2 : 0)>(P, st_smem, s, lane, C, store_what);
    }
    if (s < n_steps)
      tma_mbar_wait(bar0 + 8u * (uint32_t)(s % ST_SRC_BLOCKS), (uint32_t)(s / ST_SRC_BLOCKS) & 1u); /* source rows of this step have landed */
  };
  auto step_end = [&](int s) {
    if (is_patcher && (border_l || border_r) && s + 1 < n_steps)
      patch(s + 1);
    if (store_what >= 0 && lane < ST_B)
      st_bulk_wait_read(); /* the rows stored at the top of this step have left shared memory: their slots may be rewritten */
    __syncthreads();
  };

  /* one loop per role group: a warp only carries the registers and the code of its own roles */
  if (grp == 0)
  {
    pk2 win[2 * G::R(0) + ST_B];
#pragma unroll
    for (int i = 0; i < 2 * G::R(0) + ST_B; i++)
      win[i] = 0ull;
#pragma unroll 1
    for (int s = 0; s < n_steps; s++)
    {
      step_begin(s);
      strip_v_role<G, 0, FP16>(P, st_smem, s, cb, lane, C, win);
      if (NL == 3)
        strip_h_role<G, (NL > 1 ? 1 : 0)>(P, st_smem, s, cb, lane, C.ry_end);
      step_end(s);
    }
  }
  else if (grp == 1)
  {
    constexpr int K1 = NL > 1 ? 1 : 0;
    pk2 win[2 * G::R(K1) + ST_B];
#pragma unroll
    for (int i = 0; i < 2 * G::R(K1) + ST_B; i++)
      win[i] = 0ull;
#pragma unroll 1
    for (int s = 0; s < n_steps; s++)
    {
      step_begin(s);
      strip_v_role<G, K1, FP16>(P, st_smem, s, cb, lane, C, win);
      if (NL == 3)
        strip_h_role<G, 0>(P, st_smem, s, cb, lane, C.ry_end);
      step_end(s);
    }
  }
  else if (grp == 2)
  {
    constexpr int K2 = NL > 2 ? 2 : 0;
    pk2 win[2 * G::R(K2) + ST_B];
#pragma unroll
    for (int i = 0; i < 2 * G::R(K2) + ST_B; i++)
      win[i] = 0ull;
#pragma unroll 1
    for (int s = 0; s < n_steps; s++)
    {
      step_begin(s);
      if (NL == 3)
        strip_v_role<G, K2, FP16>(P, st_smem, s, cb, lane, C, win);
      else
        strip_h_role<G, 0>(P, st_smem, s, cb, lane, C.ry_end);
      step_end(s);
    }
  }
  else
  {
#pragma unroll 1
    for (int s = 0; s < n_steps; s++)
    {
      step_begin(s);
      strip_h_role<G, NL - 1>(P, st_smem, s, cb, lane, C.ry_end);
      step_end(s);
    }
  }
  /* the rows of the last step */
  step_begin(n_steps);
  if (store_what >= 0 && lane < ST_B)
    st_bulk_wait_all();
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
  if (first.w < 1024 || first.h < 256 || first.fp16) /* binary16 layers: per-layer launches */
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
    P.g[k] = (float *)passes[k].dst_g;
    P.d[k] = (float *)passes[k].dst_d;
    for (int j = 0; j < 14; j++)
    {
      const float v = (j <= passes[k].radius) ? passes[k].taps[j] : 0.f;
      P.taps2[k][j] = make_float2(v, v);
    }
  }
  const BlurPass &last = passes[n - 1];
  P.next = (float *)last.dst_next;
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
  static bool attr_done[2][64] = {{false}};
  int dev = 0;
  cudaGetDevice(&dev);
  const int v = P.fp16 ? 1 : 0; /* the function-local statics are per geometry: one flag per precision variant */
  if (dev < 64 && !attr_done[v][dev])
  {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
    if (e != cudaSuccess)
      return e;
    attr_done[v][dev] = true;
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
    return P.fp16 ? strip_launch_kind<GeoA>(pyramid_strip_kernel<3, 4, 6, 8, GeoA::W, true>, P, L.grid, st)
                  : strip_launch_kind<GeoA>(pyramid_strip_kernel<3, 4, 6, 8, GeoA::W, false>, P, L.grid, st);
  return P.fp16 ? strip_launch_kind<GeoC>(pyramid_strip_kernel<2, 10, 12, 0, GeoC::W, true>, P, L.grid, st)
                : strip_launch_kind<GeoC>(pyramid_strip_kernel<2, 10, 12, 0, GeoC::W, false>, P, L.grid, st);
}

} // namespace vks
