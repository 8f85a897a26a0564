/*
 * match_tc.cuh -- 2-NN matcher on the 5th-generation tensor cores (tcgen05, sm_100a).
 *
 * dot[a][b] = sum_k A[a][k]*B[b][k] is a dense u8 x u8 -> s32 contraction with
 * K = 128 (one 128-byte row = one SWIZZLE_128B atom row, four K=32 MMAs):
 *
 *   TMA (cp.async.bulk.tensor.2d, 128B swizzle)  ->  smem A tile 128x128 B (once per CTA)
 *                                                ->  smem B tiles 128x128 B + 512 B of |b|^2, 4-stage ring
 *   tcgen05.mma.cta_group::1.kind::i8  M=128 N=128 K=32 x4 -> TMEM accumulator (2 x 128 columns, double buffered)
 *   4 epilogue warps: tcgen05.ld 32x32b.x32 -> e = |b|^2 - 2*dot, running top-2 per A row in registers
 *
 * One CTA owns 128 rows of A and a contiguous range of B tiles ("split");
 * partial top-2 keys ((d^2 << 32) | pos) go to HBM and a small merge kernel
 * folds the splits, applies the shader's tie rule and takes the square root.
 * Two CTAs fit one SM (256 TMEM columns and ~83 KB smem each), so one CTA's
 * MMA overlaps the other's epilogue on top of the in-CTA double buffering.
 *
 * Reference semantics: shaders/Get2NearestNeighbors.comp:43-104.
 */
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "vksift_internal.h"

namespace vks
{

#define MT_M 128        /* A rows per CTA = TMEM lanes */
#define MT_N 128        /* B rows per MMA tile = TMEM columns per buffer */
#define MT_STAGES 4     /* B smem ring */
#define MT_TILE_BYTES (MT_N * 128)
#define MT_THREADS 192  /* warps 0-3 epilogue, warp 4 TMA, warp 5 MMA + TMEM alloc */
#define MT_TMEM_COLS 256
#define MT_SMEM_BYTES (1024 + MT_M * 128 + MT_STAGES * (MT_TILE_BYTES + MT_N * 4) + 256)

/* ---- PTX wrappers --------------------------------------------------------- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  uint32_t done;
  do
  {
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t"
                 "}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n\t"
               ".reg .pred p;\n\t"
               "setp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
               "}" ::"r"(tmem_d),
               "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int32_t (&v)[32])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                 "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                 "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                 "=r"(v[31])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ int4 lds_v4(uint32_t saddr)
{
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

/* K-major SWIZZLE_128B operand descriptor: 8-row groups 1024 B apart (SBO), LBO unused (=1),
 * descriptor version 1, layout type 2.  (cute/arch/mma_sm100_desc.hpp SmemDescriptor) */
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr)
{
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;            /* leading byte offset (ignored for swizzled K-major) */
  d |= (uint64_t)(1024u >> 4) << 32; /* stride byte offset */
  d |= (uint64_t)1 << 46;            /* version */
  d |= (uint64_t)2 << 61;            /* SWIZZLE_128B */
  return d;
}
/* kind::i8 instruction descriptor: D=s32, A=B=u8, K-major both, N=128, M=128 */
#define MT_IDESC ((2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(MT_N >> 3) << 17) | ((uint32_t)(MT_M >> 4) << 24))

__device__ __forceinline__ uint32_t mt_pos(uint32_t b) { return b < 2u ? (b ^ 1u) : b; }

/* ---- main kernel ---------------------------------------------------------- */
__global__ void __launch_bounds__(MT_THREADS, 2)
    match_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const uint32_t *__restrict__ norm_a,
                    const uint32_t *__restrict__ norm_b, uint32_t na, uint32_t nb, uint32_t tiles_per_split, uint32_t n_tiles,
                    unsigned long long *__restrict__ partial, uint32_t na_pad)
{
  extern __shared__ uint8_t smem_raw[];
  /* 1024-byte alignment required by SWIZZLE_128B */
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *s_a = smem;
  uint8_t *s_b = smem + MT_M * 128;
  uint32_t *s_nb = (uint32_t *)(s_b + MT_STAGES * MT_TILE_BYTES);
  uint64_t *s_bar = (uint64_t *)(s_nb + MT_STAGES * MT_N);
  /* barriers: [0..S) full_b, [S..2S) empty_b, 2S full_a, 2S+1..2S+2 tmem_full, 2S+3..2S+4 tmem_empty */
  uint32_t *s_tmem = (uint32_t *)(s_bar + 2 * MT_STAGES + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row0 = blockIdx.x * MT_M;
  const uint32_t split = blockIdx.y;
  const uint32_t t_begin = split * tiles_per_split;
  const uint32_t t_end = min(n_tiles, t_begin + tiles_per_split);
  const uint32_t my_tiles = (t_end > t_begin) ? (t_end - t_begin) : 0u;

  const uint32_t bar0 = smem_u32(s_bar);
#define BAR_FULL_B(i) (bar0 + 8u * (uint32_t)(i))
#define BAR_EMPTY_B(i) (bar0 + 8u * (uint32_t)(MT_STAGES + (i)))
#define BAR_FULL_A (bar0 + 8u * (uint32_t)(2 * MT_STAGES))
#define BAR_TMEM_FULL(i) (bar0 + 8u * (uint32_t)(2 * MT_STAGES + 1 + (i)))
#define BAR_TMEM_EMPTY(i) (bar0 + 8u * (uint32_t)(2 * MT_STAGES + 3 + (i)))

  if (threadIdx.x == 0)
  {
    for (int i = 0; i < MT_STAGES; i++)
    {
      mbar_init(BAR_FULL_B(i), 1);
      mbar_init(BAR_EMPTY_B(i), 1 + 4); /* MMA commit + 4 epilogue warps (they read the stage's |b|^2) */
    }
    mbar_init(BAR_FULL_A, 1);
    for (int i = 0; i < 2; i++)
    {
      mbar_init(BAR_TMEM_FULL(i), 1);
      mbar_init(BAR_TMEM_EMPTY(i), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(MT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 4)
  {
    /* ===== TMA producer ===== */
    if (lane == 0)
    {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
      mbar_expect_tx(BAR_FULL_A, MT_M * 128);
      tma_load_2d(smem_u32(s_a), &map_a, 0, (int)row0, BAR_FULL_A);
      for (uint32_t t = 0; t < my_tiles; t++)
      {
        const uint32_t st = t % MT_STAGES, ph = (t / MT_STAGES) & 1u;
        mbar_wait(BAR_EMPTY_B(st), ph ^ 1u);
        const uint32_t b0 = (t_begin + t) * MT_N;
        mbar_expect_tx(BAR_FULL_B(st), MT_TILE_BYTES + MT_N * 4);
        tma_load_2d(smem_u32(s_b + st * MT_TILE_BYTES), &map_b, 0, (int)b0, BAR_FULL_B(st));
        bulk_load_1d(smem_u32(s_nb + st * MT_N), norm_b + b0, MT_N * 4, BAR_FULL_B(st));
      }
    }
  }
  else if (warp == 5)
  {
    /* ===== MMA issuer (one thread) ===== */
    if (lane == 0)
    {
      mbar_wait(BAR_FULL_A, 0);
      const uint64_t adesc = umma_smem_desc(smem_u32(s_a));
      for (uint32_t t = 0; t < my_tiles; t++)
      {
        const uint32_t st = t % MT_STAGES, ph = (t / MT_STAGES) & 1u;
        const uint32_t buf = t & 1u, bph = (t >> 1) & 1u;
        mbar_wait(BAR_TMEM_EMPTY(buf), bph ^ 1u);
        mbar_wait(BAR_FULL_B(st), ph);
        tmem_fence_after();
        const uint64_t bdesc = umma_smem_desc(smem_u32(s_b + st * MT_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < 4; k++) /* K = 32 bytes per MMA: advance the start address by 32 B (>>4 = 2) */
          umma_i8(tmem_base + buf * MT_N, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), MT_IDESC, k > 0 ? 1u : 0u);
        umma_commit(BAR_EMPTY_B(st));     /* smem stage consumed by the tensor core */
        umma_commit(BAR_TMEM_FULL(buf)); /* accumulator ready */
      }
    }
  }
  else
  {
    /* ===== epilogue: thread = A row, warp w reads TMEM lanes 32w..32w+31 =====
     * Hot loop per accumulator: one IMAD (e = |b|^2 - 2 a.b) and a share of a min tree; only when the
     * minimum of a 16-column group beats the row's current second best does the thread rescan that group
     * with full (d^2, pos) keys.  Columns past nb carry |b|^2 = 2^30 (norms are padded), so they never win. */
    const uint32_t row = row0 + warp * 32 + lane;
    const int32_t my_na = (row < na) ? (int32_t)norm_a[row] : 0;
    unsigned long long k1 = ~0ull, k2 = ~0ull;
    int32_t thr = 0x7fffffff; /* e-threshold: second-best d^2 - |a|^2 */
    for (uint32_t t = 0; t < my_tiles; t++)
    {
      const uint32_t st = t % MT_STAGES, ph = (t / MT_STAGES) & 1u;
      const uint32_t buf = t & 1u, bph = (t >> 1) & 1u;
      const uint32_t b0 = (t_begin + t) * MT_N;
      mbar_wait(BAR_FULL_B(st), ph); /* |b|^2 of this tile landed (same barrier as the B tile) */
      mbar_wait(BAR_TMEM_FULL(buf), bph);
      tmem_fence_after();
      const uint32_t nbs = smem_u32(s_nb + st * MT_N);
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * MT_N;
#pragma unroll 1
      for (int c0 = 0; c0 < MT_N; c0 += 32)
      {
        int32_t acc[32];
        tmem_ld32(taddr + c0, acc);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 2; g++)
        {
          int32_t e[16];
#pragma unroll
          for (int q = 0; q < 4; q++)
          {
            const int4 n4 = lds_v4(nbs + (uint32_t)(c0 + g * 16 + q * 4) * 4u);
            e[4 * q + 0] = n4.x - 2 * acc[g * 16 + 4 * q + 0];
            e[4 * q + 1] = n4.y - 2 * acc[g * 16 + 4 * q + 1];
            e[4 * q + 2] = n4.z - 2 * acc[g * 16 + 4 * q + 2];
            e[4 * q + 3] = n4.w - 2 * acc[g * 16 + 4 * q + 3];
          }
          int32_t m = e[0];
#pragma unroll
          for (int q = 1; q < 16; q++)
            m = min(m, e[q]);
          if (m <= thr)
          {
#pragma unroll
            for (int q = 0; q < 16; q++)
            {
              if (e[q] <= thr)
              {
                const unsigned long long key = ((unsigned long long)(uint32_t)(e[q] + my_na) << 32) | mt_pos(b0 + c0 + g * 16 + q);
                if (key < k1)
                {
                  k2 = k1;
                  k1 = key;
                }
                else if (key < k2)
                  k2 = key;
                if (k2 != ~0ull)
                  thr = (int32_t)(uint32_t)(k2 >> 32) - my_na;
              }
            }
          }
        }
      }
      tmem_fence_before();
      __syncwarp();
      if (lane == 0)
      {
        mbar_arrive(BAR_TMEM_EMPTY(buf));
        mbar_arrive(BAR_EMPTY_B(st));
      }
    }
    if (row < na_pad)
    {
      partial[((size_t)split * na_pad + row) * 2 + 0] = k1;
      partial[((size_t)split * na_pad + row) * 2 + 1] = k2;
    }
  }

  tmem_fence_before();
  __syncthreads();
  if (warp == 5)
  {
    tmem_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(MT_TMEM_COLS) : "memory");
  }
}

/* fold the splits, undo the pos permutation, sqrt (Get2NearestNeighbors.comp:98-102) */
__global__ void match_merge_kernel(const unsigned long long *__restrict__ partial, uint32_t splits, uint32_t na, uint32_t na_pad,
                                   vksift_Match_2NN *__restrict__ out)
{
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= na)
    return;
  unsigned long long k1 = ~0ull, k2 = ~0ull;
  for (uint32_t s = 0; s < splits; s++)
  {
#pragma unroll
    for (int q = 0; q < 2; q++)
    {
      const unsigned long long key = partial[((size_t)s * na_pad + row) * 2 + q];
      if (key < k1)
      {
        k2 = k1;
        k1 = key;
      }
      else if (key < k2)
        k2 = key;
    }
  }
  vksift_Match_2NN m;
  m.idx_a = row;
  m.idx_b1 = mt_pos((uint32_t)k1);
  m.idx_b2 = mt_pos((uint32_t)k2);
  m.dist_a_b1 = vks_sqrt((float)(uint32_t)(k1 >> 32));
  m.dist_a_b2 = vks_sqrt((float)(uint32_t)(k2 >> 32));
  out[row] = m;
}

/* ---- host side ------------------------------------------------------------ */
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct MatchTc
{
  PFN_encodeTiled encode;
  unsigned long long *partial;
  size_t partial_elems;
  int sm_count;
};

static cudaError_t match_tc_create(void **out, uint32_t max_feats)
{
  MatchTc *tc = new MatchTc();
  tc->partial = nullptr;
  tc->partial_elems = 0;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess)
  {
    delete tc;
    return e != cudaSuccess ? e : cudaErrorNotSupported;
  }
  tc->encode = (PFN_encodeTiled)fn;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&tc->sm_count, cudaDevAttrMultiProcessorCount, dev);
  e = cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM_BYTES);
  if (e != cudaSuccess)
  {
    delete tc;
    return e;
  }
  /* worst case: every row block split into all B tiles is never needed; sized on demand */
  (void)max_feats;
  *out = tc;
  return cudaSuccess;
}

static void match_tc_destroy(void *p)
{
  MatchTc *tc = (MatchTc *)p;
  if (!tc)
    return;
  cudaFree(tc->partial);
  delete tc;
}

static bool mt_make_map(MatchTc *tc, CUtensorMap *map, const uint8_t *base, uint32_t rows)
{
  const cuuint64_t gdim[2] = {128, rows};
  const cuuint64_t gstride[1] = {128};
  const cuuint32_t box[2] = {128, MT_N};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = tc->encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static cudaError_t match_tc_launch(void *p, const uint8_t *da, uint32_t na, const uint32_t *norm_a, const uint8_t *db, uint32_t nb,
                                   const uint32_t *norm_b, vksift_Match_2NN *out, cudaStream_t st, uint64_t *launch_count)
{
  MatchTc *tc = (MatchTc *)p;
  const uint32_t row_blocks = (na + MT_M - 1) / MT_M;
  const uint32_t n_tiles = (nb + MT_N - 1) / MT_N;
  /* two CTAs are resident per SM: aim at one full wave of 2*SMs CTAs */
  uint32_t splits = (2u * (uint32_t)tc->sm_count) / row_blocks;
  if (splits < 1)
    splits = 1;
  if (splits > n_tiles)
    splits = n_tiles;
  const uint32_t tiles_per_split = (n_tiles + splits - 1) / splits;
  splits = (n_tiles + tiles_per_split - 1) / tiles_per_split;
  const uint32_t na_pad = row_blocks * MT_M;
  const size_t need = (size_t)splits * na_pad * 2;
  if (need > tc->partial_elems)
  {
    /* grows only when a larger problem shows up; stream-ordered with respect to earlier matches */
    cudaStreamSynchronize(st);
    cudaFree(tc->partial);
    tc->partial = nullptr;
    tc->partial_elems = 0;
    cudaError_t e = cudaMalloc(&tc->partial, need * sizeof(unsigned long long));
    if (e != cudaSuccess)
      return e;
    tc->partial_elems = need;
  }
  CUtensorMap map_a, map_b;
  if (!mt_make_map(tc, &map_a, da, na) || !mt_make_map(tc, &map_b, db, nb))
    return cudaErrorInvalidValue;
  dim3 grid(row_blocks, splits, 1);
  match_tc_kernel<<<grid, MT_THREADS, MT_SMEM_BYTES, st>>>(map_a, map_b, norm_a, norm_b, na, nb, tiles_per_split, n_tiles, tc->partial, na_pad);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return e;
  match_merge_kernel<<<(na + 255) / 256, 256, 0, st>>>(tc->partial, splits, na, na_pad, out);
  *launch_count += 2;
  return cudaGetLastError();
}

} // namespace vks
