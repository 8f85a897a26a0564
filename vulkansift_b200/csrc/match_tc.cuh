/*
 * match_tc.cuh -- 2-NN matcher on the 5th-generation tensor cores (tcgen05, sm_100a).
 *
 * dot[a][b] = sum_k A[a][k]*B[b][k] is a dense u8 x u8 -> s32 contraction with
 * K = 128 (one 128-byte row = one SWIZZLE_128B atom row, four K=32 MMAs):
 *
 *   TMA (cp.async.bulk.tensor.2d, 128B swizzle)  ->  smem A tile 128x128 B (once per CTA)
 *                                                ->  smem B tiles 128x128 B + 512 B of packed |b|^2, 4-stage ring
 *   tcgen05.mma.cta_group::1.kind::i8  M=128 N=128 K=32 x4 -> TMEM accumulators (2 x 128 columns, one per epilogue warpgroup)
 *   2 x 4 epilogue warps (one warpgroup per TMEM buffer): tcgen05.ld 32x32b.x32 -> packed key
 *     256*(|b|^2 - 2 a.b) + (pos & 255) with one IMAD, branch-free running top-2 of the keys (min/max only)
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
#define MT_N 128         /* B rows per MMA tile = TMEM columns per accumulator buffer */
#define MT_STAGES 4     /* B smem ring */
#define MT_BUFS 2       /* TMEM accumulator buffers, one per epilogue warpgroup (N=64 x 4 buffers measured 10% slower) */
#define MT_TILE_BYTES (MT_N * 128)
#define MT_THREADS 320  /* warps 0-3 and 4-7: two epilogue warpgroups (one per TMEM buffer), warp 8 TMA, warp 9 MMA + TMEM alloc */
#define MT_WARP_TMA 8
#define MT_WARP_MMA 9
#define MT_TMEM_COLS 256
#define MT_GROUP 16      /* columns per group of the epilogue (minimum per group, winner group rescanned by the merge) */
#define MT_SMEM_BYTES (1024 + MT_M * 128 + MT_STAGES * (MT_TILE_BYTES + MT_N * 4) + 512)

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
/* wait::ld that also names the destination registers of the load it completes, so the compiler cannot move
 * any use of them above the wait (tcgen05.ld fills its registers asynchronously) */
__device__ __forceinline__ void tmem_ld_wait_for(int32_t (&v)[32])
{
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]),
                 "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]),
                 "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]),
                 "+r"(v[31])
               :
               : "memory");
}

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
/* kind::i8 instruction descriptor: D=s32, A=B=u8, K-major both, N=MT_N, M=128 */
#define MT_IDESC ((2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(MT_N >> 3) << 17) | ((uint32_t)(MT_M >> 4) << 24))

__device__ __forceinline__ uint32_t mt_pos(uint32_t b) { return b < 2u ? (b ^ 1u) : b; }

/* One launch searches buffer A against up to MT_MAX_GROUPS descriptor blocks ("groups") that lie in one allocation at a
 * common stride: the all-pairs step (A against the blocks received from every peer).  A "virtual row block" v = (group g,
 * row block rb) takes the A tile of rb and the B tiles of block blk[g]; everything that was indexed by the row block
 * (segments, partial keys) is indexed by v.  A single search is one group at block 0. */
#define MT_MAX_GROUPS 64
struct MatchGroups
{
  uint32_t n_groups;
  uint32_t stride_rows;        /* rows between the starts of consecutive blocks (multiple of MT_N) */
  uint32_t out_stride;         /* records between the result lists of consecutive blocks */
  uint32_t row_blocks;         /* row blocks of A */
  uint32_t blk[MT_MAX_GROUPS]; /* block index of group g */
  uint32_t cnt[MT_MAX_GROUPS]; /* rows of that block */
};

/* ---- main kernel ----------------------------------------------------------
 * Work unit = (row block of 128 A rows, B tile of 128 rows); units are numbered row-block major and cut into
 * equal contiguous ranges, one per CTA, so that 2 CTAs per SM all carry the same load whatever nA/128 is.
 * A range may cross into the next row block: the CTA then starts a new "segment" (reloads A, flushes and
 * resets the running top-2).  Segment j of row block rb lands in partial slot (rb, j). */
__global__ void __launch_bounds__(MT_THREADS, 2)
    match_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const uint32_t *__restrict__ norm_a,
                    const uint32_t *__restrict__ norm_b, uint32_t na, uint32_t n_tiles, uint32_t units_per_cta, uint32_t total_units,
                    uint32_t max_segs, unsigned long long *__restrict__ partial, int32_t key_scale, const __grid_constant__ MatchGroups G)
{
  extern __shared__ uint8_t smem_raw[];
  /* 1024-byte alignment required by SWIZZLE_128B */
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *s_a = smem;
  uint8_t *s_b = smem + MT_M * 128;
  uint32_t *s_nb = (uint32_t *)(s_b + MT_STAGES * MT_TILE_BYTES);
  uint64_t *s_bar = (uint64_t *)(s_nb + MT_STAGES * MT_N);
  /* barriers: [0..S) full_b, [S..2S) empty_b, 2S full_a, 2S+1 empty_a, then MT_BUFS tmem_full, MT_BUFS tmem_empty */
  uint32_t *s_tmem = (uint32_t *)(s_bar + 2 * MT_STAGES + 2 + 2 * MT_BUFS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  /* programmatic dependent launch: the merge kernel's CTAs may be scheduled while this grid drains; they wait for its
   * completion (griddepcontrol.wait) before they read the partial keys */
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t u0 = blockIdx.x * units_per_cta;
  const uint32_t u1 = min(total_units, u0 + units_per_cta);
  const uint32_t my_tiles = (u1 > u0) ? (u1 - u0) : 0u;

  const uint32_t bar0 = smem_u32(s_bar);
#define BAR_FULL_B(i) (bar0 + 8u * (uint32_t)(i))
#define BAR_EMPTY_B(i) (bar0 + 8u * (uint32_t)(MT_STAGES + (i)))
#define BAR_FULL_A (bar0 + 8u * (uint32_t)(2 * MT_STAGES))
#define BAR_EMPTY_A (bar0 + 8u * (uint32_t)(2 * MT_STAGES + 1))
#define BAR_TMEM_FULL(i) (bar0 + 8u * (uint32_t)(2 * MT_STAGES + 2 + (i)))
#define BAR_TMEM_EMPTY(i) (bar0 + 8u * (uint32_t)(2 * MT_STAGES + 2 + MT_BUFS + (i)))

  if (threadIdx.x == 0)
  {
    for (int i = 0; i < MT_STAGES; i++)
    {
      mbar_init(BAR_FULL_B(i), 1);
      mbar_init(BAR_EMPTY_B(i), 1 + 4); /* MMA commit + the 4 epilogue warps that read the stage's nbk */
    }
    mbar_init(BAR_FULL_A, 1);
    mbar_init(BAR_EMPTY_A, 1);
    for (int i = 0; i < MT_BUFS; i++)
    {
      mbar_init(BAR_TMEM_FULL(i), 1);
      mbar_init(BAR_TMEM_EMPTY(i), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MT_WARP_MMA)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(MT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t rb_first = u0 / n_tiles;

  if (warp == MT_WARP_TMA)
  {
    /* ===== TMA producer ===== */
    if (lane == 0)
    {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
      for (uint32_t t = 0; t < my_tiles; t++)
      {
        const uint32_t u = u0 + t, rb = u / n_tiles, tile = u - rb * n_tiles; /* rb: VIRTUAL row block (group, row block of A) */
        const uint32_t grp = rb / G.row_blocks;
        if (t == 0 || tile == 0)
        {
          const uint32_t seg = rb - rb_first;
          mbar_wait(BAR_EMPTY_A, (seg & 1u) ^ 1u); /* previous segment's MMAs are done with the A tile */
          mbar_expect_tx(BAR_FULL_A, MT_M * 128);
          tma_load_2d(smem_u32(s_a), &map_a, 0, (int)((rb - grp * G.row_blocks) * MT_M), BAR_FULL_A);
        }
        const uint32_t st = t % MT_STAGES, ph = (t / MT_STAGES) & 1u;
        mbar_wait(BAR_EMPTY_B(st), ph ^ 1u);
        const uint32_t b0 = G.blk[grp] * G.stride_rows + tile * MT_N; /* row of the tile in the allocation of all blocks */
        mbar_expect_tx(BAR_FULL_B(st), MT_TILE_BYTES + MT_N * 4);
        tma_load_2d(smem_u32(s_b + st * MT_TILE_BYTES), &map_b, 0, (int)b0, BAR_FULL_B(st));
        bulk_load_1d(smem_u32(s_nb + st * MT_N), norm_b + b0, MT_N * 4, BAR_FULL_B(st));
      }
    }
  }
  else if (warp == MT_WARP_MMA)
  {
    /* ===== MMA issuer (one thread) ===== */
    if (lane == 0)
    {
      const uint64_t adesc = umma_smem_desc(smem_u32(s_a));
      for (uint32_t t = 0; t < my_tiles; t++)
      {
        const uint32_t u = u0 + t, rb = u / n_tiles, tile = u - rb * n_tiles;
        if (t == 0 || tile == 0)
          mbar_wait(BAR_FULL_A, (rb - rb_first) & 1u);
        const uint32_t st = t % MT_STAGES, ph = (t / MT_STAGES) & 1u;
        const uint32_t buf = t % MT_BUFS, bph = (t / MT_BUFS) & 1u;
        mbar_wait(BAR_TMEM_EMPTY(buf), bph ^ 1u);
        mbar_wait(BAR_FULL_B(st), ph);
        tmem_fence_after();
        const uint64_t bdesc = umma_smem_desc(smem_u32(s_b + st * MT_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < 4; k++) /* K = 32 bytes per MMA: advance the start address by 32 B (>>4 = 2) */
          umma_i8(tmem_base + buf * MT_N, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), MT_IDESC, k > 0 ? 1u : 0u);
        umma_commit(BAR_EMPTY_B(st));     /* smem stage consumed by the tensor core */
        umma_commit(BAR_TMEM_FULL(buf)); /* accumulator ready */
        if (tile == n_tiles - 1 || t == my_tiles - 1)
          umma_commit(BAR_EMPTY_A); /* last MMA of the segment: the A tile may be replaced */
      }
    }
  }
  else
  {
    /* ===== epilogue: thread = A row; warpgroup g (warps 4g..4g+3) drains the tiles with (t & 1) == g, i.e.
     * TMEM buffer g; warp w of a group reads TMEM lanes 32*(w%4).. =====
     * Per accumulator: key = nbk[c] - 512*acc = 256*(|b|^2 - 2 a.b) + (pos & 255)  (one IMAD; nbk comes from the
     * norms kernel), then a branch-free top-2 of the 32-bit keys with min/max.  Keys are unique inside a tile
     * (distinct low bytes) and order exactly like (d^2, pos).  After each tile the two survivors are widened to
     * 64-bit (d^2 << 32 | pos) keys and merged into the row's running top-2.  Columns past nb carry the largest
     * possible nbk, so they only win when a segment holds fewer than two real columns (the merge drops them). */
    const int wg = warp >> 2;
    const uint32_t lrow = (uint32_t)(warp & 3) * 32 + lane;
    unsigned long long k1 = ~0ull, k2 = ~0ull;
    int32_t my_na = 0;
    uint32_t cur_rb = 0xffffffffu;
    /* Every (row block, segment, warpgroup) slot of this CTA is published, also when the warpgroup gets no tile of a short
     * segment: "nothing found" first, the real results over it.  All of it happens behind griddepcontrol.wait, as late as
     * possible: when the launch is programmatically dependent on the merge kernel of the PREVIOUS search (back-to-back
     * searches, match_tc_launch), that kernel may still be reading the partial keys while this grid already runs its MMAs. */
    bool published = false;
    auto publish_init = [&]() {
      if (published)
        return;
      published = true;
      asm volatile("griddepcontrol.wait;" ::: "memory");
      if (my_tiles > 0)
        for (uint32_t rb = rb_first; rb <= (u1 - 1) / n_tiles; rb++)
        {
          const uint32_t seg_slot = blockIdx.x - (rb * n_tiles) / units_per_cta;
          const size_t slot = (((size_t)rb * max_segs + seg_slot) * 2 + wg) * MT_M + lrow;
          partial[slot * 2 + 0] = ~0ull;
          partial[slot * 2 + 1] = ~0ull;
        }
    };
    for (uint32_t t = (uint32_t)wg; t < my_tiles; t += 2)
    {
      const uint32_t u = u0 + t, rb = u / n_tiles, tile = u - rb * n_tiles;
      if (rb != cur_rb)
      {
        if (cur_rb != 0xffffffffu)
        {
          /* flush the finished segment */
          publish_init();
          const uint32_t seg_slot = blockIdx.x - (cur_rb * n_tiles) / units_per_cta;
          const size_t slot = (((size_t)cur_rb * max_segs + seg_slot) * 2 + wg) * MT_M + lrow;
          partial[slot * 2 + 0] = k1;
          partial[slot * 2 + 1] = k2;
          k1 = k2 = ~0ull;
        }
        cur_rb = rb;
        const uint32_t row = (rb % G.row_blocks) * MT_M + lrow;
        my_na = (row < na) ? (int32_t)norm_a[row] : 0;
      }
      const uint32_t st = t % MT_STAGES, ph = (t / MT_STAGES) & 1u;
      const uint32_t buf = t % MT_BUFS, bph = (t / MT_BUFS) & 1u;
      const uint32_t b0 = tile * MT_N;
      mbar_wait(BAR_FULL_B(st), ph); /* nbk of this tile landed (same barrier as the B tile) */
      mbar_wait(BAR_TMEM_FULL(buf), bph);
      tmem_fence_after();
      const uint32_t nbs = smem_u32(s_nb + st * MT_N);
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + buf * MT_N;
      int32_t c1 = 0x7fffffff, c2 = 0x7fffffff;
      /* 32 columns at a time; the load of the next 32 is in flight while the current 32 are reduced.
       * An exact running top-2 costs 2.5 min/max per accumulator and made the ALU pipe the limit of the whole kernel.
       * Instead: only the MINIMUM of every group of MT_GROUP = 16 consecutive columns (one 3-input min per two
       * accumulators) and the top-2 of the group minima (5 operations per 32 columns).  The smallest group minimum is the
       * row's nearest neighbour; the second nearest is either the second smallest group minimum or sits in the winner's
       * own group, which the merge kernel rescans (16 columns per row) -- see match_merge_kernel. */
      auto group_min = [&](const int32_t(&a)[32], int base, int c0) -> int32_t {
        int32_t m = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < MT_GROUP / 4; q++)
        {
          const int4 n4 = lds_v4(nbs + (uint32_t)(c0 + base + q * 4) * 4u);
          /* key_scale = -512 arrives as a kernel argument so that this stays one IMAD on the FMA pipe; a literal
           * power of two is strength-reduced to shift+add on the ALU pipe, which the min ops already load */
          const int32_t e0 = a[base + q * 4 + 0] * key_scale + n4.x, e1 = a[base + q * 4 + 1] * key_scale + n4.y;
          const int32_t e2 = a[base + q * 4 + 2] * key_scale + n4.z, e3 = a[base + q * 4 + 3] * key_scale + n4.w;
          m = __vimin3_s32(m, e0, e1);
          m = __vimin3_s32(m, e2, e3);
        }
        return m;
      };
      auto reduce32 = [&](const int32_t(&a)[32], int c0) {
        const int32_t g0 = group_min(a, 0, c0), g1 = group_min(a, MT_GROUP, c0);
        const int32_t lo = min(g0, g1), hi = max(g0, g1), tt = max(c1, lo);
        c1 = min(c1, lo);
        c2 = __vimin3_s32(c2, tt, hi);
      };
      {
        int32_t acc0[32], acc1[32];
        tmem_ld32(taddr, acc0);
        tmem_ld_wait_for(acc0);
#pragma unroll 1
        for (int c0 = 0; c0 < MT_N; c0 += 64)
        {
          tmem_ld32(taddr + c0 + 32, acc1);
          reduce32(acc0, c0);
          tmem_ld_wait_for(acc1);
          if (c0 + 64 < MT_N)
            tmem_ld32(taddr + c0 + 64, acc0);
          reduce32(acc1, c0 + 32);
          if (c0 + 64 < MT_N)
            tmem_ld_wait_for(acc0);
        }
      }
      tmem_fence_before();
      __syncwarp();
      if (lane == 0)
      {
        mbar_arrive(BAR_TMEM_EMPTY(buf));
        mbar_arrive(BAR_EMPTY_B(st));
      }
      /* widen the tile's two survivors and merge */
#pragma unroll
      for (int q = 0; q < 2; q++)
      {
        const int32_t ck = q ? c2 : c1;
        const uint32_t d2 = (uint32_t)((ck >> 8) + my_na);
        const uint32_t pos = (b0 & ~255u) | ((uint32_t)ck & 255u);
        const unsigned long long key = ((unsigned long long)d2 << 32) | pos;
        if (key < k1)
        {
          k2 = k1;
          k1 = key;
        }
        else if (key < k2)
          k2 = key;
      }
    }
    publish_init();
    if (cur_rb != 0xffffffffu)
    {
      const uint32_t seg_slot = blockIdx.x - (cur_rb * n_tiles) / units_per_cta;
      const size_t slot = (((size_t)cur_rb * max_segs + seg_slot) * 2 + wg) * MT_M + lrow;
      partial[slot * 2 + 0] = k1;
      partial[slot * 2 + 1] = k2;
    }
  }

  tmem_fence_before();
  __syncthreads();
  if (warp == MT_WARP_MMA)
  {
    tmem_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(MT_TMEM_COLS) : "memory");
  }
}

/* One warp per A row.  Fold the segments of the row block: the partial keys are minima of disjoint 16-column groups, so
 * the smallest one is the nearest neighbour K1 (exact) and the second smallest, K2', is the best column outside K1's
 * group.  The second nearest neighbour is min(K2', best column of K1's group other than K1): the 16 columns of that
 * group are rescanned here with exact integer arithmetic (lane = (column, half of the 128 bytes)).  Then undo the
 * position permutation and take the square roots (Get2NearestNeighbors.comp:98-102). */
/* smallest 64-bit key of the warp with two 32-bit REDUX steps (high words, then low words among the lanes that tie) */
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long k)
{
  const uint32_t hi = (uint32_t)(k >> 32), lo = (uint32_t)k;
  const uint32_t mhi = __reduce_min_sync(0xffffffffu, hi);
  const uint32_t mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
  return ((unsigned long long)mhi << 32) | mlo;
}

#define MG_WARPS 8
__global__ void __launch_bounds__(32 * MG_WARPS) match_merge_kernel(const unsigned long long *__restrict__ partial, uint32_t n_tiles,
                                                                    uint32_t units_per_cta, uint32_t max_segs, uint32_t na,
                                                                    const uint8_t *__restrict__ da, const uint8_t *__restrict__ db_all,
                                                                    const uint32_t *__restrict__ norm_a, const uint32_t *__restrict__ norm_b_all,
                                                                    vksift_Match_2NN *__restrict__ out_all, const __grid_constant__ MatchGroups G)
{
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  /* the MMA kernel of the next search may start now (it touches the partial keys only after this grid has completed) */
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t row = blockIdx.x * MG_WARPS + (uint32_t)warp;
  if (row >= na)
    return;
  /* blockIdx.y = group: the block of B this list is for */
  const uint32_t grp = blockIdx.y, blk = G.blk[grp], nb = G.cnt[grp];
  const uint8_t *__restrict__ db = db_all + (size_t)blk * G.stride_rows * 128;
  const uint32_t *__restrict__ norm_b = norm_b_all + (size_t)blk * G.stride_rows;
  vksift_Match_2NN *__restrict__ out = out_all + (size_t)blk * G.out_stride;
  const int half = lane & 1;
  /* this lane's half of the A row, for the rescan below: independent of the fold, so the loads go out first */
  uint4 va[4];
  {
    const uint4 *pa = reinterpret_cast<const uint4 *>(da + (size_t)row * 128 + half * 64);
#pragma unroll
    for (int i = 0; i < 4; i++)
      va[i] = __ldg(pa + i);
  }
  const uint32_t my_na = __ldg(norm_a + row);
  asm volatile("griddepcontrol.wait;" ::: "memory"); /* everything above only reads inputs of the match call */
  const uint32_t rb = grp * G.row_blocks + row / MT_M, lrow = row % MT_M; /* virtual row block, like the MMA kernel */
  const uint32_t first_cta = (rb * n_tiles) / units_per_cta, last_cta = ((rb + 1) * n_tiles - 1) / units_per_cta;
  const uint32_t n_keys = (last_cta - first_cta + 1) * 2 * 2; /* segments x warpgroups x (k1, k2) */
  /* every lane folds its share of the partial keys (one key per lane unless a row block has more than 8 segments) */
  unsigned long long k1 = ~0ull, k2 = ~0ull;
  for (uint32_t i = (uint32_t)lane; i < n_keys; i += 32)
  {
    const uint32_t s = i >> 1, q = i & 1u;
    const size_t slot = (((size_t)rb * max_segs) * 2 + s) * MT_M + lrow;
    const unsigned long long key = partial[slot * 2 + q];
    if (key < k1)
    {
      k2 = k1;
      k1 = key;
    }
    else if (key < k2)
      k2 = key;
  }
  /* top-2 of the warp: keys of real columns are unique, ~0 marks "none" */
  const unsigned long long best = warp_min_u64(k1);
  const unsigned long long mine = (k1 == best) ? k2 : k1; /* the winner's lane offers its runner-up */
  unsigned long long second = warp_min_u64(mine);
  /* rescan the winner's group: positions [g*16, g*16+16) <-> the same set of B rows (pos swaps rows 0 and 1 only) */
  if (best != ~0ull)
  {
    const uint32_t pos1 = (uint32_t)best;
    const uint32_t b = (pos1 & ~(uint32_t)(MT_GROUP - 1)) + (uint32_t)(lane >> 1);
    uint32_t dot = 0, nbk = 0;
    if (b < nb)
    {
      const uint4 *pb = reinterpret_cast<const uint4 *>(db + (size_t)b * 128 + half * 64);
      nbk = __ldg(norm_b + b);
#pragma unroll
      for (int i = 0; i < 4; i++)
      {
        const uint4 vb = __ldg(pb + i);
        dot = __dp4a(va[i].x, vb.x, dot);
        dot = __dp4a(va[i].y, vb.y, dot);
        dot = __dp4a(va[i].z, vb.z, dot);
        dot = __dp4a(va[i].w, vb.w, dot);
      }
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    unsigned long long cand = ~0ull;
    const uint32_t pos = mt_pos(b);
    if (b < nb && pos != pos1)
      cand = ((unsigned long long)(my_na + (nbk >> 8) - 2u * dot) << 32) | pos;
    cand = warp_min_u64(cand);
    if (cand < second)
      second = cand;
  }
  /* The shader compares the float distances sqrt(float(d^2)) with a strict '<' (Get2NearestNeighbors.comp:82-96).  Below 2^22 the
   * square root is injective on the integers, so the integer order above IS the shader's order.  From 2^22 on (only reachable
   * with descriptors that are not SIFT descriptors: |a - b| >= 2048) distinct d^2 can round to the same float and the earlier
   * scan position wins there: such a row is rescanned completely under the key (bits of sqrt(float(d^2)), pos). */
  if ((uint32_t)(second >> 32) >= (1u << 22) && second != ~0ull)
  {
    uint4 fa[8];
    {
      const uint4 *pa = reinterpret_cast<const uint4 *>(da + (size_t)row * 128);
#pragma unroll
      for (int i = 0; i < 8; i++)
        fa[i] = __ldg(pa + i);
    }
    unsigned long long s1 = ~0ull, s2 = ~0ull;
    for (uint32_t b = (uint32_t)lane; b < nb; b += 32)
    {
      const uint4 *pb = reinterpret_cast<const uint4 *>(db + (size_t)b * 128);
      uint32_t dot = 0, nbv = 0;
#pragma unroll
      for (int i = 0; i < 8; i++)
      {
        const uint4 vb = __ldg(pb + i);
        dot = __dp4a(fa[i].x, vb.x, dot);
        dot = __dp4a(fa[i].y, vb.y, dot);
        dot = __dp4a(fa[i].z, vb.z, dot);
        dot = __dp4a(fa[i].w, vb.w, dot);
        nbv = __dp4a(vb.x, vb.x, nbv);
        nbv = __dp4a(vb.y, vb.y, nbv);
        nbv = __dp4a(vb.z, vb.z, nbv);
        nbv = __dp4a(vb.w, vb.w, nbv);
      }
      const float d = vks_sqrt((float)(my_na + nbv - 2u * dot));
      const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | mt_pos(b);
      if (key < s1)
      {
        s2 = s1;
        s1 = key;
      }
      else if (key < s2)
        s2 = key;
    }
    const unsigned long long b1 = warp_min_u64(s1);
    const unsigned long long b2 = warp_min_u64(s1 == b1 ? s2 : s1);
    if (lane == 0)
    {
      vksift_Match_2NN m;
      m.idx_a = row;
      m.idx_b1 = mt_pos((uint32_t)b1);
      m.idx_b2 = mt_pos((uint32_t)b2);
      m.dist_a_b1 = __uint_as_float((uint32_t)(b1 >> 32));
      m.dist_a_b2 = __uint_as_float((uint32_t)(b2 >> 32));
      out[row] = m;
    }
    return;
  }
  if (lane == 0)
  {
    vksift_Match_2NN m;
    m.idx_a = row;
    m.idx_b1 = mt_pos((uint32_t)best);
    m.idx_b2 = mt_pos((uint32_t)second);
    m.dist_a_b1 = vks_sqrt((float)(uint32_t)(best >> 32));
    m.dist_a_b2 = vks_sqrt((float)(uint32_t)(second >> 32));
    out[row] = m;
  }
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

static bool mt_make_map(MatchTc *tc, CUtensorMap *map, const uint8_t *base, uint32_t rows, uint32_t box_rows)
{
  const cuuint64_t gdim[2] = {128, rows};
  const cuuint64_t gstride[1] = {128};
  const cuuint32_t box[2] = {128, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = tc->encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

/* groups: the blocks of B to search (MatchGroups: n_groups, blk[], cnt[], stride_rows, out_stride filled by the caller).  db_all /
 * norm_b_all / out_all address block 0.  With more than one group the rows of every block between its count and the largest
 * count rounded up to a tile MUST be zero (the tensor map spans all blocks, so TMA does not zero them): the caller's contract.
 * overlap_previous: nothing this search reads (descriptors, norms) was written by the work enqueued just before it on `st`,
 * so the MMA kernel is launched programmatically dependent: behind the merge kernel of a previous search it starts while that
 * one still runs, and waits for it only before it publishes its partial keys. */
static cudaError_t match_tc_launch_groups(void *p, const uint8_t *da, uint32_t na, const uint32_t *norm_a, const uint8_t *db_all,
                                          const uint32_t *norm_b_all, vksift_Match_2NN *out_all, MatchGroups G, cudaStream_t st, bool overlap_previous,
                                          uint64_t *launch_count)
{
  MatchTc *tc = (MatchTc *)p;
  if (G.n_groups == 0 || G.n_groups > MT_MAX_GROUPS)
    return cudaErrorInvalidValue;
  const uint32_t row_blocks = (na + MT_M - 1) / MT_M;
  uint32_t max_cnt = 0, max_blk = 0;
  for (uint32_t g = 0; g < G.n_groups; g++)
  {
    max_cnt = G.cnt[g] > max_cnt ? G.cnt[g] : max_cnt;
    max_blk = G.blk[g] > max_blk ? G.blk[g] : max_blk;
  }
  const uint32_t n_tiles = (max_cnt + MT_N - 1) / MT_N;
  G.row_blocks = row_blocks;
  const uint32_t v_blocks = G.n_groups * row_blocks; /* virtual row blocks */
  /* two CTAs are resident per SM: one full wave of equally loaded CTAs */
  const uint32_t total_units = v_blocks * n_tiles;
  uint32_t n_cta = 2u * (uint32_t)tc->sm_count;
  if (n_cta > total_units)
    n_cta = total_units;
  const uint32_t units_per_cta = (total_units + n_cta - 1) / n_cta;
  n_cta = (total_units + units_per_cta - 1) / units_per_cta;
  const uint32_t max_segs = (n_tiles + units_per_cta - 1) / units_per_cta + 1;
  const size_t need = (size_t)v_blocks * max_segs * 2 * MT_M * 2;
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
  /* one group: the map ends at the block's last row and TMA zero-fills the rest of its last tile; several groups: the map spans
   * every block up to the last tile any of them needs */
  const uint32_t map_b_rows = G.n_groups == 1 ? G.blk[0] * G.stride_rows + G.cnt[0] : max_blk * G.stride_rows + n_tiles * MT_N;
  CUtensorMap map_a, map_b;
  if (!mt_make_map(tc, &map_a, da, na, MT_M) || !mt_make_map(tc, &map_b, db_all, map_b_rows, MT_N))
    return cudaErrorInvalidValue;
  cudaError_t e;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_cta);
    cfg.blockDim = dim3(MT_THREADS);
    cfg.dynamicSmemBytes = MT_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap_previous ? 1 : 0;
    unsigned long long *partial = tc->partial;
    const int32_t key_scale = -512;
    e = cudaLaunchKernelEx(&cfg, match_tc_kernel, map_a, map_b, norm_a, norm_b_all, na, n_tiles, units_per_cta, total_units, max_segs, partial,
                           key_scale, G);
  }
  if (e != cudaSuccess)
    return e;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((na + MG_WARPS - 1) / MG_WARPS, G.n_groups);
    cfg.blockDim = dim3(32 * MG_WARPS);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const unsigned long long *partial_c = tc->partial;
    e = cudaLaunchKernelEx(&cfg, match_merge_kernel, partial_c, n_tiles, units_per_cta, max_segs, na, da, db_all, norm_a, norm_b_all, out_all, G);
  }
  *launch_count += 2;
  return e;
}

static cudaError_t match_tc_launch(void *p, const uint8_t *da, uint32_t na, const uint32_t *norm_a, const uint8_t *db, uint32_t nb,
                                   const uint32_t *norm_b, vksift_Match_2NN *out, cudaStream_t st, bool overlap_previous, uint64_t *launch_count)
{
  MatchGroups G;
  memset(&G, 0, sizeof(G));
  G.n_groups = 1;
  G.cnt[0] = nb;
  return match_tc_launch_groups(p, da, na, norm_a, db, norm_b, out, G, st, overlap_previous, launch_count);
}

} // namespace vks
