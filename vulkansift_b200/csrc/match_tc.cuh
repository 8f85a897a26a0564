/*
 * match_tc.cuh -- 2-NN matcher on the 5th-generation tensor cores (tcgen05, sm_100a).
 *
 * The distance is produced BY THE MMA, not by the epilogue:
 *
 *   d(a,b)^2 / 2 = (|a|^2 + |b|^2) / 2 - a.b
 *
 * Operands are the descriptors converted to binary16 (0..255 are exact), K = 128, plus 16 extension columns that carry the
 * two norms: A row = [a_0 .. a_127 | n2*128, n1, n0, 256, 128, 0.5, 2048, 0..], B row = [b_0 .. b_127 | 256, 128, 0.5,
 * m2*128, m1, m0, pad, 0..] with |a|^2 = n2*65536 + n1*256 + n0 (same for |b|^2 and m).  One kind::f16 MMA over the extension
 * columns starts the accumulator at (|a|^2 + |b|^2)/2, eight more with the A operand negated (instruction descriptor bit 13)
 * subtract a.b.  Every product and every partial sum is an integer multiple of 0.5 below 2^23, so the fp32 accumulation of the
 * tensor core is exact whatever its internal order: the accumulator holds d^2/2 exactly, and non-negative floats order like
 * their bit patterns.  The epilogue therefore needs no per-column operand and no multiply:
 *
 *   TMA (cp.async.bulk.tensor.2d) -> smem operand tile = 2 x (128 rows x 128 B, 128B swizzle) + extension (128 rows x 32 B, 32B swizzle);
 *     two A tiles resident, B tiles in a 4-stage ring
 *   tcgen05.mma.cta_group::1.kind::f16  M=128 N=128 K=16 x 9 -> TMEM accumulators (4 x 128 columns, one per epilogue warpgroup)
 *   4 x 4 epilogue warps: tcgen05.ld 32x32b.x32 -> minimum of every group of 8 columns with 3-input integer minima on the
 *     float bit patterns (0.5 operations per accumulator), one conversion per group, branch-free top-2 of the groups
 *
 * Binary16 operands are twice the bytes of the descriptors (three times with the extension block), and with one A row block
 * per CTA the B tiles streamed from L2 (300 MB for 10k x 10k) bound the kernel, not the tensor pipe.  A CTA therefore keeps
 * TWO A row blocks (256 rows) resident and runs every B tile against both: half the operand traffic per MMA.
 * One CTA per SM (512 TMEM columns, 193 KB of shared memory) owns a contiguous range of (row block pair, B tile) units.  The
 * partial results -- the two best GROUPS per row -- go to HBM; the merge kernel folds them and rescans the 16 columns of the
 * two best groups with exact integer arithmetic (the nearest neighbour lies in the best group, the second nearest in the
 * best or the second best one), applies the shader's tie rule and takes the square root.
 *
 * Reference semantics: shaders/Get2NearestNeighbors.comp:43-104.
 */
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "vksift_internal.h"

namespace vks
{

#define MT_M 128        /* A rows per CTA = TMEM lanes */
#define MT_N 128        /* B rows per MMA tile = TMEM columns per accumulator buffer */
#define MT_KBLK 2       /* operand blocks of 64 halves (128 B rows, SWIZZLE_128B): bytes 0-63, bytes 64-127 */
#define MT_EXT_BYTES (128 * 32) /* extension operand: 16 halves (32 B rows, SWIZZLE_32B) of the 128 rows */
#define MT_STAGES 4     /* B smem ring */
#define MT_BUFS 4       /* TMEM accumulator buffers, one per epilogue warpgroup: (tile parity, row block of the pair) */
#define MT_BLK_BYTES (128 * 128)
#define MT_TILE_BYTES (MT_KBLK * MT_BLK_BYTES + MT_EXT_BYTES)
#define MT_WARP_TMA 16
#define MT_WARP_MMA 17
#define MT_THREADS (18 * 32) /* warps 0-15: four epilogue warpgroups, warp 16 TMA, warp 17 MMA + TMEM alloc */
#define MT_TMEM_COLS 512
#define MT_GROUP 8      /* columns per group of the epilogue */
#define MT_SMEM_BYTES (1024 + 2 * MT_TILE_BYTES + MT_STAGES * MT_TILE_BYTES + 512)
/* operand memory of a descriptor set of n_pad rows: [2][n_pad][64] halves (main blocks), then [2 roles][n_pad][16] halves (extension:
 * role 0 = the rows used as A, role 1 = the rows used as B) */
#define MT_OP_ROW_BYTES (MT_KBLK * 128 + 2 * 32)
#define MT_PAD_A 2048.f
#define MT_PAD_B 4096.f /* 2048 * 4096 = 2^23 > any d^2/2: rows that do not exist lose every comparison */

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
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n\t"
               ".reg .pred p;\n\t"
               "setp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
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
/* the same for the extension operand: rows of 32 bytes, SWIZZLE_32B (layout type 6), 8-row groups 256 B apart */
__device__ __forceinline__ uint64_t umma_smem_desc_32(uint32_t saddr)
{
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}
/* kind::f16 instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = F16 (0),
 * K-major both, N at bit 17 (>> 3), M at bit 24 (>> 4); bit 13 negates A */
#define MT_IDESC ((1u << 4) | ((uint32_t)(MT_N >> 3) << 17) | ((uint32_t)(MT_M >> 4) << 24))
#define MT_IDESC_NEG_A (MT_IDESC | (1u << 13))

__device__ __forceinline__ uint32_t mt_pos(uint32_t b) { return b < 2u ? (b ^ 1u) : b; }

/* ---- operand preparation ----------------------------------------------------
 * One warp per descriptor: |x|^2 (plain, for the exact rescan and the SIMT kernel) and the binary16 operand rows of the
 * three blocks.  op = [3][n_pad][64] halves, n_pad = n rounded up to 128; rows [n, n_pad) are "absent" B rows. */
__global__ void match_prepare_kernel(const uint8_t *__restrict__ desc, uint32_t n, uint32_t n_pad, uint32_t *__restrict__ norm, __half *__restrict__ op)
{
  const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_pad)
    return;
  const bool real = row < n;
  const uint32_t v = real ? __ldg((const uint32_t *)(desc + (size_t)row * 128) + lane) : 0u;
  uint32_t s = __dp4a(v, v, 0u);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, d);
  /* bytes 4*lane .. 4*lane+3 -> four halves of block (lane >> 4) at column 4*(lane & 15) */
  {
    const __half2 lo = __floats2half2_rn((float)(v & 255u), (float)((v >> 8) & 255u));
    const __half2 hi = __floats2half2_rn((float)((v >> 16) & 255u), (float)(v >> 24));
    uint2 w;
    w.x = *reinterpret_cast<const uint32_t *>(&lo);
    w.y = *reinterpret_cast<const uint32_t *>(&hi);
    *reinterpret_cast<uint2 *>(op + ((size_t)(lane >> 4) * n_pad + row) * 64 + 4 * (lane & 15)) = w;
  }
  /* extension rows: 16 halves per role, two per lane (lanes 0-7 the A role, lanes 8-15 the B role) */
  if (lane < 16)
  {
    const float n2 = (float)((s >> 16) * 128u), n1 = (float)((s >> 8) & 255u), n0 = (float)(s & 255u);
    const int role = lane >> 3;
    float e[2];
#pragma unroll
    for (int q = 0; q < 2; q++)
    {
      const int c = 2 * (lane & 7) + q; /* column of the role's row */
      float x = 0.f;
      if (role == 0)
        x = !real ? 0.f : (c == 0 ? n2 : c == 1 ? n1 : c == 2 ? n0 : c == 3 ? 256.f : c == 4 ? 128.f : c == 5 ? 0.5f : c == 6 ? MT_PAD_A : 0.f);
      else
        x = c == 0 ? 256.f : c == 1 ? 128.f : c == 2 ? 0.5f : c == 3 ? n2 : c == 4 ? n1 : c == 5 ? n0 : c == 6 ? (real ? 0.f : MT_PAD_B) : 0.f;
      e[q] = x;
    }
    const __half2 h = __floats2half2_rn(e[0], e[1]);
    __half *ext = op + (size_t)MT_KBLK * n_pad * 64; /* behind the main blocks */
    *reinterpret_cast<__half2 *>(ext + ((size_t)role * n_pad + row) * 16 + 2 * (lane & 7)) = h;
  }
  if (lane == 0 && real)
    norm[row] = s;
}

/* ---- main kernel ----------------------------------------------------------
 * Work unit = (pair of row blocks = 256 A rows, B tile of 128 rows); units are numbered pair major and cut into
 * equal contiguous ranges, one per CTA.  A range may cross into the next pair: the CTA then starts a new "segment"
 * (reloads A, flushes and resets the running top-2).  Segment j of pair rp lands in the partial slots (2 rp + h, j, tile parity). */
__global__ void __launch_bounds__(MT_THREADS, 1)
    match_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_ea,
                    const __grid_constant__ CUtensorMap map_eb, uint32_t na_pad, uint32_t nb_pad, uint32_t n_tiles,
                    uint32_t units_per_cta, uint32_t total_units, uint32_t max_segs, unsigned long long *__restrict__ partial)
{
  extern __shared__ uint8_t smem_raw[];
  /* 1024-byte alignment required by SWIZZLE_128B */
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *s_a = smem; /* two row blocks */
  uint8_t *s_b = smem + 2 * MT_TILE_BYTES;
  uint64_t *s_bar = (uint64_t *)(s_b + MT_STAGES * MT_TILE_BYTES);
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
      mbar_init(BAR_EMPTY_B(i), 1); /* MMA commit */
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
    /* ===== TMA producer: an operand tile = the 128 rows of each of the three blocks ===== */
    if (lane == 0)
    {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ea) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_eb) : "memory");
      for (uint32_t t = 0; t < my_tiles; t++)
      {
        const uint32_t u = u0 + t, rb = u / n_tiles, tile = u - rb * n_tiles;
        if (t == 0 || tile == 0)
        {
          const uint32_t seg = rb - rb_first;
          mbar_wait(BAR_EMPTY_A, (seg & 1u) ^ 1u); /* previous segment's MMAs are done with the A tile */
          mbar_expect_tx(BAR_FULL_A, 2 * MT_TILE_BYTES);
#pragma unroll
          for (int h = 0; h < 2; h++)
          {
#pragma unroll
            for (int k = 0; k < MT_KBLK; k++)
              tma_load_2d(smem_u32(s_a + h * MT_TILE_BYTES + k * MT_BLK_BYTES), &map_a, 0, (int)(k * na_pad + (2 * rb + h) * MT_M), BAR_FULL_A);
            tma_load_2d(smem_u32(s_a + h * MT_TILE_BYTES + MT_KBLK * MT_BLK_BYTES), &map_ea, 0, (int)((2 * rb + h) * MT_M), BAR_FULL_A); /* role 0 */
          }
        }
        const uint32_t st = t % MT_STAGES, ph = (t / MT_STAGES) & 1u;
        mbar_wait(BAR_EMPTY_B(st), ph ^ 1u);
        mbar_expect_tx(BAR_FULL_B(st), MT_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < MT_KBLK; k++)
          tma_load_2d(smem_u32(s_b + st * MT_TILE_BYTES + k * MT_BLK_BYTES), &map_b, 0, (int)(k * nb_pad + tile * MT_N), BAR_FULL_B(st));
        tma_load_2d(smem_u32(s_b + st * MT_TILE_BYTES + MT_KBLK * MT_BLK_BYTES), &map_eb, 0, (int)(nb_pad + tile * MT_N), BAR_FULL_B(st)); /* role 1 */
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
        const uint32_t bph = (t >> 1) & 1u; /* every accumulator buffer is used by every second tile */
        mbar_wait(BAR_FULL_B(st), ph);
        const uint64_t bdesc = umma_smem_desc(smem_u32(s_b + st * MT_TILE_BYTES));
        const uint64_t bdesc_e = umma_smem_desc_32(smem_u32(s_b + st * MT_TILE_BYTES + MT_KBLK * MT_BLK_BYTES));
#pragma unroll
        for (int h = 0; h < 2; h++)
        {
          const uint32_t buf = 2u * (t & 1u) + (uint32_t)h;
          mbar_wait(BAR_TMEM_EMPTY(buf), bph ^ 1u);
          tmem_fence_after();
          const uint64_t ad = adesc + (uint64_t)(h * (MT_TILE_BYTES >> 4));
          const uint32_t d = tmem_base + buf * MT_N;
          /* extension operands first: D = (|a|^2 + |b|^2) / 2 */
          umma_f16(d, umma_smem_desc_32(smem_u32(s_a + h * MT_TILE_BYTES + MT_KBLK * MT_BLK_BYTES)), bdesc_e, MT_IDESC, 0u);
          /* D -= a.b: K = 16 halves per MMA = 32 bytes (>>4 = 2), four per block */
#pragma unroll
          for (int kb = 0; kb < 2; kb++)
#pragma unroll
            for (int k = 0; k < 4; k++)
              umma_f16(d, ad + (uint64_t)(kb * (MT_BLK_BYTES >> 4) + 2 * k), bdesc + (uint64_t)(kb * (MT_BLK_BYTES >> 4) + 2 * k), MT_IDESC_NEG_A, 1u);
          umma_commit(BAR_TMEM_FULL(buf)); /* accumulator ready */
        }
        umma_commit(BAR_EMPTY_B(st)); /* smem stage consumed by the tensor core */
        if (tile == n_tiles - 1 || t == my_tiles - 1)
          umma_commit(BAR_EMPTY_A); /* last MMA of the segment: the A tile may be replaced */
      }
    }
  }
  else
  {
    /* ===== epilogue: thread = A row; warpgroup g (warps 4g..4g+3) drains TMEM buffer g = the tiles of parity g >> 1 against
     * row block g & 1 of the pair; warp w of a group reads TMEM lanes 32*(w%4).. =====
     * The accumulator is d^2/2 >= 0 as a float, whose bit pattern orders like the value.  Minimum of every group of 8
     * columns with 3-input integer minima (4 operations per 8 accumulators), then one conversion per group:
     * key = d^2 * 16 + group (d^2 < 2^23), and a branch-free top-2 of the 16 group keys of a tile.  After each tile the two
     * survivors are widened to 64-bit (d^2 << 32 | global group) keys and merged into the row's running top-2. */
    const int wg = warp >> 2;
    const uint32_t par = (uint32_t)wg >> 1, hb = (uint32_t)wg & 1u;
    const uint32_t lrow = (uint32_t)(warp & 3) * 32 + lane;
    unsigned long long k1 = ~0ull, k2 = ~0ull;
    uint32_t cur_rb = 0xffffffffu;
    /* every (row block, segment, warpgroup) slot of this CTA is published, also when the warpgroup gets no
     * tile of a short segment: start from "nothing found" and overwrite with the real result below */
    if (my_tiles > 0)
      for (uint32_t rb = rb_first; rb <= (u1 - 1) / n_tiles; rb++)
      {
        const uint32_t seg_slot = blockIdx.x - (rb * n_tiles) / units_per_cta;
        const size_t slot = (((size_t)(2 * rb + hb) * max_segs + seg_slot) * 2 + par) * MT_M + lrow;
        partial[slot * 2 + 0] = ~0ull;
        partial[slot * 2 + 1] = ~0ull;
      }
    for (uint32_t t = par; t < my_tiles; t += 2)
    {
      const uint32_t u = u0 + t, rb = u / n_tiles, tile = u - rb * n_tiles;
      if (rb != cur_rb)
      {
        if (cur_rb != 0xffffffffu)
        {
          /* flush the finished segment */
          const uint32_t seg_slot = blockIdx.x - (cur_rb * n_tiles) / units_per_cta;
          const size_t slot = (((size_t)(2 * cur_rb + hb) * max_segs + seg_slot) * 2 + par) * MT_M + lrow;
          partial[slot * 2 + 0] = k1;
          partial[slot * 2 + 1] = k2;
          k1 = k2 = ~0ull;
        }
        cur_rb = rb;
      }
      const uint32_t buf = (uint32_t)wg, bph = (t >> 1) & 1u;
      mbar_wait(BAR_TMEM_FULL(buf), bph);
      tmem_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + buf * MT_N;
      int32_t c1 = 0x7fffffff, c2 = 0x7fffffff;
      auto reduce32 = [&](const int32_t(&a)[32], int g0) {
#pragma unroll
        for (int q = 0; q < 32 / MT_GROUP; q++)
        {
          int32_t m = __vimin3_s32(a[8 * q + 0], a[8 * q + 1], a[8 * q + 2]);
          m = __vimin3_s32(m, a[8 * q + 3], a[8 * q + 4]);
          m = __vimin3_s32(m, a[8 * q + 5], a[8 * q + 6]);
          m = min(m, a[8 * q + 7]);
          const float f = __int_as_float(m);
          const int32_t key = __float2int_rn(f + f) * 16 + (g0 + q); /* d^2 * 16 + group of the tile */
          const int32_t tt = max(c1, key);
          c1 = min(c1, key);
          c2 = min(c2, tt);
        }
      };
      {
        int32_t acc0[32], acc1[32];
        tmem_ld32(taddr, acc0);
        tmem_ld_wait_for(acc0);
#pragma unroll 1
        for (int c0 = 0; c0 < MT_N; c0 += 64)
        {
          tmem_ld32(taddr + c0 + 32, acc1);
          reduce32(acc0, c0 / MT_GROUP);
          tmem_ld_wait_for(acc1);
          if (c0 + 64 < MT_N)
            tmem_ld32(taddr + c0 + 64, acc0);
          reduce32(acc1, (c0 + 32) / MT_GROUP);
          if (c0 + 64 < MT_N)
            tmem_ld_wait_for(acc0);
        }
      }
      tmem_fence_before();
      __syncwarp();
      if (lane == 0)
        mbar_arrive(BAR_TMEM_EMPTY(buf));
      /* widen the tile's two survivors and merge */
#pragma unroll
      for (int q = 0; q < 2; q++)
      {
        const int32_t ck = q ? c2 : c1;
        const uint32_t d2 = (uint32_t)ck >> 4;
        const uint32_t gg = tile * (MT_N / MT_GROUP) + ((uint32_t)ck & 15u);
        const unsigned long long key = ((unsigned long long)d2 << 32) | gg;
        if (key < k1)
        {
          k2 = k1;
          k1 = key;
        }
        else if (key < k2)
          k2 = key;
      }
    }
    if (cur_rb != 0xffffffffu)
    {
      const uint32_t seg_slot = blockIdx.x - (cur_rb * n_tiles) / units_per_cta;
      const size_t slot = (((size_t)(2 * cur_rb + hb) * max_segs + seg_slot) * 2 + par) * MT_M + lrow;
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

/* smallest 64-bit key of the warp with two 32-bit REDUX steps (high words, then low words among the lanes that tie) */
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long k)
{
  const uint32_t hi = (uint32_t)(k >> 32), lo = (uint32_t)k;
  const uint32_t mhi = __reduce_min_sync(0xffffffffu, hi);
  const uint32_t mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
  return ((unsigned long long)mhi << 32) | mlo;
}

/* One warp per A row.  The partial keys are (minimum d^2, group) of the two best groups of 8 columns of every segment and
 * warpgroup; groups are disjoint, so the two smallest keys overall name the two best groups G1, G2 of the row (ties go to the
 * lower group, i.e. the lower positions).  The nearest neighbour lies in G1, the second nearest in G1 or G2: the 16 columns
 * are rescanned with exact integer arithmetic (lane = (column, half of the 128 bytes)) under the shader's order (d^2, pos),
 * pos(0) = 1, pos(1) = 0 (Get2NearestNeighbors.comp:69-96); then the square roots (:98-102). */
#define MG_WARPS 8
__global__ void __launch_bounds__(32 * MG_WARPS) match_merge_kernel(const unsigned long long *__restrict__ partial, uint32_t n_tiles,
                                                                    uint32_t units_per_cta, uint32_t max_segs, uint32_t na, uint32_t nb,
                                                                    const uint8_t *__restrict__ da, const uint8_t *__restrict__ db,
                                                                    const uint32_t *__restrict__ norm_a, const uint32_t *__restrict__ norm_b,
                                                                    vksift_Match_2NN *__restrict__ out)
{
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row = blockIdx.x * MG_WARPS + (uint32_t)warp;
  if (row >= na)
    return;
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
  const uint32_t rb = row / MT_M, lrow = row - rb * MT_M, rp = rb >> 1;
  const uint32_t first_cta = (rp * n_tiles) / units_per_cta, last_cta = ((rp + 1) * n_tiles - 1) / units_per_cta;
  const uint32_t n_keys = (last_cta - first_cta + 1) * 2 * 2; /* segments x tile parities x (k1, k2) */
  /* every lane folds its share of the partial keys */
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
  /* the two best groups of the warp: group keys are unique, ~0 marks "none" */
  const unsigned long long g1 = warp_min_u64(k1);
  const unsigned long long mine = (k1 == g1) ? k2 : k1; /* the winner's lane offers its runner-up */
  const unsigned long long g2 = warp_min_u64(mine);
  /* rescan: lanes 0-15 the 8 columns of G1, lanes 16-31 those of G2 */
  const unsigned long long gk = (lane < 16) ? g1 : g2;
  unsigned long long cand = ~0ull;
  {
    const uint32_t b = (uint32_t)gk * MT_GROUP + (uint32_t)((lane & 15) >> 1);
    const bool ok = (gk != ~0ull) && (b < nb);
    uint32_t dot = 0, nbv = 0;
    if (ok)
    {
      const uint4 *pb = reinterpret_cast<const uint4 *>(db + (size_t)b * 128 + half * 64);
      nbv = __ldg(norm_b + b);
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
    if (ok && half == 0)
      cand = ((unsigned long long)(my_na + nbv - 2u * dot) << 32) | mt_pos(b);
  }
  const unsigned long long best = warp_min_u64(cand);
  const unsigned long long second = warp_min_u64(cand == best ? ~0ull : cand);
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

/* the two main operand blocks as one matrix of 2 * n_pad rows of 128 bytes; the two extension roles as one matrix of
 * 2 * n_pad rows of 32 bytes behind them */
static bool mt_make_maps(MatchTc *tc, CUtensorMap *map, CUtensorMap *map_ext, const void *op, uint32_t n_pad)
{
  const cuuint32_t estr[2] = {1, 1};
  {
    const cuuint64_t gdim[2] = {128, (cuuint64_t)MT_KBLK * n_pad};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {128, 128};
    if (tc->encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)op, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  const cuuint64_t gdim[2] = {32, (cuuint64_t)2 * n_pad};
  const cuuint64_t gstride[1] = {32};
  const cuuint32_t box[2] = {32, 128};
  return tc->encode(map_ext, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)((const uint8_t *)op + (size_t)MT_KBLK * n_pad * 128), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static cudaError_t match_tc_launch(void *p, const uint8_t *da, uint32_t na, const uint32_t *norm_a, const void *op_a, const uint8_t *db, uint32_t nb,
                                   const uint32_t *norm_b, const void *op_b, vksift_Match_2NN *out, cudaStream_t st, uint64_t *launch_count)
{
  MatchTc *tc = (MatchTc *)p;
  const uint32_t row_pairs = (na + 2 * MT_M - 1) / (2 * MT_M);
  const uint32_t row_blocks = 2 * row_pairs;
  const uint32_t n_tiles = (nb + MT_N - 1) / MT_N;
  const uint32_t na_pad = (na + 255u) & ~255u, nb_pad = (nb + 255u) & ~255u; /* as prepared (launch_match_prepare) */
  /* one CTA per SM: one wave of equally loaded CTAs */
  const uint32_t total_units = row_pairs * n_tiles;
  uint32_t n_cta = (uint32_t)tc->sm_count;
  if (n_cta > total_units)
    n_cta = total_units;
  const uint32_t units_per_cta = (total_units + n_cta - 1) / n_cta;
  n_cta = (total_units + units_per_cta - 1) / units_per_cta;
  const uint32_t max_segs = (n_tiles + units_per_cta - 1) / units_per_cta + 1;
  const size_t need = (size_t)row_blocks * max_segs * 2 * MT_M * 2;
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
  CUtensorMap map_a, map_b, map_ea, map_eb;
  if (!mt_make_maps(tc, &map_a, &map_ea, op_a, na_pad) || !mt_make_maps(tc, &map_b, &map_eb, op_b, nb_pad))
    return cudaErrorInvalidValue;
  match_tc_kernel<<<n_cta, MT_THREADS, MT_SMEM_BYTES, st>>>(map_a, map_b, map_ea, map_eb, na_pad, nb_pad, n_tiles, units_per_cta, total_units, max_segs,
                                                           tc->partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return e;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((na + MG_WARPS - 1) / MG_WARPS);
    cfg.blockDim = dim3(32 * MG_WARPS);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const unsigned long long *partial_c = tc->partial;
    e = cudaLaunchKernelEx(&cfg, match_merge_kernel, partial_c, n_tiles, units_per_cta, max_segs, na, nb, da, db, norm_a, norm_b, out);
  }
  *launch_count += 2;
  return e;
}

} // namespace vks
