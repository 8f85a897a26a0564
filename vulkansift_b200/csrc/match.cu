/*
 * match.cu -- brute-force 2-nearest-neighbour search on 128-byte descriptors.
 *
 * Replaces shaders/Get2NearestNeighbors.comp:43-104 (dispatched by
 * sift_matcher.c:246-279).  d(a,b)^2 = |a|^2 + |b|^2 - 2 a.b with exact integer
 * arithmetic; the a.b term is a dense u8 x u8 -> s32 contraction.
 *
 * Tie rule of the shader (strict '<', b=0 and b=1 initialised specially,
 * :69-96) == stable top-2 of B under the key (d, pos) with pos(0)=1, pos(1)=0,
 * pos(b)=b.  Distances are compared as squared integers, which orders exactly
 * like the shader's float sqrt while d^2 < 2^22 (|a-b| < 2048; SIFT descriptors
 * have |a-b|^2 <= 2*512^2 < 2^20); rows whose second neighbour lies beyond that are
 * rescanned under the float key (match_merge_kernel), so the result is the shader's
 * for any byte descriptors.
 */
#include "vksift_internal.h"

#include "match_tc.cuh"

namespace vks
{

struct MatchWorkspace
{
  uint32_t max_feats;
  uint32_t *norm_a; /* |a|^2 per row */
  uint32_t *norm_b;
  unsigned long long *partial; /* [splits][na][2] packed (d2 << 32 | pos) keys */
  uint32_t partial_splits;
  void *tc; /* tensor-core path state (tensor maps) */
};

/* ---- |x|^2 per descriptor ------------------------------------------------ */
__device__ __forceinline__ uint32_t match_pos_host_device(uint32_t b) { return b < 2u ? (b ^ 1u) : b; }
/* |x|^2 per descriptor.  packed == 0: plain norms (A side).  packed == 1 (B side): nbk = |b|^2 * 256 + (pos(b) & 255),
 * the per-column constant of the tensor-core epilogue's 32-bit keys; |b|^2 <= 128*255^2 < 2^23 so nbk fits int32.
 * Rows in [n, n_padded) do not exist and get the largest nbk: they lose every comparison. */
#define NORM_PAD_VALUE 0x7fffffffu
__global__ void norms_kernel(const uint8_t *__restrict__ desc, uint32_t n, uint32_t n_padded, uint32_t *__restrict__ out, int packed)
{
  /* one warp per descriptor, 4 bytes per lane */
  const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n)
  {
    if (row < n_padded && lane == 0)
      out[row] = NORM_PAD_VALUE;
    return;
  }
  const uint32_t v = __ldg((const uint32_t *)(desc + (size_t)row * 128) + lane);
  uint32_t s = __dp4a(v, v, 0u);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, d);
  if (lane == 0)
    out[row] = packed ? (s * 256u + (match_pos_host_device(row) & 255u)) : s;
}

/* ---- SIMT cross-check kernel (verification only, vksiftx_setMatcherImpl(1)) */
#define MS_ROWS 128
#define MS_BTILE 64
__device__ __forceinline__ void top2_insert(unsigned long long key, unsigned long long &k1, unsigned long long &k2)
{
  if (key < k1)
  {
    k2 = k1;
    k1 = key;
  }
  else if (key < k2)
    k2 = key;
}
__device__ __forceinline__ uint32_t match_pos(uint32_t b) { return b < 2u ? (b ^ 1u) : b; }

__global__ void __launch_bounds__(MS_ROWS) match_simt_kernel(const uint8_t *__restrict__ da, uint32_t na, const uint32_t *__restrict__ norm_a,
                                                             const uint8_t *__restrict__ db, uint32_t nb, const uint32_t *__restrict__ norm_b,
                                                             vksift_Match_2NN *__restrict__ out)
{
  __shared__ uint32_t s_b[MS_BTILE][32];
  __shared__ uint32_t s_nb[MS_BTILE];
  const uint32_t row = blockIdx.x * MS_ROWS + threadIdx.x;
  uint32_t a[32];
  const uint32_t arow = min(row, na - 1);
#pragma unroll
  for (int i = 0; i < 32; i++)
    a[i] = __ldg((const uint32_t *)(da + (size_t)arow * 128) + i);
  const uint32_t my_na = norm_a[arow];
  unsigned long long k1 = ~0ull, k2 = ~0ull;
  for (uint32_t b0 = 0; b0 < nb; b0 += MS_BTILE)
  {
    const uint32_t cnt = min((uint32_t)MS_BTILE, nb - b0);
    for (uint32_t i = threadIdx.x; i < cnt * 32; i += MS_ROWS)
      s_b[i >> 5][i & 31] = __ldg((const uint32_t *)(db + (size_t)b0 * 128) + i);
    if (threadIdx.x < cnt)
      s_nb[threadIdx.x] = norm_b[b0 + threadIdx.x] >> 8; /* norm_b holds the packed nbk */
    __syncthreads();
    for (uint32_t j = 0; j < cnt; j++)
    {
      uint32_t dot = 0;
#pragma unroll
      for (int i = 0; i < 32; i++)
        dot = __dp4a(a[i], s_b[j][i], dot);
      const uint32_t d2 = my_na + s_nb[j] - 2u * dot;
      /* key = (bits of sqrt(float(d^2)), pos): the shader's comparison, exact for every d^2 (non-negative floats order like their bits) */
      top2_insert(((unsigned long long)__float_as_uint(vks_sqrt((float)d2)) << 32) | match_pos(b0 + j), k1, k2);
    }
    __syncthreads();
  }
  if (row < na)
  {
    vksift_Match_2NN m;
    m.idx_a = row;
    m.idx_b1 = match_pos((uint32_t)k1);
    m.idx_b2 = match_pos((uint32_t)k2);
    m.dist_a_b1 = __uint_as_float((uint32_t)(k1 >> 32));
    m.dist_a_b2 = __uint_as_float((uint32_t)(k2 >> 32));
    out[row] = m;
  }
}

cudaError_t match_workspace_create(MatchWorkspace **out, uint32_t max_feats)
{
  MatchWorkspace *ws = new MatchWorkspace();
  ws->max_feats = max_feats;
  ws->tc = nullptr;
  ws->partial = nullptr;
  ws->partial_splits = 0;
  cudaError_t e = cudaMalloc(&ws->norm_a, sizeof(uint32_t) * (size_t)(max_feats + 256));
  if (e == cudaSuccess)
    e = cudaMalloc(&ws->norm_b, sizeof(uint32_t) * (size_t)(max_feats + 256));
  if (e == cudaSuccess)
    e = match_tc_create(&ws->tc, max_feats);
  if (e != cudaSuccess)
  {
    match_workspace_destroy(ws);
    return e;
  }
  *out = ws;
  return cudaSuccess;
}

void match_workspace_destroy(MatchWorkspace *ws)
{
  if (!ws)
    return;
  match_tc_destroy(ws->tc);
  cudaFree(ws->norm_a);
  cudaFree(ws->norm_b);
  cudaFree(ws->partial);
  delete ws;
}

/* ---- mutual-nearest-neighbour + Lowe ratio filter -------------------------------------------------
 * What every caller of the reference does on the CPU right after downloading two match lists
 * (src/examples/test_sift_match.cpp:90-107, src/perf/perf_common.cpp:122-170):
 *   keep i iff m21[m12[i].idx_b1].idx_b1 == i  and  d1/d2 < ratio in both directions.
 * One CTA, ordered compaction: the pairs come out in increasing idx_a, like the reference loop produces them. */
#define FILT_THREADS 1024
__global__ void __launch_bounds__(FILT_THREADS) match_filter_kernel(const vksift_Match_2NN *__restrict__ m12, uint32_t na,
                                                                    const vksift_Match_2NN *__restrict__ m21, uint32_t nb, float ratio,
                                                                    uint32_t *__restrict__ pairs, uint32_t capacity, uint32_t *__restrict__ count)
{
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_base;
  const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
  if (tid == 0)
    s_base = 0;
  __syncthreads();
  for (uint32_t base = 0; base < na; base += FILT_THREADS)
  {
    const uint32_t i = base + tid;
    bool keep = false;
    uint32_t j = 0;
    if (i < na)
    {
      const vksift_Match_2NN a = m12[i];
      j = a.idx_b1;
      if (j < nb)
      {
        const vksift_Match_2NN b = m21[j];
        keep = (b.idx_b1 == i) && ((a.dist_a_b1 / a.dist_a_b2) < ratio) && ((b.dist_a_b1 / b.dist_a_b2) < ratio);
      }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0)
      s_warp[wi] = __popc(bal);
    __syncthreads();
    if (wi == 0)
    {
      uint32_t v = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1)
      {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d)
          v += t;
      }
      s_warp[lane] = v; /* inclusive */
    }
    __syncthreads();
    const uint32_t start = s_base + (wi ? s_warp[wi - 1] : 0u) + __popc(bal & ((1u << lane) - 1u));
    if (keep && start < capacity)
    {
      pairs[2 * start + 0] = i;
      pairs[2 * start + 1] = j;
    }
    __syncthreads();
    if (tid == 0)
      s_base += s_warp[31];
    __syncthreads();
  }
  if (tid == 0)
    *count = s_base;
}

cudaError_t launch_match_filter(const vksift_Match_2NN *m12, uint32_t na, const vksift_Match_2NN *m21, uint32_t nb, float ratio, uint32_t *pairs,
                                uint32_t capacity, uint32_t *count, cudaStream_t st)
{
  match_filter_kernel<<<1, FILT_THREADS, 0, st>>>(m12, na, m21, nb, ratio, pairs, capacity, count);
  return cudaGetLastError();
}

__global__ void norms_both_kernel(const uint8_t *__restrict__ desc, uint32_t n, uint32_t n_padded, uint32_t *__restrict__ out_plain,
                                  uint32_t *__restrict__ out_packed)
{
  const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n)
  {
    if (row < n_padded && lane == 0)
      out_packed[row] = NORM_PAD_VALUE;
    return;
  }
  const uint32_t v = __ldg((const uint32_t *)(desc + (size_t)row * 128) + lane);
  uint32_t s = __dp4a(v, v, 0u);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, d);
  if (lane == 0)
  {
    out_plain[row] = s;
    out_packed[row] = s * 256u + (match_pos_host_device(row) & 255u);
  }
}

/* packed B-side norms of several descriptor blocks (block j = counts.n[j] rows at desc + j * stride_rows * 128) in one launch:
 * out[j * stride_rows + i] like norms_kernel(packed) of block j, rows past the block's count marked absent */
__global__ void norms_blocks_kernel(const uint8_t *__restrict__ desc, const MatchBlockCounts counts, uint32_t n_blocks, uint32_t stride_rows,
                                    uint32_t *__restrict__ out_packed)
{
  const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_blocks * stride_rows)
    return;
  const uint32_t j = row / stride_rows, i = row - j * stride_rows;
  if (i >= counts.n[j])
  {
    if (lane == 0)
      out_packed[row] = NORM_PAD_VALUE;
    return;
  }
  const uint32_t v = __ldg((const uint32_t *)(desc + (size_t)row * 128) + lane);
  uint32_t s = __dp4a(v, v, 0u);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, d);
  if (lane == 0)
    out_packed[row] = s * 256u + (match_pos_host_device(i) & 255u);
}

cudaError_t launch_norms_blocks(const uint8_t *desc, const MatchBlockCounts &counts, uint32_t n_blocks, uint32_t stride_rows, uint32_t *out_packed,
                                cudaStream_t st)
{
  const size_t rows = (size_t)n_blocks * stride_rows;
  if (rows == 0)
    return cudaSuccess;
  norms_blocks_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(desc, counts, n_blocks, stride_rows, out_packed);
  return cudaGetLastError();
}

cudaError_t launch_norms(const uint8_t *desc, uint32_t n, uint32_t *out_plain, uint32_t *out_packed, cudaStream_t st)
{
  const uint32_t n_pad = (n + 127u) & ~127u;
  if (n_pad == 0)
    return cudaSuccess;
  norms_both_kernel<<<(n_pad * 32 + 255) / 256, 256, 0, st>>>(desc, n, n_pad, out_plain, out_packed);
  return cudaGetLastError();
}

cudaError_t launch_match(MatchWorkspace *ws, int impl, const uint8_t *da, uint32_t na, const uint32_t *norm_a, const uint8_t *db, uint32_t nb,
                         const uint32_t *norm_b, vksift_Match_2NN *out, cudaStream_t st, cudaEvent_t ev_after_prepare, bool inputs_settled,
                         uint64_t *launch_count)
{
  if (na == 0)
    return cudaSuccess;
  /* norms computed here are inputs written by the launch just before the search: no overlap with the previous search then */
  if (!norm_a || !norm_b)
    inputs_settled = false;
  const uint32_t nb_pad = (nb + 127u) & ~127u; /* the tensor-core path reads |b|^2 in tiles of 128 */
  if (!norm_a)
  {
    norms_kernel<<<(na * 32 + 255) / 256, 256, 0, st>>>(da, na, na, ws->norm_a, 0);
    *launch_count += 1;
    norm_a = ws->norm_a;
  }
  if (!norm_b)
  {
    norms_kernel<<<(nb_pad * 32 + 255) / 256, 256, 0, st>>>(db, nb, nb_pad, ws->norm_b, 1);
    *launch_count += 1;
    norm_b = ws->norm_b;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return e;
  if (ev_after_prepare)
    cudaEventRecord(ev_after_prepare, st);
  if (impl == 1)
  {
    match_simt_kernel<<<(na + MS_ROWS - 1) / MS_ROWS, MS_ROWS, 0, st>>>(da, na, norm_a, db, nb, norm_b, out);
    *launch_count += 1;
    return cudaGetLastError();
  }
  return match_tc_launch(ws->tc, da, na, norm_a, db, nb, norm_b, out, st, inputs_settled, launch_count);
}

cudaError_t launch_match_blocks(MatchWorkspace *ws, const uint8_t *da, uint32_t na, const uint32_t *norm_a, const uint8_t *blocks,
                                const uint32_t *norm_blocks, uint32_t stride_rows, const uint32_t *blk, const uint32_t *cnt, uint32_t n_groups,
                                vksift_Match_2NN *out, uint32_t out_stride, cudaStream_t st, bool inputs_settled, uint64_t *launch_count)
{
  if (na == 0 || n_groups == 0)
    return cudaSuccess;
  if (n_groups > MT_MAX_GROUPS || (stride_rows % MT_N) != 0)
    return cudaErrorInvalidValue;
  MatchGroups G;
  memset(&G, 0, sizeof(G));
  G.n_groups = n_groups;
  G.stride_rows = stride_rows;
  G.out_stride = out_stride;
  for (uint32_t g = 0; g < n_groups; g++)
  {
    G.blk[g] = blk[g];
    G.cnt[g] = cnt[g];
  }
  return match_tc_launch_groups(ws->tc, da, na, norm_a, blocks, norm_blocks, out, G, st, inputs_settled, launch_count);
}

} // namespace vks
