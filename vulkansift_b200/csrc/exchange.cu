/*
 * exchange.cu -- descriptor all-gather between the GPUs of one node over NVLink peer memory.
 *
 * Cross-image matching (BASELINE configs[4]: one image per GPU, every GPU matches its features against the
 * features of every other GPU) has one exchange step.  The reference has nothing of the kind (one instance = one GPU,
 * vulkansift.h:32-34); round 1 used one NCCL all-gather plus a count read-back, 0.16 ms of three host-driven steps for 4 MB.
 * Here every rank owns one cudaMalloc'ed region that all ranks of the node map (cudaIpc handles, exchanged once by the
 * host side over torch.distributed) and the step is ONE kernel per rank:
 *
 *   region of rank r:  recv[2][world][slot_rows][128]   descriptor blocks, double buffered by the parity of the epoch
 *                      counts[2][world]                 rows of each block
 *                      flags[world]                     flags[p] = last epoch peer p has pushed completely
 *
 *   publish (epoch e): every rank PUSHES its block into recv[e & 1][rank] of every peer (posted NVLink stores, 16 bytes per
 *                      thread), the last CTA to finish writes the row count and then, behind a system-scope fence, the flag.
 *   wait    (epoch e): one small kernel polls the local flags until every peer has reached e (acquire, system scope), and
 *                      copies the counts to mapped host memory.  The blocks are then matched IN PLACE in recv
 *                      (vksiftx_matchFeaturesAgainstBlocks), no copy.
 *
 * Two buffers are enough: a rank writes epoch e+2 into the half its peers read during epoch e only after its own wait of
 * epoch e+1, i.e. after every peer has published e+1, which each peer enqueues behind its searches of epoch e (same stream).
 * The poll gives up after two seconds and reports it (a missing peer must not hang the GPU).
 */
#include "vksift_internal.h"

namespace vks
{

#define XC_HANDLE_BYTES 64
#define XC_TIMEOUT_NS 2000000000ull

struct PeerExchange
{
  int rank = 0, world = 0;
  uint32_t slot_rows = 0;
  uint32_t epoch = 0;
  uint8_t *local = nullptr;            /* this rank's region */
  uint8_t *peer[VKS_MAX_PEERS] = {};   /* every rank's region as mapped here (peer[rank] == local) */
  bool opened[VKS_MAX_PEERS] = {};
  uint32_t *d_done = nullptr;          /* CTA counter of the publish kernel */
  uint32_t *h_counts = nullptr;        /* mapped pinned: counts[world], status */
  uint32_t *h_counts_dev = nullptr;
  size_t region_bytes = 0;
};

struct XcPeers
{
  uint8_t *base[VKS_MAX_PEERS];
};

__host__ __device__ inline size_t xc_recv_bytes(int world, uint32_t slot_rows) { return (size_t)2 * world * slot_rows * 128; }
__host__ __device__ inline size_t xc_block_off(int world, uint32_t slot_rows, uint32_t half, int src)
{
  return ((size_t)half * world + src) * slot_rows * 128;
}
__host__ __device__ inline size_t xc_counts_off(int world, uint32_t slot_rows, uint32_t half) { return xc_recv_bytes(world, slot_rows) + (size_t)half * 128; }
__host__ __device__ inline size_t xc_flags_off(int world, uint32_t slot_rows) { return xc_recv_bytes(world, slot_rows) + 256; }
static_assert(VKS_MAX_PEERS * 4 <= 128, "counts and flags of all peers fit one 128-byte line each");

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

/* grid (chunks, world): CTA (c, p) pushes rows [c * rows_per_cta, ...) of this rank's block to peer p */
#define XC_THREADS 256
__global__ void __launch_bounds__(XC_THREADS) exchange_publish_kernel(const XcPeers peers, const int rank, const int world, const uint32_t slot_rows,
                                                                      const uint32_t epoch, const uint8_t *__restrict__ desc, const uint32_t n,
                                                                      const uint32_t rows_per_cta, uint32_t *__restrict__ done)
{
  const int p = (int)blockIdx.y;
  const uint32_t half = epoch & 1u;
  if (p != rank) /* a rank does not match against itself: its own block only gets a count */
  {
    const uint32_t r0 = blockIdx.x * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
    if (r1 > r0)
    {
      const uint4 *src = reinterpret_cast<const uint4 *>(desc + (size_t)r0 * 128);
      uint4 *dst = reinterpret_cast<uint4 *>(peers.base[p] + xc_block_off(world, slot_rows, half, rank) + (size_t)r0 * 128);
      const uint32_t n16 = (r1 - r0) * 8u;
      for (uint32_t i = threadIdx.x; i < n16; i += XC_THREADS)
        dst[i] = __ldg(src + i);
    }
  }
  /* the stores above are visible system-wide before this CTA counts itself done */
  __threadfence_system();
  __syncthreads();
  __shared__ uint32_t s_last;
  if (threadIdx.x == 0)
    s_last = (atomicAdd(done, 1u) == gridDim.x * gridDim.y - 1u) ? 1u : 0u;
  __syncthreads();
  if (!s_last)
    return;
  __threadfence_system();
  if ((int)threadIdx.x < world)
  {
    const int q = (int)threadIdx.x;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(peers.base[q] + xc_counts_off(world, slot_rows, half)) + rank;
    uint32_t *flag = reinterpret_cast<uint32_t *>(peers.base[q] + xc_flags_off(world, slot_rows)) + rank;
    *cnt = n;
    st_release_sys(flag, epoch); /* release: the count and (by the fences and the counter above) every row are ordered before it */
  }
  if (threadIdx.x == 0)
    *done = 0; /* next epoch */
}

/* one CTA: thread p waits for peer p; counts and status go to mapped host memory */
__global__ void exchange_wait_kernel(uint8_t *__restrict__ local, const int world, const uint32_t slot_rows, const uint32_t epoch,
                                     uint32_t *__restrict__ h_counts)
{
  const int p = (int)threadIdx.x;
  __shared__ uint32_t s_timeout;
  if (p == 0)
    s_timeout = 0;
  __syncthreads();
  if (p < world)
  {
    const uint32_t *flag = reinterpret_cast<const uint32_t *>(local + xc_flags_off(world, slot_rows)) + p;
    const unsigned long long t0 = globaltimer_ns();
    bool ok = true;
    while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0)
    {
      __nanosleep(200);
      if (globaltimer_ns() - t0 > XC_TIMEOUT_NS)
      {
        ok = false;
        break;
      }
    }
    if (!ok)
      atomicOr(&s_timeout, 1u << p);
    const uint32_t *cnt = reinterpret_cast<const uint32_t *>(local + xc_counts_off(world, slot_rows, epoch & 1u)) + p;
    h_counts[p] = ok ? *reinterpret_cast<const volatile uint32_t *>(cnt) : 0u;
  }
  __syncthreads();
  if (p == 0)
  {
    h_counts[VKS_MAX_PEERS] = s_timeout; /* bit p: peer p did not arrive */
    __threadfence_system();
  }
}

/* After the wait: zero the rows of every received block between its count and the largest count rounded up to a 128-row tile.
 * One search then covers all blocks with one tensor map (launch_match_blocks), and a slot keeps rows of earlier, larger blocks. */
__global__ void __launch_bounds__(XC_THREADS) exchange_pad_kernel(uint8_t *__restrict__ local, const int rank, const int world, const uint32_t slot_rows,
                                                                  const uint32_t epoch)
{
  const int p = (int)blockIdx.x;
  if (p == rank)
    return;
  const uint32_t half = epoch & 1u;
  const uint32_t *cnt = reinterpret_cast<const uint32_t *>(local + xc_counts_off(world, slot_rows, half));
  uint32_t mx = 0;
  for (int q = 0; q < world; q++)
    if (q != rank)
      mx = max(mx, cnt[q]);
  const uint32_t pad_to = min(slot_rows, (mx + 127u) & ~127u), n = min(cnt[p], pad_to);
  uint4 *dst = reinterpret_cast<uint4 *>(local + xc_block_off(world, slot_rows, half, p) + (size_t)n * 128);
  const uint32_t n16 = (pad_to - n) * 8u;
  for (uint32_t i = threadIdx.x; i < n16; i += XC_THREADS)
    dst[i] = make_uint4(0u, 0u, 0u, 0u);
}

cudaError_t exchange_create(PeerExchange **out, int rank, int world, uint32_t slot_rows, void *handle)
{
  if (world < 1 || world > VKS_MAX_PEERS || rank < 0 || rank >= world || slot_rows == 0 || (slot_rows % 128u) != 0)
    return cudaErrorInvalidValue;
  PeerExchange *x = new PeerExchange();
  x->rank = rank;
  x->world = world;
  x->slot_rows = slot_rows;
  x->region_bytes = xc_flags_off(world, slot_rows) + 128;
  cudaError_t e = cudaMalloc(&x->local, x->region_bytes);
  if (e == cudaSuccess)
    e = cudaMemset(x->local, 0, x->region_bytes);
  if (e == cudaSuccess)
    e = cudaMalloc(&x->d_done, sizeof(uint32_t));
  if (e == cudaSuccess)
    e = cudaMemset(x->d_done, 0, sizeof(uint32_t));
  if (e == cudaSuccess)
    e = cudaHostAlloc(&x->h_counts, sizeof(uint32_t) * (VKS_MAX_PEERS + 1), cudaHostAllocMapped);
  if (e == cudaSuccess)
    e = cudaHostGetDevicePointer(&x->h_counts_dev, x->h_counts, 0);
  if (e == cudaSuccess)
    e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  static_assert(sizeof(cudaIpcMemHandle_t) == XC_HANDLE_BYTES, "IPC handle size");
  if (e == cudaSuccess)
    e = cudaIpcGetMemHandle(&h, x->local);
  if (e != cudaSuccess)
  {
    exchange_destroy(x);
    return e;
  }
  memcpy(handle, &h, XC_HANDLE_BYTES);
  x->peer[rank] = x->local;
  *out = x;
  return cudaSuccess;
}

cudaError_t exchange_connect(PeerExchange *x, const void *handles)
{
  for (int p = 0; p < x->world; p++)
  {
    if (p == x->rank || x->opened[p])
      continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const uint8_t *)handles + (size_t)p * XC_HANDLE_BYTES, XC_HANDLE_BYTES);
    void *ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return e;
    x->peer[p] = (uint8_t *)ptr;
    x->opened[p] = true;
  }
  return cudaSuccess;
}

cudaError_t exchange_allgather(PeerExchange *x, const uint8_t *desc, uint32_t n, cudaStream_t st, uint64_t *launch_count)
{
  if (n > x->slot_rows)
    return cudaErrorInvalidValue;
  for (int p = 0; p < x->world; p++)
    if (!x->peer[p])
      return cudaErrorNotReady; /* exchange_connect has not run */
  x->epoch++;
  XcPeers peers;
  for (int p = 0; p < VKS_MAX_PEERS; p++)
    peers.base[p] = x->peer[p];
  /* 64 rows (8 KB) per CTA: a 3.5 k block is 55 CTAs per peer, enough stores in flight to fill the links of a 7-peer push */
  const uint32_t rows_per_cta = 64;
  const uint32_t chunks = n > 0 ? (n + rows_per_cta - 1) / rows_per_cta : 1u;
  exchange_publish_kernel<<<dim3(chunks, (unsigned)x->world), XC_THREADS, 0, st>>>(peers, x->rank, x->world, x->slot_rows, x->epoch, desc, n, rows_per_cta,
                                                                                   x->d_done);
  exchange_wait_kernel<<<1, 32, 0, st>>>(x->local, x->world, x->slot_rows, x->epoch, x->h_counts_dev);
  exchange_pad_kernel<<<(unsigned)x->world, XC_THREADS, 0, st>>>(x->local, x->rank, x->world, x->slot_rows, x->epoch);
  *launch_count += 3;
  return cudaGetLastError();
}

const uint32_t *exchange_host_counts(const PeerExchange *x) { return x->h_counts; }
uint32_t exchange_timeout_mask(const PeerExchange *x) { return x->h_counts[VKS_MAX_PEERS]; }
const uint8_t *exchange_blocks(const PeerExchange *x, uint64_t *stride_bytes)
{
  *stride_bytes = (uint64_t)x->slot_rows * 128;
  return x->local + xc_block_off(x->world, x->slot_rows, x->epoch & 1u, 0);
}
int exchange_world(const PeerExchange *x) { return x->world; }
uint32_t exchange_slot_rows(const PeerExchange *x) { return x->slot_rows; }
int exchange_rank(const PeerExchange *x) { return x->rank; }

void exchange_destroy(PeerExchange *x)
{
  if (!x)
    return;
  for (int p = 0; p < x->world; p++)
    if (x->opened[p])
      cudaIpcCloseMemHandle(x->peer[p]);
  cudaFree(x->local);
  cudaFree(x->d_done);
  if (x->h_counts)
    cudaFreeHost(x->h_counts);
  delete x;
}

} // namespace vks
