/* TMA fetch throughput against box shape (B200): what sets the tile period of the extrema scan and the blur passes?
 *
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vulkansift_b200/csrc tools/ubench/tma_fetch.cu -o tools/ubench/tma_fetch
 *
 * Persistent CTAs walk the tiles of a [layers][H][W] fp32 tensor (3840 x 2160 x 5 = 166 MB, the DoG stack of octave 0) with a
 * ring of tile buffers; a tile is `req` TMA requests of box (bw x bh x bl) elements.  Consumers only wait for the barrier and
 * touch one word per warp, so the time is the fetch path alone.  Printed: tile geometry, bytes per box row, GB/s over the
 * tensor bytes (DRAM reads when the tensor does not fit L2 ... it does not: 166 MB > 126 MB), B/clk/SM including halos. */
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tma_util.cuh"

using namespace vks;

struct Shape
{
  const char *name;
  int elem;      /* 4 = fp32 map, 8 = 64-bit map over the same memory */
  int tw, th;    /* tile core in floats / rows (what a consumer would test) */
  int bw, bh;    /* box in floats / rows (core + halo) */
  int per_layer; /* 1: one request per layer (box depth 1); 0: one request for all layers (box depth = layers) */
  int ring;      /* tile buffers per CTA */
  int ctas;      /* CTAs per SM */
};

__global__ void __launch_bounds__(288) fetch_kernel(const __grid_constant__ CUtensorMap map, int tiles_x, int n_tiles, int tw, int th, int halo_x, int elem,
                                                    int layers, int per_layer, int ring, uint32_t tile_bytes, uint32_t layer_bytes, float *sink)
{
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) uint64_t bars[16]; /* full[8], empty[8] */
  const int tid = threadIdx.x;
  const uint32_t full = tma_smem_u32(&bars[0]), empty = tma_smem_u32(&bars[8]);
  if (tid == 0)
  {
    for (int b = 0; b < 8; b++)
    {
      tma_mbar_init(full + 8 * b, 1);
      tma_mbar_init(empty + 8 * b, 8);
    }
    tma_mbar_fence_init();
  }
  __syncthreads();
  const uint32_t buf_bytes = (tile_bytes + 127u) & ~127u;
  if (tid >= 256)
  {
    if (tid == 256)
    {
      int buf = 0, use = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
      {
        if (use >= 1)
        {
          tma_mbar_wait(empty + 8 * buf, (uint32_t)(use - 1) & 1u);
          tma_fence_proxy_async();
        }
        const int x0 = (t % tiles_x) * tw - halo_x, y0 = (t / tiles_x) * th - 1;
        const int cx = elem == 8 ? x0 / 2 : x0;
        tma_mbar_expect_tx(full + 8 * buf, tile_bytes);
        const uint32_t dst = tma_smem_u32(sm) + buf * buf_bytes;
        if (per_layer)
          for (int l = 0; l < layers; l++)
            tma_load_3d(dst + l * layer_bytes, &map, cx, y0, l, full + 8 * buf);
        else
          tma_load_3d(dst, &map, cx, y0, 0, full + 8 * buf);
        if (++buf == ring)
        {
          buf = 0;
          use++;
        }
      }
    }
    return;
  }
  int buf = 0, use = 0;
  float acc = 0.f;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
  {
    tma_mbar_wait(full + 8 * buf, (uint32_t)use & 1u);
    acc += sm[(buf * buf_bytes) / 4 + (tid & 31)];
    __syncwarp();
    if ((tid & 31) == 0)
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty + 8 * buf) : "memory");
    if (++buf == ring)
    {
      buf = 0;
      use++;
    }
  }
  if (acc == 123.456f)
    sink[0] = acc;
}

int main()
{
  const int W = 3840, H = 2160, L = 5;
  const size_t n = (size_t)W * H * L;
  float *d = nullptr, *sink = nullptr;
  cudaMalloc(&d, n * 4);
  cudaMalloc(&sink, 4);
  cudaMemset(d, 0, n * 4);
  PFN_encodeTiled enc = tma_encoder();
  if (!enc)
  {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  int sms = 148, clk_khz = 1965000;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  cudaFuncSetAttribute(fetch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const Shape shapes[] = {
      {"extrema now: 240x8 (+halo 256x10), 5 requests", 4, 240, 8, 256, 10, 1, 2, 2},
      {"same, one 3-D request per tile", 4, 240, 8, 256, 10, 0, 2, 2},
      {"same, ring of 4, one CTA per SM", 4, 240, 8, 256, 10, 1, 4, 1},
      {"16-row tiles: 240x16 (256x18), one CTA per SM", 4, 240, 16, 256, 18, 1, 2, 1},
      {"4-row tiles: 240x4 (256x6), 3 CTAs per SM", 4, 240, 4, 256, 6, 1, 2, 3},
      {"64-bit elements: 496x8 (512x10), one CTA per SM", 8, 496, 8, 512, 10, 1, 2, 1},
      {"64-bit elements: 496x4 (512x6), ring 3, one CTA per SM", 8, 496, 4, 512, 6, 1, 3, 1},
      {"narrow rows: 112x8 (128x10), 4 CTAs per SM", 4, 112, 8, 128, 10, 1, 2, 4},
      {"blur-like: 64x64 (+4: 76x72) one layer, 4 CTAs per SM", 4, 64, 64, 76, 72, 2, 2, 4},
      {"blur-like wide: 128x64 (140x72) one layer, 3 CTAs per SM", 4, 128, 64, 140, 72, 2, 2, 3},
  };
  for (const Shape &s : shapes)
  {
    const int layers = (s.per_layer == 2) ? 1 : L; /* per_layer == 2: a single layer (the blur reads one source layer) */
    CUtensorMap map;
    const int ew = s.elem == 8 ? 2 : 1;
    cuuint64_t gdim[3] = {(cuuint64_t)(W / ew), (cuuint64_t)H, (cuuint64_t)L};
    cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)(s.bw / ew), (cuuint32_t)s.bh, (cuuint32_t)((s.per_layer == 0) ? L : 1)};
    cuuint32_t es[3] = {1, 1, 1};
    /* rows of 76 / 140 floats are not multiples of 16 bytes ... round the box up */
    if ((box[0] * 4 * ew) % 16)
      box[0] = (box[0] + 3) & ~3u;
    CUresult r = enc(&map, s.elem == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
    {
      printf("%-58s  map rejected (%d)\n", s.name, (int)r);
      continue;
    }
    const int row_bytes = (int)box[0] * 4 * ew;
    const uint32_t layer_bytes = (uint32_t)row_bytes * s.bh;
    const uint32_t tile_bytes = layer_bytes * layers;
    const size_t smem = (size_t)s.ring * ((tile_bytes + 127u) & ~127u);
    if (smem * s.ctas > 220 * 1024)
    {
      printf("%-58s  does not fit (%zu KB x %d)\n", s.name, smem / 1024, s.ctas);
      continue;
    }
    const int tiles_x = (W + s.tw - 1) / s.tw, tiles_y = (H + s.th - 1) / s.th;
    const int n_tiles = tiles_x * tiles_y;
    const int halo_x = (s.bw - s.tw) / 2 >= 4 ? 4 : (s.bw - s.tw) / 2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 6; rep++)
    {
      cudaEventRecord(e0);
      fetch_kernel<<<sms * s.ctas, 288, smem>>>(map, tiles_x, n_tiles, s.tw, s.th, halo_x & ~1, s.elem, layers, s.per_layer == 0 ? 0 : 1, s.ring, tile_bytes,
                                                 layer_bytes, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best)
        best = ms;
    }
    const cudaError_t err = cudaGetLastError();
    const double tensor_bytes = (double)W * H * 4 * layers;
    const double fetched = (double)n_tiles * tile_bytes;
    const double rows = (double)n_tiles * s.bh * layers;
    const double clk = best * 1e-3 * (double)clk_khz * 1e3;
    printf("%-58s  row %4d B  %7.1f us  %5.0f GB/s of tensor  %5.1f B/clk/SM fetched  %5.1f clk per box row per SM  %s\n", s.name, row_bytes,
           best * 1e3, tensor_bytes / (best * 1e-3) / 1e9, fetched / clk / sms, clk / (rows / sms), err == cudaSuccess ? "" : cudaGetErrorName(err));
  }
  return 0;
}
