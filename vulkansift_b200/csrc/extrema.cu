/*
 * extrema.cu -- DoG extrema scan, sub-pixel refinement and deterministic ordering.
 *
 * Replaces ExtractKeypoints.comp (reference: shaders/ExtractKeypoints.comp:44-231,
 * dispatched by sift_detector.c:1106-1189).  The reference appends accepted
 * keypoints with atomicAdd, i.e. in a race-dependent order; here accepted
 * keypoints carry the id of the thread that detected them, key = (s, y, x), and
 * a rank pass places them in increasing key order -- the canonical order of
 * SURVEY.md B-D4, which the oracle produces by construction.  Section overflow
 * therefore drops the highest keys deterministically.
 *
 * All float expressions keep the association order of the shader; the library
 * is compiled with -fmad=false so nothing is contracted.
 */
#include "vksift_internal.h"

namespace vks
{

/* imageLoad on the DoG array; layers beyond ns+1 do not exist -> 0 (SURVEY B-D3) */
__device__ __forceinline__ float dog_at(const OctaveView &ov, int ns, int s, int x, int y)
{
  if (s < 0 || s >= ns + 2 || x < 0 || x >= ov.w || y < 0 || y >= ov.h)
    return 0.f;
  return __ldg(ov.D + (size_t)s * ov.layer_stride + (size_t)y * ov.pitch + x);
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
 * Here one CTA stages a 64x32 tile (+1 halo) of ALL ns+2 DoG layers of an octave in shared memory once
 * (every DoG value is read from HBM a single time), then tests the ns inner scales from smem:
 * prefilter |v| > 0.8*thr, strict 26-neighbour extremum, and for the few survivors the refinement
 * (which reads global memory, it walks outside the tile).  All octaves run in one launch. */
#define EX_TW 64
#define EX_TH 32
#define EX_SW 72 /* smem row: columns x0-4 .. x0+67 as 18 aligned float4 */
#define EX_SH (EX_TH + 2)

__global__ void __launch_bounds__(256) extrema_kernel(const __grid_constant__ DetectParams P, Candidate *__restrict__ cand,
                                                      DetectCounters *__restrict__ cnt)
{
  extern __shared__ __align__(16) float ex_smem[];
  /* linear CTA index -> (octave, tile) */
  int t = blockIdx.x, o = 0, tx = 0;
  for (; o < P.n_oct; o++)
  {
    tx = (P.oct[o].w + EX_TW - 1) / EX_TW;
    const int n = tx * ((P.oct[o].h + EX_TH - 1) / EX_TH);
    if (t < n)
      break;
    t -= n;
  }
  if (o >= P.n_oct)
    return;
  const OctaveView &ov = P.oct[o];
  const int nl = P.ns + 2;
  const int x0 = (t % tx) * EX_TW, y0 = (t / tx) * EX_TH;
  const int tid = threadIdx.x;

  /* stage: layer l, smem row r <-> image row y0-1+r, smem col c <-> image col x0-4+c.
   * Loads are issued in batches of 6 per thread before any store so their latencies overlap. */
  const int n_items = nl * EX_SH * (EX_SW / 4);
  for (int it0 = tid; it0 < n_items; it0 += 6 * 256)
  {
    float4 v[6];
#pragma unroll
    for (int k = 0; k < 6; k++)
    {
      const int it = it0 + k * 256;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (it < n_items)
      {
        const int l = it / (EX_SH * (EX_SW / 4));
        const int rem = it - l * (EX_SH * (EX_SW / 4));
        const int r = rem / (EX_SW / 4), c4 = rem - r * (EX_SW / 4);
        const int gy = y0 - 1 + r, gx = x0 - 4 + c4 * 4;
        if (gy >= 0 && gy < ov.h && gx >= 0 && gx + 3 < ov.pitch)
          v[k] = __ldg((const float4 *)(ov.D + (size_t)l * ov.layer_stride + (size_t)gy * ov.pitch + gx));
      }
    }
#pragma unroll
    for (int k = 0; k < 6; k++)
    {
      const int it = it0 + k * 256;
      if (it < n_items)
        *(float4 *)(ex_smem + it * 4) = v[k]; /* item order == smem order: (l, r, c4) row-major */
    }
  }
  __syncthreads();

  const int lx = tid & 63, ly = (tid >> 6) * 8;
  const int x = x0 + lx;
  if (x < 1 || x >= ov.w - 1)
    return;
  for (int s = 1; s <= P.ns; s++)
  {
    const float *Ls = ex_smem + (s * EX_SH) * EX_SW + (lx + 4);
#pragma unroll 1
    for (int r = 0; r < 8; r++)
    {
      const int y = y0 + ly + r;
      if (y < 1 || y >= ov.h - 1)
        continue;
      const float *pc = Ls + (ly + r + 1) * EX_SW;
      const float c = pc[0];
      if (!(fabsf(c) > P.prefilter))
        continue;
      bool gt = true, lt = true;
#pragma unroll
      for (int ds = -1; ds <= 1; ds++)
#pragma unroll
        for (int dy = -1; dy <= 1; dy++)
#pragma unroll
          for (int dx = -1; dx <= 1; dx++)
          {
            if (ds == 0 && dy == 0 && dx == 0)
              continue;
            const float n = pc[ds * (EX_SH * EX_SW) + dy * EX_SW + dx];
            gt = gt && (c > n);
            lt = lt && (c < n);
          }
      if (!(gt || lt))
        continue;
      FeatHead hd;
      if (!refine_keypoint(P, ov, o, x, y, s, &hd))
        continue;
      const uint32_t slot = atomicAdd(&cnt->n_cand[o], 1u);
      if (slot < P.cand_cap)
      {
        Candidate cd;
        cd.key = ((unsigned long long)s << 40) | ((unsigned long long)y << 20) | (unsigned long long)x;
        cd.head = hd;
        cd.pad_ = 0;
        cand[(size_t)o * P.cand_cap + slot] = cd;
      }
    }
  }
}

cudaError_t launch_extrema(const DetectParams &P, Candidate *cand, DetectCounters *cnt, cudaStream_t st)
{
  int tiles = 0;
  for (int o = 0; o < P.n_oct; o++)
    tiles += ((P.oct[o].w + EX_TW - 1) / EX_TW) * ((P.oct[o].h + EX_TH - 1) / EX_TH);
  if (tiles == 0)
    return cudaSuccess;
  const size_t smem = sizeof(float) * (size_t)(P.ns + 2) * EX_SH * EX_SW;
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev])
  {
    cudaError_t e = cudaFuncSetAttribute(extrema_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * (VKS_MAX_LAYERS - 1) * EX_SH * EX_SW));
    if (e != cudaSuccess)
      return e;
    attr_done[dev] = true;
  }
  extrema_kernel<<<tiles, 256, smem, st>>>(P, cand, cnt);
  return cudaGetLastError();
}

/* Rank pass: candidate i of octave o goes to slot #{j : key_j < key_i}; slots
 * beyond the section capacity are dropped (ExtractKeypoints.comp:208-211 keeps
 * counting past max_nb_feat, so does n_cand). */
#define ORD_THREADS 256
__global__ void __launch_bounds__(ORD_THREADS) order_primaries_kernel(const __grid_constant__ DetectParams P, const Candidate *__restrict__ cand,
                                                                      DetectCounters *__restrict__ cnt, FeatHead *__restrict__ prim)
{
  __shared__ unsigned long long s_keys[ORD_THREADS];
  const int o = blockIdx.y;
  const uint32_t n_all = cnt->n_cand[o];
  const uint32_t n = min(n_all, P.cand_cap);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    cnt->n_prim[o] = min(n, P.cap[o]);
  if (blockIdx.x * ORD_THREADS >= n)
    return;
  const Candidate *__restrict__ lst = cand + (size_t)o * P.cand_cap;
  const uint32_t i = blockIdx.x * ORD_THREADS + threadIdx.x;
  const unsigned long long my = (i < n) ? lst[i].key : ~0ull;
  uint32_t rank = 0;
  for (uint32_t base = 0; base < n; base += ORD_THREADS)
  {
    const uint32_t j = base + threadIdx.x;
    s_keys[threadIdx.x] = (j < n) ? lst[j].key : ~0ull;
    __syncthreads();
    const uint32_t m = min((uint32_t)ORD_THREADS, n - base);
    for (uint32_t k = 0; k < m; k++)
      rank += (s_keys[k] < my) ? 1u : 0u;
    __syncthreads();
  }
  if (i < n && rank < P.cap[o])
    prim[P.sec_off[o] + rank] = lst[i].head;
}

cudaError_t launch_order_primaries(const DetectParams &P, const Candidate *cand, DetectCounters *cnt, FeatHead *prim, cudaStream_t st)
{
  if (P.n_oct == 0)
    return cudaSuccess;
  /* worst case grid; CTAs beyond the candidate count exit immediately */
  uint32_t max_cap = 0;
  for (int o = 0; o < P.n_oct; o++)
    max_cap = P.cap[o] > max_cap ? P.cap[o] : max_cap;
  uint32_t bx = (P.cand_cap + ORD_THREADS - 1) / ORD_THREADS;
  dim3 grid(bx, P.n_oct, 1);
  order_primaries_kernel<<<grid, ORD_THREADS, 0, st>>>(P, cand, cnt, prim);
  return cudaGetLastError();
}

} // namespace vks
