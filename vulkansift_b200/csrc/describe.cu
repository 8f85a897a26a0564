/*
 * describe.cu -- orientation assignment, feature assembly and 4x4x8 descriptors.
 *
 *   orientation_kernel  <- shaders/ComputeOrientation.comp:52-186 (one warp per keypoint)
 *   assemble_kernel     <- the atomicAdd appends of ComputeOrientation.comp:170-183 made
 *                          deterministic: primaries first, extra orientations after them in
 *                          (parent, bin) order, sections clamped like
 *                          sift_memory.c:1083-1095, output packed in octave order like
 *                          sift_memory.c:1128-1146
 *   descriptor_kernel   <- shaders/ComputeDescriptors.comp:84-274 (one CTA per feature)
 *
 * Histograms are accumulated in uint32 fixed point exactly like the shaders,
 * so the result does not depend on the order of the shared-memory atomics.
 */
#include "vksift_internal.h"

#include "layer_io.cuh"

namespace vks
{

/* imageLoad on the Gaussian array, out of bounds -> 0 (SURVEY B-D3) */
__device__ __forceinline__ float gauss_at(const void *__restrict__ L, int fp16, int w, int h, int pitch, int x, int y)
{
  if (x < 0 || x >= w || y < 0 || y >= h)
    return 0.f;
  return layer_ld(L, (size_t)y * pitch + x, fp16);
}

/* ---- orientation --------------------------------------------------------- */
#define ORI_THREADS 64
#define ORI_TERMS 2048 /* terms of the fixed-point scale staged per chunk */

/* One CTA per keypoint.  The reference runs one 32-thread work group per keypoint in which EVERY thread
 * recomputes the (2r+1)^2-term fixed-point scale (:75-81); here 128 threads evaluate the terms once into
 * shared memory, one thread adds them in the reference's (i outer, j inner) fp32 order, and all four warps
 * share the pixel loop. */
template <bool H16>
__global__ void __launch_bounds__(ORI_THREADS) orientation_kernel(const __grid_constant__ DetectParams P, const DetectCounters *__restrict__ cnt,
                                                                  const FeatHead *__restrict__ prim, float *__restrict__ ori,
                                                                  uint32_t *__restrict__ n_ori)
{
  __shared__ uint32_t hist[36];
  __shared__ uint32_t tmp[36];
  __shared__ float s_terms[ORI_TERMS];
  __shared__ float s_m;
  __shared__ float s_part[ORI_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31;

  uint32_t total = 0;
  for (int o = P.ob; o < P.oe; o++)
    total += cnt->n_prim[o];
  for (uint32_t item = blockIdx.x; item < total; item += gridDim.x)
  {
    /* item -> (octave, index in section) */
    int o = P.ob;
    uint32_t idx = item;
    while (idx >= cnt->n_prim[o])
    {
      idx -= cnt->n_prim[o];
      o++;
    }
    const uint32_t slot = P.sec_off[o] + idx;
    const FeatHead kp = prim[slot];
    const OctaveView &ov = P.oct[o];
    const void *__restrict__ L = layer_ptr(ov.G, (size_t)kp.scale_idx * ov.layer_stride, ov.fp16);
    constexpr int f16 = H16 ? 1 : 0;

    const float sf = vks_pow2i(kp.octave_idx);
    const float lambda = 1.5f * (kp.sigma / sf);
    const int r = (int)floorf(3 * lambda);
    const float es = -1.f / (2.f * lambda * lambda);
    const int box = 2 * r + 1;
    const int n_terms = box * box;

    /* ComputeOrientation.comp:75-81: M = sum_{i,j} exp(es*(i*i+j*j))*sqrt(2) in (i outer, j inner) fp32 order, and only
     * ceil(log2(M)) is used (the fixed-point scale below).  The sequential sum is a chain of up to ~1400 dependent additions
     * on one thread (it was half of this kernel's critical path), so M is first summed in parallel: both sums are within
     * n * 2^-24 relative of the exact value (positive terms; n <= 1369 with the default configuration: 8e-5), hence they fall
     * between the same two powers of two unless the parallel sum lies within delta = 1e-4 + 4 n 2^-24 of one; only then (a few
     * keypoints in 1000) is the reference-order sum evaluated. */
    /* (row, column) of this thread's first cell of the box and the step of ORI_THREADS cells: no division in the two loops */
    const int step_r = ORI_THREADS / box, step_c = ORI_THREADS - step_r * box;
    const int row0 = tid / box, col0 = tid - row0 * box;
    float m;
    {
      float part = 0.f;
      int row = row0, col = col0;
      for (int q = tid; q < n_terms; q += ORI_THREADS)
      {
        const int i = row - r, j = col - r;
        part += vks_expf(es * (float)((i * i) + (j * j))) * VKS_SQRT2_F;
        col += step_c;
        row += step_r;
        if (col >= box)
        {
          col -= box;
          row++;
        }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1)
        part += __shfl_xor_sync(0xffffffffu, part, d);
      if (lane == 0)
        s_part[tid >> 5] = part;
      if (tid < 36)
        hist[tid] = 0;
      __syncthreads();
      m = 0.f;
#pragma unroll
      for (int k = 0; k < ORI_THREADS / 32; k++)
        m += s_part[k];
      const float frac = __uint_as_float((__float_as_uint(m) & 0x007fffffu) | 0x3f800000u); /* mantissa as a value in [1, 2) */
      const float delta = 1e-4f + (float)n_terms * 2.4e-7f;
      if (!(frac > 1.f + delta && frac < 2.f - 2.f * delta))
      {
        m = 0.f;
        for (int t0 = 0; t0 < n_terms; t0 += ORI_TERMS)
        {
          const int nt = min(ORI_TERMS, n_terms - t0);
          __syncthreads(); /* s_terms of the previous chunk (and s_m) consumed */
          for (int t = tid; t < nt; t += ORI_THREADS)
          {
            const int q = t0 + t;
            const int i = q / box - r, j = q % box - r;
            s_terms[t] = vks_expf(es * (float)((i * i) + (j * j))) * VKS_SQRT2_F;
          }
          __syncthreads();
          if (tid == 0)
          {
#pragma unroll 1
            for (int t = 0; t < nt; t++) /* rare path: rolled */
              m += s_terms[t];
            s_m = m;
          }
        }
        __syncthreads();
        m = s_m;
      }
    }
    const float fp = (float)(1u << (uint32_t)(30 - vks_ceil_log2(m)));

    const float rsx = vks_rint(kp.scale_x), rsy = vks_rint(kp.scale_y);
    const int cx = (int)rsx, cy = (int)rsy;
    const float r2 = (float)(r * r);
    int prow = row0, pcol = col0;
    for (int pix = tid; pix < n_terms; pix += ORI_THREADS)
    {
      const int dy = prow - r, dx = pcol - r;
      pcol += step_c;
      prow += step_r;
      if (pcol >= box)
      {
        pcol -= box;
        prow++;
      }
      const int gx = cx + dx, gy = cy + dy;
      const float sdx = (rsx + (float)dx) - kp.scale_x;
      const float sdy = (rsy + (float)dy) - kp.scale_y;
      const float d2 = (sdx * sdx) + (sdy * sdy);
      if ((gx < 1 || gx >= (ov.w - 1) || gy < 1 || gy >= (ov.h - 1)) && (d2 > r2))
        continue;
      const float gX = 0.5f * (gauss_at(L, f16, ov.w, ov.h, ov.pitch, gx + 1, gy) - gauss_at(L, f16, ov.w, ov.h, ov.pitch, gx - 1, gy));
      const float gY = 0.5f * (gauss_at(L, f16, ov.w, ov.h, ov.pitch, gx, gy + 1) - gauss_at(L, f16, ov.w, ov.h, ov.pitch, gx, gy - 1));
      const float mag = vks_expf(d2 * es) * vks_sqrt((gX * gX) + (gY * gY));
      float th = vks_atan2f(gY, gX);
      if (th < 0.f)
        th += VKS_TWO_PI_F;
      else if (th > VKS_TWO_PI_F)
        th -= VKS_TWO_PI_F;
      int bin = (int)((th * 36.f) / VKS_TWO_PI_F);
      if (bin < 0)
        bin += 36;
      else if (bin >= 36)
        bin -= 36;
      atomicAdd(&hist[bin], (uint32_t)(mag * fp));
    }
    __syncthreads();

    if (tid < 32)
    {
      /* :130-147 three double box smoothings in uint/float mixed arithmetic */
      for (int it = 0; it < 3; it++)
      {
        for (int i = lane; i < 36; i += 32)
          tmp[i] = (uint32_t)((float)(hist[(i + 35) % 36] + hist[i] + hist[(i + 1) % 36]) / 3.f);
        __syncwarp();
        for (int i = lane; i < 36; i += 32)
          hist[i] = (uint32_t)((float)(tmp[(i + 35) % 36] + tmp[i] + tmp[(i + 1) % 36]) / 3.f);
        __syncwarp();
      }
      uint32_t mx = hist[lane];
      if (lane < 4)
        mx = max(mx, hist[32 + lane]);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1)
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));

      /* :156-183 peaks in increasing bin order */
      uint32_t count = 0;
      float *my_ori = ori + (size_t)slot * P.ori_stride;
      for (int base = 0; base < 36; base += 32)
      {
        const int i = base + lane;
        bool peak = false;
        float theta = 0.f;
        if (i < 36)
        {
          const uint32_t hp = hist[(i + 35) % 36], hn = hist[(i + 1) % 36], hc = hist[i];
          if (((float)hc >= (0.8f * (float)mx)) && (hc > hp) && (hc > hn))
          {
            peak = true;
            /* uint32 differences wrap before the conversion (SURVEY B-D12) */
            const float num = (float)(uint32_t)(hp - hn);
            const float den = (float)(uint32_t)(hp - (2u * hc) + hn);
            const float fi = (float)i + 0.5f * (num / den);
            theta = ((fi + 0.5f) * VKS_TWO_PI_F) / 36.f;
          }
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, peak);
        if (peak)
        {
          const uint32_t k = count + __popc(mask & ((1u << lane) - 1u));
          if (k < P.ori_stride)
            my_ori[k] = theta;
        }
        count += __popc(mask);
      }
      if (lane == 0)
      {
        if (count == 0)
          my_ori[0] = 0.f; /* no peak: orientation stays 0 and the keypoint is still described */
        n_ori[slot] = count;
      }
    }
    __syncthreads();
  }
}

/* ctas_per_sm: the grid is persistent (a CTA walks the keypoints with the grid's stride).  A keypoint is a latency chain that leaves its
 * CTA's slots mostly idle, so MORE resident CTAs take issue and thread slots from the kernels of the other detections in flight:
 * measured with 8 lanes 0.3015 / 0.2968 / 0.2947 / 0.2934 / 0.2920 ms per image for 32 / 16 / 8 / 4 / 2 CTAs of 64 threads per SM; a
 * detection alone is fastest with 3-4 (0.450 ms against 0.470 with 2). */
cudaError_t launch_orientation(const DetectParams &P, DetectCounters *cnt, const FeatHead *prim, float *ori, uint32_t *n_ori, cudaStream_t st,
                               int ctas_per_sm)
{
  static int env_per_sm = -1;
  if (env_per_sm < 0)
  {
    const char *e = getenv("VKSIFT_ORI_CTAS");
    env_per_sm = e ? atoi(e) : 0;
    if (env_per_sm < 0 || env_per_sm > 32)
      env_per_sm = 0;
  }
  const int per_sm = env_per_sm ? env_per_sm : (ctas_per_sm >= 1 && ctas_per_sm <= 32 ? ctas_per_sm : 4);
  if (P.oct[P.ob].fp16)
    orientation_kernel<true><<<148 * per_sm, ORI_THREADS, 0, st>>>(P, cnt, prim, ori, n_ori);
  else
    orientation_kernel<false><<<148 * per_sm, ORI_THREADS, 0, st>>>(P, cnt, prim, ori, n_ori);
  return cudaGetLastError();
}

/* ---- assembly ------------------------------------------------------------ */
#define ASM_THREADS 1024
__device__ __forceinline__ uint32_t extra_orientations(uint32_t n, uint32_t max_ori)
{
  /* peak k >= 1 is appended iff max_ori == 0 || k < max_ori (ComputeOrientation.comp:175) */
  if (n == 0)
    return 0;
  const uint32_t lim = (max_ori == 0) ? n : min(n, max_ori);
  return lim - 1;
}

__global__ void __launch_bounds__(ASM_THREADS) assemble_kernel(const __grid_constant__ DetectParams P, DetectCounters *__restrict__ cnt,
                                                               const uint32_t *__restrict__ n_ori, uint32_t *__restrict__ feat_src,
                                                               uint32_t *__restrict__ host_counts)
{
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
  uint32_t out_off = 0;
  for (int o = 0; o < P.n_oct; o++)
  {
    const uint32_t np = cnt->n_prim[o];
    const uint32_t cap = P.cap[o];
    const uint32_t sec = P.sec_off[o];
    if (tid == 0)
      s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < np; base += ASM_THREADS)
    {
      const uint32_t i = base + tid;
      const uint32_t e = (i < np) ? extra_orientations(n_ori[sec + i], P.max_ori) : 0u;
      /* block exclusive scan of e */
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
      const uint32_t excl = carry + (wi ? s_warp[wi - 1] : 0u) + (v - e);
      if (i < np)
      {
        /* primary slot i, extras at np + excl .. */
        if (i < cap)
          feat_src[out_off + i] = (i << 6);
        for (uint32_t k = 0; k < e; k++)
        {
          const uint32_t slot = np + excl + k;
          if (slot < cap)
            feat_src[out_off + slot] = (i << 6) | (k + 1);
        }
      }
      __syncthreads();
      if (tid == ASM_THREADS - 1)
        s_carry = carry + s_warp[31];
      __syncthreads();
    }
    const uint32_t found = cnt->n_cand[o] + s_carry; /* counter semantics of nb_elem */
    const uint32_t total = np + s_carry;
    const uint32_t kept = min(total, cap);
    if (tid == 0)
    {
      cnt->n_found[o] = found;
      cnt->n_kept[o] = kept;
      cnt->out_off[o] = out_off;
      host_counts[1 + o] = found;
      host_counts[1 + VKS_MAX_OCT + o] = kept;
    }
    out_off += kept;
    __syncthreads();
  }
  if (tid == 0)
  {
    cnt->n_total = out_off;
    host_counts[0] = out_off;
  }
}

cudaError_t launch_assemble(const DetectParams &P, DetectCounters *cnt, const uint32_t *n_ori, uint32_t *feat_src, uint32_t *host_counts,
                            cudaStream_t st)
{
  assemble_kernel<<<1, ASM_THREADS, 0, st>>>(P, cnt, n_ori, feat_src, host_counts);
  return cudaGetLastError();
}

/* ---- descriptor ---------------------------------------------------------- */
#define DESC_THREADS 128
#define DESC_MAXBOX 160 /* window rows the row-interval tables hold (2R+1 <= 77 for the default configuration) */

/* m_table[hr] = M(hr) of ComputeDescriptors.comp:116-124 for hr = R/2 < VKS_DESC_M_TABLE:
 *   for i<hr { m += e(i,i)*sqrt2; for j in (i, hr) m += e(i,j)*sqrt2*2 },  e(i,j) = exp(-0.125*(i*i+j*j))
 * in the shader's sequential fp32 order.  The sum only depends on the window radius, so it is tabulated once per instance (on the
 * host, api.cu) instead of being rebuilt (2*hr block barriers and a serial sum) for every feature. */
template <bool H16>
__global__ void __launch_bounds__(DESC_THREADS) descriptor_kernel(const __grid_constant__ DetectParams P, DetectCounters *__restrict__ cnt,
                                                                  const float *__restrict__ m_table, const FeatHead *__restrict__ prim,
                                                                  const float *__restrict__ ori,
                                                                  const uint32_t *__restrict__ feat_src, FeatHead *__restrict__ out_heads,
                                                                  uint8_t *__restrict__ out_desc)
{
  __shared__ uint32_t s_desc[128];
  __shared__ uint32_t s_acc;
  __shared__ float s_terms[DESC_THREADS];
  __shared__ float s_m;
  __shared__ uint32_t s_g;
  __shared__ int s_row_lo[DESC_MAXBOX];
  __shared__ int s_row_start[DESC_MAXBOX + 1];
  const int tid = threadIdx.x;
  const uint32_t total = cnt->n_total;
  for (;;)
  {
    /* features cost between 1x and 4x (window area): hand them out one at a time instead of a fixed stride */
    if (tid == 0)
      s_g = atomicAdd(&cnt->desc_next, 1u);
    __syncthreads();
    const uint32_t g = s_g;
    if (g >= total)
      break;
    int o = 0;
    while (o + 1 < P.n_oct && g >= cnt->out_off[o + 1])
      o++;
    const uint32_t src = feat_src[g];
    const uint32_t pi = P.sec_off[o] + (src >> 6), k = src & 63u;
    FeatHead kp = prim[pi];
    kp.orientation = ori[(size_t)pi * P.ori_stride + k];
    const OctaveView &ov = P.oct[o];
    const void *__restrict__ L = layer_ptr(ov.G, (size_t)kp.scale_idx * ov.layer_stride, ov.fp16);
    constexpr int f16 = H16 ? 1 : 0; /* compile-time: the four loads of a sample carry no format branch */

    s_desc[tid] = 0;
    const float sf = vks_pow2i(kp.octave_idx);
    const float lambda = 3.0f * (kp.sigma / sf);
    const float radius = VKS_SQRT2_F * lambda * 5.f * 0.5f;
    const int R = (int)floorf(radius + 0.5f);
    float sn, cs;
    vks_sincosf(kp.orientation, &sn, &cs);
    const float kc = cs / lambda, ks = sn / lambda;
    const float es = -0.125f;

    /* ComputeDescriptors.comp:116-124: fixed-point scale from the sequential sum
     *   for i<R/2 { m += e(i,i)*sqrt2; for j in (i, R/2) m += e(i,j)*sqrt2*2 }
     * terms evaluated in parallel per row i, summed in reference order by thread 0. */
    const int hr = R / 2;
    float m = 0.f;
    if (hr < VKS_DESC_M_TABLE)
      m = __ldg(m_table + hr);
    else
    for (int i = 0; i < hr; i++)
    {
      for (int j0 = i; j0 < hr; j0 += DESC_THREADS)
      {
        const int j = j0 + tid;
        if (j < hr)
        {
          float t = vks_expf(es * (float)((i * i) + (j * j))) * VKS_SQRT2_F;
          if (j > i)
            t = t * 2.f;
          s_terms[tid] = t;
        }
        __syncthreads();
        if (tid == 0)
        {
          const int c = min(DESC_THREADS, hr - j0);
#pragma unroll 1
          for (int q = 0; q < c; q++) /* rare (windows beyond the host table): rolled, it must not take instruction-cache space */
            m += s_terms[q];
        }
        __syncthreads();
      }
    }
    if (hr >= VKS_DESC_M_TABLE)
    {
      if (tid == 0)
        s_m = m;
      __syncthreads();
      m = s_m;
    }
    __syncthreads(); /* s_desc cleared */
    const float fp = (float)(1u << (uint32_t)(16 - vks_ceil_log2(m)));

    const float rsx = vks_rint(kp.scale_x), rsy = vks_rint(kp.scale_y);
    const int cx = (int)rsx, cy = (int)rsy;
    const int box = 2 * R + 1;
    const uint32_t s_desc_addr = (uint32_t)__cvta_generic_to_shared(s_desc);
    const int pitch = ov.pitch;
    auto add_pixel = [&](int dx, int dy) {
      const int ix = cx + dx, iy = cy + dy;
      if (ix < 1 || ix >= (ov.w - 1) || iy < 1 || iy >= (ov.h - 1))
        return;
      const float sdx = (rsx + (float)dx) - kp.scale_x;
      const float sdy = (rsy + (float)dy) - kp.scale_y;
      const float ox = kc * sdx + ks * sdy;
      const float oy = kc * sdy - ks * sdx;
      /* pixels whose both spatial cells fall outside the 4x4 grid contribute to no bin (:187):
       * skip them before the transcendental work (about half of the box after rotation) */
      const float fx = ox + 2.f, fy = oy + 2.f;
      const int hx = (int)floorf(fx - 0.5f), hy = (int)floorf(fy - 0.5f);
      if (hx < -1 || hx > 3 || hy < -1 || hy > 3)
        return;
      const int off = iy * pitch + ix; /* 32-bit index: a layer has fewer than 2^31 cells */
      const float gX = 0.5f * (layer_ld(L, (size_t)(off + 1), f16) - layer_ld(L, (size_t)(off - 1), f16));
      const float gY = 0.5f * (layer_ld(L, (size_t)(off + pitch), f16) - layer_ld(L, (size_t)(off - pitch), f16));
      const float mag = vks_expf(es * ((ox * ox) + (oy * oy))) * vks_sqrt((gX * gX) + (gY * gY));
      /* Every contribution is (uint32)(((wx*wy)*wb)*mag*fp) with weights <= 1 (up to an ulp) and fp a power of two: below
       * mag*fp = 0.99 all eight truncate to zero, so the sample adds nothing and its angle is not needed (flat or weakly
       * textured pixels under the tail of the Gaussian window). */
      if (mag * fp < 0.99f)
        return;
      float th = vks_atan2f(gY, gX);
      if (th < 0.f)
        th += VKS_TWO_PI_F;
      else if (th > VKS_TWO_PI_F)
        th -= VKS_TWO_PI_F;
      th = th - kp.orientation;
      if (th < 0.f)
        th += VKS_TWO_PI_F;
      else if (th > VKS_TWO_PI_F)
        th -= VKS_TWO_PI_F;
      const float fb = P.vlfeat ? ((th * 8.f) / VKS_TWO_PI_F) : ((-th * 8.f) / VKS_TWO_PI_F);
      const int hb = (int)floorf(fb);
      const float rx = fx - ((float)hx + 0.5f), ry = fy - ((float)hy + 0.5f), rb = fb - (float)hb;
      /* trilinear weights |1 - i - r| for i = 0, 1 (the second one is |0 - r| = |r| exactly); the eight contributions are
       * computed unconditionally in the shader's order ((wx*wy)*wb)*mag, then *fp, and added by predicated shared-memory
       * reductions: integer adds commute, so the histogram does not depend on the order of the lanes */
      const float wx[2] = {fabsf(1.f - rx), fabsf(rx)}, wy[2] = {fabsf(1.f - ry), fabsf(ry)}, wb[2] = {fabsf(1.f - rb), fabsf(rb)};
      const uint32_t b0 = (uint32_t)hb & 7u, b1 = (uint32_t)(hb + 1) & 7u; /* non-negative modulo (SURVEY B-D7) */
      /* addresses of the two orientation bins in cell (hx, hy); the four cells of a sample are immediate offsets from them */
      const uint32_t a0 = s_desc_addr + (uint32_t)((hy * 32 + hx * 8) * 4) + b0 * 4u;
      const uint32_t a1 = s_desc_addr + (uint32_t)((hy * 32 + hx * 8) * 4) + b1 * 4u;
#define DESC_CELL(J, I)                                                                                                                              \
  {                                                                                                                                                  \
    const uint32_t ok = ((uint32_t)((I) + hx) < 4u && (uint32_t)((J) + hy) < 4u) ? 1u : 0u;                                                           \
    const float wxy = wx[I] * wy[J];                                                                                                                 \
    const uint32_t v0 = (uint32_t)(((wxy * wb[0]) * mag) * fp), v1 = (uint32_t)(((wxy * wb[1]) * mag) * fp);                                        \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p red.shared.add.u32 [%0+%5], %1;\n\t@p red.shared.add.u32 [%2+%5], %4;\n\t}" ::"r"(a0), \
                 "r"(v0), "r"(a1), "r"(ok), "r"(v1), "n"(((J) * 32 + (I) * 8) * 4)                                                                  \
                 : "memory");                                                                                                                        \
  }
      DESC_CELL(0, 0)
      DESC_CELL(0, 1)
      DESC_CELL(1, 0)
      DESC_CELL(1, 1)
#undef DESC_CELL
    };
    if (box <= DESC_MAXBOX)
    {
      /* The rotated 4x4 grid covers exactly half of the (2R+1)^2 box.  Per row, the pixels that can pass the grid test
       * form an interval: compute a conservative one (the exact test stays in add_pixel, so the result is unchanged),
       * and let the threads walk the concatenated intervals, so that a warp's 32 lanes are (almost) all useful. */
      for (int row = tid; row < box; row += DESC_THREADS)
      {
        const int dy = row - R, iy = cy + dy;
        int lo = max(-R, 1 - cx), hi = min(R, ov.w - 2 - cx);
        if (iy < 1 || iy >= ov.h - 1)
          hi = lo - 1;
        else
        {
          const float sdy = (rsy + (float)dy) - kp.scale_y;
          const float T = 2.5f + 0.02f;
          float lo_f = -(float)(R + 2), hi_f = (float)(R + 2);
          if (fabsf(kc) >= 1e-4f)
          {
            const float a = (-T - ks * sdy) / kc, b = (T - ks * sdy) / kc;
            lo_f = fmaxf(lo_f, fminf(a, b));
            hi_f = fminf(hi_f, fmaxf(a, b));
          }
          if (fabsf(ks) >= 1e-4f)
          {
            const float a = (kc * sdy - T) / ks, b = (kc * sdy + T) / ks;
            lo_f = fmaxf(lo_f, fminf(a, b));
            hi_f = fminf(hi_f, fmaxf(a, b));
          }
          const float fx = rsx - kp.scale_x; /* sdx = dx + fx */
          if (lo_f <= hi_f)
          {
            lo = max(lo, (int)floorf(lo_f - fx) - 1);
            hi = min(hi, (int)ceilf(hi_f - fx) + 1);
          }
          else
            hi = lo - 1;
        }
        s_row_lo[row] = lo;
        s_row_start[row + 1] = max(0, hi - lo + 1);
      }
      if (tid == 0)
        s_row_start[0] = 0;
      __syncthreads();
      /* exclusive prefix of the row lengths (box <= 160 rows: one warp, five shuffle steps per 32 rows) */
      if (tid < 32)
      {
        int carry = 0;
        for (int base = 0; base < box; base += 32)
        {
          const int i = base + tid;
          int v = (i < box) ? s_row_start[i + 1] : 0;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1)
          {
            const int n = __shfl_up_sync(0xffffffffu, v, d);
            if (tid >= d)
              v += n;
          }
          if (i < box)
            s_row_start[i + 1] = carry + v;
          carry += __shfl_sync(0xffffffffu, v, 31);
        }
      }
      __syncthreads();
      const int n_px = s_row_start[box];
      /* Each warp takes one contiguous quarter of the concatenated intervals and walks it 32 pixels at a time, so a lane
       * moves on by about one row per step (the search for its row is a short loop; with a stride of the whole CTA it cost
       * five iterations per pixel).  First row of the warp by bisection. */
      const int per_warp = (((n_px + (DESC_THREADS / 32) - 1) / (DESC_THREADS / 32)) + 31) & ~31;
      const int w_begin = (tid >> 5) * per_warp, w_end = min(n_px, w_begin + per_warp);
      int row = 0;
      {
        int lo = 0, hi = box; /* largest row with s_row_start[row] <= w_begin */
        while (hi - lo > 1)
        {
          const int mid = (lo + hi) >> 1;
          if (s_row_start[mid] <= w_begin)
            lo = mid;
          else
            hi = mid;
        }
        row = lo;
      }
      int row_begin = s_row_start[row], row_end = s_row_start[row + 1];
      for (int idx = w_begin + (tid & 31); idx < w_end; idx += 32)
      {
        while (idx >= row_end)
        {
          row++;
          row_begin = row_end;
          row_end = s_row_start[row + 1];
        }
        add_pixel(s_row_lo[row] + (idx - row_begin), row - R);
      }
    }
    else
    {
#pragma unroll 1
      for (int pix = tid; pix < box * box; pix += DESC_THREADS)
        add_pixel(pix % box - R, pix / box - R);
    }
    if (tid == 0)
      s_acc = 0;
    __syncthreads();

    /* :209-265 norm, clamp at 0.2, renorm, *512, truncate to u8 */
    uint32_t v = s_desc[tid];
    atomicAdd(&s_acc, v * v);
    __syncthreads();
    float norm = vks_sqrt((float)s_acc);
    __syncthreads();
    v = min(v, (uint32_t)(norm * 0.2f));
    if (tid == 0)
      s_acc = 0;
    __syncthreads();
    atomicAdd(&s_acc, v * v);
    __syncthreads();
    norm = vks_sqrt((float)s_acc);
    const float d = (float)v * (512.f / norm);
    uint32_t q8;
    if (!(d == d))
      q8 = 0;
    else if (d > 255.f)
      q8 = 255;
    else
      q8 = (uint32_t)d;
    out_desc[(size_t)g * 128 + tid] = (uint8_t)q8;
    if (tid == 0)
      out_heads[g] = kp;
    __syncthreads();
  }
}

cudaError_t launch_descriptors(const DetectParams &P, DetectCounters *cnt, const float *m_table, const FeatHead *prim, const float *ori,
                               const uint32_t *feat_src, FeatHead *out_heads, uint8_t *out_desc, cudaStream_t st)
{
  static int per_sm = 0;
  if (per_sm == 0)
  {
    const char *e = getenv("VKSIFT_DESC_CTAS");
    per_sm = e ? atoi(e) : 12; /* 44 registers x 128 threads: up to 11 resident; 12 measured 5 us faster than 8 for 3.4 k features */
    if (per_sm < 1 || per_sm > 16)
      per_sm = 12;
  }
  if (P.oct[0].fp16)
    descriptor_kernel<true><<<148 * per_sm, DESC_THREADS, 0, st>>>(P, cnt, m_table, prim, ori, feat_src, out_heads, out_desc);
  else
    descriptor_kernel<false><<<148 * per_sm, DESC_THREADS, 0, st>>>(P, cnt, m_table, prim, ori, feat_src, out_heads, out_desc);
  return cudaGetLastError();
}

/* ---- AoS <-> SoA for host transfers (vksift_Feature = 36 B head + 128 B descriptor) */
__global__ void pack_aos_kernel(const FeatHead *__restrict__ heads, const uint8_t *__restrict__ desc, uint32_t n, uint32_t *__restrict__ aos)
{
  /* 41 words per feature: 9 head + 32 descriptor */
  const size_t total = (size_t)n * 41;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
  {
    const uint32_t f = (uint32_t)(i / 41), wv = (uint32_t)(i % 41);
    aos[i] = (wv < 9) ? ((const uint32_t *)heads)[(size_t)f * 9 + wv] : ((const uint32_t *)desc)[(size_t)f * 32 + (wv - 9)];
  }
}
__global__ void unpack_aos_kernel(const uint32_t *__restrict__ aos, uint32_t n, FeatHead *__restrict__ heads, uint8_t *__restrict__ desc)
{
  const size_t total = (size_t)n * 41;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
  {
    const uint32_t f = (uint32_t)(i / 41), wv = (uint32_t)(i % 41);
    if (wv < 9)
      ((uint32_t *)heads)[(size_t)f * 9 + wv] = aos[i];
    else
      ((uint32_t *)desc)[(size_t)f * 32 + (wv - 9)] = aos[i];
  }
}
cudaError_t launch_pack_aos(const FeatHead *heads, const uint8_t *desc, uint32_t n, uint8_t *aos, cudaStream_t st)
{
  if (n == 0)
    return cudaSuccess;
  const int blocks = (int)min((size_t)148 * 8, ((size_t)n * 41 + 255) / 256);
  pack_aos_kernel<<<blocks, 256, 0, st>>>(heads, desc, n, (uint32_t *)aos);
  return cudaGetLastError();
}
cudaError_t launch_unpack_aos(const uint8_t *aos, uint32_t n, FeatHead *heads, uint8_t *desc, cudaStream_t st)
{
  if (n == 0)
    return cudaSuccess;
  const int blocks = (int)min((size_t)148 * 8, ((size_t)n * 41 + 255) / 256);
  unpack_aos_kernel<<<blocks, 256, 0, st>>>((const uint32_t *)aos, n, heads, desc);
  return cudaGetLastError();
}

} // namespace vks
