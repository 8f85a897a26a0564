// Cost of running straight-line code for the first time on an SM (B200): one CTA executes N independent-chain FFMAs once.
// "warm" = the same kernel launched again right away, "cold" = after another large kernel and a 512 MB memset evicted it
// from the instruction caches and L2.  Design input for pyramid.cu (unrolled kernels on one-tile grids).
#include <cstdio>
#include <cuda_runtime.h>
template <int N> __global__ void k(float *out, float a)
{
  float x0 = threadIdx.x, x1 = 1.f, x2 = 2.f, x3 = 3.f, x4 = 4.f, x5 = 5.f, x6 = 6.f, x7 = 7.f;
#pragma unroll
  for (int i = 0; i < N / 8; i++)
  {
    asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x0) : "f"(a));
    asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x1) : "f"(a));
    asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x2) : "f"(a));
    asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x3) : "f"(a));
    asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x4) : "f"(a));
    asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x5) : "f"(a));
    asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x6) : "f"(a));
    asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x7) : "f"(a));
  }
  out[threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
template <int N> void run(float *d, char *big, int threads)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float cold = 0, warm = 0;
  for (int rep = 0; rep < 5; rep++)
  {
    cudaMemset(big, rep, 512u << 20);   // evict L2 (and with it the code)
    k<65536><<<148 * 4, 128>>>(d + 4096, 1.0001f); // another large kernel through the instruction caches of every SM
    cudaDeviceSynchronize();
    cudaEventRecord(e0); k<N><<<1, threads>>>(d, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) cold += ms;
    cudaEventRecord(e0); k<N><<<1, threads>>>(d, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); if (rep) warm += ms;
  }
  printf("%6d instructions (%4d KB), %4d threads: cold %7.2f us  warm %7.2f us  -> %.2f us per KB\n", N, N * 16 / 1024, threads, cold * 250, warm * 250,
         (cold - warm) * 250 / (N * 16 / 1024.0));
}
int main()
{
  float *d; cudaMalloc(&d, 1 << 20);
  char *big; cudaMalloc(&big, 512u << 20);
  run<256>(d, big, 256); run<1024>(d, big, 256); run<4096>(d, big, 256); run<16384>(d, big, 256);
  run<1024>(d, big, 32); run<4096>(d, big, 32); run<4096>(d, big, 1024);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
