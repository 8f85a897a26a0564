// Microbenchmark: issue rate of scalar vs packed fp32 on sm_100a (design input for pyramid.cu).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp32_rate fp32_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pk_fma(pk2 a, pk2 b, pk2 c){ pk2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ pk2 pk_add(pk2 a, pk2 b){ pk2 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float s_fma(float a, float b, float c){ float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float s_add(float a, float b){ float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
#define N_ACC 8
#define ITERS 4096
template <int MODE> __global__ void __launch_bounds__(256) k(float *out, float seed)
{
  float k1 = seed, k2 = seed * 0.5f;
  if (MODE == 0) { // scalar FFMA chains
    float a[N_ACC];
    for (int i = 0; i < N_ACC; i++) a[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
      for (int i = 0; i < N_ACC; i++) a[i] = s_fma(a[i], k1, k2);
    float s = 0; for (int i = 0; i < N_ACC; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else if (MODE == 1) { // packed FFMA2 chains
    pk2 a[N_ACC]; pk2 kk1 = ((pk2)__float_as_uint(k1) << 32) | __float_as_uint(k1), kk2 = ((pk2)__float_as_uint(k2) << 32) | __float_as_uint(k2);
    for (int i = 0; i < N_ACC; i++) a[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
      for (int i = 0; i < N_ACC; i++) a[i] = pk_fma(a[i], kk1, kk2);
    pk2 s = 0; for (int i = 0; i < N_ACC; i++) s ^= a[i];
    ((pk2 *)out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else if (MODE == 2) { // scalar add+fma pairs (blur inner loop shape)
    float a[N_ACC], w[N_ACC];
    for (int i = 0; i < N_ACC; i++) { a[i] = threadIdx.x + i; w[i] = i * seed; }
    for (int it = 0; it < ITERS; it++)
#pragma unroll
      for (int i = 0; i < N_ACC; i++) a[i] = s_fma(s_add(w[i], w[(i + 1) % N_ACC]), k1, a[i]);
    float s = 0; for (int i = 0; i < N_ACC; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else if (MODE == 3) { // packed add2+fma2 pairs
    pk2 a[N_ACC], w[N_ACC]; pk2 kk1 = ((pk2)__float_as_uint(k1) << 32) | __float_as_uint(k1);
    for (int i = 0; i < N_ACC; i++) { a[i] = threadIdx.x + i; w[i] = (pk2)__float_as_uint(i * seed) * 0x100000001ull; }
    for (int it = 0; it < ITERS; it++)
#pragma unroll
      for (int i = 0; i < N_ACC; i++) a[i] = pk_fma(pk_add(w[i], w[(i + 1) % N_ACC]), kk1, a[i]);
    pk2 s = 0; for (int i = 0; i < N_ACC; i++) s ^= a[i];
    ((pk2 *)out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else if (MODE == 4) { // scalar FADD chains only
    float a[N_ACC];
    for (int i = 0; i < N_ACC; i++) a[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
      for (int i = 0; i < N_ACC; i++) a[i] = s_add(a[i], k1);
    float s = 0; for (int i = 0; i < N_ACC; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else if (MODE == 5) { // packed FADD2 chains only
    pk2 a[N_ACC]; pk2 kk1 = ((pk2)__float_as_uint(k1) << 32) | __float_as_uint(k1);
    for (int i = 0; i < N_ACC; i++) a[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
      for (int i = 0; i < N_ACC; i++) a[i] = pk_add(a[i], kk1);
    pk2 s = 0; for (int i = 0; i < N_ACC; i++) s ^= a[i];
    ((pk2 *)out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}
template <int MODE> void run(const char *name, int results_per_instr, float *d)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int grid = 148 * 8;
  k<MODE><<<grid, 256>>>(d, 1.0001f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(d, 1.0001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int per_it = (MODE == 2 || MODE == 3) ? 2 * N_ACC : N_ACC;
  double warp_instr = (double)grid * 8 * ITERS * per_it;
  double instr_per_s = warp_instr / (ms * 1e-3);
  printf("%-28s %8.3f ms  %7.1f G warp-instr/s  = %5.2f warp-instr/clk/SM @1.965GHz, %6.1f G fp32 results/s/1e3\n", name, ms, instr_per_s / 1e9,
         instr_per_s / 148 / 1.965e9, instr_per_s * 32 * results_per_instr / 1e12);
}
int main()
{
  float *d; cudaMalloc(&d, 148 * 8 * 256 * 16);
  run<0>("scalar FFMA", 1, d);
  run<1>("packed FFMA2", 2, d);
  run<2>("scalar FADD+FFMA", 1, d);
  run<3>("packed FADD2+FFMA2", 2, d);
  run<4>("scalar FADD", 1, d);
  run<5>("packed FADD2", 2, d);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
