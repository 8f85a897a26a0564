// Per-launch floor of a dependent kernel chain on one stream (B200): empty kernels, with/without big dynamic smem, with/without PDL.
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
__global__ void k_empty(float *p) { if (p && threadIdx.x == 9999) p[0] = 1.f; }
__global__ void k_pdl(float *p)
{
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p && threadIdx.x == 9999) p[0] = 1.f;
}
// a little dependent work: each launch reads what the previous wrote
__global__ void k_work(float *p, int n)
{
  extern __shared__ float sm[];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += p[i];
  sm[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) p[blockIdx.x] = sm[0] + sm[1];
}
template <typename F> float time_chain(F launch, int n, cudaStream_t st)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 20; i++) launch();
  cudaStreamSynchronize(st);
  cudaEventRecord(e0, st);
  for (int i = 0; i < n; i++) launch();
  cudaEventRecord(e1, st); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1000.f / n;
}
static void launch_ex(void (*k)(float *), int grid, int block, size_t smem, cudaStream_t st, float *p, bool pdl)
{
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; a[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = a; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, k, p);
}
static void launch_work(int grid, int block, size_t smem, cudaStream_t st, float *p, int n, bool pdl)
{
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; a[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = a; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, k_work, p, n);
}
int main()
{
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  float *d; cudaMalloc(&d, 1 << 20); cudaMemset(d, 0, 1 << 20);
  cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_pdl, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_work, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int N = 200;
  printf("empty 1x256            : %.2f us/launch\n", time_chain([&] { launch_ex(k_empty, 1, 256, 0, st, d, false); }, N, st));
  printf("empty 4x512 80KB smem  : %.2f us/launch\n", time_chain([&] { launch_ex(k_empty, 4, 512, 80 * 1024, st, d, false); }, N, st));
  printf("empty 148x512 80KB     : %.2f us/launch\n", time_chain([&] { launch_ex(k_empty, 148, 512, 80 * 1024, st, d, false); }, N, st));
  printf("empty 1020x256 97KB    : %.2f us/launch\n", time_chain([&] { launch_ex(k_empty, 1020, 256, 97 * 1024, st, d, false); }, N, st));
  printf("pdl   4x512 80KB smem  : %.2f us/launch\n", time_chain([&] { launch_ex(k_pdl, 4, 512, 80 * 1024, st, d, true); }, N, st));
  printf("pdl   1020x256 97KB    : %.2f us/launch\n", time_chain([&] { launch_ex(k_pdl, 1020, 256, 97 * 1024, st, d, true); }, N, st));
  printf("work  4x512 80KB nopdl : %.2f us/launch\n", time_chain([&] { launch_work(4, 512, 80 * 1024, st, d, 4096, false); }, N, st));
  printf("work  4x512 80KB pdl   : %.2f us/launch\n", time_chain([&] { launch_work(4, 512, 80 * 1024, st, d, 4096, true); }, N, st));
  // event record between launches (what VKSIFT_TRACE does)
  cudaEvent_t ev[2]; cudaEventCreate(&ev[0]); cudaEventCreate(&ev[1]);
  printf("work + 2 event records : %.2f us/launch\n", time_chain([&] { cudaEventRecord(ev[0], st); launch_work(4, 512, 80 * 1024, st, d, 4096, true); cudaEventRecord(ev[1], st); }, N, st));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
