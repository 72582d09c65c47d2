// micro_pdl.cu -- cost of one link in a chain of dependent kernels, with and without programmatic dependent
// launch (griddepcontrol).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o micro_pdl micro_pdl.cu
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void k_plain(double* x, int work) {
  extern __shared__ double sm[];
  double v = x[threadIdx.x & 31];
  for (int i = 0; i < work; ++i) v = v * 1.0000001 + 1e-9;
  if (threadIdx.x < 32) x[threadIdx.x] = v;
}
__global__ void k_pdl(double* x, int work) {
  extern __shared__ double sm[];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
  double v = x[threadIdx.x & 31];
  for (int i = 0; i < work; ++i) v = v * 1.0000001 + 1e-9;
  if (threadIdx.x < 32) x[threadIdx.x] = v;
}
// launch_dependents at the END (dependents only get scheduled when the primary is about to finish)
__global__ void k_pdl_late(double* x, int work) {
  extern __shared__ double sm[];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  double v = x[threadIdx.x & 31];
  for (int i = 0; i < work; ++i) v = v * 1.0000001 + 1e-9;
  if (threadIdx.x < 32) x[threadIdx.x] = v;
  asm volatile("griddepcontrol.launch_dependents;");
}

template <typename K>
float run(K k, bool pdl, int n, int grid, int thr, size_t smem, double* x, int work, cudaStream_t st) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(a, st);
    for (int i = 0; i < n; ++i) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(thr); cfg.dynamicSmemBytes = smem; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
      cudaLaunchKernelEx(&cfg, k, x, work);
    }
    cudaEventRecord(b, st);
    cudaStreamSynchronize(st);
  }
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  return 1e3f * ms / n;
}

int main() {
  double* x; cudaMalloc(&x, 256); cudaMemset(x, 0, 256);
  cudaStream_t st; cudaStreamCreate(&st);
  const int n = 4000;
  struct { int grid, thr; size_t smem; int work; const char* name; } cs[] = {
    {1, 32, 0, 0, "1 CTA x 32 thr, empty"},
    {1, 256, 140 * 1024, 0, "1 CTA x 256 thr, 140 KB smem, empty"},
    {56, 256, 150 * 1024, 0, "56 CTA x 256 thr, 150 KB smem, empty"},
    {1, 256, 140 * 1024, 2000, "1 CTA x 256 thr, 140 KB smem, ~8 us work"},
    {56, 256, 150 * 1024, 1000, "56 CTA x 256 thr, 150 KB smem, ~4 us work"},
  };
  for (auto& c : cs) {
    float t0 = run(k_plain, false, n, c.grid, c.thr, c.smem, x, c.work, st);
    float t1 = run(k_pdl, true, n, c.grid, c.thr, c.smem, x, c.work, st);
    float t2 = run(k_pdl_late, true, n, c.grid, c.thr, c.smem, x, c.work, st);
    printf("%-44s plain %.2f us/launch | PDL (early trigger) %.2f | PDL (late trigger) %.2f\n", c.name, t0, t1, t2);
  }
  return 0;
}
