// micro_lat.cu -- dependent-chain latencies of the FP64 building blocks on one warp (B200).
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void lat(double* out, long long* cyc, double x0) {
  __shared__ double sm[512];
  const int lane = threadIdx.x & 31;
  sm[threadIdx.x] = x0 + threadIdx.x;
  __syncthreads();
  double x = x0 + lane * 1e-3, y = 1.0000001;
  long long t0, t1; int k = 0;
#define T(name, N, body) t0 = clock64(); _Pragma("unroll") for (int i = 0; i < N; ++i) { body; } t1 = clock64(); \
  if (threadIdx.x == 0) cyc[k] = (t1 - t0) / N; ++k;
  T("dfma", 64, x = x * y + 1e-9)
  T("dadd", 64, x = x + y)
  T("dmul", 64, x = x * y)
  T("shfl64+dadd", 32, x += __shfl_xor_sync(0xffffffffu, x, 1))
  T("rsqrt", 16, x = rsqrt(x) + 2.0)
  T("sqrt", 16, x = sqrt(x) + 2.0)
  T("div", 16, x = 1.0 / x + 2.0)
  T("lds", 32, x += sm[((int)x + lane) & 511])
  T("ffma", 64, { float f = (float)x; f = f * 1.0001f + 1e-3f; x = f; })
  out[threadIdx.x] = x;
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 8 * 512); cudaMalloc(&c, 8 * 16);
  const char* names[] = {"dfma", "dadd", "dmul", "shfl64+dadd", "rsqrt+dadd", "sqrt+dadd", "rcp+dadd", "lds+cvt+dadd", "cvt+ffma+cvt"};
  for (int nt : {32, 512}) {
    lat<<<1, nt>>>(o, c, 1.5); cudaDeviceSynchronize();
    lat<<<1, nt>>>(o, c, 1.5); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads %d:", nt);
    for (int i = 0; i < 9; ++i) printf("  %s=%lld", names[i], h[i]);
    printf("  (%s)\n", cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
