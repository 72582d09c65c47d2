// micro_panel.cu -- clock64 instrumentation of the register-resident Householder panel (the chain of
// 32 dependent columns that bounds a d=256 compress! step).  Standalone: nvcc -arch=sm_100a, run on a B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_panel micro_panel.cu && ./micro_panel
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define RPL 16
#define LDV (32 * RPL)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int ld_flag(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ double dot16(const double (&v)[RPL], const double (&a)[RPL]) {
  double d0 = 0, d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
  for (int i = 0; i < RPL; i += 4) { d0 += v[i] * a[i]; d1 += v[i + 1] * a[i + 1]; d2 += v[i + 2] * a[i + 2]; d3 += v[i + 3] * a[i + 3]; }
  return (d0 + d1) + (d2 + d3);
}

// MODE 0: flags (as shipped)   MODE 1: same but consumers use __nanosleep backoff
// stamps[cc*4 + k]: k=0 owner start, 1 after norm reduce, 2 after rsqrt/div, 3 flag raised
template <int MODE>
__global__ void __launch_bounds__(512, 1) panel(double* X, int p, int rows, long long* stamps) {
  extern __shared__ __align__(16) double dyn[];
  double* vs = dyn;
  double* sh_tau = vs + 32 * LDV;
  int* ready = reinterpret_cast<int*>(sh_tau + 32);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double a[2][RPL];
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      const int r = lane + 32 * i, cs = wid + 16 * s2;
      a[s2][i] = r < rows ? X[r + (size_t)p * cs] : 0.0;
    }
  if (tid < 32) ready[tid] = 0;
  __syncthreads();
#pragma unroll 1
  for (int cc = 0; cc < 32; ++cc) {
    const int ow = cc & 15;
    const bool os = (cc >> 4) != 0;
    if (wid == ow) {
      long long t0 = clock64();
      double col[RPL];
#pragma unroll
      for (int i = 0; i < RPL; ++i) col[i] = os ? a[1][i] : a[0][i];
      double p0 = 0, p1 = 0, p2 = 0, p3 = 0;
#pragma unroll
      for (int i = 0; i < RPL; i += 4) {
        double x0 = (lane + 32 * i > cc) ? col[i] : 0.0, x1 = (lane + 32 * (i + 1) > cc) ? col[i + 1] : 0.0;
        double x2 = (lane + 32 * (i + 2) > cc) ? col[i + 2] : 0.0, x3 = (lane + 32 * (i + 3) > cc) ? col[i + 3] : 0.0;
        p0 += x0 * x0; p1 += x1 * x1; p2 += x2 * x2; p3 += x3 * x3;
      }
      const double xn2 = warp_sum((p0 + p1) + (p2 + p3));
      const double alpha = __shfl_sync(0xffffffffu, col[0], cc);
      long long t1 = clock64();
      double tau = 0.0, scal = 0.0, beta = alpha;
      if (xn2 > 0.0) {
        const double nn = alpha * alpha + xn2;
        const double rs = rsqrt(nn);
        const double nrm = nn * rs;
        const double sg = alpha >= 0.0 ? 1.0 : -1.0;
        beta = -sg * nrm;
        const double den = alpha - beta;
        scal = 1.0 / den;
        tau = den * sg * rs;
      }
      long long t2 = clock64();
      double* vcol = vs + (size_t)cc * LDV;
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        const int r = lane + 32 * i;
        double v = 0.0;
        if (r > cc) { v = col[i] * scal; col[i] = v; }
        else if (r == cc) { v = 1.0; col[i] = beta; }
        vcol[r] = v;
      }
#pragma unroll
      for (int i = 0; i < RPL; ++i) { if (os) a[1][i] = col[i]; else a[0][i] = col[i]; }
      if (lane == 0) sh_tau[cc] = tau;
      __threadfence_block();
      __syncwarp();
      if (lane == 0) *reinterpret_cast<volatile int*>(&ready[cc]) = 1;
      long long t3 = clock64();
      if (lane == 0) { stamps[cc * 8 + 0] = t0; stamps[cc * 8 + 1] = t1; stamps[cc * 8 + 2] = t2; stamps[cc * 8 + 3] = t3; }
    } else {
      if (lane == 0) {
        unsigned it = 0;
        while (ld_flag(&ready[cc]) == 0) {
          if (MODE == 1) __nanosleep(20);
          if (++it > (1u << 26)) __trap();
        }
      }
      __syncwarp();
      __threadfence_block();
    }
    long long t4 = clock64();
    const double tau = *reinterpret_cast<volatile double*>(&sh_tau[cc]);
    double v[RPL];
    const double* vcol = vs + (size_t)cc * LDV;
#pragma unroll
    for (int i = 0; i < RPL; ++i) v[i] = vcol[lane + 32 * i];
    const int first = ((cc + 1) >> 4) & 1;
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) {
      const int s2 = k2 == 0 ? first : 1 - first;
      const int cs = wid + 16 * s2;
      if (cs > cc) {
        double dot = s2 ? dot16(v, a[1]) : dot16(v, a[0]);
        dot = warp_sum(dot);
        if (tau != 0.0) {
          const double w = tau * dot;
          if (s2) {
#pragma unroll
            for (int i = 0; i < RPL; ++i) a[1][i] -= v[i] * w;
          } else {
#pragma unroll
            for (int i = 0; i < RPL; ++i) a[0][i] -= v[i] * w;
          }
        }
      }
      if (k2 == 0 && wid == ((cc + 1) & 15) && lane == 0) { stamps[cc * 8 + 4] = t4; stamps[cc * 8 + 5] = clock64(); }
    }
  }
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      const int r = lane + 32 * i, cs = wid + 16 * s2;
      if (r < rows) X[r + (size_t)p * cs] = a[s2][i];
    }
}

template <int MODE>
static void run(const char* name, double* dX, const std::vector<double>& h, int p, int rows, long long* dst) {
  const size_t smem = sizeof(double) * (32 * LDV + 32) + 128;
  cudaFuncSetAttribute(panel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaMemcpy(dX, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaEventRecord(e0);
    panel<MODE><<<1, 512, smem>>>(dX, p, rows, dst);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  std::vector<long long> st(32 * 8);
  cudaMemcpy(st.data(), dst, st.size() * 8, cudaMemcpyDeviceToHost);
  printf("%s: kernel %.1f us  err=%s\n", name, best * 1e3, cudaGetErrorString(cudaGetLastError()));
  printf("  col: norm+reduce | rsqrt+div | publish | owner total | next-owner wait->dot+update | ready-to-ready\n");
  for (int cc = 1; cc < 32; cc += 5) {
    long long* s = &st[cc * 8];
    long long* sp = &st[(cc - 1) * 8];
    printf("  %2d: %6lld | %6lld | %6lld | %6lld | %6lld | %6lld\n", cc, s[1] - s[0], s[2] - s[1], s[3] - s[2], s[3] - s[0],
           sp[5] - sp[4], s[3] - sp[3]);
  }
}

int main() {
  const int p = 512, rows = 512;
  std::vector<double> h((size_t)p * 32);
  srand(1);
  for (auto& x : h) x = rand() / (double)RAND_MAX;
  double* dX; long long* dst;
  cudaMalloc(&dX, h.size() * 8);
  cudaMalloc(&dst, 32 * 8 * 8);
  cudaMemset(dst, 0, 32 * 8 * 8);
  run<0>("flags/spin", dX, h, p, rows, dst);
  run<1>("flags/nanosleep", dX, h, p, rows, dst);
  return 0;
}
