// micro_panel2.cu -- times the SHIPPED panel kernel (qr_panel.cuh) phase by phase with clock64 stamps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cudart shared [-DTTB_PANEL_FINE] -DNSV=4 -I../tensortrains.jl_b200/csrc -o micro_panel2 micro_panel2.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
__device__ long long g_stamps[2048];
#define TTB_PANEL_STAMPS g_stamps
#include "common.cuh"
namespace ttb {
thread_local std::string g_err;
int fail(int code, const char*, ...) { return code; }
// mirror of the definitions at the top of sweep.cu
struct SweepCtx {
  double* slab; int64_t tt_stride; const int64_t* off; int* bond; int L1; int Q;
  double* X0; double* X1; int64_t x_stride;
  double* T; int64_t t_stride; int bw;
  double* Bv; double* Jv; int64_t bv_stride, jv_stride; int ldv, ldj;
  int* perm; double* sigma; int n_max;
  StepState* st;
  double* logz;
  double* stat_err; int* stat_rank; double* spectrum;
  int* status;
};
struct StepDesc {
  int site, nbr, dir, xin, rescale, qr_only, kind; double eps; int cap, bond_idx, bond_idx2, last, slot;
};
struct Dims { int p, n, k, mode, N; };
__device__ __forceinline__ Dims get_dims(const SweepCtx& c, const StepDesc& s, int b) {
  const int* bond = c.bond + (int64_t)b * c.L1;
  Dims d;
  if (s.dir == 0) { d.p = bond[s.site + 1] * c.Q; d.n = bond[s.site]; d.N = bond[s.nbr]; }
  else            { d.p = bond[s.site] * c.Q;     d.n = bond[s.site + 1]; d.N = bond[s.nbr + 1]; }
  d.k = min(d.p, d.n);
  d.mode = s.qr_only ? 0 : (d.p >= d.n ? 1 : 2);
  return d;
}
__device__ __forceinline__ double* xbuf(const SweepCtx& c, int which, int b) {
  return (which ? c.X1 : c.X0) + (int64_t)b * c.x_stride;
}
}  // namespace ttb
#include "tma_mma.cuh"
#include "qr_panel.cuh"
using namespace ttb;
#ifndef NSV
#define NSV 2
#endif

int main() {
#ifndef NCOLS
#define NCOLS 256
#endif
  const int p = 512, n = NCOLS;
  std::vector<double> h((size_t)p * n);
  srand(1);
  for (auto& x : h) x = rand() / (double)RAND_MAX;
  SweepCtx c = {};
  StepDesc s = {};
  int hb[3] = {NCOLS, NCOLS, 256};  // dir 0: p = bond[site+1]*Q, n = bond[site]
  cudaMalloc(&c.bond, sizeof(hb)); cudaMemcpy(c.bond, hb, sizeof(hb), cudaMemcpyHostToDevice);
  c.L1 = 3; c.Q = 2; c.bw = 32; c.x_stride = (int64_t)p * n; c.t_stride = 8 * 1024;
  cudaMalloc(&c.X0, h.size() * 8); cudaMalloc(&c.T, 8 * 1024 * 8); cudaMalloc(&c.st, sizeof(StepState));
  s.site = 1; s.nbr = 0; s.dir = 0; s.xin = 0; s.qr_only = 1;
  const size_t smem = sizeof(double) * (1024 * 16 + 1088) + 256;
  cudaFuncSetAttribute(k_qr_panel_flag<16, NSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int panel = 0; panel < (NCOLS > 32 ? 2 : 1); ++panel) {
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaMemcpy(c.X0, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
      cudaEventRecord(e0);
      k_qr_panel_flag<16, NSV><<<1, 1024 / NSV, smem>>>(c, s, panel);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    long long st[2048];
    cudaMemcpyFromSymbol(st, g_stamps, sizeof(st));
#ifdef TTB_PANEL_FINE
    printf("  col(OS): prev publish->entry | partial sums | butterfly | scalars | v + publish | own updates\n");
    for (int cc = 1; cc < 32; ++cc) {
      long long* a = &st[64 + 16 * cc];
      printf("  %2d(%d): %5lld | %5lld | %5lld | %5lld | %5lld | %5lld\n", cc, cc % 4, a[0] - st[8 + cc - 1], a[1] - a[0], a[2] - a[1],
             a[3] - a[2], a[4] - a[3], a[5] - a[4]);
    }
    printf("  next owner as consumer of reflector cc: owner entry->raw seen | early dots | wait ready (vs publish) | update | -> its owner step entry\n");
    for (int cc = 1; cc < 31; ++cc) {
      long long* a = &st[64 + 16 * cc];
      printf("  %2d: %5lld | %5lld | %5lld (%5lld after publish) | %5lld | %5lld\n", cc, a[6] - a[0], a[7] - a[6], a[8] - a[7], a[8] - st[8 + cc], a[9] - a[8],
             (cc % 4 == 3) ? st[64 + 16 * (cc + 1)] - a[9] : 0LL);
    }
#endif
    printf("panel %d: kernel %.1f us (%s)  cycles: loop %lld  Tbuild %lld  store %lld\n", panel, best * 1e3,
           cudaGetErrorString(cudaGetLastError()), st[1] - st[0], st[2] - st[1], st[3] - st[2]);
    printf("  ready-to-ready:");
    for (int cc = 1; cc < (NCOLS < 32 ? NCOLS : 32); cc += 1) printf(" %lld", st[8 + cc] - st[8 + cc - 1]);
    printf("\n");
  }
  return 0;
}
