// gemm_items.cuh -- one launch, a whole table of FP64 tensor-core products with their own shapes (mma.sync m16n8k8):
// the sandwiches of the MPS environments (mps.cu) and the merged two-site tensor (twosite.cu).
#pragma once

#include "common.cuh"
#include "tma_mma.cuh"

namespace ttb {

// one product of a table: C (m x n) = sum_{s < nsum} op(A + s sA) op(B + s sB), column-major operands
struct GemmItem {
  const double* A;
  const double* B;
  double* C;
  int m, n, k;
  int lda, ldb, ldc;
  int nsum, pad_;
  int64_t sA, sB;
};

// TA: A is stored k x m (op(A) = A^T); TB: B is stored n x k (op(B) = B^T).  64 x 64 tile per CTA, 4 warps of
// 32 x 32 (2 x 4 DMMA tiles), K staged 16 at a time through shared memory with a register prefetch of the
// next stage.  grid = (max tiles of any item, items): surplus CTAs exit.
template <bool TA, bool TB>
__global__ void __launch_bounds__(128) k_gemm_items(const GemmItem* __restrict__ items) {
  const GemmItem it = items[blockIdx.y];
  const int tiles_m = (it.m + 63) >> 6, tiles_n = (it.n + 63) >> 6;
  if ((int)blockIdx.x >= tiles_m * tiles_n) return;
  const int i0 = ((int)blockIdx.x % tiles_m) * 64, j0 = ((int)blockIdx.x / tiles_m) * 64;
  __shared__ double As[16][72], Bs[16][72];   // [k][m], [k][n]; stride 72 = 8 mod 32: fragment loads conflict-free
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const int wm = (wid & 1) * 32, wn = (wid >> 1) * 32;
  double acc[2][4][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][b][e] = 0.0;
  const int nst = (it.k + 15) >> 4, total = nst * it.nsum;
  double ra[8], rb[8];
  auto load = [&](int st) {
    const int s = st / nst, k0 = (st - s * nst) * 16;
    const double* Ap = it.A + (int64_t)s * it.sA;
    const double* Bp = it.B + (int64_t)s * it.sB;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int idx = tid + 128 * r;
      int mm, kk;
      if (TA) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
      const bool ok = (i0 + mm < it.m) && (k0 + kk < it.k);
      ra[r] = ok ? (TA ? Ap[(k0 + kk) + (int64_t)it.lda * (i0 + mm)] : Ap[(i0 + mm) + (int64_t)it.lda * (k0 + kk)]) : 0.0;
      int nn, kb;
      if (TB) { nn = idx & 63; kb = idx >> 6; } else { kb = idx & 15; nn = idx >> 4; }
      const bool okb = (j0 + nn < it.n) && (k0 + kb < it.k);
      rb[r] = okb ? (TB ? Bp[(j0 + nn) + (int64_t)it.ldb * (k0 + kb)] : Bp[(k0 + kb) + (int64_t)it.ldb * (j0 + nn)]) : 0.0;
    }
  };
  if (total > 0) load(0);
  for (int st = 0; st < total; ++st) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int idx = tid + 128 * r;
      if (TA) As[idx & 15][idx >> 4] = ra[r]; else As[idx >> 6][idx & 63] = ra[r];
      if (TB) Bs[idx >> 6][idx & 63] = rb[r]; else Bs[idx & 15][idx >> 4] = rb[r];
    }
    __syncthreads();
    if (st + 1 < total) load(st + 1);
#pragma unroll
    for (int ks = 0; ks < 16; ks += 8) {
      double af[2][4], bf[4][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r0 = wm + 16 * mt + g;
        af[mt][0] = As[ks + tq][r0]; af[mt][1] = As[ks + tq][r0 + 8];
        af[mt][2] = As[ks + tq + 4][r0]; af[mt][3] = As[ks + tq + 4][r0 + 8];
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int c0 = wn + 8 * nt + g;
        bf[nt][0] = Bs[ks + tq][c0]; bf[nt][1] = Bs[ks + tq + 4][c0];
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma_16x8x8(acc[mt][nt], af[mt], bf[nt]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = i0 + wm + 16 * mt + g + ((e & 2) ? 8 : 0), j = j0 + wn + 8 * nt + 2 * tq + (e & 1);
        if (i < it.m && j < it.n) it.C[i + (int64_t)it.ldc * j] = acc[mt][nt][e];
      }
}

static inline unsigned tiles64(int m, int n) { return (unsigned)(((m + 63) / 64) * ((n + 63) / 64)); }

template <bool TA, bool TB>
static inline void launch_gemm(ttb_tt* h, const GemmItem* d_items, int nitems, unsigned max_tiles) {
  if (nitems <= 0) return;
  TTB_LAUNCH(h, "k_gemm_items", (k_gemm_items<TA, TB><<<dim3(max_tiles, (unsigned)nitems), 128, 0, h->stream>>>(d_items)));
}

}  // namespace ttb
