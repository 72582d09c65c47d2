// jacobi_cluster.cuh -- one-sided Jacobi SVD for matrices that do not fit one CTA's shared memory
// (k >= ~115 at k = n: the full-rank randn fill of config 3, or a truncating sweep run first on wide bonds).
//
// The single-CTA kernel then works out of L2 with 32 warps: ~15 ms per 256 x 256 SVD.  Here a thread-block
// CLUSTER of 8 CTAs shares the matrix: the k rows form 16 blocks of rb rows; in block-round kr CTA c holds
// the block pair the circle method assigns to it (8 disjoint block pairs per round, 15 rounds per sweep),
// with both blocks -- rows of B and of the accumulated rotations J -- in ITS shared memory, and performs
// every rotation between the two blocks (rb sub-rounds of rb disjoint pairs, one warp per pair) and, in the
// first block-round of a sweep, every rotation inside each block.  That is exactly the 255 rounds x 128 pairs
// of the cyclic ordering, 8 SMs wide.  Blocks live at their home rows of the global workspace (L2 resident,
// 0.5 MB) between block-rounds: store, __threadfence, hardware cluster barrier, load the next pair with
// L1-bypassing loads.  Convergence is a cluster-wide rotation count exchanged through global memory.
// Reference: `svd(M)` inside the truncating functors, src/svd_trunc.jl:37,73,97,129 (LAPACK gesdd there;
// one-sided Jacobi here: same singular values to O(eps sigma_1), W orthonormal by construction).
#pragma once
#include <cooperative_groups.h>

namespace ttb {

namespace cg = cooperative_groups;

constexpr int kJcCluster = 8;            // CTAs per cluster (portable maximum)
constexpr int kJcBlocks = 2 * kJcCluster;
constexpr int kJcThreads = 512;
#ifdef TTB_JC_STAMPS
__device__ long long g_jc_stamps[8];   // cycles of CTA 0: load, norms, in-block, cross, store+fence, cluster sync, sweeps, block-rounds
#define JC_T(i) do { if (tid == 0) { const long long t_ = clock64(); acc_[i] += t_ - last_; last_ = t_; } } while (0)
#else
#define JC_T(i) do { } while (0)
#endif

__host__ __device__ inline int jc_rows_per_block(int nv) {
  int rb = (nv + kJcBlocks - 1) / kJcBlocks;
  rb = (rb + 1) & ~1;
  return rb < 2 ? 2 : rb;
}
// trains the single-CTA kernel handles when both paths are launched (same predicate as k_jacobi's SMEM)
__device__ __forceinline__ bool jc_single_takes(const SweepCtx& c, const StepDesc& s, const Dims& d) {
  return s.jc_sd > 0 && 2 * (int64_t)c.n_max + (int64_t)d.k * (d.n + d.k) <= (int64_t)s.jc_sd;
}
__host__ __device__ inline size_t jc_smem_bytes(int nv, int len) {
  const int rb = jc_rows_per_block(nv);
  return sizeof(double) * ((size_t)2 * rb * (len + nv) + 2 * rb) + 16;
}

// jacobi_pair with row i (and its rotation row) RESIDENT IN REGISTERS: in the cross-block phase warp w
// rotates its fixed row i = w against a different row j every sub-round, so only row j moves through
// shared memory -- the phase is shared-memory-bandwidth bound (16 pairs x 4 rows x 2 KB x load+store per
// sub-round) and this halves the traffic.  Requires len, nv <= 32*NE.
template <int NE>
__device__ __forceinline__ bool jacobi_pair_fixed(double (&xi)[NE], double (&jx)[NE], double& al, double* bj, double* jj,
                                                  int len, int nv, double* sqj, double tol2, int lane) {
  double yj[NE], jy[NE];
#pragma unroll
  for (int m = 0; m < NE; ++m) {
    const int e = lane + 32 * m;
    yj[m] = e < len ? bj[e] : 0.0;
    jy[m] = e < nv ? jj[e] : 0.0;
  }
  const double be = *sqj;
  double g0 = 0.0, g1 = 0.0;
#pragma unroll
  for (int m = 0; m < NE; m += 2) { g0 += xi[m] * yj[m]; g1 += xi[m + 1] * yj[m + 1]; }
  const double ga = warp_sum(g0 + g1);
  if (!(al > 0.0 && be > 0.0 && ga * ga > tol2 * al * be)) return false;
  const double da = be - al, b2 = 2.0 * ga;
  const double h2 = da * da + b2 * b2;
  double r = rsqrt_nr(h2), das = da, b2s = b2;
  if (!(h2 > 1e-280 && h2 < 1e280)) {
    const double sc = 1.0 / fmax(fabs(da), fabs(b2));
    das = da * sc; b2s = b2 * sc;
    r = rsqrt(das * das + b2s * b2s);
  }
  const double c2 = fabs(das) * r, s2 = (da < 0.0 ? -b2s : b2s) * r;
  const double u = 0.5 + 0.5 * c2;
  const double ic = rsqrt_nr(u);
  const double cs = u * ic, sn = 0.5 * s2 * ic;
#pragma unroll
  for (int m = 0; m < NE; ++m) {
    const int e = lane + 32 * m;
    const double x = xi[m], y = yj[m];
    xi[m] = cs * x - sn * y;
    if (e < len) bj[e] = sn * x + cs * y;
    const double p = jx[m], q = jy[m];
    jx[m] = cs * p - sn * q;
    if (e < nv) jj[e] = sn * p + cs * q;
  }
  const double t = sn * ic;
  const bool big = ga * ga > kJacobiBig2 * al * be;   // counts against convergence (common.cuh)
  al -= t * ga;
  *sqj = be + t * ga;
  return big;
}

// copy `rows` rows of `n` doubles between a strided global matrix and a packed shared block; 16-byte
// accesses with four requests in flight per thread when the layout allows it (L2 latency, not bandwidth,
// is what a block exchange costs).  TO_SMEM: global -> shared past L1; otherwise shared -> global.
template <bool TO_SMEM>
__device__ __forceinline__ void jc_copy_rows(double* g, int64_t ldg, double* sh, int n, int rows, int row0, int nv, int tid) {
  if (((n | (int)ldg) & 1) == 0 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(sh)) & 15) == 0) {
    const int n2 = n >> 1, tot = rows * n2;
#pragma unroll 4
    for (int i = tid; i < tot; i += kJcThreads) {
      const int r = i / n2, e = i - r * n2;
      double2* sp = reinterpret_cast<double2*>(sh + (size_t)r * n) + e;
      double2* gp = reinterpret_cast<double2*>(g + (int64_t)(row0 + r) * ldg) + e;
      if (TO_SMEM) *sp = (row0 + r < nv) ? __ldcg(gp) : make_double2(0.0, 0.0);
      else if (row0 + r < nv) *gp = *sp;
    }
  } else {
    const int tot = rows * n;
#pragma unroll 4
    for (int i = tid; i < tot; i += kJcThreads) {
      const int r = i / n, e = i - r * n;
      if (TO_SMEM) sh[(size_t)r * n + e] = (row0 + r < nv) ? __ldcg(g + (int64_t)(row0 + r) * ldg + e) : 0.0;
      else if (row0 + r < nv) g[(int64_t)(row0 + r) * ldg + e] = sh[(size_t)r * n + e];
    }
  }
}

// rows of R (tall) / X (wide), scaled, into the vector-major global workspace; J = I
__global__ void k_jacobi_init(SweepCtx c, StepDesc s) {
  pdl_enter();
  const int b = blockIdx.y;
  const Dims d = get_dims(c, s, b);
  if (d.mode == 0 || jc_single_takes(c, s, d)) return;
  const int nv = d.k, len = d.n, p = d.p;
  const double* X = xbuf(c, s.xin, b);
  const double scale_in = c.st[b].scale_in[s.slot];
  double* Bg = c.Bv + (int64_t)b * c.bv_stride;
  double* Jg = c.Jv + (int64_t)b * c.jv_stride;
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int tr = (nv + 31) / 32, tc = (len + 31) / 32;
  for (int tl = blockIdx.x; tl < tr * tc; tl += gridDim.x) {
    const int r0 = (tl % tr) * 32, j0 = (tl / tr) * 32;
    for (int jj = ty; jj < 32; jj += 8) {
      const int r = r0 + tx, j = j0 + jj;
      double v = 0.0;
      if (r < nv && j < len && !(d.mode == 1 && r > j)) v = X[r + (int64_t)p * j] * scale_in;
      tile[jj][tx] = v;
    }
    __syncthreads();
    for (int rr = ty; rr < 32; rr += 8) {
      const int r = r0 + rr, j = j0 + tx;
      if (r < nv && j < len) Bg[(int64_t)r * c.ldv + j] = tile[tx][rr];
    }
    __syncthreads();
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)nv * nv; i += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(i % nv), v = (int)(i / nv);
    Jg[(int64_t)v * c.ldj + e] = (e == v) ? 1.0 : 0.0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { c.st[b].rot[0] = 0; c.st[b].rot[1] = 0; }
}

template <int NE>
__global__ void __cluster_dims__(kJcCluster, 1, 1) __launch_bounds__(kJcThreads, 1) k_jacobi_cluster(SweepCtx c, StepDesc s) {
  pdl_enter();
  cg::cluster_group cl = cg::this_cluster();
  const int b = blockIdx.y;
  const int rank = blockIdx.x;  // gridDim.x == cluster size
  const Dims d = get_dims(c, s, b);
  if (d.mode == 0 || jc_single_takes(c, s, d)) return;      // uniform over the cluster
  const int nv = d.k, len = d.n;
  const int rb = jc_rows_per_block(nv), rb2 = 2 * rb;
  double* Bg = c.Bv + (int64_t)b * c.bv_stride;
  double* Jg = c.Jv + (int64_t)b * c.jv_stride;
  extern __shared__ __align__(16) double sm[];
  double* Bs = sm;                          // [2 rb][len]
  double* Js = Bs + (size_t)rb2 * len;      // [2 rb][nv]
  double* sq = Js + (size_t)rb2 * nv;       // [2 rb]
  __shared__ int rotated;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = kJcThreads / 32;
  const double tol = sqrt((double)max(len, 2)) * 2.220446049250313e-16;
  const double tol2 = tol * tol;
  int* rot = c.st[b].rot;
#ifdef TTB_JC_STAMPS
  long long acc_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, last_ = clock64();
#endif
  bool conv = (nv < 2);
  int sweep = 0;
  for (; sweep < s.max_sweeps && !conv; ++sweep) {
    if (tid == 0) rotated = 0;
    for (int kr = 0; kr < kJcBlocks - 1; ++kr) {
      int bi, bj;
      rr_pair(kJcBlocks, kr, rank, bi, bj);
      // ---- load the two blocks (home rows bi*rb.., bj*rb..) past L1; rows >= nv are zero padding
      for (int half = 0; half < 2; ++half) {
        const int g0 = (half ? bj : bi) * rb;
        jc_copy_rows<true>(Bg, c.ldv, Bs + (size_t)half * rb * len, len, rb, g0, nv, tid);
        jc_copy_rows<true>(Jg, c.ldj, Js + (size_t)half * rb * nv, nv, rb, g0, nv, tid);
      }
      __syncthreads();
      JC_T(0);
      for (int v = wid; v < rb2; v += NW) {
        const double* row = Bs + (size_t)v * len;
        double a0 = 0.0, a1 = 0.0;
        int e = lane;
        for (; e + 32 < len; e += 64) { a0 += row[e] * row[e]; a1 += row[e + 32] * row[e + 32]; }
        if (e < len) a0 += row[e] * row[e];
        const double al = warp_sum(a0 + a1);
        if (lane == 0) sq[v] = al;
      }
      __syncthreads();
      JC_T(1);
      bool did = false;
      // ---- rotations inside each block, once per sweep (first block-round): rb/2 + rb/2 pairs per sub-round
      if (kr == 0) {
        for (int rd = 0; rd < rb - 1; ++rd) {
          if (wid < rb) {
            const int half = wid >= rb / 2 ? 1 : 0;
            int i, j;
            rr_pair(rb, rd, wid - half * (rb / 2), i, j);
            i += half * rb; j += half * rb;
            did |= jacobi_pair<NE, NE>(Bs + (size_t)i * len, Bs + (size_t)j * len, Js + (size_t)i * nv, Js + (size_t)j * nv,
                                       len, nv, sq + i, sq + j, tol2, lane);
          }
          __syncthreads();
        }
      }
      JC_T(2);
      // ---- rotations between the two blocks: rb sub-rounds of rb disjoint pairs
      if (len <= 32 * NE && nv <= 32 * NE) {
        // warp w keeps its row i = w (and the rotation row) in registers through all rb sub-rounds
        double xi[NE], jx[NE], al = 0.0;
        if (wid < rb) {
#pragma unroll
          for (int m = 0; m < NE; ++m) {
            const int e = lane + 32 * m;
            xi[m] = e < len ? Bs[(size_t)wid * len + e] : 0.0;
            jx[m] = e < nv ? Js[(size_t)wid * nv + e] : 0.0;
          }
          al = sq[wid];
        }
        for (int rd = 0; rd < rb; ++rd) {
          if (wid < rb) {
            int j = wid + rd; if (j >= rb) j -= rb;
            j += rb;
            did |= jacobi_pair_fixed<NE>(xi, jx, al, Bs + (size_t)j * len, Js + (size_t)j * nv, len, nv, sq + j, tol2, lane);
          }
          __syncthreads();
        }
        if (wid < rb) {
#pragma unroll
          for (int m = 0; m < NE; ++m) {
            const int e = lane + 32 * m;
            if (e < len) Bs[(size_t)wid * len + e] = xi[m];
            if (e < nv) Js[(size_t)wid * nv + e] = jx[m];
          }
        }
        __syncthreads();
      } else {
        for (int rd = 0; rd < rb; ++rd) {
          if (wid < rb) {
            const int i = wid;
            int j = wid + rd; if (j >= rb) j -= rb;
            j += rb;
            did |= jacobi_pair<NE, NE>(Bs + (size_t)i * len, Bs + (size_t)j * len, Js + (size_t)i * nv, Js + (size_t)j * nv,
                                       len, nv, sq + i, sq + j, tol2, lane);
          }
          __syncthreads();
        }
      }
      if (did && lane == 0) rotated = 1;
      JC_T(3);
      // ---- blocks back to their home rows
      for (int half = 0; half < 2; ++half) {
        const int g0 = (half ? bj : bi) * rb;
        jc_copy_rows<false>(Bg, c.ldv, Bs + (size_t)half * rb * len, len, rb, g0, nv, tid);
        jc_copy_rows<false>(Jg, c.ldj, Js + (size_t)half * rb * nv, nv, rb, g0, nv, tid);
      }
      __threadfence();
      __syncthreads();
      if (kr == kJcBlocks - 2 && tid == 0) {
        if (rotated) atomicAdd(&rot[sweep & 1], 1);
        __threadfence();
      }
      JC_T(4);
      cl.sync();
      JC_T(5);
#ifdef TTB_JC_STAMPS
      if (tid == 0) acc_[7] += 1;
#endif
    }
    const int total = __ldcg(&rot[sweep & 1]);  // L2 load: identical in every CTA after the barrier
    if (rank == 0 && tid == 0) rot[(sweep + 1) & 1] = 0;  // ordered before its next use by 15 cluster barriers
    conv = (total == 0);
  }
#ifdef TTB_JC_STAMPS
  if (rank == 0 && tid == 0 && b == 0) { acc_[6] = sweep; for (int i = 0; i < 8; ++i) g_jc_stamps[i] = acc_[i]; }
#endif
  if (!conv && rank == 0 && tid == 0) atomicOr(c.status, ST_NOCONV);
}

// singular values, order, rank rule for the cluster path (the rotated rows are in the global workspace)
__global__ void __launch_bounds__(1024) k_jacobi_finish(SweepCtx c, StepDesc s) {
  pdl_enter();
  const int b = blockIdx.x;
  const Dims d = get_dims(c, s, b);
  if (d.mode == 0 || jc_single_takes(c, s, d)) return;
  extern __shared__ double sm[];
  jacobi_finish(c, s, b, c.Bv + (int64_t)b * c.bv_stride, c.ldv, d.k, d.n, sm, sm + c.n_max);
}

}  // namespace ttb
