// jacobi_lq.cuh -- k_jacobi_lq: SVD of a WIDE fold (p < n: nv = p <= 32 rows of length len <= 256 -- every step of the
// left sweep of config 3 once the bonds have collapsed: 16 x 256) as   X = L Qh  (Householder LQ, L nv x nv)  followed by
// the one-sided Jacobi on the rows of L.
//
// The Jacobi of k_jacobi_small<32,256> rotates rows of length 256: 150 rounds (15 rounds x ~10 sweeps) of ~900 cycles
// each, 69 us per step and 17.5 of the 77 ms of config 3's compress!.  Rows of length nv make a round a third cheaper
// (one element per lane group instead of eight per lane, 3-level butterflies, four pairs per warp as in k_jacobi_g8),
// and the LQ costs nv chained Householder steps once.  The rotations J are the same product of exact plane rotations
// as before (W = J[:, sel] stays orthonormal by construction); the factor that is absorbed into the neighbour is
// recomputed from the untouched input,  S = J X  (nv x nv x len FMAs), so Qh is never formed or applied, and the
// singular values are the row norms of the rotated L.  Householder LQ and Jacobi are both backward stable: the
// spectrum agrees with the direct Jacobi to O(eps sigma_1) (parity tests: bonds identical, spectra 1e-8, and the
// 40-digit arbiter on the near-threshold bonds).
// MEASURED (round 2, config 3, parity-green: all GPU tests pass with it): 19.2 ms of Jacobi per compress! against 17.35 ms
// for k_jacobi_small<32,256> -- the matrices are 12 x 256, the rounds on 12 x 12 still cost ~1000 cycles (the rotation
// chain, not the row length, sets the round), and the LQ (12 chained reflectors, 23 k cycles) plus S = J X (8 k) are
// added on top.  Opt-in (TTB_JACOBI_LQ=1), not the default.
// Reference: `svd(M)` inside the truncating functors, src/svd_trunc.jl:37,73,97,129.
#pragma once

namespace ttb {

constexpr int kJlqWL = 256, kJlqLD = 36;
__host__ __device__ constexpr size_t jlq_smem() { return sizeof(double) * (2 * 32 + 32 * (size_t)kJlqWL + 2 * 32 * kJlqLD + kJlqWL + 8); }

__global__ void __launch_bounds__(256, 1) k_jacobi_lq(SweepCtx c, StepDesc s) {
  pdl_enter();
  constexpr int WL = kJlqWL, LD = kJlqLD, W = 32, NT = 256, E2 = 2;
  const int b = blockIdx.x;
  const Dims d = get_dims(c, s, b);
  if (d.mode == 0) return;
  const int nv = d.k, len = d.n, p = d.p;   // host guarantees nv <= 32, len <= 256; mode 2 (wide) is where the LQ pays
  const double* X = xbuf(c, s.xin, b);
  const double scale_in = c.st[b].scale_in[s.slot];
  extern __shared__ __align__(16) double lsm[];
  double* sig = lsm;                 // 32
  double* srt = sig + 32;            // 32
  double* Y = srt + 32;              // [32][WL] rows of X, LQ in place
  double* B = Y + 32 * WL;           // [32][LD] rows of L (zero padded)
  double* J = B + 32 * LD;           // [32][LD]
  double* vb = J + 32 * LD;          // [WL] current reflector (v_k = 1), vb[WL] = tau
  __shared__ int rotated;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NWP = NT / 32;
  for (int i = tid; i < 32 * WL; i += NT) {
    const int r = i & 31, j = i >> 5;   // coalesced read of column j of X
    double v = 0.0;
    if (r < nv && j < len && !(d.mode == 1 && r > j)) v = X[r + (int64_t)p * j] * scale_in;
    Y[r * WL + j] = v;
  }
  for (int i = tid; i < 32 * LD; i += NT) { B[i] = 0.0; J[i] = ((i / LD) == (i % LD)) ? 1.0 : 0.0; }
  __syncthreads();
#ifdef TTB_JSYS_DEBUG
  long long t0_ = clock64(), t1_ = 0, t2_ = 0; int nsw_ = 0;
#endif
  const bool lq = d.mode == 2 && len > 32;
  const int lenb = lq ? nv : len;          // row length the Jacobi works on (len <= 32 without the LQ)
  if (lq) {
    // ---- Householder LQ on the rows: step k annihilates Y[k][k+1 ..] and updates the rows below
    for (int k = 0; k < nv; ++k) {
      if (wid == 0) {
        double* yk = Y + k * WL;
        double x[WL / 32], q2 = 0.0;
#pragma unroll
        for (int m = 0; m < WL / 32; ++m) {
          const int j = lane + 32 * m;
          x[m] = (j > k && j < len) ? yk[j] : 0.0;
          q2 = fma(x[m], x[m], q2);
        }
        q2 = warp_sum(q2);
        const double alpha = yk[k];
        double tau = 0.0, scal = 0.0, beta = alpha;
        if (q2 > 0.0) {
          const double nrm = sqrt(fma(alpha, alpha, q2));
          beta = alpha >= 0.0 ? -nrm : nrm;
          tau = (beta - alpha) / beta;
          scal = 1.0 / (alpha - beta);
        }
#pragma unroll
        for (int m = 0; m < WL / 32; ++m) {
          const int j = lane + 32 * m;
          vb[j] = j == k ? 1.0 : x[m] * scal;      // zero left of k and beyond len
        }
        if (lane == 0) { vb[WL] = tau; yk[k] = beta; }
      }
      __syncthreads();
      const double tau = vb[WL];
      if (tau != 0.0) {
        for (int i = k + 1 + wid; i < nv; i += NWP) {
          double* yi = Y + i * WL;
          double dot = 0.0;
#pragma unroll
          for (int m = 0; m < WL / 32; ++m) { const int j = lane + 32 * m; dot = fma(vb[j], (j >= k) ? yi[j] : 0.0, dot); }
          dot = warp_sum(dot) * tau;
#pragma unroll
          for (int m = 0; m < WL / 32; ++m) { const int j = lane + 32 * m; if (j >= k && j < len) yi[j] -= dot * vb[j]; }
        }
      }
      __syncthreads();
    }
    for (int i = tid; i < nv * nv; i += NT) { const int r = i / nv, k = i - r * nv; B[r * LD + k] = k <= r ? Y[r * WL + k] : 0.0; }
  } else {
    for (int i = tid; i < nv * 32; i += NT) { const int r = i >> 5, j = i & 31; B[r * LD + j] = j < len ? Y[r * WL + j] : 0.0; }
  }
  __syncthreads();
#ifdef TTB_JSYS_DEBUG
  t1_ = clock64();
#endif
  // ---- one-sided Jacobi on the rows of B (k_jacobi_g8<32>'s round: four pairs per warp, 8 lanes per pair)
  const int nvp = (nv + 1) & ~1;
  const int npairs = nvp / 2, rounds = nvp - 1;
  const double tol2 = (double)max(lenb, 2) * 2.220446049250313e-16 * 2.220446049250313e-16;
  double* sq = sig;
  const int grp = lane >> 3, e0 = (lane & 7) * 2;
  const int q = wid * 4 + grp;
  bool conv = (nv < 2);
  for (int sweep = 0; sweep < s.max_sweeps && !conv; ++sweep) {
    if (tid < W) {
      double a0 = 0.0;
      if (tid < nv) {
#pragma unroll 8
        for (int m = 0; m < W; ++m) { const double xx = B[tid * LD + m]; a0 += xx * xx; }
      }
      sq[tid] = a0;
    }
    if (tid == 0) rotated = 0;
    __syncthreads();
    for (int rd = 0; rd < rounds; ++rd) {
      int i = 0, j = 0;
      bool act = q < npairs;
      if (act) { rr_pair(nvp, rd, q, i, j); act = j < nv; }
      if (!act) { i = 0; j = 0; }
      double2* bi = reinterpret_cast<double2*>(B + i * LD + e0);
      double2* bj = reinterpret_cast<double2*>(B + j * LD + e0);
      double2* ji = reinterpret_cast<double2*>(J + i * LD + e0);
      double2* jj = reinterpret_cast<double2*>(J + j * LD + e0);
      double2 x[E2], y[E2], u[E2], v[E2];
#pragma unroll
      for (int m = 0; m < E2; ++m) { x[m] = bi[8 * m]; y[m] = bj[8 * m]; u[m] = ji[8 * m]; v[m] = jj[8 * m]; }
      const double al = sq[i], be = sq[j];
      double g0 = 0.0, g1 = 0.0;
#pragma unroll
      for (int m = 0; m < E2; ++m) { g0 += x[m].x * y[m].x; g1 += x[m].y * y[m].y; }
      double ga = g0 + g1;
      ga += __shfl_xor_sync(0xffffffffu, ga, 4);
      ga += __shfl_xor_sync(0xffffffffu, ga, 2);
      ga += __shfl_xor_sync(0xffffffffu, ga, 1);
      const bool rot = act && al > 0.0 && be > 0.0 && ga * ga > tol2 * al * be;
      const double da = be - al, b2 = 2.0 * ga;
      double h2 = da * da + b2 * b2, das = da, b2s = b2;
      if (rot && !(h2 > 1e-280 && h2 < 1e280)) {
        const double sc = 1.0 / fmax(fabs(da), fabs(b2));
        das = da * sc; b2s = b2 * sc;
        h2 = das * das + b2s * b2s;
      }
      const double r = rsqrt_nr(rot ? h2 : 1.0);
      const double c2 = fabs(das) * r, s2 = (da < 0.0 ? -b2s : b2s) * r;
      const double uu = rot ? 0.5 + 0.5 * c2 : 1.0;
      const double ic = rsqrt_nr(uu);
      const double cs = uu * ic, sn = 0.5 * s2 * ic;
      if (rot) {
#pragma unroll
        for (int m = 0; m < E2; ++m) {
          bi[8 * m] = make_double2(cs * x[m].x - sn * y[m].x, cs * x[m].y - sn * y[m].y);
          bj[8 * m] = make_double2(sn * x[m].x + cs * y[m].x, sn * x[m].y + cs * y[m].y);
          ji[8 * m] = make_double2(cs * u[m].x - sn * v[m].x, cs * u[m].y - sn * v[m].y);
          jj[8 * m] = make_double2(sn * u[m].x + cs * v[m].x, sn * u[m].y + cs * v[m].y);
        }
        if ((lane & 7) == 0) {
          const double t = sn * ic;
          sq[i] = al - t * ga;
          sq[j] = be + t * ga;
          if (ga * ga > kJacobiBig2 * al * be) rotated = 1;
        }
      }
      __syncthreads();
    }
    conv = (rotated == 0);
    __syncthreads();
#ifdef TTB_JSYS_DEBUG
    ++nsw_;
#endif
  }
#ifdef TTB_JSYS_DEBUG
  t2_ = clock64();
#endif
  if (!conv && tid == 0) atomicOr(c.status, ST_NOCONV);
  jacobi_finish(c, s, b, B, LD, nv, lenb, sig, srt);
  double* Bg = c.Bv + (int64_t)b * c.bv_stride;
  double* Jg = c.Jv + (int64_t)b * c.jv_stride;
  if (lq) {
    // S = J X from the untouched input: one column of X (nv contiguous values) per thread
    for (int e = tid; e < len; e += NT) {
      double xc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) xc[i] = i < nv ? X[i + (int64_t)p * e] * scale_in : 0.0;
      for (int v2 = 0; v2 < nv; ++v2) {
        const double* jr = J + v2 * LD;
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < 32; ++i) acc = fma(jr[i], xc[i], acc);
        Bg[(int64_t)v2 * c.ldv + e] = acc;
      }
    }
  } else {
    for (int i = tid; i < nv * len; i += NT) {
      const int e = i % len, v2 = i / len;
      Bg[(int64_t)v2 * c.ldv + e] = B[v2 * LD + e];
    }
  }
  for (int i = tid; i < nv * nv; i += NT) {
    const int e = i % nv, v2 = i / nv;
    Jg[(int64_t)v2 * c.ldj + e] = J[v2 * LD + e];
  }
#ifdef TTB_JSYS_DEBUG
  if (b == 0 && tid == 0 && s.site % 32 == 0) printf("[jlq] site %d nv=%d len=%d lq=%d: load+LQ %lld  jacobi %lld (%d sweeps)  finish+S %lld cycles\n", s.site, nv, len, (int)lq, t1_ - t0_, t2_ - t1_, nsw_, clock64() - t2_);
#endif
}

}  // namespace ttb
