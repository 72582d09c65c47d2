// jacobi_small.cuh -- one-sided Jacobi for k <= WR rows of length n <= WL, compile-time strides:
// <64,64> the truncating steps of config 2 (k = n = 64), <32,32> a config-5 batch (k = n = 32, hundreds of
// CTAs), <32,256> the left sweep of config 3 (k ~ 16 retained directions x Q, n = 256).
//
// With 8 (or, in a batch, 16) warps per SM sub-partition a Jacobi round is bound by INSTRUCTION ISSUE, not
// latency: the general kernel executes ~350 instructions per pair (bounds checks, 64-bit address math).
// Here rows are zero-padded to compile-time strides in shared memory -- no predicates, 32-bit offsets,
// WL/32 row elements and WR/32 rotation elements per lane: ~120-180 instructions per pair.  With few warps
// (config 3) the same leanness shortens the latency chain of a round, which one warp issues alone.
// Reference: `svd(M)` inside the truncating functors, src/svd_trunc.jl:37,73,97,129; the rank rule that
// follows (jacobi_finish -> rank_rule in sweep.cu) restates :38-47, :74, :98-100, :130-140 verbatim.
#pragma once

namespace ttb {

// Two pairs of one round in ONE warp, interleaved in source order (both pairs' loads, then both dots, one
// combined butterfly, both rotations, both stores) and branch-free: a pair that needs no rotation gets
// cs = 1, sn = 0.  In a batch the SM is full of warps that each walk one long dependent chain per pair;
// with half the warps and two independent chains per warp the same issue slots hide twice the latency.
template <int WL, int WR>
__device__ __forceinline__ int jacobi_two_pairs(double* B, double* J, double* sq, int i0, int j0, int i1, int j1, double tol2,
                                                int lane) {
  constexpr int EPL = WL / 32, JPL = WR / 32;
  double* bi[2] = {B + i0 * WL + lane, B + i1 * WL + lane};
  double* bj[2] = {B + j0 * WL + lane, B + j1 * WL + lane};
  double* ji[2] = {J + i0 * WR + lane, J + i1 * WR + lane};
  double* jj[2] = {J + j0 * WR + lane, J + j1 * WR + lane};
  const int ii[2] = {i0, i1}, jx[2] = {j0, j1};
  double x[2][EPL], y[2][EPL], u_[2][JPL], v_[2][JPL], al[2], be[2], ga[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
#pragma unroll
    for (int m = 0; m < EPL; ++m) { x[k][m] = bi[k][32 * m]; y[k][m] = bj[k][32 * m]; }
#pragma unroll
    for (int m = 0; m < JPL; ++m) { u_[k][m] = ji[k][32 * m]; v_[k][m] = jj[k][32 * m]; }
    al[k] = sq[ii[k]]; be[k] = sq[jx[k]];
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    double g0 = 0.0;
#pragma unroll
    for (int m = 0; m < EPL; ++m) g0 += x[k][m] * y[k][m];
    ga[k] = g0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ga[0] += __shfl_xor_sync(0xffffffffu, ga[0], o);
    ga[1] += __shfl_xor_sync(0xffffffffu, ga[1], o);
  }
  int rotated = 0;
  double cs[2], sn[2], dsq[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const bool rot = al[k] > 0.0 && be[k] > 0.0 && ga[k] * ga[k] > tol2 * al[k] * be[k];
    // same rotation as the single-pair path: cos 2th = |da|/h, sin 2th = sign(da) b2/h
    const double da = be[k] - al[k], b2 = 2.0 * ga[k];
    double h2 = da * da + b2 * b2, das = da, b2s = b2;
    if (rot && !(h2 > 1e-280 && h2 < 1e280)) {   // rare: rescale before squaring
      const double sc = 1.0 / fmax(fabs(da), fabs(b2));
      das = da * sc; b2s = b2 * sc;
      h2 = das * das + b2s * b2s;
    }
    const double r = rsqrt_nr(rot ? h2 : 1.0);
    const double c2 = fabs(das) * r, s2 = (da < 0.0 ? -b2s : b2s) * r;
    const double uu = rot ? 0.5 + 0.5 * c2 : 1.0;
    const double ic = rsqrt_nr(uu);
    cs[k] = rot ? uu * ic : 1.0;
    sn[k] = rot ? 0.5 * s2 * ic : 0.0;
    dsq[k] = sn[k] * ic * ga[k];
    rotated |= (rot && ga[k] * ga[k] > kJacobiBig2 * al[k] * be[k]) ? 1 : 0;
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (sn[k] != 0.0 || cs[k] != 1.0) {
#pragma unroll
      for (int m = 0; m < EPL; ++m) {
        bi[k][32 * m] = cs[k] * x[k][m] - sn[k] * y[k][m];
        bj[k][32 * m] = sn[k] * x[k][m] + cs[k] * y[k][m];
      }
#pragma unroll
      for (int m = 0; m < JPL; ++m) {
        ji[k][32 * m] = cs[k] * u_[k][m] - sn[k] * v_[k][m];
        jj[k][32 * m] = sn[k] * u_[k][m] + cs[k] * v_[k][m];
      }
      sq[ii[k]] = al[k] - dsq[k];
      sq[jx[k]] = be[k] + dsq[k];
    }
  }
  return rotated;
}

template <int WR, int WL>
__global__ void __launch_bounds__(16 * WR, (WR == 32 && WL == 32) ? 2 : 1) k_jacobi_small(SweepCtx c, StepDesc s) {
  pdl_enter();
  constexpr int EPL = WL / 32;      // row elements per lane
  constexpr int JPL = WR / 32;      // rotation-row elements per lane
  const int b = blockIdx.x;
  const Dims d = get_dims(c, s, b);
  if (d.mode == 0) return;
  const int nv = d.k, len = d.n, p = d.p;   // host guarantees nv <= WR, len <= WL
  const double* X = xbuf(c, s.xin, b);
  const double scale_in = c.st[b].scale_in[s.slot];
  extern __shared__ double sm[];
  double* sig = sm;                  // WR  (cached squared norms, then singular values)
  double* srt = sig + WR;            // WR
  double* B = srt + WR;              // [WR][WL] rows zero-padded
  double* J = B + WR * WL;           // [WR][WR]
  __shared__ int rotated;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  for (int i = tid; i < WR * WL; i += nt) {
    const int r = i % WR, j = i / WR;  // coalesced read of column j of X
    double v = 0.0;
    if (r < nv && j < len && !(d.mode == 1 && r > j)) v = X[r + (int64_t)p * j] * scale_in;
    B[r * WL + j] = v;
  }
  for (int i = tid; i < WR * WR; i += nt) J[i] = (i / WR == i % WR) ? 1.0 : 0.0;
  __syncthreads();
  const int nvp = (nv + 1) & ~1;
  const int npairs = nvp / 2, rounds = nvp - 1;
  const double tol2 = (double)max(len, 2) * 2.220446049250313e-16 * 2.220446049250313e-16;
  double* sq = sig;
  bool conv = (nv < 2);
  for (int sweep = 0; sweep < s.max_sweeps && !conv; ++sweep) {
    for (int v = wid; v < nv; v += nw) {
      double a0 = 0.0;
#pragma unroll
      for (int m = 0; m < EPL; ++m) { const double x = B[v * WL + lane + 32 * m]; a0 += x * x; }
      a0 = warp_sum(a0);
      sq[v] = a0;
    }
    rotated = 0;
    __syncthreads();
    for (int rd = 0; rd < rounds; ++rd) {
      for (int q = wid; q < npairs; q += nw) {
        int i, j;
        rr_pair(nvp, rd, q, i, j);
        if constexpr (WR == 32 && WL == 32) if (q + nw < npairs) {   // fewer warps than pairs: this warp's next pair rides along (two chains interleaved)
          int i2, j2;
          rr_pair(nvp, rd, q + nw, i2, j2);
          if (j < nv && j2 < nv) {
            if (jacobi_two_pairs<WL, WR>(B, J, sq, i, j, i2, j2, tol2, lane)) rotated = 1;
            q += nw;
            continue;
          }
        }
        if (j >= nv) continue;
        double* bi = B + i * WL + lane;
        double* bj = B + j * WL + lane;
        double* ji = J + i * WR + lane;
        double* jj = J + j * WR + lane;
        double x[EPL], y[EPL], u_[JPL], v_[JPL];
#pragma unroll
        for (int m = 0; m < EPL; ++m) { x[m] = bi[32 * m]; y[m] = bj[32 * m]; }
#pragma unroll
        for (int m = 0; m < JPL; ++m) { u_[m] = ji[32 * m]; v_[m] = jj[32 * m]; }
        const double al = sq[i], be = sq[j];
        double g0 = 0.0, g1 = 0.0;
#pragma unroll
        for (int m = 0; m < EPL; ++m) { if (m & 1) g1 += x[m] * y[m]; else g0 += x[m] * y[m]; }
        const double ga = warp_sum(g0 + g1);
        if (al > 0.0 && be > 0.0 && ga * ga > tol2 * al * be) {
          // same rotation as jacobi_pair (sweep.cu): cos 2th = |da|/h, sin 2th = sign(da) b2/h
          const double da = be - al, b2 = 2.0 * ga;
          const double h2 = da * da + b2 * b2;
          double r, das = da, b2s = b2;
          if (h2 > 1e-280 && h2 < 1e280) r = rsqrt_nr(h2);
          else {
            const double sc = 1.0 / fmax(fabs(da), fabs(b2));
            das = da * sc; b2s = b2 * sc;
            r = rsqrt(das * das + b2s * b2s);
          }
          const double c2 = fabs(das) * r, s2 = (da < 0.0 ? -b2s : b2s) * r;
          const double uu = 0.5 + 0.5 * c2;
          const double ic = rsqrt_nr(uu);
          const double cs = uu * ic, sn = 0.5 * s2 * ic;
#pragma unroll
          for (int m = 0; m < EPL; ++m) {
            bi[32 * m] = cs * x[m] - sn * y[m];
            bj[32 * m] = sn * x[m] + cs * y[m];
          }
#pragma unroll
          for (int m = 0; m < JPL; ++m) {
            ji[32 * m] = cs * u_[m] - sn * v_[m];
            jj[32 * m] = sn * u_[m] + cs * v_[m];
          }
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
  }
  if (!conv && tid == 0) atomicOr(c.status, ST_NOCONV);
  jacobi_finish(c, s, b, B, WL, nv, len, sig, srt);
  double* Bg = c.Bv + (int64_t)b * c.bv_stride;
  double* Jg = c.Jv + (int64_t)b * c.jv_stride;
  for (int i = tid; i < nv * len; i += nt) {
    const int e = i % len, v = i / len;
    Bg[(int64_t)v * c.ldv + e] = B[v * WL + e];
  }
  for (int i = tid; i < nv * nv; i += nt) {
    const int e = i % nv, v = i / nv;
    Jg[(int64_t)v * c.ldj + e] = J[v * WR + e];
  }
}

// ---------------------------------------------------------------------------------------------
// k_jacobi_g8<W> -- SVDs with k, n <= W (W = 32: config-5 batches; W = 64: the truncating steps of config 2)
// in the regime where a round is bound by INSTRUCTION ISSUE: thousands of trains filling every SM, or one
// train whose 32 pair-warps crowd one SM.  In k_jacobi_small a warp spends ~175 instructions on ONE pair,
// most of them the scalar rotation algebra that all 32 lanes execute redundantly.  Here a warp handles FOUR
// pairs of the round at once: 8 lanes per pair, lane e of a group holds W/8 consecutive elements of both rows
// (and of both rotation rows), the dot is a 3-level butterfly inside the group, and every instruction of the
// rotation algebra serves four different pairs.  Branch-free: a pair that needs no rotation skips only its
// stores.  4 W threads per train (W/2 pairs per round), row stride W + 4.  Same rotation, same convergence
// test, same ordering as k_jacobi_small; jacobi_finish (sort + rank rule on the device) unchanged.
// Reference: `svd(M)` inside the truncating functors, src/svd_trunc.jl:37,73,97,129.
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(4 * W, W == 32 ? 8 : 1) k_jacobi_g8(SweepCtx c, StepDesc s) {
  pdl_enter();
  constexpr int LD = W + 4, NT = 4 * W, EL = W / 8, E2 = EL / 2;
  const int b = blockIdx.x;
  const Dims d = get_dims(c, s, b);
  if (d.mode == 0) return;
  const int nv = d.k, len = d.n, p = d.p;   // host guarantees nv <= W, len <= W
  const double* X = xbuf(c, s.xin, b);
  const double scale_in = c.st[b].scale_in[s.slot];
  extern __shared__ __align__(16) double jsm[];
  double* B = jsm;                // [W][LD]
  double* J = B + W * LD;         // [W][LD]
  double* sig = J + W * LD;       // W
  double* srt = sig + W;          // W
  __shared__ int rotated;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < W * W; i += NT) {
    const int r = i % W, j = i / W;   // coalesced read of column j of X
    double v = 0.0;
    if (r < nv && j < len && !(d.mode == 1 && r > j)) v = X[r + (int64_t)p * j] * scale_in;
    B[r * LD + j] = v;
    J[r * LD + j] = (r == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  const int nvp = (nv + 1) & ~1;
  const int npairs = nvp / 2, rounds = nvp - 1;
  const double tol2 = (double)max(len, 2) * 2.220446049250313e-16 * 2.220446049250313e-16;
  double* sq = sig;
  // lane e of a group holds the element pairs {16 m + 2 e, 16 m + 2 e + 1}, m < W/16: the eight 16-byte accesses
  // of a group are contiguous (128 bytes, one conflict-free wavefront per quarter warp)
  const int grp = lane >> 3, e0 = (lane & 7) * 2;    // pair slot inside the warp, first element of this lane
  const int q = wid * 4 + grp;                       // pair index of the round handled by this lane group
  bool conv = (nv < 2);
  int sweep = 0;
  for (; sweep < s.max_sweeps && !conv; ++sweep) {
    if (tid < W) {   // cached squared norms, recomputed once per sweep (one row per thread, stride LD: conflict-free)
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
      if (!act) { i = 0; j = 0; }       // idle group: harmless loads, no stores
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
      // same rotation as k_jacobi_small: cos 2th = |da|/h, sin 2th = sign(da) b2/h
      const double da = be - al, b2 = 2.0 * ga;
      double h2 = da * da + b2 * b2, das = da, b2s = b2;
      if (rot && !(h2 > 1e-280 && h2 < 1e280)) {   // rare: rescale before squaring
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
  }
#ifdef TTB_JSYS_DEBUG
  if (b == 0 && tid == 0) printf("[g8] W=%d nv=%d len=%d sweeps=%d\n", W, nv, len, sweep);
#endif
  if (!conv && tid == 0) atomicOr(c.status, ST_NOCONV);
  jacobi_finish(c, s, b, B, LD, nv, len, sig, srt);
  double* Bg = c.Bv + (int64_t)b * c.bv_stride;
  double* Jg = c.Jv + (int64_t)b * c.jv_stride;
  for (int i = tid; i < nv * len; i += NT) {
    const int e = i % len, v2 = i / len;
    Bg[(int64_t)v2 * c.ldv + e] = B[v2 * LD + e];
  }
  for (int i = tid; i < nv * nv; i += NT) {
    const int e = i % nv, v2 = i / nv;
    Jg[(int64_t)v2 * c.ldj + e] = J[v2 * LD + e];
  }
}
__host__ __device__ constexpr size_t jg8_smem(int W) { return sizeof(double) * (2 * (size_t)W * (W + 4) + 2 * W); }

}  // namespace ttb
