// twovar_open.cuh -- twovar_marginals (src/abstract_tensor_train.jl:231-254) of OPEN trains with large bonds.
//
// For an open train d_1 = 1, so K_x = l[t-1] A_t(x) M[t,u] is a ROW VECTOR and stepping u -> u+1 is a
// vector-matrix product K_x <- K_x T_u (T_u = sum_x A_u(x), accumulate_M :119-137 streamed).  One CTA per first
// site walks that chain as a GEMV and re-reads the d x d matrix T_u for a handful of FLOPs per byte: 170 ms for
// the 32640 pairs of the config-3 train (generic kernel), 13 ms at d = 128 (k_twovar_fast).  Here a CTA owns a
// GROUP of consecutive first sites and stacks their K_x into one (group * Q) x d_u operand (up to 32 rows), so
// one pass over T_u serves the whole group:
//   * K <- K T_u is an FP64 tensor-core product (mma.sync m16n8k8): 32 x d_{u+1} x d_u, warp w owns the column
//     tiles w, w + 8, ..., A fragments from shared memory (row stride = 4 mod 32: conflict-free), B fragments
//     straight from global memory (T_u is used once per CTA; a fragment is eight fully used 32-byte sectors);
//   * b[t,u][x_t, x_u] = K_{t,x_t} . GT_u(x_u) (GT_u(x) = (A_u(x) r_{u+1})^T from k_twovar_pre) as warp dot products;
//   * every first site keeps its own max-abs scale (a common factor of b[t,u'], u' > u, which is normalised).
// Rows of first sites u has not reached yet hold their initial value l[t-1] A_t(x) and are left alone.
// Measured (B200, one train): config-3 size (L 256, d 256, q 2, 32640 pairs) 170 ms -> 7.5 ms, config-4 size
// (d 128, q 4) 13.2 -> 5.1 ms; an ablation puts 4.7 of the 7.5 ms in the product (18 us per step against a 4 us
// floor for streaming the 512 KB of T_u into one SM), the rest in L2 round trips of the small phases.
#pragma once

namespace ttb {

constexpr int kTvoRows = 32;
__host__ __device__ constexpr int tvo_ld(int dmax) { return ((dmax + 31) & ~31) + 4; }
__host__ __device__ constexpr size_t tvo_smem(int dmax, int Q) {
  return sizeof(double) * ((size_t)kTvoRows * tvo_ld(dmax) + kTvoRows * 8 + kTvoRows + 8 * kTvoRows + (size_t)Q * (dmax + 8));
}

// grid (groups, batch); group g owns first sites t0 + g tb .. t0 + g tb + tb - 1 (tb = 32 / Q)
__global__ void __launch_bounds__(256) k_twovar_open(TwoFast c, int tb, int t_end) {
  const int Q = c.Q, QQ = Q * Q;
  const int t_lo = c.t0 + (int)blockIdx.x * tb, b = blockIdx.y;
  if (t_lo >= t_end) return;
  const int nts = min(tb, t_end - t_lo);          // first sites of this group
  const int nrows = nts * Q;
  const int* bond = c.bond + (int64_t)b * c.L1;
  const double* tt = c.slab + (int64_t)b * c.tt_stride;
  const int LD = tvo_ld(c.dmax);
  extern __shared__ __align__(16) double tsm[];
  double* K = tsm;                                   // [32][LD], row r = (t - t_lo) Q + x
  double* bv = K + (size_t)kTvoRows * LD;            // [32][8]  b values of the step: row r, y
  double* rsc = bv + kTvoRows * 8;                   // [32] scale applied when the row is written back
  double* rmx = rsc + kTvoRows;                      // [8 warps][32 rows] max-abs of the new row over the warp's columns
  double* gts = rmx + 8 * kTvoRows;                  // [Q][dmax + 8] GT_u(y) of the current step
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, tq = lane & 3;
  for (int i = tid; i < kTvoRows * LD; i += 256) K[i] = 0.0;
  __syncthreads();
  // initial rows: K_{t,x} = l[t-1] A_t(x)  (1 x d_{t+1}); l[-1] = 1
  for (int o = wid; o < nrows * c.dmax; o += 8) {
    const int r = o / c.dmax, n = o - r * c.dmax;
    const int t = t_lo + r / Q, x = r - (r / Q) * Q;
    const int dl = bond[t], dr = bond[t + 1];
    if (n >= dr) continue;
    const double* A = tt + c.off[t] + (int64_t)dl * dr * x + (int64_t)dl * n;
    double acc = 0.0;
    if (t > 0) {
      const double* l = c.envL + (int64_t)b * c.env_stride + c.eoffL[t - 1];
      for (int i = lane; i < dl; i += 32) acc += l[i] * A[i];
      acc = warp_sum(acc);
    } else acc = A[0];
    if (lane == 0) K[r * LD + n] = acc;
  }
  __syncthreads();
  const int u_hi = min(c.L - 1, t_lo + nts - 1 + c.maxdist);
  const double* GTb = c.GT + (int64_t)b * c.L * Q * c.gstride;
  const double* Tb = c.Tm + (int64_t)b * c.L * c.tstride;
  for (int u = t_lo + 1; u <= u_hi; ++u) {
    const int du = bond[u], dn = u < c.L - 1 ? bond[u + 1] : 1;
    // ---- GT_u(y) into shared memory (one L2 round trip per step instead of one per dot), then
    //      b[t,u][x, y] for the rows whose first site t satisfies t < u <= t + maxdist
    const int gld = c.dmax + 8;
    for (int o = tid; o < Q * du; o += 256) {
      const int y = o / du, k = o - y * du;
      gts[y * gld + k] = GTb[((int64_t)u * Q + y) * c.gstride + k];
    }
    __syncthreads();
    for (int r = wid; r < nrows; r += 8) {
      const int t = t_lo + r / Q;
      if (!(t < u && u <= t + c.maxdist)) continue;
      const double* kr = K + r * LD;
      for (int y = 0; y < Q; ++y) {
        const double* gt = gts + y * gld;
        double acc = 0.0;
        for (int k = lane; k < du; k += 32) acc += kr[k] * gt[k];
        acc = warp_sum(acc);
        if (lane == 0) bv[r * 8 + y] = acc;
      }
    }
    // ---- K T_u for all 32 rows on the tensor cores (rows that are not active are not written back)
    const bool more = u < u_hi;
    const int ntile = (dn + 7) >> 3;
    double acc[2][4][4];
    if (more) {
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[a][j][e] = 0.0;
      const double* T = Tb + (int64_t)u * c.tstride;     // du x dn, ld du
      const int ksteps = (du + 7) >> 3;
      // B fragments come straight from global memory (L2): explicit register double buffering, two k-steps per
      // group, so one L2 round trip is in flight while the previous group's 16 DMMAs issue (a loop that loads and
      // multiplies in place serialised the round trips: 60 us per step at d = 256)
      const double* tcol[4];
      bool nok[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = (wid + 8 * j) * 8 + g;
        nok[j] = n < dn;
        tcol[j] = T + (int64_t)du * (nok[j] ? n : 0) + tq;
      }
      auto load_grp = [&](int ks0, double (&bf)[2][4][2]) {
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          const int k0 = (ks0 + s2) * 8 + tq;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            bf[s2][j][0] = (nok[j] && k0 < du) ? tcol[j][(ks0 + s2) * 8] : 0.0;
            bf[s2][j][1] = (nok[j] && k0 + 4 < du) ? tcol[j][(ks0 + s2) * 8 + 4] : 0.0;
          }
        }
      };
      auto mma_grp = [&](int ks0, const double (&bf)[2][4][2]) {
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          const int k0 = (ks0 + s2) * 8 + tq;
          double af[2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const double* ka = K + (mt * 16 + g) * LD + k0;
            af[mt][0] = ka[0]; af[mt][1] = ka[8 * LD]; af[mt][2] = ka[4]; af[mt][3] = ka[8 * LD + 4];
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            dmma_16x8x8(acc[0][j], af[0], bf[s2][j]);
            dmma_16x8x8(acc[1][j], af[1], bf[s2][j]);
          }
        }
      };
      double b0[2][4][2], b1[2][4][2];
      load_grp(0, b0);
      for (int ks = 0; ks < ksteps; ks += 4) {      // (K is zero beyond du up to the padded row: k-steps past ksteps add nothing)
        load_grp(ks + 2, b1);
        mma_grp(ks, b0);
        load_grp(ks + 4, b0);
        mma_grp(ks + 2, b1);
      }
      // per-first-site max-abs of the new rows
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          double m = 0.0;
#pragma unroll
          for (int j = 0; j < 4; ++j) m = fmax(m, fmax(fabs(acc[mt][j][2 * hh]), fabs(acc[mt][j][2 * hh + 1])));
          m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
          m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
          if (tq == 0) rmx[wid * kTvoRows + mt * 16 + g + 8 * hh] = m;
        }
    }
    __syncthreads();
    // ---- normalise and store b[t,u]; scales of the rows that move on
    if (tid < nts) {
      const int t = t_lo + tid;
      if (t < u && u <= t + c.maxdist) {
        double tot = 0.0;
        for (int x = 0; x < Q; ++x)
          for (int y = 0; y < Q; ++y) tot += bv[(tid * Q + x) * 8 + y];
        double* o = c.out + ((int64_t)b * c.npairs + c.pair_off[t] - (t + 1) + u) * QQ;
        for (int x = 0; x < Q; ++x)
          for (int y = 0; y < Q; ++y) o[x + Q * y] = bv[(tid * Q + x) * 8 + y] / tot;
      }
      double m = 0.0;
      if (more)
        for (int x = 0; x < Q; ++x)
          for (int w2 = 0; w2 < 8; ++w2) m = fmax(m, rmx[w2 * kTvoRows + tid * Q + x]);
      const double inv = (m > 0.0 && !isinf(m)) ? (m < 1e-290 ? 1e290 : 1.0 / m) : 1.0;
      for (int x = 0; x < Q; ++x) rsc[tid * Q + x] = (t < u && u < t + c.maxdist) ? inv : 0.0;   // 0: row not written back
    }
    __syncthreads();
    if (more) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int r = mt * 16 + g + 8 * hh;
          const double sc = r < nrows ? rsc[r] : 0.0;
          if (sc != 0.0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = (wid + 8 * j) * 8 + 2 * tq;
              if (n < ntile * 8) {
                K[r * LD + n] = acc[mt][j][2 * hh] * sc;
                K[r * LD + n + 1] = acc[mt][j][2 * hh + 1] * sc;
              }
            }
          }
        }
    }
    __syncthreads();
  }
}

}  // namespace ttb
