// twovar_fast.cuh -- twovar_marginals (src/abstract_tensor_train.jl:231-254) restructured so that the
// inner (t,u) loop is a Frobenius product and one small GEMM.
//
//   b[t,u][x_t,x_u] ∝ tr( r[u+1] l[t-1] A_t(x_t) M[t,u] A_u(x_u) )
//                    = < K_{x_t} , GT_u(x_u) >,   K_x  = l[t-1] A_t(x) M[t,u]        (r0 x d_u)
//                                                 GT_u(x) = (A_u(x) r[u+1])^T        (r0 x d_u)
//   and stepping u -> u+1 is K_x <- K_x T_u / max|.|,   T_u = sum_x A_u(x)  (accumulate_M, :119-137,
//   streamed: the L x L table of M matrices is never stored).
// GT_u(x) and T_u do not depend on t, so they are computed once per call (k_twovar_pre) instead of once
// per pair; the per-pair work drops from O(Q^2 r0 d^2) to Q^2 r0 d + Q r0 d^2.
#pragma once

namespace ttb {

struct TwoPre {
  const double* slab; int64_t tt_stride; const int64_t* off; const int* bond; int L, L1, Q;
  const double* envR; int64_t env_stride; const int64_t* eoffR;
  double* GT; double* Tm;          // [batch][L][Q][gstride], [batch][L][tstride]
  int64_t gstride, tstride;
  int padded;   // k_twovar_pre_small only: T_u as a zero-padded 32 x 32 (ld 32), GT_u(x) zero-padded to 32 columns
};

// one CTA per (site u, train): T_u and GT_u(x) for every x
__global__ void k_twovar_pre(TwoPre c) {
  const int u = blockIdx.x, b = blockIdx.y;
  const int* bond = c.bond + (int64_t)b * c.L1;
  const int dl = bond[u], dr = bond[u + 1], r0 = bond[0];
  const double* A = c.slab + (int64_t)b * c.tt_stride + c.off[u];
  const double* r = u < c.L - 1 ? c.envR + (int64_t)b * c.env_stride + c.eoffR[u + 1] : nullptr;  // dr x r0
  double* Tm = c.Tm + ((int64_t)b * c.L + u) * c.tstride;
  for (int i = threadIdx.x; i < dl * dr; i += blockDim.x) {
    double acc = 0.0;
    for (int x = 0; x < c.Q; ++x) acc += A[i + (int64_t)dl * dr * x];
    Tm[i] = acc;
  }
  for (int x = 0; x < c.Q; ++x) {
    double* GT = c.GT + (((int64_t)b * c.L + u) * c.Q + x) * c.gstride;  // r0 x dl, [a + r0*cc]
    const double* Ax = A + (int64_t)dl * dr * x;
    for (int o = threadIdx.x; o < r0 * dl; o += blockDim.x) {
      const int a = o % r0, cc = o / r0;
      double acc = 0.0;
      if (r) for (int e = 0; e < dr; ++e) acc += Ax[cc + (int64_t)dl * e] * r[e + (int64_t)dr * a];
      else acc = Ax[cc + (int64_t)dl * a];  // r = I (dr == r0)
      GT[o] = acc;
    }
  }
}

struct TwoFast {
  const double* slab; int64_t tt_stride; const int64_t* off; const int* bond; int L, L1, Q;
  const double* envL; int64_t env_stride; const int64_t* eoffL;
  const double* GT; const double* Tm; int64_t gstride, tstride;
  const int64_t* pair_off; int maxdist;
  double* out; int64_t npairs;
  int kcap;   // doubles reserved per K_x buffer (r0max * dmax)
  int dmax;
  int t0;     // first site of the range this launch covers (blockIdx.x = t - t0)
};

// one CTA per (t, train): walks u = t+1 .. t+maxdist
__global__ void __launch_bounds__(256) k_twovar_fast(TwoFast c) {
  const int t = blockIdx.x + c.t0, b = blockIdx.y;
  if (t >= c.L - 1) return;
  const int* bond = c.bond + (int64_t)b * c.L1;
  const double* tt = c.slab + (int64_t)b * c.tt_stride;
  const int r0 = bond[0], Q = c.Q;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  extern __shared__ double sm[];
  double* K = sm;                          // Q * kcap
  double* K2 = K + (size_t)Q * c.kcap;     // Q * kcap
  double* Ts = K2 + (size_t)Q * c.kcap;    // dmax * dmax
  double* Gs = Ts + (size_t)c.dmax * c.dmax;  // Q * kcap
  double* red = Gs + (size_t)Q * c.kcap;   // 8 warps * (Q*Q + 1)
  double* bq = red + 8 * (Q * Q + 1);      // Q*Q
  // K_x = l[t-1] A_t(x)   (r0 x d_{t+1})
  {
    const int dl = bond[t], dr = bond[t + 1];
    const double* A = tt + c.off[t];
    const double* l = t > 0 ? c.envL + (int64_t)b * c.env_stride + c.eoffL[t - 1] : nullptr;  // r0 x dl
    for (int o = tid; o < Q * r0 * dr; o += 256) {
      const int a = o % r0, cc = (o / r0) % dr, x = o / (r0 * dr);
      double acc = 0.0;
      if (l) for (int i = 0; i < dl; ++i) acc += l[a + (int64_t)r0 * i] * A[i + (int64_t)dl * (cc + (int64_t)dr * x)];
      else acc = A[a + (int64_t)dl * (cc + (int64_t)dr * x)];
      K[(size_t)x * c.kcap + a + r0 * cc] = acc;
    }
  }
  __syncthreads();
  const int uhi = min(c.L - 1, t + c.maxdist);
  for (int u = t + 1; u <= uhi; ++u) {
    const int dl = bond[u], dr = bond[u + 1];
    const int nk = r0 * dl;
    // stage GT_u(x) (and T_u when another step follows)
    for (int x = 0; x < Q; ++x) {
      const double* GT = c.GT + (((int64_t)b * c.L + u) * Q + x) * c.gstride;
      for (int i = tid; i < nk; i += 256) Gs[(size_t)x * c.kcap + i] = GT[i];
    }
    if (u < uhi) {
      const double* Tm = c.Tm + ((int64_t)b * c.L + u) * c.tstride;
      for (int i = tid; i < dl * dr; i += 256) Ts[i] = Tm[i];
    }
    __syncthreads();
    // b[xt, xu] = <K_xt, GT_xu>: all Q*Q partial sums per thread, one combined reduction
    for (int pq = 0; pq < Q * Q; ++pq) {
      const int xt = pq % Q, xu = pq / Q;
      const double* kx = K + (size_t)xt * c.kcap;
      const double* gx = Gs + (size_t)xu * c.kcap;
      double acc = 0.0;
      for (int i = tid; i < nk; i += 256) acc += kx[i] * gx[i];
      acc = warp_sum(acc);
      if (lane == 0) red[wid * (Q * Q + 1) + pq] = acc;
    }
    __syncthreads();
    if (tid < Q * Q) {
      double v = 0.0;
      for (int w2 = 0; w2 < 8; ++w2) v += red[w2 * (Q * Q + 1) + tid];
      bq[tid] = v;
    }
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int i = 0; i < Q * Q; ++i) tot += bq[i];
      double* o = c.out + ((int64_t)b * c.npairs + c.pair_off[t] + (u - t - 1)) * Q * Q;
      for (int i = 0; i < Q * Q; ++i) o[i] = bq[i] / tot;
    }
    if (u < uhi) {
      // K2_x = K_x T_u ; rescale by the max-abs over all x (a common factor of b[t,u'] for u' > u)
      double mx = 0.0;
      for (int o = tid; o < Q * r0 * dr; o += 256) {
        const int a = o % r0, e = (o / r0) % dr, x = o / (r0 * dr);
        const double* kx = K + (size_t)x * c.kcap;
        double acc = 0.0;
        for (int cc = 0; cc < dl; ++cc) acc += kx[a + r0 * cc] * Ts[cc + dl * e];
        K2[(size_t)x * c.kcap + a + r0 * e] = acc;
        mx = fmax(mx, fabs(acc));
      }
      mx = warp_max(mx);
      __syncthreads();
      if (lane == 0) red[wid] = mx;
      __syncthreads();
      double m = 0.0;
      for (int w2 = 0; w2 < 8; ++w2) m = fmax(m, red[w2]);
      if (m > 0.0 && !isinf(m)) {
        for (int x = 0; x < Q; ++x)
          for (int i = tid; i < r0 * dr; i += 256) K2[(size_t)x * c.kcap + i] /= m;
      }
      __syncthreads();
      double* tmp = K; K = K2; K2 = tmp;
    }
    __syncthreads();
  }
}

}  // namespace ttb
