// twovar_mma.cuh -- twovar_marginals (src/abstract_tensor_train.jl:231-254) for batches of small-bond
// trains (every bond <= 32, Q * d_1 <= 64: config 5), same restructuring as twovar_fast.cuh
//
//   b[t,u][x_t,x_u] ∝ < K_{x_t} , GT_u(x_u) >,   K_x <- K_x T_u / max|K|   (accumulate_M :119-137 streamed)
//
// but the per-pair product runs on the FP64 tensor cores.  The scalar kernel is bound by shared-memory
// bandwidth (two 8-byte loads per FMA: ~1 MB of LDS traffic per pair); here
//   * the Q matrices K_x are stacked into ONE (Q d_1) x d_u operand, column-major with a padded leading
//     dimension (68 = 4 mod 16) so that the m16n8k8 A fragments are bank-conflict free; T_u is staged with
//     leading dimension 36 for the B fragments; K <- K T_u is 8 DMMAs per warp, in place (the products
//     wait in registers across the barrier that carries the max-abs reduction);
//   * GT_u(x) is consumed exactly once per pair, so it never visits shared memory: the next pair's GT and
//     T are prefetched into registers while the current pair is computed;
//   * two block barriers per pair.
// One CTA per (t, train) walks u = t+1 .. t+maxdist.
#pragma once

namespace ttb {

// PAD: k_twovar_pre_small wrote T_u as a zero-padded 32 x 32 and GT_u(x) zero-padded to 32 columns, so the pair
// loop needs no dimension lookups, no predicates that change from pair to pair and no address arithmetic beyond
// two running pointers (ncu on the unpadded version: ~750 warp instructions per warp and pair, 384 of the
// kernel's 1760 SASS instructions were IMADs of address math).
template <int Q, bool PAD>
__global__ void __launch_bounds__(256, 4) k_twovar_mma(TwoFast c) {
  constexpr int LDK = 68, LDT = 36, QQ = Q * Q;
  constexpr int EPT = (8 + Q - 1) / Q;   // Frobenius elements per thread: d_1 <= 64 / Q, so d_1 d_u <= 256 EPT
  const int t = blockIdx.x + c.t0, b = blockIdx.y;
  if (t >= c.L - 1) return;
  const int* bond = c.bond + (int64_t)b * c.L1;
  const double* tt = c.slab + (int64_t)b * c.tt_stride;
  const int r0 = bond[0], R = Q * r0;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, tq = lane & 3;
  __shared__ double K[LDK * 32];
  __shared__ double Ts[LDT * 32];
  __shared__ double red[2][8][QQ];
  __shared__ double redm[8];
  extern __shared__ int sbond[];   // this train's bond dimensions (L + 1): keeps dependent global loads off the pair loop
  for (int i = tid; i < LDK * 32; i += 256) K[i] = 0.0;
  for (int i = tid; i < c.L1; i += 256) sbond[i] = bond[i];
  const int64_t out_base = ((int64_t)b * c.npairs + c.pair_off[t] - (t + 1)) * QQ;
  __syncthreads();
  // K_x = l[t-1] A_t(x)   (d_1 x d_{t+1}), stacked over x
  {
    const int dl = bond[t], dr = bond[t + 1];
    const double* A = tt + c.off[t];
    const double* l = t > 0 ? c.envL + (int64_t)b * c.env_stride + c.eoffL[t - 1] : nullptr;  // r0 x dl
    for (int o = tid; o < Q * r0 * dr; o += 256) {
      const int a = o % r0, cc = (o / r0) % dr, x = o / (r0 * dr);
      double acc = 0.0;
      if (l) for (int i = 0; i < dl; ++i) acc += l[a + (int64_t)r0 * i] * A[i + (int64_t)dl * (cc + (int64_t)dr * x)];
      else acc = A[a + (int64_t)dl * (cc + (int64_t)dr * x)];
      K[(x * r0 + a) + LDK * cc] = acc;
    }
  }
  // this thread's four elements (a, cc) of the Frobenius products: i = tid + 256 m = a + r0 cc
  int koff[EPT];
#pragma unroll
  for (int m = 0; m < EPT; ++m) {
    const int i = tid + 256 * m;
    koff[m] = (i % r0) + LDK * min(i / r0, 31);
  }
  const int uhi = min(c.L - 1, t + c.maxdist);
  double gq[EPT][Q], tp[4];
  const double* gbase = c.GT + (int64_t)b * c.L * Q * c.gstride + tid;
  const double* tbase = c.Tm + (int64_t)b * c.L * c.tstride + tid;
  const int gs = (int)c.gstride, tsd = (int)c.tstride, nkp = r0 * 32;
  auto prefetch_G = [&](int u) {
    const int nk = PAD ? nkp : r0 * sbond[u];
    const double* GT = gbase + (int64_t)u * Q * gs;
#pragma unroll
    for (int x = 0; x < Q; ++x) {
#pragma unroll
      for (int m = 0; m < EPT; ++m) gq[m][x] = (tid + 256 * m < nk) ? GT[x * gs + 256 * m] : 0.0;
    }
  };
  auto prefetch_T = [&](int u) {
    if (PAD) {
      const double* Tm = tbase + (int64_t)u * tsd;
#pragma unroll
      for (int m = 0; m < 4; ++m) tp[m] = Tm[256 * m];
    } else {
      const int dl = sbond[u], dr = sbond[u + 1];
      const double* Tm = tbase - tid + (int64_t)u * tsd;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int j = tid + 256 * m, cc = j & 31, e = j >> 5;
        tp[m] = (cc < dl && e < dr) ? Tm[cc + dl * e] : 0.0;
      }
    }
  };
  auto stage_T = [&]() {
#pragma unroll
    for (int m = 0; m < 4; ++m) { const int j = tid + 256 * m; Ts[(j & 31) + LDT * (j >> 5)] = tp[m]; }
  };
  prefetch_G(t + 1);
  if (t + 1 < uhi) { prefetch_T(t + 1); stage_T(); }
  __syncthreads();
  const int mt = wid & 3, nt0 = (wid >> 2) * 2;
  int par = 0;
  for (int u = t + 1; u <= uhi; ++u, par ^= 1) {
    const int nk = PAD ? nkp : r0 * sbond[u];
    const bool more = u < uhi;
    if (more && u + 1 < uhi) prefetch_T(u + 1);   // the next pair's T travels while this pair is computed
    // b[x_t, x_u] partial sums
    double acc[QQ];
#pragma unroll
    for (int i = 0; i < QQ; ++i) acc[i] = 0.0;
#pragma unroll
    for (int m = 0; m < EPT; ++m) {
      if (tid + 256 * m < nk) {
#pragma unroll
        for (int xt = 0; xt < Q; ++xt) {
          const double kx = K[koff[m] + xt * r0];
#pragma unroll
          for (int xu = 0; xu < Q; ++xu) acc[xt + Q * xu] += kx * gq[m][xu];
        }
      }
    }
    if constexpr (QQ == 4) {
      // four totals with 6 shuffles instead of 20: every level sends the half the partner keeps; afterwards
      // lanes 0, 8, 16, 24 hold the totals of values 0, 1, 2, 3
      const bool h16 = lane & 16, h8 = lane & 8;
      const double b0 = (h16 ? acc[2] : acc[0]) + __shfl_xor_sync(0xffffffffu, h16 ? acc[0] : acc[2], 16);
      const double b1 = (h16 ? acc[3] : acc[1]) + __shfl_xor_sync(0xffffffffu, h16 ? acc[1] : acc[3], 16);
      double cv = (h8 ? b1 : b0) + __shfl_xor_sync(0xffffffffu, h8 ? b0 : b1, 8);
      cv += __shfl_xor_sync(0xffffffffu, cv, 4);
      cv += __shfl_xor_sync(0xffffffffu, cv, 2);
      cv += __shfl_xor_sync(0xffffffffu, cv, 1);
      if ((lane & 7) == 0) red[par][wid][lane >> 3] = cv;
    } else {
#pragma unroll
      for (int i = 0; i < QQ; ++i) {
        acc[i] = warp_sum(acc[i]);
        if (lane == 0) red[par][wid][i] = acc[i];
      }
    }
    if (more) prefetch_G(u + 1);                   // ... and its GT, consumed one full step from now
    double d[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
    if (more) {
      // K T_u on the tensor cores: warp = (16-row tile mt) x (two 8-column tiles nt0, nt0+1), K = 32
      if (mt * 16 < R) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const double* ka = K + (mt * 16 + g) + LDK * (ks * 8 + tq);
          const double af[4] = {ka[0], ka[8], ka[4 * LDK], ka[8 + 4 * LDK]};
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const double* tb = Ts + (ks * 8 + tq) + LDT * ((nt0 + j) * 8 + g);
            const double bf[2] = {tb[0], tb[4]};
            dmma_16x8x8(d[j], af, bf);
          }
        }
      }
      double mx = 0.0;
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) mx = fmax(mx, fabs(d[j][i]));
      mx = warp_max(mx);
      if (lane == 0) redm[wid] = mx;
    }
    __syncthreads();
    if (wid == (u & 7)) {
      double v = 0.0;
      if (lane < QQ) {
#pragma unroll
        for (int w2 = 0; w2 < 8; ++w2) v += red[par][w2][lane];
      }
      const double tot = warp_sum(v);
      if (lane < QQ) c.out[out_base + (int64_t)u * QQ + lane] = v / tot;
    }
    if (more) {
      double m = 0.0;
#pragma unroll
      for (int w2 = 0; w2 < 8; ++w2) m = fmax(m, redm[w2]);
      // rescale by the max-abs over all x (a common factor of b[t,u'] for u' > u, which is normalised)
      const double inv = (m > 0.0 && !isinf(m)) ? (m < 1e-290 ? 1e290 : 1.0 / m) : 1.0;
      if (mt * 16 < R) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          double* ko = K + (mt * 16 + g) + LDK * ((nt0 + j) * 8 + 2 * tq);
          ko[0] = d[j][0] * inv; ko[LDK] = d[j][1] * inv;
          ko[8] = d[j][2] * inv; ko[8 + LDK] = d[j][3] * inv;
        }
      }
      if (u + 1 < uhi) stage_T();
    }
    __syncthreads();
  }
}

}  // namespace ttb
