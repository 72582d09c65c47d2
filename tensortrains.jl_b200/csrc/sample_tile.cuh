// sample_tile.cuh -- k_sample_tile: hierarchical sampling of an OPEN train, a tile of draws per CTA.
//
// Reference: sample! src/abstract_tensor_train.jl:355-381 (one rand per site, inverse CDF utils.jl:9-20).
// Per site t and draw s the work is  p_s[x] = q_s . G_t(x),  draw x_s,  q_s <- q_s A_t(x_s) / p_s[x_s].
// The last product is 2 d^2 flop per draw and site -- the whole cost -- and it is a GEMM only if draws
// that chose the same x are multiplied together.  So each CTA keeps SMAX running vectors in shared
// memory, counting-sorts them by the value just drawn into x-pure tiles of 16 rows, and every warp runs
// one 16 x DP x DP tile through FP64 tensor cores (mma.sync.m16n8k8.f64, SASS DMMA) with the rows
// gathered through the sort permutation; the result is written back in place (tiles own disjoint rows).
// A_t is pre-packed once per call as [site][k-chunk of 8][x][n][8] so that one 1-D TMA bulk copy
// (cp.async.bulk + mbarrier, SASS UBLKCP) brings the chunk of ALL x values a K step needs; two stages.
// The running vector is renormalised by 1/p_s[x_s] each site, so log p(draw) = sum_t log(p[x_t]/sum p)
// accumulates exactly the reference's tr(Q)/z and long chains cannot overflow.
#pragma once

namespace ttb {

#define ST_SMAX 96          // draw slots per CTA (6 tiles of 16); sorted tiles <= 6 + (Q-1)
#define ST_MAXTILES 16
#define ST_THREADS 256      // 8 warps

struct TileCtx {
  const double* Ap;    // [L][NKC][Q][DP][8]   zero padded
  const double* Gp;    // [L][DP][8]           G_t(m, x), zero padded
  int L, Q;
  const double* uniforms; uint64_t seed, first;
  uint8_t* x_out; double* p_out; unsigned long long* hist; double* sum_logp;
  int* status;
  int64_t nsamples;
};

template <int DP>
__global__ void __launch_bounds__(ST_THREADS, 1) k_sample_tile(TileCtx c) {
  constexpr int NKC = DP / 8;       // K chunks per site
  constexpr int NT = DP / 8;        // n-tiles of 8 columns
  constexpr int LDQ = DP + 4;       // == 4 (mod 16): conflict-free A fragments for in-order rows
  const int Q = c.Q;
  const int stage_doubles = Q * DP * 8;
  extern __shared__ __align__(128) double sm[];
  double* Qs = sm;                                   // (SMAX+1) x LDQ, last row = zeros
  double* Bst = Qs + (ST_SMAX + 1) * LDQ;            // 2 stages
  double* Gs = Bst + 2 * stage_doubles;              // DP x 8
  double* Ps = Gs + DP * 8;                          // SMAX x 8 weights
  double* scl = Ps + ST_SMAX * 8;                    // SMAX
  double* lp = scl + ST_SMAX;                        // SMAX log-probabilities
  int* xsel = reinterpret_cast<int*>(lp + ST_SMAX);  // SMAX
  int* order = xsel + ST_SMAX;                       // MAXTILES*16 positions -> slot (or SMAX = zero row)
  int* tile_x = order + ST_MAXTILES * 16;            // MAXTILES
  int* misc = tile_x + ST_MAXTILES;                  // [0] = ntiles
  int* cnt_s = misc + 4;                             // [8] draws per value at this site
  int* rank_s = cnt_s + 8;                           // SMAX: arrival rank of the slot inside its value group
  unsigned int* hsm = reinterpret_cast<unsigned int*>(rank_s + ST_SMAX);  // L*Q histogram (stats mode)
  __shared__ __align__(8) uint64_t full[2];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;

  if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); }
  if (c.hist) for (int i = tid; i < c.L * Q; i += ST_THREADS) hsm[i] = 0u;
  __syncthreads();
  int64_t cons = 0;    // chunks consumed so far (identical in every thread)
  int64_t issued = 0;  // chunks issued so far (thread 0); never more than 2 ahead of cons

  const int64_t nbatches = (c.nsamples + ST_SMAX - 1) / ST_SMAX;
  for (int64_t batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
    const int64_t s0 = batch * ST_SMAX;
    const int nact = (int)min((int64_t)ST_SMAX, c.nsamples - s0);
    // initial state: q = e_0 (the 1 x 1 left boundary), zero row, log p = 0
    for (int i = tid; i < (ST_SMAX + 1) * LDQ; i += ST_THREADS) {
      const int r = i / LDQ, k = i % LDQ;
      Qs[i] = (k == 0 && r < nact) ? 1.0 : 0.0;
    }
    for (int i = tid; i < ST_SMAX; i += ST_THREADS) lp[i] = 0.0;
    for (int i = tid; i < ST_MAXTILES * 16; i += ST_THREADS) order[i] = ST_SMAX;   // zero row
    if (tid < 8) cnt_s[tid] = 0;
    __syncthreads();

    for (int t = 0; t < c.L; ++t) {
      const double* Apt = c.Ap + (int64_t)t * NKC * stage_doubles;
      // start streaming this site's first two K chunks while the weights / draw / sort run
      if (tid == 0) {
        for (int kc = 0; kc < 2 && kc < NKC; ++kc) {
          const int buf = (int)(issued & 1);
          fence_proxy_async();
          mbar_expect_tx(&full[buf], (uint32_t)(stage_doubles * 8));
          bulk_g2s(Bst + (size_t)buf * stage_doubles, Apt + (int64_t)kc * stage_doubles, (uint32_t)(stage_doubles * 8),
                   &full[buf]);
          ++issued;
        }
      }
      {  // G_t -> smem
        const double* Gt = c.Gp + (int64_t)t * DP * 8;
        for (int i = tid; i < DP * 8; i += ST_THREADS) Gs[i] = Gt[i];
      }
      __syncthreads();
      // ---- weights p[s][x] = q_s . G(:, x): one m16n8k8 column block per 16-slot tile ----
      if (wid < ST_SMAX / 16) {
        const int r0 = wid * 16;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
        for (int kc = 0; kc < NKC; ++kc) {
          double af[4], bf[2];
          af[0] = Qs[(r0 + g) * LDQ + kc * 8 + t4];
          af[1] = Qs[(r0 + g + 8) * LDQ + kc * 8 + t4];
          af[2] = Qs[(r0 + g) * LDQ + kc * 8 + t4 + 4];
          af[3] = Qs[(r0 + g + 8) * LDQ + kc * 8 + t4 + 4];
          bf[0] = Gs[(kc * 8 + t4) * 8 + g];
          bf[1] = Gs[(kc * 8 + t4 + 4) * 8 + g];
          dmma_16x8x8(acc, af, bf);
        }
        Ps[(r0 + g) * 8 + 2 * t4] = acc[0];
        Ps[(r0 + g) * 8 + 2 * t4 + 1] = acc[1];
        Ps[(r0 + g + 8) * 8 + 2 * t4] = acc[2];
        Ps[(r0 + g + 8) * 8 + 2 * t4 + 1] = acc[3];
        __syncwarp();
        // ---- draw (lanes 0..15 own the 16 slots of this tile) ----
        if (lane < 16) {
          const int sl = r0 + lane;
          int xs = Q;  // sentinel: inactive slot
          if (sl < nact) {
            double tot = 0.0;
            bool neg = false;
            for (int x = 0; x < Q; ++x) { const double pv = Ps[sl * 8 + x]; neg |= pv < 0.0; tot += pv; }
            xs = Q - 1;
            if (neg) { atomicOr(c.status, ST_NEGATIVE); xs = 0; }
            else {
              const int64_t gs = s0 + sl;
              const double u = c.uniforms ? c.uniforms[gs * c.L + t] : philox_uniform(c.seed, c.first + (uint64_t)gs, (uint32_t)t);
              double cw = 0.0;
              for (int x = 0; x < Q; ++x) {
                cw += Ps[sl * 8 + x] / tot;
                if (cw > u) { xs = x; break; }
              }
            }
            const double px = Ps[sl * 8 + xs];
            scl[sl] = (px > 0.0) ? 1.0 / px : 0.0;
            lp[sl] += log(px / tot);
            if (c.x_out) c.x_out[(s0 + sl) * c.L + t] = (uint8_t)xs;
            if (c.hist) atomicAdd(&hsm[t * Q + xs], 1u);
            rank_s[sl] = atomicAdd(&cnt_s[xs], 1);   // position inside the value group: any order will do
          }
          xsel[sl] = xs;
        }
      }
      __syncthreads();
      // ---- counting sort by x into x-pure tiles of 16 positions: the draw left the group sizes in cnt_s and
      //      every slot's rank inside its group, so each thread places its slot itself (no serial pass) ----
      int ntiles;
      {
        int start[8], pos = 0;
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          start[x] = pos;
          pos += (x < Q) ? ((cnt_s[x] + 15) >> 4) << 4 : 0;
        }
        ntiles = pos >> 4;
        if (tid < ST_SMAX) {
          const int xv = xsel[tid];
          if (xv < Q) {
            int st = 0;
#pragma unroll
            for (int x = 0; x < 8; ++x) st = (x == xv) ? start[x] : st;
            order[st + rank_s[tid]] = tid;
          }
        }
        if (tid < ntiles) {
          int tx = 0;
#pragma unroll
          for (int x = 1; x < 8; ++x) tx += (x < Q && tid * 16 >= start[x]) ? 1 : 0;   // start[] is non-decreasing
          tile_x[tid] = tx;
        }
      }
      __syncthreads();
      const int rounds = (ntiles + 7) / 8;
      // ---- q_s <- q_s A_t(x_s) * scale_s : one x-pure 16 x DP tile per warp and round ----
      for (int rd = 0; rd < rounds; ++rd) {
        const int ti = rd * 8 + wid;
        const bool act = ti < ntiles;
        const int x = act ? tile_x[ti] : 0;
        const int rowa = act ? order[ti * 16 + g] : ST_SMAX;
        const int rowb = act ? order[ti * 16 + g + 8] : ST_SMAX;
        double acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[nt][e] = 0.0;
#pragma unroll 1
        for (int kc = 0; kc < NKC; ++kc) {
          // chunks are produced and consumed in one global order: stage = cons & 1, and the j-th use of
          // a stage completes phase j of its mbarrier
          const int buf = (int)(cons & 1);
          mbar_wait(&full[buf], (uint32_t)((cons >> 1) & 1));
          ++cons;
          if (act) {
            const double* Bx = Bst + (size_t)buf * stage_doubles + (size_t)x * DP * 8;
            double af[4];
            af[0] = Qs[rowa * LDQ + kc * 8 + t4];
            af[1] = Qs[rowb * LDQ + kc * 8 + t4];
            af[2] = Qs[rowa * LDQ + kc * 8 + t4 + 4];
            af[3] = Qs[rowb * LDQ + kc * 8 + t4 + 4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              double bf[2];
              bf[0] = Bx[(nt * 8 + g) * 8 + t4];
              bf[1] = Bx[(nt * 8 + g) * 8 + t4 + 4];
              dmma_16x8x8(acc[nt], af, bf);
            }
          }
          __syncthreads();  // everyone is done with this stage
          if (tid == 0) {
            // refill the stage just released with the chunk two steps ahead (same site, maybe next round)
            const int nxt = rd * NKC + kc + 2;
            if (nxt < rounds * NKC) {
              const int kcn = nxt % NKC;
              fence_proxy_async();
              mbar_expect_tx(&full[buf], (uint32_t)(stage_doubles * 8));
              bulk_g2s(Bst + (size_t)buf * stage_doubles, Apt + (int64_t)kcn * stage_doubles, (uint32_t)(stage_doubles * 8),
                       &full[buf]);
              ++issued;
            }
          }
        }
        if (act) {
          const double sa = rowa < ST_SMAX ? scl[rowa] : 0.0;
          const double sb = rowb < ST_SMAX ? scl[rowb] : 0.0;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            if (rowa < ST_SMAX) {
              Qs[rowa * LDQ + nt * 8 + 2 * t4] = acc[nt][0] * sa;
              Qs[rowa * LDQ + nt * 8 + 2 * t4 + 1] = acc[nt][1] * sa;
            }
            if (rowb < ST_SMAX) {
              Qs[rowb * LDQ + nt * 8 + 2 * t4] = acc[nt][2] * sb;
              Qs[rowb * LDQ + nt * 8 + 2 * t4 + 1] = acc[nt][3] * sb;
            }
          }
        }
      }
      __syncthreads();
      // next site's sort starts from an all-padding order and empty groups
      for (int i = tid; i < ST_MAXTILES * 16; i += ST_THREADS) order[i] = ST_SMAX;
      if (tid < 8) cnt_s[tid] = 0;
    }
    // ---- results of this batch ----
    for (int sl = tid; sl < nact; sl += ST_THREADS) {
      if (c.p_out) c.p_out[s0 + sl] = exp(lp[sl]);
    }
    if (c.sum_logp) {
      double v = 0.0;
      for (int sl = tid; sl < nact; sl += ST_THREADS) v += lp[sl];
      v = warp_sum(v);
      if (lane == 0 && v != 0.0) atomicAdd(c.sum_logp, v);
    }
    __syncthreads();
  }
  if (c.hist) {
    __syncthreads();
    for (int i = tid; i < c.L * Q; i += ST_THREADS)
      if (hsm[i]) atomicAdd(&c.hist[i], (unsigned long long)hsm[i]);
  }
}

}  // namespace ttb
