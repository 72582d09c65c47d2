// jacobi_sys.cuh -- k_jacobi_sys<W>: one-sided Jacobi for k, n <= W (W = 32, 64) with the matrix RESIDENT IN
// REGISTERS for the whole SVD.
//
// k_jacobi_g8 keeps B and J in shared memory and every round reads and writes every row of both: 128 KiB per round
// at W = 64, i.e. ~1100 cycles of shared-memory bandwidth under a ~600-cycle rotation chain (measured 1400 cycles
// per round, 361 us per 64 x 64 SVD -- 46 of the 52 ms of config 2's compress!).  Here W/2 "processors" of 8 lanes
// each own TWO row slots (top, bottom): both rows of B and of the accumulated rotations J live in the registers of
// the group (W/8 elements of each row per lane), the rotation of a round is register-local (dot = W/8 FMAs + a
// 3-level butterfly inside the group), and the rows then move one processor up / down in the Brent-Luk systolic
// order (top_0 stays, top_1 <- bot_0, top_k <- top_{k-1}, bot_k <- bot_{k+1}, bot_last <- top_last): neighbours in
// the same warp exchange with shuffles, only the two rows that cross a warp boundary go through shared memory
// (32 KiB per round at W = 64 instead of 256).  Every row carries its cached squared norm and its original index;
// after convergence the rows return to their own places in shared memory and the unchanged epilogue (sort + rank
// rule on the device, jacobi_finish) takes over.  Same rotation, same threshold, same stopping rule as the other
// Jacobi kernels; the visiting order is the same round-robin tournament written as a systolic array.
// Reference: `svd(M)` inside the truncating functors, src/svd_trunc.jl:37,73,97,129.
#pragma once

namespace ttb {

__host__ __device__ constexpr size_t jsys_smem(int W) {
  // B, J (epilogue / staging), sig, srt, exchange: [2 parities][W/8 warps][2 directions][8 lanes][2 W/8 + 2]
  return sizeof(double) * (2 * (size_t)W * (W + 4) + 2 * W + 2 * (size_t)(W / 8) * 2 * 8 * (2 * (W / 8) + 2));
}

template <int W>
__global__ void __launch_bounds__(4 * W, W == 32 ? 4 : 1) k_jacobi_sys(SweepCtx c, StepDesc s) {
  pdl_enter();
  constexpr int LD = W + 4, NT = 4 * W, EL = W / 8, NW = W / 8, M = W / 2;   // M processors, NW warps
  constexpr int XS = 2 * EL + 2;                                             // doubles a lane ships per row
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
  double* xch = srt + W;          // [2][NW][2][8][XS]
  __shared__ int rotated;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < W * W; i += NT) {
    const int r = i % W, j = i / W;   // coalesced read of column j of X
    double v = 0.0;
    if (r < nv && j < len && !(d.mode == 1 && r > j)) v = X[r + (int64_t)p * j] * scale_in;
    B[r * LD + j] = v;
  }
  __syncthreads();
  const int grp = lane >> 3, e = lane & 7;
  const int k = wid * 4 + grp;                 // processor index, 0 .. M-1
  double tb[EL], tj[EL], bb[EL], bj[EL];       // top / bottom row: B and J elements e + 8 i
  int tid_row = 2 * k, bid_row = 2 * k + 1;
  double tn = 0.0, bn = 0.0;
#pragma unroll
  for (int i = 0; i < EL; ++i) {
    const int col = e + 8 * i;
    tb[i] = B[tid_row * LD + col]; bb[i] = B[bid_row * LD + col];
    tj[i] = col == tid_row ? 1.0 : 0.0; bj[i] = col == bid_row ? 1.0 : 0.0;
  }
  const double tol2 = (double)max(len, 2) * 2.220446049250313e-16 * 2.220446049250313e-16;
  const int nvp = (nv + 1) & ~1;
  const int rounds = max(nvp, 2) - 1;          // rows beyond nv are zero and never rotate: W - 1 rounds visit every pair,
  (void)rounds;                                // and the systolic order needs all of them to bring the rows back in phase
  bool conv = (nv < 2);
  int par = 0, sweep = 0;
  for (; sweep < s.max_sweeps && !conv; ++sweep) {
    // exact squared norms once per sweep (the rotations update them incrementally in between)
    {
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int i = 0; i < EL; ++i) { a0 += tb[i] * tb[i]; a1 += bb[i] * bb[i]; }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
      tn = a0; bn = a1;
    }
    if (tid == 0) rotated = 0;
    __syncthreads();
    for (int rd = 0; rd < W - 1; ++rd, par ^= 1) {
      // ---- rotation of (top, bottom), register-local
      double ga = 0.0;
#pragma unroll
      for (int i = 0; i < EL; ++i) ga += tb[i] * bb[i];
      ga += __shfl_xor_sync(0xffffffffu, ga, 4);
      ga += __shfl_xor_sync(0xffffffffu, ga, 2);
      ga += __shfl_xor_sync(0xffffffffu, ga, 1);
      const double al = tn, be = bn;
      const bool rot = al > 0.0 && be > 0.0 && ga * ga > tol2 * al * be;
      if (rot) {
        // same rotation as the other kernels (i = top, j = bottom): cos 2th = |da|/h, sin 2th = sign(da) b2/h
        const double da = be - al, b2 = 2.0 * ga;
        double h2 = da * da + b2 * b2, das = da, b2s = b2;
        if (!(h2 > 1e-280 && h2 < 1e280)) {   // rare: rescale before squaring
          const double sc = 1.0 / fmax(fabs(da), fabs(b2));
          das = da * sc; b2s = b2 * sc;
          h2 = das * das + b2s * b2s;
        }
        const double r = rsqrt_nr(h2);
        const double c2 = fabs(das) * r, s2 = (da < 0.0 ? -b2s : b2s) * r;
        const double uu = 0.5 + 0.5 * c2;
        const double ic = rsqrt_nr(uu);
        const double cs = uu * ic, sn = 0.5 * s2 * ic;
#pragma unroll
        for (int i = 0; i < EL; ++i) {
          const double x = tb[i], y = bb[i], u = tj[i], v = bj[i];
          tb[i] = cs * x - sn * y; bb[i] = sn * x + cs * y;
          tj[i] = cs * u - sn * v; bj[i] = sn * u + cs * v;
        }
        const double t = sn * ic * ga;
        tn = al - t;
        bn = be + t;
        if (e == 0 && ga * ga > kJacobiBig2 * al * be) rotated = 1;
      }
      // ---- systolic move.  Rows that cross a warp boundary: group 3 ships its top up, group 0 its bottom down
      double* xo = xch + (size_t)par * NW * 2 * 8 * XS + (size_t)wid * 2 * 8 * XS;
      if (grp == 3) {
        double* q = xo + e * XS;
#pragma unroll
        for (int i = 0; i < EL; ++i) { q[i] = tb[i]; q[EL + i] = tj[i]; }
        q[2 * EL] = tn; q[2 * EL + 1] = (double)tid_row;
      }
      if (grp == 0) {
        double* q = xo + 8 * XS + e * XS;
#pragma unroll
        for (int i = 0; i < EL; ++i) { q[i] = bb[i]; q[EL + i] = bj[i]; }
        q[2 * EL] = bn; q[2 * EL + 1] = (double)bid_row;
      }
      __syncthreads();
      // what every slot receives: new top from the processor below (k-1), new bottom from the one above (k+1)
      double ntb[EL], ntj[EL], nbb[EL], nbj[EL], ntn, nbn;
      int ntid, nbid;
#pragma unroll
      for (int i = 0; i < EL; ++i) {
        ntb[i] = __shfl_up_sync(0xffffffffu, tb[i], 8); ntj[i] = __shfl_up_sync(0xffffffffu, tj[i], 8);
        nbb[i] = __shfl_down_sync(0xffffffffu, bb[i], 8); nbj[i] = __shfl_down_sync(0xffffffffu, bj[i], 8);
      }
      ntn = __shfl_up_sync(0xffffffffu, tn, 8); ntid = __shfl_up_sync(0xffffffffu, tid_row, 8);
      nbn = __shfl_down_sync(0xffffffffu, bn, 8); nbid = __shfl_down_sync(0xffffffffu, bid_row, 8);
      if (grp == 0 && wid > 0) {            // top from the last group of the previous warp
        const double* q = xch + (size_t)par * NW * 2 * 8 * XS + (size_t)(wid - 1) * 2 * 8 * XS + e * XS;
#pragma unroll
        for (int i = 0; i < EL; ++i) { ntb[i] = q[i]; ntj[i] = q[EL + i]; }
        ntn = q[2 * EL]; ntid = (int)q[2 * EL + 1];
      }
      if (grp == 3 && wid < NW - 1) {       // bottom from the first group of the next warp
        const double* q = xch + (size_t)par * NW * 2 * 8 * XS + (size_t)(wid + 1) * 2 * 8 * XS + 8 * XS + e * XS;
#pragma unroll
        for (int i = 0; i < EL; ++i) { nbb[i] = q[i]; nbj[i] = q[EL + i]; }
        nbn = q[2 * EL]; nbid = (int)q[2 * EL + 1];
      }
      if (wid == 0) {                       // warp-uniform: top_1 <- bot_0 (one more shuffle set, all lanes), top_0 stays
        double sbb[EL], sbj[EL];
#pragma unroll
        for (int i = 0; i < EL; ++i) { sbb[i] = __shfl_up_sync(0xffffffffu, bb[i], 8); sbj[i] = __shfl_up_sync(0xffffffffu, bj[i], 8); }
        const double sbn = __shfl_up_sync(0xffffffffu, bn, 8);
        const int sbid = __shfl_up_sync(0xffffffffu, bid_row, 8);
        if (k == 1) {
#pragma unroll
          for (int i = 0; i < EL; ++i) { ntb[i] = sbb[i]; ntj[i] = sbj[i]; }
          ntn = sbn; ntid = sbid;
        } else if (k == 0) {
#pragma unroll
          for (int i = 0; i < EL; ++i) { ntb[i] = tb[i]; ntj[i] = tj[i]; }
          ntn = tn; ntid = tid_row;
        }
      }
      if (k == M - 1) {                     // bot_last <- top_last
#pragma unroll
        for (int i = 0; i < EL; ++i) { nbb[i] = tb[i]; nbj[i] = tj[i]; }
        nbn = tn; nbid = tid_row;
      }
#pragma unroll
      for (int i = 0; i < EL; ++i) { tb[i] = ntb[i]; tj[i] = ntj[i]; bb[i] = nbb[i]; bj[i] = nbj[i]; }
      tn = ntn; bn = nbn; tid_row = ntid; bid_row = nbid;
    }
    __syncthreads();
    conv = (rotated == 0);
    __syncthreads();
  }
#ifdef TTB_JSYS_DEBUG
  if (b == 0 && tid == 0) printf("[jsys] W=%d nv=%d len=%d sweeps=%d\n", W, nv, len, sweep);
#endif
  if (!conv && tid == 0) atomicOr(c.status, ST_NOCONV);
  // rows back to their own places: the epilogue reads rows 0 .. nv-1
#pragma unroll
  for (int i = 0; i < EL; ++i) {
    const int col = e + 8 * i;
    B[tid_row * LD + col] = tb[i]; J[tid_row * LD + col] = tj[i];
    B[bid_row * LD + col] = bb[i]; J[bid_row * LD + col] = bj[i];
  }
  __syncthreads();
  jacobi_finish(c, s, b, B, LD, nv, len, sig, srt);
  double* Bg = c.Bv + (int64_t)b * c.bv_stride;
  double* Jg = c.Jv + (int64_t)b * c.jv_stride;
  for (int i = tid; i < nv * len; i += NT) {
    const int e2 = i % len, v2 = i / len;
    Bg[(int64_t)v2 * c.ldv + e2] = B[v2 * LD + e2];
  }
  for (int i = tid; i < nv * nv; i += NT) {
    const int e2 = i % nv, v2 = i / nv;
    Jg[(int64_t)v2 * c.ldj + e2] = J[v2 * LD + e2];
  }
}

}  // namespace ttb
