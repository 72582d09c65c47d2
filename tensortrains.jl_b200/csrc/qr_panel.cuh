// qr_panel.cuh -- k_qr_panel_flag: register-resident Householder panel (32 columns, <= 512 rows) whose
// warps hand reflectors to each other through shared-memory flags instead of block barriers.
//
// Why (r1 ncu profiles of the barrier version): the panel is a chain of 32 dependent columns; with a
// __syncthreads per column every column paid owner latency + the slowest warp + the barrier, ~2.4 us.
// Here all 32 reflectors stay in shared memory (no buffer reuse => no WAR hazard), the owner of column
// c publishes v_c and raises ready[c]; every warp consumes the reflectors in order at its own pace and
// the owner of column c+1 updates THAT column first, so the chain per column is one warp's
// norm -> rsqrt/rcp -> publish plus one dot+update.  Dots use 4 independent accumulators.
#pragma once

#ifndef TTB_PANEL_BACKOFF_NS
#define TTB_PANEL_BACKOFF_NS 200
#endif

namespace ttb {

__device__ __forceinline__ int ld_flag(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}

template <int RPL>
__device__ __forceinline__ double dot_rpl(const double (&v)[RPL], const double (&a)[RPL]) {
  double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
  for (int i = 0; i < RPL; i += 4) {
    d0 += v[i] * a[i];
    if (i + 1 < RPL) d1 += v[i + 1] * a[i + 1];
    if (i + 2 < RPL) d2 += v[i + 2] * a[i + 2];
    if (i + 3 < RPL) d3 += v[i + 3] * a[i + 3];
  }
  return (d0 + d1) + (d2 + d3);
}

// Householder reflector of column `cc` held as col[i] = row (lane + 32 i); overwrites the column with
// (R_cc on the diagonal, v below) and publishes v (unit diagonal, zeros above) to vcol.  Returns tau.
template <int RPL>
__device__ __forceinline__ double make_reflector(double (&col)[RPL], double* vcol, int cc, int lane) {
  double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    const int r = lane + 32 * i;
    const double x = (r > cc) ? col[i] : 0.0;
    if ((i & 3) == 0) p0 += x * x;
    else if ((i & 3) == 1) p1 += x * x;
    else if ((i & 3) == 2) p2 += x * x;
    else p3 += x * x;
  }
  const double xn2 = warp_sum((p0 + p1) + (p2 + p3));
  const double alpha = __shfl_sync(0xffffffffu, col[0], cc);
  double tau = 0.0, scal = 0.0, beta = alpha;
  if (xn2 > 0.0) {
    const double nn = alpha * alpha + xn2;
    const double rs = rsqrt(nn);  // 1/||x||
    const double nrm = nn * rs;
    const double sg = alpha >= 0.0 ? 1.0 : -1.0;
    beta = -sg * nrm;
    const double den = alpha - beta;  // = sg*(|alpha| + nrm), no cancellation
    scal = 1.0 / den;
    tau = den * sg * rs;  // (beta-alpha)/beta = den/(sg*nrm)
  }
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    const int r = lane + 32 * i;
    double v = 0.0;
    if (r > cc) { v = col[i] * scal; col[i] = v; }
    else if (r == cc) { v = 1.0; col[i] = beta; }
    vcol[r] = v;
  }
  return tau;
}

template <int RPL>
__global__ void __launch_bounds__(512, 1) k_qr_panel_flag(SweepCtx c, StepDesc s, int panel) {
  const int b = blockIdx.x;
  const Dims d = get_dims(c, s, b);
  if (d.mode == 2) return;
  const int j0 = panel * 32;
  if (panel == 0 && d.mode == 0 && threadIdx.x == 0) c.st[b].mprime = d.k;
  if (j0 >= d.k) return;
  const int bwc = min(32, d.k - j0);
  const int rows = d.p - j0;  // host guarantees rows <= 32*RPL
  const int64_t p = d.p;
  double* X = xbuf(c, s.xin, b) + j0 + p * j0;  // panel origin
  extern __shared__ __align__(16) double dyn[];
  double* vs = dyn;                       // [32][32*RPL] all reflectors of the panel
  double* G = vs + 32 * 32 * RPL;         // 32*32
  double* Tm = G + 1024;                  // 32*32
  double* sh_tau = Tm + 1024;             // 32
  int* ready = reinterpret_cast<int*>(sh_tau + 32);  // 32 flags
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;  // 16 warps
  constexpr int LDV = 32 * RPL;
  double a[2][RPL];
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2) {
    const int cs = wid + 16 * s2;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      const int r = lane + 32 * i;
      a[s2][i] = (cs < bwc && r < rows) ? X[r + p * cs] : 0.0;
    }
  }
  for (int i = tid; i < 1024; i += 512) { G[i] = 0.0; Tm[i] = 0.0; }
  if (tid < 32) ready[tid] = 0;
  __syncthreads();
#ifdef TTB_PANEL_STAMPS
  if (tid == 0) TTB_PANEL_STAMPS[0] = clock64();
#endif

  // apply reflector ccv (already in registers as v) to one of this warp's two columns
  auto apply_slot = [&](int ccv, const double (&v)[RPL], double tau, int s2) {
    const int cs = wid + 16 * s2;
    if (cs < bwc && cs != ccv) {  // warp-uniform
      double dot = s2 ? dot_rpl<RPL>(v, a[1]) : dot_rpl<RPL>(v, a[0]);
      dot = warp_sum(dot);
      if (cs > ccv) {
        if (tau != 0.0) {
          const double w = tau * dot;
          if (s2) {
#pragma unroll
            for (int i = 0; i < RPL; ++i) a[1][i] -= v[i] * w;
          } else {
#pragma unroll
            for (int i = 0; i < RPL; ++i) a[0][i] -= v[i] * w;
          }
        }
      } else if (lane == 0) {
        G[cs + 32 * ccv] = dot;  // V_cs^T v_ccv for the T factor
      }
    }
  };
  bool pending = false;  // this warp still owes reflector cc-1 to its non-priority column
  int pend_slot = 0;
#pragma unroll 1
  for (int cc = 0; cc < bwc; ++cc) {
    const int ow = cc & 15;
    const bool os = (cc >> 4) != 0;
    if (wid == ow) {
      // ---- owner: form reflector cc from its (fully updated) column, in place ----
      double* vcol = vs + (size_t)cc * LDV;
      const double tau = os ? make_reflector<RPL>(a[1], vcol, cc, lane) : make_reflector<RPL>(a[0], vcol, cc, lane);
      if (lane == 0) sh_tau[cc] = tau;
      // No MEMBAR here: shared-memory stores of one warp are performed in order, __syncwarp orders the
      // lanes, and the flag is the last store (a CTA fence with 16 STS in flight cost ~600 cycles per
      // column in the micro-benchmark, profiles/micro_panel.cu).
      __syncwarp();
      if (lane == 0) *reinterpret_cast<volatile int*>(&ready[cc]) = 1;
#ifdef TTB_PANEL_STAMPS
      if (lane == 0) TTB_PANEL_STAMPS[8 + cc] = clock64();
#endif
    } else {
      // ---- consumer: wait for reflector cc ----
      if (lane == 0) {
        unsigned it = 0;
        const bool hot = (wid == ((cc + 1) & 15));  // next link of the chain polls tightly
        while (ld_flag(&ready[cc]) == 0) {
          if (!hot) __nanosleep(TTB_PANEL_BACKOFF_NS);  // keep the other pollers off the shared-memory pipe
          if (++it > (1u << 26)) __trap();
        }
      }
      __syncwarp();
    }
    double v[RPL];
    if (pending) {  // catch up the column that was skipped while this warp hurried to publish
      const double taup = *reinterpret_cast<volatile double*>(&sh_tau[cc - 1]);
      const double* vp = vs + (size_t)(cc - 1) * LDV;
#pragma unroll
      for (int i = 0; i < RPL; ++i) v[i] = vp[lane + 32 * i];
      apply_slot(cc - 1, v, taup, pend_slot);
      pending = false;
    }
    const double tau = *reinterpret_cast<volatile double*>(&sh_tau[cc]);
    {
      const double* vcol = vs + (size_t)cc * LDV;
#pragma unroll
      for (int i = 0; i < RPL; ++i) v[i] = vcol[lane + 32 * i];
    }
    // the slot that holds column cc+1 goes first: its owner is the next link of the chain and
    // defers its other column until after it has published reflector cc+1
    const int first = ((cc + 1) >> 4) & 1;
    apply_slot(cc, v, tau, first);
    if (cc + 1 < bwc && wid == ((cc + 1) & 15)) {
      pending = true;
      pend_slot = 1 - first;
    } else {
      apply_slot(cc, v, tau, 1 - first);
    }
  }
  __syncthreads();
#ifdef TTB_PANEL_STAMPS
  if (tid == 0) TTB_PANEL_STAMPS[1] = clock64();
#endif
  // No T factor is built (that serial 32-step recurrence cost ~31k cycles, a quarter of the kernel):
  // consumers solve with T^{-1} = diag(1/tau) + striu(V^T V) directly, so ship G with tau on the diagonal.
  if (tid < 32) G[tid + 32 * tid] = tid < bwc ? sh_tau[tid] : 0.0;
#ifdef TTB_PANEL_STAMPS
  if (tid == 0) TTB_PANEL_STAMPS[2] = clock64();
#endif
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2) {
    const int cs = wid + 16 * s2;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      const int r = lane + 32 * i;
      if (cs < bwc && r < rows) X[r + p * cs] = a[s2][i];
    }
  }
  __syncthreads();
#ifdef TTB_PANEL_STAMPS
  if (tid == 0) TTB_PANEL_STAMPS[3] = clock64();
#endif
  double* Tg = c.T + (int64_t)b * c.t_stride + (int64_t)panel * 1024;
  for (int i = tid; i < 1024; i += 512) Tg[i] = G[i];
}

}  // namespace ttb
