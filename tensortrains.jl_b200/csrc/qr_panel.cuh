// qr_panel.cuh -- k_qr_panel_flag: register-resident Householder panel (32 columns, <= 512 rows) whose
// warps hand reflectors to each other through shared-memory flags instead of block barriers.
//
// Layout: NW = 32/NS warps, warp w owns the NS panel columns w, w+NW, ...; lane l holds rows l+32i.
// The panel is a chain of 32 dependent columns: reflector c can only be formed after reflector c-1
// has been applied to column c.  All 32 reflectors stay in shared memory (no buffer reuse => no WAR
// hazard); the owner of column c publishes v_c and raises ready[c]; every warp consumes the reflectors
// in order at its own pace.  The owner of column c+1 applies v_c to THAT column first, forms and
// publishes v_{c+1}, and only then catches up on its other columns.  The V^T V products that define the
// compact-WY T factor (T^{-1} = diag(1/tau) + striu(V^T V)) fall out of the same dots and are shipped
// as they are: no T factor is ever built (its serial recurrence cost a quarter of the kernel).
// Measured building blocks (profiles/micro_lat.cu, B200): dependent DFMA 8, shfl+DADD 35, rsqrt ~80,
// 1/x ~80, LDS ~70 cycles.
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

template <int RPL, int NS>
__global__ void __launch_bounds__(1024 / NS, 1) k_qr_panel_flag(SweepCtx c, StepDesc s, int panel) {
  constexpr int NW = 32 / NS;
  constexpr int NT = 32 * NW;
  constexpr int LDV = 32 * RPL;
  const int b = blockIdx.x;
  const Dims d = get_dims(c, s, b);
  if (d.mode == 2) return;
  const int j0 = panel * 32;
  if (panel == 0 && d.mode == 0 && threadIdx.x == 0) c.st[b].mprime[s.slot] = d.k;
  if (j0 >= d.k) return;
  const int bwc = min(32, d.k - j0);
  const int rows = d.p - j0;  // host guarantees rows <= 32*RPL
  const int64_t p = d.p;
  double* X = xbuf(c, s.xin, b) + j0 + p * j0;  // panel origin
  extern __shared__ __align__(16) double dyn[];
  double* vs = dyn;                       // [32][32*RPL] all reflectors of the panel
  double* G = vs + 32 * LDV;              // 32*32: striu = V^T V, diagonal = tau
  double* sh_tau = G + 1024;              // 32
  int* ready = reinterpret_cast<int*>(sh_tau + 32);  // 32 flags
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double a[NS][RPL];
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) {
    const int cs = wid + NW * s2;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      const int r = lane + 32 * i;
      a[s2][i] = (cs < bwc && r < rows) ? X[r + p * cs] : 0.0;
    }
  }
  for (int i = tid; i < 1024; i += NT) G[i] = 0.0;
  if (tid < 32) ready[tid] = 0;
  __syncthreads();
#ifdef TTB_PANEL_STAMPS
  if (tid == 0) TTB_PANEL_STAMPS[0] = clock64();
#endif

  // apply reflector ccv (in registers as v) to this warp's column slot `slot` (runtime index resolved
  // by predication over the unrolled slots, so `a` stays in registers)
  auto apply_slot = [&](int ccv, const double (&v)[RPL], double tau, int slot) {
#pragma unroll
    for (int s2 = 0; s2 < NS; ++s2) {
      if (s2 != slot) continue;
      const int cs = wid + NW * s2;
      if (cs < bwc && cs != ccv) {  // warp-uniform
        const double dot = warp_sum(dot_rpl<RPL>(v, a[s2]));
        if (cs > ccv) {
          if (tau != 0.0) {
            const double w = tau * dot;
#pragma unroll
            for (int i = 0; i < RPL; ++i) a[s2][i] -= v[i] * w;
          }
        } else {
          G[cs + 32 * ccv] = dot;  // V_cs^T v_ccv (same value from every lane)
        }
      }
    }
  };
  bool pending = false;  // this warp still owes reflector cc-1 to its non-priority columns
  int pend_pri = 0;
#pragma unroll 1
  for (int cc = 0; cc < bwc; ++cc) {
    const int ow = cc % NW;
    const int os = cc / NW;
    if (wid == ow) {
      // ---- owner: form reflector cc from its (fully updated) column, in place ----
      double* vcol = vs + (size_t)cc * LDV;
      double tau = 0.0;
#pragma unroll
      for (int s2 = 0; s2 < NS; ++s2)
        if (s2 == os) tau = make_reflector<RPL>(a[s2], vcol, cc, lane);
      // All lanes store the same tau / flag (no divergent region on the chain).  No MEMBAR: shared-memory
      // stores of one warp are performed in order and the flag is the last store instruction.
      sh_tau[cc] = tau;
      __syncwarp();
      *reinterpret_cast<volatile int*>(&ready[cc]) = 1;
#ifdef TTB_PANEL_STAMPS
      if (lane == 0) TTB_PANEL_STAMPS[8 + cc] = clock64();
#endif
    } else {
      // ---- consumer: wait for reflector cc ----
      // every lane polls the same word (a broadcast load): no divergence, no reconvergence barrier
      {
        unsigned it = 0;
        const bool hot = (wid == ((cc + 1) % NW));  // next link of the chain polls tightly
        while (ld_flag(&ready[cc]) == 0) {
          if (!hot) __nanosleep(TTB_PANEL_BACKOFF_NS);
          if (++it > (1u << 26)) __trap();
        }
      }
    }
    double v[RPL];
#ifdef TTB_PANEL_CHAIN_ONLY  // timing experiment: only the critical chain does work (results are wrong)
    pending = false;
#endif
    if (pending) {  // catch up the columns that were skipped while this warp hurried to publish
      const double taup = *reinterpret_cast<volatile double*>(&sh_tau[cc - 1]);
      const double* vp = vs + (size_t)(cc - 1) * LDV;
#pragma unroll
      for (int i = 0; i < RPL; ++i) v[i] = vp[lane + 32 * i];
#pragma unroll
      for (int s2 = 0; s2 < NS; ++s2)
        if (s2 != pend_pri) apply_slot(cc - 1, v, taup, s2);
      pending = false;
    }
    const double tau = *reinterpret_cast<volatile double*>(&sh_tau[cc]);
    {
      const double* vcol = vs + (size_t)cc * LDV;
#pragma unroll
      for (int i = 0; i < RPL; ++i) v[i] = vcol[lane + 32 * i];
    }
    // the slot that holds column cc+1 goes first: its owner is the next link of the chain
    const int pri = ((cc + 1) / NW) % NS;
#ifdef TTB_PANEL_CHAIN_ONLY
    if (wid == ((cc + 1) % NW)) apply_slot(cc, v, tau, pri);
    continue;
#endif
    apply_slot(cc, v, tau, pri);
    if (cc + 1 < bwc && wid == ((cc + 1) % NW)) {
      pending = true;
      pend_pri = pri;
    } else {
#pragma unroll
      for (int s2 = 0; s2 < NS; ++s2)
        if (s2 != pri) apply_slot(cc, v, tau, s2);
    }
  }
  __syncthreads();
#ifdef TTB_PANEL_STAMPS
  if (tid == 0) TTB_PANEL_STAMPS[1] = clock64();
#endif
  if (tid < 32) G[tid + 32 * tid] = tid < bwc ? sh_tau[tid] : 0.0;
#ifdef TTB_PANEL_STAMPS
  if (tid == 0) TTB_PANEL_STAMPS[2] = clock64();
#endif
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) {
    const int cs = wid + NW * s2;
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
  for (int i = tid; i < 1024; i += NT) Tg[i] = G[i];
}

}  // namespace ttb
