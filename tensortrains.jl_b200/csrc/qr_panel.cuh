// qr_panel.cuh -- k_qr_panel_flag: register-resident Householder panel (32 columns, <= 512 rows) whose
// warps hand reflectors to each other through shared-memory flags instead of block barriers.
//
// Layout: NW = 32/NS warps; warp w owns the NS CONSECUTIVE panel columns NS*w .. NS*w+NS-1, lane l holds
// rows l + 32 i.  The panel is a chain of 32 dependent columns (reflector c needs reflector c-1 applied
// to column c).  Inside a warp's block the chain never leaves the warp: the reflector it has just
// formed is applied to its own later columns straight from registers, so only every NS-th link pays the
// shared-memory hand-off (publish 32*RPL doubles, raise ready[c], next warp polls).  All 32 reflectors
// stay in shared memory (no buffer reuse => no WAR hazard); the other warps consume them in order at
// their own pace, all NS columns at once (independent dots, interleaved warp reductions).
// The V^T V products that define the compact-WY T factor (T^{-1} = diag(1/tau) + striu(V^T V)) fall out
// of the same dots and are shipped as they are: no T factor is ever built.
// Measured (profiles/micro_lat.cu, micro_panel2.cu, B200): dependent DFMA 8, shfl+DADD 35, rsqrt ~80,
// 1/x ~80, LDS ~70 cycles; round-robin ownership cost ~2450 cycles per column of which ~600 were the
// hand-off, block ownership removes it from NS-1 of NS links.
// Reference: same call sites as qr_stream.cuh (src/tensor_train.jl:161-163, :200-202); this is the 32-column
// panel of the blocked Householder QR that replaces the eps = 0 SVD and precedes the Jacobi SVD of tall steps.
#pragma once

namespace ttb {

__device__ __forceinline__ int ld_flag(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}

// NS (= 2 or 4) per-lane values -> lane l holds the warp total of value l / (32 / NS); every level of the
// first log2(NS) sends the half of the values the partner lane keeps
template <int NS>
__device__ __forceinline__ double warp_reduce_ns(double (&a)[NS], int lane) {
  static_assert(NS == 2 || NS == 4, "panel block size");
  double r;
  if (NS == 4) {
    double b2[2];
    {
      const bool hi = lane & 16;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const double send = hi ? a[k] : a[k + 2], keep = hi ? a[k + 2] : a[k];
        b2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
    }
    const bool hi = lane & 8;
    const double send = hi ? b2[0] : b2[1], keep = hi ? b2[1] : b2[0];
    r = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  } else {
    const bool hi = lane & 16;
    const double send = hi ? a[0] : a[NS - 1], keep = hi ? a[NS - 1] : a[0];
    r = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    r += __shfl_xor_sync(0xffffffffu, r, 8);
  }
  r += __shfl_xor_sync(0xffffffffu, r, 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
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

// One link of the chain, executed by the warp that owns column cc = NS*wid + OS.  ONE reduction round: the
// norm of the raw column x (rows > cc) and its dots with the warp's LATER columns ride the same butterfly;
// with v = (e_cc, x * scal) the products the reflector needs are v^T a = a[cc] + scal * (x^T a), so nothing
// waits for the normalised v.  Publishes v / tau / the ready flag, updates the later own columns, leaves
// (R_cc on the diagonal, v below) in the column.  Returns tau; v stays in registers.
// STREAM (k_qr_stream): only the first `nvi` 32-row groups exist below the panel's first row; tau is stored
// right behind them at vcol[tauoff] so that reflector + tau leave the CTA as ONE contiguous bulk copy.
template <int RPL, int NS, int OS, bool STREAM = false>
__device__ __forceinline__ void panel_owner_step(double (&a)[NS][RPL], double* vcol, double* sh_tau, int* ready, int cc,
                                                 int lane, int wid, int bwc, int nvi = RPL, int tauoff = 0) {
  constexpr int NL = NS - 1 - OS;   // later own columns
  double v[RPL];
#ifdef TTB_PANEL_FINE
  if (lane == 0) TTB_PANEL_STAMPS[64 + 16 * cc + 0] = clock64();
#endif
  double x[RPL];
#pragma unroll
  for (int i = 0; i < RPL; ++i) x[i] = a[OS][i];
  const double x0m = (lane > cc) ? x[0] : 0.0;   // rows <= cc of the first 32 do not belong to the reflector
  double red[NL + 1];
  {
    double p0 = x0m * x0m, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
    for (int i = 1; i < RPL; ++i) {
      if ((i & 3) == 0) p0 += x[i] * x[i];
      else if ((i & 3) == 1) p1 += x[i] * x[i];
      else if ((i & 3) == 2) p2 += x[i] * x[i];
      else p3 += x[i] * x[i];
    }
    red[NL] = (p0 + p1) + (p2 + p3);
  }
#pragma unroll
  for (int k = 0; k < NL; ++k) {
    const int s2 = OS + 1 + k;
    double d0 = x0m * a[s2][0], d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
    for (int i = 1; i < RPL; ++i) {
      if ((i & 3) == 0) d0 += x[i] * a[s2][i];
      else if ((i & 3) == 1) d1 += x[i] * a[s2][i];
      else if ((i & 3) == 2) d2 += x[i] * a[s2][i];
      else d3 += x[i] * a[s2][i];
    }
    red[k] = (d0 + d1) + (d2 + d3);
  }
#ifdef TTB_PANEL_FINE
  if (lane == 0 && red[0] != 1.2345) TTB_PANEL_STAMPS[64 + 16 * cc + 1] = clock64();
#endif
  double acc[NL + 1];   // element at row cc of this and the later own columns (row cc lives in lane cc, i = 0)
#pragma unroll
  for (int k = 0; k <= NL; ++k) acc[k] = __shfl_sync(0xffffffffu, a[OS + k][0], cc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k <= NL; ++k) red[k] += __shfl_xor_sync(0xffffffffu, red[k], o);
  }
  const double xn2 = red[NL];
#ifdef TTB_PANEL_FINE
  if (lane == 0 && xn2 != 1.2345) TTB_PANEL_STAMPS[64 + 16 * cc + 2] = clock64();
#endif
  const double alpha = acc[0];
  double scal = 0.0, beta = alpha, tau = 0.0;
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
#ifdef TTB_PANEL_FINE
  if (lane == 0 && scal != 1.2345 && tau != 1.2345) TTB_PANEL_STAMPS[64 + 16 * cc + 3] = clock64();
#endif
  {
    double v0 = 0.0;
    if (lane > cc) v0 = x[0] * scal;
    else if (lane == cc) v0 = 1.0;
    v[0] = v0;
    vcol[lane] = v0;
  }
#pragma unroll
  for (int i = 1; i < RPL; ++i) {
    v[i] = x[i] * scal;
    if (!STREAM || i < nvi) vcol[lane + 32 * i] = v[i];
  }
  // every lane stores the same tau / flag (no divergent region on the chain).  No MEMBAR: shared stores of
  // one warp are performed in order and the flag is the last store instruction.
  sh_tau[cc] = tau;
  if (STREAM) {
    vcol[tauoff] = tau;
    fence_proxy_async();   // writer-side proxy fence: the reflector leaves through a bulk (async-proxy) copy
  }
  __syncwarp();
  *reinterpret_cast<volatile int*>(&ready[cc]) = 1;
#ifdef TTB_PANEL_STAMPS
  if (lane == 0) TTB_PANEL_STAMPS[8 + cc] = clock64();
#endif
#ifdef TTB_PANEL_FINE
  if (lane == 0) TTB_PANEL_STAMPS[64 + 16 * cc + 4] = clock64();
#endif
  // own column: R_cc on the diagonal, v below, untouched above
  if (lane > cc) a[OS][0] = v[0];
  else if (lane == cc) a[OS][0] = beta;
#pragma unroll
  for (int i = 1; i < RPL; ++i) a[OS][i] = v[i];
#pragma unroll
  for (int k = 0; k < NL; ++k) {
    const int s2 = OS + 1 + k;
    if (wid * NS + s2 < bwc) {   // warp-uniform
      const double w = tau * (acc[k + 1] + scal * red[k]);   // tau * v_cc^T a_cs
#pragma unroll
      for (int i = 0; i < RPL; ++i) a[s2][i] -= v[i] * w;
    }
  }
#ifdef TTB_PANEL_FINE
  if (lane == 0 && a[NS - 1][0] != 1.2345) TTB_PANEL_STAMPS[64 + 16 * cc + 5] = clock64();
#endif
}

// V^T V entries between the columns of ONE finished block (s1 < s2): G = v_s1[row cs2] + sum_{r > cs2} v_s1[r] v_s2[r].
// Runs once per block after its last reflector has been published, i.e. off the chain.
template <int RPL, int NS>
__device__ __forceinline__ void panel_block_gram(const double (&a)[NS][RPL], double* G, int wid, int lane, int bwc) {
  constexpr int NPAIR = NS * (NS - 1) / 2;
  double red[NPAIR > 0 ? NPAIR : 1], diag[NPAIR > 0 ? NPAIR : 1];
  int k = 0;
#pragma unroll
  for (int s2 = 1; s2 < NS; ++s2) {
    const int cs2 = wid * NS + s2;
#pragma unroll
    for (int s1 = 0; s1 < s2; ++s1) {
      double d0 = (lane > cs2) ? a[s1][0] * a[s2][0] : 0.0, d1 = 0.0;
#pragma unroll
      for (int i = 1; i < RPL; ++i) {
        if (i & 1) d1 += a[s1][i] * a[s2][i];
        else d0 += a[s1][i] * a[s2][i];
      }
      red[k] = d0 + d1;
      diag[k] = __shfl_sync(0xffffffffu, a[s1][0], cs2 & 31);
      ++k;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int q = 0; q < NPAIR; ++q) red[q] += __shfl_xor_sync(0xffffffffu, red[q], o);
  }
  k = 0;
#pragma unroll
  for (int s2 = 1; s2 < NS; ++s2) {
    const int cs2 = wid * NS + s2;
#pragma unroll
    for (int s1 = 0; s1 < s2; ++s1) {
      if (cs2 < bwc) G[(wid * NS + s1) + 32 * cs2] = diag[k] + red[k];   // same value from every lane
      ++k;
    }
  }
}

template <int RPL, int NS>
__global__ void __launch_bounds__(1024 / NS, (RPL <= 2 ? 4 : 1)) k_qr_panel_flag(SweepCtx c, StepDesc s, int panel) {
  pdl_enter();
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
    const int cs = wid * NS + s2;
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

  // One block of NS columns at a time.  The owner warp runs its NS chain links as straight-line code (no
  // per-column dispatch: ~350 cycles of loop / switch / register shuffling per link in the previous
  // form, and the rank-1 updates of one link overlap the partial sums of the next); every other warp
  // consumes the block's reflectors as they are published.
#pragma unroll 1
  for (int blk = 0; blk * NS < bwc; ++blk) {
    const int c0 = blk * NS;
    if (wid == blk) {
      panel_owner_step<RPL, NS, 0>(a, vs + (size_t)c0 * LDV, sh_tau, ready, c0, lane, wid, bwc);
      if (NS > 1 && c0 + 1 < bwc)
        panel_owner_step<RPL, NS, (1 < NS ? 1 : 0)>(a, vs + (size_t)(c0 + 1) * LDV, sh_tau, ready, c0 + 1, lane, wid, bwc);
      if (NS > 2 && c0 + 2 < bwc)
        panel_owner_step<RPL, NS, (2 < NS ? 2 : 0)>(a, vs + (size_t)(c0 + 2) * LDV, sh_tau, ready, c0 + 2, lane, wid, bwc);
      if (NS > 3 && c0 + 3 < bwc)
        panel_owner_step<RPL, NS, (3 < NS ? 3 : 0)>(a, vs + (size_t)(c0 + 3) * LDV, sh_tau, ready, c0 + 3, lane, wid, bwc);
      panel_block_gram<RPL, NS>(a, G, wid, lane, bwc);
    } else {
      const int cend = min(c0 + NS, bwc);
#pragma unroll 1
      for (int cc = c0; cc < cend; ++cc) {
        // ---- consumer: every lane polls the same word (broadcast load, no divergence).  Warps that are
        // not about to take over the chain back off between polls: their spinning would take issue slots
        // from the owner on the same SM sub-partition
        const bool relaxed = (wid != blk + 1);
        unsigned it = 0;
        while (ld_flag(&ready[cc]) == 0) {
          if (relaxed) __nanosleep(100);
          if (++it > (1u << 26)) __trap();
        }
#ifdef TTB_PANEL_FINE
        if (lane == 0 && wid == blk + 1) TTB_PANEL_STAMPS[64 + 16 * cc + 6] = clock64();
#endif
        const double tau = *reinterpret_cast<volatile double*>(&sh_tau[cc]);
        const double* vcol = vs + (size_t)cc * LDV;
        double v[RPL];
#pragma unroll
        for (int i = 0; i < RPL; ++i) v[i] = vcol[lane + 32 * i];
        // apply reflector cc to all NS columns of this warp at once: independent dots, ONE transpose
        // butterfly (each level sends the half of the values the partner lane keeps: 6 shuffles for 4
        // values instead of 20; lane l ends up with the total of value (l >> 3) & 3).  The consumer path is
        // what the next owner must get through before it can take over the chain, so its instruction
        // count is on the critical path at every hand-off.
        double dot[NS];
#pragma unroll
        for (int s2 = 0; s2 < NS; ++s2) dot[s2] = dot_rpl<RPL>(v, a[s2]);
        const double tot = warp_reduce_ns<NS>(dot, lane);
        if (wid > blk) {
          // all own columns lie after cc (columns >= bwc are zero and stay zero): rank-1 updates
#pragma unroll
          for (int s2 = 0; s2 < NS; ++s2) {
            const double w = tau * __shfl_sync(0xffffffffu, tot, s2 * (32 / NS));
#pragma unroll
            for (int i = 0; i < RPL; ++i) a[s2][i] -= v[i] * w;
          }
        } else {
          // all own columns are finished reflectors: only the V^T V entries
          if ((lane & (32 / NS - 1)) == 0) G[(wid * NS + lane / (32 / NS)) + 32 * cc] = tot;
        }
#ifdef TTB_PANEL_FINE
        if (lane == 0 && wid == blk + 1 && a[0][0] != 1.2345) TTB_PANEL_STAMPS[64 + 16 * cc + 7] = clock64();
#endif
      }
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
    const int cs = wid * NS + s2;
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
