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

// 1/a for a normal, positive double: MUFU.RCP64H seed + two Newton iterations (the library division carries
// ~20 instructions of special-case handling; callers range-check).
__device__ __forceinline__ double rcp_nr(double a) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double e = fma(-a, y, 1.0);
  y = fma(y, e, y);
  e = fma(-a, y, 1.0);
  y = fma(y, e, y);
  return y;
}

// One link of the chain, executed by the warp that owns column cc = NS*wid + OS.
//
// UN-NORMALISED reflectors inside the factorisation: v' = (den, x) with x the raw column below the diagonal and
// den = alpha + sign(alpha) ||(alpha, x)||, H = I - tau' v' v'^T, tau' = 1 / (||.|| |den|).  Nothing on the chain
// waits for a scaled copy of the column: the raw x goes to shared memory BEFORE the reduction (raw[cc] is raised,
// consumers start their dot products while the owner is still in its butterfly and scalar arithmetic), and what
// follows the reduction is one rsqrt, one reciprocal (both MUFU seed + two Newton steps) and three stores
// (den into row cc of the published column, tau', the ready flag).  ONE reduction round per column: the norm of
// x and its dots with the warp's LATER columns ride the same butterfly; v'^T a = den a[cc] + x^T a.
// The LAPACK form (v_cc = 1, tau = tau' den^2) that the kernels downstream read is produced once per column in
// the store epilogue, off the chain (panel_normalise).
// STREAM (k_qr_stream): only the first `nvi` 32-row groups exist below the panel's first row; tau' is stored
// right behind them at vcol[tauoff] so that reflector + tau' leave the CTA as ONE contiguous bulk copy.
template <int RPL, int NS, int OS, bool STREAM = false>
__device__ __forceinline__ void panel_owner_step(double (&a)[NS][RPL], double* vcol, double* sh_tau, double* sh_den,
                                                 int* ready, int* raw, int cc, int lane, int wid, int bwc, int nvi = RPL,
                                                 int tauoff = 0) {
  constexpr int NL = NS - 1 - OS;   // later own columns
#ifdef TTB_PANEL_FINE
  if (lane == 0) TTB_PANEL_STAMPS[64 + 16 * cc + 0] = clock64();
#endif
  double x[RPL];
#pragma unroll
  for (int i = 0; i < RPL; ++i) x[i] = a[OS][i];
  const double x0m = (lane > cc) ? x[0] : 0.0;   // rows <= cc of the first 32 do not belong to the reflector
#ifdef TTB_PANEL_RAW_EARLY   // measured A/B (round 2): no gain in situ (54.9 vs 54.8 ms per compress!), kept for the record
  // the raw column goes to shared memory FIRST (it is final at entry): the next owner's dot products with it --
  // ~750 cycles of loads, FMAs and a transposed butterfly -- then finish before den / tau' are published, and what
  // is left on the hand-off is the rank-1 update (stamps: the flag used to follow the butterfly, the next owner
  // finished its dots ~1000 cycles after the publish)
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    if (!STREAM || i < nvi) vcol[lane + 32 * i] = i == 0 ? x0m : x[i];
  }
  __syncwarp();
  *reinterpret_cast<volatile int*>(&raw[cc]) = 1;
#endif
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
  // the raw column goes to shared memory in the latency bubbles of the butterfly (a quarter of it per level):
  // consumers start their dots on it while the owner is in its scalar arithmetic
  constexpr int SPL = (RPL + 3) / 4;
#pragma unroll
  for (int lv = 0; lv < 5; ++lv) {
    const int o = 16 >> lv;
    double got[NL + 1];
#pragma unroll
    for (int k = 0; k <= NL; ++k) got[k] = __shfl_xor_sync(0xffffffffu, red[k], o);
#ifndef TTB_PANEL_RAW_EARLY
#pragma unroll
    for (int i = lv * SPL; i < (lv + 1) * SPL; ++i) {
      if (i < RPL && (!STREAM || i < nvi)) vcol[lane + 32 * i] = i == 0 ? x0m : x[i];
    }
#endif
#pragma unroll
    for (int k = 0; k <= NL; ++k) red[k] += got[k];
  }
#ifndef TTB_PANEL_RAW_EARLY
  __syncwarp();
  *reinterpret_cast<volatile int*>(&raw[cc]) = 1;
#endif
  const double xn2 = red[NL];
#ifdef TTB_PANEL_FINE
  if (lane == 0 && xn2 != 1.2345) TTB_PANEL_STAMPS[64 + 16 * cc + 2] = clock64();
#endif
  const double alpha = acc[0];
  double den = 1.0, beta = alpha, taup = 0.0;
  if (xn2 > 0.0) {   // warp-uniform
    const double nn = fma(alpha, alpha, xn2);
    if (nn > 1e-280 && nn < 1e280) {
      const double rs = rsqrt_nr(nn);        // 1 / ||(alpha, x)||
      const double nrm = nn * rs;
      const double aa = fabs(alpha);
      const double ad = aa + nrm;            // |den|: no cancellation
      den = alpha >= 0.0 ? ad : -ad;
      beta = alpha >= 0.0 ? -nrm : nrm;
      taup = rs * rcp_nr(ad);                // 2 / (v'^T v') = 1 / (nrm |den|)
    } else {                                  // far outside the normal range (rare): library arithmetic
      const double nrm = sqrt(nn);
      const double ad = fabs(alpha) + nrm;
      den = alpha >= 0.0 ? ad : -ad;
      beta = alpha >= 0.0 ? -nrm : nrm;
      taup = (1.0 / nrm) / ad;
    }
  }
#ifdef TTB_PANEL_FINE
  if (lane == 0 && den != 1.2345 && taup != 1.2345) TTB_PANEL_STAMPS[64 + 16 * cc + 3] = clock64();
#endif
  // publish: every lane stores the same values (no divergent region on the chain).  No MEMBAR: shared stores
  // of one warp are performed in order and the flag is the last store instruction.
  vcol[cc] = den;
  sh_tau[cc] = taup;
  sh_den[cc] = den;
  if (STREAM) {
    vcol[tauoff] = taup;
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
  // own column: R_cc on the diagonal, the raw x stays below (normalised in the epilogue), untouched above
  const double v0 = lane > cc ? x[0] : (lane == cc ? den : 0.0);
  if (lane == cc) a[OS][0] = beta;
#pragma unroll
  for (int k = 0; k < NL; ++k) {
    const int s2 = OS + 1 + k;
    if (wid * NS + s2 < bwc) {   // warp-uniform
      const double w = taup * fma(den, acc[k + 1], red[k]);   // tau' * v'^T a_cs
      a[s2][0] -= v0 * w;
#pragma unroll
      for (int i = 1; i < RPL; ++i) a[s2][i] -= x[i] * w;
    }
  }
#ifdef TTB_PANEL_FINE
  if (lane == 0 && a[NS - 1][0] != 1.2345) TTB_PANEL_STAMPS[64 + 16 * cc + 5] = clock64();
#endif
}

// A warp that owns LATER columns applies reflector cc of the block in flight to all NS of them at once.  Two
// flags: raw[cc] says the raw column x is in shared memory -- the dots x^T a and their transpose butterfly (6
// shuffles for 4 values; lane l ends up with the total of value l / (32 / NS)) run while the owner is still
// reducing -- and ready[cc] says den / tau' are there: what is left on the hand-off is two broadcast loads, one
// FMA per column and the rank-1 update.
template <int RPL, int NS>
__device__ __forceinline__ void panel_consume_later(double (&a)[NS][RPL], const double* vcol, const double* sh_tau,
                                                    const int* ready, const int* raw, int cc, int lane, bool relaxed,
                                                    int nvi = RPL) {
  unsigned it = 0;
  while (ld_flag(&raw[cc]) == 0) {
    if (relaxed) __nanosleep(100);
    if (++it > (1u << 26)) __trap();
  }
#ifdef TTB_PANEL_FINE
  if (lane == 0 && !relaxed) TTB_PANEL_STAMPS[64 + 16 * cc + 6] = clock64();
#endif
  double v[RPL];
  v[0] = lane > cc ? vcol[lane] : 0.0;    // row cc will hold den: it enters through acc below, never through v
#pragma unroll
  for (int i = 1; i < RPL; ++i) v[i] = i < nvi ? vcol[lane + 32 * i] : 0.0;   // (stream: tau' sits behind the valid groups)
  double dot[NS], acc[NS];
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) dot[s2] = dot_rpl<RPL>(v, a[s2]);
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) acc[s2] = __shfl_sync(0xffffffffu, a[s2][0], cc);
  const double tot = warp_reduce_ns<NS>(dot, lane);
  double ts[NS];
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) ts[s2] = __shfl_sync(0xffffffffu, tot, s2 * (32 / NS));
#ifdef TTB_PANEL_FINE
  if (lane == 0 && !relaxed && ts[0] != 1.2345) TTB_PANEL_STAMPS[64 + 16 * cc + 7] = clock64();
#endif
  it = 0;
  while (ld_flag(&ready[cc]) == 0) {
    if (++it > (1u << 26)) __trap();
  }
#ifdef TTB_PANEL_FINE
  if (lane == 0 && !relaxed) TTB_PANEL_STAMPS[64 + 16 * cc + 8] = clock64();
#endif
  const double den = *reinterpret_cast<const volatile double*>(&vcol[cc]);
  const double taup = *reinterpret_cast<const volatile double*>(&sh_tau[cc]);
  const double v0 = lane == cc ? den : v[0];
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) {
    const double w = taup * fma(den, acc[s2], ts[s2]);
    a[s2][0] -= v0 * w;
#pragma unroll
    for (int i = 1; i < RPL; ++i) a[s2][i] -= v[i] * w;
  }
#ifdef TTB_PANEL_FINE
  if (lane == 0 && !relaxed && a[0][0] != 1.2345) TTB_PANEL_STAMPS[64 + 16 * cc + 9] = clock64();
#endif
}

// Two reflectors of the block in flight (cc, cc + 1) applied TOGETHER.  Stamps (profiles/r2_micro_panel.txt): applying
// one reflector costs a consuming warp ~935 cycles (16 loads, 64 FMAs of dots, a transposed butterfly, the flag wait,
// 64 FMAs of update) while the owner forms one every 680-960 -- the NEXT owner falls behind inside the block and still
// has ~1.5 reflectors of work left when the block's last reflector is published: that, not the flag latency, is the
// 1000-2400 cycles of every hand-off.  With V0 = den0 e_cc + x0, V1 = den1 e_{cc+1} + x1:
//   d0 = V0.a, d1 = V1.a (both on the ORIGINAL a), g = V1.V0 = den1 x0[cc+1] + x1.x0,
//   w0 = tau0 d0, w1 = tau1 (d1 - g w0), a -= w0 V0 + w1 V1
// -- one set of loads, one round of butterflies, one wait and one update pass for two reflectors.  The raw dots start
// when raw[cc + 1] is up (x1 is the column AFTER the owner applied reflector cc to it); both columns are streamed
// from shared memory twice instead of living in registers across the butterflies (k_qr_stream sits at 254 registers).
template <int RPL, int NS>
__device__ __forceinline__ void panel_consume_pair(double (&a)[NS][RPL], const double* v0c, const double* v1c, const double* sh_tau,
                                                   const int* ready, const int* raw, int cc, int lane, bool relaxed, int nvi = RPL) {
  unsigned it = 0;
  while (ld_flag(&raw[cc + 1]) == 0) {
    if (relaxed) __nanosleep(100);
    if (++it > (1u << 26)) __trap();
  }
  double d0[NS], d1[NS], g = 0.0;
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) { d0[s2] = 0.0; d1[s2] = 0.0; }
  const double x00 = lane > cc ? v0c[lane] : 0.0;          // rows <= cc do not belong to reflector cc
  const double x10 = lane > cc + 1 ? v1c[lane] : 0.0;
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    const double x0 = i == 0 ? x00 : (i < nvi ? v0c[lane + 32 * i] : 0.0);
    const double x1 = i == 0 ? x10 : (i < nvi ? v1c[lane + 32 * i] : 0.0);
    g = fma(x1, x0, g);
#pragma unroll
    for (int s2 = 0; s2 < NS; ++s2) { d0[s2] = fma(x0, a[s2][i], d0[s2]); d1[s2] = fma(x1, a[s2][i], d1[s2]); }
  }
  double acc0[NS], acc1[NS];
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) {
    acc0[s2] = __shfl_sync(0xffffffffu, a[s2][0], cc);
    acc1[s2] = __shfl_sync(0xffffffffu, a[s2][0], cc + 1);
  }
  const double x0c1 = __shfl_sync(0xffffffffu, x00, cc + 1);   // x0[cc + 1]
  const double t0 = warp_reduce_ns<NS>(d0, lane);
  const double t1 = warp_reduce_ns<NS>(d1, lane);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
  double ts0[NS], ts1[NS];
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) {
    ts0[s2] = __shfl_sync(0xffffffffu, t0, s2 * (32 / NS));
    ts1[s2] = __shfl_sync(0xffffffffu, t1, s2 * (32 / NS));
  }
  it = 0;
  while (ld_flag(&ready[cc + 1]) == 0) {
    if (++it > (1u << 26)) __trap();
  }
  const double den0 = *reinterpret_cast<const volatile double*>(&v0c[cc]);
  const double tau0 = *reinterpret_cast<const volatile double*>(&sh_tau[cc]);
  const double den1 = *reinterpret_cast<const volatile double*>(&v1c[cc + 1]);
  const double tau1 = *reinterpret_cast<const volatile double*>(&sh_tau[cc + 1]);
  const double gg = fma(den1, x0c1, g);
  double w0[NS], w1[NS];
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) {
    w0[s2] = tau0 * fma(den0, acc0[s2], ts0[s2]);
    w1[s2] = tau1 * (fma(den1, acc1[s2], ts1[s2]) - gg * w0[s2]);
  }
  const double v00 = lane == cc ? den0 : x00;
  const double v10 = lane == cc + 1 ? den1 : x10;
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    const double x0 = i == 0 ? v00 : (i < nvi ? v0c[lane + 32 * i] : 0.0);
    const double x1 = i == 0 ? v10 : (i < nvi ? v1c[lane + 32 * i] : 0.0);
#pragma unroll
    for (int s2 = 0; s2 < NS; ++s2) a[s2][i] -= fma(x0, w0[s2], x1 * w1[s2]);
  }
}

// V'^T V' of the whole panel on the FP64 tensor cores, once, off the chain: G[i + 32 j] = sum_r v'_i[r] v'_j[r] for
// i < j < bwc from the published reflectors (shared memory, column stride SL, rows 0 .. krows-1 valid, zeros above
// each diagonal and den on it).  While the chain runs, warps whose columns are finished do NOTHING -- their dots
// used to take FP64 issue slots from the owner and the next owner on the same SM sub-partition.  Eight 16 x 8
// tiles of mma.sync.m16n8k8.f64, one per warp (warps 0..7), K = krows.  Call between two block barriers.
__device__ __forceinline__ void panel_gram_mma(const double* vs, int SL, int krows, double* G, int wid, int lane, int bwc) {
  if (wid >= 8) return;
  const int mt = wid & 1, nt = wid >> 1;         // rows 16 mt .., columns 8 nt ..
  if (16 * mt >= 8 * nt + 8) return;             // tile entirely on or below the diagonal
  const int g = lane >> 2, t = lane & 3;
  const double* pa = vs + (size_t)(16 * mt + g) * SL + t;
  const double* pb = vs + (size_t)(8 * nt + g) * SL + t;
  double d[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k0 = 0; k0 < krows; k0 += 8) {
    const double af[4] = {pa[k0], pa[k0 + 8 * (size_t)SL], pa[k0 + 4], pa[k0 + 8 * (size_t)SL + 4]};
    const double bf[2] = {pb[k0], pb[k0 + 4]};
    dmma_16x8x8(d, af, bf);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = 16 * mt + g + ((e & 2) ? 8 : 0), j = 8 * nt + 2 * t + (e & 1);
    if (i < j && j < bwc) G[i + 32 * j] = d[e];
  }
}

// Store epilogue, off the chain: LAPACK form of the panel.  Column cs of the warp: rows below the diagonal are
// divided by den_cs; G (strict upper part = V'^T V') is scaled by 1/(den_i den_j), its diagonal receives
// tau_i = tau'_i den_i^2.  Call after a block barrier (all reflectors and G entries are in shared memory).
template <int RPL, int NS>
__device__ __forceinline__ void panel_normalise(double (&a)[NS][RPL], double* G, const double* sh_tau, const double* sh_den,
                                                int tid, int nthreads, int wid, int lane, int bwc) {
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) {
    const int cs = wid * NS + s2;
    if (cs < bwc) {
      const double scal = 1.0 / sh_den[cs];
      if (lane > cs) a[s2][0] *= scal;
#pragma unroll
      for (int i = 1; i < RPL; ++i) a[s2][i] *= scal;
    }
  }
  for (int i = tid; i < 1024; i += nthreads) {
    const int r = i & 31, cj = i >> 5;
    if (r < cj && cj < bwc) G[i] = G[i] / (sh_den[r] * sh_den[cj]);
    else if (r == cj) G[i] = cj < bwc ? sh_tau[cj] * sh_den[cj] * sh_den[cj] : 0.0;
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
  double* sh_tau = G + 1024;              // 32  tau' of the un-normalised reflectors
  double* sh_den = sh_tau + 32;           // 32  their diagonal elements
  int* ready = reinterpret_cast<int*>(sh_den + 32);  // 32 flags: den / tau' published
  int* raw = ready + 32;                  // 32 flags: raw column in shared memory
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
  if (tid < 64) ready[tid] = 0;   // ready[32] + raw[32]
  if (tid < 32) sh_den[tid] = 1.0;
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
      panel_owner_step<RPL, NS, 0>(a, vs + (size_t)c0 * LDV, sh_tau, sh_den, ready, raw, c0, lane, wid, bwc);
      if (NS > 1 && c0 + 1 < bwc)
        panel_owner_step<RPL, NS, (1 < NS ? 1 : 0)>(a, vs + (size_t)(c0 + 1) * LDV, sh_tau, sh_den, ready, raw, c0 + 1, lane, wid, bwc);
      if (NS > 2 && c0 + 2 < bwc)
        panel_owner_step<RPL, NS, (2 < NS ? 2 : 0)>(a, vs + (size_t)(c0 + 2) * LDV, sh_tau, sh_den, ready, raw, c0 + 2, lane, wid, bwc);
      if (NS > 3 && c0 + 3 < bwc)
        panel_owner_step<RPL, NS, (3 < NS ? 3 : 0)>(a, vs + (size_t)(c0 + 3) * LDV, sh_tau, sh_den, ready, raw, c0 + 3, lane, wid, bwc);
    } else if (wid * NS >= bwc) {
      // no column of this warp exists in a narrow panel: nothing to update, nothing to contribute to G
    } else if (wid > blk) {
      // all own columns lie after the block (columns >= bwc are zero and stay zero): rank-1 updates, started on
      // the raw column while the owner is still reducing.  Warps that are not about to take over the chain back
      // off between polls: their spinning would take issue slots from the owner on the same SM sub-partition
#ifdef TTB_PANEL_DEFER   // measured A/B (round 2): no effect -- the step in the block times at block 4 is NOT sub-partition sharing
      if (NS == 4) {   // the sub-partition partner of the chain warp (blk + 4) applies this block one block later (qr_stream.cuh)
        if (wid == blk + 4) continue;
        if (wid == blk + 3 && blk >= 1) {
          const int d0 = (blk - 1) * NS;
#pragma unroll 1
          for (int cc = d0; cc < d0 + NS; ++cc)
            panel_consume_later<RPL, NS>(a, vs + (size_t)cc * LDV, sh_tau, ready, raw, cc, lane, false);
        }
      }
#endif
      const int cend = min(c0 + NS, bwc);
#ifdef TTB_PANEL_PAIR   // measured (micro_panel2, round 2): hand-offs 15-30 % LONGER -- the pair's dots start at the second raw column and are twice the work
#pragma unroll 1
      for (int cc = c0; cc < cend;) {
        if (cc + 1 < cend) {
          panel_consume_pair<RPL, NS>(a, vs + (size_t)cc * LDV, vs + (size_t)(cc + 1) * LDV, sh_tau, ready, raw, cc, lane, wid != blk + 1);
          cc += 2;
        } else {
          panel_consume_later<RPL, NS>(a, vs + (size_t)cc * LDV, sh_tau, ready, raw, cc, lane, wid != blk + 1);
          cc += 1;
        }
      }
#else
#pragma unroll 1
      for (int cc = c0; cc < cend; ++cc)
        panel_consume_later<RPL, NS>(a, vs + (size_t)cc * LDV, sh_tau, ready, raw, cc, lane, wid != blk + 1);
#endif
    }
    // (warps whose columns are finished reflectors have nothing to do: V'^T V' is formed after the chain)
  }
  __syncthreads();
#ifdef TTB_PANEL_STAMPS
  if (tid == 0) TTB_PANEL_STAMPS[1] = clock64();
#endif
  for (int i = bwc * LDV + tid; i < 32 * LDV; i += NT) vs[i] = 0.0;   // columns a narrow panel never published
  __syncthreads();
  panel_gram_mma(vs, LDV, LDV, G, wid, lane, bwc);
  __syncthreads();
  panel_normalise<RPL, NS>(a, G, sh_tau, sh_den, tid, NT, wid, lane, bwc);
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
