// qr_stream.cuh -- k_qr_stream: the blocked Householder QR of ONE site as a single thread-block CLUSTER.
//
// The panel/update chain of sweep.cu costs, per 32-column panel, one panel kernel (~27 us) plus one
// trailing-update kernel that the next panel has to wait for (~10 us with its launch gap) -- and every
// column block travels registers -> global -> registers between panels.  Here CTA k of the cluster OWNS
// column block k (32 columns, all rows, in registers, for the whole factorisation):
//   * while panels 0..k-1 are being factored in the other CTAs it CONSUMES their reflectors one by one as
//     they arrive (rank-1 updates of its 4 columns per warp, the consumer path of qr_panel.cuh);
//   * then it factors panel k with the register-resident chain of qr_panel.cuh;
//   * then it stores R / V / the V^T V + tau block.
// No trailing-update kernels, no launch gaps, no round trip through global memory between panels: the
// critical path is the sum of the panel chains plus one push -> apply latency per hand-off.
//
// Transport of a reflector (distributed shared memory, no global staging, no polling of global memory):
// the owner warp writes it to the factoring CTA's shared memory as before; warp 0 of that CTA PUSHES it
// with bulk copies to every later CTA (cp.async.bulk shared::cta -> shared::cluster, lane r serves CTA r)
// into slot g mod 8 of the destination's ring, completing on the DESTINATION's mbarrier.  Flow control is
// credit based: ring slot w of a consumer CTA belongs to its warp w, which re-arms the slot's mbarrier
// (expect_tx) as soon as every warp of the CTA is done with the previous occupant and then releases a
// credit (one remote store) into the producing CTA; the pusher never blocks on credits except in its final
// flush.  Slots are in ABSOLUTE row order (zeros above the producing panel, tau behind the last row), so
// the consumer's register rotation turns into two contiguous runs with immediate offsets, and reflector
// + tau are one contiguous copy on the producer's side.
//
// Row mapping: CTA k keeps absolute row ((i + k) mod RPL) * 32 + lane in register i, so that in its own
// factor phase register 0 holds the panel's diagonal rows -- the mapping qr_panel.cuh assumes.  Rows above
// the panel are final when the factor phase starts: they are stored and zeroed at the transition.
// Requirements (host-checked): dimensions exact on the host, p >= n, bw == 32, p <= 512, 2..8 column blocks.
// Reference: the factorisation inside one step of orthogonalize_right! / orthogonalize_left! with
// TruncThresh(0.0) -- `U, lambda, V = svd_trunc(M)` at src/tensor_train.jl:161-163 (right sweep, M = d_t x
// d_{t+1}Q) and :200-202 (left sweep), src/periodic_tensor_train.jl:86,95,119,133; with eps = 0 every
// direction is kept (src/svd_trunc.jl:38-47), so W S = Q R is gauge equivalent to U (lambda V').
#pragma once

namespace ttb {

#ifndef TTB_QS_PUSH_SLEEP
#define TTB_QS_PUSH_SLEEP 100
#endif
#ifndef TTB_QS_CREDIT_LD
#define TTB_QS_CREDIT_LD ld_acquire_cluster_s32
#endif
constexpr int kQsRing = 8;         // reflector slots per consumer CTA
constexpr int kQsThreads = 256;    // 8 warps, 4 columns each
constexpr int kQsMaxBlocks = 8;    // portable cluster size

__host__ __device__ constexpr size_t qs_smem_bytes(int RPL) {
  // vs[32][SL] + ring[RING][SL] + G[1024] + tau'[32] + den[32] doubles, ready[32] + raw[32] + done[8] + credit[8][RING] ints,
  // RING mbarriers
  return sizeof(double) * ((size_t)(32 + kQsRing) * (32 * RPL + 8) + 1024 + 64) + sizeof(int) * (72 + 8 * kQsRing) + 8 * kQsRing + 64;
}

// timeline stamps (globaltimer ns) of one traced site step: [CTA][0] start, [1 + j] consumed panel j,
// [20] factor done, [21] all stored, [22 + blk] block done; written only when the step descriptor asks
__device__ unsigned long long g_qs_trace[16 * 32];
__device__ unsigned long long g_qs_trace2[4 * 64];   // panel 0 -> CTA 1: [0] pushed, [1] ready seen by the pusher, [2] arrived, [3] applied
__device__ __forceinline__ unsigned long long qs_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int RPL>
__global__ void __launch_bounds__(kQsThreads, 1) k_qr_stream(SweepCtx c, StepDesc s) {
  constexpr int NS = 4;
  constexpr int LDV = 32 * RPL;
  constexpr int SL = LDV + 8;           // slot stride; own panel: relative rows, tau right behind the valid ones;
                                        // ring: absolute rows (zeros above the producing panel), tau at [LDV]
  const int b = blockIdx.y, kb = blockIdx.x;   // kb == rank in the cluster (cluster = one row of the grid)
  const int nblk = gridDim.x;
  const Dims d = get_dims(c, s, b);     // exact (host-known) dimensions
  const int64_t p = d.p;
  const int n = d.n;
  const int j0 = 32 * kb;
  const int bwc = min(32, n - j0);      // columns of this block (p >= n: all of them are factored)
  const int nremote = j0;               // reflectors of the earlier panels
  const int rows = d.p - j0;
  const int nvi = RPL - kb;             // 32-row groups at or below the panel's first row
  double* Xa = xbuf(c, s.xin, b);
  double* Xp = Xa + j0 + p * j0;        // panel origin
  extern __shared__ __align__(16) double dyn[];
  double* vs = dyn;                                  // [32][SL] reflectors of the own panel
  double* ring = vs + 32 * SL;                       // [RING][SL] remote reflectors
  double* G = ring + kQsRing * SL;                   // 32*32
  double* sh_tau = G + 1024;                         // 32  tau' of the un-normalised reflectors (qr_panel.cuh)
  double* sh_den = sh_tau + 32;                      // 32  their diagonal elements
  int* ready = reinterpret_cast<int*>(sh_den + 32);  // [32] den / tau' published
  int* raw = ready + 32;                             // [32] raw column in shared memory
  int* done = raw + 32;                              // [8] remote reflectors consumed per warp
  int* credit = done + 8;                            // [8 consumer ranks][RING]: (global index + 1) of the reflector the slot is armed for
  uint64_t* bars = reinterpret_cast<uint64_t*>(credit + 8 * kQsRing);  // [RING]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < 1024; i += kQsThreads) G[i] = 0.0;
  for (int i = tid; i < kQsRing * SL; i += kQsThreads) ring[i] = 0.0;   // rows above a producing panel must read as zero
  if (tid < 72 + 8 * kQsRing) ready[tid] = 0;   // ready[32] + raw[32] + done[8] + credit[8][RING]
  if (tid < 32) sh_den[tid] = 1.0;
  if (tid == 0)
    for (int r = 0; r < kQsRing; ++r) mbar_init(&bars[r], 1);
  __syncthreads();
  cluster_sync_all();             // every CTA's barriers / credits exist before anybody pushes or releases credits
  // everything above touched only this CTA's shared memory (the dimensions came with the launch): with
  // programmatic dependent launch it overlaps the tail of the previous kernel; global memory from here on
  pdl_enter();
  const bool trace = s.trace && b == 0 && tid == 0;
  if (trace) g_qs_trace[kb * 32 + 0] = qs_now();

  // columns 4*wid .. 4*wid+3 of the block, all rows, rotated row mapping
  double a[NS][RPL];
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) {
    const int cs = wid * NS + s2;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      const int r = ((i + kb) & (RPL - 1)) * 32 + lane;
      a[s2][i] = (cs < bwc && r < d.p) ? Xa[r + p * (j0 + cs)] : 0.0;
    }
  }
  // ---- consume the reflectors of the earlier panels; ring slot `wid` is armed by this warp
  {
    int my_next = wid;       // next reflector this warp arms its slot for (= wid mod 8)
    auto try_arm = [&](int g_now) {
      if (my_next >= nremote || g_now + (kQsRing - 1) < my_next) return;   // its slot cannot be free yet
      if (my_next >= kQsRing) {   // slot free: every warp is done with reflector my_next - 8
        int mc = lane < 8 ? *reinterpret_cast<volatile int*>(&done[lane]) : 0x7fffffff;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) mc = min(mc, __shfl_xor_sync(0xffffffffu, mc, o));
        mc = __shfl_sync(0xffffffffu, mc, 0);
        if (mc < my_next - (kQsRing - 1)) return;
      }
      const int kprod = my_next >> 5;
      double* slotp = ring + (size_t)wid * SL;
      // first use of the slot for this producing panel: the previous panel's rows just above it become zeros
      if (kprod > 0 && (my_next & 31) < kQsRing) slotp[32 * (kprod - 1) + lane] = 0.0;
      __syncwarp();
      if (lane == 0) {
        mbar_expect_tx(&bars[wid], (uint32_t)((LDV - 32 * kprod + 2) * 8));
        st_remote_release_s32(mapa_u32(smem_u32(&credit[kb * kQsRing + wid]), (uint32_t)kprod), my_next + 1);
      }
      __syncwarp();
      my_next += kQsRing;
    };
    // A remote reflector costs this CTA ~1200 cycles applied alone (16 loads, 4 dots, one transposed butterfly, the
    // update) while the factoring CTA forms one every ~1100: the consumer fell behind over a panel and the NEXT panel
    // started ~4 us after the previous one had finished (globaltimer trace).  Whenever two reflectors have already
    // arrived they are applied TOGETHER: d0 = v0.a, d1 = v1.a, g = v1.v0 ride the same butterflies,
    // w0 = tau0 d0, w1 = tau1 (d1 - g w0), a -= w0 v0 + w1 v1 -- one latency chain for two reflectors.
#pragma unroll 1
    for (int g = 0; g < nremote;) {
      try_arm(g);
      const int slot = g & (kQsRing - 1);
      const uint32_t parity = (uint32_t)(g / kQsRing) & 1u;
      for (uint32_t it = 0; !mbar_try_wait(&bars[slot], parity); ++it) {
        try_arm(g);
        if (it > (1u << 24)) __trap();
      }
#ifdef TTB_QS_TRACE2
      if (s.trace && b == 0 && kb == 1 && tid == 0 && g < 64) g_qs_trace2[2 * 64 + g] = qs_now();
#endif
      // register i holds row group (i + kb) mod RPL => two contiguous runs, immediate offsets
      const double* vsl = ring + (size_t)slot * SL;
      const double tau = vsl[LDV];
      const double* pA = vsl + 32 * kb + lane;
      const double* pB = pA - LDV;
      bool two = false;
#ifdef TTB_QS_PAIR   // measured A/B on one box (round 2): 55.0 ms per compress! with the pair path, 54.8 without -- the consumers are not the hand-off's bottleneck
      if (g + 1 < nremote) {
        try_arm(g + 1);
        const int slot1 = (g + 1) & (kQsRing - 1);
        two = __all_sync(0xffffffffu, mbar_try_wait(&bars[slot1], (uint32_t)((g + 1) / kQsRing) & 1u));
      }
#endif
      if (two) {
        // (both vectors are streamed from shared memory twice -- dots, then update -- instead of living in
        // registers across the butterflies: the kernel sits at the 255-register limit)
        const double* vsl1 = ring + (size_t)((g + 1) & (kQsRing - 1)) * SL;
        const double tau1 = vsl1[LDV];
        const double* qA = vsl1 + 32 * kb + lane;
        const double* qB = qA - LDV;
        double d0[NS], d1[NS], gg = 0.0;
#pragma unroll
        for (int s2 = 0; s2 < NS; ++s2) { d0[s2] = 0.0; d1[s2] = 0.0; }
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
          const double vi = (i < RPL - kb ? pA : pB)[32 * i], ui = (i < RPL - kb ? qA : qB)[32 * i];
          gg = fma(ui, vi, gg);
#pragma unroll
          for (int s2 = 0; s2 < NS; ++s2) { d0[s2] = fma(vi, a[s2][i], d0[s2]); d1[s2] = fma(ui, a[s2][i], d1[s2]); }
        }
        const double t0 = warp_reduce_ns<NS>(d0, lane);
        const double t1 = warp_reduce_ns<NS>(d1, lane);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gg += __shfl_xor_sync(0xffffffffu, gg, o);
        double w0[NS], w1[NS];
#pragma unroll
        for (int s2 = 0; s2 < NS; ++s2) {
          w0[s2] = tau * __shfl_sync(0xffffffffu, t0, s2 * (32 / NS));
          w1[s2] = tau1 * fma(-gg, w0[s2], __shfl_sync(0xffffffffu, t1, s2 * (32 / NS)));
        }
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
          const double vi = (i < RPL - kb ? pA : pB)[32 * i], ui = (i < RPL - kb ? qA : qB)[32 * i];
#pragma unroll
          for (int s2 = 0; s2 < NS; ++s2) a[s2][i] -= fma(vi, w0[s2], ui * w1[s2]);
        }
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile int*>(&done[wid]) = g + 2;
#ifdef TTB_QS_TRACE2
        if (s.trace && b == 0 && kb == 1 && tid == 0 && g + 1 < 64) { g_qs_trace2[3 * 64 + g] = qs_now(); g_qs_trace2[3 * 64 + g + 1] = qs_now(); g_qs_trace2[2 * 64 + g + 1] = g_qs_trace2[2 * 64 + g]; }
#endif
        if (trace && ((g & 31) == 31 || ((g + 1) & 31) == 31)) g_qs_trace[kb * 32 + 1 + ((g + 1) >> 5)] = qs_now();
        g += 2;
        continue;
      }
      double v[RPL];
#pragma unroll
      for (int i = 0; i < RPL; ++i) v[i] = (i < RPL - kb ? pA : pB)[32 * i];
      double dot[NS];
#pragma unroll
      for (int s2 = 0; s2 < NS; ++s2) dot[s2] = dot_rpl<RPL>(v, a[s2]);
      const double tot = warp_reduce_ns<NS>(dot, lane);
#pragma unroll
      for (int s2 = 0; s2 < NS; ++s2) {
        const double w = tau * __shfl_sync(0xffffffffu, tot, s2 * (32 / NS));
#pragma unroll
        for (int i = 0; i < RPL; ++i) a[s2][i] -= v[i] * w;
      }
      __syncwarp();
      if (lane == 0) *reinterpret_cast<volatile int*>(&done[wid]) = g + 1;
      if (trace && (g & 31) == 31) g_qs_trace[kb * 32 + 1 + (g >> 5)] = qs_now();
#ifdef TTB_QS_TRACE2
      if (s.trace && b == 0 && kb == 1 && tid == 0 && g < 64) g_qs_trace2[3 * 64 + g] = qs_now();
#endif
      ++g;
    }
  }
  // ---- transition: rows above the panel are final R entries; store them, zero the registers
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) {
    const int cs = wid * NS + s2;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      if (i >= RPL - kb) {   // wrapped register: absolute rows < j0
        const int r = ((i + kb) & (RPL - 1)) * 32 + lane;
        if (cs < bwc) Xa[r + p * (j0 + cs)] = a[s2][i];
        a[s2][i] = 0.0;
      }
    }
  }
  if (kb == 0 && d.mode == 0 && tid == 0) c.st[b].mprime[s.slot] = d.k;
  // ---- factor the own panel: the chain of qr_panel.cuh; warp 0 also pushes every reflector to the CTAs
  //      that own later column blocks: lane r serves CTA r, as far as its credits allow, never blocking
  // Warp 0 pushes: lane r serves CTA r (the CTAs after this one), as far as that CTA's credits allow, never
  // blocking.  (Spreading the pushes over the warps was measured slower: every warp then carries the
  // credit polling, and the warp about to take over the chain must not be delayed at all.)
  const bool pusher = wid == 0 && lane > kb && lane < nblk;
  int np = 0;   // own reflectors already pushed to CTA `lane`
  const uint32_t ring_r = pusher ? mapa_u32(smem_u32(ring), (uint32_t)lane) : 0u;
  const uint32_t bars_r = pusher ? mapa_u32(smem_u32(bars), (uint32_t)lane) : 0u;
  auto pump = [&](int nready) {
    if (pusher) {
      while (np < nready) {
        const int slot = (j0 + np) & (kQsRing - 1);
        // relaxed read: the consumer releases the credit AFTER all its warps are done with the slot's previous
        // occupant and after re-arming the barrier; the bulk copy below is issued after this load has returned.
        // (An acquire at cluster scope cost the pusher ~0.5 us per reflector: it could not keep up with the last
        // blocks of a panel -- one reflector every 0.57 us -- and the next panel started ~3 us late.)
        if (TTB_QS_CREDIT_LD(&credit[lane * kQsRing + slot]) <= j0 + np) break;   // slot not armed for it yet
        const uint32_t bar_r = bars_r + (uint32_t)slot * 8u;
        const uint32_t dst_r = ring_r + (uint32_t)(slot * SL * 8);
        const double* src = vs + (size_t)np * SL;
        bulk_s2remote(dst_r + (uint32_t)(j0 * 8), src, (uint32_t)((LDV - j0 + 2) * 8), bar_r);   // rows j0..LDV of the slot, then tau
#ifdef TTB_QS_TRACE2
        if (s.trace && b == 0 && kb == 0 && lane == 1) g_qs_trace2[np] = qs_now();
#endif
        ++np;
      }
    }
    __syncwarp();
  };
#pragma unroll 1
  for (int blk = 0; blk * NS < bwc; ++blk) {
    const int c0 = blk * NS;
    if (wid == blk) {
      panel_owner_step<RPL, NS, 0, true>(a, vs + (size_t)c0 * SL, sh_tau, sh_den, ready, raw, c0, lane, wid, bwc, nvi, LDV - j0);
      if (c0 + 1 < bwc) panel_owner_step<RPL, NS, 1, true>(a, vs + (size_t)(c0 + 1) * SL, sh_tau, sh_den, ready, raw, c0 + 1, lane, wid, bwc, nvi, LDV - j0);
      if (c0 + 2 < bwc) panel_owner_step<RPL, NS, 2, true>(a, vs + (size_t)(c0 + 2) * SL, sh_tau, sh_den, ready, raw, c0 + 2, lane, wid, bwc, nvi, LDV - j0);
      if (c0 + 3 < bwc) panel_owner_step<RPL, NS, 3, true>(a, vs + (size_t)(c0 + 3) * SL, sh_tau, sh_den, ready, raw, c0 + 3, lane, wid, bwc, nvi, LDV - j0);
      if (s.trace && b == 0 && lane == 0) g_qs_trace[kb * 32 + 22 + blk] = qs_now();
      if (wid == 0) pump(min(c0 + NS, bwc));   // warp 0's own block: its reflectors leave once the block is done
    } else if (wid > blk) {
#ifdef TTB_PANEL_DEFER   // measured A/B (round 2): no effect -- the step in the block times at block 4 is NOT sub-partition sharing
      // Warps w and w + 4 share an SM sub-partition (issue port, FP64 pipe): while warp blk is on the chain its
      // partner blk + 4 stays OFF the pipe and applies this block's four reflectors one block later, when the
      // owner is blk + 1 (partner blk + 5).  The trace showed a step, not a slope: 3.0 us per block while the
      // partner consumed (blocks 0-3), 2.3 us once it had nothing left to do (blocks 4-7).
      if (wid == blk + 4) continue;
      if (wid == blk + 3 && blk >= 1) {
        const int d0 = (blk - 1) * NS;
#pragma unroll 1
        for (int cc = d0; cc < d0 + NS; ++cc)
          panel_consume_later<RPL, NS>(a, vs + (size_t)cc * SL, sh_tau, ready, raw, cc, lane, false, nvi);
      }
#endif
      const int cend = min(c0 + NS, bwc);
#ifdef TTB_PANEL_PAIR   // measured (micro_panel2, round 2): hand-offs 15-30 % LONGER -- the pair's dots start at the second raw column and are twice the work
#pragma unroll 1
      for (int cc = c0; cc < cend;) {     // two reflectors at a time: the consuming warps must be faster than the owner (qr_panel.cuh)
        if (cc + 1 < cend) {
          panel_consume_pair<RPL, NS>(a, vs + (size_t)cc * SL, vs + (size_t)(cc + 1) * SL, sh_tau, ready, raw, cc, lane, wid != blk + 1, nvi);
          cc += 2;
        } else {
          panel_consume_later<RPL, NS>(a, vs + (size_t)cc * SL, sh_tau, ready, raw, cc, lane, wid != blk + 1, nvi);
          cc += 1;
        }
      }
#else
#pragma unroll 1
      for (int cc = c0; cc < cend; ++cc)
        panel_consume_later<RPL, NS>(a, vs + (size_t)cc * SL, sh_tau, ready, raw, cc, lane, wid != blk + 1, nvi);
#endif
    } else if (wid == 0) {
      // warp 0 owns finished columns from the second block on: it only keeps the pushes going (V'^T V' is formed
      // after the chain, panel_gram_mma).  Lane 0 polls the flags and the whole warp follows its view (a poll
      // per lane could diverge around the __syncwarp of pump()); no sleep: the last blocks of a panel publish a
      // reflector every ~0.57 us and a pusher that lags delays the next panel by the same amount
      const int cend = min(c0 + NS, bwc);
#pragma unroll 1
      for (int cc = c0; cc < cend; ++cc) {
        unsigned it = 0;
        while (ld_flag(&ready[cc]) == 0) {
          __nanosleep(TTB_QS_PUSH_SLEEP);
          if (++it > (1u << 26)) __trap();
        }
#ifdef TTB_QS_TRACE2
        if (s.trace && b == 0 && kb == 0 && lane == 0) g_qs_trace2[64 + cc] = qs_now();
#endif
        pump(cc + 1);
      }
    }
  }
  if (trace) g_qs_trace[kb * 32 + 20] = qs_now();
  // flush: push what is left, now waiting for credits
  if (wid == 0) {
    for (unsigned it = 0;; ++it) {
      pump(bwc);
      if (!__any_sync(0xffffffffu, pusher && np < bwc)) break;
      if (it > (1u << 24)) __trap();
    }
  }
  __syncthreads();
  // K range of the Gram product = the 32 nvi valid rows: tau' sits right behind them and must stay intact (a bulk
  // push may still be reading it); columns a narrow panel never published are cleared (they are never pushed)
  for (int i = bwc * SL + tid; i < 32 * SL; i += kQsThreads) vs[i] = 0.0;
  __syncthreads();
  panel_gram_mma(vs, SL, 32 * nvi, G, wid, lane, bwc);
  __syncthreads();
  panel_normalise<RPL, NS>(a, G, sh_tau, sh_den, tid, kQsThreads, wid, lane, bwc);   // LAPACK form for the kernels downstream
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) {
    const int cs = wid * NS + s2;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      const int r = lane + 32 * i;
      if (cs < bwc && r < rows) Xp[r + p * cs] = a[s2][i];
    }
  }
  __syncthreads();
  double* Tg = c.T + (int64_t)b * c.t_stride + (int64_t)kb * 1024;
  for (int i = tid; i < 1024; i += 256) Tg[i] = G[i];
  if (trace) g_qs_trace[kb * 32 + 21] = qs_now();
  // shared memory of this CTA is a copy source / credit target for the others until they are done
  cluster_sync_all();
}

}  // namespace ttb
