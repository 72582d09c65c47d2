// qr_fast.cuh -- the bw = 32 fast path of the blocked Householder QR (included by sweep.cu).
//
//  * the panel factorisation itself is k_qr_panel_flag in qr_panel.cuh.
//  * k_qr_update_tma / k_apply_q_tma : block-reflector application with the V panel staged in shared
//    memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier), one copy per panel column.
// Reference: trailing update and thin-Q formation of the same factorisation (src/tensor_train.jl:161-166:
// C[t] = V' / :200-205: C[t] = U, the orthonormal factor that becomes the new site tensor).
#pragma once

namespace ttb {

// Stage `ncols` columns of `rows` doubles (column stride ld_g in global, `rows` in shared) into smem.
// Bulk (TMA) path when every column start is 16-byte aligned, plain loads otherwise.  Called by all
// threads; returns after the data is visible to the whole CTA.
__device__ __forceinline__ void stage_columns(double* dst, const double* src, int rows, int ncols, int64_t ld_g,
                                              uint64_t* bar, uint32_t& phase, bool aligned) {
  if (aligned) {
    if (threadIdx.x < 32) {
      // one lane arms the barrier, then the copies are issued by the whole warp (one per lane) instead of
      // one thread walking through up to 32 of them
      if (threadIdx.x == 0) {
        fence_proxy_async();
        mbar_expect_tx(bar, (uint32_t)(rows * ncols * 8));
      }
      __syncwarp();
      for (int cc = threadIdx.x; cc < ncols; cc += 32)
        bulk_g2s(dst + (size_t)rows * cc, src + ld_g * cc, (uint32_t)(rows * 8), bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
  } else {
    for (int cc = 0; cc < ncols; ++cc) {
      const double* g = src + ld_g * cc;
      double* sd = dst + (size_t)rows * cc;
      for (int r = threadIdx.x; r < rows; r += blockDim.x) sd[r] = g[r];
    }
    __syncthreads();
  }
}

// Sum 16 per-lane values over the warp with 16 shuffles instead of 80: every level sends the half of the
// values the partner keeps.  Afterwards lane l (and its neighbour l^1) hold the total of value
// index (l >> 1) & 15.
__device__ __forceinline__ double warp_reduce16(double (&a)[16], int lane) {
  double b8[8], b4[4], b2[2];
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double send = hi ? a[k] : a[k + 8], keep = hi ? a[k + 8] : a[k];
      b8[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double send = hi ? b8[k] : b8[k + 4], keep = hi ? b8[k + 4] : b8[k];
      b4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double send = hi ? b4[k] : b4[k + 2], keep = hi ? b4[k + 2] : b4[k];
      b2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  const bool hi = lane & 2;
  const double send = hi ? b2[0] : b2[1], keep = hi ? b2[1] : b2[0];
  double r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

// ---------------------------------------------------------------------------------------------
// Block reflector on a chunk held in shared memory:  Wc <- (I - V op(T) V^T) Wc,  op = T^T if TRANS.
// Vs: rows x bwc unit-lower-trapezoidal (explicit 1 / 0), ld = rows.  Ts: 32 x 32 upper, ld 32.
// Wc: rows x cwc, ld = ldw (cwc <= 4).  256 threads.  W, W2: 32*4 scratch each.
// ---------------------------------------------------------------------------------------------
template <bool TRANS>
__device__ __forceinline__ void block_reflect(const double* Vs, const double* Ts, double* Wc, int ldw, int rows, int bwc,
                                              int cwc, double* W, double* W2) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;  // 8 warps
  // W = V^T Wc : warp handles i in [4*wid, 4*wid+4), all cc
  {
    double acc[4][4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) acc[ii][cc] = 0.0;
    const int i0 = 4 * wid;
    if (i0 < bwc) {
      for (int r = lane; r < rows; r += 32) {
        double vv[4], ww[4];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) vv[ii] = (i0 + ii < bwc) ? Vs[r + (size_t)rows * (i0 + ii)] : 0.0;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) ww[cc] = (cc < cwc) ? Wc[r + (size_t)ldw * cc] : 0.0;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) acc[ii][cc] += vv[ii] * ww[cc];
      }
      double flat[16];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) flat[ii * 4 + cc] = acc[ii][cc];
      const double t = warp_reduce16(flat, lane);
      if (!(lane & 1)) {
        const int idx = lane >> 1;  // value index = ii * 4 + cc
        W[(i0 + (idx >> 2)) + 32 * (idx & 3)] = t;
      }
    }
  }
  __syncthreads();
  // W2 = op(T) W without ever forming T: T^{-1} = diag(1/tau) + striu(V^T V) (compact-WY identity), so
  // op(T) w is a 32-step triangular substitution on the Gram matrix the panel kernel produced for
  // free.  One warp per right-hand side, lane = row.  Ts[lane + 32 i] must be G[lane][i] for the
  // backward solve (T w) and G[i][lane] for the forward solve (T^T w; the caller stages G transposed);
  // the diagonal holds tau.  tau = 0 (H = I) gives y_i = 0, like a zero row/column of T.
  if (wid < 4) {
    const int cc = wid;
    double w = (lane < bwc && cc < cwc) ? W[lane + 32 * cc] : 0.0;
    double y = 0.0;
    if (TRANS) {
      for (int i = 0; i < bwc; ++i) {
        const double yi = __shfl_sync(0xffffffffu, w, i) * Ts[i + 32 * i];
        if (lane > i) w -= Ts[lane + 32 * i] * yi;
        if (lane == i) y = yi;
      }
    } else {
      for (int i = bwc - 1; i >= 0; --i) {
        const double yi = __shfl_sync(0xffffffffu, w, i) * Ts[i + 32 * i];
        if (lane < i) w -= Ts[lane + 32 * i] * yi;
        if (lane == i) y = yi;
      }
    }
    W2[lane + 32 * cc] = y;
  }
  __syncthreads();
  for (int r = tid; r < rows; r += 256) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int imax = min(r, bwc - 1);
    for (int i = 0; i <= imax; ++i) {
      const double vv = Vs[r + (size_t)rows * i];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) acc[cc] += vv * W2[i + 32 * cc];
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
      if (cc < cwc) Wc[r + (size_t)ldw * cc] -= acc[cc];
  }
  __syncthreads();
}

// make the staged panel explicit unit-lower-trapezoidal
__device__ __forceinline__ void fix_unit_lower(double* Vs, int rows, int bwc) {
  for (int i = threadIdx.x; i < bwc * bwc; i += blockDim.x) {
    const int r = i % bwc, cc = i / bwc;
    if (r < cc) Vs[r + (size_t)rows * cc] = 0.0;
    else if (r == cc) Vs[r + (size_t)rows * cc] = 1.0;
  }
}

// trailing update  A <- (I - V T^T V^T) A  for one chunk of <= 4 columns (bw == 32)
__global__ void __launch_bounds__(256) k_qr_update_tma(SweepCtx c, StepDesc s, int panel) {
  pdl_enter();
  const int b = blockIdx.y;
  const Dims d = get_dims(c, s, b);
  if (d.mode == 2) return;
  const int j0 = panel * 32;
  if (j0 >= d.k) return;
  const int bwc = min(32, d.k - j0);
  const int rows = d.p - j0;
  const int jc0 = j0 + bwc + blockIdx.x * 4;
  if (jc0 >= d.n) return;
  const int cwc = min(4, d.n - jc0);
  const int64_t p = d.p;
  double* X = xbuf(c, s.xin, b);
  extern __shared__ __align__(128) double smx[];
  double* Ts = smx;               // 1024
  double* W = Ts + 1024;          // 128
  double* W2 = W + 128;           // 128
  double* As = W2 + 128;          // rows*4
  double* Vs = As + (size_t)rows * 4;  // rows*32
  __shared__ __align__(8) uint64_t bar;
  uint32_t phase = 0;
  const int tid = threadIdx.x;
  if (tid == 0) mbar_init(&bar, 1);
  const double* Tg = c.T + (int64_t)b * c.t_stride + (int64_t)panel * 1024;
  __syncthreads();
  const bool aligned = ((p & 1) == 0);
  // one barrier phase covers both the V panel and the chunk; the T factor is fetched while TMA runs
  if (aligned) {
    if (tid < 32) {
      if (tid == 0) {
        fence_proxy_async();
        mbar_expect_tx(&bar, (uint32_t)(rows * (bwc + cwc) * 8));
      }
      __syncwarp();
      // the whole warp issues the copies: lane cc one V column, lanes 0..cwc-1 also one chunk column
      if (tid < bwc) bulk_g2s(Vs + (size_t)rows * tid, X + j0 + p * (j0 + tid), (uint32_t)(rows * 8), &bar);
      if (tid < cwc) bulk_g2s(As + (size_t)rows * tid, X + j0 + p * (jc0 + tid), (uint32_t)(rows * 8), &bar);
    }
    {
      double t4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t4[u] = Tg[tid + 256 * u];
#pragma unroll
      for (int u = 0; u < 4; ++u) {  // staged TRANSPOSED: the forward solve reads G[i][lane]
        const int idx = tid + 256 * u;
        Ts[(idx >> 5) + 32 * (idx & 31)] = t4[u];
      }
    }
    mbar_wait(&bar, phase);
    phase ^= 1u;
  } else {
    for (int i = tid; i < 1024; i += 256) Ts[(i >> 5) + 32 * (i & 31)] = Tg[i];
    stage_columns(Vs, X + j0 + p * j0, rows, bwc, p, &bar, phase, false);
    stage_columns(As, X + j0 + p * jc0, rows, cwc, p, &bar, phase, false);
  }
  fix_unit_lower(Vs, rows, bwc);
  __syncthreads();
  block_reflect<true>(Vs, Ts, As, rows, rows, bwc, cwc, W, W2);
  for (int i = tid; i < rows * cwc; i += 256) {
    const int r = i % rows, cc = i / rows;
    X[(j0 + r) + p * (jc0 + cc)] = As[i];
  }
}

// W = H_1 ... H_k E for one chunk of <= 4 columns, written into the site tensor (bw == 32)
__global__ void __launch_bounds__(256) k_apply_q_tma(SweepCtx c, StepDesc s) {
  pdl_enter();
  const int b = blockIdx.y;
  const Dims d = get_dims(c, s, b);
  const int mprime = c.st[b].mprime[s.slot];
  const int c0 = blockIdx.x * 4;
  if (c0 >= mprime) return;
  const int cwc = min(4, mprime - c0);
  const int p = d.p;
  const double* X = xbuf(c, s.xin, b);
  extern __shared__ __align__(128) double smx[];
  double* Ts = smx;               // 1024
  double* W = Ts + 1024;          // 128
  double* W2 = W + 128;           // 128
  double* Wc = W2 + 128;          // p*4
  double* Vs = Wc + (size_t)p * 4;  // rows*32
  __shared__ __align__(8) uint64_t bar;
  uint32_t phase = 0;
  const int tid = threadIdx.x;
  if (tid == 0) mbar_init(&bar, 1);
  const int* perm = c.perm + (int64_t)b * c.n_max;
  const double* Jv = c.Jv + (int64_t)b * c.jv_stride;
  for (int i = tid; i < p * cwc; i += 256) {
    const int r = i % p, cc = i / p;
    double v;
    if (d.mode == 0) v = (r == c0 + cc) ? 1.0 : 0.0;
    else v = r < d.k ? Jv[(int64_t)perm[c0 + cc] * c.ldj + r] : 0.0;
    Wc[i] = v;
  }
  __syncthreads();
  if (d.mode != 2) {
    const bool aligned = ((p & 1) == 0);
    const int np = (d.k + 31) / 32;
    for (int panel = np - 1; panel >= 0; --panel) {
      const int j0 = panel * 32;
      if (d.mode == 0 && j0 > c0 + cwc - 1) continue;
      const int bwc = min(32, d.k - j0);
      const int rows = p - j0;
      const double* Tg = c.T + (int64_t)b * c.t_stride + (int64_t)panel * 1024;
      for (int i = tid; i < 1024; i += 256) Ts[i] = Tg[i];
      stage_columns(Vs, X + j0 + (int64_t)p * j0, rows, bwc, p, &bar, phase, aligned);
      fix_unit_lower(Vs, rows, bwc);
      __syncthreads();
      block_reflect<false>(Vs, Ts, Wc + j0, p, rows, bwc, cwc, W, W2);
    }
  }
  double* site = c.slab + (int64_t)b * c.tt_stride + c.off[s.site];
  if (s.dir == 0) {
    for (int i = tid; i < p * cwc; i += 256) {
      const int cc = i % cwc, r = i / cwc;
      site[(c0 + cc) + (int64_t)mprime * r] = Wc[r + (size_t)p * cc];
    }
  } else {
    const int dl = p / c.Q;
    for (int i = tid; i < p * cwc; i += 256) {
      const int r = i % p, cc = i / p;
      const int m = r % dl, x = r / dl;
      site[m + (int64_t)dl * ((c0 + cc) + (int64_t)mprime * x)] = Wc[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// W = H_1 ... H_k E for a SMALL step (p <= 64 rows, one panel: k <= 32, m' <= 32 columns): one CTA forms
// the whole orthonormal factor of its train instead of eight CTAs staging the same panel for four columns
// each (a batch of config-5 trains launches 4096 of these per step).  Same algebra as block_reflect<false>
// -- W = E - V (T (V^T E)), T applied as a back substitution on diag(1/tau) + striu(V^T V) -- but the two
// products run on the FP64 tensor cores: operands in shared memory with leading dimensions 68 / 36
// (= 4 mod 16, conflict-free m16n8k8 fragments in both orientations), zero padded to 64 x 32.
// Reference: C[t] = V' / C[t] = U (src/tensor_train.jl:165-166, :204-205).
// ---------------------------------------------------------------------------------------------
constexpr size_t kApplySmallSmem = sizeof(double) * (2 * 68 * 32 + 1024 + 36 * 32);
__global__ void __launch_bounds__(256) k_apply_q_small(SweepCtx c, StepDesc s) {
  pdl_enter();
  constexpr int LDV = 68, LDW = 36;
  const int b = blockIdx.x;
  const Dims d = get_dims(c, s, b);
  const int mprime = c.st[b].mprime[s.slot];
  const int p = d.p;
  const double* X = xbuf(c, s.xin, b);
  extern __shared__ __align__(16) double smq[];   // kApplySmallSmem bytes
  double* Vs = smq;                 // (r, i)  at r + LDV i
  double* Wc = Vs + LDV * 32;       // (r, cc) at r + LDV cc
  double* Ts = Wc + LDV * 32;       // 32 x 32
  double* W = Ts + 1024;            // (i, cc) at i + LDW cc
  double* W2 = W;                   // the solve keeps its columns in registers: in place
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, tq = lane & 3;
  const int* perm = c.perm + (int64_t)b * c.n_max;
  const double* Jv = c.Jv + (int64_t)b * c.jv_stride;
  const int bwc = d.mode == 2 ? 0 : d.k;
  for (int i = tid; i < 64 * 32; i += 256) {
    const int r = i & 63, cc = i >> 6;
    double v = 0.0, w = 0.0;
    if (r < p && cc < mprime) {
      if (d.mode == 0) v = (r == cc) ? 1.0 : 0.0;
      else v = r < d.k ? Jv[(int64_t)perm[cc] * c.ldj + r] : 0.0;
    }
    if (r < p && cc < bwc) w = r < cc ? 0.0 : (r == cc ? 1.0 : X[r + (int64_t)p * cc]);  // explicit unit lower
    Wc[r + LDV * cc] = v;
    Vs[r + LDV * cc] = w;
  }
  if (d.mode != 2) {
    const double* Tg = c.T + (int64_t)b * c.t_stride;
    for (int i = tid; i < 1024; i += 256) Ts[i] = Tg[i];
    __syncthreads();
    {  // W (32 x 32) = V^T Wc, K = 64 rows: warp = one 16 x 8 tile
      const int mt = wid & 1, nt = wid >> 1;
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const double* pa = Vs + (mt * 16 + g) * LDV + (ks * 8 + tq);
        const double af[4] = {pa[0], pa[8 * LDV], pa[4], pa[8 * LDV + 4]};
        const double* pb = Wc + (ks * 8 + tq) + (nt * 8 + g) * LDV;
        const double bf[2] = {pb[0], pb[4]};
        dmma_16x8x8(acc, af, bf);
      }
      double* wo = W + (mt * 16 + g) + LDW * (nt * 8 + 2 * tq);
      wo[0] = acc[0]; wo[LDW] = acc[1]; wo[8] = acc[2]; wo[8 + LDW] = acc[3];
    }
    __syncthreads();
    {  // W2 = T W: four right-hand sides per warp, lane = row, interleaved for ILP
      double w[4], y[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = (lane < bwc && wid + 8 * q < mprime) ? W[lane + LDW * (wid + 8 * q)] : 0.0;
      for (int i = bwc - 1; i >= 0; --i) {
        const double tau = Ts[i + 32 * i];
        const double gi = lane < i ? Ts[lane + 32 * i] : 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double yi = __shfl_sync(0xffffffffu, w[q], i) * tau;
          w[q] -= gi * yi;
          if (lane == i) y[q] = yi;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) W2[lane + LDW * (wid + 8 * q)] = y[q];
    }
    __syncthreads();
    {  // Wc (64 x 32) -= V W2, K = 32: warp = 16 rows x 16 columns
      const int mt = wid & 3, nt0 = (wid >> 2) * 2;
      double acc[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const double* pa = Vs + (mt * 16 + g) + LDV * (ks * 8 + tq);
        const double af[4] = {pa[0], pa[8], pa[4 * LDV], pa[8 + 4 * LDV]};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const double* pb = W2 + (ks * 8 + tq) + LDW * ((nt0 + j) * 8 + g);
          const double bf[2] = {pb[0], pb[4]};
          dmma_16x8x8(acc[j], af, bf);
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double* wo = Wc + (mt * 16 + g) + LDV * ((nt0 + j) * 8 + 2 * tq);
        wo[0] -= acc[j][0]; wo[LDV] -= acc[j][1]; wo[8] -= acc[j][2]; wo[8 + LDV] -= acc[j][3];
      }
    }
  }
  __syncthreads();
  double* site = c.slab + (int64_t)b * c.tt_stride + c.off[s.site];
  if (s.dir == 0) {
    for (int i = tid; i < p * mprime; i += 256) {
      const int cc = i % mprime, r = i / mprime;
      site[cc + (int64_t)mprime * r] = Wc[r + LDV * cc];
    }
  } else {
    const int dl = p / c.Q;
    for (int i = tid; i < p * mprime; i += 256) {
      const int r = i % p, cc = i / p;
      const int m = r % dl, x = r / dl;
      site[m + (int64_t)dl * (cc + (int64_t)mprime * x)] = Wc[r + LDV * cc];
    }
  }
}

}  // namespace ttb
