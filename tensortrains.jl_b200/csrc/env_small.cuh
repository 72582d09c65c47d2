// env_small.cuh -- environments, marginals and the pair-marginal precompute for batches of small-bond
// trains with a matrix-valued boundary (every bond <= 32, d_1 > 1: periodic trains, config 5).
//
// Reference: src/abstract_tensor_train.jl  trace :87, accumulate_L :89-102, accumulate_R :104-117,
//            marginals :208-220, twovar_marginals :245-249.
//
// At this size a site is 16 KiB and the per-site product is a 32 x 32 x 32 GEMM, so the path is bound by
// HBM (every site tensor is read once) as long as enough CTAs are resident to cover the load latency.
// One 128-thread CTA per train (environments) or per (site, train) (marginals / precompute); every
// 32 x 32 operand lives in shared memory column-major with leading dimension 36 (= 4 mod 16), which makes
// the m16n8k8 FP64 fragments bank-conflict free in BOTH orientations (element (r, c) at r + 36 c read as
// A(m = r, k = c), B(k = r, n = c), or transposed), zero padded so no kernel needs a predicate in the MMA.
// ~18 KiB of shared memory and <= 64 registers per CTA keep 8-12 CTAs on an SM, i.e. 128-192 KiB of site
// loads in flight per SM, without an explicit pipeline.
#pragma once

namespace ttb {

constexpr int kSmallLD = 36;

// l[t] = (l[t-1]/max|l[t-1]|) T_t  (LEFT)   /   r[t] = T_t (r[t+1]/max|r[t+1]|)  (RIGHT),  T_t = sum_x A_t(x).
// The stored environments carry the reference's scaling (every stored matrix but the last one computed
// has max-abs 1, :93-97), c.store == 0 skips the stores (normalize! only needs z).
// PF: the raw site tensor travels global -> shared with cp.async (16-byte chunks, no registers) into a
// staging buffer; the copy of site t+1 is issued as soon as site t has been summed into Ts and overlaps the
// MMA, the max reduction and the stores of step t.  Without PF (Q too large for the staging buffer) the
// loads are issued at the top of the step and only other CTAs hide their latency.
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <bool LEFT, bool PF>
__global__ void __launch_bounds__(128, PF ? 6 : 8) k_env_small(EnvCtx c, int store) {
  constexpr int LD = kSmallLD;
  const int b = blockIdx.x;
  const int* bond = c.bond + (int64_t)b * c.L1;
  const double* tt = c.slab + (int64_t)b * c.tt_stride;
  double* env = c.env + (int64_t)b * c.env_stride;
  __shared__ double E[LD * 32];
  __shared__ double Ts[LD * 32];
  __shared__ double red[2][4];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, tq = lane & 3;
  const int mt = wid & 1, nt0 = (wid >> 1) * 2;
  const int r = LEFT ? bond[0] : bond[c.L];
  extern __shared__ __align__(16) double raw[];   // PF: one site tensor at its current dims (<= 1024 Q doubles)
  for (int i = tid; i < LD * 32; i += 128) { E[i] = ((i % LD) == (i / LD) && (i / LD) < r) ? 1.0 : 0.0; Ts[i] = 0.0; }
  auto issue = [&](int t) {
    const double* A = tt + c.off[t];
    const int tot = bond[t] * bond[t + 1] * c.Q;
    for (int i = tid; i < (tot >> 1); i += 128) cp_async16(raw + 2 * i, A + 2 * i);
    if ((tot & 1) && tid == 0) cp_async8(raw + tot - 1, A + tot - 1);
  };
  if (PF) issue(LEFT ? 0 : c.L - 1);
  double logz = 0.0;
  for (int s = 0; s < c.L; ++s) {
    const int t = LEFT ? s : c.L - 1 - s;
    const int dl = bond[t], dr = bond[t + 1];
    const double* A = tt + c.off[t];
    const int64_t nn = (int64_t)dl * dr;
    double tv[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) tv[m] = 0.0;
    if (PF) {
      cp_async_wait_all();
      __syncthreads();   // site t has landed (and the previous step's readers of Ts / writers of E are done)
      for (int x = 0; x < c.Q; ++x) {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int idx = tid + 128 * m, i = idx & 31, j = idx >> 5;
          if (i < dl && j < dr) tv[m] += raw[i + dl * j + (int)nn * x];
        }
      }
    } else {
#pragma unroll 2
      for (int x = 0; x < c.Q; ++x) {   // 8 independent loads per x in flight
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int idx = tid + 128 * m, i = idx & 31, j = idx >> 5;
          if (i < dl && j < dr) tv[m] += A[i + dl * j + nn * x];
        }
      }
    }
    __syncthreads();  // the previous step's readers of Ts / writers of E are done; PF: every thread has read raw
    if (PF && s + 1 < c.L) issue(LEFT ? s + 1 : c.L - 2 - s);
#pragma unroll
    for (int m = 0; m < 8; ++m) { const int idx = tid + 128 * m; Ts[(idx & 31) + LD * (idx >> 5)] = tv[m]; }
    __syncthreads();
    double d[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
    if (LEFT) mma32_warp<1, LD, 1, LD>(d, E, Ts, mt, nt0, g, tq);   // (r x dl) (dl x dr)
    else      mma32_warp<1, LD, 1, LD>(d, Ts, E, mt, nt0, g, tq);   // (dl x dr) (dr x rr)
    const bool last = (s == c.L - 1);
    double inv = 1.0, dv = 0.0;   // dv != 0: divide (1/nt overflows for a subnormal max)
    if (!last) {
      double mx = 0.0;
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) mx = absmax_nan(mx, d[j][i]);
      const double nt = block4_max_bits(mx, red[s & 1], lane, wid);   // one barrier: E's readers are done too
      if (nt != 0.0 && c.normalize) {  // NaN != 0 is true, like !iszero(NaN)
        inv = 1.0 / nt;
        if (isinf(inv) && nt != 0.0) dv = nt;
        logz += log(nt);
      }
    } else {
      __syncthreads();
    }
    // write the new environment: shared (next step's operand) and, scaled like the reference, global
    const int rows = LEFT ? r : dl, cols = LEFT ? dr : r;
    double* eg = env + c.eoff[t];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r0_ = mt * 16 + g, c0_ = (nt0 + j) * 8 + 2 * tq;
      double v0 = d[j][0] * inv, v1 = d[j][1] * inv, v2 = d[j][2] * inv, v3 = d[j][3] * inv;
      if (dv != 0.0) { v0 = d[j][0] / dv; v1 = d[j][1] / dv; v2 = d[j][2] / dv; v3 = d[j][3] / dv; }
      E[r0_ + LD * c0_] = v0; E[r0_ + LD * (c0_ + 1)] = v1;
      E[r0_ + 8 + LD * c0_] = v2; E[r0_ + 8 + LD * (c0_ + 1)] = v3;
      if (store) {
        if (r0_ < rows && c0_ < cols) eg[r0_ + (int64_t)rows * c0_] = v0;
        if (r0_ < rows && c0_ + 1 < cols) eg[r0_ + (int64_t)rows * (c0_ + 1)] = v1;
        if (r0_ + 8 < rows && c0_ < cols) eg[r0_ + 8 + (int64_t)rows * c0_] = v2;
        if (r0_ + 8 < rows && c0_ + 1 < cols) eg[r0_ + 8 + (int64_t)rows * (c0_ + 1)] = v3;
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    const int n = LEFT ? min(r, bond[c.L]) : min(bond[0], r);
    double tr = 0.0;
    for (int a = 0; a < n; ++a) tr += E[a + LD * a];
    c.log_out[b] = logz + log(fabs(tr));
    c.sign_out[b] = tr < 0.0 ? -1 : 1;
  }
}

// marginals: p_t[x] ∝ sum_{b,c} A_t[b,c,x] (r[t+1] l[t-1])[c,b];  one CTA per (site, train)
__global__ void __launch_bounds__(128, 8) k_marginals_small(MargCtx c) {
  constexpr int LD = kSmallLD;
  const int t = blockIdx.x, b = blockIdx.y;
  const int* bond = c.bond + (int64_t)b * c.L1;
  const int dl = bond[t], dr = bond[t + 1], r0 = bond[0], rL = bond[c.L];
  const double* A = c.slab + (int64_t)b * c.tt_stride + c.off[t];
  const double* l = t > 0 ? c.envL + (int64_t)b * c.env_stride + c.eoffL[t - 1] : nullptr;        // r0 x dl
  const double* r = t < c.L - 1 ? c.envR + (int64_t)b * c.env_stride + c.eoffR[t + 1] : nullptr;  // dr x rL
  __shared__ double Ls[LD * 32];   // (a, b) at a + LD b
  __shared__ double Rs[LD * 32];   // (c, a) at c + LD a
  extern __shared__ double psh[];  // 4 * Q partial sums, then Q
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, tq = lane & 3;
  const int mt = wid & 1, nt0 = (wid >> 1) * 2;
  const int na = min(r0, rL);
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int idx = tid + 128 * m, i = idx & 31, j = idx >> 5;
    // l = I (t == 0, dl == r0) and r = I (t == L-1, dr == rL) as explicit identities
    Ls[i + LD * j] = (i < na && j < dl) ? (l ? l[i + (int64_t)r0 * j] : (i == j ? 1.0 : 0.0)) : 0.0;
    Rs[i + LD * j] = (i < dr && j < na) ? (r ? r[i + (int64_t)dr * j] : (i == j ? 1.0 : 0.0)) : 0.0;
  }
  __syncthreads();
  double d[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
  mma32_warp<1, LD, 1, LD>(d, Rs, Ls, mt, nt0, g, tq);   // RL (dr x dl) = r (dr x na) l (na x dl)
  __syncthreads();
  // RL^T into Ls: (b, c) at b + LD c, the layout of A_t(:, :, x)
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int cr = mt * 16 + g, bc = (nt0 + j) * 8 + 2 * tq;
    Ls[bc + LD * cr] = d[j][0]; Ls[bc + 1 + LD * cr] = d[j][1];
    Ls[bc + LD * (cr + 8)] = d[j][2]; Ls[bc + 1 + LD * (cr + 8)] = d[j][3];
  }
  __syncthreads();
  const int64_t nn = (int64_t)dl * dr;
  for (int x = 0; x < c.Q; ++x) {
    double acc = 0.0;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int idx = tid + 128 * m, i = idx & 31, j = idx >> 5;
      if (i < dl && j < dr) acc += A[i + dl * j + nn * x] * Ls[i + LD * j];
    }
    acc = warp_sum(acc);
    if (lane == 0) psh[wid * c.Q + x] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int x = 0; x < c.Q; ++x) {
      const double v = psh[x] + psh[c.Q + x] + psh[2 * c.Q + x] + psh[3 * c.Q + x];
      psh[x] = v;
      tot += v;
    }
    double* o = c.out + ((int64_t)b * c.L + t) * c.Q;
    for (int x = 0; x < c.Q; ++x) o[x] = psh[x] / tot;
  }
}

// pair-marginal precompute: T_u = sum_x A_u(x) and GT_u(x) = (A_u(x) r[u+1])^T, one CTA per (site, train)
__global__ void __launch_bounds__(128, 6) k_twovar_pre_small(TwoPre c) {
  constexpr int LD = kSmallLD;
  const int u = blockIdx.x, b = blockIdx.y;
  const int* bond = c.bond + (int64_t)b * c.L1;
  const int dl = bond[u], dr = bond[u + 1], r0 = bond[0];
  const double* A = c.slab + (int64_t)b * c.tt_stride + c.off[u];
  const double* r = u < c.L - 1 ? c.envR + (int64_t)b * c.env_stride + c.eoffR[u + 1] : nullptr;  // dr x r0
  double* Tm = c.Tm + ((int64_t)b * c.L + u) * c.tstride;
  __shared__ double As[LD * 32];   // A_u(x): (cc, e) at cc + LD e
  __shared__ double Rs[LD * 32];   // r:      (e, a)  at e + LD a
  __shared__ double Tsum[LD * 32]; // sum_x A_u(x)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, tq = lane & 3;
  const int mt = wid & 1, nt0 = (wid >> 1) * 2;
  const int64_t nn = (int64_t)dl * dr;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int idx = tid + 128 * m, i = idx & 31, j = idx >> 5;
    Rs[i + LD * j] = (i < dr && j < r0) ? (r ? r[i + (int64_t)dr * j] : (i == j ? 1.0 : 0.0)) : 0.0;
  }
  for (int x = 0; x < c.Q; ++x) {
    double av[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int idx = tid + 128 * m, i = idx & 31, j = idx >> 5;
      av[m] = (i < dl && j < dr) ? A[i + dl * j + nn * x] : 0.0;
    }
    __syncthreads();  // the previous x's MMA has read As
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int o = ((tid + 128 * m) & 31) + LD * ((tid + 128 * m) >> 5);
      As[o] = av[m];
      Tsum[o] = x == 0 ? av[m] : Tsum[o] + av[m];   // each thread owns its eight entries
    }
    __syncthreads();
    // GT (a, cc) = sum_e r[e, a] A[cc, e]: A-operand r^T (m = a, k = e) at e + LD a, B-operand A^T (k = e, n = cc)
    double d[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
    mma32_warp<LD, 1, LD, 1>(d, Rs, As, mt, nt0, g, tq);
    double* GT = c.GT + (((int64_t)b * c.L + u) * c.Q + x) * c.gstride;  // r0 x dl, [a + r0 cc]
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int a = mt * 16 + g, cc = (nt0 + j) * 8 + 2 * tq;
      const int ccmax = c.padded ? 32 : dl;   // padded: columns >= dl are written as the zeros the MMA produced
      if (a < r0 && cc < ccmax) GT[a + r0 * cc] = d[j][0];
      if (a < r0 && cc + 1 < ccmax) GT[a + r0 * (cc + 1)] = d[j][1];
      if (a + 8 < r0 && cc < ccmax) GT[a + 8 + r0 * cc] = d[j][2];
      if (a + 8 < r0 && cc + 1 < ccmax) GT[a + 8 + r0 * (cc + 1)] = d[j][3];
    }
  }
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int idx = tid + 128 * m, i = idx & 31, j = idx >> 5;
    if (c.padded) Tm[i + 32 * j] = Tsum[i + LD * j];
    else if (i < dl && j < dr) Tm[i + dl * j] = Tsum[i + LD * j];
  }
}

}  // namespace ttb
