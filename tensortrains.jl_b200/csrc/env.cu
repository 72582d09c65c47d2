// env.cu -- left/right environments and everything built on them.
//
// Reference: src/abstract_tensor_train.jl  trace :87, accumulate_L :89-102, accumulate_R :104-117,
//            normalization :154-157, normalize_eachmatrix! :165-176, normalize! :184-197,
//            marginals :208-220, twovar_marginals :231-254 (+ accumulate_M :119-137).
//
// The physical-index sum (`trace`) is fused into the environment step: l_t = (l_{t-1}/max|l_{t-1}|) *
// sum_x A_t(x) reads every site tensor exactly once.  The running scale is carried on the device as
// log|z| + sign, like the reference's Logarithmic accumulator.
#include <algorithm>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tma_mma.cuh"

namespace ttb {

struct EnvCtx {
  const double* slab; int64_t tt_stride; const int64_t* off; const int* bond; int L, L1, Q;
  double* env; int64_t env_stride; const int64_t* eoff;  // eoff: L entries, capacity-based offsets
  double* log_out; int* sign_out;
  int normalize;
};

#define ENV_AC 8  // environment rows handled per pass

// l[t] (r x d_{t+1}) = (l[t-1] / nt) * sum_x A_t(:,:,x);  one CTA per train
__global__ void __launch_bounds__(1024) k_accumulate_L(EnvCtx c) {
  const int b = blockIdx.x;
  const int* bond = c.bond + (int64_t)b * c.L1;
  const double* tt = c.slab + (int64_t)b * c.tt_stride;
  double* env = c.env + (int64_t)b * c.env_stride;
  __shared__ double red[40];
  const int tid = threadIdx.x, nt_ = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt_ >> 5;
  const int r = bond[0];
  double logz = 0.0;
  int sign = 1;
  for (int t = 0; t < c.L; ++t) {
    const int dl = bond[t], dr = bond[t + 1];
    const double* A = tt + c.off[t];
    double* lprev = t > 0 ? env + c.eoff[t - 1] : nullptr;  // r x dl
    double* lcur = env + c.eoff[t];                          // r x dr
    if (t > 0) {
      double mx = 0.0;
      for (int i = tid; i < r * dl; i += nt_) mx = absmax_nan(mx, lprev[i]);
      // NaN-propagating block max through ordered bits
      unsigned long long bits = (unsigned long long)__double_as_longlong(mx);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ot = __shfl_xor_sync(0xffffffffu, bits, o);
        bits = ot > bits ? ot : bits;
      }
      __syncthreads();
      if (lane == 0) red[wid] = __longlong_as_double((long long)bits);
      __syncthreads();
      double nt = 0.0;
      {
        unsigned long long bb = 0ull;
        for (int w2 = 0; w2 < nw; ++w2) {
          const unsigned long long v = (unsigned long long)__double_as_longlong(red[w2]);
          bb = v > bb ? v : bb;
        }
        nt = __longlong_as_double((long long)bb);
      }
      if (nt != 0.0 && c.normalize) {  // NaN != 0 is true, like !iszero(NaN)
        for (int i = tid; i < r * dl; i += nt_) lprev[i] /= nt;
        logz += log(nt);
      }
      __syncthreads();
    }
    for (int a0 = 0; a0 < r; a0 += ENV_AC) {
      const int ac = min(ENV_AC, r - a0);
      for (int j = wid; j < dr; j += nw) {
        double acc[ENV_AC];
#pragma unroll
        for (int a = 0; a < ENV_AC; ++a) acc[a] = 0.0;
        for (int i = lane; i < dl; i += 32) {
          double ts = 0.0;
          for (int x = 0; x < c.Q; ++x) ts += A[i + (int64_t)dl * (j + (int64_t)dr * x)];
          if (t == 0) {
#pragma unroll
            for (int a = 0; a < ENV_AC; ++a)
              if (a < ac && a0 + a == i) acc[a] += ts;  // l[-1] = I
          } else {
#pragma unroll
            for (int a = 0; a < ENV_AC; ++a)
              if (a < ac) acc[a] += lprev[(a0 + a) + (int64_t)r * i] * ts;
          }
        }
#pragma unroll
        for (int a = 0; a < ENV_AC; ++a) {
          const double v = warp_sum(acc[a]);
          if (lane == 0 && a < ac) lcur[(a0 + a) + (int64_t)r * j] = v;
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    const double* ll = env + c.eoff[c.L - 1];
    const int dr = bond[c.L];
    double tr = 0.0;
    for (int a = 0; a < min(r, dr); ++a) tr += ll[a + (int64_t)r * a];
    if (tr < 0.0) sign = -sign;
    c.log_out[b] = logz + log(fabs(tr));
    c.sign_out[b] = sign;
  }
}

// r[t] (d_t x rr) = sum_x A_t(:,:,x) * (r[t+1] / nt);  one CTA per train
__global__ void __launch_bounds__(1024) k_accumulate_R(EnvCtx c) {
  const int b = blockIdx.x;
  const int* bond = c.bond + (int64_t)b * c.L1;
  const double* tt = c.slab + (int64_t)b * c.tt_stride;
  double* env = c.env + (int64_t)b * c.env_stride;
  extern __shared__ double part[];  // [jc][ENV_AC][dl]
  __shared__ double red[40];
  const int tid = threadIdx.x, nt_ = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt_ >> 5;
  const int rr = bond[c.L];
  double logz = 0.0;
  int sign = 1;
  for (int t = c.L - 1; t >= 0; --t) {
    const int dl = bond[t], dr = bond[t + 1];
    const double* A = tt + c.off[t];
    double* rnext = t < c.L - 1 ? env + c.eoff[t + 1] : nullptr;  // dr x rr
    double* rcur = env + c.eoff[t];                                // dl x rr
    if (t < c.L - 1) {
      double mx = 0.0;
      for (int i = tid; i < dr * rr; i += nt_) mx = absmax_nan(mx, rnext[i]);
      unsigned long long bits = (unsigned long long)__double_as_longlong(mx);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ot = __shfl_xor_sync(0xffffffffu, bits, o);
        bits = ot > bits ? ot : bits;
      }
      __syncthreads();
      if (lane == 0) red[wid] = __longlong_as_double((long long)bits);
      __syncthreads();
      unsigned long long bb = 0ull;
      for (int w2 = 0; w2 < nw; ++w2) {
        const unsigned long long v = (unsigned long long)__double_as_longlong(red[w2]);
        bb = v > bb ? v : bb;
      }
      const double nt = __longlong_as_double((long long)bb);
      if (nt != 0.0 && c.normalize) {
        for (int i = tid; i < dr * rr; i += nt_) rnext[i] /= nt;
        logz += log(nt);
      }
      __syncthreads();
    }
    // thread (i, g): row i, column group g; partial sums over the groups are reduced through smem
    const int dl32 = (dl + 31) & ~31;
    for (int a0 = 0; a0 < rr; a0 += ENV_AC) {
      const int ac = min(ENV_AC, rr - a0);
      if (dl32 <= nt_) {
        const int G = max(1, min(nt_ / dl32, dr));
        const int i = tid % dl32, g = tid / dl32;
        const bool active = (i < dl) && (g < G);
        double acc[ENV_AC];
#pragma unroll
        for (int a = 0; a < ENV_AC; ++a) acc[a] = 0.0;
        if (active) {
          const int jlo = (int)((int64_t)dr * g / G), jhi = (int)((int64_t)dr * (g + 1) / G);
          for (int j = jlo; j < jhi; ++j) {
            double ts = 0.0;
            for (int x = 0; x < c.Q; ++x) ts += A[i + (int64_t)dl * (j + (int64_t)dr * x)];
            if (t == c.L - 1) {
#pragma unroll
              for (int a = 0; a < ENV_AC; ++a)
                if (a < ac && a0 + a == j) acc[a] += ts;  // r[L] = I
            } else {
#pragma unroll
              for (int a = 0; a < ENV_AC; ++a)
                if (a < ac) acc[a] += ts * rnext[j + (int64_t)dr * (a0 + a)];
            }
          }
#pragma unroll
          for (int a = 0; a < ENV_AC; ++a) part[((size_t)g * ENV_AC + a) * dl + i] = acc[a];
        }
        __syncthreads();
        for (int o = tid; o < dl * ac; o += nt_) {
          const int ii = o % dl, a = o / dl;
          double v = 0.0;
          for (int q = 0; q < G; ++q) v += part[((size_t)q * ENV_AC + a) * dl + ii];
          rcur[ii + (int64_t)dl * (a0 + a)] = v;
        }
        __syncthreads();
      } else {
        for (int i = tid; i < dl; i += nt_) {
          double acc[ENV_AC];
#pragma unroll
          for (int a = 0; a < ENV_AC; ++a) acc[a] = 0.0;
          for (int j = 0; j < dr; ++j) {
            double ts = 0.0;
            for (int x = 0; x < c.Q; ++x) ts += A[i + (int64_t)dl * (j + (int64_t)dr * x)];
            if (t == c.L - 1) {
#pragma unroll
              for (int a = 0; a < ENV_AC; ++a)
                if (a < ac && a0 + a == j) acc[a] += ts;
            } else {
#pragma unroll
              for (int a = 0; a < ENV_AC; ++a)
                if (a < ac) acc[a] += ts * rnext[j + (int64_t)dr * (a0 + a)];
            }
          }
#pragma unroll
          for (int a = 0; a < ENV_AC; ++a)
            if (a < ac) rcur[i + (int64_t)dl * (a0 + a)] = acc[a];
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    const double* r0 = env + c.eoff[0];
    const int dl = bond[0];
    double tr = 0.0;
    for (int a = 0; a < min(dl, rr); ++a) tr += r0[a + (int64_t)dl * a];
    if (tr < 0.0) sign = -sign;
    c.log_out[b] = logz + log(fabs(tr));
    c.sign_out[b] = sign;
  }
}

// Phase 1 of the two-phase environment pass (few trains, large sites): T_t = sum_x A_t(:,:,x) for every
// site in parallel.  This is the HBM-streaming part (reads every site tensor once, fully coalesced, the
// whole GPU at once); the sequential chain then runs on the Q-times smaller T's out of L2.
__global__ void k_site_trace(const double* slab, int64_t tt_stride, const int64_t* off, const int* bond, int L1, int Q,
                             double* tr, int64_t tr_stride, const int64_t* toff) {
  const int t = blockIdx.y, b = blockIdx.z;
  const int* bd = bond + (int64_t)b * L1;
  const int64_t n = (int64_t)bd[t] * bd[t + 1];
  const double* A = slab + (int64_t)b * tt_stride + off[t];
  double* o = tr + (int64_t)b * tr_stride + toff[t];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int x = 0; x < Q; ++x) acc += A[i + n * x];
    o[i] = acc;
  }
}

// marginals: p_t[x] ∝ sum_{a,b,c} l[t-1][a,b] A_t[b,c,x] r[t+1][c,a]; one CTA per (site, train)
struct MargCtx {
  const double* slab; int64_t tt_stride; const int64_t* off; const int* bond; int L, L1, Q;
  const double* envL; const double* envR; int64_t env_stride; const int64_t* eoffL; const int64_t* eoffR;
  double* out;  // [batch][L][Q]
};

__global__ void k_marginals(MargCtx c) {
  const int t = blockIdx.x, b = blockIdx.y;
  const int* bond = c.bond + (int64_t)b * c.L1;
  const int dl = bond[t], dr = bond[t + 1], r0 = bond[0], rL = bond[c.L];
  const double* A = c.slab + (int64_t)b * c.tt_stride + c.off[t];
  const double* l = t > 0 ? c.envL + (int64_t)b * c.env_stride + c.eoffL[t - 1] : nullptr;     // r0 x dl
  const double* r = t < c.L - 1 ? c.envR + (int64_t)b * c.env_stride + c.eoffR[t + 1] : nullptr;  // dr x rL
  extern __shared__ double psh[];  // Q
  __shared__ double red[40];
  const int na = min(r0, rL);  // == r0 == rL for valid trains
  for (int x = 0; x < c.Q; ++x) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < dl * dr; i += blockDim.x) {
      const int bb = i % dl, cc = i / dl;
      double rl = 0.0;
      if (l && r) {
        for (int a = 0; a < na; ++a) rl += r[cc + (int64_t)dr * a] * l[a + (int64_t)r0 * bb];
      } else if (l) {         // r = I (dr == rL): rl = l[cc, bb]
        rl = l[cc + (int64_t)r0 * bb];
      } else if (r) {         // l = I (dl == r0): rl = r[cc, bb]
        rl = r[cc + (int64_t)dr * bb];
      } else {                // single site: trace
        rl = (bb == cc) ? 1.0 : 0.0;
      }
      acc += A[i + (int64_t)dl * dr * x] * rl;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) psh[x] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int x = 0; x < c.Q; ++x) tot += psh[x];
    double* o = c.out + ((int64_t)b * c.L + t) * c.Q;
    for (int x = 0; x < c.Q; ++x) o[x] = psh[x] / tot;
  }
}

__global__ void k_scale_trains(double* slab, int64_t tt_stride, const double* factor) {
  const int b = blockIdx.y;
  const double f = factor[b];
  if (f == 0.0) return;
  // tt_stride is a multiple of 16 doubles and the slab is 256-byte aligned: 16-byte accesses, four in flight
  double2* p = reinterpret_cast<double2*>(slab + (int64_t)b * tt_stride);
  const int64_t n2 = tt_stride >> 1, bd = blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * bd * 4 + threadIdx.x; i0 < n2; i0 += (int64_t)gridDim.x * bd * 4) {
    double2 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) if (i0 + k * bd < n2) v[k] = p[i0 + k * bd];
#pragma unroll
    for (int k = 0; k < 4; ++k) if (i0 + k * bd < n2) { v[k].x /= f; v[k].y /= f; p[i0 + k * bd] = v[k]; }
  }
}

// normalize! bookkeeping (src/abstract_tensor_train.jl:186-196): factor = |Z|^(1/L), logz_out, z <- 1
__global__ void k_normalize_prep(const double* envlog, double* logz, int* signz, double* factor, double* logz_out,
                                 int* status, int L, int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const double la = envlog[b];
  factor[b] = exp(la / L);
  logz_out[b] = la - logz[b];
  if (signz[b] < 0) atomicOr(status, 8);  // log of a negative number in the reference
  logz[b] = 0.0;
  signz[b] = 1;
}

// per-site max-abs (normalize_eachmatrix!, src/abstract_tensor_train.jl:165-176)
__global__ void k_site_maxabs(const double* slab, int64_t tt_stride, const int64_t* off, const int* bond, int L1, int Q,
                              double* mx_out) {
  const int t = blockIdx.x, b = blockIdx.y;
  const int* bd = bond + (int64_t)b * L1;
  const int64_t n = (int64_t)bd[t] * bd[t + 1] * Q;
  const double* A = slab + (int64_t)b * tt_stride + off[t];
  __shared__ double red[40];
  double mx = 0.0;
  bool nan = false;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double a = fabs(A[i]);
    nan |= (a != a);
    mx = fmax(mx, a);
  }
  mx = block_max(mx, red);
  const int anynan = __syncthreads_or(nan ? 1 : 0);
  if (threadIdx.x == 0) mx_out[(int64_t)b * (L1 - 1) + t] = anynan ? __longlong_as_double(0x7ff8000000000000LL) : mx;
}

__global__ void k_site_scale(double* slab, int64_t tt_stride, const int64_t* off, const int* bond, int L1, int Q,
                             const double* mx) {
  const int t = blockIdx.x, b = blockIdx.y;
  const int* bd = bond + (int64_t)b * L1;
  const int64_t n = (int64_t)bd[t] * bd[t + 1] * Q;
  double* A = slab + (int64_t)b * tt_stride + off[t];
  const double m = mx[(int64_t)b * (L1 - 1) + t];
  if ((m != m) || isinf(m) || m == 0.0) return;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) A[i] /= m;
}

__global__ void k_eachmatrix_z(const double* mx, double* logz, int L, int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double s = 0.0;
  for (int t = 0; t < L; ++t) {
    const double m = mx[(int64_t)b * L + t];
    if (!((m != m) || isinf(m) || m == 0.0)) s += log(m);
  }
  logz[b] -= s;
}

// ---------------------------------------------------------------------------------------------
// twovar_marginals (src/abstract_tensor_train.jl:231-254) with accumulate_M (:119-137) streamed:
// one CTA per (t, train) walks u = t+1 .. t+maxdist carrying K = rlAt(x_t) * M[t,u] in shared/global
// scratch, so the L x L table of M matrices is never materialised.
//   G_x = l[t-1] * A_t(x)            (r0 x d_{t+1})   for every x_t
//   for u: b[x_t, x_u] = sum G_x[a, c] * A_u[c, e, x_u] * r[u+1][e, a];  then G_x <- G_x * T_u / max
// (the per-(t,u) max-abs rescale of M only changes a common factor of b[t,u], which is normalised)
// ---------------------------------------------------------------------------------------------
struct TwoCtx {
  const double* slab; int64_t tt_stride; const int64_t* off; const int* bond; int L, L1, Q;
  const double* envL; const double* envR; int64_t env_stride; const int64_t* eoffL; const int64_t* eoffR;
  double* scratch; int64_t scratch_stride;  // per CTA: 2 * Q * r0max * dmax
  const int64_t* pair_off;                  // [L] index of pair (t, t+1) in the output
  int maxdist;
  double* out;                              // [batch][npairs][Q*Q]
  int64_t npairs;
  int t0;                                   // first site of the range this launch covers (blockIdx.x = t - t0)
};

__global__ void k_twovar(TwoCtx c) {
  const int t = blockIdx.x + c.t0, b = blockIdx.y;
  if (t >= c.L - 1) return;
  const int* bond = c.bond + (int64_t)b * c.L1;
  const double* tt = c.slab + (int64_t)b * c.tt_stride;
  const int r0 = bond[0], Q = c.Q;
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ double red[40];
  extern __shared__ double bsh[];  // Q*Q
  double* G = c.scratch ? c.scratch + ((int64_t)b * (c.L - 1) + t) * c.scratch_stride : bsh + Q * Q;
  double* G2 = G + c.scratch_stride / 2;
  // G_x = l[t-1] * A_t(x)
  {
    const int dl = bond[t], dr = bond[t + 1];
    const double* A = tt + c.off[t];
    const double* l = t > 0 ? c.envL + (int64_t)b * c.env_stride + c.eoffL[t - 1] : nullptr;
    for (int o = tid; o < Q * r0 * dr; o += nt) {
      const int a = o % r0, cc = (o / r0) % dr, x = o / (r0 * dr);
      double acc = 0.0;
      if (l) for (int i = 0; i < dl; ++i) acc += l[a + (int64_t)r0 * i] * A[i + (int64_t)dl * (cc + (int64_t)dr * x)];
      else acc = A[a + (int64_t)dl * (cc + (int64_t)dr * x)];
      G[o] = acc;
    }
  }
  __syncthreads();
  int dcur = bond[t + 1];
  const int uhi = min(c.L - 1, t + c.maxdist);
  for (int u = t + 1; u <= uhi; ++u) {
    const int dl = bond[u], dr = bond[u + 1];  // dl == dcur
    const double* A = tt + c.off[u];
    const double* r = u < c.L - 1 ? c.envR + (int64_t)b * c.env_stride + c.eoffR[u + 1] : nullptr;  // dr x rL
    // b[xt, xu] = sum_{a,cc,e} G[a,cc,xt] * A[cc,e,xu] * r[e,a]
    for (int pq = 0; pq < Q * Q; ++pq) {
      const int xt = pq % Q, xu = pq / Q;
      double acc = 0.0;
      for (int o = tid; o < dl * dr; o += nt) {
        const int cc = o % dl, e = o / dl;
        double gr = 0.0;
        if (r) for (int a = 0; a < r0; ++a) gr += G[a + (int64_t)r0 * (cc + (int64_t)dl * xt)] * r[e + (int64_t)dr * a];
        else gr = G[e + (int64_t)r0 * (cc + (int64_t)dl * xt)];  // r = I, a == e
        acc += gr * A[cc + (int64_t)dl * (e + (int64_t)dr * xu)];
      }
      acc = block_sum(acc, red);
      if (tid == 0) bsh[pq] = acc;
    }
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int i = 0; i < Q * Q; ++i) tot += bsh[i];
      double* o = c.out + ((int64_t)b * c.npairs + c.pair_off[t] + (u - t - 1)) * Q * Q;
      for (int i = 0; i < Q * Q; ++i) o[i] = bsh[i] / tot;
    }
    if (u < uhi) {
      // G2_x = G_x * T_u, T_u = sum_x A_u(x);  then rescale by max-abs
      double mx = 0.0;
      for (int o = tid; o < Q * r0 * dr; o += nt) {
        const int a = o % r0, e = (o / r0) % dr, x = o / (r0 * dr);
        double acc = 0.0;
        for (int cc = 0; cc < dl; ++cc) {
          double ts = 0.0;
          for (int xx = 0; xx < Q; ++xx) ts += A[cc + (int64_t)dl * (e + (int64_t)dr * xx)];
          acc += G[a + (int64_t)r0 * (cc + (int64_t)dl * x)] * ts;
        }
        G2[o] = acc;
        mx = fmax(mx, fabs(acc));
      }
      mx = block_max(mx, red);
      if (mx > 0.0 && !isinf(mx))
        for (int o = tid; o < Q * r0 * dr; o += nt) G2[o] /= mx;
      __syncthreads();
      double* tmp = G; G = G2; G2 = tmp;
      dcur = dr;
    }
    __syncthreads();
  }
  (void)dcur;
}

}  // namespace ttb
#include "twovar_fast.cuh"
#include "twovar_mma.cuh"
#include "twovar_open.cuh"
#include "env_fast.cuh"
#include "env_small.cuh"
namespace ttb {

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct EnvOffsets {
  std::vector<int64_t> eoffL, eoffR;
};

static int env_offsets(ttb_tt* h, EnvOffsets& eo) {
  const int L = (int)h->L;
  eo.eoffL.assign(L + 1, 0);
  eo.eoffR.assign(L + 1, 0);
  for (int t = 0; t < L; ++t) {
    eo.eoffL[t + 1] = eo.eoffL[t] + h->bond0[0] * h->bond0[t + 1];
    eo.eoffR[t + 1] = eo.eoffR[t] + h->bond0[t] * h->bond0[L];
  }
  return TTB_OK;
}

int ensure_env_ws(ttb_tt* h) {
  Workspace& w = h->ws;
  cudaError_t e = cudaSuccess;
  if (!w.env_ready) {
    EnvOffsets eo;
    env_offsets(h, eo);
    w.env_stride = std::max(eo.eoffL[h->L], eo.eoffR[h->L]);
    auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    A((void**)&w.envL, sizeof(double) * h->batch * w.env_stride);
    A((void**)&w.envR, sizeof(double) * h->batch * w.env_stride);
    A((void**)&w.env_log, sizeof(double) * h->batch * 4);  // [0]=L log, [1]=R log, [2]=factor, [3]=logz_out
    A((void**)&w.env_sign, sizeof(int) * h->batch * 2);
    if (e != cudaSuccess) return fail(TTB_ENOMEM, "environment workspace: %s", cudaGetErrorString(e));
    w.env_ready = true;
  }
  // the trace scratch of the two-phase pass (launch_acc) is allocated HERE, before any caller starts its timer:
  // a cudaMalloc inside the timed region of the first call on a handle cost up to 90 ms of "device time"
  const bool small = h->dmax <= 32 && h->bond0[0] > 1 && getenv("TTB_ENV_NOSMALL") == nullptr;
  if (!small && h->Q > 1 && h->batch < 296 && getenv("TTB_ENV_ONEPHASE") == nullptr) {
    const int L = (int)h->L;
    std::vector<int64_t> toff(L + 1, 0);
    for (int t = 0; t < L; ++t) toff[t + 1] = toff[t] + ((h->bond0[t] * h->bond0[t + 1] + 15) & ~int64_t(15));
    e = w.scr_trace.ensure(sizeof(double) * h->batch * toff[L]);
    if (e == cudaSuccess && !w.scr_toff.p) {
      e = w.scr_toff.ensure(sizeof(int64_t) * (L + 1));
      if (e == cudaSuccess) {
        cudaMemcpyAsync(w.scr_toff.p, toff.data(), sizeof(int64_t) * (L + 1), cudaMemcpyHostToDevice, h->stream);
        cudaStreamSynchronize(h->stream);
      }
    }
    if (e != cudaSuccess) return fail(TTB_ENOMEM, "environment traces: %s", cudaGetErrorString(e));
  }
  return TTB_OK;
}

// device copies of the environment offsets are tiny; keep them in a side allocation per handle
struct EnvSide {
  int64_t* d_eoffL = nullptr;
  int64_t* d_eoffR = nullptr;
};
static int env_side(ttb_tt* h, EnvSide& s, EnvOffsets& eo) {
  env_offsets(h, eo);
  TTB_CUDA(cudaMalloc(&s.d_eoffL, sizeof(int64_t) * (h->L + 1)));
  TTB_CUDA(cudaMalloc(&s.d_eoffR, sizeof(int64_t) * (h->L + 1)));
  TTB_CUDA(cudaMemcpyAsync(s.d_eoffL, eo.eoffL.data(), sizeof(int64_t) * (h->L + 1), cudaMemcpyHostToDevice, h->stream));
  TTB_CUDA(cudaMemcpyAsync(s.d_eoffR, eo.eoffR.data(), sizeof(int64_t) * (h->L + 1), cudaMemcpyHostToDevice, h->stream));
  TTB_CUDA(cudaStreamSynchronize(h->stream));  // host vectors may go away
  return TTB_OK;
}
static void env_side_free(EnvSide& s) {
  cudaFree(s.d_eoffL);
  cudaFree(s.d_eoffR);
}

static int env_threads(ttb_tt* h) {
  const int64_t d = h->dmax;
  if (h->batch >= 64) return d <= 32 ? 128 : 256;
  return d <= 64 ? 256 : 1024;
}

// small bonds with a matrix-valued boundary (periodic trains, config 5): the tensor-core kernels of
// env_small.cuh, one CTA per train / per (site, train)
static bool small_path(const ttb_tt* h) {
  return h->dmax <= 32 && h->bond0[0] > 1 && getenv("TTB_ENV_NOSMALL") == nullptr;
}

static int launch_acc(ttb_tt* h, bool left, int normalize, const EnvSide& s, int store = 1) {
  Workspace& w = h->ws;
  EnvCtx c;
  c.slab = h->slab; c.tt_stride = h->tt_stride; c.off = h->d_off; c.bond = h->d_bond;
  c.L = (int)h->L; c.L1 = (int)h->L + 1; c.Q = (int)h->Q;
  c.env = left ? w.envL : w.envR; c.env_stride = w.env_stride; c.eoff = left ? s.d_eoffL : s.d_eoffR;
  c.log_out = w.env_log + (left ? 0 : h->batch); c.sign_out = w.env_sign + (left ? 0 : h->batch);
  c.normalize = normalize;
  if (small_path(h)) {
    const size_t raw = sizeof(double) * 1024 * (size_t)h->Q;
    const bool pf = raw <= 64 * 1024 && getenv("TTB_ENV_NOPREFETCH") == nullptr;
    TTB_ONCE_PER_DEVICE(h, 0) {
      cudaFuncSetAttribute(k_env_small<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      cudaFuncSetAttribute(k_env_small<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    }
    if (left && pf) { TTB_LAUNCH(h, "k_env_small_L", (k_env_small<true, true><<<(unsigned)h->batch, 128, raw, h->stream>>>(c, store))); }
    else if (left) { TTB_LAUNCH(h, "k_env_small_L", (k_env_small<true, false><<<(unsigned)h->batch, 128, 0, h->stream>>>(c, store))); }
    else if (pf) { TTB_LAUNCH(h, "k_env_small_R", (k_env_small<false, true><<<(unsigned)h->batch, 128, raw, h->stream>>>(c, store))); }
    else { TTB_LAUNCH(h, "k_env_small_R", (k_env_small<false, false><<<(unsigned)h->batch, 128, 0, h->stream>>>(c, store))); }
    return TTB_OK;
  }
  const int thr = env_threads(h);
  // two-phase pass: parallel physical-index sum first (HBM streaming), then the chain on the traces
  double* d_tr = nullptr;
  int64_t* d_toff = nullptr;
  if (h->Q > 1 && h->batch < 296 && getenv("TTB_ENV_ONEPHASE") == nullptr) {
    const int L = (int)h->L;
    std::vector<int64_t> toff(L + 1, 0);
    for (int t = 0; t < L; ++t) toff[t + 1] = toff[t] + ((h->bond0[t] * h->bond0[t + 1] + 15) & ~int64_t(15));
    // trace buffer and its offset table live on the handle (grow-only scratch; the offsets depend only on
    // the creation-time bonds): no allocation, copy or synchronisation per call
    cudaError_t e = w.scr_trace.ensure(sizeof(double) * h->batch * toff[L]);
    if (e != cudaSuccess) return fail(TTB_ENOMEM, "environment traces: %s", cudaGetErrorString(e));
    d_tr = static_cast<double*>(w.scr_trace.p);
    if (!w.scr_toff.p) {
      e = w.scr_toff.ensure(sizeof(int64_t) * (L + 1));
      if (e != cudaSuccess) return fail(TTB_ENOMEM, "environment traces: %s", cudaGetErrorString(e));
      cudaMemcpyAsync(w.scr_toff.p, toff.data(), sizeof(int64_t) * (L + 1), cudaMemcpyHostToDevice, h->stream);
      cudaStreamSynchronize(h->stream);  // toff is a stack vector (first call only)
    }
    d_toff = static_cast<int64_t*>(w.scr_toff.p);
    const int64_t maxsite = h->dmax * h->dmax;
    const bool open_chain = !h->periodic && h->bond0[0] == 1 && h->bond0[L] == 1;
    if (open_chain) {
      // open train: both directions are coalesced column dots on T (left) or T^T (right)
      if (left) {
        dim3 g((unsigned)std::min<int64_t>((maxsite + 255) / 256, 64), (unsigned)L, (unsigned)h->batch);
        TTB_LAUNCH(h, "k_site_trace", k_site_trace<<<g, 256, 0, h->stream>>>(h->slab, h->tt_stride, h->d_off, h->d_bond,
                                                                       (int)h->L + 1, (int)h->Q, d_tr, toff[L], d_toff));
      } else {
        dim3 g((unsigned)std::min<int64_t>((maxsite + 1023) / 1024, 64), (unsigned)L, (unsigned)h->batch);
        TTB_LAUNCH(h, "k_site_trace_t", k_site_trace_t<<<g, 256, 0, h->stream>>>(h->slab, h->tt_stride, h->d_off, h->d_bond,
                                                                           (int)h->L + 1, (int)h->Q, d_tr, toff[L], d_toff));
      }
      ChainCtx cc;
      cc.T = d_tr; cc.t_stride = toff[L]; cc.toff = d_toff; cc.bond = h->d_bond; cc.L = L; cc.L1 = L + 1;
      cc.env = c.env; cc.env_stride = c.env_stride; cc.eoff = c.eoff; cc.log_out = c.log_out; cc.sign_out = c.sign_out;
      cc.normalize = normalize; cc.left = left ? 1 : 0; cc.dmax = (int)h->dmax;
      int sm_count = 0;
      cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
      if (h->dmax >= 64 && h->batch * kCcCluster <= sm_count && getenv("TTB_ENV_NOCLUSTER") == nullptr) {
        // 8-CTA cluster per train: the per-site trace is read by 8 SMs instead of one
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kCcCluster, (unsigned)h->batch); cfg.blockDim = dim3(kCcThreads);
        cfg.dynamicSmemBytes = sizeof(double) * (2 * (size_t)h->dmax + 2 * kCcCluster); cfg.stream = h->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = kCcCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        TTB_LAUNCH(h, "k_chain_cluster", cudaLaunchKernelEx(&cfg, k_chain_cluster, cc));
      } else {
        TTB_LAUNCH(h, "k_chain_open", k_chain_open<<<(unsigned)h->batch, 1024, sizeof(double) * 2 * h->dmax, h->stream>>>(cc));
      }
      return TTB_OK;
    }
    dim3 g((unsigned)std::min<int64_t>((maxsite + 255) / 256, 64), (unsigned)L, (unsigned)h->batch);
    TTB_LAUNCH(h, "k_site_trace", k_site_trace<<<g, 256, 0, h->stream>>>(h->slab, h->tt_stride, h->d_off, h->d_bond, (int)h->L + 1,
                                                                   (int)h->Q, d_tr, toff[L], d_toff));
    c.slab = d_tr; c.tt_stride = toff[L]; c.off = d_toff; c.Q = 1;
  }
  if (left) {
    TTB_LAUNCH(h, "k_accumulate_L", k_accumulate_L<<<(unsigned)h->batch, thr, 0, h->stream>>>(c));
  } else {
    const size_t smem = sizeof(double) * (size_t)thr * ENV_AC;
    TTB_ONCE_PER_DEVICE(h, 1) cudaFuncSetAttribute(k_accumulate_R, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    TTB_LAUNCH(h, "k_accumulate_R", k_accumulate_R<<<(unsigned)h->batch, thr, smem, h->stream>>>(c));
  }
  return TTB_OK;
}

// copy an environment to the host, repacked at the train's CURRENT dims
static int download_env(ttb_tt* h, bool left, const EnvOffsets& eo, double* out) {
  Workspace& w = h->ws;
  std::vector<double> tmp((size_t)h->batch * w.env_stride);
  TTB_CUDA(cudaMemcpyAsync(tmp.data(), left ? w.envL : w.envR, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost,
                           h->stream));
  TTB_CUDA(cudaStreamSynchronize(h->stream));
  double* p = out;
  const int L = (int)h->L;
  for (int64_t b = 0; b < h->batch; ++b) {
    const int* bd = &h->h_bond[b * (L + 1)];
    for (int t = 0; t < L; ++t) {
      const int64_t n = left ? (int64_t)bd[0] * bd[t + 1] : (int64_t)bd[t] * bd[L];
      const int64_t o = (left ? eo.eoffL[t] : eo.eoffR[t]);
      std::copy(tmp.begin() + b * w.env_stride + o, tmp.begin() + b * w.env_stride + o + n, p);
      p += n;
    }
  }
  return TTB_OK;
}

static int accumulate(ttb_tt* h, bool left, int normalize, double* env_out, double* logabs, int* sign) {
  TTB_TRY(check_handle(h));
  TTB_TRY(ensure_env_ws(h));
  EnvSide s; EnvOffsets eo;
  TTB_TRY(env_side(h, s, eo));
  CallTimer tm(h);
  int rc = launch_acc(h, left, normalize, s);
  if (rc == TTB_OK) rc = tm.finish();
  env_side_free(s);
  TTB_TRY(rc);
  Workspace& w = h->ws;
  if (logabs) TTB_CUDA(cudaMemcpy(logabs, w.env_log + (left ? 0 : h->batch), sizeof(double) * h->batch, cudaMemcpyDeviceToHost));
  if (sign) TTB_CUDA(cudaMemcpy(sign, w.env_sign + (left ? 0 : h->batch), sizeof(int) * h->batch, cudaMemcpyDeviceToHost));
  if (env_out) TTB_TRY(download_env(h, left, eo, env_out));
  return TTB_OK;
}

// l= / r= / rz=: a caller-supplied environment replaces the accumulate pass (it is copied into the workspace slot
// the kernels read: a few hundred KB device to device against one streaming pass over the whole train)
static int env_check(ttb_tt* h, const ttb_env* e, bool left) {
  if (!e) return TTB_OK;
  if (e->device != h->device) return fail(TTB_EINVAL, "environment lives on another device");
  if ((e->left != 0) != left) return fail(TTB_EINVAL, left ? "l= needs an accumulate_L environment" : "r= / rz= needs an accumulate_R environment");
  if (e->L != h->L || e->Q != h->Q || e->batch != h->batch || e->periodic != h->periodic || e->bond0 != h->bond0 ||
      e->bonds != h->h_bond)
    return fail(TTB_EINVAL, "environment does not match the train (length, physical dimension or bond dimensions differ)");
  return TTB_OK;
}

int provide_env(ttb_tt* h, bool left, const ttb_env* e, const int64_t* d_eoffL, const int64_t* d_eoffR) {
  if (!e) {
    EnvSide s;
    s.d_eoffL = const_cast<int64_t*>(d_eoffL); s.d_eoffR = const_cast<int64_t*>(d_eoffR);
    return launch_acc(h, left, 1, s);
  }
  Workspace& w = h->ws;
  TTB_CUDA(cudaMemcpyAsync(left ? w.envL : w.envR, e->env, sizeof(double) * h->batch * w.env_stride, cudaMemcpyDeviceToDevice, h->stream));
  TTB_CUDA(cudaMemcpyAsync(w.env_log + (left ? 0 : h->batch), e->logabs, sizeof(double) * h->batch, cudaMemcpyDeviceToDevice, h->stream));
  TTB_CUDA(cudaMemcpyAsync(w.env_sign + (left ? 0 : h->batch), e->sign, sizeof(int) * h->batch, cudaMemcpyDeviceToDevice, h->stream));
  return TTB_OK;
}

// used by sample.cu (rz=)
int provide_env_R(ttb_tt* h, const ttb_env* rz) {
  TTB_TRY(ensure_env_ws(h));
  TTB_TRY(env_check(h, rz, false));
  CallTimer tm(h);
  TTB_TRY(provide_env(h, false, rz, nullptr, nullptr));
  return tm.finish();
}

// accumulate_M (src/abstract_tensor_train.jl:119-137): row t of the table, M[t,t+1] = I, M[t,u] = M[t,u-1] T_{u-1}
// max-abs rescaled; T_s = sum_x A_s(x) precomputed by k_site_trace.  One CTA per t; the running matrix ping-pongs
// in global memory.  (The product path -- twovar_marginals -- streams this table and never stores it; this entry
// exists for callers of the reference API and for parity tests.)
struct AccM {
  const double* T; const int64_t* toff; const int* bond; int L;
  double* out; const int64_t* moff;   // moff[t]: offset of M[t, t+1] in out; the row's matrices follow each other
  double* scratch; int64_t sstride;   // [L][2][dmax*dmax]
  int normalize;
};
__global__ void k_accumulate_M(AccM c) {
  const int t = blockIdx.x;                  // 0-based row: pairs (t, u), u = t+1 .. L-1
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ double red[40];
  const int dr = c.bond[t + 1];              // rows of every M[t, u]
  double* o = c.out + c.moff[t];
  for (int i = tid; i < dr * dr; i += nt) o[i] = (i % dr == i / dr) ? 1.0 : 0.0;   // M[t, t+1] = I (d_{t+1} x d_{t+1})
  const double* prev = o;
  int dc = dr;
  o += (int64_t)dr * dr;
  __syncthreads();
  for (int u = t + 2; u < c.L; ++u) {
    const double* Tm = c.T + c.toff[u - 1];  // trace of site u-1: d_{u-1} x d_u, column-major
    const int dn = c.bond[u];
    double mx = 0.0;
    for (int i = tid; i < dr * dn; i += nt) {
      const int r = i % dr, cc = i / dr;
      double acc = 0.0;
      for (int k = 0; k < dc; ++k) acc += prev[r + (int64_t)dr * k] * Tm[k + (int64_t)dc * cc];
      o[i] = acc;
      mx = fmax(mx, fabs(acc));
    }
    mx = block_max(mx, red);
    if (c.normalize && mx != 0.0)
      for (int i = tid; i < dr * dn; i += nt) o[i] /= mx;
    __syncthreads();
    prev = o; dc = dn;
    o += (int64_t)dr * dn;
  }
}

static int marginals_impl(ttb_tt* h, const ttb_env* l, const ttb_env* r, double* out);
static int twovar_impl(ttb_tt* h, const ttb_env* l, const ttb_env* r, int64_t maxdist, int64_t t_lo, int64_t t_hi, double* out);

}  // namespace ttb

using namespace ttb;

extern "C" {

int ttb_env_create(ttb_handle h, int left, int normalize, ttb_env_handle* out) {
  TTB_TRY(check_handle(h));
  if (!out) return fail(TTB_EINVAL, "null argument");
  TTB_TRY(accumulate(h, left != 0, normalize, nullptr, nullptr, nullptr));
  Workspace& w = h->ws;
  ttb_env* e = new ttb_env();
  e->device = h->device; e->left = left ? 1 : 0; e->normalize = normalize;
  e->L = h->L; e->Q = h->Q; e->batch = h->batch; e->periodic = h->periodic; e->env_stride = w.env_stride;
  e->bond0 = h->bond0; e->bonds = h->h_bond;
  cudaError_t ce = cudaMalloc(&e->env, sizeof(double) * h->batch * w.env_stride);
  if (ce == cudaSuccess) ce = cudaMalloc(&e->logabs, sizeof(double) * h->batch);
  if (ce == cudaSuccess) ce = cudaMalloc(&e->sign, sizeof(int) * h->batch);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->env, left ? w.envL : w.envR, sizeof(double) * h->batch * w.env_stride, cudaMemcpyDeviceToDevice, h->stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->logabs, w.env_log + (left ? 0 : h->batch), sizeof(double) * h->batch, cudaMemcpyDeviceToDevice, h->stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->sign, w.env_sign + (left ? 0 : h->batch), sizeof(int) * h->batch, cudaMemcpyDeviceToDevice, h->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
  if (ce != cudaSuccess) {
    cudaFree(e->env); cudaFree(e->logabs); cudaFree(e->sign);
    delete e;
    return fail(TTB_ENOMEM, "environment handle: %s", cudaGetErrorString(ce));
  }
  *out = e;
  return TTB_OK;
}

int ttb_env_destroy(ttb_env_handle e) {
  if (!e) return TTB_OK;
  cudaSetDevice(e->device);
  cudaFree(e->env); cudaFree(e->logabs); cudaFree(e->sign);
  delete e;
  return TTB_OK;
}

int ttb_env_z(ttb_env_handle e, double* logabs, int* sign) {
  if (!e) return fail(TTB_EINVAL, "null environment");
  cudaSetDevice(e->device);
  if (logabs) TTB_CUDA(cudaMemcpy(logabs, e->logabs, sizeof(double) * e->batch, cudaMemcpyDeviceToHost));
  if (sign) TTB_CUDA(cudaMemcpy(sign, e->sign, sizeof(int) * e->batch, cudaMemcpyDeviceToHost));
  return TTB_OK;
}

int ttb_marginals_env(ttb_handle h, ttb_env_handle l, ttb_env_handle r, double* out) { return marginals_impl(h, l, r, out); }

int ttb_twovar_marginals_env(ttb_handle h, ttb_env_handle l, ttb_env_handle r, int64_t maxdist, int64_t t_lo, int64_t t_hi,
                             double* out) {
  if (!h) return fail(TTB_EINVAL, "null argument");
  return twovar_impl(h, l, r, maxdist, t_lo, t_hi, out);
}

int ttb_accumulate_M_size(ttb_handle h, int64_t b, int64_t* n_doubles) {
  TTB_TRY(check_handle(h));
  if (b < 0 || b >= h->batch || !n_doubles) return fail(TTB_EINVAL, "bad argument");
  const int L = (int)h->L;
  const int* bd = &h->h_bond[b * (L + 1)];
  int64_t n = 0;
  for (int t = 0; t + 1 < L; ++t)
    for (int u = t + 1; u < L; ++u) n += (int64_t)bd[t + 1] * bd[u];
  *n_doubles = n;
  return TTB_OK;
}

int ttb_accumulate_M(ttb_handle h, int64_t b, int normalize, double* out) {
  TTB_TRY(check_handle(h));
  if (b < 0 || b >= h->batch || !out) return fail(TTB_EINVAL, "bad argument");
  const int L = (int)h->L;
  if (L < 2) return TTB_OK;
  const int* bd = &h->h_bond[b * (L + 1)];
  std::vector<int64_t> toff(L + 1, 0), moff(L, 0);
  for (int t = 0; t < L; ++t) toff[t + 1] = toff[t] + (((int64_t)h->bond0[t] * h->bond0[t + 1] + 15) & ~int64_t(15));
  for (int t = 0; t + 1 < L; ++t) {
    int64_t n = 0;
    for (int u = t + 1; u < L; ++u) n += (int64_t)bd[t + 1] * bd[u];
    moff[t + 1] = moff[t] + n;
  }
  const int64_t total = moff[L - 1];
  double *d_tr = nullptr, *d_out = nullptr;
  int64_t *d_toff = nullptr, *d_moff = nullptr;
  cudaError_t e = cudaMalloc(&d_tr, sizeof(double) * toff[L]);
  if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(double) * std::max<int64_t>(total, 1));
  if (e == cudaSuccess) e = cudaMalloc(&d_toff, sizeof(int64_t) * (L + 1));
  if (e == cudaSuccess) e = cudaMalloc(&d_moff, sizeof(int64_t) * L);
  int rc = TTB_OK;
  if (e != cudaSuccess) rc = fail(TTB_ENOMEM, "accumulate_M: %s", cudaGetErrorString(e));
  if (rc == TTB_OK) {
    cudaMemcpyAsync(d_toff, toff.data(), sizeof(int64_t) * (L + 1), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_moff, moff.data(), sizeof(int64_t) * L, cudaMemcpyHostToDevice, h->stream);
    cudaStreamSynchronize(h->stream);
    CallTimer tm(h);
    const int64_t maxsite = h->dmax * h->dmax;
    dim3 g((unsigned)std::min<int64_t>((maxsite + 255) / 256, 64), (unsigned)L, 1);
    // traces of train b only: the kernel indexes trains by blockIdx.z, so hand it this train's slab / bonds
    TTB_LAUNCH(h, "k_site_trace", k_site_trace<<<g, 256, 0, h->stream>>>(h->slab + b * h->tt_stride, h->tt_stride, h->d_off,
                                                                   h->d_bond + b * (L + 1), L + 1, (int)h->Q, d_tr, toff[L], d_toff));
    AccM c;
    c.T = d_tr; c.toff = d_toff; c.bond = h->d_bond + b * (L + 1); c.L = L; c.out = d_out; c.moff = d_moff;
    c.scratch = nullptr; c.sstride = 0; c.normalize = normalize;
    TTB_LAUNCH(h, "k_accumulate_M", k_accumulate_M<<<(unsigned)(L - 1), 256, 0, h->stream>>>(c));
    rc = tm.finish();
    if (rc == TTB_OK && total > 0) {
      e = cudaMemcpy(out, d_out, sizeof(double) * total, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) rc = fail(TTB_ECUDA, "accumulate_M copy: %s", cudaGetErrorString(e));
    }
  }
  cudaFree(d_tr); cudaFree(d_out); cudaFree(d_toff); cudaFree(d_moff);
  return rc;
}

int ttb_env_size(ttb_handle h, int64_t b, int left, int64_t* n_doubles) {
  TTB_TRY(check_handle(h));
  if (b < 0 || b >= h->batch || !n_doubles) return fail(TTB_EINVAL, "bad argument");
  const int L = (int)h->L;
  const int* bd = &h->h_bond[b * (L + 1)];
  int64_t n = 0;
  for (int t = 0; t < L; ++t) n += left ? (int64_t)bd[0] * bd[t + 1] : (int64_t)bd[t] * bd[L];
  *n_doubles = n;
  return TTB_OK;
}

int ttb_accumulate_L(ttb_handle h, int normalize, double* l_out, double* logabs, int* sign) {
  return accumulate(h, true, normalize, l_out, logabs, sign);
}

int ttb_accumulate_R(ttb_handle h, int normalize, double* r_out, double* logabs, int* sign) {
  return accumulate(h, false, normalize, r_out, logabs, sign);
}

int ttb_normalization(ttb_handle h, double* logabs, int* sign) {
  TTB_TRY(check_handle(h));
  std::vector<double> la(h->batch), lz(h->batch);
  std::vector<int> sa(h->batch), sz(h->batch);
  TTB_TRY(accumulate(h, true, 1, nullptr, la.data(), sa.data()));
  TTB_CUDA(cudaMemcpy(lz.data(), h->d_logz, sizeof(double) * h->batch, cudaMemcpyDeviceToHost));
  TTB_CUDA(cudaMemcpy(sz.data(), h->d_signz, sizeof(int) * h->batch, cudaMemcpyDeviceToHost));
  for (int64_t b = 0; b < h->batch; ++b) {
    if (logabs) logabs[b] = la[b] - lz[b];
    if (sign) sign[b] = sa[b] * sz[b];
  }
  return TTB_OK;
}

int ttb_normalize(ttb_handle h, double* logz_out) {
  TTB_TRY(check_handle(h));
  TTB_TRY(ensure_env_ws(h));
  Workspace& w = h->ws;
  EnvSide s; EnvOffsets eo;
  TTB_TRY(env_side(h, s, eo));
  CallTimer tm(h);
  int rc = launch_acc(h, true, 1, s, /*store=*/0);
  if (rc == TTB_OK) {
    const int B = (int)h->batch;
    double* factor = w.env_log + 2 * h->batch;
    double* lzo = w.env_log + 3 * h->batch;
    TTB_LAUNCH(h, "k_normalize_prep", k_normalize_prep<<<(B + 127) / 128, 128, 0, h->stream>>>(w.env_log, h->d_logz, h->d_signz, factor, lzo, h->d_status,
                                                             (int)h->L, B));
    dim3 g((unsigned)std::min<int64_t>((h->tt_stride + 1023) / 1024, 4096), B);
    TTB_LAUNCH(h, "k_scale_trains", k_scale_trains<<<g, 256, 0, h->stream>>>(h->slab, h->tt_stride, factor));
    rc = tm.finish();
  }
  env_side_free(s);
  TTB_TRY(rc);
  if (logz_out)
    TTB_CUDA(cudaMemcpy(logz_out, w.env_log + 3 * h->batch, sizeof(double) * h->batch, cudaMemcpyDeviceToHost));
  int st = 0;
  TTB_TRY(read_status(h, &st));
  if (st & 8) return fail(TTB_EDOMAIN, "log of a negative number: normalize! on a train with negative z");
  return TTB_OK;
}

// every site tensor of train b divided by divisor[b] (z untouched): the rescaling step of normalize!
// (src/abstract_tensor_train.jl:189-193) on its own, used by the MPS wrapper's normalize!
// (src/MatrixProductStates/mps.jl:154-168), whose factor |Z|^(1/2L) comes from ttb_dot(psi, psi)
int ttb_scale_sites(ttb_handle h, const double* divisor) {
  TTB_TRY(check_handle(h));
  if (!divisor) return fail(TTB_EINVAL, "null argument");
  TTB_TRY(ensure_env_ws(h));
  Workspace& w = h->ws;
  double* factor = w.env_log + 2 * h->batch;
  TTB_CUDA(cudaMemcpyAsync(factor, divisor, sizeof(double) * h->batch, cudaMemcpyHostToDevice, h->stream));
  CallTimer tm(h);
  dim3 g((unsigned)std::min<int64_t>((h->tt_stride + 1023) / 1024, 4096), (unsigned)h->batch);
  TTB_LAUNCH(h, "k_scale_trains", k_scale_trains<<<g, 256, 0, h->stream>>>(h->slab, h->tt_stride, factor));
  return tm.finish();
}

int ttb_normalize_eachmatrix(ttb_handle h) {
  TTB_TRY(check_handle(h));
  double* mx = nullptr;
  TTB_CUDA(cudaMalloc(&mx, sizeof(double) * h->batch * h->L));
  CallTimer tm(h);
  dim3 g((unsigned)h->L, (unsigned)h->batch);
  TTB_LAUNCH(h, "k_site_maxabs", k_site_maxabs<<<g, 256, 0, h->stream>>>(h->slab, h->tt_stride, h->d_off, h->d_bond, (int)h->L + 1, (int)h->Q, mx));
  TTB_LAUNCH(h, "k_site_scale", k_site_scale<<<g, 256, 0, h->stream>>>(h->slab, h->tt_stride, h->d_off, h->d_bond, (int)h->L + 1, (int)h->Q, mx));
  TTB_LAUNCH(h, "k_eachmatrix_z", k_eachmatrix_z<<<((int)h->batch + 127) / 128, 128, 0, h->stream>>>(mx, h->d_logz, (int)h->L, (int)h->batch));
  int rc = tm.finish();
  cudaFree(mx);
  return rc;
}

int ttb_marginals(ttb_handle h, double* out) { return marginals_impl(h, nullptr, nullptr, out); }

}  // extern "C"

namespace ttb {
static int marginals_impl(ttb_tt* h, const ttb_env* l, const ttb_env* r, double* out) {
  TTB_TRY(check_handle(h));
  if (!out) return fail(TTB_EINVAL, "null argument");
  TTB_TRY(ensure_env_ws(h));
  TTB_TRY(env_check(h, l, true));
  TTB_TRY(env_check(h, r, false));
  Workspace& w = h->ws;
  EnvSide s; EnvOffsets eo;
  TTB_TRY(env_side(h, s, eo));
  double* d_out = nullptr;
  const size_t n = (size_t)h->batch * h->L * h->Q;
  cudaError_t e = cudaMalloc(&d_out, sizeof(double) * n);
  if (e != cudaSuccess) { env_side_free(s); return fail(TTB_ENOMEM, "marginals: %s", cudaGetErrorString(e)); }
  CallTimer tm(h);
  int rc = provide_env(h, true, l, s.d_eoffL, s.d_eoffR);
  if (rc == TTB_OK) rc = provide_env(h, false, r, s.d_eoffL, s.d_eoffR);
  if (rc == TTB_OK) {
    MargCtx c;
    c.slab = h->slab; c.tt_stride = h->tt_stride; c.off = h->d_off; c.bond = h->d_bond;
    c.L = (int)h->L; c.L1 = (int)h->L + 1; c.Q = (int)h->Q;
    c.envL = w.envL; c.envR = w.envR; c.env_stride = w.env_stride; c.eoffL = s.d_eoffL; c.eoffR = s.d_eoffR;
    c.out = d_out;
    dim3 g((unsigned)h->L, (unsigned)h->batch);
    if (small_path(h) && h->Q <= 1024) { TTB_LAUNCH(h, "k_marginals_small", k_marginals_small<<<g, 128, sizeof(double) * 4 * h->Q, h->stream>>>(c)); }
    else
    TTB_LAUNCH(h, "k_marginals", k_marginals<<<g, 256, sizeof(double) * h->Q, h->stream>>>(c));
    rc = tm.finish();
  }
  if (rc == TTB_OK) {
    e = cudaMemcpy(out, d_out, sizeof(double) * n, cudaMemcpyDefault);
    if (e != cudaSuccess) rc = fail(TTB_ECUDA, "marginals copy: %s", cudaGetErrorString(e));
  }
  cudaFree(d_out);
  env_side_free(s);
  return rc;
}
}  // namespace ttb

extern "C" {

int ttb_twovar_count(ttb_handle h, int64_t maxdist, int64_t* npairs) {
  if (!h || !npairs) return fail(TTB_EINVAL, "null argument");
  const int64_t L = h->L;
  if (maxdist <= 0) maxdist = L - 1;
  int64_t n = 0;
  for (int64_t t = 0; t < L - 1; ++t) n += std::min(L - 1, t + maxdist) - t;
  *npairs = n;
  return TTB_OK;
}

int ttb_twovar_marginals(ttb_handle h, int64_t maxdist, double* out) {
  if (!h) return fail(TTB_EINVAL, "null argument");
  return ttb_twovar_marginals_range(h, maxdist, 0, h->L - 1, out);
}

// pairs (t, u) with t_lo <= t < t_hi only (0-based first site; pairs are stored t-major, so this is a contiguous
// block of the full result): the unit one rank computes when the pairs of ONE train are sharded over GPUs
int ttb_twovar_marginals_range(ttb_handle h, int64_t maxdist, int64_t t_lo, int64_t t_hi, double* out) {
  return twovar_impl(h, nullptr, nullptr, maxdist, t_lo, t_hi, out);
}

}  // extern "C"

namespace ttb {
static int twovar_impl(ttb_tt* h, const ttb_env* lenv, const ttb_env* renv, int64_t maxdist, int64_t t_lo, int64_t t_hi, double* out) {
  TTB_TRY(check_handle(h));
  if (!out) return fail(TTB_EINVAL, "null argument");
  TTB_TRY(env_check(h, lenv, true));
  TTB_TRY(env_check(h, renv, false));
  const int64_t L = h->L, Q = h->Q;
  if (t_lo < 0 || t_hi > L - 1 || t_lo > t_hi) return fail(TTB_EINVAL, "pair range [%lld, %lld) outside [0, %lld)", (long long)t_lo, (long long)t_hi, (long long)(L - 1));
  if (maxdist <= 0) maxdist = L - 1;
  int64_t npairs = 0;
  ttb_twovar_count(h, maxdist, &npairs);
  if (npairs == 0 || t_lo == t_hi) return TTB_OK;
  // the Q x Q table of one pair lives in shared memory; large Q takes the generic kernel (k_twovar), whose only
  // Q-dependent resource is that table
  if (Q * Q * sizeof(double) > 200 * 1024) return fail(TTB_EUNSUPPORTED, "twovar_marginals supports Q <= 160");
  TTB_TRY(ensure_env_ws(h));
  Workspace& w = h->ws;
  EnvSide s; EnvOffsets eo;
  TTB_TRY(env_side(h, s, eo));
  std::vector<int64_t> poff(L, 0);
  for (int64_t t = 1; t < L; ++t) poff[t] = poff[t - 1] + (std::min(L - 1, (t - 1) + maxdist) - (t - 1));
  int64_t* d_poff = nullptr;
  double *d_out = nullptr, *d_scr = nullptr;
  const int64_t scr_stride = 2 * ((Q * h->bond0[0] * h->dmax + 15) & ~int64_t(15));
  const size_t smem_need = sizeof(double) * (Q * Q + scr_stride);
  const bool scr_in_smem = smem_need <= 200 * 1024;
  // scratch is kept on the handle between calls (grow-only, ttb_trim releases it)
  cudaError_t e = w.scr[0].ensure(sizeof(int64_t) * L);
  if (e == cudaSuccess) e = w.scr[1].ensure(sizeof(double) * h->batch * npairs * Q * Q);
  if (e == cudaSuccess && !scr_in_smem) e = w.scr[2].ensure(sizeof(double) * h->batch * (L - 1) * scr_stride);
  d_poff = static_cast<int64_t*>(w.scr[0].p); d_out = static_cast<double*>(w.scr[1].p);
  d_scr = scr_in_smem ? nullptr : static_cast<double*>(w.scr[2].p);
  int rc = TTB_OK;
  if (e != cudaSuccess) rc = fail(TTB_ENOMEM, "twovar_marginals: %s", cudaGetErrorString(e));
  // fast path: GT_u(x), T_u once per call, then Frobenius products + one small GEMM per pair
  const int64_t kcap = h->bond0[0] * h->dmax;
  const size_t fast_smem = sizeof(double) * (3 * (size_t)Q * kcap + (size_t)h->dmax * h->dmax + 8 * (Q * Q + 1) + Q * Q);
  // open trains with bonds above the small path: a group of first sites per CTA, K T_u on the tensor cores (twovar_open.cuh)
  const bool use_open = !h->periodic && h->bond0[0] == 1 && h->dmax > 32 && h->dmax <= 256 && Q <= 8 &&
                        getenv("TTB_TWOVAR_GENERIC") == nullptr && getenv("TTB_TWOVAR_NOOPEN") == nullptr;
  const bool use_fast = use_open || (fast_smem <= 200 * 1024 && getenv("TTB_TWOVAR_GENERIC") == nullptr);
  // small bonds (config 5): the per-pair product on the FP64 tensor cores (twovar_mma.cuh); with the
  // small-path precompute its operands are stored zero padded (no per-pair dimension handling)
  const bool use_mma = use_fast && h->dmax <= 32 && Q * h->bond0[0] <= 64 && Q >= 2 && Q <= 4 && getenv("TTB_TWOVAR_NOMMA") == nullptr;
  const bool padded = use_mma && small_path(h) && getenv("TTB_TWOVAR_NOPAD") == nullptr;
  const int64_t gstride = padded ? ((h->bond0[0] * 32 + 15) & ~int64_t(15)) : ((kcap + 15) & ~int64_t(15));
  const int64_t tstride = padded ? 1024 : ((h->dmax * h->dmax + 15) & ~int64_t(15));
  double *d_GT = nullptr, *d_Tm = nullptr;
  if (rc == TTB_OK && use_fast) {   // every allocation happens BEFORE the timed region (a first call on a handle pays cudaMalloc)
    e = w.scr[3].ensure(sizeof(double) * h->batch * L * Q * gstride);
    if (e == cudaSuccess) e = w.scr[4].ensure(sizeof(double) * h->batch * L * tstride);
    d_GT = static_cast<double*>(w.scr[3].p); d_Tm = static_cast<double*>(w.scr[4].p);
    if (e != cudaSuccess) rc = fail(TTB_ENOMEM, "twovar_marginals: %s", cudaGetErrorString(e));
  }
  if (rc == TTB_OK) {
    cudaMemcpyAsync(d_poff, poff.data(), sizeof(int64_t) * L, cudaMemcpyHostToDevice, h->stream);
    cudaStreamSynchronize(h->stream);
    CallTimer tm(h);
    rc = provide_env(h, true, lenv, s.d_eoffL, s.d_eoffR);
    if (rc == TTB_OK) rc = provide_env(h, false, renv, s.d_eoffL, s.d_eoffR);
    if (rc == TTB_OK && use_fast) {
      {
        TwoPre pc;
        pc.slab = h->slab; pc.tt_stride = h->tt_stride; pc.off = h->d_off; pc.bond = h->d_bond;
        pc.L = (int)L; pc.L1 = (int)L + 1; pc.Q = (int)Q;
        pc.envR = w.envR; pc.env_stride = w.env_stride; pc.eoffR = s.d_eoffR;
        pc.GT = d_GT; pc.Tm = d_Tm; pc.gstride = gstride; pc.tstride = tstride; pc.padded = padded ? 1 : 0;
        if (small_path(h)) { TTB_LAUNCH(h, "k_twovar_pre_small", k_twovar_pre_small<<<dim3((unsigned)L, (unsigned)h->batch), 128, 0, h->stream>>>(pc)); }
        else
        TTB_LAUNCH(h, "k_twovar_pre", k_twovar_pre<<<dim3((unsigned)L, (unsigned)h->batch), 256, 0, h->stream>>>(pc));
        TwoFast fc;
        fc.slab = h->slab; fc.tt_stride = h->tt_stride; fc.off = h->d_off; fc.bond = h->d_bond;
        fc.L = (int)L; fc.L1 = (int)L + 1; fc.Q = (int)Q;
        fc.envL = w.envL; fc.env_stride = w.env_stride; fc.eoffL = s.d_eoffL;
        fc.GT = d_GT; fc.Tm = d_Tm; fc.gstride = gstride; fc.tstride = tstride;
        fc.pair_off = d_poff; fc.maxdist = (int)std::min<int64_t>(maxdist, L); fc.out = d_out; fc.npairs = npairs;
        fc.kcap = (int)kcap; fc.dmax = (int)h->dmax; fc.t0 = (int)t_lo;
        TTB_ONCE_PER_DEVICE(h, 2) cudaFuncSetAttribute(k_twovar_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
        const dim3 gm((unsigned)(t_hi - t_lo), (unsigned)h->batch);
        const size_t bsm = sizeof(int) * (L + 1);
#define TTB_TV(Q_)                                                                                                        \
  do {                                                                                                                    \
    if (padded) { TTB_LAUNCH(h, "k_twovar_mma", (k_twovar_mma<Q_, true><<<gm, 256, bsm, h->stream>>>(fc))); }              \
    else { TTB_LAUNCH(h, "k_twovar_mma", (k_twovar_mma<Q_, false><<<gm, 256, bsm, h->stream>>>(fc))); }                    \
  } while (0)
        if (use_open) {
          const int tb = kTvoRows / (int)Q;
          const size_t osm = tvo_smem((int)h->dmax, (int)Q);
          TTB_ONCE_PER_DEVICE(h, 10) cudaFuncSetAttribute(k_twovar_open, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
          const dim3 go((unsigned)((t_hi - t_lo + tb - 1) / tb), (unsigned)h->batch);
          TTB_LAUNCH(h, "k_twovar_open", (k_twovar_open<<<go, 256, osm, h->stream>>>(fc, tb, (int)t_hi)));
        } else
        if (use_mma && Q == 2) TTB_TV(2);
        else if (use_mma && Q == 3) TTB_TV(3);
        else if (use_mma && Q == 4) TTB_TV(4);
        else
        TTB_LAUNCH(h, "k_twovar_fast", k_twovar_fast<<<gm, 256, fast_smem, h->stream>>>(fc));
#undef TTB_TV
        rc = tm.finish();
      }
    } else if (rc == TTB_OK) {
      TwoCtx c;
      c.slab = h->slab; c.tt_stride = h->tt_stride; c.off = h->d_off; c.bond = h->d_bond;
      c.L = (int)L; c.L1 = (int)L + 1; c.Q = (int)Q;
      c.envL = w.envL; c.envR = w.envR; c.env_stride = w.env_stride; c.eoffL = s.d_eoffL; c.eoffR = s.d_eoffR;
      c.scratch = d_scr; c.scratch_stride = scr_stride; c.pair_off = d_poff; c.maxdist = (int)std::min<int64_t>(maxdist, L);
      c.out = d_out; c.npairs = npairs; c.t0 = (int)t_lo;
      dim3 g((unsigned)(t_hi - t_lo), (unsigned)h->batch);
      TTB_ONCE_PER_DEVICE(h, 3) cudaFuncSetAttribute(k_twovar, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
      TTB_LAUNCH(h, "k_twovar", k_twovar<<<g, 256, scr_in_smem ? smem_need : sizeof(double) * Q * Q, h->stream>>>(c));
      rc = tm.finish();
    }
    if (rc == TTB_OK) {
      // the range's pairs are the block [poff[t_lo], poff[t_hi]) of every train's result
      const int64_t p_lo = poff[t_lo], p_hi = t_hi < L - 1 ? poff[t_hi] : npairs;
      const size_t row = sizeof(double) * (size_t)(p_hi - p_lo) * Q * Q;
      e = cudaMemcpy2D(out, row, d_out + p_lo * Q * Q, sizeof(double) * npairs * Q * Q, row, (size_t)h->batch, cudaMemcpyDefault);
      if (e != cudaSuccess) rc = fail(TTB_ECUDA, "twovar copy: %s", cudaGetErrorString(e));
    }
  }
  env_side_free(s);
  return rc;
}
}  // namespace ttb
