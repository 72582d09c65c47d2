// sample.cu -- evaluate and hierarchical sampling.
//
// Reference: evaluate  src/abstract_tensor_train.jl:80-83
//            sample!   src/abstract_tensor_train.jl:355-381, inverse-CDF draw src/utils.jl:9-20
//
// Sampling uses the cheaper equivalent of the reference contraction: with
// G_t(x) = A_t(x) * r_{t+1} precomputed once per call, the conditional weights are
// p[x] = sum_{k,m} Q[k,m] G_t(x)[m,k] and only the chosen product Q <- Q * A_t(x_t) is formed.
// The running product carries its own log-scale, so un-normalised long chains do not overflow
// (a superset of the reference, which needs normalize! first).  Uniforms come either from the caller
// (bit-parity tests) or from Philox4x32-10 keyed by (seed; global sample index, site).
#include <algorithm>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace ttb {

// (Philox4x32-10: common.cuh, shared with mps.cu)

// ---- evaluate: one CTA per query point ------------------------------------------------------
struct EvalCtx {
  const double* slab; int64_t tt_stride; const int64_t* off; const int* bond; int L, L1, Q;
  const int32_t* X; double* out; const double* logz; const int* signz; int b; int dmax;
  int* status;
};

__global__ void k_evaluate(EvalCtx c) {
  const int q = blockIdx.x;
  const int* bond = c.bond + (int64_t)c.b * c.L1;
  const double* tt = c.slab + (int64_t)c.b * c.tt_stride;
  const int32_t* xq = c.X + (int64_t)q * c.L;
  extern __shared__ double sm[];
  double* v0 = sm;
  double* v1 = sm + c.dmax;
  __shared__ double red[40];
  __shared__ double s_val, s_ls;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  const int r = bond[0];
  double tot = 0.0, totls = -INFINITY;  // sum_k val_k * exp(ls_k) kept as tot * exp(totls)
  for (int k = 0; k < r; ++k) {
    double* v = v0;
    double* vn = v1;
    for (int i = tid; i < r; i += nt) v[i] = (i == k) ? 1.0 : 0.0;
    double ls = 0.0;
    __syncthreads();
    for (int t = 0; t < c.L; ++t) {
      const int dl = bond[t], dr = bond[t + 1];
      int x = xq[t];
      x = min(max(x, 0), c.Q - 1);
      const double* A = tt + c.off[t] + (int64_t)dl * dr * x;
      for (int n = wid; n < dr; n += nw) {
        double acc = 0.0;
        for (int m = lane; m < dl; m += 32) acc += v[m] * A[m + (int64_t)dl * n];
        acc = warp_sum(acc);
        if (lane == 0) vn[n] = acc;
      }
      __syncthreads();
      double mx = 0.0;
      for (int n = tid; n < dr; n += nt) mx = fmax(mx, fabs(vn[n]));
      mx = block_max(mx, red);
      if (mx > 0.0 && !isinf(mx)) {
        for (int n = tid; n < dr; n += nt) vn[n] /= mx;
        ls += log(mx);
      }
      __syncthreads();
      double* tmp = v; v = vn; vn = tmp;
    }
    if (tid == 0) {
      const double val = v[k];
      if (val != 0.0) {
        if (totls == -INFINITY) { tot = val; totls = ls; }
        else if (ls > totls) { tot = tot * exp(totls - ls) + val; totls = ls; }
        else tot += val * exp(ls - totls);
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    s_val = tot; s_ls = totls;
    double res;
    if (tot == 0.0) res = 0.0;
    else res = (tot < 0 ? -1.0 : 1.0) * c.signz[c.b] * exp(log(fabs(tot)) + totls - c.logz[c.b]);
    c.out[q] = res;
  }
  (void)s_val; (void)s_ls;
}

// ---- G_t(x) = A_t(x) * r_{t+1}   (d_t x rr), all sites: grid (L, Q) ---------------------------
struct SampCtx {
  const double* slab; int64_t tt_stride; const int64_t* off; const int* bond; int L, L1, Q;
  const double* envR; const int64_t* eoffR;
  double* G; const int64_t* goff;   // goff[t] = offset of site t's Q blocks (d_t*rr each)
  const double* logzR; const int* signzR;
  int b; int dmax; int rmax;
  const double* uniforms; uint64_t seed; uint64_t first;
  uint8_t* x_out; double* p_out; unsigned long long* hist; double* sum_logp;
  int* status;
  int64_t nsamples;
};

__global__ void k_sample_prep(SampCtx c) {
  const int t = blockIdx.x, x = blockIdx.y;
  const int* bond = c.bond + (int64_t)c.b * c.L1;
  const int dl = bond[t], dr = bond[t + 1], rr = bond[c.L];
  const double* A = c.slab + (int64_t)c.b * c.tt_stride + c.off[t] + (int64_t)dl * dr * x;
  const double* r = t < c.L - 1 ? c.envR + c.eoffR[t + 1] : nullptr;  // dr x rr
  double* G = c.G + c.goff[t] + (int64_t)dl * rr * x;
  for (int o = threadIdx.x; o < dl * rr; o += blockDim.x) {
    const int m = o % dl, k = o / dl;
    double acc = 0.0;
    if (r) for (int n = 0; n < dr; ++n) acc += A[m + (int64_t)dl * n] * r[n + (int64_t)dr * k];
    else acc = A[m + (int64_t)dl * k];
    G[o] = acc;
  }
}

// ---- sampler: one CTA per draw (generic open/periodic path) -----------------------------------
__global__ void k_sample(SampCtx c, int64_t base) {
  const int64_t s = base + blockIdx.x;
  if (s >= c.nsamples) return;
  const int* bond = c.bond + (int64_t)c.b * c.L1;
  const double* tt = c.slab + (int64_t)c.b * c.tt_stride;
  extern __shared__ double sm[];
  double* Q0 = sm;                                   // r x d state (k + r*m)
  double* Q1 = Q0 + (size_t)c.rmax * c.dmax;
  double* pw = Q1 + (size_t)c.rmax * c.dmax;         // Q weights
  __shared__ double red[40];
  __shared__ int s_x;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  const int r = bond[0], rr = bond[c.L];
  double* Qs = Q0;
  double* Qn = Q1;
  for (int i = tid; i < r * r; i += nt) Qs[i] = (i % r == i / r) ? 1.0 : 0.0;
  double ls = 0.0;
  __syncthreads();
  for (int t = 0; t < c.L; ++t) {
    const int dl = bond[t], dr = bond[t + 1];
    const double* G = c.G + c.goff[t];
    for (int x = 0; x < c.Q; ++x) {
      double acc = 0.0;
      const double* Gx = G + (int64_t)dl * rr * x;
      for (int o = tid; o < r * dl; o += nt) {
        const int k = o % r, m = o / r;
        acc += Qs[o] * Gx[m + (int64_t)dl * k];
      }
      acc = block_sum(acc, red);
      if (tid == 0) pw[x] = acc;
    }
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      bool neg = false;
      for (int x = 0; x < c.Q; ++x) { neg |= pw[x] < 0.0; tot += pw[x]; }
      int xs = c.Q - 1;
      if (neg) { atomicOr(c.status, ST_NEGATIVE); xs = 0; }
      else {
        const double u = c.uniforms ? c.uniforms[s * c.L + t] : philox_uniform(c.seed, c.first + (uint64_t)s, (uint32_t)t);
        double cw = 0.0;
        for (int x = 0; x < c.Q; ++x) {
          cw += pw[x] / tot;
          if (cw > u) { xs = x; break; }
        }
      }
      s_x = xs;
      if (c.x_out) c.x_out[s * c.L + t] = (uint8_t)xs;
      if (c.hist) atomicAdd(&c.hist[(int64_t)t * c.Q + xs], 1ull);
    }
    __syncthreads();
    const int xs = s_x;
    const double* A = tt + c.off[t] + (int64_t)dl * dr * xs;
    // Qn[k, n] = sum_m Qs[k, m] A[m, n]
    for (int o = wid; o < r * dr; o += nw) {
      const int k = o % r, n = o / r;
      double acc = 0.0;
      for (int m = lane; m < dl; m += 32) acc += Qs[k + (size_t)r * m] * A[m + (int64_t)dl * n];
      acc = warp_sum(acc);
      if (lane == 0) Qn[o] = acc;
    }
    __syncthreads();
    double mx = 0.0;
    for (int o = tid; o < r * dr; o += nt) mx = fmax(mx, fabs(Qn[o]));
    mx = block_max(mx, red);
    if (mx > 0.0 && !isinf(mx)) {
      for (int o = tid; o < r * dr; o += nt) Qn[o] /= mx;
      ls += log(mx);
    }
    __syncthreads();
    double* tmp = Qs; Qs = Qn; Qn = tmp;
  }
  if (tid == 0) {
    double tr = 0.0;
    for (int k = 0; k < min(r, rr); ++k) tr += Qs[k + (size_t)r * k];
    // p = tr(Q) / z_R  (src/abstract_tensor_train.jl:379)
    const double lp = log(fabs(tr)) + ls - c.logzR[0];
    if (c.p_out) c.p_out[s] = ((tr < 0) ? -1.0 : 1.0) * c.signzR[0] * exp(lp);
    if (c.sum_logp) atomicAdd(c.sum_logp, lp);
  }
}

// ---- packing for the tiled sampler: A_t as [site][k-chunk][x][n][8] and G_t as [site][m][8], zero padded
struct PackCtx {
  const double* slab; int64_t tt_stride; const int64_t* off; const int* bond; int L, L1, Q;
  const double* envR; const int64_t* eoffR;
  double* Ap; double* Gp; int b; int DP;
};

__global__ void k_sample_pack(PackCtx c) {
  const int t = blockIdx.x, x = blockIdx.y;
  const int* bond = c.bond + (int64_t)c.b * c.L1;
  const int dl = bond[t], dr = bond[t + 1];
  const int DP = c.DP, NKC = DP / 8;
  const double* A = c.slab + (int64_t)c.b * c.tt_stride + c.off[t] + (int64_t)dl * dr * x;
  const double* r = t < c.L - 1 ? c.envR + c.eoffR[t + 1] : nullptr;  // dr x 1 (open train)
  double* Ap = c.Ap + (int64_t)t * NKC * c.Q * DP * 8;
  for (int i = threadIdx.x; i < DP * DP; i += blockDim.x) {
    const int k8 = i & 7, n = (i >> 3) % DP, kc = (i >> 3) / DP;
    const int m = kc * 8 + k8;
    Ap[(((int64_t)kc * c.Q + x) * DP + n) * 8 + k8] = (m < dl && n < dr) ? A[m + (int64_t)dl * n] : 0.0;
  }
  double* Gp = c.Gp + (int64_t)t * DP * 8;
  for (int m = threadIdx.x; m < DP; m += blockDim.x) {
    double acc = 0.0;
    if (m < dl) {
      if (r) for (int n = 0; n < dr; ++n) acc += A[m + (int64_t)dl * n] * r[n];
      else acc = A[m];  // last site: dr == 1, r = 1
    }
    Gp[m * 8 + x] = acc;
  }
}

}  // namespace ttb

#include "tma_mma.cuh"
#include "sample_tile.cuh"

using namespace ttb;

// environment plumbing shared with env.cu
extern "C" int ttb_accumulate_R(ttb_handle h, int normalize, double* r_out, double* logabs, int* sign);

namespace ttb {
int provide_env_R(ttb_tt* h, const ttb_env* rz);   // env.cu

static int sample_impl(ttb_tt* h, int64_t b, int64_t nsamples, uint64_t seed, uint64_t first, const double* uniforms,
                       uint8_t* x_out, double* p_out, int64_t* hist, double* sum_logp, const ttb_env* rz = nullptr) {
  TTB_TRY(check_handle(h));
  if (b < 0 || b >= h->batch) return fail(TTB_EINVAL, "train index out of range");
  if (nsamples < 0) return fail(TTB_EINVAL, "nsamples must be >= 0");
  if (nsamples == 0) return TTB_OK;
  if (h->Q > 256) return fail(TTB_EUNSUPPORTED, "sample supports Q <= 256 (uint8 output)");
  const int L = (int)h->L, Q = (int)h->Q;
  const int64_t rmax = h->bond0[0], dmax = h->dmax;
  const size_t smem = sizeof(double) * (2 * (size_t)rmax * dmax + Q);
  if (smem > 200 * 1024) return fail(TTB_EUNSUPPORTED, "sample: periodic state d_1*d_max too large for shared memory");
  // r = accumulate_R(A): recomputed per call like the reference's default keyword (:336) unless the caller
  // passes rz= (an environment handle: the streaming pass over the train is paid once for many calls)
  if (rz) TTB_TRY(provide_env_R(h, rz));
  else TTB_TRY(ttb_accumulate_R(h, 1, nullptr, nullptr, nullptr));
  double ms_acc = h->last_ms; int64_t l_acc = h->last_launches;
  Workspace& w = h->ws;
  std::vector<int64_t> eoffR(L + 1, 0), goff(L + 1, 0);
  for (int t = 0; t < L; ++t) {
    eoffR[t + 1] = eoffR[t] + h->bond0[t] * h->bond0[L];
    goff[t + 1] = goff[t] + h->bond0[t] * h->bond0[L] * Q;
  }
  int64_t *d_eoff = nullptr, *d_goff = nullptr;
  double *d_G = nullptr, *d_u = nullptr, *d_p = nullptr, *d_slp = nullptr;
  uint8_t* d_x = nullptr;
  unsigned long long* d_hist = nullptr;
  cudaError_t e = cudaSuccess;
  // scratch lives on the handle (grow-only): no cudaMalloc / cudaFree per call
  int slot = 0;
  auto A = [&](void** p, size_t bytes) {
    DevBuf& d = w.samp[slot++];
    if (e == cudaSuccess) e = d.ensure(bytes);
    *p = d.p;
  };
  A((void**)&d_eoff, sizeof(int64_t) * (L + 1));
  A((void**)&d_goff, sizeof(int64_t) * (L + 1));
  A((void**)&d_G, sizeof(double) * goff[L]);
  if (uniforms) A((void**)&d_u, sizeof(double) * nsamples * L); else ++slot;
  if (x_out) A((void**)&d_x, (size_t)nsamples * L); else ++slot;
  if (p_out) A((void**)&d_p, sizeof(double) * nsamples); else ++slot;
  if (hist) A((void**)&d_hist, sizeof(unsigned long long) * L * Q); else ++slot;
  if (sum_logp) A((void**)&d_slp, sizeof(double)); else ++slot;
  int rc = TTB_OK;
  if (e != cudaSuccess) rc = fail(TTB_ENOMEM, "sample buffers: %s", cudaGetErrorString(e));
  if (rc == TTB_OK) {
    cudaMemcpyAsync(d_eoff, eoffR.data(), sizeof(int64_t) * (L + 1), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_goff, goff.data(), sizeof(int64_t) * (L + 1), cudaMemcpyHostToDevice, h->stream);
    if (uniforms) cudaMemcpyAsync(d_u, uniforms, sizeof(double) * nsamples * L, cudaMemcpyHostToDevice, h->stream);
    if (hist) cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * L * Q, h->stream);
    if (sum_logp) cudaMemsetAsync(d_slp, 0, sizeof(double), h->stream);
    cudaStreamSynchronize(h->stream);
    SampCtx c;
    c.slab = h->slab; c.tt_stride = h->tt_stride; c.off = h->d_off; c.bond = h->d_bond; c.L = L; c.L1 = L + 1; c.Q = Q;
    c.envR = w.envR + b * w.env_stride; c.eoffR = d_eoff; c.G = d_G; c.goff = d_goff;
    c.logzR = w.env_log + h->batch + b; c.signzR = w.env_sign + h->batch + b;
    c.b = (int)b; c.dmax = (int)dmax; c.rmax = (int)rmax;
    c.uniforms = d_u; c.seed = seed; c.first = first;
    c.x_out = d_x; c.p_out = d_p; c.hist = d_hist; c.sum_logp = d_slp; c.status = h->d_status; c.nsamples = nsamples;
    TTB_ONCE_PER_DEVICE(h, 4) cudaFuncSetAttribute(k_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    // Tiled tensor-core sampler for open trains whose state fits (d <= 128); the generic one-CTA-per-draw
    // kernel covers periodic trains, larger bonds and tiny requests.
    const int DP = dmax <= 32 ? 32 : (dmax <= 64 ? 64 : 128);
    const size_t tile_smem = sizeof(double) * ((size_t)(ST_SMAX + 1) * (DP + 4) + 2 * (size_t)Q * DP * 8 + (size_t)DP * 8 +
                                               ST_SMAX * 10) +
                             sizeof(int) * (ST_SMAX + ST_MAXTILES * 17 + 4 + 8 + ST_SMAX) + sizeof(unsigned) * (size_t)L * Q + 64;
    const bool tiled = !h->periodic && rmax == 1 && dmax <= 128 && Q <= 8 && nsamples >= 32 && tile_smem <= 200 * 1024 &&
                       getenv("TTB_SAMPLE_GENERIC") == nullptr;
    double *d_Ap = nullptr, *d_Gp = nullptr;
    CallTimer tm(h);
    if (tiled) {
      const size_t nAp = (size_t)L * DP * DP * Q, nGp = (size_t)L * DP * 8;
      e = w.samp[8].ensure(sizeof(double) * nAp);
      if (e == cudaSuccess) e = w.samp[9].ensure(sizeof(double) * nGp);
      d_Ap = static_cast<double*>(w.samp[8].p); d_Gp = static_cast<double*>(w.samp[9].p);
      if (e != cudaSuccess) rc = fail(TTB_ENOMEM, "sample pack buffers: %s", cudaGetErrorString(e));
      if (rc == TTB_OK) {
        cudaMemsetAsync(d_Gp, 0, sizeof(double) * nGp, h->stream);
        PackCtx pc;
        pc.slab = h->slab; pc.tt_stride = h->tt_stride; pc.off = h->d_off; pc.bond = h->d_bond; pc.L = L; pc.L1 = L + 1;
        pc.Q = Q; pc.envR = w.envR + b * w.env_stride; pc.eoffR = d_eoff; pc.Ap = d_Ap; pc.Gp = d_Gp; pc.b = (int)b; pc.DP = DP;
        TTB_LAUNCH(h, "k_sample_pack", k_sample_pack<<<dim3(L, Q), 256, 0, h->stream>>>(pc));
        TileCtx tc;
        tc.Ap = d_Ap; tc.Gp = d_Gp; tc.L = L; tc.Q = Q; tc.uniforms = d_u; tc.seed = seed; tc.first = first;
        tc.x_out = d_x; tc.p_out = d_p; tc.hist = d_hist; tc.sum_logp = d_slp; tc.status = h->d_status; tc.nsamples = nsamples;
        int nsm = 148;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
        const int64_t nb = (nsamples + ST_SMAX - 1) / ST_SMAX;
        const unsigned g = (unsigned)std::min<int64_t>(nb, nsm);
        TTB_ONCE_PER_DEVICE(h, 5) {
          cudaFuncSetAttribute(k_sample_tile<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
          cudaFuncSetAttribute(k_sample_tile<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
          cudaFuncSetAttribute(k_sample_tile<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
        }
        if (DP == 32) TTB_LAUNCH(h, "k_sample_tile", k_sample_tile<32><<<g, ST_THREADS, tile_smem, h->stream>>>(tc));
        else if (DP == 64) TTB_LAUNCH(h, "k_sample_tile", k_sample_tile<64><<<g, ST_THREADS, tile_smem, h->stream>>>(tc));
        else TTB_LAUNCH(h, "k_sample_tile", k_sample_tile<128><<<g, ST_THREADS, tile_smem, h->stream>>>(tc));
      }
    } else {
      TTB_LAUNCH(h, "k_sample_prep", k_sample_prep<<<dim3(L, Q), 128, 0, h->stream>>>(c));
      const int thr = dmax <= 32 ? 64 : 128;
      for (int64_t base = 0; base < nsamples; base += (int64_t)1 << 30) {
        const unsigned g = (unsigned)std::min<int64_t>(nsamples - base, (int64_t)1 << 30);
        TTB_LAUNCH(h, "k_sample", k_sample<<<g, thr, smem, h->stream>>>(c, base));
      }
    }
    if (rc == TTB_OK) rc = tm.finish();
    h->last_ms += ms_acc; h->last_launches += l_acc;
  }
  if (rc == TTB_OK) {
    if (x_out) e = cudaMemcpy(x_out, d_x, (size_t)nsamples * L, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && p_out) e = cudaMemcpy(p_out, d_p, sizeof(double) * nsamples, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && hist) {
      std::vector<unsigned long long> hh((size_t)L * Q);
      e = cudaMemcpy(hh.data(), d_hist, sizeof(unsigned long long) * L * Q, cudaMemcpyDeviceToHost);
      for (size_t i = 0; i < hh.size(); ++i) hist[i] += (int64_t)hh[i];
    }
    if (e == cudaSuccess && sum_logp) e = cudaMemcpy(sum_logp, d_slp, sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(TTB_ECUDA, "sample copy: %s", cudaGetErrorString(e));
  }
  if (rc == TTB_OK) {
    int st = 0;
    TTB_TRY(read_status(h, &st));
    if (st & ST_NEGATIVE) return fail(TTB_ENEGATIVE, "Cannot sample from a tensor train with negative values");
  }
  return rc;
}

}  // namespace ttb

extern "C" {

int ttb_evaluate(ttb_handle h, int64_t b, const int32_t* X, int64_t n, double* out) {
  TTB_TRY(check_handle(h));
  if (b < 0 || b >= h->batch) return fail(TTB_EINVAL, "train index out of range");
  if (!X || !out || n < 0) return fail(TTB_EINVAL, "bad argument");
  if (n == 0) return TTB_OK;
  const int L = (int)h->L;
  for (int64_t i = 0; i < n * L; ++i)
    if (X[i] < 0 || X[i] >= h->Q) return fail(TTB_EINVAL, "X[%lld]=%d outside the domain 0:%lld", (long long)i, X[i], (long long)h->Q - 1);
  int32_t* d_X = nullptr;
  double* d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_X, sizeof(int32_t) * n * L);
  if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(double) * n);
  int rc = TTB_OK;
  if (e != cudaSuccess) rc = fail(TTB_ENOMEM, "evaluate: %s", cudaGetErrorString(e));
  if (rc == TTB_OK) {
    cudaMemcpyAsync(d_X, X, sizeof(int32_t) * n * L, cudaMemcpyHostToDevice, h->stream);
    EvalCtx c;
    c.slab = h->slab; c.tt_stride = h->tt_stride; c.off = h->d_off; c.bond = h->d_bond; c.L = L; c.L1 = L + 1;
    c.Q = (int)h->Q; c.X = d_X; c.out = d_out; c.logz = h->d_logz; c.signz = h->d_signz; c.b = (int)b;
    c.dmax = (int)h->dmax; c.status = h->d_status;
    CallTimer tm(h);
    const int thr = h->dmax <= 32 ? 64 : 128;
    TTB_LAUNCH(h, "k_evaluate", k_evaluate<<<(unsigned)n, thr, sizeof(double) * 2 * h->dmax, h->stream>>>(c));
    rc = tm.finish();
    if (rc == TTB_OK) {
      e = cudaMemcpy(out, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) rc = fail(TTB_ECUDA, "evaluate copy: %s", cudaGetErrorString(e));
    }
  }
  cudaFree(d_X); cudaFree(d_out);
  return rc;
}

int ttb_sample(ttb_handle h, int64_t b, int64_t nsamples, uint64_t seed, uint64_t first_index, const double* uniforms,
               uint8_t* x_out, double* p_out) {
  if (!x_out) return fail(TTB_EINVAL, "null x_out");
  return sample_impl(h, b, nsamples, seed, first_index, uniforms, x_out, p_out, nullptr, nullptr);
}

int ttb_sample_stats(ttb_handle h, int64_t b, int64_t nsamples, uint64_t seed, uint64_t first_index, int64_t* hist,
                     double* sum_logp) {
  return sample_impl(h, b, nsamples, seed, first_index, nullptr, nullptr, nullptr, hist, sum_logp);
}

int ttb_sample_env(ttb_handle h, ttb_env_handle rz, int64_t b, int64_t nsamples, uint64_t seed, uint64_t first_index,
                   const double* uniforms, uint8_t* x_out, double* p_out) {
  if (!x_out) return fail(TTB_EINVAL, "null x_out");
  return sample_impl(h, b, nsamples, seed, first_index, uniforms, x_out, p_out, nullptr, nullptr, rz);
}

int ttb_sample_stats_env(ttb_handle h, ttb_env_handle rz, int64_t b, int64_t nsamples, uint64_t seed, uint64_t first_index,
                         int64_t* hist, double* sum_logp) {
  return sample_impl(h, b, nsamples, seed, first_index, nullptr, nullptr, nullptr, hist, sum_logp, rz);
}

}  // extern "C"
