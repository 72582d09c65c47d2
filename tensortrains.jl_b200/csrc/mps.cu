// mps.cu -- the Born-machine wrapper MPS (p(x) = |psi(x)|^2) over a train psi:
//   accumulate_L / accumulate_R   src/MatrixProductStates/mps.jl:60-103   (the "d^4" environments)
//   marginals                     :176-196
//   accumulate_M / twovar_marginals  :105-126, :206-232
//   sample!                       :264-290
//
// For an open train (d_1 = d_{L+1} = 1) the reference's four-index environments Ll[a1, al, b1, bl] have
// a1 = b1 = 1: they ARE d x d matrices, and every contraction of the file is a "sandwich"
//     up:    E' = sum_x A(x)^T E A(x)        (left environment, accumulate_M transfer)
//     down:  R' = sum_x A(x) R A(x)^T        (right environment)
// i.e. two d^3 GEMMs per physical value -- dense FP64 tensor-core work (mma.sync m16n8k8, `k_gemm_items`:
// one launch serves a whole table of products with their own shapes, so all sites / all pairs of one distance
// are one launch).  Per-x sandwiches are kept:  S_t(x) = A_t(x) r_{t+1} A_t(x)^T  gives
//     marginals        p_t[x]        = <l_{t-1}, S_t(x)>_F
//     pair marginals   b_tu[x, y]    = <K_tu(x), S_u(y)>_F,  K_t,t+1(x) = A_t(x)^T l_{t-1} A_t(x),
//                                      K_t,u+1(x) = sym(sum_y A_u(y)^T K_tu(x) A_u(y))   (streamed, never the L x L table)
//     sampling         p[x | q]      = q S_t(x) q^T,  q <- q A_t(x_t)
// The reference's quirks are kept: the in-place `Ll ./= nl` also rescales the environment stored one step
// earlier (stored l[t < L] have max-abs 1, the last one is unscaled), hermiticity is restored after every
// step ((E + E^T) / 2), z = prod(nl) * trace.
// A periodic psi has a matrix boundary: the environments are genuine four-index arrays Ll[a1, a, b1, b].  They are
// kept as S = d_1^2 SLICES E^(a1,b1) (each d x d): every slice obeys the same sandwich, the id4 start is the unit
// matrix e_a1 e_b1^T, the hermiticity restoration pairs slice (a1,b1) with the transpose of slice (b1,a1), the
// max-abs scale is common to all slices, trace = sum_{a,b} E^(a,b)[a,b]; marginals and sampling sum their
// Frobenius products / quadratic forms over the slices.  twovar_marginals(::MPS) is defined for TensorTrain only
// (mps.jl:206), like the reference.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tma_mma.cuh"
#include "gemm_items.cuh"

namespace ttb {

// out = (raw + rawT^T) / 2 (d x d), max-abs of the result into mx[slot] (ordered bits, NaN-propagating).  rawT is the
// raw matrix of the transposed boundary slice (b1, a1) -- raw itself for an open train (mps.jl:77-78)
struct SymItem { const double* raw; const double* rawT; double* out; int d; int slot; };
__global__ void __launch_bounds__(256) k_mps_sym(const SymItem* __restrict__ items, unsigned long long* mx) {
  const SymItem it = items[blockIdx.y];
  const int64_t n = (int64_t)it.d * it.d;
  double m = 0.0;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(o % it.d), j = (int)(o / it.d);
    const double v = 0.5 * (it.raw[o] + it.rawT[j + (int64_t)it.d * i]);
    it.out[o] = v;
    m = absmax_nan(m, v);
  }
  unsigned long long bits = (unsigned long long)__double_as_longlong(m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long ot = __shfl_xor_sync(0xffffffffu, bits, o); bits = ot > bits ? ot : bits; }
  if ((threadIdx.x & 31) == 0 && bits) atomicMax(mx + it.slot, bits);
}

// id4 (mps.jl:56) as slices: slice s = a1 + d1 b1 is the d1 x d1 unit matrix e_a1 e_b1^T
__global__ void k_mps_id4(double* out, int d1) {
  const int64_t n = (int64_t)d1 * d1 * d1 * d1;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = o / ((int64_t)d1 * d1), e = o - s * d1 * d1;
    out[o] = (e % d1 == s % d1 && e / d1 == s / d1) ? 1.0 : 0.0;
  }
}

// buf /= mx[slot] unless the maximum is zero (`!iszero(nl)`, mps.jl:67)
struct ScaleItem { double* buf; int64_t n; int slot; int pad_; };
__global__ void __launch_bounds__(256) k_mps_scale(const ScaleItem* __restrict__ items, const unsigned long long* mx) {
  const ScaleItem it = items[blockIdx.y];
  const double m = __longlong_as_double((long long)mx[it.slot]);
  if (m == 0.0) return;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < it.n; o += (int64_t)gridDim.x * blockDim.x) it.buf[o] /= m;
}

// out[item] = sum_{s < ns} <X + s sx, Y + s sy>_F, n contiguous doubles per slice
struct FrobItem { const double* X; const double* Y; double* out; int64_t n; int ns, pad_; int64_t sx, sy; };
__global__ void __launch_bounds__(256) k_mps_frob(const FrobItem* __restrict__ items) {
  const FrobItem it = items[blockIdx.x];
  __shared__ double red[40];
  double v = 0.0;
  for (int sl = 0; sl < it.ns; ++sl) {
    const double* X = it.X + sl * it.sx;
    const double* Y = it.Y + sl * it.sy;
    for (int64_t o = threadIdx.x; o < it.n; o += blockDim.x) v += X[o] * Y[o];
  }
  v = block_sum(v, red);
  if (threadIdx.x == 0) *it.out = v;
}

// z = prod_{t < nscaled} mx[t] * trace(last) -> (log|z|, sign);  trace = sum_{a,b} E^(a,b)[a,b] (mps.jl:128-130)
__global__ void k_mps_final(const unsigned long long* mx, int nscaled, const double* last, int d1, double* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double lg = 0.0;
  for (int t = 0; t < nscaled; ++t) {
    const double m = __longlong_as_double((long long)mx[t]);
    if (m != 0.0) lg += log(m);
  }
  double v = 0.0;
  for (int a = 0; a < d1; ++a)
    for (int b = 0; b < d1; ++b) v += last[((int64_t)(a + d1 * b)) * d1 * d1 + a + d1 * b];
  out[0] = lg + log(fabs(v));
  out[1] = v < 0.0 ? -1.0 : 1.0;
  if (v != v) out[0] = v;
}

// ---- sampler: one CTA per draw.  Q (d1 x d_t, d1 = 1 for an open psi) is the running product of the chosen
//      matrices; p[x] = sum_{a1,b1} Q[a1,:] S_t^(a1,b1)(x) Q[b1,:]^T  (mps.jl:276-279 with S = A r A^T) ----
struct MpsSampCtx {
  const double* slab; int64_t tt_off; const int64_t* off; const int64_t* bond; int L, Q, dmax, d1;
  const double* S; const int64_t* soff;     // S_t^(s)(x): d_t x d_t at S + soff[t] + (x d1^2 + s) d_t^2
  const double* zlog;                       // log z of accumulate_R
  const double* uniforms; uint64_t seed, first;
  uint8_t* x_out; double* p_out; int64_t nsamples;
};
__global__ void __launch_bounds__(128) k_mps_sample(MpsSampCtx c) {
  const int64_t s = blockIdx.x;
  if (s >= c.nsamples) return;
  extern __shared__ double sm[];
  const int d1 = c.d1, S1 = d1 * d1;
  double* q = sm;                              // [d1][dmax] row a1 at q + a1 dmax
  double* qn = sm + (size_t)d1 * c.dmax;
  double* pw = qn + (size_t)d1 * c.dmax;
  __shared__ double red[40];
  __shared__ int s_x;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  const double* tt = c.slab + c.tt_off;
  for (int i = tid; i < d1 * c.dmax; i += nt) q[i] = (i % c.dmax == i / c.dmax) ? 1.0 : 0.0;   // Q = I(d1)
  double ls = 0.0;
  __syncthreads();
  for (int t = 0; t < c.L; ++t) {
    const int dl = (int)c.bond[t], dr = (int)c.bond[t + 1];
    for (int x = 0; x < c.Q; ++x) {
      const double* Sx = c.S + c.soff[t] + (int64_t)x * S1 * dl * dl;
      double acc = 0.0;
      // one warp per (slice, column j): sum_i Q[a1,i] S[i,j], times Q[b1,j]
      for (int o = wid; o < S1 * dl; o += nw) {
        const int sl = o / dl, j = o - sl * dl, a1 = sl % d1, b1 = sl / d1;
        const double* col = Sx + (int64_t)sl * dl * dl + (int64_t)dl * j;
        const double* qa = q + (size_t)a1 * c.dmax;
        double v = 0.0;
        for (int i = lane; i < dl; i += 32) v += qa[i] * col[i];
        v = warp_sum(v);
        if (lane == 0) acc += v * q[(size_t)b1 * c.dmax + j];
      }
      acc = block_sum(acc, red);
      if (tid == 0) pw[x] = acc;
    }
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int x = 0; x < c.Q; ++x) tot += pw[x];
      const double u = c.uniforms ? c.uniforms[s * c.L + t] : philox_uniform(c.seed, c.first + (uint64_t)s, (uint32_t)t);
      int xs = c.Q - 1;
      double cw = 0.0;
      for (int x = 0; x < c.Q; ++x) {
        cw += pw[x] / tot;
        if (cw > u) { xs = x; break; }
      }
      s_x = xs;
      c.x_out[s * c.L + t] = (uint8_t)xs;
    }
    __syncthreads();
    const int xs = s_x;
    const double* A = tt + c.off[t] + (int64_t)dl * dr * xs;
    for (int o = wid; o < d1 * dr; o += nw) {          // Qn[a1, n] = sum_m Q[a1, m] A[m, n]
      const int a1 = o / dr, n = o - a1 * dr;
      const double* qa = q + (size_t)a1 * c.dmax;
      double a = 0.0;
      for (int m = lane; m < dl; m += 32) a += qa[m] * A[m + (int64_t)dl * n];
      a = warp_sum(a);
      if (lane == 0) qn[(size_t)a1 * c.dmax + n] = a;
    }
    __syncthreads();
    double mx = 0.0;
    for (int o = tid; o < d1 * dr; o += nt) mx = fmax(mx, fabs(qn[(size_t)(o / dr) * c.dmax + (o % dr)]));
    mx = block_max(mx, red);
    if (mx > 0.0 && !isinf(mx)) {
      for (int o = tid; o < d1 * dr; o += nt) qn[(size_t)(o / dr) * c.dmax + (o % dr)] /= mx;
      ls += log(mx);
    }
    __syncthreads();
    double* tmp = q; q = qn; qn = tmp;
  }
  if (tid == 0 && c.p_out) {
    double tr = 0.0;
    for (int a = 0; a < d1; ++a) tr += q[(size_t)a * c.dmax + a];
    c.p_out[s] = exp(2.0 * (log(fabs(tr)) + ls) - c.zlog[0]);   // |tr Q|^2 / z (mps.jl:288)
  }
}

// ------------------------------------------------------------------------------------------------
struct MpsPlan {
  int L = 0, Q = 0, dmax = 1;
  int d1 = 1, S = 1;                   // boundary bond and number of boundary slices d1^2 (1 for an open psi)
  std::vector<int> bd;                 // current bonds of train b, L + 1
  std::vector<int64_t> eoffL, eoffR;   // slot t of l (S d_{t+1}^2) / of r (S d_t^2)
  std::vector<int64_t> soff, uoff;     // S_t(.) (Q S d_t^2) / U_t(.) (Q S d_t d_{t+1})
  const double* tt = nullptr;
};

static int mps_plan(ttb_tt* h, int64_t b, MpsPlan& p) {
  TTB_TRY(check_handle(h));
  if (b < 0 || b >= h->batch) return fail(TTB_EINVAL, "train index out of range");
  TTB_TRY(sync_bonds_to_host(h));
  p.L = (int)h->L; p.Q = (int)h->Q;
  p.bd.assign(h->h_bond.begin() + b * (p.L + 1), h->h_bond.begin() + (b + 1) * (p.L + 1));
  if (p.bd[0] != p.bd[p.L]) return fail(TTB_ESHAPE, "MPS: first and last bond of psi differ");
  p.d1 = p.bd[0]; p.S = p.d1 * p.d1;
  p.eoffL.assign(p.L + 1, 0); p.eoffR.assign(p.L + 1, 0); p.soff.assign(p.L + 1, 0); p.uoff.assign(p.L + 1, 0);
  for (int t = 0; t < p.L; ++t) {
    p.dmax = std::max(p.dmax, std::max(p.bd[t], p.bd[t + 1]));
    p.eoffL[t + 1] = p.eoffL[t] + (int64_t)p.S * p.bd[t + 1] * p.bd[t + 1];
    p.eoffR[t + 1] = p.eoffR[t] + (int64_t)p.S * p.bd[t] * p.bd[t];
    p.soff[t + 1] = p.soff[t] + (int64_t)p.Q * p.S * p.bd[t] * p.bd[t];
    p.uoff[t + 1] = p.uoff[t] + (int64_t)p.Q * p.S * p.bd[t] * p.bd[t + 1];
  }
  const int64_t cap = int64_t(1) << 31;   // doubles: 16 GiB per buffer
  if (std::max(std::max(p.eoffL[p.L], p.eoffR[p.L]), std::max(p.soff[p.L], p.uoff[p.L])) > cap)
    return fail(TTB_EUNSUPPORTED, "MPS environments: d_1^2 d^2 L Q = %lld doubles exceed the 16 GiB the entry allows per buffer",
                (long long)p.soff[p.L]);
  p.tt = h->slab + b * h->tt_stride;
  return TTB_OK;
}

// upload a host table into a handle-owned buffer (grow-only); returns the device pointer
template <typename Tp>
static Tp* upload_table(ttb_tt* h, DevBuf& buf, const std::vector<Tp>& v, cudaError_t& e) {
  if (e == cudaSuccess) e = buf.ensure(std::max<size_t>(sizeof(Tp) * v.size(), 64));
  if (e == cudaSuccess && !v.empty()) e = cudaMemcpyAsync(buf.p, v.data(), sizeof(Tp) * v.size(), cudaMemcpyHostToDevice, h->stream);
  return static_cast<Tp*>(buf.p);
}

// id4 slices (S x d1 x d1) live behind the max-abs slots in mps[4]
static double* mps_id4_ptr(ttb_tt* h, int L) { return reinterpret_cast<double*>(static_cast<unsigned long long*>(h->ws.mps[4].p) + L + 8); }
static int mps_ensure_id4(ttb_tt* h, const MpsPlan& p) {
  Workspace& w = h->ws;
  const size_t bytes = sizeof(unsigned long long) * (p.L + 8) + sizeof(double) * (size_t)p.S * p.S + 64;
  const bool fresh = bytes > w.mps[4].cap;
  if (w.mps[4].ensure(bytes) != cudaSuccess) return fail(TTB_ENOMEM, "MPS boundary identity");
  (void)fresh;
  k_mps_id4<<<(unsigned)std::min<int64_t>(((int64_t)p.S * p.S + 255) / 256, 1024), 256, 0, h->stream>>>(mps_id4_ptr(h, p.L), p.d1);
  return TTB_OK;
}

// environments of train b into ws.mps[left ? 0 : 1]; (log|z|, sign) into *d_zout (device, 2 doubles) when non-null.
// Buffers: mps[0] l, mps[1] r, mps[2] raw / temporaries, mps[3] item tables, mps[4] max-abs slots + id4 slices.
static int mps_env(ttb_tt* h, const MpsPlan& p, bool left, int normalize, double* d_zout) {
  Workspace& w = h->ws;
  const int L = p.L, Q = p.Q, S = p.S;
  const int64_t d2 = (int64_t)p.dmax * p.dmax;
  cudaError_t e = w.mps[left ? 0 : 1].ensure(sizeof(double) * std::max<int64_t>(left ? p.eoffL[L] : p.eoffR[L], 1));
  if (e == cudaSuccess) e = w.mps[2].ensure(sizeof(double) * (size_t)(Q + 1) * S * d2);
  if (e != cudaSuccess) return fail(TTB_ENOMEM, "MPS environments: %s", cudaGetErrorString(e));
  TTB_TRY(mps_ensure_id4(h, p));
  double* env = static_cast<double*>(w.mps[left ? 0 : 1].p);
  double* Tx = static_cast<double*>(w.mps[2].p);          // [Q][S] temporaries, then the S raw sums
  double* raw = Tx + (size_t)Q * S * d2;
  unsigned long long* mx = static_cast<unsigned long long*>(w.mps[4].p);
  const double* id4 = mps_id4_ptr(h, L);
  // tables: per site Q S products E A(x) (or A(x) R), S summed products, S sym, S scale items
  std::vector<GemmItem> g1((size_t)L * Q * S), g2((size_t)L * S);
  std::vector<SymItem> sy((size_t)L * S);
  std::vector<ScaleItem> sc((size_t)L * S);
  for (int step = 0; step < L; ++step) {
    const int t = left ? step : L - 1 - step;
    const int dl = p.bd[t], dr = p.bd[t + 1];
    const double* At = p.tt + h->off[t];
    const int din = left ? dl : dr, dn = left ? dr : dl;
    // incoming environment: the previous slot (already scaled), or the id4 slices (din == d1 there)
    const double* Ein = step == 0 ? id4 : (left ? env + p.eoffL[t - 1] : env + p.eoffR[t + 1]);
    double* slot = env + (left ? p.eoffL[t] : p.eoffR[t]);
    for (int sl = 0; sl < S; ++sl) {
      const double* Es = Ein + (int64_t)sl * din * din;
      for (int x = 0; x < Q; ++x) {
        GemmItem& it = g1[((size_t)step * Q + x) * S + sl];
        it = GemmItem{};
        if (left) {   // T_x = E (dl x dl) A(x) (dl x dr)
          it.A = Es; it.lda = dl; it.B = At + (int64_t)x * dl * dr; it.ldb = dl; it.m = dl; it.n = dr; it.k = dl;
        } else {      // T_x = A(x) (dl x dr) R (dr x dr)
          it.A = At + (int64_t)x * dl * dr; it.lda = dl; it.B = Es; it.ldb = dr; it.m = dl; it.n = dr; it.k = dr;
        }
        it.C = Tx + ((size_t)x * S + sl) * d2; it.ldc = dl; it.nsum = 1;
      }
      GemmItem& s2 = g2[(size_t)step * S + sl];
      s2 = GemmItem{};
      s2.nsum = Q;
      if (left) {     // E' = sum_x A(x)^T T_x : A stored k x m = dl x dr (TA)
        s2.A = At; s2.lda = dl; s2.sA = (int64_t)dl * dr; s2.B = Tx + (size_t)sl * d2; s2.ldb = dl; s2.sB = (int64_t)S * d2;
        s2.m = dr; s2.n = dr; s2.k = dl; s2.ldc = dr;
      } else {        // R' = sum_x T_x A(x)^T : B stored n x k = dl x dr (TB)
        s2.A = Tx + (size_t)sl * d2; s2.lda = dl; s2.sA = (int64_t)S * d2; s2.B = At; s2.ldb = dl; s2.sB = (int64_t)dl * dr;
        s2.m = dl; s2.n = dl; s2.k = dr; s2.ldc = dl;
      }
      s2.C = raw + (size_t)sl * d2;
      const int slT = (sl / p.d1) + p.d1 * (sl % p.d1);     // (a1, b1) -> (b1, a1)
      sy[(size_t)step * S + sl] = SymItem{raw + (size_t)sl * d2, raw + (size_t)slT * d2, slot + (int64_t)sl * dn * dn, dn, step};
      sc[(size_t)step * S + sl] = ScaleItem{slot + (int64_t)sl * dn * dn, (int64_t)dn * dn, step, 0};
    }
  }
  std::vector<unsigned char> blob(sizeof(GemmItem) * (g1.size() + g2.size()) + sizeof(SymItem) * sy.size() + sizeof(ScaleItem) * sc.size());
  unsigned char* bp = blob.data();
  const size_t o1 = 0, o2 = o1 + sizeof(GemmItem) * g1.size(), o3 = o2 + sizeof(GemmItem) * g2.size(), o4 = o3 + sizeof(SymItem) * sy.size();
  memcpy(bp + o1, g1.data(), sizeof(GemmItem) * g1.size());
  memcpy(bp + o2, g2.data(), sizeof(GemmItem) * g2.size());
  memcpy(bp + o3, sy.data(), sizeof(SymItem) * sy.size());
  memcpy(bp + o4, sc.data(), sizeof(ScaleItem) * sc.size());
  unsigned char* d_blob = upload_table(h, w.mps[3], blob, e);
  if (e == cudaSuccess) e = cudaMemsetAsync(mx, 0, sizeof(unsigned long long) * (L + 4), h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);   // blob is local storage
  if (e != cudaSuccess) return fail(TTB_ECUDA, "MPS environments: %s", cudaGetErrorString(e));
  const GemmItem* d1i = reinterpret_cast<const GemmItem*>(d_blob + o1);
  const GemmItem* d2i = reinterpret_cast<const GemmItem*>(d_blob + o2);
  const SymItem* d3 = reinterpret_cast<const SymItem*>(d_blob + o3);
  const ScaleItem* d4 = reinterpret_cast<const ScaleItem*>(d_blob + o4);
  for (int step = 0; step < L; ++step) {
    const int t = left ? step : L - 1 - step;
    const int dl = p.bd[t], dr = p.bd[t + 1], dn = left ? dr : dl;
    launch_gemm<false, false>(h, d1i + (size_t)step * Q * S, Q * S, tiles64(dl, dr));
    if (left) launch_gemm<true, false>(h, d2i + (size_t)step * S, S, tiles64(dn, dn));
    else launch_gemm<false, true>(h, d2i + (size_t)step * S, S, tiles64(dn, dn));
    const unsigned gs = (unsigned)std::min<int64_t>(((int64_t)dn * dn + 255) / 256, 64);
    TTB_LAUNCH(h, "k_mps_sym", (k_mps_sym<<<dim3(gs, (unsigned)S), 256, 0, h->stream>>>(d3 + (size_t)step * S, mx)));
    // `Ll ./= nl` at the START of the next iteration, in place on the stored array (mps.jl:65-69): every stored
    // environment but the last has max-abs 1 (one common scale for all boundary slices)
    if (normalize && step < L - 1)
      TTB_LAUNCH(h, "k_mps_scale", (k_mps_scale<<<dim3(gs, (unsigned)S), 256, 0, h->stream>>>(d4 + (size_t)step * S, mx)));
  }
  if (d_zout) {
    const double* last = env + (left ? p.eoffL[L - 1] : p.eoffR[0]);   // S slices of d1 x d1
    TTB_LAUNCH(h, "k_mps_final", (k_mps_final<<<1, 32, 0, h->stream>>>(mx, normalize ? L - 1 : 0, last, p.d1, d_zout)));
  }
  return TTB_OK;
}

// S_t^(s)(x) = A_t(x) r_{t+1}^(s) A_t(x)^T for all (t, x, s) into mps[5] (U in mps[6]); r must be in mps[1]
static int mps_sandwich_R(ttb_tt* h, const MpsPlan& p) {
  Workspace& w = h->ws;
  const int L = p.L, Q = p.Q, S = p.S;
  cudaError_t e = w.mps[5].ensure(sizeof(double) * std::max<int64_t>(p.soff[L], 1));
  if (e == cudaSuccess) e = w.mps[6].ensure(sizeof(double) * std::max<int64_t>(p.uoff[L], 1));
  if (e != cudaSuccess) return fail(TTB_ENOMEM, "MPS sandwiches: %s", cudaGetErrorString(e));
  double* Sb = static_cast<double*>(w.mps[5].p);
  double* U = static_cast<double*>(w.mps[6].p);
  const double* r = static_cast<const double*>(w.mps[1].p);
  const double* id4 = mps_id4_ptr(h, L);
  const size_t nit = (size_t)L * Q * S;
  std::vector<GemmItem> g(2 * nit);
  unsigned tu = 1, ts = 1;
  size_t k = 0;
  for (int t = 0; t < L; ++t) {
    const int dl = p.bd[t], dr = p.bd[t + 1];
    for (int x = 0; x < Q; ++x) {
      const double* Ax = p.tt + h->off[t] + (int64_t)x * dl * dr;
      for (int sl = 0; sl < S; ++sl, ++k) {
        const double* rs = (t < L - 1 ? r + p.eoffR[t + 1] : id4) + (int64_t)sl * dr * dr;
        GemmItem& a = g[k];
        a = GemmItem{};
        a.A = Ax; a.lda = dl; a.B = rs; a.ldb = dr; a.m = dl; a.n = dr; a.k = dr;
        a.C = U + p.uoff[t] + ((int64_t)x * S + sl) * dl * dr; a.ldc = dl; a.nsum = 1;
        GemmItem& b = g[nit + k];
        b = GemmItem{};
        b.A = a.C; b.lda = dl; b.B = Ax; b.ldb = dl; b.m = dl; b.n = dl; b.k = dr;
        b.C = Sb + p.soff[t] + ((int64_t)x * S + sl) * dl * dl; b.ldc = dl; b.nsum = 1;
      }
      tu = std::max(tu, tiles64(dl, dr)); ts = std::max(ts, tiles64(dl, dl));
    }
  }
  GemmItem* dg = upload_table(h, w.mps[7], g, e);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(TTB_ECUDA, "MPS sandwiches: %s", cudaGetErrorString(e));
  for (size_t o = 0; o < nit; o += 32768) launch_gemm<false, false>(h, dg + o, (int)std::min<size_t>(32768, nit - o), tu);
  for (size_t o = 0; o < nit; o += 32768) launch_gemm<false, true>(h, dg + nit + o, (int)std::min<size_t>(32768, nit - o), ts);
  return TTB_OK;
}

// the keyword arguments l= / r= / rz= of marginals, twovar_marginals and sample! for an MPS (mps.jl:176-177, 206-208,
// 264-265): a caller-supplied environment (the array ttb_mps_accumulate returned: host or device memory) is copied into
// the workspace slot instead of being recomputed; NULL = the reference's default keyword (compute it here)
static int mps_provide_env(ttb_tt* h, const MpsPlan& p, bool left, const double* given, double* d_zout) {
  if (!given) return mps_env(h, p, left, 1, d_zout);
  Workspace& w = h->ws;
  const int64_t n = left ? p.eoffL[p.L] : p.eoffR[p.L];
  if (w.mps[left ? 0 : 1].ensure(sizeof(double) * std::max<int64_t>(n, 1)) != cudaSuccess) return fail(TTB_ENOMEM, "MPS environments");
  TTB_TRY(mps_ensure_id4(h, p));
  TTB_CUDA(cudaMemcpyAsync(w.mps[left ? 0 : 1].p, given, sizeof(double) * n, cudaMemcpyDefault, h->stream));
  return TTB_OK;
}

}  // namespace ttb

using namespace ttb;

extern "C" {

int ttb_mps_env_size(ttb_handle h, int64_t b, int left, int64_t* n_doubles) {
  if (!n_doubles) return fail(TTB_EINVAL, "null argument");
  MpsPlan p;
  TTB_TRY(mps_plan(h, b, p));
  *n_doubles = left ? p.eoffL[p.L] : p.eoffR[p.L];
  return TTB_OK;
}

int ttb_mps_accumulate(ttb_handle h, int64_t b, int left, int normalize, double* env_out, double* logabs, int* sign) {
  MpsPlan p;
  TTB_TRY(mps_plan(h, b, p));
  Workspace& w = h->ws;
  if (w.mps[8].ensure(256) != cudaSuccess) return fail(TTB_ENOMEM, "MPS outputs");
  double* d_z = static_cast<double*>(w.mps[8].p);
  CallTimer tm(h);
  int rc = mps_env(h, p, left != 0, normalize, d_z);
  if (rc == TTB_OK) rc = tm.finish();
  if (rc != TTB_OK) return rc;
  double hz[2];
  TTB_CUDA(cudaMemcpy(hz, d_z, sizeof(hz), cudaMemcpyDeviceToHost));
  if (logabs) *logabs = hz[0];
  if (sign) *sign = hz[1] < 0.0 ? -1 : 1;
  if (env_out) TTB_CUDA(cudaMemcpy(env_out, w.mps[left ? 0 : 1].p, sizeof(double) * (left ? p.eoffL[p.L] : p.eoffR[p.L]), cudaMemcpyDefault));
  return TTB_OK;
}

int ttb_mps_marginals(ttb_handle h, int64_t b, double* out) { return ttb_mps_marginals_env(h, b, nullptr, nullptr, out); }

int ttb_mps_marginals_env(ttb_handle h, int64_t b, const double* l_env, const double* r_env, double* out) {
  if (!out) return fail(TTB_EINVAL, "null argument");
  MpsPlan p;
  TTB_TRY(mps_plan(h, b, p));
  Workspace& w = h->ws;
  const int L = p.L, Q = p.Q;
  if (w.mps[8].ensure(sizeof(double) * ((size_t)L * Q + 32)) != cudaSuccess) return fail(TTB_ENOMEM, "MPS outputs");
  double* d_p = static_cast<double*>(w.mps[8].p) + 32;
  CallTimer tm(h);
  int rc = mps_provide_env(h, p, true, l_env, nullptr);
  if (rc == TTB_OK) rc = mps_provide_env(h, p, false, r_env, nullptr);
  if (rc == TTB_OK) rc = mps_sandwich_R(h, p);
  if (rc != TTB_OK) return rc;
  const double* l = static_cast<const double*>(w.mps[0].p);
  const double* Sb = static_cast<const double*>(w.mps[5].p);
  const double* id4 = mps_id4_ptr(h, L);
  std::vector<FrobItem> f((size_t)L * Q);
  for (int t = 0; t < L; ++t)
    for (int x = 0; x < Q; ++x) {     // p_t[x] = sum over boundary slices of <l_{t-1}^(s), S_t^(s)(x)>_F
      const int64_t n = (int64_t)p.bd[t] * p.bd[t];
      f[(size_t)t * Q + x] = FrobItem{t > 0 ? l + p.eoffL[t - 1] : id4, Sb + p.soff[t] + (int64_t)x * p.S * n, d_p + (size_t)t * Q + x,
                                      n, p.S, 0, n, n};
    }
  cudaError_t e = cudaSuccess;
  FrobItem* df = upload_table(h, w.mps[9], f, e);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(TTB_ECUDA, "MPS marginals: %s", cudaGetErrorString(e));
  TTB_LAUNCH(h, "k_mps_frob", (k_mps_frob<<<(unsigned)(L * Q), 256, 0, h->stream>>>(df)));
  rc = tm.finish();
  if (rc != TTB_OK) return rc;
  std::vector<double> hp((size_t)L * Q);
  TTB_CUDA(cudaMemcpy(hp.data(), d_p, sizeof(double) * hp.size(), cudaMemcpyDeviceToHost));
  for (int t = 0; t < L; ++t) {     // p ./= sum(p)  (mps.jl:190)
    double s = 0.0;
    for (int x = 0; x < Q; ++x) s += hp[(size_t)t * Q + x];
    for (int x = 0; x < Q; ++x) out[(size_t)t * Q + x] = hp[(size_t)t * Q + x] / s;
  }
  return TTB_OK;
}

int ttb_mps_twovar_marginals(ttb_handle h, int64_t b, int64_t maxdist, double* out) {
  return ttb_mps_twovar_marginals_env(h, b, nullptr, nullptr, maxdist, out);
}

int ttb_mps_twovar_marginals_env(ttb_handle h, int64_t b, const double* l_env, const double* r_env, int64_t maxdist, double* out) {
  if (!out) return fail(TTB_EINVAL, "null argument");
  MpsPlan p;
  TTB_TRY(mps_plan(h, b, p));
  if (h->periodic || p.S != 1) return fail(TTB_EINVAL, "twovar_marginals(::MPS) is defined for a TensorTrain psi only (mps.jl:206)");
  Workspace& w = h->ws;
  const int L = p.L, Q = p.Q;
  if (maxdist <= 0 || maxdist > L - 1) maxdist = L - 1;
  std::vector<int64_t> poff(L + 1, 0);
  for (int t = 0; t < L - 1; ++t) poff[t + 1] = poff[t] + (std::min<int64_t>(L - 1, t + maxdist) - t);
  const int64_t npairs = poff[L - 1];
  if (npairs == 0) return TTB_OK;
  const int64_t d2 = (int64_t)p.dmax * p.dmax;
  cudaError_t e = w.mps[8].ensure(sizeof(double) * ((size_t)npairs * Q * Q + 32));
  // K (current, raw) for every (t, x), T for every (t, x, y)
  if (e == cudaSuccess) e = w.mps[2].ensure(sizeof(double) * (size_t)L * Q * d2 * (2 + Q));
  if (e != cudaSuccess) return fail(TTB_ENOMEM, "MPS pair marginals: %s", cudaGetErrorString(e));
  double* d_out = static_cast<double*>(w.mps[8].p) + 32;
  CallTimer tm(h);
  int rc = mps_provide_env(h, p, true, l_env, nullptr);
  if (rc == TTB_OK) rc = mps_provide_env(h, p, false, r_env, nullptr);
  if (rc == TTB_OK) rc = mps_sandwich_R(h, p);
  if (rc != TTB_OK) return rc;
  // (mps_env used mps[2] as its temporary: it is free again from here)
  double* K = static_cast<double*>(w.mps[2].p);
  double* Kraw = K + (size_t)L * Q * d2;
  double* Tm = Kraw + (size_t)L * Q * d2;
  const double* l = static_cast<const double*>(w.mps[0].p);
  const double* S = static_cast<const double*>(w.mps[5].p);
  unsigned long long* mx = static_cast<unsigned long long*>(w.mps[4].p);   // slots 0 .. L-1 reused per distance
  const double* one = mps_id4_ptr(h, L);   // S = 1: the 1 x 1 identity
  auto Kp = [&](double* base, int t, int x) { return base + ((size_t)t * Q + x) * d2; };
  auto Tp = [&](int t, int x, int y) { return Tm + (((size_t)t * Q + x) * Q + y) * d2; };
  // ---- tables for all distances, one upload ----
  std::vector<GemmItem> gi;
  std::vector<SymItem> si;
  std::vector<ScaleItem> ci;
  std::vector<FrobItem> fi;
  struct Step { size_t g1, n1, g2, n2, s, ns, f, nf; unsigned t1, t2; };
  std::vector<Step> steps;
  // distance 0 -> 1: K_{t,t+1}(x) = A_t(x)^T l_{t-1} A_t(x)
  {
    Step st{}; st.g1 = gi.size(); st.t1 = 1; st.t2 = 1;
    for (int t = 0; t < L - 1; ++t)
      for (int x = 0; x < Q; ++x) {
        const int dl = p.bd[t], dr = p.bd[t + 1];
        GemmItem a{};
        a.A = t > 0 ? l + p.eoffL[t - 1] : one; a.lda = dl; a.B = p.tt + h->off[t] + (int64_t)x * dl * dr; a.ldb = dl;
        a.m = dl; a.n = dr; a.k = dl; a.C = Tp(t, x, 0); a.ldc = dl; a.nsum = 1;
        gi.push_back(a); st.t1 = std::max(st.t1, tiles64(dl, dr));
      }
    st.n1 = gi.size() - st.g1; st.g2 = gi.size();
    for (int t = 0; t < L - 1; ++t)
      for (int x = 0; x < Q; ++x) {
        const int dl = p.bd[t], dr = p.bd[t + 1];
        GemmItem a{};
        a.A = p.tt + h->off[t] + (int64_t)x * dl * dr; a.lda = dl; a.B = Tp(t, x, 0); a.ldb = dl;
        a.m = dr; a.n = dr; a.k = dl; a.C = Kp(K, t, x); a.ldc = dr; a.nsum = 1;
        gi.push_back(a); st.t2 = std::max(st.t2, tiles64(dr, dr));
      }
    st.n2 = gi.size() - st.g2;
    steps.push_back(st);
  }
  for (int64_t dist = 1; dist <= maxdist; ++dist) {
    Step st{}; st.t1 = 1; st.t2 = 1;
    st.f = fi.size();
    for (int t = 0; t + dist <= L - 1; ++t) {
      const int u = t + (int)dist, du = p.bd[u];
      for (int x = 0; x < Q; ++x)
        for (int y = 0; y < Q; ++y)
          fi.push_back(FrobItem{Kp(K, t, x), S + p.soff[u] + (int64_t)y * du * du, d_out + (poff[t] + dist - 1) * Q * Q + x + (size_t)Q * y,
                                (int64_t)du * du, 1, 0, 0, 0});
    }
    st.nf = fi.size() - st.f;
    st.g1 = gi.size();
    if (dist < maxdist) {
      for (int t = 0; t + dist + 1 <= L - 1; ++t) {
        const int u = t + (int)dist, du = p.bd[u], dn = p.bd[u + 1];
        for (int x = 0; x < Q; ++x)
          for (int y = 0; y < Q; ++y) {
            GemmItem a{};
            a.A = Kp(K, t, x); a.lda = du; a.B = p.tt + h->off[u] + (int64_t)y * du * dn; a.ldb = du;
            a.m = du; a.n = dn; a.k = du; a.C = Tp(t, x, y); a.ldc = du; a.nsum = 1;
            gi.push_back(a); st.t1 = std::max(st.t1, tiles64(du, dn));
          }
      }
      st.n1 = gi.size() - st.g1; st.g2 = gi.size(); st.s = si.size();
      for (int t = 0; t + dist + 1 <= L - 1; ++t) {
        const int u = t + (int)dist, du = p.bd[u], dn = p.bd[u + 1];
        for (int x = 0; x < Q; ++x) {
          GemmItem a{};
          a.A = p.tt + h->off[u]; a.lda = du; a.sA = (int64_t)du * dn; a.B = Tp(t, x, 0); a.ldb = du; a.sB = d2;
          a.m = dn; a.n = dn; a.k = du; a.C = Kp(Kraw, t, x); a.ldc = dn; a.nsum = Q;
          gi.push_back(a); st.t2 = std::max(st.t2, tiles64(dn, dn));
          si.push_back(SymItem{Kp(Kraw, t, x), Kp(Kraw, t, x), Kp(K, t, x), dn, t});
          ci.push_back(ScaleItem{Kp(K, t, x), (int64_t)dn * dn, t, 0});
        }
      }
      st.n2 = gi.size() - st.g2; st.ns = si.size() - st.s;
    }
    steps.push_back(st);
  }
  GemmItem* dg = nullptr; SymItem* ds = nullptr; ScaleItem* dc = nullptr; FrobItem* df = nullptr;
  {
    const size_t bg = sizeof(GemmItem) * gi.size(), bs = sizeof(SymItem) * si.size(), bc = sizeof(ScaleItem) * ci.size(), bf = sizeof(FrobItem) * fi.size();
    std::vector<unsigned char> blob(bg + bs + bc + bf + 64);
    memcpy(blob.data(), gi.data(), bg);
    memcpy(blob.data() + bg, si.data(), bs);
    memcpy(blob.data() + bg + bs, ci.data(), bc);
    memcpy(blob.data() + bg + bs + bc, fi.data(), bf);
    unsigned char* d_blob = upload_table(h, w.mps[9], blob, e);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return fail(TTB_ECUDA, "MPS pair marginals: %s", cudaGetErrorString(e));
    dg = reinterpret_cast<GemmItem*>(d_blob); ds = reinterpret_cast<SymItem*>(d_blob + bg);
    dc = reinterpret_cast<ScaleItem*>(d_blob + bg + bs); df = reinterpret_cast<FrobItem*>(d_blob + bg + bs + bc);
  }
  const unsigned gsz = (unsigned)std::min<int64_t>((d2 + 255) / 256, 64);
  for (size_t i = 0; i < steps.size(); ++i) {
    const Step& st = steps[i];
    if (i == 0) {
      launch_gemm<false, false>(h, dg + st.g1, (int)st.n1, st.t1);
      launch_gemm<true, false>(h, dg + st.g2, (int)st.n2, st.t2);
      continue;
    }
    if (st.nf) TTB_LAUNCH(h, "k_mps_frob", (k_mps_frob<<<(unsigned)st.nf, 256, 0, h->stream>>>(df + st.f)));
    if (st.n1) {
      launch_gemm<false, false>(h, dg + st.g1, (int)st.n1, st.t1);
      launch_gemm<true, false>(h, dg + st.g2, (int)st.n2, st.t2);
      cudaMemsetAsync(mx, 0, sizeof(unsigned long long) * L, h->stream);
      // (K_tu(x) + K_tu(x)^T) / 2 like accumulate_M (mps.jl:118); one common scale per first site keeps b_tu exact
      TTB_LAUNCH(h, "k_mps_sym", (k_mps_sym<<<dim3(gsz, (unsigned)st.ns), 256, 0, h->stream>>>(ds + st.s, mx)));
      TTB_LAUNCH(h, "k_mps_scale", (k_mps_scale<<<dim3(gsz, (unsigned)st.ns), 256, 0, h->stream>>>(dc + st.s, mx)));
    }
  }
  rc = tm.finish();
  if (rc != TTB_OK) return rc;
  TTB_CUDA(cudaMemcpy(out, d_out, sizeof(double) * npairs * Q * Q, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < npairs; ++i) {   // b ./= sum(b)  (mps.jl:227)
    double s = 0.0;
    for (int j = 0; j < Q * Q; ++j) s += out[i * Q * Q + j];
    for (int j = 0; j < Q * Q; ++j) out[i * Q * Q + j] /= s;
  }
  return TTB_OK;
}

int ttb_mps_sample(ttb_handle h, int64_t b, int64_t nsamples, uint64_t seed, uint64_t first_index, const double* uniforms,
                   uint8_t* x_out, double* p_out) {
  return ttb_mps_sample_env(h, b, nullptr, 0.0, nsamples, seed, first_index, uniforms, x_out, p_out);
}

int ttb_mps_sample_env(ttb_handle h, int64_t b, const double* r_env, double z_logabs, int64_t nsamples, uint64_t seed,
                       uint64_t first_index, const double* uniforms, uint8_t* x_out, double* p_out) {
  if (!x_out) return fail(TTB_EINVAL, "null x_out");
  if (nsamples < 0) return fail(TTB_EINVAL, "nsamples must be >= 0");
  if (nsamples == 0) return TTB_OK;
  MpsPlan p;
  TTB_TRY(mps_plan(h, b, p));
  if (p.Q > 256) return fail(TTB_EUNSUPPORTED, "sample supports Q <= 256 (uint8 output)");
  Workspace& w = h->ws;
  const int L = p.L;
  cudaError_t e = w.mps[8].ensure(sizeof(double) * ((size_t)nsamples * (uniforms ? L + 1 : 1) + 32) + (size_t)nsamples * L + 64);
  if (e != cudaSuccess) return fail(TTB_ENOMEM, "MPS sample buffers: %s", cudaGetErrorString(e));
  double* d_z = static_cast<double*>(w.mps[8].p);
  double* d_pq = d_z + 32;
  double* d_u = uniforms ? d_pq + nsamples : nullptr;
  uint8_t* d_x = reinterpret_cast<uint8_t*>(d_pq + nsamples + (uniforms ? (size_t)nsamples * L : 0));
  if (uniforms) TTB_CUDA(cudaMemcpyAsync(d_u, uniforms, sizeof(double) * nsamples * L, cudaMemcpyHostToDevice, h->stream));
  CallTimer tm(h);
  int rc = mps_provide_env(h, p, false, r_env, d_z);          // rz = accumulate_R(p) (mps.jl:265), or the caller's (r, z)
  if (rc == TTB_OK && r_env) {
    const double hz[2] = {z_logabs, 1.0};
    cudaMemcpyAsync(d_z, hz, sizeof(hz), cudaMemcpyHostToDevice, h->stream);
    cudaStreamSynchronize(h->stream);   // hz is a stack array
  }
  if (rc == TTB_OK) rc = mps_sandwich_R(h, p);
  if (rc != TTB_OK) return rc;
  std::vector<int64_t> tab(3 * (L + 1));
  for (int t = 0; t <= L; ++t) { tab[t] = h->off[t]; tab[L + 1 + t] = p.soff[t]; tab[2 * (L + 1) + t] = p.bd[t]; }
  int64_t* d_tab = upload_table(h, w.mps[9], tab, e);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(TTB_ECUDA, "MPS sample: %s", cudaGetErrorString(e));
  MpsSampCtx c{};
  c.slab = h->slab; c.tt_off = b * h->tt_stride; c.off = d_tab; c.bond = d_tab + 2 * (L + 1); c.L = L; c.Q = p.Q; c.dmax = p.dmax; c.d1 = p.d1;
  c.S = static_cast<const double*>(w.mps[5].p); c.soff = d_tab + L + 1; c.zlog = d_z;
  c.uniforms = d_u; c.seed = seed; c.first = first_index; c.x_out = d_x; c.p_out = p_out ? d_pq : nullptr; c.nsamples = nsamples;
  const size_t smem = sizeof(double) * (2 * (size_t)p.d1 * p.dmax + p.Q + 8);
  if (smem > 48 * 1024) return fail(TTB_EUNSUPPORTED, "MPS sample: d_1 * d_max too large for the one-CTA-per-draw kernel");
  TTB_LAUNCH(h, "k_mps_sample", (k_mps_sample<<<(unsigned)nsamples, 128, smem, h->stream>>>(c)));
  rc = tm.finish();
  if (rc != TTB_OK) return rc;
  TTB_CUDA(cudaMemcpy(x_out, d_x, (size_t)nsamples * L, cudaMemcpyDeviceToHost));
  if (p_out) TTB_CUDA(cudaMemcpy(p_out, d_pq, sizeof(double) * nsamples, cudaMemcpyDeviceToHost));
  return TTB_OK;
}

}  // extern "C"
