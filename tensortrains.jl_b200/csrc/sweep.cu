// sweep.cu -- orthogonalize_right!/left! and compress! on the device.
//
// Reference: src/tensor_train.jl:151-218, src/periodic_tensor_train.jl:83-141,
//            src/abstract_tensor_train.jl:261-274, rank rules src/svd_trunc.jl:33-143.
//
// One sweep step factorises the folded site X (p x n, p = bond*Q) as X ~= W * S with W (p x m')
// orthonormal and S = W^T X (m' x n); W becomes the new site tensor, S is absorbed into the
// neighbour (D = S * C_nbr(x) or C_nbr(x) * S^T), D is max-abs rescaled and becomes the next X.
// Both sweep directions share this form: the right sweep factorises X = M^T, the left sweep X = M.
//   * eps = 0 TruncThresh steps keep every direction, so W S = Q R from a blocked Householder QR
//     (gauge-equivalent to the reference's U / lambda V^T; only gauge-invariant results are compared).
//   * truncating steps: tall X -> QR then one-sided (Hestenes) Jacobi on the rows of R with the
//     rotations accumulated in J; wide X -> Jacobi directly on the rows of X.  The rotated rows ARE
//     S = Sigma V^T, W = Q J[:, sel] (or J[:, sel]); the rank rule runs in the same kernel so the
//     retained bond never visits the host.  Jacobi keeps the tiny singular values accurate to
//     O(eps * sigma_1), which the TruncThresh(1e-10) rank decision needs (a Gram-based SVD does not).
// All kernels read the current dimensions from the device bond array; a whole sweep is enqueued with
// no host synchronisation.
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <time.h>

#include "common.cuh"
#include "tma_mma.cuh"

namespace ttb {

struct SweepCtx {
  double* slab; int64_t tt_stride; const int64_t* off; int* bond; int L1; int Q;
  double* X0; double* X1; int64_t x_stride;
  double* T; int64_t t_stride; int bw;
  double* Bv; double* Jv; int64_t bv_stride, jv_stride; int ldv, ldj;
  int* perm; double* sigma; int n_max;
  StepState* st;
  double* logz;
  double* stat_err; int* stat_rank; double* spectrum;
  int* status;
  int* hint;               // mapped pinned ring [kHintRing][batch]: retained rank of the step, for the host
};

struct StepDesc {
  int site, nbr, dir;      // dir 0 = right sweep (X = M^T), 1 = left sweep (X = M)
  int xin;                 // which X buffer holds the folded input
  int rescale;             // max-abs rescale of the absorb output (0 only on the periodic wrap step)
  int qr_only;             // TruncThresh(0): keep everything -> QR
  int kind; double eps; int cap;
  int bond_idx, bond_idx2; // bond entries that receive m'
  int last;                // absorb output goes to the neighbour's site storage
  int slot;                // maxabs ping-pong slot
  int next_site, next_nbr; // the step that follows in this sweep (-1: none): k_commit snapshots its dimensions
  int kp, kn, kN;          // >= 0: the step's dimensions are known exactly on the host (eps = 0 sweeps of a single
                           // shape): kernels skip the dependent global loads of the device snapshot
  int trace;               // k_qr_stream: record the timeline of this step (TTB_QS_TRACE=<site>)
  int hint_slot;           // ring slot k_commit reports the retained ranks into (-1: none)
  int max_sweeps;          // Jacobi sweep cap (60; TTB_JACOBI_MAXSWEEPS lowers it so tests can force ST_NOCONV)
  int jc_sd;               // > 0: both Jacobi paths are launched (host bounds cannot tell): trains whose ACTUAL
                           // 2 n_max + k (n + k) doubles fit jc_sd are done by k_jacobi, the others by the cluster
};

struct Dims { int p, n, k, mode, N; };  // mode 0 = QR, 1 = SVD tall (p>=n), 2 = SVD wide

// dimensions of the step in flight, from the snapshot k_snapshot took when the step started (the bond
// array itself is rewritten by k_commit while side-stream kernels of the same step may still run)
__device__ __forceinline__ Dims get_dims(const SweepCtx& c, const StepDesc& s, int b) {
  Dims d;
  if (s.kp >= 0) { d.p = s.kp; d.n = s.kn; d.N = s.kN; }
  else {
    const StepState& st = c.st[b];
    d.p = st.p[s.slot]; d.n = st.n[s.slot]; d.N = st.N[s.slot];
  }
  d.k = min(d.p, d.n);
  d.mode = s.qr_only ? 0 : (d.p >= d.n ? 1 : 2);
  return d;
}

__device__ __forceinline__ void snapshot_dims(const SweepCtx& c, int b, int dir, int site, int nbr, int slot) {
  const int* bond = c.bond + (int64_t)b * c.L1;
  StepState& st = c.st[b];
  if (dir == 0) { st.p[slot] = bond[site + 1] * c.Q; st.n[slot] = bond[site]; st.N[slot] = bond[nbr]; }
  else          { st.p[slot] = bond[site] * c.Q;     st.n[slot] = bond[site + 1]; st.N[slot] = bond[nbr + 1]; }
}

// dimensions of the first step of a sweep (later steps are snapshotted by the previous step's k_commit)
__global__ void k_snapshot(SweepCtx c, StepDesc s, int batch) {
  pdl_enter();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  snapshot_dims(c, b, s.dir, s.site, s.nbr, s.slot);
}

__device__ __forceinline__ double* xbuf(const SweepCtx& c, int which, int b) {
  return (which ? c.X1 : c.X0) + (int64_t)b * c.x_stride;
}

// ---------------------------------------------------------------------------------------------
// fold: site storage -> X   (src/tensor_train.jl:159 right: free reshape, we store its transpose;
//                            :198 left: M[(m,x),n] |= C[m,n,x])
// ---------------------------------------------------------------------------------------------
__global__ void k_fold(SweepCtx c, StepDesc s) {
  pdl_enter();
  const int b = blockIdx.y;
  const Dims d = get_dims(c, s, b);
  const double* site = c.slab + (int64_t)b * c.tt_stride + c.off[s.site];
  double* X = xbuf(c, s.xin, b);
  const int64_t tot = (int64_t)d.p * d.n;
  const int dl = d.p / c.Q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
    if (s.dir == 0) {
      // site = X row-major (ld n): read coalesced, write strided
      const int j = (int)(i % d.n); const int r = (int)(i / d.n);
      X[r + (int64_t)d.p * j] = site[i];
    } else {
      const int r = (int)(i % d.p); const int j = (int)(i / d.p);
      const int m = r % dl, x = r / dl;
      X[i] = site[m + (int64_t)dl * (j + (int64_t)d.n * x)];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// blocked Householder QR: panel factorisation (one CTA per train)
// ---------------------------------------------------------------------------------------------
__global__ void k_qr_panel(SweepCtx c, StepDesc s, int panel) {
  pdl_enter();
  const int b = blockIdx.x;
  const Dims d = get_dims(c, s, b);
  if (d.mode == 2) return;
  const int bw = c.bw;
  const int j0 = panel * bw;
  if (panel == 0 && d.mode == 0 && threadIdx.x == 0) c.st[b].mprime[s.slot] = d.k;
  if (j0 >= d.k) return;
  const int bwc = min(bw, d.k - j0);
  const int rows = d.p - j0;
  double* X = xbuf(c, s.xin, b);
  extern __shared__ double sm[];
  double* Tm = sm;                 // bw*bw
  double* G = Tm + bw * bw;        // bw*bw
  double* tau = G + bw * bw;       // bw
  double* w = tau + bw;            // bw
  double* red = w + bw;            // 40
  double* P = red + 40;            // rows*bwc
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;

  for (int i = tid; i < rows * bwc; i += nt) {
    const int r = i % rows, cc = i / rows;
    P[i] = X[(j0 + r) + (int64_t)d.p * (j0 + cc)];
  }
  for (int i = tid; i < bw * bw; i += nt) { Tm[i] = 0.0; G[i] = 0.0; }
  __syncthreads();

  for (int cc = 0; cc < bwc; ++cc) {
    double* col = P + (size_t)rows * cc;
    double part = 0.0;
    for (int r = cc + 1 + tid; r < rows; r += nt) part += col[r] * col[r];
    const double xn2 = block_sum(part, red);
    const double alpha = col[cc];
    double tc = 0.0;
    if (xn2 > 0.0) {
      const double beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
      tc = (beta - alpha) / beta;
      const double scal = 1.0 / (alpha - beta);
      __syncthreads();  // everyone has read alpha
      for (int r = cc + 1 + tid; r < rows; r += nt) col[r] *= scal;
      if (tid == 0) col[cc] = beta;
    }
    if (tid == 0) tau[cc] = tc;
    __syncthreads();
    if (tc != 0.0 && cc + 1 < bwc) {
      for (int c2 = cc + 1 + wid; c2 < bwc; c2 += nw) {
        const double* col2 = P + (size_t)rows * c2;
        double sacc = 0.0;
        for (int r = cc + 1 + lane; r < rows; r += 32) sacc += col[r] * col2[r];
        sacc = warp_sum(sacc);
        if (lane == 0) w[c2] = tc * (sacc + col2[cc]);
      }
      __syncthreads();
      const int hr = rows - cc, hc = bwc - cc - 1;
      for (int i = tid; i < hr * hc; i += nt) {
        const int r = cc + i % hr, c2 = cc + 1 + i / hr;
        const double vr = (r == cc) ? 1.0 : col[r];
        P[r + (size_t)rows * c2] -= vr * w[c2];
      }
      __syncthreads();
    }
  }
  // Gram of the unit-lower-trapezoidal V for the compact-WY T factor (dlarft, forward columnwise)
  for (int pr = wid; pr < bwc * bwc; pr += nw) {
    const int l = pr % bwc, c2 = pr / bwc;
    if (l >= c2) continue;
    const double* cl = P + (size_t)rows * l;
    const double* cr = P + (size_t)rows * c2;
    double sacc = 0.0;
    for (int r = c2 + 1 + lane; r < rows; r += 32) sacc += cl[r] * cr[r];
    sacc = warp_sum(sacc);
    if (lane == 0) G[l + bw * c2] = sacc + cl[c2];
  }
  __syncthreads();
  if (wid == 0) {
    for (int c2 = 0; c2 < bwc; ++c2) {
      const double tc = tau[c2];
      if (lane == c2) Tm[c2 + bw * c2] = tc;
      if (lane < c2) {
        double acc = 0.0;
        for (int l = lane; l < c2; ++l) acc += Tm[lane + bw * l] * G[l + bw * c2];
        Tm[lane + bw * c2] = -tc * acc;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = tid; i < rows * bwc; i += nt) {
    const int r = i % rows, cc = i / rows;
    X[(j0 + r) + (int64_t)d.p * (j0 + cc)] = P[i];
  }
  double* Tg = c.T + (int64_t)b * c.t_stride + (int64_t)panel * bw * bw;
  for (int i = tid; i < bw * bw; i += nt) Tg[i] = Tm[i];
}

// trailing update  A <- (I - V T^T V^T) A  for one chunk of columns
__global__ void k_qr_update(SweepCtx c, StepDesc s, int panel, int cw) {
  pdl_enter();
  const int b = blockIdx.y;
  const Dims d = get_dims(c, s, b);
  if (d.mode == 2) return;
  const int bw = c.bw;
  const int j0 = panel * bw;
  if (j0 >= d.k) return;
  const int bwc = min(bw, d.k - j0);
  const int rows = d.p - j0;
  const int jc0 = j0 + bwc + blockIdx.x * cw;
  if (jc0 >= d.n) return;
  const int cwc = min(cw, d.n - jc0);
  double* X = xbuf(c, s.xin, b);
  extern __shared__ double sm[];
  double* Ts = sm;                    // bw*bw
  double* W = Ts + bw * bw;           // bw*cw
  double* W2 = W + bw * cw;           // bw*cw
  double* Vs = W2 + bw * cw;          // rows*bwc
  double* As = Vs + (size_t)rows * bwc;  // rows*cwc
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  const double* Tg = c.T + (int64_t)b * c.t_stride + (int64_t)panel * bw * bw;
  for (int i = tid; i < bw * bw; i += nt) Ts[i] = Tg[i];
  for (int i = tid; i < rows * bwc; i += nt) {
    const int r = i % rows, cc = i / rows;
    Vs[i] = r > cc ? X[(j0 + r) + (int64_t)d.p * (j0 + cc)] : (r == cc ? 1.0 : 0.0);
  }
  for (int i = tid; i < rows * cwc; i += nt) {
    const int r = i % rows, cc = i / rows;
    As[i] = X[(j0 + r) + (int64_t)d.p * (jc0 + cc)];
  }
  __syncthreads();
  for (int o = wid; o < bwc * cwc; o += nw) {
    const int i = o % bwc, cc = o / bwc;
    const double* v = Vs + (size_t)rows * i;
    const double* a = As + (size_t)rows * cc;
    double acc = 0.0;
    for (int r = i + lane; r < rows; r += 32) acc += v[r] * a[r];
    acc = warp_sum(acc);
    if (lane == 0) W[i + bw * cc] = acc;
  }
  __syncthreads();
  for (int o = tid; o < bwc * cwc; o += nt) {
    const int i = o % bwc, cc = o / bwc;
    double acc = 0.0;
    for (int l = 0; l <= i; ++l) acc += Ts[l + bw * i] * W[l + bw * cc];
    W2[i + bw * cc] = acc;
  }
  __syncthreads();
  for (int o = tid; o < rows * cwc; o += nt) {
    const int r = o % rows, cc = o / rows;
    double acc = 0.0;
    const int imax = min(r, bwc - 1);
    for (int i = 0; i <= imax; ++i) acc += Vs[r + (size_t)rows * i] * W2[i + bw * cc];
    X[(j0 + r) + (int64_t)d.p * (jc0 + cc)] = As[o] - acc;
  }
}

// ---------------------------------------------------------------------------------------------
// W = H_1 ... H_k * E for one chunk of columns, written straight into the site tensor
// (unfold: src/tensor_train.jl:165 right, :204 left)
// ---------------------------------------------------------------------------------------------
__global__ void k_apply_q(SweepCtx c, StepDesc s, int cw) {
  pdl_enter();
  const int b = blockIdx.y;
  const Dims d = get_dims(c, s, b);
  const int mprime = c.st[b].mprime[s.slot];
  const int c0 = blockIdx.x * cw;
  if (c0 >= mprime) return;
  const int cwc = min(cw, mprime - c0);
  const int bw = c.bw, p = d.p;
  const double* X = xbuf(c, s.xin, b);
  extern __shared__ double sm[];
  double* Ts = sm;                 // bw*bw
  double* W = Ts + bw * bw;        // bw*cw
  double* W2 = W + bw * cw;        // bw*cw
  double* Wc = W2 + bw * cw;       // p*cwc
  double* Vs = Wc + (size_t)p * cw;  // rows*bwc
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  const int* perm = c.perm + (int64_t)b * c.n_max;
  const double* Jv = c.Jv + (int64_t)b * c.jv_stride;
  for (int i = tid; i < p * cwc; i += nt) {
    const int r = i % p, cc = i / p;
    double v;
    if (d.mode == 0) v = (r == c0 + cc) ? 1.0 : 0.0;
    else {
      const int nv = d.k;
      v = r < nv ? Jv[(int64_t)perm[c0 + cc] * c.ldj + r] : 0.0;
    }
    Wc[i] = v;
  }
  __syncthreads();
  if (d.mode != 2) {
    const int np = (d.k + bw - 1) / bw;
    for (int panel = np - 1; panel >= 0; --panel) {
      const int j0 = panel * bw;
      if (d.mode == 0 && j0 > c0 + cwc - 1) continue;  // e_c untouched by panels right of it
      const int bwc = min(bw, d.k - j0);
      const int rows = p - j0;
      const double* Tg = c.T + (int64_t)b * c.t_stride + (int64_t)panel * bw * bw;
      for (int i = tid; i < bw * bw; i += nt) Ts[i] = Tg[i];
      for (int i = tid; i < rows * bwc; i += nt) {
        const int r = i % rows, cc = i / rows;
        Vs[i] = r > cc ? X[(j0 + r) + (int64_t)p * (j0 + cc)] : (r == cc ? 1.0 : 0.0);
      }
      __syncthreads();
      for (int o = wid; o < bwc * cwc; o += nw) {
        const int i = o % bwc, cc = o / bwc;
        const double* v = Vs + (size_t)rows * i;
        const double* a = Wc + (size_t)p * cc + j0;
        double acc = 0.0;
        for (int r = i + lane; r < rows; r += 32) acc += v[r] * a[r];
        acc = warp_sum(acc);
        if (lane == 0) W[i + bw * cc] = acc;
      }
      __syncthreads();
      for (int o = tid; o < bwc * cwc; o += nt) {
        const int i = o % bwc, cc = o / bwc;
        double acc = 0.0;
        for (int l = i; l < bwc; ++l) acc += Ts[i + bw * l] * W[l + bw * cc];
        W2[i + bw * cc] = acc;
      }
      __syncthreads();
      for (int o = tid; o < rows * cwc; o += nt) {
        const int r = o % rows, cc = o / rows;
        double acc = 0.0;
        const int imax = min(r, bwc - 1);
        for (int i = 0; i <= imax; ++i) acc += Vs[r + (size_t)rows * i] * W2[i + bw * cc];
        Wc[(size_t)p * cc + j0 + r] -= acc;
      }
      __syncthreads();
    }
  }
  double* site = c.slab + (int64_t)b * c.tt_stride + c.off[s.site];
  if (s.dir == 0) {
    for (int i = tid; i < p * cwc; i += nt) {
      const int cc = i % cwc, r = i / cwc;
      site[(c0 + cc) + (int64_t)mprime * r] = Wc[r + (size_t)p * cc];
    }
  } else {
    const int dl = p / c.Q;
    for (int i = tid; i < p * cwc; i += nt) {
      const int r = i % p, cc = i / p;
      const int m = r % dl, x = r / dl;
      site[m + (int64_t)dl * ((c0 + cc) + (int64_t)mprime * x)] = Wc[i];
    }
  }
}

}  // namespace ttb
#ifndef TTB_PANEL_NS
#define TTB_PANEL_NS 4
#endif
#include "qr_fast.cuh"
#include "qr_panel.cuh"
#include "qr_stream.cuh"
namespace ttb {

// ---------------------------------------------------------------------------------------------
// rank rules  (src/svd_trunc.jl:38-47, :74, :98-100, :130-140); lam descending, m = length
// ---------------------------------------------------------------------------------------------
__device__ int rank_rule(const double* lam, int m, int kind, double eps, int cap, double* err_out) {
  double total = 0.0;
  for (int i = m - 1; i >= 0; --i) total += lam[i] * lam[i];
  int mprime = 1;
  if (kind == TTB_TRUNC_THRESH || kind == TTB_TRUNC_BONDTHRESH) {
    const double ln = sqrt(total) * eps;
    const double thr = ln * ln;
    double sacc = (kind == TTB_TRUNC_THRESH) ? lam[m - 1] * lam[m - 1] : 0.0;
    for (int k = m - 1; k >= 1; --k) {  // 1-based k
      sacc += lam[k - 1] * lam[k - 1];
      if (sacc >= thr) { mprime = k + 1; break; }
    }
    if (kind == TTB_TRUNC_BONDTHRESH) mprime = min(mprime, cap);
  } else {
    mprime = min(m, cap);
  }
  if (mprime < 1) mprime = 1;
  double tail = 0.0;
  for (int i = m - 1; i >= mprime; --i) tail += lam[i] * lam[i];
  *err_out = total > 0.0 ? sqrt(tail / total) : 0.0;
  return mprime;
}

// ---------------------------------------------------------------------------------------------
// one-sided (Hestenes) Jacobi: one rotation of the row pair (bi, bj) and of the accumulated rotations
// (ji, jj), executed by one warp.  Squared norms are cached in sqi/sqj (a rotation updates them exactly:
// |bi'|^2 = al - t ga, |bj'|^2 = be + t ga), so a pair costs ONE dot + one warp reduction.  The first
// 32*NE elements of both rows (32*NJ of the rotation rows) stay in registers between the dot and the
// update: a round is a latency chain (load -> dot -> 5 shuffles -> angle -> store), not a throughput
// problem, so every shared-memory round trip taken off it counts.  Returns true when it made a rotation that
// counts against convergence (kJacobiBig2, common.cuh).
// ---------------------------------------------------------------------------------------------
template <int NE, int NJ>
__device__ __forceinline__ bool jacobi_pair(double* bi, double* bj, double* ji, double* jj, int len, int nv,
                                            double* sqi, double* sqj, double tol2, int lane) {
  double xi[NE], yj[NE], jx[NJ], jy[NJ];
#pragma unroll
  for (int m = 0; m < NE; ++m) {
    const int e = lane + 32 * m;
    const bool ok = e < len;
    xi[m] = ok ? bi[e] : 0.0;
    yj[m] = ok ? bj[e] : 0.0;
  }
#pragma unroll
  for (int m = 0; m < NJ; ++m) {
    const int e = lane + 32 * m;
    const bool ok = e < nv;
    jx[m] = ok ? ji[e] : 0.0;
    jy[m] = ok ? jj[e] : 0.0;
  }
  const double al = *sqi, be = *sqj;
  double g0 = 0.0, g1 = 0.0;
#pragma unroll
  for (int m = 0; m < NE; m += 2) { g0 += xi[m] * yj[m]; g1 += xi[m + 1] * yj[m + 1]; }
  for (int e = lane + 32 * NE; e < len; e += 32) g0 += bi[e] * bj[e];
  const double ga = warp_sum(g0 + g1);
  if (!(al > 0.0 && be > 0.0 && ga * ga > tol2 * al * be)) return false;
  // rotation angle from tan(2 theta) = 2 ga / (be - al), |theta| <= pi/4, with two rsqrt and no
  // division on the critical path:  cos 2th = |da|/h, sin 2th = sign(da) b2/h,
  // cos th = sqrt((1 + cos 2th)/2), sin th = sin 2th / (2 cos th)
  const double da = be - al, b2 = 2.0 * ga;
  const double h2 = da * da + b2 * b2;
  double r = rsqrt_nr(h2), das = da, b2s = b2;
  if (!(h2 > 1e-280 && h2 < 1e280)) {  // far outside the normal range: rescale first (rare, warp-uniform)
    const double sc = 1.0 / fmax(fabs(da), fabs(b2));
    das = da * sc; b2s = b2 * sc;
    r = rsqrt(das * das + b2s * b2s);
  }
  const double c2 = fabs(das) * r, s2 = (da < 0.0 ? -b2s : b2s) * r;
  const double u = 0.5 + 0.5 * c2;
  const double ic = rsqrt_nr(u);   // = 1 / cos(theta), u in [0.5, 1]
  const double cs = u * ic, sn = 0.5 * s2 * ic;
#pragma unroll
  for (int m = 0; m < NE; ++m) {
    const int e = lane + 32 * m;
    if (e < len) {
      bi[e] = cs * xi[m] - sn * yj[m];
      bj[e] = sn * xi[m] + cs * yj[m];
    }
  }
  for (int e = lane + 32 * NE; e < len; e += 32) {
    const double x = bi[e], y = bj[e];
    bi[e] = cs * x - sn * y;
    bj[e] = sn * x + cs * y;
  }
#pragma unroll
  for (int m = 0; m < NJ; ++m) {
    const int e = lane + 32 * m;
    if (e < nv) {
      ji[e] = cs * jx[m] - sn * jy[m];
      jj[e] = sn * jx[m] + cs * jy[m];
    }
  }
  for (int e = lane + 32 * NJ; e < nv; e += 32) {
    const double x = ji[e], y = jj[e];
    ji[e] = cs * x - sn * y;
    jj[e] = sn * x + cs * y;
  }
  const double t = sn * ic;   // tan(theta) without a division
  *sqi = al - t * ga;         // same value from every lane
  *sqj = be + t * ga;
  return ga * ga > kJacobiBig2 * al * be;
}

// round-robin (circle method) pairing of n (even) players: pair q of round rd, i < j
__device__ __forceinline__ void rr_pair(int n, int rd, int q, int& i, int& j) {
  if (q == 0) { i = n - 1; j = rd; }
  else {
    i = rd + q; if (i >= n - 1) i -= n - 1;
    j = rd - q; if (j < 0) j += n - 1;
  }
  if (i > j) { const int t = i; i = j; j = t; }
}

// After convergence: singular values = row norms -> sort -> rank rule on the device (the retained rank never
// visits the host), sigma / spectrum / stats.  B: the rotated rows (shared or global), sig/srt: n_max
// doubles of shared scratch each.  Called by every thread of the CTA.
__device__ __forceinline__ void jacobi_finish(const SweepCtx& c, const StepDesc& s, int b, const double* B, int ldb,
                                              int nv, int len, double* sig, double* srt) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  for (int v = wid; v < nv; v += nw) {
    const double* bi = B + (size_t)v * ldb;
    double al = 0.0;
    for (int e = lane; e < len; e += 32) al += bi[e] * bi[e];
    al = warp_sum(al);
    if (lane == 0) sig[v] = sqrt(al);
  }
  __syncthreads();
  int* perm = c.perm + (int64_t)b * c.n_max;
  for (int i = tid; i < nv; i += nt) {
    const double si = sig[i];
    int pos = 0;
    for (int j = 0; j < nv; ++j) {
      const double sj = sig[j];
      pos += (sj > si || (sj == si && j < i)) ? 1 : 0;
    }
    srt[pos] = si;
    perm[pos] = i;
  }
  __syncthreads();
  if (tid == 0) {
    double err = 0.0;
    const int mp = rank_rule(srt, nv, s.kind, s.eps, s.cap, &err);
    c.st[b].mprime[s.slot] = mp;
    c.stat_err[(int64_t)b * c.L1 + s.bond_idx] = err;
  }
  double* sg = c.sigma + (int64_t)b * c.n_max;
  for (int i = tid; i < nv; i += nt) sg[i] = srt[i];
  if (c.spectrum) {
    double* sp = c.spectrum + ((int64_t)b * c.L1 + s.bond_idx) * c.n_max;
    for (int i = tid; i < c.n_max; i += nt) sp[i] = i < nv ? srt[i] : -1.0;
  }
}

// ---------------------------------------------------------------------------------------------
// one-sided Jacobi SVD on the rows of R (tall) or X (wide) + rank rule.  One CTA per train.
// ---------------------------------------------------------------------------------------------
// `smem_doubles` = dynamic shared memory the launch provides; the vectors live in shared memory when
// the ACTUAL (device-known) sizes fit, in the global workspace otherwise.
// MAXT: largest block the instantiation is launched with (register budget); NE: row elements per lane
// held in registers across a rotation (32*NE covers the row for the common sizes).
template <int MAXT, int NE>
__global__ void __launch_bounds__(MAXT) k_jacobi(SweepCtx c, StepDesc s, int smem_doubles) {
  pdl_enter();
  const int b = blockIdx.x;
  const Dims d = get_dims(c, s, b);
  if (d.mode == 0) return;
  const int nv = d.k, len = d.n, p = d.p;
  const double* X = xbuf(c, s.xin, b);
  const double scale_in = c.st[b].scale_in[s.slot];  // pending 1/max|D| of the previous step (lazy rescale)
  double* Bg = c.Bv + (int64_t)b * c.bv_stride;
  double* Jg = c.Jv + (int64_t)b * c.jv_stride;
  extern __shared__ double sm[];
  double* sig = sm;                 // n_max
  double* srt = sig + c.n_max;      // n_max
  const bool SMEM = 2 * (int64_t)c.n_max + (int64_t)nv * (len + nv) <= (int64_t)smem_doubles;
  if (s.jc_sd > 0 && !SMEM) return;  // the cluster kernels take this train
  double* B = SMEM ? (srt + c.n_max) : Bg;
  const int ldb = SMEM ? len : c.ldv;
  double* J = SMEM ? (B + (size_t)nv * ldb) : Jg;
  const int ldjj = SMEM ? nv : c.ldj;
  __shared__ int rotated;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;

  for (int i = tid; i < nv * len; i += nt) {
    const int r = i % nv, j = i / nv;  // coalesced read of column j of X
    double v = X[r + (int64_t)p * j] * scale_in;
    if (d.mode == 1 && r > j) v = 0.0;
    B[(size_t)r * ldb + j] = v;
  }
  for (int i = tid; i < nv * nv; i += nt) {
    const int e = i % nv, v = i / nv;
    J[(size_t)v * ldjj + e] = (e == v) ? 1.0 : 0.0;
  }
  __syncthreads();

  // Cyclic ordering: nvp-1 rounds of nvp/2 disjoint pairs, one warp per pair, one block barrier per
  // round; squared norms refreshed every sweep.  8-10 sweeps on the sweep matrices.
  const int nvp = (nv + 1) & ~1;
  const int npairs = nvp / 2, rounds = nvp - 1;
  const double tol = sqrt((double)max(len, 2)) * 2.220446049250313e-16;
  const double tol2 = tol * tol;
  double* sq = sig;  // cached squared norms live in the sigma scratch until the end
  bool conv = (nv < 2);
  for (int sweep = 0; sweep < s.max_sweeps && !conv; ++sweep) {
    for (int v = wid; v < nv; v += nw) {
      const double* bi = B + (size_t)v * ldb;
      double a0 = 0.0, a1 = 0.0;
      int e = lane;
      for (; e + 32 < len; e += 64) { a0 += bi[e] * bi[e]; a1 += bi[e + 32] * bi[e + 32]; }
      if (e < len) a0 += bi[e] * bi[e];
      const double al = warp_sum(a0 + a1);
      sq[v] = al;
    }
    rotated = 0;
    __syncthreads();
    for (int rd = 0; rd < rounds; ++rd) {
      for (int q = wid; q < npairs; q += nw) {
        int i, j;
        rr_pair(nvp, rd, q, i, j);
        if (j >= nv) continue;
        if (jacobi_pair<NE, 2>(B + (size_t)i * ldb, B + (size_t)j * ldb, J + (size_t)i * ldjj, J + (size_t)j * ldjj,
                               len, nv, sq + i, sq + j, tol2, lane))
          rotated = 1;
      }
      __syncthreads();
    }
    conv = (rotated == 0);
    __syncthreads();
  }
  if (!conv && tid == 0) atomicOr(c.status, ST_NOCONV);
  jacobi_finish(c, s, b, B, ldb, nv, len, sig, srt);
  if (SMEM) {
    for (int i = tid; i < nv * len; i += nt) {
      const int e = i % len, v = i / len;
      Bg[(int64_t)v * c.ldv + e] = B[(size_t)v * ldb + e];
    }
    for (int i = tid; i < nv * nv; i += nt) {
      const int e = i % nv, v = i / nv;
      Jg[(int64_t)v * c.ldj + e] = J[(size_t)v * ldjj + e];
    }
  }
}

}  // namespace ttb
#include "jacobi_cluster.cuh"
#include "jacobi_small.cuh"
#include "jacobi_sys.cuh"
#include "jacobi_lq.cuh"
namespace ttb {

// ---------------------------------------------------------------------------------------------
// absorb: Xn[(a + m'x), j] = sum_l S(a,l) * Nbr_x(l,j)   (src/tensor_train.jl:168 / :207)
// plus max|D| for the rescale (:169-173 / :208-212).  FP64 tile GEMM, 64x64 tile, 4x4 per thread.
// ---------------------------------------------------------------------------------------------
// commit the step for one train: z /= m, bond <- m', stats, reset the other maxabs slot, next step's
// snapshot and pending scale, rank feedback for the host
__device__ __forceinline__ void commit_train(const SweepCtx& c, const StepDesc& s, int b, int batch) {
  int* bond = c.bond + (int64_t)b * c.L1;
  const int mp = c.st[b].mprime[s.slot];
  // L2 read: the other CTAs' atomicMax results; this CTA's L1 may hold the line from its dimension loads
  const double m = __longlong_as_double((long long)__ldcg(&c.st[b].maxabs_bits[s.slot]));
  const bool bad = (m != m) || isinf(m);
  if (s.rescale && !bad && m != 0.0) c.logz[b] -= log(m);
  if (s.rescale && bad) atomicOr(c.status, ST_NONFINITE);
  c.st[b].maxabs_bits[s.slot ^ 1] = 0ull;
  c.st[b].scale_in[s.slot ^ 1] = (s.rescale && !bad && m != 0.0 && !s.last) ? 1.0 / m : 1.0;
  c.stat_rank[(int64_t)b * c.L1 + s.bond_idx] = mp;
  if (s.qr_only) c.stat_err[(int64_t)b * c.L1 + s.bond_idx] = 0.0;
  bond[s.bond_idx] = mp;
  if (s.bond_idx2 >= 0) bond[s.bond_idx2] = mp;
  if (s.next_site >= 0) snapshot_dims(c, b, s.dir, s.next_site, s.next_nbr, s.slot ^ 1);
  if (s.hint_slot >= 0) {
    c.hint[(int64_t)s.hint_slot * batch + b] = mp;
    __threadfence_system();
  }
}

// DMMA variant: 32x32 output tile per CTA (4 warps, each 16x16 = two m16n8k8 FP64 tensor-core tiles),
// K staged 32 at a time through shared memory with register prefetch of the next stage.
#define AM_T 32
#define AM_K 32
#define AM_LD 36  // row stride == 4 (mod 16): fragment loads are conflict-free per half-warp
__global__ void __launch_bounds__(128) k_absorb_mma(SweepCtx c, StepDesc s) {
  pdl_enter();
  const int b = blockIdx.z / c.Q, x = blockIdx.z % c.Q;
  const Dims d = get_dims(c, s, b);
  const int mprime = c.st[b].mprime[s.slot];
  const int i0 = blockIdx.x * AM_T, j0 = blockIdx.y * AM_T;
  const bool active = !(i0 >= mprime || j0 >= d.N);   // surplus CTAs still take a ticket below
  const int n = d.n, N = d.N;
  const int64_t p = d.p;
  const double* X = xbuf(c, s.xin, b);
  const double* Bv = c.Bv + (int64_t)b * c.bv_stride;
  const int* perm = c.perm + (int64_t)b * c.n_max;
  const double* nbr = c.slab + (int64_t)b * c.tt_stride + c.off[s.nbr];
  double* Xn = xbuf(c, s.xin ^ 1, b);
  const double scale_in = c.st[b].scale_in[s.slot];  // R of an unscaled X (lazy rescale); Bv is already scaled
  __shared__ double Ss[AM_K][AM_LD];
  __shared__ double Bs[AM_K][AM_LD];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int wm = (wid & 1) * 16, wn = (wid >> 1) * 16;
  if (active) {
    double acc[2][4];
  #pragma unroll
    for (int a = 0; a < 2; ++a)
  #pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][e] = 0.0;
    double rs[8], rb[8];  // prefetch registers: 1024 elements / 128 threads
    auto fetch = [&](int l0) {
  #pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = tid + 128 * u;
        int ii, ll;
        if (d.mode == 0) { ii = i % AM_T; ll = i / AM_T; }
        else { ll = i % AM_K; ii = i / AM_K; }
        const int gi = i0 + ii, gl = l0 + ll;
        double v = 0.0;
        if (gi < mprime && gl < n) {
          if (d.mode == 0) v = gi <= gl ? X[gi + p * gl] * scale_in : 0.0;
          else v = Bv[(int64_t)perm[gi] * c.ldv + gl];
        }
        rs[u] = v;
        int l2, jj;
        if (s.dir == 1) { l2 = i % AM_K; jj = i / AM_K; }
        else { jj = i % AM_T; l2 = i / AM_T; }
        const int gl2 = l0 + l2, gj = j0 + jj;
        double w = 0.0;
        if (gl2 < n && gj < N)
          w = s.dir == 1 ? nbr[gl2 + (int64_t)n * (gj + (int64_t)N * x)] : nbr[gj + (int64_t)N * (gl2 + (int64_t)n * x)];
        rb[u] = w;
      }
    };
    auto stash = [&]() {
  #pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = tid + 128 * u;
        int ii, ll;
        if (d.mode == 0) { ii = i % AM_T; ll = i / AM_T; }
        else { ll = i % AM_K; ii = i / AM_K; }
        Ss[ll][ii] = rs[u];
        int l2, jj;
        if (s.dir == 1) { l2 = i % AM_K; jj = i / AM_K; }
        else { jj = i % AM_T; l2 = i / AM_T; }
        Bs[l2][jj] = rb[u];
      }
    };
    // mode 0: S is upper triangular -> K range starts at the tile's first row
    const int lstart = (d.mode == 0) ? (i0 / AM_K) * AM_K : 0;
    fetch(lstart);
    for (int l0 = lstart; l0 < n; l0 += AM_K) {
      __syncthreads();
      stash();
      __syncthreads();
      if (l0 + AM_K < n) fetch(l0 + AM_K);
  #pragma unroll
      for (int kk = 0; kk < AM_K; kk += 8) {
        double af[4];
        af[0] = Ss[kk + t4][wm + g];
        af[1] = Ss[kk + t4][wm + g + 8];
        af[2] = Ss[kk + t4 + 4][wm + g];
        af[3] = Ss[kk + t4 + 4][wm + g + 8];
  #pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          double bf[2];
          bf[0] = Bs[kk + t4][wn + nt * 8 + g];
          bf[1] = Bs[kk + t4 + 4][wn + nt * 8 + g];
          dmma_16x8x8(acc[nt], af, bf);
        }
      }
    }
    const int64_t ldo = (int64_t)mprime * c.Q;
    double mx = 0.0;
  #pragma unroll
    for (int nt = 0; nt < 2; ++nt)
  #pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int gi = i0 + wm + g + ((e & 2) ? 8 : 0);
        const int gj = j0 + wn + nt * 8 + t4 * 2 + (e & 1);
        if (gi < mprime && gj < N) {
          Xn[(gi + (int64_t)mprime * x) + ldo * gj] = acc[nt][e];
          mx = absmax_nan(mx, acc[nt][e]);
        }
      }
    unsigned long long bits = (unsigned long long)__double_as_longlong(mx);
  #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o);
      bits = other > bits ? other : bits;
    }
    if (lane == 0) atomicMax(&c.st[b].maxabs_bits[s.slot], bits);
  }
  // the last CTA of the train to finish commits the step (z /= m, bond <- m', next step's snapshot and
  // scale, rank feedback): one kernel launch and one dependent-launch gap less per site step
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const int total = (int)(gridDim.x * gridDim.y) * c.Q;
    if (atomicAdd(&c.st[b].ticket, 1) == total - 1) {
      __threadfence();
      c.st[b].ticket = 0;
      commit_train(c, s, b, (int)gridDim.z / c.Q);
    }
  }
}

// Pipelined absorb for the large eps = 0 steps (mode 0, dimensions exact on the host and multiples of 32: every
// bulk step of config 3's right sweep).  k_absorb_mma walks K in stages of 32 with ONE stage of register
// prefetch: every stage pays a global-load round trip (~1 us) and two barriers, ~20 us for 67 MFLOP, on the
// critical path between two site QRs.  Here both operand tiles travel global -> shared with cp.async
// (16-byte chunks, no registers), three stages deep, so after the first round trip a stage costs its 8 DMMAs
// per warp and one barrier; the lazy scale of R (scale_in) is applied to the accumulators (the product is
// linear); the below-diagonal part of the diagonal K-block (Householder vectors sharing storage with R) is
// masked in the fragments of that one stage.
#define AP_STAGES 3
__device__ __forceinline__ void ap_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__global__ void __launch_bounds__(128) k_absorb_pipe(SweepCtx c, StepDesc s) {
  pdl_enter();
  const int b = blockIdx.z / c.Q, x = blockIdx.z % c.Q;
  const Dims d = get_dims(c, s, b);          // exact, host known
  const int n = d.n, N = d.N, mprime = d.k;  // eps = 0: every direction is kept
  const int64_t p = d.p;
  const int i0 = blockIdx.x * AM_T, j0 = blockIdx.y * AM_T;
  const double* X = xbuf(c, s.xin, b);
  const double* nbr = c.slab + (int64_t)b * c.tt_stride + c.off[s.nbr];
  double* Xn = xbuf(c, s.xin ^ 1, b);
  const double scale_in = c.st[b].scale_in[s.slot];
  extern __shared__ __align__(16) double apsm[];   // [AP_STAGES][2][32][AM_LD]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int wm = (wid & 1) * 16, wn = (wid >> 1) * 16;
  // stage = S tile [l][i] (rows l of R^T: 32 consecutive i per l) + neighbour tile ([l][j] for dir 0, [j][l] for dir 1)
  auto issue = [&](int l0, int st) {
    double* Ss = apsm + (size_t)st * 2 * 32 * AM_LD;
    double* Bs = Ss + 32 * AM_LD;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int ch = tid + 128 * u, row = ch >> 4, c2 = (ch & 15) * 2;   // 32 rows x 16 chunks of two doubles
      ap_cp16(Ss + row * AM_LD + c2, X + (i0 + c2) + p * (l0 + row));
      if (s.dir == 0) ap_cp16(Bs + row * AM_LD + c2, nbr + (j0 + c2) + (int64_t)N * ((l0 + row) + (int64_t)n * x));
      else            ap_cp16(Bs + row * AM_LD + c2, nbr + (l0 + c2) + (int64_t)n * ((j0 + row) + (int64_t)N * x));
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int lstart = i0;                 // R is upper triangular: K starts at the tile's first row
  const int nst = (n - lstart) / AM_K;
#pragma unroll
  for (int q = 0; q < AP_STAGES - 1; ++q) {
    if (q < nst) issue(lstart + q * AM_K, q);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  double acc[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
  for (int q = 0; q < nst; ++q) {
    asm volatile("cp.async.wait_group %0;" ::"n"(AP_STAGES - 2) : "memory");
    __syncthreads();                     // stage q landed for everybody; stage q-1's readers are done
    if (q + AP_STAGES - 1 < nst) issue(lstart + (q + AP_STAGES - 1) * AM_K, (q + AP_STAGES - 1) % AP_STAGES);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    const double* Ss = apsm + (size_t)(q % AP_STAGES) * 2 * 32 * AM_LD;
    const double* Bs = Ss + 32 * AM_LD;
    const bool diag = (q == 0);          // l0 == i0: entries with l < i belong to the Householder vectors
#pragma unroll
    for (int kk = 0; kk < AM_K; kk += 8) {
      double af[4];
      af[0] = Ss[(kk + t4) * AM_LD + wm + g];
      af[1] = Ss[(kk + t4) * AM_LD + wm + g + 8];
      af[2] = Ss[(kk + t4 + 4) * AM_LD + wm + g];
      af[3] = Ss[(kk + t4 + 4) * AM_LD + wm + g + 8];
      if (diag) {
        if (kk + t4 < wm + g) af[0] = 0.0;
        if (kk + t4 < wm + g + 8) af[1] = 0.0;
        if (kk + t4 + 4 < wm + g) af[2] = 0.0;
        if (kk + t4 + 4 < wm + g + 8) af[3] = 0.0;
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        double bf[2];
        if (s.dir == 0) {
          bf[0] = Bs[(kk + t4) * AM_LD + wn + nt * 8 + g];
          bf[1] = Bs[(kk + t4 + 4) * AM_LD + wn + nt * 8 + g];
        } else {
          bf[0] = Bs[(wn + nt * 8 + g) * AM_LD + kk + t4];
          bf[1] = Bs[(wn + nt * 8 + g) * AM_LD + kk + t4 + 4];
        }
        dmma_16x8x8(acc[nt], af, bf);
      }
    }
  }
  const int64_t ldo = (int64_t)mprime * c.Q;
  double mx = 0.0;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int gi = i0 + wm + g + ((e & 2) ? 8 : 0);
      const int gj = j0 + wn + nt * 8 + t4 * 2 + (e & 1);
      const double v = acc[nt][e] * scale_in;
      Xn[(gi + (int64_t)mprime * x) + ldo * gj] = v;
      mx = absmax_nan(mx, v);
    }
  unsigned long long bits = (unsigned long long)__double_as_longlong(mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o);
    bits = other > bits ? other : bits;
  }
  if (lane == 0) atomicMax(&c.st[b].maxabs_bits[s.slot], bits);
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const int total = (int)(gridDim.x * gridDim.y) * c.Q;
    if (atomicAdd(&c.st[b].ticket, 1) == total - 1) {
      __threadfence();
      c.st[b].ticket = 0;
      commit_train(c, s, b, (int)gridDim.z / c.Q);
    }
  }
}

// Small absorb (m', n, N <= 32: every step of a config-5 batch).  The tiled kernel above launches Q tiny CTAs
// per train (168 registers each, a ticket and an atomic per CTA) and is bound by CTA turnover; here ONE CTA
// per train stages S once, walks x, multiplies on the tensor cores out of zero-padded shared memory
// (leading dimension 36: conflict-free fragments), takes the block max and commits the step itself.
__global__ void __launch_bounds__(128, 8) k_absorb_small(SweepCtx c, StepDesc s) {
  pdl_enter();
  constexpr int LD = 36;
  const int b = blockIdx.x;
  const Dims d = get_dims(c, s, b);
  const int mprime = c.st[b].mprime[s.slot];
  const int n = d.n, N = d.N;
  const int64_t p = d.p;
  const double* X = xbuf(c, s.xin, b);
  const double* Bv = c.Bv + (int64_t)b * c.bv_stride;
  const int* perm = c.perm + (int64_t)b * c.n_max;
  const double* nbr = c.slab + (int64_t)b * c.tt_stride + c.off[s.nbr];
  double* Xn = xbuf(c, s.xin ^ 1, b);
  const double scale_in = c.st[b].scale_in[s.slot];
  __shared__ double Ss[LD * 32];   // S(a, l)     at a + LD l
  __shared__ double Ns[LD * 32];   // Nbr_x(l, j) at l + LD j
  __shared__ double red[4];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, tq = lane & 3;
  const int mt = wid & 1, nt0 = (wid >> 1) * 2;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int idx = tid + 128 * m;
    int a, l;
    if (d.mode == 0) { a = idx & 31; l = idx >> 5; } else { l = idx & 31; a = idx >> 5; }
    double v = 0.0;
    if (a < mprime && l < n) {
      if (d.mode == 0) v = a <= l ? X[a + p * l] * scale_in : 0.0;
      else v = Bv[(int64_t)perm[a] * c.ldv + l];
    }
    Ss[a + LD * l] = v;
  }
  const int64_t ldo = (int64_t)mprime * c.Q;
  double mx = 0.0;
  for (int x = 0; x < c.Q; ++x) {
    double w[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int idx = tid + 128 * m;
      int l, j;
      if (s.dir == 1) { l = idx & 31; j = idx >> 5; } else { j = idx & 31; l = idx >> 5; }
      w[m] = 0.0;
      if (l < n && j < N) w[m] = s.dir == 1 ? nbr[l + (int64_t)n * (j + (int64_t)N * x)] : nbr[j + (int64_t)N * (l + (int64_t)n * x)];
    }
    if (x > 0) __syncthreads();   // the previous x's MMA has read Ns
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int idx = tid + 128 * m;
      int l, j;
      if (s.dir == 1) { l = idx & 31; j = idx >> 5; } else { j = idx & 31; l = idx >> 5; }
      Ns[l + LD * j] = w[m];
    }
    __syncthreads();
    double acc[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
    mma32_warp<1, LD, 1, LD>(acc, Ss, Ns, mt, nt0, g, tq);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int gi = mt * 16 + g + ((e & 2) ? 8 : 0);
        const int gj = (nt0 + nt) * 8 + tq * 2 + (e & 1);
        if (gi < mprime && gj < N) {
          Xn[(gi + (int64_t)mprime * x) + ldo * gj] = acc[nt][e];
          mx = absmax_nan(mx, acc[nt][e]);
        }
      }
  }
  __syncthreads();
  const double m = block4_max_bits(mx, red, lane, wid);
  if (tid == 0) {
    atomicMax(&c.st[b].maxabs_bits[s.slot], (unsigned long long)__double_as_longlong(m));
    __threadfence();
    commit_train(c, s, b, (int)gridDim.x);
  }
}

// Skinny absorb (m' <= 32 retained directions: every step of a truncating sweep once the bonds are small).
// The tiled kernel above walks K in 8 synchronised stages per CTA and has only a handful of CTAs to run:
// ~16 us of pure latency for 4 MFLOP.  Here one CTA owns 8 output columns, loads ALL of S (m' x n) and its
// n x 8 neighbour tile in one shot (every load in flight at once), the four warps split K for the DMMAs and
// are summed through shared memory.
#define AS_TN 8
__host__ __device__ inline int as_ld(int n) { return ((n + 15) & ~15) + 4; }   // = 4 (mod 16): conflict-free fragments
__global__ void __launch_bounds__(128) k_absorb_skinny(SweepCtx c, StepDesc s) {
  pdl_enter();
  const int b = blockIdx.z / c.Q, x = blockIdx.z % c.Q;
  const Dims d = get_dims(c, s, b);
  const int mprime = c.st[b].mprime[s.slot];   // host guarantees <= 32
  const int j0 = blockIdx.y * AS_TN;
  const bool active = j0 < d.N;
  const int n = d.n, N = d.N;
  const int64_t p = d.p;
  const int ld = as_ld(n);
  extern __shared__ double smk[];
  double* Ss = smk;                 // [32][ld]
  double* Bs = Ss + 32 * ld;        // [AS_TN][ld]
  double* red = Bs + AS_TN * ld;    // [4 warps][2][32 lanes][4]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (active) {
    const double* X = xbuf(c, s.xin, b);
    const double* Bv = c.Bv + (int64_t)b * c.bv_stride;
    const int* perm = c.perm + (int64_t)b * c.n_max;
    const double* nbr = c.slab + (int64_t)b * c.tt_stride + c.off[s.nbr];
    const double scale_in = c.st[b].scale_in[s.slot];
    const int mt_n = (mprime + 15) >> 4;   // 16-row tiles in use
    // staging: thread = column l (coalesced over the warp), rows unrolled so that eight independent loads are
    // in flight per thread and no index needs a division (the one-shot load IS this kernel's critical path)
    for (int l = tid; l < ld; l += 128) {
      const bool lin = l < n;
      if (d.mode == 0) {
#pragma unroll 8
        for (int i = 0; i < 16 * mt_n; ++i) Ss[i * ld + l] = (lin && i < mprime && i <= l) ? X[i + p * l] * scale_in : 0.0;
      } else {
#pragma unroll 8
        for (int i = 0; i < 16 * mt_n; ++i) Ss[i * ld + l] = (lin && i < mprime) ? Bv[(int64_t)perm[i] * c.ldv + l] : 0.0;
      }
      if (s.dir == 1) {
#pragma unroll
        for (int jj = 0; jj < AS_TN; ++jj)
          Bs[jj * ld + l] = (lin && j0 + jj < N) ? nbr[l + (int64_t)n * (j0 + jj + (int64_t)N * x)] : 0.0;
      } else {
        const double* src = nbr + j0 + (int64_t)N * (l + (int64_t)n * x);   // AS_TN consecutive values
#pragma unroll
        for (int jj = 0; jj < AS_TN; ++jj) Bs[jj * ld + l] = (lin && j0 + jj < N) ? src[jj] : 0.0;
      }
    }
    __syncthreads();
    const int g = lane >> 2, t4 = lane & 3;
    double acc[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][e] = 0.0;
    for (int kk = 8 * wid; kk < n; kk += 32) {
      double bf[2];
      bf[0] = Bs[g * ld + kk + t4];
      bf[1] = Bs[g * ld + kk + t4 + 4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if (mt < mt_n) {
          double af[4];
          af[0] = Ss[(mt * 16 + g) * ld + kk + t4];
          af[1] = Ss[(mt * 16 + g + 8) * ld + kk + t4];
          af[2] = Ss[(mt * 16 + g) * ld + kk + t4 + 4];
          af[3] = Ss[(mt * 16 + g + 8) * ld + kk + t4 + 4];
          dmma_16x8x8(acc[mt], af, bf);
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int e = 0; e < 4; ++e) red[((wid * 2 + mt) * 32 + lane) * 4 + e] = acc[mt][e];
    __syncthreads();
    if (wid == 0) {
      double* Xn = xbuf(c, s.xin ^ 1, b);
      const int64_t ldo = (int64_t)mprime * c.Q;
      double mx = 0.0;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          double v = 0.0;
#pragma unroll
          for (int w2 = 0; w2 < 4; ++w2) v += red[((w2 * 2 + mt) * 32 + lane) * 4 + e];
          const int gi = mt * 16 + g + ((e & 2) ? 8 : 0);
          const int gj = j0 + t4 * 2 + (e & 1);
          if (gi < mprime && gj < N) {
            Xn[(gi + (int64_t)mprime * x) + ldo * gj] = v;
            mx = absmax_nan(mx, v);
          }
        }
      unsigned long long bits = (unsigned long long)__double_as_longlong(mx);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o);
        bits = other > bits ? other : bits;
      }
      if (lane == 0) atomicMax(&c.st[b].maxabs_bits[s.slot], bits);
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const int total = (int)(gridDim.x * gridDim.y) * c.Q;
    if (atomicAdd(&c.st[b].ticket, 1) == total - 1) {
      __threadfence();
      c.st[b].ticket = 0;
      commit_train(c, s, b, (int)gridDim.z / c.Q);
    }
  }
}

// Rescale of the absorb output (src/tensor_train.jl:169-173 / :208-212).  Between steps it is LAZY: the
// blocked QR is invariant under a uniform scale of X, so the division by max|D| is applied where the
// numbers are next consumed (k_jacobi's load, the R that k_absorb_mma reads) through StepState::scale_in,
// written by k_commit.  This kernel only runs on the last step of a sweep, where the scaled output is
// unfolded into the neighbour's site storage.
__global__ void k_rescale(SweepCtx c, StepDesc s) {
  pdl_enter();
  const int b = blockIdx.y;
  const int mprime = c.st[b].mprime[s.slot];
  const int N = c.st[b].N[s.slot];
  const double m = __longlong_as_double((long long)c.st[b].maxabs_bits[s.slot]);
  const bool ok = s.rescale && !(m != m) && !isinf(m) && m != 0.0;
  if (!ok && !s.last) return;
  const int64_t pn = (int64_t)mprime * c.Q;
  const int64_t tot = pn * N;
  double* Xn = xbuf(c, s.xin ^ 1, b);
  double* site = c.slab + (int64_t)b * c.tt_stride + c.off[s.nbr];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
    double v = Xn[i];
    if (ok) v /= m;
    if (!s.last) {
      Xn[i] = v;
    } else {
      // site layouts: right sweep site (N, m', Q): [j + N*(a + m'x)] = Xn[(a+m'x) + pn*j]
      //               left  sweep site (m', N, Q): [a + m'*(j + N*x)] = Xn[(a+m'x) + pn*j]
      const int64_t r = i % pn, j = i / pn;
      if (s.dir == 0) site[j + (int64_t)N * r] = v;
      else {
        const int a = (int)(r % mprime), x = (int)(r / mprime);
        site[a + (int64_t)mprime * (j + (int64_t)N * x)] = v;
      }
    }
  }
}

__global__ void k_reset_state(StepState* st, double* stat_err, int* stat_rank, int batch, int L1) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch) {
    st[i].mprime[0] = st[i].mprime[1] = 1; st[i].maxabs_bits[0] = 0ull; st[i].maxabs_bits[1] = 0ull;
    st[i].scale_in[0] = st[i].scale_in[1] = 1.0;
    st[i].ticket = 0;
  }
  if (i < batch * L1) { stat_err[i] = 0.0; stat_rank[i] = 0; }
}

// canonical-form residual max|A'A - I| (src/tensor_train.jl:263-273); one CTA
__global__ void k_canon_resid(const double* site, int dl, int dr, int Q, int left, double* out) {
  __shared__ double red[40];
  const int n = left ? dr : dl;
  double mx = 0.0;
  for (int o = threadIdx.x; o < n * n; o += blockDim.x) {
    const int i = o % n, j = o / n;
    double acc = 0.0;
    if (left) {
      for (int x = 0; x < Q; ++x)
        for (int k = 0; k < dl; ++k)
          acc += site[k + (int64_t)dl * (i + (int64_t)dr * x)] * site[k + (int64_t)dl * (j + (int64_t)dr * x)];
    } else {
      for (int x = 0; x < Q; ++x)
        for (int k = 0; k < dr; ++k)
          acc += site[i + (int64_t)dl * (k + (int64_t)dr * x)] * site[j + (int64_t)dl * (k + (int64_t)dr * x)];
    }
    mx = fmax(mx, fabs(acc - (i == j ? 1.0 : 0.0)));
  }
  mx = block_max(mx, red);
  if (threadIdx.x == 0) *out = mx;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static const int kMaxSmem = 208 * 1024;
static const int kCw = 4;
static constexpr int kPanelNS = TTB_PANEL_NS;  // panel columns per warp (2 -> 16 warps, 4 -> 8 warps)

int ensure_sweep_ws(ttb_tt* h) {
  Workspace& w = h->ws;
  if (w.sweep_ready) return TTB_OK;
  const int64_t p_max = h->dmax * h->Q, n_max = h->dmax;
  int bw = 32;
  while (bw > 4 && (p_max * (bw + kCw) + 3 * bw * bw + 128) * 8 > 200 * 1024) bw >>= 1;
  if ((p_max * (bw + kCw) + 3 * bw * bw + 128) * 8 > 200 * 1024)
    return fail(TTB_EUNSUPPORTED, "sweeps support bond*Q <= ~3000 (got %lld)", (long long)p_max);
  w.p_max = (int)p_max; w.n_max = (int)n_max; w.bw = bw;
  const int64_t B = h->batch;
  w.x_stride = (p_max * n_max + 15) & ~int64_t(15);
  w.t_stride = ((n_max + bw - 1) / bw) * bw * bw;
  w.ldv = (int)n_max; w.ldj = (int)n_max;
  w.bv_stride = n_max * n_max; w.jv_stride = n_max * n_max;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
  A((void**)&w.X[0], sizeof(double) * B * w.x_stride);
  A((void**)&w.X[1], sizeof(double) * B * w.x_stride);
  A((void**)&w.T, sizeof(double) * 2 * B * w.t_stride);  // two step slots
  A((void**)&w.Bv, sizeof(double) * B * w.bv_stride);
  A((void**)&w.Jv, sizeof(double) * B * w.jv_stride);
  A((void**)&w.perm, sizeof(int) * B * n_max);
  A((void**)&w.sigma, sizeof(double) * B * n_max);
  A((void**)&w.st, sizeof(StepState) * B);
  A((void**)&w.stat_err, sizeof(double) * B * (h->L + 1));
  A((void**)&w.stat_rank, sizeof(int) * B * (h->L + 1));
  if (e != cudaSuccess) return fail(TTB_ENOMEM, "sweep workspace: %s", cudaGetErrorString(e));
  TTB_ONCE_PER_DEVICE(h, 6) {
    cudaFuncSetAttribute(k_qr_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_qr_update, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_apply_q, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_qr_update_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_apply_q_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_absorb_skinny, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_qr_stream<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_qr_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_qr_stream<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_qr_stream<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_jacobi_cluster<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_jacobi_small<64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_jacobi_small<32, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_jacobi<512, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_jacobi<512, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_jacobi<1024, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_jacobi<1024, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_qr_panel_flag<2, kPanelNS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_qr_panel_flag<4, kPanelNS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_qr_panel_flag<8, kPanelNS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(k_qr_panel_flag<16, kPanelNS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  }
  w.sweep_ready = true;
  return TTB_OK;
}

static int ensure_spectrum(ttb_tt* h) {
  Workspace& w = h->ws;
  if (!h->record_spectrum) return TTB_OK;
  if (!w.spectrum) {
    const size_t n = (size_t)h->batch * (h->L + 1) * w.n_max;
    TTB_CUDA(cudaMalloc(&w.spectrum, sizeof(double) * n));
    // all bits set = a negative NaN: fails the `>= 0` scan of ttb_get_spectrum, i.e. "bond not touched"
    TTB_CUDA(cudaMemsetAsync(w.spectrum, 0xff, sizeof(double) * n, h->stream));
  }
  return TTB_OK;
}

static SweepCtx make_ctx(ttb_tt* h) {
  Workspace& w = h->ws;
  SweepCtx c;
  c.slab = h->slab; c.tt_stride = h->tt_stride; c.off = h->d_off; c.bond = h->d_bond; c.L1 = (int)h->L + 1;
  c.Q = (int)h->Q;
  c.X0 = w.X[0]; c.X1 = w.X[1]; c.x_stride = w.x_stride;
  c.T = w.T; c.t_stride = w.t_stride; c.bw = w.bw;
  c.Bv = w.Bv; c.Jv = w.Jv; c.bv_stride = w.bv_stride; c.jv_stride = w.jv_stride; c.ldv = w.ldv; c.ldj = w.ldj;
  c.perm = w.perm; c.sigma = w.sigma; c.n_max = w.n_max;
  c.st = w.st; c.logz = h->d_logz;
  c.stat_err = w.stat_err; c.stat_rank = w.stat_rank; c.spectrum = h->record_spectrum ? w.spectrum : nullptr;
  c.status = h->d_status;
  c.hint = h->hint_dev;
  return c;
}

struct HostStep { int site, nbr, rescale, bond_idx, bond_idx2, last; };

static int check_trunc(const ttb_trunc* tr) {
  if (!tr) return fail(TTB_EINVAL, "null svd_trunc");
  if (tr->kind < 0 || tr->kind > 3) return fail(TTB_EINVAL, "unknown svd_trunc kind %d", tr->kind);
  if (tr->kind != TTB_TRUNC_THRESH && tr->mprime < 1) return fail(TTB_EINVAL, "mprime must be >= 1");
  return TTB_OK;
}

// Enqueue one sweep (no host synchronisation inside).  dir 0 = right, 1 = left.
static int enqueue_sweep(ttb_tt* h, int dir, const ttb_trunc* tr, const std::vector<HostStep>& steps,
                         std::vector<int>& ub, std::vector<int>& lb) {
  TTB_TRY(ensure_sweep_ws(h));
  TTB_TRY(ensure_spectrum(h));
  Workspace& w = h->ws;
  const SweepCtx cbase = make_ctx(h);
  const int B = (int)h->batch, Q = (int)h->Q, bw = w.bw, L1 = (int)h->L + 1;
  cudaStream_t st = h->stream;
  // thin-Q formation runs on the side stream unless per-kernel profiling wants a serial timeline
  const bool overlap = !h->prof_on && getenv("TTB_NO_OVERLAP") == nullptr;
  cudaStream_t st2 = overlap ? h->stream2 : h->stream;
  bool side_pending = false;  // apply_q of the previous step may still be running on st2
  int side_slot = 0;
  const bool qr_only = (tr->kind == TTB_TRUNC_THRESH && tr->eps == 0.0);
  const bool feedback = !qr_only && getenv("TTB_NO_FEEDBACK") == nullptr;
  const bool use_stream = getenv("TTB_NO_STREAM") == nullptr;
  const int max_sweeps = getenv("TTB_JACOBI_MAXSWEEPS") ? std::max(1, atoi(getenv("TTB_JACOBI_MAXSWEEPS"))) : 60;
  int sm_count = 0;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
  // Main-stream launches use programmatic dependent launch (common.cuh): the next grid is resident and
  // parked in griddepcontrol.wait while its predecessor drains.  A launch that follows an event wait (or
  // opens the sweep) is an ordinary one.
  const bool use_pdl = !h->prof_on && getenv("TTB_NO_PDL") == nullptr;
  bool hard_dep = true;
#define SW_LAUNCH(name, kern, grid, block, smem, ...)                                              \
  do {                                                                                             \
    if (hard_dep || !use_pdl) { TTB_LAUNCH(h, name, (kern<<<grid, block, smem, st>>>(__VA_ARGS__))); } \
    else { TTB_LAUNCH(h, name, launch_pdl(kern, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)); } \
    hard_dep = false;                                                                              \
  } while (0)
  {
    const int n = std::max(B, B * L1);
    // per-bond stats and spectra describe the last CALL: the second sweep of orthogonalize_center! keeps the first one's
    SW_LAUNCH("k_reset_state", k_reset_state, (n + 255) / 256, 256, 0, w.st, w.stat_err, w.stat_rank, B, h->keep_stats ? 0 : L1);
    if (cbase.spectrum && !h->keep_stats)
      cudaMemsetAsync(w.spectrum, 0xff, sizeof(double) * (size_t)B * L1 * w.n_max, st);   // "bond not touched"
  }
  int xin = 0, slot = 0;
  for (size_t si = 0; si < steps.size(); ++si) {
    const HostStep& hs = steps[si];
    StepDesc s;
    s.site = hs.site; s.nbr = hs.nbr; s.dir = dir; s.xin = xin; s.rescale = hs.rescale; s.qr_only = qr_only ? 1 : 0;
    s.kind = tr->kind; s.eps = tr->eps;
    s.cap = (int)std::min<int64_t>(tr->mprime, 1 << 30);
    s.bond_idx = hs.bond_idx; s.bond_idx2 = hs.bond_idx2; s.last = hs.last; s.slot = slot;
    s.next_site = si + 1 < steps.size() ? steps[si + 1].site : -1;
    s.next_nbr = si + 1 < steps.size() ? steps[si + 1].nbr : -1;
    s.jc_sd = 0;
    s.max_sweeps = max_sweeps;
    s.trace = 0;
    s.hint_slot = feedback ? (int)(si % ttb_tt::kHintRing) : -1;
    SweepCtx c = cbase;
    c.T = cbase.T + (int64_t)slot * B * w.t_stride;
    // Launch-sizing feedback (never a blocking sync, and the rank DECISION never leaves the device): wait --
    // by polling an event -- until the step two before this one has committed, then tighten the bounds
    // with the ranks it reported.  Two steps of lead keep the GPU queue fed; bounds are at most two
    // doublings above the truth, so later grids shrink, whole QR phases are skipped when p < n is certain
    // and the Jacobi variant is picked for the real size.
    if (feedback && si >= 2) {
      const size_t done = si - 2;
      const int rs = (int)(done % ttb_tt::kHintRing);
      unsigned spins = 0;
      while (cudaEventQuery(h->ev_hint[rs]) == cudaErrorNotReady) {
        if (++spins > (1u << 30)) return fail(TTB_ECUDA, "rank feedback event never completed");
      }
      const volatile int* hv = h->hint_host + (size_t)rs * B;
      int v = 0, lo = 1 << 30;
      for (int bb = 0; bb < B; ++bb) { const int r = hv[bb]; v = std::max(v, r); lo = std::min(lo, r); }
      const HostStep& ds = steps[done];
      ub[ds.bond_idx] = v; lb[ds.bond_idx] = lo;
      if (ds.bond_idx2 >= 0) { ub[ds.bond_idx2] = v; lb[ds.bond_idx2] = lo; }
      // the step in between was sized from the old bound of this bond: its own bound can be tightened too
      const HostStep& ms = steps[done + 1];
      const int pm = (dir == 0 ? ub[ms.site + 1] : ub[ms.site]) * Q, nm = dir == 0 ? ub[ms.site] : ub[ms.site + 1];
      int capm = std::min(pm, nm);
      if (tr->kind != TTB_TRUNC_THRESH) capm = (int)std::min<int64_t>(capm, tr->mprime);
      if (capm < ub[ms.bond_idx]) { ub[ms.bond_idx] = capm; if (ms.bond_idx2 >= 0) ub[ms.bond_idx2] = capm; }
    }
    if (h->up_n) TTB_TRY(join_uploads(h, dir == 0 ? std::min(hs.site, hs.nbr) : -1));
    int p_ub, n_ub, N_ub, p_lb, n_lb;
    if (dir == 0) { p_ub = ub[hs.site + 1] * Q; n_ub = ub[hs.site]; N_ub = ub[hs.nbr]; p_lb = lb[hs.site + 1] * Q; n_lb = lb[hs.site]; }
    else { p_ub = ub[hs.site] * Q; n_ub = ub[hs.site + 1]; N_ub = ub[hs.nbr + 1]; p_lb = lb[hs.site] * Q; n_lb = lb[hs.site + 1]; }
    const int k_ub = std::min(p_ub, n_ub);
    {
      const int N_lb = dir == 0 ? lb[hs.nbr] : lb[hs.nbr + 1];
      const bool exact = p_ub == p_lb && n_ub == n_lb && N_ub == N_lb;
      s.kp = exact ? p_ub : -1; s.kn = exact ? n_ub : -1; s.kN = exact ? N_ub : -1;
    }
    if (si == 0) SW_LAUNCH("k_snapshot", k_snapshot, (B + 127) / 128, 128, 0, c, s, B);
    if (si == 0) {
      const int64_t tot = (int64_t)p_ub * n_ub;
      dim3 g((unsigned)std::min<int64_t>((tot + 255) / 256, 2048), B);
      SW_LAUNCH("k_fold", k_fold, g, 256, 0, c, s);
    }
    const bool maybe_qr = qr_only || !(p_ub < n_lb);  // p < n for every train => wide => Jacobi on X directly
    // register-panel / TMA kernels (they exchange V^T V + tau instead of an explicit T factor)
    const bool fast = (bw == 32 && p_ub <= 512);
    // One cooperative kernel for the whole site QR (qr_stream.cuh) when the dimensions are exact on the
    // host, the matrix is tall, there are at least two panels and all column-block CTAs can be co-resident.
    const int nblk = (n_ub + 31) / 32;
    const bool stream_qr = maybe_qr && fast && use_stream && s.kp >= 0 && p_ub >= n_ub && nblk >= 2 &&
                           nblk <= kQsMaxBlocks && (int64_t)nblk * B <= sm_count;
    if (stream_qr) {
      const int rpl = p_ub <= 64 ? 2 : (p_ub <= 128 ? 4 : (p_ub <= 256 ? 8 : 16));
      static const char* trace_env = getenv("TTB_QS_TRACE");
      s.trace = (trace_env && atoi(trace_env) == hs.site) ? 1 : 0;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(nblk, B); cfg.blockDim = dim3(kQsThreads); cfg.dynamicSmemBytes = qs_smem_bytes(rpl); cfg.stream = st;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = nblk; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = (use_pdl && !hard_dep) ? 2 : 1;
      if (rpl == 2) TTB_LAUNCH(h, "k_qr_stream", cudaLaunchKernelEx(&cfg, k_qr_stream<2>, c, s));
      else if (rpl == 4) TTB_LAUNCH(h, "k_qr_stream", cudaLaunchKernelEx(&cfg, k_qr_stream<4>, c, s));
      else if (rpl == 8) TTB_LAUNCH(h, "k_qr_stream", cudaLaunchKernelEx(&cfg, k_qr_stream<8>, c, s));
      else TTB_LAUNCH(h, "k_qr_stream", cudaLaunchKernelEx(&cfg, k_qr_stream<16>, c, s));
      hard_dep = false;
    } else if (maybe_qr) {
      const int np = (k_ub + bw - 1) / bw;
      for (int panel = 0; panel < np; ++panel) {
        const int j0 = panel * bw;
        const int rows_ub = p_ub - j0;
        if (fast) {
#define TTB_PANEL(RPL_)                                                                            \
  do {                                                                                             \
    auto kp = k_qr_panel_flag<RPL_, kPanelNS>;                                                     \
    SW_LAUNCH("k_qr_panel_flag", kp, B, 1024 / kPanelNS, sizeof(double) * (1024 * RPL_ + 1088) + 256, c, s, panel); \
  } while (0)
          if (rows_ub <= 64) TTB_PANEL(2);
          else if (rows_ub <= 128) TTB_PANEL(4);
          else if (rows_ub <= 256) TTB_PANEL(8);
          else TTB_PANEL(16);
#undef TTB_PANEL
        } else {
          const int thr = std::min(512, std::max(64, ((rows_ub + 31) / 32) * 32));
          size_t smem = sizeof(double) * ((size_t)2 * bw * bw + 2 * bw + 40 + (size_t)rows_ub * bw);
          SW_LAUNCH("k_qr_panel", k_qr_panel, B, thr, smem, c, s, panel);
        }
        // columns behind the panel: when every train is certainly tall (k = n) the panel covers min(32, n - j0)
        // columns and a single-panel step has nothing to update (it used to launch a grid of early exits)
        const int trail = (p_lb >= n_ub) ? n_ub - j0 - bw : n_ub - j0 - 1;
        if (trail > 0) {
          dim3 g((trail + kCw - 1) / kCw, B);
          if (fast) {
            size_t sm2 = sizeof(double) * ((size_t)1280 + (size_t)rows_ub * 36);
            SW_LAUNCH("k_qr_update_tma", k_qr_update_tma, g, 256, sm2, c, s, panel);
          } else {
            size_t sm2 = sizeof(double) * ((size_t)bw * bw + 2 * bw * kCw + (size_t)rows_ub * (bw + kCw));
            SW_LAUNCH("k_qr_update", k_qr_update, g, 256, sm2, c, s, panel, kCw);
          }
        }
      }
    }
    int mp_ub = k_ub;
    if (!qr_only) {
      if (tr->kind != TTB_TRUNC_THRESH) mp_ub = (int)std::min<int64_t>(mp_ub, tr->mprime);
      const int nv_ub = k_ub, len_ub = n_ub;
      const size_t need = sizeof(double) * (2 * (size_t)w.n_max + (size_t)nv_ub * (len_ub + nv_ub));
      const size_t give = std::min<size_t>(need, (size_t)200 * 1024);
      const int thr = std::min(1024, std::max(64, ((nv_ub + 1) / 2) * 32));  // one warp per pair
      if (side_pending && overlap) { cudaStreamWaitEvent(st, h->ev_aq[side_slot], 0); side_pending = false; hard_dep = true; }
      const int sd = (int)(give / sizeof(double));
      // too large for one CTA's shared memory: cluster of 8 CTAs (jacobi_cluster.cuh)
      const bool cluster_path = need > (size_t)200 * 1024 && jc_smem_bytes(nv_ub, len_ub) <= (size_t)200 * 1024 &&
                                getenv("TTB_JACOBI_NOCLUSTER") == nullptr;
      if (cluster_path) {
        // the bounds say "may not fit one CTA"; the device knows: launch both, each train takes one path
        s.jc_sd = sd;
        auto kj1 = k_jacobi<1024, 8>;
        SW_LAUNCH("k_jacobi", kj1, B, 1024, give, c, s, sd);
        const int tiles = ((nv_ub + 31) / 32) * ((len_ub + 31) / 32);
        SW_LAUNCH("k_jacobi_init", k_jacobi_init, dim3(std::min(tiles, 128), B), 256, 0, c, s);
        TTB_LAUNCH(h, "k_jacobi_cluster",
                   (k_jacobi_cluster<8><<<dim3(kJcCluster, B), kJcThreads, jc_smem_bytes(nv_ub, len_ub), st>>>(c, s)));
        SW_LAUNCH("k_jacobi_finish", k_jacobi_finish, B, 1024, sizeof(double) * 2 * (size_t)w.n_max, c, s);
      } else if (nv_ub <= 64 && len_ub <= 64 && getenv("TTB_JACOBI_GENERIC") == nullptr) {
        // many warps per SM: the issue-lean fixed-stride kernel (jacobi_small.cuh)
        static const bool no_g8 = getenv("TTB_JACOBI_NOG8") != nullptr;
        TTB_ONCE_PER_DEVICE(h, 7) cudaFuncSetAttribute(k_jacobi_g8<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jg8_smem(64));
        static const bool use_sys = getenv("TTB_JACOBI_SYS") != nullptr;
        TTB_ONCE_PER_DEVICE(h, 11) {
          cudaFuncSetAttribute(k_jacobi_sys<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jsys_smem(64));
          cudaFuncSetAttribute(k_jacobi_sys<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jsys_smem(32));
        }
        if (use_sys && nv_ub <= 32 && len_ub <= 32 && nv_ub > 16) {
          auto kg = k_jacobi_sys<32>;
          SW_LAUNCH("k_jacobi", kg, B, 128, jsys_smem(32), c, s);
        } else if (use_sys && !(nv_ub <= 32 && len_ub <= 32) && nv_ub > 32) {
          auto kg = k_jacobi_sys<64>;
          SW_LAUNCH("k_jacobi", kg, B, 256, jsys_smem(64), c, s);
        } else
        if (nv_ub <= 32 && len_ub <= 32 && B >= sm_count && !no_g8) {
          // a batch that fills the GPU is issue-bound: four pairs per warp (k_jacobi_g8)
          auto kg = k_jacobi_g8<32>;
          SW_LAUNCH("k_jacobi", kg, B, 128, jg8_smem(32), c, s);
        } else if (!(nv_ub <= 32 && len_ub <= 32) && !no_g8) {
          // k, n <= 64: 32 pair-warps of one train crowd one SM (issue-bound as well): 8 warps, four pairs each
          auto kg = k_jacobi_g8<64>;
          SW_LAUNCH("k_jacobi", kg, B, 256, jg8_smem(64), c, s);
        } else if (nv_ub <= 32 && len_ub <= 32) {
          auto ks = k_jacobi_small<32, 32>;
          // TTB_JACOBI_ILP=1: half the warps with two interleaved pairs each (jacobi_two_pairs).  Measured on the
          // 4096-train batch of config 5: 34.4 ms either way -- the round is bound by instruction issue, not by
          // the latency of one warp's chain -- so it stays off by default.
          static const bool ilp = getenv("TTB_JACOBI_ILP") != nullptr;
          const int thr32 = (ilp && B >= 2 * sm_count && thr > 256) ? 256 : std::min(512, thr);
          SW_LAUNCH("k_jacobi", ks, B, thr32, sizeof(double) * (2 * 32 + 2 * 32 * 32), c, s);
        } else {
          auto ks = k_jacobi_small<64, 64>;
          SW_LAUNCH("k_jacobi", ks, B, thr, sizeof(double) * (2 * 64 + 2 * 64 * 64), c, s);
        }
      } else if (nv_ub <= 32 && len_ub <= 256 && len_ub > 32 && getenv("TTB_JACOBI_GENERIC") == nullptr && getenv("TTB_JACOBI_LQ") != nullptr) {
        // (opt-in, measured slower) wide folds: Householder LQ first, Jacobi on the small square factor (jacobi_lq.cuh)
        TTB_ONCE_PER_DEVICE(h, 12) cudaFuncSetAttribute(k_jacobi_lq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jlq_smem());
        SW_LAUNCH("k_jacobi", k_jacobi_lq, B, 256, jlq_smem(), c, s);
      } else if (nv_ub <= 32 && len_ub <= 256 && getenv("TTB_JACOBI_GENERIC") == nullptr) {
        auto ks = k_jacobi_small<32, 256>;
        SW_LAUNCH("k_jacobi", ks, B, std::min(512, thr), sizeof(double) * (2 * 32 + 32 * 256 + 32 * 32), c, s);
      } else {
      auto kj = thr <= 512 ? (len_ub <= 64 ? k_jacobi<512, 2> : k_jacobi<512, 8>)
                           : (len_ub <= 64 ? k_jacobi<1024, 2> : k_jacobi<1024, 8>);
      SW_LAUNCH("k_jacobi", kj, B, thr, give, c, s, sd);
      }
    }
    {
      dim3 g((mp_ub + kCw - 1) / kCw, B);
      // The new site tensor W = Q E (or Q J[:,sel]) is consumed by nobody in this sweep: form it on the
      // side stream while the main stream absorbs S and factorises the next site.  It reads X[xin], the
      // T slot and Bv/Jv/perm of this step; the next step's absorb (writes X[xin]) / Jacobi wait for it.
      if (overlap) {
        cudaEventRecord(h->ev_qr[slot], st);
        cudaStreamWaitEvent(st2, h->ev_qr[slot], 0);
      }
      if (fast && p_ub <= 64 && k_ub <= 32 && getenv("TTB_APPLYQ_NOSMALL") == nullptr) {
        // small step: one CTA forms the whole orthonormal factor of its train (tensor cores)
        TTB_ONCE_PER_DEVICE(h, 8) cudaFuncSetAttribute(k_apply_q_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kApplySmallSmem);
        TTB_LAUNCH(h, "k_apply_q_small", k_apply_q_small<<<B, 256, kApplySmallSmem, st2>>>(c, s));
      } else if (fast) {
        size_t smem = sizeof(double) * ((size_t)1280 + (size_t)p_ub * 36);
        TTB_LAUNCH(h, "k_apply_q_tma", k_apply_q_tma<<<g, 256, smem, st2>>>(c, s));
      } else {
        size_t smem = sizeof(double) * ((size_t)bw * bw + 2 * bw * kCw + (size_t)p_ub * kCw + (size_t)p_ub * bw);
        TTB_LAUNCH(h, "k_apply_q", k_apply_q<<<g, 256, smem, st2>>>(c, s, kCw));
      }
      if (overlap) {
        // absorb of THIS step only reads; the wait for the PREVIOUS step's apply_q must come before it
        // because absorb writes X[xin ^ 1], which that apply_q read as its X[xin]
        if (side_pending) { cudaStreamWaitEvent(st, h->ev_aq[side_slot], 0); side_pending = false; hard_dep = true; }
        cudaEventRecord(h->ev_aq[slot], st2);
        side_pending = true;
        side_slot = slot;
      }
    }
    {
      const size_t sk_smem = sizeof(double) * ((size_t)(32 + AS_TN) * as_ld(n_ub) + 4 * 2 * 32 * 4);
      // (a batch that already fills the GPU is better off with the small-footprint tiled kernel)
      if (mp_ub <= 32 && n_ub <= 32 && N_ub <= 32 && (int64_t)B * Q >= 64 && getenv("TTB_ABSORB_NOSMALL") == nullptr) {
        SW_LAUNCH("k_absorb_small", k_absorb_small, B, 128, 0, c, s);
      } else if (mp_ub <= 32 && sk_smem <= (size_t)200 * 1024 && (int64_t)B * Q * ((N_ub + AS_TN - 1) / AS_TN) <= 2 * (int64_t)sm_count &&
          getenv("TTB_ABSORB_TILED") == nullptr) {
        dim3 g(1, (N_ub + AS_TN - 1) / AS_TN, Q * B);
        SW_LAUNCH("k_absorb_mma", k_absorb_skinny, g, 128, sk_smem, c, s);
      } else if (qr_only && s.kp >= 0 && p_ub >= n_ub && n_ub % 32 == 0 && N_ub % 32 == 0 && (p_ub & 1) == 0 &&
                 getenv("TTB_ABSORB_NOPIPE") == nullptr) {
        // exact, 32-aligned eps = 0 step: cp.async pipeline (k_absorb_pipe)
        const size_t psm = sizeof(double) * AP_STAGES * 2 * 32 * AM_LD;
        TTB_ONCE_PER_DEVICE(h, 9) cudaFuncSetAttribute(k_absorb_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
        dim3 g(n_ub / AM_T, N_ub / AM_T, Q * B);
        SW_LAUNCH("k_absorb_mma", k_absorb_pipe, g, 128, psm, c, s);
      } else {
        dim3 g((mp_ub + AM_T - 1) / AM_T, (N_ub + AM_T - 1) / AM_T, Q * B);
        SW_LAUNCH("k_absorb_mma", k_absorb_mma, g, 128, 0, c, s);
      }
    }
    {
      const int64_t tot = (int64_t)mp_ub * Q * N_ub;
      dim3 g((unsigned)std::min<int64_t>((tot + 1023) / 1024, 1024), B);
      if (hs.last) SW_LAUNCH("k_rescale", k_rescale, g, 256, 0, c, s);
      if (feedback) cudaEventRecord(h->ev_hint[s.hint_slot], st);
    }
    const int mp_lb = qr_only ? std::min(p_lb, n_lb) : 1;
    ub[hs.bond_idx] = mp_ub; lb[hs.bond_idx] = mp_lb;
    if (hs.bond_idx2 >= 0) { ub[hs.bond_idx2] = mp_ub; lb[hs.bond_idx2] = mp_lb; }
    xin ^= 1; slot ^= 1;
  }
#undef SW_LAUNCH
  if (side_pending && overlap) cudaStreamWaitEvent(st, h->ev_aq[side_slot], 0);  // join the side stream
  if (h->up_n) TTB_TRY(join_uploads(h, -1));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TTB_ECUDA, "sweep launch: %s", cudaGetErrorString(e));
  return TTB_OK;
}

static void bond_upper_bounds(ttb_tt* h, std::vector<int>& ub, std::vector<int>& lb) {
  const int L1 = (int)h->L + 1;
  ub.assign(L1, 0);
  lb.assign(L1, 1 << 30);
  for (int64_t b = 0; b < h->batch; ++b)
    for (int t = 0; t < L1; ++t) {
      const int v = h->h_bond[b * L1 + t];
      ub[t] = std::max(ub[t], v);
      lb[t] = std::min(lb[t], v);
    }
}

static int build_right(ttb_tt* h, int64_t lo, int64_t hi, std::vector<HostStep>& steps) {
  const int L = (int)h->L;
  steps.clear();
  if (h->periodic) {
    if (L < 2) return fail(TTB_EUNSUPPORTED, "periodic sweeps need L >= 2");
    steps.push_back({0, L - 1, 0, 0, L, 0});                       // wrap step, no rescale (:84-92)
    for (int t = L - 1; t >= 1; --t) steps.push_back({t, t - 1, 1, t, -1, t == 1 ? 1 : 0});
    return TTB_OK;
  }
  if (lo == 0 && hi == 0) { lo = 2; hi = L; }
  if (lo > hi) return TTB_OK;  // isempty(indices)
  if (lo < 2 || hi > L) return fail(TTB_EINVAL, "Indices not compatible with tensor train positions 1:%d: got %lld:%lld", L, (long long)lo, (long long)hi);
  if (hi != L) return fail(TTB_EUNSUPPORTED, "orthogonalize_right! always seeds from the last site (tensor_train.jl:157); ranges must end at L");
  for (int t = (int)hi; t >= (int)lo; --t) steps.push_back({t - 1, t - 2, 1, t - 1, -1, t == (int)lo ? 1 : 0});
  return TTB_OK;
}

static int build_left(ttb_tt* h, int64_t lo, int64_t hi, std::vector<HostStep>& steps) {
  const int L = (int)h->L;
  steps.clear();
  if (h->periodic) {
    if (L < 2) return fail(TTB_EUNSUPPORTED, "periodic sweeps need L >= 2");
    for (int t = 0; t <= L - 2; ++t) steps.push_back({t, t + 1, 1, t + 1, -1, 0});
    steps.push_back({L - 1, 0, 0, L, 0, 1});                       // wrap step, no rescale (:133-138)
    return TTB_OK;
  }
  if (lo == 0 && hi == 0) { lo = 1; hi = L - 1; }
  if (lo > hi) return TTB_OK;
  if (lo < 1 || hi > L - 1) return fail(TTB_EINVAL, "Indices not compatible with tensor train positions 1:%d: got %lld:%lld", L, (long long)lo, (long long)hi);
  if (lo != 1) return fail(TTB_EUNSUPPORTED, "orthogonalize_left! always seeds from the first site (tensor_train.jl:196); ranges must start at 1");
  for (int t = (int)lo; t <= (int)hi; ++t) steps.push_back({t - 1, t, 1, t, -1, t == (int)hi ? 1 : 0});
  return TTB_OK;
}

static int finish_sweeps(ttb_tt* h, CallTimer& tm) {
  TTB_TRY(tm.finish());
  TTB_TRY(sync_bonds_to_host(h));
  int st = 0;
  TTB_TRY(read_status(h, &st));
  if (st & ST_NONFINITE) return fail(TTB_ENONFINITE, "non-finite values met during the sweep");
  // the reference's LAPACK svd throws when it does not converge; a non-converged Jacobi step has wrong singular
  // values, a wrong rank decision and non-orthogonal factors, so the call must not report success
  if (st & ST_NOCONV) return fail(TTB_ENOCONV, "one-sided Jacobi SVD did not converge within the sweep cap");
  return TTB_OK;
}

}  // namespace ttb

using namespace ttb;

extern "C" {

int ttb_orthogonalize_right(ttb_handle h, const ttb_trunc* tr, int64_t lo, int64_t hi) {
  TTB_TRY(check_handle(h));
  TTB_TRY(check_trunc(tr));
  std::vector<HostStep> steps;
  TTB_TRY(build_right(h, lo, hi, steps));
  if (steps.empty()) return TTB_OK;
  std::vector<int> ub, lb;
  bond_upper_bounds(h, ub, lb);
  TTB_TRY(ensure_sweep_ws(h));  // allocate outside the timed bracket
  CallTimer tm(h);
  TTB_TRY(enqueue_sweep(h, 0, tr, steps, ub, lb));
  return finish_sweeps(h, tm);
}

int ttb_orthogonalize_left(ttb_handle h, const ttb_trunc* tr, int64_t lo, int64_t hi) {
  TTB_TRY(check_handle(h));
  TTB_TRY(check_trunc(tr));
  std::vector<HostStep> steps;
  TTB_TRY(build_left(h, lo, hi, steps));
  if (steps.empty()) return TTB_OK;
  std::vector<int> ub, lb;
  bond_upper_bounds(h, ub, lb);
  TTB_TRY(ensure_sweep_ws(h));  // allocate outside the timed bracket
  CallTimer tm(h);
  TTB_TRY(enqueue_sweep(h, 1, tr, steps, ub, lb));
  return finish_sweeps(h, tm);
}

int ttb_orthogonalize_center(ttb_handle h, const ttb_trunc* tr, int64_t l) {
  TTB_TRY(check_handle(h));
  if (h->periodic) return fail(TTB_EINVAL, "orthogonalize_center! is defined for TensorTrain only");
  if (l < 1 || l > h->L) return fail(TTB_EINVAL, "center %lld out of range", (long long)l);
  if (l - 1 >= 1) TTB_TRY(ttb_orthogonalize_left(h, tr, 1, l - 1));
  int rc = TTB_OK;
  if (l + 1 <= h->L) {
    h->keep_stats = (l - 1 >= 1);   // svd_trunc.jl:98-100 updates maxerr on every SVD of BOTH sweeps
    rc = ttb_orthogonalize_right(h, tr, l + 1, h->L);
    h->keep_stats = false;
  }
  return rc;
}

int ttb_orthogonalize_two_site_center(ttb_handle h, const ttb_trunc* tr, int64_t k) {
  TTB_TRY(check_handle(h));
  if (h->periodic) return fail(TTB_EINVAL, "orthogonalize_two_site_center! is defined for TensorTrain only");
  if (k < 1 || k >= h->L) return fail(TTB_EINVAL, "k must be between 1 and length(C)-1 for two-site update");
  if (k - 1 >= 1) TTB_TRY(ttb_orthogonalize_left(h, tr, 1, k - 1));
  int rc = TTB_OK;
  if (k + 2 <= h->L) {
    h->keep_stats = (k - 1 >= 1);
    rc = ttb_orthogonalize_right(h, tr, k + 2, h->L);
    h->keep_stats = false;
  }
  return rc;
}

int ttb_compress(ttb_handle h, const ttb_trunc* tr, int is_orthogonal) {
  TTB_TRY(check_handle_raw(h));   // a pending asynchronous upload is consumed chunk by chunk by the right sweep
  if (is_orthogonal != TTB_ORTH_NONE && h->up_n) TTB_TRY(join_uploads(h, -1));
  TTB_TRY(check_trunc(tr));
  if (is_orthogonal == TTB_ORTH_LEFT) return ttb_orthogonalize_right(h, tr, 0, 0);
  if (is_orthogonal == TTB_ORTH_RIGHT) return ttb_orthogonalize_left(h, tr, 0, 0);
  if (is_orthogonal != TTB_ORTH_NONE)
    return fail(TTB_EINVAL, "Keyword `is_orthogonal` only supports: :none, :left, :right");
  // both sweeps are enqueued back to back; the left sweep sizes its launches from the bond
  // dimensions the eps=0 right sweep is known to produce
  std::vector<HostStep> rs, ls;
  TTB_TRY(build_right(h, 0, 0, rs));
  TTB_TRY(build_left(h, 0, 0, ls));
  std::vector<int> ub, lb;
  bond_upper_bounds(h, ub, lb);
  ttb_trunc t0;
  t0.kind = TTB_TRUNC_THRESH; t0._pad = 0; t0.eps = 0.0; t0.mprime = 0;
  TTB_TRY(ensure_sweep_ws(h));  // allocate outside the timed bracket
  CallTimer tm(h);
  const bool dbg = getenv("TTB_TIMING") != nullptr;
  timespec t_a, t_b, t_c, t_d;
  if (dbg) clock_gettime(CLOCK_MONOTONIC, &t_a);
  if (!rs.empty()) TTB_TRY(enqueue_sweep(h, 0, &t0, rs, ub, lb));
  cudaEvent_t ev_mid = nullptr;
  if (dbg) { clock_gettime(CLOCK_MONOTONIC, &t_b); cudaEventCreate(&ev_mid); cudaEventRecord(ev_mid, h->stream); }
  if (!ls.empty()) TTB_TRY(enqueue_sweep(h, 1, tr, ls, ub, lb));
  if (dbg) clock_gettime(CLOCK_MONOTONIC, &t_c);
  const int rc = finish_sweeps(h, tm);
#ifdef TTB_JC_STAMPS
  if (getenv("TTB_JC_TRACE")) {
    long long js[8];
    cudaMemcpyFromSymbol(js, g_jc_stamps, sizeof(js));
    fprintf(stderr, "[jc trace] last cluster SVD, CTA 0 cycles: load %lld norms %lld in-block %lld cross %lld store+fence %lld cluster-sync %lld | sweeps %lld block-rounds %lld\n",
            js[0], js[1], js[2], js[3], js[4], js[5], js[6], js[7]);
  }
#endif
  if (getenv("TTB_QS_TRACE")) {
    unsigned long long tr_h[16 * 32];
    cudaMemcpyFromSymbol(tr_h, g_qs_trace, sizeof(tr_h));
    const unsigned long long t0 = tr_h[0];
    for (int kb = 0; kb < 8; ++kb) {
      fprintf(stderr, "[qs trace] CTA %d: start %+6.1f us | consumed panels:", kb, ((double)tr_h[kb * 32] - (double)t0) * 1e-3);
      for (int j = 0; j < kb; ++j) fprintf(stderr, " %6.1f", ((double)tr_h[kb * 32 + 1 + j] - (double)t0) * 1e-3);
      fprintf(stderr, " | factor done %6.1f | stored %6.1f | blocks done:", ((double)tr_h[kb * 32 + 20] - (double)t0) * 1e-3,
              ((double)tr_h[kb * 32 + 21] - (double)t0) * 1e-3);
      for (int blk = 0; blk < 8; ++blk) fprintf(stderr, " %5.1f", ((double)tr_h[kb * 32 + 22 + blk] - (double)t0) * 1e-3);
      fprintf(stderr, "\n");
    }
    unsigned long long t2[4 * 64];
    cudaMemcpyFromSymbol(t2, g_qs_trace2, sizeof(t2));
    fprintf(stderr, "[qs trace] panel 0 -> CTA 1, per reflector: ready seen by pusher | pushed | arrived | applied (us)\n");
    for (int g = 0; g < 32; ++g)
      fprintf(stderr, "[qs trace]   %2d: %6.2f %6.2f %6.2f %6.2f\n", g, ((double)t2[64 + g] - (double)t0) * 1e-3, ((double)t2[g] - (double)t0) * 1e-3,
              ((double)t2[128 + g] - (double)t0) * 1e-3, ((double)t2[192 + g] - (double)t0) * 1e-3);
  }
  if (dbg) {
    clock_gettime(CLOCK_MONOTONIC, &t_d);
    auto ms = [](const timespec& x, const timespec& y) { return (y.tv_sec - x.tv_sec) * 1e3 + (y.tv_nsec - x.tv_nsec) * 1e-6; };
    float g_right = 0.f, g_left = 0.f;
    cudaEventElapsedTime(&g_right, h->ev0, ev_mid);
    cudaEventElapsedTime(&g_left, ev_mid, h->ev1);
    cudaEventDestroy(ev_mid);
    fprintf(stderr, "[ttb timing] compress: host enqueue right %.1f ms, left (incl. rank-hint syncs) %.1f ms, drain %.1f ms; "
                    "device %.1f ms (right sweep %.1f, left sweep %.1f), %lld launches\n", ms(t_a, t_b), ms(t_b, t_c),
            ms(t_c, t_d), h->last_ms, g_right, g_left, (long long)h->last_launches);
  }
  return rc;
}

int ttb_sweep_stats(ttb_handle h, int64_t b, int64_t* rank, double* err) {
  TTB_TRY(check_handle(h));
  if (b < 0 || b >= h->batch) return fail(TTB_EINVAL, "train index out of range");
  const int L1 = (int)h->L + 1;
  if (!h->ws.sweep_ready) {
    for (int i = 0; i < L1; ++i) { if (rank) rank[i] = 0; if (err) err[i] = 0.0; }
    return TTB_OK;
  }
  std::vector<int> r(L1);
  std::vector<double> e(L1);
  TTB_CUDA(cudaMemcpyAsync(r.data(), h->ws.stat_rank + b * L1, sizeof(int) * L1, cudaMemcpyDeviceToHost, h->stream));
  TTB_CUDA(cudaMemcpyAsync(e.data(), h->ws.stat_err + b * L1, sizeof(double) * L1, cudaMemcpyDeviceToHost, h->stream));
  TTB_CUDA(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < L1; ++i) { if (rank) rank[i] = r[i]; if (err) err[i] = e[i]; }
  return TTB_OK;
}

int ttb_get_spectrum(ttb_handle h, int64_t b, int64_t bond_index, double* out, int64_t cap, int64_t* n) {
  TTB_TRY(check_handle(h));
  if (!h->ws.spectrum) return fail(TTB_EINVAL, "spectrum recording is off (ttb_record_spectrum)");
  if (b < 0 || b >= h->batch || bond_index < 0 || bond_index > h->L) return fail(TTB_EINVAL, "index out of range");
  const int nm = h->ws.n_max;
  std::vector<double> tmp(nm);
  TTB_CUDA(cudaMemcpyAsync(tmp.data(), h->ws.spectrum + (b * (h->L + 1) + bond_index) * nm, sizeof(double) * nm,
                           cudaMemcpyDeviceToHost, h->stream));
  TTB_CUDA(cudaStreamSynchronize(h->stream));
  int64_t cnt = 0;
  while (cnt < nm && tmp[cnt] >= 0.0) ++cnt;
  *n = cnt;   // the caller learns the length even when `out` is too small
  if (cnt > cap) return fail(TTB_EINVAL, "spectrum has %lld values, out holds %lld", (long long)cnt, (long long)cap);
  for (int64_t i = 0; i < cnt; ++i) out[i] = tmp[i];
  return TTB_OK;
}

int ttb_canonical_residual(ttb_handle h, int64_t b, int64_t t, int left, double* resid) {
  TTB_TRY(check_handle(h));
  if (b < 0 || b >= h->batch || t < 0 || t >= h->L) return fail(TTB_EINVAL, "index out of range");
  const int* bd = &h->h_bond[b * (h->L + 1)];
  double* d_out = nullptr;
  TTB_CUDA(cudaMalloc(&d_out, sizeof(double)));
  TTB_LAUNCH(h, "k_canon_resid", k_canon_resid<<<1, 256, 0, h->stream>>>(h->slab + b * h->tt_stride + h->off[t], bd[t], bd[t + 1], (int)h->Q, left,
                                          d_out));
  cudaError_t e = cudaMemcpyAsync(resid, d_out, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_out);
  if (e != cudaSuccess) return fail(TTB_ECUDA, "canonical_residual: %s", cudaGetErrorString(e));
  return TTB_OK;
}

}  // extern "C"
