// env_fast.cuh -- open-chain environment pass, phase 2 of the two-phase scheme (see k_site_trace).
//
// accumulate_L / accumulate_R of an OPEN train are chains of vector-matrix products
// (src/abstract_tensor_train.jl:89-117): v <- (v / max|v|) * T_t.  Phase 1 has already summed the
// physical index (T_t = sum_x A_t(x); written transposed for the right environments), so both directions
// are the same kernel: v_new[j] = sum_i v[i] M[i + rows*j] with M column-major, i.e. coalesced column
// dots.  One CTA walks the chain; each warp owns columns j = w, w+32, ... and issues all loads of two
// columns before reducing, the running vector lives in shared memory, and the reference's scaling
// convention (stored vectors max-abs 1 except the last, log z accumulated) is applied on the fly.
#pragma once

namespace ttb {

struct ChainCtx {
  const double* T; int64_t t_stride; const int64_t* toff;   // traces (left) or transposed traces (right)
  const int* bond; int L, L1;
  double* env; int64_t env_stride; const int64_t* eoff;
  double* log_out; int* sign_out;
  int normalize; int left; int dmax;
};

// transposed physical-index sum: Tt[j + dr*i] = sum_x A[i + dl*(j + dr*x)]   (32 x 32 tiles via smem)
__global__ void k_site_trace_t(const double* slab, int64_t tt_stride, const int64_t* off, const int* bond, int L1, int Q,
                               double* tr, int64_t tr_stride, const int64_t* toff) {
  const int t = blockIdx.y, b = blockIdx.z;
  const int* bd = bond + (int64_t)b * L1;
  const int dl = bd[t], dr = bd[t + 1];
  const double* A = slab + (int64_t)b * tt_stride + off[t];
  double* o = tr + (int64_t)b * tr_stride + toff[t];
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int tiles_i = (dl + 31) / 32, tiles_j = (dr + 31) / 32;
  for (int tl = blockIdx.x; tl < tiles_i * tiles_j; tl += gridDim.x) {
    const int i0 = (tl % tiles_i) * 32, j0 = (tl / tiles_i) * 32;
    for (int jj = ty; jj < 32; jj += 8) {
      const int i = i0 + tx, j = j0 + jj;
      double acc = 0.0;
      if (i < dl && j < dr)
        for (int x = 0; x < Q; ++x) acc += A[i + (int64_t)dl * (j + (int64_t)dr * x)];
      tile[jj][tx] = acc;
    }
    __syncthreads();
    for (int ii = ty; ii < 32; ii += 8) {
      const int j = j0 + tx, i = i0 + ii;
      if (i < dl && j < dr) o[j + (int64_t)dr * i] = tile[tx][ii];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) k_chain_open(ChainCtx c) {
  const int b = blockIdx.x;
  const int* bond = c.bond + (int64_t)b * c.L1;
  const double* T = c.T + (int64_t)b * c.t_stride;
  double* env = c.env + (int64_t)b * c.env_stride;
  extern __shared__ double sm[];
  double* v = sm;              // dmax
  double* vn = sm + c.dmax;    // dmax
  __shared__ double red[40];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;  // 32 warps
  if (tid == 0) v[0] = 1.0;
  double logz = 0.0;
  double nt = 1.0;  // max-abs of the current vector (identity start)
  __syncthreads();
  for (int st = 0; st < c.L; ++st) {
    const int t = c.left ? st : c.L - 1 - st;
    const int rows = c.left ? bond[t] : bond[t + 1];      // length of v
    const int cols = c.left ? bond[t + 1] : bond[t];      // length of v_new
    const double* M = T + c.toff[t];
    // reference quirk (:93-97): the previous vector -- already stored -- is divided by its max-abs in place
    if (st > 0 && nt != 0.0 && c.normalize) {
      double* prev = env + c.eoff[c.left ? t - 1 : t + 1];
      for (int i = tid; i < rows; i += 1024) { const double x = v[i] / nt; v[i] = x; prev[i] = x; }
      logz += log(nt);
      __syncthreads();
    }
    for (int j0 = wid; j0 < cols; j0 += 64) {
      const int ja = j0, jb = j0 + 32;
      const bool hb = jb < cols;
      double xa[8], xb[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int i = lane + 32 * m;
        xa[m] = i < rows ? M[i + (int64_t)rows * ja] : 0.0;
        xb[m] = (hb && i < rows) ? M[i + (int64_t)rows * jb] : 0.0;
      }
      double sa = 0.0, sb = 0.0;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int i = lane + 32 * m;
        const double vi = i < rows ? v[i] : 0.0;
        sa += vi * xa[m];
        sb += vi * xb[m];
      }
      for (int i = lane + 256; i < rows; i += 32) {  // bonds beyond 256
        const double vi = v[i];
        sa += vi * M[i + (int64_t)rows * ja];
        if (hb) sb += vi * M[i + (int64_t)rows * jb];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
      }
      if (lane == 0) { vn[ja] = sa; if (hb) vn[jb] = sb; }
    }
    __syncthreads();
    // max-abs of the new vector (NaN-propagating), store it, swap
    double mx = 0.0;
    double* cur = env + c.eoff[t];
    for (int j = tid; j < cols; j += 1024) { const double x = vn[j]; mx = absmax_nan(mx, x); cur[j] = x; }
    unsigned long long bits = (unsigned long long)__double_as_longlong(mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ot = __shfl_xor_sync(0xffffffffu, bits, o);
      bits = ot > bits ? ot : bits;
    }
    if (lane == 0) red[wid] = __longlong_as_double((long long)bits);
    __syncthreads();
    {
      unsigned long long bb = 0ull;
      for (int w2 = 0; w2 < 32; ++w2) {
        const unsigned long long x = (unsigned long long)__double_as_longlong(red[w2]);
        bb = x > bb ? x : bb;
      }
      nt = __longlong_as_double((long long)bb);
    }
    double* tmp = v; v = vn; vn = tmp;
    __syncthreads();
  }
  if (tid == 0) {
    const double tr = v[0];  // open train: the last vector is 1 x 1
    c.log_out[b] = logz + log(fabs(tr));
    c.sign_out[b] = tr < 0.0 ? -1 : 1;
  }
}

// ---------------------------------------------------------------------------------------------
// The same chain on a thread-block CLUSTER of 8 CTAs.  One step reads a whole d x d trace (512 KB at
// d = 256): a single SM pulls that out of L2 at ~45 GB/s, which is what bounds k_chain_open (~12 us per
// site).  Here CTA r computes the columns j = r (mod 8) of the new vector (its share of the trace: 64 KB),
// writes them into EVERY CTA's copy of the vector through distributed shared memory together with its
// share of the max-abs, and one hardware cluster barrier per site makes the new vector visible everywhere.
// ---------------------------------------------------------------------------------------------
constexpr int kCcCluster = 8;
constexpr int kCcThreads = 512;

__global__ void __launch_bounds__(kCcThreads) k_chain_cluster(ChainCtx c) {
  const int b = blockIdx.y, rank = blockIdx.x;   // cluster = one row of the grid
  const int* bond = c.bond + (int64_t)b * c.L1;
  const double* T = c.T + (int64_t)b * c.t_stride;
  double* env = c.env + (int64_t)b * c.env_stride;
  extern __shared__ double sm[];
  double* vbuf[2] = {sm, sm + c.dmax};                 // current / next vector (full copies in every CTA)
  double* mxs = sm + 2 * (size_t)c.dmax;               // [2][8] per-CTA max-abs of the next vector (ping-pong)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = kCcThreads / 32;
  if (tid == 0) vbuf[0][0] = 1.0;
  uint32_t vb_remote[2][kCcCluster], mx_remote[kCcCluster];
#pragma unroll
  for (int r = 0; r < kCcCluster; ++r) {
    vb_remote[0][r] = mapa_u32(smem_u32(vbuf[0]), (uint32_t)r);
    vb_remote[1][r] = mapa_u32(smem_u32(vbuf[1]), (uint32_t)r);
    mx_remote[r] = mapa_u32(smem_u32(mxs), (uint32_t)r);
  }
  double logz = 0.0, nt = 1.0;
  double pre_a[8], pre_b[8];
  bool have_pre = false;
  __syncthreads();
  cluster_sync_all();
  int cur = 0;
  for (int st = 0; st < c.L; ++st) {
    const int t = c.left ? st : c.L - 1 - st;
    const int rows = c.left ? bond[t] : bond[t + 1];
    const int cols = c.left ? bond[t + 1] : bond[t];
    const double* M = T + c.toff[t];
    double* v = vbuf[cur];
    // reference quirk (:93-97): the previous vector -- already stored -- is divided by its max-abs in place
    if (st > 0 && nt != 0.0 && c.normalize) {
      double* prev = env + c.eoff[c.left ? t - 1 : t + 1];
      for (int i = tid; i < rows; i += kCcThreads) {
        const double x = v[i] / nt;
        v[i] = x;
        if ((i & (kCcCluster - 1)) == rank) prev[i] = x;   // every CTA scales its copy, the owner stores
      }
      logz += log(nt);
      __syncthreads();
    }
    // columns j = rank (mod 8): warp per column, two columns' loads in flight.  The trace does not depend
    // on the running vector, so the first column pair of the NEXT site is requested before this site's
    // cluster barrier and its HBM / L2 latency overlaps the exchange.
    const int nxt = cur ^ 1;
    double mx = 0.0;
    double* curenv = env + c.eoff[t];
    for (int jj = wid; rank + kCcCluster * jj < cols; jj += 2 * NW) {
      const int ja = rank + kCcCluster * jj, jb = rank + kCcCluster * (jj + NW);
      const bool hb = jb < cols;
      double xa[8], xb[8];
      if (jj == wid && have_pre) {
#pragma unroll
        for (int m = 0; m < 8; ++m) { xa[m] = pre_a[m]; xb[m] = pre_b[m]; }
      } else {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int i = lane + 32 * m;
          xa[m] = i < rows ? M[i + (int64_t)rows * ja] : 0.0;
          xb[m] = (hb && i < rows) ? M[i + (int64_t)rows * jb] : 0.0;
        }
      }
      double sa = 0.0, sb = 0.0;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int i = lane + 32 * m;
        const double vi = i < rows ? v[i] : 0.0;
        sa += vi * xa[m];
        sb += vi * xb[m];
      }
      for (int i = lane + 256; i < rows; i += 32) {   // bonds beyond 256
        const double vi = v[i];
        sa += vi * M[i + (int64_t)rows * ja];
        if (hb) sb += vi * M[i + (int64_t)rows * jb];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
      }
      // lane r delivers the value to CTA r; lane 8 stores the environment
      if (lane < kCcCluster) {
        asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(vb_remote[nxt][lane] + (uint32_t)ja * 8u), "d"(sa) : "memory");
        if (hb) asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(vb_remote[nxt][lane] + (uint32_t)jb * 8u), "d"(sb) : "memory");
      } else if (lane == kCcCluster) {
        curenv[ja] = sa;
        if (hb) curenv[jb] = sb;
      }
      mx = absmax_nan(mx, sa);
      if (hb) mx = absmax_nan(mx, sb);
    }
    // prefetch for the next site (its dimensions are in the bond array already)
    have_pre = false;
    if (st + 1 < c.L) {
      const int t2 = c.left ? st + 1 : c.L - 2 - st;
      const int rows2 = c.left ? bond[t2] : bond[t2 + 1];
      const int cols2 = c.left ? bond[t2 + 1] : bond[t2];
      const double* M2 = T + c.toff[t2];
      const int ja = rank + kCcCluster * wid, jb = rank + kCcCluster * (wid + NW);
      if (ja < cols2) {
        have_pre = true;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int i = lane + 32 * m;
          pre_a[m] = i < rows2 ? M2[i + (int64_t)rows2 * ja] : 0.0;
          pre_b[m] = (jb < cols2 && i < rows2) ? M2[i + (int64_t)rows2 * jb] : 0.0;
        }
      }
    }
    // max-abs of this CTA's columns -> everybody (bits order like |x|, NaN on top)
    __shared__ unsigned long long wmx[NW];
    if (lane == 0) wmx[wid] = (unsigned long long)__double_as_longlong(mx);
    __syncthreads();
    if (wid == 0) {
      unsigned long long bits = lane < NW ? wmx[lane] : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ot = __shfl_xor_sync(0xffffffffu, bits, o);
        bits = ot > bits ? ot : bits;
      }
      if (lane < kCcCluster)
        asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(mx_remote[lane] + (uint32_t)(((st & 1) * kCcCluster + rank) * 8)), "l"(bits) : "memory");
    }
    cluster_sync_all();   // release / acquire: the new vector and the 8 partial maxima are visible in every CTA
    {
      unsigned long long bb = 0ull;
#pragma unroll
      for (int r = 0; r < kCcCluster; ++r) {
        const unsigned long long x = (unsigned long long)__double_as_longlong(mxs[(st & 1) * kCcCluster + r]);
        bb = x > bb ? x : bb;
      }
      nt = __longlong_as_double((long long)bb);
    }
    cur = nxt;
  }
  if (rank == 0 && tid == 0) {
    const double tr = vbuf[cur][0];  // open train: the last vector is 1 x 1
    c.log_out[b] = logz + log(fabs(tr));
    c.sign_out[b] = tr < 0.0 ? -1 : 1;
  }
  cluster_sync_all();   // nobody leaves while its shared memory can still be written remotely
}

}  // namespace ttb
