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

}  // namespace ttb
