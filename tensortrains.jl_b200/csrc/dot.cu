// dot.cu -- inner product of two tensor trains (src/abstract_tensor_train.jl:395-408), the first row
// after the hot path.  norm / norm2m / isapprox (:418-437) are host arithmetic on top of it.
//
//   C[a_l, a_1, b_1, b_l] = sum_{x, a_{l+1}, b_{l+1}} A_l[a_l, a_{l+1}, x] C[a_{l+1}, a_1, b_1, b_{l+1}] B_l[b_l, b_{l+1}, x]
//   dot = sum_{a_1, b_1} C[a_1, a_1, b_1, b_1] / (z_A z_B)
//
// Data layout: the boundary pair (a_1, b_1) is a slice index s; every slice is a d^A_l x d^B_l column-major
// matrix C_s.  One site is two batched GEMMs over all slices,
//   E_s[:, :, x] = A_l(x) C_s          (d^A_l x d^B_{l+1}, per x)
//   C_s'         = [E_s(1) .. E_s(Q)] [B_l(1) .. B_l(Q)]^T     (contraction over (b_{l+1}, x) at once)
// followed by a max-abs rescale whose logarithm is accumulated, so the value is returned as (log|.|, sign)
// and long chains do not overflow the way the reference's plain Float64 accumulation can.
#include <math.h>

#include "common.cuh"

namespace ttb {

// C(m x n) = alpha * A(m x k) * op(B); A column-major (lda); op(B) = B (k x n, ldb) or B^T with B n x k (ldb)
template <bool TRANSB>
__device__ __forceinline__ void gemm_tile32(int m, int n, int k, const double* A, int64_t lda, const double* B, int64_t ldb,
                                            double* Cout, int64_t ldc, double alpha, int tile, unsigned long long* mx_bits) {
  __shared__ double As[32][33], Bs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int tiles_m = (m + 31) / 32;
  const int i0 = (tile % tiles_m) * 32, j0 = (tile / tiles_m) * 32;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k0 = 0; k0 < k; k0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kk = ty + 8 * r;
      As[kk][tx] = (i0 + tx < m && k0 + kk < k) ? A[i0 + tx + lda * (k0 + kk)] : 0.0;
      if (TRANSB) Bs[kk][tx] = (j0 + tx < n && k0 + kk < k) ? B[j0 + tx + ldb * (k0 + kk)] : 0.0;
      else        Bs[tx][kk] = (k0 + tx < k && j0 + kk < n) ? B[k0 + tx + ldb * (j0 + kk)] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const double a = As[kk][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] += a * Bs[kk][ty + 8 * r];
    }
    __syncthreads();
  }
  double mx = 0.0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + tx, j = j0 + ty + 8 * r;
    if (i < m && j < n) { const double v = alpha * acc[r]; Cout[i + ldc * j] = v; mx = absmax_nan(mx, v); }
  }
  if (mx_bits) {
    unsigned long long bits = (unsigned long long)__double_as_longlong(mx);  // |.| and NaN order as integers
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long ot = __shfl_xor_sync(0xffffffffu, bits, o); bits = ot > bits ? ot : bits; }
    if (tx == 0 && bits) atomicMax(mx_bits, bits);
  }
}

struct DotCtx {
  const double *A, *B;            // site tensors of the two trains at the current site
  int da_l, da_r, db_l, db_r, Q;
  double *C, *Cn, *E;             // [S][capC], [S][capC], [S][Q*capC]
  int64_t capC;
  double* scal;                   // [0] = scale applied when C is read, [1] = accumulated log, [2] = sign product
  unsigned long long* mx_bits;
};

__global__ void k_dot_init(double* C, int64_t capC, int ra, int rb) {
  pdl_enter();
  const int s = blockIdx.x, a1 = s % ra, b1 = s / ra;
  for (int i = threadIdx.x; i < ra * rb; i += blockDim.x) C[(int64_t)s * capC + i] = (i == a1 + ra * b1) ? 1.0 : 0.0;
}

// E_s(x) = A(x) * (scale * C_s)
__global__ void __launch_bounds__(256) k_dot_ac(DotCtx c) {
  pdl_enter();
  const int x = blockIdx.y, s = blockIdx.z;
  gemm_tile32<false>(c.da_l, c.db_r, c.da_r, c.A + (int64_t)c.da_l * c.da_r * x, c.da_l, c.C + (int64_t)s * c.capC, c.da_r,
                     c.E + ((int64_t)s * c.Q + x) * c.capC, c.da_l, c.scal[0], blockIdx.x, nullptr);
}

// C_s' = E_s B^T over the combined (b_{l+1}, x) index.  E_s(x) blocks are capC apart, so the contraction is
// split per x and accumulated by re-reading the tile: keep it simple -- one launch per site, x looped here.
__global__ void __launch_bounds__(256) k_dot_eb(DotCtx c) {
  pdl_enter();
  const int s = blockIdx.z;
  __shared__ double As[32][33], Bs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int m = c.da_l, n = c.db_l, k = c.db_r;
  const int tiles_m = (m + 31) / 32;
  const int i0 = (blockIdx.x % tiles_m) * 32, j0 = (blockIdx.x / tiles_m) * 32;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int x = 0; x < c.Q; ++x) {
    const double* E = c.E + ((int64_t)s * c.Q + x) * c.capC;       // m x k, ld m
    const double* B = c.B + (int64_t)c.db_l * c.db_r * x;          // n x k, ld n
    for (int k0 = 0; k0 < k; k0 += 32) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int kk = ty + 8 * r;
        As[kk][tx] = (i0 + tx < m && k0 + kk < k) ? E[i0 + tx + (int64_t)m * (k0 + kk)] : 0.0;
        Bs[kk][tx] = (j0 + tx < n && k0 + kk < k) ? B[j0 + tx + (int64_t)n * (k0 + kk)] : 0.0;
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < 32; ++kk) {
        const double a = As[kk][tx];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] += a * Bs[kk][ty + 8 * r];
      }
      __syncthreads();
    }
  }
  double mx = 0.0;
  double* Cn = c.Cn + (int64_t)s * c.capC;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + tx, j = j0 + ty + 8 * r;
    if (i < m && j < n) { Cn[i + (int64_t)m * j] = acc[r]; mx = absmax_nan(mx, acc[r]); }
  }
  unsigned long long bits = (unsigned long long)__double_as_longlong(mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long ot = __shfl_xor_sync(0xffffffffu, bits, o); bits = ot > bits ? ot : bits; }
  if (tx == 0 && bits) atomicMax(c.mx_bits, bits);
}

// turn the max-abs of the step into the scale the next step applies; accumulate its log
__global__ void k_dot_scale(double* scal, unsigned long long* mx_bits) {
  pdl_enter();
  const double mx = __longlong_as_double((long long)*mx_bits);
  if (mx != 0.0 && !isnan(mx) && !isinf(mx)) { scal[0] = 1.0 / mx; scal[1] += log(mx); }
  else scal[0] = 1.0;
  *mx_bits = 0ull;
}

// d = sum_{a1,b1} C_{(a1,b1)}[a1, b1];  out = (log|d| + acc - log z_A - log z_B, sign)
__global__ void k_dot_final(const double* C, int64_t capC, int ra, int rb, const double* scal,
                            const double* logzA, const int* sgnA, const double* logzB, const int* sgnB, double* out) {
  pdl_enter();
  __shared__ double red[40];
  double v = 0.0;
  for (int s = threadIdx.x; s < ra * rb; s += blockDim.x) {
    const int a1 = s % ra, b1 = s / ra;
    v += C[(int64_t)s * capC + a1 + (int64_t)ra * b1];
  }
  v = block_sum(v, red);
  if (threadIdx.x == 0) {
    out[0] = log(fabs(v)) + scal[1] - *logzA - *logzB;
    out[1] = (v < 0.0 ? -1.0 : 1.0) * (double)(*sgnA) * (double)(*sgnB);
    if (v != v) out[1] = v;  // NaN propagates like the reference's plain arithmetic
  }
}

}  // namespace ttb

using namespace ttb;

extern "C" int ttb_dot(ttb_handle a, int64_t ba, ttb_handle b, int64_t bb, double* logabs, int* sign) {
  TTB_TRY(check_handle(a));
  TTB_TRY(check_handle(b));
  if (!logabs || !sign) return fail(TTB_EINVAL, "null output pointer");
  if (ba < 0 || ba >= a->batch || bb < 0 || bb >= b->batch) return fail(TTB_EINVAL, "train index out of range");
  if (a->device != b->device) return fail(TTB_EINVAL, "dot: the two trains live on different devices");
  if (a->L != b->L || a->Q != b->Q)
    return fail(TTB_ESHAPE, "dot: trains must have the same length and physical dimension (got L=%lld,%lld q=%lld,%lld)",
                (long long)a->L, (long long)b->L, (long long)a->Q, (long long)b->Q);
  TTB_CUDA(cudaSetDevice(a->device));
  TTB_TRY(sync_bonds_to_host(a));
  if (b != a) TTB_TRY(sync_bonds_to_host(b));
  const int L = (int)a->L, Q = (int)a->Q, L1 = L + 1;
  const int* bA = a->h_bond.data() + ba * L1;
  const int* bB = b->h_bond.data() + bb * L1;
  const int ra = bA[0], rb = bB[0];
  if (bA[L] != ra || bB[L] != rb) return fail(TTB_ESHAPE, "dot: first and last bond of a train must agree");
  int dA = 1, dB = 1;
  for (int t = 0; t <= L; ++t) { dA = bA[t] > dA ? bA[t] : dA; dB = bB[t] > dB ? bB[t] : dB; }
  const int64_t S = (int64_t)ra * rb, capC = (int64_t)dA * dB;
  const int64_t need = S * capC * (2 + Q);
  if (need > (int64_t(1) << 31)) return fail(TTB_EUNSUPPORTED, "dot: boundary*bond product too large (%lld doubles of scratch)", (long long)need);
  // scratch kept on handle `a` between calls (grow-only, ttb_trim releases it)
  if (a->ws.scr[5].ensure(sizeof(double) * (need + 8)) != cudaSuccess) return fail(TTB_ENOMEM, "dot scratch");
  double* scratch = static_cast<double*>(a->ws.scr[5].p);
  double* scal = scratch + need;
  cudaStream_t s = a->stream;
  const double h_scal[8] = {1.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  cudaMemcpyAsync(scal, h_scal, sizeof(h_scal), cudaMemcpyHostToDevice, s);
  unsigned long long* mx_bits = reinterpret_cast<unsigned long long*>(scal + 3);
  double* out_dev = scal + 4;
  CallTimer timer(a);
  DotCtx c{};
  c.Q = Q; c.capC = capC; c.scal = scal; c.mx_bits = mx_bits;
  c.C = scratch; c.Cn = scratch + S * capC; c.E = scratch + 2 * S * capC;
  TTB_LAUNCH(a, "k_dot_init", (k_dot_init<<<(unsigned)S, 128, 0, s>>>(c.C, capC, ra, rb)));
  for (int t = L - 1; t >= 0; --t) {
    c.A = a->slab + ba * a->tt_stride + a->off[t];
    c.B = b->slab + bb * b->tt_stride + b->off[t];
    c.da_l = bA[t]; c.da_r = bA[t + 1]; c.db_l = bB[t]; c.db_r = bB[t + 1];
    const unsigned t1 = (unsigned)(((c.da_l + 31) / 32) * ((c.db_r + 31) / 32));
    const unsigned t2 = (unsigned)(((c.da_l + 31) / 32) * ((c.db_l + 31) / 32));
    // a chain of 3 L short dependent kernels: programmatic dependent launch (common.cuh)
    TTB_LAUNCH(a, "k_dot_ac", launch_pdl(k_dot_ac, dim3(t1, (unsigned)Q, (unsigned)S), dim3(256), 0, s, c));
    TTB_LAUNCH(a, "k_dot_eb", launch_pdl(k_dot_eb, dim3(t2, 1, (unsigned)S), dim3(256), 0, s, c));
    if (t > 0) TTB_LAUNCH(a, "k_dot_scale", launch_pdl(k_dot_scale, dim3(1), dim3(1), 0, s, scal, mx_bits));  // the last product stays unscaled
    double* tmp = c.C; c.C = c.Cn; c.Cn = tmp;
  }
  TTB_LAUNCH(a, "k_dot_final", launch_pdl(k_dot_final, dim3(1), dim3(256), 0, s, (const double*)c.C, capC, ra, rb, (const double*)scal,
                                          (const double*)(a->d_logz + ba), (const int*)(a->d_signz + ba),
                                          (const double*)(b->d_logz + bb), (const int*)(b->d_signz + bb), out_dev));
  double h_out[2] = {0.0, 0.0};
  cudaMemcpyAsync(h_out, out_dev, sizeof(h_out), cudaMemcpyDeviceToHost, s);
  int rc = timer.finish();
  if (rc != TTB_OK) return rc;
  *logabs = h_out[0];
  *sign = h_out[1] < 0.0 ? -1 : 1;
  if (h_out[1] != h_out[1]) *logabs = h_out[1];
  return TTB_OK;
}
