// dot.cu -- inner product of two tensor trains (src/abstract_tensor_train.jl:395-408), the first row
// after the hot path.  norm / norm2m / isapprox (:418-437) are host arithmetic on top of it.
//
//   C[a_l, a_1, b_1, b_l] = sum_{x, a_{l+1}, b_{l+1}} A_l[a_l, a_{l+1}, x] C[a_{l+1}, a_1, b_1, b_{l+1}] B_l[b_l, b_{l+1}, x]
//   dot = sum_{a_1, b_1} C[a_1, a_1, b_1, b_1] / (z_A z_B)
//
// Data layout: the boundary pair (a_1, b_1) is a slice index s; every slice is a d^A_l x d^B_l column-major
// matrix C_s.  One site is two TABLES of FP64 tensor-core GEMMs (k_gemm_items) over all slices,
//   E_s[:, :, x] = A_l(x) C_s          (d^A_l x d^B_{l+1}, per x)
//   C_s'         = [E_s(1) .. E_s(Q)] [B_l(1) .. B_l(Q)]^T     (contraction over (b_{l+1}, x) at once)
// followed by a max-abs rescale whose logarithm is accumulated, so the value is returned as (log|.|, sign)
// and long chains do not overflow the way the reference's plain Float64 accumulation can.
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "gemm_items.cuh"

namespace ttb {

// Round 2: both products of a site are tables of FP64 tensor-core GEMMs (k_gemm_items, gemm_items.cuh) -- one
// item per boundary slice (and per x for the first product); the scalar 32 x 32 tiles of round 1 are gone.
struct DotScaleItem { double* buf; int64_t n; };

// max-abs over all slices of the step into mx[slot] (ordered bits, NaN-propagating)
__global__ void __launch_bounds__(256) k_dot_absmax(const DotScaleItem* __restrict__ items, unsigned long long* mx, int slot) {
  pdl_enter();
  const DotScaleItem it = items[blockIdx.y];
  double m = 0.0;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < it.n; o += (int64_t)gridDim.x * blockDim.x) m = absmax_nan(m, it.buf[o]);
  unsigned long long bits = (unsigned long long)__double_as_longlong(m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long ot = __shfl_xor_sync(0xffffffffu, bits, o); bits = ot > bits ? ot : bits; }
  if ((threadIdx.x & 31) == 0 && bits) atomicMax(mx + slot, bits);
}

// every slice divided by the step's max-abs (finite and non-zero), whose logarithm the final kernel sums
__global__ void __launch_bounds__(256) k_dot_scale(const DotScaleItem* __restrict__ items, const unsigned long long* mx, int slot) {
  pdl_enter();
  const DotScaleItem it = items[blockIdx.y];
  const double m = __longlong_as_double((long long)mx[slot]);
  if (m == 0.0 || isnan(m) || isinf(m)) return;
  const double inv = 1.0 / m;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < it.n; o += (int64_t)gridDim.x * blockDim.x) it.buf[o] *= inv;
}

__global__ void k_dot_init(double* C, int64_t capC, int ra, int rb) {
  pdl_enter();
  const int s = blockIdx.x, a1 = s % ra, b1 = s / ra;
  for (int i = threadIdx.x; i < ra * rb; i += blockDim.x) C[(int64_t)s * capC + i] = (i == a1 + ra * b1) ? 1.0 : 0.0;
}

// d = sum_{a1,b1} C_{(a1,b1)}[a1, b1];  out = (log|d| + sum of the logs of the applied scales - log z_A - log z_B, sign)
__global__ void k_dot_final(const double* C, int64_t capC, int ra, int rb, const unsigned long long* mx, int nscaled,
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
    double lg = 0.0;
    for (int t = 0; t < nscaled; ++t) {
      const double m = __longlong_as_double((long long)mx[t]);
      if (m != 0.0 && !isnan(m) && !isinf(m)) lg += log(m);
    }
    out[0] = log(fabs(v)) + lg - *logzA - *logzB;
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
  if (need > (int64_t(1) << 31) || S * Q > 60000) return fail(TTB_EUNSUPPORTED, "dot: boundary*bond product too large (%lld doubles of scratch)", (long long)need);
  // scratch kept on handle `a` between calls (grow-only, ttb_trim releases it)
  const size_t tab_bytes = (size_t)L * (sizeof(GemmItem) * (size_t)(S * Q + S) + sizeof(DotScaleItem) * (size_t)S);
  if (a->ws.scr[5].ensure(sizeof(double) * (need + 8) + sizeof(unsigned long long) * (L + 8) + tab_bytes + 256) != cudaSuccess)
    return fail(TTB_ENOMEM, "dot scratch");
  double* scratch = static_cast<double*>(a->ws.scr[5].p);
  double* out_dev = scratch + need;
  unsigned long long* mx = reinterpret_cast<unsigned long long*>(scratch + need + 8);
  unsigned char* d_tab = reinterpret_cast<unsigned char*>(mx + L + 8);
  cudaStream_t st = a->stream;
  double* Cb[2] = {scratch, scratch + S * capC};
  double* E = scratch + 2 * S * capC;
  // tables of all sites: the chain runs t = L-1 .. 0, site t reads C[cur] and writes C[cur ^ 1]
  std::vector<GemmItem> g1((size_t)L * S * Q), g2((size_t)L * S);
  std::vector<DotScaleItem> sc((size_t)L * S);
  int cur = 0;
  for (int step = 0; step < L; ++step, cur ^= 1) {
    const int t = L - 1 - step;
    const double* At = a->slab + ba * a->tt_stride + a->off[t];
    const double* Bt = b->slab + bb * b->tt_stride + b->off[t];
    const int da_l = bA[t], da_r = bA[t + 1], db_l = bB[t], db_r = bB[t + 1];
    for (int64_t sl = 0; sl < S; ++sl) {
      for (int x = 0; x < Q; ++x) {   // E_s(x) = A(x) C_s   (da_l x db_r)
        GemmItem& it = g1[((size_t)step * S + sl) * Q + x];
        it = GemmItem{};
        it.A = At + (int64_t)da_l * da_r * x; it.lda = da_l; it.B = Cb[cur] + sl * capC; it.ldb = da_r;
        it.m = da_l; it.n = db_r; it.k = da_r; it.C = E + (sl * Q + x) * capC; it.ldc = da_l; it.nsum = 1;
      }
      GemmItem& s2 = g2[(size_t)step * S + sl];   // C_s' = sum_x E_s(x) B(x)^T   (da_l x db_l), B stored n x k
      s2 = GemmItem{};
      s2.A = E + sl * Q * capC; s2.lda = da_l; s2.sA = capC; s2.B = Bt; s2.ldb = db_l; s2.sB = (int64_t)db_l * db_r;
      s2.m = da_l; s2.n = db_l; s2.k = db_r; s2.C = Cb[cur ^ 1] + sl * capC; s2.ldc = da_l; s2.nsum = Q;
      sc[(size_t)step * S + sl] = DotScaleItem{Cb[cur ^ 1] + sl * capC, (int64_t)da_l * db_l};
    }
  }
  std::vector<unsigned char> blob(tab_bytes);
  const size_t o2 = sizeof(GemmItem) * g1.size(), o3 = o2 + sizeof(GemmItem) * g2.size();
  memcpy(blob.data(), g1.data(), o2);
  memcpy(blob.data() + o2, g2.data(), sizeof(GemmItem) * g2.size());
  memcpy(blob.data() + o3, sc.data(), sizeof(DotScaleItem) * sc.size());
  cudaMemcpyAsync(d_tab, blob.data(), tab_bytes, cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(mx, 0, sizeof(unsigned long long) * (L + 8), st);
  cudaStreamSynchronize(st);   // blob is local storage
  const GemmItem* d1 = reinterpret_cast<const GemmItem*>(d_tab);
  const GemmItem* d2 = reinterpret_cast<const GemmItem*>(d_tab + o2);
  const DotScaleItem* d3 = reinterpret_cast<const DotScaleItem*>(d_tab + o3);
  CallTimer timer(a);
  TTB_LAUNCH(a, "k_dot_init", (k_dot_init<<<(unsigned)S, 128, 0, st>>>(Cb[0], capC, ra, rb)));
  for (int step = 0; step < L; ++step) {
    const int t = L - 1 - step;
    const int da_l = bA[t], db_l = bB[t], db_r = bB[t + 1];
    launch_gemm<false, false>(a, d1 + (size_t)step * S * Q, (int)(S * Q), tiles64(da_l, db_r));
    launch_gemm<false, true>(a, d2 + (size_t)step * S, (int)S, tiles64(da_l, db_l));
    if (t > 0) {   // the last product stays unscaled
      const unsigned gs = (unsigned)std::min<int64_t>(((int64_t)da_l * db_l + 255) / 256, 64);
      TTB_LAUNCH(a, "k_dot_absmax", (k_dot_absmax<<<dim3(gs, (unsigned)S), 256, 0, st>>>(d3 + (size_t)step * S, mx, step)));
      TTB_LAUNCH(a, "k_dot_scale", (k_dot_scale<<<dim3(gs, (unsigned)S), 256, 0, st>>>(d3 + (size_t)step * S, mx, step)));
    }
  }
  const double* Cfin = Cb[L & 1];
  TTB_LAUNCH(a, "k_dot_final", (k_dot_final<<<1, 256, 0, st>>>(Cfin, capC, ra, rb, mx, L - 1, (const double*)(a->d_logz + ba),
                                                                 (const int*)(a->d_signz + ba), (const double*)(b->d_logz + bb),
                                                                 (const int*)(b->d_signz + bb), out_dev)));
  double h_out[2] = {0.0, 0.0};
  cudaMemcpyAsync(h_out, out_dev, sizeof(h_out), cudaMemcpyDeviceToHost, st);
  int rc = timer.finish();
  if (rc != TTB_OK) return rc;
  *logabs = h_out[0];
  *sign = h_out[1] < 0.0 ? -1 : 1;
  if (h_out[1] != h_out[1]) *logabs = h_out[1];
  return TTB_OK;
}
