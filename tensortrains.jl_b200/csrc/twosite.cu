// twosite.cu -- the two-site pieces next to the hot path:
//   _merge_tensors        src/utils.jl:37-44            C[m,n,xA,xB] = sum_k A[m,k,xA] B[k,n,xB]
//   _split_tensor         src/abstract_tensor_train.jl:283-308   C_resh[(al,xl),(al1,xl1)] = U L V' -> A = U, B = L V' (lr = Left)
//                                                                 or A = U L, B = V' (lr = Right), with any SVDTrunc
//   precompute_left/right_environments  src/tensor_train.jl:320-334   per-datapoint partial products
//   update_environments!  src/dmrg.jl:124-151           one step of the same recursion after a site changed
// The optimisation loop of two_site_dmrg! itself (Optim / Adam, src/dmrg.jl:59-121) is host control flow and
// out of scope.
//
// _split_tensor is ONE site step of the sweep machinery: with the temporary two-site chain segment
//   lr = Left :  T1[al, (al1,xl1), xl] = C[al, al1, xl, xl1],  T2[(al1,xl1), j, x] = delta      (T1 T2 = C)
//   lr = Right:  T1[i, (al,xl), x] = delta,                    T2[(al,xl), al1, xl1] = C[al, al1, xl, xl1]
// the left-sweep step at site 1 (resp. right-sweep step at site 2) folds exactly C_resh, truncates it with the
// device rank rule and absorbs L V' (resp. U L) into the delta site, i.e. writes A and B.  The sweep's max-abs
// rescale of the absorbed factor is undone from the segment's z, so the pair (A, B) is the reference's.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "gemm_items.cuh"

namespace ttb {

__global__ void k_split_fill(const double* __restrict__ C, double* sc, double* sd, int d0, int d2, int Q, int left) {
  const int64_t nC = (int64_t)d0 * d2 * Q * Q;
  const int64_t nD = left ? (int64_t)d2 * Q * d2 * Q : (int64_t)d0 * d0 * Q * Q;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t o = i0; o < nC; o += stride) {
    const int al = (int)(o % d0);
    int64_t r = o / d0;
    const int al1 = (int)(r % d2); r /= d2;
    const int xl = (int)(r % Q), xl1 = (int)(r / Q);
    const int64_t dst = left ? al + (int64_t)d0 * ((al1 + (int64_t)d2 * xl1) + (int64_t)d2 * Q * xl)
                             : (al + (int64_t)d0 * xl) + (int64_t)d0 * Q * (al1 + (int64_t)d2 * xl1);
    sc[dst] = C[o];
  }
  for (int64_t o = i0; o < nD; o += stride) {
    double v;
    if (left) {   // T2[(al1,xl1), j, x], rows d2 Q
      const int row = (int)(o % ((int64_t)d2 * Q));
      const int64_t r = o / ((int64_t)d2 * Q);
      const int j = (int)(r % d2), x = (int)(r / d2);
      v = (row % d2 == j && row / d2 == x) ? 1.0 : 0.0;
    } else {      // T1[i, (al,xl), x], rows d0
      const int i = (int)(o % d0);
      const int64_t r = o / d0;
      const int col = (int)(r % ((int64_t)d0 * Q)), x = (int)(r / ((int64_t)d0 * Q));
      v = (col % d0 == i && col / d0 == x) ? 1.0 : 0.0;
    }
    sd[o] = v;
  }
}

// dst = src * exp(-logz[0]) (the absorbed factor, un-rescaled) or a plain copy (logz == nullptr)
__global__ void k_copy_unscale(double* dst, const double* src, int64_t n, const double* logz) {
  const double f = logz ? exp(-logz[0]) : 1.0;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) dst[o] = src[o] * f;
}

// per-datapoint partial products.  env[n][j][0 .. d_j) with stride `ds` per entry, L + 1 entries per datapoint:
//   left : entry j = A_0(x_0) ... A_{j-1}(x_{j-1})   (1 x d_j), entry 0 = [1]      (prodA_left[j],  tensor_train.jl:320-326)
//   right: entry j = A_j(x_j) ... A_{L-1}(x_{L-1})   (d_j x 1), entry L = [1]      (prodA_right[j+1], :328-334)
// t_lo <= t < t_hi: the sites whose step is (re)computed -- the whole chain, or one site (update_environments!)
struct DataEnvCtx {
  const double* tt; const int64_t* off; const int* bond; int L, Q; int64_t ds;
  const int32_t* X; double* env; int left, t_lo, t_hi;
};
__global__ void __launch_bounds__(128) k_data_env(DataEnvCtx c) {
  const int64_t n = blockIdx.x;
  double* e = c.env + n * (int64_t)(c.L + 1) * c.ds;
  const int32_t* x = c.X + n * c.L;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  if (c.left) {
    if (c.t_lo == 0 && tid == 0) e[0] = 1.0;
    __syncthreads();
    for (int t = c.t_lo; t < c.t_hi; ++t) {
      const int dl = c.bond[t], dr = c.bond[t + 1];
      const double* A = c.tt + c.off[t] + (int64_t)dl * dr * x[t];
      const double* v = e + (int64_t)t * c.ds;
      double* w = e + (int64_t)(t + 1) * c.ds;
      for (int j = wid; j < dr; j += nw) {
        double a = 0.0;
        for (int m = lane; m < dl; m += 32) a += v[m] * A[m + (int64_t)dl * j];
        a = warp_sum(a);
        if (lane == 0) w[j] = a;
      }
      __threadfence_block();
      __syncthreads();
    }
  } else {
    if (c.t_hi == c.L && tid == 0) e[(int64_t)c.L * c.ds] = 1.0;
    __syncthreads();
    for (int t = c.t_hi - 1; t >= c.t_lo; --t) {
      const int dl = c.bond[t], dr = c.bond[t + 1];
      const double* A = c.tt + c.off[t] + (int64_t)dl * dr * x[t];
      const double* v = e + (int64_t)(t + 1) * c.ds;
      double* w = e + (int64_t)t * c.ds;
      for (int m = tid; m < dl; m += blockDim.x) {
        double a = 0.0;
        for (int j = 0; j < dr; ++j) a += A[m + (int64_t)dl * j] * v[j];
        w[m] = a;
      }
      __threadfence_block();
      __syncthreads();
    }
  }
}

// a caller pointer that may be host or device memory, seen as device memory
struct DevView {
  void* dev = nullptr;
  void* user = nullptr;
  size_t bytes = 0;
  bool staged = false;
  cudaError_t open(void* p, size_t n, bool copy_in, cudaStream_t s) {
    user = p; bytes = n;
    cudaPointerAttributes at{};
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e == cudaSuccess && at.type == cudaMemoryTypeDevice) { dev = p; return cudaSuccess; }
    cudaGetLastError();
    staged = true;
    e = cudaMalloc(&dev, std::max<size_t>(n, 16));
    if (e == cudaSuccess && copy_in) e = cudaMemcpyAsync(dev, p, n, cudaMemcpyHostToDevice, s);
    return e;
  }
  cudaError_t close(bool copy_out, cudaStream_t s) {
    cudaError_t e = cudaSuccess;
    if (staged) {
      if (copy_out) e = cudaMemcpyAsync(user, dev, bytes, cudaMemcpyDeviceToHost, s);
      cudaStreamSynchronize(s);
      cudaFree(dev);
    }
    dev = nullptr; staged = false;
    return e;
  }
};

static int site_pair_dims(ttb_tt* h, int64_t b, int64_t t, int& d0, int& d1, int& d2) {
  TTB_TRY(check_handle(h));
  if (b < 0 || b >= h->batch) return fail(TTB_EINVAL, "train index out of range");
  if (t < 0 || t + 1 >= h->L) return fail(TTB_EINVAL, "site pair (%lld, %lld) outside the train", (long long)t, (long long)t + 1);
  TTB_TRY(sync_bonds_to_host(h));
  const int* bd = h->h_bond.data() + b * (h->L + 1);
  d0 = bd[t]; d1 = bd[t + 1]; d2 = bd[t + 2];
  return TTB_OK;
}

}  // namespace ttb

using namespace ttb;

extern "C" {

int ttb_merge_tensors(ttb_handle h, int64_t b, int64_t t, double* out) {
  if (!out) return fail(TTB_EINVAL, "null argument");
  int d0, d1, d2;
  TTB_TRY(site_pair_dims(h, b, t, d0, d1, d2));
  const int Q = (int)h->Q;
  const size_t nC = (size_t)d0 * d2 * Q * Q;
  DevView ov;
  cudaError_t e = ov.open(out, sizeof(double) * nC, false, h->stream);
  GemmItem* d_items = nullptr;
  if (e == cudaSuccess) e = cudaMalloc(&d_items, sizeof(GemmItem) * Q * Q);
  int rc = TTB_OK;
  if (e != cudaSuccess) rc = fail(TTB_ENOMEM, "merge_tensors: %s", cudaGetErrorString(e));
  if (rc == TTB_OK) {
    const double* A = h->slab + b * h->tt_stride + h->off[t];
    const double* B = h->slab + b * h->tt_stride + h->off[t + 1];
    std::vector<GemmItem> it((size_t)Q * Q);
    for (int xa = 0; xa < Q; ++xa)
      for (int xb = 0; xb < Q; ++xb) {
        GemmItem& g = it[xa + (size_t)Q * xb];
        g = GemmItem{};
        g.A = A + (int64_t)xa * d0 * d1; g.lda = d0; g.B = B + (int64_t)xb * d1 * d2; g.ldb = d1;
        g.m = d0; g.n = d2; g.k = d1; g.C = static_cast<double*>(ov.dev) + (size_t)(xa + (size_t)Q * xb) * d0 * d2; g.ldc = d0; g.nsum = 1;
      }
    cudaMemcpyAsync(d_items, it.data(), sizeof(GemmItem) * it.size(), cudaMemcpyHostToDevice, h->stream);
    cudaStreamSynchronize(h->stream);
    CallTimer tm(h);
    launch_gemm<false, false>(h, d_items, Q * Q, tiles64(d0, d2));
    rc = tm.finish();
  }
  e = ov.close(rc == TTB_OK, h->stream);
  cudaFree(d_items);
  if (rc == TTB_OK && e != cudaSuccess) rc = fail(TTB_ECUDA, "merge_tensors: %s", cudaGetErrorString(e));
  return rc;
}

int ttb_split_tensor(ttb_handle h, int64_t b, int64_t t, const double* C, const ttb_trunc* tr, int left, double* trunc_err) {
  if (!C || !tr) return fail(TTB_EINVAL, "null argument");
  int d0, d1, d2;
  TTB_TRY(site_pair_dims(h, b, t, d0, d1, d2));
  const int Q = (int)h->Q;
  const int64_t cap = h->bond0[t + 1];
  const size_t nC = (size_t)d0 * d2 * Q * Q;
  // the segment: bonds (d0, d2 Q, d2) for Left, (d0, d0 Q, d2) for Right
  const int64_t mid = left ? (int64_t)d2 * Q : (int64_t)d0 * Q;
  const int64_t sb[3] = {d0, mid, d2};
  ttb_tt* seg = nullptr;
  TTB_TRY(create_handle(2, sb, Q, 0, 1, &seg, false));
  DevView cv;
  cudaError_t e = cv.open(const_cast<double*>(C), sizeof(double) * nC, true, seg->stream);
  int rc = TTB_OK;
  if (e != cudaSuccess) rc = fail(TTB_ENOMEM, "split_tensor: %s", cudaGetErrorString(e));
  if (rc == TTB_OK) {
    double* s0 = seg->slab + seg->off[0];
    double* s1 = seg->slab + seg->off[1];
    const unsigned g = (unsigned)std::min<size_t>((std::max<size_t>(nC, (size_t)mid * mid) + 255) / 256, 2048);
    k_split_fill<<<g, 256, 0, seg->stream>>>(static_cast<const double*>(cv.dev), left ? s0 : s1, left ? s1 : s0, d0, d2, Q, left ? 1 : 0);
    cudaStreamSynchronize(seg->stream);
  }
  cv.close(false, seg->stream);
  if (rc == TTB_OK) rc = left ? ttb_orthogonalize_left(seg, tr, 1, 1) : ttb_orthogonalize_right(seg, tr, 2, 2);
  int64_t nb[3] = {0, 0, 0};
  if (rc == TTB_OK) rc = ttb_get_bonds(seg, 0, nb);
  if (rc == TTB_OK && nb[1] > cap)
    rc = fail(TTB_EUNSUPPORTED, "split_tensor: the new bond %lld exceeds the capacity %lld the train was created with at bond %lld",
              (long long)nb[1], (long long)cap, (long long)(t + 1));
  if (rc == TTB_OK && trunc_err) {
    int64_t ranks[3]; double errs[3] = {0.0, 0.0, 0.0};
    if (ttb_sweep_stats(seg, 0, ranks, errs) == TTB_OK) *trunc_err = errs[1];
  }
  if (rc == TTB_OK) {
    const int m1 = (int)nb[1];
    double* A = h->slab + b * h->tt_stride + h->off[t];
    double* B = h->slab + b * h->tt_stride + h->off[t + 1];
    cudaStreamSynchronize(h->stream);
    const int64_t nA = (int64_t)d0 * m1 * Q, nB = (int64_t)m1 * d2 * Q;
    // the factor that absorbed L was divided by its max-abs and the segment's z carries it: undo
    k_copy_unscale<<<(unsigned)std::min<int64_t>((nA + 255) / 256, 1024), 256, 0, seg->stream>>>(A, seg->slab + seg->off[0], nA, left ? nullptr : seg->d_logz);
    k_copy_unscale<<<(unsigned)std::min<int64_t>((nB + 255) / 256, 1024), 256, 0, seg->stream>>>(B, seg->slab + seg->off[1], nB, left ? seg->d_logz : nullptr);
    h->h_bond[b * (h->L + 1) + t + 1] = m1;
    cudaMemcpyAsync(h->d_bond + b * (h->L + 1) + t + 1, &h->h_bond[b * (h->L + 1) + t + 1], sizeof(int), cudaMemcpyHostToDevice, seg->stream);
    e = cudaStreamSynchronize(seg->stream);
    if (e != cudaSuccess) rc = fail(TTB_ECUDA, "split_tensor: %s", cudaGetErrorString(e));
    h->last_ms = seg->last_ms; h->last_launches = seg->last_launches + 3;
  }
  ttb_destroy(seg);
  return rc;
}

static int data_env(ttb_handle h, int64_t b, const int32_t* X, int64_t n, int left, double* env, int t_lo, int t_hi, bool update) {
  TTB_TRY(check_handle(h));
  if (b < 0 || b >= h->batch) return fail(TTB_EINVAL, "train index out of range");
  if (!X || !env || n < 0) return fail(TTB_EINVAL, "bad argument");
  if (h->periodic || h->bond0[0] != 1 || h->bond0[h->L] != 1) return fail(TTB_EINVAL, "data environments are defined for TensorTrain only");
  if (n == 0) return TTB_OK;
  const int L = (int)h->L;
  for (int64_t i = 0; i < n * L; ++i)
    if (X[i] < 0 || X[i] >= h->Q) return fail(TTB_EINVAL, "X[%lld]=%d outside the domain 0:%lld", (long long)i, X[i], (long long)h->Q - 1);
  const int64_t ds = h->dmax;
  const size_t bytes = sizeof(double) * (size_t)n * (L + 1) * ds;
  DevView ev;
  int32_t* d_X = nullptr;
  cudaError_t e = ev.open(env, bytes, update, h->stream);
  if (e == cudaSuccess && !update && ev.staged) e = cudaMemsetAsync(ev.dev, 0, bytes, h->stream);
  if (e == cudaSuccess) e = cudaMalloc(&d_X, sizeof(int32_t) * n * L);
  int rc = TTB_OK;
  if (e != cudaSuccess) rc = fail(TTB_ENOMEM, "data environments: %s", cudaGetErrorString(e));
  if (rc == TTB_OK) {
    cudaMemcpyAsync(d_X, X, sizeof(int32_t) * n * L, cudaMemcpyHostToDevice, h->stream);
    if (!update && !ev.staged) cudaMemsetAsync(ev.dev, 0, bytes, h->stream);
    DataEnvCtx c{};
    c.tt = h->slab + b * h->tt_stride; c.off = h->d_off; c.bond = h->d_bond + b * (L + 1); c.L = L; c.Q = (int)h->Q; c.ds = ds;
    c.X = d_X; c.env = static_cast<double*>(ev.dev); c.left = left ? 1 : 0; c.t_lo = t_lo; c.t_hi = t_hi;
    CallTimer tm(h);
    TTB_LAUNCH(h, "k_data_env", (k_data_env<<<(unsigned)n, 128, 0, h->stream>>>(c)));
    rc = tm.finish();
  }
  e = ev.close(rc == TTB_OK, h->stream);
  cudaFree(d_X);
  if (rc == TTB_OK && e != cudaSuccess) rc = fail(TTB_ECUDA, "data environments: %s", cudaGetErrorString(e));
  return rc;
}

int ttb_data_environments(ttb_handle h, int64_t b, const int32_t* X, int64_t n, int left, double* env) {
  if (!h) return fail(TTB_EINVAL, "null argument");
  return data_env(h, b, X, n, left, env, 0, (int)h->L, false);
}

int ttb_data_env_update(ttb_handle h, int64_t b, const int32_t* X, int64_t n, int64_t t, int left, double* env) {
  if (!h) return fail(TTB_EINVAL, "null argument");
  if (t < 0 || t >= h->L) return fail(TTB_EINVAL, "site %lld outside the train", (long long)t);
  return data_env(h, b, X, n, left, env, (int)t, (int)t + 1, true);
}

}  // extern "C"
