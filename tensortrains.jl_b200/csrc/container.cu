// container.cu -- device container behind TensorTrain / PeriodicTensorTrain
// (reference: src/tensor_train.jl:8-24, src/periodic_tensor_train.jl:8-27,
//  check_bond_dims src/abstract_tensor_train.jl:40-52).
#include <algorithm>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace ttb {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

int check_handle_raw(ttb_tt* h) {
  if (!h) return fail(TTB_EINVAL, "null handle");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return fail(TTB_ECUDA, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e));
  return TTB_OK;
}

int join_uploads(ttb_tt* h, int need_site) {
  while (h->up_waited < h->up_n &&
         (need_site < 0 || h->up_waited == 0 || h->up_lo[h->up_waited - 1] > need_site)) {
    cudaError_t e = cudaStreamWaitEvent(h->stream, h->up_events[h->up_waited], 0);
    if (e != cudaSuccess) return fail(TTB_ECUDA, "upload join: %s", cudaGetErrorString(e));
    ++h->up_waited;
  }
  if (h->up_waited == h->up_n) h->up_n = h->up_waited = 0;
  return TTB_OK;
}

int check_handle(ttb_tt* h) {
  TTB_TRY(check_handle_raw(h));
  if (h->up_n) TTB_TRY(join_uploads(h, -1));
  return TTB_OK;
}

void prof_pre(ttb_tt* h, const char* name) {
  if (!h->prof_on) return;
  ProfEvent e;
  e.name = name;
  cudaEventCreate(&e.a);
  cudaEventCreate(&e.b);
  cudaEventRecord(e.a, h->stream);
  h->prof.push_back(e);
}
void prof_post(ttb_tt* h) {
  if (!h->prof_on) return;
  cudaEventRecord(h->prof.back().b, h->stream);
}

int sync_bonds_to_host(ttb_tt* h) {
  TTB_CUDA(cudaMemcpyAsync(h->h_bond.data(), h->d_bond, sizeof(int) * h->h_bond.size(), cudaMemcpyDeviceToHost,
                           h->stream));
  TTB_CUDA(cudaStreamSynchronize(h->stream));
  return TTB_OK;
}

int read_status(ttb_tt* h, int* st) {
  TTB_CUDA(cudaMemcpyAsync(st, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  TTB_CUDA(cudaStreamSynchronize(h->stream));
  if (*st) TTB_CUDA(cudaMemsetAsync(h->d_status, 0, sizeof(int), h->stream));
  return TTB_OK;
}

static void free_ws(ttb_tt* h) {
  Workspace& w = h->ws;
  cudaFree(w.X[0]); cudaFree(w.X[1]); cudaFree(w.T); cudaFree(w.Bv); cudaFree(w.Jv);
  cudaFree(w.perm); cudaFree(w.sigma); cudaFree(w.st); cudaFree(w.stat_err); cudaFree(w.stat_rank);
  cudaFree(w.spectrum); cudaFree(w.envL); cudaFree(w.envR); cudaFree(w.env_log); cudaFree(w.env_sign);
  for (DevBuf& d : w.samp) d.release();
  for (DevBuf& d : w.scr) d.release();
  for (DevBuf& d : w.mps) d.release();
  w.scr_trace.release(); w.scr_toff.release();
  w = Workspace();
}

}  // namespace ttb

using namespace ttb;

extern "C" {

int ttb_version(void) { return 100; }
const char* ttb_last_error(void) { return g_err.c_str(); }

int ttb_device_count(int* n) {
  if (!n) return fail(TTB_EINVAL, "null argument");
  TTB_CUDA(cudaGetDeviceCount(n));
  return TTB_OK;
}
int ttb_set_device(int dev) {
  TTB_CUDA(cudaSetDevice(dev));
  return TTB_OK;
}
int ttb_host_alloc(void** p, size_t bytes) {
  if (!p) return fail(TTB_EINVAL, "null argument");
  TTB_CUDA(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
  return TTB_OK;
}
int ttb_host_free(void* p) {
  TTB_CUDA(cudaFreeHost(p));
  return TTB_OK;
}

int ttb_create(int64_t L, const int64_t* bond, int64_t Q, int periodic, int64_t batch, ttb_handle* out) {
  return ttb::create_handle(L, bond, Q, periodic, batch, out, true);
}

}  // extern "C"

// check_boundary = false: an open chain SEGMENT (first / last bond > 1), the temporary two-site train of
// ttb_split_tensor (twosite.cu); everything else is the public constructor
int ttb::create_handle(int64_t L, const int64_t* bond, int64_t Q, int periodic, int64_t batch, ttb_handle* out, bool check_boundary) {
  if (!bond || !out) return fail(TTB_EINVAL, "null argument");
  if (L < 1 || Q < 1 || batch < 1) return fail(TTB_EINVAL, "L, Q and batch must be >= 1 (got %lld, %lld, %lld)",
                                               (long long)L, (long long)Q, (long long)batch);
  for (int64_t t = 0; t <= L; ++t)
    if (bond[t] < 1) return fail(TTB_ESHAPE, "bond %lld is %lld, must be >= 1", (long long)t, (long long)bond[t]);
  if (check_boundary && !periodic && !(bond[0] == 1 && bond[L] == 1))
    return fail(TTB_ESHAPE, "First matrix must have 1 row, last matrix must have 1 column");
  if (periodic && bond[L] != bond[0])
    return fail(TTB_ESHAPE, "Matrix indices for matrix product non compatible (bond[L]=%lld != bond[0]=%lld)",
                (long long)bond[L], (long long)bond[0]);
  ttb_tt* h = new ttb_tt();
  cudaGetDevice(&h->device);
  h->L = L; h->Q = Q; h->batch = batch; h->periodic = periodic ? 1 : 0;
  h->bond0.assign(bond, bond + L + 1);
  h->off.resize(L + 1);
  int64_t o = 0;
  h->dmax = 0;
  for (int64_t t = 0; t < L; ++t) {
    h->off[t] = o;
    int64_t n = bond[t] * bond[t + 1] * Q;
    o += (n + 15) & ~int64_t(15);  // keep every site 128-byte aligned
    h->dmax = std::max(h->dmax, std::max(bond[t], bond[t + 1]));
  }
  h->off[L] = o;
  h->tt_stride = o;
  if (h->dmax > (1 << 20) || h->dmax * Q > (int64_t(1) << 28)) {
    delete h;
    return fail(TTB_EUNSUPPORTED, "bond*Q too large");
  }
  auto bail = [&](cudaError_t e, const char* what) {
    int rc = fail(e == cudaErrorMemoryAllocation ? TTB_ENOMEM : TTB_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    ttb_destroy(h);
    return rc;
  };
  cudaError_t e;
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
  if ((e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
  if ((e = cudaStreamCreateWithFlags(&h->stream_up, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
  for (int i = 0; i < 2; ++i) {
    if ((e = cudaEventCreateWithFlags(&h->ev_qr[i], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    if ((e = cudaEventCreateWithFlags(&h->ev_aq[i], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
  }
  for (int i = 0; i < ttb_tt::kHintRing; ++i)
    if ((e = cudaEventCreateWithFlags(&h->ev_hint[i], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
  if ((e = cudaHostAlloc((void**)&h->hint_host, sizeof(int) * ttb_tt::kHintRing * batch, cudaHostAllocMapped)) != cudaSuccess)
    return bail(e, "rank feedback buffer");
  if ((e = cudaHostGetDevicePointer((void**)&h->hint_dev, h->hint_host, 0)) != cudaSuccess) return bail(e, "rank feedback buffer");
  if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess) return bail(e, "event");
  if ((e = cudaEventCreate(&h->ev1)) != cudaSuccess) return bail(e, "event");
  if ((e = cudaMalloc(&h->slab, sizeof(double) * batch * h->tt_stride)) != cudaSuccess) return bail(e, "slab");
  if ((e = cudaMemsetAsync(h->slab, 0, sizeof(double) * batch * h->tt_stride, h->stream)) != cudaSuccess)
    return bail(e, "memset");
  if ((e = cudaMalloc(&h->d_off, sizeof(int64_t) * (L + 1))) != cudaSuccess) return bail(e, "off");
  if ((e = cudaMalloc(&h->d_bond, sizeof(int) * batch * (L + 1))) != cudaSuccess) return bail(e, "bond");
  if ((e = cudaMalloc(&h->d_logz, sizeof(double) * batch)) != cudaSuccess) return bail(e, "logz");
  if ((e = cudaMalloc(&h->d_signz, sizeof(int) * batch)) != cudaSuccess) return bail(e, "signz");
  if ((e = cudaMalloc(&h->d_status, sizeof(int))) != cudaSuccess) return bail(e, "status");
  h->h_bond.resize(batch * (L + 1));
  for (int64_t b = 0; b < batch; ++b)
    for (int64_t t = 0; t <= L; ++t) h->h_bond[b * (L + 1) + t] = (int)bond[t];
  std::vector<int> ones(batch, 1);
  cudaMemcpyAsync(h->d_off, h->off.data(), sizeof(int64_t) * (L + 1), cudaMemcpyHostToDevice, h->stream);
  cudaMemcpyAsync(h->d_bond, h->h_bond.data(), sizeof(int) * h->h_bond.size(), cudaMemcpyHostToDevice, h->stream);
  cudaMemsetAsync(h->d_logz, 0, sizeof(double) * batch, h->stream);
  cudaMemcpyAsync(h->d_signz, ones.data(), sizeof(int) * batch, cudaMemcpyHostToDevice, h->stream);
  cudaMemsetAsync(h->d_status, 0, sizeof(int), h->stream);
  if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return bail(e, "init");
  *out = h;
  return TTB_OK;
}

extern "C" {

int ttb_destroy(ttb_handle h) {
  if (!h) return TTB_OK;
  cudaSetDevice(h->device);
  if (h->stream_up) cudaStreamSynchronize(h->stream_up);
  if (h->stream2) cudaStreamSynchronize(h->stream2);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (cudaEvent_t ev : h->up_events) cudaEventDestroy(ev);
  if (h->stream_up) cudaStreamDestroy(h->stream_up);
  free_ws(h);
  cudaFree(h->slab); cudaFree(h->d_off); cudaFree(h->d_bond); cudaFree(h->d_logz); cudaFree(h->d_signz);
  cudaFree(h->d_status);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (int i = 0; i < ttb_tt::kHintRing; ++i)
    if (h->ev_hint[i]) cudaEventDestroy(h->ev_hint[i]);
  if (h->hint_host) cudaFreeHost(h->hint_host);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_qr[i]) cudaEventDestroy(h->ev_qr[i]);
    if (h->ev_aq[i]) cudaEventDestroy(h->ev_aq[i]);
  }
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return TTB_OK;
}

int ttb_get_capacity_bonds(ttb_handle h, int64_t* out) {
  TTB_TRY(check_handle(h));
  if (!out) return fail(TTB_EINVAL, "null argument");
  for (int64_t t = 0; t <= h->L; ++t) out[t] = h->bond0[t];
  return TTB_OK;
}

int ttb_clone(ttb_handle h, ttb_handle* out) {
  TTB_TRY(check_handle(h));
  if (!out) return fail(TTB_EINVAL, "null argument");
  ttb_handle c = nullptr;
  TTB_TRY(ttb_create(h->L, h->bond0.data(), h->Q, h->periodic, h->batch, &c));
  cudaStreamSynchronize(h->stream);
  cudaMemcpyAsync(c->slab, h->slab, sizeof(double) * h->batch * h->tt_stride, cudaMemcpyDeviceToDevice, c->stream);
  cudaMemcpyAsync(c->d_bond, h->d_bond, sizeof(int) * h->h_bond.size(), cudaMemcpyDeviceToDevice, c->stream);
  cudaMemcpyAsync(c->d_logz, h->d_logz, sizeof(double) * h->batch, cudaMemcpyDeviceToDevice, c->stream);
  cudaMemcpyAsync(c->d_signz, h->d_signz, sizeof(int) * h->batch, cudaMemcpyDeviceToDevice, c->stream);
  c->h_bond = h->h_bond;
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    ttb_destroy(c);
    return fail(TTB_ECUDA, "clone: %s", cudaGetErrorString(e));
  }
  *out = c;
  return TTB_OK;
}

// ---- A + B / A - B (src/tensor_train.jl:240-261, src/periodic_tensor_train.jl:62-78,
//      src/abstract_tensor_train.jl:315-322): block assembly on the device, a pure HBM copy -------------
}  // extern "C"
namespace ttb {
struct ComposeSide { const double* slab; int64_t tt_stride; const int64_t* off; const int* bond; const double* logz; const int* signz; };
struct ComposeCtx {
  ComposeSide a, b;
  double* slab; int64_t tt_stride; const int64_t* off; int* bond; double* logz; int* signz;
  int L, L1, Q, periodic, sign;
};

// one CTA column per (site, train); the first CTA of a site also writes the site's bond entries and z
__global__ void k_compose(ComposeCtx c) {
  const int t = blockIdx.y, bt = blockIdx.z;
  const int* ba = c.a.bond + (int64_t)bt * c.L1;
  const int* bb = c.b.bond + (int64_t)bt * c.L1;
  const int al = ba[t], ar = ba[t + 1], bl = bb[t], br = bb[t + 1];
  const bool open = !c.periodic;
  const bool first = open && t == 0, last = open && t == c.L - 1;
  const int ol = first ? 1 : al + bl, orr = last ? 1 : ar + br;
  // open trains: z = min(|zA|, |zB|), za = z / zA, zb = f(z / zB) scale the first site; periodic trains
  // ignore z (the reference builds the sum with z = 1) and f acts on the first site's B block
  double fa = 1.0, fb = 1.0;
  if (open) {
    const double la = c.a.logz[bt], lb = c.b.logz[bt], lz = fmin(la, lb);
    if (t == 0) { fa = exp(lz - la) * c.a.signz[bt]; fb = exp(lz - lb) * c.b.signz[bt] * c.sign; }
    if (blockIdx.x == 0 && threadIdx.x == 0 && t == 0) { c.logz[bt] = lz; c.signz[bt] = 1; }
  } else {
    if (t == 0) fb = (double)c.sign;
    if (blockIdx.x == 0 && threadIdx.x == 0 && t == 0) { c.logz[bt] = 0.0; c.signz[bt] = 1; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int* bo = c.bond + (int64_t)bt * c.L1;
    bo[t] = ol;
    if (t == c.L - 1) bo[c.L] = orr;
  }
  const double* A = c.a.slab + (int64_t)bt * c.a.tt_stride + c.a.off[t];
  const double* B = c.b.slab + (int64_t)bt * c.b.tt_stride + c.b.off[t];
  double* O = c.slab + (int64_t)bt * c.tt_stride + c.off[t];
  const int64_t n = (int64_t)ol * orr * c.Q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i % ol), cc = (int)((i / ol) % orr), x = (int)(i / ((int64_t)ol * orr));
    // row / column inside the A block or the B block (the boundary sites of an open train concatenate)
    const bool ra = first || r < al, ca = last || cc < ar;
    const int rb = first ? 0 : r - al, cb = last ? 0 : cc - ar;
    double v = 0.0;
    if (first)      v = cc < ar ? fa * A[(int64_t)al * (cc + (int64_t)ar * x)] : fb * B[(int64_t)bl * (cb + (int64_t)br * x)];
    else if (last)  v = r < al ? A[r + (int64_t)al * ((int64_t)ar * x)] : B[rb + (int64_t)bl * ((int64_t)br * x)];
    else if (ra && ca)   v = fa * A[r + (int64_t)al * (cc + (int64_t)ar * x)];
    else if (!ra && !ca) v = fb * B[rb + (int64_t)bl * (cb + (int64_t)br * x)];
    O[i] = v;
  }
}
}  // namespace ttb
extern "C" {

int ttb_compose(ttb_handle a, ttb_handle b, int sign, ttb_handle* out) {
  TTB_TRY(check_handle(a));
  TTB_TRY(check_handle(b));
  if (!out || (sign != 1 && sign != -1)) return fail(TTB_EINVAL, "bad argument");
  if (a->L != b->L) return fail(TTB_ESHAPE, "Tensor Trains must have the same length, got %lld and %lld", (long long)a->L, (long long)b->L);
  if (a->Q != b->Q) return fail(TTB_ESHAPE, "Tensor Trains must have the same physical dimensions, got %lld and %lld", (long long)a->Q, (long long)b->Q);
  if (a->periodic != b->periodic || a->batch != b->batch || a->device != b->device)
    return fail(TTB_EINVAL, "trains of different kind, batch size or device");
  const int64_t L = a->L, B = a->batch;
  if (!a->periodic && L < 2) return fail(TTB_ESHAPE, "First matrix must have 1 row, last matrix must have 1 column");
  // capacity = the largest sum over the batch; current dims = the per-train sums (written by the kernel)
  std::vector<int64_t> cap(L + 1, 1);
  std::vector<int> hb((size_t)B * (L + 1));
  for (int64_t bt = 0; bt < B; ++bt)
    for (int64_t t = 0; t <= L; ++t) {
      const bool edge = !a->periodic && (t == 0 || t == L);
      const int v = edge ? 1 : a->h_bond[bt * (L + 1) + t] + b->h_bond[bt * (L + 1) + t];
      hb[bt * (L + 1) + t] = v;
      cap[t] = std::max<int64_t>(cap[t], v);
    }
  ttb_handle c = nullptr;
  TTB_TRY(ttb_create(L, cap.data(), a->Q, a->periodic, B, &c));
  c->h_bond = hb;
  ttb::ComposeCtx cc;
  cc.a = {a->slab, a->tt_stride, a->d_off, a->d_bond, a->d_logz, a->d_signz};
  cc.b = {b->slab, b->tt_stride, b->d_off, b->d_bond, b->d_logz, b->d_signz};
  cc.slab = c->slab; cc.tt_stride = c->tt_stride; cc.off = c->d_off; cc.bond = c->d_bond; cc.logz = c->d_logz; cc.signz = c->d_signz;
  cc.L = (int)L; cc.L1 = (int)L + 1; cc.Q = (int)a->Q; cc.periodic = a->periodic; cc.sign = sign;
  cudaStreamSynchronize(a->stream);
  cudaStreamSynchronize(b->stream);
  int64_t maxsite = 0;
  for (int64_t t = 0; t < L; ++t) maxsite = std::max(maxsite, cap[t] * cap[t + 1] * a->Q);
  const dim3 g((unsigned)std::min<int64_t>((maxsite + 1023) / 1024, 256), (unsigned)L, (unsigned)B);
  ttb::CallTimer tm(c);
  TTB_LAUNCH(c, "k_compose", ttb::k_compose<<<g, 256, 0, c->stream>>>(cc));
  int rc = tm.finish();
  if (rc != TTB_OK) { ttb_destroy(c); return rc; }
  *out = c;
  return TTB_OK;
}

int ttb_info(ttb_handle h, int64_t* L, int64_t* Q, int* periodic, int64_t* batch) {
  if (!h) return fail(TTB_EINVAL, "null handle");
  if (L) *L = h->L;
  if (Q) *Q = h->Q;
  if (periodic) *periodic = h->periodic;
  if (batch) *batch = h->batch;
  return TTB_OK;
}

static int check_bt(ttb_tt* h, int64_t b, int64_t t) {
  if (b < 0 || b >= h->batch) return fail(TTB_EINVAL, "train index %lld out of range 0:%lld", (long long)b, (long long)h->batch - 1);
  if (t < 0 || t >= h->L) return fail(TTB_EINVAL, "site index %lld out of range 0:%lld", (long long)t, (long long)h->L - 1);
  return TTB_OK;
}

int ttb_get_bonds(ttb_handle h, int64_t b, int64_t* out) {
  TTB_TRY(check_handle(h));
  TTB_TRY(check_bt(h, b, 0));
  if (!out) return fail(TTB_EINVAL, "null argument");
  for (int64_t t = 0; t <= h->L; ++t) out[t] = h->h_bond[b * (h->L + 1) + t];
  return TTB_OK;
}

static int64_t site_len(ttb_tt* h, int64_t b, int64_t t) {
  const int* bd = &h->h_bond[b * (h->L + 1)];
  return (int64_t)bd[t] * bd[t + 1] * h->Q;
}

int ttb_nparams(ttb_handle h, int64_t b, int64_t* n) {
  TTB_TRY(check_handle(h));
  TTB_TRY(check_bt(h, b, 0));
  int64_t s = 0;
  for (int64_t t = 0; t < h->L; ++t) s += site_len(h, b, t);
  *n = s;
  return TTB_OK;
}

int ttb_upload_site(ttb_handle h, int64_t b, int64_t t, const double* host) {
  TTB_TRY(check_handle(h));
  TTB_TRY(check_bt(h, b, t));
  TTB_CUDA(cudaMemcpyAsync(h->slab + b * h->tt_stride + h->off[t], host, sizeof(double) * site_len(h, b, t),
                           cudaMemcpyHostToDevice, h->stream));
  TTB_CUDA(cudaStreamSynchronize(h->stream));
  return TTB_OK;
}

int ttb_download_site(ttb_handle h, int64_t b, int64_t t, double* host) {
  TTB_TRY(check_handle(h));
  TTB_TRY(check_bt(h, b, t));
  TTB_CUDA(cudaMemcpyAsync(host, h->slab + b * h->tt_stride + h->off[t], sizeof(double) * site_len(h, b, t),
                           cudaMemcpyDeviceToHost, h->stream));
  TTB_CUDA(cudaStreamSynchronize(h->stream));
  return TTB_OK;
}

// Copies between a packed host buffer and the slab, merging device-contiguous runs of sites into one
// cudaMemcpyAsync (a site is followed contiguously when its current length equals its padded capacity).
static int copy_all(ttb_tt* h, double* host, bool to_device) {
  double* p = host;
  int64_t run_dev = -1, run_len = 0;
  double* run_host = nullptr;
  auto flush = [&]() -> cudaError_t {
    if (run_len == 0) return cudaSuccess;
    cudaError_t e = to_device ? cudaMemcpyAsync(h->slab + run_dev, run_host, sizeof(double) * run_len,
                                                cudaMemcpyHostToDevice, h->stream)
                              : cudaMemcpyAsync(run_host, h->slab + run_dev, sizeof(double) * run_len,
                                                cudaMemcpyDeviceToHost, h->stream);
    run_len = 0;
    return e;
  };
  for (int64_t b = 0; b < h->batch; ++b)
    for (int64_t t = 0; t < h->L; ++t) {
      int64_t n = site_len(h, b, t);
      int64_t dev = b * h->tt_stride + h->off[t];
      if (run_len > 0 && dev == run_dev + run_len) {
        run_len += n;
      } else {
        TTB_CUDA(flush());
        run_dev = dev; run_host = p; run_len = n;
      }
      p += n;
    }
  TTB_CUDA(flush());
  TTB_CUDA(cudaStreamSynchronize(h->stream));
  return TTB_OK;
}

int ttb_upload_all(ttb_handle h, const double* host) {
  TTB_TRY(check_handle(h));
  if (!host) return fail(TTB_EINVAL, "null argument");
  return copy_all(h, const_cast<double*>(host), true);
}

// Asynchronous form of ttb_upload_all for a train at its ORIGINAL bond dimensions (fresh or ttb_reset):
// returns once the copies are enqueued; `host` (pinned for a real overlap) must stay valid until the next
// call on this handle that returns results.  Sites travel last-to-first in >= 16 MB chunks so that a right
// sweep (orthogonalize_right!, compress!) starts as soon as its first sites have landed.
int ttb_upload_all_async(ttb_handle h, const double* host) {
  TTB_TRY(check_handle(h));
  if (!host) return fail(TTB_EINVAL, "null argument");
  TTB_TRY(sync_bonds_to_host(h));
  const int L = (int)h->L, L1 = L + 1;
  bool original = true;
  for (int64_t b = 0; b < h->batch && original; ++b)
    for (int t = 0; t < L1; ++t)
      if (h->h_bond[b * L1 + t] != (int)h->bond0[t]) { original = false; break; }
  if (!original || h->batch != 1) return copy_all(h, const_cast<double*>(host), true);
  std::vector<int64_t> hoff(L1, 0);
  for (int t = 0; t < L; ++t) hoff[t + 1] = hoff[t] + site_len(h, 0, t);
  const int64_t chunk = int64_t(16) << 20;  // bytes
  h->up_lo.clear(); h->up_n = 0; h->up_waited = 0;
  int hi = L;
  while (hi > 0) {
    int lo = hi;
    while (lo > 0 && (hoff[hi] - hoff[lo]) * 8 < chunk) --lo;
    // one copy per run that is contiguous on the device (sites are 128-byte aligned there, packed on the host)
    int t = lo;
    while (t < hi) {
      int u = t + 1;
      while (u < hi && h->off[u] == h->off[u - 1] + site_len(h, 0, u - 1)) ++u;
      TTB_CUDA(cudaMemcpyAsync(h->slab + h->off[t], host + hoff[t], sizeof(double) * (hoff[u] - hoff[t]),
                               cudaMemcpyHostToDevice, h->stream_up));
      t = u;
    }
    if ((int)h->up_events.size() <= h->up_n) {
      cudaEvent_t ev;
      TTB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      h->up_events.push_back(ev);
    }
    TTB_CUDA(cudaEventRecord(h->up_events[h->up_n], h->stream_up));
    h->up_lo.push_back(lo);
    ++h->up_n;
    hi = lo;
  }
  return TTB_OK;
}

// Back to the original (capacity) bond dimensions and z = 1: the handle can take a new train of the same
// shape without reallocating the slab or the workspaces.
int ttb_reset(ttb_handle h) {
  TTB_TRY(check_handle(h));
  const int64_t L1 = h->L + 1;
  for (int64_t b = 0; b < h->batch; ++b)
    for (int64_t t = 0; t < L1; ++t) h->h_bond[b * L1 + t] = (int)h->bond0[t];
  std::vector<int> ones(h->batch, 1);
  TTB_CUDA(cudaMemcpyAsync(h->d_bond, h->h_bond.data(), sizeof(int) * h->h_bond.size(), cudaMemcpyHostToDevice, h->stream));
  TTB_CUDA(cudaMemsetAsync(h->d_logz, 0, sizeof(double) * h->batch, h->stream));
  TTB_CUDA(cudaMemcpyAsync(h->d_signz, ones.data(), sizeof(int) * h->batch, cudaMemcpyHostToDevice, h->stream));
  TTB_CUDA(cudaStreamSynchronize(h->stream));
  return TTB_OK;
}

// release the grow-only scratch the sampler / twovar_marginals / dot keep on the handle between calls
int ttb_trim(ttb_handle h) {
  TTB_TRY(check_handle(h));
  TTB_CUDA(cudaStreamSynchronize(h->stream));
  for (DevBuf& d : h->ws.samp) d.release();
  for (DevBuf& d : h->ws.scr) d.release();
  for (DevBuf& d : h->ws.mps) d.release();
  h->ws.scr_trace.release();
  return TTB_OK;
}

int ttb_download_all(ttb_handle h, double* host) {
  TTB_TRY(check_handle(h));
  if (!host) return fail(TTB_EINVAL, "null argument");
  return copy_all(h, host, false);
}

int ttb_get_z(ttb_handle h, int64_t b, double* logabs, int* sign) {
  TTB_TRY(check_handle(h));
  TTB_TRY(check_bt(h, b, 0));
  TTB_CUDA(cudaMemcpyAsync(logabs, h->d_logz + b, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  TTB_CUDA(cudaMemcpyAsync(sign, h->d_signz + b, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  TTB_CUDA(cudaStreamSynchronize(h->stream));
  return TTB_OK;
}

int ttb_set_z(ttb_handle h, int64_t b, double logabs, int sign) {
  TTB_TRY(check_handle(h));
  TTB_TRY(check_bt(h, b, 0));
  int s = sign < 0 ? -1 : 1;
  TTB_CUDA(cudaMemcpyAsync(h->d_logz + b, &logabs, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  TTB_CUDA(cudaMemcpyAsync(h->d_signz + b, &s, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  TTB_CUDA(cudaStreamSynchronize(h->stream));
  return TTB_OK;
}

// ---- is_in_domain (src/abstract_tensor_train.jl:59-64) and == (:85) --------------------------------------
int ttb_is_in_domain(ttb_handle h, const int32_t* X, int64_t n, int* ok) {
  TTB_TRY(check_handle_raw(h));
  if (!X || !ok || n < 0) return fail(TTB_EINVAL, "bad argument");
  *ok = 1;
  for (int64_t i = 0; i < n * h->L; ++i)
    if (X[i] < 0 || X[i] >= h->Q) { *ok = 0; break; }
  return TTB_OK;
}

}  // extern "C"

namespace ttb {
// isequal on doubles: NaN equals NaN, -0.0 differs from 0.0 (Julia's isequal, used by == on the tensor vectors)
__global__ void k_equal(const double* a, const double* b, int64_t n, int* differ) {
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = a[i], y = b[i];
    const bool same = (x != x && y != y) || __double_as_longlong(x) == __double_as_longlong(y);
    bad |= !same;
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(differ, 1);
}
}  // namespace ttb

extern "C" {

int ttb_equal(ttb_handle a, int64_t ba, ttb_handle b, int64_t bb, int* equal) {
  TTB_TRY(check_handle(a));
  TTB_TRY(check_handle(b));
  if (!equal) return fail(TTB_EINVAL, "null argument");
  if (ba < 0 || ba >= a->batch || bb < 0 || bb >= b->batch) return fail(TTB_EINVAL, "train index out of range");
  if (a->device != b->device) return fail(TTB_EINVAL, "== needs both trains on the same device");
  *equal = 0;
  // different types / lengths / shapes are simply not equal (isequal of the tensor vectors)
  if (a->periodic != b->periodic || a->L != b->L || a->Q != b->Q) return TTB_OK;
  const int L = (int)a->L;
  const int* da = &a->h_bond[ba * (L + 1)];
  const int* db = &b->h_bond[bb * (L + 1)];
  for (int t = 0; t <= L; ++t)
    if (da[t] != db[t]) return TTB_OK;
  int* d_flag = nullptr;
  TTB_CUDA(cudaMalloc(&d_flag, sizeof(int)));
  cudaMemsetAsync(d_flag, 0, sizeof(int), a->stream);
  // b's stream may still be writing its slab: order this stream after it
  cudaEvent_t ev;
  cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  cudaEventRecord(ev, b->stream);
  cudaStreamWaitEvent(a->stream, ev, 0);
  cudaEventDestroy(ev);
  CallTimer tm(a);
  for (int t = 0; t < L; ++t) {
    const int64_t n = (int64_t)da[t] * da[t + 1] * a->Q;
    const unsigned g = (unsigned)std::min<int64_t>((n + 255) / 256, 1184);
    TTB_LAUNCH(a, "k_equal", k_equal<<<g, 256, 0, a->stream>>>(a->slab + ba * a->tt_stride + a->off[t],
                                                               b->slab + bb * b->tt_stride + b->off[t], n, d_flag));
  }
  int rc = tm.finish();
  int differ = 1;
  if (rc == TTB_OK) {
    cudaError_t e = cudaMemcpy(&differ, d_flag, sizeof(int), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(TTB_ECUDA, "==: %s", cudaGetErrorString(e));
  }
  cudaFree(d_flag);
  if (rc == TTB_OK) *equal = differ ? 0 : 1;
  return rc;
}

int ttb_record_spectrum(ttb_handle h, int on) {
  TTB_TRY(check_handle(h));
  h->record_spectrum = on != 0;
  return TTB_OK;
}

int ttb_profile(ttb_handle h, int on) {
  TTB_TRY(check_handle(h));
  for (auto& e : h->prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  h->prof.clear();
  h->prof_on = on != 0;
  return TTB_OK;
}

int ttb_profile_report(ttb_handle h, char* out, size_t cap) {
  TTB_TRY(check_handle(h));
  if (!out || cap == 0) return fail(TTB_EINVAL, "null argument");
  TTB_CUDA(cudaStreamSynchronize(h->stream));
  struct Acc { const char* name; long n; double ms; };
  std::vector<Acc> acc;
  for (auto& e : h->prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e.a, e.b);
    bool found = false;
    for (auto& a : acc)
      if (strcmp(a.name, e.name) == 0) { a.n++; a.ms += ms; found = true; break; }
    if (!found) acc.push_back({e.name, 1, (double)ms});
  }
  std::string s;
  char line[256];
  for (auto& a : acc) {
    snprintf(line, sizeof(line), "%s %ld %.6f\n", a.name, a.n, a.ms);
    s += line;
  }
  snprintf(out, cap, "%s", s.c_str());
  return TTB_OK;
}

int ttb_last_timing(ttb_handle h, double* ms, int64_t* launches) {
  if (!h) return fail(TTB_EINVAL, "null handle");
  if (ms) *ms = h->last_ms;
  if (launches) *launches = h->last_launches;
  return TTB_OK;
}

}  // extern "C"
