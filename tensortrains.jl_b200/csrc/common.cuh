// common.cuh -- handle layout, error plumbing and small device helpers shared by the ttb200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ttb200.h"

namespace ttb {

extern thread_local std::string g_err;
int fail(int code, const char* fmt, ...);

#define TTB_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return ttb::fail(TTB_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
  } while (0)

#define TTB_TRY(call)            \
  do {                           \
    int rc__ = (call);           \
    if (rc__ != TTB_OK) return rc__; \
  } while (0)

// device status bits (ttb_tt::d_status)
enum { ST_NEGATIVE = 1, ST_NONFINITE = 2, ST_NOCONV = 4 };

// per-train state of the sweep step in flight
// Two slots (step parity): thin-Q formation of step t runs on a side stream while step t+1 is already
// being factorised, so nothing a step's kernels read may be overwritten by the next step.
struct StepState {
  int mprime[2];                     // retained rank of the step (written by the rank rule)
  unsigned long long maxabs_bits[2]; // max|D| of the absorb output, as ordered bits
  int p[2], n[2], N[2];              // dimensions of the step, snapshotted from the bond array
  double scale_in[2];                // lazy rescale: factor the step applies to its input X (1/max|D| of the previous step)
  int rot[2];                        // cluster Jacobi: rotations of the sweep in flight (ping-pong), all CTAs
  int ticket;                        // absorb: CTAs of the train that have finished (the last one commits the step)
};

// grow-only device scratch kept on the handle between calls (cudaFree of a few hundred MB costs up to
// hundreds of ms of wall clock; the sampler needs the same buffers call after call)
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Workspace {
  DevBuf samp[10];                    // sampler scratch (offsets, G, uniforms, x, p, hist, sum log p, packed A, packed G)
  DevBuf scr[6];                      // twovar_marginals (pair offsets, out, scratch, GT, Tm) and dot scratch
  DevBuf scr_trace, scr_toff;         // two-phase environment pass: per-site traces and their offset table
  DevBuf mps[10];                     // MPS (Born machine) environments, sandwiches and item tables (mps.cu)
  // sweep
  double* X[2] = {nullptr, nullptr};  // folded site / absorb output, p_max x n_max col-major
  int64_t x_stride = 0;
  double* T = nullptr;                // compact-WY T factors, n_max x bw
  int64_t t_stride = 0;
  double* Bv = nullptr;               // Jacobi vectors [nv][ldv]  (rows of R or of X)
  double* Jv = nullptr;               // accumulated rotations, vector-major [nv][ldj]
  int64_t bv_stride = 0, jv_stride = 0;
  int ldv = 0, ldj = 0;
  int* perm = nullptr;                // [n_max] order of singular values (descending)
  double* sigma = nullptr;            // [n_max]
  StepState* st = nullptr;            // [batch]
  double* stat_err = nullptr;         // [batch][L+1]
  int* stat_rank = nullptr;           // [batch][L+1]
  double* spectrum = nullptr;         // [batch][L+1][n_max] (optional)
  int p_max = 0, n_max = 0, bw = 0;
  bool sweep_ready = false;
  // environments
  double* envL = nullptr;             // [batch][env_stride]
  double* envR = nullptr;
  int64_t env_stride = 0;
  double* env_log = nullptr;          // [batch] log|z| of the last accumulate
  int* env_sign = nullptr;            // [batch]
  bool env_ready = false;
};

}  // namespace ttb

namespace ttb {
struct ProfEvent { const char* name; cudaEvent_t a, b; };
}

struct ttb_tt {
  bool prof_on = false;
  std::vector<ttb::ProfEvent> prof;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;            // side stream: thin-Q formation overlaps the next factorisation
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_qr[2] = {nullptr, nullptr}, ev_aq[2] = {nullptr, nullptr};
  // rank feedback of truncating sweeps: k_commit writes the retained rank of every train into a ring of
  // mapped pinned slots; the host enqueues at most two steps ahead of the last slot it has seen complete
  // (cudaEventQuery, never a blocking sync) and sizes later launches from it
  static constexpr int kHintRing = 4;
  cudaEvent_t ev_hint[kHintRing] = {nullptr, nullptr, nullptr, nullptr};
  int* hint_host = nullptr;    // [kHintRing][batch], cudaHostAllocMapped
  int* hint_dev = nullptr;     // device alias of hint_host
  // asynchronous upload (ttb_upload_all_async): chunks are copied on their own stream in REVERSE site order
  // with one event each; a right sweep waits chunk by chunk (it starts at the last site), every other
  // entry point joins all of them first
  cudaStream_t stream_up = nullptr;
  std::vector<cudaEvent_t> up_events;
  std::vector<int> up_lo;      // lowest site present once chunk i has landed
  int up_n = 0, up_waited = 0;
  int64_t L = 0, Q = 0, batch = 0;
  int periodic = 0;
  std::vector<int64_t> bond0;  // original (capacity) bonds, L+1
  std::vector<int64_t> off;    // per-site offset in doubles, L+1 entries (off[L] = used length)
  int64_t tt_stride = 0;       // doubles per train
  int64_t dmax = 0;            // largest original bond
  double* slab = nullptr;      // batch * tt_stride
  int64_t* d_off = nullptr;    // device copy of off
  int* d_bond = nullptr;       // [batch][L+1] current bonds (device)
  std::vector<int> h_bond;     // host mirror
  double* d_logz = nullptr;    // [batch]
  int* d_signz = nullptr;      // [batch]
  int* d_status = nullptr;     // [1]
  ttb::Workspace ws;
  bool record_spectrum = false;
  bool keep_stats = false;     // second sweep of orthogonalize_center!: per-bond stats / spectra of the first one stay
  double last_ms = 0.0;
  int64_t last_launches = 0;
  int64_t launches = 0;        // running counter inside a call
};

// accumulate_L / accumulate_R result kept on the device for the l= / r= / rz= keyword arguments of marginals,
// twovar_marginals and sample (src/abstract_tensor_train.jl:208-209, 231-233, 335-336)
struct ttb_env {
  int device = 0;
  int left = 0, normalize = 1;
  int64_t L = 0, Q = 0, batch = 0, env_stride = 0;
  int periodic = 0;
  std::vector<int64_t> bond0;   // capacity bonds of the train it was computed on (they fix the slot offsets)
  std::vector<int> bonds;       // current bonds of every train at that time
  double* env = nullptr;        // [batch][env_stride], laid out like Workspace::envL / envR
  double* logabs = nullptr;     // [batch]
  int* sign = nullptr;          // [batch]
};

namespace ttb {

// RAII-ish timing bracket used by every entry point that launches kernels
struct CallTimer {
  ttb_tt* h;
  explicit CallTimer(ttb_tt* h_) : h(h_) {
    h->launches = 0;
    cudaEventRecord(h->ev0, h->stream);
  }
  int finish() {
    cudaEventRecord(h->ev1, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return fail(TTB_ECUDA, "stream sync: %s", cudaGetErrorString(e));
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(TTB_ECUDA, "kernel: %s", cudaGetErrorString(e));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->last_ms = ms;
    h->last_launches = h->launches;
    return TTB_OK;
  }
};

int create_handle(int64_t L, const int64_t* bond, int64_t Q, int periodic, int64_t batch, ttb_tt** out, bool check_boundary);
int check_handle(ttb_tt* h);      // validates, selects the device, joins pending asynchronous uploads
int check_handle_raw(ttb_tt* h);  // same without the join (sweeps that consume the upload chunk by chunk)
int join_uploads(ttb_tt* h, int need_site);  // main stream waits for the chunks holding sites >= need_site (-1: all)
void prof_pre(ttb_tt* h, const char* name);
void prof_post(ttb_tt* h);
// every kernel launch goes through this: counts the launch and, in profiling mode, brackets it with
// CUDA events on the handle's stream (per-kernel device time for bench.py's roofline)
#define TTB_LAUNCH(h, name, ...)     \
  do {                               \
    ttb::prof_pre((h), (name));      \
    __VA_ARGS__;                     \
    ttb::prof_post((h));             \
    (h)->launches++;                 \
  } while (0)
int sync_bonds_to_host(ttb_tt* h);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is PER DEVICE: a process-wide once-flag leaves the second
// device of a process without it ("invalid argument" at launch).  True exactly once per (device, slot),
// under a lock held while the caller's attribute calls run (see TTB_ONCE_PER_DEVICE).
struct OnceGuard {
  std::unique_lock<std::mutex> lk;
  bool first;
  OnceGuard(int dev, int slot) : lk(mu()), first(false) {
    static unsigned seen[64] = {};
    const unsigned bit = 1u << slot;
    unsigned& w = seen[dev & 63];
    first = !(w & bit);
    w |= bit;
  }
  static std::mutex& mu() { static std::mutex m; return m; }
};
#define TTB_ONCE_PER_DEVICE(h, slot) if (ttb::OnceGuard og__((h)->device, (slot)); og__.first)
int ensure_sweep_ws(ttb_tt* h);
int ensure_env_ws(ttb_tt* h);
int read_status(ttb_tt* h, int* st);  // reads and clears the device status word

// ------------------------------------------------------------------------------------------
// programmatic dependent launch.  A sweep is a chain of ~20 short dependent kernels per site; the
// stream-order gap between two of them is ~4 us on B200, ~1.7 us when the next grid is already resident
// and parked in griddepcontrol.wait (profiles/micro_pdl.cu).  Contract: EVERY kernel launched through
// launch_pdl -- and every kernel that can sit between two of them in a stream -- starts with pdl_enter()
// in all threads, before it touches memory and before any early return; completion of a grid then implies
// completion of all its predecessors.  On a kernel launched the ordinary way both instructions are no-ops.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum; `red` must hold >= 33 doubles; all threads get the result
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = lane < nw ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ double block_max(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = lane < nw ? red[lane] : 0.0;
    t = warp_max(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
// 1/sqrt(a) for a normal, positive double: MUFU.RSQ64H seed (2^-22) + two Newton iterations.  No special
// cases (callers range-check): ~9 instructions instead of the library's ~25 on the Jacobi critical path.
__device__ __forceinline__ double rsqrt_nr(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double t = a * y;
  double e = fma(-t, y, 1.0);
  y = fma(0.5 * y, e, y);
  t = a * y;
  e = fma(-t, y, 1.0);
  y = fma(0.5 * y, e, y);
  return y;
}
// ---- Philox4x32-10 ---------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                             uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
  c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
}
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t sample, uint32_t site) {
  uint32_t c0 = (uint32_t)sample, c1 = (uint32_t)(sample >> 32), c2 = site, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const uint64_t bits = ((uint64_t)c0 << 21) ^ ((uint64_t)c1 >> 11);
  return (double)bits * (1.0 / 9007199254740992.0);
}

// One-sided Jacobi, stopping rule.  A sweep is the last one when it made no rotation with
// (b_i . b_j)^2 > kJacobiBig2 |b_i|^2 |b_j|^2, i.e. every rotation it DID make had a relative off-diagonal below
// 3.2e-8: convergence is quadratic, so what such a sweep leaves behind is below 1e-15 -- the rotation threshold
// itself (sqrt(n) eps) -- and the confirming sweep that finds nothing to rotate (1/8 of the SVD's time: measured
// rotation counts per sweep e.g. 2016 2015 1952 1770 1316 892 63 0, largest relative off-diagonal of the last
// rotating sweep 2e-13 .. 1e-7, profiles/jac_proto.py) is not run.  Singular values are second order in the
// residual off-diagonal: their error stays at the rounding level.
constexpr double kJacobiBig2 = 1e-15;

// max-abs with NaN propagation (fmax drops NaN, the reference's maximum(abs, .) keeps it)
__device__ __forceinline__ double absmax_nan(double acc, double v) {
  double a = fabs(v);
  return (a != a || acc != acc) ? __longlong_as_double(0x7ff8000000000000LL) : fmax(acc, a);
}

}  // namespace ttb
