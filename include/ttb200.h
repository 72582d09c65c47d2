/* ttb200.h -- C ABI of libttb200.so: the B200 (sm_100a) hot path of TensorTrains.jl.
 *
 * The reference (stecrotti/TensorTrains.jl v0.14.0, pure Julia) has no FFI of its own; the seam is
 * Julia dispatch on the storage type of TensorTrain{F,N,T,Z} / PeriodicTensorTrain{F,N,T,Z}
 * (src/tensor_train.jl:8-12, src/periodic_tensor_train.jl:8-12).  Each entry point below names the
 * reference method it replaces (file:line relative to the reference root).  INTEGRATION.md shows the
 * `ccall` stubs a maintainer adds on the Julia side and the ctypes mirror used in this repo.
 *
 * Conventions
 *  - Float64 only.  Site tensor t (0-based) has dims (d_t, d_{t+1}, Q), Q = prod(q), COLUMN-MAJOR
 *    exactly like the Julia arrays: element (m,n,x) at m + d_t*(n + d_{t+1}*x).
 *  - A handle owns the device memory of `batch` tensor trains that share L, Q and the ORIGINAL bond
 *    dimensions (capacity); current bond dimensions are tracked per train and only ever shrink.
 *  - Host buffers are caller-owned.  Every call is synchronous on return and returns a status code;
 *    nothing throws or aborts across the boundary.  ttb_last_error() gives the message (thread-local).
 *  - z (LogarithmicNumbers.Logarithmic in the reference) crosses the boundary as (log|z|, sign).
 */
#ifndef TTB200_H
#define TTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ttb_tt* ttb_handle;
typedef struct ttb_env* ttb_env_handle; /* a stored accumulate_L / accumulate_R result (see "environment handles") */

enum {
  TTB_OK = 0,
  TTB_EINVAL = 1,       /* ArgumentError in the reference (bad indices / is_orthogonal / arguments) */
  TTB_ESHAPE = 2,       /* ArgumentError: incompatible bond dimensions / boundary bonds != 1       */
  TTB_ENEGATIVE = 3,    /* ErrorException("Cannot sample from a tensor train with negative values") */
  TTB_ENONFINITE = 4,   /* NaN/Inf met where the reference would propagate it                        */
  TTB_ECUDA = 5,        /* CUDA runtime error (message has the CUDA string)                          */
  TTB_ENOMEM = 6,
  TTB_EUNSUPPORTED = 7, /* shape outside what the kernels cover (message says which limit)          */
  TTB_EDOMAIN = 8,      /* DomainError: log of a negative normalization                              */
  TTB_ENOCONV = 9       /* LAPACKException of svd (src/svd_trunc.jl:37): the Jacobi SVD did not converge */
};

/* SVD truncation policy: src/svd_trunc.jl:33-50 (THRESH), :69-77 (BOND), :91-103 (BONDMAX),
 * :123-143 (BONDTHRESH).  The rank rule runs on the device right after the factorisation. */
enum { TTB_TRUNC_THRESH = 0, TTB_TRUNC_BOND = 1, TTB_TRUNC_BONDMAX = 2, TTB_TRUNC_BONDTHRESH = 3 };
typedef struct {
  int32_t kind;
  int32_t _pad;
  double eps;     /* THRESH, BONDTHRESH */
  int64_t mprime; /* BOND, BONDMAX, BONDTHRESH */
} ttb_trunc;

/* is_orthogonal of compress! (src/abstract_tensor_train.jl:261-274) */
enum { TTB_ORTH_NONE = 0, TTB_ORTH_LEFT = 1, TTB_ORTH_RIGHT = 2 };

int ttb_version(void);
const char* ttb_last_error(void);
int ttb_device_count(int* n);
int ttb_set_device(int dev);

/* pinned host staging memory for the e2e path */
int ttb_host_alloc(void** p, size_t bytes);
int ttb_host_free(void* p);

/* ---- container: TensorTrain / PeriodicTensorTrain (src/tensor_train.jl:8-24,
 *      src/periodic_tensor_train.jl:8-27; validation = check_bond_dims,
 *      src/abstract_tensor_train.jl:40-52).  bond has L+1 entries; open trains need
 *      bond[0]==bond[L]==1, periodic ones bond[L]==bond[0].  z starts at 1. */
int ttb_create(int64_t L, const int64_t* bond, int64_t Q, int periodic, int64_t batch, ttb_handle* out);
int ttb_destroy(ttb_handle h);
int ttb_clone(ttb_handle h, ttb_handle* out); /* deepcopy */
/* A + B (sign = +1) / A - B (sign = -1): `_compose`, src/tensor_train.jl:240-261 (open: first site
 * [za A, zb B], last site [A; B], block diagonal in between, z = min(|zA|, |zB|)),
 * src/periodic_tensor_train.jl:62-78 (block diagonal everywhere, f on the first site's B block, z ignored
 * like the reference); Base.:+ / Base.:- src/abstract_tensor_train.jl:315-322.  New handle, bonds add up. */
int ttb_compose(ttb_handle a, ttb_handle b, int sign, ttb_handle* out);
int ttb_info(ttb_handle h, int64_t* L, int64_t* Q, int* periodic, int64_t* batch);
/* current bond dims of train b: L+1 entries (bond_dims(A) of the reference = the first L) */
int ttb_get_bonds(ttb_handle h, int64_t b, int64_t* out);
/* doubles needed by ttb_download_all for train b at its CURRENT dims */
int ttb_nparams(ttb_handle h, int64_t b, int64_t* n); /* src/abstract_tensor_train.jl:31-37 */
int ttb_upload_site(ttb_handle h, int64_t b, int64_t t, const double* host);
int ttb_download_site(ttb_handle h, int64_t b, int64_t t, double* host);
/* all sites of all trains, packed train-major then site-major, each at its current dims */
int ttb_upload_all(ttb_handle h, const double* host);
/* Asynchronous upload_all for a train at its original bond dimensions (fresh handle or after ttb_reset):
 * returns when the copies are enqueued; `host` (pinned memory for a real overlap) must stay valid until the
 * next call on the handle that returns results.  Sites travel last-to-first, so a following
 * orthogonalize_right!/compress! (tensor_train.jl:151, abstract_tensor_train.jl:261) starts while the
 * rest of the train is still on the bus.  Falls back to the synchronous copy for batches / shrunk bonds. */
int ttb_upload_all_async(ttb_handle h, const double* host);
/* bond dimensions back to the ones the handle was created with, z = 1: reuse the slab and the workspaces
 * for the next train of the same shape */
int ttb_reset(ttb_handle h);
/* the bond dimensions the handle was created with (L+1 entries): sweeps only shrink bonds, storage and every
 * stride derived from it keep this capacity (src/tensor_train.jl:8-24 has no such notion: Julia reallocates) */
int ttb_get_capacity_bonds(ttb_handle h, int64_t* out);
/* sample / twovar_marginals / dot keep their device scratch on the handle between calls (grow-only: a
 * cudaFree of a few hundred MB costs up to hundreds of ms of wall clock); this releases it */
int ttb_trim(ttb_handle h);
int ttb_download_all(ttb_handle h, double* host);
int ttb_get_z(ttb_handle h, int64_t b, double* logabs, int* sign);
int ttb_set_z(ttb_handle h, int64_t b, double logabs, int sign);

/* ---- sweeps.  lo..hi are 1-based inclusive site indices like the reference's `indices`
 *      (pass 0,0 for the default range).  Open right sweep needs hi==L and lo>=2
 *      (src/tensor_train.jl:151-179); open left sweep needs lo==1 and hi<=L-1 (:190-218).
 *      Periodic trains ignore lo/hi (src/periodic_tensor_train.jl:83-141). */
int ttb_orthogonalize_right(ttb_handle h, const ttb_trunc* tr, int64_t lo, int64_t hi);
int ttb_orthogonalize_left(ttb_handle h, const ttb_trunc* tr, int64_t lo, int64_t hi);
/* orthogonalize_center! / orthogonalize_two_site_center! (src/tensor_train.jl:220-236) */
int ttb_orthogonalize_center(ttb_handle h, const ttb_trunc* tr, int64_t l);
int ttb_orthogonalize_two_site_center(ttb_handle h, const ttb_trunc* tr, int64_t k);
/* compress! (src/abstract_tensor_train.jl:261-274) */
int ttb_compress(ttb_handle h, const ttb_trunc* tr, int is_orthogonal);
/* per-bond results of the LAST sweep of train b: rank[i] = retained dimension of bond i (0 where the
 * sweep did not touch bond i), err[i] = relative truncation error (the quantity _debug_svd prints,
 * src/svd_trunc.jl:8-13; max over bonds = TruncBondMax.maxerr). Both have L+1 entries. */
int ttb_sweep_stats(ttb_handle h, int64_t b, int64_t* rank, double* err);
/* keep every singular-value vector of truncating sweeps (test/diagnostic aid; costs L*dmax doubles) */
int ttb_record_spectrum(ttb_handle h, int on);
/* singular values (descending, BEFORE truncation) seen at bond i in the last truncating sweep; `out` holds
 * `cap` doubles, *n receives the count (0 for a bond that sweep did not touch; TTB_EINVAL with *n set when
 * cap is too small; the largest ORIGINAL bond dimension always suffices) */
int ttb_get_spectrum(ttb_handle h, int64_t b, int64_t bond_index, double* out, int64_t cap, int64_t* n);
/* is_left_canonical / is_right_canonical residual max|A'A - I| (src/tensor_train.jl:263-273) */
int ttb_canonical_residual(ttb_handle h, int64_t b, int64_t t, int left, double* resid);

/* ---- environments (src/abstract_tensor_train.jl:89-117).  l_out/r_out may be NULL; otherwise they
 *      receive, per train, the L matrices packed site-major: l[t] is d_1 x d_{t+1}, r[t] is
 *      d_t x d_{L+1}, column-major, with the reference's scaling (normalize!=0: stored l[t<L-1]
 *      max-abs 1, last one unscaled; mirrored for r).  logabs/sign: one entry per train. */
int ttb_env_size(ttb_handle h, int64_t b, int left, int64_t* n_doubles);
int ttb_accumulate_L(ttb_handle h, int normalize, double* l_out, double* logabs, int* sign);
int ttb_accumulate_R(ttb_handle h, int normalize, double* r_out, double* logabs, int* sign);
/* normalization(A) = accumulate_L(A)[2] / A.z   (src/abstract_tensor_train.jl:154-157) */
int ttb_normalization(ttb_handle h, double* logabs, int* sign);
/* normalize!(A): returns log(|Z|/z_old) per train (src/abstract_tensor_train.jl:184-197) */
int ttb_normalize(ttb_handle h, double* logz_out);
/* normalize_eachmatrix! (src/abstract_tensor_train.jl:165-176) */
/* every site tensor of train b divided by divisor[b] (host array, one per train), z untouched: the rescaling
 * step of normalize! (src/abstract_tensor_train.jl:189-193); the MPS wrapper's normalize!
 * (src/MatrixProductStates/mps.jl:154-168) is ttb_dot(psi, psi) + this + ttb_set_z. */
int ttb_scale_sites(ttb_handle h, const double* divisor);
int ttb_normalize_eachmatrix(ttb_handle h);
/* marginals (src/abstract_tensor_train.jl:208-220): out[b][t][x], batch*L*Q doubles.  `out` of ttb_marginals* and
 * ttb_twovar_marginals* may be HOST memory (pageable or pinned) or DEVICE memory of the handle's device (unified
 * addressing decides): a caller that all-gathers the results over NCCL keeps them on the device. */
int ttb_marginals(ttb_handle h, double* out);
/* twovar_marginals (src/abstract_tensor_train.jl:231-254): pairs t<u<=t+maxdist in lexicographic
 * order, each Q*Q doubles column-major (x_t fastest); *npairs receives the count per train. */
int ttb_twovar_count(ttb_handle h, int64_t maxdist, int64_t* npairs);
int ttb_twovar_marginals(ttb_handle h, int64_t maxdist, double* out);
/* the pairs (t, u) with t_lo <= t < t_hi only (0-based first site): a contiguous block of the full result
 * (pairs are stored t-major); out holds batch x (pairs in the range) x Q*Q doubles.  This is the unit one
 * rank computes when the pair marginals of ONE train are sharded over GPUs (third axis). */
int ttb_twovar_marginals_range(ttb_handle h, int64_t maxdist, int64_t t_lo, int64_t t_hi, double* out);

/* ---- environment handles: the keyword arguments l=, r= of marginals / twovar_marginals
 *      (src/abstract_tensor_train.jl:208-209, 231-233) and rz= of sample / sample! (:335-336, :355-356) let a
 *      caller pay accumulate_L / accumulate_R once and reuse it.  ttb_env_create runs the accumulate pass
 *      (left != 0: accumulate_L, else accumulate_R) and keeps the result on the device; the *_env entry points
 *      take NULL for "compute it" (the reference's default keyword).  An environment only fits the train it was
 *      computed on AS IT WAS THEN: a call whose train has other shapes / bond dimensions fails with TTB_EINVAL. */
int ttb_env_create(ttb_handle h, int left, int normalize, ttb_env_handle* out);
int ttb_env_destroy(ttb_env_handle e);
/* the Logarithmic z returned next to the environments (accumulate_L(A)[2]), one entry per train */
int ttb_env_z(ttb_env_handle e, double* logabs, int* sign);
int ttb_marginals_env(ttb_handle h, ttb_env_handle l, ttb_env_handle r, double* out);
int ttb_twovar_marginals_env(ttb_handle h, ttb_env_handle l, ttb_env_handle r, int64_t maxdist, int64_t t_lo,
                             int64_t t_hi, double* out);
/* accumulate_M (src/abstract_tensor_train.jl:119-137) of train b: the matrices M[t,u], t < u (0-based), packed in
 * lexicographic (t,u) order, each d_{t+1} x d_u column-major, M[u-1,u] = I.  twovar_marginals streams this table
 * and never stores it (63 GiB at config 5); the entry serves callers of the reference API (M= keyword) and tests. */
int ttb_accumulate_M_size(ttb_handle h, int64_t b, int64_t* n_doubles);
int ttb_accumulate_M(ttb_handle h, int64_t b, int normalize, double* out);

/* ---- is_in_domain (src/abstract_tensor_train.jl:59-64): X is n x L flat 0-based indices; *ok = all inside 0:Q-1.
 *      == (src/abstract_tensor_train.jl:85): isequal of the site tensors (z is NOT compared, like the reference;
 *      NaN equals NaN, -0.0 differs from 0.0); trains of different kind / length / shape are unequal. */
int ttb_is_in_domain(ttb_handle h, const int32_t* X, int64_t n, int* ok);
int ttb_equal(ttb_handle a, int64_t ba, ttb_handle b, int64_t bb, int* equal);

/* ---- evaluate (src/abstract_tensor_train.jl:80-83): X is n x L row-major int32 of 0-based flat
 *      physical indices; evaluates train b at n points; out = tr(prod A_t(x_t)) / z */
int ttb_evaluate(ttb_handle h, int64_t b, const int32_t* X, int64_t n, double* out);

/* ---- sample (src/abstract_tensor_train.jl:335-384, draw = src/utils.jl:9-20) on train b.
 *      Draw i uses uniforms[i*L + t] at site t when `uniforms` != NULL, else the Philox4x32-10 stream
 *      key=(seed), counter=(first_index+i, t).  x_out: nsamples*L flat 0-based indices (uint8 when
 *      Q<=256) ; p_out (nullable): the normalised probability of each draw, tr(Q)/z_R. */
int ttb_sample(ttb_handle h, int64_t b, int64_t nsamples, uint64_t seed, uint64_t first_index,
               const double* uniforms, uint8_t* x_out, double* p_out);
/* device-resident variant for throughput runs: draws stay in HBM, only site histograms
 * (L*Q int64, accumulated) come back.  hist may be NULL. */
int ttb_sample_stats(ttb_handle h, int64_t b, int64_t nsamples, uint64_t seed, uint64_t first_index,
                     int64_t* hist, double* sum_logp);

/* sample / sample_stats with rz= (an accumulate_R environment handle; NULL = recompute like the default keyword) */
int ttb_sample_env(ttb_handle h, ttb_env_handle rz, int64_t b, int64_t nsamples, uint64_t seed, uint64_t first_index,
                   const double* uniforms, uint8_t* x_out, double* p_out);
int ttb_sample_stats_env(ttb_handle h, ttb_env_handle rz, int64_t b, int64_t nsamples, uint64_t seed,
                         uint64_t first_index, int64_t* hist, double* sum_logp);

/* ---- MPS: the Born-machine wrapper p(x) = |psi(x)|^2 over the train psi held by `h`
 *      (src/MatrixProductStates/mps.jl).  evaluate / normalization / normalize! / the forwarded sweeps are host
 *      arithmetic on ttb_evaluate, ttb_dot(psi, psi), ttb_scale_sites and the sweep entries; the entries below are
 *      the functions that need the reference's "d^4" environments (mps.jl:60-126): the four-index arrays
 *      Ll[a1, a, b1, b] are kept as d_1^2 boundary slices of d x d (one slice, a1 = b1 = 1, for an open psi).
 *      `b` selects the train of a batch.
 *  ttb_mps_accumulate: accumulate_L (left != 0, mps.jl:60-80) / accumulate_R (:82-103) -- the stored environments
 *      (slot t: d_1^2 slices (a1 + d_1 b1) of d_{t+1} x d_{t+1} for l[t], of d_t x d_t for r[t], column-major, packed in
 *      site order; hermiticity
 *      restored after every step; with normalize every stored slot but the last computed has max-abs 1, like the
 *      reference's in-place `Ll ./= nl`) go to env_out when non-NULL, z = prod(nl) * trace to (logabs, sign).
 *  ttb_mps_marginals (:176-196): out[t][x], L*Q doubles.
 *  ttb_mps_twovar_marginals (:206-232, with accumulate_M :105-126 streamed, never stored): pairs t < u <= t + maxdist
 *      in lexicographic order, Q*Q doubles each (x_t fastest), ttb_twovar_count pairs; TensorTrain psi only (mps.jl:206).
 *  ttb_mps_sample (sample!, :264-290): like ttb_sample; p_out = |tr Q|^2 / z with z of accumulate_R. */
int ttb_mps_env_size(ttb_handle h, int64_t b, int left, int64_t* n_doubles);
int ttb_mps_accumulate(ttb_handle h, int64_t b, int left, int normalize, double* env_out, double* logabs, int* sign);
int ttb_mps_marginals(ttb_handle h, int64_t b, double* out);
int ttb_mps_twovar_marginals(ttb_handle h, int64_t b, int64_t maxdist, double* out);
int ttb_mps_sample(ttb_handle h, int64_t b, int64_t nsamples, uint64_t seed, uint64_t first_index, const double* uniforms,
                   uint8_t* x_out, double* p_out);
/* the same with the keyword arguments l= / r= / rz= (mps.jl:176-177, 206-208, 264-265): environments as
 * ttb_mps_accumulate(normalize = 1) returned them (host or device memory), NULL = compute here like the reference's
 * default keyword; z_logabs = log z of accumulate_R (the second element of rz).  M= of twovar_marginals has no
 * counterpart: accumulate_M is streamed. */
int ttb_mps_marginals_env(ttb_handle h, int64_t b, const double* l_env, const double* r_env, double* out);
int ttb_mps_twovar_marginals_env(ttb_handle h, int64_t b, const double* l_env, const double* r_env, int64_t maxdist, double* out);
int ttb_mps_sample_env(ttb_handle h, int64_t b, const double* r_env, double z_logabs, int64_t nsamples, uint64_t seed,
                       uint64_t first_index, const double* uniforms, uint8_t* x_out, double* p_out);

/* ---- two-site pieces (the data-parallel parts of 2-site DMRG, src/dmrg.jl:59-151; the optimisation loop itself is
 *      host control flow and not part of the library).  `t` is the 0-based index of the first site of the pair.
 *  ttb_merge_tensors (_merge_tensors, src/utils.jl:37-44): out[m, n, xA, xB] = sum_k A_t[m,k,xA] A_{t+1}[k,n,xB],
 *      d_t x d_{t+2} x Q x Q column-major; `out` may be host or device memory.
 *  ttb_split_tensor (_split_tensor, src/abstract_tensor_train.jl:283-308): SVD of C[(al,xl),(al1,xl1)] under `tr`,
 *      left != 0: A_t = U, A_{t+1} = L V' (lr = Left()); left == 0: A_t = U L, A_{t+1} = V' (lr = Right()); the
 *      sites and the bond between them are replaced in place (z is untouched).  The new bond must fit the capacity
 *      the handle was created with at that bond (TTB_EUNSUPPORTED otherwise).  *trunc_err (nullable) receives the
 *      relative truncation error of the step (what TruncBondMax records).
 *  ttb_data_environments (precompute_left_environments / precompute_right_environments, src/tensor_train.jl:320-334):
 *      X is n x L flat 0-based indices; env[i][j][0 .. d_j) for j = 0..L with a stride of max(bond capacity) doubles
 *      per entry: left != 0: entry j = A_0(x_0)...A_{j-1}(x_{j-1}) (entry 0 = [1]); left == 0: entry j =
 *      A_j(x_j)...A_{L-1}(x_{L-1}) (entry L = [1]).  `env` may be host or device memory.
 *  ttb_data_env_update (update_environments!, src/dmrg.jl:124-151): after site t changed, left != 0: entry t+1 =
 *      entry t * A_t(x_t); left == 0: entry t = A_t(x_t) * entry t+1, for every datapoint, in place. */
int ttb_merge_tensors(ttb_handle h, int64_t b, int64_t t, double* out);
int ttb_split_tensor(ttb_handle h, int64_t b, int64_t t, const double* C, const ttb_trunc* tr, int left, double* trunc_err);
int ttb_data_environments(ttb_handle h, int64_t b, const int32_t* X, int64_t n, int left, double* env);
int ttb_data_env_update(ttb_handle h, int64_t b, const int32_t* X, int64_t n, int64_t t, int left, double* env);

/* ---- dot (src/abstract_tensor_train.jl:395-408) of train ba of `a` with train bb of `b` (same length and
 * physical dimension, same device; open or periodic, bonds may differ).  The value is returned as
 * (log|dot|, sign) including both z factors; norm / norm2m / isapprox (:418-437) are host arithmetic on it. */
int ttb_dot(ttb_handle a, int64_t ba, ttb_handle b, int64_t bb, double* logabs, int* sign);

/* timing aid for bench.py: milliseconds spent on the device inside the last call that ran kernels
 * (CUDA events on the handle's stream) and how many kernels it launched */
int ttb_last_timing(ttb_handle h, double* ms, int64_t* launches);

/* per-kernel profiling for the roofline: after ttb_profile(h,1) every kernel launch of this handle is
 * bracketed by CUDA events; ttb_profile_report writes lines "<kernel> <launches> <total ms>". */
int ttb_profile(ttb_handle h, int on);
int ttb_profile_report(ttb_handle h, char* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* TTB200_H */
