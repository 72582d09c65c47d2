"""CPU oracle: a NumPy/SciPy restatement of the TensorTrains.jl hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``tensortrains.jl_b200/``) never does.

What it restates (reference = stecrotti/TensorTrains.jl v0.14.0, paths relative to
``/root/reference``); every function cites the lines it follows:

* ``src/svd_trunc.jl``              -> TruncThresh / TruncBond / TruncBondMax / TruncBondThresh
* ``src/tensor_train.jl``           -> TensorTrain, flat_tt/rand_tt/randn_tt, orthogonalize_*
* ``src/periodic_tensor_train.jl``  -> PeriodicTensorTrain, wrap-around sweeps
* ``src/abstract_tensor_train.jl``  -> evaluate, accumulate_L/R/M, normalization, normalize!,
                                       marginals, twovar_marginals, compress!, sample, dot, norm
* ``src/utils.jl``                  -> _reshape1, _reshapeas, sample_noalloc
* ``test/exact.jl``                 -> exact_* brute-force enumerators

Third-party arithmetic the reference delegates to and that is NOT in /root/reference
(no Manifest.toml is shipped, versions bounded by Project.toml [compat] only):
``LinearAlgebra.svd`` = LAPACK ``gesdd`` via MKL.jl (0.6.3/0.7/0.9) -> restated with
``scipy.linalg.svd(lapack_driver="gesdd")`` (same LAPACK routine, OpenBLAS build);
Tullio 0.3 contractions -> ``numpy`` matmul/einsum; TensorCast 0.4 folds -> ``order="F"``
reshapes; LogarithmicNumbers 1.4 -> :class:`LogNum` (sign, log|.|).

Parity pinning: Julia is not installed in this image, so the reference cannot be run here and
it ships no golden vectors.  The oracle is pinned by restating the reference's own test-suite
invariants (brute-force enumeration ``test/exact.jl``, gauge invariance, known answers; see
``tests/test_oracle_reference_suite.py``) -- i.e. pinned up to gauge, at the sizes the
reference tests use.  Julia RNG streams cannot be reproduced; inputs come from NumPy
``Generator(Philox(seed))`` and are fed bit-identically to oracle and GPU.

Layout: site tensors are ``numpy`` arrays of shape ``(d_t, d_{t+1}, q_1, ...)``; all folds use
column-major (``order="F"``) semantics exactly like Julia.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg as sla


# ----------------------------------------------------------------------------------------------
# LogarithmicNumbers.Logarithmic restatement: sign * exp(logabs)
# ----------------------------------------------------------------------------------------------
class LogNum:
    """Signed logarithmic number (LogarithmicNumbers.jl ``Logarithmic``; used for ``z``,
    reference ``src/tensor_train.jl:10,12``)."""

    __slots__ = ("logabs", "sign")

    def __init__(self, x=1.0, *, logabs=None, sign=None):
        if logabs is not None:
            self.logabs = float(logabs)
            self.sign = int(sign if sign is not None else 1)
        elif isinstance(x, LogNum):
            self.logabs, self.sign = x.logabs, x.sign
        else:
            x = float(x)
            if x == 0.0:
                self.logabs, self.sign = -math.inf, 1
            elif math.isnan(x):
                self.logabs, self.sign = math.nan, 1
            else:
                self.logabs, self.sign = math.log(abs(x)), (1 if x > 0 else -1)

    @staticmethod
    def _c(o):
        return o if isinstance(o, LogNum) else LogNum(o)

    def __mul__(self, o):
        o = self._c(o)
        return LogNum(logabs=self.logabs + o.logabs, sign=self.sign * o.sign)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = self._c(o)
        return LogNum(logabs=self.logabs - o.logabs, sign=self.sign * o.sign)

    def __rtruediv__(self, o):
        return self._c(o) / self

    def __abs__(self):
        return LogNum(logabs=self.logabs, sign=1)

    def __float__(self):
        return self.sign * math.exp(self.logabs)

    def log(self):
        if self.sign < 0:
            raise ValueError("DomainError: log of a negative Logarithmic")
        return self.logabs

    def __repr__(self):
        return f"LogNum({'+' if self.sign > 0 else '-'}exp({self.logabs}))"


# ----------------------------------------------------------------------------------------------
# helpers  (src/utils.jl:1-3)
# ----------------------------------------------------------------------------------------------
def _reshape1(x: np.ndarray) -> np.ndarray:
    """Collapse the physical indices into one, column-major (``src/utils.jl:1``)."""
    return x.reshape(x.shape[0], x.shape[1], -1, order="F")


def _reshapeas(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Give ``x`` the physical shape of ``y`` (``src/utils.jl:2``)."""
    return x.reshape((x.shape[0], x.shape[1]) + tuple(y.shape[2:]), order="F")


def _svd(M: np.ndarray):
    """``LinearAlgebra.svd`` default = thin SVD through LAPACK gesdd (``src/svd_trunc.jl:37``)."""
    if not np.all(np.isfinite(M)):
        # LAPACK gesdd errors on NaN input; Julia propagates NaN.  Only the rank rule matters.
        k = min(M.shape)
        return np.full((M.shape[0], k), np.nan), np.full(k, np.nan), np.full((M.shape[1], k), np.nan)
    U, s, Vh = sla.svd(M, full_matrices=False, lapack_driver="gesdd", check_finite=False)
    return U, s, Vh.T


# ----------------------------------------------------------------------------------------------
# SVD truncators  (src/svd_trunc.jl)
# ----------------------------------------------------------------------------------------------
class SVDTrunc:
    kind = -1
    capture = None      # test hook: callable(M, lam, mprime) called on every SVD (the high-precision arbiter of
                        # tests/test_gpu_full_parity.py keeps the matrices whose rank decision sits near the threshold)

    def rank(self, lam: np.ndarray) -> int:  # pragma: no cover - abstract
        raise NotImplementedError

    def __call__(self, M: np.ndarray):
        U, lam, V = _svd(M)
        mprime = self.rank(lam)
        self.last_spectrum = lam.copy()
        if self.capture is not None:
            self.capture(M, lam, mprime)
        return U[:, :mprime], lam[:mprime], V[:, :mprime]


def _thresh_rank(lam: np.ndarray, eps: float, s0: float) -> int:
    """Shared loop of ``TruncThresh`` (``src/svd_trunc.jl:38-47``, ``s0 = last(lam)^2``) and
    ``TruncBondThresh`` (``:130-138``, ``s0 = 0``): walk k = m-1 .. 1 accumulating lam_k^2 and stop
    at the first k whose running sum reaches ``(eps*||lam||)^2``; keep ``k+1`` values."""
    lam_norm = float(np.linalg.norm(lam))
    thr = (lam_norm * eps) ** 2
    mprime = 1
    s = s0
    for k in range(len(lam) - 1, 0, -1):  # 1-based k
        s += float(lam[k - 1]) ** 2
        if s >= thr:
            mprime = k + 1
            break
    return mprime


@dataclass
class TruncThresh(SVDTrunc):
    """``src/svd_trunc.jl:33-50``."""
    eps: float
    kind = 0

    def rank(self, lam):
        return _thresh_rank(lam, self.eps, float(lam[-1]) ** 2)


@dataclass
class TruncBond(SVDTrunc):
    """``src/svd_trunc.jl:69-77``."""
    mprime: int
    kind = 1

    def rank(self, lam):
        return min(len(lam), self.mprime)


@dataclass
class TruncBondMax(SVDTrunc):
    """``src/svd_trunc.jl:91-103``: TruncBond + running max of the relative truncation error."""
    mprime: int
    maxerr: List[float] = field(default_factory=lambda: [0.0])
    kind = 2

    def rank(self, lam):
        mprime = min(len(lam), self.mprime)
        err = math.sqrt(float(np.sum(lam[mprime:] ** 2)) / float(np.sum(lam ** 2)))
        self.maxerr[0] = max(self.maxerr[0], err)
        return mprime


@dataclass
class TruncBondThresh(SVDTrunc):
    """``src/svd_trunc.jl:123-143``: tail sum starts at 0 and skips lam_m, then min with cap."""
    mprime: int
    eps: float = 0.0
    kind = 3

    def rank(self, lam):
        return min(_thresh_rank(lam, self.eps, 0.0), self.mprime)


# ----------------------------------------------------------------------------------------------
# containers  (src/tensor_train.jl:8-24, src/periodic_tensor_train.jl:8-27)
# ----------------------------------------------------------------------------------------------
def check_bond_dims(tensors: Sequence[np.ndarray]) -> bool:
    """``src/abstract_tensor_train.jl:40-52`` (cyclic compatibility check)."""
    n = len(tensors)
    for t in range(n):
        if tensors[t].shape[1] != tensors[(t + 1) % n].shape[0]:
            return False
    return True


class AbstractTensorTrain:
    periodic = False

    def __init__(self, tensors: Sequence[np.ndarray], z=1.0):
        tensors = [np.asarray(a, dtype=np.float64) for a in tensors]
        if any(a.ndim <= 2 for a in tensors):
            raise ValueError("ArgumentError: Tensors should have at least 3 indices: 2 virtual and 1 physical")
        self._validate(tensors)
        self.tensors = list(tensors)
        self.z = z if isinstance(z, LogNum) else LogNum(z)

    def _validate(self, tensors):  # pragma: no cover - abstract
        raise NotImplementedError

    # forwarding (src/tensor_train.jl:27-29)
    def __len__(self):
        return len(self.tensors)

    def __getitem__(self, t):
        return self.tensors[t]

    def __setitem__(self, t, v):
        self.tensors[t] = v

    def __iter__(self):
        return iter(self.tensors)

    def copy(self):
        out = type(self)([a.copy() for a in self.tensors], z=LogNum(self.z))
        return out


class TensorTrain(AbstractTensorTrain):
    """Open-boundary TT (``src/tensor_train.jl:8-24``)."""

    def _validate(self, tensors):
        if not (tensors[0].shape[0] == 1 and tensors[-1].shape[1] == 1):
            raise ValueError("ArgumentError: First matrix must have 1 row, last matrix must have 1 column")
        if not check_bond_dims(tensors):
            raise ValueError("ArgumentError: Matrix indices for matrix product non compatible")


class PeriodicTensorTrain(AbstractTensorTrain):
    """Periodic TT (``src/periodic_tensor_train.jl:8-27``)."""
    periodic = True

    def _validate(self, tensors):
        if not check_bond_dims(tensors):
            raise ValueError("ArgumentError: Matrix indices for matrix product non compatible")


def _compose(f, A, B):
    """``A + B`` / ``A - B`` (``f`` = +1 / -1).  Open trains: ``src/tensor_train.jl:240-261`` -- first site
    ``[za A, zb B]`` with ``z = min(|zA|, |zB|)``, ``za = z/zA``, ``zb = f(z/zB)``, last site ``[A; B]``,
    block diagonal in between, ``C.z = z``.  Periodic trains: ``src/periodic_tensor_train.jl:62-78`` -- block
    diagonal everywhere, ``f`` on the first site's B block, and ``z`` is ignored (the result has z = 1)."""
    if type(A) is not type(B):
        raise ValueError("ArgumentError: cannot compose trains of different kinds")
    if len(A) != len(B):
        raise ValueError(f"ArgumentError: Tensor Trains must have the same length, got {len(A)} and {len(B)}")
    if A[0].shape[2:] != B[0].shape[2:]:
        raise ValueError("ArgumentError: Tensor Trains must have the types of physical indices")
    L = len(A)
    out = []
    if A.periodic:
        za, zb, z = 1.0, float(f), LogNum(1.0)
    else:
        z = A.z if abs(A.z).logabs <= abs(B.z).logabs else B.z
        z = abs(z)
        za, zb = float(z / A.z), f * float(z / B.z)
    for t in range(L):
        At, Bt = _reshape1(A[t]), _reshape1(B[t])
        (al, ar, Q), (bl, br, _) = At.shape, Bt.shape
        if not A.periodic and t == 0:
            C = np.concatenate([za * At, zb * Bt], axis=1)
        elif not A.periodic and t == L - 1:
            C = np.concatenate([At, Bt], axis=0)
        else:
            C = np.zeros((al + bl, ar + br, Q))
            C[:al, :ar] = At
            C[al:, ar:] = (zb * Bt) if (A.periodic and t == 0) else Bt
        out.append(C.reshape(C.shape[:2] + A[t].shape[2:], order="F"))
    return type(A)(out, z=z)


def plus(A, B):
    """``Base.:+`` (``src/abstract_tensor_train.jl:315``)"""
    return _compose(1.0, A, B)


def minus(A, B):
    """``Base.:-`` (``src/abstract_tensor_train.jl:322``)"""
    return _compose(-1.0, A, B)


def bond_dims(A) -> List[int]:
    """Left bond of every site (``src/abstract_tensor_train.jl:23``)."""
    return [a.shape[0] for a in A]


def nparams(A) -> int:
    """``src/abstract_tensor_train.jl:31-37`` (real entries)."""
    return int(sum(a.size for a in A))


def _open_bonds(d, L):
    return [1] + [d] * (L - 1) + [1]


def flat_tt(bondsizes, *q):
    """``src/tensor_train.jl:42-48``; ``flat_tt(d, L, q...)`` form via :func:`_open_bonds`."""
    return TensorTrain([np.ones((bondsizes[t], bondsizes[t + 1]) + tuple(q)) for t in range(len(bondsizes) - 1)])


def rand_tt(rng: np.random.Generator, bondsizes, *q):
    """``src/tensor_train.jl:61-77``: U[0,1) entries (NumPy stream, not Julia's)."""
    return TensorTrain([np.asfortranarray(rng.random((bondsizes[t], bondsizes[t + 1]) + tuple(q)))
                        for t in range(len(bondsizes) - 1)])


def randn_tt(rng: np.random.Generator, bondsizes, *q):
    """``src/tensor_train.jl:90-108``: N(0,1)/sqrt(max bond)."""
    sigma = 1.0 / math.sqrt(max(bondsizes))
    return TensorTrain([np.asfortranarray(sigma * rng.standard_normal((bondsizes[t], bondsizes[t + 1]) + tuple(q)))
                        for t in range(len(bondsizes) - 1)])


def flat_periodic_tt(bondsizes, *q):
    """``src/periodic_tensor_train.jl:39-44`` (``bondsizes`` has one entry per site)."""
    n = len(bondsizes)
    x = 1.0 / (float(np.prod(bondsizes, dtype=np.float64)) ** (1.0 / n) * float(np.prod(q)))
    return PeriodicTensorTrain([np.full((bondsizes[t], bondsizes[(t + 1) % n]) + tuple(q), x) for t in range(n)])


def rand_periodic_tt(rng: np.random.Generator, bondsizes, *q):
    """``src/periodic_tensor_train.jl:56-59``."""
    n = len(bondsizes)
    return PeriodicTensorTrain([np.asfortranarray(rng.random((bondsizes[t], bondsizes[(t + 1) % n]) + tuple(q)))
                                for t in range(n)])


# ----------------------------------------------------------------------------------------------
# sweeps
# ----------------------------------------------------------------------------------------------
def _finite_nonzero(m: float) -> bool:
    return not (math.isnan(m) or math.isinf(m) or m == 0.0)


def _absorb_right(Cprev: np.ndarray, U: np.ndarray, lam: np.ndarray) -> np.ndarray:
    """``D[m,n,x] = sum_k C[m,k,x] U[k,n] lam[n]`` (``src/tensor_train.jl:168``) through BLAS."""
    UL = U * lam[None, :]
    D = np.tensordot(Cprev, UL, axes=([1], [0]))  # (m, x, n)
    return np.ascontiguousarray(D.transpose(0, 2, 1))


def _absorb_left(lam: np.ndarray, V: np.ndarray, Cnext: np.ndarray) -> np.ndarray:
    """``D[m,n,x] = lam[m] sum_l V'[m,l] C[l,n,x]`` (``src/tensor_train.jl:207``)."""
    LV = lam[:, None] * V.T
    return np.tensordot(LV, Cnext, axes=([1], [0]))  # (m, n, x)


def _fold_right(C3: np.ndarray) -> np.ndarray:
    """``M[m,(n,x)] := C[m,n,x]`` (``src/tensor_train.jl:159,174``): free column-major reshape."""
    return C3.reshape(C3.shape[0], -1, order="F")


def _fold_left(C3: np.ndarray) -> np.ndarray:
    """``M[(m,x),n] |= C[m,n,x]`` (``src/tensor_train.jl:198,213``)."""
    return C3.transpose(0, 2, 1).reshape(-1, C3.shape[1], order="F")


def _unfold_right(Vt: np.ndarray, q: int) -> np.ndarray:
    """``A[m,n,x] := V'[m,(n,x)]`` (``src/tensor_train.jl:165``)."""
    return Vt.reshape(Vt.shape[0], -1, q, order="F")


def _unfold_left(U: np.ndarray, q: int) -> np.ndarray:
    """``A[m,n,x] := U[(m,x),n]`` (``src/tensor_train.jl:204``)."""
    return U.reshape(-1, q, U.shape[1], order="F").transpose(0, 2, 1)


def _check_indices(indices, lo, hi, L):
    idx = list(indices)
    if idx != sorted(idx):
        raise ValueError("ArgumentError: Indices must be sorted")
    if not all(lo <= i <= hi for i in (min(idx), max(idx))):
        raise ValueError(f"ArgumentError: Indices not compatible with tensor train positions 1:{L}: got {idx}")
    return idx


def orthogonalize_right(C, svd_trunc: Optional[SVDTrunc] = None, indices=None, stats=None):
    """``orthogonalize_right!`` for open (``src/tensor_train.jl:151-179``) and periodic
    (``src/periodic_tensor_train.jl:83-111``) trains.  ``indices`` are 1-based site numbers."""
    L = len(C)
    if C.periodic:
        svd_trunc = svd_trunc if svd_trunc is not None else TruncThresh(0.0)
        C0 = _reshape1(C[0])
        q = C0.shape[2]
        U, lam, V = svd_trunc(_fold_right(C0))
        if stats is not None:
            stats.append((1, svd_trunc.last_spectrum, len(lam)))
        C[0] = _reshapeas(_unfold_right(V.T, q), C[0])
        D = _absorb_right(_reshape1(C[L - 1]), U, lam)  # no rescale on the wrap step (:91-92)
        M = _fold_right(D)
        c = LogNum(1.0)
        for t in range(L, 1, -1):
            U, lam, V = svd_trunc(M)
            if stats is not None:
                stats.append((t, svd_trunc.last_spectrum, len(lam)))
            C[t - 1] = _reshapeas(_unfold_right(V.T, q), C[t - 1])
            D = _absorb_right(_reshape1(C[t - 2]), U, lam)
            m = float(np.max(np.abs(D)))
            if _finite_nonzero(m):
                D = D / m
                c = c * m
            M = _fold_right(D)
        C[0] = _reshapeas(D, C[0])
        C.z = C.z / c
        return C

    svd_trunc = svd_trunc if svd_trunc is not None else TruncThresh(1e-6)
    indices = range(2, L + 1) if indices is None else indices
    if len(indices) == 0:
        return C
    idx = _check_indices(indices, 2, L, L)
    CT = _reshape1(C[L - 1])  # always seeds from the LAST site (:157)
    q = CT.shape[2]
    M = _fold_right(CT)
    D = np.ones((1, 1, 1))
    c = LogNum(1.0)
    for t in reversed(idx):
        U, lam, V = svd_trunc(M)
        if stats is not None:
            stats.append((t, svd_trunc.last_spectrum, len(lam)))
        C[t - 1] = _reshapeas(_unfold_right(V.T, q), C[t - 1])
        D = _absorb_right(_reshape1(C[t - 2]), U, lam)
        m = float(np.max(np.abs(D)))
        if _finite_nonzero(m):
            D = D / m
            c = c * m
        M = _fold_right(D)
    C[idx[0] - 2] = _reshapeas(D, C[idx[0] - 2])
    C.z = C.z / c
    return C


def orthogonalize_left(C, svd_trunc: Optional[SVDTrunc] = None, indices=None, stats=None):
    """``orthogonalize_left!`` for open (``src/tensor_train.jl:190-218``) and periodic
    (``src/periodic_tensor_train.jl:113-141``) trains."""
    L = len(C)
    if C.periodic:
        svd_trunc = svd_trunc if svd_trunc is not None else TruncThresh(0.0)
        A0 = _reshape1(C[0])
        q = A0.shape[2]
        M = _fold_left(A0)
        D = np.ones((1, 1, 1))
        c = LogNum(1.0)
        for t in range(1, L):
            U, lam, V = svd_trunc(M)
            if stats is not None:
                stats.append((t, svd_trunc.last_spectrum, len(lam)))
            C[t - 1] = _reshapeas(_unfold_left(U, q), C[t - 1])
            D = _absorb_left(lam, V, _reshape1(C[t]))
            m = float(np.max(np.abs(D)))
            if _finite_nonzero(m):
                D = D / m
                c = c * m
            M = _fold_left(D)
        U, lam, V = svd_trunc(M)
        if stats is not None:
            stats.append((L, svd_trunc.last_spectrum, len(lam)))
        C[L - 1] = _reshapeas(_unfold_left(U, q), C[L - 1])
        D = _absorb_left(lam, V, _reshape1(C[0]))  # wrap step, no rescale (:136-138)
        C[0] = _reshapeas(D, C[0])
        C.z = C.z / c
        return C

    svd_trunc = svd_trunc if svd_trunc is not None else TruncThresh(1e-6)
    indices = range(1, L) if indices is None else indices
    if len(indices) == 0:
        return C
    idx = _check_indices(indices, 1, L - 1, L)
    C0 = _reshape1(C[0])  # always seeds from the FIRST site (:196)
    q = C0.shape[2]
    M = _fold_left(C0)
    D = np.ones((1, 1, 1))
    c = LogNum(1.0)
    for t in idx:
        U, lam, V = svd_trunc(M)
        if stats is not None:
            stats.append((t, svd_trunc.last_spectrum, len(lam)))
        C[t - 1] = _reshapeas(_unfold_left(U, q), C[t - 1])
        D = _absorb_left(lam, V, _reshape1(C[t]))
        m = float(np.max(np.abs(D)))
        if _finite_nonzero(m):
            D = D / m
            c = c * m
        M = _fold_left(D)
    C[idx[-1]] = _reshapeas(D, C[idx[-1]])
    C.z = C.z / c
    return C


def orthogonalize_center(C, l: int, svd_trunc=None):
    """``src/tensor_train.jl:220-223`` (1-based ``l``)."""
    svd_trunc = svd_trunc if svd_trunc is not None else TruncThresh(1e-6)
    orthogonalize_left(C, svd_trunc, range(1, l))
    orthogonalize_right(C, svd_trunc, range(l + 1, len(C) + 1))
    return C


def orthogonalize_two_site_center(C, k: int, svd_trunc=None):
    """``src/tensor_train.jl:232-236``."""
    assert 1 <= k < len(C)
    svd_trunc = svd_trunc if svd_trunc is not None else TruncThresh(1e-6)
    orthogonalize_left(C, svd_trunc, range(1, k))
    orthogonalize_right(C, svd_trunc, range(k + 2, len(C) + 1))
    return C


def compress(A, svd_trunc=None, is_orthogonal: str = "none", stats=None):
    """``compress!`` (``src/abstract_tensor_train.jl:261-274``)."""
    svd_trunc = svd_trunc if svd_trunc is not None else TruncThresh(1e-6)
    if is_orthogonal == "none":
        orthogonalize_right(A, TruncThresh(0.0))
        orthogonalize_left(A, svd_trunc, stats=stats)
    elif is_orthogonal == "left":
        orthogonalize_right(A, svd_trunc, stats=stats)
    elif is_orthogonal == "right":
        orthogonalize_left(A, svd_trunc, stats=stats)
    else:
        raise ValueError(f"ArgumentError: Keyword `is_orthogonal` only supports: :none, :left, :right, got :{is_orthogonal}")
    return A


def is_left_canonical(a: np.ndarray, atol=1e-10) -> bool:
    """``src/tensor_train.jl:263-267``."""
    a3 = _reshape1(a)
    AA = np.einsum("kix,kjx->ij", a3, a3)
    return bool(np.allclose(AA, np.eye(AA.shape[0]), atol=atol, rtol=0))


def is_right_canonical(a: np.ndarray, atol=1e-10) -> bool:
    """``src/tensor_train.jl:269-273``."""
    a3 = _reshape1(a)
    AA = np.einsum("ikx,jkx->ij", a3, a3)
    return bool(np.allclose(AA, np.eye(AA.shape[0]), atol=atol, rtol=0))


# ----------------------------------------------------------------------------------------------
# environments, normalisation, marginals  (src/abstract_tensor_train.jl:80-254)
# ----------------------------------------------------------------------------------------------
def evaluate(A, X) -> float:
    """``tr(prod A[t][:,:,x_t...]) / z`` (``src/abstract_tensor_train.jl:80-83``).
    ``X`` = per-site tuples (or ints) of 0-based physical indices."""
    P = None
    for a, x in zip(A, X):
        x = (x,) if np.isscalar(x) else tuple(x)
        m = a[(slice(None), slice(None)) + x]
        P = m if P is None else P @ m
    return float(LogNum(float(np.trace(P))) / A.z)


def trace(At: np.ndarray) -> np.ndarray:
    """``sum_x A[:,:,x]`` (``src/abstract_tensor_train.jl:87``)."""
    return _reshape1(At).sum(axis=2)


def accumulate_L(A, normalize=True):
    """``src/abstract_tensor_train.jl:89-102`` including the in-place quirk: ``Lt ./= nt``
    rescales the array already stored as the previous element of the result."""
    d1 = A[0].shape[0]
    Lt = np.eye(d1)
    z = LogNum(1.0)
    out = []
    for At in A:
        T = trace(At)
        nt = float(np.max(np.abs(Lt)))
        if nt != 0.0 and normalize:
            Lt /= nt  # in place: also rescales out[-1]
            z = z * nt
        Lt = Lt @ T
        out.append(Lt)
    z = z * float(np.trace(Lt))
    return out, z


def accumulate_R(A, normalize=True):
    """``src/abstract_tensor_train.jl:104-117``."""
    dL = A[len(A) - 1].shape[1]
    Rt = np.eye(dL)
    z = LogNum(1.0)
    out = []
    for At in reversed(list(A)):
        T = trace(At)
        nt = float(np.max(np.abs(Rt)))
        if nt != 0.0 and normalize:
            Rt /= nt
            z = z * nt
        Rt = T @ Rt
        out.append(Rt)
    out.reverse()
    z = z * float(np.trace(Rt))
    return out, z


def accumulate_M(A, normalize=True):
    """``src/abstract_tensor_train.jl:119-137``; returns a dict keyed by 0-based ``(t,u)``."""
    L = len(A)
    M = {}
    for u in range(1, L):  # 0-based u = 1..L-1  (Julia u = 2..L)
        Au = trace(A[u - 1])
        for t in range(0, u - 1):
            m = M[(t, u - 1)] @ Au
            nt = float(np.max(np.abs(m)))
            if nt != 0.0 and normalize:
                m = m / nt
            M[(t, u)] = m
        M[(u - 1, u)] = np.eye(A[u].shape[0])
    return M


def normalization(A) -> LogNum:
    """``src/abstract_tensor_train.jl:154-157``."""
    _, z = accumulate_L(A)
    return z / A.z


def lognormalization(A) -> float:
    """``src/abstract_tensor_train.jl:145-147`` (raises on negative Z)."""
    return normalization(A).log()


def normalize_eachmatrix(A):
    """``src/abstract_tensor_train.jl:165-176``."""
    c = LogNum(1.0)
    for i, m in enumerate(A):
        mm = float(np.max(np.abs(m)))
        if _finite_nonzero(mm):
            A[i] = m / mm
            c = c * mm
    A.z = A.z / c


def normalize(A) -> float:
    """``normalize!`` (``src/abstract_tensor_train.jl:184-197``); returns ``log(|Z|/z_old)``."""
    Z = accumulate_L(A)[1]
    absZ = abs(Z)
    L = len(A)
    x = math.exp(absZ.logabs / L)
    if x != 0:
        for i in range(L):
            A[i] = A[i] / x
    logz = (absZ / A.z).log()
    A.z = LogNum(1.0)
    return logz


def marginals(A, l=None, r=None):
    """``src/abstract_tensor_train.jl:208-220``."""
    l = accumulate_L(A)[0] if l is None else l
    r = accumulate_R(A)[0] if r is None else r
    L = len(A)
    out = []
    for t in range(L):
        At = _reshape1(A[t])
        R = r[t + 1] if t + 1 < L else np.eye(A[L - 1].shape[1])
        Lm = l[t - 1] if t - 1 >= 0 else np.eye(A[0].shape[0])
        lA = np.einsum("ab,bcx->acx", Lm, At)
        p = np.einsum("acx,ca->x", lA, R)
        p = p / p.sum()
        out.append(p.reshape(A[t].shape[2:], order="F"))
    return out


def twovar_marginals(A, l=None, r=None, M=None, maxdist=None):
    """``src/abstract_tensor_train.jl:231-254``; returns dict keyed by 0-based ``(t,u)``, each a
    ``(q..., q...)`` array (entries the reference leaves as 0-size arrays are absent)."""
    l = accumulate_L(A)[0] if l is None else l
    r = accumulate_R(A)[0] if r is None else r
    M = accumulate_M(A) if M is None else M
    L = len(A)
    maxdist = L - 1 if maxdist is None else maxdist
    qs = tuple(A[0].shape[2:]) * 2
    d = A[0].shape[0]
    b = {}
    for t in range(L - 1):
        lt = np.eye(d) if t == 0 else l[t - 1]
        At = _reshape1(A[t])
        for u in range(t + 1, min(L - 1, t + maxdist) + 1):
            ru = np.eye(d) if u == L - 1 else r[u + 1]
            Au = _reshape1(A[u])
            rl = ru @ lt
            rlAt = np.einsum("ua,abx->ubx", rl, At)
            rlAtM = np.einsum("ubx,bc->uxc", rlAt, M[(t, u)])
            btu = np.einsum("uxc,cuy->xy", rlAtM, Au)
            btu = btu / btu.sum()
            b[(t, u)] = btu.reshape(qs, order="F")
    return b


# ----------------------------------------------------------------------------------------------
# sampling  (src/abstract_tensor_train.jl:335-384, src/utils.jl:9-20)
# ----------------------------------------------------------------------------------------------
class NegativeTTError(RuntimeError):
    pass


def sample_noalloc(u: float, w: np.ndarray) -> int:
    """Inverse-CDF draw, strict ``cumsum > u`` (``src/utils.jl:9-20``); returns 0-based index."""
    cw = 0.0
    for i, p in enumerate(w):
        cw += float(p)
        if cw > u:
            return i
    raise AssertionError(f"{w}")


def sample(uniform: Callable[[int], float], A, rz=None):
    """``sample!`` (``src/abstract_tensor_train.jl:355-381``).  ``uniform(t)`` supplies the one
    ``rand(rng)`` consumed at 0-based site ``t``.  Returns (flat 0-based x per site, p)."""
    r, z = accumulate_R(A) if rz is None else rz
    L = len(A)
    d = A[0].shape[0]
    Q = np.eye(d)
    xs = []
    for t in range(L):
        rt = np.eye(d) if t == L - 1 else r[t + 1]
        At = _reshape1(A[t])
        QA = np.einsum("km,mnx->knx", Q, At)
        p = np.einsum("knx,nk->x", QA, rt)
        if np.any(p < 0):
            raise NegativeTTError("Cannot sample from a tensor train with negative values")
        p = p / p.sum()
        xt = sample_noalloc(uniform(t), p)
        xs.append(xt)
        Q = Q @ At[:, :, xt]
    pr = float(LogNum(float(np.trace(Q))) / z)
    return xs, pr


# ----------------------------------------------------------------------------------------------
# dot / norm  (src/abstract_tensor_train.jl:395-437)  -- "next" row, needed by the tests now
# ----------------------------------------------------------------------------------------------
def dot(A, B) -> float:
    """``src/abstract_tensor_train.jl:395-408``."""
    L = len(A)
    AL, BL = _reshape1(A[L - 1]), _reshape1(B[L - 1])
    C = np.einsum("apx,bqx->apqb", AL, BL)
    for t in range(L - 2, -1, -1):
        Al, Bl = _reshape1(A[t]), _reshape1(B[t])
        C = np.einsum("aex,epqf,bfx->apqb", Al, C, Bl, optimize=True)
    d = float(np.einsum("aabb->", C))
    return d / float(A.z * B.z)


def norm(A) -> float:
    """``src/abstract_tensor_train.jl:419``."""
    return math.sqrt(dot(A, A))


def norm2m(A, B) -> float:
    """``src/abstract_tensor_train.jl:430-432``."""
    return norm(A) ** 2 + norm(B) ** 2 - 2 * dot(A, B)


def isapprox(A, B, atol=0.0, rtol=1e-6) -> bool:
    """``src/abstract_tensor_train.jl:434-437``."""
    na, nb = norm(A), norm(B)
    return na ** 2 + nb ** 2 - 2 * dot(A, B) <= max(atol, rtol * max(na, nb)) ** 2



# ----------------------------------------------------------------------------------------------
# two-site pieces (2-site DMRG building blocks): src/utils.jl:37-44, src/abstract_tensor_train.jl:283-308,
# src/tensor_train.jl:320-334, src/dmrg.jl:124-151
# ----------------------------------------------------------------------------------------------
def merge_tensors(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """``_merge_tensors`` (``utils.jl:37-44``): ``C[m,n,xA...,xB...] = sum_k A[m,k,xA] B[k,n,xB]``"""
    sA, sB = A.shape[2:], B.shape[2:]
    assert len(sA) == len(sB)
    C = np.einsum("mka,knb->mnab", _reshape1(A), _reshape1(B))
    return C.reshape((C.shape[0], C.shape[1]) + tuple(sA) + tuple(sB), order="F")


def split_tensor(C: np.ndarray, svd_trunc: Optional[SVDTrunc] = None, lr: str = "left"):
    """``_split_tensor`` (``abstract_tensor_train.jl:283-308``): ``C_resh[(al,xl),(al1,xl1)] = U L V'``;
    ``lr="left"``: ``A = U, B = L V'``; ``lr="right"``: ``A = U L, B = V'``"""
    svd_trunc = TruncThresh(0.0) if svd_trunc is None else svd_trunc
    sC = C.shape[2:]
    h = len(sC) // 2
    sA, sB = sC[:h], sC[h:]
    qa, qb = int(np.prod(sA)), int(np.prod(sB))
    Cr = C.reshape((C.shape[0], C.shape[1], qa, qb), order="F")
    # (al, xl) with al fastest, (al1, xl1) with al1 fastest (TensorCast)
    M = np.transpose(Cr, (0, 2, 1, 3)).reshape((C.shape[0] * qa, C.shape[1] * qb), order="F")
    U, lam, V = svd_trunc(M)
    if lr == "left":
        Ar, Br = U, lam[:, None] * V.T
    else:
        Ar, Br = U * lam[None, :], V.T
    m = len(lam)
    A = Ar.reshape((C.shape[0], qa, m), order="F").transpose(0, 2, 1)          # A_[i,j,x] = A_resh[(i,x),j]
    B = Br.reshape((m, C.shape[1], qb), order="F")                             # B_[i,j,x] = B_resh[i,(j,x)]
    return (np.ascontiguousarray(A).reshape((C.shape[0], m) + tuple(sA), order="F"),
            np.ascontiguousarray(B).reshape((m, C.shape[1]) + tuple(sB), order="F"))


def precompute_left_environments(A, X):
    """``tensor_train.jl:320-326``: list indexed 0..L, entry j = ``A_1(x_1)...A_j(x_j)`` (entry 0 = ones(1,1))"""
    out = [np.ones((1, 1))]
    for Ak, xk in zip(A.tensors, X):
        out.append(out[-1] @ Ak[(slice(None), slice(None)) + tuple(xk)])
    return out


def precompute_right_environments(A, X):
    """``tensor_train.jl:328-334``: list indexed 0..L, entry j = ``A_{j+1}(x_{j+1})...A_L(x_L)`` (the reference's
    ``prodA_right[j+1]``; entry L = ones(1,1))"""
    out = [np.ones((1, 1))]
    for Ak, xk in zip(reversed(A.tensors), reversed(list(X))):
        out.append(Ak[(slice(None), slice(None)) + tuple(xk)] @ out[-1])
    return out[::-1]


def update_environments(prod_left, prod_right, A, k: int, X, lr: str):
    """``update_environments!`` (``dmrg.jl:124-151``), ``k`` the 0-based first site of the pair just split; the
    lists are those of :func:`precompute_left_environments` / :func:`precompute_right_environments` per datapoint."""
    for n, x in enumerate(X):
        if lr == "left":
            Anew = A[k][(slice(None), slice(None)) + tuple(x[k])]
            prod_left[n][k + 1] = prod_left[n][k] @ Anew
        else:
            Anew = A[k + 1][(slice(None), slice(None)) + tuple(x[k + 1])]
            prod_right[n][k + 1] = Anew @ prod_right[n][k + 2]


# ----------------------------------------------------------------------------------------------
# MPS Born-machine wrapper (src/MatrixProductStates/mps.jl), real Float64 (conj is a no-op)
# ----------------------------------------------------------------------------------------------
class MPS:
    """``p(x) = |psi(x)|^2`` over a wrapped train ``psi`` (``mps.jl:19-21``)."""

    def __init__(self, psi):
        self.psi = psi

    def __len__(self):
        return len(self.psi)


def _id4(d):
    """``id4`` (``mps.jl:56``): ``[a == a1 && b == b1]`` indexed ``[a, a1, b, b1]``"""
    e = np.eye(d)
    return np.einsum("ac,bd->acbd", e, e)


def mps_evaluate(p: MPS, X) -> float:
    """``mps.jl:54``"""
    return abs(evaluate(p.psi, X)) ** 2


def mps_accumulate_L(p: MPS, normalize=True):
    """``mps.jl:60-80``: ``Ll[a1, al, b1, bl]``, hermiticity restored after every step"""
    psi = p.psi
    Ll = _id4(psi[0].shape[0])
    z = LogNum(1.0)
    out = []
    for Al in psi:
        A = _reshape1(Al)
        nl = np.max(np.abs(Ll))
        if nl != 0 and normalize:
            Ll /= nl
            z = z * nl
        Ll = np.einsum("acbd,cex,dfx->aebf", Ll, A, A, optimize=True)
        Ll = (np.transpose(Ll, (2, 3, 0, 1)) + Ll) / 2
        out.append(Ll)
    z = z * float(np.einsum("aabb->", Ll))
    return out, z


def mps_accumulate_R(p: MPS, normalize=True):
    """``mps.jl:82-103``: ``Rl[al, aL, bl, bL]``"""
    psi = p.psi
    Rl = _id4(psi[len(psi) - 1].shape[1])
    z = LogNum(1.0)
    out = []
    for Al in reversed(psi.tensors):
        A = _reshape1(Al)
        nl = np.max(np.abs(Rl))
        if nl != 0 and normalize:
            Rl /= nl
            z = z * nl
        Rl = np.einsum("ecfd,aex,bfx->acbd", Rl, A, A, optimize=True)
        Rl = (np.transpose(Rl, (2, 3, 0, 1)) + Rl) / 2
        out.append(Rl)
    z = z * float(np.einsum("aabb->", Rl))
    return out[::-1], z


def mps_normalization(p: MPS) -> LogNum:
    """``mps.jl:139-146``: ``z / abs2(psi.z)``"""
    _, z = mps_accumulate_L(p)
    return z / (p.psi.z * p.psi.z)


def mps_normalize(p: MPS) -> float:
    """``normalize!`` (``mps.jl:154-168``): tensors divided by ``|Z|^(1/2L)``, ``psi.z = 1``, returns
    ``log(|Z| / abs2(psi.z))``"""
    Z = mps_accumulate_L(p)[1]
    L = len(p)
    x = math.exp(0.5 * Z.logabs / L)
    if x != 0:
        for a in p.psi.tensors:
            a /= x
    logz = Z.logabs - 2.0 * p.psi.z.logabs
    p.psi.z = LogNum(1.0)
    return logz


def mps_marginals(p: MPS):
    """``mps.jl:178-196``"""
    psi = p.psi
    l, _ = mps_accumulate_L(p)
    r, _ = mps_accumulate_R(p)
    d = psi[0].shape[0]
    out = []
    for t in range(len(psi)):
        A = _reshape1(psi[t])
        R = r[t + 1] if t + 1 < len(psi) else _id4(d)
        Lm = l[t - 1] if t >= 1 else _id4(d)
        pt = np.einsum("acbd,cex,dfx,eafb->x", Lm, A, A, R, optimize=True)
        pt = pt / pt.sum()
        out.append(pt.reshape(psi[t].shape[2:], order="F"))
    return out


def mps_marginals_env(p: MPS, l=None, r=None):
    """``marginals(p; l, r)`` with caller-supplied environments (``mps.jl:176-177``)"""
    psi = p.psi
    l = mps_accumulate_L(p)[0] if l is None else l
    r = mps_accumulate_R(p)[0] if r is None else r
    d = psi[0].shape[0]
    out = []
    for t in range(len(psi)):
        A = _reshape1(psi[t])
        R = r[t + 1] if t + 1 < len(psi) else _id4(d)
        Lm = l[t - 1] if t >= 1 else _id4(d)
        pt = np.einsum("acbd,cex,dfx,eafb->x", Lm, A, A, R, optimize=True)
        out.append((pt / pt.sum()).reshape(psi[t].shape[2:], order="F"))
    return out


def mps_accumulate_M(p: MPS):
    """``mps.jl:104-126``: ``M[t,u][a_{t+1}, a_u, b_{t+1}, b_u]`` (0-based keys ``(t,u)``, ``t < u``), no rescaling,
    hermiticity restored after every step, ``M[u-1,u] = id4(d_u)``"""
    L = len(p)
    M = {}
    for u in range(1, L):
        Au = _reshape1(p.psi[u - 1])
        for t in range(0, u - 1):
            Mu = np.einsum("acbd,cex,dfx->aebf", M[(t, u - 1)], Au, Au, optimize=True)
            M[(t, u)] = (np.transpose(Mu, (2, 3, 0, 1)) + Mu) / 2
        M[(u - 1, u)] = _id4(p.psi[u].shape[0])
    return M


def mps_twovar_marginals(p: MPS, l=None, r=None, M=None, maxdist=None):
    """``mps.jl:206-232`` (open trains): dict ``(t,u) -> b[x_t..., x_u...]``"""
    psi = p.psi
    L = len(psi)
    l = mps_accumulate_L(p)[0] if l is None else l
    r = mps_accumulate_R(p)[0] if r is None else r
    M = mps_accumulate_M(p) if M is None else M
    maxdist = L - 1 if maxdist is None else maxdist
    d = psi[0].shape[0]
    qs = tuple(psi[0].shape[2:])
    out = {}
    for t in range(L - 1):
        lt = _id4(d) if t == 0 else l[t - 1]
        At = _reshape1(psi[t])
        for u in range(t + 1, min(L - 1, t + maxdist) + 1):
            ru = _id4(d) if u == L - 1 else r[u + 1]
            Au = _reshape1(psi[u])
            lAAM = np.einsum("acbd,cex,dfx,egfh->abghx", lt, At, At, M[(t, u)], optimize=True)
            AAr = np.einsum("gky,hly,kalb->abghy", Au, Au, ru, optimize=True)
            btu = np.einsum("abghx,abghy->xy", lAAM, AAr, optimize=True)
            btu = btu / btu.sum()
            out[(t, u)] = btu.reshape(qs + qs, order="F")
    return out


def mps_sample(uniform: Callable[[int], float], p: MPS, rz=None):
    """``sample!`` for an MPS (``mps.jl:264-290``): ``uniform(t)`` supplies the one ``rand`` of site ``t``.
    Returns ``(x, q)`` with ``x`` 0-based multi-indices and ``q = |tr Q|^2 / z``."""
    psi = p.psi
    r, z = mps_accumulate_R(p) if rz is None else rz
    L = len(psi)
    d = psi[L - 1].shape[1]
    Qm = np.eye(d)
    x = []
    for t in range(L):
        rl = _id4(d) if t == L - 1 else r[t + 1]
        A = _reshape1(psi[t])
        QA = np.einsum("km,mnx->knx", Qm, A)
        q = np.einsum("acx,cadb,bdx->x", QA, rl, QA, optimize=True)
        q = q / q.sum()
        xt = sample_noalloc(uniform(t), q)
        x.append(tuple(int(v) for v in np.unravel_index(xt, psi[t].shape[2:], order="F")))
        Qm = QA[:, :, xt]
    return x, float(np.trace(Qm) ** 2 / float(z))


def mps_exact_prob(p: MPS) -> np.ndarray:
    """``test/exact.jl`` for an MPS: enumerate ``|psi(x)|^2``"""
    return np.abs(_exact_values(p.psi)) ** 2


def _exact_values(A) -> np.ndarray:
    """all values ``tr(prod A_t(x_t)) / z`` as an array over the flattened configurations (site-major)"""
    L = len(A)
    Ts = [_reshape1(a) for a in A]
    cur = np.moveaxis(Ts[0], 2, 0)  # (x, a, b)
    for t in range(1, L):
        nxt = np.moveaxis(Ts[t], 2, 0)
        cur = np.einsum("kab,xbc->kxac", cur, nxt).reshape(-1, cur.shape[1], nxt.shape[2])
    return np.einsum("kaa->k", cur) / float(A.z)

# ----------------------------------------------------------------------------------------------
# brute force  (test/exact.jl:1-50)
# ----------------------------------------------------------------------------------------------
def exact_prob(A) -> np.ndarray:
    """``test/exact.jl:10-16``: all ``Q^L`` evaluations, array indexed by flat x per site."""
    Qs = [int(np.prod(a.shape[2:])) for a in A]
    out = np.zeros(Qs)
    A3 = [_reshape1(a) for a in A]
    for x in itertools.product(*[range(Q) for Q in Qs]):
        P = None
        for a, xi in zip(A3, x):
            P = a[:, :, xi] if P is None else P @ a[:, :, xi]
        out[x] = float(LogNum(float(np.trace(P))) / A.z)
    return out


def exact_normalization(A) -> float:
    """``test/exact.jl:1-8``."""
    return float(exact_prob(A).sum())


def exact_marginals(A, p=None):
    """``test/exact.jl:18-26`` (flat physical index)."""
    p = exact_prob(A) if p is None else p
    out = []
    for l in range(len(A)):
        v = p.sum(axis=tuple(i for i in range(len(A)) if i != l))
        out.append(v / v.sum())
    return out


def exact_twovar_marginals(A, p=None):
    """``test/exact.jl:28-41`` (flat physical indices), dict keyed by 0-based ``(l,m)``, l<m."""
    p = exact_prob(A) if p is None else p
    L = len(A)
    out = {}
    for l in range(L):
        for m in range(l + 1, L):
            v = p.sum(axis=tuple(i for i in range(L) if i not in (l, m)))
            out[(l, m)] = v / v.sum()
    return out


def exact_norm(A, p=None) -> float:
    """``test/exact.jl:43-45``."""
    p = exact_prob(A) if p is None else p
    return math.sqrt(float(np.sum(p ** 2)))


def exact_dot(A, B) -> float:
    """``test/exact.jl:47-50``."""
    return float(np.sum(exact_prob(A) * exact_prob(B)))


# ----------------------------------------------------------------------------------------------
# Philox4x32-10 uniforms, the stream the CUDA sampler consumes (our harness, not the reference):
# key = (seed_lo, seed_hi), counter = (sample_lo, sample_hi, site, 0); u = 53 random bits / 2^53.
# ----------------------------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al. 2011).  All inputs uint32 arrays (broadcastable)."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)]
    k0 = np.asarray(k0, dtype=np.uint32)
    k1 = np.asarray(k1, dtype=np.uint32)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PHILOX_M0 * c0.astype(np.uint64)
            p1 = _PHILOX_M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = (k0 + _PHILOX_W0).astype(np.uint32)
            k1 = (k1 + _PHILOX_W1).astype(np.uint32)
    return c0, c1, c2, c3


def philox_uniform(seed: int, sample_index, site) -> np.ndarray:
    """The uniform in [0,1) the CUDA sampler draws for (seed, global sample index, site)."""
    s = np.asarray(sample_index, dtype=np.uint64)
    t = np.asarray(site, dtype=np.uint32)
    r0, r1, _, _ = philox4x32_10((s & np.uint64(0xFFFFFFFF)).astype(np.uint32),
                                 (s >> np.uint64(32)).astype(np.uint32), t, np.uint32(0),
                                 np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF))
    bits = (r0.astype(np.uint64) << np.uint64(21)) ^ (r1.astype(np.uint64) >> np.uint64(11))
    return bits.astype(np.float64) * (1.0 / 9007199254740992.0)


# ----------------------------------------------------------------------------------------------
# is_in_domain, ==  (src/abstract_tensor_train.jl:59-64, :85)
# ----------------------------------------------------------------------------------------------
def is_in_domain(A, X) -> bool:
    """``src/abstract_tensor_train.jl:59-64`` with 0-based indices: ``X`` is a list over sites of index tuples."""
    return all(all(0 <= int(xi) < Ax.shape[i + 2] for i, xi in enumerate(np.atleast_1d(x))) for Ax, x in zip(A, X))


def equal(A, B) -> bool:
    """``==`` (``src/abstract_tensor_train.jl:85``): ``isequal(A.tensors, B.tensors)`` for trains of the same type --
    elementwise ``isequal`` (NaN equals NaN, -0.0 differs from 0.0); ``z`` is not compared."""
    if type(A) is not type(B) or len(A) != len(B):
        return False
    for a, b in zip(A, B):
        if a.shape != b.shape:
            return False
        same = (np.isnan(a) & np.isnan(b)) | (a.view(np.int64) == b.view(np.int64)) if a.size else np.array([True])
        if not bool(np.all(same)):
            return False
    return True
