"""ttb200 -- host-side mirror of the TensorTrains.jl hot-path API over libttb200.so (sm_100a).

The names, argument meaning and error behaviour follow the reference (stecrotti/TensorTrains.jl
v0.14.0; file:line citations are relative to its root).  Julia's ``f!`` becomes ``f_``.  Site
tensors live in device memory behind an opaque handle of the C ABI declared in ``include/ttb200.h``;
every function here is a thin ctypes call -- there is NO CPU fallback: importing this package fails
if the CUDA library is missing, and every call fails if no CUDA device is present.

Conventions carried over from the reference:
  * a site tensor has shape ``(d_t, d_{t+1}, q_1, ...)``, column-major like Julia arrays;
  * physical indices passed to/returned from this module are 0-based (Julia's are 1-based);
  * ``z`` is a :class:`Logarithmic` (sign, log|z|) like ``LogarithmicNumbers.Logarithmic``.
A handle may hold a *batch* of independent trains with identical shapes (the reference would hold a
``Vector`` of trains); batched objects return arrays with a leading batch axis.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

__all__ = [
    "TensorTrain", "PeriodicTensorTrain", "Logarithmic",
    "TruncThresh", "TruncBond", "TruncBondMax", "TruncBondThresh",
    "flat_tt", "rand_tt", "randn_tt", "flat_periodic_tt", "rand_periodic_tt",
    "bond_dims", "nparams", "evaluate", "accumulate_L", "accumulate_R", "normalization", "lognormalization",
    "normalize_", "normalize_eachmatrix_", "marginals", "twovar_marginals", "twovar_marginals_block", "compress_",
    "orthogonalize_right_", "orthogonalize_left_", "orthogonalize_center_", "orthogonalize_two_site_center_",
    "is_left_canonical", "is_right_canonical", "is_canonical", "sample", "sample_stats", "lib", "TTBError",
    "Environment", "accumulate_M", "is_in_domain",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libttb200.so")
if not os.path.exists(_LIB_PATH):
    raise ImportError(f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback)")
lib = C.CDLL(_LIB_PATH)

_h = C.c_void_p
_i64 = C.c_int64
_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)
_pi64 = C.POINTER(C.c_int64)


class _Trunc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("eps", C.c_double), ("mprime", C.c_int64)]


def _sig(name, *args):
    f = getattr(lib, name)
    f.restype = C.c_int
    f.argtypes = list(args)
    return f


lib.ttb_last_error.restype = C.c_char_p
_sig("ttb_device_count", _pi)
_sig("ttb_set_device", C.c_int)
_sig("ttb_host_alloc", C.POINTER(C.c_void_p), C.c_size_t)
_sig("ttb_host_free", C.c_void_p)
_sig("ttb_create", _i64, _pi64, _i64, C.c_int, _i64, C.POINTER(_h))
_sig("ttb_destroy", _h)
_sig("ttb_clone", _h, C.POINTER(_h))
_sig("ttb_compose", _h, _h, C.c_int, C.POINTER(_h))
_sig("ttb_info", _h, _pi64, _pi64, _pi, _pi64)
_sig("ttb_get_bonds", _h, _i64, _pi64)
_sig("ttb_get_capacity_bonds", _h, _pi64)
_sig("ttb_nparams", _h, _i64, _pi64)
_sig("ttb_upload_site", _h, _i64, _i64, _pd)
_sig("ttb_download_site", _h, _i64, _i64, _pd)
_sig("ttb_upload_all", _h, _pd)
_sig("ttb_download_all", _h, _pd)
_sig("ttb_get_z", _h, _i64, _pd, _pi)
_sig("ttb_set_z", _h, _i64, C.c_double, C.c_int)
_sig("ttb_orthogonalize_right", _h, C.POINTER(_Trunc), _i64, _i64)
_sig("ttb_orthogonalize_left", _h, C.POINTER(_Trunc), _i64, _i64)
_sig("ttb_orthogonalize_center", _h, C.POINTER(_Trunc), _i64)
_sig("ttb_orthogonalize_two_site_center", _h, C.POINTER(_Trunc), _i64)
_sig("ttb_compress", _h, C.POINTER(_Trunc), C.c_int)
_sig("ttb_sweep_stats", _h, _i64, _pi64, _pd)
_sig("ttb_record_spectrum", _h, C.c_int)
_sig("ttb_get_spectrum", _h, _i64, _i64, _pd, _i64, _pi64)
_sig("ttb_canonical_residual", _h, _i64, _i64, C.c_int, _pd)
_sig("ttb_env_create", _h, C.c_int, C.c_int, C.POINTER(_h))
_sig("ttb_env_destroy", _h)
_sig("ttb_env_z", _h, _pd, _pi)
_sig("ttb_marginals_env", _h, _h, _h, _pd)
_sig("ttb_twovar_marginals_env", _h, _h, _h, _i64, _i64, _i64, _pd)
_sig("ttb_accumulate_M_size", _h, _i64, _pi64)
_sig("ttb_accumulate_M", _h, _i64, C.c_int, _pd)
_sig("ttb_is_in_domain", _h, C.POINTER(C.c_int32), _i64, _pi)
_sig("ttb_equal", _h, _i64, _h, _i64, _pi)
_sig("ttb_sample_env", _h, _h, _i64, _i64, C.c_uint64, C.c_uint64, _pd, C.POINTER(C.c_uint8), _pd)
_sig("ttb_sample_stats_env", _h, _h, _i64, _i64, C.c_uint64, C.c_uint64, _pi64, _pd)
_sig("ttb_env_size", _h, _i64, C.c_int, _pi64)
_sig("ttb_accumulate_L", _h, C.c_int, _pd, _pd, _pi)
_sig("ttb_accumulate_R", _h, C.c_int, _pd, _pd, _pi)
_sig("ttb_normalization", _h, _pd, _pi)
_sig("ttb_normalize", _h, _pd)
_sig("ttb_normalize_eachmatrix", _h)
_sig("ttb_scale_sites", _h, _pd)
_sig("ttb_twovar_marginals_range", _h, _i64, _i64, _i64, _pd)
_sig("ttb_marginals", _h, _pd)
_sig("ttb_twovar_count", _h, _i64, _pi64)
_sig("ttb_twovar_marginals", _h, _i64, _pd)
_sig("ttb_evaluate", _h, _i64, C.POINTER(C.c_int32), _i64, _pd)
_sig("ttb_sample", _h, _i64, _i64, C.c_uint64, C.c_uint64, _pd, C.POINTER(C.c_uint8), _pd)
_sig("ttb_sample_stats", _h, _i64, _i64, C.c_uint64, C.c_uint64, _pi64, _pd)
_sig("ttb_upload_all_async", _h, _pd)
_sig("ttb_reset", _h)
_sig("ttb_trim", _h)
_sig("ttb_dot", _h, _i64, _h, _i64, _pd, _pi)
_sig("ttb_merge_tensors", _h, _i64, _i64, _pd)
_sig("ttb_split_tensor", _h, _i64, _i64, _pd, C.POINTER(_Trunc), C.c_int, _pd)
_sig("ttb_data_environments", _h, _i64, C.POINTER(C.c_int32), _i64, C.c_int, _pd)
_sig("ttb_data_env_update", _h, _i64, C.POINTER(C.c_int32), _i64, _i64, C.c_int, _pd)
_sig("ttb_mps_env_size", _h, _i64, C.c_int, _pi64)
_sig("ttb_mps_accumulate", _h, _i64, C.c_int, C.c_int, _pd, _pd, _pi)
_sig("ttb_mps_marginals", _h, _i64, _pd)
_sig("ttb_mps_marginals_env", _h, _i64, _pd, _pd, _pd)
_sig("ttb_mps_twovar_marginals_env", _h, _i64, _pd, _pd, _i64, _pd)
_sig("ttb_mps_sample_env", _h, _i64, _pd, C.c_double, _i64, C.c_uint64, C.c_uint64, _pd, C.POINTER(C.c_uint8), _pd)
_sig("ttb_mps_twovar_marginals", _h, _i64, _i64, _pd)
_sig("ttb_mps_sample", _h, _i64, _i64, C.c_uint64, C.c_uint64, _pd, C.POINTER(C.c_uint8), _pd)
_sig("ttb_last_timing", _h, _pd, _pi64)
_sig("ttb_profile", _h, C.c_int)
_sig("ttb_profile_report", _h, C.c_char_p, C.c_size_t)
lib.ttb_version.restype = C.c_int


class TTBError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class NegativeValuesError(TTBError):
    """ErrorException("Cannot sample from a tensor train with negative values")
    (src/abstract_tensor_train.jl:371)."""


_EINVAL, _ESHAPE, _ENEGATIVE, _ENONFINITE, _ECUDA, _ENOMEM, _EUNSUPPORTED, _EDOMAIN, _ENOCONV = 1, 2, 3, 4, 5, 6, 7, 8, 9


class SVDNotConvergedError(TTBError):
    """LAPACKException thrown by ``svd`` (src/svd_trunc.jl:37) when it fails to converge."""


def _ck(rc):
    if rc == 0:
        return
    msg = lib.ttb_last_error().decode("utf-8", "replace")
    if rc in (_EINVAL, _ESHAPE):
        raise ValueError("ArgumentError: " + msg)          # Julia ArgumentError
    if rc == _EDOMAIN:
        raise ValueError("DomainError: " + msg)
    if rc == _ENEGATIVE:
        raise NegativeValuesError(rc, msg)
    if rc == _ENOCONV:
        raise SVDNotConvergedError(rc, msg)
    if rc == _ENOMEM:
        raise MemoryError(msg)
    if rc == _EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise TTBError(rc, msg)


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_pd)


def device_count() -> int:
    n = C.c_int(0)
    _ck(lib.ttb_device_count(C.byref(n)))
    return n.value


def set_device(dev: int):
    _ck(lib.ttb_set_device(int(dev)))


# ------------------------------------------------------------------------------------------------
class Logarithmic:
    """sign * exp(logabs); mirrors LogarithmicNumbers.Logarithmic as used for ``z``
    (src/tensor_train.jl:10,12)."""
    __slots__ = ("logabs", "sign")

    def __init__(self, x=1.0, *, logabs=None, sign=1):
        if logabs is not None:
            self.logabs, self.sign = float(logabs), (1 if sign >= 0 else -1)
        else:
            x = float(x)
            self.sign = -1 if x < 0 else 1
            self.logabs = -math.inf if x == 0 else math.log(abs(x))

    def __float__(self):
        return self.sign * math.exp(self.logabs)

    def __repr__(self):
        return f"{'+' if self.sign > 0 else '-'}exp({self.logabs})"


# ---- SVD truncators (src/svd_trunc.jl) -----------------------------------------------------------
class SVDTrunc:
    def _c(self) -> _Trunc:  # pragma: no cover
        raise NotImplementedError

    def _after(self, A):
        pass


@dataclass
class TruncThresh(SVDTrunc):
    """src/svd_trunc.jl:29-50"""
    eps: float

    def _c(self):
        return _Trunc(0, 0, float(self.eps), 0)

    def summary(self):
        return f"SVD truncation with threshold ε={self.eps}"


@dataclass
class TruncBond(SVDTrunc):
    """src/svd_trunc.jl:66-77"""
    mprime: int

    def _c(self):
        return _Trunc(1, 0, 0.0, int(self.mprime))

    def summary(self):
        return f"SVD truncation to bond size m'={self.mprime}"


@dataclass
class TruncBondMax(SVDTrunc):
    """src/svd_trunc.jl:91-103: also keeps the largest truncation error seen so far."""
    mprime: int
    maxerr: List[float] = field(default_factory=lambda: [0.0])

    def _c(self):
        return _Trunc(2, 0, 0.0, int(self.mprime))

    def _after(self, A):
        for b in range(A.batch):
            _, err = A.sweep_stats(b)
            self.maxerr[0] = max(self.maxerr[0], float(np.max(err)))

    def summary(self):
        return f"SVD truncation to bond size m'={self.mprime}. Max error {self.maxerr[0]}"


@dataclass
class TruncBondThresh(SVDTrunc):
    """src/svd_trunc.jl:118-143"""
    mprime: int
    eps: float = 0.0

    def _c(self):
        return _Trunc(3, 0, float(self.eps), int(self.mprime))

    def summary(self):
        return ("SVD truncation with truncation to bond size given by the minimum of threshold "
                f"ε={self.eps} and m'={self.mprime}")


# ---- containers ----------------------------------------------------------------------------------
class AbstractTensorTrain:
    """Device-resident train(s).  ``tensors``: list over sites of arrays ``(d_t, d_{t+1}, q...)``
    (or, with ``batched=True``, ``(B, d_t, d_{t+1}, q...)``)."""
    periodic = False

    def __init__(self, tensors: Sequence[np.ndarray], z=1.0, batched: bool = False, _handle=None, _q=None):
        self._hd = _h()
        if _handle is not None:
            self._hd, self.q, self.batched = _handle, tuple(_q), batched
            self._refresh()
            return
        tensors = [np.asarray(a, dtype=np.float64) for a in tensors]
        if not batched:
            tensors = [a[None] for a in tensors]
        if any(a.ndim <= 3 for a in tensors):
            raise ValueError("ArgumentError: Tensors should have at least 3 indices: 2 virtual and 1 physical")
        B = tensors[0].shape[0]
        self.q = tuple(tensors[0].shape[3:])
        L = len(tensors)
        if any(tuple(a.shape[3:]) != self.q or a.shape[0] != B for a in tensors):
            raise ValueError("ArgumentError: all sites must share the physical dimensions and the batch size")
        for t in range(L):
            if tensors[t].shape[2] != tensors[(t + 1) % L].shape[1] and (self.periodic or t + 1 < L):
                raise ValueError("ArgumentError: Matrix indices for matrix product non compatible")
        bond = np.array([a.shape[1] for a in tensors] + [tensors[-1].shape[2]], dtype=np.int64)
        self.batched = batched
        _ck(lib.ttb_create(L, bond.ctypes.data_as(_pi64), int(np.prod(self.q)), int(self.periodic), B,
                           C.byref(self._hd)))
        self._refresh()
        # pack train-major, site-major, each site column-major
        Qf = int(np.prod(self.q))
        buf = np.concatenate([np.reshape(tensors[t][b].reshape(tensors[t].shape[1], tensors[t].shape[2], Qf, order="F"),
                                         -1, order="F") for b in range(B) for t in range(L)])
        _ck(lib.ttb_upload_all(self._hd, _dp(np.ascontiguousarray(buf))))
        if not (isinstance(z, (int, float)) and z == 1.0):
            zz = z if isinstance(z, Logarithmic) else Logarithmic(z)
            for b in range(B):
                _ck(lib.ttb_set_z(self._hd, b, zz.logabs, zz.sign))

    def _dmax(self) -> int:
        """largest bond the handle was created with (the stride of data-environment entries)"""
        out = np.zeros(self.L + 1, dtype=np.int64)
        _ck(lib.ttb_get_capacity_bonds(self._hd, out.ctypes.data_as(_pi64)))
        return int(out.max())

    def _refresh(self):
        L, Q, per, B = _i64(), _i64(), C.c_int(), _i64()
        _ck(lib.ttb_info(self._hd, C.byref(L), C.byref(Q), C.byref(per), C.byref(B)))
        self.L, self.Q, self.batch = L.value, Q.value, B.value

    def __del__(self):
        try:
            if self._hd:
                lib.ttb_destroy(self._hd)
                self._hd = _h()
        except Exception:
            pass

    def __len__(self):
        return self.L

    def copy(self):
        """deepcopy(A)"""
        nh = _h()
        _ck(lib.ttb_clone(self._hd, C.byref(nh)))
        return type(self)(None, batched=self.batched, _handle=nh, _q=self.q)

    __deepcopy__ = lambda self, memo: self.copy()

    # -- metadata
    def bonds(self, b: int = 0) -> np.ndarray:
        out = np.zeros(self.L + 1, dtype=np.int64)
        _ck(lib.ttb_get_bonds(self._hd, b, out.ctypes.data_as(_pi64)))
        return out

    @property
    def z(self):
        zs = []
        for b in range(self.batch):
            la, sg = C.c_double(), C.c_int()
            _ck(lib.ttb_get_z(self._hd, b, C.byref(la), C.byref(sg)))
            zs.append(Logarithmic(logabs=la.value, sign=sg.value))
        return zs if self.batched else zs[0]

    @z.setter
    def z(self, v):
        vs = v if isinstance(v, (list, tuple)) else [v] * self.batch
        for b, x in enumerate(vs):
            x = x if isinstance(x, Logarithmic) else Logarithmic(x)
            _ck(lib.ttb_set_z(self._hd, b, x.logabs, x.sign))

    # -- site access (getindex / setindex!, src/tensor_train.jl:27-29); copies through the host
    def site(self, t: int, b: int = 0) -> np.ndarray:
        bd = self.bonds(b)
        out = np.zeros(int(bd[t] * bd[t + 1] * self.Q))
        _ck(lib.ttb_download_site(self._hd, b, t, _dp(out)))
        return out.reshape((int(bd[t]), int(bd[t + 1])) + self.q, order="F")

    def __getitem__(self, t):
        return self.site(t % self.L)

    def set_site(self, t: int, a: np.ndarray, b: int = 0):
        bd = self.bonds(b)
        a = np.asarray(a, dtype=np.float64)
        if a.shape[:2] != (bd[t], bd[t + 1]) or a.size != bd[t] * bd[t + 1] * self.Q:
            raise ValueError("ArgumentError: site shape does not match the current bond dimensions")
        _ck(lib.ttb_upload_site(self._hd, b, t, _dp(np.ascontiguousarray(a.reshape(-1, order="F")))))

    def tensors(self, b: int = 0) -> List[np.ndarray]:
        return [self.site(t, b) for t in range(self.L)]

    def sweep_stats(self, b: int = 0):
        rank = np.zeros(self.L + 1, dtype=np.int64)
        err = np.zeros(self.L + 1)
        _ck(lib.ttb_sweep_stats(self._hd, b, rank.ctypes.data_as(_pi64), _dp(err)))
        return rank, err

    def record_spectrum(self, on: bool = True):
        _ck(lib.ttb_record_spectrum(self._hd, int(on)))

    def _compose(self, other, sign: int):
        if type(self) is not type(other):
            raise ValueError("ArgumentError: cannot compose trains of different kinds")
        if tuple(self.q) != tuple(other.q):
            raise ValueError("ArgumentError: Tensor Trains must have the types of physical indices")
        nh = _h()
        _ck(lib.ttb_compose(self._hd, other._hd, sign, C.byref(nh)))
        return type(self)(None, batched=self.batched, _handle=nh, _q=self.q)

    def __add__(self, other):
        """``A + B`` (src/abstract_tensor_train.jl:315; `_compose` src/tensor_train.jl:240-261,
        src/periodic_tensor_train.jl:62-78): bond dimensions add up, assembled on the device."""
        return self._compose(other, 1)

    def __sub__(self, other):
        """``A - B`` (src/abstract_tensor_train.jl:322)"""
        return self._compose(other, -1)

    def __eq__(self, other):
        """``==`` (src/abstract_tensor_train.jl:85): ``isequal`` of the site tensors; ``z`` is not compared."""
        if not isinstance(other, AbstractTensorTrain) or type(self) is not type(other) or self.batch != other.batch \
                or self.q != other.q:
            return False
        eq = C.c_int(0)
        for b in range(self.batch):
            _ck(lib.ttb_equal(self._hd, b, other._hd, b, C.byref(eq)))
            if not eq.value:
                return False
        return True

    __hash__ = None

    def spectrum(self, bond_index: int, b: int = 0) -> np.ndarray:
        out, n = np.zeros(64), _i64()
        rc = lib.ttb_get_spectrum(self._hd, b, bond_index, _dp(out), out.size, C.byref(n))
        if rc != 0 and n.value > out.size:      # *n is filled in even when `out` was too small
            out = np.zeros(n.value)
            rc = lib.ttb_get_spectrum(self._hd, b, bond_index, _dp(out), out.size, C.byref(n))
        _ck(rc)
        return out[: n.value].copy()

    def last_timing(self):
        ms, n = C.c_double(), _i64()
        _ck(lib.ttb_last_timing(self._hd, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def profile(self, on: bool = True):
        """Bracket every kernel launch of this handle with CUDA events (bench.py's roofline)."""
        _ck(lib.ttb_profile(self._hd, int(on)))

    def profile_report(self):
        """``{kernel: (launches, total_ms)}`` since :meth:`profile` was switched on."""
        buf = C.create_string_buffer(1 << 16)
        _ck(lib.ttb_profile_report(self._hd, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.split()
            out[name] = (int(n), float(ms))
        return out

    def upload_all(self, packed: np.ndarray, async_: bool = False):
        """Raw packed upload (train-major, site-major, column-major sites at their current dims).
        ``async_=True``: return once the copies are enqueued (``packed`` must stay alive, pinned for a real
        overlap, until the next call that returns results); a following ``compress_`` starts on the sites
        that have already landed."""
        if async_:
            self._keepalive = packed
            _ck(lib.ttb_upload_all_async(self._hd, _dp(packed)))
        else:
            _ck(lib.ttb_upload_all(self._hd, _dp(packed)))

    def trim(self):
        """Release the device scratch ``sample`` / ``twovar_marginals`` / ``dot`` keep between calls."""
        _ck(lib.ttb_trim(self._hd))

    def reset(self):
        """Back to the bond dimensions the train was created with, z = 1 (slab and workspaces are kept)."""
        _ck(lib.ttb_reset(self._hd))

    def download_all(self, out: np.ndarray):
        _ck(lib.ttb_download_all(self._hd, _dp(out)))


    @classmethod
    def empty(cls, bonds: Sequence[int], q: Sequence[int], batch: int = 1):
        """Allocate device storage only (zeros); fill it with :meth:`upload_all`."""
        bond = np.array(list(bonds), dtype=np.int64)
        hd = _h()
        _ck(lib.ttb_create(len(bond) - 1, bond.ctypes.data_as(_pi64), int(np.prod(q)), int(cls.periodic), int(batch),
                           C.byref(hd)))
        return cls(None, batched=batch > 1, _handle=hd, _q=tuple(q))


class PinnedBuffer:
    """Page-locked host staging buffer (cudaHostAlloc) viewed as a float64 NumPy array."""

    def __init__(self, n_doubles: int):
        self._p = C.c_void_p()
        _ck(lib.ttb_host_alloc(C.byref(self._p), int(n_doubles) * 8))
        self.array = np.ctypeslib.as_array(C.cast(self._p, _pd), shape=(int(n_doubles),))

    def free(self):
        if self._p:
            lib.ttb_host_free(self._p)
            self._p = C.c_void_p()


class TensorTrain(AbstractTensorTrain):
    """Open-boundary train (src/tensor_train.jl:8-24): first bond and last bond must be 1."""
    periodic = False


class PeriodicTensorTrain(AbstractTensorTrain):
    """Periodic train (src/periodic_tensor_train.jl:8-27)."""
    periodic = True


# ---- constructors (host-side init + upload; src/tensor_train.jl:42-108,
#      src/periodic_tensor_train.jl:39-59) -----------------------------------------------------------
def _open_bonds(args):
    if len(args) >= 2 and np.isscalar(args[0]) and np.isscalar(args[1]):
        d, L = int(args[0]), int(args[1])
        return [1] + [d] * (L - 1) + [1], args[2:]
    return list(args[0]), args[1:]


def flat_tt(*args):
    bonds, q = _open_bonds(args)
    return TensorTrain([np.ones((bonds[t], bonds[t + 1]) + tuple(q)) for t in range(len(bonds) - 1)])


def rand_tt(rng: np.random.Generator, *args):
    bonds, q = _open_bonds(args)
    return TensorTrain([rng.random((bonds[t], bonds[t + 1]) + tuple(q)) for t in range(len(bonds) - 1)])


def randn_tt(rng: np.random.Generator, *args):
    bonds, q = _open_bonds(args)
    s = 1.0 / math.sqrt(max(bonds))
    return TensorTrain([s * rng.standard_normal((bonds[t], bonds[t + 1]) + tuple(q)) for t in range(len(bonds) - 1)])


def _per_bonds(args):
    if len(args) >= 2 and np.isscalar(args[0]) and np.isscalar(args[1]):
        # reference quirk: (d, L, q...) builds fill(d, L-1), i.e. L-1 sites (periodic_tensor_train.jl:44,59)
        return [int(args[0])] * (int(args[1]) - 1), args[2:]
    return list(args[0]), args[1:]


def flat_periodic_tt(*args):
    bonds, q = _per_bonds(args)
    n = len(bonds)
    x = 1.0 / (float(np.prod(np.array(bonds, dtype=np.float64))) ** (1.0 / n) * float(np.prod(q)))
    return PeriodicTensorTrain([np.full((bonds[t], bonds[(t + 1) % n]) + tuple(q), x) for t in range(n)])


def rand_periodic_tt(rng: np.random.Generator, *args):
    bonds, q = _per_bonds(args)
    n = len(bonds)
    return PeriodicTensorTrain([rng.random((bonds[t], bonds[(t + 1) % n]) + tuple(q)) for t in range(n)])


# ---- metadata ------------------------------------------------------------------------------------
def bond_dims(A: AbstractTensorTrain, b: int = 0) -> List[int]:
    """Left bond of every site (src/abstract_tensor_train.jl:23)."""
    return [int(v) for v in A.bonds(b)[:-1]]


def nparams(A: AbstractTensorTrain, b: int = 0) -> int:
    """src/abstract_tensor_train.jl:31-37"""
    n = _i64()
    _ck(lib.ttb_nparams(A._hd, b, C.byref(n)))
    return n.value


def _flat_x(A, X) -> np.ndarray:
    """Accepts per-site ints or per-site tuples of 0-based physical indices; returns flat indices."""
    out = np.zeros(len(X), dtype=np.int32)
    for t, x in enumerate(X):
        if np.isscalar(x):
            out[t] = int(x)
        else:
            x = tuple(int(v) for v in x)
            if len(x) != len(A.q) or any(not (0 <= xi < qi) for xi, qi in zip(x, A.q)):
                raise ValueError("ArgumentError: X outside the domain")
            out[t] = int(np.ravel_multi_index(x, A.q, order="F"))
    return out


def evaluate(A: AbstractTensorTrain, X, b: int = 0):
    """``tr(prod A[t][:,:,x_t...]) / z`` (src/abstract_tensor_train.jl:80-83).  ``X`` is one point
    (sequence over sites) or an ``(n, L)`` int array of flat indices; returns a float or ``n`` floats."""
    if isinstance(X, np.ndarray) and X.ndim == 2:
        Xf = np.ascontiguousarray(X, dtype=np.int32)
        single = False
    else:
        Xf = _flat_x(A, X)[None]
        single = True
    if Xf.shape[1] != A.L:
        raise ValueError("ArgumentError: X must have one entry per site")
    out = np.zeros(Xf.shape[0])
    _ck(lib.ttb_evaluate(A._hd, b, Xf.ctypes.data_as(C.POINTER(C.c_int32)), Xf.shape[0], _dp(out)))
    return float(out[0]) if single else out


# ---- environments --------------------------------------------------------------------------------
def _accumulate(A, left: bool, normalize: bool):
    fn = lib.ttb_accumulate_L if left else lib.ttb_accumulate_R
    sizes = []
    for b in range(A.batch):
        n = _i64()
        _ck(lib.ttb_env_size(A._hd, b, int(left), C.byref(n)))
        sizes.append(n.value)
    buf = np.zeros(sum(sizes))
    la = np.zeros(A.batch)
    sg = np.zeros(A.batch, dtype=np.int32)
    _ck(fn(A._hd, int(normalize), _dp(buf), _dp(la), sg.ctypes.data_as(_pi)))
    envs, o = [], 0
    for b in range(A.batch):
        bd = A.bonds(b)
        per = []
        for t in range(A.L):
            shp = (int(bd[0]), int(bd[t + 1])) if left else (int(bd[t]), int(bd[A.L]))
            n = shp[0] * shp[1]
            per.append(buf[o:o + n].reshape(shp, order="F").copy())
            o += n
        envs.append(per)
    zs = [Logarithmic(logabs=la[b], sign=int(sg[b])) for b in range(A.batch)]
    return (envs, zs) if A.batched else (envs[0], zs[0])


class Environment:
    """A stored ``accumulate_L`` / ``accumulate_R`` result kept on the device: what the reference's ``l=`` / ``r=``
    keywords of ``marginals`` / ``twovar_marginals`` (src/abstract_tensor_train.jl:208-209, 231-233) and ``rz=`` of
    ``sample`` (:335-336) take.  Only valid for the train it was computed on, as it was then."""

    def __init__(self, A, left: bool, normalize: bool = True):
        self._hd = _h()
        self.left, self.batch, self.batched = bool(left), A.batch, A.batched
        _ck(lib.ttb_env_create(A._hd, int(left), int(normalize), C.byref(self._hd)))

    @property
    def z(self):
        la = np.zeros(self.batch)
        sg = np.zeros(self.batch, dtype=np.int32)
        _ck(lib.ttb_env_z(self._hd, _dp(la), sg.ctypes.data_as(_pi)))
        zs = [Logarithmic(logabs=la[b], sign=int(sg[b])) for b in range(self.batch)]
        return zs if self.batched else zs[0]

    def __del__(self):
        try:
            if self._hd:
                lib.ttb_env_destroy(self._hd)
                self._hd = _h()
        except Exception:
            pass


def _env_arg(e, left: bool, name: str):
    if e is None:
        return None
    if isinstance(e, tuple) and len(e) == 2 and isinstance(e[0], Environment):   # the (env, z) pair of accumulate_*
        e = e[0]
    if not isinstance(e, Environment):
        raise TypeError(f"{name}= takes the device-resident result of accumulate_{'L' if left else 'R'}(A, keep=True)")
    if e.left != left:
        raise ValueError(f"ArgumentError: {name}= needs an accumulate_{'L' if left else 'R'} environment")
    return e._hd


def accumulate_L(A, normalize: bool = True, keep: bool = False):
    """src/abstract_tensor_train.jl:89-102 -> ``(l, z)``.  ``keep=True`` leaves ``l`` on the device
    (an :class:`Environment`, to be passed as ``l=``) instead of copying the matrices to the host."""
    if keep:
        e = Environment(A, True, normalize)
        return e, e.z
    return _accumulate(A, True, normalize)


def accumulate_R(A, normalize: bool = True, keep: bool = False):
    """src/abstract_tensor_train.jl:104-117 -> ``(r, z)``; ``keep=True``: device-resident, for ``r=`` / ``rz=``."""
    if keep:
        e = Environment(A, False, normalize)
        return e, e.z
    return _accumulate(A, False, normalize)


def accumulate_M(A, normalize: bool = True, b: int = 0):
    """src/abstract_tensor_train.jl:119-137: the ``L x L`` table ``M[t][u]`` (0-based) of matrices
    ``d_{t+1} x d_u`` for ``t < u``, ``M[u-1][u] = I``; the other entries are empty ``0 x 0`` arrays like the
    reference's.  (``twovar_marginals`` streams this table on the device and never stores it.)"""
    n = _i64()
    _ck(lib.ttb_accumulate_M_size(A._hd, b, C.byref(n)))
    buf = np.zeros(max(n.value, 1))
    _ck(lib.ttb_accumulate_M(A._hd, b, int(normalize), _dp(buf)))
    bd = A.bonds(b)
    M = [[np.zeros((0, 0)) for _ in range(A.L)] for _ in range(A.L)]
    o = 0
    for t in range(A.L - 1):
        for u in range(t + 1, A.L):
            shp = (int(bd[t + 1]), int(bd[u]))
            M[t][u] = buf[o:o + shp[0] * shp[1]].reshape(shp, order="F").copy()
            o += shp[0] * shp[1]
    return M


def is_in_domain(A, X) -> bool:
    """src/abstract_tensor_train.jl:59-64: every index of every point of ``X`` lies inside the physical dimensions
    (0-based here).  ``X``: one point (list over sites of indices / multi-indices) or an ``(n, L)`` array."""
    Xa = X if isinstance(X, np.ndarray) else None
    if Xa is None:
        try:
            pts = [tuple(np.atleast_1d(x)) for x in X]
        except TypeError:
            return False
        if len(pts) != A.L:
            return False
        for x in pts:
            if len(x) != len(A.q) or any(int(xi) < 0 or int(xi) >= qi for xi, qi in zip(x, A.q)):
                return False
        return True
    Xa = np.ascontiguousarray(Xa, dtype=np.int32).reshape(-1, A.L)
    ok = C.c_int(0)
    _ck(lib.ttb_is_in_domain(A._hd, Xa.ctypes.data_as(C.POINTER(C.c_int32)), Xa.shape[0], C.byref(ok)))
    return bool(ok.value)


def normalization(A):
    """``accumulate_L(A)[2] / A.z`` (src/abstract_tensor_train.jl:154-157)"""
    la = np.zeros(A.batch)
    sg = np.zeros(A.batch, dtype=np.int32)
    _ck(lib.ttb_normalization(A._hd, _dp(la), sg.ctypes.data_as(_pi)))
    zs = [Logarithmic(logabs=la[b], sign=int(sg[b])) for b in range(A.batch)]
    return zs if A.batched else zs[0]


def lognormalization(A):
    """src/abstract_tensor_train.jl:145-147 (DomainError on negative Z)"""
    zs = normalization(A)
    zl = zs if A.batched else [zs]
    if any(z.sign < 0 for z in zl):
        raise ValueError("DomainError: log of a negative normalization")
    out = np.array([z.logabs for z in zl])
    return out if A.batched else float(out[0])


def normalize_(A):
    """``normalize!`` (src/abstract_tensor_train.jl:184-197): returns ``log(|Z|/z_old)``"""
    out = np.zeros(A.batch)
    _ck(lib.ttb_normalize(A._hd, _dp(out)))
    return out if A.batched else float(out[0])


def normalize_eachmatrix_(A):
    """src/abstract_tensor_train.jl:165-176"""
    _ck(lib.ttb_normalize_eachmatrix(A._hd))


def marginals(A, l=None, r=None):
    """src/abstract_tensor_train.jl:208-220: list over sites of arrays shaped ``q``.  ``l`` / ``r``: environments
    from ``accumulate_L(A, keep=True)`` / ``accumulate_R(A, keep=True)`` (default: computed here, like the
    reference's default keywords)."""
    out = np.zeros((A.batch, A.L, A.Q))
    _ck(lib.ttb_marginals_env(A._hd, _env_arg(l, True, "l"), _env_arg(r, False, "r"), _dp(out)))
    res = [[out[b, t].reshape(A.q, order="F") for t in range(A.L)] for b in range(A.batch)]
    return res if A.batched else res[0]


def twovar_marginals(A, maxdist: Optional[int] = None, l=None, r=None, M=None):
    """src/abstract_tensor_train.jl:231-254: dict ``(t,u) -> (q..., q...)`` array, 0-based t<u<=t+maxdist
    (the reference's other entries are empty arrays).  ``l`` / ``r`` as in :func:`marginals`.  ``M`` (the table of
    :func:`accumulate_M`) is accepted for signature parity and not read: the kernels stream ``M[t,u]`` along ``u``."""
    md = A.L - 1 if maxdist is None else int(maxdist)
    n = _i64()
    _ck(lib.ttb_twovar_count(A._hd, md, C.byref(n)))
    out = np.zeros((A.batch, n.value, A.Q * A.Q))
    if n.value:
        _ck(lib.ttb_twovar_marginals_env(A._hd, _env_arg(l, True, "l"), _env_arg(r, False, "r"), md, 0, A.L - 1, _dp(out)))
    res = []
    for b in range(A.batch):
        d, i = {}, 0
        for t in range(A.L - 1):
            for u in range(t + 1, min(A.L - 1, t + md) + 1):
                d[(t, u)] = out[b, i].reshape(A.q + A.q, order="F")
                i += 1
        res.append(d)
    return res if A.batched else res[0]


def _out_ptr(out, n_doubles: int):
    """Result destination of the marginal entry points: a float64 NumPy array (pageable or a PinnedBuffer view)
    or DEVICE memory -- anything with ``data_ptr()`` / ``numel()`` (a torch CUDA tensor on the handle's device);
    the C ABI takes either (unified addressing)."""
    if isinstance(out, np.ndarray):
        if out.dtype != np.float64 or not out.flags.c_contiguous or out.size < n_doubles:
            raise ValueError("ArgumentError: out must be a C-contiguous float64 array with room for the result")
        return _dp(out)
    if hasattr(out, "data_ptr"):
        if out.numel() < n_doubles or out.element_size() != 8 or not out.is_contiguous():
            raise ValueError("ArgumentError: out must be a contiguous float64 tensor with room for the result")
        return C.cast(C.c_void_p(out.data_ptr()), _pd)
    raise TypeError("out: NumPy array or device tensor expected")


def twovar_marginals_block(A, t_lo: int, t_hi: int, maxdist: Optional[int] = None, out=None):
    """The pairs ``(t, u)`` with ``t_lo <= t < t_hi`` (0-based) of :func:`twovar_marginals` as one array
    ``(batch, pairs in the range, Q*Q)``, t-major like the full result: the unit one rank computes when the
    pair marginals of a train are sharded over GPUs (``ttb200.parallel.twovar_marginals_sharded``).
    ``out``: optional destination (NumPy array, pinned or not, or a device tensor: the result then stays on the
    GPU for an NCCL all-gather); returned as given."""
    md = A.L - 1 if maxdist is None else int(maxdist)
    if not (0 <= int(t_lo) <= int(t_hi) <= A.L - 1):
        raise ValueError(f"ArgumentError: pair range [{t_lo}, {t_hi}) outside [0, {A.L - 1})")
    n = sum(min(A.L - 1, t + md) - t for t in range(int(t_lo), int(t_hi)))
    if out is None:
        out = np.zeros((A.batch, n, A.Q * A.Q))
    ptr = _out_ptr(out, A.batch * n * A.Q * A.Q)
    if n:
        _ck(lib.ttb_twovar_marginals_range(A._hd, md, int(t_lo), int(t_hi), ptr))
    return out


# ---- sweeps --------------------------------------------------------------------------------------
def _range(indices):
    if indices is None:
        return 0, 0
    idx = list(indices)
    if len(idx) == 0:
        return 1, 0
    if idx != sorted(idx):
        raise ValueError("ArgumentError: Indices must be sorted")
    if idx != list(range(idx[0], idx[-1] + 1)):
        raise NotImplementedError("only contiguous index ranges are supported")
    return idx[0], idx[-1]


def orthogonalize_right_(A, svd_trunc: Optional[SVDTrunc] = None, indices=None):
    """``orthogonalize_right!`` (src/tensor_train.jl:151-179, periodic: src/periodic_tensor_train.jl:83-111).
    ``indices`` are 1-based site numbers like the reference's."""
    tr = svd_trunc if svd_trunc is not None else (TruncThresh(0.0) if A.periodic else TruncThresh(1e-6))
    lo, hi = _range(indices)
    if lo > hi:
        return A
    ct = tr._c()
    _ck(lib.ttb_orthogonalize_right(A._hd, C.byref(ct), lo, hi))
    tr._after(A)
    return A


def orthogonalize_left_(A, svd_trunc: Optional[SVDTrunc] = None, indices=None):
    """``orthogonalize_left!`` (src/tensor_train.jl:190-218, periodic :113-141)"""
    tr = svd_trunc if svd_trunc is not None else (TruncThresh(0.0) if A.periodic else TruncThresh(1e-6))
    lo, hi = _range(indices)
    if lo > hi:
        return A
    ct = tr._c()
    _ck(lib.ttb_orthogonalize_left(A._hd, C.byref(ct), lo, hi))
    tr._after(A)
    return A


def orthogonalize_center_(A, l: int, svd_trunc: Optional[SVDTrunc] = None):
    """src/tensor_train.jl:220-223 (1-based ``l``)"""
    tr = svd_trunc if svd_trunc is not None else TruncThresh(1e-6)
    ct = tr._c()
    _ck(lib.ttb_orthogonalize_center(A._hd, C.byref(ct), int(l)))
    tr._after(A)
    return A


def orthogonalize_two_site_center_(A, k: int, svd_trunc: Optional[SVDTrunc] = None):
    """src/tensor_train.jl:232-236"""
    tr = svd_trunc if svd_trunc is not None else TruncThresh(1e-6)
    ct = tr._c()
    _ck(lib.ttb_orthogonalize_two_site_center(A._hd, C.byref(ct), int(k)))
    tr._after(A)
    return A


_ORTH = {"none": 0, "left": 1, "right": 2}


def compress_(A, svd_trunc: Optional[SVDTrunc] = None, is_orthogonal: str = "none"):
    """``compress!`` (src/abstract_tensor_train.jl:261-274)"""
    tr = svd_trunc if svd_trunc is not None else TruncThresh(1e-6)
    if is_orthogonal not in _ORTH:
        raise ValueError(f"ArgumentError: Keyword `is_orthogonal` only supports: :none, :left, :right, got :{is_orthogonal}")
    ct = tr._c()
    _ck(lib.ttb_compress(A._hd, C.byref(ct), _ORTH[is_orthogonal]))
    tr._after(A)
    return A


def _resid(A, t, left, b=0):
    r = C.c_double()
    _ck(lib.ttb_canonical_residual(A._hd, b, t, int(left), C.byref(r)))
    return r.value


def is_left_canonical(A, t: int, atol: float = 1e-10, b: int = 0) -> bool:
    """src/tensor_train.jl:263-267 applied to site ``t`` (0-based)"""
    return _resid(A, t, True, b) <= atol


def is_right_canonical(A, t: int, atol: float = 1e-10, b: int = 0) -> bool:
    """src/tensor_train.jl:269-273"""
    return _resid(A, t, False, b) <= atol


def is_canonical(A, central_idx: int, atol: float = 1e-10, b: int = 0) -> bool:
    """src/tensor_train.jl:275-280 (1-based ``central_idx``)"""
    return (all(is_left_canonical(A, t, atol, b) for t in range(0, central_idx - 1)) and
            all(is_right_canonical(A, t, atol, b) for t in range(central_idx, A.L)))


# ---- sampling ------------------------------------------------------------------------------------
def sample(A, nsamples: Optional[int] = None, seed: int = 0, first_index: int = 0, uniforms=None, b: int = 0, rz=None):
    """``sample`` (src/abstract_tensor_train.jl:335-384).  Without ``nsamples`` returns one draw
    ``(x, p)`` like the reference (x = list over sites of 0-based multi-indices); with ``nsamples``
    returns ``(X, p)`` with ``X`` an ``(nsamples, L)`` uint8 array of flat 0-based indices.
    ``uniforms`` (``(nsamples, L)`` doubles) replaces the Philox stream (one ``rand`` per site).
    ``rz``: ``accumulate_R(A, keep=True)`` computed once for many calls (the reference's ``rz`` keyword, :336)."""
    n = 1 if nsamples is None else int(nsamples)
    up = None
    if uniforms is not None:
        u = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64).reshape(n, A.L))
        up = _dp(u)
    X = np.zeros((n, A.L), dtype=np.uint8)
    p = np.zeros(n)
    _ck(lib.ttb_sample_env(A._hd, _env_arg(rz, False, "rz"), b, n, C.c_uint64(seed), C.c_uint64(first_index), up,
                           X.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(p)))
    if nsamples is None:
        x = [list(np.unravel_index(int(xi), A.q, order="F")) for xi in X[0]]
        return x, float(p[0])
    return X, p


def sample_stats(A, nsamples: int, seed: int = 0, first_index: int = 0, b: int = 0, rz=None):
    """Throughput variant: draws stay on the device; returns the ``(L, Q)`` histogram of the drawn
    values and the sum of log-probabilities."""
    hist = np.zeros((A.L, A.Q), dtype=np.int64)
    slp = C.c_double(0.0)
    _ck(lib.ttb_sample_stats_env(A._hd, _env_arg(rz, False, "rz"), b, int(nsamples), C.c_uint64(seed), C.c_uint64(first_index),
                                 hist.ctypes.data_as(_pi64), C.byref(slp)))
    return hist, slp.value


# ---- two-site pieces: _merge_tensors / _split_tensor / data environments ------------------
def merge_tensors(A, t: int, b: int = 0) -> np.ndarray:
    """``_merge_tensors(A[t], A[t+1])`` (src/utils.jl:37-44), ``t`` 0-based: array ``(d_t, d_{t+2}) + q + q``"""
    bd = A.bonds(b)
    if not (0 <= int(t) < A.L - 1):
        raise ValueError(f"ArgumentError: site pair ({t}, {t + 1}) outside the train")
    out = np.zeros(int(bd[t]) * int(bd[t + 2]) * A.Q * A.Q)
    _ck(lib.ttb_merge_tensors(A._hd, b, int(t), _dp(out)))
    return out.reshape((int(bd[t]), int(bd[t + 2])) + tuple(A.q) + tuple(A.q), order="F")


def split_tensor_(A, t: int, Cm: np.ndarray, svd_trunc: Optional[SVDTrunc] = None, lr: str = "left", b: int = 0):
    """``A[t], A[t+1] = _split_tensor(C; svd_trunc, lr)`` (src/abstract_tensor_train.jl:283-308) in place on the device
    train; ``lr`` = "left" (``A[t] = U``) or "right" (``A[t+1] = V'``).  Returns the relative truncation error."""
    if lr not in ("left", "right"):
        raise ValueError("ArgumentError: lr must be \"left\" or \"right\"")
    tr = (svd_trunc or TruncThresh(0.0))
    bd = A.bonds(b)
    Cf = np.ascontiguousarray(np.reshape(np.asarray(Cm, dtype=np.float64), -1, order="F"))
    if Cf.size != int(bd[t]) * int(bd[t + 2]) * A.Q * A.Q:
        raise ValueError("DimensionMismatch: C does not have the shape (d_t, d_{t+2}, q..., q...)")
    err = C.c_double(0.0)
    c = tr._c()
    _ck(lib.ttb_split_tensor(A._hd, b, int(t), _dp(Cf), C.byref(c), 1 if lr == "left" else 0, C.byref(err)))
    A._refresh()
    if isinstance(tr, TruncBondMax):
        tr.maxerr[0] = max(tr.maxerr[0], err.value)
    return err.value


def _x_flat_rows(A, X) -> np.ndarray:
    return np.ascontiguousarray(np.stack([_flat_x(A, x) for x in X]).astype(np.int32))


def data_environments(A, X, left: bool = True, b: int = 0) -> np.ndarray:
    """``[precompute_left_environments(A, x) for x in X]`` (``left``) / ``precompute_right_environments``
    (src/tensor_train.jl:320-334) as one array ``(len(X), L + 1, dmax)``: entry ``j`` holds ``d_j`` values (the
    product of the first ``j`` matrices, resp. of matrices ``j..L-1``), zero padded."""
    Xf = _x_flat_rows(A, X)
    env = np.zeros((Xf.shape[0], A.L + 1, A._dmax()))
    _ck(lib.ttb_data_environments(A._hd, b, Xf.ctypes.data_as(C.POINTER(C.c_int32)), Xf.shape[0], int(bool(left)), _dp(env)))
    return env


def update_environments_(env: np.ndarray, A, t: int, X, left: bool = True, b: int = 0) -> np.ndarray:
    """``update_environments!`` (src/dmrg.jl:124-151) after site ``t`` (0-based) changed: entry ``t+1`` (``left``) or
    entry ``t`` of every datapoint's environment list is recomputed in place."""
    Xf = _x_flat_rows(A, X)
    if env.shape != (Xf.shape[0], A.L + 1, A._dmax()) or not env.flags.c_contiguous:
        raise ValueError("ArgumentError: env must be the array data_environments returned")
    _ck(lib.ttb_data_env_update(A._hd, b, Xf.ctypes.data_as(C.POINTER(C.c_int32)), Xf.shape[0], int(t), int(bool(left)), _dp(env)))
    return env


# ---- dot / norm / norm2m / isapprox (first row after the hot path) -------------------------
def dot_log(A, B, ba: int = 0, bb: int = 0) -> Logarithmic:
    """``dot(A, B)`` as a ``Logarithmic`` (sign, log|.|): long chains do not overflow."""
    la = C.c_double(0.0)
    sg = C.c_int(1)
    _ck(lib.ttb_dot(A._hd, ba, B._hd, bb, C.byref(la), C.byref(sg)))
    return Logarithmic(logabs=float(la.value), sign=int(sg.value))


def dot(A, B, ba: int = 0, bb: int = 0) -> float:
    """``LinearAlgebra.dot(A, B)`` (src/abstract_tensor_train.jl:395-408), z factors included."""
    return float(dot_log(A, B, ba, bb))


def norm(A, b: int = 0) -> float:
    """``LinearAlgebra.norm(A)`` = sqrt(A . A) (src/abstract_tensor_train.jl:418-420)"""
    return math.sqrt(abs(dot(A, A, b, b)))


def norm2m(A, B, ba: int = 0, bb: int = 0) -> float:
    """``norm2m(A, B)`` = |A|^2 + |B|^2 - 2 A.B (src/abstract_tensor_train.jl:430-432)"""
    return norm(A, ba) ** 2 + norm(B, bb) ** 2 - 2.0 * dot(A, B, ba, bb)


def isapprox(A, B, atol: float = 0.0, rtol: float = 1e-6, ba: int = 0, bb: int = 0) -> bool:
    """``Base.isapprox`` for trains (src/abstract_tensor_train.jl:434-437)"""
    na, nb = norm(A, ba), norm(B, bb)
    return na * na + nb * nb - 2.0 * dot(A, B, ba, bb) <= max(atol, rtol * max(na, nb)) ** 2


__all__ += ["dot", "dot_log", "norm", "norm2m", "isapprox"]


# ---- MPS Born-machine wrapper (src/MatrixProductStates/mps.jl) ------------------
class MPS:
    """``p(x) = |psi(x)|^2`` over a device train ``psi`` (src/MatrixProductStates/mps.jl:19-25).

    Built so far on the device: ``evaluate`` (:54), ``normalization`` (:139-146, the ``d^4`` environment
    chain of ``accumulate_L`` is the ``dot(psi, psi)`` chain of ``dot.cu``), ``normalize_`` (:154-168) and the
    sweeps, which the reference forwards to the wrapped train (:234-252).  ``marginals`` /
    ``twovar_marginals`` / ``sample`` of an MPS need the stored ``d^4`` environments and are not built yet:
    they raise ``NotImplementedError`` instead of falling back to the CPU."""

    def __init__(self, psi: AbstractTensorTrain):
        if getattr(psi, "batched", False):
            raise NotImplementedError("MPS over a batched handle")
        self.psi = psi

    def __len__(self):
        return self.psi.L

    def copy(self):
        return MPS(self.psi.copy())


def mps_evaluate(p: MPS, X) -> float:
    """``abs2(evaluate(p.psi, X...))`` (mps.jl:54)"""
    v = evaluate(p.psi, X)
    return v * v


def mps_normalization(p: MPS) -> Logarithmic:
    """``accumulate_L(p)[2] / abs2(psi.z)`` (mps.jl:139-146) = ``dot(psi, psi)``"""
    return dot_log(p.psi, p.psi)


def mps_normalize_(p: MPS) -> float:
    """``normalize!(p::MPS)`` (mps.jl:154-168): tensors divided by ``|Z|^(1/2L)``, ``psi.z = 1``; returns
    ``log(|Z| / abs2(psi.z))``"""
    zn = dot_log(p.psi, p.psi)                     # = Z / z^2
    zpsi = p.psi.z
    raw = zn.logabs + 2.0 * zpsi.logabs            # log|Z| of the environment chain itself
    x = math.exp(0.5 * raw / p.psi.L)
    if x != 0.0 and math.isfinite(x):
        _ck(lib.ttb_scale_sites(p.psi._hd, _dp(np.array([x], dtype=np.float64))))
    _ck(lib.ttb_set_z(p.psi._hd, 0, 0.0, 1))
    return zn.logabs


def mps_compress_(p: MPS, svd_trunc=None, is_orthogonal: str = "none"):
    """forwarded to the wrapped train (mps.jl:250-252)"""
    compress_(p.psi, svd_trunc, is_orthogonal=is_orthogonal)
    return p


def mps_orthogonalize_left_(p: MPS, svd_trunc=None):
    """mps.jl:238-240"""
    orthogonalize_left_(p.psi, svd_trunc)
    return p


def mps_orthogonalize_right_(p: MPS, svd_trunc=None):
    """mps.jl:234-236"""
    orthogonalize_right_(p.psi, svd_trunc)
    return p


def _mps_accumulate(p: MPS, left: bool, normalize: bool):
    A = p.psi
    n = _i64()
    _ck(lib.ttb_mps_env_size(A._hd, 0, int(left), C.byref(n)))
    buf = np.zeros(max(n.value, 1))
    la, sg = C.c_double(), C.c_int()
    _ck(lib.ttb_mps_accumulate(A._hd, 0, int(left), int(bool(normalize)), _dp(buf), C.byref(la), C.byref(sg)))
    bd = [int(v) for v in A.bonds(0)]
    d1 = bd[0]
    envs, o = [], 0
    for t in range(A.L):
        d = bd[t + 1] if left else bd[t]
        n = d1 * d1 * d * d
        # the device keeps one d x d slice per boundary pair (a1, b1): [a, b, a1, b1] column-major; the reference's
        # arrays are Ll[a1, a, b1, b] (mps.jl:60-80) and Rl[a, aL, b, bL] (:82-103); an open train has a1 = b1 = 1
        blk = buf[o:o + n].reshape((d, d, d1, d1), order="F")
        envs.append(np.ascontiguousarray(blk.transpose(2, 0, 3, 1) if left else blk.transpose(0, 2, 1, 3)))
        o += n
    return envs, Logarithmic(logabs=la.value, sign=sg.value)


def mps_accumulate_L(p: MPS, normalize: bool = True):
    """``accumulate_L(p::MPS)`` (mps.jl:60-80): ``(l, z)``; ``l[t]`` indexed ``[a1, a_{t+1}, b1, b_{t+1}]`` (open psi:
    ``a1 = b1 = 1``; periodic psi: the full ``d_1 x d x d_1 x d`` array), hermiticity restored after every step, every
    ``l[t < L]`` max-abs 1 with ``normalize``."""
    return _mps_accumulate(p, True, normalize)


def mps_accumulate_R(p: MPS, normalize: bool = True):
    """``accumulate_R(p::MPS)`` (mps.jl:82-103): ``(r, z)``; ``r[t]`` indexed ``[a_t, aL, b_t, bL]`` -- for an open psi
    ``aL = bL = 1``, shape ``(d, 1, d, 1)``."""
    return _mps_accumulate(p, False, normalize)


def _mps_pack_env(envs, left: bool):
    """list of the reference's four-index arrays (what mps_accumulate_L / _R returned) -> the device layout"""
    if envs is None:
        return None
    parts = []
    for e in envs:
        e = np.asarray(e, dtype=np.float64)
        if e.ndim != 4:
            raise ValueError("ArgumentError: an MPS environment is a list of four-index arrays")
        blk = e.transpose(1, 3, 0, 2) if left else e.transpose(0, 2, 1, 3)      # -> [a, b, a1, b1]
        parts.append(np.reshape(blk, -1, order="F"))
    return np.ascontiguousarray(np.concatenate(parts))


def _mps_check_env(A, buf, left: bool, name: str):
    if buf is None:
        return None
    n = _i64()
    _ck(lib.ttb_mps_env_size(A._hd, 0, int(left), C.byref(n)))
    if buf.size != n.value:
        raise ValueError(f"ArgumentError: {name}= does not fit this train (expected {n.value} values, got {buf.size})")
    return _dp(buf)


def mps_marginals(p: MPS, l=None, r=None):
    """``marginals(p::MPS; l, r)`` (mps.jl:176-196): list over sites of arrays shaped ``q``; ``l`` / ``r``: the lists
    ``mps_accumulate_L(p)[0]`` / ``mps_accumulate_R(p)[0]`` computed once for many calls"""
    A = p.psi
    lb, rb = _mps_pack_env(l, True), _mps_pack_env(r, False)
    out = np.zeros((A.L, A.Q))
    _ck(lib.ttb_mps_marginals_env(A._hd, 0, _mps_check_env(A, lb, True, "l"), _mps_check_env(A, rb, False, "r"), _dp(out)))
    return [out[t].reshape(A.q, order="F") for t in range(A.L)]


def mps_twovar_marginals(p: MPS, maxdist: Optional[int] = None, l=None, r=None, M=None):
    """``twovar_marginals(p::MPS; l, r, M, maxdist)`` (mps.jl:206-232): dict ``(t, u) -> (q..., q...)`` array, 0-based
    ``t < u <= t + maxdist``.  ``M`` is accepted for signature parity and not read (``accumulate_M`` is streamed)."""
    A = p.psi
    md = A.L - 1 if maxdist is None else int(maxdist)
    n = _i64()
    _ck(lib.ttb_twovar_count(A._hd, md, C.byref(n)))
    out = np.zeros((max(n.value, 1), A.Q * A.Q))
    lb, rb = _mps_pack_env(l, True), _mps_pack_env(r, False)
    _ck(lib.ttb_mps_twovar_marginals_env(A._hd, 0, _mps_check_env(A, lb, True, "l"), _mps_check_env(A, rb, False, "r"), md, _dp(out)))
    d, i = {}, 0
    for t in range(A.L - 1):
        for u in range(t + 1, min(A.L - 1, t + md) + 1):
            d[(t, u)] = out[i].reshape(A.q + A.q, order="F")
            i += 1
    return d


def mps_sample(p: MPS, nsamples: Optional[int] = None, seed: int = 0, first_index: int = 0, uniforms=None, rz=None):
    """``sample(p::MPS; rz)`` (mps.jl:264-290, 300-313): ``(X, q)`` like :func:`sample`; ``q = |tr Q|^2 / z``;
    ``rz = mps_accumulate_R(p)`` computed once for many calls"""
    A = p.psi
    single = nsamples is None and uniforms is None
    if uniforms is not None:
        uniforms = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64).reshape(-1, A.L))
        n = uniforms.shape[0]
    else:
        n = 1 if nsamples is None else int(nsamples)
    X = np.zeros((max(n, 1), A.L), dtype=np.uint8)
    q = np.zeros(max(n, 1))
    rb, zl = None, 0.0
    if rz is not None:
        renv, z = rz
        rb = _mps_pack_env(renv, False)
        zl = z.logabs if isinstance(z, Logarithmic) else math.log(abs(float(z)))
    _ck(lib.ttb_mps_sample_env(A._hd, 0, _mps_check_env(A, rb, False, "rz"), zl, n, seed, first_index,
                               _dp(uniforms) if uniforms is not None else None, X.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(q)))
    if single:
        return X[0].astype(np.int64), float(q[0])
    return X[:n].astype(np.int64), q[:n]


__all__ += ["MPS", "mps_evaluate", "mps_normalization", "mps_normalize_", "mps_compress_", "mps_orthogonalize_left_",
            "mps_orthogonalize_right_", "mps_accumulate_L", "mps_accumulate_R", "mps_marginals", "mps_twovar_marginals",
            "mps_sample"]
