"""GPU parity tests: the CUDA path (through the C ABI / ctypes mirror) against the CPU oracle on the
same seeded inputs.  Tolerances are the ones BASELINE.json's north_star states:
  evaluate, log Z, marginals: rel 1e-10;  singular values |Δσ_k| <= 1e-8 σ_1;  post-truncation evaluate
  rel 1e-8;  retained bond dimensions identical;  sampling: identical draws for injected uniforms,
  chi-square for the Philox stream."""
import math

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

from oracle import tt_oracle as O

T = None


@pytest.fixture(scope="module", autouse=True)
def _load():
    global T
    import ttb200
    T = ttb200
    T.set_device(0)
    yield


RNG = lambda s: np.random.Generator(np.random.Philox(s))
REL = 1e-10


def close(a, b, rel=REL, abs_=0.0):
    return abs(a - b) <= rel * max(abs(a), abs(b)) + abs_


def pair(tensors, periodic=False, z=1.0):
    cls_g = T.PeriodicTensorTrain if periodic else T.TensorTrain
    cls_o = O.PeriodicTensorTrain if periodic else O.TensorTrain
    return cls_g(tensors, z=z), cls_o([a.copy() for a in tensors], z=z)


def rand_open(rng, bonds, q, fill="rand"):
    f = (lambda s: rng.random(s)) if fill == "rand" else (lambda s: rng.standard_normal(s) / math.sqrt(max(bonds)))
    return [f((bonds[t], bonds[t + 1]) + tuple(q)) for t in range(len(bonds) - 1)]


def rand_per(rng, bonds, q):
    n = len(bonds)
    return [rng.random((bonds[t], bonds[(t + 1) % n]) + tuple(q)) for t in range(n)]


def rand_points(rng, Ao, n):
    Qs = [int(np.prod(a.shape[2:])) for a in Ao]
    return np.stack([rng.integers(0, Q, n) for Q in Qs], axis=1).astype(np.int32)


def o_eval_flat(Ao, xf):
    return O.evaluate(Ao, [np.unravel_index(int(x), a.shape[2:], order="F") for x, a in zip(xf, Ao)])


def check_evals(A, Ao, X, rel, signed=False):
    """signed=True (randn fills): values cancel, so the error is measured against the largest |value|
    of the point set instead of each value."""
    g = T.evaluate(A, X)
    e = np.array([o_eval_flat(Ao, X[i]) for i in range(X.shape[0])])
    if signed:
        assert np.max(np.abs(g - e)) <= rel * np.max(np.abs(e)), (g, e)
    else:
        for i in range(X.shape[0]):
            assert close(g[i], e[i], rel, 1e-300), (i, g[i], e[i])


def spectra_close(lam_g, lam_o):
    """|Δσ_k| <= 1e-8 σ_1 on the scale-normalised spectrum.  The absolute scale of the matrix a sweep
    step factorises is NOT gauge invariant: the reference divides the previous step's D by max|D|
    (tensor_train.jl:169-173), and max|.| changes under the orthogonal gauge freedom of the eps=0
    sweep (QR here, U*lambda there).  Ratios sigma_k/sigma_1 -- all the rank rules use -- are."""
    return len(lam_g) == len(lam_o) and np.max(np.abs(lam_g / lam_g[0] - lam_o / lam_o[0])) <= 1e-8


# ---- container -------------------------------------------------------------------------------------
def test_container_roundtrip_and_metadata():
    rng = RNG(100)
    ts = rand_open(rng, [1, 3, 4, 10, 1], (2, 2))
    A, Ao = pair(ts, z=12.3)
    assert T.bond_dims(A) == [1, 3, 4, 10] == O.bond_dims(Ao)
    assert T.nparams(A) == O.nparams(Ao)
    for t in range(4):
        assert np.array_equal(A.site(t), ts[t])
    assert close(float(A.z), 12.3)
    B = A.copy()
    B.set_site(1, 2 * ts[1])
    assert np.array_equal(A.site(1), ts[1]) and np.array_equal(B.site(1), 2 * ts[1])
    with pytest.raises(ValueError):
        T.TensorTrain([rng.random((2, 3, 2)), rng.random((3, 1, 2))])          # first bond != 1
    with pytest.raises(ValueError):
        T.TensorTrain([rng.random((1, 4, 2)), rng.random((3, 1, 2))])          # incompatible bonds
    with pytest.raises(ValueError):
        T.PeriodicTensorTrain([rng.random((2, 3, 2)), rng.random((3, 4, 2))])  # ring does not close


# ---- evaluate / environments / marginals -------------------------------------------------------------
@pytest.mark.parametrize("q", [(2,), (2, 3), (3, 1, 2)])
def test_evaluate_and_environments_open(q):
    rng = RNG(101 + len(q))
    ts = rand_open(rng, [1, 5, 7, 3, 6, 1], q)
    A, Ao = pair(ts, z=-0.37)
    check_evals(A, Ao, rand_points(rng, Ao, 16), REL)
    for norm in (True, False):
        (l, zl), (lo, zlo) = T.accumulate_L(A, norm), O.accumulate_L(Ao, norm)
        (r, zr), (ro, zro) = T.accumulate_R(A, norm), O.accumulate_R(Ao, norm)
        for a, b in zip(l + r, lo + ro):
            assert np.allclose(a, b, rtol=1e-10, atol=0)
        assert close(zl.logabs, zlo.logabs, REL, 1e-12) and zl.sign == zlo.sign
        assert close(zr.logabs, zro.logabs, REL, 1e-12) and zr.sign == zro.sign
    zg, zo = T.normalization(A), O.normalization(Ao)
    assert zg.sign == zo.sign == -1 and close(zg.logabs, zo.logabs, REL, 1e-12)
    with pytest.raises(ValueError):
        T.lognormalization(A)
    for a, b in zip(T.marginals(A), O.marginals(Ao)):
        assert np.allclose(a, b, rtol=1e-10, atol=0)
    m2, m2o = T.twovar_marginals(A), O.twovar_marginals(Ao)
    assert set(m2) == set(m2o)
    for k in m2o:
        assert np.allclose(m2[k], m2o[k], rtol=1e-10, atol=0), k
    m2 = T.twovar_marginals(A, maxdist=2)
    m2o = O.twovar_marginals(Ao, maxdist=2)
    assert set(m2) == set(m2o)
    for k in m2o:
        assert np.allclose(m2[k], m2o[k], rtol=1e-10, atol=0), k


def test_evaluate_and_environments_periodic():
    rng = RNG(110)
    ts = rand_per(rng, [3, 5, 2, 4], (2, 2))
    A, Ao = pair(ts, periodic=True)
    check_evals(A, Ao, rand_points(rng, Ao, 16), REL)
    (l, zl), (lo, zlo) = T.accumulate_L(A), O.accumulate_L(Ao)
    (r, zr), (ro, zro) = T.accumulate_R(A), O.accumulate_R(Ao)
    for a, b in zip(l + r, lo + ro):
        assert np.allclose(a, b, rtol=1e-10, atol=0)
    assert close(zl.logabs, zlo.logabs, REL) and close(zr.logabs, zro.logabs, REL)
    assert close(float(T.normalization(A)), O.exact_normalization(Ao), 1e-10)
    for a, b in zip(T.marginals(A), O.marginals(Ao)):
        assert np.allclose(a, b, rtol=1e-10, atol=0)
    m2, m2o = T.twovar_marginals(A), O.twovar_marginals(Ao)
    for k in m2o:
        assert np.allclose(m2[k], m2o[k], rtol=1e-10, atol=0), k
    A2 = T.flat_periodic_tt([3, 5, 2, 4], 2, 3)
    assert close(float(T.normalization(A2)), 1.0, 1e-12)


def test_flat_known_normalization_and_normalize():
    q, bonds = (2, 4), [1, 3, 5, 2, 1]
    A = T.flat_tt(bonds, *q)
    assert close(float(T.normalization(A)), np.prod(bonds) * np.prod(q) ** 4, 1e-12)
    rng = RNG(111)
    ts = rand_open(rng, [1, 6, 9, 4, 1], (3,))
    A, Ao = pair(ts, z=5.0)
    lg, lo = T.normalize_(A), O.normalize(Ao)
    assert close(lg, lo, REL)
    assert close(float(T.normalization(A)), 1.0, 1e-10) and close(float(A.z), 1.0)
    for t in range(4):
        assert np.allclose(A.site(t), Ao[t], rtol=1e-12)
    A, Ao = pair(ts, z=-3.0)
    T.normalize_eachmatrix_(A)
    O.normalize_eachmatrix(Ao)
    for t in range(4):
        assert np.allclose(A.site(t), Ao[t], rtol=1e-14)
    assert close(A.z.logabs, Ao.z.logabs, 1e-12) and A.z.sign == Ao.z.sign == -1


def test_long_chain_stays_finite():
    """test/tensor_train.jl:232-258 (L = 10 000, bonds 10..15)."""
    rng = RNG(112)
    L = 10000
    bonds = [1] + list(rng.integers(10, 16, L - 1)) + [1]
    ts = rand_open(rng, bonds, (2, 2))
    A, Ao = pair(ts)
    assert close(T.lognormalization(A), O.lognormalization(Ao), REL)
    T.normalize_(A)
    assert close(float(T.normalization(A)), 1.0, 1e-8)
    A, Ao = pair([a * 1e-50 for a in ts])
    assert close(T.lognormalization(A), O.lognormalization(Ao), REL)


# ---- sweeps ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fill", ["rand", "randn"])
def test_orthogonalize_eps0_gauge_invariance(fill):
    rng = RNG(120)
    ts = rand_open(rng, [1, 3, 4, 10, 7, 1], (2, 2), fill)
    A, Ao = pair(ts, z=12.3)
    X = rand_points(rng, Ao, 12)
    z0 = O.normalization(Ao)
    sg = fill == "randn"
    T.orthogonalize_right_(A, T.TruncThresh(0.0))
    check_evals(A, Ao, X, 1e-10, sg)
    assert all(T.is_right_canonical(A, t) for t in range(1, 5))
    Bo = Ao.copy(); O.orthogonalize_right(Bo, O.TruncThresh(0.0))
    assert T.bond_dims(A) == O.bond_dims(Bo)
    zg = T.normalization(A)
    assert zg.sign == z0.sign and close(zg.logabs, z0.logabs, REL, 1e-12)
    T.orthogonalize_left_(A, T.TruncThresh(0.0))
    O.orthogonalize_left(Bo, O.TruncThresh(0.0))
    check_evals(A, Ao, X, 1e-10, sg)
    assert all(T.is_left_canonical(A, t) for t in range(0, 4))
    assert T.bond_dims(A) == O.bond_dims(Bo)


@pytest.mark.parametrize("trunc", ["thresh1e-3", "thresh1e-8", "bond3", "bondmax2", "bondthresh", "thresh1e-15"])
@pytest.mark.parametrize("fill", ["rand", "randn"])
def test_compress_matches_oracle(trunc, fill):
    rng = RNG(130)
    ts = rand_open(rng, [1, 4, 9, 12, 8, 5, 1], (3, 2), fill)
    A, Ao = pair(ts)
    mk = {"thresh1e-3": (T.TruncThresh(1e-3), O.TruncThresh(1e-3)), "thresh1e-8": (T.TruncThresh(1e-8), O.TruncThresh(1e-8)),
          "bond3": (T.TruncBond(3), O.TruncBond(3)), "bondmax2": (T.TruncBondMax(2), O.TruncBondMax(2)),
          "bondthresh": (T.TruncBondThresh(4, 1e-4), O.TruncBondThresh(4, 1e-4)),
          "thresh1e-15": (T.TruncThresh(1e-15), O.TruncThresh(1e-15))}[trunc]
    X = rand_points(rng, Ao, 12)
    A.record_spectrum(True)
    stats = []
    T.compress_(A, mk[0])
    O.compress(Ao, mk[1], stats=stats)
    assert T.bond_dims(A) == O.bond_dims(Ao), (T.bond_dims(A), O.bond_dims(Ao))
    for (t, lam_o, kept) in stats:      # left sweep at 1-based site t decides bond index t
        lam_g = A.spectrum(t)
        assert spectra_close(lam_g, lam_o), (t, lam_g, lam_o)
    check_evals(A, Ao, X, 1e-8, fill == "randn")
    if trunc != "thresh1e-15":
        assert all(T.is_left_canonical(A, t) for t in range(0, 5))
    zg, zo = T.normalization(A), O.normalization(Ao)
    assert zg.sign == zo.sign and close(zg.logabs, zo.logabs, 1e-8, 1e-10)
    if trunc == "bondmax2":
        assert close(mk[0].maxerr[0], mk[1].maxerr[0], 1e-6, 1e-14)


def test_compress_routes_and_errors():
    """test/tensor_train.jl:217-230."""
    rng = RNG(131)
    ts = rand_open(rng, [1, 4, 4, 4, 4, 1], (3, 2), "randn")
    tr = T.TruncThresh(1e-2)
    A, Ao = pair(ts)
    X = rand_points(rng, Ao, 8)
    B = A.copy(); Cc = A.copy()
    Bo = Ao.copy(); Co = Ao.copy()
    T.compress_(A, tr); O.compress(Ao, O.TruncThresh(1e-2))
    T.orthogonalize_right_(B, T.TruncThresh(0.0)); T.compress_(B, tr, is_orthogonal="right")
    O.orthogonalize_right(Bo, O.TruncThresh(0.0)); O.compress(Bo, O.TruncThresh(1e-2), is_orthogonal="right")
    T.orthogonalize_left_(Cc, T.TruncThresh(0.0)); T.compress_(Cc, tr, is_orthogonal="left")
    O.orthogonalize_left(Co, O.TruncThresh(0.0)); O.compress(Co, O.TruncThresh(1e-2), is_orthogonal="left")
    assert T.bond_dims(A) == O.bond_dims(Ao) and T.bond_dims(B) == O.bond_dims(Bo) and T.bond_dims(Cc) == O.bond_dims(Co)
    check_evals(A, Ao, X, 1e-8, True); check_evals(B, Bo, X, 1e-8, True); check_evals(Cc, Co, X, 1e-8, True)
    with pytest.raises(ValueError):
        T.compress_(A, tr, is_orthogonal="something")
    with pytest.raises(ValueError):
        T.orthogonalize_right_(A, tr, indices=[1, 2])
    with pytest.raises(ValueError):
        T.orthogonalize_left_(A, tr, indices=[3, 2])


def test_orthogonalize_center():
    """test/tensor_train.jl:198-215."""
    rng = RNG(132)
    ts = rand_open(rng, [1, 3, 3, 3, 1], (2, 2), "randn")
    A, Ao = pair(ts)
    T.orthogonalize_center_(A, 2, T.TruncBond(3))
    O.orthogonalize_center(Ao, 2, O.TruncBond(3))
    assert T.is_canonical(A, 2)
    assert T.bond_dims(A) == O.bond_dims(Ao)
    zg, zo = T.normalization(A), O.normalization(Ao)
    assert zg.sign == zo.sign and close(zg.logabs, zo.logabs, 1e-8, 1e-10)
    T.orthogonalize_two_site_center_(A, 2, T.TruncThresh(0.0))
    assert T.is_right_canonical(A, 3)


def test_periodic_sweeps_match_oracle():
    """src/periodic_tensor_train.jl:83-141 incl. the wrap-around step order."""
    rng = RNG(140)
    ts = rand_per(rng, [4, 3, 5, 2, 6], (2, 3))
    A, Ao = pair(ts, periodic=True)
    X = rand_points(rng, Ao, 12)
    T.orthogonalize_right_(A); Bo = Ao.copy(); O.orthogonalize_right(Bo)
    assert T.bond_dims(A) == O.bond_dims(Bo)
    check_evals(A, Ao, X, 1e-9)
    T.orthogonalize_left_(A); O.orthogonalize_left(Bo)
    assert T.bond_dims(A) == O.bond_dims(Bo)
    check_evals(A, Ao, X, 1e-9)
    zg, zo = T.normalization(A), O.normalization(Ao)
    assert close(zg.logabs, zo.logabs, 1e-9, 1e-10)
    A, Ao = pair(ts, periodic=True)
    T.compress_(A, T.TruncThresh(1e-6)); O.compress(Ao, O.TruncThresh(1e-6))
    assert T.bond_dims(A) == O.bond_dims(Ao)
    check_evals(A, Ao, X, 1e-8)
    F = T.flat_periodic_tt([4, 4, 4, 4], 2); Fo = O.flat_periodic_tt([4, 4, 4, 4], 2)
    T.compress_(F, T.TruncThresh(3e-2)); O.compress(Fo, O.TruncThresh(3e-2))
    assert T.bond_dims(F) == O.bond_dims(Fo) and max(T.bond_dims(F)) < 4
    assert close(float(T.normalization(F)), 1.0, 1e-8)


def test_truncation_error_bound():
    """test/svd_trunc.jl:1-24: norm2m(A,B) < (L eps)^2 (dot/norm evaluated by the oracle on the
    downloaded tensors -- dot is a `next` row)."""
    rng = RNG(141)
    L = 10
    ts = rand_open(rng, [1] + list(rng.integers(1, 4, L - 1)) + [1], (4, 4))
    Bo = O.TensorTrain([a.copy() for a in ts]); O.normalize(Bo)
    for eps in 10.0 ** np.arange(0.0, -5.5, -1.0):
        A = T.TensorTrain([a.copy() for a in Bo.tensors])
        T.compress_(A, T.TruncThresh(eps))
        Ad = O.TensorTrain(A.tensors(), z=O.LogNum(logabs=A.z.logabs, sign=A.z.sign))
        assert O.norm2m(Ad, Bo) < (L * eps) ** 2 + 1e-13


@pytest.mark.parametrize("d,q,L,B", [(32, (2,), 10, 3), (16, (4,), 7, 2), (21, (3,), 6, 2), (32, (2,), 5, 1),
                                     (6, (3, 4), 4, 2)])   # Q = 12: no cp.async staging buffer, scalar pair kernel
def test_twovar_tensor_core_path_config5_shapes(d, q, L, B):
    """twovar_mma.cuh (bonds <= 32, Q d_1 <= 64): full bonds, then the ragged bonds compress! leaves."""
    rng = RNG(160 + d)
    tsb = [rng.random((B, d, d) + q) for _ in range(L)]
    A = T.PeriodicTensorTrain(tsb, batched=True) if B > 1 else T.PeriodicTensorTrain([a[0] for a in tsb])
    for stage in range(2):
        m2 = T.twovar_marginals(A)
        m2 = m2 if B > 1 else [m2]
        m3 = T.twovar_marginals(A, maxdist=3)
        m3 = m3 if B > 1 else [m3]
        envs = {(left, nrm): (T.accumulate_L if left else T.accumulate_R)(A, nrm) for left in (True, False) for nrm in (True, False)}
        m1 = T.marginals(A)
        m1 = m1 if B > 1 else [m1]
        zs = T.normalization(A)
        zs = zs if B > 1 else [zs]
        for b in range(B):
            zb = A.z[b] if B > 1 else A.z
            Ao = O.PeriodicTensorTrain([a.copy() for a in A.tensors(b)], z=O.LogNum(logabs=zb.logabs, sign=zb.sign))
            for (left, nrm), (e, z) in envs.items():
                eo, zo = (O.accumulate_L if left else O.accumulate_R)(Ao, nrm)
                e, z = (e[b], z[b]) if B > 1 else (e, z)
                for a_, b_ in zip(e, eo):
                    assert a_.shape == b_.shape and np.allclose(a_, b_, rtol=1e-10, atol=0), (stage, left, nrm)
                assert close(z.logabs, zo.logabs, REL, 1e-12) and z.sign == zo.sign
            zo = O.normalization(Ao)
            assert close(zs[b].logabs, zo.logabs, REL, 1e-12) and zs[b].sign == zo.sign
            for a_, b_ in zip(m1[b], O.marginals(Ao)):
                assert np.allclose(a_, b_, rtol=1e-10, atol=0)
            m2o = O.twovar_marginals(Ao)
            assert set(m2[b]) == set(m2o)
            for k in m2o:
                assert np.allclose(m2[b][k], m2o[k], rtol=1e-10, atol=0), (stage, b, k)
            m3o = O.twovar_marginals(Ao, maxdist=3)
            assert set(m3[b]) == set(m3o)
            for k in m3o:
                assert np.allclose(m3[b][k], m3o[k], rtol=1e-10, atol=0), (stage, b, k)
        T.compress_(A, T.TruncThresh(1e-4))


def test_batched_periodic_matches_single():
    rng = RNG(150)
    B, bonds, q = 5, [4, 6, 3, 5], (2,)
    n = len(bonds)
    tsb = [rng.random((B, bonds[t], bonds[(t + 1) % n]) + q) for t in range(n)]
    A = T.PeriodicTensorTrain(tsb, batched=True)
    lz = T.normalize_(A)
    m2 = T.twovar_marginals(A)
    T.compress_(A, T.TruncThresh(1e-6))
    for b in range(B):
        Ao = O.PeriodicTensorTrain([a[b].copy() for a in tsb])
        assert close(lz[b], O.normalize(Ao), REL)
        m2o = O.twovar_marginals(Ao)
        for k in m2o:
            assert np.allclose(m2[b][k], m2o[k], rtol=1e-10, atol=0)
        O.compress(Ao, O.TruncThresh(1e-6))
        assert T.bond_dims(A, b) == O.bond_dims(Ao)
        X = rand_points(rng, Ao, 4)
        g = T.evaluate(A, X, b=b)
        for i in range(4):
            assert close(g[i], o_eval_flat(Ao, X[i]), 1e-8)


# ---- sampling ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("periodic", [False, True])
def test_sampling_injected_uniforms_bit_parity(periodic):
    rng = RNG(160)
    if periodic:
        ts = rand_per(rng, [3, 4, 2, 3, 3], (3,))
    else:
        ts = rand_open(rng, [1, 5, 7, 3, 6, 1], (2, 2))
    A, Ao = pair(ts, periodic=periodic)
    n = 64
    us = rng.random((n, len(ts)))
    X, p = T.sample(A, n, uniforms=us)
    rz = O.accumulate_R(Ao)
    for i in range(n):
        xo, po = O.sample(lambda t: us[i, t], Ao, rz=rz)
        assert list(X[i]) == xo, i
        assert close(p[i], po, 1e-9)
    T.normalize_(A); O.normalize(Ao)
    X2, p2 = T.sample(A, n, uniforms=us)
    assert np.array_equal(X, X2)
    g = T.evaluate(A, X2.astype(np.int32))
    assert np.allclose(g, p2, rtol=1e-9)
    x1, p1 = T.sample(A, uniforms=us[:1])
    assert [int(np.ravel_multi_index(tuple(v), A.q, order="F")) for v in x1] == list(X[0])


def test_sampling_philox_stream_and_chi_square():
    rng = RNG(161)
    ts = rand_open(rng, [1, 4, 6, 5, 1], (3,))
    A, Ao = pair(ts)
    T.normalize_(A); O.normalize(Ao)
    n, seed, first = 20000, 1234, 77
    X, p = T.sample(A, n, seed=seed, first_index=first)
    # exact stream parity on a subset
    rz = O.accumulate_R(Ao)
    for i in range(32):
        xo, _ = O.sample(lambda t: float(O.philox_uniform(seed, first + i, t)), Ao, rz=rz)
        assert list(X[i]) == xo
    # independence of the launch split: same global indices, two calls
    Xa, _ = T.sample(A, 100, seed=seed, first_index=first)
    Xb, _ = T.sample(A, 60, seed=seed, first_index=first + 40)
    assert np.array_equal(X[:100], Xa) and np.array_equal(X[40:100], Xb)
    # chi-square on single- and two-site marginals (critical values at p = 1e-6)
    m1 = O.marginals(Ao)
    for t in range(4):
        cnt = np.bincount(X[:, t], minlength=3)
        chi = np.sum((cnt - n * m1[t]) ** 2 / (n * m1[t]))
        assert chi < 27.6, (t, chi)          # dof 2
    m2 = O.twovar_marginals(Ao)
    for (t, u), b in m2.items():
        cnt = np.zeros((3, 3))
        np.add.at(cnt, (X[:, t], X[:, u]), 1)
        chi = np.sum((cnt - n * b) ** 2 / (n * b))
        assert chi < 44.0, (t, u, chi)       # dof 8
    hist, slp = T.sample_stats(A, n, seed=seed, first_index=first)
    assert np.array_equal(hist, np.stack([np.bincount(X[:, t], minlength=3) for t in range(4)]))
    assert close(slp, float(np.sum(np.log(p))), 1e-9)


def test_sampling_negative_values_raise():
    rng = RNG(162)
    ts = rand_open(rng, [1, 6, 7, 5, 1], (2, 2))
    ts[0] = -ts[0]
    A, Ao = pair(ts)
    zg = T.normalization(A)
    assert zg.sign == -1 and close(float(zg), O.exact_normalization(Ao), 1e-10)
    with pytest.raises(T.TTBError):
        T.sample(A, 4, seed=1)


# ---- the named configurations (reduced where the CPU oracle would take minutes) -----------------------
def test_config2_pipeline():
    """C2: L=128, q=2, d=64: orthogonalize_right!(eps=0) + compress!(TruncBond(32), :right) +
    normalize! + marginals."""
    rng = RNG(2)
    L, d = 128, 64
    ts = rand_open(rng, [1] + [d] * (L - 1) + [1], (2,))
    A, Ao = pair(ts)
    X = rand_points(rng, Ao, 6)
    T.orthogonalize_right_(A, T.TruncThresh(0.0)); O.orthogonalize_right(Ao, O.TruncThresh(0.0))
    assert T.bond_dims(A) == O.bond_dims(Ao)
    A.record_spectrum(True)
    stats = []
    T.compress_(A, T.TruncBond(32), is_orthogonal="right")
    O.compress(Ao, O.TruncBond(32), is_orthogonal="right", stats=stats)
    assert T.bond_dims(A) == O.bond_dims(Ao)
    for (t, lam_o, kept) in stats:
        lam_g = A.spectrum(t)
        assert spectra_close(lam_g, lam_o), (t, lam_g, lam_o)
    lg, lo = T.normalize_(A), O.normalize(Ao)
    assert close(lg, lo, 1e-8)
    for a, b in zip(T.marginals(A), O.marginals(Ao)):
        assert np.allclose(a, b, rtol=1e-8, atol=1e-12)
    g = T.evaluate(A, X)
    for i in range(X.shape[0]):
        assert close(g[i], o_eval_flat(Ao, X[i]), 1e-7, 1e-300)


@pytest.mark.parametrize("fill", ["rand", "randn"])
def test_config3_reduced(fill):
    """C3 at reduced length: L=12, q=2, d=256, compress!(TruncThresh(1e-10))."""
    rng = RNG(3)
    L, d = 12, 256
    bonds = [1] + [d] * (L - 1) + [1]
    ts = rand_open(rng, bonds, (2,), fill)
    A, Ao = pair(ts)
    X = rand_points(rng, Ao, 6)
    A.record_spectrum(True)
    stats = []
    T.compress_(A, T.TruncThresh(1e-10)); O.compress(Ao, O.TruncThresh(1e-10), stats=stats)
    assert T.bond_dims(A) == O.bond_dims(Ao), (T.bond_dims(A), O.bond_dims(Ao))
    for (t, lam_o, kept) in stats:
        lam_g = A.spectrum(t)
        assert spectra_close(lam_g, lam_o), (t, lam_g, lam_o)
    check_evals(A, Ao, X, 1e-8, fill == "randn")
    zg, zo = T.normalization(A), O.normalization(Ao)
    if fill == "rand":
        assert zg.sign == zo.sign and close(zg.logabs, zo.logabs, 1e-8, 1e-10)


def test_config4_reduced_sampling():
    """C4 at reduced length: L=24, q=4, d=128; normalize! then sample."""
    rng = RNG(4)
    L, d = 24, 128
    ts = rand_open(rng, [1] + [d] * (L - 1) + [1], (4,))
    A, Ao = pair(ts)
    T.normalize_(A); O.normalize(Ao)
    n = 64          # >= 32 draws: the tiled DMMA sampler (DP = 128)
    us = rng.random((n, L))
    X, p = T.sample(A, n, uniforms=us)
    rz = O.accumulate_R(Ao)
    for i in range(n):
        xo, po = O.sample(lambda t: us[i, t], Ao, rz=rz)
        assert list(X[i]) == xo and close(p[i], po, 1e-9)


# ---- dot / norm / norm2m (first row after the hot path, src/abstract_tensor_train.jl:395-437) --------
@pytest.mark.parametrize("periodic", [False, True])
def test_dot_norm_match_oracle(periodic):
    rng = RNG(700 + periodic)
    q = (2, 3)
    if periodic:
        ba, bb = [3, 5, 4, 2], [2, 3, 6, 4]
        ta, tb = rand_per(rng, ba, q), rand_per(rng, bb, q)
    else:
        ba, bb = [1, 5, 40, 7, 1], [1, 3, 33, 6, 1]
        ta, tb = rand_open(rng, ba, q, "randn"), rand_open(rng, bb, q, "randn")
    A, Ao = pair(ta, periodic, z=-2.5)
    B, Bo = pair(tb, periodic, z=0.75)
    assert close(T.dot(A, B), O.dot(Ao, Bo), 1e-10)
    assert close(T.dot(B, A), O.dot(Ao, Bo), 1e-10)
    assert close(T.norm(A), O.norm(Ao), 1e-10)
    assert close(T.norm2m(A, B), O.norm2m(Ao, Bo), 1e-9)
    # brute force: the dot is the sum over all configurations of the product of the two evaluations
    assert close(T.dot(A, B), O.exact_dot(Ao, Bo), 1e-10)


def test_dot_tracks_truncation_error_and_isapprox():
    """norm2m(A, compress(A)) is the squared truncation error (test/tensor_train.jl:88-106 uses it as
    the accuracy measure); isapprox follows (:434-437)."""
    rng = RNG(710)
    bonds = [1, 6, 20, 20, 6, 1]
    A, Ao = pair(rand_open(rng, bonds, (3,)))
    B = A.copy()
    T.compress_(B, T.TruncBond(4))
    e2 = T.norm2m(A, B)
    assert 0.0 <= e2 <= 1e-2 * T.norm(A) ** 2
    assert T.isapprox(A, A.copy()) and T.isapprox(A, B, rtol=1.0)
    C2 = A.copy()
    T.compress_(C2, T.TruncThresh(1e-12))
    assert T.isapprox(A, C2)              # default rtol 1e-6; the difference of squares resolves ~1e-8
    # long chain: the plain value would overflow, the logarithmic one does not
    L = 400
    tl = [rng.random((1 if t == 0 else 4, 1 if t == L - 1 else 4, 2)) * 50.0 for t in range(L)]
    G, Go = pair(tl)
    d = T.dot_log(G, G)
    assert d.sign == 1 and math.isfinite(d.logabs) and d.logabs > 709.0
    n = T.normalization(G)
    assert d.logabs < 2 * n.logabs + 1e-6     # sum of squares <= (sum)^2 for a positive train


def test_dot_shape_errors():
    rng = RNG(720)
    A = T.TensorTrain(rand_open(rng, [1, 3, 1], (2,)))
    B = T.TensorTrain(rand_open(rng, [1, 3, 3, 1], (2,)))
    with pytest.raises(ValueError):
        T.dot(A, B)


# ---- matrices too large for one CTA's shared memory: the cluster Jacobi (jacobi_cluster.cuh) -----------
@pytest.mark.parametrize("case", ["randn150_bond97", "rand136_left_first"])
def test_large_bond_cluster_jacobi(case):
    rng = RNG(150)
    if case == "randn150_bond97":
        L, d = 10, 150                      # k = 150: 16 blocks of 10 rows, the last one padded
        ts = rand_open(rng, [1] + [d] * (L - 1) + [1], (2,), "randn")
        A, Ao = pair(ts)
        X = rand_points(rng, Ao, 6)
        A.record_spectrum(True)
        stats = []
        T.compress_(A, T.TruncBond(97)); O.compress(Ao, O.TruncBond(97), stats=stats)
        assert T.bond_dims(A) == O.bond_dims(Ao)
        for (t, lam_o, kept) in stats:
            assert spectra_close(A.spectrum(t), lam_o), t
        g = T.evaluate(A, X)
        e = np.array([o_eval_flat(Ao, X[i]) for i in range(X.shape[0])])
        assert np.max(np.abs(g - e)) <= 1e-8 * np.max(np.abs(e))
        assert all(T.is_left_canonical(A, t) for t in range(0, L - 1))
    else:
        L, d = 9, 136                       # truncating LEFT sweep first: full-rank 136-wide SVDs of a rand fill
        ts = rand_open(rng, [1] + [d] * (L - 1) + [1], (2,), "rand")
        A, Ao = pair(ts)
        X = rand_points(rng, Ao, 6)
        A.record_spectrum(True)
        stats = []
        T.orthogonalize_left_(A, T.TruncThresh(1e-7)); O.orthogonalize_left(Ao, O.TruncThresh(1e-7), stats=stats)
        assert T.bond_dims(A) == O.bond_dims(Ao)
        for (t, lam_o, kept) in stats:
            assert spectra_close(A.spectrum(t), lam_o), t
        check_evals(A, Ao, X, 1e-6)


def test_async_upload_and_reset_reuse_handle():
    """ttb_reset + ttb_upload_all_async: a long-lived handle takes a stream of trains; the result is the one
    a fresh handle gives, for the chunked path (single train) and the fallback (batch)."""
    rng = RNG(900)
    L, d, q = 40, 96, 2                       # 2.8 MB per bulk site: several 16 MB chunks
    bonds = [1] + [d] * (L - 1) + [1]
    H = T.TensorTrain.empty(bonds, (q,))
    nd = sum(bonds[t] * bonds[t + 1] * q for t in range(L))
    pin = T.PinnedBuffer(nd)
    X = rng.integers(0, q, (5, L)).astype(np.int32)
    for rep in range(3):
        ts = [rng.random((bonds[t], bonds[t + 1], q)) for t in range(L)]
        pin.array[:] = np.concatenate([a.reshape(-1, order="F") for a in ts])
        F = T.TensorTrain(ts)
        T.compress_(F, T.TruncThresh(1e-9))
        H.reset()
        assert T.bond_dims(H) == bonds[:L]
        H.upload_all(pin.array, async_=True)
        T.compress_(H, T.TruncThresh(1e-9))
        assert T.bond_dims(H) == T.bond_dims(F)
        assert np.allclose(T.evaluate(H, X), T.evaluate(F, X), rtol=1e-12, atol=0.0)
        zh, zf = T.normalization(H), T.normalization(F)
        assert zh.sign == zf.sign and close(zh.logabs, zf.logabs, 1e-12, 1e-12)
    # async upload followed by something that is not a right sweep: joined before use
    H.reset()
    H.upload_all(pin.array, async_=True)
    assert np.allclose(T.evaluate(H, X), T.evaluate(T.TensorTrain(ts), X), rtol=1e-13)


@pytest.mark.parametrize("d,q", [(33, 2), (64, 2), (65, 3), (100, 2), (130, 3), (255, 2), (170, 3)])
def test_eps0_sweeps_odd_sizes(d, q):
    """eps = 0 sweeps over bond sizes that are not multiples of the 32-column panel / 32-row register group
    (single-cluster QR with a partial last block, the panel/update path beyond 8 blocks or 512 rows, the
    shared-memory fallbacks): gauge invariance, canonical form, exact bond dimensions."""
    rng = RNG(1000 + d)
    L = 6
    bonds = [1, min(q, d), d, d, min(q * q, d), min(q, d), 1]
    ts = rand_open(rng, bonds, (q,), "randn")
    A, Ao = pair(ts)
    X = rand_points(rng, Ao, 8)
    e0 = np.array([o_eval_flat(Ao, X[i]) for i in range(X.shape[0])])
    T.orthogonalize_right_(A, T.TruncThresh(0.0))
    g = T.evaluate(A, X)
    assert np.max(np.abs(g - e0)) <= 1e-10 * np.max(np.abs(e0))
    assert all(T.is_right_canonical(A, t, atol=1e-10) for t in range(1, L))
    O.orthogonalize_right(Ao, O.TruncThresh(0.0))
    assert T.bond_dims(A) == O.bond_dims(Ao)
    T.orthogonalize_left_(A, T.TruncThresh(0.0))
    g = T.evaluate(A, X)
    assert np.max(np.abs(g - e0)) <= 1e-10 * np.max(np.abs(e0))
    assert all(T.is_left_canonical(A, t, atol=1e-10) for t in range(0, L - 1))


# ---- A + B / A - B (section 8f rank 3): test/tensor_train.jl:302-345, test/periodic_tensor_train.jl:176-200
@pytest.mark.parametrize("periodic", [False, True])
@pytest.mark.parametrize("q", [(2,), (3, 2)])
def test_sum_difference_then_compress(periodic, q):
    rng = RNG(900 + len(q) + 2 * periodic)
    L = 6
    if periodic:
        ba, bb = list(rng.integers(1, 7, L)), list(rng.integers(1, 7, L))
        ta, tb = rand_per(rng, ba, q), rand_per(rng, bb, q)
    else:
        ba, bb = [1] + list(rng.integers(1, 8, L - 1)) + [1], [1] + list(rng.integers(1, 8, L - 1)) + [1]
        ta, tb = rand_open(rng, ba, q), rand_open(rng, bb, q)
    A, Ao = pair(ta, periodic=periodic, z=-0.6)
    B, Bo = pair(tb, periodic=periodic, z=2.5)
    for Cg, Co in ((A + B, O.plus(Ao, Bo)), (A - B, O.minus(Ao, Bo))):
        assert T.bond_dims(Cg) == O.bond_dims(Co)
        for g, o in zip(Cg.tensors(), Co.tensors):
            assert g.shape == o.shape and np.allclose(g, o, rtol=1e-14, atol=0)
        assert close(Cg.z.logabs, Co.z.logabs, 1e-14, 1e-15) and Cg.z.sign == Co.z.sign
        X = rand_points(rng, Co, 8)
        check_evals(Cg, Co, X, REL, signed=True)
    # the canonical producer of over-sized bonds: A + A compresses back to the bonds of A
    S, So = A + A, O.plus(Ao, Ao)
    T.compress_(S, T.TruncThresh(1e-9))
    O.compress(So, O.TruncThresh(1e-9))
    assert T.bond_dims(S) == O.bond_dims(So)
    check_evals(S, So, rand_points(rng, So, 8), 1e-8, signed=True)
    with pytest.raises(ValueError):
        A + T.TensorTrain(rand_open(rng, [1, 2, 1], q)) if not periodic else A + T.PeriodicTensorTrain(rand_per(rng, [2, 2], q))


def test_sum_of_batched_trains_after_compress():
    """ragged per-train bonds (after a truncating sweep) add up train by train"""
    rng = RNG(930)
    Bn, bonds, q = 4, [5, 6, 4, 5, 6], (2,)
    n = len(bonds)
    tsa = [rng.random((Bn, bonds[t], bonds[(t + 1) % n]) + q) for t in range(n)]
    tsb = [rng.random((Bn, bonds[t], bonds[(t + 1) % n]) + q) for t in range(n)]
    A, B = T.PeriodicTensorTrain(tsa, batched=True), T.PeriodicTensorTrain(tsb, batched=True)
    T.compress_(A, T.TruncThresh(1e-3))
    S = A + B
    for b in range(Bn):
        Ao = O.PeriodicTensorTrain([a.copy() for a in A.tensors(b)])
        Bo = O.PeriodicTensorTrain([a[b].copy() for a in tsb])
        So = O.plus(Ao, Bo)
        assert T.bond_dims(S, b) == O.bond_dims(So)
        for g, o in zip(S.tensors(b), So.tensors):
            assert g.shape == o.shape and np.array_equal(g, o)


@pytest.mark.parametrize("periodic", [True, False])
def test_batched_small_step_kernels(periodic):
    """batches with B*Q >= 64 take the one-CTA-per-train small-step kernels (k_absorb_small, k_apply_q_small)"""
    rng = RNG(950 + periodic)
    Bn, q = 36, (2,)
    bonds = [4, 6, 3, 5, 7] if periodic else [1, 4, 7, 6, 3, 1]
    n = len(bonds) if periodic else len(bonds) - 1
    ts = [rng.random((Bn, bonds[t], bonds[(t + 1) % len(bonds)]) + q) for t in range(n)]
    cls, ocls = (T.PeriodicTensorTrain, O.PeriodicTensorTrain) if periodic else (T.TensorTrain, O.TensorTrain)
    for trunc, otrunc in ((T.TruncThresh(1e-5), O.TruncThresh(1e-5)), (T.TruncBond(3), O.TruncBond(3))):
        A = cls(ts, batched=True)
        T.compress_(A, trunc)
        for b in range(0, Bn, 5):
            Ao = ocls([a[b].copy() for a in ts])
            O.compress(Ao, otrunc)
            assert T.bond_dims(A, b) == O.bond_dims(Ao), (b, T.bond_dims(A, b), O.bond_dims(Ao))
            X = rand_points(rng, Ao, 4)
            gv = T.evaluate(A, X, b=b)
            for i in range(4):
                assert close(gv[i], o_eval_flat(Ao, X[i]), 1e-8)


# ---- MPS wrapper (section 8f rank 2, the part that maps onto existing kernels): test/mps.jl:18-85 ------------
@pytest.mark.parametrize("z", [1.0, 2.0, -0.5])
def test_mps_wrapper_matches_oracle(z):
    rng = RNG(960)
    ts = [rng.standard_normal(s) for s in [(1, 5, 2, 2), (5, 4, 2, 2), (4, 10, 2, 2), (10, 1, 2, 2)]]
    psi, psio = pair(ts, z=z)
    p, po = T.MPS(psi), O.MPS(psio)
    zn, zno = T.mps_normalization(p), O.mps_normalization(po)
    assert zn.sign == zno.sign == 1 and close(zn.logabs, zno.logabs, REL, 1e-12)
    X = [tuple(int(rng.integers(0, 2)) for _ in range(2)) for _ in range(4)]
    assert close(T.mps_evaluate(p, X), O.mps_evaluate(po, X), REL)
    px = T.mps_evaluate(p, X)
    T.mps_orthogonalize_left_(p, T.TruncThresh(0.0))                      # test/mps.jl:71-75
    assert close(T.mps_evaluate(p, X), px, 1e-9)
    T.mps_orthogonalize_right_(p, T.TruncThresh(0.0))
    assert close(T.mps_evaluate(p, X), px, 1e-9)
    logz, logzo = T.mps_normalize_(p), O.mps_normalize(po)
    assert close(logz, logzo, REL, 1e-12)
    one = T.mps_normalization(p)
    assert abs(one.logabs) <= 1e-12 and one.sign == 1
    assert close(T.mps_evaluate(p, X), O.mps_evaluate(po, X), 1e-9)


def _mps_case(name, rng):
    if name == "ref":                   # test/mps.jl:14-15 (real Float64)
        shapes = [(1, 5, 2, 2), (5, 4, 2, 2), (4, 10, 2, 2), (10, 1, 2, 2)]
    elif name == "wide":                # bonds above one 64 x 64 GEMM tile, ragged, Q = 3
        shapes = [(1, 4, 3), (4, 66, 3), (66, 70, 3), (70, 5, 3), (5, 3, 3), (3, 1, 3)]
    else:                               # long chain: the per-step rescaling matters
        shapes = [(1, 6, 2)] + [(6, 6, 2)] * 28 + [(6, 1, 2)]
    return [rng.standard_normal(s) for s in shapes]


@pytest.mark.parametrize("case", ["ref", "wide", "long"])
def test_mps_environments_marginals_sampling_match_oracle(case):
    """accumulate_L/R(::MPS) stored, marginals, twovar_marginals, sample! (src/MatrixProductStates/mps.jl:60-126,
    176-232, 264-290) against the oracle; brute force on the reference's own test case (test/mps.jl:31-33)."""
    rng = RNG(970)
    ts = _mps_case(case, rng)
    psi, psio = pair(ts)
    p, po = T.MPS(psi), O.MPS(psio)
    for normalize in (True, False) if case != "long" else (True,):
        for fg, fo in ((T.mps_accumulate_L, O.mps_accumulate_L), (T.mps_accumulate_R, O.mps_accumulate_R)):
            eg, zg = fg(p, normalize)
            eo, zo = fo(po, normalize)
            assert zg.sign == zo.sign and close(zg.logabs, zo.logabs, 1e-10, 1e-12)
            for a, b in zip(eg, eo):
                assert a.shape == b.shape and np.allclose(a, b, rtol=1e-10, atol=1e-12 * np.abs(b).max())
    zn = T.mps_normalization(p)
    zL = T.mps_accumulate_L(p)[1]
    assert close(zn.logabs, zL.logabs, 1e-10, 1e-12)            # z / abs2(psi.z), psi.z = 1
    for a, b in zip(T.mps_marginals(p), O.mps_marginals(po)):
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-10, atol=1e-13)
    bo = O.mps_twovar_marginals(po)
    bg = T.mps_twovar_marginals(p)
    assert set(bg) == set(bo)
    for k in bo:
        assert np.allclose(bg[k], bo[k], rtol=1e-9, atol=1e-12), k
    b3 = T.mps_twovar_marginals(p, maxdist=2)
    assert set(b3) == {k for k in bo if k[1] - k[0] <= 2} and all(np.array_equal(b3[k], bg[k]) for k in b3)
    if case == "ref":                   # exact_marginals / exact_twovar_marginals (test/exact.jl)
        qq = O.mps_exact_prob(po).reshape([4] * 4)
        mg = T.mps_marginals(p)
        for t in range(4):
            pt = qq.sum(axis=tuple(i for i in range(4) if i != t))
            assert np.allclose(pt / pt.sum(), mg[t].reshape(-1, order="F"), rtol=1e-10, atol=0)
        for (t, u), b in bg.items():
            ptu = qq.sum(axis=tuple(i for i in range(4) if i not in (t, u)))
            assert np.allclose(ptu / ptu.sum(), b.reshape(4, 4, order="F"), rtol=1e-10, atol=0)
    n = 24
    us = rng.random((n, len(ts)))
    X, qg = T.mps_sample(p, uniforms=us)
    zR = O.mps_accumulate_R(po)
    for i in range(n):
        xo, qo = O.mps_sample(lambda t: us[i, t], po, rz=zR)
        flat = [int(np.ravel_multi_index(x, psi.q, order="F")) for x in xo]
        assert list(X[i]) == flat and close(qg[i], qo, 1e-9)
    # keyword arguments l=, r=, rz= (mps.jl:176-177, 206-208, 264-265): same results, the accumulate passes paid once
    lg, _ = T.mps_accumulate_L(p)
    rg, zr = T.mps_accumulate_R(p)
    for a, b in zip(T.mps_marginals(p), T.mps_marginals(p, l=lg, r=rg)):
        assert np.allclose(a, b, rtol=1e-13, atol=0)
    b2 = T.mps_twovar_marginals(p, l=lg, r=rg)
    assert all(np.allclose(b2[k], bg[k], rtol=1e-12, atol=0) for k in bg)
    Xr, qr = T.mps_sample(p, uniforms=us, rz=(rg, zr))
    assert np.array_equal(Xr, X) and np.allclose(qr, qg, rtol=1e-12)
    with pytest.raises(ValueError):
        T.mps_marginals(p, l=lg[:-1])
    X1, q1 = T.mps_sample(p, 300, seed=5)
    X2, q2 = T.mps_sample(p, 100, seed=5, first_index=200)
    assert np.array_equal(X1[200:], X2) and np.allclose(q1[200:], q2, rtol=1e-13)     # Philox counter = global draw index


def test_mps_orthogonal_environments_and_unsupported():
    """test/mps.jl:71-80: after orthogonalize_left!/right! the unnormalised environments are identities;
    a periodic psi is refused loudly (no CPU path)."""
    rng = RNG(971)
    psi, _ = pair(_mps_case("ref", rng))
    p = T.MPS(psi)
    T.mps_orthogonalize_left_(p, T.TruncThresh(0.0))
    for l in T.mps_accumulate_L(p, False)[0][:-1]:
        d = l.shape[1]
        assert np.allclose(l.reshape(d, d), np.eye(d), atol=1e-10)
    T.mps_orthogonalize_right_(p, T.TruncThresh(0.0))
    for r in T.mps_accumulate_R(p, False)[0][1:]:
        d = r.shape[0]
        assert np.allclose(r.reshape(d, d), np.eye(d), atol=1e-10)


@pytest.mark.parametrize("bonds,q", [([3, 4, 5, 2], (2,)), ([2, 6, 6, 6, 6, 6, 3], (3,))])
def test_mps_periodic_psi_matches_oracle(bonds, q):
    """A periodic psi: the environments are the reference's genuine four-index arrays (mps.jl:60-103), kept on the
    device as d_1^2 boundary slices; accumulate_L/R, marginals and sample! against the oracle, marginals against
    brute force; twovar_marginals(::MPS) is TensorTrain-only like the reference (mps.jl:206)."""
    rng = RNG(972)
    ts = [rng.standard_normal((bonds[t], bonds[(t + 1) % len(bonds)]) + q) for t in range(len(bonds))]
    psi, psio = pair(ts, periodic=True)
    p, po = T.MPS(psi), O.MPS(psio)
    for normalize in (True, False):
        for fg, fo in ((T.mps_accumulate_L, O.mps_accumulate_L), (T.mps_accumulate_R, O.mps_accumulate_R)):
            eg, zg = fg(p, normalize)
            eo, zo = fo(po, normalize)
            assert zg.sign == zo.sign and close(zg.logabs, zo.logabs, 1e-10, 1e-12)
            for a, b in zip(eg, eo):
                assert a.shape == b.shape and np.allclose(a, b, rtol=1e-10, atol=1e-12 * np.abs(b).max())
    zn = T.mps_normalization(p)
    assert close(zn.logabs, T.mps_accumulate_L(p)[1].logabs, 1e-10, 1e-12)
    mg = T.mps_marginals(p)
    for a, b in zip(mg, O.mps_marginals(po)):
        assert np.allclose(a, b, rtol=1e-10, atol=1e-13)
    Qf = int(np.prod(q))
    qq = O.mps_exact_prob(po).reshape([Qf] * len(ts))
    for t in range(len(ts)):
        pt = qq.sum(axis=tuple(i for i in range(len(ts)) if i != t))
        assert np.allclose(pt / pt.sum(), mg[t].reshape(-1, order="F"), rtol=1e-10, atol=0)
    n = 16
    us = rng.random((n, len(ts)))
    X, qg = T.mps_sample(p, uniforms=us)
    zR = O.mps_accumulate_R(po)
    for i in range(n):
        xo, qo = O.mps_sample(lambda t: us[i, t], po, rz=zR)
        flat = [int(np.ravel_multi_index(x, psi.q, order="F")) for x in xo]
        assert list(X[i]) == flat and close(qg[i], qo, 1e-9)
    with pytest.raises(ValueError):
        T.mps_twovar_marginals(p)


@pytest.mark.parametrize("case", ["open_d40", "periodic_d20", "batched"])
def test_twovar_marginals_block_is_a_slice_of_the_full_result(case):
    """ttb_twovar_marginals_range: the unit of the pair-sharded multi-GPU path (all three kernels)"""
    rng = RNG(970)
    if case == "open_d40":
        A = T.TensorTrain(rand_open(rng, [1, 40, 40, 33, 40, 12, 1], (2,)))       # k_twovar_fast
    elif case == "periodic_d20":
        A = T.PeriodicTensorTrain(rand_per(rng, [20, 16, 20, 9, 14], (2,)))       # k_twovar_mma
    else:
        bonds = [4, 6, 3, 5]
        A = T.PeriodicTensorTrain([rng.random((3, bonds[t], bonds[(t + 1) % 4], 2)) for t in range(4)], batched=True)
    L = A.L
    for md in (None, 2):
        full = T.twovar_marginals(A, md)
        full = full if A.batched else [full]
        mdv = L - 1 if md is None else md
        for lo, hi in ((0, L - 1), (0, 1), (1, 3), (L - 2, L - 1), (2, 2)):
            blk = T.twovar_marginals_block(A, lo, hi, md)
            i = 0
            for t in range(lo, hi):
                for u in range(t + 1, min(L - 1, t + mdv) + 1):
                    for b in range(A.batch):
                        assert np.array_equal(blk[b, i], full[b][(t, u)].reshape(-1, order="F")), (case, md, lo, hi, t, u)
                    i += 1
            assert i == blk.shape[1]


def test_sharded_pair_marginals_single_process_path():
    """parallel.twovar_marginals_sharded without a process group = one rank owning every pair (the N > 1 logic is
    covered by the gloo test; the per-rank block by test_twovar_marginals_block_is_a_slice_of_the_full_result)"""
    from ttb200 import parallel as P
    rng = RNG(980)
    A = T.TensorTrain(rand_open(rng, [1, 6, 9, 7, 5, 1], (3,)))
    for md in (None, 2):
        full, sh = T.twovar_marginals(A, md), P.twovar_marginals_sharded(A, md)
        assert set(full) == set(sh)
        for k in full:
            assert np.array_equal(full[k], sh[k])


def test_edge_cases_of_the_round_1b_entries():
    """shortest trains, argument errors and empty ranges of ttb_compose / ttb_twovar_marginals_range /
    ttb_scale_sites (the reference's ArgumentError paths: src/tensor_train.jl:241-242,
    src/periodic_tensor_train.jl:63-64)"""
    rng = RNG(990)
    # L = 2 open trains: first site [za A, zb B], last site [A; B], nothing in between
    ta, tb = rand_open(rng, [1, 3, 1], (2,)), rand_open(rng, [1, 5, 1], (2,))
    A, Ao = pair(ta, z=0.5)
    B, Bo = pair(tb, z=-3.0)
    for Cg, Co in ((A + B, O.plus(Ao, Bo)), (A - B, O.minus(Ao, Bo))):
        assert T.bond_dims(Cg) == O.bond_dims(Co) == [1, 8]
        check_evals(Cg, Co, rand_points(rng, Co, 4), REL, signed=True)
    # L = 1 periodic train (a single matrix-valued site): block diagonal, trace adds up
    pa, pb = rand_per(rng, [3], (2,)), rand_per(rng, [2], (2,))
    P1, P1o = pair(pa, periodic=True)
    P2, P2o = pair(pb, periodic=True)
    check_evals(P1 + P2, O.plus(P1o, P2o), rand_points(rng, P1o, 2), REL, signed=True)
    # mismatched physical dimensions / kinds
    with pytest.raises(ValueError):
        A + T.TensorTrain(rand_open(rng, [1, 3, 1], (3,)))
    with pytest.raises(ValueError):
        A + P1
    # pair ranges: empty and out-of-range
    C4 = T.TensorTrain(rand_open(rng, [1, 3, 4, 3, 1], (2,)))
    assert T.twovar_marginals_block(C4, 2, 2).shape == (1, 0, 4)
    with pytest.raises(ValueError):
        T.twovar_marginals_block(C4, 0, 4)          # first sites run over [0, L-1)
    with pytest.raises(ValueError):
        T.twovar_marginals_block(C4, 3, 1)
    # scale_sites: every tensor divided by the factor, z untouched, evaluate scales by factor^-L
    X = rand_points(rng, O.TensorTrain(C4.tensors()), 1)
    e0, z0 = float(T.evaluate(C4, X)[0]), C4.z
    T._ck(T.lib.ttb_scale_sites(C4._hd, T._dp(np.array([2.0]))))
    assert close(float(T.evaluate(C4, X)[0]), e0 / 2.0 ** 4, 1e-14) and C4.z.logabs == z0.logabs


# ---- round-2 boundary fixes ------------------------------------------------------------------------------
def test_jacobi_non_convergence_is_an_error(monkeypatch):
    """The reference's LAPACK svd throws when it does not converge (src/svd_trunc.jl:37); a Jacobi SVD that hits
    its sweep cap must not return TTB_OK.  TTB_JACOBI_MAXSWEEPS=1 forces the cap on every Jacobi variant."""
    rng = RNG(900)
    cases = [([1, 8, 8, 8, 1], (2,), 1),        # k_jacobi_small / g8 sized steps
             ([1, 48, 48, 1], (2,), 1),         # k_jacobi_g8<64> / k_jacobi_small<64,64>
             ([1, 160, 160, 1], (2,), 1)]       # cluster Jacobi
    for bonds, q, _ in cases:
        ts = rand_open(rng, bonds, q, "randn")
        A = T.TensorTrain(ts)
        monkeypatch.setenv("TTB_JACOBI_MAXSWEEPS", "1")
        with pytest.raises(T.SVDNotConvergedError):
            T.compress_(A, T.TruncThresh(1e-12))
        monkeypatch.delenv("TTB_JACOBI_MAXSWEEPS")
        B = T.TensorTrain(ts)
        T.compress_(B, T.TruncThresh(1e-12))     # the status word was cleared: the next call is clean
        assert T.bond_dims(B)[0] == 1


def test_truncbondmax_maxerr_covers_both_sweeps_of_center_calls():
    """src/svd_trunc.jl:98-100 updates maxerr on every SVD: orthogonalize_center! runs a left AND a right sweep."""
    rng = RNG(901)
    ts = rand_open(rng, [1, 6, 9, 9, 9, 6, 1], (3,), "randn")
    for l, two in ((3, False), (4, False), (3, True)):
        A, Ao = pair(ts)
        tg, to = T.TruncBondMax(2), O.TruncBondMax(2)
        if two:
            T.orthogonalize_two_site_center_(A, l, tg); O.orthogonalize_two_site_center(Ao, l, to)
        else:
            T.orthogonalize_center_(A, l, tg); O.orthogonalize_center(Ao, l, to)
        assert T.bond_dims(A) == O.bond_dims(Ao)
        assert close(tg.maxerr[0], to.maxerr[0], 1e-6), (l, two, tg.maxerr, to.maxerr)
        # the errors of the left half must be in it: the left sweep alone already gives a non-zero error
        Bo, tl = O.TensorTrain([a.copy() for a in ts]), O.TruncBondMax(2)
        O.orthogonalize_left(Bo, tl, range(1, l))
        assert tl.maxerr[0] > 0 and tg.maxerr[0] >= tl.maxerr[0] * (1 - 1e-6)


def test_spectrum_of_untouched_bonds_is_empty_and_capacity_checked():
    import ctypes as C
    rng = RNG(902)
    ts = rand_open(rng, [1, 12, 12, 12, 1], (2,))
    A = T.TensorTrain(ts)
    A.record_spectrum(True)
    T.orthogonalize_left_(A, T.TruncThresh(1e-9), indices=range(1, 3))     # touches bonds 1 and 2 only
    assert len(A.spectrum(1)) == 2 and len(A.spectrum(2)) > 2
    assert len(A.spectrum(3)) == 0 and len(A.spectrum(0)) == 0 and len(A.spectrum(4)) == 0
    out, n = np.zeros(1), C.c_int64()
    rc = T.lib.ttb_get_spectrum(A._hd, 0, 2, T._dp(out), 1, C.byref(n))     # too small: error, length reported
    assert rc == 1 and n.value == len(A.spectrum(2))
    T.orthogonalize_right_(A, T.TruncThresh(0.0))                           # eps = 0 sweep records nothing
    assert all(len(A.spectrum(i)) == 0 for i in range(5))


# ---- environment handles (l=, r=, rz=), accumulate_M, is_in_domain, == ---------------------------------------
@pytest.mark.parametrize("periodic", [False, True])
def test_environment_keywords_reuse_accumulate(periodic):
    """marginals / twovar_marginals with l=, r= and sample with rz= (src/abstract_tensor_train.jl:208-209, 231-233,
    335-336): same results as the default keywords, the accumulate pass paid once."""
    rng = RNG(910 + periodic)
    q = (2, 3)
    ts = rand_per(rng, [3, 5, 4, 2, 6], q) if periodic else rand_open(rng, [1, 4, 6, 5, 3, 1], q)
    A, Ao = pair(ts, periodic=periodic)
    T.normalize_(A); O.normalize(Ao)
    l, zl = T.accumulate_L(A, keep=True)
    r, zr = T.accumulate_R(A, keep=True)
    _, zo = O.accumulate_L(Ao)
    assert zl.sign == zo.sign and close(zl.logabs, zo.logabs, 1e-10, 1e-12)
    assert zr.sign == zo.sign and close(zr.logabs, zo.logabs, 1e-10, 1e-12)
    m0, m1 = T.marginals(A), T.marginals(A, l=l, r=r)
    for a, b, c in zip(m0, m1, O.marginals(Ao)):
        assert np.array_equal(a, b) and np.allclose(a, c, rtol=1e-10, atol=1e-14)
    m1 = T.marginals(A, l=l)            # one of the two given
    assert all(np.array_equal(a, b) for a, b in zip(m0, m1))
    p0, p1 = T.twovar_marginals(A), T.twovar_marginals(A, l=l, r=r, M=T.accumulate_M(A))
    po = O.twovar_marginals(Ao)
    for k in po:
        assert np.array_equal(p0[k], p1[k]) and np.allclose(p0[k], po[k], rtol=1e-10, atol=1e-14)
    us = rng.random((40, len(ts)))
    X0, q0 = T.sample(A, 40, uniforms=us)
    X1, q1 = T.sample(A, 40, uniforms=us, rz=r)
    assert np.array_equal(X0, X1) and np.array_equal(q0, q1)
    h0, s0 = T.sample_stats(A, 500, seed=3)
    h1, s1 = T.sample_stats(A, 500, seed=3, rz=(r, zr))      # the (env, z) pair itself is accepted too
    assert np.array_equal(h0, h1) and close(s0, s1, 1e-12, 0.0)   # sum of log p: atomics, order may differ
    # wrong side / stale environment
    with pytest.raises(ValueError):
        T.marginals(A, l=r)
    with pytest.raises(TypeError):
        T.marginals(A, l=[np.eye(3)])
    T.compress_(A, T.TruncBond(2))
    with pytest.raises(ValueError):
        T.marginals(A, l=l, r=r)


@pytest.mark.parametrize("periodic", [False, True])
def test_accumulate_M_matches_oracle(periodic):
    """src/abstract_tensor_train.jl:119-137: the table itself, and l[0] M[0,L-1] r[L-1] == Z for open trains
    (test/tensor_train.jl:185-196)."""
    rng = RNG(920 + periodic)
    ts = rand_per(rng, [3, 5, 4, 2, 6, 3], (2, 2)) if periodic else rand_open(rng, [1, 4, 6, 5, 3, 2, 1], (3,))
    A, Ao = pair(ts, periodic=periodic)
    L = len(ts)
    for normalize in (True, False):
        M, Mo = T.accumulate_M(A, normalize=normalize), O.accumulate_M(Ao, normalize=normalize)
        for t in range(L):
            for u in range(L):
                if t < u:
                    assert M[t][u].shape == Mo[(t, u)].shape
                    assert np.allclose(M[t][u], Mo[(t, u)], rtol=1e-12, atol=1e-300), (t, u)
                else:
                    assert M[t][u].shape == (0, 0)
    if not periodic:
        M = T.accumulate_M(A, normalize=False)
        l, z = T.accumulate_L(A, normalize=False)
        r, _ = T.accumulate_R(A, normalize=False)
        assert close((l[0] @ M[0][L - 1] @ r[L - 1]).item(), float(z), 1e-10)


def test_is_in_domain_and_equality():
    """src/abstract_tensor_train.jl:59-64, :85."""
    rng = RNG(930)
    ts = rand_open(rng, [1, 3, 4, 1], (2, 3))
    A, Ao = pair(ts)
    good = [(1, 2), (0, 0), (1, 1)]
    bad = [(1, 3), (0, 0), (1, 1)]
    assert T.is_in_domain(A, good) and O.is_in_domain(Ao, good)
    assert not T.is_in_domain(A, bad) and not O.is_in_domain(Ao, bad)
    assert not T.is_in_domain(A, [(2, 0), (0, 0), (0, 0)]) and not T.is_in_domain(A, good[:2])
    assert T.is_in_domain(A, np.array([[5, 0, 3], [0, 1, 2]])) and not T.is_in_domain(A, np.array([[6, 0, 0]]))
    B = A.copy()
    assert A == B and O.equal(Ao, Ao.copy())
    B.z = 3.0                                   # z is not part of == (isequal of the tensors only)
    assert A == B
    s1 = ts[1].copy(); s1[2, 1, 0, 2] = np.nextafter(s1[2, 1, 0, 2], 1.0)
    B.set_site(1, s1)
    Bo = Ao.copy(); Bo[1] = s1
    assert not (A == B) and A != B and not O.equal(Ao, Bo)
    # isequal semantics: NaN equals NaN, -0.0 differs from 0.0
    n1 = ts[1].copy(); n1[0, 0, 0, 0] = np.nan
    Cn, Dn = T.TensorTrain([ts[0], n1, ts[2]]), T.TensorTrain([ts[0], n1.copy(), ts[2]])
    assert Cn == Dn and not (Cn == A)
    z0, z1 = ts[1].copy(), ts[1].copy()
    z0[0, 0, 0, 0], z1[0, 0, 0, 0] = 0.0, -0.0
    assert not (T.TensorTrain([ts[0], z0, ts[2]]) == T.TensorTrain([ts[0], z1, ts[2]]))
    assert not O.equal(O.TensorTrain([ts[0], z0, ts[2]]), O.TensorTrain([ts[0], z1, ts[2]]))
    # other shapes / kinds are simply unequal
    assert not (A == T.TensorTrain(rand_open(rng, [1, 3, 5, 1], (2, 3))))
    P = T.PeriodicTensorTrain(rand_per(rng, [2, 3, 2], (2, 3)))
    assert not (A == P) and P == P.copy()
    # after a sweep the comparison runs on the CURRENT dims
    C1, C2 = A.copy(), A.copy()
    T.compress_(C1, T.TruncBond(2)); T.compress_(C2, T.TruncBond(2))
    assert C1 == C2 and not (C1 == A)


# ---- two-site pieces: src/utils.jl:37-44, src/abstract_tensor_train.jl:283-308, src/dmrg.jl:124-151 ----
@pytest.mark.parametrize("q", [(2,), (2, 3)])
def test_merge_and_split_tensor_match_oracle(q):
    rng = RNG(980)
    bonds = [1, 6, 9, 6, 4, 1]
    ts = rand_open(rng, bonds, q, "randn")
    A, Ao = pair(ts)
    Qf = int(np.prod(q))
    X = rand_points(rng, Ao, 6)
    e0 = [T.evaluate(A, x) for x in X]
    for t in range(len(ts) - 1):
        Cg, Co = T.merge_tensors(A, t), O.merge_tensors(Ao[t], Ao[t + 1])
        assert Cg.shape == Co.shape and np.allclose(Cg, Co, rtol=1e-12, atol=1e-13)
    with pytest.raises(ValueError):
        T.merge_tensors(A, len(ts) - 1)
    # exact split of the pair (1, 2): rank = the bond 9 it was merged over (<= capacity)
    for lr in ("left", "right"):
        B = A.copy()
        Cm = T.merge_tensors(B, 1)
        T.split_tensor_(B, 1, Cm, T.TruncBond(9), lr)      # (TruncThresh keeps one value more than the rank: svd_trunc.jl:38-47)
        Ao1, Ao2 = O.split_tensor(O.merge_tensors(Ao[1], Ao[2]), O.TruncBond(9), lr)
        assert T.bond_dims(B)[2] == Ao1.shape[1] == 9
        assert np.allclose(T.merge_tensors(B, 1), Cm, rtol=1e-10, atol=1e-12)          # A * B = C
        assert all(close(T.evaluate(B, x), e, 1e-9) for x, e in zip(X, e0)) and float(B.z) == float(A.z)
        a1, a2 = B.site(1), B.site(2)
        if lr == "left":        # A = U: left-orthonormal; B = L V': rows carry the singular values
            assert T.is_left_canonical(B, 1)
            sv = np.sort(np.linalg.norm(a2.reshape(a2.shape[0], -1), axis=1))[::-1]
            svo = np.sort(np.linalg.norm(Ao2.reshape(Ao2.shape[0], -1), axis=1))[::-1]
        else:
            assert T.is_right_canonical(B, 2)
            sv = np.sort(np.linalg.norm(a1.transpose(1, 0, *range(2, a1.ndim)).reshape(a1.shape[1], -1), axis=1))[::-1]
            svo = np.sort(np.linalg.norm(Ao1.transpose(1, 0, *range(2, Ao1.ndim)).reshape(Ao1.shape[1], -1), axis=1))[::-1]
        assert np.allclose(sv, svo, rtol=1e-9, atol=1e-12 * svo[0])
    # truncating split of a perturbed merged tensor (what a DMRG step hands back)
    B = A.copy()
    Cm = T.merge_tensors(B, 1) + 0.05 * rng.standard_normal((6, 6) + q + q)
    tb = T.TruncBondMax(4)
    err = T.split_tensor_(B, 1, Cm, tb, "left")
    So = O.TruncBondMax(4)
    Ao1, Ao2 = O.split_tensor(Cm, So, "left")
    assert T.bond_dims(B)[2] == 4 and close(err, So.maxerr[0], 1e-6) and close(tb.maxerr[0], So.maxerr[0], 1e-6)
    assert np.allclose(T.merge_tensors(B, 1), O.merge_tensors(Ao1, Ao2), rtol=1e-8, atol=1e-10)
    # a split that would need more than the capacity of the bond is refused, the train is untouched
    B = A.copy()
    with pytest.raises(NotImplementedError):
        T.split_tensor_(B, 1, Cm, T.TruncThresh(0.0), "left")
    assert T.bond_dims(B) == T.bond_dims(A) and np.array_equal(B.site(1), A.site(1))
    assert Qf >= 2


def test_data_environments_and_updates_match_oracle():
    rng = RNG(981)
    q = (2, 2)
    ts = rand_open(rng, [1, 4, 7, 5, 3, 1], q, "randn")
    A, Ao = pair(ts)
    L = len(ts)
    X = [[tuple(int(rng.integers(0, v)) for v in q) for _ in range(L)] for _ in range(9)]
    envL, envR = T.data_environments(A, X, left=True), T.data_environments(A, X, left=False)
    pl = [O.precompute_left_environments(Ao, x) for x in X]
    pr = [O.precompute_right_environments(Ao, x) for x in X]
    def check(env, ref):
        for n in range(len(X)):
            for j in range(L + 1):
                v = np.reshape(ref[n][j], -1)
                assert np.allclose(env[n, j, :v.size], v, rtol=1e-12, atol=1e-14) and not env[n, j, v.size:].any()
    check(envL, pl); check(envR, pr)
    # one DMRG-like step at the pair (1, 2): perturb, split keeping the sweep direction's gauge, update the environments
    for lr, left in (("left", True), ("right", False)):
        Cm = T.merge_tensors(A, 1) + 0.01 * rng.standard_normal((4, 5) + q + q)
        T.split_tensor_(A, 1, Cm, T.TruncBond(7), lr)
        Ao.tensors[1], Ao.tensors[2] = O.split_tensor(Cm, O.TruncBond(7), lr)
        O.update_environments(pl, pr, Ao, 1, X, lr)
        if left:
            T.update_environments_(envL, A, 1, X, left=True)
        else:
            T.update_environments_(envR, A, 2, X, left=False)
        # the two SVDs may differ by a gauge (signs): compare evaluations, and the updated entry with a fresh pass
        eg = [T.evaluate(A, x) for x in X]
        eo = [O.evaluate(Ao, x) for x in X]
        assert np.allclose(eg, eo, rtol=1e-9)
        full = T.data_environments(A, X, left=left)
        upd = envL if left else envR
        j = 2
        assert np.allclose(upd[:, j], full[:, j], rtol=1e-12, atol=1e-14)
        if left:
            envR = T.data_environments(A, X, left=False)
        else:
            envL = T.data_environments(A, X, left=True)


@pytest.mark.parametrize("case", ["ragged_q3", "d96_q2", "maxdist", "batched_q4"])
def test_twovar_open_group_kernel_matches_oracle(case):
    """k_twovar_open (open trains, bonds above 32): a group of first sites per CTA, K T_u on the tensor cores --
    against the oracle's twovar_marginals (src/abstract_tensor_train.jl:231-254) and against the generic kernel."""
    rng = RNG(990)
    md = None
    if case == "ragged_q3":
        bonds, q = [1, 3, 9, 27, 40, 37, 50, 33, 8, 3, 1], (3,)
    elif case == "d96_q2":
        bonds, q = [1] + [96] * 22 + [1], (2,)          # 23 sites: more first sites than one group of 16
    elif case == "maxdist":
        bonds, q, md = [1] + [48] * 19 + [1], (2,), 5
    else:
        bonds, q = [1, 4, 16, 40, 40, 16, 4, 1], (2, 2)
    if case == "batched_q4":
        B = 3
        tss = [rand_open(rng, bonds, q) for _ in range(B)]
        A = T.TensorTrain([np.stack([tss[b][t] for b in range(B)]) for t in range(len(bonds) - 1)], batched=True)
        got = T.twovar_marginals(A)
        for b in range(B):
            Ao = O.TensorTrain([a.copy() for a in tss[b]])
            po = O.twovar_marginals(Ao)
            assert set(got[b]) == set(po)
            for k in po:
                assert np.allclose(got[b][k], po[k], rtol=1e-10, atol=1e-14), (b, k)
        return
    ts = rand_open(rng, bonds, q)
    A, Ao = pair(ts)
    T.normalize_(A); O.normalize(Ao)
    got = T.twovar_marginals(A, maxdist=md)
    po = O.twovar_marginals(Ao, maxdist=md)
    assert set(got) == set(po)
    for k in po:
        assert np.allclose(got[k], po[k], rtol=1e-10, atol=1e-14), k
    L = len(ts)
    mdv = L - 1 if md is None else md
    blk = T.twovar_marginals_block(A, 3, L - 2, mdv)[0]      # a range that starts inside a group
    i = 0
    for t in range(3, L - 2):
        for u in range(t + 1, min(L - 1, t + mdv) + 1):
            assert np.array_equal(blk[i].reshape(q + q, order="F"), got[(t, u)])
            i += 1
    assert i == blk.shape[0]


def test_twovar_marginals_large_physical_dimension():
    """Q = 90 (> 71, the round-1 limit): the Q x Q table of a pair goes through the generic kernel"""
    rng = RNG(991)
    ts = rand_open(rng, [1, 3, 4, 2, 1], (90,))
    A, Ao = pair(ts)
    got, po = T.twovar_marginals(A), O.twovar_marginals(Ao)
    assert set(got) == set(po)
    for k in po:
        assert got[k].shape == (90, 90) and np.allclose(got[k], po[k], rtol=1e-10, atol=1e-15), k
    per, pero = pair(rand_per(rng, [2, 3, 2], (75,)), periodic=True)
    g2, p2 = T.twovar_marginals(per), O.twovar_marginals(pero)
    for k in p2:
        assert np.allclose(g2[k], p2[k], rtol=1e-10, atol=1e-15), k


def test_new_entries_argument_errors():
    """argument checking of the round-2 entries: errors come back as status codes -> the reference's exception classes"""
    rng = RNG(992)
    ts = rand_open(rng, [1, 3, 4, 1], (2,))
    A, _ = pair(ts)
    with pytest.raises(ValueError):
        T.merge_tensors(A, 2)                                  # no site pair (2, 3)
    with pytest.raises(ValueError):
        T.split_tensor_(A, 0, np.zeros((1, 4, 2)))             # wrong shape
    with pytest.raises(ValueError):
        T.split_tensor_(A, 0, T.merge_tensors(A, 0), lr="up")
    with pytest.raises(ValueError):
        T.data_environments(A, [[0, 1, 2]])                    # x = 2 outside 0:1
    with pytest.raises(ValueError):
        T.update_environments_(np.zeros((1, 2, 3)), A, 0, [[0, 1, 0]])
    with pytest.raises(ValueError):
        T.twovar_marginals_block(A, 2, 1)                      # empty / reversed range
    with pytest.raises(ValueError):
        T.twovar_marginals_block(A, 0, 2, out=np.zeros(3))     # destination too small
    per, _ = pair(rand_per(rng, [2, 3, 2], (2,)), periodic=True)
    with pytest.raises(ValueError):
        T.data_environments(per, [[0, 1, 0]])                  # TensorTrain only (src/tensor_train.jl:320)
    batched = T.TensorTrain([np.stack([a, a]) for a in ts], batched=True)
    with pytest.raises(NotImplementedError):
        T.MPS(batched)
    assert T.lib.ttb_merge_tensors(A._hd, 0, 0, None) != 0 and T.lib.ttb_mps_marginals(A._hd, 0, None) != 0
    assert T.lib.ttb_split_tensor(A._hd, 0, 0, None, None, 1, None) != 0


def test_mps_entries_on_a_train_of_a_batch():
    """the C entries take the train index b of a batched handle (the Python MPS mirror wraps single trains only)"""
    import ctypes as C
    rng = RNG(993)
    bonds, q = [1, 4, 6, 3, 1], (2,)
    tss = [[rng.standard_normal((bonds[t], bonds[t + 1]) + q) for t in range(4)] for _ in range(3)]
    A = T.TensorTrain([np.stack([tss[b][t] for b in range(3)]) for t in range(4)], batched=True)
    for b in (0, 2):
        po = O.MPS(O.TensorTrain([a.copy() for a in tss[b]]))
        out = np.zeros((4, 2))
        T._ck(T.lib.ttb_mps_marginals(A._hd, b, T._dp(out)))
        for t, m in enumerate(O.mps_marginals(po)):
            assert np.allclose(out[t], m.reshape(-1, order="F"), rtol=1e-10, atol=1e-13)
        la, sg = C.c_double(), C.c_int()
        T._ck(T.lib.ttb_mps_accumulate(A._hd, b, 1, 1, None, C.byref(la), C.byref(sg)))
        zo = O.mps_accumulate_L(po)[1]
        assert sg.value == zo.sign and close(la.value, zo.logabs, 1e-10, 1e-12)
        n = C.c_int64()
        T._ck(T.lib.ttb_twovar_count(A._hd, 3, C.byref(n)))
        pairs = np.zeros((n.value, 4))
        T._ck(T.lib.ttb_mps_twovar_marginals(A._hd, b, 3, T._dp(pairs)))
        bo = O.mps_twovar_marginals(po)
        i = 0
        for t in range(3):
            for u in range(t + 1, 4):
                assert np.allclose(pairs[i], bo[(t, u)].reshape(-1, order="F"), rtol=1e-9, atol=1e-12)
                i += 1
    assert T.lib.ttb_mps_marginals(A._hd, 3, T._dp(np.zeros((4, 2)))) != 0      # train index out of range
