"""Pins the CPU oracle with the reference's OWN test-suite invariants (the reference ships no
golden vectors and Julia is absent, see oracle/tt_oracle.py header).  Each test cites the reference
test it restates (paths relative to /root/reference)."""
import math

import numpy as np
import pytest

from oracle import tt_oracle as O

RNG = lambda s: np.random.Generator(np.random.Philox(s))


def _rand_x(rng, A):
    return [tuple(int(rng.integers(0, q)) for q in a.shape[2:]) for a in A]


# ---- svd_trunc rank rules: src/svd_trunc.jl known answers -------------------------------------
def test_rank_rules_known_answers():
    lam = np.array([1.0, 1e-1, 1e-2, 1e-3])
    # TruncThresh keeps one more than the minimal tail rule (svd_trunc.jl:38-47)
    assert O.TruncThresh(0.0).rank(lam) == 4
    assert O.TruncThresh(1.0).rank(lam) == 2          # tail from k=1 reaches ||lam||^2 -> k+1 = 2
    assert O.TruncThresh(0.05).rank(lam) == 3         # tail(k=2)=1e-2^2+..>=? (0.05*1.005)^2=2.5e-3 no; k=2 -> .0101>=.0025 -> 3
    assert O.TruncThresh(2.0).rank(lam) == 1          # never reached
    assert O.TruncThresh(0.0).rank(np.array([3.0])) == 1
    assert O.TruncThresh(1e-3).rank(np.array([np.nan, np.nan])) == 1
    # TruncBondThresh starts the tail at 0 and skips lam_m (svd_trunc.jl:130-138)
    assert O.TruncBondThresh(10, 0.0).rank(lam) == 4
    assert O.TruncBondThresh(2, 0.0).rank(lam) == 2
    assert O.TruncBondThresh(10, 0.0099).rank(lam) == 4   # lam_3^2=1e-4 >= (0.0099*1.005)^2 -> k=3 -> 4
    assert O.TruncBondThresh(10, 0.0101).rank(lam) == 3
    assert O.TruncBond(3).rank(lam) == 3 and O.TruncBond(9).rank(lam) == 4
    tb = O.TruncBondMax(2)
    assert tb.rank(lam) == 2
    assert math.isclose(tb.maxerr[0], math.sqrt((1e-4 + 1e-6) / np.sum(lam ** 2)))


# ---- README example (config 1): README.md:83-94 ------------------------------------------------
def test_readme_example():
    rng = RNG(1)
    A = O.rand_tt(rng, [1, 5, 5, 1], 2, 3)
    x = [(0, 1), (1, 2), (0, 0)]
    e1 = O.evaluate(A, x)
    O.compress(A, O.TruncThresh(1e-8))
    assert math.isclose(O.evaluate(A, x), e1, rel_tol=1e-12)


# ---- test/tensor_train.jl:78-183 gauge invariance ----------------------------------------------
@pytest.mark.parametrize("q", [(2,), (2, 2)])
def test_gauge_invariance_open(q):
    rng = RNG(7)
    C = O.TensorTrain([rng.random((1, 3) + q), rng.random((3, 4) + q), rng.random((4, 10) + q),
                       rng.random((10, 1) + q)], z=12.3)
    assert O.bond_dims(C) == [1, 3, 4, 10]
    x = _rand_x(rng, C)
    e1, z1 = O.evaluate(C, x), float(O.normalization(C))
    O.orthogonalize_right(C, O.TruncThresh(0.0))
    assert math.isclose(O.evaluate(C, x), e1, rel_tol=1e-10)
    assert math.isclose(float(O.normalization(C)), z1, rel_tol=1e-10)
    O.orthogonalize_left(C, O.TruncBondMax(20))
    assert math.isclose(O.evaluate(C, x), e1, rel_tol=1e-10)
    O.compress(C, O.TruncThresh(1e-15))
    assert math.isclose(O.evaluate(C, x), e1, rel_tol=1e-10)
    O.orthogonalize_right(C, O.TruncBondThresh(20, 0.0))
    assert math.isclose(O.evaluate(C, x), e1, rel_tol=1e-10)


def test_flat_known_normalization():
    """test/tensor_train.jl:168-183."""
    q, bonds = (2, 4), [1, 3, 5, 2, 1]
    C = O.flat_tt(bonds, *q)
    assert math.isclose(float(O.normalization(C)), np.prod(bonds) * np.prod(q) ** len(C), rel_tol=1e-12)
    x = _rand_x(RNG(3), C)
    e1 = O.evaluate(C, x)
    O.orthogonalize_right(C, O.TruncThresh(0.0))
    assert math.isclose(O.evaluate(C, x), e1, rel_tol=1e-9)
    O.orthogonalize_left(C, O.TruncThresh(0.0))
    assert math.isclose(O.evaluate(C, x), e1, rel_tol=1e-9)


def test_accumulators():
    """test/tensor_train.jl:185-196."""
    A = O.randn_tt(RNG(11), [1, 3, 3, 3, 1], 2, 2)
    l, _ = O.accumulate_L(A, normalize=False)
    r, _ = O.accumulate_R(A, normalize=False)
    m = O.accumulate_M(A, normalize=False)
    Z = float(O.normalization(A))
    assert math.isclose(Z, O.exact_normalization(A), rel_tol=1e-10)
    assert math.isclose(float(O.accumulate_R(A)[1]), Z, rel_tol=1e-10)
    assert math.isclose(l[-1].item(), Z, rel_tol=1e-10)
    assert math.isclose(r[0].item(), Z, rel_tol=1e-10)
    assert math.isclose((l[0] @ m[(0, len(A) - 1)] @ r[-1]).item(), Z, rel_tol=1e-10)


def test_accumulate_scaling_quirk():
    """abstract_tensor_train.jl:93-98: stored l[t<L] end up max-abs 1, l[L] unscaled."""
    A = O.rand_tt(RNG(5), [1, 4, 4, 4, 1], 3)
    l, z = O.accumulate_L(A)
    for lt in l[:-1]:
        assert math.isclose(np.max(np.abs(lt)), 1.0)
    r, z2 = O.accumulate_R(A)
    for rt in r[1:]:
        assert math.isclose(np.max(np.abs(rt)), 1.0)
    assert math.isclose(z.logabs, z2.logabs, rel_tol=1e-12)


def test_orthogonalization_canonical():
    """test/tensor_train.jl:198-215."""
    rng = RNG(13)
    B0 = O.randn_tt(rng, [1, 3, 3, 3, 1], 2, 2)
    B = B0.copy()
    O.orthogonalize_right(B, O.TruncThresh(1e-3))
    assert all(O.is_right_canonical(a) for a in B.tensors[1:])
    B = B0.copy()
    O.orthogonalize_left(B, O.TruncThresh(1e-3))
    assert all(O.is_left_canonical(a) for a in B.tensors[:-1])
    B = B0.copy()
    O.orthogonalize_center(B, 2, O.TruncBond(3))
    assert all(O.is_left_canonical(a) for a in B.tensors[:1]) and all(O.is_right_canonical(a) for a in B.tensors[2:])
    z = float(O.normalization(B))
    O.orthogonalize_center(B, 2, O.TruncThresh(0.0))
    assert math.isclose(float(O.normalization(B)), z, rel_tol=1e-10)
    O.orthogonalize_two_site_center(B, 2, O.TruncThresh(0.0))
    assert all(O.is_right_canonical(a) for a in B.tensors[3:])


def test_compression_routes_agree():
    """test/tensor_train.jl:217-230."""
    rng = RNG(17)
    A = O.randn_tt(rng, [1, 4, 4, 4, 4, 1], 3, 2)
    x = _rand_x(rng, A)
    B, C = A.copy(), A.copy()
    tr = O.TruncThresh(1e-2)
    O.compress(A, tr)
    O.orthogonalize_right(B, O.TruncThresh(0.0))
    O.compress(B, tr, is_orthogonal="right")
    O.orthogonalize_left(C, O.TruncThresh(0.0))
    O.compress(C, tr, is_orthogonal="left")
    assert math.isclose(O.evaluate(A, x), O.evaluate(B, x), rel_tol=1e-8)
    # route C truncates in the other direction; same tolerance class as the reference's `≈`
    assert math.isclose(O.evaluate(A, x), O.evaluate(C, x), rel_tol=0.2, abs_tol=1e-3)
    with pytest.raises(ValueError):
        O.compress(A, tr, is_orthogonal="something")


def test_long_chains_stay_finite():
    """test/tensor_train.jl:232-258 (L = 10 000)."""
    rng = RNG(19)
    L = 10000
    bonds = [1] + list(rng.integers(10, 16, L - 1)) + [1]
    A = O.rand_tt(rng, bonds, 2, 2)
    with np.errstate(over="ignore", invalid="ignore"):
        l, _ = O.accumulate_L(A, normalize=False)
    assert np.isinf(l[-1]).any()
    assert math.isfinite(O.lognormalization(A))
    O.normalize(A)
    assert math.isclose(float(O.normalization(A)), 1.0, rel_tol=1e-8)
    B = O.TensorTrain([a * 1e-50 for a in A.tensors])
    l, _ = O.accumulate_L(B, normalize=False)
    assert (l[-1] == 0).any()
    assert math.isfinite(O.lognormalization(B))
    O.normalize(B)
    assert math.isclose(float(O.normalization(B)), 1.0, rel_tol=1e-8)


def test_long_chain_compress_finite():
    rng = RNG(23)
    L = 1500
    A = O.rand_tt(rng, [1] + list(rng.integers(10, 16, L - 1)) + [1], 2, 2)
    O.normalize(A)
    O.compress(A, O.TruncThresh(0.0))
    assert math.isfinite(O.normalization(A).logabs)


def test_negative_values():
    """test/tensor_train.jl:260-273."""
    rng = RNG(29)
    A = O.rand_tt(rng, [1, 6, 7, 5, 1], 2, 2)
    A[0] = -A[0]
    Z = O.exact_normalization(A)
    assert Z < 0 and math.isclose(float(O.normalization(A)), Z, rel_tol=1e-10)
    with pytest.raises(O.NegativeTTError):
        O.sample(lambda t: 0.5, A)
    with pytest.raises(ValueError):
        O.lognormalization(A)


def test_normalize_eachmatrix():
    """test/tensor_train.jl:275-287."""
    rng = RNG(31)
    A = O.rand_tt(rng, [1, 11, 12, 10, 1], 2, 4)
    A.z = O.LogNum(-37.0)
    us = rng.random(len(A))
    x, p = O.sample(lambda t: us[t], A)
    xt = [np.unravel_index(xi, a.shape[2:], order="F") for xi, a in zip(x, A)]
    e, z = O.evaluate(A, xt), float(O.normalization(A))
    O.normalize_eachmatrix(A)
    assert math.isclose(O.evaluate(A, xt), e, rel_tol=1e-10)
    assert math.isclose(float(O.normalization(A)), z, rel_tol=1e-10)


@pytest.mark.parametrize("N,q", [(n, q) for n in (1, 2, 3) for q in (1, 2, 3)])
def test_sampling(N, q):
    """test/tensor_train.jl:346-364."""
    rng = RNG(37 + 3 * N + q)
    L = 6
    A = O.rand_tt(rng, [1] + list(rng.integers(1, 8, L - 1)) + [1], *([q] * N))
    us = rng.random(L)
    x, p = O.sample(lambda t: us[t], A)
    assert all(0 <= xi < q ** N for xi in x)
    y, _ = O.sample(lambda t: us[t], A)
    assert x == y
    O.normalize(A)
    xt = [np.unravel_index(xi, a.shape[2:], order="F") for xi, a in zip(x, A)]
    assert math.isclose(O.evaluate(A, xt), p, rel_tol=1e-9)


@pytest.mark.parametrize("N,q", [(n, q) for n in (1, 2) for q in (1, 2, 3)])
def test_marginals_vs_brute_force(N, q):
    """test/tensor_train.jl:366-380."""
    rng = RNG(41 + 3 * N + q)
    A = O.rand_tt(rng, [1] + list(rng.integers(1, 8, 2)) + [1], *([q] * N))
    m, me = O.marginals(A), O.exact_marginals(A)
    for a, b in zip(m, me):
        assert np.allclose(a.reshape(-1, order="F"), b, rtol=1e-10)
    m2, m2e = O.twovar_marginals(A), O.exact_twovar_marginals(A)
    for k, v in m2e.items():
        Q = q ** N
        assert np.allclose(m2[k].reshape(Q, Q, order="F"), v, rtol=1e-10)


def test_norm_and_dot():
    """test/tensor_train.jl:382-387 (without `-`, which is out of scope)."""
    rng = RNG(43)
    A = O.rand_tt(rng, [1, 2, 3, 1], 2, 2)
    B = O.rand_tt(rng, [1, 3, 3, 1], 2, 2)
    assert math.isclose(O.norm(A), O.exact_norm(A), rel_tol=1e-10)
    assert math.isclose(O.dot(A, B), O.exact_dot(A, B), rel_tol=1e-10)
    pa, pb = O.exact_prob(A), O.exact_prob(B)
    assert math.isclose(O.norm2m(A, B), float(np.sum((pa - pb) ** 2)), rel_tol=1e-8)


@pytest.mark.parametrize("kind", ["thresh", "bondthresh"])
def test_truncation_error_bound(kind):
    """test/svd_trunc.jl:1-24: norm2m(A,B) < (L*eps)^2."""
    rng = RNG(47)
    L, qs = 10, (4, 4)
    hi = 4 if kind == "thresh" else 6
    B = O.rand_tt(rng, [1] + list(rng.integers(1, hi, L - 1)) + [1], *qs)
    O.normalize(B)
    for eps in 10.0 ** np.arange(0.0, -5.5, -0.5):
        A = B.copy()
        O.compress(A, O.TruncThresh(eps) if kind == "thresh" else O.TruncBondThresh(2, eps))
        if kind == "thresh":
            assert O.norm2m(A, B) < (L * eps) ** 2 + 1e-13
        else:
            assert math.isfinite(O.norm2m(A, B))


# ---- periodic: test/periodic_tensor_train.jl ---------------------------------------------------
def test_periodic_accumulators_and_marginals():
    """test/periodic_tensor_train.jl:22-34, 242-259."""
    rng = RNG(53)
    A = O.rand_periodic_tt(rng, [3, 2, 4], 2, 2)
    Z = float(O.normalization(A))
    assert math.isclose(Z, O.exact_normalization(A), rel_tol=1e-10)
    assert math.isclose(float(O.accumulate_R(A)[1]), Z, rel_tol=1e-10)
    l, _ = O.accumulate_L(A, normalize=False)
    r, _ = O.accumulate_R(A, normalize=False)
    assert math.isclose(np.trace(l[-1]), Z, rel_tol=1e-10) and math.isclose(np.trace(r[0]), Z, rel_tol=1e-10)
    for a, b in zip(O.marginals(A), O.exact_marginals(A)):
        assert np.allclose(a.reshape(-1, order="F"), b, rtol=1e-10)
    m2, m2e = O.twovar_marginals(A), O.exact_twovar_marginals(A)
    for k, v in m2e.items():
        assert np.allclose(m2[k].reshape(4, 4, order="F"), v, rtol=1e-10)


def test_periodic_flat_is_normalised():
    """test/periodic_tensor_train.jl:113-119."""
    A = O.flat_periodic_tt([3, 5, 2, 4], 2, 3)
    assert math.isclose(float(O.normalization(A)), 1.0, rel_tol=1e-12)


def test_periodic_gauge_invariance_and_routes():
    """test/periodic_tensor_train.jl:36-51, 121-160."""
    rng = RNG(59)
    A = O.rand_periodic_tt(rng, [4, 3, 5, 2], 2, 3)
    x = _rand_x(rng, A)
    e1, z1 = O.evaluate(A, x), float(O.normalization(A))
    B = A.copy()
    O.orthogonalize_right(B)
    assert math.isclose(O.evaluate(B, x), e1, rel_tol=1e-9)
    O.orthogonalize_left(B)
    assert math.isclose(O.evaluate(B, x), e1, rel_tol=1e-9)
    assert math.isclose(float(O.normalization(B)), z1, rel_tol=1e-9)
    O.compress(B, O.TruncThresh(1e-14))
    assert math.isclose(O.evaluate(B, x), e1, rel_tol=1e-9)
    with pytest.raises(ValueError):
        O.compress(B, is_orthogonal="bad")


def test_periodic_bond_reduction_on_flat():
    """test/periodic_tensor_train.jl:203-229: TruncThresh(3e-2) reduces flat bonds."""
    A = O.flat_periodic_tt([4, 4, 4, 4], 2)
    O.compress(A, O.TruncThresh(3e-2))
    assert max(O.bond_dims(A)) < 4
    assert math.isclose(float(O.normalization(A)), 1.0, rel_tol=1e-8)


def test_periodic_sampling_invariant_under_orthogonalization():
    """test/periodic_tensor_train.jl:162-174, 231-240: same uniforms, same sample."""
    rng = RNG(61)
    A = O.rand_periodic_tt(rng, [3, 4, 2, 3, 3], 3)
    us = rng.random(len(A))
    x, p = O.sample(lambda t: us[t], A)
    B = A.copy()
    O.orthogonalize_left(B)
    y, p2 = O.sample(lambda t: us[t], B)
    assert x == y and math.isclose(p, p2, rel_tol=1e-8)
    O.normalize(A)
    assert math.isclose(O.evaluate(A, [(xi,) for xi in x]), p, rel_tol=1e-9)


def test_philox_known_answer():
    """Random123 known-answer vectors for Philox4x32-10 (kat_vectors: zero and pi cases)."""
    r = O.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(v) for v in r] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    r = O.philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(v) for v in r] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    r = O.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)
    assert [int(v) for v in r] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    u = O.philox_uniform(4, np.arange(1000), 3)
    assert ((u >= 0) & (u < 1)).all() and abs(u.mean() - 0.5) < 0.05


# ---- A + B / A - B: test/tensor_train.jl:302-345, test/periodic_tensor_train.jl:176-200 -----------
@pytest.mark.parametrize("N,q", [(n, q) for n in (1, 2, 3) for q in (1, 2, 3)])
def test_sum_and_difference_of_trains(N, q):
    rng = RNG(700 + 10 * N + q)
    L = 6
    qs = (q,) * N
    A = O.rand_tt(rng, [1] + list(rng.integers(1, 8, L - 1)) + [1], *qs)
    B = O.rand_tt(rng, [1] + list(rng.integers(1, 8, L - 1)) + [1], *qs)
    A.z, B.z = O.LogNum(-0.7), O.LogNum(3.1)
    x = _rand_x(rng, A)
    assert math.isclose(O.evaluate(A, x) + O.evaluate(B, x), O.evaluate(O.plus(A, B), x), rel_tol=1e-12)
    assert math.isclose(O.evaluate(A, x) - O.evaluate(B, x), O.evaluate(O.minus(A, B), x), rel_tol=1e-12)
    assert O.bond_dims(O.plus(A, B)) == [1] + [a + b for a, b in zip(O.bond_dims(A)[1:], O.bond_dims(B)[1:])]
    A.z = O.LogNum(0.8)
    assert O.isapprox(O.minus(O.plus(A, B), B), A, rtol=1e-4)       # (A + B) - B ~ A
    C = A.copy()
    C.z = C.z / 2
    assert O.isapprox(C, O.plus(A, A))                              # A + A = A with z / 2
    # periodic: block diagonal everywhere, z ignored (src/periodic_tensor_train.jl:62-78)
    Ap = O.rand_periodic_tt(rng, list(rng.integers(1, 6, 4)), *qs)
    Bp = O.rand_periodic_tt(rng, list(rng.integers(1, 6, 4)), *qs)
    xp = _rand_x(rng, Ap)
    assert math.isclose(O.evaluate(Ap, xp) + O.evaluate(Bp, xp), O.evaluate(O.plus(Ap, Bp), xp), rel_tol=1e-12)
    assert math.isclose(O.evaluate(Ap, xp) - O.evaluate(Bp, xp), O.evaluate(O.minus(Ap, Bp), xp), rel_tol=1e-12)
    with pytest.raises(ValueError):
        O.plus(A, O.rand_tt(rng, [1, 2, 1], *qs))


# ---- MPS wrapper: test/mps.jl:18-67 (real Float64; the reference test uses ComplexF64) ----------------
@pytest.mark.parametrize("z", [1.0, 2.0, -0.5])
def test_mps_normalization_and_marginals_vs_brute_force(z):
    rng = RNG(810)
    psi = O.TensorTrain([rng.standard_normal(s) for s in [(1, 5, 2, 2), (5, 4, 2, 2), (4, 6, 2, 2), (6, 1, 2, 2)]], z=z)
    p = O.MPS(psi)
    q = O.mps_exact_prob(p)
    zn = float(O.mps_normalization(p))
    assert math.isclose(zn, q.sum(), rel_tol=1e-12)                       # z ~ sum(q)
    assert math.isclose(zn, O.norm(psi) ** 2, rel_tol=1e-12)              # z ~ abs2(norm(psi))
    (Lm, zL), (Rm, zR) = O.mps_accumulate_L(p, False), O.mps_accumulate_R(p, False)
    assert math.isclose(float(zL), float(zR), rel_tol=1e-12)
    assert math.isclose(float(np.einsum("aabb->", Lm[-1])), float(np.einsum("aabb->", Rm[0])), rel_tol=1e-12)
    x = _rand_x(rng, psi)
    assert math.isclose(O.mps_evaluate(p, x), O.evaluate(psi, x) ** 2, rel_tol=1e-14)
    qq = q.reshape([4] * 4)
    for t, mt in enumerate(O.mps_marginals(p)):                           # exact_marginals(p) ~ marginals(p)
        pt = qq.sum(axis=tuple(i for i in range(4) if i != t))
        assert np.allclose(pt / pt.sum(), mt.reshape(-1, order="F"), rtol=1e-11, atol=0)
    pc = O.MPS(psi.copy())
    logz = O.mps_normalize(pc)
    assert math.isclose(float(O.mps_normalization(pc)), 1.0, rel_tol=1e-12)
    assert math.isclose(math.exp(logz), zn, rel_tol=1e-12)
    O.mps_normalize(pc)
    assert math.isclose(float(O.mps_normalization(pc)), 1.0, rel_tol=1e-12)


# ---- two-site pieces (src/utils.jl:37-44, src/abstract_tensor_train.jl:283-308, src/tensor_train.jl:320-349) --------
@pytest.mark.parametrize("lr", ["left", "right"])
def test_merge_split_and_data_environments_invariants(lr):
    """_split_tensor(_merge_tensors(A, B)) reproduces the product and leaves the requested factor orthonormal; the
    per-datapoint environments give back evaluate (grad_evaluate_two_site's `val`, tensor_train.jl:339-349)."""
    rng = RNG(820)
    q = (2, 3)
    ts = [rng.standard_normal(s + q) for s in [(1, 4), (4, 6), (6, 3), (3, 1)]]
    A = O.TensorTrain([a.copy() for a in ts])
    C = O.merge_tensors(A[1], A[2])
    assert C.shape == (4, 3) + q + q
    x1, x2 = (1, 2), (0, 1)
    assert np.allclose(C[(slice(None), slice(None)) + x1 + x2], A[1][(slice(None), slice(None)) + x1] @ A[2][(slice(None), slice(None)) + x2])
    a, b = O.split_tensor(C, O.TruncBond(6), lr)
    assert np.allclose(O.merge_tensors(a, b), C, rtol=1e-12, atol=1e-13)
    if lr == "left":
        assert O.is_left_canonical(a)
    else:
        assert O.is_right_canonical(b)
    a4, b4 = O.split_tensor(C, O.TruncBond(2), lr)
    assert a4.shape[1] == b4.shape[0] == 2
    X = [tuple(int(rng.integers(0, v)) for v in q) for _ in range(4)]
    pl, pr = O.precompute_left_environments(A, X), O.precompute_right_environments(A, X)
    assert pl[0].shape == (1, 1) and pr[4].shape == (1, 1)
    for k in range(3):   # val = Ax_left * Ax_center * Ax_right / z
        center = O.merge_tensors(A[k], A[k + 1])[(slice(None), slice(None)) + X[k] + X[k + 1]]
        assert math.isclose((pl[k] @ center @ pr[k + 2]).item(), O.evaluate(A, X), rel_tol=1e-12)
    A.tensors[1], A.tensors[2] = a, b
    O.update_environments([pl], [pr], A, 1, [X], lr)
    assert math.isclose(O.evaluate(A, X), O.evaluate(O.TensorTrain(ts), X), rel_tol=1e-10)       # the split changed the gauge only
    if lr == "left":
        assert np.allclose(pl[2], pl[1] @ a[(slice(None), slice(None)) + X[1]])
    else:
        assert np.allclose(pr[2], b[(slice(None), slice(None)) + X[2]] @ pr[3])
