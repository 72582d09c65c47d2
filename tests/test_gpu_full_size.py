"""GPU tests at BASELINE.json's FULL sizes, through size-independent properties the domain offers
(the CPU oracle would need minutes there): gauge invariance of evaluate / log Z under compress!,
canonical form, idempotence of compress!, p(draw) == evaluate(draw), marginal consistency
sum_{x_u} b[t,u](x_t, x_u) == p_t(x_t), chi-square of draws against the marginals."""
import math

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

T = None


@pytest.fixture(scope="module", autouse=True)
def _load():
    global T
    import ttb200
    T = ttb200
    T.set_device(0)
    yield


RNG = lambda s: np.random.Generator(np.random.Philox(s))


def open_tt(seed, L, d, q, fill="rand"):
    rng = RNG(seed)
    bonds = [1] + [d] * (L - 1) + [1]
    if fill == "rand":
        return T.TensorTrain([rng.random((bonds[t], bonds[t + 1], q)) for t in range(L)])
    return T.TensorTrain([rng.standard_normal((bonds[t], bonds[t + 1], q)) / math.sqrt(d) for t in range(L)])


def test_config3_full_compress_properties():
    """C3: L=256, q=2, d=256, rand_tt, compress!(TruncThresh(1e-10))."""
    L, d, q = 256, 256, 2
    A = open_tt(3, L, d, q)
    T.normalize_(A)                                   # keeps evaluate in range on a 256-site chain
    X = RNG(33).integers(0, q, (24, L)).astype(np.int32)
    e0 = T.evaluate(A, X)
    lz0 = T.normalization(A)
    B = A.copy()
    T.compress_(B, T.TruncThresh(1e-10))
    bd = T.bond_dims(B)
    assert bd[0] == 1 and max(bd) <= 16, bd          # rand_tt is numerically rank ~ 6-8 across every cut
    e1 = T.evaluate(B, X)
    assert np.max(np.abs(e1 - e0) / np.abs(e0)) <= 1e-8
    lz1 = T.normalization(B)
    assert lz1.sign == lz0.sign and abs(lz1.logabs - lz0.logabs) <= 1e-8
    assert all(T.is_left_canonical(B, t, atol=1e-10) for t in range(0, L - 1, 17))
    # a second compress! never grows a bond and keeps the values (the reference's TruncThresh rule is
    # not exactly idempotent: it keeps one value more than the tail criterion needs, svd_trunc.jl:38-47)
    C = B.copy()
    T.compress_(C, T.TruncThresh(1e-10))
    assert all(c <= b for c, b in zip(T.bond_dims(C), bd))
    e2 = T.evaluate(C, X)
    assert np.max(np.abs(e2 - e1) / np.abs(e1)) <= 1e-8
    # eps = 0 sweeps alone change nothing observable and keep every bond at min(p, n)
    D = A.copy()
    T.orthogonalize_right_(D, T.TruncThresh(0.0))
    assert T.bond_dims(D) == [1] + [min(d, q ** (L - t)) for t in range(1, L)]   # reduced from the right only
    e3 = T.evaluate(D, X)
    assert np.max(np.abs(e3 - e0) / np.abs(e0)) <= 1e-9
    assert all(T.is_right_canonical(D, t, atol=1e-10) for t in range(1, L, 31))


def test_config3_randn_both_sweeps_heavy():
    """C3 shape with the randn fill (full rank, both sweeps heavy), shorter chain to bound the time."""
    L, d, q = 24, 256, 2
    A = open_tt(3, L, d, q, "randn")
    X = RNG(34).integers(0, q, (24, L)).astype(np.int32)
    e0 = T.evaluate(A, X)
    B = A.copy()
    T.compress_(B, T.TruncThresh(1e-10))
    e1 = T.evaluate(B, X)
    assert np.max(np.abs(e1 - e0)) <= 1e-8 * np.max(np.abs(e0))
    assert max(T.bond_dims(B)) == d                   # nothing to truncate in a random train
    assert all(T.is_left_canonical(B, t, atol=1e-10) for t in range(0, L - 1, 5))


def test_config4_full_sampling_properties():
    """C4: L=256, q=4, d=128: normalize!, 40 000 draws with the tiled tensor-core sampler."""
    L, d, q = 256, 128, 4
    A = open_tt(4, L, d, q)
    lz = T.normalize_(A)
    assert abs(T.normalization(A).logabs) <= 1e-8 and math.isfinite(lz)
    n = 40000
    X, p = T.sample(A, n, seed=2024, first_index=5)
    assert X.shape == (n, L) and X.max() < q
    # the probability returned with a draw is the train evaluated at that draw (:379)
    ev = T.evaluate(A, X[:64].astype(np.int32))
    assert np.max(np.abs(np.log(ev) - np.log(p[:64]))) <= 1e-8
    # launch-split independence of the Philox stream
    Xb, _ = T.sample(A, 500, seed=2024, first_index=5 + 1000)
    assert np.array_equal(Xb, X[1000:1500])
    # chi-square of every site's histogram against marginals(A) (dof 3, p = 1e-7 cut, 256 sites)
    m = np.stack([mt.reshape(-1) for mt in T.marginals(A)])
    cnt = np.stack([np.bincount(X[:, t], minlength=q) for t in range(L)])
    chi = np.sum((cnt - n * m) ** 2 / (n * m), axis=1)
    assert chi.max() < 45.0, chi.max()
    hist, slp = T.sample_stats(A, n, seed=2024, first_index=5)
    assert np.array_equal(hist, cnt)
    assert abs(slp - float(np.sum(np.log(p)))) <= 1e-6 * abs(slp)


def test_config5_batch_properties():
    """C5 shape (periodic, 64 sites, bond 32, q=2), 256 of the 4096 trains."""
    B, L, d, q = 256, 64, 32, 2
    rng = RNG(5)
    A = T.PeriodicTensorTrain([rng.random((B, d, d, q)) for _ in range(L)], batched=True)
    T.normalize_(A)
    z = T.normalization(A)
    assert all(zz.sign == 1 and abs(zz.logabs) <= 1e-9 for zz in z)
    m1 = T.marginals(A)
    m2 = T.twovar_marginals(A, maxdist=3)
    for b in (0, 100, 255):
        for (t, u), bt in m2[b].items():
            assert np.allclose(bt.sum(axis=1), m1[b][t], rtol=1e-9, atol=1e-13)
            assert np.allclose(bt.sum(axis=0), m1[b][u], rtol=1e-9, atol=1e-13)
    X = rng.integers(0, q, (8, L)).astype(np.int32)
    e0 = [T.evaluate(A, X, b=b) for b in (0, 100, 255)]
    T.compress_(A, T.TruncThresh(1e-6))
    z2 = T.normalization(A)
    assert all(abs(zz.logabs) <= 1e-6 for zz in z2)
    for i, b in enumerate((0, 100, 255)):
        e1 = T.evaluate(A, X, b=b)
        assert np.max(np.abs(e1 - e0[i]) / np.abs(e0[i])) <= 1e-3      # TruncThresh(1e-6) per bond, 64 bonds
        assert max(T.bond_dims(A, b)) < d


def test_config5_large_batch_matches_oracle():
    """A batch that fills the GPU (>= 2 trains per SM) takes the batch-tuned kernels -- two interleaved Jacobi
    pairs per warp, one-CTA-per-train absorb / thin Q, three panel CTAs per SM: bond dimensions and values of
    sampled trains against the oracle (C5 shape, shortened to 12 sites)."""
    from oracle import tt_oracle as O
    B, L, d, q = 320, 12, 32, 2
    rng = RNG(55)
    ts = [rng.random((B, d, d, q)) for _ in range(L)]
    A = T.PeriodicTensorTrain(ts, batched=True)
    lz = T.normalize_(A)
    m2 = T.twovar_marginals(A)
    T.compress_(A, T.TruncThresh(1e-6))
    X = rng.integers(0, q, (6, L)).astype(np.int32)
    for b in (0, 1, 159, 160, 318, 319):
        Ao = O.PeriodicTensorTrain([a[b].copy() for a in ts])
        lzo = O.normalize(Ao)
        assert abs(lz[b] - lzo) <= 1e-10 * abs(lzo)
        m2o = O.twovar_marginals(Ao)
        for k in m2o:
            assert np.allclose(m2[b][k], m2o[k], rtol=1e-10, atol=0), (b, k)
        O.compress(Ao, O.TruncThresh(1e-6))
        assert T.bond_dims(A, b) == O.bond_dims(Ao), (b, T.bond_dims(A, b), O.bond_dims(Ao))
        g = T.evaluate(A, X, b=b)
        for i in range(X.shape[0]):
            o = O.evaluate(Ao, [(int(v),) for v in X[i]])
            assert abs(g[i] - o) <= 1e-7 * abs(o), (b, i, g[i], o)
