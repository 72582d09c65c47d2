"""Committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU oracle; the
reference itself ships no golden vectors and cannot run here -- see the generator's docstring).
CPU: the oracle still reproduces them.  GPU: the CUDA path, through the C ABI, matches them."""
import glob
import os

import numpy as np
import pytest

from oracle import tt_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
IDS = [os.path.basename(f)[:-4] for f in FILES]


def _load(path):
    g = np.load(path)
    bonds, q, periodic = [int(v) for v in g["bonds"]], tuple(int(v) for v in g["q"]), bool(int(g["periodic"]))
    L = len(bonds) if periodic else len(bonds) - 1
    ts = [g[f"site{t}"] for t in range(L)]
    trunc = (str(g["trunc_kind"]), float(g["trunc_val"]))
    return g, ts, q, periodic, trunc


def _unflat(xf, q):
    return [tuple(int(v) for v in np.unravel_index(int(x), q, order="F")) for x in xf]


def test_fixtures_exist():
    assert len(FILES) >= 5


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_oracle_reproduces_golden(path):
    g, ts, q, periodic, trunc = _load(path)
    A = (O.PeriodicTensorTrain if periodic else O.TensorTrain)([a.copy() for a in ts], z=float(g["z"]))
    ev = np.array([O.evaluate(A, _unflat(x, q)) for x in g["X"]])
    assert np.allclose(ev, g["evaluate"], rtol=1e-12, atol=0)
    zn = O.normalization(A)
    assert abs(zn.logabs - float(g["logZ"])) <= 1e-12 * max(1.0, abs(float(g["logZ"]))) and zn.sign == int(g["signZ"])
    assert np.allclose(np.stack([m.reshape(-1, order="F") for m in O.marginals(A)]), g["marginals"], rtol=1e-12, atol=0)
    tr = O.TruncThresh(trunc[1]) if trunc[0] == "thresh" else O.TruncBond(int(trunc[1]))
    O.compress(A, tr)
    assert O.bond_dims(A) == [int(v) for v in g["bonds_after"]]
    ev1 = np.array([O.evaluate(A, _unflat(x, q)) for x in g["X"]])
    assert np.allclose(ev1, g["evaluate_after"], rtol=1e-9, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_cuda_path_matches_golden(path):
    import sys
    root = os.path.dirname(HERE)
    for p in (root, os.path.join(root, "tensortrains.jl_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ttb200 as T
    T.set_device(0)
    g, ts, q, periodic, trunc = _load(path)
    A = (T.PeriodicTensorTrain if periodic else T.TensorTrain)([a.copy() for a in ts], z=float(g["z"]))
    X = np.ascontiguousarray(g["X"], dtype=np.int32)
    assert np.allclose(T.evaluate(A, X), g["evaluate"], rtol=1e-10, atol=0)
    zn = T.normalization(A)
    assert abs(zn.logabs - float(g["logZ"])) <= 1e-10 * max(1.0, abs(float(g["logZ"]))) and zn.sign == int(g["signZ"])
    assert np.allclose(np.stack([m.reshape(-1, order="F") for m in T.marginals(A)]), g["marginals"], rtol=1e-10, atol=0)
    m2 = T.twovar_marginals(A)
    for k, ref in zip(g["pair_keys"], g["pair_marginals"]):
        assert np.allclose(m2[(int(k[0]), int(k[1]))].reshape(-1, order="F"), ref, rtol=1e-10, atol=0)
    if g["uniforms"].shape[0]:
        An = A.copy()
        T.normalize_(An)
        Xs, ps = T.sample(An, g["uniforms"].shape[0], uniforms=np.ascontiguousarray(g["uniforms"]))
        for i in range(g["uniforms"].shape[0]):
            assert list(Xs[i]) == [int(v) for v in g["draws"][i]]
            assert abs(ps[i] - g["draw_probs"][i]) <= 1e-8 * g["draw_probs"][i]
    tr = T.TruncThresh(trunc[1]) if trunc[0] == "thresh" else T.TruncBond(int(trunc[1]))
    T.compress_(A, tr)
    assert T.bond_dims(A) == [int(v) for v in g["bonds_after"]]
    assert np.allclose(T.evaluate(A, X), g["evaluate_after"], rtol=1e-8, atol=0)
