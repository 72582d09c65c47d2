"""Regenerates tests/golden/*.npz:  python tests/golden/make_golden.py

The reference (Julia) cannot run in this image and ships no golden vectors, so these
fixtures are produced by the CPU oracle (oracle/tt_oracle.py), which is itself pinned by the reference's own
test-suite invariants (tests/test_oracle_reference_suite.py).  They freeze the oracle's answers on seeded
inputs -- a change of the oracle or of the CUDA path that moves a gauge-invariant number shows up as a diff
against a committed file -- they are NOT outputs of the reference implementation.
Every case stores the inputs (site tensors, z) and gauge-invariant outputs only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tt_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def case(name, seed, bonds, q, periodic, trunc, z=1.0, fill="rand"):
    rng = np.random.Generator(np.random.Philox(seed))
    L = len(bonds) if periodic else len(bonds) - 1
    shp = [(bonds[t], bonds[(t + 1) % len(bonds)]) + tuple(q) for t in range(L)]
    ts = [rng.random(s) if fill == "rand" else rng.standard_normal(s) / np.sqrt(max(bonds)) for s in shp]
    cls = O.PeriodicTensorTrain if periodic else O.TensorTrain
    A = cls([a.copy() for a in ts], z=z)
    Q = int(np.prod(q))
    X = rng.integers(0, Q, (8, L)).astype(np.int32)            # flat 0-based physical indices
    unflat = lambda xf: [tuple(int(v) for v in np.unravel_index(int(x), q, order="F")) for x in xf]
    ev0 = np.array([O.evaluate(A, unflat(x)) for x in X])
    zn = O.normalization(A)
    marg = np.stack([m.reshape(-1, order="F") for m in O.marginals(A)])
    m2 = O.twovar_marginals(A)
    keys = sorted(m2)
    pair = np.stack([m2[k].reshape(-1, order="F") for k in keys]) if keys else np.zeros((0, Q * Q))
    us = rng.random((4, L))
    draws, probs = [], []
    if fill == "rand":                                         # a train with negative values cannot be sampled
        An = A.copy()
        O.normalize(An)
        for i in range(4):
            x, pr = O.sample(lambda t: us[i, t], An)
            draws.append(x); probs.append(pr)
    else:
        us = us[:0]
    B = A.copy()
    stats = {}
    tr = {"thresh": O.TruncThresh, "bond": O.TruncBond}[trunc[0]](trunc[1])
    O.compress(B, tr)
    ev1 = np.array([O.evaluate(B, unflat(x)) for x in X])
    out = {"bonds": np.array(bonds), "q": np.array(q), "periodic": np.array(int(periodic)), "z": np.array(z),
           "trunc_kind": np.array(trunc[0]), "trunc_val": np.array(float(trunc[1])),
           "X": X, "evaluate": ev0, "logZ": np.array(zn.logabs), "signZ": np.array(zn.sign),
           "marginals": marg, "pair_keys": np.array(keys).reshape(-1, 2), "pair_marginals": pair,
           "uniforms": us, "draws": np.array(draws, dtype=np.int32).reshape(len(draws), L), "draw_probs": np.array(probs),
           "bonds_after": np.array(O.bond_dims(B)), "evaluate_after": ev1}
    for t, a in enumerate(ts):
        out[f"site{t}"] = a
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "bonds after", O.bond_dims(B))


if __name__ == "__main__":
    case("readme_config1", 1, [1, 5, 5, 1], (2, 3), False, ("thresh", 1e-8))                 # README.md:83-94
    case("open_q2_bond", 21, [1, 6, 9, 12, 9, 6, 1], (2,), False, ("bond", 4), z=0.37)
    case("open_randn", 22, [1, 4, 8, 8, 4, 1], (3,), False, ("thresh", 1e-2), fill="randn")
    case("periodic_q2", 23, [4, 6, 3, 5], (2,), True, ("thresh", 5e-2))
    case("periodic_q22", 24, [3, 5, 2, 4, 3], (2, 2), True, ("bond", 2), z=2.5)
