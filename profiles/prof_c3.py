"""Small driver for ncu captures: one compress! on a short chain of the config-3 shape (q=2, d=256)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tensortrains.jl_b200")]
import numpy as np
import ttb200 as T
L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
fill = sys.argv[2] if len(sys.argv) > 2 else "rand"
rng = np.random.Generator(np.random.Philox(3))
bonds = [1] + [256] * (L - 1) + [1]
f = (lambda s: rng.random(s)) if fill == "rand" else (lambda s: rng.standard_normal(s) / 16.0)
A = T.TensorTrain([f((bonds[t], bonds[t + 1], 2)) for t in range(L)])
T.compress_(A, T.TruncThresh(1e-10))
print("bonds", T.bond_dims(A), "ms", A.last_timing())
