"""Repeated full config-3 compress! with the library's host/device timing breakdown (TTB_TIMING=1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tensortrains.jl_b200")]
import numpy as np
import ttb200 as T
L = 256
rng = np.random.Generator(np.random.Philox(3))
bonds = [1] + [256] * (L - 1) + [1]
A = T.TensorTrain([rng.random((bonds[t], bonds[t + 1], 2)) for t in range(L)])
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    B = A.copy()
    T.compress_(B, T.TruncThresh(1e-10))
    print("ms", B.last_timing())
