"""Timing / ncu driver for the sampling half of the metric: config 4 (L=256, q=4, d=128)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tensortrains.jl_b200")]
import numpy as np
import ttb200 as T
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rng = np.random.Generator(np.random.Philox(4))
bonds = [1] + [128] * (L - 1) + [1]
A = T.TensorTrain([rng.random((bonds[t], bonds[t + 1], 4)) for t in range(L)])
T.normalize_(A)
T.sample_stats(A, 1000, seed=1)
for rep in range(2):
    A.profile(True)
    t0 = time.perf_counter()
    hist, slp = T.sample_stats(A, n, seed=1)
    wall = time.perf_counter() - t0
    ms, nl = A.last_timing()
    rep_ = A.profile_report()
    A.profile(False)
    print(f"n={n} device_ms={ms:.2f} draws/s={n/(ms*1e-3):.0f} wall_draws/s={n/wall:.0f} mean_logp={slp/n:.4f}", rep_)
