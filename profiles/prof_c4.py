"""Config 4 sampling: wall clock vs device time of ttb_sample_stats / ttb_sample."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tensortrains.jl_b200")]
import numpy as np
import ttb200 as T
L, d, q = 256, 128, 4
rng = np.random.Generator(np.random.Philox(4))
bonds = [1] + [d] * (L - 1) + [1]
A = T.TensorTrain([rng.random((bonds[t], bonds[t + 1], q)) for t in range(L)])
T.normalize_(A)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
T.sample_stats(A, 20000, seed=1)
for rep in range(3):
    t0 = time.perf_counter(); hist, slp = T.sample_stats(A, n, seed=1 + rep); w = time.perf_counter() - t0
    ms, nl = A.last_timing()
    print(f"sample_stats n={n}: wall {w*1e3:.1f} ms, device {ms:.1f} ms, launches {nl}, {n/w:.0f} draws/s wall, {n/(ms*1e-3):.0f} device")
m = 200_000
t0 = time.perf_counter(); X, p = T.sample(A, m, seed=5); w = time.perf_counter() - t0
ms, nl = A.last_timing()
print(f"sample (x, p to host) n={m}: wall {w*1e3:.1f} ms, device {ms:.1f} ms, {m/w:.0f} draws/s wall")
