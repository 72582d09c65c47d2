"""Environment-pass timing: accumulate_L/R, normalize!, marginals on configs 2/3/4 (single open trains)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tensortrains.jl_b200")]
import numpy as np
import ttb200 as T
for name, L, d, q in (("C2", 128, 64, 2), ("C4", 256, 128, 4), ("C3", 256, 256, 2)):
    rng = np.random.Generator(np.random.Philox(2))
    bonds = [1] + [d] * (L - 1) + [1]
    A = T.TensorTrain([rng.random((bonds[t], bonds[t + 1], q)) for t in range(L)])
    nbytes = T.nparams(A) * 8
    for rep in range(2):
        out = []
        for fn, label, passes in ((lambda: T.normalization(A), "accumulate_L", 1), (lambda: T.normalize_(A.copy()), "normalize!(incl copy)", 3),
                                  (lambda: T.marginals(A), "marginals", 3)):
            A.profile(True); r = fn(); rep_ = A.profile_report(); A.profile(False)
            ms, nl = A.last_timing()
            out.append(f"{label}: {ms:.3f} ms {rep_}")
        if rep: print(name, f"{nbytes/1e6:.0f} MB", " | ".join(out))
