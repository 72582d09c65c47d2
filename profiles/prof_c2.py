"""Config 2 (L=128, q=2, d=64): per-kernel CUDA-event times of orthogonalize_right!(eps=0) and compress!(TruncBond(32), :right)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tensortrains.jl_b200")]
import numpy as np
import ttb200 as T
L, d = 128, 64
bonds = [1] + [d] * (L - 1) + [1]
rng = np.random.Generator(np.random.Philox(2))
A = T.TensorTrain([rng.random((bonds[t], bonds[t + 1], 2)) for t in range(L)])
W = A.copy(); T.orthogonalize_right_(W, T.TruncThresh(0.0)); T.compress_(W, T.TruncBond(32), is_orthogonal="right")
B = A.copy()
B.profile(True)
T.orthogonalize_right_(B, T.TruncThresh(0.0)); r1 = B.profile_report()
T.compress_(B, T.TruncBond(32), is_orthogonal="right"); r2 = B.profile_report()
B.profile(False)
for name, rep in (("orthogonalize_right!(eps=0)", r1), ("compress!(TruncBond(32), :right)", r2)):
    print(name, "total %.2f ms" % sum(v[1] for v in rep.values()))
    for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
        print("    %-22s %4d launches %8.3f ms  (%.1f us each)" % (k, n, ms, 1e3 * ms / max(n, 1)))
