"""Config 5: batch of B periodic trains (64 sites, bond 32, q=2): normalize! + twovar_marginals + compress!.
Also config 2 (L=128, q=2, d=64 pipeline).  Prints device times (parity against the CPU restatement: tests/test_gpu_full_size.py::test_config5_large_batch_matches_reference restatement)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tensortrains.jl_b200")]
import numpy as np
import ttb200 as T

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
L, d, q = 64, 32, 2
rng = np.random.Generator(np.random.Philox(5))
ts = [rng.random((B, d, d, q)) for _ in range(L)]
t0 = time.perf_counter()
A = T.PeriodicTensorTrain(ts, batched=True)
print(f"C5 B={B}: upload {time.perf_counter()-t0:.2f}s")
def timed(name, f):
    t0 = time.perf_counter(); r = f(); w = time.perf_counter() - t0
    ms, nl = A.last_timing()
    print(f"  {name}: device {ms:.1f} ms, wall {w*1e3:.1f} ms, launches {nl}")
    return r
lz = timed("normalize!", lambda: T.normalize_(A))
m2 = timed("twovar_marginals", lambda: T.twovar_marginals(A))
timed("compress!(TruncThresh(1e-6))", lambda: T.compress_(A, T.TruncThresh(1e-6)))
ms, _ = A.last_timing()
print(f"  compress: {B*2*L/(ms*1e-3):.0f} site-steps/s, {B/(ms*1e-3):.0f} trains/s")
# config 2
L2, d2 = 128, 64
bonds = [1] + [d2] * (L2 - 1) + [1]
rng = np.random.Generator(np.random.Philox(2))
A = T.TensorTrain([rng.random((bonds[t], bonds[t + 1], 2)) for t in range(L2)])
print("C2:")
for rep in range(2):
    Bc = A.copy()
    def tm(name, f):
        r = f(); ms, nl = Bc.last_timing(); print(f"  {name}: device {ms:.2f} ms launches {nl}"); return ms
    t = tm("orthogonalize_right!(eps=0)", lambda: T.orthogonalize_right_(Bc, T.TruncThresh(0.0)))
    t += tm("compress!(TruncBond(32), :right)", lambda: T.compress_(Bc, T.TruncBond(32), is_orthogonal="right"))
    print(f"  sweeps: {2*(L2-1)/(t*1e-3):.0f} site-steps/s")
    tm("normalize!", lambda: T.normalize_(Bc))
    tm("marginals", lambda: T.marginals(Bc))
