"""Per-kernel device time of the truncating LEFT sweep (config 3 after an eps=0 right sweep), of config 2's
TruncBond(32) compress and of config 5's batched compress, from the library's per-launch events."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tensortrains.jl_b200")]
import numpy as np
import ttb200 as T


def report(A, title):
    rep = A.profile_report()
    rows = sorted(rep.items(), key=lambda kv: -kv[1][1])
    print(title)
    tot = sum(v[1] for _, v in rows)
    for k, (n, ms) in rows:
        print(f"  {k:24s} n={n:6d} ms={ms:9.3f} us/launch={1e3 * ms / max(n, 1):8.2f} share={ms / tot:6.1%}")
    print(f"  total {tot:.2f} ms")


rng = np.random.Generator(np.random.Philox(3))
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
if which == "c3":
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    bonds = [1] + [256] * (L - 1) + [1]
    A = T.TensorTrain([rng.random((bonds[t], bonds[t + 1], 2)) for t in range(L)])
    T.orthogonalize_right_(A, T.TruncThresh(0.0))
    B = A.copy()
    T.orthogonalize_left_(B, T.TruncThresh(1e-10))
    print("left sweep unprofiled ms", B.last_timing())
    A.profile(True)
    T.orthogonalize_left_(A, T.TruncThresh(1e-10))
    report(A, f"C3 left sweep L={L}")
elif which == "c2":
    L, d, q = 128, 64, 2
    bonds = [1] + [d] * (L - 1) + [1]
    A = T.TensorTrain([rng.random((bonds[t], bonds[t + 1], q)) for t in range(L)])
    B = A.copy(); T.compress_(B, T.TruncBond(32)); print("compress unprofiled ms", B.last_timing())
    A.profile(True)
    T.compress_(A, T.TruncBond(32))
    report(A, "C2 compress TruncBond(32)")
elif which == "c5":
    Bn, L, d, q = 512, 64, 32, 2
    A = T.PeriodicTensorTrain([rng.random((Bn, d, d, q)) for _ in range(L)], batched=True)
    T.normalize_(A)
    B = A.copy(); T.compress_(B, T.TruncThresh(1e-6)); print("compress unprofiled ms", B.last_timing())
    A.profile(True)
    T.compress_(A, T.TruncThresh(1e-6))
    report(A, "C5 compress B=512")
if which == "c3right":
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    bonds = [1] + [256] * (L - 1) + [1]
    A = T.TensorTrain([rng.random((bonds[t], bonds[t + 1], 2)) for t in range(L)])
    B = A.copy()
    T.orthogonalize_right_(B, T.TruncThresh(0.0))
    print("right sweep unprofiled ms", B.last_timing())
    A.profile(True)
    T.orthogonalize_right_(A, T.TruncThresh(0.0))
    report(A, f"C3 right sweep L={L}")
