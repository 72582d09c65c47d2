"""Config 5: wall clock against device time of every call of the pipeline, twice on the same handle (the first
round pays workspace / scratch allocation), results into a device tensor."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tensortrains.jl_b200")]
import numpy as np, torch
import ttb200 as T
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L, d, q = 64, 32, 2
packed = np.random.Generator(np.random.Philox(5)).random(B * L * d * d * q)
A = T.PeriodicTensorTrain.empty([d] * (L + 1), (q,), batch=B)
dev_out = torch.empty((B, L * (L - 1) // 2, q * q), dtype=torch.float64, device="cuda")
for rnd in range(3):
    A.reset(); A.upload_all(packed); torch.cuda.synchronize()
    print(f"round {rnd} (B={B})")
    def timed(name, f):
        t0 = time.perf_counter(); r = f(); w = time.perf_counter() - t0
        ms, nl = A.last_timing()
        print(f"  {name}: device {ms:.2f} ms, wall {w*1e3:.2f} ms, launches {nl}")
        return r
    timed("normalize!", lambda: T.normalize_(A))
    timed("twovar_marginals -> device", lambda: T.twovar_marginals_block(A, 0, L - 1, L - 1, out=dev_out))
    timed("compress!(1e-6)", lambda: T.compress_(A, T.TruncThresh(1e-6)))
    timed("bonds of all trains", lambda: [A.bonds(b) for b in range(B)])
