"""Round-2 workloads for `ncu --set full` captures, one kernel family per invocation:
   qr      config 3 (L=64 slice of the chain): k_qr_stream / k_absorb_* / k_jacobi_small
   mps     MPS environments + marginals of an open train, d = 128, q = 2, L = 24: k_gemm_items
   tvopen  pair marginals of one open train, L = 128, d = 256, q = 2: k_twovar_open
   tvmma   config 5 at 512 trains: k_twovar_mma, k_env_small
Timing only (parity lives in tests/)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tensortrains.jl_b200")]
import numpy as np
import ttb200 as T
what = sys.argv[1]
rng = np.random.Generator(np.random.Philox(3))
def open_tt(L, d, q, fill="rand"):
    bonds = [1] + [d] * (L - 1) + [1]
    f = rng.random if fill == "rand" else (lambda s: rng.standard_normal(s) / np.sqrt(d))
    return T.TensorTrain([f((bonds[t], bonds[t + 1], q)) for t in range(L)])
if what == "qr":
    A = open_tt(64, 256, 2)
    T.compress_(A, T.TruncThresh(1e-10))
    print("compress ms", A.last_timing())
elif what == "mps":
    p = T.MPS(open_tt(24, 128, 2, "randn"))
    m = T.mps_marginals(p)
    print("mps marginals ms", p.psi.last_timing(), float(m[3].sum()))
elif what == "tvopen":
    A = open_tt(128, 256, 2)
    T.normalize_(A)
    o = T.twovar_marginals_block(A, 0, 127)
    print("twovar ms", A.last_timing(), o.shape)
elif what == "tvmma":
    B, L, d, q = 512, 64, 32, 2
    A = T.PeriodicTensorTrain.empty([d] * (L + 1), (q,), batch=B)
    A.upload_all(rng.random(B * L * d * d * q))
    T.normalize_(A)
    o = T.twovar_marginals_block(A, 0, L - 1)
    print("twovar ms", A.last_timing(), o.shape)
